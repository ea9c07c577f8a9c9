"""Bring-up check of the tiled any-N attention kernels (arx_tcn.cu) against the CPU oracle on every tiling shape:
single tile (T=8; T=16 pairs forced onto the tiled kernels), even tile count (T=32 pairs, 4 tiles), odd tile count
(T=16 triples, 5 tiles), 39 tiles (T=32 triples), the open-set head, and the ROWMAX variant (large LayerNorm gain)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle.synth import Cfg, make_episode  # noqa: E402
from oracle.trx_oracle import TrxOracle  # noqa: E402
from tests.util import make_model, rel_err  # noqa: E402


def pairs_case(name, cfg, B, variant=0, ln_gain=None):
    m, sd = make_model(cfg, 0)
    if ln_gain is not None:
        with torch.no_grad():
            m.transformers[0].norm_k.weight.fill_(ln_gain)
        sd = dict(sd)
        sd["transformers.0.norm_k.weight"] = np.full((128,), ln_gain, np.float32)
    if variant:
        m.debug_set(0, variant)
    support, labels, query, planted = make_episode(cfg, B, 5, "structured")
    m.set_support(poses=torch.from_numpy(support[0]).cuda())
    t0 = time.time()
    logits, is_true = m.score(torch.from_numpy(query).cuda())
    torch.cuda.synchronize()
    dt = time.time() - t0
    lo, it = TrxOracle(cfg, sd).score(support, labels, query, chunk=16)
    e1, e2 = rel_err(logits.cpu(), lo).max(), rel_err(is_true.cpu(), it).max()
    print(f"{name}: path={m.last_path()} logits err {e1:.2e} is_true err {e2:.2e} argmax ok {np.array_equal(logits.argmax(1).cpu().numpy(), lo.argmax(1))} ({dt*1e3:.1f} ms)", flush=True)
    return e1 < 1e-3 and e2 < 1e-3


def triples_case(name, cfg, B):
    m, sd = make_model(cfg, 0)
    support, labels, query, _ = make_episode(cfg, B, 9, "structured")
    m.set_support(poses=torch.from_numpy(support[0]).cuda())
    qf = m.embed(torch.from_numpy(query).cuda())
    t0 = time.time()
    logits = m.score_features(1, qf)
    torch.cuda.synchronize()
    dt = time.time() - t0
    o = TrxOracle(cfg, sd)
    with torch.no_grad():
        ssf = o.embed(torch.from_numpy(support))
        ref = o.cross_transformer(ssf.expand(B, -1, -1, -1), torch.from_numpy(labels).long(), o.embed(torch.from_numpy(query)).unsqueeze(1), ti=1)
    e = rel_err(logits.cpu(), ref["logits"].numpy()).max()
    print(f"{name}: path={m.last_path()} logits err {e:.2e} ({dt*1e3:.1f} ms)", flush=True)
    return e < 1e-3


CASES = [
    lambda: triples_case("T=16 triples N=560 (5 tiles)", Cfg(way=5, seq_len=16, temp_set=[2, 3]), 9),
    lambda: pairs_case("T=16 pairs N=120 forced onto the tiled kernels", Cfg(), 131, variant=4096),
    lambda: pairs_case("T=8 pairs N=28 (1 tile)", Cfg(seq_len=8), 67),
    lambda: pairs_case("T=32 pairs N=496 20-way (4 tiles)", Cfg(way=20, seq_len=32, temp_set=[2, 3]), 16),
    lambda: pairs_case("T=16 pairs, LayerNorm gain 3 (ROWMAX)", Cfg(), 64, ln_gain=3.0),
    lambda: triples_case("T=32 triples N=4960 (39 tiles)", Cfg(way=20, seq_len=32, temp_set=[2, 3]), 2),
]
if len(sys.argv) > 1:           # one case per process: a CUDA fault in one does not hide the others
    sys.exit(0 if CASES[int(sys.argv[1])]() else 1)
ok = all([c() for c in CASES])
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
