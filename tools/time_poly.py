"""Attention stage time and parity error against the FMA-pipe exp2 share of k_attn_tc3 (debug key 4), on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode
from oracle.trx_oracle import TrxOracle
from tests.util import make_model
cfg = Cfg(); m, sd = make_model(cfg, 0)
support, labels, query, _ = make_episode(cfg, 4096, 1, "structured")
m.set_support(poses=torch.from_numpy(support[0]).cuda()); Q = torch.from_numpy(query).cuda()
o = TrxOracle(cfg, sd)
ref = o.forward({"sk": support}, labels, {"sk": query[:256]})
for poly in [int(a) for a in sys.argv[1:]] or [0, 4, 3, 2]:
    m.debug_set(4, poly)
    for _ in range(5): lg, it = m.score(Q)
    m.profile(True); m.profile_read(reset=True)
    for _ in range(20): m.score(Q)
    torch.cuda.synchronize(); ms, n = m.profile_read(); m.profile(False)
    a = lg[:256].cpu().numpy(); r = np.asarray(ref["logits"])
    err = np.abs(a - r).max() / np.abs(r).max()
    print("poly", poly, "attention %.4f ms" % (ms["cross_attention"] / 20), "head %.4f" % (ms["open_set_head"] / 20), "max rel err %.2e" % err,
          "argmax agree", float((a.argmax(1) == r.argmax(1)).mean()))
