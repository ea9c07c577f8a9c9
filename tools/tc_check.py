"""Bring-up check of the tcgen05 attention kernel against the fp32 CUDA path and the CPU oracle (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode
from oracle.trx_oracle import TrxOracle
from tests.util import make_model

def run(B, way, variant, affine=False, kind="structured"):
    cfg = Cfg(way=way)
    m32, sd = make_model(cfg, 5 if affine else 0, affine=affine, force_path=1)
    mtc, _ = make_model(cfg, 5 if affine else 0, affine=affine, force_path=2)
    support, labels, query, planted = make_episode(cfg, B, 3, kind)
    S, Q = torch.from_numpy(support[0]).cuda(), torch.from_numpy(query).cuda()
    m32.set_support(poses=S); mtc.set_support(poses=S)
    mtc.debug_set(0, variant)
    l32, t32 = m32.score(Q)
    torch.cuda.synchronize()
    t0 = time.time()
    ltc, ttc = mtc.score(Q)
    torch.cuda.synchronize()
    dt = time.time() - t0
    rel = ((ltc - l32).abs() / l32.abs()).max().item()
    relt = ((ttc - t32).abs() / t32.abs()).max().item()
    agree = (ltc.argmax(1) == l32.argmax(1)).float().mean().item()
    lo, it = TrxOracle(cfg, sd).score(support, labels, query[:64])
    relo = np.abs(ltc[:64].cpu().numpy() / lo - 1).max()
    print(f"B={B} way={way} variant={variant} affine={affine} {kind}: path={mtc.last_path()} rel(tc,fp32)={rel:.3e} is_true={relt:.3e} "
          f"argmax agree={agree:.4f} rel(tc,oracle64)={relo:.3e} nan={torch.isnan(ltc).any().item()} t={dt*1e3:.1f}ms", flush=True)
    return rel

if __name__ == "__main__":
    for variant in (1, 0):
        for B, way in [(1, 5), (2, 5), (3, 1), (64, 5), (301, 3), (1024, 5), (600, 60)]:
            run(B, way, variant)
        run(512, 5, variant, affine=True)
        run(512, 5, variant, kind="iid")
