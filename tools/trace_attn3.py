"""Timeline of CTA 0 of the third-generation attention kernel (per-role clock stamps), run on the GPU box.
index k = class iteration (two tiles each). role 0: MMA issue stamps (mma1, mma2 w0, mma2 w1); roles 1/2: softmax group 0/1."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode
from tests.util import make_model
from isbfsar_b200 import _lib
cfg = Cfg()
m, sd = make_model(cfg, 0)
support, labels, query, _ = make_episode(cfg, 4096, 1, "structured")
m.set_support(poses=torch.from_numpy(support[0]).cuda())
Q = torch.from_numpy(query).cuda()
for _ in range(3): m.score(Q)
m.debug_set(1, 1)
m.score(Q); torch.cuda.synchronize()
buf = (C.c_longlong * (3 * 64 * 8))()
_lib.check(_lib.load().arx_debug_read_trace(m._h, buf), m._h, "trace")
t = np.array(buf[:], dtype=np.int64).reshape(3, 64, 8)
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
print("  k | MMA: mma1 mma2_w0 mma2_w1 | G0: start sfull ld_done exp_done pempty_ok stored arrived | G1: same")
for f in range(0, 36):
    print(f"{f:3d} | " + " ".join(f"{x:7d}" for x in rel[0, f, :3]) + " | " + " ".join(f"{x:7d}" for x in rel[1, f, :7]) + " | " + " ".join(f"{x:7d}" for x in rel[2, f, :7]))
for g in (1, 2):
    d = np.diff(rel[g, 8:36, 6]); print("group %d period per class (clk): mean %.0f min %d max %d" % (g - 1, d.mean(), d.min(), d.max()))
    s = rel[g, 8:36]
    print("   phases mean: wait_sfull %.0f ld %.0f exp %.0f wait_pempty %.0f store %.0f fence+arrive %.0f" % tuple((s[:, k + 1] - s[:, k]).mean() for k in range(6)))
