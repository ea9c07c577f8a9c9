"""Timeline of CTA 0 of the tcgen05 attention kernel (per-role clock stamps), run on the GPU box."""
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
print("tile | MMA: mma1_issue(f) mma2_wait_start pfull_ok oempty_ok | SOFTMAX: start sfull ld_done exp_done pempty_ok stored arrived | EPI: start ofull done")
for f in range(0, 40):
    print(f"{f:3d} | " + " ".join(f"{x:7d}" for x in rel[0, f, :4]) + " | " + " ".join(f"{x:7d}" for x in rel[1, f, :7]) + " | " + " ".join(f"{x:7d}" for x in rel[2, f, :3]))
d = np.diff(rel[1, 8:40, 6]); print("softmax period (clk) tiles 8..40: mean %.0f min %d max %d" % (d.mean(), d.min(), d.max()))
s = rel[1, 8:40]
print("softmax phases mean: wait_sfull %.0f ld %.0f exp %.0f wait_pempty %.0f store %.0f fence+arrive %.0f" % tuple((s[:, k + 1] - s[:, k]).mean() for k in range(6)))
e = rel[2, 8:40]; print("epilogue phases mean: wait_ofull %.0f work %.0f" % ((e[:, 1] - e[:, 0]).mean(), (e[:, 2] - e[:, 1]).mean()))
