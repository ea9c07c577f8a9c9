"""Timeline of CTA 0 of the fused frame-MLP kernel (clock stamps per tile and role), run on the GPU box.
MMA issuer: 0 top, 1 X full, 2 fc1 issued, 3 fc2 accumulator free, 4-6 hidden sub-tile 0-2 ready, 7 fc2 issued;
epilogue: 0 top, 1 fc1 accumulator full, 2 hidden image written, 3 fc2 accumulator full, 4 stores issued;
loader: 0 top (loads issued), 1 data arrived + converted, 2 operand image written, 3 published."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode
from tests.util import make_model
from isbfsar_b200 import _lib
cfg = Cfg(); m, sd = make_model(cfg, 0)
support, labels, query, _ = make_episode(cfg, 4096, 5, "structured")
m.set_support(poses=torch.from_numpy(support[0]).cuda()); Q = torch.from_numpy(query).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2): m.score(Q)
m.debug_set(1, 2)
flush.zero_(); m.score(Q); torch.cuda.synchronize()
buf = (C.c_longlong * (3 * 64 * 8))()
_lib.check(_lib.load().arx_debug_read_trace(m._h, buf), m._h, "trace")
t = np.array(buf[:], dtype=np.int64).reshape(3, 64, 8)
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
print("tile | MMA: top xfull fc1 a2free h0 h1 h2 fc2 | EPI: top a1full hdone a2full stored | LOAD: top arrived written published")
for f in range(5):
    print(f"{f:3d} | " + " ".join(f"{x:6d}" for x in rel[0, f, :8]) + " | " + " ".join(f"{x:6d}" for x in rel[1, f, :5]) + " | " + " ".join(f"{x:6d}" for x in rel[2, f, :4]))
