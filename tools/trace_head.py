"""Timeline of CTA 0 of the open-set head kernel k_head2_tc (clock stamps per window and role), run on the GPU box.
producer: 0 top, 1 ring stage free, 2 Kq/Kc copies issued, 3 Uc stage free | MMA1 issuer: 4 top, 5 operands landed, 6 S free -> issue;
softmax group (even windows group 0, odd group 1): 0 top, 1 S full, 2 loaded, 3 exps issued, 4 sums done, 5 P buffer free, 6 P published;
epilogue: 0 top, 1 Y full."""
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
m.debug_set(1, 3)
flush.zero_(); m.score(Q); torch.cuda.synchronize()
buf = (C.c_longlong * (3 * 64 * 8))()
_lib.check(_lib.load().arx_debug_read_trace(m._h, buf), m._h, "trace")
t = np.array(buf[:], dtype=np.int64).reshape(3, 64, 8)
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
print("win | PROD: top stagefree issued ucfree | MMA1: top landed issue | SOFTMAX: top sfull loaded exps sums pfree published | EPI: top yfull")
for f in range(28):
    print(f"{f:3d} | " + " ".join(f"{x:6d}" for x in rel[0, f, :4]) + " | " + " ".join(f"{x:6d}" for x in rel[0, f, 4:7]) + " | "
          + " ".join(f"{x:6d}" for x in rel[1, f, :7]) + " | " + " ".join(f"{x:6d}" for x in rel[2, f, :2]))
