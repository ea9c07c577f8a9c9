"""Timeline of CTA 0 of the tiled attention kernel (per-role clock stamps over its first 64 steps), run on the GPU box.
role 0: MMA1 issuer (0 step start, 1 Kc arrived, 2 S free -> issue; 3/4 Kq wait start/end at a tile-pair boundary);
roles 1/2: softmax group 0/1 (0 start, 1 S full, 2 loaded + token, 3 exps done, 4 sums done, 5 P buffer free, 6 P published)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode
from tests.util import make_model
from isbfsar_b200 import _lib
triples = len(sys.argv) > 1 and sys.argv[1] == "triples"
cfg = Cfg(way=20, seq_len=32, temp_set=[2, 3])
m, sd = make_model(cfg, 0)
support, labels, query, _ = make_episode(cfg, 1024, 71, "structured")
m.set_support(poses=torch.from_numpy(support[0]).cuda())
Q = torch.from_numpy(query).cuda()
qf = m.embed(Q[:37])
run = (lambda: m.score_features(1, qf)) if triples else (lambda: m.score_features(0, m.embed(Q)))
for _ in range(2): run()
m.debug_set(1, 1)
run(); torch.cuda.synchronize()
buf = (C.c_longlong * (3 * 64 * 8))()
_lib.check(_lib.load().arx_debug_read_trace(m._h, buf), m._h, "trace")
t = np.array(buf[:], dtype=np.int64).reshape(3, 64, 8)
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
print("step | MMA1: start kc_ok issue kqwait kq_ok | G0: start sfull ld+tok exps sums pfree published | G1: same")
for f in range(0, 52):
    print(f"{f:3d} | " + " ".join(f"{x:7d}" for x in rel[0, f, :5]) + " | " + " ".join(f"{x:7d}" for x in rel[1, f, :7]) + " | " + " ".join(f"{x:7d}" for x in rel[2, f, :7]))
