"""Per-stage timers of the cfg4 shapes (T=32, 20-way; pairs with the open-set head, triples logits) on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode
from tests.util import make_model
from tools.bench_cfgs import timed
torch.cuda.set_stream(torch.cuda.Stream())
cfg = Cfg(way=20, seq_len=32, temp_set=[2, 3]); m, sd = make_model(cfg, 0); B = 2048
support, labels, query, _ = make_episode(cfg, B, 71, "structured")
m.set_support(poses=torch.from_numpy(support[0]).cuda()); Q = torch.from_numpy(query).cuda()
qf = m.embed(Q[:37])
ms_p = timed(lambda: m.score(Q), 3); ms_t = timed(lambda: m.score_features(1, qf), 2)
print(f"pairs {B / ms_p * 1e3:.0f} windows/s ({ms_p:.2f} ms), triples {37 / ms_t * 1e3:.0f} windows/s ({ms_t:.2f} ms)")
m.profile(True); m.profile_read()
for _ in range(3): m.score(Q)
torch.cuda.synchronize()
st, n = m.profile_read()
print("pairs stages (ms per score):", {k: round(v / 3, 3) for k, v in st.items()}, "chunks", n)
for _ in range(3): m.score_features(1, qf)
torch.cuda.synchronize()
st, n = m.profile_read(); m.profile(False)
print("triples stages (ms per score):", {k: round(v / 3, 3) for k, v in st.items()}, "chunks", n)
