"""One arx_score of cfg4 pairs (T=32, 20-way, 2048 windows) for an ncu launch list (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import Cfg, make_episode
from tests.util import make_model
cfg = Cfg(way=20, seq_len=32, temp_set=[2, 3]); m, sd = make_model(cfg, 0); B = 2048
support, labels, query, _ = make_episode(cfg, B, 71, "structured")
m.set_support(poses=torch.from_numpy(support[0]).cuda()); Q = torch.from_numpy(query).cuda()
for _ in range(2): m.score(Q)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.score(Q)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
