"""One arx_set_support + arx_score step at the bench workload, for `ncu --set full` (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import Cfg, make_episode
from tests.util import make_model
cfg = Cfg(); m, sd = make_model(cfg, 0)
support, labels, query, _ = make_episode(cfg, 4096, 1, "structured")
S = torch.from_numpy(support[0]).cuda(); Q = torch.from_numpy(query).cuda()
for _ in range(3):
    m.set_support(poses=S); m.score(Q)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.set_support(poses=S); m.score(Q)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
