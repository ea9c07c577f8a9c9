"""Small-batch smoke of arx_score under a debug variant / stagger (bring-up aid), on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode
from tests.util import make_model
variant, stagger, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = Cfg(); m, sd = make_model(cfg, 0)
support, labels, query, _ = make_episode(cfg, max(n, 8), 1, "structured")
m.debug_set(0, variant); m.debug_set(3, stagger)
m.set_support(poses=torch.from_numpy(support[0]).cuda()); Q = torch.from_numpy(query[:n]).cuda()
lg, it = m.score(Q); torch.cuda.synchronize()
print("variant", variant, "stagger", stagger, "n", n, "ok", lg.shape, float(lg.abs().max()))
