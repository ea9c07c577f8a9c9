"""Attention stage time of arx_score against the softmax-group stagger of k_attn_tc3 (debug key 3), on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import Cfg, make_episode
from tests.util import make_model
cfg = Cfg(); m, sd = make_model(cfg, 0)
support, labels, query, _ = make_episode(cfg, 4096, 1, "structured")
m.set_support(poses=torch.from_numpy(support[0]).cuda()); Q = torch.from_numpy(query).cuda()
for stg in [int(a) for a in sys.argv[1:]] or [0, 400, 800, 1000, 1200, 1500, 2000]:
    m.debug_set(3, stg)
    for _ in range(5): m.score(Q)
    m.profile(True); m.profile_read(reset=True)
    for _ in range(20): m.score(Q)
    torch.cuda.synchronize(); ms, n = m.profile_read()
    print("stagger", stg, {k: round(v / 20, 4) for k, v in ms.items()})
