"""Stage timings of arx_score under a debug variant (timing only; results may be invalid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import Cfg, make_episode
from tests.util import make_model
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cfg = Cfg(); m, sd = make_model(cfg, 0)
support, labels, query, _ = make_episode(cfg, 4096, 1, "structured")
m.set_support(poses=torch.from_numpy(support[0]).cuda()); Q = torch.from_numpy(query).cuda()
m.debug_set(0, variant)
for _ in range(5): m.score(Q)
m.profile(True); m.profile_read(reset=True)
for _ in range(20): m.score(Q)
torch.cuda.synchronize(); ms, n = m.profile_read()
print("variant", variant, {k: round(v / 20, 4) for k, v in ms.items()})
