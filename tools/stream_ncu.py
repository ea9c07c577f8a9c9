"""A few streaming frames for an ncu launch list (eager launches: ARX_GRAPHS=0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg
from tests.util import make_model
cfg = Cfg()
m, sd = make_model(cfg, 0)
rng = np.random.default_rng(7)
m.set_support(poses=torch.from_numpy((0.17 * rng.standard_normal((5, 16, 90))).astype(np.float32)).cuda())
frames = (0.17 * rng.standard_normal((24, 90))).astype(np.float32)
for f in frames:
    m.stream_push(f)
torch.cuda.synchronize()
