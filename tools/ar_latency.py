"""Per-frame latency of ActionRecognizer.inference (B=1 streaming path, ar.py:30-84), on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_state_dict
from tests.util import Args, torch_sd
from isbfsar_b200 import ActionRecognizer
cfg = Cfg()
ar = ActionRecognizer(Args(cfg), state_dict=torch_sd(make_state_dict(cfg, 0)))
rng = np.random.default_rng(7)
for i, n in enumerate(["a", "b", "c", "d", "e"]):
    ar.train({"flag": n, "data": {"poses": (0.17 * rng.standard_normal((16, 90))).astype(np.float32)}, "requires_focus": False})
frames = (0.17 * rng.standard_normal((600, 90))).astype(np.float32)
for f in range(100):
    ar.inference({"sk": frames[f]})
torch.cuda.synchronize()
t = []
for f in range(100, 600):
    t0 = time.perf_counter()
    res, o, _ = ar.inference({"sk": frames[f]})
    t.append(time.perf_counter() - t0)
t = np.array(t) * 1e6
print("ActionRecognizer.inference per frame: median %.1f us, p90 %.1f us, p99 %.1f us (5-way, T=16, one H2D + one D2H + one sync per frame)" % (np.median(t), np.percentile(t, 90), np.percentile(t, 99)))
