"""Per-frame latency of the resident streaming path, on the GPU box: arx_stream_push through the thin model wrapper and
through ActionRecognizer.inference (ar.py:30-84), plus the device time of one frame's CUDA graph."""
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
frames = (0.17 * rng.standard_normal((1100, 90))).astype(np.float32)
for f in range(100):
    ar.inference({"sk": frames[f]})
torch.cuda.synchronize()
t = []
for f in range(100, 600):
    t0 = time.perf_counter()
    res, o, _ = ar.inference({"sk": frames[f]})
    t.append(time.perf_counter() - t0)
t = np.array(t) * 1e6
print("ActionRecognizer.inference per frame: median %.1f us, p90 %.1f us, p99 %.1f us" % (np.median(t), np.percentile(t, 90), np.percentile(t, 99)))
m = ar.ar
t = []
for f in range(600, 1100):
    t0 = time.perf_counter()
    m.stream_push(frames[f])
    t.append(time.perf_counter() - t0)
t = np.array(t) * 1e6
print("TRXOS.stream_push per frame:           median %.1f us, p90 %.1f us, p99 %.1f us" % (np.median(t), np.percentile(t, 90), np.percentile(t, 99)))
m.profile(True)
m.profile_read(reset=True)
for f in range(100):
    m.stream_push(frames[f])
ms, n = m.profile_read(reset=True)
m.profile(False)
print("eager launches, stage timers per frame (us):", {k: round(1e3 * v / max(n, 1), 1) for k, v in ms.items()}, "chunks", n)
