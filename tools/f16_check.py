import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode
from tests.util import make_model
cfg = Cfg(); m, sd = make_model(cfg, 0)
support, labels, query, _ = make_episode(cfg, 4096, 1, "structured")
m.set_support(poses=torch.from_numpy(support[0]).cuda())
q32 = torch.from_numpy(query).pin_memory(); q16 = torch.from_numpy(query).to(torch.float16).pin_memory()
outs = [(torch.empty((4096, 5)).pin_memory(), torch.empty((4096, 1)).pin_memory()) for _ in range(3)]
for name, q in [("f32", q32), ("f16", q16), ("f32", q32), ("f16", q16)]:
    for k in range(3):
        m.score_host_async(q, out=outs[k]).result()
    torch.cuda.synchronize()
    l0 = m.launch_count()
    t0 = time.perf_counter(); pend = []
    for k in range(100):
        pend.append(m.score_host_async(q, out=outs[k % 3]))
        if len(pend) == 2: pend.pop(0).result()
    for t in pend: t.result()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(name, "streaming ms/step", dt * 10, "launches/step", (m.launch_count() - l0) / 100)
    t0 = time.perf_counter()
    for k in range(50):
        m.score_host(q, out=outs[0])
    print(name, "blocking ms/step", (time.perf_counter() - t0) * 20)
