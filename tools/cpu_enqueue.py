"""Host-side enqueue cost of one bench step (set_support + score) vs its GPU time, on the GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import Cfg, make_episode
from tests.util import make_model
cfg = Cfg(); m, sd = make_model(cfg, 0)
support, labels, query, _ = make_episode(cfg, 4096, 1, "structured")
S = torch.from_numpy(support[0]).cuda(); Q = torch.from_numpy(query).cuda()
out = (torch.empty((4096, 5), device="cuda"), torch.empty((4096, 1), device="cuda"))
torch.cuda.set_stream(torch.cuda.Stream())
for prof in (False, True):
    m.profile(prof)
    for _ in range(5):
        m.set_support(poses=S); m.score(Q, out=out)
    torch.cuda.synchronize()
    K = 200
    t0 = time.perf_counter()
    for _ in range(K):
        m.set_support(poses=S); m.score(Q, out=out)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("profile", prof, "host enqueue %.3f ms/step, total %.3f ms/step" % ((t1 - t0) / K * 1e3, (t2 - t0) / K * 1e3))
print("host cores", os.cpu_count())
