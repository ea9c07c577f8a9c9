"""One pass of the other BASELINE shapes for `ncu --set full --profile-from-start off` (run on the GPU box):
cfg4 pairs (T=32, 20-way, N=496) and triples (N=4960) on the tiled kernels, cfg3 (60-way), the decode kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode, make_heatmaps
from tests.util import make_model
from isbfsar_b200 import HeatmapDecoder

cfg4 = Cfg(way=20, seq_len=32, temp_set=[2, 3])
m4, _ = make_model(cfg4, 0)
s4, _, q4, _ = make_episode(cfg4, 148, 71, "structured")
m4.set_support(poses=torch.from_numpy(s4[0]).cuda())
Q4 = torch.from_numpy(q4).cuda()
qf = m4.embed(Q4[:8])
cfg3 = Cfg(way=60)
m3, _ = make_model(cfg3, 0)
s3, _, q3, _ = make_episode(cfg3, 2048, 61, "structured")
m3.set_support(poses=torch.from_numpy(s3[0]).cuda())
Q3 = torch.from_numpy(q3).cuda()
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "decode_64.npz"))
m5, _ = make_model(Cfg(), 0)
dec = HeatmapDecoder(m5, g["expand30"], None, g["new_K"], g["homo_inv"])
hm = torch.from_numpy(make_heatmaps(1024, seed=2)).cuda()
hm = torch.cat([hm] * 4)


def once():
    m4.score(Q4)
    m4.score_features(1, qf)
    m3.score(Q3)
    dec.decode(hm)


once(); once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
