"""A/B of the FMA-pipe exponentials in the normaliser pass of the tiled attention kernels (debug key 6), cfg4 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode
from oracle.trx_oracle import TrxOracle
from tests.util import make_model, rel_err
from tools.bench_cfgs import timed
torch.cuda.set_stream(torch.cuda.Stream())
cfg = Cfg(way=20, seq_len=32, temp_set=[2, 3]); m, sd = make_model(cfg, 0); B = 2048
support, labels, query, _ = make_episode(cfg, B, 71, "structured")
m.set_support(poses=torch.from_numpy(support[0]).cuda()); Q = torch.from_numpy(query).cuda()
qf = m.embed(Q[:37])
lo, it = TrxOracle(cfg, sd).score(support, labels, query[:16], chunk=16)
for poly, free in ((0, 0), (1, 0), (0, 1), (1, 1)):
    m.debug_set(6, poly)
    m.debug_set(7, free)
    ms_p = timed(lambda: m.score(Q), 3)
    ms_t = timed(lambda: m.score_features(1, qf), 2)
    lg, t = m.score(Q[:16])
    print(f"poly={poly} free_a={free}: pairs {B / ms_p * 1e3:.0f} windows/s ({ms_p:.2f} ms), triples {37 / ms_t * 1e3:.0f} windows/s ({ms_t:.2f} ms), "
          f"logit err {rel_err(lg.cpu(), lo).max():.2e}", flush=True)
