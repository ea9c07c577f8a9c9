"""A/B of variants of the cfg2 step on the GPU box (debug key 0 bits): fused frame MLP (8192 = off), L2 prefetch ahead of the
head kernel's operand ring (16384 = off).
Every variant: whole arx_score of 4096 windows timed with CUDA events, L2 flushed before every call; bit-equality of
the scores against the default variant; then the per-stage timers of an eager pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode
from oracle.trx_oracle import TrxOracle
from tests.util import make_model, rel_err

torch.cuda.set_stream(torch.cuda.Stream())
cfg = Cfg(); m, sd = make_model(cfg, 0); B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
support, labels, query, _ = make_episode(cfg, B, 5, "structured")
m.set_support(poses=torch.from_numpy(support[0]).cuda()); Q = torch.from_numpy(query).cuda()
res = (torch.empty((B, cfg.way), device="cuda"), torch.empty((B, 1), device="cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
lo, it = TrxOracle(cfg, sd).score(support, labels, query[:256])


def timed(reps=30):
    for _ in range(3): m.score(Q, out=res)
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); m.score(Q, out=res); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


ref = None
for name, bits in (("default", 0), ("mlp unfused", 8192), ("no L2 prefetch in the head kernel", 16384), ("both off (round-2 mid state)", 8192 | 16384)):
    m.debug_set(0, bits)
    ms = timed()
    lg, t = res[0].clone(), res[1].clone()
    if ref is None:
        ref = (lg, t)
        print(f"parity vs oracle (256 windows): logits {rel_err(lg[:256].cpu(), lo).max():.2e} is_true {rel_err(t[:256].cpu(), it).max():.2e}")
    same = bool((lg == ref[0]).all()) and bool((t == ref[1]).all())
    m.profile(True); m.profile_read()
    for _ in range(10):
        flush.zero_(); m.score(Q, out=res)
    torch.cuda.synchronize()
    st, n = m.profile_read(); m.profile(False)
    print(f"{name:26s}: {ms * 1e3:7.1f} us/score  {B / ms * 1e3 / 1e6:6.2f} M windows/s  bit-identical to default: {same}  stages(us): "
          + " ".join(f"{k}={v / n * 1e3:.1f}" for k, v in st.items()), flush=True)
m.debug_set(0, 0)
