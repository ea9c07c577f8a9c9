"""Throughput of the other BASELINE configs on one GPU (cfg3 60-way, cfg4 T=32 pairs / triples, cfg5 decode), device-timed
with CUDA events on an explicit stream after warm-up; prints one JSON line per config.  Run on the GPU box."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.synth import Cfg, make_episode, make_heatmaps
from tests.util import make_model


def timed(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    torch.cuda.set_stream(torch.cuda.Stream())
    out = []
    # cfg3: 60-way, T=16 pairs, one rank's shard of the 65 536 windows
    try:
        cfg = Cfg(way=60); m, sd = make_model(cfg, 0); B = 8192
        support, labels, query, _ = make_episode(cfg, B, 61, "structured")
        m.set_support(poses=torch.from_numpy(support[0]).cuda()); Q = torch.from_numpy(query).cuda()
        res = (torch.empty((B, 60), device="cuda"), torch.empty((B, 1), device="cuda"))
        ms = timed(lambda: m.score(Q, out=res), 5)
        out.append({"config": "cfg3 60-way T=16 pairs, B=8192", "path": m.last_path(), "ms": ms, "windows_per_s": B / ms * 1e3,
                    "attention_tflops": 4 * 60 * 120 * 120 * 128 * B / ms / 1e9})
    except Exception as e:
        out.append({"config": "cfg3", "error": repr(e)})
    # cfg4 (i): 20-way, T=32, pairs N=496 (forward: logits + is_true)
    try:
        cfg = Cfg(way=20, seq_len=32, temp_set=[2, 3]); m, sd = make_model(cfg, 0); B = 2048
        support, labels, query, _ = make_episode(cfg, B, 71, "structured")
        m.set_support(poses=torch.from_numpy(support[0]).cuda()); Q = torch.from_numpy(query).cuda()
        ms = timed(lambda: m.score(Q), 3)
        out.append({"config": f"cfg4 20-way T=32 pairs N=496, B={B}", "path": m.last_path(), "ms": ms, "windows_per_s": B / ms * 1e3,
                    "attention_tflops": 4 * 20 * 496 * 496 * 128 * B / ms / 1e9})
        # cfg4 (ii): triples N=4960 through transformers[1]
        Bt = 37
        qf = m.embed(Q[:Bt])
        ms = timed(lambda: m.score_features(1, qf), 2)
        out.append({"config": f"cfg4 20-way T=32 triples N=4960, B={Bt}", "path": m.last_path(), "ms": ms, "windows_per_s": Bt / ms * 1e3,
                    "attention_tflops": 4 * 20 * 4960 * 4960 * 128 * Bt / ms / 1e9})
    except Exception as e:
        out.append({"config": "cfg4", "error": repr(e)})
    # cfg5: heatmap decode of 1024 frames
    try:
        from isbfsar_b200 import HeatmapDecoder
        g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "decode_64.npz"))
        m, _ = make_model(Cfg(), 0)
        dec = HeatmapDecoder(m, g["expand30"], None, g["new_K"], g["homo_inv"])
        hm = torch.from_numpy(make_heatmaps(1024, seed=2)).cuda()
        ms = timed(lambda: dec.decode(hm), 10)
        out.append({"config": "cfg5 decode 1024 frames", "ms": ms, "frames_per_s": 1024 / ms * 1e3, "GBps": 1024 * 73728 / ms / 1e6})
    except Exception as e:
        out.append({"config": "cfg5", "error": repr(e)})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
