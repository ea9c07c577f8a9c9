#!/usr/bin/env python
"""Benchmark of the AR scoring hot path (BASELINE.json metric: query windows/sec, 5-way 1-shot, T=16).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A step = one pass of the hot path over one batch: BASELINE cfg2, 4096 synthetic query windows
(T=16 x 30 joints x 3) scored against a 5-way 1-shot support set with pair tuples, per GPU (weak scaling:
windows are independent and shard with no data-path collective; with N>1 every step also does the one
broadcast of the support operands from rank 0 and the all-gather of the scores).

Prints ONE JSON line (rank 0).  `value` is device-timed with the inputs resident in HBM; `e2e` goes
through the host-buffer entry point (arx_score_host) with the pinned H2D/D2H copies inside the timed
region; `roofline` is for the dominant kernel stage (cross-attention), timed live with CUDA events on the
launching stream; `cpu_baseline` is the CPU oracle port timed on this box's host cores.
`--impl reference` times the CPU oracle port (the reference is pure Python/torch and is not shipped to the
GPU box) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "query windows/sec (5-way 1-shot, T=16)"
UNIT = "windows/s"
WINDOWS_PER_GPU = 4096
WAY, T, J3, N_TUP, D = 5, 16, 90, 120, 128
# algorithmic attention work per query window = 4*W*N^2*D FLOP
ATTN_FLOP_PER_WINDOW = 4 * WAY * N_TUP * N_TUP * D


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML; nvidia-smi fields equivalent)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def ncu_traffic_bytes():
    """DRAM bytes per launch of the attention kernel from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")
    try:
        for d in json.load(open(p)):
            if "k_attn_tc" in d["Kernel Name"]:
                scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
                rd = float(d["dram__bytes_read.sum"]) * scale[d["units"]["dram__bytes_read.sum"]]
                wr = float(d["dram__bytes_write.sum"]) * scale[d["units"]["dram__bytes_write.sum"]]
                return rd + wr
    except Exception:
        pass
    return None


def cpu_port(cfg, sd, support, labels, query, seconds, threads):
    """Time the CPU oracle port (torch CPU fp32, all host threads) on a bounded sample of the workload."""
    import torch
    from oracle.trx_oracle import TrxOracle
    torch.set_num_threads(threads)
    o = TrxOracle(cfg, sd)
    ssf = o.embed(torch.from_numpy(support))
    chunk = 512
    o.score(None, labels, query[:64], ss_features=ssf)                       # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        s = done % query.shape[0]
        q = query[s:s + chunk]
        o.score(None, labels, q, chunk=chunk, ss_features=ssf)
        done += q.shape[0]
        el = time.perf_counter() - t0
        if el >= seconds or done >= 4 * query.shape[0]:
            break
    return done / el, done, el


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (torch CPU)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.synth import Cfg, make_episode, make_state_dict
    cfg = Cfg()
    sd = make_state_dict(cfg, 0)
    support, labels, query, _ = make_episode(cfg, WINDOWS_PER_GPU, 1, "structured")
    threads = os.cpu_count() or 1
    import torch
    from oracle.trx_oracle import TrxOracle
    torch.set_num_threads(threads)
    o = TrxOracle(cfg, sd)
    ssf = o.embed(torch.from_numpy(support))
    sample = 1024                                                            # windows per step (bounded sample)
    for _ in range(max(1, args.warmup)):
        o.score(None, labels, query[:256], chunk=256, ss_features=ssf)
    t0 = time.perf_counter()
    for k in range(args.steps):
        s = (k * sample) % WINDOWS_PER_GPU
        o.score(None, labels, query[s:s + sample], chunk=512, ss_features=ssf)
    el = time.perf_counter() - t0
    val = args.steps * sample / el
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: 5-way 1-shot, T=16, J=30, pair tuples (N=120); each step scores a bounded "
                                   f"sample of {sample} of the {WINDOWS_PER_GPU} query windows on the host CPU"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} steps x {sample} windows, torch CPU fp32, {threads} threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=WINDOWS_PER_GPU, help="query windows per GPU per step")
    ap.add_argument("--force-path", type=int, default=0, help="0 auto, 1 fp32 kernels, 2 tcgen05 kernels")
    ap.add_argument("--chunk", type=int, default=0, help="windows per internal pass of arx_score (0 = library default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--static-support", action="store_true",
                    help="diagnostic: set/broadcast the support set once before timing instead of in every step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from oracle.synth import Cfg, make_episode, make_state_dict
    from tests.util import make_model
    from isbfsar_b200.dist import ScoreGatherer, broadcast_support

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run for N>1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the only collective is a 24-byte-per-window all-gather that overlaps the next batch's persistent kernels (one
        # CTA per SM, whole register file): keep NCCL's footprint to a couple of CTAs and let them in first
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "2")
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = Cfg()
    B = args.windows
    model, sd = make_model(cfg, 0, force_path=args.force_path, max_chunk=args.chunk)
    support, labels, query, planted = make_episode(cfg, B, 1 + rank, "structured")
    # every rank scores ITS OWN B windows (weak scaling); the support set is rank 0's
    support0 = make_episode(cfg, 1, 1, "structured")[0]
    q_dev = torch.from_numpy(query).to(dev)
    q_pin = torch.from_numpy(query).pin_memory()
    s_dev = torch.from_numpy(support0[0]).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)           # > 126 MB L2

    # everything below runs on an explicit stream: the legacy default stream cannot be captured, and arx_score replays
    # its kernel chain as CUDA graphs when its arguments recur
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    gatherer = ScoreGatherer(B, WAY, True, dev, depth=2)
    in_flight = []

    def step():
        # every rank processes the (replicated) support poses itself -- asynchronously on the scorer's side stream,
        # exactly as at N=1 -- scores its own shard, and the scores are all-gathered (one NCCL call per batch).  The
        # collective of batch k runs on a communication stream and is joined after batch k+1 has been scored, so that
        # it (and the rank skew it absorbs) overlaps the next batch's kernels; `drain()` joins the last one.
        if not args.static_support:
            model.set_support(poses=s_dev)
        model.score(q_dev, out=gatherer.out())
        in_flight.append(gatherer.gather_async())
        return gatherer.wait(in_flight.pop(0)) if len(in_flight) > 1 else None

    def drain():
        res = None
        while in_flight:
            res = gatherer.wait(in_flight.pop(0))
        return res

    # once, before timing (north_star: "broadcast the support-set tuple embeddings once"): rank 0 processes the support
    # set and NCCL-broadcasts the tuple embeddings; the other ranks import them and must reproduce their own local result
    model.set_support(poses=s_dev)
    ref_logits, ref_true = model.score(q_dev)
    if world > 1:
        broadcast_support(model, WAY, src=0, device=dev)
        torch.cuda.synchronize()
        got_logits, got_true = model.score(q_dev)
        if not (torch.equal(got_logits, ref_logits) and torch.equal(got_true, ref_true)):
            raise SystemExit("bench: scores with NCCL-broadcast support embeddings differ from locally computed ones")
        model.set_support(poses=s_dev)

    # correctness guard on the exact tensors being timed (oracle = checker only, small subset)
    step()
    per_rank = drain()
    torch.cuda.synchronize()
    logits = per_rank[rank][0]
    mine = logits[:32].cpu().numpy()
    from oracle.trx_oracle import TrxOracle
    lo, it = TrxOracle(cfg, sd).score(support0, labels, query[:32])
    err = float(np.abs(mine / lo - 1).max())
    if not err < 1e-3:
        raise SystemExit(f"bench: parity check failed before timing (max rel err {err:.3e})")

    for _ in range(args.warmup):
        flush.fill_(1)
        step()
    drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:            # one sampler per job: the ranks share few host cores
        sampler.start()
    # stage timers: a second pass of the same steps right after the timed loop, under identical conditions.  With the
    # timers armed arx_score launches its kernels one by one (events between stages); unarmed it replays the chain as
    # CUDA graphs, which is what a user gets -- so the throughput loop runs unarmed.
    prof_in_loop = False
    if prof_in_loop:
        model.profile(True)
        model.profile_read(reset=True)
    l0 = model.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                       # L2 flush between timed iterations (outside the events)
        ev[k][0].record()
        step()
        if k == args.steps - 1:
            drain()                                 # the last collective is joined inside the last timed interval
        ev[k][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = model.launch_count() - l0
    clocks = sampler.stop()
    if not prof_in_loop:
        model.profile(True)
        model.profile_read(reset=True)
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            step()
        drain()
        torch.cuda.synchronize()
    stage_ms, chunks = model.profile_read(reset=True)
    model.profile(False)

    model.set_support(poses=s_dev)
    # end to end through the host-buffer entry points: pinned H2D of the windows + D2H of the scores, every step.
    #  (a) streaming: arx_score_host_submit/_wait with two requests in flight (what a frame-streaming caller does);
    #  (b) blocking:  one arx_score_host call at a time.
    outs = [(torch.empty((B, WAY), dtype=torch.float32).pin_memory(), torch.empty((B, 1), dtype=torch.float32).pin_memory())
            for _ in range(3)]
    for _ in range(2):
        model.score_host(q_pin, out=outs[0])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(5, min(args.steps, 40))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.score_host(q_pin, out=outs[0])
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    for k in range(2):
        model.score_host_async(q_pin, out=outs[k]).result()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pending = []
    for k in range(e2e_steps):
        pending.append(model.score_host_async(q_pin, out=outs[k % 3]))
        if len(pending) == 2:
            pending.pop(0).result()
    for tk in pending:
        tk.result()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_check = float((outs[(e2e_steps - 1) % 3][0] - ref_logits.cpu()).abs().max())

    t = torch.tensor([total_ms, e2e_s, float(launches), e2e_sync_s], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_s, launches, e2e_sync_s = float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tmax[3])
    value = world * B * args.steps / (total_ms * 1e-3)
    e2e_val = world * B * e2e_steps / e2e_s

    if rank == 0:
        peaks = measured_peaks()
        path = model.last_path()
        attn_ms = stage_ms["cross_attention"] / max(1, args.steps)
        achieved = B * ATTN_FLOP_PER_WINDOW / (attn_ms * 1e-3) / 1e12 if attn_ms > 0 else 0.0
        peak = peaks["bf16_tflops"]
        roofline = {"bound": "tensor", "kernel": "cross_attention (" + ("tcgen05 fp16" if path == 2 else "fp32 CUDA-core") + ")",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": ncu_traffic_bytes() if (path == 2 and B == WINDOWS_PER_GPU) else None, "traffic_unit": "bytes per launch (ncu)",
                    "peak_source": peaks["source"] + " bf16 burst",
                    "algorithmic_flop_per_launch": B * ATTN_FLOP_PER_WINDOW, "ms_per_launch": attn_ms,
                    "stage_ms_per_step": {k: v / max(1, args.steps) for k, v in stage_ms.items()}}
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, done, el = cpu_port(cfg, sd, support0, labels, query, args.cpu_seconds, threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{done} of the step's windows in {el:.1f} s, oracle port (torch CPU fp32, {threads} threads)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16" if path == 2 else "f32", "data": "synthetic",
                "config": {"workload": f"cfg2: {B} query windows per GPU x 5-way 1-shot, T=16, J=30, pair tuples (N=120); "
                                       + ("step = score shard + all-gather scores (support set processed once, --static-support)" if args.static_support
                                          else "step = process support set (every rank, replicated poses) + score shard + all-gather scores (collective of batch k joined after batch k+1 is scored; the last one inside the last timed step); "
                                               "NCCL broadcast of the support tuple embeddings done and verified once before timing"),
                           "l2": "flushed between timed steps (256 MiB write)", "path": path,
                           "timing": "CUDA events per step on the launching stream, summed; max over ranks; stage timers "
                                     + "from a second pass of the same steps right after (kernels launched one by one, events between stages)"},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": world * B * T * J3 * 4,
                        "d2h_bytes_per_step": world * B * (WAY + 1) * 4, "steps": e2e_steps,
                        "mode": "streaming: arx_score_host_submit/_wait, two requests in flight, pinned host buffers",
                        "blocking_value": world * B * e2e_steps / e2e_sync_s,
                        "blocking_mode": "one synchronous arx_score_host call at a time",
                        "max_abs_diff_vs_device_path": e2e_check,
                        "timing": "host wall clock from first submit to last result, max over ranks"},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "parity_check_max_rel_err": err}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
