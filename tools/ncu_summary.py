"""Condense an `ncu --set full` report into the JSON summary kept under profiles/ (one record per kernel launch).
usage: python tools/ncu_summary.py gpurun_out/step_full.ncu-rep profiles/r01_ncu_full_summary.json"""
import csv, json, subprocess, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
recs = []
for r in rows[2:]:
    d = {"Kernel Name": r[idx["Kernel Name"]], "Grid Size": r[idx["Grid Size"]], "Block Size": r[idx["Block Size"]], "units": {}}
    for k in KEEP:
        if k in idx:
            d[k] = r[idx[k]]
            d["units"][k] = units[idx[k]]
    recs.append(d)
json.dump(recs, open(out, "w"), indent=1)
print(len(recs), "launches ->", out)
