import sys, json
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "blocking", round(d["e2e"].get("blocking_value", 0)), "e2e_diff", d["e2e"].get("max_abs_diff_vs_device_path"), "ms/step", round(d["ms_per_step"], 3), "attn frac", round(r.get("frac", 0), 3),
          {k: round(v, 3) for k, v in r.get("stage_ms_per_step", {}).items()}, "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "err", d.get("parity_check_max_rel_err"))
