"""Aggregates a step-level ncu CSV (tools/step_traffic.py) per kernel: launches, time, DRAM bytes, achieved HBM GB/s.
    python tools/ncu_step_summary.py launches.csv profiles/r2_step_traffic      -> .md and .json"""
import csv
import json
import re
import sys

src, out = sys.argv[1], sys.argv[2]
rows = []
with open(src) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
per = {}
for r in rd:
    name = r.get("Kernel Name") or ""
    metric, unit, val = r.get("Metric Name"), r.get("Metric Unit"), r.get("Metric Value")
    if not name or metric is None:
        continue
    v = float(val.replace(",", ""))
    scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    k = per.setdefault((r["ID"], name), {})
    k[metric] = v * scale
agg = {}
for (_, name), m in per.items():
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("sg::", "")
    a = agg.setdefault(short, dict(launches=0, time_s=0.0, rd=0.0, wr=0.0))
    a["launches"] += 1
    a["time_s"] += m.get("gpu__time_duration.sum", 0.0)
    a["rd"] += m.get("dram__bytes_read.sum", 0.0)
    a["wr"] += m.get("dram__bytes_write.sum", 0.0)
tot_t = sum(a["time_s"] for a in agg.values())
tot_b = sum(a["rd"] + a["wr"] for a in agg.values())
order = sorted(agg.items(), key=lambda kv: -kv[1]["time_s"])
with open(out + ".md", "w") as f:
    f.write("# One SEGAN+ G+D train step (batch 300, eager, one stream) under ncu: per-kernel time and DRAM traffic\n\n")
    f.write("ncu serialises and cold-starts every launch: use the SHARES; `--clock-control none`.\n\n")
    f.write("Total: %d launches, %.2f ms summed kernel time, %.2f GB DRAM traffic (read %.2f + write %.2f)\n\n"
            % (sum(a["launches"] for a in agg.values()), tot_t * 1e3, tot_b / 1e9,
               sum(a["rd"] for a in agg.values()) / 1e9, sum(a["wr"] for a in agg.values()) / 1e9))
    f.write("| kernel | launches | time ms | share | DRAM read MB | DRAM write MB | HBM GB/s |\n|---|---|---|---|---|---|---|\n")
    for name, a in order:
        f.write("| `%s` | %d | %.3f | %.1f %% | %.1f | %.1f | %.0f |\n"
                % (name, a["launches"], a["time_s"] * 1e3, 100 * a["time_s"] / tot_t, a["rd"] / 1e6, a["wr"] / 1e6,
                   (a["rd"] + a["wr"]) / a["time_s"] / 1e9 if a["time_s"] > 0 else 0))
js = {"total_launches": sum(a["launches"] for a in agg.values()), "total_kernel_ms": tot_t * 1e3,
      "total_dram_bytes": tot_b,
      "kernels": {n: dict(launches=a["launches"], ms=a["time_s"] * 1e3, dram_read_bytes=a["rd"], dram_write_bytes=a["wr"],
                          dram_bytes_per_launch=(a["rd"] + a["wr"]) / a["launches"]) for n, a in order}}
with open(out + ".json", "w") as f:
    json.dump(js, f, indent=1)
print(open(out + ".md").read())
