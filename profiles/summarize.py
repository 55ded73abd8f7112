#!/usr/bin/env python
"""Turns the raw ncu output of profiles/capture.sh (gpurun_out/) into the small tracked summaries under profiles/:

    python profiles/summarize.py r1

  launches_<tag>_summary.csv   per kernel: launches, total device time, share of all profiled launches (ncu launch list)
  metrics_<tag>_summary.csv    per profiled launch: duration, DRAM bytes, DRAM %, tensor-pipe %, issue %, occupancy, registers
  traffic.json                 kernel family -> mean DRAM bytes (read+write) per launch, read by bench.py's roofline.traffic
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src = os.path.join(ROOT, "gpurun_out")
out = os.path.join(ROOT, "profiles")


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("b200::", "")
    return re.sub(r"(<unnamed>|\(anonymous namespace\))::", "", name).strip()


def read_ncu_csv(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[hi], rows[hi + 1:]


# ---- launch list ----
hdr, rows = read_ncu_csv(os.path.join(src, f"launches_{tag}.csv"))
ix = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict()
for r in rows:
    if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    k = short(r[ix["Kernel Name"]])
    unit = r[ix["Metric Unit"]]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out, f"launches_{tag}_summary.csv"), "w") as f:
    f.write("kernel,launches,total_us,share\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{us:.1f},{us / tot:.4f}\n")
print("launch list:", len(rows), "rows,", len(agg), "kernels")

# ---- full metrics ----
want = OrderedDict([
    ("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
])
mpath = os.path.join(src, f"metrics_{tag}.csv")
traffic = defaultdict(list)
if os.path.exists(mpath):
    rows = list(csv.reader(open(mpath)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in want if c in ix]
    with open(os.path.join(out, f"metrics_{tag}_summary.csv"), "w") as f:
        f.write("kernel," + ",".join(f"{want[c]}" for c in cols) + "\n")
        for r in data:
            k = short(r[ix["Kernel Name"]])
            vals = []
            for c in cols:
                v = r[ix[c]].replace(",", "")
                u = units[ix[c]]
                try:
                    x = float(v)
                    if c.startswith("dram__bytes"):
                        x = x * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    if c == "gpu__time_duration.sum":
                        x = x * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
                    vals.append(f"{x:.3f}")
                except ValueError:
                    vals.append(v)
            f.write(k + "," + ",".join(vals) + "\n")
            try:
                rd = float(vals[cols.index("dram__bytes_read.sum")]); wr = float(vals[cols.index("dram__bytes_write.sum")])
                traffic[k].append((rd + wr) * 1e6)
            except Exception:
                pass
    print("metrics:", len(data), "launches")
# bench.py names kernels by what they compute (gemm_tcgen05_conv1x1 / _linear / _qkv ...), ncu by symbol (k_gemm_tcgen05<EPI>).
# The profiled launches are the first forward pass in plan order, so gpurun_out/layers_<tag>.txt (tests/profile_layers.py,
# same plan) maps launch i to its bench name; without that file the symbol name is used.
order = []
lpath = os.path.join(src, f"layers_{tag}.txt")
if os.path.exists(lpath):
    for line in open(lpath):
        m = re.match(r"\s*[\d.]+ us\s+\d+ GB/s\s+[\d.]+ TF/s\s+[\d.]+ MB\s+(\S+)", line)
        if m:
            order.append(m.group(1))
    import shutil
    shutil.copy(lpath, os.path.join(out, f"layers_{tag}.txt"))
tj = defaultdict(list)
flat = []  # (symbol, bytes) in launch order
if os.path.exists(mpath):
    for r in data:
        try:
            rd = float(r[ix["dram__bytes_read.sum"]].replace(",", "")); wr = float(r[ix["dram__bytes_write.sum"]].replace(",", ""))
            sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            flat.append((short(r[ix["Kernel Name"]]), rd * sc.get(units[ix["dram__bytes_read.sum"]], 1.0) + wr * sc.get(units[ix["dram__bytes_write.sum"]], 1.0)))
        except Exception:
            pass
# the full capture only profiles the hot kernels (capture.sh -k regex): drop the plan's other launches from the order, and map ONE forward
order = [o for o in order if o not in ("nhwc_to_ggml_layout", "pool_mean", "classifier_head_f32", "residual_add")]
for i, (sym, b) in enumerate(flat[:len(order)] if order else flat):
    tj[order[i] if order else sym].append(b)
json.dump({k: sum(v) / len(v) for k, v in tj.items()}, open(os.path.join(out, "traffic.json"), "w"), indent=1)
print({k: (len(v), round(sum(v) / len(v) / 1e6, 1)) for k, v in tj.items()})
