"""Per-kernel DRAM traffic / time from an ncu CSV (`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
[,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed] --csv`):

    python tools/traffic_summary.py in.csv out.json "<command that was profiled>"
"""
import csv, json, re, sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ""
with open(src, newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
K = defaultdict(lambda: {"launch_ids": set(), "dram_bytes_total": 0.0, "time_s_total": 0.0, "tensor": []})
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0}
for r in rd:
    name = re.sub(r"^void\s+", "", r[ik])
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"[<(].*", "", name)
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    k = K[name]
    k["launch_ids"].add(r[iid])
    m = r[im]
    if m.startswith("dram__bytes"):
        k["dram_bytes_total"] += v * unit.get(r[iu], 1.0)
    elif m.startswith("gpu__time_duration"):
        k["time_s_total"] += v * unit.get(r[iu], 1e-9)
    elif "pipe_tensor" in m:
        k["tensor"].append(v)
out = {"source": cmd, "kernels": {}}
for name, k in sorted(K.items(), key=lambda kv: -kv[1]["time_s_total"]):
    n = len(k["launch_ids"])
    out["kernels"][name] = {"launches": n, "dram_bytes_total": k["dram_bytes_total"], "dram_bytes_per_launch": k["dram_bytes_total"] / max(n, 1),
                            "time_s_total": k["time_s_total"], "gbs_while_running": k["dram_bytes_total"] / max(k["time_s_total"], 1e-12) / 1e9,
                            "tensor_pct_avg": sum(k["tensor"]) / len(k["tensor"]) if k["tensor"] else 0.0}
json.dump(out, open(dst, "w"), indent=1)
for name, v in out["kernels"].items():
    print(f"{name:28s} {v['launches']:5d} launches  {v['dram_bytes_total'] / 1e9:9.3f} GB  {v['time_s_total'] * 1e3:9.3f} ms  {v['gbs_while_running']:8.1f} GB/s  tensor {v['tensor_pct_avg']:.1f}%")
