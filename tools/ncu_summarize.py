"""Summarise gpurun_out/*.ncu-rep + launches.csv into profiles/ (run in the build container)."""
import csv
import json
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fp64.sum']


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    res = {"kernels": [r[hdr.index("Kernel Name")][:80] for r in rows[2:]]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            res[k] = {"unit": rows[1][i], "values": [r[i] for r in rows[2:]]}
    return res


for name in ("render", "sim"):
    d = raw(f"gpurun_out/prof_{name}.ncu-rep")
    json.dump(d, open(f"profiles/{tag}_k_{name}_ncu_full_summary.json", "w"), indent=1)
    print(name, d.get("gpu__time_duration.sum"), d.get("dram__bytes_write.sum"), d.get("launch__registers_per_thread"))

mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
d = json.load(open(f"profiles/{tag}_k_render_ncu_full_summary.json"))
w, r = d["dram__bytes_write.sum"], d["dram__bytes_read.sum"]
tr = sum(float(v) for v in w["values"]) / len(w["values"]) * mul[w["unit"]] + \
    sum(float(v) for v in r["values"]) / len(r["values"]) * mul[r["unit"]]
json.dump({"dram_bytes_per_launch": tr,
           "source": f"profiles/{tag}_k_render_ncu_full_summary.json (ncu --set full, bench.py --steps 6 --warmup 3 "
                     "--no-cpu-baseline --pool 512, 4096 envs)",
           "note": "dram__bytes_read.sum + dram__bytes_write.sum averaged over the captured launches"},
          open("profiles/render_traffic.json", "w"), indent=1)

rows = [r for r in csv.reader(l for l in open("gpurun_out/launches.csv") if not l.startswith("=="))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(list)
for r in rows[1:]:
    try:
        agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
    except Exception:
        pass
with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
    for k, v in agg.items():
        line = f"{k:72s} launches {len(v):3d}  avg {sum(v) / len(v) / 1e3:9.2f} us"
        print(line)
        f.write(line + "\n")
