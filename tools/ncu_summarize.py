"""Summarise the raw-page CSV exports of tools/r2_ncu.sh (gpurun_out/r2n/*_raw.csv + launches_c2.csv) into profiles/
(run in the build container):  python tools/ncu_summarize.py r02"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/r2n"
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__occupancy_limit_blocks', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fp64.sum']
MUL = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    return re.sub(r"\((RenderParams|SimParams|PoolDev|float const).*", "", name).replace("void ", "")


def load(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    out = []
    for r in rows[2:]:
        d = {"kernel": short(r[ki])}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = float(r[i].replace(",", ""))
                    if units[i]:
                        d[k + "__unit"] = units[i]
                except ValueError:
                    pass
        stalls = {}
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio|"
                         r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", h)
            if m:
                try:
                    stalls[m.group(1) or m.group(2)] = float(r[i])
                except ValueError:
                    pass
        d["top_stalls"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        out.append(d)
    return out


def si(d, key):
    return d[key] * MUL.get(d.get(key + "__unit", ""), 1.0) if key in d else None


summary = {}
for stem in ("prof_c2", "prof_c3", "prof_c5", "prof_variants", "prof_f4"):
    p = os.path.join(src, f"{stem}_raw.csv")
    if not os.path.exists(p):
        continue
    rows = load(p)
    json.dump(rows, open(f"profiles/{tag}_{stem}_ncu_full.json", "w"), indent=1)
    for d in rows:
        t, w, r = si(d, 'gpu__time_duration.sum'), si(d, 'dram__bytes_write.sum'), si(d, 'dram__bytes_read.sum')
        print(f"{stem:14s} {d['kernel'][:44]:44s} {t * 1e6:8.1f} us  dram w {w / 1e6:7.1f} MB r {r / 1e6:6.1f} MB  regs "
              f"{d.get('launch__registers_per_thread', 0):.0f} occ smem/regs {d.get('launch__occupancy_limit_shared_mem', 0):.0f}/"
              f"{d.get('launch__occupancy_limit_registers', 0):.0f}  l1 {d.get('l1tex__throughput.avg.pct_of_peak_sustained_active', 0):.0f}% "
              f"bankconf {d.get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 0) / 1e6:.1f}M of "
              f"{d.get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 0) / 1e6:.1f}M  stalls {d['top_stalls']}")
    summary[stem] = rows

# steady-state DRAM traffic of the raster kernel per workload -> roofline.traffic of bench.py
traffic = {}
for stem, wl, cmd in (("prof_c2", "c2", "bench.py --workload c2 --steps 8 --warmup 130 (launches 125-126 of k_render: steady "
                                         "state with reset frames and ring-wrap mirrors)"),
                      ("prof_c3", "c3", "bench.py --workload c3 --steps 6 --warmup 30 --pool 512 (launch 30 of k_render)"),
                      ("prof_c5", "c5", "bench.py --workload c5 --steps 6 --warmup 30 --pool 512 (launch 30 of k_render)")):
    rr = [d for d in summary.get(stem, []) if d["kernel"].startswith("k_render")]
    if rr:
        tr = sum(si(d, 'dram__bytes_write.sum') + si(d, 'dram__bytes_read.sum') for d in rr) / len(rr)
        traffic[wl] = {"dram_bytes_per_launch": tr, "launches": len(rr),
                       "source": f"profiles/{tag}_{stem}_ncu_full.json: ncu --set full --clock-control none, {cmd}"}
for stem, wl, cmd in (("prof_f4", "f4", "bench.py --workload f4 --steps 4 --warmup 12 --pool 128 (launches 12-13 of k_render_any)"),):
    rr = [d for d in summary.get(stem, []) if d["kernel"].startswith("k_render")]
    if rr:
        tr = sum(si(d, 'dram__bytes_write.sum') + si(d, 'dram__bytes_read.sum') for d in rr) / len(rr)
        traffic[wl] = {"dram_bytes_per_launch": tr, "launches": len(rr),
                       "source": f"profiles/{tag}_{stem}_ncu_full.json: ncu --set full --clock-control none, {cmd}"}
if traffic:
    old = {}
    if os.path.exists("profiles/render_traffic.json"):
        old = json.load(open("profiles/render_traffic.json"))
    old.update(traffic)  # workloads not re-captured in this pass keep their entry
    json.dump(old, open("profiles/render_traffic.json", "w"), indent=1)

lp = os.path.join(src, "launches_c2.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[1:]:
        try:
            agg[short(r[ki])[:60]].append(float(r[vi].replace(",", "")))
        except Exception:  # noqa: BLE001
            pass
    with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400: bench.py --workload c2 --steps 6 --warmup 6 "
                "(cold-cache, serialised launches: compare SHARES)\n")
        for k, v in agg.items():
            line = f"{k:62s} launches {len(v):3d}  avg {sum(v) / len(v) / 1e3:9.2f} us"
            print(line)
            f.write(line + "\n")
    os.replace(lp, f"profiles/{tag}_launches_c2.csv") if False else None
