"""Per-source-line instruction counts and stall samples of one kernel from an `ncu --set full --import-source on` report
(build container): python tools/ncu_lines.py <report.ncu-rep> <kernel substring of the mangled name> [min percent]
The SASS page of the report is joined with `nvdisasm -g` line information of the in-tree libcbev.so (same build)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "carlabev_env_b200", "libcbev.so")], cwd=tmp, check=True,
               stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("render")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l][0]
addr2line, cur = {}, None
for l in dis[start + 1:]:
    if l.startswith("//-----") and ".text." in l:
        break
    m = re.search(r'//## File ".*render.cu", line (\d+)', l)
    if m:
        cur = int(m.group(1))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur is not None:
        addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
idx = {n: i for i, n in enumerate(hdr)}
S, IE, A = idx["# Samples"], idx["Instructions Executed"], idx["Address"]
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
base, ts, ti = None, 0.0, 0.0
for r in rows[2:]:
    if len(r) <= IE:
        continue
    a = int(r[A], 16) if r[A].startswith("0x") else int(r[A])
    base = a if base is None else base
    ln = addr2line.get(a - base)
    s_, i_ = float(r[S] or 0), float(r[IE] or 0)
    agg[ln][0] += s_
    agg[ln][1] += i_
    ts += s_
    ti += i_
    for n in stalls:
        v = r[idx[n]] if idx[n] < len(r) else ""
        if v:
            agg[ln][2][n[6:]] += float(v)
src = open(os.path.join(root, "carlabev_env_b200", "csrc", "render.cu")).read().split("\n")
print(f"{rows[0][1][:80]}: {ts:.0f} samples, {ti:.0f} warp instructions")
for ln in sorted(k for k in agg if k is not None):
    s_, i_, st = agg[ln]
    if s_ > ts * minpct / 100 or i_ > ti * minpct / 100:
        top = ", ".join(f"{k} {v / max(s_, 1) * 100:.0f}%" for k, v in st.most_common(2))
        print(f"{ln:5d} smp {s_ / ts * 100:5.1f}% inst {i_ / ti * 100:5.1f}%  [{top}]  {src[ln - 1].strip()[:90]}")
