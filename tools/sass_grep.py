"""profiles/<tag>_sass_grep.txt: per-kernel counts of the SASS mnemonics that prove the Blackwell paths
(UTMALDG = TMA tensor load, UBLKCP = bulk copy, SYNCS = mbarrier, STG.E.*.128 = 16-byte global stores, ...).
Run in the build container:  python tools/sass_grep.py r02"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = "carlabev_env_b200/libcbev.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
PATTERNS = OrderedDict([
    ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("STG.*.128", r"\bSTG\.[A-Z.]*128"),
    ("STG (all)", r"\bSTG\b|\bSTG\."), ("LDG.*.128", r"\bLDG\.[A-Z.]*128"), ("LDS.U8", r"\bLDS\.U8"), ("LDS.128", r"\bLDS\.128"),
    ("STS", r"\bSTS\b|\bSTS\."), ("SHFL", r"\bSHFL"), ("VOTE", r"\bVOTE"), ("DFMA", r"\bDFMA"), ("DMUL", r"\bDMUL"),
    ("DADD", r"\bDADD"), ("BAR.SYNC", r"\bBAR\.SYNC"), ("ATOM / RED", r"\bATOM|\bRED\b|\bRED\."), ("IMAD.WIDE", r"\bIMAD\.WIDE"),
])
kernels = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = Counter()
        kernels[cur]["instructions"] = 0
        continue
    if cur is None or not re.search(r"/\*[0-9a-f]{4,6}\*/", line):
        continue
    kernels[cur]["instructions"] += 1
    for name, pat in PATTERNS.items():
        if re.search(pat, line):
            kernels[cur][name] += 1
demangled = {}
try:
    names = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    demangled = dict(zip(kernels, names))
except Exception:  # noqa: BLE001
    pass
out = [f"cuobjdump -sass {lib} (sm_100a): mnemonic counts per kernel\n"]
cols = ["instructions"] + list(PATTERNS)
for k, c in kernels.items():
    name = demangled.get(k, k)
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"\((RenderParams|SimParams|PoolDev).*", "", name)
    out.append(name)
    out.append("    " + "  ".join(f"{col}={c[col]}" for col in cols if c[col]))
text = "\n".join(out) + "\n"
open(f"profiles/{tag}_sass_grep.txt", "w").write(text)
print(text)
