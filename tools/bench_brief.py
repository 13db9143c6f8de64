"""One line per bench JSON file: value, ms/step, dominant-kernel time and roofline fractions, e2e."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f"{f}: {d['value']:.3e} {d['unit']}, {d['ms_per_step']:.4f} ms/step, {r['kernel']} {r['kernel_ms']:.4f} ms "
              f"frac {r['frac']:.3f} step_frac {r['step_frac']:.3f}, e2e {d['e2e']['value']:.3e}")
        for w in d.get("extra_workloads", []):
            rr = w.get("roofline", {})
            print(f"    {w['name']}: {w.get('value', 0):.3e}, frac {rr.get('frac')}, step_frac {rr.get('step_frac')}"
                  f" {w.get('skipped') or ''}{w.get('error') or ''}")
    except Exception as ex:  # noqa: BLE001
        print(f, "unreadable:", ex)
