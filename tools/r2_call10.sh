show () { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$1', 'ms %.4f'%d['ms_per_step'], 'render %.4f move %.4f judge %.4f'%(r['kernel_ms'],r['sim_kernel_ms'],r['judge_kernel_ms']), 'frac %.3f step %.3f'%(r['frac'],r['step_frac']), 'e2e %.3e venv %.3e'%(d['e2e']['value'], d['e2e_vector_env']['value']), 'blocks', ['%.2f'%b for b in d['blocks_ms']])"; }
B="python bench.py --steps 300 --warmup 60 --no-cpu-baseline --no-extras"
for v in 0 256 0 256; do timeout 300 $B --workload c2 --debug-flags $v 2>/dev/null | show "c2 flags $v"; done
for v in 0 256; do timeout 300 $B --workload c3 --pool 512 --debug-flags $v 2>/dev/null | show "c3 flags $v"; done
for v in 0 256; do timeout 300 $B --workload c5 --pool 512 --debug-flags $v 2>/dev/null | show "c5 flags $v"; done
