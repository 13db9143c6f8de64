# usage: bash tools/quick_bench.sh "<label>" [bench args...]   -> one compact line
label="$1"; shift
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --pool 1024 "$@" 2>&1 | tail -1 | python -c "
import json,sys
line=sys.stdin.read()
try:
    d=json.loads(line); r=d['roofline']
    print('$label', 'envs', d['config']['envs_per_gpu'], 'steps/s %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'render %.4f' % r['kernel_ms'], 'sim %.4f' % r['sim_kernel_ms'], 'frac %.3f' % r['frac'], 'e2e %.3e' % d['e2e']['value'])
except Exception as ex:
    print('$label', 'FAILED', ex, line[-400:])
"
