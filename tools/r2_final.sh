# final validation with the driver's own commands
mkdir -p gpurun_out/r2z
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2z/pytest.log 2>&1; echo "pytest rc $?"; grep -E "^E  " gpurun_out/r2z/pytest.log | head; tail -3 gpurun_out/r2z/pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2z/smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2z/smoke.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z/bench_ref.json 2> gpurun_out/r2z/bench_ref.err ) 2>&1 | grep real; echo "ref rc $?"; cut -c1-300 gpurun_out/r2z/bench_ref.json
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z/bench_20.json 2> gpurun_out/r2z/bench_20.err ) 2>&1 | grep real; echo "bench20 rc $?"
( time timeout 900 python bench.py > gpurun_out/r2z/bench_default.json 2> gpurun_out/r2z/bench_default.err ) 2>&1 | grep real; echo "bench default rc $?"
timeout 300 python bench.py --no-extras --no-cpu-baseline --device-pool --steps 100 > gpurun_out/r2z/bench_devpool.json 2> gpurun_out/r2z/bench_devpool.err; echo "devpool rc $?"
python - <<'PY'
import json
for f in ('bench_20','bench_default','bench_devpool'):
    try:
        d=json.loads(open(f'gpurun_out/r2z/{f}.json').read().strip().splitlines()[-1]); r=d['roofline']
        print(f, 'val %.3e'%d['value'], 'ms %.4f'%d['ms_per_step'], 'frac %.3f step %.3f'%(r['frac'],r['step_frac']), 'e2e %.3e venv %.3e'%(d['e2e']['value'], d['e2e_vector_env']['value']), 'clocks', d['clocks'], 'traffic', r['traffic'])
        for w in d.get('extra_workloads',[]): print('   ', w['name'], '%.3e'%w.get('value',0), w.get('skipped'), w.get('error'))
        if d.get('cpu_baseline'): print('    cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
        print('    pool:', d['config'].get('pool'))
    except Exception as ex: print(f, 'parse failed', ex)
PY
