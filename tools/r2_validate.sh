# round-2 validation call: full GPU test-suite, smoke, default bench line (with extras), launch list
mkdir -p gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/pytest.log 2>&1; echo "pytest rc $?"; tail -15 gpurun_out/r2b/pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2b/smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r2b/smoke.log
timeout 900 python bench.py --steps 100 --warmup 20 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err; echo "bench rc $?"; tail -5 gpurun_out/r2b/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2b/bench.json').read().strip().splitlines()[-1])
    def show(n,w):
        r=w.get('roofline',{})
        print(n, 'val %.3e'%w['value'], 'ms %.4f'%w['ms_per_step'], 'render %.4f move %.4f judge %.4f'%(r.get('kernel_ms',0),r.get('sim_kernel_ms',0),r.get('judge_kernel_ms',0)), 'frac %.3f step_frac %.3f'%(r.get('frac',0),r.get('step_frac',0)), 'e2e %.3e venv %.3e pipe %.3e'%(w['e2e']['value'], w['e2e_vector_env']['value'], w['e2e_vector_env']['pipelined_value']))
    show('c2',d)
    for w in d.get('extra_workloads',[]):
        if 'value' in w: show(w['name'],w)
        else: print(w)
    print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('kind'))
except Exception as ex:
    print('parse failed', ex)
PY
