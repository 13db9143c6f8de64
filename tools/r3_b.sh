# direct RGB path + row f4 bench workload
mkdir -p gpurun_out/r3b
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_round2.py tests/test_gpu_engine.py -q -m gpu -k "make_env_at_other or raw_rgb or rgb" > gpurun_out/r3b/new.log 2>&1; echo "new rc $?"; grep -E "^E  |passed|failed" gpurun_out/r3b/new.log | head -40
for w in c5 f4; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --steps 200 --warmup 100 > gpurun_out/r3b/bench_$w.json 2> gpurun_out/r3b/bench_$w.err; echo "bench $w rc $?"
done
python - <<'PY'
import json
for w in ('c5','f4'):
    try:
        d=json.loads(open(f'gpurun_out/r3b/bench_{w}.json').read().strip().splitlines()[-1]); r=d['roofline']
        print(w, 'val %.3e'%d['value'], 'ms %.4f'%d['ms_per_step'], 'kernel %.4f move %.4f'%(r['kernel_ms'], r['sim_kernel_ms']), 'frac %.3f step %.3f'%(r['frac'],r['step_frac']), 'e2e %.3e'%d['e2e']['value'], d['config'].get('workload','')[:60])
    except Exception as ex: print(w, 'failed', ex)
PY
tail -3 gpurun_out/r3b/bench_f4.err
