# k_render_any v2 (swizzled strips, fast rotate, uniform-footprint resize): parity, then the f4 bench
mkdir -p gpurun_out/r3c
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_round2.py tests/test_gpu_engine.py -q -m gpu -k "other_map_scales or any_size_kernel or other_observation_sizes or observation_size_errors or raw_rgb" > gpurun_out/r3c/new.log 2>&1; echo "new rc $?"; grep -E "^E  |passed|failed" gpurun_out/r3c/new.log | head -30
timeout 600 python bench.py --workload f4 --no-cpu-baseline --steps 100 --warmup 50 > gpurun_out/r3c/bench_f4.json 2> gpurun_out/r3c/bench_f4.err; echo "bench f4 rc $?"
python - <<'PY'
import json
for w in ('f4',):
    try:
        d=json.loads(open(f'gpurun_out/r3c/bench_{w}.json').read().strip().splitlines()[-1]); r=d['roofline']
        print(w, 'val %.3e'%d['value'], 'ms %.4f'%d['ms_per_step'], 'kernel %.4f move %.4f'%(r['kernel_ms'], r['sim_kernel_ms']), 'frac %.3f step %.3f'%(r['frac'],r['step_frac']), 'e2e %.3e'%d['e2e']['value'])
    except Exception as ex: print(w, 'failed', ex)
PY
