# row f4 (other map scales) on the GPU: new tests first, then the whole suite and a short bench
mkdir -p gpurun_out/r3a
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_round2.py tests/test_gpu_engine.py -q -m gpu -k "other_map_scales or any_size_kernel or other_observation_sizes or observation_size_errors" > gpurun_out/r3a/new.log 2>&1; echo "new rc $?"; grep -E "^E  |passed|failed" gpurun_out/r3a/new.log | head -40
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r3a/pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r3a/pytest.log
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r3a/bench.json 2> gpurun_out/r3a/bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3a/bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('val %.3e'%d['value'], 'ms %.4f'%d['ms_per_step'], 'frac %.3f step %.3f'%(r['frac'],r['step_frac']), 'e2e %.3e'%d['e2e']['value'])
PY
