# A/B timing probes on c2 (one GPU): shipped / serial judge / identity CTA order, then the ncu evidence pass
mkdir -p gpurun_out/r2d
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r2d/pytest.log
B="python bench.py --workload c2 --steps 300 --warmup 60 --no-cpu-baseline --no-extras"
for v in 0 2 32 0 2 32; do
  timeout 300 $B --debug-flags $v 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('flags $v', 'ms %.4f'%d['ms_per_step'], 'render %.4f move %.4f judge %.4f'%(r['kernel_ms'],r['sim_kernel_ms'],r['judge_kernel_ms']), 'frac %.3f step %.3f'%(r['frac'],r['step_frac']), 'e2e %.3e venv %.3e'%(d['e2e']['value'], d['e2e_vector_env']['value']), 'blocks', ['%.2f'%b for b in d['blocks_ms']])"
done
bash tools/r2_ncu.sh
