# 2-GPU call: the two-device engine test and a 2-rank bench line (torchrun, NCCL)
mkdir -p gpurun_out/r2m
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k two_engines > gpurun_out/r2m/pytest2.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2m/pytest2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/r2m/bench2.json 2> gpurun_out/r2m/bench2.err; echo "bench2 rc $?"; tail -c 600 gpurun_out/r2m/bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m/bench2.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value %.3e'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'])
for w in d.get('extra_workloads', []): print(w['name'], '%.3e'%w.get('value', 0), w.get('skipped'))
PY
