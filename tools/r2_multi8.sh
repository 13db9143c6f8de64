# 8-GPU call: the scaling bench line (torchrun, NCCL) with the extra workloads (c3 = 65536 envs = BASELINE configs[2])
mkdir -p gpurun_out/r2m
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 100 --warmup 20 > gpurun_out/r2m/bench8.json 2> gpurun_out/r2m/bench8.err; echo "bench8 rc $?"; grep -v "^W\|^\[W\|^$" gpurun_out/r2m/bench8.err | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m/bench8.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value %.3e'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'venv %.3e'%d['e2e_vector_env']['value'])
for w in d.get('extra_workloads', []): print(w['name'], '%.3e'%w.get('value', 0), 'ms', w.get('ms_per_step'), w.get('skipped'), 'envs', w.get('config',{}).get('envs_total'))
PY
