mkdir -p gpurun_out/r2e
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k c4 > gpurun_out/r2e/c4.log 2>&1; echo "c4 rc $?"; grep -E "^E  " gpurun_out/r2e/c4.log | head -20; tail -3 gpurun_out/r2e/c4.log
bash tools/r2_ncu.sh
