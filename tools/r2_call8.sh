mkdir -p gpurun_out/r2i gpurun_out/r2n
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/r2i/pytest.log 2>&1; echo "pytest rc $?"; grep -E "^E  " gpurun_out/r2i/pytest.log | head -12; tail -3 gpurun_out/r2i/pytest.log
bash tools/r2_ncu.sh
