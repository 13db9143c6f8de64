# usage: bash tools/ncu_render.sh  (on the GPU box; writes gpurun_out/)
mkdir -p gpurun_out
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --pool 512"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render -s 6 -c 2 -o gpurun_out/prof_render -f $CMD > gpurun_out/ncu_full.log 2>&1
timeout 300 $CMD > gpurun_out/plain3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sim -s 6 -c 2 -o gpurun_out/prof_sim -f $CMD > gpurun_out/ncu_full_sim.log 2>&1
ls -la gpurun_out
