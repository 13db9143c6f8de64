# round-2 first GPU call: sanitizer on smoke(), fuzz, RGB (c5) ncu --set full
mkdir -p gpurun_out/r2a
SM='import __graft_entry__ as g; g.smoke()'
timeout 300 python -c "$SM" > gpurun_out/r2a/smoke.log 2>&1; echo "smoke rc $?"
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/r2a/$tool.log python -c "$SM" > gpurun_out/r2a/$tool.out 2>&1; echo "$tool rc $?"
  tail -3 gpurun_out/r2a/$tool.log
done
timeout 400 python tools/gpu_fuzz.py 6 1 250 > gpurun_out/r2a/fuzz.log 2>&1; echo "fuzz rc $?"; tail -3 gpurun_out/r2a/fuzz.log
C5="python bench.py --workload c5 --steps 30 --warmup 10 --no-cpu-baseline --pool 256"
timeout 400 $C5 > gpurun_out/r2a/c5.json 2> gpurun_out/r2a/c5.err && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render -s 10 -c 2 -o gpurun_out/r2a/prof_render_rgb -f $C5 > gpurun_out/r2a/ncu_rgb.log 2>&1
echo "c5 rc $?"; cat gpurun_out/r2a/c5.json | cut -c1-600
ls -la gpurun_out/r2a
