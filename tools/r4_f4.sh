# k_render_any: 8:3 block shortcut + class table -- parity first, then timeline and the row-f4 workload
O=gpurun_out/${1:-r4g}
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_round2.py tests/test_gpu_engine.py -q -m gpu -k "other_map_scales or any_size_kernel or other_observation_sizes or block_shortcut" > $O/new.log 2>&1; echo "new rc $?"; grep -E "^E  |passed|failed" $O/new.log | head -20
python tools/render_trace_any.py 256 4096 > $O/trace256.log 2>&1; tail -n 3 $O/trace256.log; python tools/render_trace_any.py 64 4096 > $O/trace64.log 2>&1; tail -n 3 $O/trace64.log
timeout 600 python bench.py --workload f4 --no-cpu-baseline --no-extras > $O/bench_f4.json 2> $O/bench_f4.err; echo "bench f4 rc $?"
python tools/bench_brief.py $O/bench_f4.json
