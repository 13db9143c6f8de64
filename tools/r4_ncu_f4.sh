O=gpurun_out/r4f
mkdir -p $O
F4="python bench.py --workload f4 --steps 4 --warmup 12 --no-cpu-baseline --no-extras --pool 128"
timeout 300 $F4 > $O/f4_plain.json 2> $O/f4_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_render' -s 12 -c 1 -o $O/prof_f4 -f $F4 > $O/ncu_f4.log 2>&1
echo "f4 ncu rc $?"; ls -la $O
