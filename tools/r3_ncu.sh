# ncu of the raw-RGB raster kernel and of k_render_any at size 256 (one GPU; each command exits 0 without ncu first)
mkdir -p gpurun_out/r3n
O=gpurun_out/r3n
C5="python bench.py --workload c5 --steps 6 --warmup 30 --no-cpu-baseline --pool 512"
F4="python bench.py --workload f4 --steps 4 --warmup 12 --no-cpu-baseline --pool 128"
export_rep () { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2> /dev/null; rm -f $O/$1.ncu-rep; }
timeout 300 $C5 > $O/c5_plain.json 2> $O/c5_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_render' -s 30 -c 2 -o $O/prof_c5 -f $C5 > $O/ncu_c5.log 2>&1
echo "c5 full rc $?"; export_rep prof_c5
timeout 300 $F4 > $O/f4_plain.json 2> $O/f4_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_render' -s 12 -c 2 -o $O/prof_f4 -f $F4 > $O/ncu_f4.log 2>&1
echo "f4 full rc $?"; export_rep prof_f4
ls -la $O
