# round-2 ncu evidence (one GPU).  Every command first exits 0 WITHOUT ncu; numbers printed under ncu are never bench values.
# The .ncu-rep files are summarised ON THE BOX (raw-page CSV) and removed: gpurun_out/ may not exceed 64 MiB.
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
C2="python bench.py --workload c2 --steps 8 --warmup 130 --no-cpu-baseline --no-extras"
C2S="python bench.py --workload c2 --steps 6 --warmup 6 --no-cpu-baseline --no-extras"
C5="python bench.py --workload c5 --steps 6 --warmup 30 --no-cpu-baseline --pool 512"
C3="python bench.py --workload c3 --steps 6 --warmup 30 --no-cpu-baseline --pool 512"
export_rep () {  # $1 = report stem
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2> /dev/null
  ncu -i $O/$1.ncu-rep --page details --csv > $O/$1_details.csv 2> /dev/null
  [ "$2" = keep ] || rm -f $O/$1.ncu-rep
}
timeout 300 $C2S > $O/c2_plain.json 2> $O/c2_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv $C2S > $O/ncu_launch.log 2>&1
echo "launch list rc $?"
# steady state: skip 125 steps x 3 kernels, then two steps' k_move, k_judge, k_render
timeout 300 $C2 > $O/c2_plain2.json 2> $O/c2_plain2.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(render|move|judge)' -s 375 -c 6 -o $O/prof_c2 -f $C2 > $O/ncu_c2.log 2>&1
echo "c2 full rc $?"; export_rep prof_c2 keep
timeout 300 $C5 > $O/c5_plain.json 2> $O/c5_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(render|move|judge)' -s 90 -c 3 -o $O/prof_c5 -f $C5 > $O/ncu_c5.log 2>&1
echo "c5 full rc $?"; export_rep prof_c5
timeout 300 $C3 > $O/c3_plain.json 2> $O/c3_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(render|move|judge)' -s 90 -c 3 -o $O/prof_c3 -f $C3 > $O/ncu_c3.log 2>&1
echo "c3 full rc $?"; export_rep prof_c3
timeout 300 python tools/ncu_variants.py > $O/variants_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none -k regex:'k_(render|move|fuse)' -o $O/prof_variants -f python tools/ncu_variants.py > $O/ncu_variants.log 2>&1
echo "variants rc $?"; export_rep prof_variants
du -sh $O; ls -la $O
