# validation with the driver's own commands + ncu evidence for k_render_any (row f4)
O=gpurun_out/r4z
mkdir -p $O
timeout 1800 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc $?"; grep -E "^E  " $O/pytest.log | head; tail -3 $O/pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke.log 2>&1; echo "smoke rc $?"; tail -1 $O/smoke.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err ) 2>&1 | grep real; cut -c1-200 $O/bench_ref.json
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_20.json 2> $O/bench_20.err ) 2>&1 | grep real; echo "bench20 rc $?"
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2>&1 | grep real; echo "bench default rc $?"
F4="python bench.py --workload f4 --steps 4 --warmup 12 --no-cpu-baseline --pool 128"
timeout 300 $F4 > $O/f4_plain.json 2> $O/f4_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_render' -s 12 -c 2 -o $O/prof_f4 -f $F4 > $O/ncu_f4.log 2>&1
echo "f4 ncu rc $?"; ncu -i $O/prof_f4.ncu-rep --page raw --csv > $O/prof_f4_raw.csv 2> /dev/null; rm -f $O/prof_f4.ncu-rep
python - <<'PY'
import json
for f in ('bench_20','bench_default'):
    try:
        d=json.loads(open(f'gpurun_out/r4z/{f}.json').read().strip().splitlines()[-1]); r=d['roofline']
        print(f, 'val %.3e'%d['value'], 'ms %.4f'%d['ms_per_step'], 'frac %.3f step %.3f'%(r['frac'],r['step_frac']), 'e2e %.3e venv %.3e'%(d['e2e']['value'], d['e2e_vector_env']['value']), 'clocks', d['clocks'])
        for w in d.get('extra_workloads',[]): print('   ', w['name'], '%.3e'%w.get('value',0), 'frac', w.get('roofline',{}).get('frac'), w.get('skipped'), w.get('error'), (w.get('cpu_baseline') or {}).get('value'))
        if d.get('cpu_baseline'): print('    cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
    except Exception as ex: print(f, 'parse failed', ex)
PY
