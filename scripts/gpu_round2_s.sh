# 1 GPU, release build: memcheck of the round-2 kernels, full GPU tests, smoke, default bench, launch list of the bench command
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitizer_round2.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02_sanitizer_memcheck.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1_g.json 2> gpurun_out/r02_bench_n1_g.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n1_g.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02_bench_n1_g.json').read().strip().splitlines()[-1])
print('value', b['value'], 'frac', b['roofline']['frac'], 'e2e', b['e2e']['value'], 'exact', b['exact_math'], 'single', b.get('single_body'))
print({k:(round(v['frac'],3), round(v['us_per_tick'],1)) for k,v in b['effector_sets'].items()}, b.get('egm08_degree_64'))
m=b['multi_gpu']; print({k: (v.get('us_per_tick'), v.get('replicas')) for k,v in m.items() if isinstance(v, dict)})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_default_cmd.csv python bench.py --steps 50 --warmup 3 > gpurun_out/b_under_ncu.log 2>&1; wc -l gpurun_out/r02_launches_bench_default_cmd.csv
