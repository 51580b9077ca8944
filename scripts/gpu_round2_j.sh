# 1 GPU, release build: GPU tests, default bench, bench --configs, the opt-in TMA-ring variant next to the default kernel
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_n1_e.json 2> gpurun_out/r02_bench_n1_e.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n1_e.err
python bench.py --configs --steps 200 > gpurun_out/r02_bench_n1_configs.json 2> gpurun_out/r02_bench_n1_configs.err; echo "bench --configs rc=$?"; tail -c 300 gpurun_out/r02_bench_n1_configs.err
for cfg in 3 13 14; do B200_BODY_CFG=$cfg B200_NO_SPEC=$([ $cfg = 3 ] && echo 0 || echo 1) python bench.py --kernel-only --steps 200 --warmup 5 | tail -1 | sed "s/^/BODY_CFG=$cfg /"; done
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02_bench_n1_e.json').read().strip().splitlines()[-1])
print('value', b['value'], 'frac', b['roofline']['frac'], 'e2e', b['e2e']['value'])
print({k:(round(v['frac'],3), round(v['us_per_tick'],1)) for k,v in b['effector_sets'].items()}, b['telemetry_every_tick'])
c=json.loads(open('gpurun_out/r02_bench_n1_configs.json').read().strip().splitlines()[-1])['baseline_configs']
for k,v in c.items(): print(k, {a:(round(x,3) if isinstance(x,float) else x) for a,x in v.items() if a!='note'})
PY
