# 1 GPU: streaming trajectory stores, grouped Newton-pair divisions — parity, telemetry timing, the BASELINE configs
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python scripts/telemetry_perf.py 2>&1 | tail -3
python bench.py --configs --steps 200 > gpurun_out/r02_bench_n1_configs_b.json 2> gpurun_out/r02_bench_n1_configs_b.err; echo "bench --configs rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02_bench_n1_configs_b.json').read().strip().splitlines()[-1])
print('telemetry', b['telemetry_every_tick']['us_per_tick'], b['telemetry_every_tick']['frac'])
for k,v in b['baseline_configs'].items(): print(k, {a:(round(x,3) if isinstance(x,float) else x) for a,x in v.items() if a!='note'})
PY
