# 1 GPU: tuning sweeps of the secondary kernels (tuning build), then release build: full GPU tests, bench, reference arm
make -s -C elodin_b200/csrc TUNE=1 2>&1 | tail -2
python scripts/tune_misc.py 2>&1 | tail -40
make -s -C elodin_b200/csrc 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_n1_b.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_ref_arm_b.json 2>&1
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02_bench_n1_b.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/r02_ref_arm_b.json').read().strip().splitlines()[-1])
print('cpu_baseline (gpu arm)', b['cpu_baseline']['value'], b['cpu_baseline']['cores'], '| reference arm', r['value'], r['cpu_baseline']['cores'], '| ratio', b['cpu_baseline']['value']/r['value'])
print('value', b['value'], 'e2e', b['e2e']['value'], 'exact', b['exact_math']['value'], 'nbody', b['multi_gpu']['nbody_1024_sharded_worlds']['us_per_tick'], b['multi_gpu']['nbody_1024_sharded_worlds']['roofline']['pipe_frac'])
PY
