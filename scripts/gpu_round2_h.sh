# 1 GPU, release build: the driver's round-end sequence (GPU tests, smoke, default bench, reference arm)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n1_d.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_ref_arm_d.json 2>&1
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02_bench_n1_d.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/r02_ref_arm_d.json').read().strip().splitlines()[-1])
print('value', b['value'], 'frac', b['roofline']['frac'], 'dram_frac', b['roofline']['dram_frac'], 'e2e', b['e2e']['value'], 'dirty', b['e2e']['dirty_inputs_only']['value'], 'exact', b['exact_math']['value'])
print({k:(round(v['frac'],3), v.get('dram_frac')) for k,v in b['effector_sets'].items()})
m=b['multi_gpu']; print('nbody', m['nbody_1024_sharded_worlds']['us_per_tick'], m['nbody_1024_sharded_worlds']['roofline']['pipe_frac'], m['nbody_1024_sharded_worlds']['saturated_batch']['roofline']['pipe_frac'], 'err' if 'error' in m else '')
print('cpu', b['cpu_baseline']['value'], r['value'], 'ratio e2e/ref', b['e2e']['value']/r['value'], 'clocks', b['clocks'])
PY
