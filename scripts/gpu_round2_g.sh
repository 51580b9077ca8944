# 1 GPU (release build): fused pair kernel on/off, tests with the fused path forced, e2e range-size sweep,
# ncu of the cube-sat signature kernel, full tests, bench
python scripts/tune_fused.py 2>&1 | tail -14
B200_NBODY_FUSED=3 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "nbody or world_resident or fused or pipelined" 2>&1 | tail -3
python scripts/tune_e2e.py 2>&1 | tail -12
ncu --set full --clock-control none --import-source on -k regex:body_fast --launch-skip 6 --launch-count 1 -f -o gpurun_out/r02_cube_sat python scripts/rocket_kernel_run.py cube_sat > gpurun_out/ncu_cube_sat.log 2>&1
ncu -i gpurun_out/r02_cube_sat.ncu-rep --page raw --csv > gpurun_out/r02_cube_sat_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_cube_sat.ncu-rep --page details > gpurun_out/r02_cube_sat_details.txt 2>/dev/null
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n1_c.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02_bench_n1_c.json').read().strip().splitlines()[-1])
print('value', b['value'], 'e2e', b['e2e']['value'], 'exact', b['exact_math']['value'])
print({k:(v['frac'], v['us_per_tick']) for k,v in b['effector_sets'].items()})
m=b['multi_gpu']['nbody_1024_sharded_worlds']; print('nbody', m['us_per_tick'], m['roofline']['pipe_frac'], m['saturated_batch']['roofline']['pipe_frac'])
PY
