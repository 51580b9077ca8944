# 8 GPUs of one node: the driver's scaling commands (N = 8, 4) + the CPU reference arm; nvidia-smi topology for the record
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_gpus" 2>&1 | tail -2
for N in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench n$N rc=$?"; tail -c 300 gpurun_out/r02_bench_n$N.err | tail -3
done
python bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_ref_arm_8gpu_box.json 2>&1
python - <<'PY'
import json
for N in (8, 4):
    d=json.loads(open(f'gpurun_out/r02_bench_n{N}.json').read().strip().splitlines()[-1])
    e=d['e2e']; m=d['multi_gpu']
    print(N, 'value', d['value'], 'e2e', e['value'], 'ms/call', e['ms_per_call'], 'pcie', e['pcie']['h2d_GBps_rank0'], e['pcie']['d2h_GBps_rank0'], 'bound', e['pcie']['copy_bound_ms_per_call'])
    print('   nbody8', m['nbody_1024_sharded_worlds']['us_per_tick'], 'single', m['nbody_1024_single_world']['replicas']['us_per_tick'], m['nbody_1024_single_world']['row_shards'].get('us_per_tick'), 'big', m['nbody_8192_single_world']['replicas']['us_per_tick'], m['nbody_8192_single_world']['row_shards'].get('us_per_tick'))
    f=m['falcon9_mc_rollouts']; print('   falcon9', f['seconds_steps'], f['seconds_gather'], f['gather']['recv_GBps_per_gpu'])
PY
