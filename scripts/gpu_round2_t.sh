# N GPUs (NG=2|4|8): default bench under torchrun
NG=${NG:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG > gpurun_out/r02_bench_n${NG}_c.json 2> gpurun_out/r02_bench_n${NG}_c.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_n${NG}_c.err
python - <<PY
import json
b=json.loads(open('gpurun_out/r02_bench_n${NG}_c.json').read().strip().splitlines()[-1])
print('value', b['value'], 'e2e', b['e2e']['value'], 'n', b['n_gpus'])
m=b['multi_gpu']
for k,v in m.items(): print(k, json.dumps(v)[:900])
PY
