# 2 GPUs: the two-rank library tests (NCCL gather, row shards over NCCL and through peer windows), then the three routes timed
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_gpus" 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 scripts/row_shard_perf.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -8
