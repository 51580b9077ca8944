# 2-GPU checks of the library's NCCL paths + the N=2 bench line + the CPU reference arm on the same box
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_gpus or two_devices" 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/r02_bench_n2.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_ref_arm.json 2>&1; tail -c 600 gpurun_out/r02_ref_arm.json
