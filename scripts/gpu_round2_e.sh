# 1 GPU: more pair-kernel shapes (tuning build), then release build: sanitizer on the round-2 kernels, GPU tests,
# launch list of the default bench command
make -s -C elodin_b200/csrc TUNE=1 2>&1 | tail -2
WORLD_CFGS=0,9,10,11,12,13,14 python scripts/tune_world.py 2>&1 | tail -24
cp gpurun_out/tune_world.json gpurun_out/tune_world_b.json
make -s -C elodin_b200/csrc 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitizer_round2.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench_default_cmd.csv python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err; echo "ncu bench rc=$?"
wc -l gpurun_out/r02_launches_bench_default_cmd.csv
