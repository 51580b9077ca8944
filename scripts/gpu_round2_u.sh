python scripts/tune_e2e_taper.py 2>&1 | grep -v "^$" | tee gpurun_out/tune_e2e_taper.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "invoke_batch or pipelined or null" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_kernel_only.csv python bench.py --kernel-only --steps 200 --warmup 5 > gpurun_out/b_under_ncu.log 2>&1; wc -l gpurun_out/r02_launches_bench_kernel_only.csv
