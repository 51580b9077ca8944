python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_n1.err
for s in free rocket falcon9; do
  ncu --set full --clock-control none --import-source on -k regex:body_fast --launch-skip 6 --launch-count 1 -f -o gpurun_out/r02_$s python scripts/rocket_kernel_run.py $s > gpurun_out/ncu_$s.log 2>&1
  ncu -i gpurun_out/r02_$s.ncu-rep --page raw --csv > gpurun_out/r02_${s}_raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -15
