ncu --set full --clock-control none --import-source on -k regex:body_exact --launch-skip 3 --launch-count 1 -o gpurun_out/r02_exact python scripts/exact_kernel_run.py > /dev/null 2>&1
ncu -i gpurun_out/r02_exact.ncu-rep --page details --csv > gpurun_out/r02_exact_details.csv 2>/dev/null
ncu -i gpurun_out/r02_exact.ncu-rep --page raw --csv > gpurun_out/r02_exact_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_exact.ncu-rep --page source --csv > gpurun_out/r02_exact_source.csv 2>/dev/null
ls -la gpurun_out/r02_exact*
