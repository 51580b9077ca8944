# 1 GPU: shape sweep + ncu of the persistent pair kernel (tuning build), then back to the release build
make -s -C elodin_b200/csrc TUNE=1 2>&1 | tail -2
python scripts/tune_world.py 2>&1 | tail -32
for M in 8 296; do
  ncu --set full --clock-control none --import-source on -k regex:graph_dense_world --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_world_M$M python scripts/tune_world.py ncu $M > gpurun_out/ncu_world_$M.log 2>&1
  ncu -i gpurun_out/r02_world_M$M.ncu-rep --page raw --csv > gpurun_out/r02_world_M${M}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_world_M$M.ncu-rep --page details > gpurun_out/r02_world_M${M}_details.txt 2>/dev/null
done
make -s -C elodin_b200/csrc 2>&1 | tail -2
ls gpurun_out | tail -8
