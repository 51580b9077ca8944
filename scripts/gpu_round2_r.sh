python scripts/tune_nbody_small.py 2>&1 | tee gpurun_out/tune_nbody_small3.txt | cut -c1-110
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "world_resident or nbody or graph" 2>&1 | tail -3
