# 1 GPU: vectorised per-tick trajectory store of the pair kernel — parity, timing
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python scripts/telemetry_perf.py 2>&1 | tail -3
