# 1 GPU: shared-divisor divisions in EXACT mode — self-test, every EXACT parity test, timing of the three effector sets
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "shared_divisor" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/tune_exact_seq.py 2>&1 | tail -12
