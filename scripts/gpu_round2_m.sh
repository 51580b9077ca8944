# 1 GPU: EGM08 term-stream kernel — parity, timing at degree 64 (2^16 and 2^18 worlds) and 8
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "egm08 or smoke or j2" 2>&1 | tail -3
python scripts/egm08_perf.py 64 16; python scripts/egm08_perf.py 64 18; python scripts/egm08_perf.py 8 20
