# 1 GPU: EXACT effector divisions grouped — parity (release build), then launch bounds of the compiled sequences (tuning build on the box)
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python scripts/tune_exact_seq.py 2>&1 | grep "^{" | head -3
cp elodin_b200/libb200_sixdof.so /tmp/release.so
make -C elodin_b200/csrc TUNE=1 2>&1 | grep -E "error" | head; python scripts/tune_exact_seq.py 1 2 3 4 2>&1 | grep "^{\|seq_cfg" | tee gpurun_out/tune_exact_seq_cfg.txt
cp /tmp/release.so elodin_b200/libb200_sixdof.so
