"""Alternating traversal direction (B200_SNAKE=1, default) vs always forward (=0) for the FAST body kernel:
free / rocket / falcon9 at 2^22 worlds, 1 tick per launch, 100 back-to-back launches.  One subprocess per setting
(the switch is read once per process).  Results: profiles/r02_tune_snake.txt."""
import os, subprocess, sys
for v in ("0", "1"):
    out = subprocess.run([sys.executable, "scripts/effector_perf.py"], env=dict(os.environ, B200_SNAKE=v), capture_output=True, text=True)
    print("B200_SNAKE=" + v); print(out.stdout.strip()); print(out.stderr.strip()[-300:])
