import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = """Round-1 sweep (profiles/r01_tuning.md).  Since round 2: scripts/tune_misc.py (shapes, tuning build) and scripts/tune_exact_seq.py (compiled effector sequences vs the interpreter).

import sys; sys.path.insert(0, %r)
import numpy as np, torch, elodin_b200 as el, bench
M = 1 << 20
pos, vel, ine = bench.synth_world(M, 1)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    ex = el.B200Exec(1, M, 1e-3, None, [], "rk4", "exact"); ex.set_stream(st.cuda_stream); ex.set_state(pos, vel, ine)
    ex.step(3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(st); ex.step(10); e1.record(st); torch.cuda.synchronize()
    print("exact entity-steps/s %%.3e" %% (M * 10 / (e0.elapsed_time(e1) * 1e-3)))
""" % root
for cfg in ["0", "1", "2", "3"]:
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B200_EXACT_CFG=cfg), capture_output=True, text=True)
    print("exact cfg", cfg, out.stdout.strip(), out.stderr.strip()[-200:])
