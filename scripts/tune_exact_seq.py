"""EXACT body kernel: compiled effector sequence (B200_EXACT_CFG=3, default) vs the run-time interpreter (=12),
2^20 worlds, RK4, for the free body, the rocket Monte-Carlo list and the falcon9 list.  One subprocess per setting
(the variable is read once in a release build).  Also checks that both produce the same bits."""
import json, os, subprocess, sys
code = r'''
import sys, os, json, hashlib
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import elodin_b200 as el, bench
from elodin_b200.executor import WORLD_POS, WORLD_VEL, FORCE
M = 1 << 20
pos, vel, ine = bench.synth_world(M, 1)
rng = np.random.default_rng(0)
sets = {"free": ([], {}),
        "rocket": ([el.GravityConst(), el.ThrustBody((-1.0, 0, 0), "thrust"), el.DragQuadratic(0.6, 0.01, "wind")],
                   {"thrust": rng.uniform(50, 100, (M, 1, 1)), "wind": rng.normal(0, 1, (M, 1, 3))}),
        "falcon9": ([el.GravityFrame(), el.WrenchBody("body_wrench", "linear_first")], {"body_wrench": rng.normal(0, 1e3, (M, 1, 6))})}
st = torch.cuda.Stream()
for name, (effs, cols) in sets.items():
    p = pos.copy()
    if name == "falcon9": p[..., 4:] += np.array([6.4e6, 0, 0])
    ex = el.B200Exec(1, M, 1e-3, None, effs, "rk4", "exact")
    ex.set_stream(st.cuda_stream); ex.set_state(p, vel, ine, **cols)
    with torch.cuda.stream(st):
        ex.step(3); torch.cuda.synchronize()
        h = hashlib.sha1(ex.download(WORLD_POS).tobytes() + ex.download(WORLD_VEL).tobytes() + ex.download(FORCE).tobytes()).hexdigest()[:12]
        best = 1e30
        for _ in range(3):
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record(st); ex.step(10); b.record(st); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 10)
    ex.close()
    print(json.dumps({"cfg": os.environ.get("B200_EXACT_CFG", "3"), "set": name, "us_per_tick": best * 1e3, "entity_steps_per_s": M / (best * 1e-3), "sha1_after_3_ticks": h}), flush=True)
'''
# B200_EXACT_SEQ_CFG (tuning build only): launch bounds of the compiled sequences — 0 = 128 x 4 (default), 1 = 128 x 3,
# 2 = 128 x 2, 3 = 64 x 6, 4 = 128 x 5
for cfg, seq_cfg in (("3", "0"), ("12", "0")) + tuple(("3", c) for c in sys.argv[1:]):
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B200_EXACT_CFG=cfg, B200_EXACT_SEQ_CFG=seq_cfg), capture_output=True, text=True)
    print("seq_cfg", seq_cfg); print(out.stdout.strip()); print(out.stderr.strip()[-300:])
