"""One to four n-body worlds of 1024 bodies (FAST, RK4): the small-grid fused kernel vs the world-resident pair kernel with
the integration fused in, spread over the SMs (B200_NBODY_WORLD_MIN=0 forces it; B200_NBODY_WORLD_ROUNDS = rounds of items
per warp).  One subprocess per setting."""
import json, os, subprocess, sys
code = r'''
import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import elodin_b200 as el
from elodin_b200.executor import WORLD_POS
rng = np.random.default_rng(7)
st = torch.cuda.Stream()
for N, Mw in ((64, 1), (64, 16), (200, 1), (256, 1), (256, 4), (512, 1), (512, 2), (1024, 1), (1024, 2), (1024, 3), (1024, 4)):
    g = el.GravityEdges("softened", k_squared=2.9591220828e-4 / 86400.0 ** 2, softening=1e-10, edges=el.all_pairs_edges(N))
    p = np.zeros((Mw, N, 7)); p[..., 3] = 1.0; p[..., 4:] = rng.uniform(-30, 30, (Mw, N, 3))
    v = np.zeros((Mw, N, 6)); v[..., 3:] = rng.normal(0, 1e-4, (Mw, N, 3))
    m = 10 ** rng.uniform(-10, -3, (Mw, N)); m[:, 0] = 1.0
    I = np.zeros((Mw, N, 7)); I[..., :3] = m[..., None]; I[..., 6] = m
    ex = el.B200Exec(N, Mw, 3600.0, None, [g], "rk4", "fast")
    ex.set_stream(st.cuda_stream); ex.set_state(p, v, I)
    with torch.cuda.stream(st):
        ex.step(5); torch.cuda.synchronize()
        chk = float(np.sum(ex.download(WORLD_POS)[..., 4:]))
        best = 1e30
        for _ in range(3):
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record(st); ex.step(400); b.record(st); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 400)
    ex.close()
    print(json.dumps({"world_min": os.environ.get("B200_NBODY_WORLD_MIN", "444"), "rounds": os.environ.get("B200_NBODY_WORLD_ROUNDS", "1"), "spread": os.environ.get("B200_NBODY_WORLD_SPREAD", "16"),
                      "N": N, "worlds": Mw, "us_per_tick": round(best * 1e3, 2), "checksum_after_5_ticks": chk}), flush=True)
'''
for wm, rounds, spread in (("0", "1", "16"), ("0", "1", "11"), ("0", "1", "8"), ("0", "1", "4")):
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B200_NBODY_WORLD_MIN=wm, B200_NBODY_WORLD_ROUNDS=rounds, B200_NBODY_WORLD_SPREAD=spread), capture_output=True, text=True)
    print(out.stdout.strip()); print(out.stderr.strip()[-400:])
