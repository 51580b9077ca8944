"""n-body 1024: separate gravity + body launches vs the pair kernel with the integration fused in
(B200_NBODY_FUSED=3, read once per process -> one subprocess per setting)."""
import json, os, subprocess, sys
code = r'''
import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import elodin_b200 as el
from elodin_b200 import _lib
from elodin_b200.executor import WORLD_POS
N = 1024
rng = np.random.default_rng(7)
g = el.GravityEdges("softened", k_squared=2.9591220828e-4 / 86400.0 ** 2, softening=1e-10, edges=el.all_pairs_edges(N))
fp64 = _lib.lib().b200_probe_fp64_gflops(0, 20000)
st = torch.cuda.Stream()
for Mw in (8, 64, 296):
    p = np.zeros((Mw, N, 7)); p[..., 3] = 1.0; p[..., 4:] = rng.uniform(-30, 30, (Mw, N, 3))
    v = np.zeros((Mw, N, 6)); v[..., 3:] = rng.normal(0, 1e-4, (Mw, N, 3))
    m = 10 ** rng.uniform(-10, -3, (Mw, N)); m[:, 0] = 1.0
    I = np.zeros((Mw, N, 7)); I[..., :3] = m[..., None]; I[..., 6] = m
    ex = el.B200Exec(N, Mw, 3600.0, None, [g], "rk4", "fast")
    ex.set_stream(st.cuda_stream); ex.set_state(p, v, I)
    with torch.cuda.stream(st):
        ex.step(5); torch.cuda.synchronize()
        chk = float(np.sum(ex.download(WORLD_POS)[..., 4:]))
        ticks = 100 if Mw <= 64 else 20
        best = 1e30
        for _ in range(3):
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record(st); ex.step(ticks); b.record(st); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / ticks)
    launches = ex.timings()["kernel_launches"]
    ex.close()
    pair = 3.0 * N * (N - 1) * Mw / (best * 1e-3)
    print(json.dumps({"fused": os.environ.get("B200_NBODY_FUSED", "1"), "worlds": Mw, "us_per_tick": best * 1e3, "pipe_frac": pair * 18.0 / (fp64 * 1e9 / 2.0),
                      "checksum_after_5_ticks": chk, "launches": launches}), flush=True)
'''
for fused in ("1", "3"):
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B200_NBODY_FUSED=fused), capture_output=True, text=True)
    print(out.stdout.strip()); print(out.stderr.strip()[-400:])
