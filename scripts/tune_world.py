"""Shape sweep of graph_dense_world_kernel (tuning build; B200_WORLD_CFG re-read per launch): n-body 1024 at
M = 8 / 64 / 296 worlds, us per tick, FP64-pipe fraction, max difference of the positions against cfg 0 after 3 ticks.
`python scripts/tune_world.py ncu <M>` runs 6 ticks of the default shape for an ncu capture."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import elodin_b200 as el
from elodin_b200 import _lib
from elodin_b200.executor import WORLD_POS

N = 1024
rng = np.random.default_rng(7)
g = el.GravityEdges("softened", k_squared=2.9591220828e-4 / 86400.0 ** 2, softening=1e-10, edges=el.all_pairs_edges(N))

def world(Mw):
    p = np.zeros((Mw, N, 7)); p[..., 3] = 1.0; p[..., 4:] = rng.uniform(-30, 30, (Mw, N, 3))
    v = np.zeros((Mw, N, 6)); v[..., 3:] = rng.normal(0, 1e-4, (Mw, N, 3))
    m = 10 ** rng.uniform(-10, -3, (Mw, N)); m[:, 0] = 1.0
    I = np.zeros((Mw, N, 7)); I[..., :3] = m[..., None]; I[..., 6] = m
    return p, v, I

if sys.argv[1:2] == ["ncu"]:
    Mw = int(sys.argv[2])
    p, v, I = world(Mw)
    ex = el.B200Exec(N, Mw, 3600.0, None, [g], "rk4", "fast")
    ex.set_state(p, v, I)
    ex.step(6, sync=True)
    sys.exit(0)

L = _lib.lib()
fp64 = L.b200_probe_fp64_gflops(0, 20000)
st = torch.cuda.Stream()
rows = []
for Mw in (8, 64, 296):
    p, v, I = world(Mw)
    ref = None
    for cfg in [int(c) for c in os.environ.get("WORLD_CFGS", "0,1,2,3,4,5,6,7,8").split(",")]:
        os.environ["B200_WORLD_CFG"] = str(cfg)
        ex = el.B200Exec(N, Mw, 3600.0, None, [g], "rk4", "fast")
        ex.set_stream(st.cuda_stream)
        ex.set_state(p, v, I)
        with torch.cuda.stream(st):
            ex.step(3); torch.cuda.synchronize()
            got = ex.download(WORLD_POS)[..., 4:]
            if ref is None:
                ref = got
            err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
            best = 1e30
            ticks = 60 if Mw <= 64 else 12
            for _ in range(3):
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record(st); ex.step(ticks); b.record(st); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) / ticks)
        ex.close()
        pair = 3.0 * N * (N - 1) * Mw / (best * 1e-3)
        rows.append({"worlds": Mw, "cfg": cfg, "us_per_tick": best * 1e3, "pair_evals_per_s": pair,
                     "pipe_frac": pair * 18.0 / (fp64 * 1e9 / 2.0), "max_rel_vs_cfg0": err})
        print(json.dumps(rows[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"fp64_probe_GFLOPs": fp64, "rows": rows}, open("gpurun_out/tune_world.json", "w"), indent=1)
