"""ncu target: small_world_kernel only (FAST / EXACT, N = 3 and N = 8), 16 ticks per launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import elodin_b200 as el

rng = np.random.default_rng(0)
for Nw, Mw in ((3, 1 << 18), (8, 1 << 16)):
    p = np.zeros((Mw, Nw, 7)); p[..., 3] = 1.0; p[..., 4:] = rng.uniform(-30, 30, (Mw, Nw, 3))
    v = np.zeros((Mw, Nw, 6)); v[..., 3:] = rng.normal(0, 1e-3, (Mw, Nw, 3))
    m = 10 ** rng.uniform(-3, 0, (Mw, Nw)); I = np.zeros((Mw, Nw, 7)); I[..., :3] = m[..., None]; I[..., 6] = m
    g = el.GravityEdges("softened", k_squared=1e-3, softening=1e-6, edges=el.all_pairs_edges(Nw))
    for math in ("fast", "exact"):
        with el.B200Exec(Nw, Mw, 0.01, None, [g], "rk4", math, max_fused_ticks=16) as ex:
            ex.set_state(p, v, I); ex.step(32, sync=True)
print("done")
