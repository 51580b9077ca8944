"""Small graph worlds (N <= 32 bodies, all-pairs gravity): the one-warp-per-world multi-tick kernel vs the
generic gravity-launch + body-launch route (B200_SMALL_WORLD=0).  Prints us/tick and entity-steps/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import elodin_b200 as el

cases = [(3, 1, 2000), (3, 1 << 18, 64), (8, 1 << 16, 64), (16, 1 << 14, 64), (32, 1 << 13, 32)]
if os.environ.get("SW_QUICK"):
    cases = [(3, 1 << 18, 64), (8, 1 << 16, 64), (32, 1 << 13, 32)]
rng = np.random.default_rng(3)
tag = "small_world=" + os.environ.get("B200_SMALL_WORLD", "1") + " cfg=" + os.environ.get("B200_SMALL_WORLD_CFG", "2")
for math in ("fast", "exact"):
    for N, M, T in cases:
        p = np.zeros((M, N, 7)); p[..., 3] = 1.0; p[..., 4:] = rng.uniform(-30, 30, (M, N, 3))
        v = np.zeros((M, N, 6)); v[..., 3:] = rng.normal(0, 1e-3, (M, N, 3))
        m = 10 ** rng.uniform(-3, 0, (M, N))
        I = np.zeros((M, N, 7)); I[..., :3] = m[..., None]; I[..., 6] = m
        g = el.GravityEdges("softened", k_squared=1e-3, softening=1e-6, edges=el.all_pairs_edges(N))
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            with el.B200Exec(N, M, 0.01, None, [g], "rk4", math, max_fused_ticks=32) as ex:
                ex.set_stream(st.cuda_stream); ex.set_state(p, v, I)
                ex.step(8); torch.cuda.synchronize()
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record(st); ex.step(T); b.record(st); torch.cuda.synchronize()
                us = a.elapsed_time(b) * 1e3 / T
                print(f"{tag} {math:5s} N={N:3d} M={M:7d}: {us:9.2f} us/tick  {N * M / us * 1e6:.3e} entity-steps/s", flush=True)
