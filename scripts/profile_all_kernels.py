"""Exercise every kernel of libb200_sixdof.so once at a representative size (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import elodin_b200 as el
import bench
from elodin_b200.executor import WORLD_POS

rng = np.random.default_rng(0)
M = 1 << 20
pos, vel, ine = bench.synth_world(M, 1)
# body_exact_kernel (RK4) + aos_to_soa / soa_to_aos at 2^20 bodies
with el.B200Exec(1, M, 1e-3, None, [], "rk4", "exact") as ex:
    ex.set_state(pos, vel, ine); ex.step(3, sync=True); ex.download(WORLD_POS)
# body_fast_kernel with the rocket effector set (gravity + thrust + drag) and with the falcon9 set
effs = [el.GravityConst(), el.ThrustBody((-1.0, 0, 0), "thrust"), el.DragQuadratic(0.6, 0.01, "wind")]
with el.B200Exec(1, M, 1e-3, None, effs, "rk4", "fast") as ex:
    ex.set_state(pos, vel, ine, thrust=rng.uniform(50, 100, (M, 1, 1)), wind=rng.normal(0, 1, (M, 1, 3))); ex.step(3, sync=True)
# semi-implicit fast / exact
for math in ("fast", "exact"):
    with el.B200Exec(1, M, 1e-3, None, [], "semi_implicit", math) as ex:
        ex.set_state(pos, vel, ine); ex.step(3, sync=True)
# n-body: graph_dense_fast_kernel (split and unsplit), graph_dense_kernel (exact), graph_csr_kernel
N = 1024
for Mw, math in ((1, "fast"), (8, "fast"), (8, "exact")):
    p = np.zeros((Mw, N, 7)); p[..., 3] = 1.0; p[..., 4:] = rng.uniform(-30, 30, (Mw, N, 3))
    v = np.zeros((Mw, N, 6)); m = 10 ** rng.uniform(-10, -3, (Mw, N))
    I = np.zeros((Mw, N, 7)); I[..., :3] = m[..., None]; I[..., 6] = m
    g = el.GravityEdges("softened", k_squared=2.9591220828e-4 / 86400.0 ** 2, softening=1e-10, edges=el.all_pairs_edges(N))
    with el.B200Exec(N, Mw, 3600.0, None, [g], "rk4", math) as ex:
        ex.set_state(p, v, I); ex.step(3, sync=True)
edges = np.array([(i, j) for i in range(0, 512, 2) for j in rng.permutation(512)[:16] if i != j])
p = np.zeros((64, 512, 7)); p[..., 3] = 1.0; p[..., 4:] = rng.uniform(-30, 30, (64, 512, 3))
I = np.ones((64, 512, 7))
with el.B200Exec(512, 64, 0.01, None, [el.GravityEdges("softened", k_squared=1e-3, softening=1e-6, edges=edges)], "rk4", "fast") as ex:
    ex.set_state(p, np.zeros((64, 512, 6)), I); ex.step(3, sync=True)
# small graph worlds: small_world_kernel (one warp per floor(32/N) worlds, 16 ticks per launch), FAST and EXACT
for Nw, Mw in ((3, 1 << 18), (8, 1 << 16)):
    p = np.zeros((Mw, Nw, 7)); p[..., 3] = 1.0; p[..., 4:] = rng.uniform(-30, 30, (Mw, Nw, 3))
    v = np.zeros((Mw, Nw, 6)); v[..., 3:] = rng.normal(0, 1e-3, (Mw, Nw, 3))
    m = 10 ** rng.uniform(-3, 0, (Mw, Nw)); I = np.zeros((Mw, Nw, 7)); I[..., :3] = m[..., None]; I[..., 6] = m
    g = el.GravityEdges("softened", k_squared=1e-3, softening=1e-6, edges=el.all_pairs_edges(Nw))
    for math in ("fast", "exact"):
        with el.B200Exec(Nw, Mw, 0.01, None, [g], "rk4", math, max_fused_ticks=16) as ex:
            ex.set_state(p, v, I); ex.step(32, sync=True)
# full-telemetry trajectory ring: body_fast_kernel<.., TRAJ = true> recording 25 planes every 4th tick + its read-back
with el.B200Exec(1, M, 1e-3, None, [], "rk4", "fast", max_fused_ticks=16, trajectory_every=4, trajectory_capacity=4,
                 trajectory_full=True) as ex:
    ex.set_state(pos, vel, ine); ex.step(16, sync=True); ex.trajectory()
print("done")
