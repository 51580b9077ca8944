"""compute-sanitizer target for the round-2 kernels: signature-specialised body kernels (one body and body pairs per
thread, odd tails, odd world-range offsets through invoke_batch), the persistent pair kernel with source-row ranges,
the gravity-only small-world signature, the new effector kinds, NULL columns, and the world-major trajectory layout
kernel.  Small sizes: the tool slows every kernel by 10-50x (the pair path needs >= 113 664 bodies: one short run)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import elodin_b200 as el
from elodin_b200.executor import FORCE, INERTIA, WORLD_ACCEL, WORLD_POS, WORLD_VEL
from tests.util import random_world

rng = np.random.default_rng(3)
for M in (1, 257, 113665):  # the last one takes the body-pair kernel with an odd tail
    pos, vel, ine = random_world(M, M, 1)
    cols = {"thrust": rng.uniform(0, 5, (M, 1, 1)), "wind": rng.normal(0, 1, (M, 1, 5)) ** 2, "aero_force": rng.normal(0, 1, (M, 1, 6))}
    for effs, names in (([], ()), ([el.GravityConst(), el.ThrustBody(), el.DragQuadratic(column="wind", per_body_params=True)], ("thrust", "wind")),
                        ([el.GravityFrame(), el.WrenchBody("aero_force", "linear_first")], ("aero_force",)),
                        ([el.GravityJ2(), el.WrenchWorld("aero_force")], ("aero_force",))):
        for integ in ("rk4", "semi_implicit"):
            with el.B200Exec(1, M, 0.01, None, effs, integ, "fast", max_fused_ticks=2, trajectory_every=2, trajectory_capacity=2) as ex:
                ex.set_state(pos + np.array([0, 0, 0, 0, 6.9e6, 0, 0]), vel, ine, **{k: cols[k] for k in names})
                ex.step(4, sync=True)
                assert ex.trajectory().shape[0] == 2
# invoke_batch over odd world ranges (plane bases not 16-byte aligned -> one body per thread) with NULL columns
M, N = 301, 3
pos, vel, ine = random_world(9, M, N)
with el.B200Exec(N, M, 0.01, None, [el.GravityConst()], "rk4", "fast", invoke_chunk_bodies=7 * N) as ex:
    tick, dt = el.component_id("tick"), el.component_id("simulation_time_step")
    table = {tick: np.array([0], dtype=np.uint64), FORCE: np.zeros((M, N, 6)), INERTIA: ine, WORLD_POS: pos, WORLD_ACCEL: np.zeros((M, N, 6)),
             dt: np.array([0.01]), WORLD_VEL: vel}
    ex.invoke_batch([table[c] for c in ex.input_ids], 2, out_cols=[None] * len(ex.output_ids))
    outs = [None if c == WORLD_ACCEL else np.empty(ex.column_shape(c), dtype=np.uint64 if c == tick else np.float64) for c in ex.output_ids]
    ex.invoke_batch([None] * len(ex.input_ids), 2, out_cols=outs)
# persistent pair kernel (64 <= N <= 1024), ragged N, few and many worlds; wheel fold + J2 through the interpreter
for N, M in ((64, 3), (100, 2), (257, 1), (96, 300)):
    p, v, I = random_world(N, M, N)
    p[..., 4:] *= 1e-2
    g = el.GravityEdges("softened", k_squared=0.3, softening=1e-4, edges=el.all_pairs_edges(N))
    for integ in ("rk4", "semi_implicit"):
        with el.B200Exec(N, M, 0.01, None, [g], integ, "fast") as ex:
            ex.set_state(p, v, I)
            ex.step(3, sync=True)
for N in (3, 7):
    M = 41
    p, v, I = random_world(N, M, N)
    with el.B200Exec(N, M, 0.01, None, [el.GravityEdges("newton", G=1e-3, edges=el.all_pairs_edges(N))], "rk4", "fast", max_fused_ticks=3) as ex:
        ex.set_state(p, v, I)
        ex.step(7, sync=True)
with el.B200Exec(2, 5, 0.01, None, [el.TorqueBodyFold("wheel_torques", 3), el.GravityJ2()], "semi_implicit", "exact") as ex:
    p, v, I = random_world(1, 5, 2)
    ex.set_state(p + np.array([0, 0, 0, 0, 6.9e6, 0, 0]), v, I, wheel_torques=rng.normal(0, 1e-3, (5, 2, 9)))
    ex.step(3, sync=True)
print("done")
# EGM08 stage-force kernel (degree 12, masked, RK4 and semi-implicit)
from tests.test_oracle_golden import _egm08_random_tables
c, s = _egm08_random_tables(12, rng)
for integ in ("rk4", "semi_implicit"):
    with el.B200Exec(3, 37, 0.05, None, [el.GravityEGM08(c, s, 12).with_mask(np.array([1, 0, 1], dtype=np.uint8))], integ, "exact") as ex:
        p, v, I = random_world(2, 37, 3)
        p[..., 4:] += 6.9e6
        ex.set_state(p, v, I)
        ex.step(2, sync=True)
print("egm done")
# late round 2: EGM08 term stream at the degrees that exercise every column shape, the EXACT tick's shared-divisor
# divisions (zero torque, zero force, ordinary operands), the division self-test kernel, worlds spread over CTA slices
import ctypes as C
from elodin_b200 import _lib
for Ld in (0, 1, 2, 5):
    cd, sd = np.tril(rng.normal(0, 1e-5, (Ld + 1, Ld + 1))), np.tril(rng.normal(0, 1e-5, (Ld + 1, Ld + 1)), -1)
    cd[0, 0] = 1.0
    with el.B200Exec(1, 33, 0.05, None, [el.GravityEGM08(cd, sd, Ld)], "rk4", "fast") as ex:
        p, v, I = random_world(4, 33, 1)
        p[..., 4:] += 6.9e6
        ex.set_state(p, v, I)
        ex.step(2, sync=True)
for effs, names in (([], ()), ([el.GravityConst(), el.ThrustBody(), el.DragQuadratic(column="wind", per_body_params=True)], ("thrust", "wind"))):
    M = 1000
    pos, vel, ine = random_world(5, M, 1)
    cols = {"thrust": rng.uniform(0, 5, (M, 1, 1)), "wind": rng.normal(0, 1, (M, 1, 5)) ** 2}
    for integ in ("rk4", "semi_implicit"):
        with el.B200Exec(1, M, 0.01, None, effs, integ, "exact") as ex:
            ex.set_state(pos, vel, ine, **{k: cols[k] for k in names})
            ex.step(3, sync=True)
out = (C.c_uint64 * 2)()
_lib.check(_lib.lib().b200_selftest_shared_divisor(0, 7, 1 << 16, out))
assert out[0] == 0
for N, M in ((1024, 1), (333, 2), (64, 1)):
    p, v, I = random_world(N, M, N)
    p[..., 4:] *= 1e-2
    with el.B200Exec(N, M, 0.01, None, [el.GravityEdges("softened", k_squared=0.3, softening=1e-4, edges=el.all_pairs_edges(N))], "rk4", "fast") as ex:
        ex.set_state(p, v, I)
        ex.step(3, sync=True)
print("late round 2 done")
