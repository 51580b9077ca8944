"""Pin the CPU oracle against the reference's own golden vectors (CPU only).

Golden telemetry: scripts/ci/baseline/{three-body,rocket,ball}-csv of the
reference, repacked by tests/golden/make_golden.py.  Known-answer vectors:
libs/nox/src/spatial.rs:630-676, libs/nox/src/quaternion.rs:352-388,
libs/nox-py/python/tests/test_all.py:67-83,228-291,342-366.
"""

import numpy as np
import pytest

THREE_BODY_EDGES = np.array([[0, 1], [1, 0], [0, 2], [1, 2], [2, 0], [2, 1]])  # main.py:82-89
G_NEWTON = 6.6743e-11


def _three_body_cols(golden):
    def col(c):
        return np.stack([golden[f"three_body.{e}.{c}"] for e in "abc"], 1)

    return [col(c) for c in ("world_pos", "world_vel", "world_accel", "force", "inertia")]


def test_three_body_100_ticks_bit_exact(golden, oracle):
    """RK4 (incl. the v0-stage behaviour) + edge_fold gravity: every one of the
    100 recorded ticks is reproduced bit for bit (pos, vel, force, accel)."""
    O = oracle
    pos, vel, acc, frc, ine = _three_body_cols(golden)
    dt = float(golden["three_body.simulation_time_step"][0, 0])
    assert dt == 0.008333333  # Duration-quantised 1/120 s (world_builder.rs:221)
    w = O.World(pos[0], vel[0], ine[0], acc[0], frc[0])
    eff = [O.Effector(O.EFF_GRAVITY_EDGES_NEWTON, p=(G_NEWTON,), edges=THREE_BODY_EDGES)]
    for t in range(1, 101):
        w.rk4(dt, 1, eff)
        assert np.array_equal(w.pos[0], pos[t]), t
        assert np.array_equal(w.vel[0], vel[t]), t
        assert np.array_equal(w.force[0], frc[t]), t
        assert np.array_equal(w.accel[0], acc[t]), t
    # entity ids of the recorded edges: a=1, b=2, c=3 (entity 0 = Globals, world.rs:174-183)
    assert np.array_equal(golden["three_body.edges_entity_ids"], THREE_BODY_EDGES + 1)
    assert np.array_equal(golden["three_body.tick"][:, 0], np.arange(101))


def test_textbook_rk4_is_not_the_reference(golden, oracle):
    """Guard against 'fixing' the scheme: advancing stage positions with the
    stage velocity (textbook RK4) misses the golden by ~5e-8 after one tick."""
    O = oracle
    pos, vel, acc, frc, ine = _three_body_cols(golden)
    dt = float(golden["three_body.simulation_time_step"][0, 0])
    eff = [O.Effector(O.EFF_GRAVITY_EDGES_NEWTON, p=(G_NEWTON,), edges=THREE_BODY_EDGES)]

    def A(x, v):
        w = O.World(x, v, ine[0])
        return w.eval_stage(0, eff)[1]

    x0, v0 = pos[0], vel[0]
    k1v, k1a = v0, A(x0, v0)
    x2 = x0.copy(); x2[:, 4:] += 0.5 * dt * k1v[:, 3:]
    k2v, k2a = v0 + 0.5 * dt * k1a, A(x2, v0 + 0.5 * dt * k1a)
    x3 = x0.copy(); x3[:, 4:] += 0.5 * dt * k2v[:, 3:]
    k3v, k3a = v0 + 0.5 * dt * k2a, A(x3, v0 + 0.5 * dt * k2a)
    x4 = x0.copy(); x4[:, 4:] += dt * k3v[:, 3:]
    k4v = v0 + dt * k3a
    x1 = x0[:, 4:] + dt / 6 * (k1v + 2 * k2v + 2 * k3v + k4v)[:, 3:]
    err = np.max(np.abs(x1 - pos[1][:, 4:]))
    assert 1e-11 < err < 1e-6


def _rocket_step(O, golden, t):
    g = golden
    w = O.World(g["rocket.world_pos"][t][None], g["rocket.world_vel"][t][None],
                g["rocket.inertia"][t][None], g["rocket.world_accel"][t][None],
                g["rocket.force"][t][None])
    eff = [  # examples/rocket/main.py:575  effectors = gravity | apply_thrust | apply_aero_forces
        O.Effector(O.EFF_GRAVITY_CONST, p=(0.0, 0.0, -9.81)),
        O.Effector(O.EFF_THRUST_BODY, p=(-1.0, 0.0, 0.0), column=g["rocket.thrust"][t + 1].reshape(1, 1, 1)),
        O.Effector(O.EFF_WRENCH_BODY, column=g["rocket.aero_force"][t + 1].reshape(1, 1, 6)),
    ]
    w.rk4(float(g["rocket.simulation_time_step"][0, 0]), 1, eff)
    return w


def test_rocket_one_step_predictions(golden, oracle):
    """Attitude kinematics, torque / diagonal inertia, body->world rotation with
    the recorded thrust / aero_force as stage-constant inputs.

    Canonical (plain IEEE) oracle: within 1e-15 x max|value| of every recorded vector.
    Golden-host mode (FMA-contracted `dot` where the reference JIT's pointer-ABI
    runtime contracts it, tensor_rt.rs:1141-1163): 100/100 steps bit-exact."""
    O = oracle
    g = golden
    names = ("world_pos", "world_vel", "force", "world_accel")
    try:
        O.set_dot_mode(0)
        exact = 0
        for t in range(100):
            w = _rocket_step(O, g, t)
            got = (w.pos[0, 0], w.vel[0, 0], w.force[0, 0], w.accel[0, 0])
            ok = True
            for name, a in zip(names, got):
                ref = g[f"rocket.{name}"][t + 1]
                scale = np.max(np.abs(ref))
                assert np.max(np.abs(a - ref)) <= 1e-15 * scale, (t, name)
                ok &= np.array_equal(a, ref)
            exact += ok
        assert exact >= 60
        O.set_dot_mode(1)
        for t in range(100):
            w = _rocket_step(O, g, t)
            got = (w.pos[0, 0], w.vel[0, 0], w.force[0, 0], w.accel[0, 0])
            for name, a in zip(names, got):
                assert np.array_equal(a, g[f"rocket.{name}"][t + 1]), (t, name)
    finally:
        O.set_dot_mode(0)


def test_ball_one_step_predictions_bit_exact(golden, oracle):
    """const-g + quadratic drag evaluated on the *stage* velocity
    (examples/ball/sim.py:56-58,99-116) with the recorded wind; `bounce`
    (sim.py:64-72) edits vel before six_dof and is replayed on the host."""
    O = oracle
    g = golden
    dt = float(g["ball.simulation_time_step"][0, 0])
    for t in range(100):
        v = g["ball.world_vel"][t].copy()
        if max(g["ball.world_pos"][t][6], v[5]) < 0.0:
            v = np.concatenate([np.zeros(3), v[3:] * np.array([1.0, 1.0, -1.0]) * 0.85])
        w = O.World(g["ball.world_pos"][t][None], v[None], g["ball.inertia"][t][None],
                    g["ball.world_accel"][t][None], g["ball.force"][t][None])
        eff = [O.Effector(O.EFF_GRAVITY_CONST, p=(0.0, 0.0, -9.81)),
               O.Effector(O.EFF_DRAG_QUADRATIC, p=(0.5 * 1.225, 2 * 3.1415 * 0.2 ** 2),
                          column=g["ball.wind"][t + 1].reshape(1, 1, 3))]
        w.rk4(dt, 1, eff)
        assert np.array_equal(w.pos[0, 0], g["ball.world_pos"][t + 1]), t
        assert np.array_equal(w.vel[0, 0], g["ball.world_vel"][t + 1]), t
        assert np.array_equal(w.force[0, 0], g["ball.force"][t + 1]), t
        assert np.array_equal(w.accel[0, 0], g["ball.world_accel"][t + 1]), t


def test_cube_sat_earth_semi_implicit_100_ticks_bit_exact(golden, oracle):
    """Integrator.SemiImplicit (semi_implicit.rs:42-62) pinned by a golden: the `earth` entity of
    scripts/ci/baseline/cube-sat-csv is a free body spinning at the sidereal rate; all 100
    recorded rows (pos, vel, accel) are reproduced bit for bit from row 0."""
    O = oracle
    g = golden
    dt = float(g["cube_sat.simulation_time_step"][0, 0])
    assert not np.any(g["cube_sat.earth.force"])
    w = O.World(g["cube_sat.earth.world_pos"][0][None, None], g["cube_sat.earth.world_vel"][0][None, None],
                g["cube_sat.earth.inertia"][0][None, None])
    for t in range(1, 101):
        w.semi_implicit(dt, 1)
        assert np.array_equal(w.pos[0, 0], g["cube_sat.earth.world_pos"][t]), t
        assert np.array_equal(w.vel[0, 0], g["cube_sat.earth.world_vel"][t]), t
        assert np.array_equal(w.accel[0, 0], g["cube_sat.earth.world_accel"][t]), t
    assert w.pos[0, 0, 2] > 3e-5  # the attitude really moved


# ---- known-answer tests of the primitives -------------------------------------


def _axis_angle(axis, angle):
    axis = np.asarray(axis, float)
    axis = axis / np.sqrt(axis @ axis)
    return np.concatenate([axis * np.sin(angle / 2), [np.cos(angle / 2)]])


def test_quat_known_answers(oracle):
    O = oracle
    # quaternion.rs:352-361 test_quat_mult
    out = O.qmul(_axis_angle([1, 0, 0], 3.0), _axis_angle([1, 0, 0], 1.0))
    assert np.array_equal(out, [0.9092974268256817, 0.0, 0.0, -0.4161468365471424])
    # quaternion.rs:363-370 test_quat_inverse
    out = O.qinv(_axis_angle([1, 0, 0], 3.0))
    assert np.array_equal(out, [-0.9974949866040544, -0.0, -0.0, 0.0707372016677029])
    # quaternion.rs:372-381 test_quat_vec_mult
    out = O.qrot(_axis_angle([1, 0, 0], 3.0), [1.0, 2.0, 3.0])
    assert np.allclose(out, [1.0, -2.4033450173804924, -2.6877374736816018], atol=1e-6)
    # quaternion.rs:383-388 convention i*j = k
    assert np.array_equal(O.qmul([1.0, 0, 0, 0], [0, 1.0, 0, 0]), [0, 0, 1.0, 0])


def test_spatial_transform_add_known_answers(oracle):
    O = oracle
    # spatial.rs:630-650 test_spatial_transform_add (assert_eq! => exact)
    out = O.transform_add_motion([0, 0, 0, 1.0, 0, 0, 0], [0, 0, 1.0, 0, 0, 0])
    assert np.array_equal(out, [0.0, 0.0, 0.4472135954999579, 0.8944271909999159, 0.0, 0.0, 0.0])
    # spatial.rs:652-676 test_spatial_transform_integrate
    p = np.array([0, 0, 0, 1.0, 0, 0, 0])
    for _ in range(20):
        p = O.transform_add_motion(p, [0, 0, 0.25 / 20.0, 0, 0, 0])
    assert np.allclose(p, [0, 0, 0.12467473338522769, 0.992197667229329, 0, 0, 0], atol=1e-5)


def _body(O, vel, inertia=None):
    pos = np.array([[0, 0, 0, 1.0, 0, 0, 0]])
    ine = np.array([[1.0, 1.0, 1.0, 0, 0, 0, 1.0]]) if inertia is None else inertia
    return O.World(pos, np.array([vel], float), ine)


def test_six_dof_kats_from_test_all(oracle):
    O = oracle
    # test_all.py:67-83 test_six_dof: x = dt after one tick of six_dof(1/60)
    w = _body(O, [0, 0, 0, 1.0, 0, 0]).rk4(0.01, 1, dt_final=1.0 / 60.0)
    assert np.allclose(w.pos[0, 0, :4], [0, 0, 0, 1.0])
    assert np.allclose(w.pos[0, 0, 4:], [0.01666667, 0, 0])
    # test_all.py:228-291 test_six_dof_ang_vel_int (Julia/Simulink values, rtol 1e-5)
    dt = 0.008333333
    for omega, want in [
        ([0, 0, 1.0], [0.0, 0.0, 0.479425538604203, 0.8775825618903728]),
        ([0, 1.0, 0], [0.0, 0.479425538604203, 0.0, 0.8775825618903728]),
        ([1.0, 1.0, 0], [0.45936268493243, 0.45936268493243, 0.0, 0.76024459707606]),
    ]:
        w = _body(O, omega + [0, 0, 0]).rk4(dt, 120, dt_final=1.0 / 120.0)
        assert np.isclose(w.pos[0, 0, :4], want, rtol=1e-5).all()
    # test_all.py:342-366 test_six_dof_force: x = 0.5 after 1 s of unit force
    class ConstForce:  # constant world-frame force == GRAVITY_CONST with g = F/m, m = 1
        pass

    w = _body(O, [0, 0, 0, 0, 0, 0]).rk4(dt, 120, [O.Effector(O.EFF_GRAVITY_CONST, p=(1.0, 0, 0))],
                                         dt_final=1.0 / 120.0)
    assert np.isclose(w.pos[0, 0], [0, 0, 0, 1.0, 0.5, 0, 0], rtol=1e-5).all()


def test_semi_implicit_matches_definition(oracle):
    """semi_implicit.rs:42-62: v' = v + dt*a ; x' = x (+) dt*v'."""
    O = oracle
    rng = np.random.default_rng(3)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    pos = np.concatenate([q, rng.normal(size=3)])[None]
    vel = rng.normal(size=(1, 6))
    ine = np.array([[0.5, 2.0, 3.0, 0, 0, 0, 4.0]])
    eff = [O.Effector(O.EFF_GRAVITY_CONST, p=(0.1, -0.2, -9.81))]
    w = O.World(pos, vel, ine)
    F, A = w.eval_stage(0, eff)
    v1 = vel[0] + 0.01 * A[0]
    x1 = O.transform_add_motion(pos[0], 0.01 * v1)
    w.semi_implicit(0.01, 1, eff)
    assert np.array_equal(w.vel[0, 0], v1)
    assert np.array_equal(w.pos[0, 0], x1)
    assert np.array_equal(w.force[0, 0], F[0]) and np.array_equal(w.accel[0, 0], A[0])


def test_oracle_world_axis_and_threads(oracle):
    """M stacked worlds == M independent runs; thread count does not change bits."""
    O = oracle
    rng = np.random.default_rng(11)
    M, N = 6, 5
    q = rng.normal(size=(M, N, 4)); q /= np.linalg.norm(q, axis=-1, keepdims=True)
    pos = np.concatenate([q, rng.normal(size=(M, N, 3))], -1)
    vel = rng.normal(size=(M, N, 6))
    ine = np.concatenate([rng.uniform(0.1, 10, (M, N, 3)), np.zeros((M, N, 3)), rng.uniform(0.5, 50, (M, N, 1))], -1)
    edges = np.array([[i, j] for i in range(N) for j in range(N) if i != j])
    thrust = rng.uniform(0, 5, (M, N, 1))
    eff = [O.Effector(O.EFF_GRAVITY_EDGES_SOFTENED, p=(1e-3, 1e-10), edges=edges),
           O.Effector(O.EFF_THRUST_BODY, p=(-1.0, 0, 0), column=thrust)]
    a = O.World(pos, vel, ine).rk4(0.01, 7, eff, threads=1)
    b = O.World(pos, vel, ine).rk4(0.01, 7, eff, threads=4)
    assert np.array_equal(a.pos, b.pos) and np.array_equal(a.vel, b.vel)
    for m in range(M):
        e1 = [O.Effector(O.EFF_GRAVITY_EDGES_SOFTENED, p=(1e-3, 1e-10), edges=edges),
              O.Effector(O.EFF_THRUST_BODY, p=(-1.0, 0, 0), column=thrust[m:m + 1])]
        c = O.World(pos[m], vel[m], ine[m]).rk4(0.01, 7, e1)
        assert np.array_equal(c.pos[0], a.pos[m]) and np.array_equal(c.vel[0], a.vel[m])


def test_oracle_entity_masks(oracle):
    """An effector with an entity mask equals the unmasked effector on members and no effector on the rest."""
    O = oracle
    rng = np.random.default_rng(4)
    q = rng.normal(size=(1, 3, 4)); q /= np.linalg.norm(q, axis=-1, keepdims=True)
    pos = np.concatenate([q, rng.normal(size=(1, 3, 3))], -1)
    vel = rng.normal(size=(1, 3, 6))
    ine = np.concatenate([rng.uniform(1, 2, (1, 3, 3)), np.zeros((1, 3, 3)), rng.uniform(1, 2, (1, 3, 1))], -1)
    thrust = rng.uniform(1, 5, (1, 3, 1))
    masked = O.World(pos, vel, ine).rk4(0.01, 3, [O.Effector(O.EFF_THRUST_BODY, p=(1.0, 0, 0), column=thrust, mask=[1, 0, 1])])
    full = O.World(pos, vel, ine).rk4(0.01, 3, [O.Effector(O.EFF_THRUST_BODY, p=(1.0, 0, 0), column=thrust)])
    free = O.World(pos, vel, ine).rk4(0.01, 3)
    assert np.array_equal(masked.pos[0, [0, 2]], full.pos[0, [0, 2]]) and np.array_equal(masked.vel[0, 1], free.vel[0, 1])
    assert not np.array_equal(masked.vel[0, 1], full.vel[0, 1])


# --------------------------------------------------------------------------- cube-sat `ore_sat` (round 2)
def _max_ulp_rel(a, b):
    scale = np.maximum(np.max(np.abs(b), axis=-1, keepdims=True), 1e-300)
    return float(np.max(np.abs(a - b) / scale))


def test_cube_sat_reaction_wheel_fold_golden(golden, oracle):
    """rw_effector (examples/cube-sat/main.py:492-505): Force.torque of tick t = fold over the three wheels, in edge
    spawn order, of q_{t-1} @ rw_force_t[k].torque.  The golden was produced by backend="jax-cpu" (main.py:714), whose
    LLVM code generator contracts multiply-adds the IEEE-plain oracle does not: 54/100 rows are bit-identical, the
    rest differ in the last bit.  Stated bar: 5e-16 vector-relative (2 ulp); the reference's own gate is 1e-4."""
    O = oracle
    pos, frc = golden["cube_sat.ore_sat.world_pos"], golden["cube_sat.ore_sat.force"]
    rw = np.concatenate([golden[f"cube_sat.rw_{k}.rw_force"][:, :3] for k in (1, 2, 3)], -1)  # [T, 9]
    T = len(pos)
    eff = O.Effector(O.EFF_TORQUE_BODY_FOLD, column=rw[1:].reshape(T - 1, 1, 9))
    w = O.World(pos[:-1].reshape(T - 1, 1, 7), np.zeros((T - 1, 1, 6)), np.tile(golden["cube_sat.ore_sat.inertia"][0], (T - 1, 1, 1)))
    got = np.stack([w.eval_stage(m, [eff])[0][0] for m in range(T - 1)])
    assert np.all(got[:, 3:] == 0.0)
    exact = int(np.sum(np.all(got[:, :3] == frc[1:, :3], axis=-1)))
    assert exact >= 50, exact
    assert _max_ulp_rel(got[:, :3], frc[1:, :3]) <= 5e-16


def test_cube_sat_semi_implicit_with_recorded_wrench_golden(golden, oracle):
    """semi_implicit.rs:42-62 with a full wrench: from row t-1's state and row t's recorded Force (wheel torques + EGM08
    gravity, fed through the WRENCH_WORLD effector) the oracle reproduces row t's WorldAccel / WorldVel / WorldPos to
    <= 2e-15 vector-relative (XLA-CPU's FMA contraction again; 92/100 velocities and 98/100 poses are bit-identical)."""
    O = oracle
    g = golden
    pos, vel, acc, frc = (g[f"cube_sat.ore_sat.{c}"] for c in ("world_pos", "world_vel", "world_accel", "force"))
    ine = g["cube_sat.ore_sat.inertia"]
    dt = float(g["cube_sat.simulation_time_step"][0, 0])
    T = len(pos)
    eff = O.Effector(O.EFF_WRENCH_WORLD, column=frc[1:].reshape(T - 1, 1, 6))
    w = O.World(pos[:-1].reshape(T - 1, 1, 7).copy(), vel[:-1].reshape(T - 1, 1, 6).copy(), np.tile(ine[0], (T - 1, 1, 1)))
    w.semi_implicit(dt, 1, [eff])
    assert np.array_equal(w.force[:, 0], frc[1:])
    assert _max_ulp_rel(w.accel[:, 0], acc[1:]) <= 2e-15
    assert _max_ulp_rel(w.vel[:, 0], vel[1:]) <= 2e-15
    assert _max_ulp_rel(w.pos[:, 0, :4], pos[1:, :4]) <= 2e-15 and _max_ulp_rel(w.pos[:, 0, 4:], pos[1:, 4:]) <= 2e-15
    assert int(np.sum(np.all(w.pos[:, 0] == pos[1:], axis=-1))) >= 90


def test_j2_field_matches_an_independent_closed_form(oracle):
    """GRAVITY_J2 (python/elodin/j2.py:5-29) against the textbook J2 acceleration written a different way
    (a = -mu r/n^3 - 1.5 J2 mu R^2/n^5 [(1 - 5 z^2/n^2) r + 2 z e_z]); parity unpinned: no reference golden uses J2."""
    O = oracle
    rng = np.random.default_rng(2)
    M = 64
    r = rng.normal(size=(M, 3)); r *= (6.8e6 + rng.uniform(0, 4e5, (M, 1))) / np.linalg.norm(r, axis=-1, keepdims=True)
    pos = np.zeros((M, 1, 7)); pos[..., 3] = 1.0; pos[:, 0, 4:] = r
    m = rng.uniform(1, 500, M)
    ine = np.zeros((M, 1, 7)); ine[:, 0, :3] = 1.0; ine[:, 0, 6] = m
    mu, J2, R = 3.986004418e14, 1.08262668e-3, 6.378e6
    w = O.World(pos, np.zeros((M, 1, 6)), ine)
    got = np.stack([w.eval_stage(i, [O.Effector(O.EFF_GRAVITY_J2, p=(mu, J2, R))])[0][0] for i in range(M)])
    n = np.linalg.norm(r, axis=-1, keepdims=True)
    ez = np.array([0.0, 0.0, 1.0])
    a = -mu * r / n**3 - 1.5 * J2 * mu * R**2 / n**5 * ((1 - 5 * r[:, 2:3] ** 2 / n**2) * r + 2 * r[:, 2:3] * ez)
    assert np.all(got[:, :3] == 0.0)
    assert _max_ulp_rel(got[:, 3:], a * m[:, None]) <= 5e-15
    # the J2 part is ~1e-3 of the field: make sure it is there with the right sign (pulls toward the equator plane)
    pm = -mu * r / n**3 * m[:, None]
    assert 2e-4 < np.max(np.linalg.norm(got[:, 3:] - pm, axis=-1) / np.linalg.norm(pm, axis=-1)) < 3e-3


# --------------------------------------------------------------------------- EGM08 (round 2)
def _egm08_array_form(x, y, z, mass, c_bar, s_bar, L, mu=3.986004418e14, r_ref=6.378e6):
    """python/elodin/egm08.py:84-216 restated in its own array formulation (scans -> loops, rolls and `.at[].set()`
    kept as such, `jnp.sum(jnp.sum(.., axis=1), axis=0)` as numpy sums) — an independent second reading of the source
    that the column-wise oracle (oracle/sixdof_oracle.c:eff_gravity_egm08) is checked against."""
    kd = lambda d: 1.0 if d == 0 else 2.0
    l_arr, m_arr = np.arange(L + 1), np.arange(L + 1)
    diag, cur = np.zeros(L + 1), 1.0
    for l in range(L + 1):                                         # compute_a_bar_diagonal
        cur = cur if l == 0 else cur * np.sqrt(((2 * l + 1) * kd(l)) / ((2 * l) * kd(l - 1)))
        diag[l] = cur
    a = np.diag(diag)
    r = np.sqrt(x * x + y * y + z * z)
    s, t, u = x / r, y / r, z / r
    off = np.array([0.0 if l == 0 else a[l, l] * np.sqrt(((2 * l) * kd(l - 1)) / kd(l)) * u for l in range(L + 1)])
    a = np.roll(np.diag(off), -1, axis=1) + a                      # pre_compute_parameters
    n1 = lambda l, m: np.sqrt(((2 * l + 1) * (2 * l - 1)) / ((l + m) * (l - m))) if l >= m + 2 else 0.0
    n2 = lambda l, m: np.sqrt(((l + m - 1) * (l - m - 1) * (2 * l + 1)) / ((2 * l - 3) * (l + m) * (l - m))) if l >= m + 2 else 0.0
    full = np.zeros((L + 1, L + 1))
    for m in range(L + 1):                                         # compute_a_bar_full_m, vmapped over m
        c0, c1 = a[0, 0], (a[1, 0] if L >= 1 else 0.0)
        for l in range(L + 1):
            alm = u * n1(l, m) * c0 - n2(l, m) * c1 if l >= m + 2 else a[l, m]
            c0, c1 = alm, c0
            full[m, l] = alm
    a = full.T
    im, rm, ci, cr = np.zeros(L + 1), np.zeros(L + 1), 0.0, 1.0
    for m in range(L + 1):                                         # compute_i_r_m
        if m > 0:
            ci, cr = s * ci + t * cr, s * cr - t * ci
        im[m], rm[m] = ci, cr
    rho = (mu / r) * ((r_ref / r) ** l_arr)
    nq1, nq2 = np.zeros((L + 1, L + 1)), np.zeros((L + 1, L + 1))
    for m in range(L + 1):
        for l in range(L + 1):
            num = (l - m) * kd(m) * (l + m + 1)
            nq1[l, m] = 0.0 if num < 0 else np.sqrt(num / kd(m + 1))
            nq2[l, m] = np.sqrt((l + m + 2) * (l + m + 1) * (2 * l + 1) * kd(m) / ((2 * l + 3) * kd(m + 1)))
    rho1 = np.roll(rho, -1); rho1[-1] = 0.0                        # compute_components
    rm1 = np.roll(rm, 1); rm1[0] = 0.0
    im1 = np.roll(im, 1); im1[0] = 0.0
    e = c_bar * rm1 + s_bar * im1
    mp = np.roll(m_arr, -1).astype(float); mp[-1] = 0.0
    a1 = np.sum(np.sum(((rho1 / r_ref) * a.T).T * mp * e, axis=1), axis=0)
    f = s_bar * rm1 - c_bar * im1
    a2 = np.sum(np.sum(((rho1 / r_ref) * a.T).T * mp * f, axis=1), axis=0)
    d = c_bar * rm + s_bar * im
    ab1 = np.roll(a, -1, axis=1); ab1[:, -1] = 0.0
    a3 = np.sum(np.sum(((rho1 / r_ref) * ab1.T).T * mp * nq1 * d, axis=1), axis=0)
    ab2 = np.roll(ab1, -1, axis=0); ab2[-1, :] = 0.0
    a4 = np.sum(np.sum(((rho1 / r_ref) * ab2.T).T * mp * nq2 * d * (-1), axis=1), axis=0)
    return mass * np.array([a1 + s * a4, a2 + t * a4, a3 + u * a4])


def _egm08_random_tables(L, rng):
    """EGM-like normalised coefficients: C00 = 1, degree-1 terms 0, C20 = -J2/sqrt(5), the rest ~ 1e-5 / l^2 (Kaula)."""
    c, s = np.zeros((L + 1, L + 1)), np.zeros((L + 1, L + 1))
    for l in range(2, L + 1):
        for m in range(l + 1):
            c[l, m] = rng.normal(0, 1e-5 / l**2)
            s[l, m] = 0.0 if m == 0 else rng.normal(0, 1e-5 / l**2)
    c[0, 0], c[2, 0] = 1.0, -1.08262668e-3 / np.sqrt(5.0)
    return c, s


def test_egm08_oracle_against_the_array_form_and_j2(oracle):
    """GRAVITY_EGM08 (egm08.py): (1) with C00 and C20 alone the field IS j2.py's (2.3e-16) — the reference's own closed
    form pins the zonal path; (2) with full random tables of degree 8 and 64 the column-wise oracle equals the array
    formulation of the source (different summation order: <= 1e-13).  Parity unpinned against a golden: the coefficient
    tables are a run-time download of the reference (egm08.py:27-40), and cube-sat's recorded field cannot be recomputed."""
    O = oracle
    rng = np.random.default_rng(12)
    r = np.array([-4302097.779462299, -3609888.660035412, 3795167.739124752])  # the cube-sat golden's first position
    mass = 2.8252
    pos = np.zeros((1, 1, 7)); pos[0, 0, 3] = 1.0; pos[0, 0, 4:] = r
    ine = np.zeros((1, 1, 7)); ine[0, 0, :3] = 1.0; ine[0, 0, 6] = mass
    w = O.World(pos, np.zeros((1, 1, 6)), ine)
    c, s = np.zeros((5, 5)), np.zeros((5, 5))
    c[0, 0], c[2, 0] = 1.0, -1.08262668e-3 / np.sqrt(5.0)
    f_egm = w.eval_stage(0, [O.Effector(O.EFF_GRAVITY_EGM08, p=(3.986004418e14, 6.378e6, 4), tables=(c, s))])[0][0]
    f_j2 = w.eval_stage(0, [O.Effector(O.EFF_GRAVITY_J2, p=(3.986004418e14, 1.08262668e-3, 6.378e6))])[0][0]
    assert np.all(f_egm[:3] == 0.0) and np.max(np.abs(f_egm - f_j2)) <= 1e-15 * np.max(np.abs(f_j2))
    # the source zeroes rho_{L+1} (`rho_l_1 = roll(rho_l, -1).at[-1].set(0)`, egm08.py:150): the terms of the top degree L
    # drop out, so max_degree = 2 leaves the point mass alone
    f_l2 = w.eval_stage(0, [O.Effector(O.EFF_GRAVITY_EGM08, p=(3.986004418e14, 6.378e6, 2), tables=(c[:3, :3], s[:3, :3]))])[0][0]
    pm0 = -3.986004418e14 * mass * r / np.linalg.norm(r) ** 3
    assert np.max(np.abs(f_l2[3:] - pm0)) <= 1e-15 * np.max(np.abs(pm0))
    for L in (8, 64):
        c, s = _egm08_random_tables(L, rng)
        got = w.eval_stage(0, [O.Effector(O.EFF_GRAVITY_EGM08, p=(3.986004418e14, 6.378e6, L), tables=(c, s))])[0][0][3:]
        want = _egm08_array_form(*r, mass, c, s, L)
        assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want)), L
        pm = -3.986004418e14 * mass * r / np.linalg.norm(r) ** 3
        assert 1e-4 < np.linalg.norm(got - pm) / np.linalg.norm(pm) < 5e-3  # the harmonics are really in there
