"""GPU parity tests: the sm_100a kernels, called through the C ABI
(libb200_sixdof.so), against the CPU oracle and the reference's golden telemetry.

Bars:
  * B200_MATH_EXACT: bit-identical to the oracle (which reproduces the reference's
    three-body / ball golden CSVs bit for bit) — `np.array_equal`.
  * B200_MATH_FAST: <= FAST_TOL_TICK vector-relative per tick vs EXACT/oracle
    (stated tolerance 1e-12), and <= FAST_TOL_1000 after 1000 ticks.
"""

import numpy as np
import pytest

import elodin_b200 as el
from elodin_b200.executor import FORCE, INERTIA, WORLD_ACCEL, WORLD_POS, WORLD_VEL
from tests.util import effector_pair, max_rel, random_world

pytestmark = pytest.mark.gpu

FAST_TOL_TICK = 1e-12
FAST_TOL_1000 = 1e-9
THREE_BODY_EDGES = np.array([[0, 1], [1, 0], [0, 2], [1, 2], [2, 0], [2, 1]])


def _run_gpu(pos, vel, ine, effs, cols, dt, n_ticks, math="exact", integrator="rk4", time_step=None, fused=1,
             accel=None):
    M, N, _ = pos.shape
    with el.B200Exec(N, M, dt, time_step, effs, integrator, math, max_fused_ticks=fused) as ex:
        ex.set_state(pos, vel, ine, accel=accel, **cols)
        ex.step(n_ticks, sync=True)
        return (ex.download(WORLD_POS), ex.download(WORLD_VEL), ex.download(WORLD_ACCEL), ex.download(FORCE))


def _run_oracle(O, pos, vel, ine, effs, dt, n_ticks, integrator="rk4", time_step=None, accel=None):
    w = O.World(pos, vel, ine, accel)
    if integrator == "rk4":
        w.rk4(dt, n_ticks, effs, dt_final=time_step, threads=4)
    else:
        w.semi_implicit(dt if time_step is None else time_step, n_ticks, effs, threads=4)
    return w.pos, w.vel, w.accel, w.force


def _assert_exact(got, want, what=""):
    for name, a, b in zip(("pos", "vel", "accel", "force"), got, want):
        assert np.array_equal(a, b), f"{what} {name}: max abs diff {np.max(np.abs(a - b))}"


def _assert_close(got, want, tol, what="", check_force=True):
    names = ("pos_q", "pos_x", "vel", "accel", "force")
    pairs = [(got[0][..., :4], want[0][..., :4]), (got[0][..., 4:], want[0][..., 4:]), (got[1], want[1]),
             (got[2], want[2]), (got[3], want[3])]
    for name, (a, b) in zip(names, pairs):
        if name == "force" and not check_force:
            continue
        # accelerations / forces can cancel to ~0: scale by the batch-wide magnitude
        scale = max(np.max(np.abs(b)), 1e-300)
        err = float(np.max(np.abs(a - b)) / scale)
        assert err <= tol, f"{what} {name}: rel err {err:.3e} > {tol}"


# --------------------------------------------------------------------------- golden


def test_three_body_golden_bit_exact(golden, oracle):
    """The GPU EXACT path reproduces all 100 recorded ticks of
    scripts/ci/baseline/three-body-csv bit for bit, through invoke_batch."""
    def col(c):
        return np.stack([golden[f"three_body.{e}.{c}"] for e in "abc"], 1)

    pos, vel, acc, frc, ine = [col(c) for c in ("world_pos", "world_vel", "world_accel", "force", "inertia")]
    dt = float(golden["three_body.simulation_time_step"][0, 0])
    eff = el.GravityEdges("newton", G=6.6743e-11, edges=THREE_BODY_EDGES)
    with el.B200Exec(3, 1, dt, None, [eff], "rk4", "exact") as ex:
        state = {WORLD_POS: pos[0][None], WORLD_VEL: vel[0][None], WORLD_ACCEL: acc[0][None], FORCE: frc[0][None],
                 INERTIA: ine[0][None]}
        tick = 0
        for t in range(1, 101):
            ins = []
            for cid in ex.input_ids:
                if cid == el.component_id("tick"):
                    ins.append(np.array([tick], dtype=np.uint64))
                elif cid == el.component_id("simulation_time_step"):
                    ins.append(np.array([dt]))
                else:
                    ins.append(state[cid])
            outs = dict(zip(ex.output_ids, ex.invoke_batch(ins, 1)))
            tick = int(outs[el.component_id("tick")][0])
            assert tick == t
            for cid in state:
                state[cid] = outs[cid]
            assert np.array_equal(state[WORLD_POS][0], pos[t]), t
            assert np.array_equal(state[WORLD_VEL][0], vel[t]), t
            assert np.array_equal(state[FORCE][0], frc[t]), t
            assert np.array_equal(state[WORLD_ACCEL][0], acc[t]), t
            assert np.array_equal(state[INERTIA][0], ine[t]), t  # pass-through output
    # FAST: within tolerance of the golden after 100 ticks
    got = _run_gpu(pos[0][None], vel[0][None], ine[0][None], [eff], {}, dt, 100, "fast")
    assert max_rel(got[0][0][:, 4:], pos[100][:, 4:]) < 1e-11
    assert max_rel(got[1][0][:, 3:], vel[100][:, 3:]) < 1e-11


def test_ball_golden_one_step_bit_exact(golden):
    g = golden
    dt = float(g["ball.simulation_time_step"][0, 0])
    effs = [el.GravityConst((0.0, 0.0, -9.81)), el.DragQuadratic(0.5 * 1.225, 2 * 3.1415 * 0.2 ** 2, "wind")]
    with el.B200Exec(1, 1, dt, None, effs, "rk4", "exact") as ex:
        for t in range(100):
            v = g["ball.world_vel"][t].copy()
            if max(g["ball.world_pos"][t][6], v[5]) < 0.0:  # bounce, examples/ball/sim.py:64-72 (host side)
                v = np.concatenate([np.zeros(3), v[3:] * np.array([1.0, 1.0, -1.0]) * 0.85])
            ex.set_state(g["ball.world_pos"][t], v, g["ball.inertia"][t], accel=g["ball.world_accel"][t],
                         wind=g["ball.wind"][t + 1])
            ex.step(1, sync=True)
            assert np.array_equal(ex.download(WORLD_POS)[0, 0], g["ball.world_pos"][t + 1]), t
            assert np.array_equal(ex.download(WORLD_VEL)[0, 0], g["ball.world_vel"][t + 1]), t
            assert np.array_equal(ex.download(FORCE)[0, 0], g["ball.force"][t + 1]), t
            assert np.array_equal(ex.download(WORLD_ACCEL)[0, 0], g["ball.world_accel"][t + 1]), t


def test_rocket_golden_one_step(golden, oracle):
    """EXACT == canonical oracle bit for bit; both within 1e-15 of the recorded rocket
    telemetry (whose host JIT contracted `dot` with FMA, see oracle/sixdof_oracle.c)."""
    g = golden
    dt = float(g["rocket.simulation_time_step"][0, 0])
    effs = [el.GravityConst((0.0, 0.0, -9.81)), el.ThrustBody((-1.0, 0.0, 0.0), "thrust"), el.WrenchBody("aero_force")]
    O = oracle
    with el.B200Exec(1, 1, dt, None, effs, "rk4", "exact") as ex, el.B200Exec(1, 1, dt, None, effs, "rk4", "fast") as fx:
        for t in range(100):
            for e in (ex, fx):
                e.set_state(g["rocket.world_pos"][t], g["rocket.world_vel"][t], g["rocket.inertia"][t],
                            accel=g["rocket.world_accel"][t], thrust=g["rocket.thrust"][t + 1],
                            aero_force=g["rocket.aero_force"][t + 1])
                e.step(1, sync=True)
            w = O.World(g["rocket.world_pos"][t][None], g["rocket.world_vel"][t][None], g["rocket.inertia"][t][None],
                        g["rocket.world_accel"][t][None])
            w.rk4(dt, 1, [O.Effector(O.EFF_GRAVITY_CONST, p=(0, 0, -9.81)),
                          O.Effector(O.EFF_THRUST_BODY, p=(-1.0, 0, 0), column=g["rocket.thrust"][t + 1].reshape(1, 1, 1)),
                          O.Effector(O.EFF_WRENCH_BODY, column=g["rocket.aero_force"][t + 1].reshape(1, 1, 6))])
            got = (ex.download(WORLD_POS), ex.download(WORLD_VEL), ex.download(WORLD_ACCEL), ex.download(FORCE))
            _assert_exact(got, (w.pos, w.vel, w.accel, w.force), f"rocket t={t}")
            for name, a in (("world_pos", got[0]), ("world_vel", got[1]), ("world_accel", got[2]), ("force", got[3])):
                ref = g[f"rocket.{name}"][t + 1]
                assert np.max(np.abs(a[0, 0] - ref)) <= 1e-15 * np.max(np.abs(ref)), (t, name)
            fgot = (fx.download(WORLD_POS), fx.download(WORLD_VEL), fx.download(WORLD_ACCEL), fx.download(FORCE))
            _assert_close(fgot, got, FAST_TOL_TICK, f"rocket fast t={t}")


# --------------------------------------------------------------------------- random worlds vs oracle


@pytest.mark.parametrize("M,N", [(1, 1), (3, 5), (64, 7), (2, 257), (1000, 1)])
@pytest.mark.parametrize("combo", ["free", "rocket", "ball", "falcon9", "wrench_then_drag"])
def test_effector_combos_exact_and_fast(oracle, M, N, combo):
    O = oracle
    pos, vel, ine = random_world(100 + M + N, M, N)
    rng = np.random.default_rng(7)
    specs = {
        "free": [],
        "rocket": [("gravity", {}), ("thrust", {"thrust": rng.uniform(0, 400, (M, N, 1))}),
                   ("wrench", {"wrench": rng.normal(0, 3, (M, N, 6))})],
        "ball": [("gravity", {}), ("drag", {"wind": rng.normal(0, 1, (M, N, 3))})],
        "falcon9": [("frame", {}), ("wrench", {"wrench": rng.normal(0, 1e3, (M, N, 6)), "linear_first": True})],
        "wrench_then_drag": [("wrench", {"wrench": rng.normal(0, 3, (M, N, 6))}), ("drag", {"wind": rng.normal(0, 1, (M, N, 3))}),
                             ("thrust", {"thrust": rng.uniform(0, 40, (M, N, 1))})],
    }[combo]
    if combo == "falcon9":  # near the Earth's surface, ECEF
        pos[..., 4:] = pos[..., 4:] * 1e2 + np.array([6.4e6, 0, 0])
    oeffs, geffs, cols = [], [], {}
    for kind, kw in specs:
        o, g, c = effector_pair(O, kind, **kw)
        oeffs.append(o); geffs.append(g); cols.update(c)
    dt = 0.008333333
    acc0 = rng.normal(0, 1, (M, N, 6))
    want = _run_oracle(O, pos, vel, ine, oeffs, dt, 5, accel=acc0)
    got = _run_gpu(pos, vel, ine, geffs, cols, dt, 5, "exact", accel=acc0)
    _assert_exact(got, want, f"{combo} M={M} N={N}")
    fast = _run_gpu(pos, vel, ine, geffs, cols, dt, 5, "fast", accel=acc0)
    _assert_close(fast, want, 5 * FAST_TOL_TICK, f"{combo} fast M={M} N={N}")


@pytest.mark.parametrize("integrator", ["rk4", "semi_implicit"])
@pytest.mark.parametrize("time_step", [None, 1.0 / 60.0])
def test_integrators_and_dt_override(oracle, integrator, time_step):
    O = oracle
    M, N = 4, 9
    pos, vel, ine = random_world(5, M, N, unit_q=(integrator == "rk4"))
    rng = np.random.default_rng(1)
    o1, g1, c1 = effector_pair(O, "gravity")
    o2, g2, c2 = effector_pair(O, "wrench", wrench=rng.normal(0, 2, (M, N, 6)))
    want = _run_oracle(O, pos, vel, ine, [o1, o2], 0.01, 7, integrator, time_step)
    got = _run_gpu(pos, vel, ine, [g1, g2], {**c1, **c2}, 0.01, 7, "exact", integrator, time_step)
    _assert_exact(got, want, f"{integrator} ts={time_step}")
    fast = _run_gpu(pos, vel, ine, [g1, g2], {**c1, **c2}, 0.01, 7, "fast", integrator, time_step)
    _assert_close(fast, want, 7 * FAST_TOL_TICK, f"{integrator} fast")


@pytest.mark.parametrize("kind", ["softened", "newton"])
@pytest.mark.parametrize("M,N", [(1, 2), (3, 130), (2, 300)])
def test_nbody_dense_gravity(oracle, kind, M, N):
    """All-pairs edge_fold (examples/n-body/sim.py:334-369): the tiled kernel keeps the
    reference's ascending fold order, so EXACT is bit-exact at any N."""
    O = oracle
    pos, vel, ine = random_world(17, M, N)
    rng = np.random.default_rng(3)
    pos[..., 4:] = rng.uniform(-30, 30, (M, N, 3))
    vel[..., 3:] = rng.normal(0, 1e-2, (M, N, 3))
    ine[..., 6] = 10 ** rng.uniform(-6, -3, (M, N))
    edges = el.all_pairs_edges(N)
    kw = {"k2": 2.9591220828e-4, "soft": 1e-10} if kind == "softened" else {"G": 1e-3}
    o, g, _ = effector_pair(O, kind, edges=edges, **kw)
    want = _run_oracle(O, pos, vel, ine, [o], 0.05, 4)
    got = _run_gpu(pos, vel, ine, [g], {}, 0.05, 4, "exact")
    _assert_exact(got, want, f"{kind} dense N={N}")
    fast = _run_gpu(pos, vel, ine, [g], {}, 0.05, 4, "fast")
    _assert_close(fast, want, 1e-11, f"{kind} dense fast N={N}")


def test_sparse_graph_csr_and_order(oracle):
    """Irregular edge list: bodies without out-edges keep the forces of earlier
    effectors; fold order is the spawn order (deliberately shuffled here)."""
    O = oracle
    M, N = 3, 40
    pos, vel, ine = random_world(23, M, N)
    rng = np.random.default_rng(9)
    pos[..., 4:] = rng.uniform(-5, 5, (M, N, 3))
    edges = np.array([(i, j) for i in range(0, N, 2) for j in rng.permutation(N)[:7] if i != j])
    rng.shuffle(edges)
    og, gg, _ = effector_pair(O, "gravity")
    o, g, _ = effector_pair(O, "softened", edges=edges, k2=0.3, soft=1e-6)
    want = _run_oracle(O, pos, vel, ine, [og, o], 0.01, 3)
    got = _run_gpu(pos, vel, ine, [gg, g], {}, 0.01, 3, "exact")
    _assert_exact(got, want, "sparse graph")
    # FAST requires the graph effector first; put gravity after it
    want2 = _run_oracle(O, pos, vel, ine, [o, og], 0.01, 3)
    fast = _run_gpu(pos, vel, ine, [g, gg], {}, 0.01, 3, "fast")
    _assert_close(fast, want2, 1e-11, "sparse graph fast")
    with pytest.raises(el.B200Error):
        _run_gpu(pos, vel, ine, [gg, g], {}, 0.01, 1, "fast")


@pytest.mark.parametrize("N", [1, 2, 3, 7, 16, 31, 32])
@pytest.mark.parametrize("integrator", ["rk4", "semi_implicit"])
def test_small_world_kernel_sparse_and_dense(oracle, N, integrator):
    """Worlds of <= 32 bodies run whole ticks inside one warp (small_world_kernel): gravity through warp
    shuffles in CSR = spawn order, several ticks per launch, ragged worlds per warp (32 % N != 0), bodies
    without out-edges keeping the other effectors' force, a thrust column next to the graph effector.
    EXACT stays bit-identical to the oracle; FAST within tolerance."""
    O = oracle
    M = 41
    pos, vel, ine = random_world(100 + N, M, N)
    rng = np.random.default_rng(N)
    pos[..., 4:] = rng.uniform(-5, 5, (M, N, 3))
    thrust = rng.uniform(0, 3, (M, N, 1))
    graphs = {"dense": el.all_pairs_edges(N)}
    if N >= 3:
        e = np.array([(i, j) for i in range(0, N, 2) for j in rng.permutation(N)[:5] if i != j])
        rng.shuffle(e)
        graphs["sparse"] = e
    for name, edges in graphs.items():
        if len(edges) == 0:
            continue
        og, gg, _ = effector_pair(O, "gravity")
        ot, gt, cols = effector_pair(O, "thrust", thrust=thrust)
        o, g, _ = effector_pair(O, "softened", edges=edges, k2=0.3, soft=1e-6)
        # EXACT keeps array order (gravity first: bodies without edges keep it); FAST needs the graph effector first
        want = _run_oracle(O, pos, vel, ine, [og, o, ot], 0.01, 7, integrator)
        with el.B200Exec(N, M, 0.01, None, [gg, g, gt], integrator, "exact", max_fused_ticks=3) as ex:
            ex.set_state(pos, vel, ine, **cols)
            ex.step(7, sync=True)  # 3 + 3 + 1 ticks per launch
            got = (ex.download(WORLD_POS), ex.download(WORLD_VEL), ex.download(WORLD_ACCEL), ex.download(FORCE))
        _assert_exact(got, want, f"small world N={N} {name} {integrator}")
        want2 = _run_oracle(O, pos, vel, ine, [o, og, ot], 0.01, 7, integrator)
        with el.B200Exec(N, M, 0.01, None, [g, gg, gt], integrator, "fast", max_fused_ticks=4) as ex:
            ex.set_state(pos, vel, ine, **cols)
            ex.step(7, sync=True)
            fast = (ex.download(WORLD_POS), ex.download(WORLD_VEL), ex.download(WORLD_ACCEL), ex.download(FORCE))
        _assert_close(fast, want2, 7 * 1e-11, f"small world fast N={N} {name} {integrator}")


def test_semi_implicit_nbody(oracle):
    O = oracle
    M, N = 2, 33
    pos, vel, ine = random_world(31, M, N)
    pos[..., 4:] *= 1e-2
    o, g, _ = effector_pair(O, "softened", edges=el.all_pairs_edges(N), k2=0.1, soft=1e-4)
    want = _run_oracle(O, pos, vel, ine, [o], 0.01, 6, "semi_implicit")
    got = _run_gpu(pos, vel, ine, [g], {}, 0.01, 6, "exact", "semi_implicit")
    _assert_exact(got, want, "semi-implicit n-body")


# --------------------------------------------------------------------------- structure / plumbing


def test_fused_ticks_equal_single_ticks():
    pos, vel, ine = random_world(2, 50, 3)
    rng = np.random.default_rng(0)
    effs = [el.GravityConst(), el.ThrustBody((0.0, 0.0, 1.0), "thrust")]
    cols = {"thrust": rng.uniform(0, 300, (50, 3, 1))}
    for math in ("exact", "fast"):
        a = _run_gpu(pos, vel, ine, effs, cols, 1e-3, 37, math, fused=1)
        b = _run_gpu(pos, vel, ine, effs, cols, 1e-3, 37, math, fused=16)
        for x, y in zip(a[:2], b[:2]):
            assert np.array_equal(x, y), math
        # Force / WorldAccel leave the batch holding the last tick's stage-4 values either way
        for x, y in zip(a[2:], b[2:]):
            assert np.array_equal(x, y), math


def test_trajectory_ring_matches_states():
    M, N = 6, 4
    pos, vel, ine = random_world(4, M, N)
    effs = [el.GravityConst()]
    with el.B200Exec(N, M, 0.01, None, effs, "rk4", "exact", max_fused_ticks=8, trajectory_every=5,
                     trajectory_capacity=10) as ex:
        ex.set_state(pos, vel, ine)
        snaps = []
        for _ in range(12):
            ex.step(5, sync=True)
            snaps.append(np.concatenate([ex.download(WORLD_POS), ex.download(WORLD_VEL)], -1))
        traj = ex.trajectory()
        assert traj.shape == (10, M, N, 13)  # capacity caps the ring
        for s in range(10):
            assert np.array_equal(traj[s], snaps[s]), s
        assert ex.tick == 60


@pytest.mark.parametrize("effectors", [False, True])
def test_trajectory_sample_on_every_tick_of_the_body_pair_kernel(effectors):
    """Telemetry on every tick at a batch large enough for the body-pair kernel (two bodies per thread): the pair writes
    its (WorldPos, WorldVel) sample as one 16-byte store per plane after the tick; the odd tail body and the 25-plane
    ring keep the per-body stores.  Every sample must equal the state a download after that tick returns."""
    M = 2 * 128 * 3 * 148 + 3  # one full wave of pairs + an odd tail
    pos, vel, ine = random_world(21, M, 1)
    effs, cols = [], {}
    if effectors:
        effs = [el.GravityConst(), el.ThrustBody((-1.0, 0, 0), "thrust")]
        cols = {"thrust": np.random.default_rng(2).uniform(0, 5, (M, 1, 1))}
    for full in (False, True):
        with el.B200Exec(1, M, 0.01, None, effs, "rk4", "fast", max_fused_ticks=1, trajectory_every=1, trajectory_capacity=3,
                         trajectory_full=full) as ex:
            ex.set_state(pos, vel, ine, **cols)
            snaps = []
            for _ in range(4):
                ex.step(1, sync=True)
                snaps.append(np.concatenate([ex.download(WORLD_POS), ex.download(WORLD_VEL)], -1))
            traj = ex.trajectory()
            assert traj.shape == (3, M, 1, 25 if full else 13)
            for k in range(3):
                assert np.array_equal(traj[k][..., :13], snaps[k]), (full, k)
            assert not np.array_equal(snaps[0], snaps[1])


@pytest.mark.parametrize("math", ["exact", "fast"])
@pytest.mark.parametrize("case", ["effectors", "nbody", "semi_implicit"])
def test_full_trajectory_ring_carries_accel_and_force(math, case):
    """B200_TRAJ_FULL: a sample holds all five Body columns' worth of telemetry — (pos, vel, accel, force)
    exactly as a download after that tick returns them — whatever the launch fusing, for free bodies with
    effectors, for the n-body tick (one-launch variant in FAST) and for the semi-implicit integrator."""
    integ = "semi_implicit" if case == "semi_implicit" else "rk4"
    if case == "nbody":
        M, N = 2, 24
        pos, vel, ine = random_world(11, M, N)
        effs, cols = [el.GravityEdges("softened", k_squared=1e-3, softening=1e-6, edges=el.all_pairs_edges(N))], {}
    else:
        M, N = 5, 3
        pos, vel, ine = random_world(12, M, N)
        rng = np.random.default_rng(5)
        effs = [el.GravityConst(), el.ThrustBody(), el.WrenchBody()]
        cols = {"thrust": rng.uniform(1, 5, (M, N, 1)), "aero_force": rng.normal(0, 1, (M, N, 6))}
    every, n_cycles = 3, 7
    def snapshots(full, fused):
        with el.B200Exec(N, M, 0.01, None, effs, integ, math, max_fused_ticks=fused, trajectory_every=every,
                         trajectory_capacity=n_cycles, trajectory_full=full) as ex:
            ex.set_state(pos, vel, ine, **cols)
            assert ex.trajectory_width() == (25 if full else 13)
            snaps = []
            for _ in range(n_cycles):
                ex.step(every, sync=True)
                snaps.append(np.concatenate([ex.download(c) for c in (WORLD_POS, WORLD_VEL, WORLD_ACCEL, FORCE)], -1))
            return np.stack(snaps), ex.trajectory()
    snaps, traj = snapshots(True, 1)
    assert traj.shape == (n_cycles, M, N, 25)
    assert np.array_equal(traj, snaps)
    # one launch for the whole run (ticks fused in registers where the path allows it): same samples
    with el.B200Exec(N, M, 0.01, None, effs, integ, math, max_fused_ticks=32, trajectory_every=every,
                     trajectory_capacity=n_cycles, trajectory_full=True) as ex:
        ex.set_state(pos, vel, ine, **cols)
        ex.step(every * n_cycles, sync=True)
        assert np.array_equal(ex.trajectory(), snaps)
    # and the 13-wide ring is the same run's (pos, vel)
    _, traj13 = snapshots(False, 8)
    assert np.array_equal(traj13, snaps[..., :13])


def test_unknown_trajectory_flags_are_rejected():
    from elodin_b200 import _lib
    import ctypes as C

    d = _lib.Desc()
    d.abi_version, d.n_entities, d.n_worlds, d.sim_time_step, d.time_step, d.device = _lib.ABI_VERSION, 1, 1, 0.01, float("nan"), -1
    d.trajectory_flags = 2
    h = C.c_void_p()
    assert _lib.lib().b200_sixdof_create(C.byref(d), C.byref(h)) == _lib.ERR_INVALID_ARGUMENT
    d.trajectory_flags, d.abi_version = 0, 1
    assert _lib.lib().b200_sixdof_create(C.byref(d), C.byref(h)) == _lib.ERR_INVALID_ARGUMENT  # ABI v1 callers are refused


@pytest.mark.parametrize("B", [1, 31, 32, 33, 255, 257, 1000, 4097])
def test_layout_roundtrip_ragged(B):
    """K6 aos<->soa: upload then download returns the same bytes for ragged sizes."""
    rng = np.random.default_rng(B)
    with el.B200Exec(B, 1, 0.01, None, [el.WrenchBody("aero_force"), el.ThrustBody((1, 0, 0), "thrust")], "rk4", "exact") as ex:
        for cid, w in ((WORLD_POS, 7), (WORLD_VEL, 6), (INERTIA, 7), ("aero_force", 6), ("thrust", 1)):
            a = rng.normal(size=(1, B, w))
            ex.upload(cid, a)
            assert np.array_equal(ex.download(cid), a)


def test_world_axis_is_independent_worlds(oracle):
    """[M, N] batch == M separate executors (the reference runs one process per world)."""
    M, N = 5, 6
    pos, vel, ine = random_world(77, M, N)
    pos[..., 4:] *= 1e-2
    eff = lambda: [el.GravityEdges("softened", k_squared=0.2, softening=1e-5, edges=el.all_pairs_edges(N))]
    both = _run_gpu(pos, vel, ine, eff(), {}, 0.01, 5, "exact")
    for m in range(M):
        one = _run_gpu(pos[m:m + 1], vel[m:m + 1], ine[m:m + 1], eff(), {}, 0.01, 5, "exact")
        for a, b in zip(one, both):
            assert np.array_equal(a[0], b[m])


def test_abi_errors_match_reference_semantics():
    with el.B200Exec(4, 1, 0.01, None, [], "rk4", "exact") as ex:
        with pytest.raises(ValueError):  # Error::ValueSizeMismatch -> ValueError (error.rs:46-58)
            ex.upload(WORLD_POS, np.zeros((1, 3, 7)))
        with pytest.raises(ValueError):  # Error::ComponentNotFound -> ValueError
            ex.upload("no_such_component", np.zeros(4))
        assert ex.input_ids == [el.component_id(n) for n in
                                ("tick", "force", "inertia", "world_pos", "world_accel", "simulation_time_step", "world_vel")]
        assert ex.output_ids == sorted(ex.input_ids)
    with pytest.raises(el.B200Error):
        el.B200Exec(4, 1, -1.0, None, [], "rk4", "exact")  # Error::InvalidTimeStep
    with pytest.raises(el.B200Error):
        el.B200Exec(4, 1, 0.01, None, [el.GravityConst()] * 9, "rk4", "exact")
    # empty world: legal, ticks still advance
    with el.B200Exec(0, 1, 0.01, None, [], "rk4", "exact") as ex:
        ex.step(3, sync=True)
        assert ex.tick == 3


@pytest.mark.parametrize("math", ["exact", "fast"])
def test_invoke_batch_pipelined_world_ranges(oracle, math):
    """invoke_batch splits the world axis into ranges whose PCIe transfers overlap the ticks of
    their neighbours; the result must not depend on the split (ragged last range, graph worlds,
    effector columns, trajectory)."""
    O = oracle
    M, N = 37, 5
    pos, vel, ine = random_world(55, M, N)
    pos[..., 4:] *= 1e-2
    rng = np.random.default_rng(5)
    thrust = rng.uniform(0, 10, (M, N, 1))
    o1, g1, _ = effector_pair(O, "softened", edges=el.all_pairs_edges(N), k2=0.2, soft=1e-5)
    o2, g2, _ = effector_pair(O, "thrust", thrust=thrust)
    acc0 = rng.normal(0, 1, (M, N, 6))
    want = _run_oracle(O, pos, vel, ine, [o1, o2], 0.01, 6, accel=acc0)
    outs = {}
    for chunk in (0, 5 * N, 7 * N, 1000 * N):
        with el.B200Exec(N, M, 0.01, None, [g1, g2], "rk4", math, invoke_chunk_bodies=chunk, trajectory_every=3,
                         trajectory_capacity=4) as ex:
            table = {el.component_id("tick"): np.array([0], dtype=np.uint64), FORCE: rng.normal(size=(M, N, 6)), INERTIA: ine,
                     WORLD_POS: pos, WORLD_ACCEL: acc0, el.component_id("simulation_time_step"): np.array([0.01]),
                     WORLD_VEL: vel, el.component_id("thrust"): thrust}
            o = dict(zip(ex.output_ids, ex.invoke_batch([table[c] for c in ex.input_ids], 6)))
            traj = ex.trajectory()
        got = (o[WORLD_POS], o[WORLD_VEL], o[WORLD_ACCEL], o[FORCE])
        if math == "exact":
            _assert_exact(got, want, f"chunk={chunk}")
        else:
            _assert_close(got, want, 6 * FAST_TOL_TICK, f"chunk={chunk}")
        assert int(o[el.component_id("tick")][0]) == 6
        assert np.array_equal(o[INERTIA], ine) and np.array_equal(o[el.component_id("thrust")], thrust)  # pass-through
        assert np.array_equal(traj[1], np.concatenate([got[0], got[1]], -1))
        outs[chunk] = got
    for chunk, got in outs.items():
        for a, b in zip(got, outs[0]):
            assert np.array_equal(a, b), chunk  # the split never changes a bit


def test_tickfn_shaped_entry():
    pos, vel, ine = random_world(8, 1, 3)
    with el.B200Exec(3, 1, 0.01, None, [], "rk4", "exact") as ex:
        zeros6 = np.zeros((1, 3, 6))
        table = {el.component_id("tick"): np.array([41], dtype=np.uint64), FORCE: zeros6, INERTIA: ine, WORLD_POS: pos,
                 WORLD_ACCEL: zeros6, el.component_id("simulation_time_step"): np.array([0.01]), WORLD_VEL: vel}
        ins = [table[c] for c in ex.input_ids]
        outs = [np.empty_like(table[c]) for c in ex.output_ids]
        ex.tick_fn(ins, outs)
        o = dict(zip(ex.output_ids, outs))
        assert int(o[el.component_id("tick")][0]) == 42
        ref = _run_gpu(pos, vel, ine, [], {}, 0.01, 1, "exact")
        assert np.array_equal(o[WORLD_POS], ref[0]) and np.array_equal(o[WORLD_VEL], ref[1])


# --------------------------------------------------------------------------- full-size properties


def test_full_size_properties_free_body(oracle):
    """BASELINE configs[1] batched (2^20 worlds x 1 body, dt = 1e-3): properties that do
    not need the oracle at full size + an oracle check on a slice."""
    O = oracle
    M = 1 << 20
    pos, vel, ine = random_world(2026, M, 1)
    for math in ("exact", "fast"):
        got = _run_gpu(pos, vel, ine, [], {}, 1e-3, 20, math, fused=20)
        q = got[0][..., :4]
        assert np.max(np.abs(np.linalg.norm(q, axis=-1) - 1.0)) < 4e-16 * 4  # renormalised every tick
        assert np.array_equal(got[1], vel)  # free body: velocity is constant, bit for bit
        lin = pos[..., 4:] + 20 * 1e-3 * vel[..., 3:]
        assert max_rel(got[0][..., 4:], lin) < 1e-13
        sl = slice(0, 4096)
        want = _run_oracle(O, pos[sl], vel[sl], ine[sl], [], 1e-3, 20)
        if math == "exact":
            _assert_exact([g[sl] for g in got], want, "full-size slice")
        else:
            _assert_close([g[sl] for g in got], want, 20 * FAST_TOL_TICK, "full-size slice fast", check_force=False)
    # replicated worlds stay replicated (no cross-talk across the batch axis)
    rep = np.broadcast_to(pos[:1], pos.shape).copy(), np.broadcast_to(vel[:1], vel.shape).copy(), np.broadcast_to(ine[:1], ine.shape).copy()
    got = _run_gpu(*rep, [el.GravityConst()], {}, 1e-3, 10, "fast")
    assert np.all(got[0] == got[0][:1]) and np.all(got[1] == got[1][:1])


def test_fast_drift_over_1000_ticks(oracle):
    """FAST vs oracle after 1000 ticks of the rocket-style world (<= 1e-9)."""
    O = oracle
    M, N = 256, 1
    pos, vel, ine = random_world(42, M, N)
    rng = np.random.default_rng(42)
    specs = [("gravity", {}), ("thrust", {"thrust": rng.uniform(50, 100, (M, N, 1))}),
             ("drag", {"wind": rng.normal(0, 1, (M, N, 3)), "cd_rho": 0.6, "area": 0.01}),
             ("wrench", {"wrench": rng.normal(0, 0.05, (M, N, 6))})]
    # drag resets torque, so put the wrench after it to keep the attitude dynamics alive
    oeffs, geffs, cols = [], [], {}
    for kind, kw in specs:
        o, g, c = effector_pair(O, kind, **kw)
        oeffs.append(o); geffs.append(g); cols.update(c)
    want = _run_oracle(O, pos, vel, ine, oeffs, 0.008333333, 1000)
    exact = _run_gpu(pos, vel, ine, geffs, cols, 0.008333333, 1000, "exact", fused=50)
    _assert_exact(exact, want, "1000 ticks exact")
    fast = _run_gpu(pos, vel, ine, geffs, cols, 0.008333333, 1000, "fast", fused=50)
    _assert_close(fast, want, FAST_TOL_1000, "1000 ticks fast")


# --------------------------------------------------------------------------- the ECS mirror (reads like test_all.py)


def test_six_dof_like_reference_test_all():
    """libs/nox-py/python/tests/test_all.py:67-83 — same script against `el = elodin_b200`."""
    w = el.World()
    w.spawn(el.Body(world_pos=el.SpatialTransform(linear=np.array([0.0, 0.0, 0.0])),
                    world_vel=el.SpatialMotion(linear=np.array([1.0, 0.0, 0.0])),
                    inertia=el.SpatialInertia(1.0)), "e1")
    sys = el.six_dof(1.0 / 60.0)
    exec = w.build(sys)
    exec.run()
    df = exec.history("e1.world_pos")
    x = df["e1.world_pos"][-1]
    assert np.allclose(x.to_numpy()[:4], np.array([0.0, 0.0, 0.0, 1.0]))
    assert np.allclose(x.to_numpy()[4:], np.array([0.01666667, 0.0, 0.0]))


@pytest.mark.parametrize("omega,want", [
    ([0.0, 0.0, 1.0], [0.0, 0.0, 0.479425538604203, 0.8775825618903728, 0.0, 0.0, 0.0]),
    ([0.0, 1.0, 0.0], [0.0, 0.479425538604203, 0.0, 0.8775825618903728, 0.0, 0.0, 0.0]),
    ([1.0, 1.0, 0.0], [0.45936268493243, 0.45936268493243, 0.0, 0.76024459707606, 0.0, 0.0, 0.0]),
])
@pytest.mark.parametrize("math", ["exact", "fast"])
def test_six_dof_ang_vel_int_like_reference(omega, want, math):
    """test_all.py:228-291 (Julia/Simulink values, rtol 1e-5)."""
    w = el.World()
    w.spawn(el.Body(world_pos=el.SpatialTransform(linear=np.array([0.0, 0.0, 0.0])),
                    world_vel=el.SpatialMotion(angular=np.array(omega)), inertia=el.SpatialInertia(1.0)), "e1")
    exec = w.build(el.six_dof(1.0 / 120.0), math=math)
    exec.run(120)
    x = exec.history("e1.world_pos")["e1.world_pos"][-1]
    assert np.isclose(x.to_numpy(), np.array(want), rtol=1e-5).all()


def test_six_dof_force_like_reference():
    """test_all.py:342-366: unit force for 1 s -> x = 0.5."""
    w = el.World()
    w.spawn(el.Body(world_pos=el.SpatialTransform(linear=np.array([0.0, 0.0, 0.0])),
                    world_vel=el.SpatialMotion(angular=np.array([0.0, 0.0, 0.0])), inertia=el.SpatialInertia(1.0)), "e1")
    constant_force = el.GravityConst((1.0, 0.0, 0.0))  # unit mass: g*m == the reference's constant_force
    exec = w.build(el.six_dof(1.0 / 120.0, constant_force))
    exec.run(120)
    df = exec.history(["e1.world_pos", "e1.world_vel", "e1.world_accel"])
    assert np.isclose(df["e1.world_pos"][-1].to_numpy(), np.array([0.0, 0.0, 0.0, 1.0, 0.5, 0.0, 0.0]), rtol=1e-5).all()
    assert np.isclose(df["e1.world_accel"][-1].to_numpy(), [0, 0, 0, 1.0, 0, 0]).all()


def test_three_body_example_through_ecs_mirror(golden):
    """examples/three-body/main.py rebuilt on the mirror: spawn order, entity ids, edges
    and the 100-tick golden, end to end through World.build / Exec.run / history."""
    G = 6.6743e-11
    w = el.World()
    ics = {"a": ([0.8920281421, 0.0, 0.0], [0.0, 0.9957939373, 0.0]),
           "b": ([-0.6628498947, 0.0, 0.0], [0.0, -1.6191613336, 0.0]),
           "c": ([-0.2291782474, 0, 0], [0, 0.6233673964, 0.0])}
    ids = {}
    for k, (x, v) in ics.items():
        ids[k] = w.spawn([el.Body(world_pos=el.WorldPos(linear=np.array(x)),
                                  world_vel=el.WorldVel(linear=np.array(v)),
                                  inertia=el.SpatialInertia(1.0 / G))], name=k.upper())
    assert [int(ids[k]) for k in "abc"] == [1, 2, 3]  # entity 0 = Globals
    GravityEdge = el.Annotated[el.Edge, el.Component("gravity_edge", el.ComponentType.Edge)]

    @el.dataclass
    class GravityConstraint(el.Archetype):
        a: GravityEdge

    for s, d in (("a", "b"), ("b", "a"), ("a", "c"), ("b", "c"), ("c", "a"), ("c", "b")):
        w.spawn(GravityConstraint(el.Edge(ids[s], ids[d])), name=f"{s.upper()} -> {d.upper()}")
    assert np.array_equal(w.edge_rows(), THREE_BODY_EDGES)
    exec = w.build(el.six_dof(sys=el.GravityEdges("newton", G=G)), simulation_rate=120.0)
    assert exec.sim_time_step == 0.008333333
    exec.run(100)
    h = exec.history(["A.world_pos", "B.world_vel", "C.force", "Globals.tick"])
    assert np.array_equal(h["A.world_pos"], golden["three_body.a.world_pos"])
    assert np.array_equal(h["B.world_vel"], golden["three_body.b.world_vel"])
    assert np.array_equal(h["C.force"][1:], golden["three_body.c.force"][1:])
    assert np.array_equal(h["Globals.tick"], np.arange(101))


def test_monte_carlo_campaign_is_one_executor(oracle):
    """Plan -> per-world parameter columns -> one executor == one run per plan row
    (the reference's process-per-world model, libs/monte-carlo/src/lib.rs:2083)."""
    from elodin_b200 import monte_carlo as mc

    O = oracle
    spec = {"monte_carlo": {"n_samples": 12, "seed": 42, "method": "lhs", "variables": {
        "thrust_gain": {"dist": "uniform", "min": 0.8, "max": 1.2}, "mass": {"dist": "uniform", "min": 2.5, "max": 3.5}}}}
    rows = mc.materialize(spec)
    Thrust = el.Annotated[np.ndarray, el.Component("thrust", el.ComponentType.F64)]

    @el.dataclass
    class Motor(el.Archetype):
        thrust: Thrust

    def world():
        w = el.World()
        w.spawn([el.Body(world_pos=el.SpatialTransform(angular=el.Quaternion.from_euler([0.0, np.radians(70.0), 0.0])),
                         inertia=el.SpatialInertia(3.0, np.array([0.1, 1.0, 1.0]))), Motor(np.array([88.426]))], name="rocket")
        return w

    system = lambda: el.six_dof(sys=el.GravityConst((0.0, 0.0, -9.81)) | el.ThrustBody((-1.0, 0.0, 0.0), "thrust"))
    cols = mc.world_params(rows, 1, {"thrust": lambda p: [88.426 * p["thrust_gain"]],
                                     "inertia": lambda p: [0.1, 1.0, 1.0, 0, 0, 0, p["mass"]]})
    campaign = world().build(system(), n_worlds=12, world_params=cols, telemetry_rate=12.0)
    campaign.run(30)
    batch = campaign.history_worlds("rocket.world_pos")
    assert batch.shape == (4, 12, 7)  # initial row + 3 telemetry cycles of 10 ticks
    for k in (0, 5, 11):
        single = world().build(system(), world_params={name: a[k:k + 1] for name, a in cols.items()}, telemetry_rate=12.0)
        single.run(30)
        assert np.array_equal(single.history_worlds("rocket.world_pos")[:, 0], batch[:, k])
        w0 = O.World(campaign._history[el.component_id("world_pos")][0][k:k + 1], np.zeros((1, 1, 6)), cols["inertia"][k:k + 1])
        w0.rk4(campaign.sim_time_step, 30, [O.Effector(O.EFF_GRAVITY_CONST, p=(0, 0, -9.81)),
                                           O.Effector(O.EFF_THRUST_BODY, p=(-1.0, 0, 0), column=cols["thrust"][k:k + 1])])
        assert np.array_equal(w0.pos[0, 0], batch[-1, k])


def test_three_body_csv_export_passes_the_reference_regression_gate(golden, tmp_path):
    """World.run -> export_csv writes the directory `elodin-db export --format csv --flatten`
    would, and it satisfies scripts/ci/compare_baseline_csv.py's checks against the reference's
    own three-body baseline: same file set, same headers (time ignored), same row count, every
    numeric cell within tolerances.json (1e-4) — here in fact bit-identical."""
    import csv
    import json
    import math
    import os

    from elodin_b200.export import export_csv

    G = 6.6743e-11
    w = el.World()
    ids = {}
    for k, x, v in (("A", [0.8920281421, 0.0, 0.0], [0.0, 0.9957939373, 0.0]), ("B", [-0.6628498947, 0.0, 0.0], [0.0, -1.6191613336, 0.0]),
                    ("C", [-0.2291782474, 0, 0], [0, 0.6233673964, 0.0])):
        ids[k] = w.spawn([el.Body(world_pos=el.WorldPos(linear=np.array(x)), world_vel=el.WorldVel(linear=np.array(v)),
                                  inertia=el.Inertia(1.0 / G))], name=k)
    GravityEdge = el.Annotated[el.Edge, el.Component("gravity_edge", el.ComponentType.Edge)]

    @el.dataclass
    class GravityConstraint(el.Archetype):
        a: GravityEdge

    for s_, d in (("A", "B"), ("B", "A"), ("A", "C"), ("B", "C"), ("C", "A"), ("C", "B")):
        w.spawn(GravityConstraint(el.Edge(ids[s_], ids[d])), name=f"{s_} -> {d}")
    ex = w.run(el.six_dof(sys=el.GravityEdges("newton", G=G)), simulation_rate=120.0, max_ticks=100)
    out = os.environ.get("B200_EXPORT_DIR") or str(tmp_path / "three-body-csv")
    export_csv(ex, out)

    with open(os.path.join(os.path.dirname(__file__), "golden", "three_body_csv_layout.json")) as f:
        layout = json.load(f)
    got_files = sorted(fn for fn in os.listdir(out) if fn.endswith(".csv"))
    assert got_files == sorted(layout)  # same file set (incl. a_to_b.gravity_edge.csv, globals.*.csv)
    for fn, want_header in layout.items():
        with open(os.path.join(out, fn), newline="") as f:
            rows = list(csv.reader(f))
        assert rows[0] == want_header, fn  # identical header incl. the greek element names
        assert len(rows) - 1 == 101, fn
        stem = fn[:-4]
        if stem.startswith("globals."):
            ref = golden["three_body." + stem.split(".", 1)[1]]
        elif "gravity_edge" in stem:
            ref = golden[f"three_body.file.{fn}"]
        else:
            ref = golden[f"three_body.{stem}"]
        for r, want in zip(rows[1:], ref):
            for cell, b in zip(r[1:], np.atleast_1d(want)):
                a = float(cell)
                assert math.isclose(a, float(b), rel_tol=1e-4, abs_tol=1e-4)  # the reference gate
                assert a == float(b), (fn, cell, b)                          # and in fact exact

    # the same run through the elodin-db directory format: GPU history -> db -> `export`
    from elodin_b200 import db_sink

    db = str(tmp_path / "three-body-db")
    ex.write_db(db)
    with pytest.raises(FileExistsError):
        ex.write_db(db)  # create_new semantics: never overwrites a database
    _, series, _ = db_sink.read_db(db)
    assert sorted(db_sink._safe_name(n) + ".csv" for n in series) == got_files
    assert np.array_equal(series["b.world_vel"].values, golden["three_body.b.world_vel"])
    assert series["a.world_pos"].timestamps[:3].tolist() == [s0 := int(series["a.world_pos"].timestamps[0]), s0, s0 + 8333]
    out2 = str(tmp_path / "three-body-csv-from-db")
    db_sink.export_db_csv(db, out2)
    for fn in got_files:
        a = [r.split(",")[1:] for r in open(os.path.join(out, fn)).read().splitlines()]
        b = [r.split(",")[1:] for r in open(os.path.join(out2, fn)).read().splitlines()]
        assert a == b, fn                                                    # identical apart from `time`


# --------------------------------------------------------------------------- BASELINE.json configs at full size


def test_config_nbody_1024_full_size(oracle):
    """configs[3]: 1024 bodies, all-pairs softened gravity + 6DOF (SURVEY §8d C4 inputs).  EXACT is
    bit-identical to the oracle at full N (the tiled kernel keeps the sequential fold order); FAST is
    within tolerance; and the size-independent property holds: total linear momentum is conserved
    (pairwise forces cancel) to rounding."""
    O = oracle
    N = 1024
    rng = np.random.default_rng(7)
    pos = np.zeros((1, N, 7)); pos[..., 3] = 1.0; pos[..., 4:] = rng.uniform(-30, 30, (1, N, 3))
    vel = np.zeros((1, N, 6)); vel[..., 3:] = rng.normal(0, 1e-7, (1, N, 3))
    m = 10 ** rng.uniform(-10, -3, (1, N)); m[:, 0] = 1.0
    ine = np.zeros((1, N, 7)); ine[..., :3] = m[..., None]; ine[..., 6] = m
    k2 = 2.9591220828e-4 / 86400.0 ** 2
    edges = el.all_pairs_edges(N)
    o, g, _ = effector_pair(O, "softened", edges=edges, k2=k2, soft=1e-10)
    dt = 3600.0
    want = _run_oracle(O, pos, vel, ine, [o], dt, 2)
    got = _run_gpu(pos, vel, ine, [g], {}, dt, 2, "exact")
    _assert_exact(got, want, "n-body 1024 exact")
    fast = _run_gpu(pos, vel, ine, [g], {}, dt, 2, "fast")
    _assert_close(fast, want, 1e-10, "n-body 1024 fast")
    p0 = np.sum(m[0, :, None] * vel[0, :, 3:], axis=0)
    for name, res in (("exact", got), ("fast", _run_gpu(pos, vel, ine, [g], {}, dt, 50, "fast"))):
        p1 = np.sum(m[0, :, None] * res[1][0, :, 3:], axis=0)
        scale = np.sum(m[0, :, None] * np.abs(res[1][0, :, 3:]))
        assert np.max(np.abs(p1 - p0)) <= 1e-12 * scale, name


def test_config_rocket_10k_worlds_full_size(oracle):
    """configs[2]: 10 000 Monte-Carlo worlds of the rocket (const-g + body thrust + quadratic drag),
    q = euler(0, 70 deg, 0), m = 3, I = [0.1, 1, 1], 120 Hz.  A strided sample of worlds is checked
    against the oracle (EXACT bit-exact, FAST <= 1e-9 after 1000 ticks); worlds with identical
    parameters must produce identical bits wherever they sit in the batch."""
    O = oracle
    M = 10000
    rng = np.random.default_rng(42)
    q = el.Quaternion.from_euler([0.0, np.radians(70.0), 0.0]).arr
    pos = np.tile(np.concatenate([q, [0, 0, 1.0]]), (M, 1, 1))
    vel = np.zeros((M, 1, 6)); vel[..., 3:] = rng.normal(0, 0.1, (M, 1, 3))
    ine = np.tile(np.array([0.1, 1.0, 1.0, 0, 0, 0, 3.0]), (M, 1, 1))
    thrust = 88.426 * rng.uniform(0.8, 1.2, (M, 1, 1))
    wind = rng.normal(0, 2.0, (M, 1, 3))
    # duplicate world 17 at the far end of the batch
    for a in (pos, vel, ine, thrust, wind):
        a[M - 3] = a[17]
    specs = [("gravity", {}), ("thrust", {"thrust": thrust}), ("drag", {"wind": wind, "cd_rho": 0.6125, "area": 0.0025})]
    oeffs, geffs, cols = [], [], {}
    for kind, kw in specs:
        o, g, c = effector_pair(O, kind, **kw)
        oeffs.append(o); geffs.append(g); cols.update(c)
    dt = 0.008333333
    idx = np.arange(0, M, 157)
    sub = lambda a: np.ascontiguousarray(a[idx])
    sub_effs = [O.Effector(O.EFF_GRAVITY_CONST, p=(0, 0, -9.81)), O.Effector(O.EFF_THRUST_BODY, p=(-1.0, 0, 0), column=sub(thrust)),
                O.Effector(O.EFF_DRAG_QUADRATIC, p=(0.6125, 0.0025), column=sub(wind))]
    want = _run_oracle(O, sub(pos), sub(vel), sub(ine), sub_effs, dt, 1000)
    exact = _run_gpu(pos, vel, ine, geffs, cols, dt, 1000, "exact", fused=100)
    _assert_exact([a[idx] for a in exact], want, "rocket 10k exact sample")
    fast = _run_gpu(pos, vel, ine, geffs, cols, dt, 1000, "fast", fused=100)
    _assert_close([a[idx] for a in fast], want, FAST_TOL_1000, "rocket 10k fast sample")
    for res in (exact, fast):
        for a in res:
            assert np.array_equal(a[M - 3], a[17])
        assert np.all(np.isfinite(res[0])) and np.all(np.isfinite(res[1]))
        assert np.max(np.abs(np.linalg.norm(res[0][..., :4], axis=-1) - 1.0)) < 1e-15 * 8


def test_config_falcon9_style_worlds(oracle):
    """configs[4] per-GPU shard: 12 500 worlds, dt = 1e-3, rotating-frame gravity + body wrench
    ([f, tau] layout); strided sample vs the oracle."""
    O = oracle
    M = 12500
    rng = np.random.default_rng(20170814)
    pos = np.tile(np.array([0, 0, 0, 1.0, 6.4e6, 0, 0]), (M, 1, 1))
    pos[..., 4:] += rng.normal(0, 10, (M, 1, 3))
    vel = np.concatenate([rng.normal(0, 0.01, (M, 1, 3)), rng.normal(0, 50, (M, 1, 3))], -1)
    ine = np.tile(np.array([4e6, 4e6, 1e5, 0, 0, 0, 3e4]), (M, 1, 1))
    wrench = rng.normal(0, 1e4, (M, 1, 6))
    idx = np.arange(0, M, 211)
    sub = lambda a: np.ascontiguousarray(a[idx])
    of, gf, _ = effector_pair(O, "frame")
    ow = O.Effector(O.EFF_WRENCH_BODY, flags=O.FLAG_WRENCH_LINEAR_FIRST, column=sub(wrench))
    gw = el.WrenchBody("aero_force", "linear_first")
    want = _run_oracle(O, sub(pos), sub(vel), sub(ine), [of, ow], 1e-3, 500)
    exact = _run_gpu(pos, vel, ine, [gf, gw], {"aero_force": wrench}, 1e-3, 500, "exact", fused=100)
    _assert_exact([a[idx] for a in exact], want, "falcon9 exact sample")
    fast = _run_gpu(pos, vel, ine, [gf, gw], {"aero_force": wrench}, 1e-3, 500, "fast", fused=100)
    _assert_close([a[idx] for a in fast], want, FAST_TOL_1000, "falcon9 fast sample")


def test_host_system_feeds_per_tick_inputs(oracle):
    """`non_effectors | six_dof(...)` (examples/rocket/main.py:560-576): a per-tick host system
    (here a thrust curve indexed by tick, like rocket/main.py:416-426) updates an effector input
    column between GPU ticks; the result equals the oracle stepped tick by tick with the same values."""
    O = oracle
    Thrust = el.Annotated[np.ndarray, el.Component("thrust", el.ComponentType.F64)]

    @el.dataclass
    class Motor(el.Archetype):
        thrust: Thrust

    w = el.World()
    q = el.Quaternion.from_euler([0.0, np.radians(70.0), 0.0])
    w.spawn([el.Body(world_pos=el.SpatialTransform(angular=q, linear=np.array([0.0, 0.0, 1.0])),
                     inertia=el.SpatialInertia(3.0, np.array([0.1, 1.0, 1.0]))), Motor(np.array([0.0]))], name="rocket")
    curve = lambda tick: 300.0 * np.exp(-0.05 * tick)

    @el.host_system
    def thrust(ctx):
        ctx.column("thrust")[...] = curve(ctx.tick)

    effectors = el.GravityConst((0.0, 0.0, -9.81)) | el.ThrustBody((-1.0, 0.0, 0.0), "thrust")
    ex = w.build(thrust | el.six_dof(sys=effectors, integrator=el.Integrator.Rk4), simulation_rate=120.0)
    ex.run(25)
    ow = O.World(np.concatenate([q.arr, [0, 0, 1.0]])[None, None], np.zeros((1, 1, 6)), np.array([[[0.1, 1.0, 1.0, 0, 0, 0, 3.0]]]))
    for t in range(25):
        ow.rk4(ex.sim_time_step, 1, [O.Effector(O.EFF_GRAVITY_CONST, p=(0, 0, -9.81)),
                                     O.Effector(O.EFF_THRUST_BODY, p=(-1.0, 0, 0), column=np.array([[[curve(t)]]]))])
    h = ex.history(["rocket.world_pos", "rocket.world_vel", "rocket.thrust", "Globals.tick"])
    assert np.array_equal(h["rocket.world_pos"][-1], ow.pos[0, 0]) and np.array_equal(h["rocket.world_vel"][-1], ow.vel[0, 0])
    assert h["rocket.thrust"][-1][0] == curve(24) and int(h["Globals.tick"][-1]) == 25


def test_per_world_drag_parameters(oracle):
    """quadratic drag with per-world Cd*rho*A — the drag column carries
    [wind(3), Cd*rho, area] per body."""
    O = oracle
    M, N = 40, 2
    pos, vel, ine = random_world(91, M, N)
    rng = np.random.default_rng(2)
    col = np.concatenate([rng.normal(0, 1, (M, N, 3)), rng.uniform(0.3, 0.9, (M, N, 1)), rng.uniform(0.001, 0.01, (M, N, 1))], -1)
    og = O.Effector(O.EFF_GRAVITY_CONST, p=(0, 0, -9.81))
    od = O.Effector(O.EFF_DRAG_QUADRATIC, p=(9e9, 9e9), column=col)  # constants must be ignored
    want = _run_oracle(O, pos, vel, ine, [og, od], 0.01, 8)
    effs = [el.GravityConst(), el.DragQuadratic(9e9, 9e9, "wind", per_body_params=True)]
    got = _run_gpu(pos, vel, ine, effs, {"wind": col}, 0.01, 8, "exact")
    _assert_exact(got, want, "per-world drag")
    fast = _run_gpu(pos, vel, ine, effs, {"wind": col}, 0.01, 8, "fast")
    _assert_close(fast, want, 8 * FAST_TOL_TICK, "per-world drag fast")
    # and it differs from constant parameters (the column values are really used)
    other = _run_gpu(pos, vel, ine, [el.GravityConst(), el.DragQuadratic(0.6, 0.005, "wind")], {"wind": col[..., :3]}, 0.01, 8, "exact")
    assert not np.array_equal(other[1], got[1])


@pytest.mark.parametrize("n_ticks", [1, 5, 6])
def test_fused_nbody_tick_ping_pong(oracle, n_ticks):
    """Small-grid FAST n-body runs gravity + integration in one launch with ping-pong pose/velocity
    planes: odd and even tick counts, step() and chunked invoke_batch, repeated calls."""
    O = oracle
    M, N = 9, 70
    pos, vel, ine = random_world(61, M, N)
    pos[..., 4:] *= 1e-2
    o, g, _ = effector_pair(O, "softened", edges=el.all_pairs_edges(N), k2=0.3, soft=1e-5)
    want = _run_oracle(O, pos, vel, ine, [o], 0.01, 2 * n_ticks)
    with el.B200Exec(N, M, 0.01, None, [g], "rk4", "fast", invoke_chunk_bodies=4 * N) as ex:
        ex.set_state(pos, vel, ine)
        ex.step(n_ticks, sync=True)                      # first half through step()
        mid = (ex.download(WORLD_POS), ex.download(WORLD_VEL))
        table = {el.component_id("tick"): np.array([n_ticks], dtype=np.uint64), FORCE: np.zeros((M, N, 6)), INERTIA: ine,
                 WORLD_POS: mid[0], WORLD_ACCEL: np.zeros((M, N, 6)), el.component_id("simulation_time_step"): np.array([0.01]),
                 WORLD_VEL: mid[1]}
        out = dict(zip(ex.output_ids, ex.invoke_batch([table[c] for c in ex.input_ids], n_ticks)))  # second half, chunked
        got = (out[WORLD_POS], out[WORLD_VEL], out[WORLD_ACCEL], out[FORCE])
        _assert_close(got, want, 1e-11, f"fused n-body {n_ticks}")
        # the device-resident state agrees with what invoke_batch returned
        assert np.array_equal(ex.download(WORLD_POS), out[WORLD_POS]) and np.array_equal(ex.download(WORLD_VEL), out[WORLD_VEL])
        assert int(out[el.component_id("tick")][0]) == 2 * n_ticks


def test_cube_sat_earth_semi_implicit_golden(golden):
    """SemiImplicit on the GPU vs the reference's cube-sat golden (`earth` entity, 100 rows):
    EXACT bit for bit through the trajectory ring, FAST within tolerance."""
    g = golden
    dt = float(g["cube_sat.simulation_time_step"][0, 0])
    p0, v0, i0 = (g[f"cube_sat.earth.{c}"][0][None, None] for c in ("world_pos", "world_vel", "inertia"))
    with el.B200Exec(1, 1, dt, None, [], "semi_implicit", "exact", max_fused_ticks=10, trajectory_every=1,
                     trajectory_capacity=100) as ex:
        ex.set_state(p0, v0, i0)
        ex.step(100, sync=True)
        traj = ex.trajectory()
        acc = ex.download(WORLD_ACCEL)
    assert np.array_equal(traj[:, 0, 0, :7], g["cube_sat.earth.world_pos"][1:])
    assert np.array_equal(traj[:, 0, 0, 7:], g["cube_sat.earth.world_vel"][1:])
    assert np.array_equal(acc[0, 0], g["cube_sat.earth.world_accel"][100])
    fast = _run_gpu(p0, v0, i0, [], {}, dt, 100, "fast", "semi_implicit")
    assert max_rel(fast[0][0], g["cube_sat.earth.world_pos"][100][None]) < 1e-13


# --------------------------------------------------------------------------- fuzz + edge cases


def _random_program(rng, M, N):
    """A random ordered effector list (kinds may repeat in EXACT) with its columns."""
    kinds = ["gravity", "thrust", "wrench", "drag", "frame", "wrench_lin"]
    n = int(rng.integers(0, 6))
    specs, have_cols = [], set()
    for k in rng.choice(kinds, size=n, replace=True):
        if k == "gravity":
            specs.append(("gravity", {"g": tuple(rng.normal(0, 5, 3))}))
        elif k == "thrust" and "thrust" not in have_cols:
            have_cols.add("thrust")
            specs.append(("thrust", {"thrust": rng.uniform(-50, 300, (M, N, 1)), "axis": tuple(rng.normal(0, 1, 3))}))
        elif k in ("wrench", "wrench_lin") and "aero_force" not in have_cols:
            have_cols.add("aero_force")
            specs.append(("wrench", {"wrench": rng.normal(0, 4, (M, N, 6)), "linear_first": k == "wrench_lin"}))
        elif k == "drag" and "wind" not in have_cols:
            have_cols.add("wind")
            specs.append(("drag", {"wind": rng.normal(0, 2, (M, N, 3)), "cd_rho": float(rng.uniform(0.1, 1)), "area": float(rng.uniform(0.001, 0.1))}))
        elif k == "frame" and "frame" not in have_cols:
            have_cols.add("frame")
            specs.append(("frame", {"mu": float(rng.uniform(1e3, 1e5)), "omega": tuple(rng.normal(0, 1e-2, 3))}))
    return specs


@pytest.mark.parametrize("case", range(24))
def test_fuzz_random_effector_programs(oracle, case):
    """Random ordered effector programs x sizes x integrator x tick fusion x invoke chunking:
    EXACT == oracle bit for bit; FAST within tolerance."""
    O = oracle
    rng = np.random.default_rng(9000 + case)
    M, N = int(rng.integers(1, 40)), int(rng.integers(1, 9))
    ticks = int(rng.integers(1, 9))
    integrator = "rk4" if rng.random() < 0.7 else "semi_implicit"
    time_step = None if rng.random() < 0.6 else float(rng.uniform(0.002, 0.02))
    fused = int(rng.choice([1, 2, 7, 64]))
    pos, vel, ine = random_world(7000 + case, M, N, unit_q=rng.random() < 0.8)
    pos[..., 4:] = pos[..., 4:] * 0.1 + 50.0  # keep |r| away from 0 for the frame effector
    specs = _random_program(rng, M, N)
    oeffs, geffs, cols = [], [], {}
    for kind, kw in specs:
        o, g, c = effector_pair(O, kind, **kw)
        oeffs.append(o); geffs.append(g); cols.update(c)
    acc0 = rng.normal(0, 1, (M, N, 6))
    dt = float(rng.uniform(1e-3, 1e-2))
    want = _run_oracle(O, pos, vel, ine, oeffs, dt, ticks, integrator, time_step, accel=acc0)
    got = _run_gpu(pos, vel, ine, geffs, cols, dt, ticks, "exact", integrator, time_step, fused, accel=acc0)
    _assert_exact(got, want, f"fuzz {case} {[k for k, _ in specs]} M={M} N={N} {integrator}")
    fast = _run_gpu(pos, vel, ine, geffs, cols, dt, ticks, "fast", integrator, time_step, fused, accel=acc0)
    _assert_close(fast, want, ticks * 2e-12, f"fuzz fast {case}")


def test_edge_cases_nan_inf_zero(oracle):
    """Degenerate inputs behave like the reference arithmetic: zero quaternion (0/0), zero mass and
    zero inertia (x/0), zero relative wind in the drag (0/0), huge and denormal values.  EXACT must
    agree with the oracle including where the NaNs / infs land."""
    O = oracle
    M, N = 1, 8
    pos, vel, ine = random_world(3, M, N)
    pos[0, 0, :4] = 0.0                      # zero quaternion -> NaN attitude
    ine[0, 1, 6] = 0.0                       # zero mass -> inf / NaN linear accel
    ine[0, 2, 0] = 0.0                       # zero Ixx
    vel[0, 3, 3:] = [1.0, 2.0, 3.0]          # wind == velocity -> drag direction 0/0
    pos[0, 4, 4:] = 1e300                    # huge
    vel[0, 5, :] = 5e-324                    # denormal
    pos[0, 6, 4:] = -0.0                     # signed zeros
    vel[0, 6, 3:] = -0.0
    wind = np.zeros((M, N, 3)); wind[0, 3] = [1.0, 2.0, 3.0]; wind[0, [0, 1, 2, 4, 5, 6, 7]] = 0.5
    og, gg, _ = effector_pair(O, "gravity")
    od, gd, c1 = effector_pair(O, "drag", wind=wind, cd_rho=0.6, area=0.01)
    ow, gw, c2 = effector_pair(O, "wrench", wrench=np.full((M, N, 6), 0.25))
    want = _run_oracle(O, pos, vel, ine, [og, od, ow], 0.01, 3)
    got = _run_gpu(pos, vel, ine, [gg, gd, gw], {**c1, **c2}, 0.01, 3, "exact")
    for name, a, b in zip(("pos", "vel", "accel", "force"), got, want):
        assert np.array_equal(a, b, equal_nan=True), name
        assert np.array_equal(np.signbit(a[np.isfinite(b)]), np.signbit(b[np.isfinite(b)])), name  # signed zeros too
    assert np.isnan(got[0][0, 0]).any() and np.isnan(got[1][0, 3, 3:]).all()  # the degenerate bodies really are degenerate
    assert np.isfinite(got[0][0, 7]).all()                                      # and they do not contaminate a healthy one
    fast = _run_gpu(pos, vel, ine, [gg, gd, gw], {**c1, **c2}, 0.01, 3, "fast")
    assert np.isnan(fast[1][0, 3, 3:]).all() and np.isnan(fast[0][0, 0, :4]).all()  # FAST keeps the 0/0 semantics
    ok = [6, 7]
    _assert_close([a[:, ok] for a in fast], [b[:, ok] for b in want], 3 * FAST_TOL_TICK, "healthy bodies, fast")


def test_effectors_follow_the_query_join(oracle):
    """Heterogeneous world, like the reference's mixed archetypes (ball: only the ball owns `wind`;
    cube-sat: earth / satellite / wheels): an effector runs only on entities that own its input
    component (query.rs:672-710).  ECS mirror -> entity masks -> kernels == oracle with the same masks."""
    O = oracle
    Wind = el.Annotated[np.ndarray, el.Component("wind", el.ComponentType(el.PrimitiveType.F64, (3,)))]
    Thrust = el.Annotated[np.ndarray, el.Component("thrust", el.ComponentType.F64)]

    @el.dataclass
    class WindData(el.Archetype):
        wind: Wind

    @el.dataclass
    class Motor(el.Archetype):
        thrust: Thrust

    rng = np.random.default_rng(12)
    w = el.World()
    pos, vel, ine = random_world(12, 1, 4)
    arch = lambda i: el.Body(world_pos=el.SpatialTransform(arr=pos[0, i]), world_vel=el.SpatialMotion(angular=vel[0, i, :3], linear=vel[0, i, 3:]),
                             inertia=el.SpatialInertia(ine[0, i, 6], ine[0, i, :3]))
    w.spawn([arch(0)], name="plain")                                        # neither wind nor thrust
    w.spawn([arch(1), WindData(np.array([1.0, -2.0, 0.5]))], name="ball")   # drag only
    w.spawn([arch(2), Motor(np.array([40.0]))], name="rocket")              # thrust only
    w.spawn([arch(3), WindData(np.array([0.0, 3.0, 0.0])), Motor(np.array([15.0]))], name="both")
    system = el.six_dof(sys=el.GravityConst((0.0, 0.0, -9.81)) | el.ThrustBody((-1.0, 0.0, 0.0), "thrust") | el.DragQuadratic(0.6, 0.05, "wind"))
    wind = np.zeros((1, 4, 3)); wind[0, 1] = [1.0, -2.0, 0.5]; wind[0, 3] = [0.0, 3.0, 0.0]
    thrust = np.zeros((1, 4, 1)); thrust[0, 2] = 40.0; thrust[0, 3] = 15.0
    oeffs = [O.Effector(O.EFF_GRAVITY_CONST, p=(0, 0, -9.81)),
             O.Effector(O.EFF_THRUST_BODY, p=(-1.0, 0, 0), column=thrust, mask=[0, 0, 1, 1]),
             O.Effector(O.EFF_DRAG_QUADRATIC, p=(0.6, 0.05), column=wind, mask=[0, 1, 0, 1])]
    for math in ("exact", "fast"):
        ex = w.build(system, simulation_rate=120.0, math=math)
        ex.run(12)
        ow = O.World(pos, vel, ine).rk4(ex.sim_time_step, 12, oeffs)
        for i, name in enumerate(("plain", "ball", "rocket", "both")):
            h = ex.history([f"{name}.world_pos", f"{name}.world_vel", f"{name}.force"])
            got = (h[f"{name}.world_pos"][-1], h[f"{name}.world_vel"][-1], h[f"{name}.force"][-1])
            want = (ow.pos[0, i], ow.vel[0, i], ow.force[0, i])
            for a, b in zip(got, want):
                if math == "exact":
                    assert np.array_equal(a, b), (name, math)
                else:
                    assert np.max(np.abs(a - b)) <= 12 * FAST_TOL_TICK * max(np.max(np.abs(b)), 1e-300), (name, math)
        # the plain body saw gravity only: its force is exactly m*g with zero torque
        f_plain = ex.history("plain.force")["plain.force"][-1]
        assert np.array_equal(f_plain[:3], [0, 0, 0]) and np.isclose(f_plain[5], -9.81 * ine[0, 0, 6], rtol=1e-14, atol=0)
        if math == "exact":
            assert f_plain[5] == -9.81 * ine[0, 0, 6]
        # the drag quirk (torque reset) only hits members: "rocket" keeps no torque anyway, "both" is zeroed by drag
        assert ex.history("ball.wind")["ball.wind"].shape == (13, 3)


@pytest.mark.parametrize("math", ["exact", "fast"])
@pytest.mark.parametrize("telemetry_rate", [None, 24.0])
def test_resident_run_equals_invoke_batch_run(math, telemetry_rate):
    """Exec.run without host callbacks keeps the state on the device and reads all telemetry back from the
    full trajectory ring; the rows it records — every component, every cycle, plus a ragged tail that goes
    through invoke_batch — are the rows the one-invoke-per-cycle route records, bit for bit (heterogeneous
    world with a partial-membership effector column, and an n-body world)."""
    Thrust = el.Annotated[np.ndarray, el.Component("thrust", el.ComponentType.F64)]

    @el.dataclass
    class Motor(el.Archetype):
        thrust: Thrust

    pos, vel, ine = random_world(21, 1, 5)
    arch = lambda i: el.Body(world_pos=el.SpatialTransform(arr=pos[0, i]), world_vel=el.SpatialMotion(angular=vel[0, i, :3], linear=vel[0, i, 3:]),
                             inertia=el.SpatialInertia(ine[0, i, 6], ine[0, i, :3]))

    def mixed():
        w = el.World()
        w.spawn([arch(0)], name="plain")
        w.spawn([arch(1), Motor(np.array([40.0]))], name="rocket")
        w.spawn([arch(2)], name="third")
        return w, el.six_dof(sys=el.GravityConst((0.0, 0.0, -9.81)) | el.ThrustBody((-1.0, 0.0, 0.0), "thrust")), ("plain", "rocket", "third")

    def nbody():
        w = el.World()
        for i in range(5):
            w.spawn([arch(i)], name=f"p{i}")
        return w, el.six_dof(sys=el.GravityEdges("softened", k_squared=1e-2, softening=1e-6, edges=el.all_pairs_edges(5))), tuple(f"p{i}" for i in range(5))

    for make in (mixed, nbody):
        runs = []
        for resident in (True, False):
            w, system, names = make()
            ex = w.build(system, simulation_rate=120.0, telemetry_rate=telemetry_rate, math=math, n_worlds=3, resident=resident)
            assert bool(ex._ring_cap) == resident
            ex.run(23)            # telemetry_rate 24 -> 5 ticks per cycle: 4 resident cycles + a 3-tick tail
            ex.run(10)            # a second run() re-uploads the host columns and carries on
            runs.append((ex, names))
        (a, names), (b, _) = runs
        assert a.tick == b.tick == 33
        if telemetry_rate is None:  # one cycle per tick: the resident route has no per-cycle layout launches
            assert a.backend.timings()["kernel_launches"] < b.backend.timings()["kernel_launches"]
        for name in names:
            for comp in ("world_pos", "world_vel", "world_accel", "force", "inertia"):
                ha, hb = a.history_worlds(f"{name}.{comp}"), b.history_worlds(f"{name}.{comp}")
                assert ha.shape == hb.shape and np.array_equal(ha, hb), (make.__name__, name, comp)
        assert np.array_equal(a.history("globals.tick")["globals.tick"], b.history("globals.tick")["globals.tick"])
        for cid, col in a.world.columns.items():
            assert np.array_equal(col.buffer, b.world.columns[cid].buffer)


def test_two_devices_in_one_process():
    """One handle per GPU inside a single process (the C ABI selects the device per call): both
    produce the bits of a single-GPU run, including the >48 KB dynamic-shared-memory kernels."""
    if el.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    M, N = 2, 96
    pos, vel, ine = random_world(71, M, N)
    pos[..., 4:] *= 1e-2
    mk = lambda dev, math: el.B200Exec(N, M, 0.01, None, [el.GravityEdges("softened", k_squared=0.2, softening=1e-5,
                                                                           edges=el.all_pairs_edges(N))], "rk4", math, device=dev)
    for math in ("exact", "fast"):
        a, b = mk(0, math), mk(1, math)
        for ex in (a, b):
            ex.set_state(pos, vel, ine)
        for _ in range(3):  # interleave the two devices
            a.step(2); b.step(2)
        a.sync(); b.sync()
        ra, rb = (a.download(WORLD_POS), a.download(WORLD_VEL)), (b.download(WORLD_POS), b.download(WORLD_VEL))
        a.close(); b.close()
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]), math


def test_fast_math_is_run_to_run_deterministic():
    """FAST changes the summation order of the gravity fold (warp-shuffle butterfly) but the order is
    fixed: two executors fed the same inputs return the same bits, whatever the tick fusion or the
    invoke range size."""
    M, N = 5, 150
    pos, vel, ine = random_world(81, M, N)
    pos[..., 4:] *= 1e-2
    rng = np.random.default_rng(8)
    thrust = rng.uniform(0, 3, (M, N, 1))
    effs = lambda: [el.GravityEdges("softened", k_squared=0.2, softening=1e-5, edges=el.all_pairs_edges(N)),
                    el.ThrustBody((0.0, 1.0, 0.0), "thrust")]
    runs = []
    for chunk in (0, 2 * N):
        with el.B200Exec(N, M, 0.01, None, effs(), "rk4", "fast", invoke_chunk_bodies=chunk) as ex:
            ex.set_state(pos, vel, ine, thrust=thrust)
            ex.step(7, sync=True)
            runs.append((ex.download(WORLD_POS), ex.download(WORLD_VEL), ex.download(FORCE)))
    for a, b in zip(*runs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("lang", ["c", "cpp"])
def test_compiled_consumers_of_the_abi_run_on_the_gpu(lang, tmp_path):
    """The plain-C consumer (tests/c/abi_smoke.c: 64 ticks of a free body, x = v t) and the C++17 host mirror
    (tests/cpp/world_exec_test.cpp over include/b200_world.hpp: the reference's test_six_dof, three-body ticks
    1 and 100 against the golden telemetry bit for bit, the error mapping) — built here, run on the device."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "elodin_b200")
    exe = tmp_path / f"consumer_{lang}"
    if lang == "c":
        cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", os.path.join(root, "tests", "c", "abi_smoke.c")]
        want = "C ABI ok"
    else:
        cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", os.path.join(root, "tests", "cpp", "world_exec_test.cpp")]
        want = "C++ host mirror ok"
    subprocess.run(cmd + ["-I", os.path.join(root, "include"), "-L", lib_dir, "-lb200_sixdof", "-lm", "-Wl,-rpath," + lib_dir,
                          "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and want in out.stdout, out.stdout + out.stderr


# --------------------------------------------------------------------------- round 2: dirty inputs / unread outputs
@pytest.mark.parametrize("M,N", [(3, 2), (3000, 1)])  # packed small path and pipelined world ranges
@pytest.mark.parametrize("math", ["exact", "fast"])
def test_invoke_batch_null_columns_mean_not_dirty_and_not_read(oracle, math, M, N):
    """in_cols[i] = None: the column is not dirty (World::dirty_components, world.rs:43,249-252), the device copy
    stands; out_cols[j] = None: not read back.  Two calls that only hand over what changed must equal two full calls."""
    O = oracle
    pos, vel, ine = random_world(91, M, N)
    rng = np.random.default_rng(3)
    thrust = rng.uniform(0, 10, (M, N, 1))
    o2, g2, _ = effector_pair(O, "thrust", thrust=thrust)
    want = _run_oracle(O, pos, vel, ine, [o2], 0.01, 7)
    tick, dt = el.component_id("tick"), el.component_id("simulation_time_step")
    with el.B200Exec(N, M, 0.01, None, [g2], "rk4", math, invoke_chunk_bodies=0 if M < 100 else 1024) as ex:
        table = {tick: np.array([0], dtype=np.uint64), FORCE: np.zeros((M, N, 6)), INERTIA: ine, WORLD_POS: pos,
                 WORLD_ACCEL: np.zeros((M, N, 6)), dt: np.array([0.01]), WORLD_VEL: vel, el.component_id("thrust"): thrust}
        # call 1: everything uploaded, nothing read back
        ex.invoke_batch([table[c] for c in ex.input_ids], 3, out_cols=[None] * len(ex.output_ids))
        # call 2: nothing is dirty; read back the state, Force and the pass-through Inertia, not WorldAccel / thrust
        outs = [None if c in (WORLD_ACCEL, el.component_id("thrust")) else np.full(ex.column_shape(c), np.nan, dtype=np.uint64 if c == tick else np.float64)
                for c in ex.output_ids]
        ex.invoke_batch([None] * len(ex.input_ids), 4, out_cols=outs)
        o = dict(zip(ex.output_ids, outs))
        acc = ex.download(WORLD_ACCEL)
    got = (o[WORLD_POS], o[WORLD_VEL], acc, o[FORCE])
    if math == "exact":
        _assert_exact(got, want)
    else:
        _assert_close(got, want, 7 * FAST_TOL_TICK)
    assert int(o[tick][0]) == 7
    assert np.array_equal(o[INERTIA], ine)  # pass-through of a non-dirty input = the device-resident column


@pytest.mark.parametrize("integrator", ["rk4", "semi_implicit"])
def test_signature_kernels_match_the_interpreter_kernel(oracle, integrator):
    """Every effector signature with a compiled FAST kernel (one body and body pairs per thread, odd and even range
    lengths) against the oracle, and masked / repeated lists that must take the run-time interpreter."""
    O = oracle
    rng = np.random.default_rng(17)
    for M in (1, 2, 257, 120001):  # 120001 >= one wave of body pairs: the double2 kernel with an odd tail
        pos, vel, ine = random_world(400 + M, M, 1)
        pos[..., 4:] += np.array([6.4e6, 0.0, 0.0])
        wind = rng.normal(0, 3, (M, 1, 3))
        thrust = rng.uniform(0, 10, (M, 1, 1))
        wrench = rng.normal(0, 5, (M, 1, 6))
        wheels = rng.normal(0, 2e-2, (M, 1, 9))
        lists = {
            "free": [],
            "g": [("gravity", {})],
            "drag": [("gravity", {}), ("drag", dict(wind=wind))],
            "rocket": [("gravity", {}), ("thrust", dict(thrust=thrust)), ("drag", dict(wind=wind))],
            "rocket_golden": [("gravity", {}), ("thrust", dict(thrust=thrust)), ("wrench", dict(wrench=wrench))],
            "falcon9": [("frame", {}), ("wrench", dict(wrench=wrench, linear_first=True))],
            "frame": [("frame", {})],
            "wrench": [("wrench", dict(wrench=wrench))],
            "thrust": [("thrust", dict(thrust=thrust))],
            "two_g (summed)": [("gravity", {}), ("gravity", dict(g=(0.5, 0.0, 1.0)))],
            "j2": [("j2", {})],
            "cube_sat (wheel fold first, then J2)": [("wheels", dict(torques=wheels)), ("j2", {})],
            "g then wheels (interpreter: the fold overwrites)": [("gravity", {}), ("wheels", dict(torques=wheels))],
            "external world wrench": [("wrench_world", dict(wrench=wrench))],
            "wheels + external gravity (the cube-sat golden's replay shape)": [("wheels", dict(torques=wheels)), ("wrench_world", dict(wrench=wrench))],
            "wrench_then_drag (interpreter: torque reset)": [("wrench", dict(wrench=wrench)), ("drag", dict(wind=wind))],
        }
        for name, spec in lists.items():
            oe, ge, cols = [], [], {}
            for kind, kw in spec:
                a, b, c = effector_pair(O, kind, **kw)
                oe.append(a); ge.append(b); cols.update(c)
            want = _run_oracle(O, pos, vel, ine, oe, 0.01, 3, integrator)
            got = _run_gpu(pos, vel, ine, ge, cols, 0.01, 3, "fast", integrator)
            _assert_close(got, want, 3 * FAST_TOL_TICK, f"M={M} {name}")


# --------------------------------------------------------------------------- round 2: §8f-4 effectors
@pytest.mark.parametrize("integrator", ["rk4", "semi_implicit"])
def test_wrench_world_wheel_fold_and_j2_effectors(oracle, integrator):
    """WRENCH_WORLD, TORQUE_BODY_FOLD (cube-sat/main.py:492-505) and GRAVITY_J2 (j2.py:5-29): EXACT bit-identical to
    the oracle for the first two (J2's `norm**6.0` is a pow() call: <= 1e-14), FAST within tolerance; the fold
    overwrites what earlier effectors accumulated, as every edge_fold does."""
    O = oracle
    rng = np.random.default_rng(23)
    M, N = 5, 3
    pos, vel, ine = random_world(77, M, N)
    pos[..., 4:] = rng.normal(size=(M, N, 3))
    pos[..., 4:] *= 6.9e6 / np.linalg.norm(pos[..., 4:], axis=-1, keepdims=True)
    wr = rng.normal(0, 3, (M, N, 6))
    tq = rng.normal(0, 2e-3, (M, N, 9))
    combos = {
        "wrench_world": [("wrench_world", dict(wrench=wr))],
        "wheels": [("wheels", dict(torques=tq))],
        "gravity then wheels (fold overwrites) then wrench_world": [("gravity", {}), ("wheels", dict(torques=tq)), ("wrench_world", dict(wrench=wr))],
        "j2": [("j2", {})],
        "wheels + j2 (cube-sat shape)": [("wheels", dict(torques=tq)), ("j2", {})],
    }
    n = 4
    for name, spec in combos.items():
        oe, ge, cols = [], [], {}
        for kind, kw in spec:
            a, b, c = effector_pair(O, kind, **kw)
            oe.append(a); ge.append(b); cols.update(c)
        want = _run_oracle(O, pos, vel, ine, oe, 0.01, n, integrator)
        got = _run_gpu(pos, vel, ine, ge, cols, 0.01, n, "exact", integrator)
        if "j2" in name:
            _assert_close(got, want, 1e-14, f"exact {name}")
        else:
            _assert_exact(got, want, f"exact {name}")
        fast = _run_gpu(pos, vel, ine, ge, cols, 0.01, n, "fast", integrator)
        _assert_close(fast, want, n * FAST_TOL_TICK, f"fast {name}")


def test_cube_sat_ore_sat_golden_on_gpu(golden):
    """The reference's cube-sat golden, satellite entity: (1) the reaction-wheel fold on the recorded wheel commands,
    (2) one semi-implicit tick per recorded row with the recorded Force as a world-frame wrench.  Same bars as the
    oracle-side tests (tests/test_oracle_golden.py): the golden came from XLA-CPU, whose FMA contraction the IEEE-plain
    EXACT path does not imitate — <= 5e-16 (fold) / 2e-15 (tick) vector-relative, most rows bit-identical."""
    g = golden
    pos, vel, acc, frc = (g[f"cube_sat.ore_sat.{c}"] for c in ("world_pos", "world_vel", "world_accel", "force"))
    ine = g["cube_sat.ore_sat.inertia"]
    dt = float(g["cube_sat.simulation_time_step"][0, 0])
    T = len(pos)
    P, V, I = pos[:-1].reshape(T - 1, 1, 7), vel[:-1].reshape(T - 1, 1, 6), np.tile(ine[0], (T - 1, 1, 1))
    rw = np.concatenate([g[f"cube_sat.rw_{k}.rw_force"][:, :3] for k in (1, 2, 3)], -1)[1:].reshape(T - 1, 1, 9)
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.max(np.abs(b), axis=-1, keepdims=True), 1e-300)))
    # (1) every recorded row as one world of a batch: Force after one tick = the fold at the row's start pose
    got = _run_gpu(P, V, I, [el.TorqueBodyFold("wheel_torques", 3)], {"wheel_torques": rw}, dt, 1, "exact", "semi_implicit")
    assert np.all(got[3][:, 0, 3:] == 0.0)
    assert rel(got[3][:, 0, :3], frc[1:, :3]) <= 5e-16
    assert int(np.sum(np.all(got[3][:, 0, :3] == frc[1:, :3], axis=-1))) >= 50
    # (2) semi-implicit tick with the recorded wrench
    for math, tol in (("exact", 2e-15), ("fast", 1e-12)):
        p, v, a, f = _run_gpu(P, V, I, [el.WrenchWorld("external_force")], {"external_force": frc[1:].reshape(T - 1, 1, 6)}, dt, 1,
                              math, "semi_implicit")
        assert rel(a[:, 0], acc[1:]) <= tol and rel(v[:, 0], vel[1:]) <= tol
        assert rel(p[:, 0, :4], pos[1:, :4]) <= tol and rel(p[:, 0, 4:], pos[1:, 4:]) <= tol
        if math == "exact":
            assert np.array_equal(f[:, 0], frc[1:])
            assert int(np.sum(np.all(p[:, 0] == pos[1:], axis=-1))) >= 90


# --------------------------------------------------------------------------- round 2: the library's own NCCL paths (needs 2 GPUs)
def _two_rank_worker(rank, uid, q):
    import numpy as np

    import elodin_b200 as el
    from elodin_b200.executor import FORCE, WORLD_ACCEL, WORLD_POS, WORLD_VEL
    from elodin_b200.sharding import Comm, shard_sizes, shard_worlds
    from oracle import oracle as O
    from tests.util import random_world

    try:
        comm = Comm(uid, 2, rank, rank)
        # (1) world-sharded trajectory all-gather, ragged shards (3 + 2 worlds), host destination
        total = 5
        pos, vel, ine = random_world(5, total, 2)
        w0, w1 = shard_worlds(total, rank, 2)
        with el.B200Exec(2, w1 - w0, 0.01, None, [], "rk4", "exact", device=rank, trajectory_every=2, trajectory_capacity=3) as ex:
            ex.set_state(pos[w0:w1], vel[w0:w1], ine[w0:w1])
            ex.step(6, sync=True)
            full = comm.trajectory_allgather(ex, shard_sizes(total, 2))
        want = O.World(pos.copy(), vel.copy(), ine)
        rows = []
        for _ in range(3):
            want.rk4(0.01, 2)
            rows.append(np.concatenate([want.pos, want.vel], -1))
        ok_gather = bool(np.array_equal(full, np.stack(rows, 1)))  # [world][sample][entity][13], global world order
        # (2) one world, rows split over the two GPUs: EXACT folds are sequential per source, so the result is
        # bit-identical to the oracle's whole-world run
        N = 96
        p, v, I = random_world(9, 1, N)
        p[..., 4:] *= 1e-2
        edges = el.all_pairs_edges(N)
        o = O.World(p.copy(), v.copy(), I).rk4(0.01, 5, [O.Effector(O.EFF_GRAVITY_EDGES_SOFTENED, p=(0.3, 1e-4), edges=edges)])
        res = {}
        for math in ("exact", "fast"):
            with el.B200Exec(N, 1, 0.01, None, [el.GravityEdges("softened", k_squared=0.3, softening=1e-4, edges=edges)], "rk4", math,
                             device=rank) as ex:
                ex.set_state(p, v, I)
                comm.step_row_sharded(ex, 5)
                ex.sync()
                got = [ex.download(c) for c in (WORLD_POS, WORLD_VEL, WORLD_ACCEL, FORCE)]
                tick = ex.tick
            wants = (o.pos, o.vel, o.accel, o.force)
            if math == "exact":
                res[math] = all(np.array_equal(a, b) for a, b in zip(got, wants)) and tick == 5
            else:
                res[math] = max(float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) for a, b in zip(got, wants))
        comm.close()
        q.put((rank, ok_gather, res["exact"], res["fast"], None))
    except Exception as e:  # surface the failure in the parent
        import traceback

        q.put((rank, False, False, 1.0, traceback.format_exc()))


def test_library_nccl_gather_and_row_sharded_world_on_two_gpus():
    """b200_sixdof_trajectory_allgather (ragged world shards, global order) and b200_sixdof_step_row_sharded
    (one world, source rows over two GPUs: EXACT bit-identical to the oracle) — two processes, one GPU each."""
    if el.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from elodin_b200.sharding import Comm

    uid = Comm.unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_two_rank_worker, args=(r, uid, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok_gather, ok_exact, fast_err, err in got:
        assert err is None, err
        assert ok_gather, f"rank {rank}: gathered trajectory differs from the oracle"
        assert ok_exact, f"rank {rank}: row-sharded EXACT differs from the oracle"
        assert fast_err <= 5 * FAST_TOL_TICK * 10, fast_err


def _peer_window_worker(rank, uid, q):
    import numpy as np

    import elodin_b200 as el
    from elodin_b200 import _lib
    from elodin_b200.executor import FORCE, WORLD_ACCEL, WORLD_POS, WORLD_VEL
    from elodin_b200.sharding import Comm
    from oracle import oracle as O
    from tests.util import random_world

    try:
        comm = Comm(uid, 2, rank, rank)
        cols = (WORLD_POS, WORLD_VEL, WORLD_ACCEL, FORCE)
        res = {"unsupported": None}
        # N = 96: one shared-memory tile set (world kernel in FAST); N = 1280: the tiled fold kernels
        for N, ticks in ((96, (3, 2)), (1280, (1, 2))):
            p, v, I = random_world(9 + N, 1, N)
            p[..., 4:] *= 1e-2
            edges = el.all_pairs_edges(N)
            o = O.World(p.copy(), v.copy(), I).rk4(0.01, sum(ticks), [O.Effector(O.EFF_GRAVITY_EDGES_SOFTENED, p=(0.3, 1e-4), edges=edges)])
            for math in ("exact", "fast"):
                got = {}
                for route in ("nccl", "peer"):
                    with el.B200Exec(N, 1, 0.01, None, [el.GravityEdges("softened", k_squared=0.3, softening=1e-4, edges=edges)], "rk4",
                                     math, device=rank) as ex:
                        ex.set_state(p, v, I)
                        if route == "peer":
                            try:
                                comm.peer_attach(ex)
                            except _lib.B200Error as e:
                                if e.code != _lib.ERR_UNSUPPORTED:
                                    raise
                                res["unsupported"] = str(e)
                                break
                            assert comm.peer_attached
                        for n in ticks:  # two calls: the second one starts from the window state the first one left
                            comm.step_row_sharded(ex, n)
                        ex.sync()
                        got[route] = [ex.download(c) for c in cols]
                        assert ex.tick == sum(ticks)
                        if route == "peer":
                            comm.peer_detach()
                            assert not comm.peer_attached
                if res["unsupported"]:
                    break
                res[(N, math, "same")] = all(np.array_equal(a, b) for a, b in zip(got["nccl"], got["peer"]))
                if math == "exact":
                    res[(N, math, "oracle")] = all(np.array_equal(a, b) for a, b in zip(got["peer"], (o.pos, o.vel, o.accel, o.force)))
            if res["unsupported"]:
                break
        comm.close()
        q.put((rank, res, None))
    except Exception:
        import traceback

        q.put((rank, {}, traceback.format_exc()))


def test_row_sharded_world_through_peer_windows_on_two_gpus():
    """b200_comm_peer_attach: the row-sharded world exchanges its rows with NVLink stores into CUDA-IPC windows and
    counter releases instead of a collective per tick — bit-identical to the NCCL route in both math modes, and to the
    oracle in EXACT; two processes, one GPU each, two step calls per executor."""
    if el.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from elodin_b200.sharding import Comm

    uid = Comm.unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_window_worker, args=(r, uid, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, res, err in got:
        assert err is None, err
    if any(res["unsupported"] for _, res, _ in got):
        pytest.skip("CUDA IPC between the two processes is not available here: " + str(got[0][1]["unsupported"]))
    for rank, res, _ in got:
        checks = {k: v for k, v in res.items() if k != "unsupported"}
        assert len(checks) == 6 and all(checks.values()), (rank, checks)


def test_exact_shared_divisor_divisions():
    """EXACT mode divides groups of dividends by one divisor (quaternion / norm, force / mass, torque / inertia) with the
    divisor part of ptxas's div.rn.f64 expansion computed once (ex::rcp_prep / ex::div_rcp) and falls back to __ddiv_rn
    outside the expansion's own range test: 2^26 groups (2.7e8 divisions) over every encoding class — any encoding,
    ordinary magnitudes, the thresholds of the range test, overflow / denormal quotients, zeros, powers of two and
    all-ones significands, exact quotients, unit quaternions over their norm — must agree with the GPU's IEEE division
    in every bit, and the ordinary classes must not need the fallback."""
    import ctypes as C

    from elodin_b200 import _lib

    L = _lib.lib()
    out = (C.c_uint64 * 2)()
    n = 1 << 26
    for seed in (1, 0xB200):
        _lib.check(L.b200_selftest_shared_divisor(0, seed, n, out))
        assert out[0] == 0, f"{out[0]} of {4 * n} divisions differ from div.rn.f64"
        assert out[1] >= 0.35 * n  # classes 1, 7, 8, 9 (4 of 10) stay inside the window


def test_numa_local_pinned_buffers_and_pcie_probe():
    """b200_host_alloc_local: page-locked memory bound to the NUMA node of the GPU's PCIe root (falls back to plain
    pinned memory when the node is unknown), usable as invoke_batch column buffers; b200_probe_pcie_gbs reports both
    directions."""
    import ctypes as C

    from elodin_b200 import _lib

    L = _lib.lib()
    a = el.pinned_empty((1 << 16, 1, 7), np.float64, device=0)
    a[...] = 1.5
    node_gpu, node_buf = int(L.b200_device_numa_node(0)), int(L.b200_host_node_of(C.c_void_p(a.ctypes.data)))
    if node_gpu >= 0 and node_buf >= 0:
        assert node_buf == node_gpu
    pos, vel, ine = random_world(3, 1 << 16, 1)
    a[...] = pos
    with el.B200Exec(1, 1 << 16, 0.01, None, [], "rk4", "fast") as ex:
        ex.upload(WORLD_POS, a)
        ex.upload(WORLD_VEL, vel)
        ex.upload(INERTIA, ine)
        ex.step(2, sync=True)
        got = ex.download(WORLD_POS, out=a)
    assert np.isfinite(got).all() and not np.array_equal(got, pos)
    out = (C.c_double * 2)()
    scratch = el.pinned_empty(1 << 21, np.float64, device=0)  # 16 MB
    _lib.check(L.b200_probe_pcie_gbs(0, C.c_void_p(scratch.ctypes.data), 8 << 20, 8 << 20, 3, out))
    assert out[0] > 1.0 and out[1] > 1.0  # GB/s, both directions at once
    el.pinned_free(scratch)
    el.pinned_free(a)


@pytest.mark.parametrize("n_ticks", [1, 4])
@pytest.mark.parametrize("extra", [False, True])
@pytest.mark.parametrize("shape", [(41, 97), (1, 64), (2, 333), (3, 1024)])
def test_world_resident_pair_kernel_with_fused_integration(oracle, n_ticks, extra, shape):
    """Worlds of 64..1024 bodies run through the persistent pair kernel (graph_dense_world_kernel): gravity and
    integration in one launch per tick, ping-pong planes; odd / even tick counts, step() then chunked invoke_batch,
    gravity alone (compiled signature) and with another effector (interpreter), ragged N; a batch of many worlds (CTAs
    take whole worlds) and one to three worlds (every CTA a slice of a world's sources)."""
    O = oracle
    M, N = shape
    if N == 1024 and (extra or n_ticks > 1):
        pytest.skip("the 1024-body world is checked once (oracle cost)")
    pos, vel, ine = random_world(67, M, N)
    pos[..., 4:] *= 1e-2
    o, g, _ = effector_pair(O, "softened", edges=el.all_pairs_edges(N), k2=0.3, soft=1e-5)
    oe, ge, cols = [o], [g], {}
    if extra:
        thrust = np.random.default_rng(4).uniform(0, 3, (M, N, 1))
        o2, g2, c2 = effector_pair(O, "thrust", thrust=thrust)
        oe.append(o2); ge.append(g2); cols.update(c2)
    want = _run_oracle(O, pos, vel, ine, oe, 0.01, 2 * n_ticks)
    with el.B200Exec(N, M, 0.01, None, ge, "rk4", "fast", invoke_chunk_bodies=max(1, (M * 3) // 4) * N) as ex:
        ex.set_state(pos, vel, ine, **cols)
        ex.step(n_ticks, sync=True)
        mid = (ex.download(WORLD_POS), ex.download(WORLD_VEL))
        table = {el.component_id("tick"): np.array([n_ticks], dtype=np.uint64), FORCE: np.zeros((M, N, 6)), INERTIA: ine,
                 WORLD_POS: mid[0], WORLD_ACCEL: np.zeros((M, N, 6)), el.component_id("simulation_time_step"): np.array([0.01]),
                 WORLD_VEL: mid[1]}
        table.update({el.component_id(k): v for k, v in cols.items()})
        out = dict(zip(ex.output_ids, ex.invoke_batch([table[c] for c in ex.input_ids], n_ticks)))
        got = (out[WORLD_POS], out[WORLD_VEL], out[WORLD_ACCEL], out[FORCE])
        _assert_close(got, want, 1e-11, f"world kernel fused {n_ticks} extra={extra}")
        assert np.array_equal(ex.download(WORLD_POS), out[WORLD_POS]) and np.array_equal(ex.download(WORLD_VEL), out[WORLD_VEL])


@pytest.mark.parametrize("integrator", ["rk4", "semi_implicit"])
def test_egm08_gravity_effector(oracle, integrator):
    """GRAVITY_EGM08 (python/elodin/egm08.py): evaluated by egm08_force_kernel at the tick's three stage positions with
    the oracle's arithmetic — EXACT bit-identical to the oracle, FAST within tolerance; alone, with the wheel fold ahead
    of it (the cube-sat pipeline), masked to some entities, and through chunked invoke_batch."""
    from tests.test_oracle_golden import _egm08_random_tables

    O = oracle
    rng = np.random.default_rng(31)
    M, N, L = 4, 3, 12
    c, s = _egm08_random_tables(L, rng)
    pos, vel, ine = random_world(88, M, N)
    pos[..., 4:] = rng.normal(size=(M, N, 3))
    pos[..., 4:] *= 6.9e6 / np.linalg.norm(pos[..., 4:], axis=-1, keepdims=True)
    vel[..., 3:] = rng.normal(0, 7.6e3 / np.sqrt(3), (M, N, 3))
    tq = rng.normal(0, 2e-3, (M, N, 9))
    combos = {"egm08": [("egm08", dict(c_bar=c, s_bar=s, L=L))],
              "wheels + egm08 (cube-sat)": [("wheels", dict(torques=tq)), ("egm08", dict(c_bar=c, s_bar=s, L=L))]}
    n = 3
    for name, spec in combos.items():
        oe, ge, cols = [], [], {}
        for kind, kw in spec:
            a, b, cc = effector_pair(O, kind, **kw)
            oe.append(a); ge.append(b); cols.update(cc)
        want = _run_oracle(O, pos, vel, ine, oe, 0.05, n, integrator)
        got = _run_gpu(pos, vel, ine, ge, cols, 0.05, n, "exact", integrator)
        _assert_exact(got, want, f"exact {name}")
        fast = _run_gpu(pos, vel, ine, ge, cols, 0.05, n, "fast", integrator)
        _assert_close(fast, want, n * FAST_TOL_TICK, f"fast {name}")
    # the degrees at the edges of the term stream (one term; two-term columns) and the bench's degree 64, one tick each
    for Ld in (0, 1, 2, 64):
        cd, sd = np.tril(rng.normal(0, 1e-5, (Ld + 1, Ld + 1))), np.tril(rng.normal(0, 1e-5, (Ld + 1, Ld + 1)), -1)
        cd[0, 0] = 1.0
        a, b, cc = effector_pair(O, "egm08", c_bar=cd, s_bar=sd, L=Ld)
        want = _run_oracle(O, pos, vel, ine, [a], 0.05, 1, integrator)
        _assert_exact(_run_gpu(pos, vel, ine, [b], cc, 0.05, 1, "exact", integrator), want, f"exact degree {Ld}")
    # C20 alone == GRAVITY_J2 (the reference's own closed form).  max_degree 3: the source zeroes rho_{L+1}
    # (egm08.py:150), so the terms of the top degree L drop out — degree 2 needs L >= 3
    c2, s2 = np.zeros((4, 4)), np.zeros((4, 4))
    c2[0, 0], c2[2, 0] = 1.0, -1.08262668e-3 / np.sqrt(5.0)
    a = _run_gpu(pos, vel, ine, [el.GravityEGM08(c2, s2, 3)], {}, 0.05, 1, "exact", integrator)
    b = _run_gpu(pos, vel, ine, [el.GravityJ2()], {}, 0.05, 1, "exact", integrator)
    assert max_rel(a[3][..., 3:], b[3][..., 3:]) <= 1e-14
    # entity mask (only entity 1 feels the field) and chunked invoke_batch
    mask = np.array([0, 1, 0], dtype=np.uint8)
    oe = [O.Effector(O.EFF_GRAVITY_EGM08, p=(3.986004418e14, 6.378e6, L), tables=(c, s), mask=mask)]
    ge = [el.GravityEGM08(c, s, L).with_mask(mask)]
    want = _run_oracle(O, pos, vel, ine, oe, 0.05, 2, integrator)
    with el.B200Exec(N, M, 0.05, None, ge, integrator, "exact", invoke_chunk_bodies=2 * N) as ex:
        tick, dt = el.component_id("tick"), el.component_id("simulation_time_step")
        table = {tick: np.array([0], dtype=np.uint64), FORCE: np.zeros((M, N, 6)), INERTIA: ine, WORLD_POS: pos,
                 WORLD_ACCEL: np.zeros((M, N, 6)), dt: np.array([0.05]), WORLD_VEL: vel}
        o = dict(zip(ex.output_ids, ex.invoke_batch([table[k] for k in ex.input_ids], 2)))
    _assert_exact((o[WORLD_POS], o[WORLD_VEL], o[WORLD_ACCEL], o[FORCE]), want, "masked, chunked")
    with pytest.raises(el.B200Error):
        el.B200Exec(N, M, 0.05, None, [el.GravityEGM08(c, s, L), el.GravityEGM08(c, s, L)], integrator, "exact")
