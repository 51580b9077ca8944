"""Shared helpers for the parity tests: seeded synthetic worlds and
the oracle <-> C-ABI effector translation."""

import numpy as np

import elodin_b200 as el


def random_world(seed, M, N, unit_q=True):
    """q ~ normalised N(0,1)^4, x ~ U(-1e3,1e3), omega ~ N(0,.5), v ~ N(0,10),
    inertia diag ~ U(.1,10), m ~ U(.5,50)  (SURVEY §8d synthetic inputs)."""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(M, N, 4))
    if unit_q:
        q /= np.linalg.norm(q, axis=-1, keepdims=True)
    pos = np.concatenate([q, rng.uniform(-1e3, 1e3, (M, N, 3))], -1)
    vel = np.concatenate([rng.normal(0, 0.5, (M, N, 3)), rng.normal(0, 10, (M, N, 3))], -1)
    ine = np.concatenate([rng.uniform(0.1, 10, (M, N, 3)), np.zeros((M, N, 3)), rng.uniform(0.5, 50, (M, N, 1))], -1)
    return np.ascontiguousarray(pos), np.ascontiguousarray(vel), np.ascontiguousarray(ine)


def effector_pair(O, kind, **kw):
    """(oracle effector, elodin_b200 effector, {column name: array}) for one built-in kind."""
    if kind == "gravity":
        g = kw.get("g", (0.0, 0.0, -9.81))
        return O.Effector(O.EFF_GRAVITY_CONST, p=g), el.GravityConst(g), {}
    if kind == "drag":
        wind = kw["wind"]
        cd, area = kw.get("cd_rho", 0.6125), kw.get("area", 0.25)
        return (O.Effector(O.EFF_DRAG_QUADRATIC, p=(cd, area), column=wind), el.DragQuadratic(cd, area, "wind"),
                {"wind": wind})
    if kind == "thrust":
        thrust = kw["thrust"]
        axis = kw.get("axis", (-1.0, 0.0, 0.0))
        return (O.Effector(O.EFF_THRUST_BODY, p=axis, column=thrust), el.ThrustBody(axis, "thrust"),
                {"thrust": thrust})
    if kind == "wrench":
        wr = kw["wrench"]
        lin_first = kw.get("linear_first", False)
        return (O.Effector(O.EFF_WRENCH_BODY, flags=O.FLAG_WRENCH_LINEAR_FIRST if lin_first else 0, column=wr),
                el.WrenchBody("aero_force", "linear_first" if lin_first else "torque_first"), {"aero_force": wr})
    if kind == "frame":
        mu, om = kw.get("mu", 3.986004418e14), kw.get("omega", (0.0, 0.0, 7.292115e-5))
        return O.Effector(O.EFF_GRAVITY_FRAME, p=(mu, *om)), el.GravityFrame(mu, om), {}
    if kind == "wrench_world":
        wr = kw["wrench"]
        return O.Effector(O.EFF_WRENCH_WORLD, column=wr), el.WrenchWorld("external_force"), {"external_force": wr}
    if kind == "wheels":
        tq = kw["torques"]
        return (O.Effector(O.EFF_TORQUE_BODY_FOLD, column=tq), el.TorqueBodyFold("wheel_torques", tq.shape[-1] // 3),
                {"wheel_torques": tq})
    if kind == "j2":
        mu, j2, rr = kw.get("mu", 3.986004418e14), kw.get("j2", 1.08262668e-3), kw.get("r_ref", 6.378e6)
        return O.Effector(O.EFF_GRAVITY_J2, p=(mu, j2, rr)), el.GravityJ2(mu, j2, rr), {}
    if kind == "egm08":
        c, s, L = kw["c_bar"], kw["s_bar"], kw["L"]
        mu, rr = kw.get("mu", 3.986004418e14), kw.get("r_ref", 6.378e6)
        return O.Effector(O.EFF_GRAVITY_EGM08, p=(mu, rr, L), tables=(c, s)), el.GravityEGM08(c, s, L, mu, rr), {}
    if kind == "newton":
        return (O.Effector(O.EFF_GRAVITY_EDGES_NEWTON, p=(kw.get("G", 6.6743e-11),), edges=kw["edges"]),
                el.GravityEdges("newton", G=kw.get("G", 6.6743e-11), edges=kw["edges"]), {})
    if kind == "softened":
        k2, soft = kw.get("k2", 1e-3), kw.get("soft", 1e-10)
        return (O.Effector(O.EFF_GRAVITY_EDGES_SOFTENED, p=(k2, soft), edges=kw["edges"]),
                el.GravityEdges("softened", k_squared=k2, softening=soft, edges=kw["edges"]), {})
    raise KeyError(kind)


def max_rel(a, b):
    """max |a-b| / max(|b|) per trailing vector block — the vector-scaled relative error."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.max(np.abs(b), axis=-1, keepdims=True), 1e-300)
    return float(np.max(np.abs(a - b) / scale))
