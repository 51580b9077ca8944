"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes binding of oracle/libsixdof_oracle.so (the plain-C restatement of the
reference arithmetic, see sixdof_oracle.h).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsixdof_oracle.so")

EFF_GRAVITY_CONST = 1
EFF_DRAG_QUADRATIC = 2
EFF_THRUST_BODY = 3
EFF_WRENCH_BODY = 4
EFF_GRAVITY_FRAME = 5
EFF_GRAVITY_EDGES_NEWTON = 6
EFF_GRAVITY_EDGES_SOFTENED = 7
EFF_WRENCH_WORLD = 8
EFF_TORQUE_BODY_FOLD = 9
EFF_GRAVITY_J2 = 10
EFF_GRAVITY_EGM08 = 11
FLAG_WRENCH_LINEAR_FIRST = 1


class _Effector(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32),
        ("flags", C.c_uint32),
        ("p", C.c_double * 8),
        ("column", C.c_void_p),
        ("column_width", C.c_uint32),
        ("reserved", C.c_uint32),
        ("n_edges", C.c_uint64),
        ("edge_from", C.c_void_p),
        ("edge_to", C.c_void_p),
        ("entity_mask", C.c_void_p),
        ("table0", C.c_void_p),
        ("table1", C.c_void_p),
        ("table_len", C.c_uint64),
    ]


class _World(C.Structure):
    _fields_ = [
        ("n", C.c_uint64),
        ("n_worlds", C.c_uint64),
        ("pos", C.c_void_p),
        ("vel", C.c_void_p),
        ("accel", C.c_void_p),
        ("force", C.c_void_p),
        ("inertia", C.c_void_p),
    ]


def build() -> str:
    """Compile the oracle in place (gcc, no FMA contraction)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        L.orc_qmul.argtypes = [dp, dp, dp]
        L.orc_qinv.argtypes = [dp, dp]
        L.orc_qrot.argtypes = [dp, dp, dp]
        L.orc_qnormalize.argtypes = [dp, dp]
        L.orc_transform_add_motion.argtypes = [dp, dp, dp]
        L.orc_calc_accel.argtypes = [dp, dp, dp, dp]
        L.orc_rk4_ticks.argtypes = [C.POINTER(_World), C.c_uint32, C.POINTER(_Effector), C.c_double,
                                    C.c_double, C.c_uint64, C.c_int]
        L.orc_semi_implicit_ticks.argtypes = [C.POINTER(_World), C.c_uint32, C.POINTER(_Effector),
                                              C.c_double, C.c_uint64, C.c_int]
        L.orc_eval_stage.argtypes = [C.POINTER(_World), C.c_uint64, C.c_uint32, C.POINTER(_Effector),
                                     dp, dp, dp, dp]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _vec(fn, n_out, *ins):
    ins = [np.ascontiguousarray(x, dtype=np.float64) for x in ins]
    out = np.zeros(n_out)
    fn(*[_dp(x) for x in ins], _dp(out))
    return out


def qmul(l, r):
    return _vec(lib().orc_qmul, 4, l, r)


def qinv(q):
    return _vec(lib().orc_qinv, 4, q)


def qrot(q, v):
    return _vec(lib().orc_qrot, 3, q, v)


def qnormalize(q):
    return _vec(lib().orc_qnormalize, 4, q)


def transform_add_motion(pos, motion):
    return _vec(lib().orc_transform_add_motion, 7, pos, motion)


def calc_accel(pos, force, inertia):
    return _vec(lib().orc_calc_accel, 6, pos, force, inertia)


@dataclass
class Effector:
    kind: int
    p: tuple = ()
    flags: int = 0
    column: np.ndarray | None = None  # [M, N, width]
    edges: np.ndarray | None = None  # [E, 2] (from, to) entity rows
    mask: np.ndarray | None = None  # [N] bool: entity rows the effector applies to (query join)
    tables: tuple | None = None  # GRAVITY_EGM08: (c_bar, s_bar), each [(L+1), (L+1)]

    _keep: list = field(default_factory=list, repr=False)

    def to_c(self) -> _Effector:
        e = _Effector()
        e.kind = self.kind
        e.flags = self.flags
        for i, v in enumerate(self.p):
            e.p[i] = float(v)
        if self.column is not None:
            col = np.ascontiguousarray(self.column, dtype=np.float64)
            self._keep.append(col)
            e.column = col.ctypes.data
            e.column_width = col.shape[-1] if col.ndim > 1 else 1
        if self.edges is not None:
            ed = np.asarray(self.edges, dtype=np.uint32).reshape(-1, 2)
            f = np.ascontiguousarray(ed[:, 0])
            t = np.ascontiguousarray(ed[:, 1])
            self._keep += [f, t]
            e.n_edges = len(f)
            e.edge_from = f.ctypes.data
            e.edge_to = t.ctypes.data
        if self.mask is not None:
            m = np.ascontiguousarray(np.asarray(self.mask, dtype=np.uint8))
            self._keep.append(m)
            e.entity_mask = m.ctypes.data
        if self.tables is not None:
            c, sbar = (np.ascontiguousarray(t, dtype=np.float64) for t in self.tables)
            self._keep += [c, sbar]
            e.table0, e.table1, e.table_len = c.ctypes.data, sbar.ctypes.data, c.size
        return e


class World:
    """AoS columns [M, N, width] (f64), the host layout of the C ABI."""

    def __init__(self, pos, vel, inertia, accel=None, force=None):
        self.pos = np.array(pos, dtype=np.float64, order="C")
        if self.pos.ndim == 2:
            self.pos = self.pos[None]
        M, N, _ = self.pos.shape
        self.M, self.N = M, N
        self.vel = np.array(vel, dtype=np.float64, order="C").reshape(M, N, 6)
        self.inertia = np.array(inertia, dtype=np.float64, order="C").reshape(M, N, 7)
        self.accel = (np.zeros((M, N, 6)) if accel is None
                      else np.array(accel, dtype=np.float64, order="C").reshape(M, N, 6))
        self.force = (np.zeros((M, N, 6)) if force is None
                      else np.array(force, dtype=np.float64, order="C").reshape(M, N, 6))

    def _c(self) -> _World:
        w = _World()
        w.n, w.n_worlds = self.N, self.M
        w.pos, w.vel = self.pos.ctypes.data, self.vel.ctypes.data
        w.accel, w.force = self.accel.ctypes.data, self.force.ctypes.data
        w.inertia = self.inertia.ctypes.data
        return w

    def _effs(self, effectors):
        effectors = list(effectors or [])
        arr = (_Effector * max(len(effectors), 1))()
        for i, e in enumerate(effectors):
            arr[i] = e.to_c()
        return arr, len(effectors)

    def rk4(self, dt, n_ticks=1, effectors=None, dt_final=None, threads=1):
        arr, n = self._effs(effectors)
        w = self._c()
        lib().orc_rk4_ticks(C.byref(w), n, arr, float(dt), float(dt if dt_final is None else dt_final),
                            int(n_ticks), int(threads))
        return self

    def semi_implicit(self, dt, n_ticks=1, effectors=None, threads=1):
        arr, n = self._effs(effectors)
        w = self._c()
        lib().orc_semi_implicit_ticks(C.byref(w), n, arr, float(dt), int(n_ticks), int(threads))
        return self

    def eval_stage(self, world=0, effectors=None):
        """force, accel of the effector pipe + calc_accel evaluated on the current state."""
        arr, n = self._effs(effectors)
        w = self._c()
        F = np.zeros((self.N, 6))
        A = np.zeros((self.N, 6))
        lib().orc_eval_stage(C.byref(w), int(world), n, arr, _dp(self.pos[world]), _dp(self.vel[world]),
                             _dp(F), _dp(A))
        return F, A


def set_dot_mode(mode: int) -> None:
    """0 = plain IEEE (canonical); 1 = golden-host FMA-contracted dot (see sixdof_oracle.c)."""
    lib().orc_set_dot_mode(int(mode))


def max_threads() -> int:
    return int(lib().orc_max_threads())
