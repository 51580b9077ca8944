"""FAST vs EXACT drift over long runs (rocket-style worlds with live attitude dynamics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import elodin_b200 as el
from elodin_b200.executor import WORLD_POS, WORLD_VEL
from tests.util import random_world
M = 4096
pos, vel, ine = random_world(42, M, 1)
rng = np.random.default_rng(42)
effs = lambda: [el.GravityConst(), el.ThrustBody((-1.0, 0, 0), "thrust"), el.DragQuadratic(0.6, 0.01, "wind"), el.WrenchBody("aero_force")]
cols = {"thrust": rng.uniform(50, 100, (M, 1, 1)), "wind": rng.normal(0, 1, (M, 1, 3)), "aero_force": rng.normal(0, 0.05, (M, 1, 6))}
def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
ex = el.B200Exec(1, M, 0.008333333, None, effs(), "rk4", "exact", max_fused_ticks=100); ex.set_state(pos, vel, ine, **cols)
fx = el.B200Exec(1, M, 0.008333333, None, effs(), "rk4", "fast", max_fused_ticks=100); fx.set_state(pos, vel, ine, **cols)
done = 0
for n in (1, 9, 90, 900, 9000):
    ex.step(n, sync=True); fx.step(n, sync=True); done += n
    pe, pf, ve, vf = ex.download(WORLD_POS), fx.download(WORLD_POS), ex.download(WORLD_VEL), fx.download(WORLD_VEL)
    print(f"ticks {done:6d}: q {rel(pf[..., :4], pe[..., :4]):.2e}  x {rel(pf[..., 4:], pe[..., 4:]):.2e}  omega {rel(vf[..., :3], ve[..., :3]):.2e}  v {rel(vf[..., 3:], ve[..., 3:]):.2e}")
