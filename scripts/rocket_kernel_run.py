"""FAST body kernel with the free / rocket / falcon9 effector set at 2^22 worlds (ncu target):
    ncu --set full --clock-control none --import-source on -k regex:body_fast --launch-skip 6 --launch-count 1 \
        -o gpurun_out/r02_<set> python scripts/rocket_kernel_run.py <free|rocket|falcon9>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import elodin_b200 as el, bench
M = 1 << 22
pos, vel, ine = bench.synth_world(M, 1)
rng = np.random.default_rng(0)
which = sys.argv[1] if len(sys.argv) > 1 else "rocket"
if which == "free":
    effs, cols = [], {}
elif which == "rocket":
    effs = [el.GravityConst(), el.ThrustBody((-1.0, 0, 0), "thrust"), el.DragQuadratic(0.6, 0.01, "wind")]
    cols = {"thrust": rng.uniform(50, 100, (M, 1, 1)), "wind": rng.normal(0, 1, (M, 1, 3))}
elif which == "cube_sat":
    pos[..., 4:] += np.array([6.778e6, 0, 0])
    effs = [el.TorqueBodyFold("wheel_torques", 3), el.GravityJ2()]
    cols = {"wheel_torques": rng.normal(0, 2e-3, (M, 1, 9))}
else:
    pos[..., 4:] += np.array([6.4e6, 0, 0])
    effs = [el.GravityFrame(), el.WrenchBody("body_wrench", "linear_first")]
    cols = {"body_wrench": rng.normal(0, 1e3, (M, 1, 6))}
ex = el.B200Exec(1, M, 1e-3, None, effs, "rk4", "fast")
ex.set_state(pos, vel, ine, **cols)
ex.step(8, sync=True)
print("done")
