"""A rocket-style Monte-Carlo campaign as ONE executor (BASELINE configs[2] shape).

    python examples/rocket_monte_carlo.py [n_samples] [ticks]

The reference would spawn one Python process per plan row (elodin monte-carlo run); here the plan
becomes the world axis: spec -> plan (bit-identical sampler) -> per-world parameter columns ->
one B200 executor -> per-run result.json.
"""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import elodin_b200 as el
from elodin_b200 import monte_carlo as mc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
spec = {"monte_carlo": {"n_samples": n, "seed": 42, "method": "lhs", "variables": {
    "thrust_gain": {"dist": "uniform", "min": 0.8, "max": 1.2},
    "mass": {"dist": "uniform", "min": 2.5, "max": 3.5},
    "wind_x": {"dist": "normal", "mean": 0.0, "std": 2.0}}}}
rows = mc.materialize(spec)

Thrust = el.Annotated[np.ndarray, el.Component("thrust", el.ComponentType.F64)]
Wind = el.Annotated[np.ndarray, el.Component("wind", el.ComponentType(el.PrimitiveType.F64, (3,)))]


@el.dataclass
class Rocket(el.Archetype):
    thrust: Thrust
    wind: Wind


w = el.World()
w.spawn([el.Body(world_pos=el.SpatialTransform(angular=el.Quaternion.from_euler([0.0, np.radians(70.0), 0.0]),
                                               linear=np.array([0.0, 0.0, 1.0])),
                 inertia=el.SpatialInertia(3.0, np.array([0.1, 1.0, 1.0]))),
         Rocket(np.array([88.426]), np.zeros(3))], name="rocket")
effectors = el.GravityConst((0.0, 0.0, -9.81)) | el.ThrustBody((-1.0, 0.0, 0.0), "thrust") | el.DragQuadratic(0.6125, 0.0025, "wind")
params = mc.world_params(rows, 1, {
    "thrust": lambda p: [88.426 * p["thrust_gain"]],
    "wind": lambda p: [p["wind_x"], 0.0, 0.0],
    "inertia": lambda p: [0.1, 1.0, 1.0, 0, 0, 0, p["mass"]],
})
exec = w.build(el.six_dof(sys=effectors), simulation_rate=120.0, telemetry_rate=120.0 / ticks, math="fast",
               n_worlds=n, world_params=params)
exec.run(ticks)
pos = exec.history_worlds("rocket.world_pos")[-1]  # [n_worlds, 7]
print(f"{n} worlds x {ticks} ticks; downrange x: mean {pos[:, 4].mean():.2f} m, std {pos[:, 4].std():.2f} m; kernel stats {exec.profile()}")
out = mc.write_results(rows[:5], tempfile.mkdtemp(), {"world_pos": exec.world.columns[el.component_id('world_pos')].buffer[:5]})
print("result.json for the first runs:", out[0])
