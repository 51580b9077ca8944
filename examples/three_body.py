"""examples/three-body/main.py of the reference, on the B200 backend.

    python examples/three_body.py [ticks]
    python examples/three_body.py bench --ticks 1000 [--detail]      # the reference's bench CLI

Same script shape as the reference (spawn three bodies, six gravity edges, el.six_dof(sys=gravity),
run); the only changes are the import and that `gravity` is the built-in edge_fold effector
instead of a traced JAX function.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import elodin_b200 as el

SIM_TIME_STEP = 1.0 / 120.0
G = 6.6743e-11

w = el.World()
# figure-eight-like initial conditions of the reference example (positions on the x axis, velocities along y)
INITIAL = {"A": (0.8920281421, 0.9957939373), "B": (-0.6628498947, -1.6191613336), "C": (-0.2291782474, 0.6233673964)}
body = {name: w.spawn([el.Body(world_pos=el.WorldPos(linear=np.array([x, 0.0, 0.0])),
                               world_vel=el.WorldVel(linear=np.array([0.0, vy, 0.0])),
                               inertia=el.Inertia(1.0 / G))], name=name)
        for name, (x, vy) in INITIAL.items()}
a, b, c = body["A"], body["B"], body["C"]

GravityEdge = el.Annotated[el.Edge, el.Component("gravity_edge", el.ComponentType.Edge)]


@el.dataclass
class GravityConstraint(el.Archetype):
    a: GravityEdge

    def __init__(self, a: el.EntityId, b: el.EntityId):
        self.a = el.Edge(a, b)


# directed gravity edges in the reference's spawn order (it fixes the edge_fold order per source body)
for src, dst in (("A", "B"), ("B", "A"), ("A", "C"), ("B", "C"), ("C", "A"), ("C", "B")):
    w.spawn(GravityConstraint(body[src], body[dst]), name=f"{src} -> {dst}")

gravity = el.GravityEdges("newton", G=G)  # edge_fold over the spawned GravityEdge components
sys_ = el.six_dof(sys=gravity)
ticks = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1000
sim = w.run(sys_, simulation_rate=1.0 / SIM_TIME_STEP, max_ticks=ticks)  # `python three_body.py bench --ticks N` also works
if sys.argv[1:2] == ["bench"]:
    raise SystemExit(0)
h = sim.history(["A.world_pos", "B.world_pos", "C.world_pos"])
for k, v in h.items():
    print(k, "after", ticks, "ticks:", v[-1][4:])
