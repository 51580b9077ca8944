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
a = w.spawn([el.Body(world_pos=el.WorldPos(linear=np.array([0.8920281421, 0.0, 0.0])),
                     world_vel=el.WorldVel(linear=np.array([0.0, 0.9957939373, 0.0])),
                     inertia=el.Inertia(1.0 / G))], name="A")
b = w.spawn([el.Body(world_pos=el.WorldPos(linear=np.array([-0.6628498947, 0.0, 0.0])),
                     world_vel=el.WorldVel(linear=np.array([0.0, -1.6191613336, 0.0])),
                     inertia=el.Inertia(1.0 / G))], name="B")
c = w.spawn([el.Body(world_pos=el.WorldPos(linear=np.array([-0.2291782474, 0, 0])),
                     world_vel=el.WorldVel(linear=np.array([0, 0.6233673964, 0.0])),
                     inertia=el.Inertia(1.0 / G))], name="C")

GravityEdge = el.Annotated[el.Edge, el.Component("gravity_edge", el.ComponentType.Edge)]


@el.dataclass
class GravityConstraint(el.Archetype):
    a: GravityEdge

    def __init__(self, a: el.EntityId, b: el.EntityId):
        self.a = el.Edge(a, b)


w.spawn(GravityConstraint(a, b), name="A -> B")
w.spawn(GravityConstraint(b, a), name="B -> A")
w.spawn(GravityConstraint(a, c), name="A -> C")
w.spawn(GravityConstraint(b, c), name="B -> C")
w.spawn(GravityConstraint(c, a), name="C -> A")
w.spawn(GravityConstraint(c, b), name="C -> B")

gravity = el.GravityEdges("newton", G=G)  # edge_fold over the spawned GravityEdge components
sys_ = el.six_dof(sys=gravity)
ticks = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1000
sim = w.run(sys_, simulation_rate=1.0 / SIM_TIME_STEP, max_ticks=ticks)  # `python three_body.py bench --ticks N` also works
if sys.argv[1:2] == ["bench"]:
    raise SystemExit(0)
h = sim.history(["A.world_pos", "B.world_pos", "C.world_pos"])
for k, v in h.items():
    print(k, "after", ticks, "ticks:", v[-1][4:])
