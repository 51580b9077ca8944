"""Ad-hoc device timing of the body kernels (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import elodin_b200 as el
from elodin_b200 import _lib
from tests.util import random_world

L = _lib.lib()
print("copy GB/s", L.b200_probe_copy_gbs(0, 1 << 30, 5), "fp64 GFLOP/s", L.b200_probe_fp64_gflops(0, 20000))
torch.cuda.init()
st = torch.cuda.current_stream()
def bench(M, math, effs, cols, fused, ticks, reps=5):
    pos, vel, ine = random_world(1, M, 1)
    with el.B200Exec(1, M, 1e-3, None, effs, "rk4", math, max_fused_ticks=fused) as ex:
        ex.set_stream(st.cuda_stream)
        ex.set_state(pos, vel, ine, **cols)
        ex.step(ticks); torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); ex.step(ticks); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        es = M * ticks / (best * 1e-3)
        print(f"M={M:>8} math={math:5} neff={len(effs)} fused={fused:3} ticks={ticks:4}: {best/ticks*1e3:9.2f} us/tick  {es:.3e} entity-steps/s  {es*264/1e9:8.1f} GB/s-alg")
rng = np.random.default_rng(0)
for M in (1 << 20, 1 << 22):
    rocket = ([el.GravityConst(), el.ThrustBody((-1., 0, 0), "thrust"), el.DragQuadratic(0.6, 0.01, "wind")],
              {"thrust": rng.uniform(50, 100, (M, 1, 1)), "wind": rng.normal(0, 1, (M, 1, 3))})
    for math in ("fast", "exact"):
        bench(M, math, [], {}, 1, 20)
        bench(M, math, [], {}, 20, 20)
        bench(M, math, *rocket, 1, 20)
bench(1, "fast", [], {}, 1000, 100000, reps=2)
bench(1, "exact", [], {}, 1000, 100000, reps=2)
bench(10000, "fast", *([el.GravityConst(), el.ThrustBody((-1., 0, 0), "thrust"), el.DragQuadratic(0.6, 0.01, "wind")],
      {"thrust": rng.uniform(50, 100, (10000, 1, 1)), "wind": rng.normal(0, 1, (10000, 1, 3))}), 100, 5000, reps=2)
