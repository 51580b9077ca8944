"""Per-range timeline of one invoke_batch call (B200_PIPE_TRACE=1): 2^20 worlds x 100 ticks, state outputs, pinned columns."""
import sys, os, time
os.environ["B200_PIPE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import elodin_b200 as el, bench
from elodin_b200.executor import FORCE, INERTIA, WORLD_ACCEL, WORLD_POS, WORLD_VEL
eM, T = 1 << 20, 100
pos, vel, ine = bench.synth_world(eM, 2000)
host = {WORLD_POS: pos, WORLD_VEL: vel, INERTIA: ine, WORLD_ACCEL: np.zeros((eM, 1, 6)), FORCE: np.zeros((eM, 1, 6)),
        el.component_id("tick"): np.zeros(1, dtype=np.uint64), el.component_id("simulation_time_step"): np.array([1e-3])}
ee = el.B200Exec(1, eM, 1e-3, None, [], "rk4", "fast", max_fused_ticks=25)
pin_in, pin_out = [], []
for cid in ee.input_ids:
    a = el.pinned_empty(host[cid].shape, host[cid].dtype, device=0); a[...] = host[cid]; pin_in.append(a)
for cid in ee.output_ids:
    pin_out.append(el.pinned_empty(host[cid].shape, host[cid].dtype, device=0))
keep = (WORLD_POS, WORLD_VEL, el.component_id("tick"))
ip, op = [a.ctypes.data for a in pin_in], [a.ctypes.data if c in keep else None for c, a in zip(ee.output_ids, pin_out)]
for i in range(3):
    t0 = time.perf_counter(); ee.invoke_batch_ptrs(ip, op, T); print(f"call {i}: {(time.perf_counter()-t0)*1e3:.3f} ms", file=sys.stderr)
