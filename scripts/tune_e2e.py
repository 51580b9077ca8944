"""Sweep invoke_batch world-range size / fused ticks for the e2e path (2^20 worlds x 100 ticks per call)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import elodin_b200 as el, bench
from elodin_b200.executor import FORCE, INERTIA, WORLD_ACCEL, WORLD_POS, WORLD_VEL
eM, T = 1 << 20, 100
pos, vel, ine = bench.synth_world(eM, 2000)
host = {WORLD_POS: pos, WORLD_VEL: vel, INERTIA: ine, WORLD_ACCEL: np.zeros((eM, 1, 6)), FORCE: np.zeros((eM, 1, 6)),
        el.component_id("tick"): np.zeros(1, dtype=np.uint64), el.component_id("simulation_time_step"): np.array([1e-3])}
for chunk in (16384, 32768, 65536, 131072, 262144):
    for fuse in (25, 100):
        ee = el.B200Exec(1, eM, 1e-3, None, [], "rk4", "fast", max_fused_ticks=fuse, invoke_chunk_bodies=chunk)
        pin_in, pin_out = [], []
        for cid in ee.input_ids:
            a = el.pinned_empty(host[cid].shape, host[cid].dtype, device=0); a[...] = host[cid]; pin_in.append(a)
        for cid in ee.output_ids:
            pin_out.append(el.pinned_empty(host[cid].shape, host[cid].dtype, device=0))
        keep = (WORLD_POS, WORLD_VEL, el.component_id("tick"))  # state outputs only (bench.py's headline e2e point)
        ip, op = [a.ctypes.data for a in pin_in], [a.ctypes.data if c in keep else None for c, a in zip(ee.output_ids, pin_out)]
        ee.invoke_batch_ptrs(ip, op, T)
        t0 = time.perf_counter()
        for _ in range(5): ee.invoke_batch_ptrs(ip, op, T)
        ms = (time.perf_counter() - t0) / 5 * 1e3
        print(f"chunk {chunk:8d} fuse {fuse:3d}: {ms:6.2f} ms/call  {eM*T/ms/1e3:.3e} entity-steps/s")
        ee.close()
        for a in pin_in + pin_out: el.pinned_free(a)
