"""Free body with a (WorldPos, WorldVel) sample written to the device trajectory ring on every tick: us per tick at 2^22 bodies."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import elodin_b200 as el, bench
M = 1 << 22
pos, vel, ine = bench.synth_world(M, 1)
st = torch.cuda.Stream()
for every, planes in ((1, 13), (1, 25), (4, 13)):
    with torch.cuda.stream(st):
        ex = el.B200Exec(1, M, 1e-3, None, [], "rk4", "fast", trajectory_every=every, trajectory_capacity=32 // every, trajectory_full=(planes == 25))
        ex.set_stream(st.cuda_stream); ex.set_state(pos, vel, ine)
        ex.step(8); torch.cuda.synchronize()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(st); ex.step(24); b.record(st); torch.cuda.synchronize()  # the ring holds 32 ticks' samples: 8 warm + 24 timed
        us = a.elapsed_time(b) * 1e3 / 24
        by = 264 + 8 * planes / every
        print(f"sample every {every} tick(s), {planes} planes: {us:7.1f} us/tick  {by*M/us/1e3:7.0f} GB/s algorithmic ({by:.0f} B/entity-step)")
        ex.close()
