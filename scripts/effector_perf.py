"""Device timing of the FAST body kernel with the rocket / falcon9 effector sets (2^22 worlds)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import elodin_b200 as el, bench
M = 1 << 22
pos, vel, ine = bench.synth_world(M, 1)
rng = np.random.default_rng(0)
sets = {
 "free": ([], {}),
 "rocket(g+thrust+drag)": ([el.GravityConst(), el.ThrustBody((-1.0, 0, 0), "thrust"), el.DragQuadratic(0.6, 0.01, "wind")],
            {"thrust": rng.uniform(50, 100, (M, 1, 1)), "wind": rng.normal(0, 1, (M, 1, 3))}),
 "falcon9(frame+wrench)": ([el.GravityFrame(), el.WrenchBody("body_wrench", "linear_first")], {"body_wrench": rng.normal(0, 1e3, (M, 1, 6))}),
}
st = torch.cuda.Stream()
for name, (effs, cols) in sets.items():
    p = pos.copy()
    if name.startswith("falcon9"): p[..., 4:] += np.array([6.4e6, 0, 0])
    with torch.cuda.stream(st):
        ex = el.B200Exec(1, M, 1e-3, None, effs, "rk4", "fast"); ex.set_stream(st.cuda_stream); ex.set_state(p, vel, ine, **cols)
        ex.step(5); torch.cuda.synchronize()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(st); ex.step(100); b.record(st); torch.cuda.synchronize()
        us = a.elapsed_time(b) * 10
        bytes_ = 264 + 8 * sum({"thrust": 1, "wind": 3, "body_wrench": 6}[c] for c in cols)
        print(f"{name:24s} {us:7.1f} us/tick  {M/us*1e6:.3e} entity-steps/s  {bytes_*M/us/1e3:7.0f} GB/s ({bytes_} B/entity-step incl. effector columns)")
        ex.close()
