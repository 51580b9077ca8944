"""egm08_force_kernel timing: wheel fold + degree-L EGM08 (the cube-sat effector shape), M worlds, RK4, FAST.
    python scripts/egm08_perf.py [L=64] [log2 M=16]      ncu target: -k regex:egm08_force"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import elodin_b200 as el, bench
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 16)
pos, vel, ine = bench.synth_world(M, 1)
rng = np.random.default_rng(8)
cb, sb = np.zeros((L + 1, L + 1)), np.zeros((L + 1, L + 1))
for l in range(2, L + 1):
    cb[l, : l + 1] = rng.normal(0, 1e-5 / l**2, l + 1); sb[l, 1: l + 1] = rng.normal(0, 1e-5 / l**2, l)
cb[0, 0], cb[2, 0] = 1.0, -1.08262668e-3 / np.sqrt(5.0)
pos[..., 4:] += np.array([6.778e6, 0.0, 0.0])
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    ex = el.B200Exec(1, M, 1e-3, None, [el.TorqueBodyFold("wheel_torques", 3), el.GravityEGM08(cb, sb, L)], "rk4", "fast")
    ex.set_stream(st.cuda_stream); ex.set_state(pos, vel, ine, wheel_torques=rng.normal(0, 2e-3, (M, 1, 9)))
    ex.step(2); torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record(st); ex.step(10); b.record(st); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 100
terms = (L + 1) * (L + 2) // 2
print(f"L={L} M={M}: {us:.1f} us/tick  {M/us*1e6:.3e} entity-steps/s  {3*M/us*1e6:.3e} field evals/s  {3*M*terms/us*1e6:.3e} terms/s")
ex.close()
