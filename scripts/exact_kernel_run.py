"""EXACT body kernel, free bodies, 2^20 worlds, RK4 (ncu target: -k regex:body_exact)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elodin_b200 as el, bench
M = 1 << 20
pos, vel, ine = bench.synth_world(M, 1)
ex = el.B200Exec(1, M, 1e-3, None, [], "rk4", "exact")
ex.set_state(pos, vel, ine)
ex.step(6, sync=True)
print("done")
