"""n-body 1024 (BASELINE configs[3]) mini driver for profiling: M worlds, T ticks, FAST or EXACT."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import elodin_b200 as el
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
math = sys.argv[3] if len(sys.argv) > 3 else "fast"
N = 1024
rng = np.random.default_rng(7)
p = np.zeros((M, N, 7)); p[..., 3] = 1.0; p[..., 4:] = rng.uniform(-30, 30, (M, N, 3))
v = np.zeros((M, N, 6)); v[..., 3:] = rng.normal(0, 1e-7, (M, N, 3))
m = 10 ** rng.uniform(-10, -3, (M, N)); m[:, 0] = 1.0
I = np.zeros((M, N, 7)); I[..., :3] = m[..., None]; I[..., 6] = m
g = el.GravityEdges("softened", k_squared=2.9591220828e-4 / 86400.0 ** 2, softening=1e-10, edges=el.all_pairs_edges(N))
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    ex = el.B200Exec(N, M, 3600.0, None, [g], "rk4", math)
    ex.set_stream(st.cuda_stream); ex.set_state(p, v, I)
    ex.step(5); torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record(st); ex.step(T); b.record(st); torch.cuda.synchronize()
    print(f"nbody N={N} M={M} {math}: {a.elapsed_time(b)*1e3/T:.2f} us/tick")
