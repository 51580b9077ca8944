"""Launch-shape sweep of the specialised FAST kernels (tuning build: make -C elodin_b200/csrc TUNE=1).
One process; B200_SPEC_CFG is re-read per launch.  Prints us/tick, achieved GB/s (algorithmic bytes incl. effector
columns) and the fraction of the measured HBM copy peak, and checks every shape against the interpreter kernel."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import elodin_b200 as el, bench
from elodin_b200 import _lib
from elodin_b200.executor import WORLD_POS, WORLD_VEL

M = int(os.environ.get("TUNE_WORLDS", 1 << 22))
TICKS = int(os.environ.get("TUNE_TICKS", 60))
cfgs = [int(c) for c in (sys.argv[1:] or ["-1", "1", "2", "3", "8", "4", "5", "6", "7"])]
peak, _ = bench.measured_peak()
L = _lib.lib()
print("copy GB/s", L.b200_probe_copy_gbs(0, 1 << 30, 5), "fp64 GFLOP/s", L.b200_probe_fp64_gflops(0, 20000), "peak used", peak, flush=True)
pos, vel, ine = bench.synth_world(M, 1)
rng = np.random.default_rng(0)
sets = {
    "free": ([], {}, 264),
    "rocket": ([el.GravityConst(), el.ThrustBody((-1.0, 0, 0), "thrust"), el.DragQuadratic(0.6, 0.01, "wind")],
               {"thrust": rng.uniform(50, 100, (M, 1, 1)), "wind": rng.normal(0, 1, (M, 1, 3))}, 264 + 32),
    "falcon9": ([el.GravityFrame(), el.WrenchBody("body_wrench", "linear_first")], {"body_wrench": rng.normal(0, 1e3, (M, 1, 6))}, 264 + 48),
}
st = torch.cuda.Stream()
rows = []
for name, (effs, cols, nbytes) in sets.items():
    p = pos.copy()
    if name == "falcon9":
        p[..., 4:] += np.array([6.4e6, 0, 0])
    with torch.cuda.stream(st):
        ex = el.B200Exec(1, M, 1e-3, None, effs, "rk4", "fast")
        ex.set_stream(st.cuda_stream)
        ref = None
        for cfg in ["generic"] + cfgs:
            os.environ["B200_NO_SPEC"] = "1" if cfg == "generic" else "0"
            os.environ["B200_SPEC_CFG"] = str(cfg if cfg != "generic" else -1)
            ex.set_state(p, vel, ine, **cols)
            ex.step(10)
            torch.cuda.synchronize()
            got = np.concatenate([ex.download(WORLD_POS)[:4096].ravel(), ex.download(WORLD_VEL)[:4096].ravel()])
            if ref is None:
                ref = got
            err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300))) if np.isfinite(got).all() else float("nan")
            best = 1e30
            for _ in range(3):
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record(st); ex.step(TICKS); b.record(st); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) / TICKS)
            gbs = nbytes * M / (best * 1e-3) / 1e9
            row = {"set": name, "cfg": cfg, "us_per_tick": best * 1e3, "GBps": gbs, "frac": gbs / peak, "max_rel_vs_generic": err}
            rows.append(row)
            print(json.dumps(row), flush=True)
        ex.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/tune_spec.json", "w"), indent=1)
