"""Generate Monte-Carlo plan goldens by running the REFERENCE's own sampler.

    python tests/golden/make_mc_golden.py        (build container only: needs /root/reference)

Imports /root/reference/libs/nox-py/python/elodin/monte_carlo/sample.py (pure stdlib; the
`elodin` package itself cannot be imported here, so the module is loaded by path) and
materialises every spec below, writing tests/golden/mc_plans.json = {name: {"spec": toml text,
"plan_csv": the reference's plan.csv text}}.
"""

import importlib.util
import json
import os
import tempfile
from pathlib import Path

REF = os.environ.get("ELODIN_REFERENCE", "/root/reference")
HERE = Path(__file__).resolve().parent

EXTRA_SPECS = {
    "mixed_dists": """
[monte_carlo]
n_samples = 17
seed = 7
method = "lhs"
[monte_carlo.variables]
a = { dist = "normal", mean = 1.5, std = 0.25 }
b = { dist = "loguniform", lo = 1e-3, hi = 10.0 }
c = { dist = "choice", values = [1, 2, 3, 5, 8] }
d = { dist = "fixed", value = 3.25 }
e = { dist = "uniform", low = -1.0, high = 1.0 }
[sim_sweep]
gain = [0.5, 1.0]
[meta_sweep]
label = ["x", "y", "z"]
""",
    "random_method": """
[monte_carlo]
n_samples = 9
seed = 123456
method = "random"
[monte_carlo.variables]
thrust = { dist = "uniform", min = 80.0, max = 100.0 }
cd = { dist = "normal", mean = 0.5, std = 0.05 }
""",
    "sweep_only": """
[sim_sweep]
mass = [1.0, 2.0, 3.0]
k = [10, 20]
""",
}


def main():
    spec = importlib.util.spec_from_file_location(
        "ref_sample", os.path.join(REF, "libs/nox-py/python/elodin/monte_carlo/sample.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    specs = {
        "falcon9_spec": Path(REF, "examples/falcon9/spec.toml").read_text(),
        "monte_carlo_example": Path(REF, "examples/monte-carlo/spec.toml").read_text(),
        **EXTRA_SPECS,
    }
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name, text in specs.items():
            sp, pl = Path(td, f"{name}.toml"), Path(td, f"{name}.csv")
            sp.write_text(text)
            ref.materialize(sp, pl)
            out[name] = {"spec": text, "plan_csv": pl.read_text()}
    (HERE / "mc_plans.json").write_text(json.dumps(out, indent=0))
    print("wrote", HERE / "mc_plans.json", {k: v["plan_csv"].count("\n") - 1 for k, v in out.items()})


if __name__ == "__main__":
    main()
