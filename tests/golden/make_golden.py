"""Repack the reference's own CI golden telemetry into compact fixtures.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

Reads  /root/reference/scripts/ci/baseline/{three-body,rocket,ball,cube-sat}-csv/*.csv
       (101 rows each, written by the reference's `bench --ticks 100` +
        `elodin-db export --format csv --flatten`, scripts/ci/regress.sh)
Writes tests/golden/elodin_ci_baseline.npz   (f64 arrays, bit-exact copies of
       the CSV values; the `time` column is dropped, as compare_baseline_csv.py does).

Nothing under tests/ reads /root/reference at run time: only this script does.
"""

import csv
import os
import sys

import numpy as np

REF = os.environ.get("ELODIN_REFERENCE", "/root/reference")
BASE = os.path.join(REF, "scripts", "ci", "baseline")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "elodin_ci_baseline.npz")


def read(sim: str, stem: str) -> np.ndarray:
    path = os.path.join(BASE, f"{sim}-csv", f"{stem}.csv")
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    data = np.array([[float(x) for x in r[1:]] for r in rows[1:]], dtype=np.float64)
    return data


def main() -> int:
    out = {}
    for ent in "abc":
        for comp in ("world_pos", "world_vel", "world_accel", "force", "inertia"):
            out[f"three_body.{ent}.{comp}"] = read("three-body", f"{ent}.{comp}")
    edges = []
    for name in ("a_to_b", "b_to_a", "a_to_c", "b_to_c", "c_to_a", "c_to_b"):  # spawn order, main.py:82-89
        edges.append(read("three-body", f"{name}.gravity_edge")[0])
    out["three_body.edges_entity_ids"] = np.array(edges, dtype=np.float64)
    out["three_body.simulation_time_step"] = read("three-body", "globals.simulation_time_step")
    out["three_body.tick"] = read("three-body", "globals.tick")
    for comp in ("world_pos", "world_vel", "world_accel", "force", "inertia", "thrust", "aero_force"):
        out[f"rocket.{comp}"] = read("rocket", f"rocket.{comp}")
    out["rocket.simulation_time_step"] = read("rocket", "globals.simulation_time_step")
    for comp in ("world_pos", "world_vel", "world_accel", "force", "inertia", "wind"):
        out[f"ball.{comp}"] = read("ball", f"ball.{comp}")
    out["ball.simulation_time_step"] = read("ball", "globals.simulation_time_step")
    # cube-sat runs Integrator.SemiImplicit (examples/cube-sat/main.py:710); its `earth` entity is a free
    # spinning body (zero force), which pins semi_implicit.rs:42-62 without needing the EGM08 / reaction-wheel models
    for comp in ("world_pos", "world_vel", "world_accel", "force", "inertia"):
        out[f"cube_sat.earth.{comp}"] = read("cube-sat", f"earth.{comp}")
    out["cube_sat.simulation_time_step"] = read("cube-sat", "globals.simulation_time_step")
    # the satellite itself: Force = reaction-wheel edge fold (main.py:492-505) + EGM08 gravity (main.py:516-527), integrated
    # by Integrator.SemiImplicit.  The recorded wheel commands pin the fold, the recorded Force pins semi-implicit with a
    # full wrench (the EGM08 coefficient tables are a download the reference tree does not hold)
    for comp in ("world_pos", "world_vel", "world_accel", "force", "inertia"):
        out[f"cube_sat.ore_sat.{comp}"] = read("cube-sat", f"ore_sat.{comp}")
    for k in (1, 2, 3):  # edge spawn order sat_to_rw_1..3
        out[f"cube_sat.rw_{k}.rw_force"] = read("cube-sat", f"rw_{k}.rw_force")
    # layout of the exported directory (file stems + header rows), for the CSV-export parity test
    layout = {}
    d = os.path.join(BASE, "three-body-csv")
    for fn in sorted(os.listdir(d)):
        if fn.endswith(".csv"):
            with open(os.path.join(d, fn), newline="") as f:
                layout[fn] = next(csv.reader(f))
            if "gravity_edge" in fn:
                out[f"three_body.file.{fn}"] = read("three-body", fn[:-4])
    import json
    with open(os.path.join(os.path.dirname(OUT), "three_body_csv_layout.json"), "w") as f:
        json.dump(layout, f, indent=0, ensure_ascii=False)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT)} bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
