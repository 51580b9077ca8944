"""Round-1 sweep (profiles/r01_tuning.md).  Since round 2 the default FAST route is the signature-specialised kernel, swept by scripts/tune_spec.py in a tuning build; B200_BODY_CFG now only selects the opt-in TMA-ring variants 10..14 (any other value = default).
Sweep B200_BODY_CFG launch variants of the FAST RK4 body kernel (one process per variant)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for cfg in sys.argv[1:] or ["0", "1", "2", "3", "4", "5", "10", "11", "12"]:
    env = dict(os.environ, B200_BODY_CFG=cfg)
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--kernel-only", "--steps", "100", "--warmup", "5"],
                         env=env, capture_output=True, text=True)
    print("cfg", cfg, out.stdout.strip()[-200:], out.stderr.strip()[-300:])
    chk = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests/test_parity_gpu.py"), "-x", "-q", "-k",
                          "effector_combos or fused or full_size or trajectory"], env=env, capture_output=True, text=True, cwd=root)
    print("   parity:", chk.stdout.strip().splitlines()[-1] if chk.stdout.strip() else chk.stderr[-300:])
