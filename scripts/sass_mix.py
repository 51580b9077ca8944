"""Instruction mix of the kernels in the built .so whose demangled name contains every filter word:
python scripts/sass_mix.py <word> [<word> ...]"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "elodin_b200", "libb200_sixdof.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, mixes = None, {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void b200::", "")
        mixes[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        mixes[cur][m.group(1).split(".")[0] + ("." + m.group(1).split(".")[2] if m.group(1).startswith(("LDG", "STG")) and len(m.group(1).split(".")) > 2 else "")] += 1
for name, c in mixes.items():
    if all(w in name for w in sys.argv[1:]):
        tot = sum(c.values())
        fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
        print(f"{name}: {tot} instr, fp64-pipe {fp64}")
        print("   " + " ".join(f"{k}:{v}" for k, v in c.most_common(24)))
