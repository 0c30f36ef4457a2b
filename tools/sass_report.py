"""Opcode histograms (and a few telling lines) of named kernels in sqlp_b200/libsqlp_b200.so, from `cuobjdump -sass`.
usage: python tools/sass_report.py  >> profiles/r02_sass_contract.txt"""
import collections
import re
import subprocess
import sys

SO = "sqlp_b200/libsqlp_b200.so"
KERNELS = [("screening pass, end of round 2 (NX = 2, scalar adds)", "8k_screenILi2ELi0"),
           ("exact decision, rows staged in shared memory (NX = 2)", "15k_screen_decideILi2"),
           ("warm start (NX = 2)", "13k_screen_seedILi2"),
           ("per-vertex weight sums in fixed point (NX = 2)", "13k_cut_hist_fxILi2")]
SHOW = ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "LDGSTS", "DMMA", "FADD2", "FMNMX3", "ATOMS", "REDG", "RED.", "MATCH", "ACQBULK", "PREEXIT")

sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"(?m)^\s*Function : ", sass)[1:]
print("\n== end of round 2: kernels added or rewritten after the sections above (tools/sass_report.py) ==")
for title, key in KERNELS:
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if key not in name:
            continue
        ops = collections.Counter()
        lines = []
        for ln in f.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if not m:
                continue
            op = m.group(1)
            ops[op.split(".")[0]] += 1
            if any(op.startswith(s) for s in SHOW) and len(lines) < 400:
                lines.append(ln.split("/*", 2)[1].split("*/")[0] + "  " + ln.split("*/", 1)[1].split(";")[0].strip())
        print(f"\n-- {title}: {name} --")
        print("  " + ", ".join(f"{o} {c}" for o, c in ops.most_common(28)))
        seen = collections.Counter()
        for ln in lines:
            op = ln.split()[1].split(".")[0] if len(ln.split()) > 1 else ""
            seen[op] += 1
            if seen[op] <= 3:
                print("      " + ln)
        break
