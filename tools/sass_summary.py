#!/usr/bin/env python
"""Summarise the SASS of the shipped library (cuobjdump -sass): per kernel the static instruction count, code size and
the mnemonic histogram - which memory / tensor / TMA instructions the product really contains.
usage: sass_summary.py [lib.so] > profiles/rNN_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "quaternion_mpc_b200", "libqmpc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).startswith(("LDGSTS", "CCTL", "UBLKCP", "SYNCS", "BAR")) and "." in m.group(1) else "")] += 1
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print(f"# SASS summary of {os.path.basename(lib)} (cuobjdump -sass; arch {', '.join(arch)})\n")
print("Static instruction counts per kernel; 16 bytes per instruction. DFMA/DMUL/DADD = FP64 vector pipe; LDS/STS = shared memory;\n"
      "LDGSTS = cp.async (global -> shared, asynchronous); CCTL.E.RML2 = discard.global.L2; UBLKCP = cp.async.bulk (TMA 1-D);\n"
      "DMMA / HMMA / UTCMMA = tensor pipes (none: see DESIGN.md section 4 for the measured decision).\n")
print("| kernel | instructions | KB | DFMA | DMUL | DADD | LDS | STS | LDG/LD | STG/ST | LDL | STL | LDGSTS | CCTL.RML2 | UBLKCP | DMMA | MUFU | BAR |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
    h = hist[k]
    n = sum(h.values())
    g = lambda *names: sum(v for kk, v in h.items() if kk.split(".")[0] in names)
    print(f"| `{k}` | {n} | {n * 16 / 1024:.1f} | {g('DFMA')} | {g('DMUL')} | {g('DADD')} | {g('LDS')} | {g('STS')} | {g('LDG', 'LD')} | {g('STG', 'ST')} | "
          f"{g('LDL')} | {g('STL')} | {g('LDGSTS')} | {sum(v for kk, v in h.items() if kk.startswith('CCTL'))} | {g('UBLKCP')} | {g('DMMA')} | {g('MUFU')} | {g('BAR')} |")
