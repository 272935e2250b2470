#!/usr/bin/env python
"""Top source lines of an ncu report by stall samples (source page, cuda+sass): file:line samples% inst% text."""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
S = defaultdict(lambda: [0, 0, 0]); hdr = None; cur = None; files = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name": continue
    if r[0] != "" and r[0].isdigit(): cur = (cur_file, int(r[0])); continue
    if r[0] == "" and len(r) > 5 and r[2].startswith("0x"):
        d = dict(zip(hdr[4:], r[4:]))
        S[cur][0] += int(d["# Samples"]); S[cur][1] += int(d["Instructions Executed"]); S[cur][2] += 1
tot = sum(v[0] for v in S.values()); toti = sum(v[1] for v in S.values())
for (f, l), v in sorted(S.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in files:
        try: files[f] = open(f.replace("/root/repo", "/root/repo")).read().split("\n")
        except Exception: files[f] = []
    text = files[f][l - 1].strip()[:90] if 0 < l <= len(files[f]) else ""
    print(f"{f.split('/')[-1]}:{l:5d} {100*v[0]/tot:5.2f}% t {100*v[1]/toti:5.2f}% i {v[2]:4d} sass | {text}")
