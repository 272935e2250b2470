#!/usr/bin/env python
"""Summarise an ncu report of the coop kernel: headline metrics + per-source-region samples / instruction mix.
usage: ncu_regions.py report.ncu-rep   (regions are found from marker comments in qmpc_coop.cuh)"""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum"]
for h, v in zip(hdr, vals):
    if h in want or re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", h):
        try:
            if float(v.replace(",", "")) < 0.03: continue
        except Exception: pass
        print(f"{h:95s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# region markers from the source file itself
lines = open("/root/repo/quaternion_mpc_b200/csrc/qmpc_coop.cuh").read().split("\n")
marks = []
pats = [("blk helpers", "// ---- 3x3 block kernels"), ("layout", "struct CoopRow"), ("linearize/At/Mt", "// ---- model dispatch: linearisation"),
        ("knot_merit/hphi", "// stage cost + AL terms of one knot"), ("bulk/cp.async helpers", "// 1-D bulk copy global -> shared"),
        ("rollout", "// One roll-out of the whole horizon"), ("ctx", "// Per-problem context of the phase functions"),
        ("setup", "// ------------------------------------------------------------------ set-up + nominal roll-out"),
        ("expansions", "// ---------------- expansions, lane k <- knot k"), ("stationarity", "// ---------------- stationarity"),
        ("dual update", "// dual update (row-parallel)"), ("AL terms", "// ---------------- AL terms of every (knot, foot)"),
        ("bp init", "// ------------------------------------------------------------------ Riccati backward pass"),
        ("phase B", "// ---- phase B"), ("phase C", "// ---- phase C"), ("phase D", "// ---- phase D"), ("phase E", "// ---- phase E"),
        ("chol+solves", "// ---- Cholesky + both triangular solves"), ("phase F", "// ---- phase F"),
        ("linesearch", "// ------------------------------------------------------------------ forward pass"),
        ("accept", "// ---------------- accepted step"), ("epilogue", "// ------------------------------------------------------------------ result + warm-start buffer"),
        ("fused loop", "// ------------------------------------------------------------------ the whole solve, fused")]
for name, pat in pats:
    p0 = pat.split("\n")[0]
    for i, l in enumerate(lines):
        if p0 in l:
            marks.append((i + 1, name)); break
marks.sort()
def region(f, l):
    if f != "qmpc_coop.cuh": return f
    name = "head"
    for ln, n in marks:
        if l >= ln: name = n
    return name
hdr = None; cur_file = None; cur_line = None
S = defaultdict(lambda: defaultdict(float)); seen = set()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name": continue
    if r[0] != "" and r[0].isdigit(): cur_line = int(r[0]); continue
    if r[0] == "" and len(r) > 5 and r[2].startswith("0x"):
        d = dict(zip(hdr[4:], r[4:]))
        reg = region(cur_file, cur_line)
        op = r[3].strip().split()
        op = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
        n = int(d["Instructions Executed"])
        S[reg]["samples"] += int(d["# Samples"]); S[reg]["inst"] += n
        for k in ("stall_wait", "stall_short_sb", "stall_long_sb", "stall_no_inst", "stall_branch_resolving", "stall_mio", "stall_lg", "stall_math"):
            S[reg][k] += int(d[k])
        if op in ("LDS",): S[reg]["LDS"] += n
        elif op in ("STS",): S[reg]["STS"] += n
        elif op in ("LD", "LDG", "LDL"): S[reg]["LDg"] += n
        elif op in ("ST", "STG", "STL"): S[reg]["STg"] += n
        elif op in ("DFMA", "DMUL", "DADD"): S[reg]["FP64"] += n
        if r[2] not in seen: seen.add(r[2]); S[reg]["static"] += 1
tot = sum(v["samples"] for v in S.values()); toti = sum(v["inst"] for v in S.values())
print(f"\ntotal samples {tot:.0f}  instructions {toti/1e6:.0f}M  static {sum(v['static'] for v in S.values()):.0f} instrs")
print(f"{'region':20s} {'time%':>6s} {'inst%':>6s} {'instM':>7s} {'FP64':>6s} {'LDS':>6s} {'STS':>5s} {'LDg':>5s} {'STg':>5s} {'KB':>5s} | wait shortsb longsb noinst branch mio")
for k, v in sorted(S.items(), key=lambda kv: -kv[1]["samples"]):
    print(f"{k:20s} {100*v['samples']/tot:6.1f} {100*v['inst']/toti:6.1f} {v['inst']/1e6:7.1f} {v['FP64']/1e6:6.1f} {v['LDS']/1e6:6.1f} {v['STS']/1e6:5.1f} {v['LDg']/1e6:5.1f} {v['STg']/1e6:5.1f} {v['static']*16/1024:5.1f} | "
          f"{100*v['stall_wait']/tot:4.1f} {100*v['stall_short_sb']/tot:4.1f} {100*v['stall_long_sb']/tot:4.1f} {100*v['stall_no_inst']/tot:4.1f} {100*v['stall_branch_resolving']/tot:4.1f} {100*v['stall_mio']/tot:4.1f}")
