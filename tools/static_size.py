#!/usr/bin/env python
"""Static SASS size per source line range of the coop<4> kernel (+ its out-of-line callees) from `nvdisasm -g -c`."""
import re, sys
from collections import defaultdict
sass = sys.argv[1]
src = open("/root/repo/quaternion_mpc_b200/csrc/qmpc_coop.cuh").read().split("\n")
infunc = None; cur = None
cnt = defaultdict(int); per_func = defaultdict(int)
for line in open(sass, errors="replace"):
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m:
        infunc = m.group(1); continue
    if line.startswith("\t.section") or line.startswith(".section"):
        continue
    if infunc is None: continue
    keep = ("qmpc_coop_kernelINS_9QuatModelILi4" in infunc) or ("coop_kernelILi4" in infunc) or ("ILi4E" in infunc and ("rollout" in infunc or "cost_expand" in infunc)) or "gemm3" in infunc
    if not keep: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        cnt[cur] += 1; per_func[infunc[:60]] += 1
print({k: v for k, v in per_func.items()})
# group coop.cuh lines into 25-line buckets labelled by the nearest preceding '// ----' comment
def label(f, l):
    if f != "qmpc_coop.cuh": return f
    lab = "head"
    for i in range(l - 1, -1, -1):
        t = src[i].strip()
        if t.startswith("// ----") or t.startswith("// ----------------") or "QMPC_NOINLINE void" in t or t.startswith("QMPC_HD inline"):
            lab = t[:70]; break
    return lab
agg = defaultdict(int)
for (k, v) in cnt.items():
    if k is None: agg["?"] += v; continue
    agg[label(*k)] += v
tot = sum(agg.values())
print("total", tot, "instrs", tot * 16 / 1024, "KB")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{v:6d} {v*16/1024:6.1f} KB  {k}")
