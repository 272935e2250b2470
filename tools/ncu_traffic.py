#!/usr/bin/env python
"""Measure the DRAM bytes one solve launch actually moves (dram__bytes_read.sum + dram__bytes_write.sum, ncu) and record
it in profiles/ncu_traffic.json, which bench.py reads for `roofline.traffic` (labelled with the capture's commit and batch).
Run on the GPU box:  python tools/ncu_traffic.py [--batch 16384] [--model quat] [--horizon 10] [--kernel auto]"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--model", default="quat")
    ap.add_argument("--horizon", type=int, default=10)
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--commit", default=os.environ.get("QMPC_COMMIT", "unknown"))
    a = ap.parse_args()
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct",
           "--clock-control", "none", "-k", "regex:qmpc_(coop|phased)", "-c", "64", "--csv", sys.executable, os.path.join(ROOT, "bench.py"),
           "--steps", "1", "--warmup", "1", "--batch", str(a.batch), "--model", a.model, "--horizon", str(a.horizon), "--kernel", a.kernel,
           "--no-cpu-baseline", "--no-aux", "--no-config1"]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout
    rows = [r for r in csv.reader(io.StringIO(out)) if len(r) > 14 and r[0].isdigit()]
    # launches: warm-up step, timed step, e2e warm-ups ...: group by launch id, take the launches of ONE step (the second)
    by_id = {}
    for r in rows:
        by_id.setdefault(int(r[0]), {"name": r[4]})[r[12]] = float(r[14].replace(",", ""))
    ids = sorted(by_id)
    names = [by_id[i]["name"].split("(")[0] for i in ids]
    per_step = len(ids) // 4 if a.kernel == "phased" else 1          # bench: 1 warm-up + 1 timed + 2 e2e warm-ups + 1 e2e
    if a.kernel == "phased":
        per_step = 1 + 2 * (5 if a.model == "convex" else 10)
    step = ids[per_step:2 * per_step]
    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    # ncu prints bytes with a unit column (r[13]); re-read units
    units = {(int(r[0]), r[12]): r[13] for r in rows}
    tot = 0.0
    for i in step:
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += by_id[i][m] * unit.get(units[(i, m)], 1.0)
    rec = {"model": a.model, "horizon": a.horizon, "batch": a.batch, "kernel": "kernel=phased" if a.kernel == "phased" else "kernel=coop",
           "dram_bytes_per_solve": tot / a.batch, "launches": len(step), "kernels": sorted(set(names)),
           "l2_hit_pct": by_id[step[0]].get("lts__t_sector_hit_rate.pct"), "commit": a.commit,
           "capture": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on bench.py --batch {a.batch} --model {a.model} --horizon {a.horizon} --kernel {a.kernel}"}
    path = os.path.join(ROOT, "gpurun_out", "ncu_traffic.json")
    recs = json.load(open(path)) if os.path.exists(path) else []
    recs = [r for r in recs if not (r["model"] == rec["model"] and r["horizon"] == rec["horizon"] and r["kernel"] == rec["kernel"])] + [rec]
    json.dump(recs, open(path, "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
