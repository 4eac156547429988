#!/usr/bin/env python3
"""Regenerate profiles/inst_table.json: per-ply warp-instruction counts of bench.py's headline workload, for the CURRENT
build, from an ncu capture of the very command bench.py runs (seeded => the same plies, the same work).

    python tools/make_inst_table.py [--plies 26]          # on the GPU box (gpurun); ~10-15 min under ncu

Runs `ncu --profile-from-start off --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,
gpu__time_duration.sum python bench.py --steps <plies> --warmup 0 --ncu-range
...`, splits the launch list into plies at every qz_mcts_choose_kernel launch (one per ply) and sums per kernel.
bench.py divides these counts by its own CUDA-event time for the same plies: `roofline.frac`.
"""
import argparse
import csv
import datetime
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

# one ncu pass: more metrics (e.g. dram__bytes_*) force kernel replay, and replay saves / restores every writable
# allocation around each of the ~15,000 launches (the first attempt ran 27 minutes and died)
METRICS = ["smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "gpu__time_duration.sum"]


def short(name):
    m = re.match(r"(?:void\s+)?([A-Za-z_0-9]+)", name)
    return m.group(1) if m else name


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


def parse_csv(path):
    """ncu --csv (one row per kernel x metric) -> ordered list of dict(name, metric values)."""
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    cur = None
    for r in rd:
        kid = r.get("ID")
        if cur is None or cur["id"] != kid:
            cur = {"id": kid, "name": short(r.get("Kernel Name", "")), "m": {}}
            rows.append(cur)
        val = to_float(r.get("Metric Value", "0"))
        unit = (r.get("Metric Unit") or "").lower()
        name = r.get("Metric Name")
        if name == "gpu__time_duration.sum":
            val *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3,
                    "second": 1e3}.get(unit, 1e-6)
        if name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            val *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        cur["m"][name] = val
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plies", default="0,3,5,8,11,14,17,20,23,25",
                    help="self-play plies to capture (ncu costs ~0.15 s per launch, ~400 launches per ply); bench.py "
                         "interpolates the plies in between")
    ap.add_argument("--csv", default=os.path.join(ROOT, "gpurun_out", "inst_table_launches.csv"))
    ap.add_argument("--parse-only", action="store_true")
    a = ap.parse_args()
    args = bench.parse([])
    plies = sorted(int(x) for x in a.plies.split(",") if x.strip())
    os.makedirs(os.path.dirname(a.csv), exist_ok=True)
    if not a.parse_only:
        # -k regex:qz_ : only the library's own kernels (a torch elementwise kernel once failed to profile and took the
        # whole capture down; torch's share of the step is < 0.1 % of the instructions)
        cmd = ["ncu", "--profile-from-start", "off", "--clock-control", "none", "-k", "regex:qz_", "--metrics", ",".join(METRICS), "--csv",
               "--log-file", a.csv, sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(max(plies) + 1), "--warmup", "0",
               "--ncu-range", "--ncu-plies", a.plies, "--no-kernels", "--no-az", "--no-parity", "--no-cpu-baseline",
               "--games-plies", "0"]
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    rows = parse_csv(a.csv)
    per_ply, cur = [], None
    for r in rows:
        if cur is None:
            cur = {"warp_inst": 0.0, "thread_inst": 0.0, "time_ms": 0.0, "dram_bytes": 0.0, "launches": 0, "kernels": {}}
        m = r["m"]
        wi, ti = m.get("smsp__inst_executed.sum", 0.0), m.get("smsp__thread_inst_executed.sum", 0.0)
        ms = m.get("gpu__time_duration.sum", 0.0)
        db = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        cur["warp_inst"] += wi
        cur["thread_inst"] += ti
        cur["time_ms"] += ms
        cur["dram_bytes"] += db
        cur["launches"] += 1
        k = cur["kernels"].setdefault(r["name"], {"launches": 0, "warp_inst": 0.0, "thread_inst": 0.0, "time_ms": 0.0,
                                                  "dram_bytes": 0.0})
        k["launches"] += 1
        k["warp_inst"] += wi
        k["thread_inst"] += ti
        k["time_ms"] += ms
        k["dram_bytes"] += db
        if r["name"] == "qz_mcts_choose_kernel":
            per_ply.append(cur)
            cur = None
    if cur is not None and cur["launches"] > 8:                # launches after the last move (restart bookkeeping)
        for k, v in cur["kernels"].items():
            t = per_ply[-1]["kernels"].setdefault(k, {"launches": 0, "warp_inst": 0.0, "thread_inst": 0.0, "time_ms": 0.0,
                                                      "dram_bytes": 0.0})
            for f in v:
                t[f] += v[f]
        for f in ("warp_inst", "thread_inst", "time_ms", "dram_bytes", "launches"):
            per_ply[-1][f] += cur[f]
    if len(per_ply) != len(plies):
        print("WARNING: captured %d plies, expected %d (%s): keeping the first ones" % (len(per_ply), len(plies), plies))
    per_ply = {str(p): row for p, row in zip(plies, per_ply)}
    table = {"when": datetime.datetime.utcnow().isoformat() + "Z", "build_hash": bench.build_hash(),
             "args": {k: getattr(args, k) for k in ("games", "playouts", "leaves", "seed", "defer")},
             "metrics": METRICS, "how": "ncu --profile-from-start off --clock-control none, kernel replay; time_ms is "
             "serialised and cold-cache (compare shares, not absolutes)", "plies": list(per_ply), "per_ply": per_ply}
    out = os.path.join(ROOT, "profiles", "inst_table.json")
    with open(out, "w") as f:
        json.dump(table, f, indent=0)
    rows_all = list(per_ply.values())
    tot = sum(p["warp_inst"] for p in rows_all)
    print("wrote %s: plies %s, %d launches, %.2f G warp instructions per ply (mean), %.1f lanes/instruction"
          % (out, list(per_ply), sum(p["launches"] for p in rows_all), tot / max(len(rows_all), 1) / 1e9,
             sum(p["thread_inst"] for p in rows_all) / max(tot, 1)))
    agg = {}
    for p in rows_all:
        for k, v in p["kernels"].items():
            g = agg.setdefault(k, {"launches": 0, "warp_inst": 0.0, "thread_inst": 0.0, "time_ms": 0.0})
            for f in g:
                g[f] += v[f]
    tms = sum(v["time_ms"] for v in agg.values()) or 1.0
    with open(os.path.join(ROOT, "profiles", "inst_table_summary.txt"), "w") as f:
        f.write("# %s  build %s  plies %s of `python bench.py` (ncu, serialised)\n" % (table["when"], table["build_hash"], list(per_ply)))
        f.write("%-36s %8s %14s %8s %8s %8s\n" % ("kernel", "launches", "warp_inst", "inst %", "time %", "lanes"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["warp_inst"]):
            f.write("%-36s %8d %14.4g %8.2f %8.2f %8.1f\n" % (k, v["launches"], v["warp_inst"], 100 * v["warp_inst"] / max(tot, 1),
                                                            100 * v["time_ms"] / tms, v["thread_inst"] / max(v["warp_inst"], 1)))
    print(open(os.path.join(ROOT, "profiles", "inst_table_summary.txt")).read())


if __name__ == "__main__":
    main()
