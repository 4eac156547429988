#!/usr/bin/env python3
"""Per-source-line instruction counts of one profiled kernel.

  python tools/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTR [FILE_TO_ATTRIBUTE_TO [inner]]

Joins `ncu --page source --csv` (SASS view: executed instructions per SASS instruction) with
`nvdisasm -gi` of the cubin embedded in libqzb200.so (file:line + inlined-at chain per instruction), and prints
the share of warp-level instructions per source line.  With FILE given, an instruction is attributed to the
outermost frame of its inline chain that lies in FILE (so qz_warp.cuh shows which of ITS lines the time
belongs to, helpers included); otherwise to the innermost frame.
"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "alphazero_quoridor_b200", "libqzb200.so")],
                   cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
        m = re.search(r"^(_Z\w*%s\w*):$" % re.escape(kernel_substr), dis, re.M)
        if not m:
            continue
        body = dis[m.end():]
        table, frames = [], []
        pending = []
        for ln in body.split("\n"):
            f = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if f:
                pending.append((os.path.basename(f.group(1)), int(f.group(2))))
                continue
            i = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if i:
                if pending:
                    frames = pending
                pending = []
                table.append((int(i.group(1), 16), i.group(2).strip(), list(frames)))
                continue
            if re.match(r"\s*\.section|^//-+ \.text", ln) and table:
                break
        return table
    raise SystemExit("kernel not found in any cubin")


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    attr_file = sys.argv[3] if len(sys.argv) > 3 else None
    inner = len(sys.argv) > 4 and sys.argv[4] == "inner"       # innermost frame inside FILE instead of outermost
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # the report may hold several kernels: take the first section whose kernel name matches
    start = next(i for i, r in enumerate(rows) if len(r) > 1 and r[0] == "Kernel Name" and kern in r[1])
    hdr = next(i for i in range(start, len(rows)) if rows[i] and rows[i][0] == "Address")
    h = rows[hdr]
    ia, ii, it = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    isamp = h.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in h else None
    insts = []
    for r in rows[hdr + 1:]:
        if len(r) != len(h) or r[0] in ("Address", "Kernel Name"):
            break
        insts.append((int(r[ia], 16), int(r[ii]), int(r[it]), int(r[isamp]) if isamp is not None and r[isamp].isdigit() else 0))
    base = insts[0][0]
    table = {off: (txt, fr) for off, txt, fr in line_table(kern)}
    agg, tot, ttot, stot = {}, 0, 0, 0
    for addr, n, tn, sm in insts:
        txt, fr = table.get(addr - base, ("?", []))
        key = ("?", 0)
        if fr:
            key = fr[0]
            if attr_file:
                for f in fr:                      # innermost first; keep the outermost frame inside attr_file
                    if f[0] == attr_file:
                        key = f
                        if inner:
                            break
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += n
        a[1] += tn
        a[2] += sm
        tot += n
        ttot += tn
        stot += sm
    print("# %s: %d warp instructions, %.2f threads/instruction" % (kern, tot, ttot / max(tot, 1)))
    src_cache = {}
    by = 2 if os.environ.get("SORT_BY_STALL") else 0      # SORT_BY_STALL=1: rank by warp-stall samples (latency-bound kernels)
    for key, (n, tn, sm) in sorted(agg.items(), key=lambda kv: -kv[1][by])[:40]:
        f, l = key
        if f not in src_cache:
            p = glob.glob(os.path.join(ROOT, "alphazero_quoridor_b200", "csrc", f))
            src_cache[f] = open(p[0]).read().split("\n") if p else []
        text = src_cache[f][l - 1].strip()[:90] if 0 < l <= len(src_cache[f]) else ""
        print("%5.1f%% inst  %5.1f%% stall  %5.1f thr  %s:%d  %s" % (100.0 * n / tot, 100.0 * sm / max(stot, 1), tn / max(n, 1), f, l, text))


if __name__ == "__main__":
    main()
