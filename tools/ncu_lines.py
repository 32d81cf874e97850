#!/usr/bin/env python
"""Per-source-line view of an ncu capture: joins `ncu --page source --csv` (SASS rows with stall samples and executed
instruction counts) with the line table of the cubin (`nvdisasm -g`), by instruction offset inside the kernel.

    ncu -i X.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libba_cuda.so            # -> demod.sm_100a.cubin ...
    python tools/ncu_lines.py src.csv demod.sm_100a.cubin demod_full_kernel [--top 40] [--ranges 330-585,589-668]

Prints the hottest source lines (stall samples, warp instructions executed) and optional totals per line range.
"""
import argparse
import csv
import re
import subprocess
import sys


def line_table(cubin, kernel):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    table, inside, line = {}, False, 0
    for ln in out:
        if ln.startswith("//-") and ".text." in ln:
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if m.group(1).endswith(".cu"):
                line = int(m.group(2))
            elif m.group(1).endswith(".h") and "/csrc/" in m.group(1):
                line = -int(m.group(2))  # a line of one of the kernel's own headers (ba_port.h ...): negative, so that ranges over the .cu leave it out
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (line, m.group(2).strip())
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("cubin")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--ranges", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv, errors="replace")))
    hdr = rows[1]
    ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    body = [r for r in rows[2:] if len(r) > iex and r[ia].startswith("0x")]
    base = int(body[0][ia], 16)
    for i in range(1, len(body)):  # the page lists the kernel once per view (SASS, PTX-correlated ...): keep the first pass over the addresses
        if int(body[i][ia], 16) < int(body[i - 1][ia], 16):
            body = body[:i]
            break
    table = line_table(a.cubin, a.kernel)
    per = {}
    tot_s = tot_e = 0
    for r in body:
        off = int(r[ia], 16) - base
        line = table.get(off, (0, ""))[0]
        s, e = int(r[isamp] or 0), int(r[iex] or 0)
        p = per.setdefault(line, [0, 0, 0])
        p[0] += s
        p[1] += e
        p[2] += 1
        tot_s += s
        tot_e += e
    print("total samples %d, warp instructions %d, SASS rows %d" % (tot_s, tot_e, len(body)))
    print("%6s %8s %6s %12s %6s" % ("line", "samples", "%", "warp-inst", "sass"))
    for line, (s, e, n) in sorted(per.items(), key=lambda kv: -kv[1][0])[: a.top]:
        print("%6d %8d %5.1f%% %12d %6d" % (line, s, 100.0 * s / max(1, tot_s), e, n))
    if a.ranges:
        print("ranges:")
        for rg in a.ranges.split(","):
            lo, hi = [int(x) for x in rg.split("-")]
            s = sum(v[0] for k, v in per.items() if lo <= k <= hi)
            e = sum(v[1] for k, v in per.items() if lo <= k <= hi)
            print("  %5d-%-5d samples %8d (%5.1f%%)  warp-inst %12d (%5.1f%%)" % (lo, hi, s, 100.0 * s / max(1, tot_s), e, 100.0 * e / max(1, tot_e)))


if __name__ == "__main__":
    main()
