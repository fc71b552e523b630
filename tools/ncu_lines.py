#!/usr/bin/env python
"""Per-source-line totals (stall samples, warp instructions) of one kernel from an ncu report captured with
--import-source on:   python tools/ncu_lines.py report.ncu-rep k_descriptor [min_pct]"""
import csv
import subprocess
import sys


def main(rep, kernel, min_pct=1.0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    lines = []
    for r in rows:
        if len(r) >= 8 and r[0].isdigit():
            try:
                lines.append((int(r[0]), r[1], float(r[6]), float(r[7])))
            except ValueError:
                pass
    ts = sum(l[2] for l in lines) or 1.0
    ti = sum(l[3] for l in lines) or 1.0
    print("samples %.0f  warp-instructions %.0f" % (ts, ti))
    for ln, src, s, n in sorted(lines):
        if 100 * s / ts >= min_pct or 100 * n / ti >= min_pct:
            print("%5d  smp %5.1f%%  inst %5.1f%%  %s" % (ln, 100 * s / ts, 100 * n / ti, src.strip()[:120]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
