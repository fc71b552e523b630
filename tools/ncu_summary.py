#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export (or a launch list made with --metrics gpu__time_duration.sum) into one
row per kernel: launches, total/average duration, share of the captured time, DRAM traffic, issue and occupancy
figures. Used to produce the summaries committed under profiles/.

    python tools/ncu_summary.py gpurun_out/r1w_full_raw.csv > profiles/r1w_ncu_summary.csv
"""
import collections
import csv
import re
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
]
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3,
              "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def short_name(n):
    n = re.sub(r"\(anonymous namespace\)::|<?unnamed>::|akz::|void ", "", n)
    m = re.match(r"([A-Za-z0-9_]+)(<[^(]*>)?", n)
    return (m.group(1) + (m.group(2) or "")) if m else n[:60]


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    # skip ncu's ==PROF== lines in launch lists written with --log-file
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[start], rows[start + 1], rows[start + 2:]
    idx = {h: i for i, h in enumerate(hdr)}
    if "Metric Name" in idx:  # long format (one row per metric): pivot
        piv = collections.OrderedDict()
        for r in data:
            if len(r) <= idx["Metric Value"]:
                continue
            key = r[idx["ID"]]
            d = piv.setdefault(key, {"Kernel Name": r[idx["Kernel Name"]]})
            d[r[idx["Metric Name"]]] = (r[idx["Metric Value"]], r[idx["Metric Unit"]])
        recs = list(piv.values())
        get = lambda rec, m: rec.get(m, ("nan", ""))
    else:
        recs = [{"Kernel Name": r[idx["Kernel Name"]], **{m: (r[idx[m]], units[idx[m]]) for m, _ in METRICS if m in idx}} for r in data if len(r) == len(hdr)]
        get = lambda rec, m: rec.get(m, ("nan", ""))
    agg = collections.OrderedDict()
    for rec in recs:
        agg.setdefault(short_name(rec["Kernel Name"]), []).append(rec)

    def val(rec, m):
        v, u = get(rec, m)
        x = fnum(v)
        if m == "gpu__time_duration.sum" or m.startswith("dram__bytes"):
            x *= UNIT_SCALE.get(u, 1.0)
        return x

    total = sum(val(r, "gpu__time_duration.sum") for rs in agg.values() for r in rs)
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "total_us", "share_pct", "avg_us", "dram_MB_per_launch", "dram_pct_of_peak", "issue_active_pct",
                "warps_active_pct", "l1tex_pct", "l2_pct", "tensor_pipe_pct", "regs", "warp_inst_per_launch", "smem_wavefronts_per_launch"])
    for k, rs in sorted(agg.items(), key=lambda kv: -sum(val(r, "gpu__time_duration.sum") for r in kv[1])):
        t = sum(val(r, "gpu__time_duration.sum") for r in rs)

        def wavg(m):
            xs = [(val(r, m), val(r, "gpu__time_duration.sum")) for r in rs]
            xs = [(a, b) for a, b in xs if a == a]
            return sum(a * b for a, b in xs) / sum(b for _, b in xs) if xs else float("nan")

        def avg(m):
            xs = [val(r, m) for r in rs]
            xs = [a for a in xs if a == a]
            return sum(xs) / len(xs) if xs else float("nan")

        w.writerow([k, len(rs), "%.1f" % t, "%.2f" % (100 * t / total), "%.2f" % (t / len(rs)),
                    "%.2f" % (avg("dram__bytes_read.sum") + avg("dram__bytes_write.sum")), "%.1f" % wavg(METRICS[3][0]), "%.1f" % wavg(METRICS[4][0]),
                    "%.1f" % wavg(METRICS[5][0]), "%.1f" % wavg(METRICS[6][0]), "%.1f" % wavg(METRICS[7][0]), "%.1f" % wavg(METRICS[9][0]),
                    "%.0f" % avg("launch__registers_per_thread"), "%.0f" % avg("smsp__inst_executed.sum"), "%.0f" % avg(METRICS[12][0])])


if __name__ == "__main__":
    main(sys.argv[1])
