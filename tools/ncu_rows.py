#!/usr/bin/env python
"""Print one line per captured launch of an `ncu --page raw --csv` export with the metrics that decide what bounds a
streaming kernel: duration, issue-active, resident warps, L1/shared pipe, DRAM, pipe shares, top stall reasons."""
import csv
import sys

SEL = [("gpu__time_duration.sum", "ms"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
       ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%"),
       ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wf%"),
       ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
       ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
       ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
       ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
       ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmaH%"), ("sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active", "fmaL%"),
       ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
       ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensI%"), ("smsp__inst_executed.sum", "Minst"),
       ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_smem"), ("launch__occupancy_limit_registers", "occ_reg")]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
REASONS = ["long_scoreboard", "short_scoreboard", "mio_throttle", "lg_throttle", "wait", "math_pipe_throttle", "not_selected", "no_instruction",
           "barrier", "dispatch_stall", "branch_resolving", "imc_miss", "selected"]


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernel".ljust(34), " ".join(n.rjust(8) for _, n in SEL), " stalls (warps per issue)")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("(anonymous namespace)::", "").replace("unnamed>::", "").replace("void ", "")
        vals = []
        for m, n in SEL:
            v = r[idx[m]] if m in idx else "nan"
            try:
                x = float(v.replace(",", ""))
                if n in ("rdMB", "wrMB"):
                    u = units[idx[m]]
                    x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                if n == "Minst":
                    x /= 1e6
                if n == "ms":
                    x *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}.get(units[idx[m]], 1.0)
                vals.append(("%.3f" % x if n == "ms" else "%.1f" % x).rjust(8))
            except ValueError:
                vals.append(v.rjust(8))
        st = []
        for rs in REASONS:
            k = STALL % rs
            if k in idx:
                try:
                    st.append((float(r[idx[k]]), rs))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print(name[:34].ljust(34), " ".join(vals), "  ".join("%s=%.2f" % (b, a) for a, b in st[:5]))


if __name__ == "__main__":
    main(sys.argv[1])
