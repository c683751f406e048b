"""Summarise an `ncu --set full` report as a markdown table, one row per captured launch.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-name-substring] > profiles/rNN_ncu_x.md

Columns: duration, DRAM bytes read / written, DRAM / L2 / L1 throughput (% of peak), tensor pipe busy %, issue slots
busy %, executed warp instructions, registers, dynamic shared memory, achieved occupancy.  Cold-cache, serialised replays:
compare shares and ratios, not absolute times (B200_PROFILING.md)."""
import csv
import subprocess
import sys

COLS = [
    ("time us", "gpu__time_duration.sum", lambda v, u: "%.1f" % (v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v)),
    ("dram rd MB", "dram__bytes_read.sum", None),
    ("dram wr MB", "dram__bytes_write.sum", None),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", lambda v, u: "%.1f" % v),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", lambda v, u: "%.1f" % v),
    ("L1 %", "l1tex__throughput.avg.pct_of_peak_sustained_active", lambda v, u: "%.1f" % v),
    ("tensor pipe %", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", lambda v, u: "%.1f" % v),
    ("issue busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active", lambda v, u: "%.1f" % v),
    ("warp instr M", "smsp__inst_executed.sum", lambda v, u: "%.2f" % (v / 1e6)),
    ("regs", "launch__registers_per_thread", lambda v, u: "%d" % v),
    ("dyn smem KB", "launch__shared_mem_per_block_dynamic", None),
    ("occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active", lambda v, u: "%.1f" % v),
]
UNIT_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
UNIT_KB = {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}


def main():
    rep = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {c: i for i, c in enumerate(hdr)}
    print("| # | kernel | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    n = 0
    for r in data:
        name = r[idx["Kernel Name"]]
        if flt and flt not in name:
            continue
        cells = []
        for label, key, fmt in COLS:
            if key not in idx or r[idx[key]] == "":
                cells.append("-")
                continue
            v = float(r[idx[key]].replace(",", ""))
            u = units[idx[key]]
            if label.endswith("MB"):
                cells.append("%.1f" % (v * UNIT_MB.get(u, 1.0)))
            elif label.endswith("KB"):
                cells.append("%.1f" % (v * UNIT_KB.get(u, 1.0)))
            else:
                cells.append(fmt(v, u))
        short = name.split("(")[0].replace("io::", "")
        print("| %d | %s | %s |" % (n, short, " | ".join(cells)))
        n += 1


if __name__ == "__main__":
    main()
