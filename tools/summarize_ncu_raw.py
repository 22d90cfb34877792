"""Turns `ncu -i report.ncu-rep --page raw --csv` into the per-launch markdown table kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv && python tools/summarize_ncu_raw.py raw.csv "title" > profiles/x.md
"""
import csv
import re
import sys

COLS = [("grid", "launch__grid_size", "%d"), ("time us", "gpu__time_duration.sum", "%.1f"), ("dram rd MB", "dram__bytes_read.sum", "%.1f"),
        ("dram wr MB", "dram__bytes_write.sum", "%.1f"), ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%.1f"),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "%.1f"), ("L1 %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "%.1f"),
        ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active", "%.1f"),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "%.1f"),
        ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active", "%.1f"), ("regs", "launch__registers_per_thread", "%d")]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main(path, title):
    rows = list(csv.reader(ln for ln in open(path) if not ln.startswith("==")))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {}
    for _, metric, _ in COLS:
        hit = [i for i, h in enumerate(hdr) if h == metric] or [i for i, h in enumerate(hdr) if h.endswith("." + metric)]
        idx[metric] = hit[0] if hit else None
    name_i = hdr.index("Kernel Name")
    print("# " + title + "\n")
    print("| # | kernel | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    total = 0.0
    for n, r in enumerate(body, 1):
        name = re.sub(r"\(.*", "", r[name_i]).replace("void ", "").replace("effimvs::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        cells = []
        for _, metric, fmt in COLS:
            i = idx[metric]
            try:
                v = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            except (TypeError, ValueError, IndexError):
                cells.append("-")
                continue
            if metric == "gpu__time_duration.sum":
                total += v
            cells.append(fmt % v)
        print("| %d | %s | %s |" % (n, name, " | ".join(cells)))
    print("\ntotal %.1f us over %d launches (under the profiler: cold caches, serialised)" % (total, len(body)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu --set full")
