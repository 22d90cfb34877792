"""Digest of one kernel of an .ncu-rep (run where ncu is installed, no GPU needed): headline counters, stall mix and the
SASS opcode histogram / hottest source lines from the source page.

    python tools/ncu_digest.py gpurun_out/x.ncu-rep [launch index] [--lines]
"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, which, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO("\n".join(ln for ln in out.splitlines() if not ln.startswith("==")))))


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
    rows = page(rep, "raw")
    hdr, units, r = rows[0], rows[1], rows[2 + which]
    idx = {h: i for i, h in enumerate(hdr)}
    print(r[idx["Kernel Name"]][:100])
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__t_sector_hit_rate.pct",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem"]
    for k in keys:
        if k in idx:
            print("  {:70s} {} {}".format(k, r[idx[k]], units[idx[k]]))
    print("  stalls (warps per issue):")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[idx[h]])
            except ValueError:
                continue
            if v >= 0.05:
                print("    {:28s} {:.2f}".format(h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    src = page(rep, "source", ("--print-source", "sass"))
    # the source page lists the kernels one after the other: take block number `which`
    blocks, cur = [], None
    for row in src:
        if row and row[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None:
            cur.append(row)
    if which >= len(blocks):
        return
    b = blocks[which]
    h = b[0]
    ia, isrc, ismp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    body = [x for x in b[1:] if len(x) > ia and x[ia].isdigit()]
    tot = sum(int(x[ia]) for x in body)
    ops = collections.Counter()
    for x in body:
        t = x[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += int(x[ia])
    print("  static SASS {}, executed {}".format(len(body), tot))
    print("  " + ", ".join("{} {:.1f}%".format(o, 100.0 * c / tot) for o, c in ops.most_common(24)))
    if "--lines" in sys.argv:
        top = sorted(body, key=lambda x: -int(x[ismp]))[:40]
        for x in top:
            print("   {:>8s} smp {:>10s} exe  {}".format(x[ismp], x[ia], x[isrc].strip()[:100]))


if __name__ == "__main__":
    main()
