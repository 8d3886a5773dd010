"""Condenses an ncu report (--set full --import-source on) into the text summary kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.summary.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
    "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print("kernel", col.get("Kernel Name", ("", "?"))[1])
    for m in METRICS:
        if m in col:
            print(m, col[m][0], col[m][1])
    stalls = {}
    for h, (u, v) in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or \
           h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio"):
            name = h.split("stalled_")[1].replace("_per_issue_active.ratio", "").replace(".ratio", "")
            try:
                stalls[name] = float(v.replace(",", ""))
            except ValueError:
                pass
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:8]
    print("stalls (warp-cycles per issue):", ", ".join("%s=%.2f" % kv for kv in top))
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    h = src[1]
    si, ei, wi = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    data = [(int(r[wi]), int(r[ei]), r[si].strip()) for r in src[2:] if len(r) > wi]
    total_w = sum(d[0] for d in data)
    total_e = sum(d[1] for d in data)
    print("instructions executed (warp level):", total_e, " static:", len(data))
    print("top stall instructions (of %d samples):" % total_w)
    for w, e, s in sorted(data, key=lambda d: -d[0])[:12]:
        print("  %5d %5.1f%%  %s | execs %d" % (w, 100.0 * w / max(total_w, 1), s, e))
    hist = {}
    for w, e, s in data:
        hist.setdefault(e, [0, 0])
        hist[e][0] += 1
        hist[e][1] += w
    print("execution-count groups (execs x static instructions = share of executed, share of stall samples):")
    for e, (c, w) in sorted(hist.items(), key=lambda kv: -kv[0] * kv[1][0])[:10]:
        print("  %9d x %4d = %5.1f%%  stalls %5.1f%%" % (e, c, 100.0 * e * c / max(total_e, 1), 100.0 * w / max(total_w, 1)))


if __name__ == "__main__":
    main()
