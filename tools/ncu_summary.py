"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + the hottest source lines.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--lines 15]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 15
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("==", d.get("Kernel Name", "?")[:100])
        for k in KEYS:
            if k in d:
                print("  %-90s %s %s" % (k, d[k], units[hdr.index(k)]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) < 3:
        return
    hdr = rows[0]
    def col(name):
        for i, h in enumerate(hdr):
            if h.strip() == name:
                return i
        return None
    c_src, c_samp, c_inst = col("Source"), col("# Samples") or col("Warp Stall Sampling (All Samples)"), col("Instructions Executed")
    if c_samp is None:
        print("columns:", hdr[:20]); return
    body = [r for r in rows[1:] if len(r) > max(c_src, c_samp) and r[c_samp].replace(".", "").isdigit()]
    tot = sum(float(r[c_samp]) for r in body) or 1
    body.sort(key=lambda r: -float(r[c_samp]))
    print("-- hottest source lines (stall samples)")
    for r in body[:nlines]:
        print("  %5.1f%%  %s" % (100 * float(r[c_samp]) / tot, r[c_src].strip()[:150]))


main()
