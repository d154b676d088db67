"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/<name>.txt
Prints the headline raw metrics of the first kernel in the report and the top CUDA source
lines by executed warp instructions / stall samples (needs -lineinfo + --import-source on).
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    print("# " + title)
    print("# source: {} (ncu --set full --clock-control none --import-source on)".format(rep))
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("kernel:", d.get("Kernel Name", ("?",))[0])
    for k in KEYS:
        if k in d:
            print("  {:<88s} {} {}".format(k, d[k][0], d[k][1]))
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hi]
    i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    lines, stalls = [], {h: 0 for _, h in stall_cols}
    for r in rows[hi + 1:]:
        if len(r) == len(hdr) and r[0].isdigit():
            try:
                lines.append((int(r[0]), r[1].strip(), int(r[i_s]), int(r[i_i]), int(r[i_t])))
                for i, h in stall_cols:
                    stalls[h] += int(r[i])
            except ValueError:
                pass
    ts, ti = sum(l[2] for l in lines) or 1, sum(l[3] for l in lines) or 1
    print("\nwarp-stall samples by reason (all source lines):")
    for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
        print("  {:<28s} {:5.1f}%".format(h, 100.0 * v / ts))
    print("\ntop source lines by executed warp instructions (total {:.3e}):".format(ti))
    for l in sorted(lines, key=lambda x: -x[3])[:25]:
        print("  L{:<5d} inst {:5.1f}%  samples {:5.1f}%  thr/inst {:4.1f} | {}".format(
            l[0], 100.0 * l[3] / ti, 100.0 * l[2] / ts, l[4] / max(1, l[3]), l[1][:96]))


if __name__ == "__main__":
    main()
