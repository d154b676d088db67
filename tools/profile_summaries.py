"""Turn the two ncu artefacts of a bench run into the small text files kept under profiles/.

usage: python tools/profile_summaries.py <launch_list.csv> <full.ncu-rep> <tag> [first_launch_of_step]
  launch_list.csv : ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file ...
  full.ncu-rep    : ncu --set full --import-source on -k regex:wfl_pipe_ ...
Writes profiles/<tag>_launches_summary.csv, profiles/<tag>_kernels_ncu.txt and profiles/roofline_traffic.json.
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    out = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        d = out.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("unnamed>::", "")})
        v, unit = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
        else:
            d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return list(out.values())


def main():
    lst, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    L = launches(lst)
    prep = [i for i, l in enumerate(L) if l["name"] == "wfl_pipe_prepare"]
    first = int(sys.argv[4]) if len(sys.argv) > 4 else prep[4]   # 5th whole-batch pass: after the warm-up steps
    end = next(i for i in range(first, len(L)) if L[i]["name"] == "wfl_compact_scatter") + 1
    step = L[first:end]
    tot = sum(s["ns"] for s in step)
    agg = collections.OrderedDict()
    for s_ in step:
        a = agg.setdefault(s_["name"], [0, 0, 0, 0])
        a[0] += s_["ns"]; a[1] += s_["dram__bytes_read.sum"]; a[2] += s_["dram__bytes_write.sum"]; a[3] += 1
    with open("profiles/%s_launches_summary.csv" % tag, "w") as f:
        f.write("# ncu launch list of python bench.py --steps 2 --warmup 3 --no-cpu-baseline (cfg2, 100k contigs, pipeline mode)\n")
        f.write("# one resident step (launches %d..%d); ncu --metrics gpu__time_duration.sum,dram__bytes_* --clock-control none\n" % (first, end - 1))
        f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("kernel, launches, total_ns, share, dram_read_bytes, dram_write_bytes\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write("%s, %d, %d, %.4f, %d, %d\n" % (k, a[3], a[0], a[0] / tot, a[1], a[2]))
        f.write("TOTAL, %d, %d, 1.0, %d, %d\n" % (len(step), tot, sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values())))
    dram = sum(a[1] + a[2] for k, a in agg.items() if k.startswith("wfl_pipe"))
    json.dump({"workload": "cfg2", "contigs": 100000, "kernel": "wfl_pipe_* (all launches of one step)",
               "dram_bytes_per_launch": int(dram),
               "source": "profiles/%s_launches_summary.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the "
                         "wfl_pipe_* launches of one resident step of bench.py)" % tag},
              open("profiles/roofline_traffic.json", "w"), indent=1)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open("profiles/%s_kernels_ncu.txt" % tag, "w") as f:
        f.write("# python bench.py --steps 1 --warmup 3 --no-cpu-baseline (cfg2, 100000 contigs), first whole-batch pass\n")
        f.write("# ncu --set full --clock-control none --import-source on -k regex:wfl_pipe_\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("kernel: %s\n" % d.get("Kernel Name"))
            for k in KEYS:
                if k in d:
                    f.write("  %-88s %s %s\n" % (k, d[k], units[hdr.index(k)]))
            f.write("\n")
    print(open("profiles/%s_launches_summary.csv" % tag).read())


if __name__ == "__main__":
    main()
