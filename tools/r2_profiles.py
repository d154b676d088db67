"""Round-2 profile artefacts for profiles/ from one GPU visit:

  python tools/r2_profiles.py <launch_list.csv> <full.ncu-rep> <sass_listing.txt or -> <tag>

  launch_list.csv : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python bench.py --steps 2 --warmup 1 ...
  full.ncu-rep    : ncu --set full --import-source on -k regex:wfl_fast_contigs -c 1 ... (the resident first pass at 100k contigs)
Writes profiles/<tag>_launches_raw.csv (copy), profiles/<tag>_launches_summary.csv (per kernel: launches, total ms, share),
profiles/<tag>_fast_kernel_ncu.txt and profiles/roofline_traffic.json (DRAM bytes of the dominant kernel, per launch).
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    lst, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(open(lst)) if r and r[0].isdigit()]
    shutil.copy(lst, os.path.join(ROOT, "profiles", tag + "_launches_raw.csv"))
    per = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").replace("unnamed>::", "").strip()
        per.setdefault(name, [0, 0.0])
        per[name][0] += 1
        per[name][1] += float(r[-1]) / 1e6
    tot = sum(v[1] for v in per.values())
    with open(os.path.join(ROOT, "profiles", tag + "_launches_summary.csv"), "w") as fh:
        fh.write("# {}: every launch of `ncu --metrics gpu__time_duration.sum` over the command in the raw file's header\n".format(tag))
        fh.write("kernel,launches,total_ms,share\n")
        for k, (n, ms) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            fh.write("{},{},{:.4f},{:.4f}\n".format(k, n, ms, ms / tot))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep,
                          tag + ": wfl_fast_contigs, cfg2 100k contigs resident (first pass over all contigs)"],
                         capture_output=True, text=True).stdout
    with open(os.path.join(ROOT, "profiles", tag + "_fast_kernel_ncu.txt"), "w") as fh:
        fh.write(out)
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}

    def nbytes(key):
        v, u = d[key]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    traffic = int(nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"))
    with open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w") as fh:
        json.dump({"workload": "cfg2", "contigs": 100000, "mode": "fast", "kernel": "wfl_fast_contigs (first pass, all contigs)",
                   "dram_bytes_per_launch": traffic,
                   "source": "profiles/{}_fast_kernel_ncu.txt (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of the "
                             "resident launch over 100 000 cfg2 contigs, wide 29 B/hit layout)".format(tag)}, fh, indent=1)
    print("traffic", traffic, "launch kernels", len(per))


if __name__ == "__main__":
    main()
