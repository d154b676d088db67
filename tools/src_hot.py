"""Top source lines of an `ncu --page source --csv --print-source cuda,sass` dump: python tools/src_hot.py file.csv [n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
blocks, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Function Name" and cur is not None:
        cur["fn"] = r[1]
    elif r and r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur:
        cur["rows"].append(r)
grand = 0
per_block = []
for b in blocks:
    hdr = b["hdr"]
    ii, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    per, line = collections.OrderedDict(), None
    for r in b["rows"]:
        if len(r) != len(hdr):
            continue
        if r[0].isdigit():
            line = int(r[0])
            per.setdefault(line, [r[1], 0, 0, 0])
            continue
        if r[ii].isdigit() and line is not None:
            per[line][1] += int(r[ii]); per[line][2] += int(r[it]); per[line][3] += int(r[isamp])
    tot = sum(v[1] for v in per.values())
    grand += tot
    per_block.append((b, per, tot))
for b, per, tot in per_block:
    if tot < 0.02 * grand:
        continue
    ts = sum(v[3] for v in per.values()) or 1
    print("%s  %s: %.1f%% of all warp instructions" % (b["file"].split("/")[-1], b.get("fn", "?")[:40], 100.0 * tot / grand))
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:top_n]:
        print("   %5d %5.1f%% inst %5.1f%% smp  thr %4.1f  %s" % (k, 100.0 * v[1] / grand, 100.0 * v[3] / ts, v[2] / max(v[1], 1), v[0][:110]))
