"""Instruction / stall-sample shares per source-line bin of an .ncu-rep: python tools/src_bins.py rep [binsize] [file-substring]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; binsz = int(sys.argv[2]) if len(sys.argv) > 2 else 20; sub = sys.argv[3] if len(sys.argv) > 3 else "wfl_fast"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "File Path": cur = {"file": r[1], "rows": []}; blocks.append(cur)
    elif r and r[0] == "Function Name" and cur is not None: cur["fn"] = r[1]
    elif r and r[0] == "Line No" and cur is not None: cur["hdr"] = r
    elif cur is not None and "hdr" in cur: cur["rows"].append(r)
grand = gs = 0; res = []
for b in blocks:
    hdr = b["hdr"]; ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); it = hdr.index("Thread Instructions Executed")
    per = collections.OrderedDict(); line = None
    for r in b["rows"]:
        if len(r) != len(hdr): continue
        if r[0].isdigit(): line = int(r[0]); per.setdefault(line, [r[1], 0, 0, 0]); continue
        if r[ii].isdigit() and line is not None:
            per[line][1] += int(r[ii]); per[line][2] += int(r[isamp]); per[line][3] += int(r[it])
    tot = sum(v[1] for v in per.values()); ts = sum(v[2] for v in per.values())
    grand += tot; gs += ts; res.append((b, per, tot, ts))
for b, per, tot, ts in res:
    print("%-70s inst %5.1f%% samples %5.1f%%" % (b["file"][-70:], 100 * tot / grand, 100 * ts / max(gs, 1)))
    if sub in b["file"]:
        bins = collections.Counter(); sb = collections.Counter(); tb = collections.Counter()
        for l, v in per.items(): bins[l // binsz * binsz] += v[1]; sb[l // binsz * binsz] += v[2]; tb[l // binsz * binsz] += v[3]
        for k in sorted(bins):
            if bins[k] > 0.004 * grand or sb[k] > 0.004 * gs:
                first = next((per[l][0] for l in range(k, k + binsz) if l in per and per[l][1] > 0), "")
                print("  L%4d-%4d inst %5.1f%% samples %5.1f%% thr/inst %4.1f | %s" % (k, k + binsz - 1, 100 * bins[k] / grand, 100 * sb[k] / max(gs, 1), tb[k] / max(bins[k], 1), first.strip()[:90]))
