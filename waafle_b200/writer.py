"""Result unpacking and the three main TSV outputs.

Turns the engine's compacted result arrays back into the rows the reference writes in
`write_main_output_files` (waafle/waafle_orgscorer.py:814-894): same columns, same field
encodings (OS:750-764, 793-800), same float formatting and `--` for empty cells
(waafle/utils.py:137-139), rows sorted by contig name (OS:838).
"""

import os

import numpy as np

from .utils import format_field, say, write_row

CALLS = ("unclassified", "no_lgt", "lgt")
FLAG_RETAINED, FLAG_IGNORED = 1, 2

C_ANNOTATION_PREFIX = "ANNOTATIONS:"   # OS:60
C_MISSING_ANNOTATION = "None"          # OS:61
DIRECTIONS = ("A?B", "B>A")            # OS:485, OS:543

# OS:78-114
FORMATS = {
    "lgt": ["contig_name", "call", "contig_length", "min_max_score", "avg_max_score", "synteny",
            "direction", "clade_A", "clade_B", "lca", "melded_A", "melded_B", "taxonomy_A",
            "taxonomy_B", "loci"],
    "no_lgt": ["contig_name", "call", "contig_length", "min_score", "avg_score", "synteny",
               "clade", "melded", "taxonomy", "loci"],
    "unclassified": ["contig_name", "call", "contig_length", "loci"],
}


def _tails_field(taxonomy, members, lca):
    """OS:750-759 on the distinct melded clades."""
    items = set()
    for m in members:
        t = taxonomy.get_tail(int(m), int(lca))
        if len(t) > 0:
            items.add("|".join(t))
    return "; ".join(sorted(items))


def build_records(batch, loci, hits, taxonomy, res):
    """One dict per contig (FASTA order) with the reference's row values, unformatted."""
    names = taxonomy.names
    S = res["ann_winner"].shape[1] if res["ann_winner"].ndim == 2 else 0
    systems = list(hits.systems)[:S] if hits is not None else []
    records = []
    for c in range(batch.n_contigs):
        l0, l1 = int(batch.locus_off[c]), int(batch.locus_off[c + 1])
        kept = [j for j in range(l0, l1) if res["locus_flags"][j] & FLAG_RETAINED]
        rec = dict(
            contig_name=batch.contig_names[c], call=CALLS[int(res["call"][c])],
            contig_length=int(batch.contig_lengths[c]),
            loci="|".join(loci.code(int(batch.locus_row[j])) for j in kept),
            ignore=[bool(res["locus_flags"][j] & FLAG_IGNORED) for j in kept],
            lifts=int(res["lifts"][c]))
        ann = []
        for j in kept:
            d = {}
            for s in range(S):
                w = int(res["ann_winner"][j, s])
                if w >= 0:
                    sid = int(hits.sseqid_id[int(batch.hit_row[w])])
                    d[systems[s]] = hits.sseqid_annotations[sid][systems[s]]
            ann.append(d)
        rec["annotations"] = ann
        call = int(res["call"][c])
        if call:
            m0, m1 = int(res["member_off"][c]), int(res["member_off"][c + 1])
            na = int(res["n_members_a"][c])
            mem = res["members"][m0:m1]
            c1 = int(res["clade1"][c])
            syn = bytes(res["synteny"][kept]).decode("ascii") if kept else ""
            if call == 1:
                rec.update(min_score=float(res["crit"][c]), avg_score=float(res["rank"][c]),
                           synteny=syn, clade=names[c1],
                           melded=_tails_field(taxonomy, mem[:na], c1),
                           taxonomy="|".join(taxonomy.get_lineage(c1)))
            else:
                c2 = int(res["clade2"][c])
                rec.update(min_max_score=float(res["crit"][c]),
                           avg_max_score=float(res["rank"][c]), synteny=syn,
                           direction=DIRECTIONS[int(res["direction"][c])],
                           clade_A=names[c1], clade_B=names[c2], lca=names[int(res["lca"][c])],
                           melded_A=_tails_field(taxonomy, mem[:na], c1),
                           melded_B=_tails_field(taxonomy, mem[na:], c2),
                           taxonomy_A="|".join(taxonomy.get_lineage(c1)),
                           taxonomy_B="|".join(taxonomy.get_lineage(c2)))
        records.append(rec)
    return records


def write_main_output_files(records, outdir, basename):
    """Write <basename>.{lgt,no_lgt,unclassified}.tsv (OS:814-894)."""
    say("Initializing outputs.")
    # annotation systems that were actually transferred to some locus (OS:824-828)
    systems = sorted({s for r in records for d in r["annotations"] for s in d})
    handles = {}
    for option in ("lgt", "no_lgt", "unclassified"):
        handles[option] = open(os.path.join(outdir, ".".join([basename, option, "tsv"])), "w")
        cols = FORMATS[option] + [C_ANNOTATION_PREFIX + s for s in systems]
        write_row([k.upper() for k in cols], handles[option])
    for r in sorted(records, key=lambda r: r["contig_name"]):
        option = r["call"]
        cells = [format_field(r[f]) for f in FORMATS[option]]
        for s in systems:
            cells.append(format_field("|".join(d.get(s, C_MISSING_ANNOTATION)
                                               for d in r["annotations"])))
        write_row(cells, handles[option])
    for h in handles.values():
        h.close()
