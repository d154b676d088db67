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


def _winner_annotations(batch, hits, res, systems):
    """{(locus row j, system index s): value} for every transferred annotation.  The subject headers of all
    winning hits are fetched and parsed once (they live in an Arrow array on the fast front end)."""
    aw = np.asarray(res["ann_winner"])
    if aw.ndim != 2 or aw.shape[1] == 0 or hits is None or not systems:
        return {}
    jj, ss = np.nonzero(aw >= 0)
    if len(jj) == 0:
        return {}
    rows = np.asarray(batch.hit_row)[aw[jj, ss]]
    sid = rows if hits.sseqid_id is None else np.asarray(hits.sseqid_id)[rows]   # None: one header per row (GPU parser)
    uniq = np.unique(sid)
    parsed = {int(u): hits.sseqid_annotations[int(u)] for u in uniq}
    return {(int(j), int(s)): parsed[int(i)][systems[int(s)]] for j, s, i in zip(jj.tolist(), ss.tolist(), sid.tolist())}


def build_records(batch, loci, hits, taxonomy, res):
    """One dict per contig (FASTA order) with the reference's row values, unformatted."""
    names = taxonomy.names
    S = res["ann_winner"].shape[1] if res["ann_winner"].ndim == 2 else 0
    systems = list(hits.systems)[:S] if hits is not None else []
    # plain Python lists: the loop below touches every element once, and numpy scalar access is ~10x slower
    locus_off = np.asarray(batch.locus_off).tolist()
    flags = np.asarray(res["locus_flags"]).tolist()
    locus_row = np.asarray(batch.locus_row).tolist()
    calls = np.asarray(res["call"]).tolist()
    lifts = np.asarray(res["lifts"]).tolist()
    lengths = np.asarray(batch.contig_lengths).tolist()
    member_off = np.asarray(res["member_off"]).tolist()
    n_members_a = np.asarray(res["n_members_a"]).tolist()
    members = np.asarray(res["members"]).tolist()
    clade1, clade2, lca = (np.asarray(res[k]).tolist() for k in ("clade1", "clade2", "lca"))
    crit, rank = np.asarray(res["crit"]).tolist(), np.asarray(res["rank"]).tolist()
    direction = np.asarray(res["direction"]).tolist()
    synteny = bytes(np.asarray(res["synteny"], dtype=np.uint8)).decode("latin-1")
    winners = _winner_annotations(batch, hits, res, systems)
    lineage_cache = {}

    def lineage(c):
        if c not in lineage_cache:
            lineage_cache[c] = "|".join(taxonomy.get_lineage(c))
        return lineage_cache[c]

    records = []
    for c in range(batch.n_contigs):
        l0, l1 = locus_off[c], locus_off[c + 1]
        kept = [j for j in range(l0, l1) if flags[j] & FLAG_RETAINED]
        rec = dict(
            contig_name=batch.contig_names[c], call=CALLS[calls[c]],
            contig_length=lengths[c],
            loci="|".join(loci.code(locus_row[j]) for j in kept),
            ignore=[bool(flags[j] & FLAG_IGNORED) for j in kept],
            lifts=lifts[c])
        ann = []
        for j in kept:
            d = {}
            for s in range(S):
                v = winners.get((j, s))
                if v is not None:
                    d[systems[s]] = v
            ann.append(d)
        rec["annotations"] = ann
        call = calls[c]
        if call:
            m0, m1 = member_off[c], member_off[c + 1]
            na = n_members_a[c]
            mem = members[m0:m1]
            c1 = clade1[c]
            syn = "".join(synteny[j] for j in kept)
            if call == 1:
                rec.update(min_score=crit[c], avg_score=rank[c],
                           synteny=syn, clade=names[c1],
                           melded=_tails_field(taxonomy, mem[:na], c1),
                           taxonomy=lineage(c1))
            else:
                c2 = clade2[c]
                rec.update(min_max_score=crit[c],
                           avg_max_score=rank[c], synteny=syn,
                           direction=DIRECTIONS[direction[c]],
                           clade_A=names[c1], clade_B=names[c2], lca=names[lca[c]],
                           melded_A=_tails_field(taxonomy, mem[:na], c1),
                           melded_B=_tails_field(taxonomy, mem[na:], c2),
                           taxonomy_A=lineage(c1),
                           taxonomy_B=lineage(c2))
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


# ---------------------------------------------------------------------------------------------------------------
# --write-details (OS:766-812)
# ---------------------------------------------------------------------------------------------------------------

DETAILS_FORMAT = ["contig_name", "iteration", "clade", "gene_scores", "gene_spans"]   # OS:116-122
C_PRECISION = 3                                                                       # OS:59
C_DELIM3 = ":"                                                                        # OS:66


def _calc_overlap(a1, a2, b1, b2):
    """UT:487-500."""
    a1, a2 = sorted([a1, a2])
    b1, b2 = sorted([b1, b2])
    if b1 > a2 or a1 > b2:
        return 0
    _, inleft, inright, _ = sorted([a1, a2, b1, b2])
    return (inright - inleft + 1) / float(min(a2 - a1 + 1, b2 - b1 + 1))


def _contig_records(batch, params, c):
    """(retained locus rows, [(taxon index, retained locus k, python slice start, stop)]) of contig c: the hits that
    attach_hits (OS:359-369) scores and the site slices score_hit (OS:371-382) raises above zero (start == stop for a hit
    that creates the site array but leaves it at zero)."""
    l0, l1 = int(batch.locus_off[c]), int(batch.locus_off[c + 1])
    kept = [j for j in range(l0, l1)
            if abs(int(batch.locus_end[j]) - int(batch.locus_start[j])) + 1 >= params.min_gene_length]
    recs = []
    for h in range(int(batch.hit_off[c]), int(batch.hit_off[c + 1])):
        if not batch.hit_scov[h] >= params.min_scov:
            continue
        for k, j in enumerate(kept):
            if params.stranded and int(batch.hit_strand[h]) != int(batch.locus_strand[j]):
                continue
            qs, qe = int(batch.hit_qstart[h]), int(batch.hit_qend[h])
            ls, le = int(batch.locus_start[j]), int(batch.locus_end[j])
            if not _calc_overlap(qs, qe, ls, le) >= params.min_overlap:
                continue
            lo, n = min(ls, le), abs(le - ls) + 1
            h1 = max(0, min(qs, qe) - lo)
            h2 = min(n - 1, max(qs, qe) - lo)
            a, b, _ = slice(h1, h2 + 1).indices(n)   # python slice semantics incl. the wrap of a negative stop
            if not (b > a and batch.hit_score[h] > 0.0):
                b = a                                # the site array exists but this hit leaves it at zero
            recs.append((int(batch.hit_taxon[h]), k, a, b))
    return kept, recs


def _spans_field(runs):
    """make_gene_spans_field (OS:773-791) of one site array given its non-zero runs [a, b) (0-based): the 1-based first and
    last site of every run of at least two sites."""
    out = []
    for a, b in runs:
        if b - a >= 2:
            out += [a + 1, b]
    return C_DELIM3.join(str(k) for k in out)


def write_details_file(path, batch, taxonomy, params, res, det, contig_order):
    """<basename>.details.tsv.gz (OS:931-937, 802-812): per contig (blastout order), evaluated level and clade (ascending
    name -- the reference walks a set) the clade's gene scores (3 decimals) and the non-zero spans of its site arrays.
    `det`: the engine's dump (Engine.score_batch_details).  Iterations are numbered like upstream: 1, 1, 2, 3, ...
    (OS:566-579 writes the level after the first raise before incrementing)."""
    import gzip
    names = taxonomy.names
    parent = taxonomy.tables()["parent"]
    order = np.lexsort((det["locus"], det["clade"], det["iteration"], det["contig"]))
    dc, di, dl, dk, ds = (det[k][order] for k in ("contig", "iteration", "clade", "locus", "score"))
    starts = np.flatnonzero(np.r_[True, dc[1:] != dc[:-1]]) if len(dc) else np.zeros(0, np.int64)
    seg = {int(dc[s]): (int(s), int(e)) for s, e in zip(starts, np.r_[starts[1:], len(dc)])}
    jump = int(params.jump_taxonomy)
    with gzip.open(path, "wt") as fh:
        write_row([k.upper() for k in DETAILS_FORMAT], fh)
        for c in contig_order:
            if c not in seg:
                continue   # no hits, or every locus ignored: the contig is never evaluated (OS:959)
            s, e = seg[c]
            kept, recs = _contig_records(batch, params, c)
            l0 = int(batch.locus_off[c])
            slot = {j - l0: k for k, j in enumerate(kept)}
            G = len(kept)
            # clade of every record at the current level
            cur = [t for t, _, _, _ in recs]
            for _ in range(jump):
                cur = [int(parent[t]) for t in cur]
            it = 0
            p = s
            while p < e:
                q = p
                while q < e and di[q] == di[p]:
                    q += 1
                level = int(di[p])
                while it < level:   # raise_taxonomy (OS:431-445)
                    cur = [int(parent[t]) for t in cur]
                    it += 1
                rows, spans = {}, {}
                for x in range(p, q):
                    rows.setdefault(int(dl[x]), [0.0] * G)[slot[int(dk[x])]] = float(ds[x])
                for (t0, k, a, b), t in zip(recs, cur):
                    spans.setdefault(t, {}).setdefault(k, []).append((a, b))
                for t in sorted(rows):
                    cells = []
                    for k in range(G):
                        if k not in spans.get(t, {}):
                            cells.append(C_MISSING_ANNOTATION)   # no site array for this clade / locus (OS:776-781)
                            continue
                        iv = sorted(x for x in spans[t][k] if x[1] > x[0])
                        if not iv:
                            cells.append("")
                            continue
                        runs, (ca, cb) = [], iv[0]
                        for a, b in iv[1:]:
                            if a <= cb:
                                cb = max(cb, b)
                            else:
                                runs.append((ca, cb))
                                ca, cb = a, b
                        runs.append((ca, cb))
                        cells.append(_spans_field(runs))
                    write_row([format_field(v) for v in (
                        batch.contig_names[c], max(1, level), names[t],
                        "|".join("{:.{p}f}".format(v, p=C_PRECISION) for v in rows[t]), "|".join(cells))], fh)
                p = q
