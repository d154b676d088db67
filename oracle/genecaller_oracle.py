"""CPU oracle of `waafle_genecaller` (TEST INFRASTRUCTURE ONLY -- never imported by waafle_b200/).

Plain-Python restatement of the reference's gene calling from BLAST hits, waafle/waafle_genecaller.py:107-170
(hits2ints, overlap_intervals, merge_inodes) with waafle/utils.py:455-500 (INode, calc_overlap):

  * hits with scov_modified >= --min-scov become intervals [min(qstart, qend), max(..)] with the hit's strand (:107-113);
  * sorted by start (stable); every later interval that overlaps an earlier one by >= --min-overlap of the SHORTER of the
    two is linked to it; the scan of later intervals stops at the first one that does not overlap at all (:146-157);
  * connected components are merged: [min start, max stop], strand of the longest member (ties: '-' beats '+', the last
    of sorted([len, strand])) (:121-135); components come out in the order of their first member (:158-162);
  * genes shorter than --min-gene-length are dropped (:218-219).
`--stranded` never takes effect upstream (main compares the store_true flag with the string "on", :215); the oracle
keeps that.  Pinned against the unmodified reference functions and the reference's own demo GFF by
oracle/validate_genecaller.py and tests/test_genecaller.py.
"""


def calc_overlap(a1, a2, b1, b2):
    """waafle/utils.py:487-500 (normalize=True)."""
    a1, a2 = sorted([a1, a2])
    b1, b2 = sorted([b1, b2])
    if b1 > a2 or a1 > b2:
        return 0
    _, inleft, inright, _ = sorted([a1, a2, b1, b2])
    return (inright - inleft + 1) / float(min(a2 - a1 + 1, b2 - b1 + 1))


def call_genes(intervals, min_overlap=0.1, min_gene_length=200.0):
    """`intervals`: [(qstart, qend, strand)] of ONE contig's hits that passed the scov filter, in file order.
    Returns [(start, stop, strand)] like overlap_intervals (:137-168) followed by the length filter (:217-219)."""
    nodes = sorted(((min(s, e), max(s, e), st) for s, e, st in intervals), key=lambda t: t[0])
    n = len(nodes)
    nbr = [set() for _ in range(n)]
    for i in range(n):
        for j in range(i + 1, n):
            score = calc_overlap(nodes[i][0], nodes[i][1], nodes[j][0], nodes[j][1])
            if score >= min_overlap:
                nbr[i].add(j)
                nbr[j].add(i)
            elif score == 0:
                break
    seen, genes = [False] * n, []
    for i in range(n):
        if seen[i]:
            continue
        comp, front = {i}, [i]
        seen[i] = True
        while front:
            nxt = []
            for u in front:
                for v in nbr[u]:
                    if not seen[v]:
                        seen[v] = True
                        comp.add(v)
                        nxt.append(v)
            front = nxt
        start = min(nodes[k][0] for k in comp)
        stop = max(nodes[k][1] for k in comp)
        strand = sorted([nodes[k][1] - nodes[k][0] + 1, nodes[k][2]] for k in comp)[-1][1]
        if stop - start + 1 >= min_gene_length:
            genes.append((start, stop, strand))
    return genes


def call_genes_blocks(block_off, qstart, qend, strand, keep, min_overlap=0.1, min_gene_length=200.0):
    """Batch form over contig blocks of a blastout (iter_contig_hits, waafle/utils.py:255-270): per block the gene list."""
    out = []
    for b in range(len(block_off) - 1):
        iv = [(int(qstart[h]), int(qend[h]), chr(strand[h])) for h in range(block_off[b], block_off[b + 1]) if keep[h]]
        out.append(call_genes(iv, min_overlap, min_gene_length))
    return out
