"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the waafle_orgscorer per-contig engine.

This file is a numpy restatement of the reference algorithm (file:line cites are
relative to /root/reference/waafle/, OS = waafle_orgscorer.py, UT = utils.py).
Nothing in the product package (waafle_b200/) may import it; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
and there only as the checker / the timed CPU baseline.

Parity pin: `oracle/validate_against_reference.py` runs the *unmodified* reference
classes (imported from /root/reference, canonical clade order, see below) and this
restatement on the same inputs and demands equality of every output field
(bit-exact floats).  The reference's own golden vectors (demo/output*/ TSVs) are
checked through the front end in tests/test_demo_golden.py.

Arithmetic: the reference uses numpy float64 `np.mean`/`np.min`/`np.maximum` on
per-site arrays (OS:381-382,403,441-442,456-460).  This oracle calls the very same
numpy functions on the very same per-site arrays, so numpy's pairwise summation
order is reproduced by construction (numpy is the pinned third-party dependency:
setup.py requires numpy>=1.13.0; the oracle of record is numpy 2.3.x).

Determinism contract: the reference iterates a Python `set`
of clade names (OS:587,603-607) so exact rank ties are hash-seed dependent.  The
oracle, the validation harness and the CUDA engine all use the canonical order
"ascending clade name (Python str order)" == ascending node index.

Input / output layout is the packed CSR/SoA layout of include/waafle_b200.h so the
same arrays feed the oracle and the engine.
"""

import re

import numpy as np

C_EPS = 1e-6  # OS:58

CALL_UNCLASSIFIED, CALL_NO_LGT, CALL_LGT = 0, 1, 2
FLAG_RETAINED, FLAG_IGNORED = 1, 2

DISAMBIG_ONE = {"report-best": 0, "meld": 1}
DISAMBIG_TWO = {"report-best": 0, "jump": 1, "meld": 2}
WEAK_LOCI = {"ignore": 0, "penalize": 1, "assign-unknown": 2}
TRI = {"off": 0, "lenient": 1, "strict": 2}


def default_params(**over):
    """Reference CLI defaults (OS:188-296, waafle_genecaller.py:83-101)."""
    p = dict(
        k1=0.5, k2=0.8, range=0.05, ambiguous_fraction=0.1, min_overlap=0.1,
        min_scov=0.75, min_gene_length=200.0,
        disambiguate_one=1, disambiguate_two=2, weak_loci=0,
        ambiguous_threshold=1, sister_penalty=2, annotation_threshold=1,
        allow_lca=0, stranded=0, jump_taxonomy=0, clade_genes=-1, clade_leaves=-1,
        n_systems=0,
    )
    p.update(over)
    return p


def _tri_threshold(mode, k1, k2):
    """off -> c_eps, lenient -> min(k1,k2), strict -> max(k1,k2) (OS:341-346, OS:515-516)."""
    return (C_EPS, min(k1, k2), max(k1, k2))[mode]


class _Tax:
    """Integer-index view of UT.Taxonomy (UT:374-447)."""

    def __init__(self, tax):
        self.parent = np.asarray(tax["parent"], dtype=np.int64)
        self.leaf_count = np.asarray(tax["leaf_count"], dtype=np.int64)
        self.listed = np.asarray(tax["listed"], dtype=np.uint8)
        self.root = int(tax["root_idx"])
        self.unknown = int(tax["unknown_idx"])

    def get_parent(self, c):  # UT:386-387 (missing -> root is baked into the table)
        return int(self.parent[c])

    def get_lineage(self, c):  # UT:392-399
        l = [c]
        while l[-1] != self.root:
            l.append(self.get_parent(l[-1]))
        l.reverse()
        return l

    def get_lca(self, *clades):  # UT:401-411
        lca = self.root
        lineages = [self.get_lineage(c) for c in clades]
        min_depth = min(len(l) for l in lineages)
        for i in range(min_depth):
            level = {l[i] for l in lineages}
            if len(level) == 1:
                lca = list(level)[0]
            else:
                break
        return lca


def _calc_overlap(a1, a2, b1, b2):
    """UT:487-500 (normalize=True)."""
    a1, a2 = sorted([a1, a2])
    b1, b2 = sorted([b1, b2])
    if b1 > a2 or a1 > b2:
        return 0
    _, inleft, inright, _ = sorted([a1, a2, b1, b2])
    overlap = inright - inleft + 1
    denom = min(a2 - a1 + 1, b2 - b1 + 1)
    return overlap / float(denom)


class _Option:
    """OS:467-493."""

    def __init__(self):
        self.ok = True
        self.crit = None
        self.rank = None
        self.clade1 = None
        self.clade2 = None
        self.best1 = None
        self.best2 = None
        self.synteny = None
        self.direction = 0
        self.recip = None
        self.members1 = []
        self.members2 = []


class _Contig:
    """One contig's state; methods follow OS:309-461 one-to-one."""

    def __init__(self, P, T, lstart, lend, lstrand):
        self.P, self.T = P, T
        self.k1, self.k2 = P["k1"], P["k2"]
        self.min_threshold = min(self.k1, self.k2)  # OS:338
        self.max_threshold = max(self.k1, self.k2)  # OS:339
        self.annotation_threshold = _tri_threshold(P["annotation_threshold"], self.k1, self.k2)
        # attach_loci OS:348-352: retained loci in GFF order
        self.raw_idx = [j for j in range(len(lstart))
                        if abs(int(lend[j]) - int(lstart[j])) + 1 >= P["min_gene_length"]]
        self.lstart = [int(lstart[j]) for j in self.raw_idx]
        self.lend = [int(lend[j]) for j in self.raw_idx]
        self.lstrand = [int(lstrand[j]) for j in self.raw_idx]
        self.llen = [abs(e - s) + 1 for s, e in zip(self.lstart, self.lend)]  # UT:321-322
        self.G = len(self.raw_idx)
        self.ignore = [False] * self.G  # UT:319
        self.mask = None
        self.site_scores = {}
        self.gene_scores = {}
        self.clades = []
        S = P["n_systems"]
        self.ann_score = [[self.annotation_threshold] * S for _ in range(self.G)]
        self.ann_winner = [[-1] * S for _ in range(self.G)]

    # OS:359-369
    def attach_hits(self, hq1, hq2, htax, hscore, hscov, hstrand, hsys, base):
        P = self.P
        for h in range(len(hq1)):
            if hscov[h] >= P["min_scov"]:
                for i in range(self.G):
                    if P["stranded"] and int(hstrand[h]) != self.lstrand[i]:
                        continue
                    ov = _calc_overlap(int(hq1[h]), int(hq2[h]), self.lstart[i], self.lend[i])
                    if ov >= P["min_overlap"]:
                        self.score_hit(i, int(hq1[h]), int(hq2[h]), int(htax[h]),
                                       float(hscore[h]), int(hsys[h]) if hsys is not None else 0,
                                       base + h)
        self.clades = sorted(self.site_scores)

    # OS:371-392
    def score_hit(self, i, q1, q2, taxon, score, sysmask, hit_index):
        l1, l2 = sorted([self.lstart[i], self.lend[i]])
        h1, h2 = sorted([q1, q2])
        h1 = max(0, h1 - l1)
        h2 = min(self.llen[i] - 1, h2 - l1)
        ldict = self.site_scores.setdefault(taxon, {})
        if i not in ldict:
            ldict[i] = np.zeros(self.llen[i])
        ldict[i][h1:h2 + 1] = np.maximum(ldict[i][h1:h2 + 1], score)
        for s in range(self.P["n_systems"]):
            if sysmask >> s & 1:
                if score >= self.ann_score[i][s]:
                    self.ann_winner[i][s] = hit_index
                    self.ann_score[i][s] = score

    # OS:394-429
    def update_gene_scores(self):
        self.gene_scores = {}
        for clade in sorted(self.site_scores):
            ldict = self.site_scores[clade]
            scores = []
            for i in range(self.G):
                if i in ldict:
                    scores.append(np.mean(ldict[i]))
                else:
                    scores.append(0)
            self.gene_scores[clade] = np.array(scores, dtype=np.float64)
        maxes = np.zeros(self.G)
        for clade, values in self.gene_scores.items():
            if clade != self.T.unknown:
                maxes = np.maximum(maxes, values)
        wl = self.P["weak_loci"]
        if wl == 1:
            pass
        elif wl == 2:
            self.gene_scores[self.T.unknown] = 1 - maxes
        elif wl == 0:
            ok_list = []
            for index, value in enumerate(maxes):
                self.ignore[index] = True
                if value >= self.min_threshold:
                    ok_list.append(index)
                    self.ignore[index] = False
            self.mask = None if len(ok_list) == self.G else np.array(ok_list, dtype=np.int64)
        self.clades = sorted(self.gene_scores)

    # OS:431-445
    def raise_taxonomy(self):
        new_site_scores = {}
        for clade in sorted(self.site_scores):
            ldict = self.site_scores[clade]
            parent = self.T.get_parent(clade)
            inner = new_site_scores.setdefault(parent, {})
            for i in ldict:
                if i not in inner:
                    inner[i] = np.zeros(self.llen[i])
                inner[i] = np.maximum(inner[i], ldict[i])
        self.site_scores = new_site_scores
        self.update_gene_scores()

    # OS:447-461
    def score(self, clade1, clade2=None):
        maxes = self.gene_scores[clade1]
        if clade2 is not None:
            maxes = np.maximum(maxes, self.gene_scores[clade2])
        maxes = maxes if self.mask is None else maxes[self.mask]
        return np.min(maxes), np.mean(maxes)


def _set_synteny_one(opt, C):  # OS:495-509
    scores = C.gene_scores[opt.clade1]
    syn = ""
    for s, ig in zip(scores, C.ignore):
        if ig:
            syn += "~"
        elif s >= C.k1:
            syn += "A"
        else:
            syn += "!"
    opt.synteny = syn


def _set_synteny_two(opt, C):  # OS:511-545
    k_amb = _tri_threshold(C.P["ambiguous_threshold"], C.k1, C.k2)
    s1s, s2s = C.gene_scores[opt.clade1], C.gene_scores[opt.clade2]
    unknown_involved = C.T.unknown in (opt.clade1, opt.clade2)
    syn = ""
    for s1, s2, ig in zip(s1s, s2s, C.ignore):
        if ig:
            syn += "~"
        elif min(s1, s2) >= k_amb and not unknown_involved:
            syn += "*"
        elif s1 >= C.k2:
            syn += "A"
        elif s2 >= C.k2:
            syn += "B"
        else:
            syn += "!"
    if re.search("^[^A]*B", syn):
        opt.clade1, opt.clade2 = opt.clade2, opt.clade1
        syn = syn.translate(str.maketrans("AB", "BA"))
    opt.synteny = syn
    if re.search("^A+B+A+$", syn.replace("~", "")):
        opt.direction = 1
        opt.recip = opt.clade2  # OS:544-545 (donor=clade1, recip=clade2, as coded)


def _explain_one(C):  # OS:585-597 + meld_one OS:621-631
    P = C.P
    options = []
    for clade in C.clades:
        crit, rank = C.score(clade)
        if crit >= C.k1:
            o = _Option()
            o.crit, o.rank, o.clade1 = crit, rank, clade
            _set_synteny_one(o, C)
            options.append(o)
    if not options:
        return None
    options = sorted(options, key=lambda x: x.rank)
    best = options[-1]
    options = [k for k in options if best.rank - k.rank <= P["range"]]
    best.best1 = best.clade1
    if P["disambiguate_one"] == 1:
        to_meld = [k.clade1 for k in options]
        best.clade1 = C.T.get_lca(*to_meld)
        best.members1 = sorted(set(to_meld))
    return best


def _check_lgt(o, C):  # OS:678-744
    P, T = C.P, C.T
    # OS:693-702
    total_len = amb_len = 0
    for ch, ln in zip(o.synteny, C.llen):
        if ch in "AB*":
            total_len += ln
            amb_len += ln if ch == "*" else 0
    if amb_len / float(total_len) > P["ambiguous_fraction"]:
        o.ok = False
    # OS:704-708
    if P["clade_genes"] >= 0:
        if min(o.synteny.count("A"), o.synteny.count("B")) < P["clade_genes"]:
            o.ok = False
    # OS:710-715
    if P["clade_leaves"] >= 0:
        to_check = [o.recip] if o.recip is not None else [o.clade1, o.clade2]
        if min(int(T.leaf_count[c]) for c in to_check) < P["clade_leaves"]:
            o.ok = False
    # OS:717-744
    if P["sister_penalty"] != 0:
        thr = C.max_threshold if P["sister_penalty"] == 1 else C.min_threshold

        def sisters(c, other):  # UT:428-434 minus the partner clade
            p = T.get_parent(c)
            return [x for x in C.gene_scores
                    if x != c and x != other and T.listed[x] and T.get_parent(x) == p]

        sis = {"B": sisters(o.clade1, o.clade2), "A": sisters(o.clade2, o.clade1)}
        bad = {"A": False, "B": False}
        for i, ch in enumerate(o.synteny):
            if ch in sis:
                for c in sis[ch]:
                    if C.gene_scores[c][i] >= thr:
                        bad[ch] = True
        to_check = "B" if o.recip is not None else "AB"
        if any(bad[ch] for ch in to_check):
            o.ok = False


def _explain_two(C):  # OS:599-619 + meld_two OS:633-676
    P, T = C.P, C.T
    potential = [c for c in C.clades if max(C.gene_scores[c]) >= C.k2]
    options = []
    for c1 in potential:
        for c2 in potential:
            if c1 < c2:
                crit, rank = C.score(c1, c2)
                if crit >= C.k2:
                    o = _Option()
                    o.rank, o.crit, o.clade1, o.clade2 = rank, crit, c1, c2
                    _set_synteny_two(o, C)
                    options.append(o)
    if not options:
        return None
    options = sorted(options, key=lambda o: o.rank)
    best = options[-1]
    options = [k for k in options if best.rank - k.rank <= P["range"]]
    for o in options:
        _check_lgt(o, C)
    best.best1, best.best2 = best.clade1, best.clade2
    if len(options) == 1:
        pass
    elif P["disambiguate_two"] == 0:
        pass
    elif P["disambiguate_two"] == 1:
        best = None
    else:
        if not all(o.ok and o.synteny == options[0].synteny for o in options):
            best = None
        else:
            c1s = [o.clade1 for o in options]
            c2s = [o.clade2 for o in options]
            best.clade1 = T.get_lca(*c1s)
            best.clade2 = T.get_lca(*c2s)
            best.members1 = sorted(set(c1s))
            best.members2 = sorted(set(c2s))
            if not P["allow_lca"]:
                if T.get_lca(best.clade1, best.clade2) in (best.clade1, best.clade2):
                    best = None
    return best


def _is_ok(o):
    return o is not None and o.ok


def score_batch(params, tax, batch, want_gene_scores=False):
    """Score and classify every contig of a packed batch.

    params: dict (see default_params); tax: dict(parent, leaf_count, listed, root_idx,
    unknown_idx); batch: dict of the SoA/CSR arrays of include/waafle_b200.h.
    Returns a dict of arrays with the engine's output layout.
    """
    P = dict(params)
    T = _Tax(tax)
    n = len(batch["hit_off"]) - 1
    S = P["n_systems"]
    nL = int(batch["locus_off"][-1])
    out = dict(
        call=np.zeros(n, np.uint8), direction=np.zeros(n, np.uint8),
        lifts=np.zeros(n, np.int32),
        clade1=np.full(n, -1, np.int32), clade2=np.full(n, -1, np.int32),
        lca=np.full(n, -1, np.int32), best1=np.full(n, -1, np.int32),
        best2=np.full(n, -1, np.int32),
        crit=np.zeros(n, np.float64), rank=np.zeros(n, np.float64),
        synteny=np.zeros(nL, np.uint8), locus_flags=np.zeros(nL, np.uint8),
        ann_winner=np.full((nL, max(S, 1)), -1, np.int32)[:, :S],
        member_off=np.zeros(n + 1, np.int64), n_members_a=np.zeros(n, np.int32),
    )
    members = []
    gene_scores = [] if want_gene_scores else None
    hsys = batch.get("hit_sysmask")
    for c in range(n):
        h0, h1 = int(batch["hit_off"][c]), int(batch["hit_off"][c + 1])
        l0, l1 = int(batch["locus_off"][c]), int(batch["locus_off"][c + 1])
        C = _Contig(P, T, batch["locus_start"][l0:l1], batch["locus_end"][l0:l1],
                    batch["locus_strand"][l0:l1])
        best_one = best_two = None
        lifts = 0
        if h1 > h0:  # contigs absent from the blastout are never visited (OS:943-960)
            C.attach_hits(batch["hit_qstart"][h0:h1], batch["hit_qend"][h0:h1],
                          batch["hit_taxon"][h0:h1], batch["hit_score"][h0:h1],
                          batch["hit_scov"][h0:h1], batch["hit_strand"][h0:h1],
                          hsys[h0:h1] if (hsys is not None and S > 0) else None, h0)
            C.update_gene_scores()
            if want_gene_scores:
                gene_scores.append({k: v.copy() for k, v in C.gene_scores.items()})
            for _ in range(P["jump_taxonomy"]):  # OS:955-957
                C.raise_taxonomy()
                lifts += 1
            if not all(C.ignore):  # OS:959
                # evaluate_contig OS:566-583
                best_one = _explain_one(C)
                best_two = _explain_two(C) if not _is_ok(best_one) else None
                it = 1
                while (len(C.clades) > 0 and T.root not in C.clades
                       and not _is_ok(best_one) and not _is_ok(best_two)):
                    C.raise_taxonomy()
                    lifts += 1
                    best_one = _explain_one(C)
                    best_two = _explain_two(C) if not _is_ok(best_one) else None
                    it += 1
                    if it > 100:
                        raise RuntimeError("Runaway taxonomic recursion")
        elif want_gene_scores:
            gene_scores.append({})
        out["lifts"][c] = lifts
        for gi, j in enumerate(C.raw_idx):
            out["locus_flags"][l0 + j] = FLAG_RETAINED | (FLAG_IGNORED if C.ignore[gi] else 0)
            for s in range(S):
                out["ann_winner"][l0 + j, s] = C.ann_winner[gi][s]
        best = None
        if _is_ok(best_one):  # OS:853
            out["call"][c] = CALL_NO_LGT
            best = best_one
        elif _is_ok(best_two):  # OS:870
            out["call"][c] = CALL_LGT
            best = best_two
            out["clade2"][c] = best.clade2
            out["best2"][c] = best.best2
            out["lca"][c] = T.get_lca(best.clade1, best.clade2)  # OS:882
            out["direction"][c] = best.direction
        if best is not None:
            out["clade1"][c] = best.clade1
            out["best1"][c] = best.best1
            out["crit"][c] = best.crit
            out["rank"][c] = best.rank
            for gi, j in enumerate(C.raw_idx):
                out["synteny"][l0 + j] = ord(best.synteny[gi])
            members += best.members1 + best.members2
            out["n_members_a"][c] = len(best.members1)
        out["member_off"][c + 1] = len(members)
    out["members"] = np.array(members, dtype=np.int32)
    # K10-equivalent compaction: contig indices grouped lgt / no_lgt / unclassified
    order = [np.nonzero(out["call"] == k)[0] for k in (CALL_LGT, CALL_NO_LGT, CALL_UNCLASSIFIED)]
    out["call_counts"] = np.array([len(o) for o in order], dtype=np.int64)
    out["call_index"] = np.concatenate(order).astype(np.int64)
    if want_gene_scores:
        out["gene_scores"] = gene_scores
    return out
