"""
TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference classes in-process.

Imports `waafle.utils` / `waafle.waafle_orgscorer` from /root/reference (read-only, present
only in the build container, never on the GPU box) and replays the body of the reference's
`main()` (waafle/waafle_orgscorer.py:900-960) on text inputs, returning one plain record
per contig.  Used to (a) pin oracle/orgscorer_oracle.py and (b) generate tests/golden/.

Determinism: the reference iterates a Python set of clade names (OS:587, OS:603-607), so
exact rank ties depend on PYTHONHASHSEED.  `CanonContig` only changes the *iteration order*
of `Contig.clades` to ascending name; every line of reference logic runs as is.
"""

import argparse
import os
import sys

REFERENCE_ROOT = os.environ.get("WAAFLE_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "waafle"))


def _import():
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    argv, sys.argv = sys.argv, ["waafle_orgscorer"]   # `describe` formats sys.argv[0]
    try:
        from waafle import utils as wu
        from waafle import waafle_orgscorer as wo
    finally:
        sys.argv = argv
    return wu, wo


class _OrderedClades:
    def __init__(self, it=()):
        self._s = set(it)

    def add(self, x):
        self._s.add(x)

    def __len__(self):
        return len(self._s)

    def __contains__(self, x):
        return x in self._s

    def __iter__(self):
        return iter(sorted(self._s))


def make_args(**over):
    """Namespace with the reference CLI defaults (OS:188-296, GC:83-101)."""
    d = dict(one_clade_threshold=0.5, two_clade_threshold=0.8, disambiguate_one="meld",
             disambiguate_two="meld", range=0.05, jump_taxonomy=None, allow_lca=False,
             ambiguous_fraction=0.1, ambiguous_threshold="lenient", sister_penalty="strict",
             clade_genes=None, clade_leaves=None, weak_loci="ignore",
             annotation_threshold="lenient", min_overlap=0.1, min_gene_length=200.0,
             min_scov=0.75, stranded=False, quiet=True, write_details=False)
    d.update(over)
    return argparse.Namespace(**d)


def run_reference(contigs, blastout, gff, taxonomy, args, canonical=True):
    """Replay OS:900-960; returns {contig_name: record} in FASTA order."""
    wu, wo = _import()

    if canonical:
        class CanonContig(wo.Contig):
            @property
            def clades(self):
                return self._clades

            @clades.setter
            def clades(self, v):
                self._clades = _OrderedClades(v)
        Contig = CanonContig
    else:
        Contig = wo.Contig

    tax = wu.Taxonomy(taxonomy)
    cs = {}
    for name, length in wu.read_contig_lengths(contigs).items():
        C = Contig(name, args)
        C.length = length
        cs[name] = C
    for name, loci in wu.iter_contig_loci(gff, attach_annotations=False):
        if name in cs:
            cs[name].attach_loci(loci)
    level0 = {}
    for name, hits in wu.iter_contig_hits(blastout):
        if name not in cs:
            continue
        C = cs[name]
        C.attach_hits(hits)
        C.update_gene_scores()
        level0[name] = {k: [float(x) for x in v] for k, v in C.gene_scores.items()}
        if args.jump_taxonomy is not None:
            for _ in range(args.jump_taxonomy):
                C.raise_taxonomy(tax)
        if not all([L.ignore for L in C.loci]):
            wo.evaluate_contig(C, tax, None, args)
    out = {}
    for name, C in cs.items():
        b1, b2 = C.best_one, C.best_two
        rec = dict(length=C.length, loci=[L.code for L in C.loci],
                   ignore=[bool(L.ignore) for L in C.loci],
                   annotations=[dict(L.annotations) for L in C.loci],
                   gene_scores0=level0.get(name, {}))
        if wo.is_ok(b1):
            rec.update(call="no_lgt", crit=float(b1.crit), rank=float(b1.rank),
                       synteny=b1.synteny, clade1=b1.clade1,
                       melded1=wo.make_tails_field(b1.tails1),
                       taxonomy1=tax.get_lineage(b1.clade1))
        elif wo.is_ok(b2):
            rec.update(call="lgt", crit=float(b2.crit), rank=float(b2.rank),
                       synteny=b2.synteny, direction=b2.direction, clade1=b2.clade1,
                       clade2=b2.clade2, lca=tax.get_lca(b2.clade1, b2.clade2),
                       melded1=wo.make_tails_field(b2.tails1),
                       melded2=wo.make_tails_field(b2.tails2))
        else:
            rec.update(call="unclassified")
        out[name] = rec
    return out
