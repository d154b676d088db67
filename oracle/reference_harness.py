"""
TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference classes in-process.

Imports `waafle.utils` / `waafle.waafle_orgscorer` from /root/reference (read-only, present
only in the build container) or, on the GPU box, from oracle/_ref/ (the same files, pip-installed
there by oracle/build_ref.py; git-ignored) and replays the body of the reference's `main()`
(waafle/waafle_orgscorer.py:900-960) on text inputs, returning one plain record per contig.
Used to (a) pin oracle/orgscorer_oracle.py, (b) generate tests/golden/ and (c) time the real
reference on the host cores beside the GPU numbers (`time_reference_sharded`, bench.py).

Determinism: the reference iterates a Python set of clade names (OS:587, OS:603-607), so
exact rank ties depend on PYTHONHASHSEED.  `CanonContig` only changes the *iteration order*
of `Contig.clades` to ascending name; every line of reference logic runs as is.
"""

import argparse
import os
import sys
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("WAAFLE_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "waafle")):
    REFERENCE_ROOT = os.path.join(_HERE, "_ref")   # installed copy of the unmodified reference (oracle/build_ref.py)


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


def run_reference(contigs, blastout, gff, taxonomy, args, canonical=True, timers=None, details=None):
    """Replay OS:900-960; returns {contig_name: record} in FASTA order.  `timers` (dict) receives the seconds spent
    inside the engine region OS:952-960 ("engine") and in the whole replay including the parsers ("total").
    `details`: a TEXT file object that receives what --write-details writes (OS:931-937, 802-812; upstream opens its
    gzip in binary mode and dies on Python 3 -- the rows themselves are produced by the unmodified write_details)."""
    wu, wo = _import()
    t_total = time.perf_counter()
    t_engine = 0.0

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
    if details is not None:
        wu.write_rowdict(None, wo.c_formats["details"], file=details)
    for name, hits in wu.iter_contig_hits(blastout):
        if name not in cs:
            continue
        C = cs[name]
        t0 = time.perf_counter()
        C.attach_hits(hits)
        C.update_gene_scores()
        if timers is None:
            level0[name] = {k: [float(x) for x in v] for k, v in C.gene_scores.items()}
        if args.jump_taxonomy is not None:
            for _ in range(args.jump_taxonomy):
                C.raise_taxonomy(tax)
        if not all([L.ignore for L in C.loci]):
            wo.evaluate_contig(C, tax, details, args)
        t_engine += time.perf_counter() - t0
    if timers is not None:
        timers["engine"] = t_engine
        timers["total"] = time.perf_counter() - t_total
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


def _shard_worker(job):
    files, flags = job
    timers = {}
    out = run_reference(files["contigs"], files["blastout"], files["gff"], files["taxonomy"], make_args(**flags),
                        canonical=False, timers=timers)
    calls = [0, 0, 0]
    for r in out.values():
        calls[{"lgt": 0, "no_lgt": 1, "unclassified": 2}[r["call"]]] += 1
    return timers["engine"], timers["total"], calls


def time_reference_sharded(shard_files, flags=None):
    """The UNMODIFIED reference on `shard_files` (one dict of text-file paths per process; contigs are independent and
    the inputs are grouped by contig, UT:255-270, 341-355).  Returns (wall seconds, max engine-only seconds over the
    processes, summed call counts)."""
    import multiprocessing as mp
    jobs = [(f, flags or {}) for f in shard_files]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(jobs)) as pool:
        res = pool.map(_shard_worker, jobs)
    wall = time.perf_counter() - t0
    calls = [sum(r[2][i] for r in res) for i in range(3)]
    return wall, max(r[0] for r in res), max(r[1] for r in res), calls
