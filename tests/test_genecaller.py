"""waafle_genecaller: the CPU oracle against the reference (committed fixture made from the unmodified
reference functions by oracle/validate_genecaller.py, and the demo GFF the reference ships), and -- on the GPU -- the CUDA
gene caller through the C ABI against the oracle, plus the drop-in CLI's GFF bytes."""
import gzip
import json
import os

import numpy as np
import pytest

import helpers
from oracle import genecaller_oracle as oracle

DEMO = os.path.join(helpers.GOLDEN, "demo")


def fixture_cases():
    with gzip.open(os.path.join(helpers.GOLDEN, "genecaller_cases.json.gz"), "rt") as fh:
        return json.load(fh)


def demo_hits():
    from waafle_b200 import parsers
    return parsers.read_blast_hits(os.path.join(DEMO, "demo_contigs.blastout"))


def oracle_gff_rows(hits, min_overlap=0.1, min_len=200.0, min_scov=0.75):
    from waafle_b200 import genecaller
    off, names = genecaller.blocks_of(hits)
    keep = np.asarray(hits.scov_modified) >= min_scov
    genes = oracle.call_genes_blocks(off, hits.qstart, hits.qend, hits.strand, keep, min_overlap, min_len)
    return [[nm, "waafle_genecaller", "gene", str(s), str(e), ".", st, "0", "."] for nm, gl in zip(names, genes) for s, e, st in gl]


def test_oracle_matches_reference_fixture():
    cases = fixture_cases()
    assert len(cases) >= 300
    for c in cases:
        got = oracle.call_genes([tuple(t) for t in c["intervals"]], c["min_overlap"], c["min_gene_length"])
        assert got == [tuple(g) for g in c["genes"]]


def test_oracle_reproduces_the_shipped_demo_gff():
    rows = oracle_gff_rows(demo_hits())
    with open(os.path.join(DEMO, "demo_contigs.gff"), newline="") as fh:
        want = [ln.rstrip("\r\n").split("\t") for ln in fh if ln.strip()]
    assert rows == want


def run_device(intervals_per_block, thr, min_len):
    from waafle_b200 import genecaller
    class H:   # the columns call_genes reads
        pass
    h = H()
    flat = [t for blk in intervals_per_block for t in blk]
    h.qstart = np.array([t[0] for t in flat], dtype=np.int32)
    h.qend = np.array([t[1] for t in flat], dtype=np.int32)
    h.strand = np.array([ord(t[2]) for t in flat], dtype=np.int8)
    h.scov_modified = np.ones(len(flat))
    h.block_starts = np.cumsum([0] + [len(b) for b in intervals_per_block])[:-1]
    h.block_names = ["c{}".format(k) for k in range(len(intervals_per_block))]
    h.__class__.__len__ = lambda self: len(flat)
    names, goff, gs, ge, gst, ms = genecaller.call_genes(h, thr, min_len, 0.75, 0)
    return [[(int(gs[g]), int(ge[g]), chr(gst[g])) for g in range(goff[b], goff[b + 1])] for b in range(len(names))]


@pytest.mark.gpu
def test_device_gene_calls_equal_reference_fixture():
    cases = fixture_cases()
    by_param = {}
    for c in cases:
        by_param.setdefault((c["min_overlap"], c["min_gene_length"]), []).append(c)
    for (thr, ml), cs in by_param.items():
        got = run_device([[tuple(t) for t in c["intervals"]] for c in cs], thr, ml)
        for c, g in zip(cs, got):
            assert g == [tuple(x) for x in c["genes"]], (thr, ml, c["intervals"][:4])


@pytest.mark.gpu
def test_cli_writes_the_shipped_demo_gff(tmp_path):
    from waafle_b200 import genecaller
    for extra in ([], ["--cpu-parse"]):
        out = tmp_path / "demo{}.gff".format(len(extra))
        genecaller.main([os.path.join(DEMO, "demo_contigs.blastout"), "--gff", str(out)] + extra)
        assert out.read_bytes() == open(os.path.join(DEMO, "demo_contigs.gff"), "rb").read()


@pytest.mark.gpu
def test_device_gene_calls_on_a_large_synthetic_blastout(tmp_path):
    """~500k hits / 2000 contigs: every contig block against the oracle, default and non-default thresholds."""
    from waafle_b200 import genecaller, parsers, synth
    data = synth.generate_config("cfg2", n_contigs=2000, seed=17)
    files = data.write_files(str(tmp_path), "g")
    hits = parsers.read_blast_hits(files["blastout"], device=0)
    for thr, ml, scov in ((0.1, 200.0, 0.75), (0.5, 0.0, 0.0), (0.0, 200.0, 0.9)):
        names, goff, gs, ge, gst, ms = genecaller.call_genes(hits, thr, ml, scov, 0)
        off, _ = genecaller.blocks_of(hits)
        keep = np.asarray(hits.scov_modified) >= scov
        want = oracle.call_genes_blocks(off, hits.qstart, hits.qend, hits.strand, keep, thr, ml)
        got = [[(int(gs[g]), int(ge[g]), chr(gst[g])) for g in range(goff[b], goff[b + 1])] for b in range(len(names))]
        assert got == want, (thr, ml, scov)
        print("genecaller kernel: {} hits, {} genes, {:.2f} ms".format(len(hits), goff[-1], ms))
