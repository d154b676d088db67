"""CPU: host-side mirror of the reference interface (parsers, taxonomy tables, packing, writer)."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers
from waafle_b200 import packing, parsers, taxonomy
from waafle_b200.params import CParams, OrgscorerParams


def test_hit_arithmetic_matches_reference_formulae():
    # utils.py:214-229 re-derived by hand for a plus and a minus hit
    h = parsers.hits_from_columns(
        ["c1", "c1"], ["G1|s__A|UniProt=X1", "G2|s__B"], [5000, 5000], [1347, 1440], [3675, 830],
        [5021, 2236], [1347, 1], [1, 1407], [92.725, 87.207], ["minus", "plus"])
    # minus: sstart' = 1347-1347+1 = 1, send' = 1347; ltrim = max(0, 1-3675)=0; rtrim = max(0,1347-1-5000+3675)=21
    assert h.scov_modified[0] == (1347 - 1 + 1) / float(1347 - 0 - 21)
    assert h.score[0] == h.scov_modified[0] * 92.725 / 100.0
    assert h.scov_modified[1] == (1407 - 1 + 1) / float(1440 - 0 - 0)
    assert list(h.strand) == [ord("-"), ord("+")]
    assert list(h.taxon) == ["s__A", "s__B"]
    assert h.systems == ["UniProt"] and list(h.sysmask) == [1, 0]


def test_demo_parse_and_pack(tmp_path):
    batch, loci, hits, tax = helpers.frontend_load(helpers.demo_files(tmp_path))
    assert batch.n_contigs == 115 and batch.n_hits == 1416 and batch.n_loci == len(loci)
    assert np.all(np.diff(batch.hit_off) >= 0) and batch.hit_off[-1] == 1416
    assert hits.scov_modified.max() > 1.0     # the demo has scov_modified up to 1.04
    # node order == python str order, root/unknown present
    assert tax.names == sorted(tax.names)
    assert tax.names[tax.root_idx] == "r__Root" and tax.names[tax.unknown_idx] == "Unknown"
    assert tax.parent[tax.root_idx] == tax.root_idx and tax.depth[tax.root_idx] == 0


def test_taxonomy_tables():
    t = taxonomy.Taxonomy(edges=[list(e) for e in helpers.TAX8]).build(["s__zz_unlisted"])
    ix = t.index
    assert t.names[t.parent[ix["s__A"]]] == "g__G1"
    assert t.parent[ix["s__zz_unlisted"]] == t.root_idx and not t.listed[ix["s__zz_unlisted"]]
    assert t.depth[ix["s__A"]] == 7 and t.depth[ix["t__A1"]] == 8 and t.depth[ix["Unknown"]] == 1
    # leaf counts (utils.py:436-447)
    assert t.leaf_count[ix["s__A"]] == 2 and t.leaf_count[ix["g__G1"]] == 4
    assert t.leaf_count[ix["k__B"]] == 8 and t.leaf_count[ix["s__zz_unlisted"]] == 1
    assert t.get_lineage(ix["s__A"])[0] == "r__Root" and t.get_lineage(ix["s__A"])[-1] == "s__A"
    assert t.get_tail(ix["s__A"], ix["f__F1"]) == ["g__G1", "s__A"]
    with pytest.raises(SystemExit):
        taxonomy.Taxonomy(edges=[["a", "b"], ["a", "c"]])
    with pytest.raises(SystemExit):
        taxonomy.Taxonomy(edges=[["a", "b"], ["b", "a"]]).build()


def test_params_struct_matches_header():
    """ctypes layout of wfl_params == field order in include/waafle_b200.h."""
    hdr = open(os.path.join(helpers.GOLDEN, "..", "..", "include", "waafle_b200.h")).read()
    body = hdr[hdr.index("typedef struct {", hdr.index("Engine-relevant CLI flags")):hdr.index("} wfl_params;")]
    names = re.findall(r"(?:double|int32_t)\s+(\w+);", body)
    assert names == [f[0] for f in CParams._fields_]
    p = OrgscorerParams(k1=0.3, clade_genes=2).as_ctypes()
    assert p.k1 == 0.3 and p.clade_genes == 2 and p.clade_leaves == -1
    assert ctypes.sizeof(CParams) == 7 * 8 + 12 * 4


def test_pack_slices_and_unknown_contigs(tmp_path, capsys):
    data = helpers.synth_case(dict(config="cfg2", n_contigs=20, seed=5, over={}))
    b = data.to_batch()
    s = b.slice(5, 12)
    assert s.n_contigs == 7 and s.hit_off[0] == 0 and s.n_hits == b.hit_off[12] - b.hit_off[5]
    assert np.array_equal(s.hit_score, b.hit_score[b.hit_off[5]:b.hit_off[12]])
    # rows naming contigs absent from the FASTA are warned about and dropped (OS:921-923, 944-946)
    files = data.write_files(str(tmp_path), "s")
    with open(files["contigs"]) as fh:
        lines = fh.read().split("\n")
    with open(files["contigs"], "w") as fh:
        fh.write("\n".join(lines[2:]))          # drop the first contig
    batch, loci, hits, tax = helpers.frontend_load(files)
    assert batch.n_contigs == 19
    err = capsys.readouterr().err
    assert "Unknown contig in <gff> file" in err and "Unknown contig in <blastout> file" in err


def test_bad_rows_die(tmp_path):
    p = tmp_path / "bad.blastout"
    p.write_text("c1\tG|s__A\t10\n")
    with pytest.raises(SystemExit):
        parsers.read_blast_hits(str(p))
    g = tmp_path / "bad.gff"
    g.write_text("c1\tx\tgene\t1\n")
    with pytest.raises(SystemExit):
        parsers.read_gff_loci(str(g))


def test_arrow_and_pandas_front_ends_agree(tmp_path, monkeypatch):
    """The Arrow reader / string kernels (fast path) and the pandas + per-header Python path (fallback) give
    the same hit table and the same packed batch, for plain and gzip-compressed blastout files."""
    import gzip
    import shutil
    from waafle_b200 import utils
    if parsers.pa is None:
        pytest.skip("pyarrow not importable")
    data = helpers.synth_case(dict(config="cfg5", n_contigs=60, seed=9, over={}))
    files = data.write_files(str(tmp_path), "a")
    gz = dict(files, blastout=files["blastout"] + ".gz")
    with open(files["blastout"], "rb") as a, gzip.open(gz["blastout"], "wb") as b:
        shutil.copyfileobj(a, b)

    def load(f):
        hits = parsers.read_blast_hits(f["blastout"])
        loci = parsers.read_gff_loci(f["gff"])
        tax = taxonomy.Taxonomy(f["taxonomy"]).build(set(hits.taxon))
        return hits, packing.pack(utils.read_contig_lengths(f["contigs"]), loci, hits, tax)

    fast, fast_gz = load(files), load(gz)
    assert fast[0].taxon_codes is not None and fast[0].qseqid_codes is not None
    for name in ("pa", "pc", "pacsv"):
        monkeypatch.setattr(parsers, name, None)
    slow = load(files)
    assert slow[0].taxon_codes is None
    for hits, batch in (fast, fast_gz):
        for k in ("qstart", "qend", "score", "scov_modified", "strand", "sysmask"):
            assert np.array_equal(getattr(hits, k), getattr(slow[0], k)), k
        assert list(hits.taxon) == list(slow[0].taxon) and list(hits.qseqid) == list(slow[0].qseqid)
        assert hits.systems == slow[0].systems
        for i in range(0, len(hits), 37):
            assert hits.sseqid_annotations[int(hits.sseqid_id[i])] == slow[0].sseqid_annotations[int(slow[0].sseqid_id[i])]
            assert hits.sseqid_names[int(hits.sseqid_id[i])] == slow[0].sseqid_names[int(slow[0].sseqid_id[i])]
        for k, v in batch.arrays().items():
            assert np.array_equal(v, slow[1].arrays()[k]), k


def test_bad_subject_header_dies_on_both_paths(tmp_path, monkeypatch):
    row = "c1\tGENE_WITHOUT_TAXON\t5000\t1000\t900\t1\t900\t1\t900\t95.0\t855\t0\t0.0\t1500\tplus\n"
    p = tmp_path / "h.blastout"
    p.write_text(row)
    with pytest.raises(SystemExit):
        parsers.read_blast_hits(str(p))
    for name in ("pa", "pc", "pacsv"):
        monkeypatch.setattr(parsers, name, None)
    with pytest.raises(SystemExit):
        parsers.read_blast_hits(str(p))


def test_sort_hits_composite_key_equals_lexsort():
    """packing.Batch.sort_hits: the composite-key integer sort gives the (contig, taxon, descending score, file order)
    order of np.lexsort; batches with annotation systems stay in file order; taxon indices too wide for the key fall
    back to lexsort."""
    import numpy as np
    from waafle_b200 import synth
    data = synth.generate_config("cfg2", n_contigs=300, seed=12)
    tax = data.taxonomy()
    b = data.to_batch(tax)
    s = b.sort_hits()
    contig = np.repeat(np.arange(b.n_contigs), np.diff(b.hit_off))
    o = np.lexsort((np.arange(len(contig)), -b.hit_score, b.hit_taxon, contig))
    for k in ("hit_qstart", "hit_qend", "hit_taxon", "hit_score", "hit_scov", "hit_strand"):
        assert np.array_equal(getattr(s, k), getattr(b, k)[o]), k
    assert np.array_equal(s.hit_off, b.hit_off)
    wide = data.to_batch(tax)
    wide.hit_taxon = (wide.hit_taxon.astype(np.int64) * 400000 % (1 << 31)).astype(np.int32)   # 31-bit taxon indices
    sw = wide.sort_hits()
    ow = np.lexsort((np.arange(len(contig)), -wide.hit_score, wide.hit_taxon, contig))
    assert np.array_equal(sw.hit_score, wide.hit_score[ow]) and np.array_equal(sw.hit_taxon, wide.hit_taxon[ow])
    ann = synth.generate_config("cfg5", n_contigs=50, seed=13, annotations=True)
    ba = ann.to_batch(ann.taxonomy())
    assert ba.sort_hits() is ba
