"""CPU: the streaming driver (waafle_b200/streaming.py) around a stand-in engine backed by the numpy oracle -- chunk
scanner, per-chunk packing against the loci index, shard files and the k-way merge -- produces the same bytes as the
single-pass writer (waafle_orgscorer.py:814-894 semantics: rows sorted by contig name, one header with the union of the
transferred annotation systems)."""
import os

import numpy as np
import pytest

import helpers
from oracle import orgscorer_oracle as oracle
from waafle_b200 import packing, parsers, streaming, synth, taxonomy, writer
from waafle_b200.params import OrgscorerParams
from waafle_b200.utils import read_contig_lengths


class OracleEngine:
    """Same calls as waafle_b200.engine.Engine, results from the CPU oracle (test infrastructure only)."""

    def __init__(self, device, params, tax):
        self.params, self.tax = params, tax

    def set_params(self, params):
        self.params = params

    def set_taxonomy(self, tax):
        self.tax = tax

    def score_batch(self, batch):
        return oracle.score_batch(self.params.as_dict(), self.tax.tables(), batch.arrays())

    def stats(self):
        return dict(ms_kernels=0.0)

    def close(self):
        pass


def single_pass(files, outdir, flags):
    tax = taxonomy.Taxonomy(files["taxonomy"])
    lengths = read_contig_lengths(files["contigs"])
    loci = parsers.read_gff_loci(files["gff"])
    hits = parsers.read_blast_hits(files["blastout"])
    tax.build(hits.distinct_taxa())
    batch = packing.pack(lengths, loci, hits, tax)
    P = helpers.params_for(flags, len(hits.systems))
    res = OracleEngine(0, P, tax).score_batch(batch)
    writer.write_main_output_files(writer.build_records(batch, loci, hits, tax, res), outdir, "run")


def streamed(files, outdir, flags, chunk_bytes):
    tax = taxonomy.Taxonomy(files["taxonomy"])
    lengths = read_contig_lengths(files["contigs"])
    loci = parsers.read_gff_loci(files["gff"])
    index = streaming.LociIndex(loci, lengths)
    scorer = streaming.ChunkScorer(0, None, tax, lengths, index, lambda n: helpers.params_for(flags, n), cpu_parse=True,
                                   engine_factory=OracleEngine)
    prefixes, seen = [], set()
    chunks = streaming.scan_blast_chunks(files["blastout"], chunk_bytes)
    with open(files["blastout"], "rb") as fh:
        for cid, (off, length) in enumerate(chunks):
            fh.seek(off)
            hits = scorer.parse(fh.read(length))
            records, names = scorer.score(streaming.chunk_contigs(hits, lengths), hits)
            assert not seen.intersection(names)
            seen.update(names)
            prefixes.append(os.path.join(outdir, "c{}".format(cid)))
            streaming.write_shard(records, prefixes[-1])
    rest = [nm for nm in lengths if nm not in seen]
    if rest:
        records, _ = scorer.score(rest, parsers.hits_from_columns(*([[]] * 10)))
        prefixes.append(os.path.join(outdir, "rest"))
        streaming.write_shard(records, prefixes[-1])
    streaming.merge_shards(prefixes, outdir, "run")
    assert not [f for f in os.listdir(outdir) if f.endswith(".shard")]
    return len(chunks)


def same_outputs(d1, d2):
    for kind in ("lgt", "no_lgt", "unclassified"):
        with open(os.path.join(d1, "run.{}.tsv".format(kind))) as a, open(os.path.join(d2, "run.{}.tsv".format(kind))) as b:
            assert a.read() == b.read(), kind


@pytest.mark.parametrize("prodigal", [False, True])
def test_demo_streamed_equals_single_pass(tmp_path, prodigal):
    files = helpers.demo_files(tmp_path, prodigal)
    for k, (flags, chunk) in enumerate([({}, 20000), (dict(weak_loci="assign-unknown"), 3000), (dict(weak_loci="penalize"), 1 << 30)]):
        one, two = tmp_path / "one{}".format(k), tmp_path / "two{}".format(k)
        one.mkdir()
        two.mkdir()
        single_pass(files, str(one), flags)
        n = streamed(files, str(two), flags, chunk)
        assert n >= 1
        same_outputs(str(one), str(two))


def test_synthetic_annotated_streamed_equals_single_pass(tmp_path):
    """cfg5 shape (Prodigal-style loci, annotations) plus contigs without hits; a chunk whose hits carry no annotation at
    all still prints `None` per locus under the global header."""
    data = synth.generate_config("cfg5", n_contigs=90, seed=77, annotations=True)
    files = data.write_files(str(tmp_path), "s")
    # strip the annotations of the last third of the rows (a chunk without systems) and add hit-less contigs
    rows = open(files["blastout"]).read().split("\n")
    cut = 2 * len(rows) // 3
    rows = rows[:cut] + [r.replace("|UniProt=", "_U") for r in rows[cut:]]
    with open(files["blastout"], "w") as fh:
        fh.write("\n".join(rows))
    with open(files["contigs"], "a") as fh:
        fh.write(">zz_nohits_1 x\nACGT\n>aa_nohits_2\nACGTACGT\n")   # no hits, no loci: every cell prints "--"
    flags = dict(weak_loci="assign-unknown")
    one, two = tmp_path / "one", tmp_path / "two"
    one.mkdir()
    two.mkdir()
    single_pass(files, str(one), flags)
    n = streamed(files, str(two), flags, os.path.getsize(files["blastout"]) // 7)
    assert n >= 5
    same_outputs(str(one), str(two))


def test_chunks_are_contig_aligned(tmp_path):
    data = synth.generate_config("cfg2", n_contigs=200, seed=5)
    files = data.write_files(str(tmp_path), "s")
    raw = open(files["blastout"], "rb").read()
    for target in (1 << 13, 1 << 16, len(raw) + 10):
        chunks = streaming.scan_blast_chunks(files["blastout"], target)
        assert chunks[0][0] == 0 and sum(n for _, n in chunks) == len(raw)
        seen = set()
        for off, n in chunks:
            assert off == 0 or raw[off - 1:off] == b"\n"
            names = {r.split(b"\t")[0] for r in raw[off:off + n].split(b"\n") if r}
            assert not names & seen
            seen |= names


def test_run_streaming_two_worker_processes(tmp_path):
    """The whole driver -- chunk queue, two forked workers, results queue, hit-less contigs, merge -- with the stand-in
    engine; same bytes as the single pass.  An ungrouped blastout is caught across chunks."""
    import argparse
    import functools
    from waafle_b200 import orgscorer
    data = synth.generate_config("cfg5", n_contigs=120, seed=78, annotations=True)
    files = data.write_files(str(tmp_path), "s")
    with open(files["contigs"], "a") as fh:
        fh.write(">zz_nohits\nACGT\n")
    flags = dict(weak_loci="penalize")
    one, two = tmp_path / "one", tmp_path / "two"
    one.mkdir()
    two.mkdir()
    single_pass(files, str(one), flags)
    args = orgscorer.get_args([files["contigs"], files["blastout"], files["gff"], files["taxonomy"], "--outdir", str(two),
                               "--basename", "run", "--weak-loci", "penalize", "--cpu-parse", "--quiet"])
    tax = taxonomy.Taxonomy(files["taxonomy"])
    lengths = read_contig_lengths(files["contigs"])
    loci = parsers.read_gff_loci(files["gff"])
    stats, n_chunks = streaming.run_streaming(args, tax, lengths, loci, functools.partial(orgscorer.params_from_args, args),
                                              [0, 0], os.path.getsize(files["blastout"]) // 9, engine_factory=OracleEngine)
    assert n_chunks >= 6 and sum(st["chunks"] for k, st in stats.items() if k != "rest") == n_chunks
    same_outputs(str(one), str(two))
    assert not [f for f in os.listdir(str(two)) if f.startswith("wfl_shards_")]
    # the same contig's hits in two places of the file
    rows = open(files["blastout"]).read().split("\n")
    with open(files["blastout"], "w") as fh:
        fh.write("\n".join(rows + rows[:3]))
    with pytest.raises(SystemExit):
        streaming.run_streaming(args, taxonomy.Taxonomy(files["taxonomy"]), lengths, loci,
                                functools.partial(orgscorer.params_from_args, args), [0, 0],
                                os.path.getsize(files["blastout"]) // 9, engine_factory=OracleEngine)
