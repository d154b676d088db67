"""GPU: the drop-in CLI (`python -m waafle_b200.orgscorer`, reference flags verbatim) end to end -- parsers, packer,
engine through the C ABI, writer -- against the TSVs the reference ships for its demo, and the
chunked scoring path (`--chunk-contigs`) against the single-call path."""
import csv
import os

import pytest

import helpers

pytestmark = pytest.mark.gpu


def read_tsv(path):
    with open(path) as fh:
        rows = list(csv.reader(fh, delimiter="\t"))
    return rows[0], rows[1:]


@pytest.mark.parametrize("prodigal", [False, True])
def test_cli_reproduces_shipped_demo_tsvs(tmp_path, prodigal):
    from waafle_b200 import orgscorer
    files = helpers.demo_files(tmp_path, prodigal)
    outs = {}
    for tag, extra in (("one", []), ("chunked", ["--chunk-contigs", "17"])):
        outdir = tmp_path / tag
        outdir.mkdir()
        orgscorer.main([files["contigs"], files["blastout"], files["gff"], files["taxonomy"],
                        "--outdir", str(outdir), "--basename", "run", "--quiet",
                        "--sister-penalty", "off", "--ambiguous-threshold", "strict"] + extra)
        outs[tag] = outdir
    stem = "demo_contigs.prodigal" if prodigal else "demo_contigs"
    for kind, score_cols in (("lgt", (3, 4)), ("no_lgt", (3, 4)), ("unclassified", ())):
        hdr_a, rows_a = read_tsv(os.path.join(str(outs["one"]), "run.{}.tsv".format(kind)))
        hdr_b, rows_b = read_tsv(os.path.join(helpers.GOLDEN, "demo", "{}.{}.tsv".format(stem, kind)))
        assert hdr_a == hdr_b and len(rows_a) == len(rows_b)
        for a, b in zip(rows_a, rows_b):
            for col, (x, y) in enumerate(zip(a, b)):
                if col in score_cols:
                    assert abs(float(x) - float(y)) <= 5.1e-4, (kind, a[0], col, x, y)   # shipped files: 3 decimals
                else:
                    assert x == y, (kind, a[0], hdr_a[col], x, y)
        with open(os.path.join(str(outs["one"]), "run.{}.tsv".format(kind))) as f1, \
                open(os.path.join(str(outs["chunked"]), "run.{}.tsv".format(kind))) as f2:
            assert f1.read() == f2.read(), kind


def run_cli(files, outdir, extra):
    from waafle_b200 import orgscorer
    os.makedirs(str(outdir), exist_ok=True)
    orgscorer.main([files["contigs"], files["blastout"], files["gff"], files["taxonomy"],
                    "--outdir", str(outdir), "--basename", "demo_contigs", "--quiet"] + extra)
    return {kind: open(os.path.join(str(outdir), "demo_contigs.{}.tsv".format(kind))).read()
            for kind in ("lgt", "no_lgt", "unclassified")}


@pytest.mark.parametrize("gff", ["genecaller", "prodigal"])
def test_cli_bytes_equal_reference_cli(tmp_path, gff):
    """TSV bytes of the drop-in CLI == bytes the CURRENT unmodified reference CLI writes (tests/golden/demo_cli, made by
    tests/golden/make_golden_tsv.py) for the demo under default and non-default flags; fast path and exact pipeline."""
    import json
    root = os.path.join(helpers.GOLDEN, "demo_cli")
    flag_sets = json.load(open(os.path.join(root, "flag_sets.json")))
    files = helpers.demo_files(tmp_path, gff == "prodigal")
    for k, flags in enumerate(flag_sets):
        want = {kind: open(os.path.join(root, "{}_{}".format(gff, k), "demo_contigs.{}.tsv".format(kind))).read()
                for kind in ("lgt", "no_lgt", "unclassified")}
        for tag, extra in (("fast", []), ("exact", ["--exact-scores"]), ("stream", ["--stream-mb", "0.004", "--devices", "0,0"])):
            got = run_cli(files, tmp_path / "{}_{}".format(tag, k), flags + extra)
            for kind in want:
                assert got[kind] == want[kind], (gff, flags, tag, kind)


def test_streamed_two_workers_equal_single_call_on_synthetic(tmp_path):
    """BASELINE configs[4] shape (Prodigal-style loci, annotation transfer, --weak-loci) as text files: the streamed run
    (contig-aligned chunks, two worker processes, GPU parser, per-chunk shards, k-way merge) writes the same bytes as
    the single call, under each --weak-loci mode."""
    from waafle_b200 import synth
    data = synth.generate_config("cfg5", n_contigs=1500, seed=401, annotations=True)
    files = data.write_files(str(tmp_path), "demo_contigs")
    with open(files["contigs"], "a") as fh:
        fh.write(">zzz_no_hits\nACGTACGTAC\n")
    for mode in ("ignore", "penalize", "assign-unknown"):
        one = run_cli(files, tmp_path / ("one_" + mode), ["--weak-loci", mode])
        two = run_cli(files, tmp_path / ("two_" + mode), ["--weak-loci", mode, "--stream-mb", "2", "--devices", "0,0"])
        assert one == two, mode
        assert one["no_lgt"].count("\n") > 100


@pytest.mark.parametrize("gff", ["genecaller", "prodigal"])
def test_write_details_equals_reference_write_details(tmp_path, gff):
    """--write-details (OS:766-812): <basename>.details.tsv.gz == what the unmodified write_details prints through the
    reference harness (tests/golden/demo_cli/details_*.tsv.gz; canonical clade order), default flags, a spiked Unknown
    and --jump-taxonomy."""
    import gzip
    import json
    root = os.path.join(helpers.GOLDEN, "demo_cli")
    flag_sets = json.load(open(os.path.join(root, "details_flag_sets.json")))
    cli = [[], ["--weak-loci", "assign-unknown"], ["--jump-taxonomy", "1", "--weak-loci", "penalize"]]
    assert len(cli) == len(flag_sets)
    files = helpers.demo_files(tmp_path, gff == "prodigal")
    for k, extra in enumerate(cli):
        outdir = tmp_path / "d{}".format(k)
        run_cli(files, outdir, extra + ["--write-details"])
        with gzip.open(os.path.join(str(outdir), "demo_contigs.details.tsv.gz"), "rt") as fh:
            got = fh.read()
        with gzip.open(os.path.join(root, "details_{}_{}.tsv.gz".format(gff, k)), "rt") as fh:
            want = fh.read()
        if got != want:
            g, w = got.split("\n"), want.split("\n")
            bad = next((i for i in range(min(len(g), len(w))) if g[i] != w[i]), min(len(g), len(w)) - 1)
            raise AssertionError((gff, extra, len(g), len(w), g[bad][:300], w[bad][:300]))
