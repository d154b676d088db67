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
