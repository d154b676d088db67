"""CPU: front end + oracle + writer against the reference's own shipped demo outputs.

The six TSVs under demo/output*/ are the only golden vectors the reference ships.
They were produced by an older CLI: they pin every non-score column under
`--sister-penalty off --ambiguous-threshold strict`, and the two score columns to 3 decimals.
"""
import csv
import os

import pytest

import helpers
from oracle import orgscorer_oracle as oracle
from waafle_b200 import writer


def read_tsv(path):
    with open(path) as fh:
        rows = list(csv.reader(fh, delimiter="\t"))
    return rows[0], rows[1:]


@pytest.mark.parametrize("prodigal", [False, True])
def test_shipped_demo_outputs(tmp_path, prodigal):
    files = helpers.demo_files(tmp_path, prodigal)
    batch, loci, hits, tax = helpers.frontend_load(files)
    P = helpers.params_for(dict(sister_penalty="off", ambiguous_threshold="strict"), len(hits.systems))
    res = oracle.score_batch(P.as_dict(), tax.tables(), batch.arrays())
    records = writer.build_records(batch, loci, hits, tax, res)
    writer.write_main_output_files(records, str(tmp_path), "run")
    stem = "demo_contigs.prodigal" if prodigal else "demo_contigs"
    for kind, score_cols in (("lgt", (3, 4)), ("no_lgt", (3, 4)), ("unclassified", ())):
        hdr_a, rows_a = read_tsv(os.path.join(str(tmp_path), "run.{}.tsv".format(kind)))
        hdr_b, rows_b = read_tsv(os.path.join(helpers.GOLDEN, "demo", "{}.{}.tsv".format(stem, kind)))
        assert hdr_a == hdr_b
        assert len(rows_a) == len(rows_b)
        for a, b in zip(rows_a, rows_b):
            for col, (x, y) in enumerate(zip(a, b)):
                if col in score_cols:
                    assert abs(float(x) - float(y)) <= 5.1e-4, (kind, a[0], col, x, y)
                else:
                    assert x == y, (kind, a[0], hdr_a[col], x, y)
