"""CPU: the numpy oracle against golden records produced by the UNMODIFIED reference.

The golden records (tests/golden/*.json.gz) were generated in the build container by
tests/golden/make_golden.py, which drives the reference classes from /root/reference
(oracle/reference_harness.py).  Floats are compared bit-exactly.
"""
import pytest

import helpers
from oracle import orgscorer_oracle as oracle
from oracle.validate_against_reference import compare_records, records_from_results

DEMO = helpers.load_json("demo_records.json.gz")
SYNTH = helpers.load_json("synth_records.json.gz")


def _check(batch, loci, hits, tax, flags, golden):
    P = helpers.params_for(flags, len(hits.systems) if hits is not None else 0)
    res = oracle.score_batch(P.as_dict(), tax.tables(), batch.arrays())
    recs = records_from_results(batch, loci, hits, tax, res)
    diffs = compare_records(helpers.decode_golden(golden), recs)
    assert not diffs, diffs[:5]


@pytest.mark.parametrize("gff", ["genecaller", "prodigal"])
@pytest.mark.parametrize("fi", [0, 1, 3, 4, 6, 8, 11, 14])
def test_demo_against_reference_records(tmp_path_factory, gff, fi):
    files = helpers.demo_files(tmp_path_factory.getbasetemp(), prodigal=(gff == "prodigal"))
    batch, loci, hits, tax = helpers.frontend_load(files)
    _check(batch, loci, hits, tax, DEMO["flag_sets"][fi], DEMO["records"]["{}:{}".format(gff, fi)])


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg5"])
@pytest.mark.parametrize("fi", ["0", "3", "7"])
def test_synthetic_against_reference_records(tmp_path, name, fi):
    entry = SYNTH["cases"][name]
    data = helpers.synth_case(entry["case"])
    if helpers.batch_checksum(data.to_batch()) != entry["checksum"]:
        pytest.skip("synthetic generator drifted from the recorded golden inputs")
    files = data.write_files(str(tmp_path), name)
    batch, loci, hits, tax = helpers.frontend_load(files)
    _check(batch, loci, hits, tax, SYNTH["flag_sets"][int(fi)], entry["records"][fi])
