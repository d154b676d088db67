"""CPU: the C restatement (oracle/orgscorer_oracle.c) against the numpy oracle of record and the
reference-generated golden records.  Everything bit-exact, crit/rank included.  The C restatement is the
checker used at BASELINE.json's full sizes (tests/test_engine_gpu.py) and bench.py's multi-threaded CPU arm.
"""
import numpy as np
import pytest

import helpers
from oracle import c_oracle
from oracle import orgscorer_oracle as oracle
from oracle.validate_against_reference import compare_records, records_from_results

DEMO = helpers.load_json("demo_records.json.gz")
SYNTH = helpers.load_json("synth_records.json.gz")


def _same(params, tax, batch, threads=None):
    ref = oracle.score_batch(params.as_dict(), tax.tables(), batch.arrays())
    got = c_oracle.score_batch(params, tax, batch, threads=threads)
    diffs = helpers.compare_results(ref, got)
    assert not diffs, diffs[:5]


@pytest.mark.parametrize("name", ["knife_edge", "ties", "multiword", "odd_inputs"])
def test_adversarial_suites(name):
    batch, tax = helpers.adversarial_batches()[name]
    flag_sets = helpers.ADVERSARIAL_FLAGS if name != "knife_edge" else helpers.ADVERSARIAL_FLAGS[:4]
    for flags in flag_sets:
        for S in (0, 1):
            _same(helpers.params_for(flags, S), tax, batch, threads=3)


@pytest.mark.parametrize("cfg,n", [("cfg2", 300), ("cfg3", 120), ("cfg5", 120), ("cfg4", 2)])
def test_synthetic_shapes(cfg, n):
    from waafle_b200 import synth
    data = synth.generate_config(cfg, n_contigs=n, seed=11, annotations=(cfg != "cfg4"))
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    S = 1 if batch.hit_sysmask is not None else 0
    for flags in ({}, helpers.ADVERSARIAL_FLAGS[2], helpers.ADVERSARIAL_FLAGS[5], helpers.ADVERSARIAL_FLAGS[8]):
        _same(helpers.params_for(flags, S), tax, batch)


def test_thread_count_does_not_change_results():
    from waafle_b200 import synth
    data = synth.generate_config("cfg2", n_contigs=500, seed=5)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    P = helpers.params_for({}, 0)
    a = c_oracle.score_batch(P, tax, batch, threads=1)
    b = c_oracle.score_batch(P, tax, batch, threads=7)
    assert not helpers.compare_results(a, b)


@pytest.mark.parametrize("gff", ["genecaller", "prodigal"])
@pytest.mark.parametrize("fi", [0, 3, 8, 14])
def test_demo_against_reference_records(tmp_path_factory, gff, fi):
    files = helpers.demo_files(tmp_path_factory.getbasetemp(), prodigal=(gff == "prodigal"))
    batch, loci, hits, tax = helpers.frontend_load(files)
    P = helpers.params_for(DEMO["flag_sets"][fi], len(hits.systems))
    res = c_oracle.score_batch(P, tax, batch)
    recs = records_from_results(batch, loci, hits, tax, res)
    diffs = compare_records(helpers.decode_golden(DEMO["records"]["{}:{}".format(gff, fi)]), recs)
    assert not diffs, diffs[:5]
