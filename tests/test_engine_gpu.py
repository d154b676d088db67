"""GPU parity tests proper: the CUDA engine, called through the C ABI, against the oracle.

Two modes, both run by every comparison helper here:
  * default (fused fast-path kernel with guard bands on its decisions and on the 4-decimal print edge of the scores it
    reports; exact pipeline for what it hands back): every integer / byte / index output bit-exact, crit and rank
    within SCORE_RTOL = 1e-12 relative (BASELINE.json north_star tolerance) -- and the TSV bytes identical
    (tests/test_cli_gpu.py);
  * exact (`set_option("exact", 1)`, the numpy-pairwise pipeline for every contig): crit and rank bit-exact as well.
"""
import numpy as np
import pytest

import helpers
from oracle import c_oracle
from oracle import orgscorer_oracle as oracle
from oracle.validate_against_reference import FLAG_SETS, compare_records, records_from_results

pytestmark = pytest.mark.gpu
SCORE_RTOL = 1e-12
MODES = (("fast", 0, SCORE_RTOL), ("exact", 1, 0.0))

DEMO = helpers.load_json("demo_records.json.gz")
SYNTH = helpers.load_json("synth_records.json.gz")


def run_engine(engine, P, tax, batch):
    engine.set_params(P)
    engine.set_taxonomy(tax)
    return engine.score_batch(batch)


def check_vs_ref(engine, P, tax, batch, ref, what=None):
    """Both engine modes against reference results `ref`; leaves the engine in the default (fast) mode."""
    got = None
    try:
        for mode, exact, rtol in MODES:
            engine.set_option("exact", exact)
            got = run_engine(engine, P, tax, batch)
            diffs = helpers.compare_results(ref, got, score_rtol=rtol)
            assert not diffs, (mode, what, diffs[:4])
    finally:
        engine.set_option("exact", 0)
    return got


def check_vs_oracle(engine, batch, tax, flags, n_systems):
    P = helpers.params_for(flags, n_systems)
    ref = oracle.score_batch(P.as_dict(), tax.tables(), batch.arrays())
    return check_vs_ref(engine, P, tax, batch, ref, flags)


@pytest.mark.parametrize("gff", ["genecaller", "prodigal"])
def test_demo_golden_records_through_c_abi(engine, tmp_path, gff):
    """Engine output -> writer records == records of the unmodified reference (all flag sets)."""
    batch, loci, hits, tax = helpers.frontend_load(helpers.demo_files(tmp_path, gff == "prodigal"))
    for fi, flags in enumerate(DEMO["flag_sets"]):
        P = helpers.params_for(flags, len(hits.systems))
        for mode, exact, rtol in MODES:
            engine.set_option("exact", exact)
            res = run_engine(engine, P, tax, batch)
            recs = records_from_results(batch, loci, hits, tax, res)
            diffs = compare_records(helpers.decode_golden(DEMO["records"]["{}:{}".format(gff, fi)]), recs, exact_scores=(rtol == 0.0), rtol=SCORE_RTOL)
            assert not diffs, (mode, flags, diffs[:4])
    engine.set_option("exact", 0)


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg5", "cfg4"])
def test_synthetic_golden_records_through_c_abi(engine, tmp_path, name):
    entry = SYNTH["cases"][name]
    data = helpers.synth_case(entry["case"])
    if helpers.batch_checksum(data.to_batch()) != entry["checksum"]:
        pytest.skip("synthetic generator drifted from the recorded golden inputs")
    batch, loci, hits, tax = helpers.frontend_load(data.write_files(str(tmp_path), name))
    for fi, golden in entry["records"].items():
        P = helpers.params_for(SYNTH["flag_sets"][int(fi)], len(hits.systems))
        for mode, exact, rtol in MODES:
            engine.set_option("exact", exact)
            res = run_engine(engine, P, tax, batch)
            recs = records_from_results(batch, loci, hits, tax, res)
            diffs = compare_records(helpers.decode_golden(golden), recs, exact_scores=(rtol == 0.0), rtol=SCORE_RTOL)
            assert not diffs, (mode, name, fi, diffs[:4])
    engine.set_option("exact", 0)


@pytest.mark.parametrize("config,n,seed", [("cfg2", 400, 21), ("cfg3", 150, 22), ("cfg5", 150, 23)])
def test_synthetic_vs_oracle_all_flag_sets(engine, config, n, seed):
    from waafle_b200 import synth
    data = synth.generate_config(config, n_contigs=n, seed=seed)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    for flags in FLAG_SETS:
        check_vs_oracle(engine, batch, tax, flags, 1 if batch.hit_sysmask is not None else 0)


def test_long_contigs_multiword_masks_vs_oracle(engine):
    """cfg4 shape (>=100 genes, hundreds of taxa): multi-word masks, pair blow-up, slab workspace."""
    from waafle_b200 import synth
    data = synth.generate_config("cfg4", n_contigs=3, seed=31, hits_per_gene=14.0)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    for flags in ({}, dict(weak_loci="assign-unknown", range=0.3), dict(sister_penalty="off")):
        check_vs_oracle(engine, batch, tax, flags, 0)
    assert engine.stats()["smem_contigs"] < batch.n_contigs   # these spill to the global slab


@pytest.mark.parametrize("name", ["knife_edge", "ties", "multiword", "odd_inputs"])
def test_adversarial_vs_oracle(engine, name):
    batch, tax = helpers.adversarial_batches()[name]
    for flags in helpers.ADVERSARIAL_FLAGS:
        got = check_vs_oracle(engine, batch, tax, flags, 1)
    if name == "knife_edge":
        # rounding alone splits this fixture between calls (finding 2): the fast path must
        # recompute those gene scores in numpy's summation order (guard band) to get every call right
        P = helpers.params_for({}, 1)
        got = run_engine(engine, P, tax, batch)
        assert got["call_counts"][0] > 100 and got["call_counts"][1] > 100
        st = engine.stats()
        assert st["refined_groups"] > 500 and st["smem_contigs"] + st["fallback_contigs"] == batch.n_contigs


def test_min_overlap_zero_python_slice_quirk(engine):
    """--min-overlap 0: disjoint hits 'match' and a hit left of the locus wraps the python slice."""
    batch, tax = helpers.adversarial_batches()["odd_inputs"]
    check_vs_oracle(engine, batch, tax, dict(min_overlap=0.0, min_scov=0.0), 1)
    check_vs_oracle(engine, batch, tax, dict(min_overlap=-1.0, min_gene_length=0.0), 1)


def test_level0_gene_scores_bit_exact(engine):
    """K1+K2 unit test: every level-0 (clade, locus) gene score equals np.mean of the site array."""
    from waafle_b200 import synth
    data = synth.generate_config("cfg2", n_contigs=40, seed=41)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    P = helpers.params_for(dict(weak_loci="penalize"), 0)
    ref = oracle.score_batch(P.as_dict(), tax.tables(), batch.arrays(), want_gene_scores=True)
    engine.set_option("exact", 1)   # the dump comes from the exact pipeline
    engine.set_params(P)
    engine.set_taxonomy(tax)
    engine.upload(batch)
    n_checked = 0
    for c in range(batch.n_contigs):
        cl, lo, sc = engine.debug_gene_scores(c)
        want = {(k, i): float(v[i]) for k, v in ref["gene_scores"][c].items() for i in range(len(v))}
        got = {(int(a), int(b)): float(s) for a, b, s in zip(cl, lo, sc)}
        for key, s in got.items():
            assert want[key].hex() == s.hex(), (c, key)
        # entries the engine does not list are loci without a matched hit: score 0
        assert all(v == 0.0 for k, v in want.items() if k not in got)
        n_checked += len(got)
    engine.set_option("exact", 0)
    assert n_checked > 1000


def test_error_paths(engine):
    from waafle_b200.engine import EngineError
    batch, tax = helpers.adversarial_batches()["ties"]
    engine.set_params(helpers.params_for({}, 1))
    engine.set_taxonomy(tax)
    bad = dict(batch.arrays())
    bad["hit_taxon"] = bad["hit_taxon"].copy()
    bad["hit_taxon"][0] = 10 ** 6
    with pytest.raises(EngineError):
        engine.score_batch(bad)
    bad = dict(batch.arrays())
    bad["hit_off"] = bad["hit_off"].copy()
    bad["hit_off"][-1] += 1
    with pytest.raises(EngineError):
        engine.score_batch(bad)
    with pytest.raises(EngineError):
        engine.set_params(dict(weak_loci=7))


def test_full_size_properties(engine):
    """BASELINE configs[1] at full size (100k contigs): size-independent properties.

    (a) idempotence / determinism: two runs give identical bytes; (b) contigs are independent:
    scoring a permuted batch permutes the results; (c) shards concatenate to the whole;
    (d) checksum of the compacted outputs is consistent (counts partition the contigs);
    (e) a random sample of contigs agrees with the oracle.
    """
    from waafle_b200 import synth
    data = synth.generate_config("cfg2", seed=1000)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    P = helpers.params_for({}, 0)
    a = run_engine(engine, P, tax, batch)
    b = engine.score_batch(batch)
    st = engine.stats()
    assert st["smem_contigs"] > 0.9 * batch.n_contigs   # the fused shared-memory kernel is the path that ran
    assert st["smem_contigs"] + st["fallback_contigs"] == batch.n_contigs
    for k in helpers.EXACT_FIELDS + ["crit", "rank"]:
        assert np.array_equal(a[k], b[k]), k
    n = batch.n_contigs
    assert a["call_counts"].sum() == n
    assert np.array_equal(np.sort(a["call_index"]), np.arange(n))
    assert np.all(a["call"][a["call_index"][:a["call_counts"][0]]] == 2)
    assert a["member_off"][-1] == len(a["members"])
    # (c) two shards
    cut = n // 3
    s1, s2 = engine.score_batch(batch.slice(0, cut)), engine.score_batch(batch.slice(cut, n))
    for k in ("call", "clade1", "clade2", "lca", "lifts", "crit", "rank", "synteny", "locus_flags"):
        assert np.array_equal(np.concatenate([s1[k], s2[k]]), a[k]), k
    # (e) sample vs oracle
    rng = np.random.default_rng(0)
    for c0 in rng.integers(0, n - 25, size=8):
        sub = batch.slice(int(c0), int(c0) + 25)
        ref = oracle.score_batch(P.as_dict(), tax.tables(), sub.arrays())
        got = engine.score_batch(sub)
        assert not helpers.compare_results(ref, got, score_rtol=SCORE_RTOL)
    # (f) the compact wire format (14 B/hit) gives the same bytes as the wide one
    c = engine.score_batch(batch.to_packed(P.min_scov))
    for k in helpers.EXACT_FIELDS + ["crit", "rank"]:
        assert np.array_equal(a[k], c[k]), k


@pytest.mark.parametrize("config,n,flags", [
    ("cfg2", None, {}),                                        # BASELINE configs[1] at full size: 100k contigs
    ("cfg2", 30000, dict(weak_loci="penalize", range=0.3, allow_lca=True)),
    ("cfg3", 30000, {}),                                       # configs[2] shape, 8 levels
    ("cfg5", 20000, dict(weak_loci="assign-unknown")),         # configs[4] shape, annotations
    ("cfg4", 24, dict(sister_penalty="off")),                  # configs[3] shape: 100+ genes, 550 taxa
])
def test_full_size_bit_exact_vs_c_oracle(engine, config, n, flags):
    """Every contig of the full-size workloads, every output byte, against the C restatement of the
    reference (oracle/orgscorer_oracle.c, itself pinned to the numpy oracle in tests/test_c_oracle.py)."""
    from waafle_b200 import synth
    data = synth.generate_config(config, n_contigs=n, seed=1000, annotations=(config == "cfg5"))
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    P = helpers.params_for(flags, 1 if config == "cfg5" else 0)
    ref = c_oracle.score_batch(P, tax, batch)
    got = check_vs_ref(engine, P, tax, batch, ref, (config, flags))
    assert got["call_counts"].sum() == batch.n_contigs
    if batch.can_pack(len(tax.tables()["parent"]), P.n_systems):
        got = engine.score_batch(batch.to_packed(P.min_scov))
        assert not helpers.compare_results(ref, got, score_rtol=SCORE_RTOL), "packed"


@pytest.mark.parametrize("config,n,flags", [
    ("cfg2", 30000, {}),
    ("cfg3", 20000, dict(weak_loci="penalize", range=0.2)),
    ("cfg5", 20000, dict(weak_loci="assign-unknown")),
])
def test_presorted_hits_bit_exact_vs_c_oracle(engine, config, n, flags):
    """What the front end's packer delivers: every contig's hits in descending score order (Batch.sort_hits).  The fast
    kernel then skips its per-locus sort; results -- annotation winners included -- equal the C restatement's on the
    same batch, and (but for the hit indices) the unsorted batch's."""
    from waafle_b200 import synth
    data = synth.generate_config(config, n_contigs=n, seed=1001, annotations=(config == "cfg5"))
    tax = data.taxonomy()
    raw = data.to_batch(tax)
    batch = raw.sort_hits()
    P = helpers.params_for(flags, 1 if config == "cfg5" else 0)
    ref = c_oracle.score_batch(P, tax, batch)
    check_vs_ref(engine, P, tax, batch, ref, (config, flags, "sorted"))
    got = run_engine(engine, P, tax, batch)
    unsorted = run_engine(engine, P, tax, raw)
    for k in ("call", "clade1", "clade2", "lca", "lifts", "synteny", "locus_flags", "members", "crit", "rank"):
        assert np.array_equal(got[k], unsorted[k]), k


def test_fast_path_capacity_fallbacks():
    """Tiny slice capacities: most contigs overflow the fast kernel's shared-memory slice and are handed to the exact
    pipeline; results do not change (calls bit-exact, scores within tolerance)."""
    from waafle_b200 import synth
    from waafle_b200.engine import Engine
    data = synth.generate_config("cfg3", n_contigs=600, seed=63, annotations=True)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    P = helpers.params_for(dict(weak_loci="assign-unknown", range=0.3), 1)
    ref = c_oracle.score_batch(P, tax, batch)
    for opts in (dict(fast_kcap=64, fast_tcap=64, fast_ncap=96), dict(fast_tcap=64), dict(fast_ncap=96), {}):
        eng = Engine(0, P, tax)
        for k, v in opts.items():
            eng.set_option(k, v)
        got = eng.score_batch(batch)
        st = eng.stats()
        eng.close()
        assert not helpers.compare_results(ref, got, score_rtol=SCORE_RTOL), opts
        assert st["smem_contigs"] + st["fallback_contigs"] == batch.n_contigs
        if opts:   # overflows of the first pass go to the second one (3x the slice), what that cannot hold to the exact pipeline
            assert st["second_pass_contigs"] > 0, opts
        if "fast_tcap" in opts:
            assert st["fallback_contigs"] > 0, opts


def test_many_small_chunks_on_two_streams(monkeypatch):
    """Plugin call cut into ~50 chunks of 1 MB that alternate between the two compute streams (and, with a
    64 MB workspace pool, into even smaller sub-batches): same bytes as the C restatement and as one chunk."""
    from waafle_b200 import synth
    from waafle_b200.engine import Engine
    data = synth.generate_config("cfg3", n_contigs=3000, seed=71)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    P = helpers.params_for(dict(weak_loci="penalize", range=0.2), 0)
    ref = c_oracle.score_batch(P, tax, batch)
    for env in (dict(WFL_CHUNK_MB="1"), dict(WFL_CHUNK_MB="1", WFL_STREAMS="1"), dict(WFL_CHUNK_MB="2", WFL_POOL_MB="64"),
                dict(WFL_CHUNK_MB="4096")):
        for k in ("WFL_CHUNK_MB", "WFL_STREAMS", "WFL_POOL_MB"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for exact in (0, 1):
            eng = Engine(0, P, tax)
            eng.set_option("exact", exact)
            got = eng.score_batch(batch)
            again = eng.score_batch(batch)
            st = eng.stats()
            eng.close()
            assert not helpers.compare_results(ref, got, score_rtol=0.0 if exact else SCORE_RTOL), env
            assert not helpers.compare_results(ref, again, score_rtol=0.0 if exact else SCORE_RTOL), env
            if env.get("WFL_CHUNK_MB") == "1":
                assert st["kernel_launches"] > (20 * 20 if exact else 20)   # many chunks really ran


def test_group_list_overflow_replays(monkeypatch):
    """Exact pipeline with a K2 group list that is far too small (WFL_K2_CAP): contigs that do not fit are replayed
    with a larger list, the list never holds unwritten items, and the results do not change."""
    from waafle_b200 import synth
    from waafle_b200.engine import Engine
    data = synth.generate_config("cfg2", n_contigs=600, seed=77)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    P = helpers.params_for(dict(weak_loci="assign-unknown"), 0)
    ref = c_oracle.score_batch(P, tax, batch)
    for cap in ("3000", "1"):
        monkeypatch.setenv("WFL_K2_CAP", cap)
        eng = Engine(0, P, tax)
        eng.set_option("exact", 1)
        got = eng.score_batch(batch)
        st = eng.stats()
        eng.close()
        assert st["workspace_retries"] > 0
        assert not helpers.compare_results(ref, got), cap


def test_small_workspace_pool_replays(monkeypatch):
    """A workspace pool too small for a long contig: it is replayed with a larger pool (workspace_retries > 0)
    and the results do not change (members included)."""
    from waafle_b200 import synth
    from waafle_b200.engine import Engine
    data = synth.generate_config("cfg4", n_contigs=4, seed=62)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    P = helpers.params_for(dict(sister_penalty="off", range=0.3), 0)
    eng = Engine(0, P, tax)
    full = eng.score_batch(batch)
    assert eng.stats()["workspace_retries"] == 0
    eng.close()
    monkeypatch.setenv("WFL_POOL_MB", "0")
    eng = Engine(0, P, tax)
    small = eng.score_batch(batch)
    assert eng.stats()["workspace_retries"] > 0
    eng.close()
    assert not helpers.compare_results(full, small)
