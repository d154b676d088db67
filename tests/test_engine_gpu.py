"""GPU parity tests proper: the CUDA engine, called through the C ABI, against the oracle.

Bar: every integer / byte / index output bit-exact; crit and rank bit-exact as well (the engine
reproduces numpy's pairwise summation order), which is stricter than the 1e-12 relative
tolerance BASELINE.json's north_star allows.
"""
import numpy as np
import pytest

import helpers
from oracle import c_oracle
from oracle import orgscorer_oracle as oracle
from oracle.validate_against_reference import FLAG_SETS, compare_records, records_from_results

pytestmark = pytest.mark.gpu

DEMO = helpers.load_json("demo_records.json.gz")
SYNTH = helpers.load_json("synth_records.json.gz")


def run_engine(engine, P, tax, batch):
    engine.set_params(P)
    engine.set_taxonomy(tax)
    return engine.score_batch(batch)


def check_vs_oracle(engine, batch, tax, flags, n_systems):
    P = helpers.params_for(flags, n_systems)
    got = run_engine(engine, P, tax, batch)
    ref = oracle.score_batch(P.as_dict(), tax.tables(), batch.arrays())
    diffs = helpers.compare_results(ref, got)
    assert not diffs, (flags, diffs[:4])
    return got


@pytest.mark.parametrize("gff", ["genecaller", "prodigal"])
def test_demo_golden_records_through_c_abi(engine, tmp_path, gff):
    """Engine output -> writer records == records of the unmodified reference (all flag sets)."""
    batch, loci, hits, tax = helpers.frontend_load(helpers.demo_files(tmp_path, gff == "prodigal"))
    for fi, flags in enumerate(DEMO["flag_sets"]):
        P = helpers.params_for(flags, len(hits.systems))
        res = run_engine(engine, P, tax, batch)
        recs = records_from_results(batch, loci, hits, tax, res)
        diffs = compare_records(helpers.decode_golden(DEMO["records"]["{}:{}".format(gff, fi)]), recs)
        assert not diffs, (flags, diffs[:4])


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg5", "cfg4"])
def test_synthetic_golden_records_through_c_abi(engine, tmp_path, name):
    entry = SYNTH["cases"][name]
    data = helpers.synth_case(entry["case"])
    if helpers.batch_checksum(data.to_batch()) != entry["checksum"]:
        pytest.skip("synthetic generator drifted from the recorded golden inputs")
    batch, loci, hits, tax = helpers.frontend_load(data.write_files(str(tmp_path), name))
    for fi, golden in entry["records"].items():
        P = helpers.params_for(SYNTH["flag_sets"][int(fi)], len(hits.systems))
        res = run_engine(engine, P, tax, batch)
        recs = records_from_results(batch, loci, hits, tax, res)
        diffs = compare_records(helpers.decode_golden(golden), recs)
        assert not diffs, (name, fi, diffs[:4])


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
        # rounding alone splits this fixture between calls (finding 2)
        P = helpers.params_for({}, 1)
        got = run_engine(engine, P, tax, batch)
        assert got["call_counts"][0] > 100 and got["call_counts"][1] > 100


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
        assert not helpers.compare_results(ref, got)


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
    got = run_engine(engine, P, tax, batch)
    ref = c_oracle.score_batch(P, tax, batch)
    diffs = helpers.compare_results(ref, got)
    assert not diffs, diffs[:4]
    assert got["call_counts"].sum() == batch.n_contigs


def test_tree_walk_gene_scores_bit_exact(monkeypatch):
    """WFL_K2=tree: the gene-score sums by tree walk with constant-subtree skipping (experimental,
    slower on B200) must reproduce the flat leaf-plan walk bit for bit."""
    from waafle_b200 import synth
    from waafle_b200.engine import Engine
    data = synth.generate_config("cfg2", n_contigs=500, seed=63)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    P = helpers.params_for({}, 0)
    outs = []
    # "global": K2 over the sorted global group list (default); "contig": K2 per contig; "tree": per contig, tree walk
    for k2 in ("global", "contig", "tree"):
        monkeypatch.setenv("WFL_K2", k2)
        eng = Engine(0, P, tax)
        outs.append(eng.score_batch(batch))
        eng.close()
    assert not helpers.compare_results(outs[0], outs[1])
    assert not helpers.compare_results(outs[0], outs[2])
    batch2, tax2 = helpers.adversarial_batches()["knife_edge"]
    eng = Engine(0, helpers.params_for({}, 1), tax2)
    got = eng.score_batch(batch2)
    eng.close()
    ref = oracle.score_batch(helpers.params_for({}, 1).as_dict(), tax2.tables(), batch2.arrays())
    assert not helpers.compare_results(ref, got)


@pytest.mark.parametrize("mode", ["v2", "v1"])
def test_alternate_kernels_stay_bit_exact(mode, monkeypatch):
    """The monolithic warp kernel (v2: also the replay path for contigs that outgrow the pipeline's
    workspace) and the first CTA-per-contig kernel (v1) are selectable with WFL_KERNEL and must give
    the same bytes as the default multi-kernel pipeline."""
    from waafle_b200 import synth
    from waafle_b200.engine import Engine
    data = synth.generate_config("cfg3", n_contigs=300, seed=61, annotations=True)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    outs = {}
    for m in ("pipe", mode):
        monkeypatch.setenv("WFL_KERNEL", m)
        for flags in ({}, dict(weak_loci="assign-unknown", range=0.3), dict(jump_taxonomy=1)):
            P = helpers.params_for(flags, 1)
            eng = Engine(0, P, tax)
            outs.setdefault(m, []).append(eng.score_batch(batch))
            eng.close()
    for a, b in zip(outs["pipe"], outs[mode]):
        assert not helpers.compare_results(a, b)
    ref = oracle.score_batch(helpers.params_for({}, 1).as_dict(), tax.tables(), batch.arrays())
    assert not helpers.compare_results(ref, outs[mode][0])


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
        eng = Engine(0, P, tax)
        got = eng.score_batch(batch)
        again = eng.score_batch(batch)
        st = eng.stats()
        eng.close()
        assert not helpers.compare_results(ref, got), env
        assert not helpers.compare_results(ref, again), env
        if env.get("WFL_CHUNK_MB") == "1":
            assert st["kernel_launches"] > 20 * 20   # many chunks really ran


def test_group_list_overflow_replays(monkeypatch):
    """A K2 group list that is far too small (WFL_K2_CAP): contigs that do not fit are replayed by the warp
    kernel, the list never holds unwritten items, and the results do not change."""
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
        got = eng.score_batch(batch)
        st = eng.stats()
        eng.close()
        assert st["workspace_retries"] > 0
        assert not helpers.compare_results(ref, got), cap


def test_small_workspace_pool_replays(monkeypatch):
    """A pool too small for a long contig: it is replayed by the monolithic kernel
    (workspace_retries > 0) and the results do not change (members included)."""
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
