import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import numpy as np
from waafle_b200 import synth
from waafle_b200.engine import Engine
from waafle_b200.params import OrgscorerParams
from oracle import orgscorer_oracle as oracle
from helpers import compare_results
for name, n, seed, kw in [("cfg2", 200, 1, {}), ("cfg3", 100, 2, {}), ("cfg5", 100, 3, {})]:
    data = synth.generate_config(name, n_contigs=n, seed=seed)
    tax = data.taxonomy(); batch = data.to_batch(tax)
    for flags in [dict(), dict(weak_loci=2), dict(weak_loci=1), dict(disambiguate_one=0, disambiguate_two=0), dict(sister_penalty=0, ambiguous_threshold=2), dict(jump_taxonomy=1, clade_genes=2, clade_leaves=2)]:
        P = OrgscorerParams(n_systems=1 if batch.hit_sysmask is not None else 0, **flags)
        eng = Engine(0, P, tax)
        t=time.time(); got = eng.score_batch(batch); dt=time.time()-t
        ref = oracle.score_batch(P.as_dict(), tax.tables(), batch.arrays())
        d = compare_results(ref, got)
        print(name, flags, "counts", got["call_counts"].tolist(), "diffs", len(d), "t=%.3f"%dt, eng.stats())
        for x in d[:6]: print("   ", x)
        eng.close()
