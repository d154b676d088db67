"""End-to-end (plugin call, pinned host buffers) timing under different streaming knobs.

usage: python tools/e2e_sweep.py [workload] [contigs] [steps] -- "WFL_STREAMS=1" "WFL_STREAMS=2 WFL_CHUNK_MB=64" ...
Each quoted group is a set of environment variables read by wfl_create; one engine per group, same batch.
Results of every group are compared with the first group's (must be identical bytes).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from waafle_b200 import synth                                  # noqa: E402
from waafle_b200.engine import Engine, PinnedArena             # noqa: E402
from waafle_b200.params import OrgscorerParams                 # noqa: E402
from helpers import compare_results                            # noqa: E402


def main():
    argv = sys.argv[1:]
    groups = [""]
    if "--" in argv:
        i = argv.index("--")
        argv, groups = argv[:i], argv[i + 1:]
    workload = argv[0] if len(argv) > 0 else "cfg2"
    n = int(argv[1]) if len(argv) > 1 else None
    steps = int(argv[2]) if len(argv) > 2 else 8
    data = synth.generate_config(workload, n_contigs=n, seed=1000)
    tax = data.taxonomy()
    batch = data.to_batch(tax).sort_hits()   # what the front end's packer delivers
    P = OrgscorerParams(n_systems=1 if batch.hit_sysmask is not None else 0)
    pin = PinnedArena()
    wide = os.environ.get("WFL_SWEEP_WIDE") == "1" or not batch.can_pack(len(tax.tables()["parent"]), P.n_systems)
    wire = batch.arrays() if wide else batch.to_packed(P.min_scov)
    harr = {k: pin.like(np.ascontiguousarray(v)) for k, v in wire.items()}
    first = None
    for g in groups:
        env = dict(kv.split("=", 1) for kv in g.split()) if g else {}
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        eng = Engine(0, P, tax)
        eng.use_pinned_results(True)
        for _ in range(3):
            out = eng.score_batch(harr)
        ts = []
        for _ in range(steps):
            t = time.perf_counter()
            out = eng.score_batch(harr)
            ts.append(1e3 * (time.perf_counter() - t))
        st = eng.stats()
        res = {k: np.array(v) for k, v in out.items()}
        same = "ref" if first is None else ("same" if not compare_results(first, res) else "DIFFERENT")
        if first is None:
            first = res
        print("%-44s e2e median %7.3f ms  min %7.3f  (h2d %6.2f, kernels window %6.2f, d2h %5.2f)  -> %.2f M contigs/s  [%s]" % (
            g or "(defaults)", float(np.median(ts)), min(ts), st["ms_h2d"], st["ms_kernels"], st["ms_d2h"],
            batch.n_contigs / float(np.median(ts)) / 1e3, same), flush=True)
        eng.close()
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    pin.close()


if __name__ == "__main__":
    main()
