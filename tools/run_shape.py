"""One BASELINE shape on one GPU with full parity: python tools/run_shape.py <workload> <contigs> [steps]
Prints one JSON line: device-timed contigs/s (resident), end-to-end contigs/s (plugin call, pinned host buffers), parity of
EVERY contig against the C restatement of the reference (oracle/orgscorer_oracle.c), engine stats."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from waafle_b200 import synth                                  # noqa: E402
from waafle_b200.engine import Engine, PinnedArena             # noqa: E402
from waafle_b200.params import OrgscorerParams                 # noqa: E402
from helpers import compare_results                            # noqa: E402
from oracle import c_oracle                                    # noqa: E402


def main():
    workload, n = sys.argv[1], int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    data = synth.generate_config(workload, n_contigs=n, seed=1000)
    tax = data.taxonomy()
    batch = data.to_batch(tax).sort_hits()
    P = OrgscorerParams(n_systems=1 if batch.hit_sysmask is not None else 0)
    t = time.perf_counter()
    ref = c_oracle.score_batch(P, tax, batch, threads=os.cpu_count() or 1)
    t_c = time.perf_counter() - t
    eng = Engine(0, P, tax)
    eng.upload(batch)
    for _ in range(2):
        eng.run_resident()
    ms = 0.0
    for _ in range(steps):
        eng.run_resident()
        ms += eng.stats()["ms_kernels"]
    res = eng.download()
    st = eng.stats()
    diffs = compare_results(ref, res, score_rtol=1e-12)
    pin = PinnedArena()
    packed = batch.can_pack(len(tax.tables()["parent"]), P.n_systems)
    wire = batch.to_packed(P.min_scov) if packed else batch.arrays()
    harr = {k: pin.like(np.ascontiguousarray(v)) for k, v in wire.items()}
    eng.use_pinned_results(True)
    for _ in range(2):
        out = eng.score_batch(harr)
    t = time.perf_counter()
    for _ in range(steps):
        out = eng.score_batch(harr)
    e2e_ms = 1e3 * (time.perf_counter() - t) / steps
    diffs2 = compare_results(ref, {k: np.array(v) for k, v in out.items()}, score_rtol=1e-12)
    print(json.dumps({
        "workload": workload, "contigs": batch.n_contigs, "hits": int(batch.n_hits), "loci": int(batch.n_loci),
        "value": batch.n_contigs / (ms / steps * 1e-3), "ms_per_step": ms / steps,
        "e2e": {"value": batch.n_contigs / (e2e_ms * 1e-3), "ms_per_step": e2e_ms,
                "wire_format": "packed" if packed else "wide 29 B/hit", "h2d_bytes": int(sum(v.nbytes for v in harr.values()))},
        "parity": {"checked_contigs": batch.n_contigs, "against": "oracle/orgscorer_oracle.c", "bit_exact": not diffs and not diffs2,
                   "score_rtol": 1e-12, "diffs": [str(d)[:120] for d in (diffs + diffs2)[:3]]},
        "c_port": {"value": batch.n_contigs / t_c, "seconds": t_c, "cores": os.cpu_count()},
        "engine_stats": {k: st[k] for k in ("levels", "pairs_tested", "pairs_scored", "smem_contigs", "fallback_contigs",
                                            "second_pass_contigs", "guard_trips", "refined_groups", "workspace_retries",
                                            "kernel_launches")}}))
    eng.close()
    pin.close()


if __name__ == "__main__":
    main()
