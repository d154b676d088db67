"""Profiling driver: one resident batch, a few runs; prints per-phase cycle shares.

usage: python tools/prof_run.py [workload] [contigs] [runs] [name=value engine options ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waafle_b200 import synth                     # noqa: E402
from waafle_b200.engine import Engine             # noqa: E402
from waafle_b200.params import OrgscorerParams    # noqa: E402

# per-phase cycles of the exact pipeline; only filled when the library is built with -DWFL_PROFILE
PHASES = ["prepare/match+fill", "clade table", "regroup", "scores (K2..masks)", "weak/masks", "one-clade",
          "two-clade(+lift)", "lift", "output"]


def main():
    a = sys.argv[1:]
    workload = a[0] if len(a) > 0 else "cfg2"
    n = int(a[1]) if len(a) > 1 else 5000
    runs = int(a[2]) if len(a) > 2 else 3
    opts = [x.split("=") for x in a[3:]]
    data = synth.generate_config(workload, n_contigs=n, seed=1000)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    if os.environ.get("WFL_PROF_UNSORTED") != "1":   # default: what the front end's packer delivers
        batch = batch.sort_hits()
    P = OrgscorerParams(n_systems=1 if batch.hit_sysmask is not None else 0)
    eng = Engine(0, P, tax)
    for k, v in opts:
        eng.set_option(k, int(v))
    eng.upload(batch)
    for _ in range(runs):
        eng.run_resident()
        st = eng.stats()
    tot = sum(st["phase_cycles"]) or 1
    print("contigs %d hits %d  score kernel %.3f ms  all kernels %.3f ms  -> %.0f contigs/s" % (
        batch.n_contigs, batch.n_hits, st["ms_score_kernel"], st["ms_kernels"],
        batch.n_contigs / st["ms_kernels"] * 1e3))
    print("cycles/contig %.0f" % (tot / batch.n_contigs))
    for name, cyc in zip(PHASES, st["phase_cycles"]):
        print("  %-20s %6.2f%%  %10.0f cyc/contig" % (name, 100.0 * cyc / tot, cyc / batch.n_contigs))
    print({k: v for k, v in st.items() if k != "phase_cycles"})
    eng.close()


if __name__ == "__main__":
    main()
