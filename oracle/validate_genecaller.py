"""Pin oracle/genecaller_oracle.py to the UNMODIFIED reference (build container only: needs /root/reference):
random interval sets through waafle_genecaller.overlap_intervals + the length filter, and the reference's own demo GFF.
Also writes the committed fixture tests/golden/genecaller_cases.json.gz (inputs + the reference's outputs)."""
import gzip
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import genecaller_oracle as oracle   # noqa: E402


def reference_call(intervals, min_overlap, min_len):
    from waafle import waafle_genecaller as gc
    merged = gc.overlap_intervals([list(t) for t in intervals], min_overlap, False)
    return [tuple(m) for m in merged if m[1] - m[0] + 1 >= min_len]


def random_case(rng):
    n = rng.choice([0, 1, 2, 5, 20, 60, 150])
    genes = [(rng.randint(1, 8000), rng.randint(150, 1500)) for _ in range(rng.randint(1, 6))]
    iv = []
    for _ in range(n):
        g0, gl = rng.choice(genes)
        a = g0 + rng.randint(-200, 200)
        b = a + max(1, gl + rng.randint(-gl + 1, 300))
        if rng.random() < 0.3:
            a, b = b, a
        iv.append((max(1, a), max(1, b), rng.choice("+-")))
    return iv


def main():
    rng = random.Random(20261018)
    cases, diffs = [], 0
    for k in range(400):
        iv = random_case(rng)
        thr = rng.choice([0.1, 0.1, 0.1, 0.0, 0.5, 0.9, 1.0])
        ml = rng.choice([200.0, 200.0, 0.0, 1000.0])
        want = reference_call(iv, thr, ml)
        got = oracle.call_genes(iv, thr, ml)
        diffs += want != got
        cases.append(dict(intervals=iv, min_overlap=thr, min_gene_length=ml, genes=want))
    print("random cases:", len(cases), "diffs:", diffs)
    with gzip.open(os.path.join(ROOT, "tests", "golden", "genecaller_cases.json.gz"), "wt") as fh:
        json.dump(cases, fh)
    print("TOTAL DIFFS", diffs)
    return diffs


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
