"""
TEST INFRASTRUCTURE ONLY -- pins oracle/orgscorer_oracle.py against the unmodified reference.

Run here (build container, /root/reference present):
    python oracle/validate_against_reference.py            # demo + synthetic, many flag sets
It feeds the same text files to (a) the reference classes (oracle/reference_harness.py) and
(b) the product front end (parsers -> packer) followed by the numpy oracle, and demands
equality of every output field with bit-exact floats.  `compare_records` is also what
tests/ uses to check the CUDA engine's records against golden reference records.
"""

import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import orgscorer_oracle as oracle   # noqa: E402
from oracle import reference_harness as ref     # noqa: E402

FLAG_SETS = [
    {},
    dict(sister_penalty="off", ambiguous_threshold="strict"),
    dict(weak_loci="penalize"),
    dict(weak_loci="assign-unknown"),
    dict(jump_taxonomy=1, clade_genes=2, clade_leaves=2),
    dict(jump_taxonomy=2),
    dict(one_clade_threshold=0.9, two_clade_threshold=0.6),
    dict(disambiguate_one="report-best", disambiguate_two="report-best"),
    dict(disambiguate_two="jump", range=0.2),
    dict(allow_lca=True, range=0.3),
    dict(stranded=True, annotation_threshold="strict"),
    dict(sister_penalty="lenient", ambiguous_threshold="off", annotation_threshold="off"),
    dict(min_overlap=0.5, min_scov=0.9, min_gene_length=500.0),
    dict(range=0.5, ambiguous_fraction=0.5),
    dict(weak_loci="assign-unknown", range=0.3, sister_penalty="off"),
    dict(weak_loci="penalize", one_clade_threshold=0.3, two_clade_threshold=0.5, range=0.2),
]


def params_from_args(args, n_systems):
    from waafle_b200.params import OrgscorerParams
    return OrgscorerParams.from_args(args, n_systems)


def frontend_pack(files):
    """Parse text inputs with the product front end and pack them."""
    from waafle_b200 import packing, parsers, taxonomy, utils
    hits = parsers.read_blast_hits(files["blastout"])
    loci = parsers.read_gff_loci(files["gff"])
    tax = taxonomy.Taxonomy(files["taxonomy"]).build(set(hits.taxon))
    lengths = utils.read_contig_lengths(files["contigs"])
    batch = packing.pack(lengths, loci, hits, tax)
    return batch, loci, hits, tax


def records_from_results(batch, loci, hits, tax, res):
    from waafle_b200 import writer
    return writer.build_records(batch, loci, hits, tax, res)


def compare_records(ref_records, records, exact_scores=True, rtol=1e-12):
    """Compare reference-harness records with front-end records; returns a list of diffs."""
    diffs = []
    for r in records:
        name = r["contig_name"]
        g = ref_records[name]

        def chk(field, a, b):
            if a != b:
                diffs.append((name, field, a, b))

        chk("call", g["call"], r["call"])
        chk("length", g["length"], r["contig_length"])
        chk("loci", "|".join(g["loci"]), r["loci"])
        chk("annotations", g["annotations"], r["annotations"])
        if g["call"] != r["call"]:
            continue
        if g["call"] != "unclassified":
            chk("ignore", g["ignore"], r["ignore"])
            chk("synteny", g["synteny"], r["synteny"])
            for gk, rk in (("crit", "min_score"), ("rank", "avg_score")):
                b = r.get(rk, r.get(rk.replace("_score", "_max_score")))
                a = g[gk]
                if exact_scores:
                    chk(gk, a.hex(), b.hex())
                elif abs(a - b) > rtol * max(abs(a), abs(b)):
                    diffs.append((name, gk, a, b))
        if g["call"] == "no_lgt":
            chk("clade", g["clade1"], r["clade"])
            chk("melded", g["melded1"], r["melded"])
            chk("taxonomy", "|".join(g["taxonomy1"]), r["taxonomy"])
        elif g["call"] == "lgt":
            chk("clade_A", g["clade1"], r["clade_A"])
            chk("clade_B", g["clade2"], r["clade_B"])
            chk("lca", g["lca"], r["lca"])
            chk("direction", g["direction"], r["direction"])
            chk("melded_A", g["melded1"], r["melded_A"])
            chk("melded_B", g["melded2"], r["melded_B"])
    if len(records) != len(ref_records):
        diffs.append(("*", "n_contigs", len(ref_records), len(records)))
    return diffs


def compare_gene_scores(ref_records, batch, tax, res):
    """Level-0 gene-score matrices, bit-exact (K2 check)."""
    diffs = []
    for c, name in enumerate(batch.contig_names):
        g = ref_records[name]["gene_scores0"]
        o = {tax.names[k]: [float(x) for x in v] for k, v in res["gene_scores"][c].items()}
        if g != o:
            diffs.append((name, "gene_scores0"))
    return diffs


def validate(files, flags, verbose=True):
    args = ref.make_args(**flags)
    ref_records = ref.run_reference(files["contigs"], files["blastout"], files["gff"],
                                    files["taxonomy"], args)
    batch, loci, hits, tax = frontend_pack(files)
    P = params_from_args(args, len(hits.systems))
    res = oracle.score_batch(P.as_dict(), tax.tables(), batch.arrays(), want_gene_scores=True)
    records = records_from_results(batch, loci, hits, tax, res)
    diffs = compare_records(ref_records, records) + compare_gene_scores(ref_records, batch, tax, res)
    if verbose:
        calls = [r["call"] for r in records]
        print("  flags={} contigs={} lgt/no_lgt/uncl={}/{}/{} max_lifts={} diffs={}".format(
            flags, len(records), calls.count("lgt"), calls.count("no_lgt"),
            calls.count("unclassified"), max(r["lifts"] for r in records), len(diffs)))
        for d in diffs[:10]:
            print("    DIFF", d)
    return diffs


def demo_files(prodigal=False):
    d = os.path.join(ref.REFERENCE_ROOT, "demo")
    return dict(contigs=os.path.join(d, "input", "demo_contigs.fna"),
                blastout=os.path.join(d, "output", "demo_contigs.blastout"),
                gff=os.path.join(d, "output_prodigal", "demo_contigs.prodigal.gff") if prodigal
                else os.path.join(d, "output", "demo_contigs.gff"),
                taxonomy=os.path.join(d, "input", "demo_taxonomy.tsv"))


def main():
    from waafle_b200 import synth
    total = 0
    for prodigal in (False, True):
        print("demo (prodigal GFF)" if prodigal else "demo (genecaller GFF)")
        for flags in FLAG_SETS:
            total += len(validate(demo_files(prodigal), flags))
    with tempfile.TemporaryDirectory() as tmp:
        cases = [
            ("cfg2-shaped", synth.generate_config("cfg2", n_contigs=150, seed=11)),
            ("cfg3-shaped", synth.generate_config("cfg3", n_contigs=80, seed=12)),
            ("cfg5-shaped", synth.generate_config("cfg5", n_contigs=80, seed=13)),
            ("cfg4-shaped", synth.generate_config("cfg4", n_contigs=2, seed=14,
                                                  genes=(64, 70), hits_per_gene=12.0)),
        ]
        for name, data in cases:
            print(name)
            files = data.write_files(tmp, name)
            for flags in FLAG_SETS[:8] if name != "cfg4-shaped" else FLAG_SETS[:2]:
                total += len(validate(files, flags))
    print("TOTAL DIFFS", total)
    return 1 if total else 0


if __name__ == "__main__":
    sys.exit(main())
