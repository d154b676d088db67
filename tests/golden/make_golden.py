"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes
  demo/                     the reference's demo inputs (blastout, two GFFs, taxonomy, contig
                            lengths) and its six shipped expected TSVs (demo/output*/)
  demo_records.json.gz       reference-harness records for the demo under several flag sets
  synth_records.json.gz      same for seeded synthetic inputs (regenerated at test time from
                            waafle_b200.synth with the recorded seeds; a checksum guards drift)
Floats are stored as hex strings (bit-exact).
"""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_harness as ref                      # noqa: E402
from oracle.validate_against_reference import FLAG_SETS          # noqa: E402
from waafle_b200 import synth                                    # noqa: E402

SYNTH_CASES = [
    dict(name="cfg2", config="cfg2", n_contigs=120, seed=101, over={}),
    dict(name="cfg3", config="cfg3", n_contigs=60, seed=102, over={}),
    dict(name="cfg5", config="cfg5", n_contigs=60, seed=103, over={}),
    dict(name="cfg4", config="cfg4", n_contigs=2, seed=104,
         over=dict(genes=[66, 70], hits_per_gene=10.0)),
]
SYNTH_FLAGS = [0, 1, 2, 3, 4, 6, 7, 14]      # indices into FLAG_SETS
DEMO_FLAGS = list(range(len(FLAG_SETS)))


def enc(rec):
    r = dict(rec)
    for k in ("crit", "rank"):
        if k in r:
            r[k] = float(r[k]).hex()
    r.pop("gene_scores0", None)
    return r


def checksum(batch):
    h = hashlib.sha256()
    for k in sorted(batch.arrays()):
        h.update(batch.arrays()[k].tobytes())
    return h.hexdigest()


def main():
    demo_src = os.path.join(ref.REFERENCE_ROOT, "demo")
    demo_dst = os.path.join(HERE, "demo")
    os.makedirs(demo_dst, exist_ok=True)
    for rel in ["output/demo_contigs.blastout", "output/demo_contigs.gff", "input/demo_taxonomy.tsv",
                "output_prodigal/demo_contigs.prodigal.gff",
                "output/demo_contigs.lgt.tsv", "output/demo_contigs.no_lgt.tsv",
                "output/demo_contigs.unclassified.tsv",
                "output_prodigal/demo_contigs.prodigal.lgt.tsv",
                "output_prodigal/demo_contigs.prodigal.no_lgt.tsv",
                "output_prodigal/demo_contigs.prodigal.unclassified.tsv"]:
        shutil.copy(os.path.join(demo_src, rel), os.path.join(demo_dst, os.path.basename(rel)))
        os.chmod(os.path.join(demo_dst, os.path.basename(rel)), 0o644)
    sys.path.insert(0, ref.REFERENCE_ROOT)
    wu, _ = ref._import()
    lengths = wu.read_contig_lengths(os.path.join(demo_src, "input", "demo_contigs.fna"))
    with open(os.path.join(demo_dst, "demo_contigs.lengths.tsv"), "w") as fh:
        for k, v in lengths.items():
            fh.write("{}\t{}\n".format(k, v))

    out = {}
    for prodigal in (False, True):
        files = dict(contigs=os.path.join(demo_src, "input", "demo_contigs.fna"),
                     blastout=os.path.join(demo_src, "output", "demo_contigs.blastout"),
                     gff=os.path.join(demo_src, "output_prodigal" if prodigal else "output",
                                      "demo_contigs.prodigal.gff" if prodigal else "demo_contigs.gff"),
                     taxonomy=os.path.join(demo_src, "input", "demo_taxonomy.tsv"))
        for fi in DEMO_FLAGS:
            recs = ref.run_reference(files["contigs"], files["blastout"], files["gff"],
                                     files["taxonomy"], ref.make_args(**FLAG_SETS[fi]))
            out["{}:{}".format("prodigal" if prodigal else "genecaller", fi)] = {
                k: enc(v) for k, v in recs.items()}
    with gzip.open(os.path.join(HERE, "demo_records.json.gz"), "wt") as fh:
        json.dump(dict(flag_sets=FLAG_SETS, records=out), fh, sort_keys=True)

    sout = {}
    with tempfile.TemporaryDirectory() as tmp:
        for case in SYNTH_CASES:
            over = {k: tuple(v) if isinstance(v, list) else v for k, v in case["over"].items()}
            data = synth.generate_config(case["config"], n_contigs=case["n_contigs"],
                                         seed=case["seed"], **over)
            files = data.write_files(tmp, case["name"])
            entry = dict(case=case, checksum=checksum(data.to_batch()), records={})
            for fi in SYNTH_FLAGS if case["name"] != "cfg4" else SYNTH_FLAGS[:3]:
                recs = ref.run_reference(files["contigs"], files["blastout"], files["gff"],
                                         files["taxonomy"], ref.make_args(**FLAG_SETS[fi]))
                entry["records"][str(fi)] = {k: enc(v) for k, v in recs.items()}
            sout[case["name"]] = entry
    with gzip.open(os.path.join(HERE, "synth_records.json.gz"), "wt") as fh:
        json.dump(dict(flag_sets=FLAG_SETS, cases=sout), fh, sort_keys=True)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
