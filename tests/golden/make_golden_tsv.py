"""Golden TSV BYTES from the CURRENT, unmodified reference CLI (waafle/waafle_orgscorer.py run as a script).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_tsv.py
Writes tests/golden/demo_cli/<gff>_<k>/demo_contigs.{lgt,no_lgt,unclassified}.tsv for the demo under both GFFs and
the flag sets below (tests/test_cli_gpu.py::test_cli_bytes_equal_reference_cli compares the drop-in CLI with them).
Rank ties are hash-seed dependent upstream; the demo has none, which the generator checks by running
every case under two PYTHONHASHSEEDs.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

REF = "/root/reference"
CLI_FLAG_SETS = [
    [],
    ["--weak-loci", "assign-unknown"],
    ["--weak-loci", "penalize", "--range", "0.2"],
    ["--sister-penalty", "lenient", "--ambiguous-threshold", "strict", "--disambiguate-two", "jump"],
]


DETAILS_FLAG_SETS = [{}, dict(weak_loci="assign-unknown"), dict(jump_taxonomy=1, weak_loci="penalize")]
DETAILS_CLI = [[], ["--weak-loci", "assign-unknown"], ["--jump-taxonomy", "1", "--weak-loci", "penalize"]]


def main():
    out_root = os.path.join(HERE, "demo_cli")
    shutil.rmtree(out_root, ignore_errors=True)
    os.makedirs(out_root)
    tmp = tempfile.mkdtemp()
    for gff in ("genecaller", "prodigal"):
        files = helpers.demo_files(tmp, gff == "prodigal")
        for k, flags in enumerate(CLI_FLAG_SETS):
            texts = []
            for seed in ("1", "2"):
                outdir = tempfile.mkdtemp()
                env = dict(os.environ, PYTHONPATH=REF, PYTHONHASHSEED=seed)
                subprocess.run([sys.executable, os.path.join(REF, "waafle", "waafle_orgscorer.py"), files["contigs"],
                                files["blastout"], files["gff"], files["taxonomy"], "--outdir", outdir,
                                "--basename", "demo_contigs"] + flags, check=True, env=env, capture_output=True)
                texts.append({kind: open(os.path.join(outdir, "demo_contigs.{}.tsv".format(kind))).read()
                              for kind in ("lgt", "no_lgt", "unclassified")})
                shutil.rmtree(outdir)
            assert texts[0] == texts[1], "hash-seed dependent output: {} {}".format(gff, flags)
            dst = os.path.join(out_root, "{}_{}".format(gff, k))
            os.makedirs(dst)
            for kind, text in texts[0].items():
                with open(os.path.join(dst, "demo_contigs.{}.tsv".format(kind)), "w") as fh:
                    fh.write(text)
    # --write-details: the unmodified write_details through the harness (text handle; canonical clade order)
    import gzip
    import io
    sys.path.insert(0, ROOT)
    from oracle import reference_harness as rh
    for gff in ("genecaller", "prodigal"):
        files = helpers.demo_files(tmp, gff == "prodigal")
        for k, over in enumerate(DETAILS_FLAG_SETS):
            buf = io.StringIO()
            rh.run_reference(files["contigs"], files["blastout"], files["gff"], files["taxonomy"], rh.make_args(**over),
                             details=buf)
            with gzip.open(os.path.join(out_root, "details_{}_{}.tsv.gz".format(gff, k)), "wt") as fh:
                fh.write(buf.getvalue())
    with open(os.path.join(out_root, "details_flag_sets.json"), "w") as fh:
        json.dump(DETAILS_FLAG_SETS, fh)
    with open(os.path.join(out_root, "flag_sets.json"), "w") as fh:
        json.dump(CLI_FLAG_SETS, fh)
    print("wrote", out_root)


if __name__ == "__main__":
    main()
