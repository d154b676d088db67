"""GPU: BLAST outfmt-6 parsing on the device (csrc/wfl_parse.cu) against the CPU reader
`parsers.read_blast_hits` -- itself pinned to the reference's Hit objects in tests/test_frontend.py -- bit for bit."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def assert_same_hits(cpu, gpu, n_ann=200):
    assert len(cpu) == len(gpu)
    for k in ("qstart", "qend", "strand", "sysmask"):
        assert np.array_equal(getattr(cpu, k), getattr(gpu, k)), k
    for k in ("score", "scov_modified"):   # same doubles, not just close
        assert np.array_equal(getattr(cpu, k).view(np.int64), getattr(gpu, k).view(np.int64)), k
    assert list(cpu.systems) == list(gpu.systems)
    assert np.array_equal(np.asarray(cpu.taxon, dtype=object), np.asarray(gpu.taxon, dtype=object))
    assert np.array_equal(np.asarray(cpu.qseqid, dtype=object), np.asarray(gpu.qseqid, dtype=object))
    assert cpu.distinct_taxa() == gpu.distinct_taxa()
    rng = np.random.default_rng(0)
    for r in rng.integers(0, len(cpu), size=min(n_ann, len(cpu))):
        assert cpu.sseqid_annotations[int(cpu.sseqid_id[r])] == gpu.sseqid_annotations[int(r)]
        assert cpu.sseqid_names[int(cpu.sseqid_id[r])] == gpu.sseqid_names[int(r)]


def test_demo_blastout_parsed_on_gpu():
    from waafle_b200 import parsers
    path = os.path.join(helpers.GOLDEN, "demo", "demo_contigs.blastout")
    cpu = parsers.read_blast_hits(path)
    gpu = parsers.read_blast_hits(path, device=0)
    assert gpu.block_starts is not None, "the GPU parser did not run"
    assert_same_hits(cpu, gpu)


@pytest.mark.parametrize("annotations", [False, True])
def test_million_row_synthetic_blastout(tmp_path, annotations):
    """~1M rows incl. minus-strand hits, scov_modified > 1, hits failing --min-scov, unlisted taxa, integer pident."""
    from waafle_b200 import parsers, synth
    data = synth.generate_config("cfg2", n_contigs=4000, seed=91, annotations=annotations)
    files = data.write_files(str(tmp_path), "big")
    # every 97th row gets a subject shorter than its alignment: scov_modified > 1 (utils.py:227)
    rows = open(files["blastout"]).read().split("\n")
    for r in range(0, len(rows) - 1, 97):
        f = rows[r].split("\t")
        f[3] = str(max(1, abs(int(f[8]) - int(f[7])) - 2))
        rows[r] = "\t".join(f)
    with open(files["blastout"], "w") as fh:
        fh.write("\n".join(rows))
    cpu = parsers.read_blast_hits(files["blastout"])
    assert len(cpu) > 900_000 and (cpu.scov_modified > 1).any() and (cpu.strand == ord("-")).any()
    gpu = parsers.read_blast_hits(files["blastout"], device=0)
    assert gpu.block_starts is not None and len(gpu.block_starts) <= 4000
    assert_same_hits(cpu, gpu)
    # timing of a second parse: the first one pays the CUDA module load of the parser kernels
    t = parsers.read_blast_hits(files["blastout"], device=0).parse_times
    rate = t["rows"] / ((t["ms_h2d"] + t["ms_kernels"] + t["ms_d2h"]) * 1e-3)
    print("GPU parse: {:.1f} M hits/s (H2D {:.1f} ms, kernels {:.1f} ms, D2H {:.1f} ms, {} rows)".format(
        rate / 1e6, t["ms_h2d"], t["ms_kernels"], t["ms_d2h"], t["rows"]))
    assert rate > 20e6


def test_rows_the_device_cannot_reproduce_fall_back(tmp_path):
    """Quoted fields / exponent floats / short rows flag the file: read_blast_hits then takes the CPU path (which has the
    reference's error behaviour); the GPU parser itself returns None."""
    from waafle_b200 import gpu_parse, parsers
    good = open(os.path.join(helpers.GOLDEN, "demo", "demo_contigs.blastout"), "rb").read()
    rows = good.split(b"\n")
    p = gpu_parse.BlastParser(0)
    f = rows[3].split(b"\t")
    f[9] = b"9.5e1"
    assert p.parse(b"\n".join(rows[:3] + [b"\t".join(f)] + rows[4:])) is None
    assert p.parse(b"\n".join(rows[:3] + [b"\t".join(rows[3].split(b"\t")[:14])] + rows[4:])) is None
    ok = p.parse(good)
    assert ok is not None and len(ok) == len([r for r in rows if r])
    assert p.parse(good.rstrip(b"\n")) is not None   # unterminated last row
    p.close()
    path = tmp_path / "exp.blastout"
    path.write_bytes(b"\n".join(rows[:3] + [b"\t".join(f)] + rows[4:]))
    hits = parsers.read_blast_hits(str(path), device=0)
    assert hits.block_starts is None and len(hits) == len(ok)   # CPU reader took over
    assert hits.score[3] == ok.scov_modified[3] * 95.0 / 100.0   # pident "9.5e1" read as 95.0 (utils.py:229)
    assert np.array_equal(np.delete(hits.score, 3), np.delete(ok.score, 3))


def test_cli_gpu_parse_equals_cpu_parse(tmp_path):
    from waafle_b200 import orgscorer
    files = helpers.demo_files(tmp_path, True)
    outs = {}
    for tag, extra in (("gpu", []), ("cpu", ["--cpu-parse"])):
        outdir = tmp_path / tag
        outdir.mkdir()
        orgscorer.main([files["contigs"], files["blastout"], files["gff"], files["taxonomy"],
                        "--outdir", str(outdir), "--basename", "run", "--quiet"] + extra)
        outs[tag] = outdir
    for kind in ("lgt", "no_lgt", "unclassified"):
        with open(os.path.join(str(outs["gpu"]), "run.{}.tsv".format(kind))) as f1, \
                open(os.path.join(str(outs["cpu"]), "run.{}.tsv".format(kind))) as f2:
            assert f1.read() == f2.read(), kind


def test_gff_parsed_on_gpu(tmp_path):
    """GFF rows on the device == the CPU reader's LocusTable: the demo GFFs (CRLF rows of waafle_genecaller, Prodigal rows
    with '#' comments) and a large synthetic one; a row the device does not reproduce falls back to the CPU reader."""
    from waafle_b200 import gpu_parse, parsers, synth
    paths = [os.path.join(helpers.GOLDEN, "demo", "demo_contigs.gff"), os.path.join(helpers.GOLDEN, "demo", "demo_contigs.prodigal.gff")]
    data = synth.generate_config("cfg5", n_contigs=20000, seed=3)
    paths.append(data.write_files(str(tmp_path), "big")["gff"])
    for path in paths:
        cpu = parsers.read_gff_loci(path)
        gpu = parsers.read_gff_loci(path, device=0)
        assert hasattr(gpu, "parse_times"), "the GPU parser did not run"
        assert len(cpu) == len(gpu) and len(cpu) > 0
        for k in ("start", "end", "strand"):
            assert np.array_equal(getattr(cpu, k), getattr(gpu, k)), (path, k)
        assert np.array_equal(cpu.seqname, gpu.seqname) and np.array_equal(cpu.strand_str, gpu.strand_str)
        assert [cpu.code(j) for j in range(0, len(cpu), 97)] == [gpu.code(j) for j in range(0, len(gpu), 97)]
    t = gpu.parse_times
    print("GPU GFF parse: {} rows, H2D {:.2f} ms, kernels {:.2f} ms, D2H {:.2f} ms".format(t["rows"], t["ms_h2d"], t["ms_kernels"], t["ms_d2h"]))
    p = gpu_parse.BlastParser(0)
    good = open(paths[0], "rb").read()
    rows = good.split(b"\n")
    f = rows[2].split(b"\t")
    f[3] = b"."
    assert p.parse_gff(b"\n".join(rows[:2] + [b"\t".join(f)] + rows[3:])) is None      # non-integer coordinate
    assert p.parse_gff(b"\n".join(rows[:2] + [b"\t".join(f[:8])] + rows[3:])) is None    # 8 fields
    assert len(p.parse_gff(good)) == len([r for r in rows if r.strip()])
    p.close()
