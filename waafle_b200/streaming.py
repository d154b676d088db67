"""Streaming, multi-GPU drive of the orgscorer engine (SURVEY 7.2 step 8 / BASELINE configs[4]).

The reference holds every contig, hit and output row in memory and scores one contig after the other
(waafle/waafle_orgscorer.py:905-960).  Contigs are independent (:943-960) and the blastout is grouped by query
(waafle/utils.py:255-258), so a run of any size is cut into contig-aligned byte ranges of the blastout file:

    parent   scan_blast_chunks()            byte ranges that start where the query name changes
    workers  one process per GPU            read a range -> parse it on the GPU -> pack with the contigs' loci -> Engine.score_batch
                                            -> rows -> per-chunk TSV shards (sorted by contig name)
    parent   merge_shards()                 k-way merge of the shards by contig name (:838) under ONE header whose
                                            annotation columns are the union over all chunks (:824-831)

The engine, the packer and the row builder are the ones of the single-call path (orgscorer.main); the outputs are
byte-identical to it (tests/test_streaming.py on CPU for the merge, tests/test_cli_gpu.py on the GPU end to end).
"""

import heapq
import os

import numpy as np

from . import packing, parsers, writer
from .utils import die, say, write_row

WINDOW = 1 << 20


# ---------------------------------------------------------------------------------------------------------------
# contig-aligned chunks of a blastout file
# ---------------------------------------------------------------------------------------------------------------

def _first_query_change(buf, carry_name):
    """Offset in `buf` (which starts at a row start) of the first row whose query differs from the row before it
    (`carry_name`: query of the row just before `buf`, or None).  -1 if there is none in the complete rows of `buf`."""
    pos, prev = 0, carry_name
    while True:
        nl = buf.find(b"\n", pos)
        if nl < 0:
            return -1, prev
        tab = buf.find(b"\t", pos, nl)
        name = buf[pos:tab if tab >= 0 else nl]
        if prev is not None and name != prev and nl > pos:
            return pos, prev
        if nl > pos:
            prev = name
        pos = nl + 1


def scan_blast_chunks(path, target_bytes):
    """[(offset, length)] covering the file; every range but the first starts at a row whose query name differs from the
    previous row's, and ranges are ~target_bytes long (a single contig's block is never split)."""
    size = os.path.getsize(path)
    chunks, start = [], 0
    with open(path, "rb") as fh:
        while start < size:
            want = start + target_bytes
            if want >= size:
                chunks.append((start, size - start))
                break
            # row start at or after `want`
            fh.seek(want - 1)
            buf = fh.read(WINDOW)
            nl = buf.find(b"\n")
            while nl < 0 and len(buf) < size - want + 1:
                more = fh.read(WINDOW)
                if not more:
                    break
                buf += more
                nl = buf.find(b"\n")
            if nl < 0:
                chunks.append((start, size - start))
                break
            row0 = want - 1 + nl + 1          # absolute offset of a row start
            # query of the row before row0
            back = min(row0 - start, WINDOW)
            fh.seek(row0 - back)
            tail = fh.read(back)
            prev_start = tail.rfind(b"\n", 0, len(tail) - 1) + 1
            prev_row = tail[prev_start:]
            tab = prev_row.find(b"\t")
            carry = prev_row[:tab] if tab >= 0 else prev_row.rstrip(b"\n")
            # first change of query at or after row0
            fh.seek(row0)
            pos_abs, cut = row0, -1
            while pos_abs < size:
                buf = fh.read(WINDOW)
                if not buf:
                    break
                last_nl = buf.rfind(b"\n")
                if last_nl < 0:
                    more = fh.read(WINDOW)
                    while more and b"\n" not in more:
                        buf += more
                        more = fh.read(WINDOW)
                    buf += more
                    last_nl = buf.rfind(b"\n")
                    if last_nl < 0:
                        break
                body = buf[:last_nl + 1]
                off, carry = _first_query_change(body, carry)
                if off >= 0:
                    cut = pos_abs + off
                    break
                pos_abs += len(body)
                fh.seek(pos_abs)
            if cut < 0 or cut >= size:
                chunks.append((start, size - start))
                break
            chunks.append((start, cut - start))
            start = cut
    return chunks


# ---------------------------------------------------------------------------------------------------------------
# loci by contig, built once
# ---------------------------------------------------------------------------------------------------------------

class LociIndex:
    """Rows of a LocusTable grouped by contig (file order kept inside a contig): the streaming packer looks up the loci
    of a chunk's contigs instead of scanning the whole GFF per chunk."""

    def __init__(self, loci, contig_lengths):
        import pandas as pd
        self.loci = loci
        codes, uniq = pd.factorize(np.asarray(loci.seqname, dtype=object)) if len(loci) else (np.zeros(0, np.int64), [])
        self.order = np.argsort(codes, kind="stable")
        self.off = np.zeros(len(uniq) + 1, dtype=np.int64)
        if len(loci):
            np.cumsum(np.bincount(codes, minlength=len(uniq)), out=self.off[1:])
        self.slot = {nm: k for k, nm in enumerate(uniq)}
        for nm in uniq:
            if nm not in contig_lengths:
                say("  Unknown contig in <{}> file".format("gff"), nm)   # OS:921-923

    def rows(self, names):
        parts = []
        for nm in names:
            k = self.slot.get(nm)
            if k is not None:
                parts.append(self.order[self.off[k]:self.off[k + 1]])
        return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)

    def subset(self, names):
        """(LocusTable of the contigs' loci, their row numbers in the full table)."""
        rows = self.rows(names)
        lt = self.loci
        sub = parsers.LocusTable.__new__(parsers.LocusTable)
        sub.seqname, sub.start, sub.end = lt.seqname[rows], lt.start[rows], lt.end[rows]
        sub.strand_str, sub.strand = lt.strand_str[rows], lt.strand[rows]
        return sub, rows


def chunk_contigs(hits, contig_lengths):
    """Names of the contigs a parsed chunk holds hits for (block order), known or not."""
    if getattr(hits, "block_names", None) is not None:
        return list(hits.block_names)
    q = hits.qseqid
    if len(q) == 0:
        return []
    starts = np.r_[True, q[1:] != q[:-1]]
    return list(q[starts])


def pack_chunk(names, contig_lengths, loci_index, hits, tax):
    """packing.pack restricted to the contigs `names` (FASTA-known ones); loci rows refer to the FULL locus table."""
    known = [nm for nm in dict.fromkeys(names) if nm in contig_lengths]
    sub_lengths = {nm: contig_lengths[nm] for nm in known}
    sub_loci, rows = loci_index.subset(known)
    batch = packing.pack(sub_lengths, sub_loci, hits, tax)
    batch.locus_row = rows[batch.locus_row] if len(rows) else batch.locus_row
    return batch


# ---------------------------------------------------------------------------------------------------------------
# shards and their merge
# ---------------------------------------------------------------------------------------------------------------

SHARD_SEP = "\x1f"   # shard rows: name, number of retained loci, annotation cells follow the fixed cells


def write_shard(records, prefix):
    """Rows of one chunk, sorted by contig name, as <prefix>.<call>.shard files.  First line: the chunk's annotation
    systems; every row: contig name, number of retained loci, the formatted fixed cells, then one cell per system."""
    systems = sorted({s for r in records for d in r["annotations"] for s in d})
    handles = {opt: open("{}.{}.shard".format(prefix, opt), "w") for opt in writer.FORMATS}
    for fh in handles.values():
        fh.write(SHARD_SEP.join(systems) + "\n")
    for r in sorted(records, key=lambda r: r["contig_name"]):
        opt = r["call"]
        cells = [writer.format_field(r[f]) for f in writer.FORMATS[opt]]
        ann = [writer.format_field("|".join(d.get(s, writer.C_MISSING_ANNOTATION) for d in r["annotations"])) for s in systems]
        handles[opt].write(SHARD_SEP.join([r["contig_name"], str(len(r["annotations"]))] + cells + ann) + "\n")
    for fh in handles.values():
        fh.close()
    return systems


def _shard_rows(path, n_fixed, all_systems):
    """(name, row cells) of one shard in name order; the annotation cells are laid out for the run-wide system list."""
    with open(path) as fh:
        head = fh.readline().rstrip("\n")
        systems = head.split(SHARD_SEP) if head else []
        same = systems == all_systems
        for line in fh:
            parts = line.rstrip("\n").split(SHARD_SEP)
            if same:   # the common case: every chunk saw the same systems
                yield parts[0], parts[2:]
                continue
            n_loci = int(parts[1])
            ann = dict(zip(systems, parts[2 + n_fixed:]))
            missing = writer.format_field("|".join([writer.C_MISSING_ANNOTATION] * n_loci))
            yield parts[0], parts[2:2 + n_fixed] + [ann.get(s, missing) for s in all_systems]


def shard_systems(prefixes):
    out = set()
    for p in prefixes:
        with open("{}.{}.shard".format(p, "lgt")) as fh:
            head = fh.readline().rstrip("\n")
            out.update(head.split(SHARD_SEP) if head else [])
    return sorted(out)


def merge_shards(prefixes, outdir, basename, remove=True):
    """k-way merge of the chunk shards by contig name into <basename>.{lgt,no_lgt,unclassified}.tsv (OS:814-894): one
    header with the union of the transferred annotation systems; a system a chunk never saw prints `None` per locus."""
    say("Initializing outputs.")
    systems = shard_systems(prefixes)
    for opt, fmt in writer.FORMATS.items():
        with open(os.path.join(outdir, ".".join([basename, opt, "tsv"])), "w") as out:
            write_row([k.upper() for k in fmt + [writer.C_ANNOTATION_PREFIX + s for s in systems]], out)
            streams = [_shard_rows("{}.{}.shard".format(p, opt), len(fmt), systems) for p in prefixes]
            last = None
            for name, cells in heapq.merge(*streams, key=lambda t: t[0]):
                if name == last:
                    die("blastout is not grouped by query sequence")   # a contig in two chunks
                last = name
                write_row(cells, out)
    if remove:
        for p in prefixes:
            for opt in writer.FORMATS:
                os.remove("{}.{}.shard".format(p, opt))


# ---------------------------------------------------------------------------------------------------------------
# workers
# ---------------------------------------------------------------------------------------------------------------

class ChunkScorer:
    """Everything one device needs to turn a piece of blastout text into shard rows."""

    def __init__(self, device, args, tax, contig_lengths, loci_index, params_factory, cpu_parse=False, engine_factory=None):
        if engine_factory is None:
            from .engine import Engine
            engine_factory = Engine
        self.device, self.args, self.tax = device, args, tax
        self.contig_lengths, self.loci_index = contig_lengths, loci_index
        self.params_factory = params_factory
        self.cpu_parse = cpu_parse
        self.engine = None
        self.Engine = engine_factory   # (device, params, taxonomy) -> object with score_batch / set_params / set_taxonomy / stats / close
        self.parser = None
        self.known_taxa = set()
        self.n_systems = None
        self.stats = dict(contigs=0, hits=0, ms_kernels=0.0, chunks=0)

    def parse(self, text):
        hits = None
        if not self.cpu_parse:
            from . import gpu_parse
            if self.parser is None:
                self.parser = gpu_parse.BlastParser(self.device)
            hits = self.parser.parse(text)
        if hits is None:   # rows the device parser does not reproduce: the CPU reader (reference error behaviour)
            import tempfile
            with tempfile.NamedTemporaryFile(suffix=".blastout", delete=False) as tmp:
                tmp.write(text)
            try:
                hits = parsers.read_blast_hits(tmp.name)
            finally:
                os.remove(tmp.name)
        return hits

    def score(self, names, hits):
        """records of the contigs `names` given their parsed hits (possibly none)."""
        if hits.sysmask is None:
            die("more than 32 annotation systems in the subject headers")
        taxa = hits.distinct_taxa()
        if self.tax.index is None or not taxa <= self.known_taxa:
            self.known_taxa |= taxa
            self.tax.build(self.known_taxa)
            tax_changed = True
        else:
            tax_changed = False
        batch = pack_chunk(names, self.contig_lengths, self.loci_index, hits, self.tax)
        n_sys = len(hits.systems)
        if self.engine is None:
            self.engine = self.Engine(self.device, self.params_factory(n_sys), self.tax)
            if getattr(self.args, "exact_scores", False):
                self.engine.set_option("exact", 1)
        else:
            if n_sys != self.n_systems:
                self.engine.set_params(self.params_factory(n_sys))
            if tax_changed:
                self.engine.set_taxonomy(self.tax)
        self.n_systems = n_sys
        res = self.engine.score_batch(batch)
        st = self.engine.stats()
        self.stats["contigs"] += batch.n_contigs
        self.stats["hits"] += batch.n_hits
        self.stats["ms_kernels"] += st["ms_kernels"]
        self.stats["chunks"] += 1
        return writer.build_records(batch, self.loci_index.loci, hits, self.tax, res), batch.contig_names

    def close(self):
        if self.engine is not None:
            self.engine.close()
        if self.parser is not None:
            self.parser.close()


def _worker(wid, device, path, tasks, results, shard_dir, args, tax, contig_lengths, loci_index, params_factory, engine_factory=None):
    try:
        scorer = ChunkScorer(device, args, tax, contig_lengths, loci_index, params_factory, cpu_parse=args.cpu_parse,
                             engine_factory=engine_factory)
        with open(path, "rb") as fh:
            while True:
                task = tasks.get()
                if task is None:
                    break
                cid, off, length = task
                fh.seek(off)
                text = fh.read(length)
                hits = scorer.parse(text)
                names = chunk_contigs(hits, contig_lengths)
                records, scored = scorer.score(names, hits)
                prefix = os.path.join(shard_dir, "chunk{:06d}".format(cid))
                write_shard(records, prefix)
                results.put(("chunk", cid, prefix, list(scored), None))
        results.put(("done", wid, None, None, dict(scorer.stats, device=device)))
        scorer.close()
    except SystemExit as exc:
        results.put(("error", device, None, None, "worker on cuda:{} exited: {}".format(device, exc)))
    except BaseException as exc:   # noqa: B902 -- the parent must hear about every failure
        import traceback
        results.put(("error", device, None, None, traceback.format_exc() or repr(exc)))


def run_streaming(args, tax, contig_lengths, loci, params_factory, devices, chunk_bytes, engine_factory=None):
    """The whole streamed run: returns (per-worker stats keyed by worker index, plus "rest"; number of chunks).  `devices`: CUDA device indices, one worker process each.  The
    CLI's parent process never initialises CUDA before this point, so the workers are plain forks sharing the contig /
    loci tables copy-on-write; a caller that already used CUDA in this process gets spawned workers (pickled tables)."""
    import multiprocessing as mp
    import shutil
    import tempfile
    from . import engine as _engine
    loci_index = LociIndex(loci, contig_lengths)
    chunks = scan_blast_chunks(args.blastout, chunk_bytes) if os.path.getsize(args.blastout) > 0 else []
    shard_dir = tempfile.mkdtemp(prefix="wfl_shards_", dir=args.outdir)
    ctx = mp.get_context("spawn" if _engine.cuda_touched() else "fork")
    tasks, results = ctx.Queue(), ctx.Queue()
    for cid, (off, length) in enumerate(chunks):
        tasks.put((cid, off, length))
    for _ in devices:
        tasks.put(None)
    procs = [ctx.Process(target=_worker, args=(w, d, args.blastout, tasks, results, shard_dir, args, tax, contig_lengths,
                                               loci_index, params_factory, engine_factory), daemon=True)
             for w, d in enumerate(devices)]
    for p in procs:
        p.start()
    prefixes, seen, stats, done = {}, set(), {}, 0
    try:
        import queue
        while done < len(devices):
            try:
                kind, a, b, c, d = results.get(timeout=5.0)
            except queue.Empty:
                dead = [p for p in procs if not p.is_alive() and p.exitcode not in (0, None)]
                if dead:   # a worker that died without a word (killed, crashed in native code)
                    die("streaming worker exited with code {}".format(dead[0].exitcode))
                continue
            if kind == "error":
                die(d)
            if kind == "done":
                stats[a] = d
                done += 1
                continue
            dup = seen.intersection(c)
            if dup:
                die("blastout is not grouped by query sequence")   # UT:255-258
            seen.update(c)
            prefixes[a] = b
        for p in procs:
            p.join()
        # contigs without a single hit: scored like the others (unclassified rows with their loci), in the parent's
        # process on the first device -- after the workers are gone
        rest = [nm for nm in contig_lengths if nm not in seen]
        if rest:
            scorer = ChunkScorer(devices[0], args, tax, contig_lengths, loci_index, params_factory, cpu_parse=True,
                                 engine_factory=engine_factory)
            empty = parsers.hits_from_columns(*([[]] * 10))
            step = max(1, int(args.chunk_contigs))
            for k in range(0, len(rest), step):
                records, _ = scorer.score(rest[k:k + step], empty)
                prefix = os.path.join(shard_dir, "rest{:06d}".format(k // step))
                write_shard(records, prefix)
                prefixes[len(chunks) + k // step] = prefix
            stats["rest"] = dict(scorer.stats, device=devices[0])
            scorer.close()
        merge_shards([prefixes[k] for k in sorted(prefixes)], args.outdir, args.basename)
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
        shutil.rmtree(shard_dir, ignore_errors=True)
    return stats, len(chunks)
