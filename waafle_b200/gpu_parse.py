"""BLAST outfmt-6 text parsed ON THE GPU (csrc/wfl_parse.cu) into the hit columns the packer ships.

Same values as `parsers.read_blast_hits` -- i.e. as the reference's `Hit` objects (waafle/utils.py:192-241) -- but the
per-row work (field split, integer / decimal conversion, scov_modified and waafle_score arithmetic, dictionary coding
of the taxon and of the annotation systems, contig block detection) runs as CUDA kernels over the raw bytes; the host
only touches the DISTINCT names (a few thousand taxa, a handful of systems, one qseqid per contig).  Rows the device
parser cannot reproduce exactly (quoting, exponent floats, malformed rows) make `parse` return None: the caller then
uses the CPU reader, which has the reference's error behaviour.  No CPU fallback is hidden in here.
"""

import ctypes

import numpy as np

from .engine import EngineError, load_library, mark_cuda_touched

TAX_SLOTS, SYS_SLOTS = 1 << 20, 64
_ready = False


def _lib():
    global _ready
    lib = load_library()
    if not _ready:
        P = ctypes.POINTER
        lib.wfl_parser_create.argtypes = [ctypes.c_int, P(ctypes.c_void_p)]
        lib.wfl_parser_destroy.argtypes = [ctypes.c_void_p]
        lib.wfl_parser_destroy.restype = None
        lib.wfl_parser_last_error.argtypes = [ctypes.c_void_p]
        lib.wfl_parser_last_error.restype = ctypes.c_char_p
        lib.wfl_parse_blast.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64, P(ctypes.c_int32), P(ctypes.c_int64)]
        lib.wfl_parse_blast.restype = ctypes.c_int64
        lib.wfl_parse_distinct.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_int32]
        lib.wfl_parse_fetch.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 13
        lib.wfl_parser_times.argtypes = [ctypes.c_void_p, P(ctypes.c_float), P(ctypes.c_float), P(ctypes.c_float)]
        lib.wfl_parse_gff.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64, P(ctypes.c_int32), P(ctypes.c_int64)]
        lib.wfl_parse_gff.restype = ctypes.c_int64
        lib.wfl_parse_gff_fetch.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 7
        _ready = True
    return lib


class _TextNames:
    """Subject headers by row, sliced out of the file text on demand (the writer only asks for the winning hits)."""

    def __init__(self, text, off, length):
        self._text, self._off, self._len = text, off, length

    def __len__(self):
        return len(self._off)

    def __getitem__(self, i):
        o = int(self._off[int(i)])
        return self._text[o:o + int(self._len[int(i)])].decode()


class BlastParser:
    """One parser handle on one GPU."""

    def __init__(self, device=0):
        self._lib = _lib()
        h = ctypes.c_void_p()
        mark_cuda_touched()
        rc = self._lib.wfl_parser_create(int(device), ctypes.byref(h))
        if rc != 0 or not h:
            raise EngineError("wfl_parser_create(device={}) failed with {}: no usable CUDA device".format(device, rc))
        self._h = h
        self.times = {}

    def close(self):
        if getattr(self, "_h", None):
            self._lib.wfl_parser_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise EngineError("waafle_b200 parser error {}: {}".format(rc, self._lib.wfl_parser_last_error(self._h).decode()))

    def parse(self, text):
        """`text`: bytes of a whole blastout file (or of whole rows of it).  Returns a parsers.HitTable, or None if some
        row needs the CPU reader."""
        from . import parsers
        if not isinstance(text, (bytes, bytearray)):
            text = bytes(text)
        if len(text.strip()) == 0:
            return parsers.hits_from_columns(*([[]] * 10))
        flagged, first = ctypes.c_int32(), ctypes.c_int64()
        n = self._lib.wfl_parse_blast(self._h, text, len(text), ctypes.byref(flagged), ctypes.byref(first))
        self._check(n)
        if flagged.value:
            return None
        n = int(n)
        # ---- the distinct names: taxa (2^20 slots) and annotation systems (64 slots) ----
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        th, to, tl = np.empty(TAX_SLOTS, np.uint64), np.empty(TAX_SLOTS, np.int64), np.empty(TAX_SLOTS, np.int32)
        self._check(self._lib.wfl_parse_distinct(self._h, 0, vp(th), vp(to), vp(tl), TAX_SLOTS))
        tslots = np.flatnonzero(th)
        taxon_names = [text[int(o):int(o) + int(l)].decode() for o, l in zip(to[tslots], tl[tslots])]
        if len(set(taxon_names)) != len(taxon_names):
            return None   # two slots, one name: a 64-bit hash collision -- let the CPU reader handle the file
        tlut = np.full(TAX_SLOTS, -1, np.int32)
        tlut[tslots] = np.arange(len(tslots), dtype=np.int32)
        sh, so, sl = np.empty(SYS_SLOTS, np.uint64), np.empty(SYS_SLOTS, np.int64), np.empty(SYS_SLOTS, np.int32)
        self._check(self._lib.wfl_parse_distinct(self._h, 1, vp(sh), vp(so), vp(sl), SYS_SLOTS))
        sslots = np.flatnonzero(sh)
        snames = [text[int(o):int(o) + int(l)].decode() for o, l in zip(so[sslots], sl[sslots])]
        systems = sorted(set(snames))
        if len(systems) != len(snames) or len(systems) > 32:
            return None
        perm = np.full(SYS_SLOTS, -1, np.int32)
        for slot, name in zip(sslots, snames):
            perm[slot] = systems.index(name)
        # ---- the columns ----
        cols = dict(qstart=np.empty(n, np.int32), qend=np.empty(n, np.int32), score=np.empty(n, np.float64),
                    scov=np.empty(n, np.float64), strand=np.empty(n, np.int8), tcode=np.empty(n, np.int32),
                    sysmask=np.empty(n, np.uint32), ss_off=np.empty(n, np.int64), ss_len=np.empty(n, np.int32),
                    newblock=np.empty(n, np.uint8), q_off=np.empty(n, np.int64), q_len=np.empty(n, np.int32))
        self._check(self._lib.wfl_parse_fetch(self._h, vp(perm), *[vp(cols[k]) for k in (
            "qstart", "qend", "score", "scov", "strand", "tcode", "sysmask", "ss_off", "ss_len", "newblock", "q_off", "q_len")]))
        t = [ctypes.c_float() for _ in range(3)]
        self._lib.wfl_parser_times(self._h, *[ctypes.byref(x) for x in t])
        self.times = dict(ms_h2d=t[0].value, ms_kernels=t[1].value, ms_d2h=t[2].value, rows=n, bytes=len(text))
        starts = np.flatnonzero(cols["newblock"]).astype(np.int64)
        block_names = [text[int(o):int(o) + int(l)].decode() for o, l in zip(cols["q_off"][starts], cols["q_len"][starts])]
        names = _TextNames(text, cols["ss_off"], cols["ss_len"])
        return parsers.HitTable(
            qseqid=None, qstart=cols["qstart"], qend=cols["qend"], taxon=None, score=cols["score"],
            scov_modified=cols["scov"], strand=cols["strand"], sseqid_id=None, sseqid_names=names,
            sseqid_annotations=parsers._LazyAnnotations(names), systems=systems, sysmask=cols["sysmask"],
            taxon_codes=tlut[cols["tcode"]], taxon_names=taxon_names, block_starts=starts, block_names=block_names)


    def parse_gff(self, text):
        """`text`: bytes of a GFF file.  Returns a parsers.LocusTable (same content as parsers.read_gff_loci), or None if
        some row needs the CPU reader (waafle/utils.py:298-355: 9 fields, integer coordinates, '#' rows skipped)."""
        from . import parsers
        if not isinstance(text, (bytes, bytearray)):
            text = bytes(text)
        if len(text.strip()) == 0:
            return parsers.LocusTable([], [], [], [])
        flagged, first = ctypes.c_int32(), ctypes.c_int64()
        n = self._lib.wfl_parse_gff(self._h, text, len(text), ctypes.byref(flagged), ctypes.byref(first))
        self._check(n)
        if flagged.value:
            return None
        n = int(n)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        cols = dict(start=np.empty(n, np.int32), end=np.empty(n, np.int32), strand=np.empty(n, np.int8),
                    skip=np.empty(n, np.uint8), newblock=np.empty(n, np.uint8), s_off=np.empty(n, np.int64),
                    s_len=np.empty(n, np.int32))
        self._check(self._lib.wfl_parse_gff_fetch(self._h, *[vp(cols[k]) for k in (
            "start", "end", "strand", "skip", "newblock", "s_off", "s_len")]))
        t = [ctypes.c_float() for _ in range(3)]
        self._lib.wfl_parser_times(self._h, *[ctypes.byref(x) for x in t])
        self.times = dict(ms_h2d=t[0].value, ms_kernels=t[1].value, ms_d2h=t[2].value, rows=n, bytes=len(text))
        keep = np.flatnonzero(cols["skip"] == 0)
        nb = cols["newblock"][keep]
        starts = np.flatnonzero(nb)
        names = [text[int(o):int(o) + int(l)].decode() for o, l in zip(cols["s_off"][keep][starts], cols["s_len"][keep][starts])]
        counts = np.diff(np.r_[starts, len(keep)])
        lt = parsers.LocusTable.__new__(parsers.LocusTable)
        lt.seqname = np.repeat(np.array(names, dtype=object), counts) if len(names) else np.zeros(0, dtype=object)
        lt.start = np.ascontiguousarray(cols["start"][keep])
        lt.end = np.ascontiguousarray(cols["end"][keep])
        lt.strand = np.ascontiguousarray(cols["strand"][keep])
        lt.strand_str = lt.strand.view("S1").astype(str).astype(object)
        return lt
