"""BLAST outfmt-6 and GFF readers producing column arrays (the values the packer ships).

Same field lists, derived quantities and error behaviour as the reference's row
objects -- `Hit` (waafle/utils.py:192-241) and `Locus` (waafle/utils.py:298-322) --
but columnar: one numpy array per attribute instead of one Python object per row.
The arithmetic keeps the reference's operation order so that `scov_modified` and
`waafle_score` are bit-identical (UT:216-229).
"""

import csv
import io

import numpy as np
import pandas as pd

from .utils import die, try_open

# UT:167-183
BLAST_FIELDS = ["qseqid", "sseqid", "qlen", "slen", "length", "qstart", "qend", "sstart",
                "send", "pident", "positive", "gaps", "evalue", "bitscore", "sstrand"]
_BLAST_STR = {"qseqid", "sseqid", "sstrand"}
_BLAST_FLOAT = {"pident", "evalue", "bitscore"}
# UT:282-292
GFF_FIELDS = ["seqname", "source", "feature", "start", "end", "score", "strand", "frame",
              "attribute"]

STRAND_PLUS, STRAND_MINUS = ord("+"), ord("-")


class HitTable:
    """All hits of a blastout file, in file order."""

    def __init__(self, qseqid, qstart, qend, taxon, score, scov_modified, strand,
                 sseqid_id, sseqid_names, sseqid_annotations, systems, sysmask):
        self.qseqid = qseqid                # object array of contig names
        self.qstart = qstart                # int32
        self.qend = qend                    # int32
        self.taxon = taxon                  # object array of taxon names (sseqid field 1)
        self.score = score                  # float64  waafle_score      UT:229
        self.scov_modified = scov_modified  # float64                    UT:227
        self.strand = strand                # int8 '+' / '-'             UT:214
        self.sseqid_id = sseqid_id          # int32 index into the unique-sseqid tables
        self.sseqid_names = sseqid_names
        self.sseqid_annotations = sseqid_annotations   # list of {system: value}  UT:237-241
        self.systems = systems              # sorted annotation systems seen in the hits
        self.sysmask = sysmask              # uint32 bit s = hit carries systems[s]

    def __len__(self):
        return len(self.qstart)


def hits_from_columns(qseqid, sseqid, qlen, slen, qstart, qend, sstart, send, pident, sstrand):
    """Derive the engine-facing hit columns from raw BLAST columns (UT:207-241)."""
    qlen, slen, qstart, qend, sstart, send = (
        np.asarray(a, dtype=np.int64) for a in (qlen, slen, qstart, qend, sstart, send))
    pident = np.asarray(pident, dtype=np.float64)
    minus = np.asarray(sstrand) == "minus"                       # UT:214
    s1 = np.where(minus, slen - sstart + 1, sstart)              # UT:219-224
    s2 = np.where(minus, slen - send + 1, send)
    ltrim = np.maximum(0, s1 - qstart)                           # UT:225
    rtrim = np.maximum(0, slen - s1 - qlen + qstart)             # UT:226
    denom = slen - ltrim - rtrim
    if np.any(denom == 0):
        raise ZeroDivisionError("float division by zero")       # what UT:227 raises
    scov_modified = (s2 - s1 + 1) / denom.astype(np.float64)     # UT:227
    score = scov_modified * pident / 100.0                       # UT:229
    # subject header: geneid|taxon|system=value...   UT:231-241
    sseqid = np.asarray(sseqid, dtype=object)
    uniq, inv = np.unique(sseqid.astype(str), return_inverse=True)
    taxa, anns = [], []
    for name in uniq:
        items = name.split("|")
        if len(items) < 2:
            die("bad subject id header:", name)
        taxa.append(items[1])
        d = {}
        for k in items[2:]:
            system, value = k.split("=")
            d[system] = value
        anns.append(d)
    systems = sorted({s for d in anns for s in d})
    sysbit = {s: i for i, s in enumerate(systems)}
    umask = np.array([sum(1 << sysbit[s] for s in d) for d in anns] or [0], dtype=np.uint64)
    taxa = np.array(taxa, dtype=object)
    lim = np.iinfo(np.int32)
    if len(qstart) and (max(qstart.max(), qend.max()) > lim.max or min(qstart.min(), qend.min()) < lim.min):
        die("hit coordinates exceed 32 bits")
    return HitTable(
        qseqid=np.asarray(qseqid, dtype=object),
        qstart=qstart.astype(np.int32), qend=qend.astype(np.int32),
        taxon=taxa[inv] if len(inv) else np.array([], dtype=object),
        score=score, scov_modified=scov_modified,
        strand=np.where(minus, STRAND_MINUS, STRAND_PLUS).astype(np.int8),
        sseqid_id=inv.astype(np.int32), sseqid_names=list(uniq), sseqid_annotations=anns,
        systems=systems,
        sysmask=(umask[inv] if len(inv) else np.array([], dtype=np.uint64)).astype(np.uint32)
        if len(systems) <= 32 else None,
    )


def read_blast_hits(path):
    """Parse a waafle_search blastout file (15-field outfmt 6, UT:167-186)."""
    with try_open(path) as fh:
        text = fh.read()
    if not text.strip():
        return hits_from_columns(*([[]] * 10))
    dtypes = {f: (str if f in _BLAST_STR else np.float64 if f in _BLAST_FLOAT else np.int64)
              for f in BLAST_FIELDS}
    try:
        df = pd.read_csv(io.StringIO(text), sep="\t", header=None, names=BLAST_FIELDS,
                         dtype=dtypes, float_precision="round_trip", na_filter=False,
                         quoting=csv.QUOTE_MINIMAL, index_col=False)
        ok = not (df["sstrand"] == "").any()
    except Exception:
        ok = False
    if not ok:
        for line in text.split("\n"):
            if line and line.count("\t") != len(BLAST_FIELDS) - 1:
                die("inconsistent blast row: {}".format(str(line.split("\t"))))   # UT:208-209
        die("unparseable blast file:", path)
    return hits_from_columns(
        df["qseqid"].to_numpy(dtype=object), df["sseqid"].to_numpy(dtype=object),
        df["qlen"].to_numpy(), df["slen"].to_numpy(), df["qstart"].to_numpy(),
        df["qend"].to_numpy(), df["sstart"].to_numpy(), df["send"].to_numpy(),
        df["pident"].to_numpy(), df["sstrand"].to_numpy(dtype=object))


class LocusTable:
    """All loci of a GFF file, in file order."""

    def __init__(self, seqname, start, end, strand_str):
        self.seqname = np.asarray(seqname, dtype=object)
        self.start = np.asarray(start, dtype=np.int32)
        self.end = np.asarray(end, dtype=np.int32)
        self.strand_str = np.asarray(strand_str, dtype=object)
        self.strand = np.array([ord(s) if len(s) == 1 else ord("?") for s in self.strand_str],
                               dtype=np.int8)

    def __len__(self):
        return len(self.start)

    def code(self, j):
        """`start:end:strand` (UT:317)."""
        return "{}:{}:{}".format(int(self.start[j]), int(self.end[j]), self.strand_str[j])


def read_gff_loci(path):
    """Parse a GFF (waafle_genecaller or Prodigal style); `#` rows are skipped (UT:345-346)."""
    with try_open(path) as fh:
        lines = [ln for ln in fh.read().split("\n") if ln and ln[0] != "#"]
    for ln in lines:
        if ln.count("\t") != len(GFF_FIELDS) - 1:
            die("Bad GFF row:", ln.split("\t"))   # UT:302-303
    if not lines:
        return LocusTable([], [], [], [])
    df = pd.read_csv(io.StringIO("\n".join(lines)), sep="\t", header=None, names=GFF_FIELDS,
                     dtype=str, na_filter=False, quoting=csv.QUOTE_NONE, index_col=False)
    return LocusTable(df["seqname"].to_numpy(dtype=object), df["start"].astype(np.int64),
                      df["end"].astype(np.int64), df["strand"].to_numpy(dtype=object))
