"""Seeded synthetic contigs / hits / taxonomy of the BASELINE.json shapes.

The generator works at the level of raw BLAST columns (qlen, slen, qstart, ..., pident,
sstrand) so that the same data can be (a) written as blastout / GFF / taxonomy / FASTA text
for the unmodified reference CLI and (b) packed straight into the engine's SoA arrays with
the front end's vectorised hit arithmetic (parsers.hits_from_columns).
"""

import os
from dataclasses import dataclass

import numpy as np

from .packing import Batch
from .parsers import STRAND_MINUS, STRAND_PLUS
from .taxonomy import Taxonomy

LEVELS_2 = [("g", 500), ("s", 5000)]
LEVELS_8 = [("k", 1), ("p", 4), ("c", 10), ("o", 30), ("f", 100), ("g", 500), ("s", 5000)]

CONFIGS = {
    # BASELINE.json configs[1]
    "cfg2": dict(n_contigs=100_000, genes=(2, 8), hits_per_gene=50.0, levels=LEVELS_2),
    # configs[2]
    "cfg3": dict(n_contigs=1_000_000, genes=(2, 20), hits_per_gene=50.0, levels=LEVELS_8),
    # configs[3]: long-contig stress
    "cfg4": dict(n_contigs=20_000, genes=(100, 130), hits_per_gene=48.0, levels=LEVELS_8,
                 gene_len=(500, 1100), gap=(0, 150), pool_size=550, island_genes=5,
                 frac_home=0.08, frac_relative=0.12),
    # configs[4]: Prodigal-style GFF (short loci, loci without hits), annotations
    "cfg5": dict(n_contigs=250_000, genes=(2, 20), hits_per_gene=50.0, levels=LEVELS_8,
                 frac_short_loci=0.07, frac_empty_loci=0.10, annotations=True),
}


@dataclass
class Synth:
    """Raw synthetic data; see to_batch() / write_files()."""
    tax_names: list            # names of taxonomy nodes (excluding r__Root), per node id
    tax_parent: np.ndarray     # parent node id, -1 = r__Root
    species_node: np.ndarray   # node id of species k
    contig_len: np.ndarray
    gene_off: np.ndarray       # [n+1]
    gene_start: np.ndarray
    gene_end: np.ndarray
    gene_strand: np.ndarray    # int8 '+'/'-'
    hit_off: np.ndarray        # [n+1]
    hit_gene: np.ndarray       # global gene id of each hit
    hit_species: np.ndarray    # species id, or -1-k for the k-th taxon missing from the taxonomy
    qstart: np.ndarray
    qend: np.ndarray
    slen: np.ndarray
    sstart: np.ndarray
    send: np.ndarray
    pident: np.ndarray
    minus: np.ndarray          # bool
    annotations: bool
    n_missing: int

    @property
    def n_contigs(self):
        return len(self.contig_len)

    def contig_name(self, i):
        return "contig{:08d}".format(i)

    def missing_name(self, k):
        return "s__unlisted_{:03d}".format(k)

    # ------------------------------------------------------------------
    def taxonomy(self):
        edges = [(nm, "r__Root" if p < 0 else self.tax_names[p])
                 for nm, p in zip(self.tax_names, self.tax_parent)]
        extra = [self.missing_name(k) for k in range(self.n_missing)]
        return Taxonomy(edges=edges).build(extra)

    def to_batch(self, taxonomy=None):
        """Pack directly (no text round trip); same arithmetic as parsers.hits_from_columns."""
        tax = taxonomy or self.taxonomy()
        node_index = np.array([tax.index[nm] for nm in self.tax_names], dtype=np.int32)
        miss_index = np.array([tax.index[self.missing_name(k)] for k in range(self.n_missing)]
                              or [0], dtype=np.int32)
        sp = self.hit_species
        taxon = np.where(sp >= 0, node_index[self.species_node[np.maximum(sp, 0)]],
                         miss_index[np.maximum(-1 - sp, 0)]).astype(np.int32)
        qlen = np.repeat(self.contig_len, np.diff(self.hit_off))
        slen, qstart = self.slen, self.qstart.astype(np.int64)
        s1 = np.where(self.minus, slen - self.sstart + 1, self.sstart)
        s2 = np.where(self.minus, slen - self.send + 1, self.send)
        ltrim = np.maximum(0, s1 - qstart)
        rtrim = np.maximum(0, slen - s1 - qlen + qstart)
        scov = (s2 - s1 + 1) / (slen - ltrim - rtrim).astype(np.float64)
        score = scov * self.pident / 100.0
        return Batch(
            hit_off=self.hit_off.astype(np.int64), locus_off=self.gene_off.astype(np.int64),
            hit_qstart=self.qstart.astype(np.int32), hit_qend=self.qend.astype(np.int32),
            hit_taxon=taxon, hit_score=score, hit_scov=scov,
            hit_strand=np.where(self.minus, STRAND_MINUS, STRAND_PLUS).astype(np.int8),
            locus_start=self.gene_start.astype(np.int32), locus_end=self.gene_end.astype(np.int32),
            locus_strand=self.gene_strand.astype(np.int8),
            hit_sysmask=np.ones(len(sp), dtype=np.uint32) if self.annotations else None,
            contig_names=[self.contig_name(i) for i in range(self.n_contigs)],
            contig_lengths=self.contig_len.astype(np.int64))

    def write_files(self, outdir, stem="synth", c0=0, c1=None):
        """Write <stem>.fna/.blastout/.gff/.taxonomy.tsv for the reference CLI / front end (contigs [c0, c1))."""
        os.makedirs(outdir, exist_ok=True)
        c1 = self.n_contigs if c1 is None else c1
        p = lambda ext: os.path.join(outdir, stem + ext)
        with open(p(".taxonomy.tsv"), "w") as fh:
            for nm, par in zip(self.tax_names, self.tax_parent):
                fh.write("{}\t{}\n".format(nm, "r__Root" if par < 0 else self.tax_names[par]))
        with open(p(".fna"), "w") as fh:
            for i in range(c0, c1):
                fh.write(">{} synthetic\n".format(self.contig_name(i)))
                fh.write("N" * int(self.contig_len[i]) + "\n")
        with open(p(".gff"), "w") as fh:
            fh.write("##gff-version  3\n")
            for i in range(c0, c1):
                for g in range(self.gene_off[i], self.gene_off[i + 1]):
                    fh.write("{}\tsynth\tCDS\t{}\t{}\t.\t{}\t0\tID={}_{}\n".format(
                        self.contig_name(i), self.gene_start[g], self.gene_end[g],
                        chr(self.gene_strand[g]), i, g))
        with open(p(".blastout"), "w") as fh:
            for i in range(c0, c1):
                for h in range(self.hit_off[i], self.hit_off[i + 1]):
                    sp = int(self.hit_species[h])
                    tname = (self.tax_names[self.species_node[sp]] if sp >= 0
                             else self.missing_name(-1 - sp))
                    sseqid = "GENE{:09d}|{}".format(h, tname)
                    if self.annotations:
                        sseqid += "|UniProt=U{:07d}".format(h % 9999991)
                    length = abs(int(self.qend[h]) - int(self.qstart[h])) + 1
                    fh.write("\t".join(str(x) for x in (
                        self.contig_name(i), sseqid, int(self.contig_len[i]), int(self.slen[h]),
                        length, int(self.qstart[h]), int(self.qend[h]), int(self.sstart[h]),
                        int(self.send[h]), "{:.3f}".format(self.pident[h]),
                        int(length * self.pident[h] / 100), 0, "0.0", 2 * length,
                        "minus" if self.minus[h] else "plus")) + "\n")
        return dict(contigs=p(".fna"), blastout=p(".blastout"), gff=p(".gff"),
                    taxonomy=p(".taxonomy.tsv"))


def _seg_cumsum(x, off):
    """Cumulative sum restarting at every segment start."""
    c = np.cumsum(x)
    base = np.repeat(c[off[:-1]] - x[off[:-1]], np.diff(off))
    return c - base


def generate(n_contigs, genes=(2, 8), gene_len=(300, 2500), gap=(0, 200), hits_per_gene=50.0,
             levels=LEVELS_2, frac_home=0.30, frac_relative=0.40, lgt_fraction=0.20,
             island_genes=1, pool_size=None, frac_low_scov=0.06, frac_missing_taxa=0.002,
             n_missing=5, frac_short_loci=0.0, frac_empty_loci=0.0, frac_integer_pident=0.05,
             annotations=False, seed=0):
    """Generate `n_contigs` synthetic contigs (SURVEY.md 8d recipe; all draws from `seed`)."""
    rng = np.random.default_rng(seed)
    # ---- taxonomy: balanced tree, level i node j hangs under node j*count[i-1]//count[i]
    names, parent, first = [], [], []
    for li, (prefix, count) in enumerate(levels):
        first.append(len(names))
        for j in range(count):
            names.append("{}__{}{:05d}".format(prefix, prefix.upper(), j))
            parent.append(-1 if li == 0 else first[li - 1] + j * levels[li - 1][1] // count)
    n_species = levels[-1][1]
    species_node = np.arange(first[-1], first[-1] + n_species)
    sp_parent = np.array(parent, dtype=np.int64)[species_node]
    # siblings of a species = the contiguous block sharing its parent
    blk_lo = np.searchsorted(sp_parent, sp_parent, side="left")
    blk_hi = np.searchsorted(sp_parent, sp_parent, side="right")

    # ---- contigs and genes
    G = rng.integers(genes[0], genes[1] + 1, size=n_contigs)
    gene_off = np.zeros(n_contigs + 1, dtype=np.int64)
    np.cumsum(G, out=gene_off[1:])
    nG = int(gene_off[-1])
    glen = rng.integers(gene_len[0], gene_len[1] + 1, size=nG)
    if frac_short_loci > 0:
        short = rng.random(nG) < frac_short_loci
        glen = np.where(short, rng.integers(60, 200, size=nG), glen)
    ggap = rng.integers(gap[0], gap[1] + 1, size=nG)
    gend = _seg_cumsum(glen + ggap, gene_off)
    gstart = gend - glen + 1
    contig_len = gend[gene_off[1:] - 1] + rng.integers(gap[0], gap[1] + 1, size=n_contigs)
    gstrand = np.where(rng.random(nG) < 0.5, STRAND_PLUS, STRAND_MINUS).astype(np.int8)
    gene_contig = np.repeat(np.arange(n_contigs), G)

    # ---- who lives on each gene: home species, optional donor island
    home = rng.integers(0, n_species, size=n_contigs)
    gene_home = home[gene_contig]
    is_lgt = rng.random(n_contigs) < lgt_fraction
    donor = rng.integers(0, n_species, size=n_contigs)
    isl0 = gene_off[:-1] + (rng.random(n_contigs) * np.maximum(1, G - island_genes + 1)).astype(np.int64)
    gidx = np.arange(nG)
    in_island = is_lgt[gene_contig] & (gidx >= isl0[gene_contig]) & (gidx < isl0[gene_contig] + island_genes)
    gene_home = np.where(in_island, donor[gene_contig], gene_home)

    # ---- hits
    nh = rng.poisson(hits_per_gene, size=nG)
    if frac_empty_loci > 0:
        nh = np.where(rng.random(nG) < frac_empty_loci, 0, nh)
    H = int(nh.sum())
    hit_gene = np.repeat(gidx, nh)
    hit_contig = gene_contig[hit_gene]
    hit_off = np.zeros(n_contigs + 1, dtype=np.int64)
    np.cumsum(np.bincount(hit_contig, minlength=n_contigs), out=hit_off[1:])
    u = rng.random(H)
    hh = gene_home[hit_gene]
    rel = blk_lo[hh] + (rng.random(H) * (blk_hi[hh] - blk_lo[hh])).astype(np.int64)
    if pool_size:
        # per-contig pool of species: pool member k of contig c is (base[c] + k*stride) % n
        base = rng.integers(0, n_species, size=n_contigs)
        k = rng.integers(0, pool_size, size=H)
        rnd = (base[hit_contig] + k * 7) % n_species
    else:
        rnd = rng.integers(0, n_species, size=H)
    is_home = u < frac_home
    species = np.where(is_home, hh, np.where(u < frac_home + frac_relative, rel, rnd))
    missing = rng.random(H) < frac_missing_taxa
    species = np.where(missing, -1 - rng.integers(0, n_missing, size=H), species)
    pident = np.where(is_home, rng.uniform(92, 100, size=H), rng.uniform(60, 95, size=H))
    pident = np.where(rng.random(H) < frac_integer_pident, np.rint(pident), np.round(pident, 3))
    # query span: gene trimmed (or slightly extended) at both ends
    gl = glen[hit_gene]
    lo = gstart[hit_gene] + rng.integers(-20, 61, size=H)
    hi = gend[hit_gene] - rng.integers(-20, 61, size=H)
    full = rng.random(H) < 0.10     # exact full-gene hits (knife-edge material)
    lo = np.where(full, gstart[hit_gene], lo)
    hi = np.where(full, gend[hit_gene], hi)
    lo = np.clip(lo, 1, contig_len[hit_contig])
    hi = np.clip(hi, 1, contig_len[hit_contig])
    bad = hi - lo < 30
    lo = np.where(bad, gstart[hit_gene], lo)
    hi = np.where(bad, gend[hit_gene], hi)
    length = hi - lo + 1
    # subject: slen >= length; a tail of partial-coverage hits fails --min-scov
    extra = (length * rng.uniform(0, 0.08, size=H)).astype(np.int64)
    extra = np.where(full, 0, extra)
    low = rng.random(H) < frac_low_scov
    extra = np.where(low, (length * rng.uniform(0.4, 1.5, size=H)).astype(np.int64), extra)
    slen = length + extra
    s_lo = 1 + (rng.random(H) * (extra + 1)).astype(np.int64)
    s_hi = s_lo + length - 1
    minus = rng.random(H) < 0.5
    flip = rng.random(H) < 0.5      # query coordinates reported descending for half the minus hits
    sstart = np.where(minus, s_hi, s_lo)
    send = np.where(minus, s_lo, s_hi)
    qstart, qend = lo, hi
    del flip
    return Synth(tax_names=names, tax_parent=np.array(parent, dtype=np.int64),
                 species_node=species_node, contig_len=contig_len, gene_off=gene_off,
                 gene_start=gstart, gene_end=gend, gene_strand=gstrand, hit_off=hit_off,
                 hit_gene=hit_gene, hit_species=species, qstart=qstart, qend=qend, slen=slen,
                 sstart=sstart, send=send, pident=pident, minus=minus, annotations=annotations,
                 n_missing=n_missing)


def generate_config(name, n_contigs=None, seed=0, **over):
    """One of the named BASELINE.json shapes, optionally at a reduced contig count."""
    kw = dict(CONFIGS[name])
    kw.update(over)
    if n_contigs is not None:
        kw["n_contigs"] = n_contigs
    return generate(seed=seed, **kw)
