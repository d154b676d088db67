"""Engine parameters: the POD mirror of the reference's argparse namespace.

Every field corresponds to a `waafle_orgscorer` flag read inside the engine
(reference waafle/waafle_orgscorer.py:188-296 and the shared flags of
waafle/waafle_genecaller.py:83-101).  The ctypes layout matches `wfl_params`
in include/waafle_b200.h field for field.
"""

import ctypes
from dataclasses import dataclass, fields

DISAMBIGUATE_ONE = {"report-best": 0, "meld": 1}
DISAMBIGUATE_TWO = {"report-best": 0, "jump": 1, "meld": 2}
WEAK_LOCI = {"ignore": 0, "penalize": 1, "assign-unknown": 2}
OFF_LENIENT_STRICT = {"off": 0, "lenient": 1, "strict": 2}
MAX_SYSTEMS = 32


class CParams(ctypes.Structure):
    """`wfl_params` (include/waafle_b200.h)."""
    _fields_ = [
        ("k1", ctypes.c_double),
        ("k2", ctypes.c_double),
        ("range", ctypes.c_double),
        ("ambiguous_fraction", ctypes.c_double),
        ("min_overlap", ctypes.c_double),
        ("min_scov", ctypes.c_double),
        ("min_gene_length", ctypes.c_double),
        ("disambiguate_one", ctypes.c_int32),
        ("disambiguate_two", ctypes.c_int32),
        ("weak_loci", ctypes.c_int32),
        ("ambiguous_threshold", ctypes.c_int32),
        ("sister_penalty", ctypes.c_int32),
        ("annotation_threshold", ctypes.c_int32),
        ("allow_lca", ctypes.c_int32),
        ("stranded", ctypes.c_int32),
        ("jump_taxonomy", ctypes.c_int32),
        ("clade_genes", ctypes.c_int32),
        ("clade_leaves", ctypes.c_int32),
        ("n_systems", ctypes.c_int32),
    ]


@dataclass
class OrgscorerParams:
    """Defaults are the reference CLI defaults."""
    k1: float = 0.5                    # -k1 / --one-clade-threshold   OS:188-194
    k2: float = 0.8                    # -k2 / --two-clade-threshold   OS:195-201
    range: float = 0.05                # --range                       OS:216-222
    ambiguous_fraction: float = 0.1    # --ambiguous-fraction          OS:238-244
    min_overlap: float = 0.1           # --min-overlap                 OS:290-296
    min_scov: float = 0.75             # --min-scov                    GC:90-96
    min_gene_length: float = 200.0     # --min-gene-length             GC:83-89
    disambiguate_one: int = 1          # --disambiguate-one            OS:202-208
    disambiguate_two: int = 2          # --disambiguate-two            OS:209-215
    weak_loci: int = 0                 # --weak-loci                   OS:276-282
    ambiguous_threshold: int = 1       # --ambiguous-threshold         OS:245-251
    sister_penalty: int = 2            # --sister-penalty              OS:252-258
    annotation_threshold: int = 1      # --annotation-threshold        OS:283-289
    allow_lca: int = 0                 # --allow-lca                   OS:233-237
    stranded: int = 0                  # --stranded                    GC:97-101
    jump_taxonomy: int = 0             # --jump-taxonomy (None -> 0)   OS:223-229
    clade_genes: int = -1              # --clade-genes   (None -> -1)  OS:259-265
    clade_leaves: int = -1             # --clade-leaves  (None -> -1)  OS:266-272
    n_systems: int = 0                 # annotation systems seen in the hits (UT:236-241)

    @classmethod
    def from_args(cls, args, n_systems=0):
        """Build from the argparse namespace of the reference CLI (same attribute names)."""
        return cls(
            k1=float(args.one_clade_threshold),
            k2=float(args.two_clade_threshold),
            range=float(args.range),
            ambiguous_fraction=float(args.ambiguous_fraction),
            min_overlap=float(args.min_overlap),
            min_scov=float(args.min_scov),
            min_gene_length=float(args.min_gene_length),
            disambiguate_one=DISAMBIGUATE_ONE[args.disambiguate_one],
            disambiguate_two=DISAMBIGUATE_TWO[args.disambiguate_two],
            weak_loci=WEAK_LOCI[args.weak_loci],
            ambiguous_threshold=OFF_LENIENT_STRICT[args.ambiguous_threshold],
            sister_penalty=OFF_LENIENT_STRICT[args.sister_penalty],
            annotation_threshold=OFF_LENIENT_STRICT[args.annotation_threshold],
            allow_lca=int(bool(args.allow_lca)),
            stranded=int(bool(args.stranded)),
            jump_taxonomy=0 if args.jump_taxonomy is None else int(args.jump_taxonomy),
            clade_genes=-1 if args.clade_genes is None else int(args.clade_genes),
            clade_leaves=-1 if args.clade_leaves is None else int(args.clade_leaves),
            n_systems=int(n_systems),
        )

    def as_dict(self):
        return {f.name: getattr(self, f.name) for f in fields(self)}

    def as_ctypes(self):
        if not 0 <= self.n_systems <= MAX_SYSTEMS:
            raise ValueError("at most %d annotation systems are supported" % MAX_SYSTEMS)
        return CParams(**self.as_dict())
