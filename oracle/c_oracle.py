"""TEST INFRASTRUCTURE ONLY -- ctypes front of oracle/orgscorer_oracle.c (the literal C restatement).

Same call shape as `orgscorer_oracle.score_batch(params, tax, batch)` (the numpy oracle of record).  The
C restatement exists so that parity can be checked bit-for-bit at BASELINE.json's full sizes, and as the
multi-threaded CPU arm of bench.py; tests/test_c_oracle.py pins it to the numpy oracle and to the
reference-generated golden records.  Nothing under waafle_b200/ imports this module.
"""

import ctypes
import os
import subprocess

import numpy as np

from waafle_b200.engine import CBatch, CResults, _as, _ptr
from waafle_b200.params import CParams, OrgscorerParams

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liborgscorer_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "orgscorer_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        i32p, u8p = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8)
        lib.wfl_oracle_score_batch.argtypes = [ctypes.POINTER(CParams), ctypes.c_int32, i32p, i32p, i32p, u8p,
                                               ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(CBatch),
                                               ctypes.POINTER(CResults)]
        lib.wfl_oracle_score_batch.restype = ctypes.c_int
        _lib = lib
    return _lib


def score_batch(params, tax, batch, threads=None):
    """params: dict or OrgscorerParams; tax: Taxonomy or its tables(); batch: Batch or its arrays()."""
    lib = load()
    if isinstance(params, dict):
        params = OrgscorerParams(**params)
    S = params.n_systems
    cp = params.as_ctypes()
    t = tax.tables() if hasattr(tax, "tables") else tax
    parent, depth = _as(t["parent"], np.int32), _as(t["depth"], np.int32)
    leaf, listed = _as(t["leaf_count"], np.int32), _as(t["listed"], np.uint8)
    a = batch.arrays() if hasattr(batch, "arrays") else batch
    k = dict(
        hit_off=_as(a["hit_off"], np.int64), locus_off=_as(a["locus_off"], np.int64),
        hit_qstart=_as(a["hit_qstart"], np.int32), hit_qend=_as(a["hit_qend"], np.int32),
        hit_taxon=_as(a["hit_taxon"], np.int32), hit_score=_as(a["hit_score"], np.float64),
        hit_scov=_as(a["hit_scov"], np.float64), hit_strand=_as(a["hit_strand"], np.int8),
        locus_start=_as(a["locus_start"], np.int32), locus_end=_as(a["locus_end"], np.int32),
        locus_strand=_as(a["locus_strand"], np.int8))
    n, nh, nl = len(k["hit_off"]) - 1, len(k["hit_qstart"]), len(k["locus_start"])
    cb = CBatch(n_contigs=n, n_hits=nh, n_loci=nl)
    ftypes = dict(CBatch._fields_)
    for name, arr in k.items():
        setattr(cb, name, _ptr(arr, ftypes[name]._type_))
    if S > 0:
        k["hit_sysmask"] = _as(a["hit_sysmask"], np.uint32)
        cb.hit_sysmask = _ptr(k["hit_sysmask"], ctypes.c_uint32)
    cap = max(4 * n, 1024)
    old = os.environ.get("WFL_ORACLE_THREADS")
    if threads:
        os.environ["WFL_ORACLE_THREADS"] = str(int(threads))
    try:
        for _ in range(2):
            r = dict(
                call=np.zeros(n, np.uint8), direction=np.zeros(n, np.uint8),
                lifts=np.zeros(n, np.int32), clade1=np.zeros(n, np.int32),
                clade2=np.zeros(n, np.int32), lca=np.zeros(n, np.int32), best1=np.zeros(n, np.int32),
                best2=np.zeros(n, np.int32), crit=np.zeros(n, np.float64), rank=np.zeros(n, np.float64),
                member_off=np.zeros(n + 1, np.int64), n_members_a=np.zeros(n, np.int32),
                members=np.zeros(max(1, cap), np.int32),
                synteny=np.zeros(nl, np.uint8), locus_flags=np.zeros(nl, np.uint8),
                ann_winner=np.zeros((nl, S), np.int32),
                call_counts=np.zeros(3, np.int64), call_index=np.zeros(n, np.int64))
            cr = CResults(members_capacity=len(r["members"]), members_used=0)
            rtypes = dict(CResults._fields_)
            for name, arr in r.items():
                setattr(cr, name, _ptr(arr, rtypes[name]._type_))
            rc = lib.wfl_oracle_score_batch(ctypes.byref(cp), len(parent), _ptr(parent, ctypes.c_int32),
                                            _ptr(depth, ctypes.c_int32), _ptr(leaf, ctypes.c_int32),
                                            _ptr(listed, ctypes.c_uint8), int(t["root_idx"]),
                                            int(t["unknown_idx"]), ctypes.byref(cb), ctypes.byref(cr))
            if rc == -4:
                cap = int(cr.members_used)
                continue
            if rc != 0:
                raise RuntimeError("wfl_oracle_score_batch failed: {}".format(rc))
            r["members"] = r["members"][:cr.members_used]
            return r
        raise RuntimeError("wfl_oracle_score_batch: capacity retry failed")
    finally:
        if threads:
            if old is None:
                os.environ.pop("WFL_ORACLE_THREADS", None)
            else:
                os.environ["WFL_ORACLE_THREADS"] = old
