"""Taxonomy front end: the reference's child->parent file flattened to index tables.

Mirrors `waafle.utils.Taxonomy` (reference waafle/utils.py:374-447) for the queries the
orgscorer path makes, but on integer node indices.  Node index order is the Python
`str` order of the clade names, so that the reference's `clade1 < clade2` test
(waafle/waafle_orgscorer.py:608) and the canonical tie-break become
integer comparisons on the device.
"""

import csv

import numpy as np

from .utils import die, try_open

C_UNKNOWN = "Unknown"   # UT:367
C_ROOT = "r__Root"      # UT:368


class Taxonomy:
    """Parsed taxonomy file plus the flattened node tables handed to the engine."""

    def __init__(self, path=None, edges=None):
        self.parents = {}
        self.children = {}
        if path is not None:
            with try_open(path) as fh:
                edges = [row for row in csv.reader(fh, csv.excel_tab)]
        for row in edges or []:
            if len(row) != 2:
                die("bad taxonomy row:", row)
            clade, parent = row
            if clade in self.parents and self.parents[clade] != parent:
                die("clade listed under two parents in taxonomy:", clade)
            self.parents[clade] = parent                         # UT:381
            self.children.setdefault(parent, set()).add(clade)   # UT:382
        self.names = None
        self.index = None

    # ---- index tables -------------------------------------------------

    def build(self, extra_names=()):
        """Assign node indices (sorted names) and compute parent/depth/leaf_count/listed.

        extra_names: taxa seen in the BLAST hits; unlisted ones become depth-1 children of the
        root exactly as `get_parent`'s default does (UT:386-387).
        """
        names = set(self.parents) | set(self.children) | set(extra_names) | {C_ROOT, C_UNKNOWN}
        self.names = sorted(names)
        self.index = {n: i for i, n in enumerate(self.names)}
        n = len(self.names)
        root = self.index[C_ROOT]
        parent = np.full(n, root, dtype=np.int32)
        listed = np.zeros(n, dtype=np.uint8)
        for clade, par in self.parents.items():
            if clade == C_ROOT:
                continue   # get_lineage stops at the root (UT:394); its own row is never read
            parent[self.index[clade]] = self.index[par]
            listed[self.index[clade]] = 1
        # depth below the root; also detects cycles (the reference would loop forever)
        depth = np.full(n, -1, dtype=np.int32)
        depth[root] = 0
        for i in range(n):
            path = []
            j = i
            while depth[j] < 0:
                path.append(j)
                j = int(parent[j])
                if len(path) > n:
                    die("cycle in taxonomy at", self.names[i])
            d = int(depth[j])
            for k in reversed(path):
                d += 1
                depth[k] = d
        # leaf counts (UT:436-447): a clade that is nobody's parent counts 1
        leaf = np.zeros(n, dtype=np.int64)
        order = np.argsort(-depth, kind="stable")
        is_parent = np.zeros(n, dtype=bool)
        for par in self.children:
            is_parent[self.index[par]] = True
        for i in order:
            if not is_parent[i]:
                leaf[i] = 1
            if listed[i] and i != root:
                leaf[parent[i]] += leaf[i]
        self.parent = parent
        self.depth = depth
        self.leaf_count = np.minimum(leaf, np.iinfo(np.int32).max).astype(np.int32)
        self.listed = listed
        self.root_idx = root
        self.unknown_idx = self.index[C_UNKNOWN]
        return self

    def tables(self):
        """Plain dict of the arrays (the oracle and the C ABI take the same tables)."""
        return dict(parent=self.parent, depth=self.depth, leaf_count=self.leaf_count,
                    listed=self.listed, root_idx=self.root_idx, unknown_idx=self.unknown_idx)

    # ---- string-side queries used by the writer ------------------------

    def get_lineage_idx(self, c):   # UT:392-399
        l = [int(c)]
        while l[-1] != self.root_idx:
            l.append(int(self.parent[l[-1]]))
        l.reverse()
        return l

    def get_lineage(self, c):
        return [self.names[i] for i in self.get_lineage_idx(c)]

    def get_tail(self, c, lca):     # UT:413-426, one clade
        t = []
        for c2 in reversed(self.get_lineage_idx(c)):
            if c2 == lca:
                break
            t.append(self.names[c2])
        t.reverse()
        return t
