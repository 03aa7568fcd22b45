"""Gene trees at the marshalling boundary.

The reference passes ``Network`` objects to ``FelsensteinCalculator.log_likelihood``
and reads them through ``root()``, ``get_children(node)``, ``get_edge(node, child)``
(may be a list), ``edge.get_length()`` (may be ``None``) and ``node.label``
(reference ``src/_seq_likelihood.py:391,407-420``).  The device wants flat arrays.
This module provides

* :class:`FlatTree` -- parent / branch-length / label arrays (the C ABI's layout),
* :func:`flatten`   -- any object with the reference's read API -> :class:`FlatTree`,
* :class:`GeneTree` -- a small tree with that same read API plus a Newick reader,
  so tests and examples do not need the reference's ``Network``.

The reference's graph layer (``src/Network.py``) itself is out of scope.
"""
from __future__ import annotations

import math
from typing import Any, List, Optional, Sequence

import numpy as np


class FlatTree:
    """Rooted tree as arrays: ``parent[i]`` (-1 = root), ``blen[i]`` = length of the
    edge above node ``i`` (NaN = the reference's ``None`` -> 0), ``label[i]``."""

    __slots__ = ("parent", "blen", "label", "_n_children", "_tip_rows")

    def __init__(self, parent: Sequence[int], blen: Sequence[Optional[float]], label: Sequence[Optional[str]]):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.blen = np.ascontiguousarray(
            [math.nan if b is None else float(b) for b in blen] if not isinstance(blen, np.ndarray) else blen,
            dtype=np.float64)
        self.label = list(label)
        n = self.parent.shape[0]
        if self.blen.shape[0] != n or len(self.label) != n:
            raise ValueError("FlatTree: parent, blen and label must have the same length")
        self._n_children = None
        self._tip_rows = None  # (id(row_of_label), array): shared by every with_lengths() copy

    @property
    def n_nodes(self) -> int:
        return int(self.parent.shape[0])

    def n_children(self) -> np.ndarray:
        if self._n_children is None:
            p = self.parent
            ok = (p >= 0) & (p < p.shape[0])  # out-of-range parents are reported by the library, not here
            self._n_children = np.bincount(p[ok], minlength=p.shape[0]).astype(np.int32)
        return self._n_children

    def with_lengths(self, blen: np.ndarray) -> "FlatTree":
        t = FlatTree.__new__(FlatTree)
        t.parent = self.parent
        t.blen = np.ascontiguousarray(blen, dtype=np.float64)
        t.label = self.label
        t._n_children = self._n_children
        t._tip_rows = self._tip_rows
        return t

    def tip_rows(self, row_of_label: dict) -> np.ndarray:
        """Row of the locus' code matrix per node: leaf with a known label -> row;
        leaf absent from the alignment -> -1 (all-ones partial, reference :409-412);
        internal nodes -> -2 (ignored by the library)."""
        cached = self._tip_rows
        if cached is not None and cached[0] is row_of_label:
            return cached[1]
        nc = self.n_children()
        out = np.full(self.n_nodes, -2, dtype=np.int32)
        for i in np.nonzero(nc == 0)[0]:
            out[i] = row_of_label.get(self.label[i], -1)
        self._tip_rows = (row_of_label, out)  # topology and labels are immutable: safe to reuse
        return out


def flatten(tree: Any) -> FlatTree:
    """Marshal a reference-style tree object (or a FlatTree / GeneTree) into arrays."""
    if isinstance(tree, FlatTree):
        return tree
    if isinstance(tree, GeneTree):
        return tree.flat()
    root = tree.root()
    parent: List[int] = [-1]
    blen: List[float] = [math.nan]
    label: List[Optional[str]] = [getattr(root, "label", None)]
    stack = [(root, 0)]
    while stack:
        node, idx = stack.pop()
        for child in tree.get_children(node):
            edges = tree.get_edge(node, child)
            edge = edges[0] if isinstance(edges, list) else edges  # reference :417-418
            t = edge.get_length()
            parent.append(idx)
            blen.append(math.nan if t is None else float(t))  # reference :419-420
            label.append(getattr(child, "label", None))
            stack.append((child, len(parent) - 1))
    return FlatTree(parent, np.asarray(blen, dtype=np.float64), label)


# ---------------------------------------------------------------------------
class Node:
    __slots__ = ("label", "_idx")

    def __init__(self, label: Optional[str], idx: int):
        self.label = label
        self._idx = idx

    def __repr__(self) -> str:
        return "Node(%r)" % (self.label,)


class Edge:
    __slots__ = ("src", "dest", "_length")

    def __init__(self, src: Node, dest: Node, length: Optional[float]):
        self.src = src
        self.dest = dest
        self._length = length

    def get_length(self) -> Optional[float]:
        return self._length

    def set_length(self, length: Optional[float]) -> None:
        self._length = length


class GeneTree:
    """A rooted tree exposing the read API the reference's Felsenstein path uses."""

    def __init__(self, parent: Sequence[int], blen: Sequence[Optional[float]], label: Sequence[Optional[str]]):
        n = len(parent)
        self._nodes = [Node(label[i], i) for i in range(n)]
        self._parent = [int(p) for p in parent]
        self._children: List[List[Node]] = [[] for _ in range(n)]
        self._edge_above: List[Optional[Edge]] = [None] * n
        roots = [i for i, p in enumerate(self._parent) if p < 0]
        if len(roots) != 1:
            raise ValueError("GeneTree needs exactly one root, found %d" % len(roots))
        self._root = self._nodes[roots[0]]
        for i, p in enumerate(self._parent):
            if p >= 0:
                self._children[p].append(self._nodes[i])
                self._edge_above[i] = Edge(self._nodes[p], self._nodes[i], blen[i])
        self._flat: Optional[FlatTree] = None

    # -- the reference's read API ------------------------------------------
    def root(self) -> Node:
        return self._root

    def get_children(self, node: Node) -> List[Node]:
        return self._children[node._idx]

    def get_edge(self, node: Node, child: Node) -> Edge:
        e = self._edge_above[child._idx]
        if e is None or e.src is not node:
            raise KeyError("no edge %r -> %r" % (node, child))
        return e

    def get_leaves(self) -> List[Node]:
        return [n for n in self._nodes if not self._children[n._idx]]

    def V(self) -> List[Node]:
        return list(self._nodes)

    def E(self) -> List[Edge]:
        return [e for e in self._edge_above if e is not None]

    # -- flat view ------------------------------------------------------------
    def flat(self) -> FlatTree:
        blen = [math.nan if e is None or e.get_length() is None else float(e.get_length()) for e in self._edge_above]
        return FlatTree(self._parent, np.asarray(blen, dtype=np.float64), [n.label for n in self._nodes])

    @classmethod
    def from_flat(cls, ft: FlatTree) -> "GeneTree":
        return cls(ft.parent.tolist(), [None if math.isnan(b) else float(b) for b in ft.blen], ft.label)

    @classmethod
    def from_newick(cls, newick: str) -> "GeneTree":
        """Parse plain Newick (nested parentheses, labels, ``:length``)."""
        s = newick.strip()
        if s.endswith(";"):
            s = s[:-1]
        parent: List[int] = [-1]
        blen: List[Optional[float]] = [None]
        label: List[Optional[str]] = [None]
        cur = 0
        open_nodes: List[int] = []
        i, n = 0, len(s)
        while i < n:
            ch = s[i]
            if ch == "(":
                open_nodes.append(cur)
                parent.append(cur); blen.append(None); label.append(None)
                cur = len(parent) - 1
                i += 1
            elif ch == ",":
                if not open_nodes:
                    raise ValueError("Newick: ',' outside parentheses")
                parent.append(open_nodes[-1]); blen.append(None); label.append(None)
                cur = len(parent) - 1
                i += 1
            elif ch == ")":
                if not open_nodes:
                    raise ValueError("Newick: unbalanced ')'")
                cur = open_nodes.pop()
                i += 1
            elif ch == ":":
                j = i + 1
                while j < n and s[j] not in ",()[":
                    j += 1
                blen[cur] = float(s[i + 1:j])
                i = j
            elif ch == "[":  # comments / annotations are skipped
                j = s.find("]", i)
                if j < 0:
                    raise ValueError("Newick: unterminated '['")
                i = j + 1
            elif ch.isspace():
                i += 1
            else:
                j = i
                while j < n and s[j] not in ":,()[":
                    j += 1
                label[cur] = s[i:j].strip()
                i = j
        if open_nodes:
            raise ValueError("Newick: unbalanced '('")
        return cls(parent, blen, label)
