"""Timed MSNC density log P(g | Psi) on the GPU -- the coalescent factor of the MCMC_SEQ likelihood.

Mirrors the public surface of the reference's ``src/_msnc_density.py`` for the timed
(event-based) density MCMC_SEQ uses:

* :func:`gene_tree_msnc_log_density` (reference ``:496-526``) -- one gene tree;
* :func:`msnc_log_density_batch` -- all loci, or all candidates of all loci, against one network in
  ONE device call (the loops of ``src/_mcmc_seq.py:750-760`` and ``:1935-1957``);
* :class:`SpeciesNetwork` -- the flat form of a species network (the role of ``_NetworkIndex`` +
  ``build_network_msnc_index``, reference ``:838-935``, ``:392-417``), built from arrays, from any
  object with the reference's ``Network`` read API, or from extended Newick with
  ``[&gamma=...]`` annotations.

The arithmetic runs in kernel K-m behind ``pnb_msnc_eval`` (``include/phynet_b200.h``); there is no
CPU fallback.  Fixed per-branch thetas (``branch_thetas``, label-keyed as in ``src/_units.py``) are supported.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .device import DeviceContext, default_context
from .tree import FlatTree, flatten

LOG_FLOOR = math.log(1e-200)   # reference _LOG_FLOOR, src/_msnc_density.py:543


class SpeciesNetwork:
    """Rooted species network as arrays.

    ``edge_src[e] -> edge_dst[e]`` (parent -> child) in the reference's ``E()`` order,
    ``edge_gamma[e]`` (NaN = absent), ``height[v]`` above the present, ``leaf_label[v]`` for leaves
    (``None`` for internal nodes).  A node with two parents is a reticulation.  Species ids are
    the ranks of the leaves in node order.
    """

    def __init__(self, edge_src: Sequence[int], edge_dst: Sequence[int], edge_gamma: Sequence[Optional[float]],
                 height: Sequence[float], leaf_label: Sequence[Optional[str]],
                 node_label: Optional[Sequence[Optional[str]]] = None):
        self.edge_src = np.ascontiguousarray(edge_src, dtype=np.int32)
        self.edge_dst = np.ascontiguousarray(edge_dst, dtype=np.int32)
        self.edge_gamma = np.ascontiguousarray(
            [math.nan if g is None else float(g) for g in edge_gamma], dtype=np.float64)
        self.height = np.ascontiguousarray(height, dtype=np.float64)
        self.leaf_label = list(leaf_label)
        # labels of ALL nodes: only needed to resolve label-keyed per-branch thetas
        self.node_label = list(node_label) if node_label is not None else list(leaf_label)
        n = self.height.shape[0]
        if len(self.leaf_label) != n:
            raise ValueError("SpeciesNetwork: height and leaf_label must have one entry per node")
        if not (self.edge_src.shape == self.edge_dst.shape == self.edge_gamma.shape):
            raise ValueError("SpeciesNetwork: edge arrays must have the same length")
        has_child = np.zeros(n, dtype=bool)
        ok = (self.edge_src >= 0) & (self.edge_src < n)
        has_child[self.edge_src[ok]] = True
        self.species: List[str] = []
        self.node_species = np.full(n, -1, dtype=np.int32)
        for v in range(n):
            if not has_child[v]:
                self.node_species[v] = len(self.species)
                self.species.append(self.leaf_label[v])
        self.species_id: Dict[Any, int] = {}
        for i, s in enumerate(self.species):
            self.species_id.setdefault(s, i)
        self._handles: Dict[Tuple[int, int], "_NetHandle"] = {}

    # -- constructors ----------------------------------------------------------------------
    @classmethod
    def from_dict(cls, d: dict) -> "SpeciesNetwork":
        """From the layout of ``tests/golden/msnc_cases.json`` (``network`` entries)."""
        return cls(d["edge_src"], d["edge_dst"], d["edge_gamma"], d["height"], d["leaf_label"], d.get("node_label"))

    def replaced(self, height: Optional[np.ndarray] = None, edge_gamma: Optional[np.ndarray] = None) -> "SpeciesNetwork":
        """A new network with other node heights and / or inheritance probabilities, same topology and labels
        (what a height or gamma move of the sampler produces; the old object stays untouched)."""
        n = SpeciesNetwork.__new__(SpeciesNetwork)
        n.__dict__.update(self.__dict__)
        n.height = self.height if height is None else np.ascontiguousarray(height, dtype=np.float64)
        n.edge_gamma = self.edge_gamma if edge_gamma is None else np.ascontiguousarray(edge_gamma, dtype=np.float64)
        n._handles = {}
        if n.height.shape != self.height.shape or n.edge_gamma.shape != self.edge_gamma.shape:
            raise ValueError("SpeciesNetwork.replaced: array shapes must not change")
        return n

    def to_dict(self) -> dict:
        """The layout :meth:`from_dict` reads (plain lists; ``None`` for an absent gamma)."""
        return {"n_nodes": self.n_nodes, "edge_src": self.edge_src.tolist(), "edge_dst": self.edge_dst.tolist(),
                "edge_gamma": [None if g != g else g for g in self.edge_gamma.tolist()],
                "height": self.height.tolist(), "leaf_label": list(self.leaf_label), "node_label": list(self.node_label),
                "is_retic": (np.bincount(self.edge_dst, minlength=self.n_nodes) >= 2).tolist()}

    @classmethod
    def from_network(cls, net: Any) -> "SpeciesNetwork":
        """From an object with the reference's read API: ``V()``, ``E()``, ``edge.src`` / ``.dest`` /
        ``.get_gamma()`` / ``.get_length()``, ``node.label``.  Heights follow ``_node_height``
        (``src/GraphUtils.py:786-819``): max over children of child height + edge length, leaves at 0."""
        nodes = list(net.V())
        pos = {id(n): i for i, n in enumerate(nodes)}
        edges = list(net.E())
        src = [pos[id(e.src)] for e in edges]
        dst = [pos[id(e.dest)] for e in edges]
        gam = [e.get_gamma() if hasattr(e, "get_gamma") else None for e in edges]
        length = [e.get_length() for e in edges]
        n = len(nodes)
        kids: List[List[Tuple[int, float]]] = [[] for _ in range(n)]
        for s, d, ln in zip(src, dst, length):
            kids[s].append((d, 0.0 if ln is None else float(ln)))
        h = [None] * n

        def height(v: int) -> float:
            stack = [v]
            while stack:
                x = stack[-1]
                if h[x] is not None:
                    stack.pop()
                    continue
                pend = [c for c, _ in kids[x] if h[c] is None]
                if pend:
                    stack.extend(pend)
                    continue
                h[x] = max((h[c] + ln for c, ln in kids[x]), default=0.0)
                stack.pop()
            return h[v]
        for v in range(n):
            height(v)
        labels = [nodes[v].label if not kids[v] else None for v in range(n)]
        return cls(src, dst, gam, h, labels, [getattr(nd, "label", None) for nd in nodes])

    @classmethod
    def from_newick(cls, newick: str) -> "SpeciesNetwork":
        """Extended Newick: a hybrid node appears once with its subtree and once as a leaf, both
        named ``#H..``; ``[&gamma=x]`` after a branch length is the inheritance probability of that
        parent edge (the syntax the reference's tests use, ``tests/test_mcmc_seq.py:224-227``)."""
        s = newick.strip().rstrip(";")
        labels: List[Optional[str]] = []
        kids: List[List[int]] = []
        hybrid: Dict[str, int] = {}
        edges: List[Tuple[int, int, Optional[float], Optional[float]]] = []   # src, dst, length, gamma
        pos = 0

        def new_node() -> int:
            labels.append(None)
            kids.append([])
            return len(labels) - 1

        def parse() -> Tuple[int, Optional[float], Optional[float]]:
            nonlocal pos
            children = []
            if pos < len(s) and s[pos] == "(":
                pos += 1
                while True:
                    children.append(parse())
                    if pos >= len(s):
                        raise ValueError("extended Newick: unbalanced '('")
                    if s[pos] == ",":
                        pos += 1
                        continue
                    if s[pos] == ")":
                        pos += 1
                        break
                    raise ValueError("extended Newick: unexpected %r" % s[pos])
            j = pos
            while j < len(s) and s[j] not in ":,()[":
                j += 1
            name = s[pos:j].strip() or None
            pos = j
            length = gamma = None
            if pos < len(s) and s[pos] == ":":
                j = pos + 1
                while j < len(s) and s[j] not in ",()[":
                    j += 1
                length = float(s[pos + 1:j])
                pos = j
            while pos < len(s) and s[pos] == "[":
                j = s.find("]", pos)
                if j < 0:
                    raise ValueError("extended Newick: unterminated '['")
                body = s[pos + 1:j]
                if body.startswith("&gamma="):
                    gamma = float(body[7:])
                pos = j + 1
            if name is not None and name.startswith("#"):
                v = hybrid.get(name)
                if v is None:
                    v = hybrid[name] = new_node()
                    labels[v] = name
            else:
                v = new_node()
                labels[v] = name
            for c, ln, g in children:
                kids[v].append(c)
                edges.append((v, c, ln, g))
            return v, length, gamma

        parse()
        if pos != len(s):
            raise ValueError("extended Newick: trailing text %r" % s[pos:])
        n = len(labels)
        lens: List[List[Tuple[int, float]]] = [[] for _ in range(n)]
        for a, b, ln, _ in edges:
            lens[a].append((b, 0.0 if ln is None else ln))
        h: List[Optional[float]] = [None] * n
        order = list(range(n))
        for _ in range(n + 1):
            pending = False
            for v in order:
                if h[v] is None:
                    if all(h[c] is not None for c, _ in lens[v]):
                        h[v] = max((h[c] + ln for c, ln in lens[v]), default=0.0)
                    else:
                        pending = True
            if not pending:
                break
        else:
            raise ValueError("extended Newick: the network has a cycle")
        leaf_label = [labels[v] if not kids[v] else None for v in range(n)]
        return cls([e[0] for e in edges], [e[1] for e in edges], [e[3] for e in edges], h, leaf_label, labels)

    # -- device handle -------------------------------------------------------------------------
    @property
    def n_nodes(self) -> int:
        return int(self.height.shape[0])

    def n_reticulations(self) -> int:
        return int((np.bincount(self.edge_dst, minlength=self.n_nodes) >= 2).sum())

    def handle(self, ctx: DeviceContext, branch_thetas: Optional[dict] = None) -> C.c_void_p:
        key = (os.getpid(), id(ctx))
        h = self._handles.get(key)
        if h is None or h.ctx is not ctx or ctx._h is None:
            h = _NetHandle(self, ctx)
            self._handles[key] = h
        h.set_branch_thetas(self, branch_thetas)
        return h.h

    def branch_theta_arrays(self, branch_thetas: Optional[dict]) -> Tuple[Optional[np.ndarray], float]:
        """Label-keyed overrides -> (theta per edge, theta above the root), NaN = the default theta.
        Precedence as ``resolve_branch_theta`` (``src/_units.py:90-125``): ``(parent_label, child_label)``
        before ``child_label``; above the root ``("__root__", label)``, ``"__root__"``, then the label."""
        if not branch_thetas:
            return None, math.nan
        validate_branch_thetas(branch_thetas)
        lab = self.node_label
        edge_theta = np.full(self.edge_src.shape[0], math.nan, dtype=np.float64)
        for e, (s, d) in enumerate(zip(self.edge_src.tolist(), self.edge_dst.tolist())):
            for key in ((lab[s], lab[d]), lab[d]):
                if key in branch_thetas:
                    edge_theta[e] = float(branch_thetas[key])
                    break
        root_theta = math.nan
        has_parent = np.zeros(self.n_nodes, dtype=bool)
        has_parent[self.edge_dst] = True
        roots = np.nonzero(~has_parent)[0]
        if roots.shape[0]:
            r = lab[int(roots[0])]
            for key in (("__root__", r), "__root__", r):
                if key in branch_thetas:
                    root_theta = float(branch_thetas[key])
                    break
        return edge_theta, root_theta

    def info(self, ctx: Optional[DeviceContext] = None) -> Dict[str, int]:
        """Shape of the compiled step program (open-edge slots, steps, reticulations, species)."""
        ctx = ctx or default_context()
        arr = (C.c_int32 * 4)()
        _lib.check(_lib.load().pnb_msnc_net_info(self.handle(ctx), arr))
        return {"slots": arr[0], "steps": arr[1], "reticulations": arr[2], "species": arr[3]}

    def __getstate__(self):   # device handles never cross a process boundary
        d = self.__dict__.copy()
        d["_handles"] = {}
        return d

    # -- gene-tree side -------------------------------------------------------------------------
    def species_ids_of(self, ft: FlatTree, species_of: Dict[str, str]) -> np.ndarray:
        """Per gene-tree node: species id of the species a leaf's allele maps to (-1: unmapped or
        internal) -- ``species_to_bits`` of the reference (``src/_msnc_density.py:262-266``)."""
        out = np.full(ft.n_nodes, -1, dtype=np.int32)
        nk = ft.n_children()
        sid = self.species_id
        for i in np.nonzero(nk == 0)[0]:
            sp = species_of.get(ft.label[i])
            if sp is not None:
                out[i] = sid.get(sp, -1)
        return out


class _NetHandle:
    def __init__(self, net: SpeciesNetwork, ctx: DeviceContext):
        self.ctx = ctx
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.pnb_msnc_net_create(
            ctx.handle, net.n_nodes, int(net.edge_src.shape[0]), _lib.ptr(net.edge_src), _lib.ptr(net.edge_dst),
            _lib.ptr(net.edge_gamma), _lib.ptr(net.height), _lib.ptr(net.node_species), len(net.species), C.byref(h)))
        self.h = h
        self._pid = os.getpid()
        self._thetas: Any = None      # the branch_thetas mapping currently installed on the handle (by content)

    def set_branch_thetas(self, net: "SpeciesNetwork", branch_thetas: Optional[dict]) -> None:
        want = tuple(sorted(branch_thetas.items(), key=repr)) if branch_thetas else None
        if want == self._thetas:
            return
        edge_theta, root_theta = net.branch_theta_arrays(branch_thetas)
        _lib.check(self._lib.pnb_msnc_net_set_branch_thetas(
            self.h, _lib.ptr(edge_theta) if edge_theta is not None else None, C.c_double(root_theta)))
        self._thetas = want

    def __del__(self):
        try:
            if self.h is not None and self._pid == os.getpid() and self.ctx._h is not None:
                self._lib.pnb_msnc_net_destroy(self.h)
        except Exception:
            pass
        self.h = None


def as_species_network(species_net: Any) -> SpeciesNetwork:
    if isinstance(species_net, SpeciesNetwork):
        return species_net
    if isinstance(species_net, str):
        return SpeciesNetwork.from_newick(species_net)
    if isinstance(species_net, dict):
        return SpeciesNetwork.from_dict(species_net)
    return SpeciesNetwork.from_network(species_net)


def validate_branch_thetas(branch_thetas: Optional[dict]) -> None:
    """``validate_branch_thetas`` of the reference (``src/_units.py:71-87``)."""
    for key, value in (branch_thetas or {}).items():
        if not (isinstance(key, str) or (isinstance(key, tuple) and len(key) == 2 and all(isinstance(x, str) for x in key))):
            raise ValueError("branch theta keys must be child labels or (parent_label, child_label) pairs; got %r." % (key,))
        try:
            v = float(value)
        except (TypeError, ValueError) as exc:
            raise ValueError("branch theta for %r must be numeric; got %r." % (key, value)) from exc
        if not math.isfinite(v) or v <= 0.0:
            raise ValueError("branch theta for %r must be finite and positive; got %s." % (key, v))


def msnc_eval_packed(net: SpeciesNetwork, tree_off: np.ndarray, parent: np.ndarray, blen: np.ndarray,
                     node_species: np.ndarray, theta: float, context: Optional[DeviceContext] = None,
                     branch_thetas: Optional[dict] = None) -> np.ndarray:
    """``pnb_msnc_eval`` on CSR-packed gene trees (the form candidate batches are built in)."""
    ctx = context or default_context()
    n = int(tree_off.shape[0]) - 1
    assert tree_off.dtype == np.int64 and parent.dtype == np.int32 and blen.dtype == np.float64
    assert node_species.dtype == np.int32
    out = np.empty(n, dtype=np.float64)
    _lib.check(_lib.load().pnb_msnc_eval(ctx.handle, net.handle(ctx, branch_thetas), n, _lib.ptr(tree_off), _lib.ptr(parent),
                                         _lib.ptr(blen), _lib.ptr(node_species), float(theta), _lib.ptr(out)))
    return out


def msnc_log_density_batch(gene_trees: Sequence[Any], species_net: Any, species_of: Dict[str, str], theta: float,
                           *, branch_thetas: Optional[dict] = None, context: Optional[DeviceContext] = None) -> np.ndarray:
    """log P(g_j | Psi) for every gene tree in one device call; ``-inf`` where g_j does not embed."""
    _check_theta(theta)
    validate_branch_thetas(branch_thetas)
    net = as_species_network(species_net)
    flats = [flatten(g) for g in gene_trees]
    if not flats:
        return np.zeros(0, dtype=np.float64)
    sizes = np.fromiter((t.n_nodes for t in flats), dtype=np.int64, count=len(flats))
    off = np.zeros(len(flats) + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    parent = np.ascontiguousarray(np.concatenate([t.parent for t in flats]), dtype=np.int32)
    blen = np.ascontiguousarray(np.concatenate([t.blen for t in flats]), dtype=np.float64)
    nsp = np.ascontiguousarray(np.concatenate([net.species_ids_of(t, species_of) for t in flats]), dtype=np.int32)
    return msnc_eval_packed(net, off, parent, blen, nsp, theta, context, branch_thetas)


def gene_tree_msnc_log_density(gene_tree: Any, species_net: Any, species_of: Dict[str, str], *, theta: float = 0.02,
                               branch_thetas: Optional[dict] = None, context: Optional[DeviceContext] = None) -> float:
    """Log timed MSNC density log P(g | Psi) (reference ``gene_tree_msnc_log_density``, ``:496-526``)."""
    return float(msnc_log_density_batch([gene_tree], species_net, species_of, theta, branch_thetas=branch_thetas,
                                        context=context)[0])


def _check_theta(theta: Any) -> float:
    # _positive_theta(theta, "default theta"), src/_units.py:59-68
    try:
        value = float(theta)
    except (TypeError, ValueError) as exc:
        raise ValueError("default theta must be numeric; got %r." % (theta,)) from exc
    if not math.isfinite(value) or value <= 0.0:
        raise ValueError("default theta must be finite and positive; got %s." % (value,))
    return value
