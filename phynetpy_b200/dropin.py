"""Run the reference's own ``MCMC_SEQ`` on top of phynetpy_b200, without editing the reference.

``install()`` applies, to an already importable reference package (``phynetpy``), exactly the two
changes INTEGRATION.md sections 2 and 3 describe as source edits:

* ``FelsensteinCalculator`` -- every module of the reference that imported the class
  (``src/_mcmc_seq.py:86-90``, ``src/infer.py:105``, ...) gets the B200 class instead, so
  ``_SeqLikelihoodEngine.__init__`` (``src/_mcmc_seq.py:673``) builds device-resident loci;
* ``_network_only_log_score`` (``src/_mcmc_seq.py:1163-1214``) -- the placement scan of the informed add / delete /
  relocate moves scores every candidate network with ``n_loci x 5`` host densities (94 % of the reference's chain time
  on 30 loci); the replacement asks the engine for the same sum through ONE ``pnb_msnc_eval`` per candidate
  (``SeqLikelihoodEngine.network_only_msnc_score``) and adds the reference's own prior;
* ``_SeqLikelihoodEngine`` (``src/_mcmc_seq.py:647-772``) -- replaced by
  :class:`B200SeqLikelihoodEngine`, the same seam (``log_likelihood`` / ``invalidate_locus`` /
  ``invalidate_network`` / ``clone_caches`` / ``restore_caches``, ``_felsen`` / ``_msnc`` lists)
  with every ``None`` entry filled by ONE ``pnb_eval`` / ONE ``pnb_msnc_eval``.

Everything above the seam -- ``SeqState``, every ``op_*`` and its undo closure, ``MCMCSeqKernel``,
``MCMC_SEQ.score`` / ``search`` / ``_search_mc3``, ``SeqBayesEngine`` and ``infer(...)`` -- runs as
the reference's unmodified Python.  The substitution-model classes are left alone: reference model
objects are recognised by ``_seq_likelihood._device_kind`` and their eigen-system is handed to K-a.

``run_parallel_chains`` (``src/_mcmc_seq.py:4076-4274``) pickles the sampler into ``spawn``ed worker
processes, which import a fresh, unpatched reference.  ``install()`` therefore also replaces the
worker entry point ``_chain_worker`` (``:3981``) by :func:`_chain_worker` below: the child imports
this module (the function is pickled by name), installs the drop-in, pins its GPU --
**device policy: chain ``c`` runs on ``devices[c % len(devices)]``** (``devices`` from ``install`` /
``PHYNET_B200_DEVICES``, default: every visible GPU) -- and then runs the reference's worker.

There is no CPU fallback behind any of this: without the built library and a CUDA device the
first calculator construction raises.
"""
from __future__ import annotations

import os
import sys
from typing import Any, Optional, Sequence

from . import _lib
from ._engine import SeqLikelihoodEngine
from ._seq_likelihood import FelsensteinCalculator

__all__ = ["install", "uninstall", "installed", "B200SeqLikelihoodEngine"]

_ENV_DEVICES = "PHYNET_B200_DEVICES"
_ENV_PRELOAD = "PHYNETPY_B200_PRELOAD"   # "module:function" run when this module is imported in a chain worker
_state: dict = {}


def _preload() -> None:
    """A spawned chain worker unpickles this module's ``_chain_worker`` BEFORE the sampler, i.e. before it
    imports the reference.  An installation whose reference needs preparation before import (a path entry,
    stand-ins for optional I/O dependencies) names a callable here; it runs once per process."""
    spec = os.environ.get(_ENV_PRELOAD, "")
    if spec and os.environ.get("PHYNETPY_B200_DROPIN"):
        import importlib
        mod, _, fn = spec.partition(":")
        getattr(importlib.import_module(mod), fn)()


_preload()


class _EngineCalc:
    """What ``engine._calcs[i]`` hands to the reference's coupled moves.

    ``_candidate_logweights`` / ``_enumerate_scored_candidates`` (``src/_mcmc_seq.py:1882,1957``) call
    ``eng._calcs[i].log_likelihood(tree, model)`` for every candidate tree of a locus, between two
    engine evaluations.  Those calls must not disturb the engine's device state (the last accepted
    proposal may still be pending; the committed partials are what the next dirty-path evaluation
    starts from), so they are routed to the engine's score-only entry: committed partials are read,
    nothing is written (``PNB_EVAL_NOSTORE``).  Everything else is the calculator's own.
    """

    __slots__ = ("_engine", "_i", "_calc")

    def __init__(self, engine: "B200SeqLikelihoodEngine", i: int, calc: FelsensteinCalculator) -> None:
        self._engine = engine
        self._i = i
        self._calc = calc

    def __getattr__(self, name: str) -> Any:
        return getattr(self._calc, name)

    def log_likelihood(self, gene_tree: Any, model: Any, *, site_rate: float = 1.0) -> float:
        eng = self._engine
        if model is not eng._model:   # not the engine's model: a plain evaluation of a settled state
            eng._commit_pending()
            return self._calc.log_likelihood(gene_tree, model, site_rate=site_rate)
        return float(eng.score_candidates(self._i, [gene_tree], site_rate=site_rate)[0])


class B200SeqLikelihoodEngine(SeqLikelihoodEngine):
    """``_SeqLikelihoodEngine`` of the reference (``src/_mcmc_seq.py:647-772``) on the device:
    same constructor arguments, same attributes read by the operators (``n_loci``, ``_calcs``,
    ``_species_of``, ``_model``, ``_felsen``, ``_msnc``)."""

    def __init__(self, loci: Sequence[Any], species_of: Optional[dict], model: Any,
                 branch_thetas: Optional[dict] = None) -> None:
        super().__init__(loci, species_of, model, branch_thetas)
        self._calcs = [_EngineCalc(self, i, c) for i, c in enumerate(self._calcs)]
        self._gt_obj: list = [None] * self.n_loci
        self._gt_index: list = [None] * self.n_loci

    def _gene_index(self, i: int, gene_tree: Any):
        """Locus ``i``'s gene-tree view + coalescence events in the REFERENCE's own representation
        (``_SeqLikelihoodEngine._gene_index``, ``src/_mcmc_seq.py:714-721``).  The placement scan of the
        informed add / delete moves (``_network_only_log_score``, ``:1196-1199``) asks the engine for it
        and feeds it to the reference's host density; it is built by the reference's own
        ``build_gene_tree_msnc_index`` and cached by tree identity exactly as there."""
        if gene_tree is not self._gt_obj[i]:
            self._gt_index[i] = _state["build_gene_tree_msnc_index"](gene_tree, self._species_of)
            self._gt_obj[i] = gene_tree
        return self._gt_index[i]


def _network_only_log_score(state: Any, net: Any) -> float:
    """``_network_only_log_score`` of the reference (``src/_mcmc_seq.py:1163-1214``) with the densities on the device.
    Anything this path cannot take (an engine that is not ours, a network the flat form rejects) goes to the
    reference's own function."""
    import math
    eng = getattr(state, "_engine", None)
    ref_fn = _state["ref_network_only_log_score"]
    if not isinstance(eng, B200SeqLikelihoodEngine) or getattr(state, "branch_thetas", None) != eng._branch_thetas:
        return ref_fn(state, net)
    mcmc = _state["mcmc"]
    try:
        total = eng.network_only_msnc_score(state.gene_trees, net, state.theta, mcmc._HEIGHT_SCALE_GRID)
    except Exception:   # whatever the flat form or the device cannot take: the reference's own host loop decides
        return ref_fn(state, net)
    if not math.isfinite(total):
        return float("-inf")
    total += mcmc.log_prior_seq(net, state.theta, state.priors)
    return total if math.isfinite(total) else float("-inf")


def _reference_modules(ref_pkg: Any):
    prefix = ref_pkg.__name__ + "."
    for name, mod in list(sys.modules.items()):
        if mod is not None and (name == ref_pkg.__name__ or name.startswith(prefix)):
            yield mod


def installed() -> bool:
    return bool(_state)


def install(ref_pkg: Any = None, *, devices: Optional[Sequence[int]] = None) -> None:
    """Patch the imported reference package (default: ``sys.modules['phynetpy']``).  Idempotent."""
    if _state:
        return
    _lib.load()  # fail now, loudly, if the native library has not been built
    if ref_pkg is None:
        import importlib
        ref_pkg = importlib.import_module("phynetpy")
    import importlib
    seq = importlib.import_module(ref_pkg.__name__ + "._seq_likelihood")
    mcmc = importlib.import_module(ref_pkg.__name__ + "._mcmc_seq")
    ref_calc = seq.FelsensteinCalculator
    undo = []
    for mod in _reference_modules(ref_pkg):
        for attr, val in list(vars(mod).items()):
            if val is ref_calc:
                undo.append((mod, attr, val))
                setattr(mod, attr, FelsensteinCalculator)
    undo.append((mcmc, "_SeqLikelihoodEngine", mcmc._SeqLikelihoodEngine))
    mcmc._SeqLikelihoodEngine = B200SeqLikelihoodEngine
    undo.append((mcmc, "_network_only_log_score", mcmc._network_only_log_score))
    _state["ref_network_only_log_score"] = mcmc._network_only_log_score
    _state["mcmc"] = mcmc
    mcmc._network_only_log_score = _network_only_log_score
    undo.append((mcmc, "_chain_worker", mcmc._chain_worker))
    mcmc._chain_worker = _chain_worker
    if devices is not None:
        os.environ[_ENV_DEVICES] = ",".join(str(int(d)) for d in devices)
    os.environ["PHYNETPY_B200_DROPIN"] = ref_pkg.__name__   # inherited by spawned chain workers
    _state["undo"] = undo
    _state["build_gene_tree_msnc_index"] = mcmc.build_gene_tree_msnc_index
    _state["pkg"] = ref_pkg
    _state["ref_chain_worker"] = undo[-1][2]


def uninstall() -> None:
    """Restore the reference's own classes (objects built while installed keep theirs)."""
    for mod, attr, val in reversed(_state.get("undo", [])):
        setattr(mod, attr, val)
    _state.clear()
    os.environ.pop("PHYNETPY_B200_DROPIN", None)


def chain_device(chain_id: int) -> int:
    """Device policy of ``run_parallel_chains``: chain c -> devices[c % len(devices)]."""
    spec = os.environ.get(_ENV_DEVICES, "").strip()
    if spec:
        devices = [int(x) for x in spec.split(",") if x.strip() != ""]
    else:
        import ctypes as C
        n = C.c_int(0)
        _lib.check(_lib.load().pnb_device_count(C.byref(n)))
        devices = list(range(max(n.value, 1)))
    return devices[chain_id % len(devices)]


def _chain_worker(sampler: Any, chain_id: int, *args: Any, **kwargs: Any) -> None:
    """Worker entry point of ``run_parallel_chains`` (replaces ``src/_mcmc_seq.py:3981``): runs in a
    spawned process, installs the drop-in there, pins the chain's GPU, then runs the reference's worker."""
    import importlib
    pkg = importlib.import_module(os.environ.get("PHYNETPY_B200_DROPIN") or type(sampler).__module__.split(".")[0])
    install(pkg)
    os.environ["PHYNET_B200_DEVICE"] = str(chain_device(int(chain_id)))
    return _state["ref_chain_worker"](sampler, chain_id, *args, **kwargs)
