"""GPU: the drop-in proof.  The UNMODIFIED reference (``baseline/_ref``, built by
``baseline/install_reference.sh``) runs its own ``MCMC_SEQ.score()`` / ``search()`` / ``run_parallel_chains``
with ``phynetpy_b200.dropin.install()`` applied -- the two changes of INTEGRATION.md sections 2 + 3 -- so every
Felsenstein pruning and every MSNC density of the chain is computed by the kernels (reference call sites:
``src/_mcmc_seq.py:3274-3276`` score, ``:3369`` search, ``:647-772`` the engine seam, ``:1882/:1957`` the coupled
moves' direct calls, ``:4076`` parallel chains).  See ``tests/dropin_common.py`` for why equality of accept /
reject decisions is checked inside one chain rather than between two runs."""
import math
import os

import numpy as np
import pytest

import dropin_common as DC

pytestmark = pytest.mark.gpu


@pytest.fixture
def world():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    ref = DC.require_reference()
    import phynetpy._mcmc_seq as mcmc
    from phynetpy_b200 import dropin
    dropin.install(ref)
    import types
    yield types.SimpleNamespace(ref=ref, mcmc=mcmc, dropin=dropin)
    dropin.uninstall()


@pytest.mark.parametrize("taxa,n_loci,n_sites", [("ABC", 2, 150), ("ABCD", 4, 200), ("ABCDEF", 6, 300)])
def test_reference_score_on_the_device(world, taxa, n_loci, n_sites):
    """MCMC_SEQ(...).score() with the drop-in == the unpatched reference, 1e-10 relative."""
    loci, mapping = DC.make_loci(seed=3, n_sites=n_sites, n_loci=n_loci, taxa=taxa)
    pri = world.mcmc.MCMCSeqPriors(max_reticulations=1)
    got = world.mcmc.MCMC_SEQ(loci, mapping, priors=pri).score()
    world.dropin.uninstall()
    want = world.mcmc.MCMC_SEQ(loci, mapping, priors=pri).score()
    world.dropin.install(world.ref)
    assert math.isfinite(want)
    assert abs(got - want) <= 1e-10 * abs(want), (got, want)


def test_reference_gtr_model_objects_are_served_by_the_device(world):
    """A reference GTR model object (its eigen-system goes to K-a) and the reference's HKY85."""
    import phynetpy._seq_likelihood as rsl
    loci, mapping = DC.make_loci(seed=4, n_sites=180, n_loci=3)
    for model in (rsl.GTR([0.1, 0.2, 0.3, 0.4], [1.0, 2.5, 0.7, 1.3, 3.1, 0.9]), rsl.HKY85(2.5, [0.3, 0.2, 0.2, 0.3])):
        got = world.mcmc.MCMC_SEQ(loci, mapping, model=model).score()
        world.dropin.uninstall()
        want = world.mcmc.MCMC_SEQ(loci, mapping, model=model).score()
        world.dropin.install(world.ref)
        assert abs(got - want) <= 1e-10 * abs(want), (type(model).__name__, got, want)


def test_reference_search_runs_on_the_device(world):
    """200 iterations of the reference's own search() (src/_mcmc_seq.py:3369) with every likelihood on the GPU."""
    from phynetpy_b200 import default_context
    loci, mapping = DC.make_loci(seed=1, n_loci=3)
    inf = world.mcmc.MCMC_SEQ(loci, mapping, priors=world.mcmc.MCMCSeqPriors(max_reticulations=1))
    res = inf.search(num_iter=200, burn_in=20, sample_freq=20, seed=7)
    assert math.isfinite(res.map_log_posterior) and 0.0 < res.acceptance_rate < 1.0 and len(res.samples) == 9
    assert default_context().last_stats()["launches"] >= 0
    # the MAP state the chain reports is reproduced by the unpatched reference
    world.dropin.uninstall()
    again = world.mcmc.MCMC_SEQ(loci, mapping, species_net=res.map_network, theta=res.map_theta,
                                priors=world.mcmc.MCMCSeqPriors(max_reticulations=1))
    assert math.isfinite(again.score())
    world.dropin.install(world.ref)


@pytest.mark.parametrize("coupled", [False, True])
def test_lockstep_chain_same_values_same_decisions(world, coupled):
    """Every likelihood the chain evaluates, by both implementations on the same proposal: values within 1e-10,
    every per-locus cache entry within 1e-10, identical accept / reject decisions, no exception out of the drop-in,
    cached == from-scratch at the end."""
    loci, mapping = DC.make_loci(seed=2, n_loci=3)
    n_iter = 200
    log = DC.run_lockstep_chain(world.mcmc, loci, mapping, n_iter, seed=5, max_reticulations=2, coupled=coupled)
    DC.check_lockstep_log(log, n_iter)
    if coupled:
        assert len(log["candidates"]) > 100


def test_mc3_ladder_runs_on_the_device(world):
    loci, mapping = DC.make_loci(seed=6, n_loci=2)
    inf = world.mcmc.MCMC_SEQ(loci, mapping, priors=world.mcmc.MCMCSeqPriors(max_reticulations=1))
    res = inf.search(num_iter=60, burn_in=10, sample_freq=10, seed=3, temperatures=[1.0, 2.0, 4.0])
    assert math.isfinite(res.map_log_posterior) and len(res.samples) == 5


def test_run_parallel_chains_device_policy(world, monkeypatch):
    """run_parallel_chains (src/_mcmc_seq.py:4076): the sampler is pickled into spawned workers; each worker
    installs the drop-in, takes device chain_id % n_devices and runs the reference's chain on it."""
    monkeypatch.setenv("PHYNETPY_B200_PRELOAD", "baseline.ref_loader:load_reference")
    monkeypatch.setenv("PYTHONPATH", DC.ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    loci, mapping = DC.make_loci(seed=8, n_loci=2)
    inf = world.mcmc.MCMC_SEQ(loci, mapping, priors=world.mcmc.MCMCSeqPriors(max_reticulations=1))
    res = world.mcmc.run_parallel_chains(inf, n_chains=2, num_iter=60, burn_in=10, sample_freq=10, seed=11,
                                         check_every=20, progress=False)
    assert not res.errors, res.errors
    assert len(res.chains) == 2
    for r in res.chains:
        assert math.isfinite(r.map_log_posterior)
