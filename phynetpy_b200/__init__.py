"""phynetpy_b200 -- B200-native sequence-likelihood hot path of PhyNetPy.

Felsenstein pruning under GTR-family substitution models, as evaluated by
MCMC_SEQ for every locus on every proposal, in hand-written sm_100a CUDA behind a
C ABI (``include/phynet_b200.h``).  The Python surface mirrors the reference's
``phynetpy._seq_likelihood`` / ``phynetpy.GTR`` / ``_SeqLikelihoodEngine`` for this
path only; see DESIGN.md for scope and INTEGRATION.md for wiring into PhyNetPy.
"""
from ._lib import PhyNetB200Error
from ._seq_likelihood import (
    DNA_STATES, FelsensteinCalculator, HKY85, JC69, SubstitutionModel, tip_partials_for_char,
)
from ._seq_likelihood import GTR as SeqGTR  # `phynetpy_b200.GTR` is the public-model MODULE (as in the reference)
from ._engine import SeqLikelihoodEngine
from ._msnc_density import SpeciesNetwork, gene_tree_msnc_log_density, msnc_log_density_batch
from .device import DeviceContext, default_context
from .distributed import ShardPlan, balanced_ranges
from .tree import FlatTree, GeneTree, flatten

__version__ = "0.1.0"

__all__ = [
    "PhyNetB200Error", "DNA_STATES", "FelsensteinCalculator", "SeqGTR", "HKY85", "JC69", "SubstitutionModel",
    "tip_partials_for_char", "SeqLikelihoodEngine", "DeviceContext", "default_context", "ShardPlan",
    "balanced_ranges", "FlatTree", "GeneTree", "flatten", "SpeciesNetwork", "gene_tree_msnc_log_density",
    "msnc_log_density_batch",
]
