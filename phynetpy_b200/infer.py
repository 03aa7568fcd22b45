"""Public re-exports for the sequence-likelihood path, mirroring what the reference
exposes from ``phynetpy.infer`` (``src/infer.py:105,134,229``): the Felsenstein
calculator and the ``_seq_likelihood`` substitution models.  (``phynetpy_b200.GTR``
is the *module* of public ``expt``-style models, as ``phynetpy.GTR`` is.)"""
from ._seq_likelihood import (  # noqa: F401
    DNA_STATES, FelsensteinCalculator, GTR, HKY85, JC69, SubstitutionModel, tip_partials_for_char,
)
from ._engine import SeqLikelihoodEngine  # noqa: F401
from ._msnc_density import SpeciesNetwork, gene_tree_msnc_log_density, msnc_log_density_batch  # noqa: F401
from ._lib import PhyNetB200Error  # noqa: F401
from .device import DeviceContext, default_context  # noqa: F401
from .tree import FlatTree, GeneTree, flatten  # noqa: F401
