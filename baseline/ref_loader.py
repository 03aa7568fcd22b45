"""Import the UNMODIFIED reference (PhyNetPy v0.6.0) from ``baseline/_ref``.

``baseline/_ref`` is produced by ``baseline/install_reference.sh`` (pip install of a scratch copy of
``/root/reference`` with its four Cython extensions compiled; git-ignored, but it travels to the GPU
box with the tree).  It is used by

* ``bench.py --impl reference``  -- the reference arm (kind "reference"),
* ``tests/test_gpu_dropin.py``   -- the reference's own MCMC_SEQ driven on top of phynetpy_b200,
* the golden generators under ``tests/golden/``.

``Bio`` (biopython) and ``nexus`` (python-nexus) are imported at module scope by reference I/O
modules that are not on the numeric path (``MSA.py``, ``IO.py``, ``GeneTrees.py``); neither is
installed in this image, so inert stand-ins are registered before the import.  Nothing of the
reference is modified.
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "phynetpy"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _stub_optional_io_deps() -> None:
    try:
        import Bio  # noqa: F401
    except ImportError:
        class NexusError(Exception):
            pass

        bio = _stub("Bio")
        bio.__path__ = []
        nx = _stub("Bio.Nexus")
        nx.__path__ = []
        nx.Nexus = _stub("Bio.Nexus.Nexus", NexusError=NexusError)
        bio.Nexus = nx
        bio.AlignIO = _stub("Bio.AlignIO", MultipleSeqAlignment=type("MultipleSeqAlignment", (), {}))
        _stub("Bio.Align", MultipleSeqAlignment=type("MultipleSeqAlignment", (), {}))
        bio.SeqRecord = _stub("Bio.SeqRecord", SeqRecord=type("SeqRecord", (), {}))
        bio.Seq = _stub("Bio.Seq", Seq=type("Seq", (), {}))
        bio.Phylo = _stub("Bio.Phylo")
        bio.SeqIO = _stub("Bio.SeqIO")
    try:
        import nexus  # noqa: F401
    except ImportError:
        _stub("nexus", NexusReader=type("NexusReader", (), {}), NexusWriter=type("NexusWriter", (), {}))


def load_reference():
    """``import phynetpy`` from ``baseline/_ref`` (raises ImportError when it has not been installed)."""
    if "phynetpy" in sys.modules:
        return sys.modules["phynetpy"]
    if not available():
        raise ImportError("baseline/_ref/phynetpy is missing: run baseline/install_reference.sh in the build container")
    _stub_optional_io_deps()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import phynetpy  # noqa: F401
    return phynetpy
