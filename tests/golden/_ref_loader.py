"""Import the UNMODIFIED reference (PhyNetPy v0.6.0) for golden-vector generation.

Only used by ``make_golden.py`` in the build container; never at test/bench time
(``/root/reference`` does not exist on the GPU box).  Recipe = SURVEY.md section 8c:

    cp -r /root/reference /tmp/ref/refcopy && chmod -R u+w /tmp/ref/refcopy
    (cd /tmp/ref/refcopy && python setup.py build_ext --inplace)
    ln -sfn refcopy/src /tmp/ref/phynetpy

``Bio`` / ``nexus`` are imported at module scope by reference I/O modules that are
not on the numeric path, so they are stubbed before import.
"""
import sys
import types

REF_ROOT = "/tmp/ref"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference():
    if "phynetpy" in sys.modules:
        return sys.modules["phynetpy"]

    class NexusError(Exception):
        pass

    bio = _stub("Bio")
    bio.__path__ = []
    nx = _stub("Bio.Nexus")
    nx.__path__ = []
    nxn = _stub("Bio.Nexus.Nexus", NexusError=NexusError)
    nx.Nexus = nxn
    bio.Nexus = nx
    bio.AlignIO = _stub("Bio.AlignIO", MultipleSeqAlignment=type("MultipleSeqAlignment", (), {}))
    _stub("Bio.Align", MultipleSeqAlignment=type("MultipleSeqAlignment", (), {}))
    bio.SeqRecord = _stub("Bio.SeqRecord", SeqRecord=type("SeqRecord", (), {}))
    bio.Seq = _stub("Bio.Seq", Seq=type("Seq", (), {}))
    bio.Phylo = _stub("Bio.Phylo")
    bio.SeqIO = _stub("Bio.SeqIO")
    _stub("nexus", NexusReader=type("NexusReader", (), {}), NexusWriter=type("NexusWriter", (), {}))
    sys.path.insert(0, REF_ROOT)
    import phynetpy  # noqa: F401
    return phynetpy
