"""Alignment text formats -> {label: sequence} (host logic, no device)."""
import pytest

from phynetpy_b200 import ingest as I

ALN = {"A": "ACGTACGTAC", "B": "ACGTACGAAC", "C_long_name": "ACG-TCGTNG"}


def test_fasta_wrapped_and_commented(tmp_path):
    text = ";comment\n>A first\nACGTA\nCGTAC\n\n>B\nACGTACGAAC\n>C_long_name x y\nACG-T\nCGTNG\n"
    assert I.parse_fasta(text) == ALN
    p = tmp_path / "x.fasta"
    p.write_text(text)
    assert I.read_alignment(p) == ALN and I.read_alignment(str(p), "fasta") == ALN


def test_fasta_errors():
    with pytest.raises(I.AlignmentFormatError, match="before the first"):
        I.parse_fasta("ACGT\n>A\nAC\n")
    with pytest.raises(I.AlignmentFormatError, match="duplicate"):
        I.parse_fasta(">A\nAC\n>A\nAC\n")
    with pytest.raises(I.AlignmentFormatError, match="differ in length"):
        I.parse_fasta(">A\nAC\n>B\nACG\n")
    with pytest.raises(I.AlignmentFormatError, match="no sequences"):
        I.parse_fasta("\n\n")


NEXUS = """#NEXUS
[a comment]
BEGIN TAXA; DIMENSIONS NTAX=3; TAXLABELS A B C_long_name; END;
Begin data;
  Dimensions ntax=3 nchar=10;
  Format datatype=dna missing=? gap=-;
  Matrix
    A            ACGTA [inline] CGTAC
    B            ACGTACGAAC
    'C_long_name' ACG-TCGTNG
  ;
End;
begin sets;
  charset locus1 = 1-4;
  charset locus2 = 5 - 10;
end;
"""


def test_nexus_sequential_interleaved_and_charsets(tmp_path):
    assert I.parse_nexus(NEXUS) == ALN
    inter = NEXUS.replace("""    A            ACGTA [inline] CGTAC
    B            ACGTACGAAC
    'C_long_name' ACG-TCGTNG""", """    A ACGTA
    B ACGTA
    C_long_name ACG-T

    A CGTAC
    B CGAAC
    C_long_name CGTNG""")
    assert I.parse_nexus(inter) == ALN
    cs = I.parse_nexus_charsets(NEXUS)
    assert cs == {"locus1": (0, 4), "locus2": (4, 10)}
    loci = I.split_loci(ALN, [cs["locus1"], cs["locus2"]])
    assert loci[0]["A"] == "ACGT" and loci[1]["C_long_name"] == "TCGTNG"
    p = tmp_path / "x.nex"
    p.write_text(NEXUS)
    assert I.read_alignment(p) == ALN and I.read_nexus_charsets(p) == cs
    q = tmp_path / "x.unknown_ext"
    q.write_text(NEXUS)
    assert I.read_alignment(q) == ALN            # the reference defaults to NEXUS (src/data/_sequence.py:277-283)


def test_nexus_errors():
    with pytest.raises(I.AlignmentFormatError, match="#NEXUS"):
        I.parse_nexus("begin data; matrix A AC; end;")
    with pytest.raises(I.AlignmentFormatError, match="no DATA"):
        I.parse_nexus("#NEXUS\nbegin trees; tree t = (A,B); end;")
    with pytest.raises(I.AlignmentFormatError, match="NCHAR"):
        I.parse_nexus(NEXUS.replace("nchar=10", "nchar=11"))
    with pytest.raises(ValueError, match="outside"):
        I.split_loci(ALN, [(0, 11)])


def test_phylip_sequential_and_interleaved(tmp_path):
    seq = "3 10\nA ACGTACGTAC\nB ACGTA CGAAC\nC_long_name ACG-TCGTNG\n"
    assert I.parse_phylip(seq) == ALN
    inter = "3 10\nA ACGTA\nB ACGTA\nC_long_name ACG-T\n\nCGTAC\nCGAAC\nCGTNG\n"
    assert I.parse_phylip(inter) == ALN
    p = tmp_path / "x.phy"
    p.write_text(seq)
    assert I.read_alignment(p) == ALN
    with pytest.raises(I.AlignmentFormatError, match="ntax nchar"):
        I.parse_phylip("A ACGT\n")
    with pytest.raises(I.AlignmentFormatError, match="columns announced"):
        I.parse_phylip("2 5\nA ACGT\nB ACGT\n")
    with pytest.raises(I.AlignmentFormatError, match="unknown alignment format"):
        I.read_alignment(p, "vcf")
