"""Alignment files -> the ``{label: sequence}`` dictionaries ``FelsensteinCalculator`` takes
("FASTA/NEXUS -> codes ingest").

The reference reads alignments through Biopython / python-nexus (``_read_alignment``,
``src/data/_sequence.py:258-300`` -> ``src/IO.py``); neither library is on the numeric path and
neither is needed here: the formats are plain text.  Sequences are returned as written (case and
IUPAC / gap characters untouched) -- upper-casing and the character -> 4-bit mask table are applied
by pattern compression (K-d), exactly as ``FelsensteinCalculator.__init__`` does
(``src/_seq_likelihood.py:344-367``).

* :func:`read_fasta`   -- ``>label`` records, sequences possibly wrapped over several lines;
* :func:`read_nexus`   -- the MATRIX of a DATA / CHARACTERS block, sequential or interleaved;
* :func:`read_phylip`  -- relaxed PHYLIP (``ntax nchar`` header), sequential or interleaved;
* :func:`read_alignment` -- dispatch on the extension like the reference (default NEXUS);
* :func:`split_loci`   -- cut a concatenated alignment into loci by ``(start, end)`` column ranges
  (NEXUS ``CHARSET`` ranges are returned by :func:`read_nexus_charsets`).
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Sequence, Tuple


class AlignmentFormatError(ValueError):
    """The file does not parse as the format it was read as."""


def _check_equal_length(aln: Dict[str, str], what: str) -> Dict[str, str]:
    if not aln:
        raise AlignmentFormatError("%s: no sequences found" % what)
    lengths = {len(s) for s in aln.values()}
    if len(lengths) != 1:
        raise AlignmentFormatError("%s: sequences differ in length (%s)" % (what, sorted(lengths)))
    return aln


def parse_fasta(text: str) -> Dict[str, str]:
    aln: Dict[str, List[str]] = {}
    cur: Optional[str] = None
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line.startswith(";"):
            continue
        if line.startswith(">"):
            cur = line[1:].split()[0] if line[1:].split() else ""
            if not cur:
                raise AlignmentFormatError("FASTA: record without a label")
            if cur in aln:
                raise AlignmentFormatError("FASTA: duplicate label %r" % cur)
            aln[cur] = []
        else:
            if cur is None:
                raise AlignmentFormatError("FASTA: sequence data before the first '>' line")
            aln[cur].append("".join(line.split()))
    return _check_equal_length({k: "".join(v) for k, v in aln.items()}, "FASTA")


_NEXUS_COMMENT = re.compile(r"\[[^\]]*\]")


def _nexus_blocks(text: str) -> List[Tuple[str, str]]:
    text = _NEXUS_COMMENT.sub("", text)
    if not text.lstrip().upper().startswith("#NEXUS"):
        raise AlignmentFormatError("NEXUS: file does not start with #NEXUS")
    return [(m.group(1).upper(), m.group(2))
            for m in re.finditer(r"(?is)\bbegin\s+(\w+)\s*;(.*?)\bend\s*;", text)]


def parse_nexus(text: str) -> Dict[str, str]:
    for name, body in _nexus_blocks(text):
        if name not in ("DATA", "CHARACTERS"):
            continue
        m = re.search(r"(?is)\bmatrix\b(.*?);", body)
        if not m:
            raise AlignmentFormatError("NEXUS: %s block without a MATRIX" % name)
        aln: Dict[str, List[str]] = {}
        for raw in m.group(1).splitlines():
            parts = raw.split()
            if len(parts) < 2:
                continue
            label = parts[0].strip("'\"")
            aln.setdefault(label, []).append("".join(parts[1:]))   # interleaved blocks append
        out = _check_equal_length({k: "".join(v) for k, v in aln.items()}, "NEXUS")
        dim = re.search(r"(?i)\bnchar\s*=\s*(\d+)", body)
        if dim and int(dim.group(1)) != len(next(iter(out.values()))):
            raise AlignmentFormatError("NEXUS: NCHAR=%s but the matrix has %d columns"
                                       % (dim.group(1), len(next(iter(out.values())))))
        return out
    raise AlignmentFormatError("NEXUS: no DATA or CHARACTERS block")


def parse_nexus_charsets(text: str) -> Dict[str, Tuple[int, int]]:
    """``CHARSET name = a-b;`` declarations (SETS / ASSUMPTIONS / MRBAYES blocks) as 0-based half-open
    column ranges ``(a-1, b)``."""
    out: Dict[str, Tuple[int, int]] = {}
    for _, body in _nexus_blocks(text):
        for m in re.finditer(r"(?i)\bcharset\s+([^\s=]+)\s*=\s*(\d+)\s*-\s*(\d+)\s*;", body):
            out[m.group(1)] = (int(m.group(2)) - 1, int(m.group(3)))
    return out


def parse_phylip(text: str) -> Dict[str, str]:
    lines = [ln for ln in text.splitlines() if ln.strip()]
    if not lines:
        raise AlignmentFormatError("PHYLIP: empty file")
    head = lines[0].split()
    if len(head) < 2 or not (head[0].isdigit() and head[1].isdigit()):
        raise AlignmentFormatError("PHYLIP: first line must be 'ntax nchar'")
    ntax, nchar = int(head[0]), int(head[1])
    body = lines[1:]
    if len(body) < ntax:
        raise AlignmentFormatError("PHYLIP: %d taxa announced, %d lines found" % (ntax, len(body)))
    labels: List[str] = []
    seqs: List[List[str]] = []
    for ln in body[:ntax]:
        parts = ln.split()
        if len(parts) < 2:
            raise AlignmentFormatError("PHYLIP: line without sequence data: %r" % ln)
        labels.append(parts[0])
        seqs.append(["".join(parts[1:])])
    rest = body[ntax:]
    if rest:
        if sum(len(s[0]) for s in seqs) == ntax * nchar:
            raise AlignmentFormatError("PHYLIP: trailing lines after a complete sequential matrix")
        if len(rest) % ntax != 0:
            raise AlignmentFormatError("PHYLIP: interleaved blocks must have one line per taxon")
        for k, ln in enumerate(rest):                    # interleaved continuation blocks carry no labels
            seqs[k % ntax].append("".join(ln.split()))
    out = _check_equal_length({lab: "".join(s) for lab, s in zip(labels, seqs)}, "PHYLIP")
    if len(out) != ntax:
        raise AlignmentFormatError("PHYLIP: duplicate labels")
    if len(next(iter(out.values()))) != nchar:
        raise AlignmentFormatError("PHYLIP: %d columns announced, %d found" % (nchar, len(next(iter(out.values())))))
    return out


def _read(path) -> str:
    with open(os.fspath(path), "r") as f:
        return f.read()


def read_fasta(path) -> Dict[str, str]:
    return parse_fasta(_read(path))


def read_nexus(path) -> Dict[str, str]:
    return parse_nexus(_read(path))


def read_nexus_charsets(path) -> Dict[str, Tuple[int, int]]:
    return parse_nexus_charsets(_read(path))


def read_phylip(path) -> Dict[str, str]:
    return parse_phylip(_read(path))


def read_alignment(path, format: Optional[str] = None) -> Dict[str, str]:
    """Dispatch like the reference's ``_read_alignment`` (``src/data/_sequence.py:258-300``): ``format``
    or the file extension; unknown extensions mean NEXUS."""
    path_str = os.fspath(path)
    if format is None:
        suffix = os.path.splitext(path_str)[1].lower()
        if suffix in (".fa", ".fasta", ".fas", ".fna"):
            format = "fasta"
        elif suffix in (".phy", ".phylip"):
            format = "phylip"
        else:
            format = "nexus"
    fmt = format.lower()
    if fmt == "fasta":
        return read_fasta(path_str)
    if fmt == "nexus":
        return read_nexus(path_str)
    if fmt == "phylip":
        return read_phylip(path_str)
    raise AlignmentFormatError("unknown alignment format %r; expected 'nexus', 'fasta' or 'phylip'." % (format,))


def split_loci(aln: Dict[str, str], ranges: Sequence[Tuple[int, int]]) -> List[Dict[str, str]]:
    """Cut a concatenated alignment into per-locus alignments by 0-based half-open column ranges."""
    n = len(next(iter(aln.values()))) if aln else 0
    out = []
    for a, b in ranges:
        if not (0 <= a < b <= n):
            raise ValueError("split_loci: range (%d, %d) outside the %d columns" % (a, b, n))
        out.append({k: s[a:b] for k, s in aln.items()})
    return out
