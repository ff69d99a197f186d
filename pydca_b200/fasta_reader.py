"""FASTA ingest with the semantics of pydca/fasta_reader/fasta_reader.py, without
Biopython: records are parsed, upper-cased, encoded and de-duplicated (first
occurrence kept, file order) by the C++ reader in csrc/fasta.cpp.

Integer convention: the public functions return states 1..q like the reference's
Python reader (fasta_reader.py:34-45); the CUDA side works on 0..q-1 (the C++
reader's convention, plmdca_numerics.cpp:708-731) -- ``*_zero_based`` variants
return that directly as a uint8 array.
"""
import numpy as np

from . import _lib

__all__ = ["get_alignment_from_fasta_file", "get_alignment_int_form", "alignment_letter2int",
           "first_record_length", "get_alignment_zero_based", "alignment_zero_based",
           "res_to_char", "get_alignment_char_form", "sequences_to_char_form"]

RES_TO_INT_ALL = {
    "PROTEIN": {c: i + 1 for i, c in enumerate("ACDEFGHIKLMNPQRSTVWY")},
    "RNA": {c: i + 1 for i, c in enumerate("ACGU")},
}
for _m, _q in (("PROTEIN", 21), ("RNA", 5)):
    for _g in "-.~":
        RES_TO_INT_ALL[_m][_g] = _q


class FastaReaderError(Exception):
    """Raised for unreadable or empty alignments."""


def _records(file_name):
    name, chunks = None, []
    with open(file_name) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if name is not None:
                    yield name, "".join(chunks)
                name, chunks = line[1:], []
            else:
                if name is None:
                    name = ""
                chunks.append(line)
    if name is not None:
        yield name, "".join(chunks)


def get_alignment_from_fasta_file(file_name):
    """List of upper-cased sequence strings (fasta_reader.py:79-119)."""
    alignment = [seq.upper() for _, seq in _records(file_name) if seq]
    if not alignment:
        raise ValueError("no sequences found in %s" % file_name)
    return alignment


def first_record_length(file_name):
    """(number of records, length of record 0): what PlmDCA needs (plmdca.py:177-180)."""
    n, length = 0, None
    for _, seq in _records(file_name):
        if not seq:
            continue
        if length is None:
            length = len(seq)
        n += 1
    if length is None:
        raise ValueError("no sequences found in %s" % file_name)
    return n, length


def first_sequence_length(file_name):
    """Length of record 0 without walking the whole file (the C reader does the rest)."""
    length, started = 0, False
    with open(file_name) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if started and length:
                    break
                started = True
            else:
                started = True
                length += len(line)
    if not length:
        raise ValueError("no sequences found in %s" % file_name)
    return length


def _num_states(biomolecule):
    b = biomolecule.strip().upper()
    if b not in ("PROTEIN", "RNA"):
        raise ValueError("%s entered. Biomolecule must be either PROTEIN or RNA" % biomolecule)
    return b, (21 if b == "PROTEIN" else 5)


def get_alignment_zero_based(file_name, biomolecule="protein"):
    """uint8[N, L], states 0..q-1, unique rows in first-occurrence order."""
    b, _ = _num_states(biomolecule)
    return _lib.read_fasta(file_name, b, first_sequence_length(file_name))


def alignment_zero_based(alignment, biomolecule="protein"):
    """Same as get_alignment_zero_based for an in-memory list of sequences
    (fasta_reader.py:122-163 semantics: unknown letters -> gap, first occurrence kept)."""
    b, q = _num_states(biomolecule)
    table = np.full(256, q - 1, dtype=np.uint8)
    for ch, v in RES_TO_INT_ALL[b].items():
        table[ord(ch)] = v - 1
        table[ord(ch.lower())] = v - 1
    rows, seen = [], set()
    length = None
    for seq in alignment:
        s = str(seq)
        if length is None:
            length = len(s)
        if len(s) != length:
            raise ValueError("sequences of unequal length in alignment (%d vs %d)" % (len(s), length))
        enc = table[np.frombuffer(s.encode("latin-1"), dtype=np.uint8)]
        key = enc.tobytes()
        if key not in seen:
            seen.add(key)
            rows.append(enc)
    if not rows:
        raise ValueError("No data found in alignment in integer representation")
    return np.ascontiguousarray(np.stack(rows))


def get_alignment_int_form(file_name, biomolecule="protein"):
    """List of lists with states 1..q (fasta_reader.py:166-188)."""
    return (get_alignment_zero_based(file_name, biomolecule).astype(np.int64) + 1).tolist()


def alignment_letter2int(alignment, biomolecule="protein"):
    """fasta_reader.py:122-163."""
    return (alignment_zero_based(alignment, biomolecule).astype(np.int64) + 1).tolist()


def res_to_char(biomolecule):
    """{state (1..q): letter}, the inverse of the residue table with '-' as the one gap
    symbol (fasta_reader.py:53-76)."""
    table = RES_TO_INT_ALL[biomolecule.strip().upper()]
    return {val: key for key, val in table.items() if key not in (".", "~")}


def sequences_to_char_form(seqs_lst, biomolecule):
    """Integer sequences (states 1..q) -> strings (fasta_reader.py:227-249)."""
    to_char = res_to_char(biomolecule)
    return ["".join(to_char[int(res)] for res in seq) for seq in seqs_lst]


def get_alignment_char_form(file_name, biomolecule="PROTEIN"):
    """The unique sequences of a FASTA file as strings over the standard alphabet: unknown
    letters and all gap aliases come back as '-' (fasta_reader.py:191-224)."""
    return sequences_to_char_form(get_alignment_int_form(file_name, biomolecule=biomolecule), biomolecule)
