"""SequenceBackmapper -- the class surface of pydca/sequence_backmapper/sequence_backmapper.py:25-466
without Biopython.

What the reference does (:228-283, :339-466): pairwise-align the reference sequence with EVERY
sequence of the alignment (Biopython ``pairwise2.align.localds``, score only) to find the best
matching "template" sequence, align that one pair once more with a traceback, and map MSA
columns to reference positions through the template's gap pattern.

Here the N score-only alignments -- the data-parallel part -- run on the GPU
(``dca_sw_scores``, csrc/align.cu: one thread per sequence, integer Smith-Waterman / Gotoh,
BLOSUM62 -10/-1 for proteins, +5/-4 and -8/0 for RNA as at :207-214); the one traceback and
the column bookkeeping (:339-466, restated) are host code.  Parity status: scores are the published
recurrence (pinned against oracle/sw_oracle.py); where several alignments share the best score
Biopython's choice among them is not reproducible without Biopython (absent from the image and
from /root/reference): the traceback here prefers a substitution over a gap in the template
over a gap in the reference and ends at the first best cell.  When the reference sequence is a
member of the alignment (the documented use, and the notebook example) the best alignment is
unique and gap-free.
"""
import logging

import numpy as np

from . import _lib, fasta_reader

logger = logging.getLogger(__name__)

_PROTEIN = "ACDEFGHIKLMNPQRSTVWY"
_RNA = "ACGU"

# BLOSUM62 on the 20 standard residues in NCBI order (public table; csrc/align.cu holds the same
# numbers in the readers' alphabet order and tests/test_seqsum_model.py-style CPU tests compare them)
_BLOSUM_ORDER = "ARNDCQEGHILKMFPSTWYV"
_BLOSUM62 = np.array([
    [4, -1, -2, -2, 0, -1, -1, 0, -2, -1, -1, -1, -1, -2, -1, 1, 0, -3, -2, 0],
    [-1, 5, 0, -2, -3, 1, 0, -2, 0, -3, -2, 2, -1, -3, -2, -1, -1, -3, -2, -3],
    [-2, 0, 6, 1, -3, 0, 0, 0, 1, -3, -3, 0, -2, -3, -2, 1, 0, -4, -2, -3],
    [-2, -2, 1, 6, -3, 0, 2, -1, -1, -3, -4, -1, -3, -3, -1, 0, -1, -4, -3, -3],
    [0, -3, -3, -3, 9, -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1],
    [-1, 1, 0, 0, -3, 5, 2, -2, 0, -3, -2, 1, 0, -3, -1, 0, -1, -2, -1, -2],
    [-1, 0, 0, 2, -4, 2, 5, -2, 0, -3, -3, 1, -2, -3, -1, 0, -1, -3, -2, -2],
    [0, -2, 0, -1, -3, -2, -2, 6, -2, -4, -4, -2, -3, -3, -2, 0, -2, -2, -3, -3],
    [-2, 0, 1, -1, -3, 0, 0, -2, 8, -3, -3, -1, -2, -1, -2, -1, -2, -2, 2, -3],
    [-1, -3, -3, -3, -1, -3, -3, -4, -3, 4, 2, -3, 1, 0, -3, -2, -1, -3, -1, 3],
    [-1, -2, -3, -4, -1, -2, -3, -4, -3, 2, 4, -2, 2, 0, -3, -2, -1, -2, -1, 1],
    [-1, 2, 0, -1, -3, 1, 1, -2, -1, -3, -2, 5, -1, -3, -1, 0, -1, -3, -2, -2],
    [-1, -1, -2, -3, -1, 0, -2, -3, -2, 1, 2, -1, 5, 0, -2, -1, -1, -1, -1, 1],
    [-2, -3, -3, -3, -2, -3, -3, -3, -1, 0, 0, -3, 0, 6, -4, -2, -2, 1, 3, -1],
    [-1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4, 7, -1, -1, -4, -3, -2],
    [1, -1, 1, 0, -1, 0, 0, 0, -1, -2, -2, 0, -1, -2, -1, 4, 1, -3, -2, -2],
    [0, -1, 0, -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 1, 5, -2, -2, 0],
    [-3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1, 1, -4, -3, -2, 11, 2, -3],
    [-2, -2, -2, -3, -2, -1, -2, -3, 2, -1, -1, -2, -1, 3, -3, -2, -2, 2, 7, -1],
    [0, -3, -3, -3, -1, -2, -2, -3, -3, 3, 1, -2, 1, -1, -2, -2, 0, -3, -1, 4]], dtype=np.int32)


def substitution_matrix(biomolecule):
    """int32[q-1, q-1] in the readers' alphabet order, and (gap open, gap extend)
    (sequence_backmapper.py:207-214)."""
    if biomolecule.strip().upper() == "PROTEIN":
        idx = [_BLOSUM_ORDER.index(c) for c in _PROTEIN]
        return _BLOSUM62[np.ix_(idx, idx)].copy(), (-10, -1)
    return np.where(np.eye(4, dtype=bool), 5, -4).astype(np.int32), (-8, 0)


def local_alignment(ref_codes, seq_codes, sub, gap_open, gap_extend):
    """Smith-Waterman / Gotoh with traceback on two code arrays (no gaps inside).  Returns
    (score, pairs) where pairs is the aligned region as a list of (i, j), i / j = index into
    ref / seq or None for a gap there."""
    n, m = len(ref_codes), len(seq_codes)
    NEG = -10 ** 9
    H = np.zeros((n + 1, m + 1), dtype=np.int64)
    E = np.full((n + 1, m + 1), NEG, dtype=np.int64)      # gap in the reference (consumes seq)
    F = np.full((n + 1, m + 1), NEG, dtype=np.int64)      # gap in the sequence (consumes ref)
    for i in range(1, n + 1):
        srow = sub[ref_codes[i - 1]]
        for j in range(1, m + 1):
            E[i, j] = max(H[i, j - 1] + gap_open, E[i, j - 1] + gap_extend)
            F[i, j] = max(H[i - 1, j] + gap_open, F[i - 1, j] + gap_extend)
            H[i, j] = max(0, H[i - 1, j - 1] + srow[seq_codes[j - 1]], E[i, j], F[i, j])
    score = int(H.max())
    if score <= 0:
        return 0, []
    i, j = (int(v) for v in np.argwhere(H == score)[0])   # first best cell (row-major)
    pairs, state = [], "H"
    while i > 0 and j > 0:
        if state == "H":
            if H[i, j] == 0:
                break
            if H[i, j] == H[i - 1, j - 1] + sub[ref_codes[i - 1], seq_codes[j - 1]]:
                pairs.append((i - 1, j - 1))
                i, j = i - 1, j - 1
            elif H[i, j] == F[i, j]:
                state = "F"
            else:
                state = "E"
        elif state == "F":
            pairs.append((i - 1, None))
            state = "H" if F[i, j] == H[i - 1, j] + gap_open else "F"
            i -= 1
        else:
            pairs.append((None, j - 1))
            state = "H" if E[i, j] == H[i, j - 1] + gap_open else "E"
            j -= 1
    pairs.reverse()
    return score, pairs


class SequenceBackmapper:
    """sequence_backmapper.py:25-466."""

    def __init__(self, msa_file=None, alignment_data=None, ref_seq=None, refseq_file=None, biomolecule=None):
        self.__biomolecule = biomolecule.strip().upper()
        if self.__biomolecule not in ("PROTEIN", "RNA"):
            raise ValueError("Unknown biomolecule type {}".format(biomolecule))
        self.__q = 21 if self.__biomolecule == "PROTEIN" else 5
        if msa_file:
            self.__msa0 = fasta_reader.get_alignment_zero_based(msa_file, self.__biomolecule)
        elif alignment_data:                                  # states 1..q, as the reference passes them (:72-79)
            rows, seen = [], set()
            for seq in alignment_data:
                key = tuple(seq)
                if key not in seen:
                    seen.add(key)
                    rows.append(seq)
            self.__msa0 = (np.asarray(rows, dtype=np.int64) - 1).astype(np.uint8)
        else:
            logger.error("\n\tPlease provide alignment file or a list of alignments")
            raise ValueError
        alphabet = (_PROTEIN if self.__q == 21 else _RNA) + "-"
        lut = np.frombuffer(alphabet.encode(), dtype=np.uint8)
        self.__alignment = [row.tobytes().decode() for row in lut[self.__msa0]]
        if refseq_file:
            self.__ref_sequence = self._reference_sequence(refseq_file)
        elif ref_seq:
            self.__ref_sequence = ref_seq.strip().upper()
        else:
            logger.error("\n\tPlease provide a reference sequence or a FASTA file containing a reference sequence")
            raise ValueError
        self._validate_refseq()

    @property
    def alignment(self):
        return self.__alignment

    @property
    def ref_sequence(self):
        return self.__ref_sequence

    def __str__(self):
        return "<A sequence backmapper object of biomolecule type {}>".format(self.__biomolecule)

    def _validate_refseq(self):
        """:133-156: standard residues only."""
        standard = _PROTEIN if self.__q == 21 else _RNA
        for res in self.__ref_sequence:
            if res not in standard:
                logger.error("\n\tReference sequence should only contain standard residues")
                raise ValueError

    def _reference_sequence(self, refseq_file):
        """:159-186: the first record of the file, upper-cased."""
        seqs = fasta_reader.get_alignment_from_fasta_file(refseq_file)
        if len(seqs) > 1:
            logger.warning("\n\tFound multiple reference sequences in file {}.\n\tThe first sequence is taken".format(refseq_file))
        return seqs[0].strip().upper()

    def _codes(self, seq):
        alphabet = _PROTEIN if self.__q == 21 else _RNA
        return np.array([alphabet.index(c) for c in seq], dtype=np.uint8)

    def align_pairs_local(self, ref_seq, other_seq, score_only=False):
        """:188-226.  score_only: the score; else [(ref_aligned, other_aligned, score, begin, end)] with the
        two full sequences laid out like pairwise2 does for a local alignment (unaligned ends padded
        with gaps), begin / end delimiting the aligned region."""
        sub, (go, ge) = substitution_matrix(self.__biomolecule)
        a, b = self._codes(ref_seq), self._codes(other_seq)
        score, pairs = local_alignment(a, b, sub, go, ge)
        if score_only:
            return score
        if not pairs:
            return []
        i0 = next(p[0] for p in pairs if p[0] is not None)
        j0 = next(p[1] for p in pairs if p[1] is not None)
        i1 = max(p[0] for p in pairs if p[0] is not None) + 1
        j1 = max(p[1] for p in pairs if p[1] is not None) + 1
        lead = max(i0, j0)
        ra = "-" * (lead - i0) + ref_seq[:i0]
        oa = "-" * (lead - j0) + other_seq[:j0]
        for i, j in pairs:
            ra += ref_seq[i] if i is not None else "-"
            oa += other_seq[j] if j is not None else "-"
        end = len(ra)
        tail = max(len(ref_seq) - i1, len(other_seq) - j1)
        ra += ref_seq[i1:] + "-" * (tail - (len(ref_seq) - i1))
        oa += other_seq[j1:] + "-" * (tail - (len(other_seq) - j1))
        return [(ra, oa, score, lead, end)]

    def find_matching_seqs_from_alignment(self):
        """:228-283: the sequence(s) of the alignment with the best local alignment score against the
        reference; the N score-only alignments run on the GPU."""
        logger.info("\n\tSearching for sequence(s) that match best with the reference sequence")
        first = self.__alignment[0]
        if first.replace("-", "") == self.__ref_sequence:
            logger.info("\n\tFirst sequence in alignment (gaps removed) matches reference,"
                        "\n\tSkipping regorous search for matching sequence")
            return [first]
        scores = _lib.sw_scores(self.__msa0, self.__q, self._codes(self.__ref_sequence))
        best = int(scores.max())
        matching = [self.__alignment[k] for k in np.flatnonzero(scores == best)]
        if len(matching) > 1:
            logger.warning("\n\tFound {} sequences in MSA that match the reference"
                           "\n\tThe first sequence is taken as matching".format(len(matching)))
        return matching

    @staticmethod
    def align_subsequences(ref_middle_subseq=None, template_subseq_in_msa=None, num_res_middle_template=None):
        """:286-336: thread the template's MSA gaps into the matched part of the reference."""
        mapped, res_count, pos = [], 0, 0
        for site in template_subseq_in_msa:
            if res_count == num_res_middle_template:
                break
            if site != "-":
                mapped.append(ref_middle_subseq[pos])
                pos += 1
                res_count += 1
                if pos == len(ref_middle_subseq):
                    break
            else:
                if ref_middle_subseq[pos] != "-":
                    mapped.append("-")
                else:
                    mapped.append(ref_middle_subseq[pos])
                    pos += 1
        mapped.extend(list(ref_middle_subseq[pos:]))
        return "".join(mapped)

    def map_to_reference_sequence(self):
        """:339-466 -> {MSA column: reference position}, both 0-based."""
        logger.info("\n\tBackmapping reference sequence to MSA")
        template_seq_in_msa = self.find_matching_seqs_from_alignment()[0]
        template_gaps_removed = template_seq_in_msa.replace("-", "")
        aligned = self.align_pairs_local(self.__ref_sequence, template_gaps_removed)
        if not aligned:
            raise ValueError("the reference sequence does not align with any sequence of the alignment")
        ref_seq_aligned, template_seq_aligned, _, start_indx, end_indx = aligned[0]
        ref_middle_subseq = ref_seq_aligned[start_indx:end_indx]
        template_middle_subseq = template_seq_aligned[start_indx:end_indx]
        num_leading_res_template = len(template_seq_aligned[:start_indx].replace("-", ""))
        num_leading_res_ref = len(ref_seq_aligned[:start_indx].replace("-", ""))
        num_res_middle_template = len(template_middle_subseq.replace("-", ""))
        res_count, start_indx_in_msa = 0, 0
        for k, site in enumerate(template_seq_in_msa):
            if res_count == num_leading_res_template:
                start_indx_in_msa = k
                break
            if site != "-":
                res_count += 1
        template_subseq_in_msa = template_seq_in_msa[start_indx_in_msa:]
        backmapped = self.align_subsequences(ref_middle_subseq=ref_middle_subseq,
                                             template_subseq_in_msa=template_subseq_in_msa,
                                             num_res_middle_template=num_res_middle_template)
        mapped_sites, mapped_res_count = dict(), 0
        for k, site in enumerate(backmapped):
            if k == len(template_seq_in_msa) - start_indx_in_msa:
                break
            if site != "-":
                mapped_sites[mapped_res_count + num_leading_res_ref] = start_indx_in_msa + k
                mapped_res_count += 1
        logger.info("\n\tNumber of residues mapped: {}\n\tNumber of residues in the (original) reference sequence: {}"
                    .format(len(mapped_sites), len(self.__ref_sequence)))
        return {value: key for key, value in mapped_sites.items()}
