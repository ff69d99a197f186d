"""MSATrimmer -- pydca/msa_trimmer/msa_trimmer.py:17-207 without Biopython: which columns of an
alignment to drop, by gap fraction or by the gaps of the sequence that matches a reference."""
import logging

from . import fasta_reader
from .sequence_backmapper import SequenceBackmapper

logger = logging.getLogger(__name__)


class MSATrimmerException(Exception):
    """Raised for invalid trimming requests."""


class MSATrimmer:
    def __init__(self, msa_file, biomolecule=None, max_gap=None, refseq_file=None):
        self.__msa_file = msa_file
        self.__refseq_file = refseq_file
        self.__max_gap = 0.5 if max_gap is None else max_gap
        if self.__max_gap > 1.0 or self.__max_gap < 0.0:
            logger.error("\n\tThe value of max_gap should be between 0 and 1")
            raise MSATrimmerException
        if biomolecule is not None:
            self.__biomolecule = biomolecule.strip().upper()
        else:
            self.__biomolecule = biomolecule
        # records as Biopython's AlignIO would give them: (id, sequence) in file order, all of them
        self.__alignment_data = [(name.split()[0] if name.split() else "", seq) for name, seq in fasta_reader._records(msa_file)
                                 if seq]
        if not self.__alignment_data:
            raise ValueError("no sequences found in %s" % msa_file)

    @property
    def alignment_data(self):
        return self.__alignment_data

    def compute_msa_columns_gap_size(self):
        """:47-84: fraction of '.' / '-' per column."""
        seqs_len = len(self.__alignment_data[0][1])
        num_seqs = len(self.__alignment_data)
        out = []
        for i in range(seqs_len):
            num_gaps = sum(1 for _, seq in self.__alignment_data if seq[i] in ".-")
            out.append(float(num_gaps) / float(num_seqs))
        return tuple(out)

    def msa_columns_beyond_max_gap(self):
        """:87-108."""
        sizes = self.compute_msa_columns_gap_size()
        return tuple(i for i in range(len(sizes)) if sizes[i] > self.__max_gap)

    def trim_by_gap_size(self):
        """:111-127."""
        return tuple(self.msa_columns_beyond_max_gap())

    def trim_by_refseq(self, remove_all_gaps=False):
        """:139-195: columns to remove, judged on the sequence that best matches the reference."""
        seqbackmapper = SequenceBackmapper(msa_file=self.__msa_file, refseq_file=self.__refseq_file,
                                           biomolecule=self.__biomolecule)
        first_matching_seq = seqbackmapper.find_matching_seqs_from_alignment()[0]
        gap_symbols = ["-", "."]
        if not remove_all_gaps:
            candidates = self.msa_columns_beyond_max_gap()
            columns_to_remove = [i for i in candidates if first_matching_seq[i] in gap_symbols]
        else:
            seqs_len = len(self.__alignment_data[0][1])
            columns_to_remove = [i for i in range(seqs_len) if first_matching_seq[i] in gap_symbols]
        return tuple(columns_to_remove)

    def get_msa_trimmed_by_refseq(self, remove_all_gaps=False):
        """:198-207 -> [(id, trimmed sequence)]."""
        drop = set(self.trim_by_refseq(remove_all_gaps=remove_all_gaps))
        return [(sid, "".join(c for i, c in enumerate(seq) if i not in drop)) for sid, seq in self.__alignment_data]

    def get_msa_trimmed_by_gap_size(self):
        drop = set(self.trim_by_gap_size())
        return [(sid, "".join(c for i, c in enumerate(seq) if i not in drop)) for sid, seq in self.__alignment_data]
