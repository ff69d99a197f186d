"""MeanFieldDCA -- the class surface of pydca/meanfield_dca/meanfield_dca.py:43-988
over the B200 backend.

Same constructor (``msa`` is a FASTA path or any iterable of records / strings --
Biopython's MultipleSeqAlignment works but is not required), same properties and
method names.  Each method calls the operator with the reference's name in
``pydca_b200.msa_numerics`` (meanfield_dca.py:371-548 call sites); the scoring
methods use the fused device path: the MSA is uploaded once and frequencies,
pseudocount, covariance assembly, inverse and Frobenius norms stay on the GPU.
"""
import logging

import numpy as np

from . import _lib, fasta_reader, msa_numerics
from .plmdca import _sorted_pairs, average_product_correction

logger = logging.getLogger(__name__)


class MeanFieldDCAException(Exception):
    """Raised for invalid MeanFieldDCA requests."""


class MeanFieldDCA:
    """Mean-field DCA on one B200 (meanfield_dca.py:43)."""

    def __init__(self, msa, biomolecule, pseudocount=None, seqid=None):
        self.__pseudocount = pseudocount if pseudocount is not None else 0.5
        self.__seqid = seqid if seqid is not None else 0.8
        if self.__pseudocount >= 1.0 or self.__pseudocount < 0:
            logger.error("\n\tValue of relative pseudo-count must be between 0 and 1.0. Typical value is 0.5")
            raise ValueError
        if self.__seqid > 1.0 or self.__seqid <= 0.0:
            logger.error("\n\tValue of sequence-identity must not exceed 1 nor less than 0.")
            raise ValueError
        biomolecule = biomolecule.strip().upper()
        self.__msa = msa
        if biomolecule == "RNA":
            self.__num_site_states = 5
        elif biomolecule == "PROTEIN":
            self.__num_site_states = 21
        else:
            logger.error("\n\tUnknown biomolecule ... must be protein (PROTEIN) or rna (RNA)")
            raise ValueError
        if isinstance(msa, str):
            self.__msa0 = fasta_reader.get_alignment_zero_based(msa, biomolecule)
        else:
            try:
                seqs = [str(getattr(rec, "seq", rec)).strip().upper() for rec in msa]
            except TypeError:
                raise ValueError("Alignment input parameter is invalid")
            seqs = [s for s in seqs if s]
            if not seqs:
                raise ValueError("Alignment input parameter is invalid")
            self.__msa0 = fasta_reader.alignment_zero_based(seqs, biomolecule)
        self.__num_sequences, self.__sequences_len = self.__msa0.shape
        self.__biomolecule = biomolecule
        if self.__seqid < 1.0:
            self.__sequences_weight = self.compute_sequences_weight()
        else:       # meanfield_dca.py:113-115
            self.__sequences_weight = np.ones((self.__num_sequences,), dtype=np.float64)
        self.__effective_num_sequences = np.sum(self.__sequences_weight)
        self.__couplings = None
        self.__refseq_mapping_dict = None
        #: copy the n x n couplings matrix to the host when scoring (8 n^2 bytes: 288 MB at L=300).
        #: The scoring methods do not need it; compute_params (the only consumer, meanfield_dca.py:
        #: 661-752) switches it on for its own call
        self.keep_couplings = False
        self.last_stage_ms = None
        logger.info("Created a MeanFieldDCA object: %s, q %d, pseudocount %s, seqid %s, L %d, N %d, Meff %s",
                    biomolecule, self.__num_site_states, self.__pseudocount, self.__seqid,
                    self.__sequences_len, self.__num_sequences, self.__effective_num_sequences)

    def __str__(self):
        return "<instance of MeanFieldDCA>"

    def __call__(self, pseudocount=0.5, seqid=0.8):
        """meanfield_dca.py:160-190 (parameters are reset; weights are not recomputed there either)."""
        self.__pseudocount = pseudocount
        self.__seqid = seqid
        logger.warning("\n\tYou have changed one of the parameters (pseudo count or sequence identity)")
        return None

    # ---- properties (meanfield_dca.py:193-347) ----------------------------------------
    @property
    def alignment(self):
        """Sequences in the reference's integer form (states 1..q)."""
        return (self.__msa0.astype(np.int64) + 1).tolist()

    @property
    def biomolecule(self):
        return self.__biomolecule

    @property
    def sequences_len(self):
        return self.__sequences_len

    @property
    def num_site_states(self):
        return self.__num_site_states

    @property
    def num_sequences(self):
        return self.__num_sequences

    @property
    def sequence_identity(self):
        return self.__seqid

    @property
    def pseudocount(self):
        return self.__pseudocount

    @property
    def sequences_weight(self):
        return self.__sequences_weight

    @property
    def effective_num_sequences(self):
        return np.sum(self.__sequences_weight)

    # ---- operators, one per reference call site ----------------------------------------------
    def _alignment_data(self):
        return self.__msa0.astype(np.int32) + 1

    def compute_sequences_weight(self):
        """meanfield_dca.py:350-375 (float64 compare, msa_numerics.py:46); the zero-based uint8
        alignment goes to the device as it is."""
        logger.info("\n\tComputing sequences weights")
        counts = _lib.sequence_counts(self.__msa0, self.__seqid, compare_f64=True)
        return 1.0 / counts.astype(np.float64)

    def get_single_site_freqs(self):
        logger.info("\n\tComputing single site frequencies")
        return msa_numerics.compute_single_site_freqs(alignment_data=self._alignment_data(),
                                                      num_site_states=self.__num_site_states,
                                                      seqs_weight=self.__sequences_weight)

    def get_reg_single_site_freqs(self):
        return msa_numerics.get_reg_single_site_freqs(single_site_freqs=self.get_single_site_freqs(),
                                                      seqs_len=self.__sequences_len,
                                                      num_site_states=self.__num_site_states,
                                                      pseudocount=self.__pseudocount)

    def get_pair_site_freqs(self):
        logger.info("\n\tComputing pair site frequencies")
        return msa_numerics.compute_pair_site_freqs(alignment_data=self._alignment_data(),
                                                    num_site_states=self.__num_site_states,
                                                    seqs_weight=self.__sequences_weight)

    def get_reg_pair_site_freqs(self):
        return msa_numerics.get_reg_pair_site_freqs(pair_site_freqs=self.get_pair_site_freqs(),
                                                    seqs_len=self.__sequences_len,
                                                    num_site_states=self.__num_site_states,
                                                    pseudocount=self.__pseudocount)

    def construct_corr_mat(self, reg_fi, reg_fij):
        logger.info("\n\tConstructing the correlation matrix")
        return msa_numerics.construct_corr_mat(reg_fi=reg_fi, reg_fij=reg_fij, seqs_len=self.__sequences_len,
                                               num_site_states=self.__num_site_states)

    def compute_couplings(self, corr_mat):
        """meanfield_dca.py:523-553."""
        logger.info("\n\tComputing couplings")
        try:
            couplings = msa_numerics.compute_couplings(corr_mat=corr_mat)
        except Exception as e:
            logger.error("\n\tCorrelation {}\n\tYou set the pseudocount {}. You might need to increase it."
                         .format(e, self.__pseudocount))
            raise
        self.__couplings = couplings
        return couplings

    def compute_two_site_model_fields(self, couplings, reg_fi):
        return msa_numerics.compute_two_site_model_fields(couplings=couplings, reg_fi=reg_fi,
                                                          seqs_len=self.__sequences_len,
                                                          num_site_states=self.__num_site_states)

    # ---- fused device pipeline (what the scoring methods run) -----------------------------------------
    def _reg_fi_and_couplings(self):
        reg_fi, corr = msa_numerics.corr_mat_from_alignment(self.__msa0, self.__num_site_states,
                                                            self.__sequences_weight, self.__pseudocount)
        return reg_fi, self.compute_couplings(corr)

    def compute_fields(self, couplings=None):
        """meanfield_dca.py:588-633 -> {site: float64[q-1]}."""
        if couplings is None:
            reg_fi, couplings = self._reg_fi_and_couplings()
        else:
            reg_fi = self.get_reg_single_site_freqs()
        q, l = self.__num_site_states, self.__sequences_len
        qm1 = q - 1
        p_res = reg_fi[:, :qm1].reshape(l * qm1)
        fields = dict()
        for i in range(l):
            rows = couplings[i * qm1:(i + 1) * qm1]
            total = rows @ p_res - rows[:, i * qm1:(i + 1) * qm1] @ reg_fi[i, :qm1]
            fields[i] = np.log(reg_fi[i, :qm1] / reg_fi[i, -1]) - total
        return fields

    def shift_couplings(self, couplings_ij):
        qm1 = self.__num_site_states - 1
        c = np.reshape(couplings_ij, (qm1, qm1))
        return c - np.mean(c, axis=1).reshape(qm1, 1) - np.mean(c, axis=0).reshape(1, qm1) + np.mean(c)

    @staticmethod
    def _stage_dict(stage_ms):
        """CUDA-event milliseconds of the device pipeline stages of the last scoring call."""
        names = ("pair_freqs_and_h2d", "covariance", "inverse", "frobenius_norm", "direct_info")
        return {k: float(v) for k, v in zip(names, stage_ms)}

    def get_mapped_site_pairs_dca_scores(self, sorted_dca_scores, seqbackmapper):
        """meanfield_dca.py:755-790."""
        mapping_dict = seqbackmapper.map_to_reference_sequence()
        self.__refseq_mapping_dict = mapping_dict
        mapped = [((mapping_dict[p[0]], mapping_dict[p[1]]), s) for p, s in sorted_dca_scores
                  if p[0] in mapping_dict and p[1] in mapping_dict]
        mapped = sorted(mapped, key=lambda k: k[1], reverse=True)
        logger.info("\n\tTotal number of mapped sites: {}".format(len(mapped)))
        return tuple(mapped)

    def get_site_pair_di_score(self):
        """meanfield_dca.py:793-830 -> {(i, j): DI}."""
        logger.info("\n\tComputing direct information")
        di = self._di_scores()
        iu, ju = np.triu_indices(self.__sequences_len, k=1)
        return {(int(i), int(j)): s for i, j, s in zip(iu, ju, di.tolist())}

    def _di_scores(self):
        res = msa_numerics.mean_field_scores(self.__msa0, self.__num_site_states, self.__sequences_weight,
                                             self.__pseudocount, want_couplings=self.keep_couplings, want_di=True)
        self.__couplings = res.get("couplings")
        self.last_stage_ms = self._stage_dict(res["stage_ms"])
        return res["di"]

    def compute_sorted_DI(self, seqbackmapper=None, _materialize=True):
        """meanfield_dca.py:833-868: the (i, j)-ordered dictionary sorted by score, descending, stable."""
        logger.info("\n\tComputing direct information")
        sorted_di = _sorted_pairs(self._di_scores(), self.__sequences_len,
                                  _materialize or seqbackmapper is not None)
        if seqbackmapper is not None:
            sorted_di = self.get_mapped_site_pairs_dca_scores(sorted_di, seqbackmapper)
        return sorted_di

    def compute_sorted_DI_APC(self, seqbackmapper=None):
        sorted_di = self.compute_sorted_DI(_materialize=False)
        logger.info("\n\tPerforming average product correction (APC) of DI scores")
        out = average_product_correction(sorted_di, self.__sequences_len)
        if seqbackmapper is not None:
            out = self.get_mapped_site_pairs_dca_scores(out, seqbackmapper)
        return out

    def compute_sorted_FN(self, seqbackmapper=None, _materialize=True):
        """meanfield_dca.py:902-943."""
        logger.info("\n\tComputing Frobenius norm of couplings")
        res = msa_numerics.mean_field_scores(self.__msa0, self.__num_site_states, self.__sequences_weight,
                                             self.__pseudocount, want_couplings=self.keep_couplings)
        self.__couplings = res.get("couplings")
        self.last_stage_ms = self._stage_dict(res["stage_ms"])
        fn = res["fn"]
        fn_sorted = _sorted_pairs(fn, self.__sequences_len, _materialize or seqbackmapper is not None)
        if seqbackmapper is not None:
            fn_sorted = self.get_mapped_site_pairs_dca_scores(fn_sorted, seqbackmapper)
        return fn_sorted

    def compute_sorted_FN_APC(self, seqbackmapper=None):
        """meanfield_dca.py:946-988."""
        raw_fn = self.compute_sorted_FN(_materialize=False)
        logger.info("\n\tPerforming average product correction (APC) to Frobenius norm of couplings.")
        out = average_product_correction(raw_fn, self.__sequences_len)
        if seqbackmapper is not None:
            out = self.get_mapped_site_pairs_dca_scores(out, seqbackmapper)
        return out

    def compute_params(self, seqbackmapper=None, ranked_by=None, linear_dist=None, num_site_pairs=None):
        """meanfield_dca.py:661-752."""
        ranked_by = "FN_APC" if ranked_by is None else ranked_by.strip().upper()
        linear_dist = 4 if linear_dist is None else linear_dist
        methods = {"FN": self.compute_sorted_FN, "FN_APC": self.compute_sorted_FN_APC,
                   "DI": self.compute_sorted_DI, "DI_APC": self.compute_sorted_DI_APC}
        if ranked_by not in methods:
            logger.error("\n\tInvalid ranking criterion {}.\nChoose from {}".format(ranked_by, tuple(methods)))
            raise MeanFieldDCAException
        keep, self.keep_couplings = self.keep_couplings, True
        try:
            dca_scores = methods[ranked_by](seqbackmapper=seqbackmapper)
        finally:
            self.keep_couplings = keep
        fields = self.compute_fields(couplings=self.__couplings)
        qm1 = self.__num_site_states - 1
        if seqbackmapper is not None:
            mapping = {v: k for k, v in self.__refseq_mapping_dict.items()}
        else:
            mapping = {i: i for i in range(self.__sequences_len)}
        if num_site_pairs is None:
            num_site_pairs = len(seqbackmapper.ref_sequence) if seqbackmapper is not None else len(mapping)
        fields_mapped = [(i, fields[mapping[i]]) for i in mapping]
        ranked, count = [], 0
        for pair, _ in dca_scores:
            if abs(pair[0] - pair[1]) > linear_dist:
                count += 1
                if count > num_site_pairs:
                    break
                i, j = mapping[pair[0]], mapping[pair[1]]
                if i > j:
                    logger.error("\n\tInvalid site pair. Site pair (i, j) should be ordered in i < j")
                    raise MeanFieldDCAException
                block = self.__couplings[i * qm1:(i + 1) * qm1, j * qm1:(j + 1) * qm1]
                ranked.append((pair, self.shift_couplings(block).reshape(qm1 * qm1)))
        if count < num_site_pairs:
            logger.warning("\n\tObtained couplings for only {} ranked site pairs.".format(count))
        return tuple(fields_mapped), tuple(ranked)
