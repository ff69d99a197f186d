"""PlmDCA -- the class surface of pydca/plmdca/plmdca.py:25-790 over the B200 backend.

Constructor keywords, defaults, validation, properties and method names are the
reference's; the numerics go through the same C entry point the reference binds
(``plmdcaBackend`` / ``freeFieldsAndCouplings``, plmdca.py:79-89, :216-230), now
exported by ``_plmdcaBackend_b200.so``.  Post-processing that the reference does
in Python loops over P parameters (:225-227, :262-267, :461-477, :508) is array
code here; Frobenius norms, APC and DI run on the device.
"""
import ctypes
import gc
import logging

import numpy as np

from . import _lib, fasta_reader, msa_numerics

logger = logging.getLogger(__name__)


class PlmDCAException(Exception):
    """Implements exceptions related to PlmDCA computation"""


_PAIR_TUPLES = {}
_PAIR_OBJECTS = {}


def _pair_tuples(seqs_len):
    """[(i, j)] for i < j in pair order, built once per L."""
    pt = _PAIR_TUPLES.get(seqs_len)
    if pt is None:
        iu, ju = np.triu_indices(seqs_len, k=1)
        pt = _PAIR_TUPLES[seqs_len] = list(zip(iu.tolist(), ju.tolist()))
    return pt


class _ScoreList(list):
    """A ranked [((i, j), score), ...] list that also carries its content as arrays (pair index in
    pair order, score), so that the APC step does not have to parse 10^5 tuples back."""
    __slots__ = ("pair_idx", "values", "seqs_len", "arrays_only")


def _pair_objects(seqs_len):
    """The (i, j) tuples of _pair_tuples as an object array: fancy indexing then builds the ranked
    list of pairs at C speed."""
    pa = _PAIR_OBJECTS.get(seqs_len)
    if pa is None:
        pt = _pair_tuples(seqs_len)
        pa = np.empty(len(pt), dtype=object)
        pa[:] = pt
        _PAIR_OBJECTS[seqs_len] = pa
    return pa


def _ranked(pair_idx, values, seqs_len, materialize=True):
    """materialize=False: arrays only (an intermediate result handed straight to the APC step; building
    4.5e4 tuples costs as much as the whole device pipeline at L = 300)."""
    if materialize:
        # 1e5 new tuples in a row trigger the cyclic collector dozens of times (60 % of the time of
        # this function); none of them can be garbage, so it is paused for the construction
        was_enabled = gc.isenabled()
        gc.disable()
        try:
            out = _ScoreList(zip(_pair_objects(seqs_len)[pair_idx].tolist(), values.tolist()))
        finally:
            if was_enabled:
                gc.enable()
    else:
        out = _ScoreList()
    out.pair_idx, out.values, out.seqs_len, out.arrays_only = pair_idx, values, seqs_len, not materialize
    return out


def _stable_descending_order(scores):
    """argsort(-scores, kind="stable") -- what sorted(..., reverse=True) does to a list in (i, j)
    order -- as an unstable sort (5x faster) followed by an index sort inside each run of ties."""
    order = np.argsort(-scores)
    vals = scores[order]
    tie = vals[1:] == vals[:-1]
    if tie.any():
        run_start = np.flatnonzero(np.concatenate(([True], ~tie)))
        run_id = np.repeat(np.arange(run_start.size), np.diff(np.append(run_start, vals.size)))
        order = order[np.lexsort((order, run_id))]
    return order


def _sorted_pairs(scores, seqs_len, materialize=True):
    """[((i, j), score), ...] sorted like ``sorted(..., reverse=True)``: descending,
    ties keep (i, j) order (plmdca.py:478)."""
    scores = np.asarray(scores, dtype=np.float64)
    order = _stable_descending_order(scores)
    return _ranked(order, scores[order], seqs_len, materialize)


def average_product_correction(sorted_scores, seqs_len):
    """plmdca.py:503-521 / meanfield_dca.py:969-984 on a [((i,j),score)] list.  The reference sums,
    for every site, the scores of the pairs containing it in the order of the sorted list; the
    unbuffered np.add.at below adds in exactly that order, with plain float64 additions as
    sum() did on the Pythons the reference ran on (3.12's sum() is compensated: <= 1 ulp apart)."""
    n = seqs_len
    iu, ju = np.triu_indices(n, k=1)
    if isinstance(sorted_scores, _ScoreList) and sorted_scores.seqs_len == n and (
            sorted_scores.arrays_only or len(sorted_scores) == len(sorted_scores.values)):
        idx, vals = sorted_scores.pair_idx, sorted_scores.values
    else:                                   # a foreign list (e.g. back-mapped scores): parse it
        first = np.full((n, n), -1, dtype=np.int64)
        first[iu, ju] = np.arange(iu.size)
        pairs = np.array([p for p, _ in sorted_scores], dtype=np.int64).reshape(-1, 2)
        idx = first[pairs[:, 0], pairs[:, 1]]
        vals = np.array([s for _, s in sorted_scores], dtype=np.float64)
        if (idx < 0).any():
            raise ValueError("average_product_correction expects site pairs (i, j) with i < j < %d" % n)
    pi, pj = iu[idx], ju[idx]
    site_sum = np.zeros(n, dtype=np.float64)
    np.add.at(site_sum, np.stack([pi, pj], axis=1).ravel(), np.repeat(vals, 2))
    av = site_sum / float(n - 1)
    tot = 0.0
    for v in av.tolist():          # plain left-to-right float sum (the reference's Python 3.5-3.7 sum())
        tot += v
    av_all = tot / float(n)
    corrected = vals - av[pi] * (av[pj] / av_all)
    order = _stable_descending_order(corrected)
    return _ranked(idx[order], corrected[order], n)


class PlmDCA:
    """Pseudolikelihood-maximisation DCA on one B200 (plmdca.py:25)."""

    def __init__(self, msa_file, biomolecule, seqid=None, lambda_h=None, lambda_J=None,
                 max_iterations=None, num_threads=None, verbose=False, arithmetic=None, estimator=None):
        # arithmetic (not in the reference): 'accurate' (float64 reductions, fastest), 'reference' (the
        # reference's sequential float32 sums bit for bit -> its L-BFGS trajectory, FN within ~1e-6;
        # one GPU) or None = the DCA_B200_ARITHMETIC environment variable (default accurate)
        if arithmetic not in _lib.ARITHMETIC_MODES:
            logger.error("\n\tInvalid arithmetic mode {}".format(arithmetic))
            raise PlmDCAException
        self.__arithmetic = arithmetic
        # estimator (not in the reference): 'joint' = the reference's fit of one symmetric J (default);
        # 'asymmetric' = every site fits its own couplings independently (sites spread over the GPUs with
        # no exchange), symmetrised at the end -- a different estimator, close scores, not the reference's;
        # None = the DCA_B200_ESTIMATOR environment variable
        if estimator not in _lib.ESTIMATORS:
            logger.error("\n\tInvalid estimator {}".format(estimator))
            raise PlmDCAException
        self.__estimator = estimator
        self.__biomolecule = biomolecule.strip().upper()
        if self.__biomolecule not in ("PROTEIN", "RNA"):
            logger.error("\n\tInvalid biomolecule type {}".format(self.__biomolecule))
            raise PlmDCAException
        self.__msa_file = msa_file
        self.__biomolecule_int = 1 if self.__biomolecule == "PROTEIN" else 2
        self.__num_site_states = 21 if self.__biomolecule == "PROTEIN" else 5
        self.__num_seqs, self.__seqs_len = self._get_num_and_len_of_seqs()
        self.__seqid = 0.8 if seqid is None else seqid
        if self.__seqid <= 0 or self.__seqid > 1.0:
            logger.error("\n\t{} is an invalid value of sequences identity (seqid) parameter".format(self.__seqid))
            raise PlmDCAException
        self.__lambda_h = 0.2 * (self.__seqs_len - 1) if lambda_h is None else lambda_h
        if self.__lambda_h < 0:
            logger.error("\n\tlambda_h must be a positive number. You passed lambda_h={}".format(self.__lambda_h))
            raise PlmDCAException
        self.__lambda_J = 0.2 * (self.__seqs_len - 1) if lambda_J is None else lambda_J
        if self.__lambda_J < 0:
            logger.error("\n\tlambda_J must be a positive number. You passed lambda_J={}".format(self.__lambda_J))
            raise PlmDCAException
        self.__max_iterations = max_iterations if max_iterations is not None else 100
        self.__num_threads = 1 if num_threads is None else num_threads
        self.__verbose = True if verbose else False
        q, l = self.__num_site_states, self.__seqs_len
        self.__data_size = l * q + l * (l - 1) // 2 * q * q
        self.__fields_and_couplings_all = None
        self.__refseq_mapping_dict = None
        self.__last_fit = None
        logger.info("Created plmDCA instance: biomolecule %s, L %d, sequences %d, seqid %s, lambda_h %s, "
                    "lambda_J %s, iterations %d", self.__biomolecule, l, self.__num_seqs, self.__seqid,
                    self.__lambda_h, self.__lambda_J, self.__max_iterations)

    # ---- properties (plmdca.py:107-160) -------------------------------------------
    @property
    def biomolecule(self):
        return self.__biomolecule

    @property
    def sequence_identity(self):
        return self.__seqid

    @property
    def lambda_h(self):
        return self.__lambda_h

    @property
    def lambda_J(self):
        return self.__lambda_J

    @property
    def max_iterations(self):
        return self.__max_iterations

    @property
    def sequences_len(self):
        return self.__seqs_len

    @property
    def num_sequences(self):
        return self.__num_seqs

    @property
    def effective_num_sequences(self):
        raise NotImplementedError

    @property
    def last_fit(self):
        """L-BFGS status / iteration / evaluation counts of the last backend call
        (the reference discards them, plmdcaBackend.cpp:197)."""
        return self.__last_fit

    def _get_num_and_len_of_seqs(self):
        """(records in the file, length of record 0), plmdca.py:163-180."""
        return fasta_reader.first_record_length(self.__msa_file)

    def map_index_couplings(self, i, j, a, b):
        """plmdca.py:183-199."""
        q, l = self.__num_site_states, self.__seqs_len
        site = ((l * (l - 1)) // 2 - ((l - i) * (l - i - 1)) // 2 + j - i - 1) * q * q
        return l * q + site + b + a * q

    # ---- backend call (plmdca.py:202-243) -------------------------------------------
    def get_fields_and_couplings_from_backend(self):
        """float32[P]: fields then couplings, through the reference's C ABI."""
        logger.info("\n\tComputing fields and couplings using gradient decent")
        lib = _lib.lib()
        _lib.set_backend_arithmetic(self.__arithmetic)
        _lib.set_estimator(self.__estimator)
        h_j_ptr = lib.plmdcaBackend(
            self.__biomolecule_int, self.__num_site_states, self.__msa_file.encode("utf-8"),
            self.__seqs_len, self.__seqid, self.__lambda_h, self.__lambda_J, self.__max_iterations,
            self.__num_threads, self.__verbose)
        if not h_j_ptr:
            logger.error("\n\tplmDCA backend failed: %s", _lib.last_error())
            raise PlmDCAException(_lib.last_error())
        view = np.ctypeslib.as_array(ctypes.cast(h_j_ptr, ctypes.POINTER(ctypes.c_float)),
                                     shape=(self.__data_size,))
        fields_and_couplings = np.array(view, dtype=np.float32, copy=True)
        lib.freeFieldsAndCouplings(h_j_ptr)
        self.__last_fit = _lib.last_fit()
        return fields_and_couplings

    # ---- gap-state stripping (plmdca.py:246-318) -----------------------------------------
    def get_couplings_no_gap_state(self, fields_and_couplings_all):
        q, l = self.__num_site_states, self.__seqs_len
        x = np.asarray(fields_and_couplings_all)
        blocks = x[l * q:].reshape(l * (l - 1) // 2, q, q)[:, :q - 1, :q - 1]
        return np.ascontiguousarray(blocks).reshape(-1)

    def get_fields_no_gap_state(self, fields_and_couplings_all):
        q, l = self.__num_site_states, self.__seqs_len
        x = np.asarray(fields_and_couplings_all)
        return x[:l * q].reshape(l, q)[:, :q - 1].reshape(-1).tolist()

    def get_fields_and_couplings_no_gap_state(self, fields_and_couplings_all):
        return (self.get_fields_no_gap_state(fields_and_couplings_all),
                self.get_couplings_no_gap_state(fields_and_couplings_all))

    def shift_couplings(self, couplings_ij):
        """plmdca.py:320-342."""
        qm1 = self.__num_site_states - 1
        c = np.reshape(couplings_ij, (qm1, qm1))
        return c - np.mean(c, axis=1).reshape(qm1, 1) - np.mean(c, axis=0).reshape(1, qm1) + np.mean(c)

    # ---- scores ---------------------------------------------------------------------------
    def _fit_if_needed(self, refit=True):
        if refit or self.__fields_and_couplings_all is None:
            self.__fields_and_couplings_all = self.get_fields_and_couplings_from_backend()
        return self.__fields_and_couplings_all

    def compute_sorted_FN(self, seqbackmapper=None, _materialize=True):
        """plmdca.py:437-481."""
        x = self._fit_if_needed()
        logger.info("\n\tComputing non-APC sorted DCA score")
        fn, _ = _lib.plm_fn_apc(x, self.__seqs_len, self.__num_site_states)
        scores = _sorted_pairs(fn, self.__seqs_len, _materialize or seqbackmapper is not None)
        if seqbackmapper is not None:
            scores = self.get_mapped_site_pairs_dca_scores(scores, seqbackmapper)
        return scores

    def compute_sorted_FN_APC(self, seqbackmapper=None):
        """plmdca.py:484-524."""
        scores_plmdca = self.compute_sorted_FN(_materialize=False)
        logger.info("\n\tPerforming average product correction (APC) of FN  of DCA scores")
        sorted_fn_apc = average_product_correction(scores_plmdca, self.__seqs_len)
        if seqbackmapper is not None:
            sorted_fn_apc = self.get_mapped_site_pairs_dca_scores(sorted_fn_apc, seqbackmapper)
        return sorted_fn_apc

    def get_mapped_site_pairs_dca_scores(self, sorted_dca_scores, seqbackmapper):
        """plmdca.py:527-562 (seqbackmapper: any object with map_to_reference_sequence())."""
        mapping_dict = seqbackmapper.map_to_reference_sequence()
        self.__refseq_mapping_dict = mapping_dict
        mapped = [((mapping_dict[p[0]], mapping_dict[p[1]]), s) for p, s in sorted_dca_scores
                  if p[0] in mapping_dict and p[1] in mapping_dict]
        mapped = sorted(mapped, key=lambda k: k[1], reverse=True)
        logger.info("\n\tTotal number of mapped sites: {}".format(len(mapped)))
        return tuple(mapped)

    # ---- DI path (plmdca.py:565-790) ------------------------------------------------------
    def _alignment(self):
        return fasta_reader.get_alignment_zero_based(self.__msa_file, self.__biomolecule)

    def compute_seqs_weight(self):
        """float64 weights recomputed as the reference does for DI (plmdca.py:565-591)."""
        counts = _lib.sequence_counts(self._alignment(), self.__seqid, compare_f64=True)
        seqs_weight = 1.0 / counts.astype(np.float64)
        self.__seqs_weight = seqs_weight
        self.__eff_num_seqs = np.sum(seqs_weight)
        return seqs_weight

    def get_single_site_freqs(self):
        msa = self._alignment()
        w = 1.0 / _lib.sequence_counts(msa, self.__seqid, compare_f64=True).astype(np.float64)
        return msa_numerics.compute_single_site_freqs(alignment_data=msa.astype(np.int64) + 1,
                                                      num_site_states=self.__num_site_states, seqs_weight=w)

    def get_reg_single_site_freqs(self):
        """pseudocount 0.5 is hard-coded by the reference for plmDCA DI (plmdca.py:647)."""
        fi = self.get_single_site_freqs()
        return msa_numerics.get_reg_single_site_freqs(single_site_freqs=fi, seqs_len=self.__seqs_len,
                                                      num_site_states=self.__num_site_states, pseudocount=0.5)

    def compute_two_site_model_fields(self, couplings=None):
        reg_fi = self.get_reg_single_site_freqs()
        x = self._fit_if_needed(refit=False)
        return msa_numerics.plm_direct_info(x, reg_fi, self.__seqs_len, self.__num_site_states,
                                            want_fields=True)[1]

    def compute_direct_info_unsorted_DI(self):
        """plmdca.py:683-720."""
        x = self._fit_if_needed()
        reg_fi = self.get_reg_single_site_freqs()
        logger.info("\n\tComputing direct information")
        return msa_numerics.plm_direct_info(x, reg_fi, self.__seqs_len, self.__num_site_states)

    def compute_sorted_DI(self, seqbackmapper=None, _materialize=True):
        di = self.compute_direct_info_unsorted_DI()
        sorted_di = _sorted_pairs(di, self.__seqs_len, _materialize or seqbackmapper is not None)
        if seqbackmapper is not None:
            sorted_di = self.get_mapped_site_pairs_dca_scores(sorted_di, seqbackmapper)
        return sorted_di

    def compute_sorted_DI_APC(self, seqbackmapper=None):
        sorted_di = self.compute_sorted_DI(_materialize=False)
        logger.info("\n\tPerforming average product correction (APC) of DI scores")
        out = average_product_correction(sorted_di, self.__seqs_len)
        if seqbackmapper is not None:
            out = self.get_mapped_site_pairs_dca_scores(out, seqbackmapper)
        return out

    # ---- compute_params (plmdca.py:345-434) ----------------------------------------------------
    def compute_params(self, seqbackmapper=None, ranked_by=None, linear_dist=None, num_site_pairs=None):
        ranked_by = "FN_APC" if ranked_by is None else ranked_by.strip().upper()
        linear_dist = 4 if linear_dist is None else linear_dist
        methods = {"FN": self.compute_sorted_FN, "FN_APC": self.compute_sorted_FN_APC,
                   "DI": self.compute_sorted_DI, "DI_APC": self.compute_sorted_DI_APC}
        if ranked_by not in methods:
            logger.error("\n\tInvalid ranking criterion {}.\nChoose from {}".format(ranked_by, tuple(methods)))
            raise PlmDCAException
        dca_scores = methods[ranked_by](seqbackmapper=seqbackmapper)
        x = self.__fields_and_couplings_all
        q, l = self.__num_site_states, self.__seqs_len
        qm1 = q - 1
        fields = np.asarray(self.get_fields_no_gap_state(x))
        couplings = self.get_couplings_no_gap_state(x).reshape(-1, qm1 * qm1)
        if seqbackmapper is not None:
            mapping = {v: k for k, v in self.__refseq_mapping_dict.items()}
        else:
            mapping = {i: i for i in range(l)}
        if num_site_pairs is None:
            num_site_pairs = len(seqbackmapper.ref_sequence) if seqbackmapper is not None else len(mapping)
        fields_mapped = [(i, fields[qm1 * mapping[i]: qm1 * mapping[i] + qm1]) for i in mapping]
        ranked, count = [], 0
        for pair, _ in dca_scores:
            if abs(pair[0] - pair[1]) > linear_dist:
                count += 1
                if count > num_site_pairs:
                    break
                i, j = mapping[pair[0]], mapping[pair[1]]
                if i > j:
                    logger.error("\n\tInvalid site pair. Site pair (i, j) should be ordered in i < j")
                    raise PlmDCAException
                k = (l * (l - 1)) // 2 - ((l - i) * (l - i - 1)) // 2 + j - i - 1
                ranked.append((pair, self.shift_couplings(couplings[k]).reshape(qm1 * qm1)))
        if count < num_site_pairs:
            logger.warning("\n\tObtained couplings for only {} ranked site pairs.".format(count))
        return tuple(fields_mapped), tuple(ranked)
