#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

* C++ side (plmdca_numerics.cpp) through oracle/_ref/libpydca_ref.so (ref_shim.cpp)
* numba side by importing pydca/meanfield_dca/msa_numerics.py and
  pydca/plmdca/msa_numerics.py BY FILE PATH (they need numba + numpy only)
* FN / APC by executing the loops of meanfield_dca.py:902-988 and
  plmdca.py:437-524 re-typed here on the reference's arrays (Biopython, which
  the class constructors need, is not installed).

The .npz files carry inputs (integer MSAs) and reference outputs, so the tests
never touch /root/reference.
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

import oracle  # noqa: E402
from pydca_b200.synthetic import synthetic_msa, write_fasta  # noqa: E402


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


mf_num = load_by_path("ref_mf_numerics", os.path.join(REF, "pydca/meanfield_dca/msa_numerics.py"))
plm_num = load_by_path("ref_plm_numerics", os.path.join(REF, "pydca/plmdca/msa_numerics.py"))


def ref_mf(msa0, q, seqid, pseudocount, with_di):
    """Reference mean-field pipeline on a 0-based uint8 MSA (shifted to 1..q)."""
    msa1 = (msa0.astype(np.int32) + 1)
    L = msa1.shape[1]
    w = mf_num.compute_sequences_weight(alignment_data=msa1, seqid=seqid)
    fi = mf_num.compute_single_site_freqs(alignment_data=msa1, num_site_states=q, seqs_weight=w)
    fi_raw = fi.copy()
    rfi = mf_num.get_reg_single_site_freqs(single_site_freqs=fi, seqs_len=L, num_site_states=q,
                                           pseudocount=pseudocount)
    fij = mf_num.compute_pair_site_freqs(alignment_data=msa1, num_site_states=q, seqs_weight=w)
    fij_raw = fij.copy()
    rfij = mf_num.get_reg_pair_site_freqs(pair_site_freqs=fij, seqs_len=L, num_site_states=q,
                                          pseudocount=pseudocount)
    C = mf_num.construct_corr_mat(reg_fi=rfi, reg_fij=rfij, seqs_len=L, num_site_states=q)
    J = mf_num.compute_couplings(corr_mat=C)
    # meanfield_dca.py:926-939
    fn = []
    for i in range(L):
        rs, re = i * (q - 1), (i + 1) * (q - 1)
        for j in range(i + 1, L):
            cs, ce = j * (q - 1), (j + 1) * (q - 1)
            cij = J[rs:re, cs:ce]
            m1 = np.reshape(np.mean(cij, axis=0), (1, q - 1))
            m2 = np.reshape(np.mean(cij, axis=1), (q - 1, 1))
            new = cij - m1 - m2 + np.mean(cij)
            fn.append(((i, j), np.sqrt(np.sum(new * new))))
    raw = sorted(fn, key=lambda x: x[1], reverse=True)
    fn_apc_sorted = ref_apc(raw, L)
    out = dict(w=w, fi=fi_raw, reg_fi=rfi, fij=fij_raw, corr=C, couplings=J,
               fn=np.array([s for _, s in fn]),
               fn_sorted_pairs=np.array([p for p, _ in raw], dtype=np.int32),
               fn_apc_sorted_pairs=np.array([p for p, _ in fn_apc_sorted], dtype=np.int32),
               fn_apc_sorted_scores=np.array([s for _, s in fn_apc_sorted]))
    if with_di:
        fields = mf_num.compute_two_site_model_fields(couplings=J, reg_fi=rfi, seqs_len=L, num_site_states=q)
        out["two_site_fields"] = fields
        out["di"] = mf_num.compute_direct_info(couplings=J, fields_ij=fields, reg_fi=rfi, seqs_len=L,
                                               num_site_states=q)
    return out


def ref_apc(sorted_scores, N):
    """meanfield_dca.py:969-984 == plmdca.py:503-521, executed verbatim in spirit."""
    av = []
    for i in range(N):
        i_scores = [score for pair, score in sorted_scores if i in pair]
        assert len(i_scores) == N - 1
        av.append(sum(i_scores) / float(N - 1))
    av_all = sum(av) / float(N)
    out = [(pair, score - av[pair[0]] * (av[pair[1]] / av_all)) for pair, score in sorted_scores]
    return sorted(out, key=lambda k: k[1], reverse=True)


def ref_plm_scores(x, L, q):
    """plmdca.py:246-268 + :437-524 on a parameter vector x (float32)."""
    qm1 = q - 1

    def idx(i, j, a, b):
        site = int(((L * (L - 1) / 2) - (L - i) * ((L - i) - 1) / 2 + j - i - 1) * q * q)
        return L * q + site + b + a * q

    couplings = []
    for i in range(L - 1):
        for j in range(i + 1, L):
            for a in range(qm1):
                for b in range(qm1):
                    couplings.append(x[idx(i, j, a, b)])
    couplings = np.array(couplings)
    fn = []
    for i in range(L - 1):
        for j in range(i + 1, L):
            s = int(((L * (L - 1) / 2) - (L - i) * ((L - i) - 1) / 2 + j - i - 1) * qm1 * qm1)
            cij = np.reshape(couplings[s:s + qm1 * qm1], (qm1, qm1))
            avx = np.reshape(np.mean(cij, axis=1), (qm1, 1))
            avy = np.reshape(np.mean(cij, axis=0), (1, qm1))
            cij = cij - avx - avy + np.mean(cij)
            fn.append(((i, j), np.sqrt(np.sum(cij * cij))))
    raw = sorted(fn, key=lambda k: k[1], reverse=True)
    # the reference's pinned numpy (<=1.15) promotes float32 scalar sums with
    # Python floats to float64; reproduce that promotion explicitly
    raw64 = [(p, float(s)) for p, s in raw]
    apc_sorted = ref_apc(raw64, L)
    return dict(couplings_nogap=couplings, fn=np.array([s for _, s in fn], dtype=np.float32),
                fn_sorted_pairs=np.array([p for p, _ in raw], dtype=np.int32),
                fn_apc_sorted_pairs=np.array([p for p, _ in apc_sorted], dtype=np.int32),
                fn_apc_sorted_scores=np.array([s for _, s in apc_sorted]))


def ref_plm_di(x, msa0, L, q, seqid):
    """plmdca.py:683-720 on a parameter vector (weights recomputed in float64
    by plmdca/msa_numerics.py, pseudocount 0.5 hard-coded at plmdca.py:647)."""
    msa1 = msa0.astype(np.int64) + 1
    sc = ref_plm_scores(x, L, q)
    w = plm_num.compute_sequences_weight(alignment_data=msa1, sequence_identity=seqid)
    fi = plm_num.compute_single_site_freqs(alignment_data=msa1, num_site_states=q, seqs_weight=w)
    rfi = plm_num.get_reg_single_site_freqs(single_site_freqs=fi, seqs_len=L, num_site_states=q,
                                            pseudocount=0.5)
    c64 = sc["couplings_nogap"].astype(np.float64)
    fields = plm_num.compute_two_site_model_fields(couplings=c64, reg_fi=rfi, seqs_len=L, num_site_states=q)
    di = plm_num.compute_direct_info(couplings=c64, fields_ij=fields, reg_fi=rfi, seqs_len=L,
                                     num_site_states=q)
    return dict(reg_fi=rfi, two_site_fields=fields, di=di)


def plm_case(tag, msa, q, seqid, lh, lJ, seed, tmp, iters=0):
    N, L = msa.shape
    bio = "protein" if q == 21 else "rna"
    path = write_fasta(msa, os.path.join(tmp, tag + ".fa"), q)
    R = oracle.ref()
    rp = R.plm(path, bio, L, q, seqid, lh, lJ, 1)
    assert (rp.seqs() == msa).all()
    w = rp.weights()
    x0 = rp.init()
    rng = np.random.default_rng(seed)
    x1 = (x0 + rng.normal(0.0, 0.1, x0.size)).astype(np.float32)
    f0, g0 = rp.gradient(x0)
    f1, g1 = rp.gradient(x1)
    out = dict(msa=msa, q=q, seqid=seqid, lambda_h=lh, lambda_J=lJ, w=w, meff=rp.meff(),
               x0=x0, x1=x1, fx0=f0, g0=g0, fx1=f1, g1=g1)
    sc = ref_plm_scores(x1, L, q)
    out.update({"x1_" + k: v for k, v in sc.items() if k != "couplings_nogap"})
    di = ref_plm_di(x1 * 0.05, msa, L, q, seqid)
    out.update({"x1s_" + k: v for k, v in di.items()})
    if iters:
        xf = R.plmdca_backend(path, bio, L, q, seqid, lh, lJ, iters, 1, False)
        out["x_fit_nt1"] = xf
        out["fit_iters"] = iters
    rp.close()
    return out


def main():
    P = oracle.port()
    tmp = tempfile.mkdtemp()

    # ---- K4/K5: RF00167 untrimmed (BASELINE config 1) ---------------------
    rf = os.path.join(REF, "examples/MSA_RF00167.fa")
    msa = P.read_fasta(rf, "rna", 102)
    R = oracle.ref()
    rp = R.plm(rf, "rna", 102, 5, 0.8, 1.0, 1.0, 8)
    assert (rp.seqs() == msa).all()
    w32 = rp.weights()
    counts = np.rint(1.0 / w32.astype(np.float64)).astype(np.uint32)
    rp.close()
    mf = ref_mf(msa, 5, 0.8, 0.5, with_di=True)
    np.savez_compressed(os.path.join(HERE, "rf00167_full.npz"), msa=msa, counts=counts, w_f32=w32,
                        mf_w=mf["w"], mf_fi=mf["fi"], mf_fn=mf["fn"], mf_di=mf["di"],
                        mf_fn_apc_sorted_pairs=mf["fn_apc_sorted_pairs"],
                        mf_fn_apc_sorted_scores=mf["fn_apc_sorted_scores"])
    print("rf00167_full: N", msa.shape, "sum c", counts.sum(), "Meff", mf["w"].sum(),
          "top5", mf["fn_apc_sorted_pairs"][:5].tolist(), mf["fn_apc_sorted_scores"][:5])

    # ---- K1/K2: RF00167 trimmed to the non-gap columns of record 18 ----------
    raw = []
    with open(rf) as f:
        for line in f:
            line = line.strip()
            if line and not line.startswith(">"):
                raw.append(line)
    refrow = raw[18]
    cols = [k for k, c in enumerate(refrow) if c not in "-.~"]
    assert len(cols) == 71
    trimmed_path = os.path.join(tmp, "rf00167_trimmed.fa")
    with open(trimmed_path, "w") as f:
        for n, s in enumerate(raw):
            f.write(">s%d\n%s\n" % (n, "".join(s[k] for k in cols)))
    msat = P.read_fasta(trimmed_path, "rna", 71)
    mft = ref_mf(msat, 5, 0.8, 0.5, with_di=False)
    xfit = R.plmdca_backend(trimmed_path, "rna", 71, 5, 0.8, 1.0, 20.0, 500, 1, False)
    sc = ref_plm_scores(xfit, 71, 5)
    np.savez_compressed(os.path.join(HERE, "rf00167_trimmed.npz"), msa=msat, mf_w=mft["w"],
                        mf_fn=mft["fn"], mf_fn_apc_sorted_pairs=mft["fn_apc_sorted_pairs"],
                        mf_fn_apc_sorted_scores=mft["fn_apc_sorted_scores"],
                        plm_x_fit_nt1=xfit, plm_fn=sc["fn"],
                        plm_fn_apc_sorted_pairs=sc["fn_apc_sorted_pairs"],
                        plm_fn_apc_sorted_scores=sc["fn_apc_sorted_scores"])
    print("rf00167_trimmed: N", msat.shape, "Meff", mft["w"].sum(),
          "mf top5", mft["fn_apc_sorted_pairs"][:5].tolist(), mft["fn_apc_sorted_scores"][:5])
    print("   plm top5", sc["fn_apc_sorted_pairs"][:5].tolist(), sc["fn_apc_sorted_scores"][:5])

    # ---- synthetic plm cases (reference C++ gradient at two points) ----------
    for tag, N, L, q, lh, lJ, iters in [("toy_rna", 60, 7, 5, 1.0, 2.0, 30),
                                        ("toy_protein", 300, 20, 21, 3.8, 3.8, 20),
                                        ("mid_protein", 1000, 40, 21, 7.8, 7.8, 0)]:
        m = synthetic_msa(N, L, q, seed=len(tag) * 7 + N)
        out = plm_case(tag, m, q, 0.8, lh, lJ, seed=N + L, tmp=tmp, iters=iters)
        np.savez_compressed(os.path.join(HERE, "plm_%s.npz" % tag), **out)
        print("plm_%s: fx0 %.6f fx1 %.6f meff %.4f" % (tag, out["fx0"], out["fx1"], out["meff"]))

    # ---- synthetic mf cases (reference numba pipeline incl. DI) --------------
    for tag, N, L, q in [("toy_rna", 80, 9, 5), ("toy_protein", 400, 16, 21)]:
        m = synthetic_msa(N, L, q, seed=N * 3 + L)
        out = ref_mf(m, q, 0.8, 0.5, with_di=True)
        out["msa"] = m
        out["q"] = q
        np.savez_compressed(os.path.join(HERE, "mf_%s.npz" % tag), **out)
        print("mf_%s: meff %.6f fn[0] %.12f di[0] %.12f" % (tag, out["w"].sum(), out["fn"][0], out["di"][0]))


if __name__ == "__main__":
    main()
