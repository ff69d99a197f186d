/*
 * dca_oracle.c -- CPU restatement of pydca's plmDCA numerics.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pydca_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function
 * here against the unmodified reference compiled into oracle/_ref/ (bit-exact
 * at num_threads=1) and against tests/golden/*.npz generated from it.
 *
 * Each function cites the reference lines it restates (paths relative to
 * /root/reference/).  Arithmetic is float32 in exactly the reference's
 * operation order so that the single-thread reference is reproduced bit for
 * bit (no FMA contraction on baseline x86-64, no -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------------- */
/* residue encoding: pydca/plmdca/plmdca_numerics.cpp:699-732 (states      */
/* 0..q-1, gap = q-1).  Characters the C++ map lacks would throw there     */
/* (:752); we follow pydca/fasta_reader/fasta_reader.py:142-151 and map     */
/* them to the gap state (the Python reader's behaviour, a superset).       */
/* ---------------------------------------------------------------------- */
int orc_encode_char(int biomolecule, int c)
{
    static const char prot[] = "ACDEFGHIKLMNPQRSTVWY";
    static const char rna[] = "ACGU";
    const char *alpha = (biomolecule == 1) ? prot : rna;
    int q = (biomolecule == 1) ? 21 : 5;
    int u = toupper(c);
    const char *p = (u != 0) ? strchr(alpha, u) : NULL;
    return p ? (int)(p - alpha) : q - 1;
}

/*
 * FASTA -> integer MSA with first-occurrence de-duplication.
 * plmdca_numerics.cpp:685-767 reads the first L characters of every
 * non-empty line that does not start with '>' (:748-752) and keeps a
 * sequence only if it was not seen before (:756).  For unwrapped FASTA this
 * is identical to fasta_reader.py:79-163; for wrapped records we join the
 * record's lines first (Biopython semantics) and then take L characters.
 * Returns the number of unique sequences, or -1 on error.  *out is malloc'd.
 */
long orc_read_fasta(const char *path, int biomolecule, int L, uint8_t **out)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    size_t cap = 1024, n = 0;
    uint8_t *msa = (uint8_t *)malloc(cap * (size_t)L);
    char *rec = (char *)malloc((size_t)L + 1);
    int fill = -1; /* -1: no record open */
    char *line = NULL;
    size_t lcap = 0;
    ssize_t len;
    int eof = 0;
    while (!eof) {
        len = getline(&line, &lcap, f);
        if (len < 0) { eof = 1; }
        int header = eof || (len > 0 && line[0] == '>');
        if (header) {
            if (fill > 0) {
                if (fill < L) { free(msa); free(rec); free(line); fclose(f); return -2; }
                /* encode + dedup */
                if (n == cap) { cap *= 2; msa = (uint8_t *)realloc(msa, cap * (size_t)L); }
                uint8_t *row = msa + n * (size_t)L;
                for (int i = 0; i < L; ++i) row[i] = (uint8_t)orc_encode_char(biomolecule, rec[i]);
                int dup = 0;
                for (size_t k = 0; k < n && !dup; ++k)
                    dup = memcmp(msa + k * (size_t)L, row, (size_t)L) == 0;
                if (!dup) ++n;
            }
            fill = 0;
            continue;
        }
        if (fill < 0) fill = 0; /* headerless file: treat lines as one record each */
        for (ssize_t k = 0; k < len; ++k) {
            char c = line[k];
            if (c == '\n' || c == '\r' || c == ' ' || c == '\t') continue;
            if (fill < L) rec[fill] = c;
            ++fill;
        }
    }
    free(rec); free(line); fclose(f);
    *out = msa;
    return (long)n;
}

void orc_free(void *p) { free(p); }

/* ---------------------------------------------------------------------- */
/* Sequence reweighting: neighbour counts.                                  */
/* plmdca_numerics.cpp:611-671 (float32 compare, self included, strict >)   */
/* meanfield_dca/msa_numerics.py:14-50 (float64 compare).                   */
/* f64 != 0 selects the numba variant.                                      */
/* ---------------------------------------------------------------------- */
void orc_reweight_counts(const uint8_t *msa, long N, int L, double seqid,
                         int f64, uint32_t *counts)
{
    const float seqid_f = (float)seqid;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) {
        const uint8_t *si = msa + i * (long)L;
        uint32_t c = 0;
        for (long j = 0; j < N; ++j) {
            const uint8_t *sj = msa + j * (long)L;
            unsigned ident = 0;
            for (int s = 0; s < L; ++s) ident += (si[s] == sj[s]);
            if (f64) {
                if ((double)ident / (double)L > seqid) ++c;
            } else {
                float sim = (float)ident / (float)L;
                if (sim > seqid_f) ++c;
            }
        }
        counts[i] = c;
    }
}

/* the same loop for a caller-chosen subset of rows i (full-size spot checks of deep
 * alignments: work is linear in the number of rows): counts[k] = c_{rows[k]}          */
void orc_reweight_counts_rows(const uint8_t *msa, long N, int L, double seqid, int f64,
                              const long *rows, long nrows, uint32_t *counts)
{
    const float seqid_f = (float)seqid;
#pragma omp parallel for schedule(static)
    for (long k = 0; k < nrows; ++k) {
        const uint8_t *si = msa + rows[k] * (long)L;
        uint32_t c = 0;
        for (long j = 0; j < N; ++j) {
            const uint8_t *sj = msa + j * (long)L;
            unsigned ident = 0;
            for (int s = 0; s < L; ++s) ident += (si[s] == sj[s]);
            if (f64) {
                if ((double)ident / (double)L > seqid) ++c;
            } else {
                float sim = (float)ident / (float)L;
                if (sim > seqid_f) ++c;
            }
        }
        counts[k] = c;
    }
}

/* plmdca_numerics.cpp:669  w = 1.f / count (count held as float there). */
void orc_plm_weights(const uint32_t *counts, long N, float *w)
{
    for (long i = 0; i < N; ++i) w[i] = 1.f / (float)counts[i];
}

/* plmdca_numerics.cpp:46  Meff = std::accumulate(w, 0.f) (sequential f32). */
float orc_plm_meff(const float *w, long N)
{
    float m = 0.f;
    for (long i = 0; i < N; ++i) m += w[i];
    return m;
}

/* ---------------------------------------------------------------------- */
/* Initial point: plmdca_numerics.cpp:51-81 (single-site freqs) + :207-249. */
/* ---------------------------------------------------------------------- */
void orc_plm_init(const uint8_t *msa, const float *w, long N, int L, int q,
                  float *x)
{
    const float meff = orc_plm_meff(w, N);
    const long nf = (long)L * q;
    const long P = nf + (long)L * (L - 1) / 2 * q * q;
    float *fi = (float *)malloc(sizeof(float) * (size_t)q);
    for (int i = 0; i < L; ++i) {
        for (int a = 0; a < q; ++a) fi[a] = 0.f;
        for (long n = 0; n < N; ++n) fi[msa[n * (long)L + i]] += w[n];
        for (int a = 0; a < q; ++a) {
            float fr = fi[a] / meff;                 /* :77 */
            x[a + q * i] = logf(fr * meff + 1.f);    /* :227 */
        }
    }
    for (int i = 0; i < L; ++i) {                    /* :231-241 */
        float s = 0.f;
        for (int a = 0; a < q; ++a) s += x[a + q * i];
        float av = s / (float)q;
        for (int a = 0; a < q; ++a) x[a + q * i] -= av;
    }
    for (long k = nf; k < P; ++k) x[k] = 0.f;        /* :244-247 */
    free(fi);
}

static inline long pair_index(long L, long i, long j)
{   /* plmdca_numerics.cpp:340 */
    return (L * (L - 1) / 2) - (L - i) * ((L - i) - 1) / 2 + j - i - 1;
}

/* ---------------------------------------------------------------------- */
/* Objective + gradient: plmdca_numerics.cpp:436-607.                       */
/* carry != 0 reproduces the reference exactly: prob_ni is allocated once   */
/* per site (:492) and only += at :499, so sequence n's logits start from   */
/* sequence n-1's probabilities.  carry == 0 is textbook plmDCA.            */
/* Sites are merged in ascending order (== the reference at num_threads=1). */
/* site_lo/site_hi restrict the data term to a site range (multi-GPU tests  */
/* and bounded CPU-baseline samples); pass 0, L for the full objective.     */
/* ---------------------------------------------------------------------- */
float orc_plm_gradient_range(const uint8_t *msa, const float *w, long N, int L,
                             int q, float lh, float lJ, const float *x,
                             float *g, int carry, int site_lo, int site_hi)
{
    const long nf = (long)L * q;
    const long qq = (long)q * q;
    float fx = 0.f;

    for (int i = 0; i < L; ++i)                       /* :463-471 */
        for (int a = 0; a < q; ++a) {
            const float h = x[a + i * q];
            g[a + q * i] = 2.f * lh * h;
            fx += lh * h * h;
        }
    for (int i = 0; i < L - 1; ++i)                   /* :473-486 */
        for (int j = i + 1; j < L; ++j) {
            const long k = nf + pair_index(L, i, j) * qq;
            for (int a = 0; a < q; ++a)
                for (int b = 0; b < q; ++b) {
                    const float J = x[k + b + q * a];
                    g[k + b + q * a] = 2.f * lJ * J;
                    fx += lJ * J * J;
                }
        }

    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    /* per-site partial results are merged in site order => deterministic and
       identical to the reference at num_threads=1 regardless of our threads */
    float *fxs = (float *)calloc((size_t)L, sizeof(float));
    const int chunk = nthreads;
    float *fg_all = (float *)malloc(sizeof(float) * (size_t)chunk * q);
    float *cg_all = (float *)malloc(sizeof(float) * (size_t)chunk * L * qq);

    for (int base = site_lo; base < site_hi; base += chunk) {
        const int cnt = (site_hi - base < chunk) ? site_hi - base : chunk;
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < cnt; ++t) {
            const int i = base + t;
            float *fg = fg_all + (size_t)t * q;
            float *cg = cg_all + (size_t)t * L * qq;
            float prob[64];
            for (int a = 0; a < q; ++a) { prob[a] = 0.f; fg[a] = 0.f; }
            memset(cg, 0, sizeof(float) * (size_t)L * qq);
            float fxi = 0.f;
            for (long n = 0; n < N; ++n) {            /* :496 */
                const uint8_t *s = msa + n * (long)L;
                if (!carry) for (int a = 0; a < q; ++a) prob[a] = 0.f;
                for (int a = 0; a < q; ++a) prob[a] += x[a + i * q];   /* :499 */
                for (int j = 0; j < i; ++j) {          /* :501-508 */
                    const long k = nf + pair_index(L, j, i) * qq;
                    for (int a = 0; a < q; ++a) prob[a] += x[k + a + s[j] * q];
                }
                for (int j = i + 1; j < L; ++j) {      /* :510-517 */
                    const long k = nf + pair_index(L, i, j) * qq;
                    for (int a = 0; a < q; ++a) prob[a] += x[k + s[j] + a * q];
                }
                float mx = prob[0];                    /* :519-524 */
                for (int a = 0; a < q; ++a) if (prob[a] > mx) mx = prob[a];
                for (int a = 0; a < q; ++a) prob[a] = expf(prob[a] - mx);
                float z = 0.f;
                for (int a = 0; a < q; ++a) z += prob[a];
                z = 1.f / z;
                for (int a = 0; a < q; ++a) prob[a] *= z;

                const float wn = w[n];
                const int r = s[i];
                fxi -= wn * logf(prob[r]);             /* :536 */
                fg[r] -= wn;                           /* :538-539 */
                for (int a = 0; a < q; ++a) fg[a] += wn * prob[a];
                for (int j = 0; j < i; ++j) cg[r + q * (s[j] + q * j)] -= wn;      /* :541-545 */
                for (int j = i + 1; j < L; ++j) cg[s[j] + q * (r + q * j)] -= wn;  /* :546-550 */
                for (int j = 0; j < i; ++j) {          /* :552-559 */
                    const long k = q * (s[j] + q * j);
                    for (int a = 0; a < q; ++a) cg[a + k] += wn * prob[a];
                }
                for (int j = i + 1; j < L; ++j)        /* :560-567 */
                    for (int a = 0; a < q; ++a) cg[s[j] + q * (a + q * j)] += wn * prob[a];
            }
            fxs[i] = fxi;
        }
        for (int t = 0; t < cnt; ++t) {               /* merge, :570-602 */
            const int i = base + t;
            const float *fg = fg_all + (size_t)t * q;
            const float *cg = cg_all + (size_t)t * L * qq;
            fx += fxs[i];
            for (int a = 0; a < q; ++a) g[a + q * i] += fg[a];
            for (int j = 0; j < i; ++j) {
                const long k = nf + pair_index(L, j, i) * qq;
                for (long ab = 0; ab < qq; ++ab) g[k + ab] += cg[ab + qq * j];
            }
            for (int j = i + 1; j < L; ++j) {
                const long k = nf + pair_index(L, i, j) * qq;
                for (long ab = 0; ab < qq; ++ab) g[k + ab] += cg[ab + qq * j];
            }
        }
    }
    free(fxs); free(fg_all); free(cg_all);
    return fx;
}

float orc_plm_gradient(const uint8_t *msa, const float *w, long N, int L, int q,
                       float lh, float lJ, const float *x, float *g, int carry)
{
    return orc_plm_gradient_range(msa, w, N, L, q, lh, lJ, x, g, carry, 0, L);
}

/* Same objective/gradient in float64 with the carry, used to put a
   precision floor under the float32 comparisons. */
double orc_plm_gradient_f64(const uint8_t *msa, const float *w, long N, int L,
                            int q, double lh, double lJ, const float *x,
                            double *g, int carry)
{
    const long nf = (long)L * q;
    const long qq = (long)q * q;
    const long P = nf + (long)L * (L - 1) / 2 * qq;
    double fx = 0.0;
    for (long k = 0; k < P; ++k) {
        const double lam = (k < nf) ? lh : lJ;
        g[k] = 2.0 * lam * x[k];
        fx += lam * (double)x[k] * x[k];
    }
    double *fxs = (double *)calloc((size_t)L, sizeof(double));
    /* rows[(i,a)] accumulated per site privately, then merged in order */
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < L; ++i) {
        double prob[64], fg[64];
        double *cg = (double *)calloc((size_t)L * qq, sizeof(double));
        for (int a = 0; a < q; ++a) { prob[a] = 0.0; fg[a] = 0.0; }
        double fxi = 0.0;
        for (long n = 0; n < N; ++n) {
            const uint8_t *s = msa + n * (long)L;
            if (!carry) for (int a = 0; a < q; ++a) prob[a] = 0.0;
            for (int a = 0; a < q; ++a) prob[a] += x[a + i * q];
            for (int j = 0; j < i; ++j) {
                const long k = nf + pair_index(L, j, i) * qq;
                for (int a = 0; a < q; ++a) prob[a] += x[k + a + s[j] * q];
            }
            for (int j = i + 1; j < L; ++j) {
                const long k = nf + pair_index(L, i, j) * qq;
                for (int a = 0; a < q; ++a) prob[a] += x[k + s[j] + a * q];
            }
            double mx = prob[0];
            for (int a = 0; a < q; ++a) if (prob[a] > mx) mx = prob[a];
            double z = 0.0;
            for (int a = 0; a < q; ++a) { prob[a] = exp(prob[a] - mx); z += prob[a]; }
            for (int a = 0; a < q; ++a) prob[a] /= z;
            const double wn = w[n];
            const int r = s[i];
            fxi -= wn * log(prob[r]);
            fg[r] -= wn;
            for (int a = 0; a < q; ++a) fg[a] += wn * prob[a];
            for (int j = 0; j < L; ++j) {
                if (j == i) continue;
                if (j < i) {
                    cg[r + q * (s[j] + q * j)] -= wn;
                    for (int a = 0; a < q; ++a) cg[a + q * (s[j] + q * j)] += wn * prob[a];
                } else {
                    cg[s[j] + q * (r + q * j)] -= wn;
                    for (int a = 0; a < q; ++a) cg[s[j] + q * (a + q * j)] += wn * prob[a];
                }
            }
        }
        fxs[i] = fxi;
#pragma omp critical
        {
            for (int a = 0; a < q; ++a) g[a + q * i] += fg[a];
            for (int j = 0; j < L; ++j) {
                if (j == i) continue;
                const long k = nf + ((j < i) ? pair_index(L, j, i) : pair_index(L, i, j)) * qq;
                for (long ab = 0; ab < qq; ++ab) g[k + ab] += cg[ab + qq * j];
            }
        }
        free(cg);
    }
    for (int i = 0; i < L; ++i) fx += fxs[i];
    free(fxs);
    return fx;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
