// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin C ABI over the UNMODIFIED reference class PlmDCA
// (/root/reference/pydca/plmdca/include/plmdca.h:16-84) so that tests and the
// CPU-baseline leg of bench.py can call computeSeqsWeight / initFieldsAndCouplings
// / gradient on caller-supplied buffers.  It is compiled together with the
// reference's own sources, from where they lie, into oracle/_ref/ by
// oracle/Makefile; no reference source is copied into this repository.
//
// The reference keeps its state private; the shim re-opens the class with the
// access keyword redefined (layout is unaffected under the Itanium ABI).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>
#if defined(_OPENMP)
#include <omp.h>
#endif

#define private public
#include "include/plmdca.h"
#undef private

extern "C" {

void *ref_plm_create(const char *msa_file, unsigned biomolecule, unsigned seqs_len,
                     unsigned q, float seqid, float lambda_h, float lambda_J,
                     unsigned num_threads)
{
    try {
        return new PlmDCA(msa_file, biomolecule, seqs_len, q, seqid, lambda_h,
                          lambda_J, num_threads);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_plm_create: %s\n", e.what());
        return nullptr;
    }
}

void ref_plm_destroy(void *p) { delete static_cast<PlmDCA *>(p); }

unsigned ref_plm_num_seqs(void *p) { return static_cast<PlmDCA *>(p)->num_seqs; }

void ref_plm_get_seqs(void *p, unsigned char *out)
{
    auto *o = static_cast<PlmDCA *>(p);
    size_t k = 0;
    for (auto const &s : o->seqs_int_form)
        for (auto v : s) out[k++] = static_cast<unsigned char>(v);
}

void ref_plm_get_weights(void *p, float *out)
{
    auto *o = static_cast<PlmDCA *>(p);
    std::copy(o->seqs_weight.begin(), o->seqs_weight.end(), out);
}

float ref_plm_meff(void *p) { return static_cast<PlmDCA *>(p)->Meff; }

void ref_plm_set_threads(void *p, unsigned n) { static_cast<PlmDCA *>(p)->num_threads = n; }

void ref_plm_init(void *p, float *x) { static_cast<PlmDCA *>(p)->initFieldsAndCouplings(x); }

float ref_plm_gradient(void *p, const float *x, float *g)
{
    return static_cast<PlmDCA *>(p)->gradient(x, g);
}

// re-run the reweighting only (timing leg)
void ref_plm_reweight(void *p, float *out)
{
    auto *o = static_cast<PlmDCA *>(p);
    auto w = o->computeSeqsWeight();
    if (out) std::copy(w.begin(), w.end(), out);
}

// The reference's own lbfgs() (lbfgs/lib/lbfgs.cpp:248) on the reference's own
// PlmDCA::gradient, parameterised exactly as ObjectiveFunction::run does
// (plmdcaBackend.cpp:61-81), but counting the objective evaluations, which
// plmdcaBackend() itself does not report.  x receives the result (P floats).
struct ShimFitCtx { PlmDCA *obj; int evals; };

static lbfgsfloatval_t shim_evaluate(void *instance, const lbfgsfloatval_t *x, lbfgsfloatval_t *g,
                                     const int, const lbfgsfloatval_t)
{
    auto *c = static_cast<ShimFitCtx *>(instance);
    ++c->evals;
    return c->obj->gradient(x, g);
}

int ref_plm_fit(void *p, float *x, unsigned max_iterations, int *evals_out, float *fx_out)
{
    auto *o = static_cast<PlmDCA *>(p);
    const int n = static_cast<int>(o->num_fields_and_couplings);
    lbfgsfloatval_t *m_x = lbfgs_malloc(n);
    if (!m_x) return 1;
    lbfgs_parameter_t param;
    lbfgs_parameter_init(&param);
    param.epsilon = 1E-3;
    param.max_iterations = static_cast<int>(max_iterations);
    param.max_linesearch = 5;
    param.ftol = 1E-4;
    param.m = 5;
    o->initFieldsAndCouplings(m_x);
    ShimFitCtx ctx{o, 0};
    lbfgsfloatval_t fx = 0;
    const int ret = lbfgs(n, m_x, &fx, shim_evaluate, nullptr, &ctx, &param);
    std::copy(m_x, m_x + n, x);
    lbfgs_free(m_x);
    if (evals_out) *evals_out = ctx.evals;
    if (fx_out) *fx_out = fx;
    return ret;
}

// The same with the reference's progress callback hooked (lbfgs.cpp:496): per iteration
// trace[6*k .. 6*k+5] = {fx, xnorm, gnorm, step, line-search evaluations, evaluations so far}
// for the first max_trace iterations; *iters_out = iterations reported.
struct ShimTraceCtx { PlmDCA *obj; int evals; float *trace; int max_trace; int iters; };

static lbfgsfloatval_t shim_evaluate_t(void *instance, const lbfgsfloatval_t *x, lbfgsfloatval_t *g,
                                       const int, const lbfgsfloatval_t)
{
    auto *c = static_cast<ShimTraceCtx *>(instance);
    ++c->evals;
    return c->obj->gradient(x, g);
}

static int shim_progress_t(void *instance, const lbfgsfloatval_t *, const lbfgsfloatval_t *,
                           const lbfgsfloatval_t fx, const lbfgsfloatval_t xnorm, const lbfgsfloatval_t gnorm,
                           const lbfgsfloatval_t step, int, int k, int ls)
{
    auto *c = static_cast<ShimTraceCtx *>(instance);
    if (c->trace && k >= 1 && k <= c->max_trace) {
        float *t = c->trace + 6 * (k - 1);
        t[0] = fx; t[1] = xnorm; t[2] = gnorm; t[3] = step; t[4] = (float)ls; t[5] = (float)c->evals;
    }
    c->iters = k;
    return 0;
}

int ref_plm_fit_trace(void *p, float *x, unsigned max_iterations, int *evals_out, float *fx_out,
                      float *trace, int max_trace, int *iters_out)
{
    auto *o = static_cast<PlmDCA *>(p);
    const int n = static_cast<int>(o->num_fields_and_couplings);
    lbfgsfloatval_t *m_x = lbfgs_malloc(n);
    if (!m_x) return 1;
    lbfgs_parameter_t param;
    lbfgs_parameter_init(&param);
    param.epsilon = 1E-3;
    param.max_iterations = static_cast<int>(max_iterations);
    param.max_linesearch = 5;
    param.ftol = 1E-4;
    param.m = 5;
    o->initFieldsAndCouplings(m_x);
    ShimTraceCtx ctx{o, 0, trace, max_trace, 0};
    lbfgsfloatval_t fx = 0;
    const int ret = lbfgs(n, m_x, &fx, shim_evaluate_t, shim_progress_t, &ctx, &param);
    std::copy(m_x, m_x + n, x);
    lbfgs_free(m_x);
    if (evals_out) *evals_out = ctx.evals;
    if (fx_out) *fx_out = fx;
    if (iters_out) *iters_out = ctx.iters;
    return ret;
}

}  // extern "C"
