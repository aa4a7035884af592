/*
 * vq_oracle.c -- CPU restatement of the reference vector-quantisation step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file.  The shipped path is the CUDA library in
 * vq-vae-transformer-arc-welding_b200/csrc and has no CPU fallback.
 *
 * What is restated (reference = tmdt-buw/VQ-VAE-Transformer-Arc-Welding):
 *   forward   model/vector_quantizer.py:88-119
 *   gather    model/vector_quantizer.py:121-131
 *   backward  the autograd of :103-111 (closed form)
 *
 * Parity pin: the reference ships no golden vectors.
 * This restatement is pinned against outputs of the unmodified reference
 * module executed in the authoring container; see oracle/make_golden.py and
 * tests/golden/.  tests/test_oracle_golden.py holds the comparison.
 *
 * Canonical fp32 evaluation order ("oracle order").  The reference evaluates
 *     d[i][k] = (sum(z_i^2) + sum(E_k^2)) - 2 * (z_i . E_k)
 * in fp32 through MKL/cuBLAS, whose summation order is not specified.  The
 * oracle fixes one: every sum is a single ascending chain of fused
 * multiply-adds starting from +0, i.e. acc = fmaf(a[j], b[j], acc) for
 * j = 0..D-1, and the three terms are combined exactly as the reference
 * associates them: fl(fl(zz + ee_k) - fl(2*dot_k)).  The CUDA kernels evaluate
 * the identical chain, so kernel-vs-oracle indices are bit-exact; oracle-vs-
 * reference indices can differ only on fp32 near-ties, which the tests explain
 * row by row with the fp64 distances computed here.
 *
 * argmin semantics follow torch.argmin (model/vector_quantizer.py:96): the
 * lowest index among equal minima, and a NaN distance beats every number
 * (first NaN wins).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VQO_LANES 8

/* ascending fused chain: sum_j a[j]*b[j] */
static inline float chain_dot(const float *a, const float *b, int d)
{
    float acc = 0.0f;
    for (int j = 0; j < d; ++j)
        acc = fmaf(a[j], b[j], acc);
    return acc;
}

/* torch.argmin ordering: NaN is smaller than everything, first one wins. */
static inline int better(float cand, float best)
{
    if (best != best)
        return 0; /* a NaN already holds the slot */
    if (cand != cand)
        return 1;
    return cand < best;
}

/* fp32 oracle-order distances of one row against the whole codebook.
 * Et is the codebook transposed to (D, Kpad) so that VQO_LANES codes advance
 * in lock-step; every lane still runs its own ascending fmaf chain, so the
 * numbers equal the scalar chain bit for bit. */
static void row_distances(const float *zrow, int d, const float *Et, int kpad,
                          const float *ee, int k, float *dist)
{
    const float zz = chain_dot(zrow, zrow, d);
    for (int k0 = 0; k0 < k; k0 += VQO_LANES) {
        float acc[VQO_LANES];
        for (int l = 0; l < VQO_LANES; ++l)
            acc[l] = 0.0f;
        for (int j = 0; j < d; ++j) {
            const float zj = zrow[j];
            const float *e = Et + (size_t)j * kpad + k0;
#pragma omp simd
            for (int l = 0; l < VQO_LANES; ++l)
                acc[l] = fmaf(zj, e[l], acc[l]);
        }
        for (int l = 0; l < VQO_LANES && k0 + l < k; ++l) {
            const float t = zz + ee[k0 + l]; /* model/vector_quantizer.py:91-92 */
            const float u = 2.0f * acc[l];   /* exact */
            dist[k0 + l] = t - u;            /* :92-93 */
        }
    }
}

static float *transpose_codebook(const float *E, int k, int d, int *kpad_out)
{
    const int kpad = (k + VQO_LANES - 1) / VQO_LANES * VQO_LANES;
    float *Et = (float *)calloc((size_t)d * kpad + VQO_LANES, sizeof(float));
    if (!Et)
        return NULL;
    for (int c = 0; c < k; ++c)
        for (int j = 0; j < d; ++j)
            Et[(size_t)j * kpad + c] = E[(size_t)c * d + j];
    *kpad_out = kpad;
    return Et;
}

/*
 * The reference gathers codebook rows with a GEMM, one_hot(n,k) @ E(k,d)
 * (model/vector_quantizer.py:103).  For a finite codebook that is a plain row
 * copy.  With a non-finite entry E[c][j] the GEMM also multiplies it by the
 * zeros of every row that did not pick c, and 0*inf = 0*NaN = NaN, so column j
 * of the gathered matrix is NaN for all those rows.  colflag[j] = 0 when column
 * j is clean, c+1 when exactly code c is non-finite there (rows that picked c
 * keep E[c][j]), -1 when several codes are.
 */
static int *column_poison(const float *E, int k, int d)
{
    int *flag = (int *)calloc((size_t)d, sizeof(int));
    if (!flag)
        return NULL;
    for (int c = 0; c < k; ++c)
        for (int j = 0; j < d; ++j)
            if (!isfinite(E[(size_t)c * d + j]))
                flag[j] = (flag[j] == 0) ? c + 1 : -1;
    return flag;
}

/*
 * Forward.  z: (n, d) contiguous fp32, E: (k, d) contiguous fp32.
 * Outputs (any may be NULL): idx (n) int64, zq (n, d) = z + (E[idx] - z),
 * counts (k) int64, loss_out[0] = m + beta*m with m = mean((E[idx]-z)^2),
 * perplexity_out[0] = exp(-sum p log(p + 1e-10)), p = counts/n.
 * Returns 0, or -1 on bad arguments / allocation failure.
 */
int vq_oracle_forward(const float *z, int64_t n, int d, const float *E, int k,
                      float beta, int64_t *idx, float *zq, int64_t *counts,
                      float *loss_out, float *perplexity_out, int nthreads)
{
    if (n < 0 || d <= 0 || k <= 0 || !E || (n > 0 && !z))
        return -1;
    int kpad = 0;
    float *Et = transpose_codebook(E, k, d, &kpad);
    float *ee = (float *)malloc(sizeof(float) * (size_t)k);
    int64_t *cnt = (int64_t *)calloc((size_t)k, sizeof(int64_t));
    if (!Et || !ee || !cnt) {
        free(Et); free(ee); free(cnt);
        return -1;
    }
    for (int c = 0; c < k; ++c)
        ee[c] = chain_dot(E + (size_t)c * d, E + (size_t)c * d, d);
    int *colflag = column_poison(E, k, d);
    if (!colflag) {
        free(Et); free(ee); free(cnt);
        return -1;
    }

    double sq_total = 0.0;
#ifdef _OPENMP
    if (nthreads > 0)
        omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        float *dist = (float *)malloc(sizeof(float) * (size_t)(kpad + VQO_LANES));
        int64_t *lcnt = (int64_t *)calloc((size_t)k, sizeof(int64_t));
        double lsq = 0.0;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            const float *zr = z + (size_t)i * d;
            row_distances(zr, d, Et, kpad, ee, k, dist);
            int best = 0;
            for (int c = 1; c < k; ++c) /* model/vector_quantizer.py:96 */
                if (better(dist[c], dist[best]))
                    best = c;
            if (idx)
                idx[i] = best;
            lcnt[best] += 1;
            const float *e = E + (size_t)best * d; /* :98-103, one-hot GEMM == row gather */
            double rsq = 0.0;
            for (int j = 0; j < d; ++j) {
                const float ej = (colflag[j] == 0 || colflag[j] == best + 1) ? e[j] : NAN;
                const float diff = ej - zr[j];
                const float sq = diff * diff; /* :107-108, fp32 square */
                rsq += (double)sq;
                if (zq)
                    zq[(size_t)i * d + j] = zr[j] + diff; /* :111 straight-through value */
            }
            lsq += rsq;
        }
#pragma omp critical
        {
            sq_total += lsq;
            for (int c = 0; c < k; ++c)
                cnt[c] += lcnt[c];
        }
        free(dist);
        free(lcnt);
    }

    if (loss_out) {
        const float m = (n > 0) ? (float)(sq_total / ((double)n * (double)d)) : NAN;
        loss_out[0] = m + beta * m; /* :107-108: both means are the same number */
    }
    if (perplexity_out) {
        double h = 0.0;
        for (int c = 0; c < k; ++c) { /* :114-115 */
            const float p = (float)cnt[c] / (float)n;
            h += (double)(p * logf(p + 1e-10f));
        }
        perplexity_out[0] = expf((float)(-h));
    }
    if (counts)
        memcpy(counts, cnt, sizeof(int64_t) * (size_t)k);
    free(Et); free(ee); free(cnt); free(colflag);
    return 0;
}

/*
 * Distances for near-tie analysis.  dist32 (n,k): oracle-order fp32 distances;
 * dist64 (n,k): the same expression in fp64 from the fp32 inputs.  Either may
 * be NULL.
 */
int vq_oracle_distances(const float *z, int64_t n, int d, const float *E, int k,
                        float *dist32, double *dist64)
{
    if (n < 0 || d <= 0 || k <= 0 || !z || !E)
        return -1;
    int kpad = 0;
    float *Et = transpose_codebook(E, k, d, &kpad);
    float *ee = (float *)malloc(sizeof(float) * (size_t)k);
    double *ee64 = (double *)malloc(sizeof(double) * (size_t)k);
    if (!Et || !ee || !ee64) {
        free(Et); free(ee); free(ee64);
        return -1;
    }
    for (int c = 0; c < k; ++c) {
        const float *e = E + (size_t)c * d;
        ee[c] = chain_dot(e, e, d);
        double s = 0.0;
        for (int j = 0; j < d; ++j)
            s += (double)e[j] * (double)e[j];
        ee64[c] = s;
    }
#pragma omp parallel
    {
        float *dist = (float *)malloc(sizeof(float) * (size_t)(kpad + VQO_LANES));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            const float *zr = z + (size_t)i * d;
            if (dist32) {
                row_distances(zr, d, Et, kpad, ee, k, dist);
                memcpy(dist32 + (size_t)i * k, dist, sizeof(float) * (size_t)k);
            }
            if (dist64) {
                double zz = 0.0;
                for (int j = 0; j < d; ++j)
                    zz += (double)zr[j] * (double)zr[j];
                for (int c = 0; c < k; ++c) {
                    const float *e = E + (size_t)c * d;
                    double dot = 0.0;
                    for (int j = 0; j < d; ++j)
                        dot += (double)zr[j] * (double)e[j];
                    dist64[(size_t)i * k + c] = (zz + ee64[c]) - 2.0 * dot;
                }
            }
        }
        free(dist);
    }
    free(Et); free(ee); free(ee64);
    return 0;
}

/* model/vector_quantizer.py:121-131 -- rows of E selected by idx. */
int vq_oracle_gather(const int64_t *idx, int64_t n, const float *E, int k, int d,
                     float *out)
{
    if (n < 0 || d <= 0 || k <= 0 || !E || (n > 0 && (!idx || !out)))
        return -1;
    for (int64_t i = 0; i < n; ++i) {
        if (idx[i] < 0 || idx[i] >= k)
            return -2;
        memcpy(out + (size_t)i * d, E + (size_t)idx[i] * d, sizeof(float) * (size_t)d);
    }
    return 0;
}

/*
 * Backward in closed form, evaluated in fp64:
 *   grad_z[i]  = g_zq[i] + g_loss * 2 * (z_i - E[idx_i]) / M          M = n*d
 *   grad_E[c]  = g_loss * beta * 2 / M * sum_{i: idx_i = c} (E[c] - z_i)
 * g_zq may be NULL (treated as zeros).  g_zq never reaches E: the value path
 * of z_q is detached at model/vector_quantizer.py:111.
 */
int vq_oracle_backward(const float *g_zq, float g_loss, const float *z,
                       const int64_t *idx, const float *E, int64_t n, int d, int k,
                       float beta, float *grad_z, float *grad_E)
{
    if (n <= 0 || d <= 0 || k <= 0 || !z || !idx || !E)
        return -1;
    const double M = (double)n * (double)d;
    const double cz = (double)g_loss * 2.0 / M;
    const double ce = (double)g_loss * (double)beta * 2.0 / M;
    double *acc = (double *)calloc((size_t)k * d, sizeof(double));
    if (!acc)
        return -1;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t c = idx[i];
        if (c < 0 || c >= k) {
            free(acc);
            return -2;
        }
        for (int j = 0; j < d; ++j) {
            const double diff = (double)E[(size_t)c * d + j] - (double)z[(size_t)i * d + j];
            acc[(size_t)c * d + j] += diff;
            if (grad_z) {
                const double g = g_zq ? (double)g_zq[(size_t)i * d + j] : 0.0;
                grad_z[(size_t)i * d + j] = (float)(g - cz * diff);
            }
        }
    }
    if (grad_E)
        for (size_t t = 0; t < (size_t)k * d; ++t)
            grad_E[t] = (float)(ce * acc[t]);
    free(acc);
    return 0;
}

int vq_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
