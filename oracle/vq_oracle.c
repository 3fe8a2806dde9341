/*
 * vq_oracle.c -- plain-C CPU restatement of the VectorQuantizer hot path.
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg as the checker.  Never part of the product path.
 *
 * Follows /root/reference/src/acoustic_locating_vq_vae/vq_vae/vector_quantizer.py:29-58 and the
 * gradients autograd derives from it.  Where the reference leaves the
 * floating-point summation order to its BLAS (the N x K GEMM at :36, the reductions at :34-35),
 * this file FIXES one order, and the CUDA "exact" kernels implement the same order, so that code
 * indices can be compared bit for bit:
 *
 *   norm2(x)  = fma chain over d = 0..D-1, acc = fmaf(x[d], x[d], acc), acc0 = +0      (:34,:35)
 *   dot(z,e)  = fma chain over d = 0..D-1, acc = fmaf(z[d], e[d], acc), acc0 = +0      (:36 matmul)
 *   dist(n,k) = fmaf(-2, dot, fl(a_n + b_k))  == fl(fl(a_n + b_k) - 2*dot)             (:34-36 order)
 *   idx(n)    = first k with the minimal dist (strict '<' scan in ascending k)          (:38)
 *
 * Parity pinning: tests/test_oracle.py checks this file against the golden fixtures generated
 * from the real reference class (tests/golden/make_golden.py): indices identical except rows
 * that are fp32 near-ties in fp64 arithmetic, loss/perplexity/quantized/dz/dE within 1e-5 rel.
 *
 * Build: see oracle/Makefile (gcc -O2 -fPIC -shared; -ffp-contract=off so that only the explicit
 * fmaf calls fuse).  Optional OpenMP over rows (the arithmetic per row is order-fixed, so the
 * thread count does not change any result except the documented double-precision sums).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define VQ_ORACLE_ABI 1

int vq_oracle_abi_version(void) { return VQ_ORACLE_ABI; }

int vq_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static float norm2_chain(const float* x, int D) {
    float acc = 0.0f;
    for (int d = 0; d < D; ++d) acc = fmaf(x[d], x[d], acc);
    return acc;
}

static float dot_chain(const float* z, const float* e, int D) {
    float acc = 0.0f;
    for (int d = 0; d < D; ++d) acc = fmaf(z[d], e[d], acc);
    return acc;
}

/* vector_quantizer.py:35 -- |E_k|^2 for every codeword. */
void vq_oracle_code_norms(const float* E, int K, int D, float* e_norm2) {
    for (int k = 0; k < K; ++k) e_norm2[k] = norm2_chain(E + (size_t)k * D, D);
}

/* vector_quantizer.py:32-38 -- flatten (rows = consecutive D-float chunks), distances, argmin.
 * dist_out (N*K floats) may be NULL.  Blocked over codes so the codebook block stays in cache;
 * the per-(n,k) arithmetic is independent of the blocking. */
void vq_oracle_argmin(const float* z, const float* E, int64_t N, int K, int D,
                      int32_t* idx, float* dist_out) {
    float* b = (float*)malloc(sizeof(float) * (size_t)K);
    vq_oracle_code_norms(E, K, D, b);
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
        const float* zr = z + (size_t)n * D;
        const float a = norm2_chain(zr, D);
        float best = INFINITY;
        int32_t bi = 0;
        for (int k = 0; k < K; ++k) {
            const float c = dot_chain(zr, E + (size_t)k * D, D);
            const float t = a + b[k];
            const float d = fmaf(-2.0f, c, t);
            if (dist_out) dist_out[(size_t)n * K + k] = d;
            if (d < best) { best = d; bi = k; }
        }
        idx[n] = bi;
    }
    free(b);
}

/* vector_quantizer.py:39-56 given the indices: gather, straight-through value, losses, usage
 * histogram, perplexity.  onehot (N*K) may be NULL.
 *   q_out = fl(z + fl(E[idx] - z))                       (:43,:54)
 *   m     = mean((E[idx]-z)^2)  accumulated in double, rounded to fp32 once   (:46-50)
 *   loss  = fl(m + fl(beta*m))                            (:52)
 *   p_k   = fl(count_k / N);  perplexity = exp(-sum p_k*log(p_k + 1e-10))     (:55-56)
 */
void vq_oracle_quantize(const float* z, const float* E, const int32_t* idx, int64_t N, int K, int D,
                        float beta, float* q_out, float* onehot, float* hist,
                        float* loss, float* perplexity, double* sse_out) {
    double sse = 0.0;
    for (int k = 0; k < K; ++k) hist[k] = 0.0f;
    if (onehot) memset(onehot, 0, sizeof(float) * (size_t)N * (size_t)K);
    for (int64_t n = 0; n < N; ++n) {
        const float* zr = z + (size_t)n * D;
        const float* er = E + (size_t)idx[n] * D;
        for (int d = 0; d < D; ++d) {
            const float diff = er[d] - zr[d];
            q_out[(size_t)n * D + d] = zr[d] + diff;
            sse += (double)diff * (double)diff;
        }
        hist[idx[n]] += 1.0f;
        if (onehot) onehot[(size_t)n * K + idx[n]] = 1.0f;
    }
    const float m = (float)(sse / ((double)N * (double)D));
    *loss = m + beta * m;
    double ent = 0.0;
    for (int k = 0; k < K; ++k) {
        const float p = hist[k] / (float)N;
        ent += (double)(p * logf(p + 1e-10f));
    }
    *perplexity = expf((float)(-ent));
    if (sse_out) *sse_out = sse;
}

/* Autograd of vector_quantizer.py:46-54.
 *   dz[n,:]      = g_q[n,:] + g_loss*beta*2*(z - q)/(n_rows_dz*D)      g_q may be NULL (= 0)
 *   dE[idx[n],:]+= g_loss*2*(q - z)/(n_rows_dE*D)                      dE may be NULL (frozen)
 * The straight-through output sends no gradient to the codebook.  Rows are accumulated into dE
 * in ascending n (a fixed order; the GPU default path uses atomics and differs in the last bits). */
void vq_oracle_backward(const float* g_q, float g_loss, const float* z, const float* E,
                        const int32_t* idx, int64_t N, int64_t n_rows_dz, int64_t n_rows_dE,
                        int K, int D, float beta, float* dz, float* dE) {
    const float cz = g_loss * beta * 2.0f / (float)((double)n_rows_dz * (double)D);
    const float ce = g_loss * 2.0f / (float)((double)n_rows_dE * (double)D);
    if (dE) memset(dE, 0, sizeof(float) * (size_t)K * (size_t)D);
    for (int64_t n = 0; n < N; ++n) {
        const float* zr = z + (size_t)n * D;
        const float* er = E + (size_t)idx[n] * D;
        for (int d = 0; d < D; ++d) {
            const float diff = er[d] - zr[d];
            const float g = g_q ? g_q[(size_t)n * D + d] : 0.0f;
            dz[(size_t)n * D + d] = fmaf(-cz, diff, g);
            if (dE) dE[(size_t)idx[n] * D + d] += ce * diff;
        }
    }
}

/* One whole step (forward + backward) on the CPU: the "port" CPU baseline bench.py times. */
void vq_oracle_step(const float* z, const float* E, const float* g_q, float g_loss,
                    int64_t N, int K, int D, float beta, int train_vq,
                    int32_t* idx, float* q_out, float* hist, float* loss, float* perplexity,
                    float* dz, float* dE) {
    vq_oracle_argmin(z, E, N, K, D, idx, NULL);
    vq_oracle_quantize(z, E, idx, N, K, D, beta, q_out, NULL, hist, loss, perplexity, NULL);
    vq_oracle_backward(g_q, g_loss, z, E, idx, N, N, N, K, D, beta, dz, train_vq ? dE : NULL);
}
