// kernels_simt.cuh -- CUDA-core kernels of the VQ hot path:
//   prep_codebook_kernel   |E_k|^2 and the tf32 hi/lo split of E        (vector_quantizer.py:35)
//   argmin_simt_kernel     exact fp32 distances + first-index argmin    (vector_quantizer.py:34-38)
//   quantize_rows_kernel   gather, straight-through value, SSE, usage histogram, one-hot,
//                          loss / perplexity                             (vector_quantizer.py:39-56)
//   finalize_stats_kernel  loss / perplexity from all-reduced statistics (data parallel)
//   onehot_kernel          dense one-hot from indices                    (vector_quantizer.py:39-40)
//   backward_kernel        dz and the scatter-add into dE, one red per element (autograd of :46-54; see kernels_bwd.cuh)
// Arithmetic order is the one oracle/vq_oracle.c documents.
#pragma once
#include "common.cuh"

namespace b200vq {

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ p, float v, long long n) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        p[i] = v;
}

// zeroes the usage histogram and the last-CTA counter ahead of the fused forward kernel
__global__ void __launch_bounds__(256) zero_state_kernel(float* __restrict__ hist, int K, unsigned int* __restrict__ counter) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) hist[k] = 0.0f;
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0u;
}

// ---------------------------------------------------------------------------------------------
// prep: one warp per codeword
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_codebook_kernel(const float* __restrict__ E, int K, int D,
                                                            float* __restrict__ e_norm2, float* __restrict__ E_hi,
                                                            float* __restrict__ E_lo, float* __restrict__ hist_zero,
                                                            unsigned int* __restrict__ counter_zero,
                                                            float* __restrict__ dE_zero) {
    pdl_launch_dependents();     // the forward kernel may set itself up while this one runs
    pdl_wait_prior_grids();      // ... but the buffers reset below may still be in use by the previous step
    {   // optional step-state reset riding along: usage histogram, last-CTA counter, dE accumulator
        const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
        if (hist_zero != nullptr)
            for (int k = tid; k < K; k += nth) hist_zero[k] = 0.0f;
        if (counter_zero != nullptr && tid == 0) *counter_zero = 0u;
        if (dE_zero != nullptr)
            for (int i = tid; i < K * D; i += nth) dE_zero[i] = 0.0f;
    }
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= K) return;
    const float* row = E + static_cast<size_t>(warp) * D;
    // |E_k|^2 as ONE sequential FMA chain over d (the order oracle/vq_oracle.c:norm2_chain fixes): the row is
    // loaded coalesced, each element is broadcast by shuffle, and every lane carries the same chain.
    float acc = 0.0f;
    for (int d0 = 0; d0 < D; d0 += 32) {
        const int d = d0 + lane;
        const float v = d < D ? row[d] : 0.0f;
        if (E_hi != nullptr && d < D) {
            const float hi = tf32_rna(v);
            E_hi[static_cast<size_t>(warp) * D + d] = hi;
            if (E_lo != nullptr) E_lo[static_cast<size_t>(warp) * D + d] = tf32_rna(v - hi);
        }
        const int n = min(32, D - d0);
        for (int l = 0; l < n; ++l) {
            const float x = __shfl_sync(0xffffffffu, v, l);
            acc = fmaf(x, x, acc);
        }
    }
    if (lane == 0) e_norm2[warp] = acc;
}

// ---------------------------------------------------------------------------------------------
// exact argmin: 64 rows x 64 codes per CTA tile, 4 x 4 per thread, D streamed in chunks of 32.
// Every (row, code) dot product is ONE sequential fmaf chain over d = 0..D-1 (zero padding past D
// adds exact zeros), so results are bit-identical to oracle/vq_oracle.c for any tiling.
// grid = (ceil(N/64), splits); with splits > 1 partial minima meet in `keys` via 64-bit atomicMin.
// ---------------------------------------------------------------------------------------------
constexpr int S_TM = 64, S_TN = 64, S_TD = 32, S_LD = 68;

__global__ void __launch_bounds__(256) argmin_simt_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                          const float* __restrict__ e_norm2, long long N, int K, int D,
                                                          int codes_per_split, int* __restrict__ idx_out,
                                                          unsigned long long* __restrict__ keys,
                                                          float* __restrict__ hist_to_zero,
                                                          unsigned int* __restrict__ counter_to_zero) {
    __shared__ __align__(16) float zs[S_TD][S_LD];
    __shared__ __align__(16) float es[S_TD][S_LD];
    __shared__ float a_s[S_TM];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long row0 = static_cast<long long>(blockIdx.x) * S_TM;
    const int k_begin = blockIdx.y * codes_per_split;
    const int k_end = min(K, k_begin + codes_per_split);

    if (blockIdx.x == 0 && blockIdx.y == 0) {  // forward-state reset rides along (consumed by the next kernel)
        for (int k = tid; k < K; k += blockDim.x) hist_to_zero[k] = 0.0f;
        if (tid == 0) *counter_to_zero = 0u;
    }
    if (tid < S_TM) {
        const long long r = row0 + tid;
        float a = 0.0f;
        if (r < N) {
            const float* zr = z + r * D;
            for (int d = 0; d < D; ++d) {
                const float v = zr[d];
                a = fmaf(v, v, a);
            }
        }
        a_s[tid] = a;
    }
    float best[4];
    int bi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        best[i] = INFINITY;
        bi[i] = k_begin;
    }
    for (int k0 = k_begin; k0 < k_end; k0 += S_TN) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
        for (int d0 = 0; d0 < D; d0 += S_TD) {
            __syncthreads();
            {
                const int d = tid & 31, gd = d0 + d;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = i * 8 + (tid >> 5);
                    const long long gr = row0 + r;
                    zs[d][r] = (gr < N && gd < D) ? z[gr * D + gd] : 0.0f;
                    const int k = k0 + r;
                    es[d][r] = (k < k_end && gd < D) ? E[static_cast<size_t>(k) * D + gd] : 0.0f;
                }
            }
            __syncthreads();
#pragma unroll
            for (int d = 0; d < S_TD; ++d) {
                const float4 zv = *reinterpret_cast<const float4*>(&zs[d][ty * 4]);
                const float4 ev = *reinterpret_cast<const float4*>(&es[d][tx * 4]);
                const float zr[4] = {zv.x, zv.y, zv.z, zv.w};
                const float er[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(zr[i], er[j], acc[i][j]);
            }
        }
        // dist = fl(fl(a_n + b_k) - 2 c): vector_quantizer.py:34-36 evaluation order (2c is exact)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < k_end) {
                const float b = __ldg(e_norm2 + k);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float t = a_s[ty * 4 + i] + b;
                    const float dist = fmaf(-2.0f, acc[i][j], t);
                    if (dist < best[i]) {
                        best[i] = dist;
                        bi[i] = k;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        unsigned long long key = pack_key(best[i], bi[i]);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        const long long r = row0 + ty * 4 + i;
        if (tx == 0 && r < N) {
            if (keys != nullptr)
                atomicMin(keys + r, key);
            else
                idx_out[r] = static_cast<int>(key & 0xffffffffu);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// rows: HBM-bound streaming kernel.  A CTA takes groups of R consecutive rows; every thread owns one
// 16-byte (or 4-byte) element of the group, so all lanes are busy for any D and the loads of a whole
// group are in flight together:
//   phase 1  q_out = fl(z + fl(E[idx] - z)); sse += (E[idx]-z)^2; hist[idx] += 1
//   phase 2  the R one-hot rows (K floats each), coalesced 16-byte streaming stores
// The last CTA to finish reduces the per-CTA partial sums in a fixed order and (unless deferred)
// writes loss and perplexity, so a single-GPU forward needs no further launch.
// ---------------------------------------------------------------------------------------------
constexpr int ROWS_MAX_R = 256;

template <bool ONEHOT, bool QUANT, int VEC>
__global__ void __launch_bounds__(256) quantize_rows_kernel(
    const float* __restrict__ z, const float* __restrict__ E, const int* __restrict__ idx_in,
    const unsigned long long* __restrict__ keys, long long N, int K, int D, float beta, float* __restrict__ q_out,
    int* __restrict__ idx_out, float* __restrict__ onehot, float* __restrict__ hist, double* __restrict__ partials,
    unsigned int* __restrict__ counter, float* __restrict__ sse_out, float* __restrict__ loss,
    float* __restrict__ perplexity, int finalize, int R, int onehot_vec_ok) {
    __shared__ double red[8];
    __shared__ bool is_last;
    __shared__ int s_idx[ROWS_MAX_R];
    // small codebooks: reds on the same address queue up in L2 (N / K per code), so the counts are kept per CTA in shared
    // memory and flushed once (integers: exact in any order)
    constexpr int ROWS_SMEM_HIST = 2048;
    __shared__ int s_hist[ROWS_SMEM_HIST];
    const bool smem_hist = K <= ROWS_SMEM_HIST;
    if (smem_hist) {
        for (int k = threadIdx.x; k < K; k += 256) s_hist[k] = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int DV = D / VEC;                 // elements (of VEC floats) per row
    const long long n_groups = (N + R - 1) / R;
    float sse = 0.0f;
    for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const long long row0 = grp * R;
        const int rows_here = static_cast<int>(min(static_cast<long long>(R), N - row0));
        if (threadIdx.x < rows_here) {
            const long long r = row0 + threadIdx.x;
            int code;
            if (keys != nullptr) {
                code = static_cast<int>(keys[r] & 0xffffffffu);
                idx_out[r] = code;
            } else {
                code = idx_in[r];
            }
            s_idx[threadIdx.x] = code;
            if (smem_hist) {
                if (static_cast<unsigned>(code) < static_cast<unsigned>(K)) atomicAdd(s_hist + code, 1);
            } else {
                atomicAdd(hist + code, 1.0f);
            }
        }
        __syncthreads();
        if (QUANT) {
            const int n_el = rows_here * DV;
            if (VEC == 4) {
                // rows are contiguous: element e of the group sits at float4 offset row0 * DV + e.  Four independent
                // 16-byte elements (z and E[idx]) in flight per thread; sse accumulates in element order per thread.
                const float4* z4 = reinterpret_cast<const float4*>(z) + row0 * DV;
                float4* q4 = reinterpret_cast<float4*>(q_out) + row0 * DV;
                for (int e0 = threadIdx.x; e0 < n_el; e0 += 4 * 256) {
                    float4 zv[4], ev[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = e0 + u * 256;
                        if (e < n_el) {
                            const int rr = e / DV, c = e - rr * DV;
                            zv[u] = __ldcs(z4 + e);
                            ev[u] = __ldg(reinterpret_cast<const float4*>(E + static_cast<size_t>(s_idx[rr]) * D) + c);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = e0 + u * 256;
                        if (e < n_el) {
                            float4 df, qv;
                            df.x = ev[u].x - zv[u].x; df.y = ev[u].y - zv[u].y; df.z = ev[u].z - zv[u].z; df.w = ev[u].w - zv[u].w;
                            qv.x = zv[u].x + df.x; qv.y = zv[u].y + df.y; qv.z = zv[u].z + df.z; qv.w = zv[u].w + df.w;
                            __stcs(q4 + e, qv);
                            sse = fmaf(df.x, df.x, sse); sse = fmaf(df.y, df.y, sse);
                            sse = fmaf(df.z, df.z, sse); sse = fmaf(df.w, df.w, sse);
                        }
                    }
                }
            }
            for (int e = threadIdx.x; VEC != 4 && e < n_el; e += 256) {
                const int rr = e / DV, c = e - rr * DV;
                const int code = s_idx[rr];
                if (VEC == 4) {
                } else {
                    const float zv = z[(row0 + rr) * D + c];
                    const float df = __ldg(E + static_cast<size_t>(code) * D + c) - zv;
                    q_out[(row0 + rr) * D + c] = zv + df;
                    sse = fmaf(df, df, sse);
                }
            }
        }
        if (ONEHOT) {
            if (onehot_vec_ok) {
                const int KV = K >> 2;
                for (int rr = 0; rr < rows_here; ++rr) {
                    const int code = s_idx[rr];
                    float4* orow = reinterpret_cast<float4*>(onehot + (row0 + rr) * K);
                    for (int c = threadIdx.x; c < KV; c += 256) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c == (code >> 2)) {
                            const int w = code & 3;
                            v.x = w == 0 ? 1.f : 0.f; v.y = w == 1 ? 1.f : 0.f;
                            v.z = w == 2 ? 1.f : 0.f; v.w = w == 3 ? 1.f : 0.f;
                        }
                        __stcs(orow + c, v);
                    }
                }
            } else {
                for (int rr = 0; rr < rows_here; ++rr) {
                    const int code = s_idx[rr];
                    float* orow = onehot + (row0 + rr) * K;
                    for (int k = threadIdx.x; k < K; k += 256) orow[k] = (k == code) ? 1.0f : 0.0f;
                }
            }
        }
        __syncthreads();   // s_idx is reused by the next group
    }
    if (smem_hist) {
        for (int k = threadIdx.x; k < K; k += 256) {
            const int c = s_hist[k];
            if (c != 0) atomicAdd(hist + k, static_cast<float>(c));
        }
    }
    // block partial (double), then last-CTA-done reduction in a fixed order
    double s = warp_sum_d(static_cast<double>(sse));
    if (lane == 0) red[wib] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        partials[blockIdx.x] = t;
        __threadfence();
        const unsigned int done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double t = 0.0;
    for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += blockDim.x) t += __ldcg(partials + i);
    t = warp_sum_d(t);
    __syncthreads();
    if (lane == 0) red[wib] = t;
    __syncthreads();
    double total = 0.0;
    for (int w = 0; w < 8; ++w) total += red[w];
    if (threadIdx.x == 0) {
        *sse_out = static_cast<float>(total);
        *counter = 0u;
    }
    if (finalize) {
        if (QUANT && threadIdx.x == 0) {
            const float m = static_cast<float>(total / (static_cast<double>(N) * static_cast<double>(D)));
            *loss = __fadd_rn(m, __fmul_rn(beta, m));   // vector_quantizer.py:52 with q_latent == e_latent == m
        }
        double ent = 0.0;
        const float nf = static_cast<float>(N);
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const float p = __fdiv_rn(__ldcg(hist + k), nf);   // vector_quantizer.py:55
            ent += static_cast<double>(p * logf(p + 1e-10f));   // :56
        }
        ent = warp_sum_d(ent);
        __syncthreads();
        if (lane == 0) red[wib] = ent;
        __syncthreads();
        if (threadIdx.x == 0) {
            double e = 0.0;
            for (int w = 0; w < 8; ++w) e += red[w];
            *perplexity = expf(static_cast<float>(-e));
        }
    }
}

// loss / perplexity from (all-reduced) statistics: one CTA
__global__ void __launch_bounds__(256) finalize_stats_kernel(const float* __restrict__ hist,
                                                             const float* __restrict__ sse, long long N_global, int K,
                                                             int D, float beta, float* __restrict__ loss,
                                                             float* __restrict__ perplexity) {
    __shared__ double red[8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double ent = 0.0;
    const float nf = static_cast<float>(N_global);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float p = __fdiv_rn(hist[k], nf);
        ent += static_cast<double>(p * logf(p + 1e-10f));
    }
    ent = warp_sum_d(ent);
    if (lane == 0) red[wib] = ent;
    __syncthreads();
    if (threadIdx.x == 0) {
        double e = 0.0;
        for (int w = 0; w < 8; ++w) e += red[w];
        *perplexity = expf(static_cast<float>(-e));
        if (loss != nullptr && sse != nullptr) {
            const float m =
                static_cast<float>(static_cast<double>(*sse) / (static_cast<double>(N_global) * static_cast<double>(D)));
            *loss = __fadd_rn(m, __fmul_rn(beta, m));
        }
    }
}

// dense one-hot from indices: one warp per row, write-only
__global__ void __launch_bounds__(256) onehot_kernel(const int* __restrict__ idx, long long N, int K,
                                                     float* __restrict__ onehot, int vec_ok) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const long long nwarps = static_cast<long long>(gridDim.x) * 8;
    for (long long r = warp0; r < N; r += nwarps) {
        const int code = idx[r];            // a value outside [0, K) leaves an all-zero row
        float* orow = onehot + r * K;
        if (vec_ok) {
            for (int c = lane; c < (K >> 2); c += 32) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c == (code >> 2)) {
                    const int w = code & 3;
                    v.x = w == 0 ? 1.f : 0.f; v.y = w == 1 ? 1.f : 0.f;
                    v.z = w == 2 ? 1.f : 0.f; v.w = w == 3 ? 1.f : 0.f;
                }
                __stcs(reinterpret_cast<float4*>(orow) + c, v);
            }
        } else {
            for (int k = lane; k < K; k += 32) orow[k] = (k == code) ? 1.0f : 0.0f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward: flat grid-stride over 16-byte elements (thread = one float4 of one row), HBM-bound.
//   dz = g_q - cz*(q - z);  dE[idx] += ce*(q - z)  (red.global.add, 16 bytes per request)
//   cz = g_loss*beta*2/(n_rows_dz*D), ce = g_loss*2/(n_rows_dE*D)   -- oracle/vq_oracle.c order
// ---------------------------------------------------------------------------------------------
template <bool TRAIN_VQ, bool HAS_GQ, int VEC>
__global__ void __launch_bounds__(256) backward_kernel(const float* __restrict__ g_q, const float* __restrict__ g_loss,
                                                       const float* __restrict__ z, const float* __restrict__ E,
                                                       const int* __restrict__ idx, long long N, float denom_dz,
                                                       float denom_dE, int D, float beta, float* __restrict__ dz,
                                                       float* __restrict__ dE, const unsigned int* ready,
                                                       float* __restrict__ repl, int n_repl, long long repl_stride) {
    // n_repl > 1: the CTAs spread their reds over n_repl zeroed copies of dE (reduce_replicas_kernel folds them into dE
    // afterwards).  With few codes the flat scatter is bound by reds queueing on the same L2 addresses (N * D / 4 reds on
    // K * D / 4 addresses: 60 - 90 G reds/s below 64 k addresses against >= 120 G/s, i.e. HBM-bound, above).
    if (TRAIN_VQ && n_repl > 1) dE = repl + static_cast<size_t>(blockIdx.x % static_cast<unsigned>(n_repl)) * repl_stride;
    pdl_launch_dependents();
    // ready != NULL (vq_step_backward right behind the fused forward): start as soon as the forward's last CTA has raised
    // the workspace's ready word -- everything this kernel reads is complete then; the forward's serial statistics tail
    // (loss / perplexity) overlaps this kernel.  Bounded: after a few thousand polls fall back to the full dependency.
    if (ready != nullptr) {
        __shared__ int s_ok;
        if (threadIdx.x == 0) {
            unsigned int polls = 0;
            while (ld_acquire_gpu_u32(ready) != 1u && ++polls < 4096u) {
            }
            s_ok = polls < 4096u ? 1 : 0;
        }
        __syncthreads();
        if (!s_ok) pdl_wait_prior_grids();
    } else {
        pdl_wait_prior_grids();
    }
    const float gl = g_loss != nullptr ? __ldg(g_loss) : 1.0f;
    const float cz = gl * beta * 2.0f / denom_dz;
    const float ce = gl * 2.0f / denom_dE;
    const int DV = D / VEC;
    const long long n_el = N * DV;
    const long long stride = static_cast<long long>(gridDim.x) * 256;
    for (long long e = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; e < n_el; e += stride) {
        const long long r = e / DV;
        const int c = static_cast<int>(e - r * DV);
        const int code = __ldg(idx + r);
        if (VEC == 4) {
            const float4 zv = __ldcs(reinterpret_cast<const float4*>(z) + e);
            const float4 ev = __ldg(reinterpret_cast<const float4*>(E + static_cast<size_t>(code) * D) + c);
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (HAS_GQ) gv = __ldcs(reinterpret_cast<const float4*>(g_q) + e);
            float4 df, o;
            df.x = ev.x - zv.x; df.y = ev.y - zv.y; df.z = ev.z - zv.z; df.w = ev.w - zv.w;
            o.x = fmaf(-cz, df.x, gv.x); o.y = fmaf(-cz, df.y, gv.y);
            o.z = fmaf(-cz, df.z, gv.z); o.w = fmaf(-cz, df.w, gv.w);
            __stcs(reinterpret_cast<float4*>(dz) + e, o);
            if (TRAIN_VQ) {
                float4 a;
                a.x = ce * df.x; a.y = ce * df.y; a.z = ce * df.z; a.w = ce * df.w;
                atomicAdd(reinterpret_cast<float4*>(dE + static_cast<size_t>(code) * D) + c, a);
            }
        } else {
            const float zv = z[e];
            const float df = __ldg(E + static_cast<size_t>(code) * D + c) - zv;
            const float gv = HAS_GQ ? g_q[e] : 0.0f;
            dz[e] = fmaf(-cz, df, gv);
            if (TRAIN_VQ) atomicAdd(dE + static_cast<size_t>(code) * D + c, ce * df);
        }
    }
    // order this kernel behind the forward's completion before it exits: "previous kernel complete" stays transitive
    if (ready != nullptr) pdl_wait_prior_grids();
}

// dE += sum of the n_repl copies the backward spread its reds over (16-byte elements; fixed order over the copies)
__global__ void __launch_bounds__(256) reduce_replicas_kernel(const float* __restrict__ repl, int n_repl, long long n4, float* __restrict__ dE) {
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
        float4 a = reinterpret_cast<const float4*>(dE)[i];
        for (int r = 0; r < n_repl; ++r) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(repl) + static_cast<long long>(r) * n4 + i);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        reinterpret_cast<float4*>(dE)[i] = a;
    }
}

// codebook gradient only (vq_backward with dz == NULL): dE[idx] += ce * (E[idx] - z)
template <int VEC>
__global__ void __launch_bounds__(256) backward_dE_kernel(const float* __restrict__ g_loss, const float* __restrict__ z,
                                                          const float* __restrict__ E, const int* __restrict__ idx, long long N,
                                                          float denom_dE, int D, float* __restrict__ dE, const unsigned int* ready) {
    pdl_launch_dependents();
    if (ready != nullptr) {          // right behind the fused forward: start on its ready word (see backward_kernel)
        __shared__ int s_ok;
        if (threadIdx.x == 0) {
            unsigned int polls = 0;
            while (ld_acquire_gpu_u32(ready) != 1u && ++polls < 4096u) {
            }
            s_ok = polls < 4096u ? 1 : 0;
        }
        __syncthreads();
        if (!s_ok) pdl_wait_prior_grids();
    } else {
        pdl_wait_prior_grids();
    }
    const float gl = g_loss != nullptr ? __ldg(g_loss) : 1.0f;
    const float ce = gl * 2.0f / denom_dE;
    const int DV = D / VEC;
    const long long n_el = N * DV;
    const long long stride = static_cast<long long>(gridDim.x) * 256;
    for (long long e = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; e < n_el; e += stride) {
        const long long r = e / DV;
        const int c = static_cast<int>(e - r * DV);
        const int code = __ldg(idx + r);
        if (VEC == 4) {
            const float4 zv = __ldg(reinterpret_cast<const float4*>(z) + e);
            const float4 ev = __ldg(reinterpret_cast<const float4*>(E + static_cast<size_t>(code) * D) + c);
            float4 a;
            a.x = ce * (ev.x - zv.x); a.y = ce * (ev.y - zv.y); a.z = ce * (ev.z - zv.z); a.w = ce * (ev.w - zv.w);
            atomicAdd(reinterpret_cast<float4*>(dE + static_cast<size_t>(code) * D) + c, a);
        } else {
            atomicAdd(dE + static_cast<size_t>(code) * D + c, ce * (__ldg(E + static_cast<size_t>(code) * D + c) - z[e]));
        }
    }
    if (ready != nullptr) pdl_wait_prior_grids();      // behind the forward's completion before it exits (transitivity)
}

// ---------------------------------------------------------------------------------------------
// SURVEY 8(f) rank 1 -- the consumer of the dense one-hot as an index gather.
// LocationModule.fc_1 (location_model.py:10,21) multiplies flatten(one_hot(B, T, K)) by a (O, T*K) weight: a
// 205 824 x 1024 GEMM on a matrix that is 99.9 % zeros.  With the weight stored transposed, Wt (T*K, O),
//     y[b, :] = bias + sum_t Wt[t*K + idx[b, t], :]
// is T coalesced row reads per sample.  One CTA per (sample, 1024-wide slab of O); thread = one float4 of O.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_sum_rows_kernel(const int* __restrict__ idx, const float* __restrict__ Wt,
                                                              const float* __restrict__ bias, float* __restrict__ y, int T,
                                                              int K, int O) {
    extern __shared__ int s_rows[];                       // [T] row of Wt for every position of this sample
    const int b = blockIdx.x;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const int code = idx[b * T + t];
        s_rows[t] = static_cast<unsigned>(code) < static_cast<unsigned>(K) ? t * K + code : -1;   // -1: not a code, contributes nothing
    }
    __syncthreads();
    const int o4 = blockIdx.y * blockDim.x + threadIdx.x;  // float4 index along O
    if (o4 * 4 >= O) return;
    const float4* W4 = reinterpret_cast<const float4*>(Wt);
    const int O4 = O >> 2;
    float4 acc = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias) + o4) : make_float4(0.f, 0.f, 0.f, 0.f);
    int t = 0;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto row_at = [&](int r) { return r >= 0 ? __ldg(W4 + static_cast<size_t>(r) * O4 + o4) : zero4; };
    for (; t + 4 <= T; t += 4) {                           // 4 independent row reads in flight
        const float4 v0 = row_at(s_rows[t]);
        const float4 v1 = row_at(s_rows[t + 1]);
        const float4 v2 = row_at(s_rows[t + 2]);
        const float4 v3 = row_at(s_rows[t + 3]);
        acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
        acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
        acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
        acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
    }
    for (; t < T; ++t) {
        const float4 v = row_at(s_rows[t]);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(y + static_cast<size_t>(b) * O)[o4] = acc;
}

// backward of the above w.r.t. Wt (dense gradient): dWt[t*K + idx[b,t], :] += g[b, :]  (rows collide across samples)
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const int* __restrict__ idx, const float* __restrict__ g,
                                                               float* __restrict__ dWt, int T, int K, int O) {
    const int b = blockIdx.x, t = blockIdx.y;
    const int code = idx[b * T + t];
    if (static_cast<unsigned>(code) >= static_cast<unsigned>(K)) return;      // not a code: no gradient row
    const size_t row = static_cast<size_t>(t) * K + code;
    const int O4 = O >> 2;
    for (int o4 = threadIdx.x; o4 < O4; o4 += blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g + static_cast<size_t>(b) * O) + o4);
        atomicAdd(reinterpret_cast<float4*>(dWt + row * O) + o4, v);
    }
}

// ---------------------------------------------------------------------------------------------
// SURVEY 8(f) rank 2 (first half) -- the time-mean variant in front of the quantizer
// (convolutional_vq_vae.py:96-97: `z = torch.mean(z, dim=2, keepdim=True)` between `_pre_vq_conv` and `_vq`).
// x is (rows = B*D, T); z[row] = sum_t x[row, t] / T.  One warp per row: coalesced 16-byte loads, a fixed summation
// order (per-lane partial sums over t = lane*4 + 128*i + j in increasing i, then the xor-shuffle tree), so the result
// is reproducible run to run; it differs from torch.mean's order by fp32 rounding only.
// The backward spreads dz[row] / T over the T positions.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) time_mean_kernel(const float* __restrict__ x, long long rows, int T, float* __restrict__ z) {
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    const int lane = threadIdx.x & 31;
    const long long warp0 = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const long long nwarps = static_cast<long long>(gridDim.x) * 8;
    const bool vec = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
    for (long long r = warp0; r < rows; r += nwarps) {
        const float* xr = x + r * T;
        float acc = 0.0f;
        if (vec) {
            const float4* x4 = reinterpret_cast<const float4*>(xr);
            for (int i = lane; i < (T >> 2); i += 32) {
                const float4 v = __ldcs(x4 + i);
                acc += v.x; acc += v.y; acc += v.z; acc += v.w;
            }
        } else {
            for (int t = lane; t < T; t += 32) acc += xr[t];
        }
        acc = warp_sum(acc);
        if (lane == 0) z[r] = acc / static_cast<float>(T);
    }
}

__global__ void __launch_bounds__(256) time_mean_backward_kernel(const float* __restrict__ dz, long long rows, int T, float* __restrict__ dx) {
    const long long n = rows * T;
    const float inv = 1.0f / static_cast<float>(T);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        dx[i] = __ldg(dz + i / T) * inv;
}

// ---------------------------------------------------------------------------------------------
// SURVEY 8(f) rank 3 -- Jitter (modules/jitter.py:47-70) as one gather along time.
// q is (rows = B*D, T); column t becomes the ORIGINAL column src[t] (src[t] in {t-1, t, t+1}).  One CTA per row:
// the row is staged in shared memory first, so the edit is in place without the reference's full clone.
// `backward` zeroes the gradient of replaced columns (their values came from a detached clone).
// ---------------------------------------------------------------------------------------------
// whole row staged in shared memory (T floats), then rewritten in place: no clone of the tensor is needed
__global__ void __launch_bounds__(256) jitter_gather_kernel(float* __restrict__ q, const int* __restrict__ src, int T) {
    extern __shared__ float jrow[];
    float* qr = q + static_cast<long long>(blockIdx.x) * T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) jrow[t] = qr[t];
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const int sc = src[t];
        if (sc != t && static_cast<unsigned>(sc) < static_cast<unsigned>(T)) qr[t] = jrow[sc];
    }
}
__global__ void __launch_bounds__(256) jitter_backward_kernel(float* __restrict__ g, const int* __restrict__ src, long long rows, int T) {
    const long long n = rows * T;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(i % T);
        if (src[t] != t) g[i] = 0.0f;
    }
}

}  // namespace b200vq
