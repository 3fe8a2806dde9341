// kernels_bwd.cuh -- backward of the VQ hot path (autograd of vector_quantizer.py:46-54).
//
//   dz[n,:]  = g_q[n,:] - cz * (E[idx[n]] - z[n])          cz = g_loss * beta * 2 / (n_rows_dz * D)
//   dE[k,:]  = ce * sum_{n : idx[n] == k} (E[k] - z[n])     ce = g_loss * 2 / (n_rows_dE * D)
//
// The flat kernel in kernels_simt.cuh issues one 16-byte red.global.add per 16-byte element of z (823 k atomics on
// 16 k addresses at the RIR-256 shape: 11.0 us back to back against 7.7 us for the dz pass alone).
//   backward_private_kernel  N >> K with a small codebook (the sweep's K = 512 corner): persistent CTAs keep a PRIVATE
//       copy of (a column slice of) dE in shared memory; rows are bucketed by code inside a window so that every
//       table row has one owning warp (plain shared-memory read-modify-write, no atomics) and the table is flushed
//       once per CTA with 16-byte reds.
// Measured and dropped (round 2):
//   * code-owner CTAs that scan idx, sort their rows and sum them in registers next to the dz streamers in one grid (no
//     atomics, single writer per dE element): the owners' latency-bound chains (idx scan, sort, row batches) stall while
//     the streamers saturate the memory system -- 45 - 65 us against 11 us flat;
//   * accumulating the code sums S_k = sum (E_k - z_n) in the FORWARD's row epilogue (E[idx] - z is in registers there)
//     and streaming dz only in the backward: the per-SM red issue rate (three worker warps per SM) puts 18 - 22 us on
//     the forward's critical tail to save 1.5 us in the backward.
#pragma once
#include "common.cuh"

namespace b200vq {

// ---------------------------------------------------------------------------------------------
// dz pass over elements [e_begin, e_end) with a stride: two independent 16-byte elements in flight per thread
// ---------------------------------------------------------------------------------------------
template <bool HAS_GQ>
__device__ __forceinline__ void dz_stream(const float* __restrict__ g_q, const float* __restrict__ z, const float* __restrict__ E,
                                          const int* __restrict__ idx, float* __restrict__ dz, int K, int D, float cz, long long first,
                                          long long n_el, long long stride) {
    const int DV = D >> 2;
    for (long long e0 = first; e0 < n_el; e0 += 2 * stride) {
        float4 zv[2], ev[2], gv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long e = e0 + u * stride;
            if (e < n_el) {
                const long long r = e / DV;
                const int c = static_cast<int>(e - r * DV);
                const int code = __ldg(idx + r);
                zv[u] = __ldcs(reinterpret_cast<const float4*>(z) + e);
                // a code outside [0, K) (never produced by vq_forward) contributes nothing: dz = g_q
                ev[u] = static_cast<unsigned>(code) < static_cast<unsigned>(K)
                            ? __ldg(reinterpret_cast<const float4*>(E + static_cast<size_t>(code) * D) + c) : zv[u];
                gv[u] = HAS_GQ ? __ldcs(reinterpret_cast<const float4*>(g_q) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long e = e0 + u * stride;
            if (e < n_el) {
                float4 df, o;
                df.x = ev[u].x - zv[u].x; df.y = ev[u].y - zv[u].y; df.z = ev[u].z - zv[u].z; df.w = ev[u].w - zv[u].w;
                o.x = fmaf(-cz, df.x, gv[u].x); o.y = fmaf(-cz, df.y, gv[u].y);
                o.z = fmaf(-cz, df.z, gv[u].z); o.w = fmaf(-cz, df.w, gv[u].w);
                __stcs(reinterpret_cast<float4*>(dz) + e, o);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// code owner: see the header.  NI = ceil(D / 32) floats per lane and row (lane l holds d = l + 32 i).
// The owner code runs once per CTA, so its size is what it costs (instruction fetch): the rare paths are kept
// out of line and there is a single flush site.
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// backward_private_kernel (N >> K): grid = (row chunks, column slices); a CTA owns the rows
// [chunk * rows_per_cta, ...) and the columns [slice * 32 * NC, (slice + 1) * 32 * NC) of z / g_q / dz / dE and keeps
// its share of dE, s_tab[K][32 * NC], in shared memory.  Rows are taken PV_WIN at a time: their codes are read once,
// bucketed by (code % warps) with a counting sort, and warp w then walks the rows whose code it owns: one coalesced
// 128 * NC-byte read of z (and g_q), the dz store, and a plain shared-memory read-modify-write of the table row
// (lane l holds columns l + 32 i; no two warps ever touch the same table row -> no atomics, no bank conflicts).
// The table is flushed once with red.global.add, so global atomics drop from N * D to (row chunks) * K * D.
// ---------------------------------------------------------------------------------------------
constexpr int PV_THREADS = 512;
constexpr int PV_WARPS = PV_THREADS / 32;
constexpr int PV_WIN = 2048;          // rows bucketed at a time (PV_THREADS * 4)

__host__ __device__ constexpr int pv_smem_bytes(int K, int nc) {
    return K * 32 * nc * 4 /* table */ + PV_WIN * 4 /* sorted rows */ + PV_WIN * 4 /* codes */ + 3 * (PV_WARPS + 1) * 4 + 64;
}

template <bool HAS_GQ, int NC>
__global__ void __launch_bounds__(PV_THREADS, 1)
backward_private_kernel(const float* __restrict__ g_q, const float* __restrict__ g_loss, const float* __restrict__ z,
                        const float* __restrict__ E, const int* __restrict__ idx, long long N, long long rows_per_cta,
                        float denom_dz, float denom_dE, int K, int D, float beta, float* __restrict__ dz, float* __restrict__ dE) {
    extern __shared__ __align__(16) uint8_t pv_smem[];
    constexpr int W = 32 * NC;                                  // columns of this CTA's slice
    float* s_tab = reinterpret_cast<float*>(pv_smem);           // [K][W]
    int* s_rows = reinterpret_cast<int*>(s_tab + static_cast<size_t>(K) * W);   // [PV_WIN] window-local row, grouped by owning warp
    int* s_code = s_rows + PV_WIN;                              // [PV_WIN] code of window-local row
    int* s_wcnt = s_code + PV_WIN;                              // [PV_WARPS + 1]
    int* s_woff = s_wcnt + PV_WARPS + 1;                        // [PV_WARPS + 1]
    int* s_wfill = s_woff + PV_WARPS + 1;                       // [PV_WARPS + 1]
    pdl_launch_dependents();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < K * W / 4; i += PV_THREADS) reinterpret_cast<float4*>(s_tab)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    pdl_wait_prior_grids();
    const float gl = g_loss != nullptr ? __ldg(g_loss) : 1.0f;
    const float cz = gl * beta * 2.0f / denom_dz;
    const float ce = gl * 2.0f / denom_dE;
    const int col0 = blockIdx.y * W;
    const long long row_lo = static_cast<long long>(blockIdx.x) * rows_per_cta;
    const long long row_hi = min(N, row_lo + rows_per_cta);
    __syncthreads();
    // codes of a window are fetched one window ahead, so their DRAM latency hides behind the previous window's walk
    int nxt[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const long long r = row_lo + tid + u * PV_THREADS;
        nxt[u] = r < row_hi ? __ldg(idx + r) : -1;
    }
    for (long long w0 = row_lo; w0 < row_hi; w0 += PV_WIN) {
        const int nwin = static_cast<int>(min(static_cast<long long>(PV_WIN), row_hi - w0));
        // ---- bucket the window's rows by owning warp (code % PV_WARPS) ----
        if (tid <= PV_WARPS) s_wcnt[tid] = 0;
        __syncthreads();
        int codes[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = tid + u * PV_THREADS;
            codes[u] = nxt[u];
            if (static_cast<unsigned>(codes[u]) >= static_cast<unsigned>(K)) codes[u] = -1;   // never index outside the table
            if (codes[u] >= 0) atomicAdd(&s_wcnt[codes[u] & (PV_WARPS - 1)], 1);
            if (r < nwin) s_code[r] = codes[u];
            const long long rn = w0 + PV_WIN + r;
            nxt[u] = rn < row_hi ? __ldg(idx + rn) : -1;
        }
        __syncthreads();
        if (tid == 0) {
            int a = 0;
            for (int w = 0; w < PV_WARPS; ++w) {
                s_woff[w] = a;
                s_wfill[w] = 0;
                a += s_wcnt[w];
            }
            s_woff[PV_WARPS] = a;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (codes[u] >= 0) {
                const int w = codes[u] & (PV_WARPS - 1);
                s_rows[s_woff[w] + atomicAdd(&s_wfill[w], 1)] = tid + u * PV_THREADS;
            }
        }
        __syncthreads();
        // ---- warp `warp` walks its rows: RB rows (RB * NC * 256 bytes of z and g_q) in flight ----
        constexpr int RB = NC == 1 ? 16 : 8;
        const int lo = s_woff[warp], hi = s_woff[warp + 1];
        for (int p0 = lo; p0 < hi; p0 += RB) {
            float zv[RB][NC], gv[RB][NC], ev[RB][NC];
            int rl[RB], cd[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                rl[u] = p0 + u < hi ? s_rows[p0 + u] : -1;
                if (rl[u] >= 0) {
                    cd[u] = s_code[rl[u]];
                    const long long off = (w0 + rl[u]) * D + col0 + lane;
                    const float* er = E + static_cast<size_t>(cd[u]) * D + col0 + lane;
#pragma unroll
                    for (int i = 0; i < NC; ++i) {
                        zv[u][i] = __ldcs(z + off + 32 * i);
                        gv[u][i] = HAS_GQ ? __ldcs(g_q + off + 32 * i) : 0.0f;
                        ev[u][i] = __ldg(er + 32 * i);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                if (rl[u] >= 0) {
                    const long long off = (w0 + rl[u]) * D + col0 + lane;
                    float* trow = s_tab + static_cast<size_t>(cd[u]) * W + lane;
#pragma unroll
                    for (int i = 0; i < NC; ++i) {
                        const float df = ev[u][i] - zv[u][i];
                        if (dz != nullptr) __stcs(dz + off + 32 * i, fmaf(-cz, df, gv[u][i]));
                        trow[32 * i] += df;                 // this warp owns the table row: plain read-modify-write
                    }
                }
            }
        }
        // rows whose code fell outside [0, K) (never produced by vq_forward): dz = g_q, no dE contribution
        if (dz != nullptr) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = tid + u * PV_THREADS;
                if (r < nwin && codes[u] < 0) {
                    for (int c = 0; c < W; ++c) {
                        const long long off = (w0 + r) * D + col0 + c;
                        dz[off] = HAS_GQ ? g_q[off] : 0.0f;
                    }
                }
            }
        }
        __syncthreads();
    }
    // ---- flush: one 16-byte red per 4 table entries that are not zero ----
    for (int i = tid; i < K * (W / 4); i += PV_THREADS) {
        const int k = i / (W / 4), c4 = i - k * (W / 4);
        float4 v = reinterpret_cast<const float4*>(s_tab)[i];
        if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) {
            v.x *= ce; v.y *= ce; v.z *= ce; v.w *= ce;
            atomicAdd(reinterpret_cast<float4*>(dE + static_cast<size_t>(k) * D + col0) + c4, v);
        }
    }
}

}  // namespace b200vq
