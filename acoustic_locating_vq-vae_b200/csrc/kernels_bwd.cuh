// kernels_bwd.cuh -- backward of the VQ hot path (autograd of vector_quantizer.py:46-54) without one global
// atomic per element.
//
//   dz[n,:]  = g_q[n,:] - cz * (E[idx[n]] - z[n])          cz = g_loss * beta * 2 / (n_rows_dz * D)
//   dE[k,:]  = ce * sum_{n : idx[n] == k} (E[k] - z[n])     ce = g_loss * 2 / (n_rows_dE * D)
//
// The flat kernel in kernels_simt.cuh issues one 16-byte red.global.add per 16-byte element of z: at the RIR-256
// shape that is 823 k atomics on 16 k addresses and the LSU's atomic issue rate (not HBM) sets the time.  Two
// kernels replace it, chosen by the launcher:
//
//   backward_bucket_kernel   N up to a few 100 k rows.  The grid has two roles that share nothing but inputs:
//       * code owners (the first BK_NB CTAs): CTA j owns the codes k with k % BK_NB == j.  It scans idx once
//         (L2-resident, 4 bytes per row), collects the rows of its codes in shared memory, sorts them by code
//         (counting sort) and sums (E[k] - z[n]) per code with warp-wide coalesced row reads from L2 into
//         REGISTERS; every dE element has exactly one writer -> plain stores, no zeroing, no global atomics, and
//         under data parallelism the owner can send its rows straight to the peers (kernels_dp.cuh).
//       * dz streamers (the other CTAs): the HBM-bound pass over z / g_q / dz, no atomics.
//   backward_private_kernel  N >> K.  Persistent CTAs keep a PRIVATE copy of (a column slice of) dE in shared
//       memory; rows are bucketed by code inside a chunk so that every table row has one owning warp (plain
//       shared-memory read-modify-write, no atomics) and the table is flushed once per CTA with 16-byte reds.
#pragma once
#include "common.cuh"

namespace b200vq {

constexpr int BK_THREADS = 512;
constexpr int BK_WARPS = BK_THREADS / 32;
constexpr int BK_NB = 128;        // code-owner CTAs; code k belongs to CTA (k & 127), local code = k >> 7
constexpr int BK_NB_LOG2 = 7;
constexpr int BK_LIST = 4096;     // rows collected before a flush (a scan pass adds at most BK_THREADS * 4)
constexpr int BK_TABLE = 2048;    // floats: ceil(K / BK_NB) * D must fit
constexpr int BK_MAXLC = 64;      // local codes per owner CTA (K <= 8192)
constexpr int BK_DZ_EPT = 4;      // 16-byte elements per dz-streamer thread

// what an owner does with a finished dE row: `vals[i]` is element d = lane + 32 * i of row `code`
struct StoreDE {
    float* dE;
    int D;
    bool accumulate;              // false: dE is overwritten (VQ_FLAG_ZERO_DE semantics without a memset)
    template <int NI>
    __device__ __forceinline__ void row(int code, int lane, const float (&vals)[NI]) const {
        float* dst = dE + static_cast<size_t>(code) * D;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int d = lane + 32 * i;
            if (d < D) dst[d] = accumulate ? dst[d] + vals[i] : vals[i];
        }
    }
    __device__ __forceinline__ void finish(int /*owner*/, int /*tid*/) const {}
};

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
// ---------------------------------------------------------------------------------------------
template <int NI, typename Sink>
__device__ __forceinline__ void bucket_owner(const float* __restrict__ z, const float* __restrict__ E, const int* __restrict__ idx,
                                             long long N, int K, int D, float ce, int owner, const Sink& sink) {
    __shared__ int s_list[BK_LIST];        // row | local code << 24, in arrival order
    __shared__ int s_sorted[BK_LIST];      // rows, grouped by local code
    __shared__ float s_table[BK_TABLE];    // [local code][D] running sums
    __shared__ int s_cnt[BK_MAXLC];        // entries per local code in s_list
    __shared__ int s_off[BK_MAXLC + 1];    // exclusive prefix of s_cnt
    __shared__ int s_fill[BK_MAXLC];
    __shared__ int s_n;
    __shared__ int s_flush;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncl = (K + BK_NB - 1) >> BK_NB_LOG2;
    for (int i = tid; i < ncl * D; i += BK_THREADS) s_table[i] = 0.0f;
    if (tid < BK_MAXLC) s_cnt[tid] = 0;
    if (tid == 0) s_n = 0;
    __syncthreads();

    auto append = [&](int code, long long row) {
        if ((code & (BK_NB - 1)) == owner && static_cast<unsigned>(code) < static_cast<unsigned>(K)) {
            const int lc = code >> BK_NB_LOG2;
            const int pos = atomicAdd(&s_n, 1);
            s_list[pos] = static_cast<int>(row) | (lc << 24);
            atomicAdd(&s_cnt[lc], 1);
        }
    };

    // sums the rows collected so far into s_table and empties the list; called by the whole CTA
    auto flush = [&]() {
        const int n = s_n;
        if (warp == 0) {                                    // exclusive scan over <= 64 local codes
            int a = lane < ncl ? s_cnt[lane] : 0;
            int b = lane + 32 < ncl ? s_cnt[lane + 32] : 0;
            int ia = a, ib = b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
                if (lane >= o) { ia += ta; ib += tb; }
            }
            const int tot_a = __shfl_sync(0xffffffffu, ia, 31);
            s_off[lane] = ia - a;
            s_off[lane + 32] = tot_a + ib - b;
            s_fill[lane] = 0;
            s_fill[lane + 32] = 0;
            if (lane == 0) s_off[BK_MAXLC] = n;
        }
        __syncthreads();
        for (int i = tid; i < n; i += BK_THREADS) {         // counting sort: scatter by local code
            const int en = s_list[i];
            const int lc = en >> 24;
            s_sorted[s_off[lc] + atomicAdd(&s_fill[lc], 1)] = en;
        }
        __syncthreads();
        // every warp takes an equal share of the sorted rows; a run of one code accumulates in registers
        const int lo = static_cast<int>(static_cast<long long>(n) * warp / BK_WARPS);
        const int hi = static_cast<int>(static_cast<long long>(n) * (warp + 1) / BK_WARPS);
        constexpr int RB = NI <= 2 ? 8 : (NI <= 4 ? 4 : 2);   // rows in flight per warp
        float acc[NI], er[NI];
        int cur = -1;
#pragma unroll
        for (int i = 0; i < NI; ++i) acc[i] = er[i] = 0.0f;
        auto spill = [&]() {
            if (cur >= 0) {
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int d = lane + 32 * i;
                    if (d < D) atomicAdd(&s_table[cur * D + d], acc[i]);
                }
            }
        };
        for (int p0 = lo; p0 < hi; p0 += RB) {
            float zv[RB][NI];
            int ent[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                ent[u] = p0 + u < hi ? s_sorted[p0 + u] : -1;
                if (ent[u] >= 0) {
                    const float* zr = z + static_cast<long long>(ent[u] & 0xffffff) * D;
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        const int d = lane + 32 * i;
                        zv[u][i] = d < D ? __ldg(zr + d) : 0.0f;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                if (ent[u] >= 0) {
                    const int lc = ent[u] >> 24;
                    if (lc != cur) {                        // warp-uniform: a new run starts
                        spill();
                        cur = lc;
                        const float* erow = E + static_cast<size_t>((lc << BK_NB_LOG2) + owner) * D;
#pragma unroll
                        for (int i = 0; i < NI; ++i) {
                            const int d = lane + 32 * i;
                            er[i] = d < D ? __ldg(erow + d) : 0.0f;
                            acc[i] = 0.0f;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < NI; ++i) acc[i] += er[i] - zv[u][i];
                }
            }
        }
        spill();
        __syncthreads();
        if (tid < BK_MAXLC) s_cnt[tid] = 0;
        if (tid == 0) s_n = 0;
        __syncthreads();
    };

    // ---- scan idx: 4 rows per thread and pass, the next pass's load in flight while this one is filed ----
    const long long n4 = N >> 2;
    const int4* idx4 = reinterpret_cast<const int4*>(idx);
    long long p = tid;
    int4 nxt = p < n4 ? __ldg(idx4 + p) : make_int4(-1, -1, -1, -1);
    for (long long base = 0; base < n4; base += BK_THREADS) {
        const int4 cur4 = nxt;
        const long long pn = base + BK_THREADS + tid;
        nxt = pn < n4 ? __ldg(idx4 + pn) : make_int4(-1, -1, -1, -1);
        const long long r0 = (base + tid) << 2;
        if (base + tid < n4) {
            append(cur4.x, r0);
            append(cur4.y, r0 + 1);
            append(cur4.z, r0 + 2);
            append(cur4.w, r0 + 3);
        }
        __syncthreads();
        if (tid == 0) s_flush = s_n > BK_LIST - 4 * BK_THREADS ? 1 : 0;   // the next pass could overflow the list
        __syncthreads();
        if (s_flush) flush();                               // uniform: written once between two barriers
    }
    {
        const long long r = (n4 << 2) + tid;                // the N % 4 tail rows
        if (r < N) append(__ldg(idx + r), r);
        __syncthreads();
    }
    flush();

    // ---- hand the owned rows over (every element exactly once) ----
    for (int lc = warp; lc < ncl; lc += BK_WARPS) {
        const int code = (lc << BK_NB_LOG2) + owner;
        if (code < K) {
            float vals[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int d = lane + 32 * i;
                vals[i] = d < D ? ce * s_table[lc * D + d] : 0.0f;
            }
            sink.template row<NI>(code, lane, vals);
        }
    }
    sink.finish(owner, tid);
}

template <bool HAS_GQ, int NI, typename Sink>
__global__ void __launch_bounds__(BK_THREADS, 2)
backward_bucket_kernel(const float* __restrict__ g_q, const float* __restrict__ g_loss, const float* __restrict__ z,
                       const float* __restrict__ E, const int* __restrict__ idx, long long N, float denom_dz, float denom_dE, int K,
                       int D, float beta, float* __restrict__ dz, const Sink sink) {
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    const float gl = g_loss != nullptr ? __ldg(g_loss) : 1.0f;
    if (blockIdx.x < BK_NB) {
        bucket_owner<NI>(z, E, idx, N, K, D, gl * 2.0f / denom_dE, static_cast<int>(blockIdx.x), sink);
    } else if (dz != nullptr) {
        const long long n_el = N * (D >> 2);
        const long long nthreads = static_cast<long long>(gridDim.x - BK_NB) * BK_THREADS;
        dz_stream<HAS_GQ>(g_q, z, E, idx, dz, K, D, gl * beta * 2.0f / denom_dz,
                          static_cast<long long>(blockIdx.x - BK_NB) * BK_THREADS + threadIdx.x, n_el, nthreads);
    }
}

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

// ---------------------------------------------------------------------------------------------
// prep: |E_k|^2 as ONE sequential fmaf chain per code (the order oracle/vq_oracle.c:norm2_chain fixes), one THREAD per
// code with all of its row loads in flight (the warp-per-code version serialised 64 shuffles per code); the tf32
// hi / lo split of E and the per-step state reset are a coalesced grid-stride pass of the same launch.
// Needs D % 4 == 0 and 16-byte aligned E / E_hi / E_lo / dE.
// ---------------------------------------------------------------------------------------------
constexpr int PREP_THREADS = 128;

__global__ void __launch_bounds__(PREP_THREADS)
prep_codebook_fast_kernel(const float* __restrict__ E, int K, int D, float* __restrict__ e_norm2, float* __restrict__ E_hi,
                          float* __restrict__ E_lo, float* __restrict__ hist_zero, unsigned int* __restrict__ counter_zero,
                          float* __restrict__ dE_zero) {
    pdl_launch_dependents();
    const int tid = blockIdx.x * PREP_THREADS + threadIdx.x, nth = gridDim.x * PREP_THREADS;
    const int DV = D >> 2;
    pdl_wait_prior_grids();      // E may have just been written (optimizer step); the outputs may still be in use
    float acc = 0.0f;
    const int code = blockIdx.x * 32 + threadIdx.x;             // warp 0 of every CTA: one code per lane
    const bool owns = threadIdx.x < 32 && code < K;
    if (owns) {
        const float4* row = reinterpret_cast<const float4*>(E + static_cast<size_t>(code) * D);
        for (int i0 = 0; i0 < DV; i0 += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = i0 + u < DV ? __ldg(row + i0 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {                       // zero padding adds exact zeros
                acc = fmaf(v[u].x, v[u].x, acc);
                acc = fmaf(v[u].y, v[u].y, acc);
                acc = fmaf(v[u].z, v[u].z, acc);
                acc = fmaf(v[u].w, v[u].w, acc);
            }
        }
    }
    if (owns) e_norm2[code] = acc;
    if (E_hi != nullptr) {
        for (int i = tid; i < K * DV; i += nth) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(E) + i);
            float4 h, l;
            h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
            reinterpret_cast<float4*>(E_hi)[i] = h;
            if (E_lo != nullptr) {
                l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
                reinterpret_cast<float4*>(E_lo)[i] = l;
            }
        }
    }
    if (hist_zero != nullptr)
        for (int k = tid; k < K; k += nth) hist_zero[k] = 0.0f;
    if (counter_zero != nullptr && tid == 0) *counter_zero = 0u;
    if (dE_zero != nullptr)
        for (int i = tid; i < K * DV; i += nth) reinterpret_cast<float4*>(dE_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace b200vq
