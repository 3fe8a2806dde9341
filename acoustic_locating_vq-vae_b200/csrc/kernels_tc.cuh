// kernels_tc.cuh -- the tensor-core distance/argmin kernel (sm_100a: TMA + tcgen05 + TMEM).
//
// Computes, for a tile of 128 rows of z and a range of the codebook, the first-index argmin of
//     dist(n,k) = fl( fl(|z_n|^2 + |E_k|^2) - 2 * <z_n, E_k> )          (vector_quantizer.py:34-38)
// with the contraction <z_n, E_k> on the 5th-gen tensor cores.  tcgen05 has no fp32 input kind, so the
// fp32 product is rebuilt from three TF32 MMAs (3xTF32 split):
//     z = z_hi + z_lo,  E = E_hi + E_lo   (hi = tf32_rna(x), lo = tf32_rna(x - hi))
//     <z,E> ~= z_hi.E_lo + z_lo.E_hi + z_hi.E_hi      (the lo.lo term, ~2^-22 relative, is dropped)
// all three accumulated into ONE fp32 TMEM accumulator, small terms first.
//
// CTA = 128 rows x (codes_per_split) codes, processed as code tiles of 128:
//   warp 0     TMA producer: z tile once (NSLAB boxes of 128 rows x 32 fp32, 128B-swizzled), then a
//              ring of NSTAGE 16 KB stages streaming E_lo / E_hi slabs of every code tile
//   warp 1     MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=128, K=8)
//   warp 2     TMEM allocator (256 columns = two 128-column accumulators, double buffered)
//   warps 4-7  epilogue: tcgen05.ld the accumulator (thread = row), form the distance in the
//              reference's evaluation order, keep a running (min, index) with strict '<' in
//              ascending k (first-index tie-break); the N x K matrix never leaves TMEM/registers
// All 8 warps first split the raw z tile in place into z_hi / z_lo (generic-proxy writes followed by
// fence.proxy.async) while the first E stages are already in flight.
#pragma once
#include "common.cuh"

namespace b200vq {

constexpr int TC_ROWS = 128;                 // rows of z per CTA  (UMMA M)
constexpr int TC_CODES = 128;                // codes per accumulator tile (UMMA N)
constexpr int TC_SLAB_FLOATS = 32;           // 128 bytes of fp32 along D = one swizzle row
constexpr int TC_SLAB_BYTES = TC_ROWS * 128; // 16 KB: 128 rows x 128 B
constexpr int TC_THREADS = 256;
constexpr int TC_TMEM_COLS = 256;

__host__ __device__ constexpr int tc_smem_bytes(int nslab, int nstage) {
    // z_hi + z_lo + stages + b tile (2 x 128 fp32) + barriers/tmem slot (256 B) + 1 KB alignment slack
    return 2 * nslab * TC_SLAB_BYTES + nstage * TC_SLAB_BYTES + 2 * TC_CODES * 4 + 256 + 1024;
}

template <int NSLAB, int NSTAGE>
__global__ void __launch_bounds__(TC_THREADS, 1)
argmin_tc_kernel(const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_ehi,
                 const __grid_constant__ CUtensorMap tm_elo, const float* __restrict__ e_norm2, long long N, int K,
                 int codes_per_split, int* __restrict__ idx_out, unsigned long long* __restrict__ keys,
                 float* __restrict__ hist_to_zero, unsigned int* __restrict__ counter_to_zero) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (128B-swizzle atoms) by OFFSET, so the compiler still knows this is shared memory
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* z_hi = smem;
    uint8_t* z_lo = z_hi + NSLAB * TC_SLAB_BYTES;
    uint8_t* stages = z_lo + NSLAB * TC_SLAB_BYTES;
    float* b_tile = reinterpret_cast<float*>(stages + NSTAGE * TC_SLAB_BYTES);   // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_tile + 2 * TC_CODES);
    uint64_t* bar_z = bars;                       // z tile landed
    uint64_t* bar_full = bars + 1;                // [NSTAGE] E slab landed
    uint64_t* bar_empty = bar_full + NSTAGE;      // [NSTAGE] E slab consumed by the MMAs
    uint64_t* bar_acc_full = bar_empty + NSTAGE;  // [2] accumulator complete
    uint64_t* bar_acc_empty = bar_acc_full + 2;   // [2] accumulator drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row_tile = blockIdx.x;
    const int k_begin = blockIdx.y * codes_per_split;
    const int n_ctiles = codes_per_split / TC_CODES;
    const int n_loads = n_ctiles * 2 * NSLAB;     // per code tile: NSLAB slabs of E_lo, then NSLAB of E_hi

    if (blockIdx.x == 0 && blockIdx.y == 0) {     // forward-state reset rides along
        for (int k = threadIdx.x; k < K; k += TC_THREADS) hist_to_zero[k] = 0.0f;
        if (threadIdx.x == 0) *counter_to_zero = 0u;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_z);
        tma_prefetch_desc(&tm_ehi);
        tma_prefetch_desc(&tm_elo);
        mbar_init(bar_z, 1);
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_acc_full + b, 1);
            mbar_init(bar_acc_empty + b, 128);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, TC_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- producer prologue: z tile + the first ring of E stages (no waits needed yet) -------------
    auto issue_e_load = [&](int L) {
        const int stage = L % NSTAGE;
        const int ct = L / (2 * NSLAB), i = L % (2 * NSLAB);
        const bool lo = i < NSLAB;
        const int slab = lo ? i : i - NSLAB;
        mbar_arrive_expect_tx(bar_full + stage, TC_SLAB_BYTES);
        tma_load_2d(stages + stage * TC_SLAB_BYTES, lo ? &tm_elo : &tm_ehi, bar_full + stage, slab * TC_SLAB_FLOATS,
                    k_begin + ct * TC_CODES);
    };
    if (warp == 0 && lane == 0) {
        mbar_arrive_expect_tx(bar_z, NSLAB * TC_SLAB_BYTES);
        for (int s = 0; s < NSLAB; ++s)
            tma_load_2d(z_hi + s * TC_SLAB_BYTES, &tm_z, bar_z, s * TC_SLAB_FLOATS, row_tile * TC_ROWS);
        const int first = n_loads < NSTAGE ? n_loads : NSTAGE;
        for (int L = 0; L < first; ++L) issue_e_load(L);
    }

    // ---- z tile: |z_n|^2 (epilogue threads, thread = row) then in-place hi/lo split (all threads) ---
    mbar_wait(bar_z, 0);
    float a_n = 0.0f;
    if (warp >= 4) {
        const int r = threadIdx.x - 128;
        const uint8_t* rowp = z_hi + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int s = 0; s < NSLAB; ++s) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {   // logical 16-byte chunk c lives at physical chunk c ^ (r & 7)
                const float4 v = *reinterpret_cast<const float4*>(rowp + s * TC_SLAB_BYTES + ((c ^ (r & 7)) << 4));
                a_n = fmaf(v.x, v.x, a_n);
                a_n = fmaf(v.y, v.y, a_n);
                a_n = fmaf(v.z, v.z, a_n);
                a_n = fmaf(v.w, v.w, a_n);
            }
        }
    }
    __syncthreads();
    {
        float4* hi4 = reinterpret_cast<float4*>(z_hi);
        float4* lo4 = reinterpret_cast<float4*>(z_lo);
        constexpr int NV = NSLAB * TC_SLAB_BYTES / 16;
        for (int i = threadIdx.x; i < NV; i += TC_THREADS) {
            const float4 v = hi4[i];
            float4 h, l;
            h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
            l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
            hi4[i] = h;
            lo4[i] = l;
        }
    }
    fence_proxy_async_smem();
    __syncthreads();

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int L = NSTAGE; L < n_loads; ++L) {
                const int stage = L % NSTAGE;
                mbar_wait(bar_empty + stage, ((L / NSTAGE) & 1) ^ 1);
                issue_e_load(L);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(TC_ROWS, TC_CODES);
            const uint32_t zhi_addr = smem_u32(z_hi), zlo_addr = smem_u32(z_lo);
            int L = 0;
            for (int ct = 0; ct < n_ctiles; ++ct) {
                const int buf = ct & 1;
                mbar_wait(bar_acc_empty + buf, ((ct >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * TC_CODES;
                uint32_t accumulate = 0;
                for (int i = 0; i < 2 * NSLAB; ++i, ++L) {
                    const int stage = L % NSTAGE;
                    mbar_wait(bar_full + stage, (L / NSTAGE) & 1);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(stages + stage * TC_SLAB_BYTES);
                    if (i < NSLAB) {   // z_hi . E_lo
                        const uint32_t a_addr = zhi_addr + i * TC_SLAB_BYTES;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            tc_mma_tf32(d_tmem, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32),
                                        idesc, accumulate);
                            accumulate = 1;
                        }
                    } else {           // z_lo . E_hi, then z_hi . E_hi
                        const int s = i - NSLAB;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            tc_mma_tf32(d_tmem, umma_desc_sw128(zlo_addr + s * TC_SLAB_BYTES + kk * 32),
                                        umma_desc_sw128(b_addr + kk * 32), idesc, 1);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            tc_mma_tf32(d_tmem, umma_desc_sw128(zhi_addr + s * TC_SLAB_BYTES + kk * 32),
                                        umma_desc_sw128(b_addr + kk * 32), idesc, 1);
                    }
                    tc_commit(bar_empty + stage);     // stage reusable once these MMAs have read it
                }
                tc_commit(bar_acc_full + buf);        // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread = row (TMEM lane) =====
        const int t = threadIdx.x - 128;              // 0..127
        const uint32_t lane_base = static_cast<uint32_t>((warp - 4) * 32) << 16;
        float best = INFINITY;
        int best_k = k_begin;
        for (int ct = 0; ct < n_ctiles; ++ct) {
            const int buf = ct & 1;
            const int k0 = k_begin + ct * TC_CODES;
            b_tile[buf * TC_CODES + t] = __ldg(e_norm2 + k0 + t);
            named_bar_sync(1, 128);
            mbar_wait(bar_acc_full + buf, (ct >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < TC_CODES / 32; ++cc) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + lane_base + buf * TC_CODES + cc * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float b = b_tile[buf * TC_CODES + cc * 32 + j];
                    const float tsum = a_n + b;                                  // fl(|z|^2 + |E|^2)
                    const float dist = fmaf(-2.0f, __uint_as_float(v[j]), tsum); // fl(tsum - 2c), 2c exact
                    if (dist < best) {
                        best = dist;
                        best_k = k0 + cc * 32 + j;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_acc_empty + buf);
        }
        const long long r = static_cast<long long>(row_tile) * TC_ROWS + t;
        if (r < N) {
            if (keys != nullptr)
                atomicMin(keys + r, pack_key(best, best_k));
            else
                idx_out[r] = best_k;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}


// =========================================================================================================
// Persistent CTA-pair kernel (cta_group::2): the production tensor path.
//
// Why a pair: with both operands in shared memory one M=128,N=128,K=8 TF32 MMA reads 8 KB in 64 cycles =
// 128 B/clk, i.e. ALL of one SM's shared-memory bandwidth, before the TMA writes that refill the ring
// (measured: tensor pipe 38 % active).  In a pair, one UMMA is M=256 (128 rows of z from each CTA) x N=256
// (128 codes of E from each CTA): per SM 8 KB per 128 cycles, and every E tile leaves L2 once per 256 rows.
// Why persistent: at K=1024, D=64 a row tile is only ~6 us of MMA work; launching a CTA per tile paid
// ~9 us of prologue/epilogue around it.  Here 74 pairs loop over work items (row-tile pair x codebook split)
// and every stage of the pipeline runs ahead across items.
//
//   cluster (2,1,1): CTA r of a pair owns rows [(2*rtp + r)*128, +128) of the item and, per code tile of
//   256, stages the codes [k0 + r*128, +128).
//   warp 0        E producer: TMA ring of NSTAGE 16 KB slabs (E_lo slabs, then E_hi slabs, per code tile)
//   warp 1        MMA issuer (leader CTA only): tcgen05.mma.cta_group::2.kind::tf32, M=256 N=256 K=8
//   warps 2-3     z pipeline: TMA the raw z tile, |z_n|^2 chain, in-place tf32 hi/lo split, publish
//   warps 4-11    epilogue: warp w reads TMEM lanes 32*(w%4).. and columns [128*(w>=8), +128) with
//                 double-buffered tcgen05.ld; distance in the reference's order; 4 running minima per thread
//   barriers      z_full/z_free/a_ready (per CTA), z_ready (leader, both CTAs' converters arrive),
//                 full (leader; both CTAs' TMA complete_tx), empty / acc_full (per CTA, multicast commit),
//                 acc_empty (leader; one arrival per epilogue warp of both CTAs)
//   TMEM          2 x 256 columns (double-buffered accumulator), cta_group::2 allocation by both CTAs
// =========================================================================================================
constexpr int TC2_CODES = 256;     // codes per accumulator tile (UMMA N); each CTA stages 128 of them
constexpr int TC2_TMEM_COLS = 512;
constexpr int TC2_THREADS = 384;   // warps 0-3 control, 4-11 epilogue
constexpr int TC2_THREADS_FUSED = 512;   // + warps 12-15: row writers (fused forward)
constexpr int TC2_ARING = 4;       // ring of |z_n|^2 vectors (one per in-flight item)
constexpr int TC2_ZERO_BYTES = 4096;

__host__ __device__ constexpr int tc2_smem_bytes(int nslab, int nstage, int zbuf) {
    return zbuf * 2 * nslab * TC_SLAB_BYTES + nstage * TC_SLAB_BYTES + 2 * TC2_CODES * 4 /* b tile */ +
           TC2_ARING * TC_ROWS * 4 /* a ring */ + 2 * TC_ROWS * 8 /* merge */ + 2 * TC_ROWS * 4 /* idx handoff */ +
           TC2_ZERO_BYTES /* zero row: source of the one-hot bulk stores */ + 512 /* barriers + scratch */ + 1024 /* align */;
}

// in-place tf32 hi/lo split of `nv` float4 (hi stays, lo goes to the twin buffer), strided over `nthreads`
__device__ __forceinline__ void split_tf32_inplace(float4* hi4, float4* lo4, int nv, int first, int nthreads) {
#pragma unroll 4
    for (int i = first; i < nv; i += nthreads) {
        const float4 v = hi4[i];
        float4 h, l;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
        hi4[i] = h;
        lo4[i] = l;
    }
}

// Arguments of the fused row epilogue (vector_quantizer.py:39-56), used when FUSE.
struct FusedRowArgs {
    const float* z;            // (N, D) exact fp32 rows (re-read from L2; the smem copy is split into tf32 halves)
    const float* E;            // (K, D)
    float* q_out;              // (N, D)
    float* onehot;             // (N, K) or nullptr
    float* hist;               // (K), zeroed before the launch
    double* partials;          // one SSE partial per CTA
    unsigned int* counter;     // last-CTA-done counter, zeroed before the launch
    float* sse_out;
    float* loss;
    float* perplexity;
    float beta;
    int finalize;              // 0 under VQ_FLAG_DEFER_STATS
    int onehot_evict_first;    // 1: one-hot stores carry an L2 evict_first policy (tuning knob)
    int rows_later;            // screen kernel: indices only -- usage counts, q_out, SSE, loss / perplexity are left to
                               // quantize_rows_kernel behind it (few code tiles per item: the row workers would set the pace)
    int4* spill;               // screen kernel: [grid][4][SC_SPILL] {row, score, code | chain-instance, type} (workspace)
    long long* trace;          // VQ_TRACE builds only: [cta][role 0..7][64] clock64 stamps
};

// The '1's of the dense one-hot (vector_quantizer.py:40).  The filler thread zero-fills whole rows with bulk copies;
// once they have all landed (`zeros_done`) every row's code is re-read from idx_out (written by this CTA, L2-resident)
// and patched in.  The codes of the first items are fetched BEFORE the wait, and later ones a batch ahead, so the
// patch costs stores only -- it sits on the kernel's critical tail.
template <typename Row0Of>
__device__ __forceinline__ void patch_onehot_ones(float* __restrict__ onehot, const int* __restrict__ idx_out, long long N, int K,
                                                  int first_item, int n_items, int item_stride, int wt, int NW, Row0Of row0_of,
                                                  const int* zeros_done) {
    constexpr int IB = 4;
    int codes[IB][2];
    auto load = [&](int w0) {
#pragma unroll
        for (int u = 0; u < IB; ++u) {
            const int w = w0 + u * item_stride;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = wt + h * NW;
                const long long gr = row0_of(w) + r;
                codes[u][h] = (w < n_items && r < TC_ROWS && gr < N) ? __ldcg(idx_out + gr) : -1;
            }
        }
    };
    load(first_item);
    while (ld_acquire_cta(zeros_done) == 0) {
    }
    for (int w0 = first_item; w0 < n_items; w0 += IB * item_stride) {
#pragma unroll
        for (int u = 0; u < IB; ++u) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (codes[u][h] >= 0) {
                    const long long gr = row0_of(w0 + u * item_stride) + (wt + h * NW);
                    __stcs(onehot + gr * K + codes[u][h], 1.0f);
                }
            }
        }
        if (w0 + IB * item_stride < n_items) load(w0 + IB * item_stride);
    }
}

// Per-CTA SSE partial, then the last CTA to finish reduces all partials in a fixed order and writes SSE, loss
// (vector_quantizer.py:46-52) and perplexity (:55-56).  Called by all `nw` row-worker threads of a CTA (named barrier
// `bar_id`); `red` is 5 doubles of shared scratch.  This is the tail of the kernel's critical path: the partials and
// the first 12 histogram entries per thread are fetched in ONE round trip, later batches 12 at a time.
__device__ __forceinline__ void publish_and_finalize(const FusedRowArgs& fr, float sse, long long N, int K, int D, int wt, int nw,
                                                     int lane, int wwarp, int n_warps, double* red, int bar_id) {
    auto red_total = [&]() {
        double t = red[0];
        for (int i = 1; i < n_warps; ++i) t += red[i];
        return t;
    };
    const double sd = warp_sum_d(static_cast<double>(sse));
    if (lane == 0) red[wwarp] = sd;
    named_bar_sync(bar_id, nw);
    volatile int* last_flag = reinterpret_cast<volatile int*>(red + 4);
    if (wt == 0) {
        fr.partials[blockIdx.x] = red_total();
        __threadfence();
        const unsigned int done = atomicAdd(fr.counter, 1u);
        const bool last = done == gridDim.x - 1;
        *last_flag = last ? 1 : 0;
        if (last) {
            // every CTA has published (fence + atomic above): indices, q_out, one-hot and the usage counts are complete.
            // Raise the workspace's ready word NOW: a backward launched behind this kernel (vq_step_backward) starts on it and
            // overlaps the serial statistics tail below (loss / perplexity, ~5 us), which it does not need.
            __threadfence();
            st_release_gpu_u32(fr.counter + 2, 1u);
        }
    }
    named_bar_sync(bar_id, nw);
    if (!*last_flag) return;
    __threadfence();
    constexpr int HB = 12;
    const int g = static_cast<int>(gridDim.x);
    const bool fin = fr.finalize != 0;
    const double p0 = wt < g ? __ldcg(fr.partials + wt) : 0.0;
    const double p1 = wt + nw < g ? __ldcg(fr.partials + wt + nw) : 0.0;
    float h[HB];
#pragma unroll
    for (int u = 0; u < HB; ++u) h[u] = (fin && wt + u * nw < K) ? __ldcg(fr.hist + wt + u * nw) : 0.0f;
    double t = p0 + p1;
    for (int i = wt + 2 * nw; i < g; i += nw) t += __ldcg(fr.partials + i);
    t = warp_sum_d(t);
    named_bar_sync(bar_id, nw);
    if (lane == 0) red[wwarp] = t;
    named_bar_sync(bar_id, nw);
    const double total = red_total();
    if (wt == 0) {
        *fr.sse_out = static_cast<float>(total);
        *fr.counter = 0u;
    }
    if (!fin) return;
    if (wt == 0 && fr.q_out != nullptr) {
        const float m = static_cast<float>(total / (static_cast<double>(N) * static_cast<double>(D)));
        *fr.loss = __fadd_rn(m, __fmul_rn(fr.beta, m));                      // vector_quantizer.py:52
    }
    double ent = 0.0;
    const float nf = static_cast<float>(N);
    for (int kb = wt; kb < K; kb += HB * nw) {
        if (kb != wt) {
#pragma unroll
            for (int u = 0; u < HB; ++u) h[u] = kb + u * nw < K ? __ldcg(fr.hist + kb + u * nw) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < HB; ++u) {
            if (kb + u * nw < K) {
                const float p = __fdiv_rn(h[u], nf);                         // vector_quantizer.py:55
                ent += static_cast<double>(p * logf(p + 1e-10f));            // :56
            }
        }
    }
    ent = warp_sum_d(ent);
    named_bar_sync(bar_id, nw);
    if (lane == 0) red[wwarp] = ent;
    named_bar_sync(bar_id, nw);
    if (wt == 0) *fr.perplexity = expf(static_cast<float>(-red_total()));
}

#ifdef VQ_TRACE
#define VQ_TR(role, slot)                                                                              \
    do {                                                                                               \
        if (fr.trace != nullptr && (slot) < 64) fr.trace[(blockIdx.x * 8 + (role)) * 64 + (slot)] = clock64(); \
    } while (0)
#else
#define VQ_TR(role, slot) \
    do {                  \
    } while (0)
#endif

template <int NSLAB, int NSTAGE, int ZBUF, bool FUSE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FUSE ? TC2_THREADS_FUSED : TC2_THREADS, 1)
argmin_tc2_kernel(const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_ehi,
                  const __grid_constant__ CUtensorMap tm_elo, const float* __restrict__ e_norm2, long long N, int K,
                  int codes_per_split, int splits, int n_items, int* __restrict__ idx_out,
                  unsigned long long* __restrict__ keys, float* __restrict__ hist_to_zero,
                  unsigned int* __restrict__ counter_to_zero, const FusedRowArgs fr) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (128B-swizzle atoms) by OFFSET, so the compiler still knows this is shared memory
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int NTHREADS = FUSE ? TC2_THREADS_FUSED : TC2_THREADS;
    constexpr int ZBYTES = 2 * NSLAB * TC_SLAB_BYTES;          // one z buffer: hi then lo
    uint8_t* zbufs = smem;
    uint8_t* stages = zbufs + ZBUF * ZBYTES;
    float* b_tile = reinterpret_cast<float*>(stages + NSTAGE * TC_SLAB_BYTES);   // [2][256]
    float* a_ring = b_tile + 2 * TC2_CODES;                                      // [TC2_ARING][128]
    unsigned long long* merge = reinterpret_cast<unsigned long long*>(a_ring + TC2_ARING * TC_ROWS);   // [2][128]
    int* s_idx = reinterpret_cast<int*>(merge + 2 * TC_ROWS);                    // [2][128] epilogue -> writers
    float* zero_row = reinterpret_cast<float*>(s_idx + 2 * TC_ROWS);             // 4 KB of zeros (bulk-store source)
    uint64_t* bars = reinterpret_cast<uint64_t*>(zero_row + TC2_ZERO_BYTES / 4);
    uint64_t* bar_z_full = bars;                        // [ZBUF]
    uint64_t* bar_z_free = bar_z_full + ZBUF;           // [ZBUF]
    uint64_t* bar_z_ready = bar_z_free + ZBUF;          // [ZBUF]  (leader side)
    uint64_t* bar_a_ready = bar_z_ready + ZBUF;         // [TC2_ARING]
    uint64_t* bar_full = bar_a_ready + TC2_ARING;       // [NSTAGE] (leader side)
    uint64_t* bar_empty = bar_full + NSTAGE;            // [NSTAGE]
    uint64_t* bar_acc_full = bar_empty + NSTAGE;        // [2]
    uint64_t* bar_acc_empty = bar_acc_full + 2;         // [2]      (leader side)
    uint64_t* bar_idx_ready = bar_acc_empty + 2;        // [2]
    uint64_t* bar_idx_free = bar_idx_ready + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_idx_free + 2);
    double* red = reinterpret_cast<double*>(tmem_slot + 2);   // [4] + flag
    int* zeros_done = reinterpret_cast<int*>(red + 5);        // 1 once every one-hot row of this CTA is zero-filled

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_ctiles = codes_per_split / TC2_CODES;

    if (!FUSE && blockIdx.x == 0) {   // fused: hist/counter are zeroed before the launch (they are written here)
        for (int k = threadIdx.x; k < K; k += NTHREADS) hist_to_zero[k] = 0.0f;
        if (threadIdx.x == 0) *counter_to_zero = 0u;
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_z);
        tma_prefetch_desc(&tm_ehi);
        tma_prefetch_desc(&tm_elo);
        for (int i = 0; i < ZBUF; ++i) {
            mbar_init(bar_z_full + i, 1);
            mbar_init(bar_z_free + i, 1);
            mbar_init(bar_z_ready + i, 128);      // 64 converter threads of each CTA
        }
        for (int i = 0; i < TC2_ARING; ++i) mbar_init(bar_a_ready + i, 64);
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_acc_full + b, 1);
            mbar_init(bar_acc_empty + b, 16);     // 8 epilogue warps of each CTA
            mbar_init(bar_idx_ready + b, 4);      // 4 epilogue warps (column half 0) publish the item's indices
            mbar_init(bar_idx_free + b, (FUSE && fr.onehot != nullptr) ? 3 : 4);   // the row-worker warps release them
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_2sm(tmem_slot, TC2_TMEM_COLS);
        tmem_relinquish_2sm();
    }
    if (FUSE && warp >= 12) {
        for (int i = threadIdx.x - 384; i < TC2_ZERO_BYTES / 16; i += 128)
            reinterpret_cast<float4*>(zero_row)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (threadIdx.x == 384) *zeros_done = 0;
        fence_proxy_async_smem();     // the bulk copies read the zero row through the async proxy
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait_prior_grids();      // everything above overlapped the previous kernel (PDL); its outputs are visible from here
    if (threadIdx.x == 0) {
        VQ_TR(7, 0);
#ifdef VQ_TRACE
        if (fr.trace != nullptr) fr.trace[(blockIdx.x * 8 + 7) * 64 + 3] = static_cast<long long>(global_timer_ns());
#endif
    }

    if (warp == 0) {
        // ===== E producer (both CTAs: each loads its 128-code half of every tile) =====
        if (lane == 0) {
            int L = 0;
            for (int w = pair; w < n_items; w += n_pairs) {
                const int k_begin = (w % splits) * codes_per_split;
                for (int ct = 0; ct < n_ctiles; ++ct) {
                    for (int i = 0; i < 2 * NSLAB; ++i, ++L) {
                        const int stage = L % NSTAGE;
                        mbar_wait(bar_empty + stage, ((L / NSTAGE) & 1) ^ 1);
                        const bool lo = i < NSLAB;
                        const int slab = lo ? i : i - NSLAB;
                        if (leader) mbar_arrive_expect_tx(bar_full + stage, 2 * TC_SLAB_BYTES);
                        tma_load_2d_2sm(stages + stage * TC_SLAB_BYTES, lo ? &tm_elo : &tm_ehi, bar_full + stage,
                                        slab * TC_SLAB_FLOATS, k_begin + ct * TC2_CODES + static_cast<int>(cta_rank) * TC_ROWS);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(2 * TC_ROWS, TC2_CODES);
            int L = 0, ctg = 0, it = 0;
            for (int w = pair; w < n_items; w += n_pairs, ++it) {
                const int zb = it % ZBUF;
                VQ_TR(1, 3 * it);
                mbar_wait(bar_z_ready + zb, (it / ZBUF) & 1);
                VQ_TR(1, 3 * it + 1);
                tc_fence_after();
                const uint32_t zhi_addr = smem_u32(zbufs + zb * ZBYTES);
                const uint32_t zlo_addr = zhi_addr + NSLAB * TC_SLAB_BYTES;
                for (int ct = 0; ct < n_ctiles; ++ct, ++ctg) {
                    const int buf = ctg & 1;
#ifdef VQ_TRACE
                    long long tw0 = clock64();
#endif
                    mbar_wait(bar_acc_empty + buf, ((ctg >> 1) & 1) ^ 1);
#ifdef VQ_TRACE
                    if (fr.trace != nullptr) fr.trace[(blockIdx.x * 8 + 6) * 64 + 32] += clock64() - tw0;   // acc_empty wait cycles
#endif
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * TC2_CODES;
                    uint32_t accumulate = 0;
                    for (int i = 0; i < 2 * NSLAB; ++i, ++L) {
                        const int stage = L % NSTAGE;
#ifdef VQ_TRACE
                        tw0 = clock64();
#endif
                        mbar_wait(bar_full + stage, (L / NSTAGE) & 1);
#ifdef VQ_TRACE
                        if (fr.trace != nullptr) fr.trace[(blockIdx.x * 8 + 6) * 64 + 33] += clock64() - tw0;   // full (TMA) wait cycles
#endif
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(stages + stage * TC_SLAB_BYTES);
                        if (i < NSLAB) {   // z_hi . E_lo
                            const uint32_t a_addr = zhi_addr + i * TC_SLAB_BYTES;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                tc_mma_tf32_2sm(d_tmem, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32),
                                                idesc, accumulate);
                                accumulate = 1;
                            }
                        } else {           // z_lo . E_hi, then z_hi . E_hi
                            const int s = i - NSLAB;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                tc_mma_tf32_2sm(d_tmem, umma_desc_sw128(zlo_addr + s * TC_SLAB_BYTES + kk * 32),
                                                umma_desc_sw128(b_addr + kk * 32), idesc, 1);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                tc_mma_tf32_2sm(d_tmem, umma_desc_sw128(zhi_addr + s * TC_SLAB_BYTES + kk * 32),
                                                umma_desc_sw128(b_addr + kk * 32), idesc, 1);
                        }
                        tc_commit_2sm(bar_empty + stage);
                    }
                    tc_commit_2sm(bar_acc_full + buf);
                }
                tc_commit_2sm(bar_z_free + zb);    // every MMA that reads this z buffer has completed
                VQ_TR(1, 3 * it + 2);
            }
        }
    } else if (warp < 4) {
        // ===== z pipeline (64 threads per CTA): TMA, |z_n|^2, hi/lo split, publish =====
        const int c = threadIdx.x - 64;            // 0..63
        int it = 0;
        for (int w = pair; w < n_items; w += n_pairs, ++it) {
            const int zb = it % ZBUF;
            const int row_tile = 2 * (w / splits) + static_cast<int>(cta_rank);
            uint8_t* z_hi = zbufs + zb * ZBYTES;
            uint8_t* z_lo = z_hi + NSLAB * TC_SLAB_BYTES;
            if (c == 0) {
                mbar_wait(bar_z_free + zb, ((it / ZBUF) & 1) ^ 1);
                VQ_TR(2, 3 * it);
                mbar_arrive_expect_tx(bar_z_full + zb, NSLAB * TC_SLAB_BYTES);
                for (int s = 0; s < NSLAB; ++s)
                    tma_load_2d(z_hi + s * TC_SLAB_BYTES, &tm_z, bar_z_full + zb, s * TC_SLAB_FLOATS, row_tile * TC_ROWS);
            }
            mbar_wait(bar_z_full + zb, (it / ZBUF) & 1);
            if (c == 0) VQ_TR(2, 3 * it + 1);
            // |z_n|^2: thread c owns rows c and c+64 (sequential FMA chains over d)
            float* a_dst = a_ring + (it % TC2_ARING) * TC_ROWS;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = c + h * 64;
                const uint8_t* rowp = z_hi + (r >> 3) * 1024 + (r & 7) * 128;
                float a = 0.0f;
#pragma unroll
                for (int s = 0; s < NSLAB; ++s) {
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {   // logical 16-byte chunk ch lives at physical chunk ch ^ (r & 7)
                        const float4 v = *reinterpret_cast<const float4*>(rowp + s * TC_SLAB_BYTES + ((ch ^ (r & 7)) << 4));
                        a = fmaf(v.x, v.x, a);
                        a = fmaf(v.y, v.y, a);
                        a = fmaf(v.z, v.z, a);
                        a = fmaf(v.w, v.w, a);
                    }
                }
                a_dst[r] = a;
            }
            float4* hi4 = reinterpret_cast<float4*>(z_hi);
            float4* lo4 = reinterpret_cast<float4*>(z_lo);
            constexpr int NV = NSLAB * TC_SLAB_BYTES / 16;
            if (it == 0) {
                // first item: nothing else is running yet, so the 256 epilogue threads split the tile with us
                named_bar_sync(5, 320);             // all |z|^2 reads of the raw tile are done before it is overwritten
                split_tf32_inplace(hi4, lo4, NV, c, 320);
                fence_proxy_async_smem();
                named_bar_sync(5, 320);             // ... and everybody's share is written (and proxy-fenced)
            } else {
                named_bar_sync(3, 64);
                split_tf32_inplace(hi4, lo4, NV, c, 64);
            }
            fence_proxy_async_smem();               // generic-proxy writes -> visible to the MMA's async-proxy reads
            mbar_arrive(bar_a_ready + (it % TC2_ARING));
            mbar_arrive_cluster(bar_z_ready + zb, 0);
            if (c == 0) VQ_TR(2, 3 * it + 2);
        }
        // drain: the leader's last multicast commits on z_free must have landed before this CTA may exit
        if (c == 0) {
            for (int j = it > ZBUF ? it - ZBUF : 0; j < it; ++j) mbar_wait(bar_z_free + (j % ZBUF), (j / ZBUF) & 1);
        }
    } else if (warp < 12) {
        // ===== epilogue (8 warps per CTA): thread = (row, column half) =====
        const int ew = warp - 4;                    // 0..7
        const int half = ew >> 2;                   // columns [128*half, +128) of every code tile
        const int row = (ew & 3) * 32 + lane;       // TMEM lane == row of the CTA's 128-row tile
        const int et = threadIdx.x - 128;           // 0..255
        const uint32_t lane_base = static_cast<uint32_t>((ew & 3) * 32) << 16;
        int it = 0, ctg = 0;
        {   // help the z pipeline with the first tile (see there)
            mbar_wait(bar_z_full + 0, 0);
            named_bar_sync(5, 320);
            split_tf32_inplace(reinterpret_cast<float4*>(zbufs), reinterpret_cast<float4*>(zbufs + NSLAB * TC_SLAB_BYTES),
                               NSLAB * TC_SLAB_BYTES / 16, 64 + et, 320);
            fence_proxy_async_smem();
            named_bar_sync(5, 320);
        }
        for (int w = pair; w < n_items; w += n_pairs, ++it) {
            const int k_begin = (w % splits) * codes_per_split;
            const int row_tile = 2 * (w / splits) + static_cast<int>(cta_rank);
            mbar_wait(bar_a_ready + (it % TC2_ARING), (it / TC2_ARING) & 1);
            if (et == 0) VQ_TR(3, 4 * it);
            const float a_n = a_ring[(it % TC2_ARING) * TC_ROWS + row];
            float best[4];
            int best_k[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                best[q] = INFINITY;
                best_k[q] = k_begin;
            }
            for (int ct = 0; ct < n_ctiles; ++ct, ++ctg) {
                const int buf = ctg & 1;
                const int k0 = k_begin + ct * TC2_CODES;
                b_tile[buf * TC2_CODES + et] = __ldg(e_norm2 + k0 + et);
                named_bar_sync(1, 256);
                if (et == 0) VQ_TR(0, 3 * ctg);              // role 0 slots: per tile (wait start | acc full | scanned)
                mbar_wait(bar_acc_full + buf, (ctg >> 1) & 1);
                if (et == 0) VQ_TR(0, 3 * ctg + 1);
                if (et == 0 && ct == 0) VQ_TR(3, 4 * it + 1);
                if (et == 0 && ct == n_ctiles - 1) VQ_TR(3, 4 * it + 2);
                tc_fence_after();
                const uint32_t t_addr = tmem_base + lane_base + buf * TC2_CODES + half * 128;
                const float4* b4 = reinterpret_cast<const float4*>(b_tile + buf * TC2_CODES + half * 128);
                uint32_t va[16], vb[16];
                auto consume = [&](const uint32_t (&v)[16], int cc) {   // 16 columns starting at cc*16
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 b = b4[cc * 4 + j4];
                        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float tsum = a_n + bb[q];                                         // fl(|z|^2 + |E|^2)
                            const float dist = fmaf(-2.0f, __uint_as_float(v[j4 * 4 + q]), tsum);   // fl(tsum - 2c)
                            if (dist < best[q]) {
                                best[q] = dist;
                                best_k[q] = k0 + half * 128 + cc * 16 + j4 * 4 + q;
                            }
                        }
                    }
                };
                tmem_ld_32x32b_x16(t_addr, va);
#pragma unroll
                for (int cc = 0; cc < 8; cc += 2) {     // double-buffered: the next 16 columns load while these are scanned
                    tmem_ld_wait();
                    tmem_ld_32x32b_x16(t_addr + (cc + 1) * 16, vb);
                    consume(va, cc);
                    tmem_ld_wait();
                    if (cc + 2 < 8) {
                        tmem_ld_32x32b_x16(t_addr + (cc + 2) * 16, va);
                    } else {
                        // this warp's share of the accumulator is in registers: hand the buffer back early
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster_relaxed(bar_acc_empty + buf, 0);
                    }
                    consume(vb, cc + 1);
                }
                if (et == 0) VQ_TR(0, 3 * ctg + 2);
            }
            // merge the 4 chains (smallest distance, then smallest index), then the two column halves
            float bd = best[0];
            int bk = best_k[0];
#pragma unroll
            for (int q = 1; q < 4; ++q) {
                if (best[q] < bd || (best[q] == bd && best_k[q] < bk)) {
                    bd = best[q];
                    bk = best_k[q];
                }
            }
            unsigned long long key = pack_key(bd, bk);
            unsigned long long* mslot = merge + (it & 1) * TC_ROWS;
            if (half == 1) mslot[row] = key;
            if (FUSE && half == 0) mbar_wait(bar_idx_free + (it & 1), ((it >> 1) & 1) ^ 1);   // writers are done with this slot
            named_bar_sync(2, 256);
            if (half == 0) {
                const unsigned long long other = mslot[row];
                key = other < key ? other : key;
                const int code = static_cast<int>(key & 0xffffffffu);
                const long long r = static_cast<long long>(row_tile) * TC_ROWS + row;
                if (r < N) {
                    if (keys != nullptr)
                        atomicMin(keys + r, key);
                    else
                        idx_out[r] = code;
                }
                if (FUSE) {
                    s_idx[(it & 1) * TC_ROWS + row] = code;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_idx_ready + (it & 1));
                    if (et == 0) VQ_TR(3, 4 * it + 3);
                }
            }
        }
    } else if (FUSE && warp == 12 && fr.onehot != nullptr) {
        // ===== one-hot zero filler (1 thread): vector_quantizer.py:39 =====
        // The zeros of the dense one-hot do not depend on the argmin, so they start streaming at the first cycle of
        // the kernel: bulk async copies (TMA engine) from one shared zero row -- no registers, and not subject to a
        // warp's limit on outstanding stores -- in step with HBM rather than with the MMAs.  Up to three items are
        // in flight; `zeros_done` tells the row workers which items may receive their '1's.
        if (lane == 0) {
            const uint32_t row_bytes = static_cast<uint32_t>(K) * 4u;
            const uint32_t chunk = row_bytes < TC2_ZERO_BYTES ? row_bytes : TC2_ZERO_BYTES;   // K % 256 == 0
            const uint64_t pol = l2_policy_evict_first();    // 4K bytes per row of write-once output: keep z / E in L2
            int it = 0;
            for (int w = pair; w < n_items; w += n_pairs, ++it) {
                const int row_tile = 2 * (w / splits) + static_cast<int>(cta_rank);
                const long long row0 = static_cast<long long>(row_tile) * TC_ROWS;
                const long long left = N - row0;
                const int rows_here = left <= 0 ? 0 : (left < TC_ROWS ? static_cast<int>(left) : TC_ROWS);
                VQ_TR(4, 2 * it);
                // Never throttled on completion (that would put the long write-completion latency of a saturated
                // HBM into the loop); the TMA queue's own back-pressure paces this thread.
                uint8_t* obase = reinterpret_cast<uint8_t*>(fr.onehot + row0 * K);
                const size_t total = static_cast<size_t>(rows_here) * row_bytes;
                if (fr.onehot_evict_first) {
                    for (size_t off = 0; off < total; off += chunk) bulk_store_s2g_hint(obase + off, zero_row, chunk, pol);
                } else {
                    for (size_t off = 0; off < total; off += chunk) bulk_store_s2g(obase + off, zero_row, chunk);
                }
                bulk_commit();
                VQ_TR(4, 2 * it + 1);
            }
            bulk_wait_all();                 // every zero row of this CTA is in place ...
            fence_proxy_async_all();
            st_release_cta(zeros_done, 1);   // ... so the workers may now patch in the '1's
            VQ_TR(6, 0);
        }
    } else if (FUSE) {
        // ===== row workers (3 warps, or 4 without a one-hot): vector_quantizer.py:40-56 for the item's 128 rows =====
        constexpr int DV = NSLAB * 8;               // float4 per row
        const bool have_oh = fr.onehot != nullptr;
        const int NW = have_oh ? 96 : 128;          // worker threads
        const int wt = have_oh ? threadIdx.x - 416 : threadIdx.x - 384;
        const int D = NSLAB * TC_SLAB_FLOATS;
        float sse = 0.0f;
        int it = 0;
        for (int w = pair; w < n_items; w += n_pairs, ++it) {
            const int row_tile = 2 * (w / splits) + static_cast<int>(cta_rank);
            const long long row0 = static_cast<long long>(row_tile) * TC_ROWS;
            const long long left = N - row0;
            const int rows_here = left <= 0 ? 0 : (left < TC_ROWS ? static_cast<int>(left) : TC_ROWS);
            mbar_wait(bar_idx_ready + (it & 1), (it >> 1) & 1);
            if (wt == 0) VQ_TR(5, 3 * it);
            const int* sidx = s_idx + (it & 1) * TC_ROWS;
            for (int r = wt; r < rows_here; r += NW) atomicAdd(fr.hist + sidx[r], 1.0f);
            // q_out = fl(z + fl(E[idx] - z)), sse += (E[idx] - z)^2 : a thread's z loads are issued ahead of the gathers
            const float4* z4 = reinterpret_cast<const float4*>(fr.z + row0 * D);
            float4* q4 = reinterpret_cast<float4*>(fr.q_out + row0 * D);
            const int n_el = rows_here * DV;
            constexpr int PER_T = (TC_ROWS * DV + 95) / 96;    // covers NW = 96; with 128 workers the tail is masked off
            constexpr int UB = 8;                   // elements per pass: 8 z loads + 8 gathers in flight per thread
#pragma unroll 1
            for (int ub = 0; ub < PER_T; ub += UB) {
                if (wt + ub * NW >= n_el) break;
                float4 zv[UB], ev[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const int e = wt + (ub + u) * NW;
                    if (e < n_el) {
                        zv[u] = __ldg(z4 + e);
                        ev[u] = __ldg(reinterpret_cast<const float4*>(fr.E + static_cast<size_t>(sidx[e / DV]) * D) + (e % DV));
                    }
                }
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const int e = wt + (ub + u) * NW;
                    if (e < n_el) {
                        const float4 zz = zv[u];
                        float4 df, qv;
                        df.x = ev[u].x - zz.x; df.y = ev[u].y - zz.y; df.z = ev[u].z - zz.z; df.w = ev[u].w - zz.w;
                        qv.x = zz.x + df.x; qv.y = zz.y + df.y; qv.z = zz.z + df.z; qv.w = zz.w + df.w;
                        __stcs(q4 + e, qv);
                        sse = fmaf(df.x, df.x, sse); sse = fmaf(df.y, df.y, sse);
                        sse = fmaf(df.z, df.z, sse); sse = fmaf(df.w, df.w, sse);
                    }
                }
            }
            if (wt == 0) VQ_TR(5, 3 * it + 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_idx_free + (it & 1));
            if (wt == 0) VQ_TR(5, 3 * it + 2);
        }
        if (have_oh) {
            // the '1's of the one-hot (vector_quantizer.py:40): once all zero rows of this CTA have landed, every
            // row's code is re-read from idx_out (written by this CTA's epilogue, L2-resident) and patched in
            patch_onehot_ones(fr.onehot, idx_out, N, K, pair, n_items, n_pairs, wt, NW,
                              [&](int w) { return static_cast<long long>(2 * (w / splits) + static_cast<int>(cta_rank)) * TC_ROWS; },
                              zeros_done);
            if (wt == 0) VQ_TR(5, 60);
        }
        // per-CTA SSE partial, then last-CTA-done reduction in a fixed order (+ loss / perplexity)
        publish_and_finalize(fr, sse, N, K, D, wt, NW, lane, have_oh ? warp - 13 : warp - 12, have_oh ? 3 : 4, red, 4);
    }

    if (threadIdx.x == 0) VQ_TR(7, 1);
    pdl_launch_dependents();
    tc_fence_before();
    cluster_sync_all();      // no CTA leaves (or frees TMEM) while its peer may still signal it
    if (threadIdx.x == 0) {
        VQ_TR(7, 2);
#ifdef VQ_TRACE
        if (fr.trace != nullptr) fr.trace[(blockIdx.x * 8 + 7) * 64 + 4] = static_cast<long long>(global_timer_ns());
#endif
    }
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, TC2_TMEM_COLS);
    }
}

}  // namespace b200vq
