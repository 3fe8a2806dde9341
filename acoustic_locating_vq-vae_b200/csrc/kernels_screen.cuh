// kernels_screen.cuh -- fused forward, "screen + refine" variant (sm_100a: TMA + tcgen05 + TMEM).
//
// The 3xTF32 kernel (kernels_tc.cuh) rebuilds an fp32-grade product from three tensor-core passes.  This kernel
// spends ONE pass and restores exactness where it matters:
//
//   screen   s(n,k) = |E_k|^2 - 2 * <tf32(z_n), tf32(E_k)>          one tcgen05 TF32 pass, fp32 accumulate
//            |s(n,k) + |z_n|^2 - dist(n,k)| <= eta_n                 eta_n = 2*eps_n + rounding,
//                                                                    eps_n = (1.02 * 2^-10 + 2 D 2^-24) |z_n| max_k|E_k|
//            (each operand carries <= 2^-11 relative rounding error; Cauchy-Schwarz over the D products; the second
//            term bounds the fp32 accumulation error of the tensor core and of the oracle's own fmaf chain)
//   collect  every code with s(n,k) <= min_k s(n,k) + 2*eta_n is a CANDIDATE; the oracle's argmin -- including all
//            of its exact ties -- is provably among them.  ~93 % of rows have exactly one candidate.
//   refine   the row workers evaluate the candidates of the remaining rows in EXACT fp32 in the oracle's order,
//            dist = fmaf(-2, fma-chain <z_n, E_k>, fl(|z_n|^2 + |E_k|^2)), first index on ties
//            (vector_quantizer.py:34-38; oracle/vq_oracle.c).  Indices are therefore bit-identical to the C oracle
//            for any input, not merely equal up to fp32 near-ties.
//
// Candidate bookkeeping in the epilogue (thread = (row, column half)): per 256-code tile four running chains
// (columns j mod 4) keep (best, index, second-best value).  At the end of the tile every chain minimum within the
// margin of the running minimum is APPENDED to the thread's 4-entry candidate log in shared memory (score, code) --
// the running minimum only decreases, so everything that can matter at the end is logged; a new minimum that beats
// the old one by more than the margin empties the log, and a full log is first compacted against the running minimum.
// A chain's second best within the margin is remembered as `lost` together with the 32 columns it hides in.  The
// logs of the two halves are filtered against the final minimum once per item.  Rows whose `lost` ends up within the
// margin of the final minimum rescan those 32 columns, or -- when that is not enough (~0.01 % of rows) -- the whole
// codebook, exactly.
//
// Structure (persistent CTA pairs, cta_group::2, M=256 N=256 K=8) and the fused row epilogue are those of
// argmin_tc2_kernel<..., FUSE=true>; only E_hi is streamed and only tf32(z) is kept in shared memory, which is
// what makes D = 256 fit.
#pragma once
#include "common.cuh"
#include "kernels_tc.cuh"

namespace b200vq {

constexpr int SC_THREADS = 512;
constexpr int SC_NC_OVERFLOW = 255;   // status: candidate set could not be bounded -> exact scan of the whole codebook
constexpr int SC_PMAX = 384;          // (row, code) pairs evaluated exactly per 128-row item
constexpr int SC_LOG = 4;             // candidate log entries per epilogue thread (compacted when full)
constexpr int SC_SPILL = 128;         // per CTA and in-flight item: candidates that fit neither the log nor the handoff (global memory)

__host__ __device__ constexpr int sc_smem_bytes(int nslab, int nstage, int zbuf) {
    return zbuf * nslab * TC_SLAB_BYTES + nstage * TC_SLAB_BYTES + 2 * TC2_CODES * 4 /* b tile */ +
           TC2_ARING * TC_ROWS * 4 /* a ring */ + TC_ROWS * 16 /* half merge */ + SC_LOG * 256 * 8 /* candidate logs */ +
           2 * TC_ROWS * 9 * 4 /* handoff: 4 candidates, status, a_n, threshold, 2 rescan locations */ +
           2 * TC_ROWS * 8 /* row keys, per worker group */ + 2 * TC_ROWS * 4 /* final idx */ +
           2 * TC_ROWS * 4 /* full-rescan lists */ + 2 * TC_ROWS * 4 /* row states */ + SC_PMAX * 4 /* pair lists */ +
           TC2_ZERO_BYTES + 512 /* barriers + scratch */ + 1024 /* align */;
}

// in-place tf32 rounding of `nv` float4, strided over `nthreads`
__device__ __forceinline__ void round_tf32_inplace(float4* p4, int nv, int first, int nthreads) {
#pragma unroll 4
    for (int i = first; i < nv; i += nthreads) {
        float4 v = p4[i];
        v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w);
        p4[i] = v;
    }
}

// exact fp32 dot product in the oracle's order: one fmaf chain over d = 0..D-1
template <int D>
__device__ __forceinline__ float dot_chain_exact(const float* __restrict__ zr, const float* __restrict__ er) {
    float acc = 0.0f;
    const float4* z4 = reinterpret_cast<const float4*>(zr);
    const float4* e4 = reinterpret_cast<const float4*>(er);
    constexpr int B = 8;                                   // 8 + 8 independent 16-byte loads in flight per batch
#pragma unroll 1
    for (int i0 = 0; i0 < D / 4; i0 += B) {
        float4 a[B], b[B];
#pragma unroll
        for (int u = 0; u < B; ++u) {
            a[u] = __ldg(z4 + i0 + u);
            b[u] = __ldg(e4 + i0 + u);
        }
#pragma unroll
        for (int u = 0; u < B; ++u) {
            acc = fmaf(a[u].x, b[u].x, acc);
            acc = fmaf(a[u].y, b[u].y, acc);
            acc = fmaf(a[u].z, b[u].z, acc);
            acc = fmaf(a[u].w, b[u].w, acc);
        }
    }
    return acc;
}

template <int NSLAB, int NSTAGE, int ZBUF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SC_THREADS, 1)
vq_screen_kernel(const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_ehi,
                 const float* __restrict__ e_norm2, long long N, int K, int n_items, int* __restrict__ idx_out,
                 const FusedRowArgs fr) {
    extern __shared__ uint8_t smem_raw[];
#ifdef VQ_TRACE
    if (threadIdx.x == 0 && fr.trace != nullptr) fr.trace[(blockIdx.x * 8 + 7) * 64 + 5] = static_cast<long long>(global_timer_ns());
#endif
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int D = NSLAB * TC_SLAB_FLOATS;
    constexpr int ZBYTES = NSLAB * TC_SLAB_BYTES;              // one z buffer: tf32(z) only
    uint8_t* zbufs = smem;
    uint8_t* stages = zbufs + ZBUF * ZBYTES;
    float* b_tile = reinterpret_cast<float*>(stages + NSTAGE * TC_SLAB_BYTES);   // [2][256]
    float* a_ring = b_tile + 2 * TC2_CODES;                                      // [TC2_ARING][128]
    float4* mrg = reinterpret_cast<float4*>(a_ring + TC2_ARING * TC_ROWS);       // [128] half 1 -> half 0: min, n, lost, loc
    float2* clog = reinterpret_cast<float2*>(mrg + TC_ROWS);                     // [SC_LOG][256] candidate logs: score, code
    int* h_cand = reinterpret_cast<int*>(clog + SC_LOG * 256);                   // [2][128][4]
    int* h_nc = h_cand + 2 * TC_ROWS * 4;                                        // [2][128] status: nc | nloc << 8, or 255
    float* h_an = reinterpret_cast<float*>(h_nc + 2 * TC_ROWS);                  // [2][128]
    float* h_thr = h_an + 2 * TC_ROWS;                                           // [2][128] final minimum + margin
    int* h_loc = reinterpret_cast<int*>(h_thr + 2 * TC_ROWS);                    // [2][128][2] chain-instances to rescan
    // worker-group state ([2]: the row workers run as two independent groups when no one-hot is emitted)
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(h_loc + 2 * TC_ROWS * 2);   // [2][128] (distance, code) minima
    int* s_idx = reinterpret_cast<int*>(s_key + 2 * TC_ROWS);                    // [2][128] final codes of the group's item
    int* s_ovf = s_idx + 2 * TC_ROWS;                                            // [2][128] rows that need the full rescan
    int* s_rng = s_ovf + 2 * TC_ROWS;                                            // [2][128] row state: 0 decided, 1 has pairs, 2 full rescan
    int* pair_rc = s_rng + 2 * TC_ROWS;                                          // [SC_PMAX] row << 20 | code (split between the groups)
    float* zero_row = reinterpret_cast<float*>(pair_rc + SC_PMAX);
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
    int* zeros_done = reinterpret_cast<int*>(red + 5);
    unsigned long long* ovf_key = reinterpret_cast<unsigned long long*>(zeros_done + 2);   // [2]
    float* s_bmax = reinterpret_cast<float*>(ovf_key + 2);    // [8] warp maxima, then [0] = max_k |E_k|^2
    int* pair_count = reinterpret_cast<int*>(s_bmax + 8);     // [2]
    int* ovf_count = pair_count + 2;                          // [2]
    int* spill_cnt = ovf_count + 2;                           // [4] entries spilled for item it & 3

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_ctiles = K / TC2_CODES;
    const bool have_oh = fr.onehot != nullptr;
    const bool quant = fr.q_out != nullptr;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_z);
        tma_prefetch_desc(&tm_ehi);
        for (int i = 0; i < ZBUF; ++i) {
            mbar_init(bar_z_full + i, 1);
            mbar_init(bar_z_free + i, 1);
            mbar_init(bar_z_ready + i, 128);
        }
        for (int i = 0; i < TC2_ARING; ++i) mbar_init(bar_a_ready + i, 64);
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_acc_full + b, 1);
            mbar_init(bar_acc_empty + b, 16);
            mbar_init(bar_idx_ready + b, 4);
            mbar_init(bar_idx_free + b, have_oh ? 3 : 2);    // one arrival per warp of the worker group that owns the slot
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_2sm(tmem_slot, TC2_TMEM_COLS);
        tmem_relinquish_2sm();
    }
    if (warp >= 12) {
        for (int i = threadIdx.x - 384; i < TC2_ZERO_BYTES / 16; i += 128)
            reinterpret_cast<float4*>(zero_row)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (threadIdx.x == 384) {
            *zeros_done = 0;
            for (int i = 0; i < 2; ++i) ovf_count[i] = pair_count[i] = 0;
            for (int i = 0; i < 4; ++i) spill_cnt[i] = 0;
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait_prior_grids();
    // the workspace's ready word (raised by the last CTA once everything but loss / perplexity is complete): lowered
    // before any dependent can be launched -- the trigger below is issued by this very thread after the store
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        fr.counter[2] = 0u;
        __threadfence();
    }
    __syncthreads();
    // The next kernel of the stream may be LAUNCHED from here on: its CTAs only become resident as ours exit (this
    // kernel owns the SM's registers and shared memory) and they order themselves with griddepcontrol.wait, so the
    // trigger costs nothing and takes the launch latency off the step's critical path.
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        VQ_TR(7, 0);
#ifdef VQ_TRACE
        if (fr.trace != nullptr) fr.trace[(blockIdx.x * 8 + 7) * 64 + 3] = static_cast<long long>(global_timer_ns());
#endif
    }

    if (warp == 0) {
        // ===== E producer: tf32(E) slabs only =====
        if (lane == 0) {
            int L = 0;
            for (int w = pair; w < n_items; w += n_pairs) {
                for (int ct = 0; ct < n_ctiles; ++ct) {
                    for (int i = 0; i < NSLAB; ++i, ++L) {
                        const int stage = L % NSTAGE;
                        mbar_wait(bar_empty + stage, ((L / NSTAGE) & 1) ^ 1);
                        if (leader) mbar_arrive_expect_tx(bar_full + stage, 2 * TC_SLAB_BYTES);
                        tma_load_2d_2sm(stages + stage * TC_SLAB_BYTES, &tm_ehi, bar_full + stage, i * TC_SLAB_FLOATS,
                                        ct * TC2_CODES + static_cast<int>(cta_rank) * TC_ROWS);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only): one TF32 pass =====
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(2 * TC_ROWS, TC2_CODES);
            int L = 0, ctg = 0, it = 0;
            for (int w = pair; w < n_items; w += n_pairs, ++it) {
                const int zb = it % ZBUF;
                VQ_TR(1, 3 * it);
                mbar_wait(bar_z_ready + zb, (it / ZBUF) & 1);
                VQ_TR(1, 3 * it + 1);
                tc_fence_after();
                const uint32_t z_addr = smem_u32(zbufs + zb * ZBYTES);
                for (int ct = 0; ct < n_ctiles; ++ct, ++ctg) {
                    const int buf = ctg & 1;
#ifdef VQ_TRACE
                    long long tw0 = clock64();
#endif
                    mbar_wait(bar_acc_empty + buf, ((ctg >> 1) & 1) ^ 1);
#ifdef VQ_TRACE
                    if (fr.trace != nullptr) fr.trace[(blockIdx.x * 8 + 6) * 64 + 32] += clock64() - tw0;
#endif
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * TC2_CODES;
                    uint32_t accumulate = 0;
                    for (int i = 0; i < NSLAB; ++i, ++L) {
                        const int stage = L % NSTAGE;
#ifdef VQ_TRACE
                        tw0 = clock64();
#endif
                        mbar_wait(bar_full + stage, (L / NSTAGE) & 1);
#ifdef VQ_TRACE
                        if (fr.trace != nullptr) fr.trace[(blockIdx.x * 8 + 6) * 64 + 33] += clock64() - tw0;
#endif
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(stages + stage * TC_SLAB_BYTES);
                        const uint32_t a_addr = z_addr + i * TC_SLAB_BYTES;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            tc_mma_tf32_2sm(d_tmem, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32), idesc,
                                            accumulate);
                            accumulate = 1;
                        }
                        tc_commit_2sm(bar_empty + stage);
                    }
                    tc_commit_2sm(bar_acc_full + buf);
                }
                tc_commit_2sm(bar_z_free + zb);
                VQ_TR(1, 3 * it + 2);
            }
        }
    } else if (warp < 4) {
        // ===== z pipeline (64 threads per CTA): TMA, |z_n|^2 chain, in-place tf32 rounding, publish =====
        const int c = threadIdx.x - 64;
        int it = 0;
        for (int w = pair; w < n_items; w += n_pairs, ++it) {
            const int zb = it % ZBUF;
            const int row_tile = 2 * w + static_cast<int>(cta_rank);
            uint8_t* zt = zbufs + zb * ZBYTES;
            if (c == 0) {
                mbar_wait(bar_z_free + zb, ((it / ZBUF) & 1) ^ 1);
                mbar_arrive_expect_tx(bar_z_full + zb, NSLAB * TC_SLAB_BYTES);
                for (int s = 0; s < NSLAB; ++s)
                    tma_load_2d(zt + s * TC_SLAB_BYTES, &tm_z, bar_z_full + zb, s * TC_SLAB_FLOATS, row_tile * TC_ROWS);
            }
            mbar_wait(bar_z_full + zb, (it / ZBUF) & 1);
            float* a_dst = a_ring + (it % TC2_ARING) * TC_ROWS;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = c + h * 64;
                const uint8_t* rowp = zt + (r >> 3) * 1024 + (r & 7) * 128;
                float a = 0.0f;
#pragma unroll
                for (int s = 0; s < NSLAB; ++s) {
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const float4 v = *reinterpret_cast<const float4*>(rowp + s * TC_SLAB_BYTES + ((ch ^ (r & 7)) << 4));
                        a = fmaf(v.x, v.x, a);
                        a = fmaf(v.y, v.y, a);
                        a = fmaf(v.z, v.z, a);
                        a = fmaf(v.w, v.w, a);
                    }
                }
                a_dst[r] = a;
            }
            float4* z4 = reinterpret_cast<float4*>(zt);
            constexpr int NV = NSLAB * TC_SLAB_BYTES / 16;
            if (it == 0) {
                named_bar_sync(5, 320);
                round_tf32_inplace(z4, NV, c, 320);
                fence_proxy_async_smem();
                named_bar_sync(5, 320);
            } else {
                named_bar_sync(3, 64);
                round_tf32_inplace(z4, NV, c, 64);
            }
            fence_proxy_async_smem();
            mbar_arrive(bar_a_ready + (it % TC2_ARING));
            mbar_arrive_cluster(bar_z_ready + zb, 0);
        }
        if (c == 0) {
            for (int j = it > ZBUF ? it - ZBUF : 0; j < it; ++j) mbar_wait(bar_z_free + (j % ZBUF), (j / ZBUF) & 1);
        }
    } else if (warp < 12) {
        // ===== epilogue (8 warps): screening scores -> candidates =====
        const int ew = warp - 4;
        const int half = ew >> 2;
        const int row = (ew & 3) * 32 + lane;
        const int et = threadIdx.x - 128;
        const uint32_t lane_base = static_cast<uint32_t>((ew & 3) * 32) << 16;
        // max_k |E_k|^2 (for the margin): 256 threads scan e_norm2 once
        {
            float bm = 0.0f;
            for (int k = et; k < K; k += 256) bm = fmaxf(bm, __ldg(e_norm2 + k));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
            if (lane == 0) s_bmax[ew] = bm;
            named_bar_sync(1, 256);
            bm = s_bmax[0];
#pragma unroll
            for (int i = 1; i < 8; ++i) bm = fmaxf(bm, s_bmax[i]);
            named_bar_sync(1, 256);
            if (et == 0) s_bmax[0] = bm;
            named_bar_sync(1, 256);
        }
        const float b_max = s_bmax[0];
        {   // help the z pipeline with the first tile
            mbar_wait(bar_z_full + 0, 0);
            named_bar_sync(5, 320);
            round_tf32_inplace(reinterpret_cast<float4*>(zbufs), NSLAB * TC_SLAB_BYTES / 16, 64 + et, 320);
            fence_proxy_async_smem();
            named_bar_sync(5, 320);
        }
        int it = 0, ctg = 0;
        float b_next = __ldg(e_norm2 + et);
        for (int w = pair; w < n_items; w += n_pairs, ++it) {
            mbar_wait(bar_a_ready + (it % TC2_ARING), (it / TC2_ARING) & 1);
            const float a_n = a_ring[(it % TC2_ARING) * TC_ROWS + row];
            // margin on the score scale: 2*eta = 4*eps + rounding slack (see the header)
            // + the fp32 accumulation error of the tensor core's and of the oracle's D-term sums: 2 * D * 2^-24 |z||E|
            // together, x 4 on this scale (negligible next to the operand term at D <= 256, but part of the bound)
            constexpr float margin_c = 0.00398438f + static_cast<float>(D) * 4.7683716e-7f;
            const float margin = 1.0001f * (margin_c * sqrtf(a_n * b_max) + 1.9073486e-6f * (a_n + b_max));
            float rmin = INFINITY;   // running minimum of the screening scores of this (row, half)
            int n = 0;               // entries in this thread's candidate log
            float lost = INFINITY;   // smallest score that is within the margin but is not in the log ...
            int lost_loc = -1;       // ... and where it hides: chain-instance ct*8 + half*4 + q, or -2 = anywhere
            float2* my_log = clog + et;
            bool spilled = false;    // this row has entries in the CTA's spill list (global memory)
            // what fits neither the log nor the registers goes to a small per-item list in global memory (rare)
            auto spill = [&](float v, int x, int type) -> bool {
                const int pos = atomicAdd(spill_cnt + (it & 3), 1);
                if (pos >= SC_SPILL) return false;
                VQ_ASSERT(pos >= 0 && row >= 0 && row < TC_ROWS);
                fr.spill[(static_cast<size_t>(blockIdx.x) * 4 + (it & 3)) * SC_SPILL + pos] = make_int4(row, __float_as_int(v), x, type);
                spilled = true;
                return true;
            };
            for (int ct = 0; ct < n_ctiles; ++ct, ++ctg) {
                const int buf = ctg & 1;
                const int k0 = ct * TC2_CODES;
                b_tile[buf * TC2_CODES + et] = b_next;
                named_bar_sync(1, 256);
                b_next = __ldg(e_norm2 + (ct + 1 == n_ctiles ? 0 : k0 + TC2_CODES) + et);   // for the next tile
                if (et == 0) VQ_TR(0, 3 * ctg);
                mbar_wait(bar_acc_full + buf, (ctg >> 1) & 1);
                if (et == 0) VQ_TR(0, 3 * ctg + 1);
                tc_fence_after();
                const uint32_t t_addr = tmem_base + lane_base + buf * TC2_CODES + half * 128;
                const float4* b4 = reinterpret_cast<const float4*>(b_tile + buf * TC2_CODES + half * 128);
                float m1[4], m2[4], i1f[4];     // per chain: best, runner-up (value only), column of the best (as a float)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    m1[q] = INFINITY;
                    m2[q] = INFINITY;
                    i1f[q] = 0.0f;
                }
                uint32_t va[32], vb[32];
                // The scan is bound by the ALU pipe (compare / select / min-max issue every 2 cycles per warp), so the
                // two conditional updates are done as predicated FMA-pipe moves (d = s*1 + 0) and the column index is
                // carried as a float: per element 3 FMA-pipe ops (score, 2 moves) and 2.5 ALU-pipe ops (setp, max, half of a
                // 3-input min).
                auto consume = [&](const uint32_t (&v)[32], int cc) {       // 32 columns starting at cc*32
#pragma unroll
                    for (int j4 = 0; j4 < 8; j4 += 2) {
                        const float4 ba = b4[cc * 8 + j4], bb = b4[cc * 8 + j4 + 1];
                        const float bq[2][4] = {{ba.x, ba.y, ba.z, ba.w}, {bb.x, bb.y, bb.z, bb.w}};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float t[2];                    // max(score, best so far): what the runner-up can fall to
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const float s = fmaf(-2.0f, __uint_as_float(v[(j4 + h) * 4 + q]), bq[h][q]);
                                t[h] = fmaxf(s, m1[q]);
                                const float colf = static_cast<float>(cc * 32 + (j4 + h) * 4 + q);
                                asm("{\n"
                                    ".reg .pred p;\n"
                                    "setp.lt.f32 p, %2, %0;\n"
                                    "@p fma.rn.f32 %0, %2, 0f3F800000, 0f00000000;\n"
                                    "@p fma.rn.f32 %1, %3, 0f3F800000, 0f00000000;\n"
                                    "}\n"
                                    : "+f"(m1[q]), "+f"(i1f[q])
                                    : "f"(s), "f"(colf));
                            }
                            m2[q] = fminf(fminf(m2[q], t[0]), t[1]);   // one 3-input FMNMX3 per two elements
                        }
                    }
                };
                // 32-column chunks, double buffered: every tcgen05.wait::ld is covered by a whole chunk of scanning
                tmem_ld_32x32b_x32(t_addr, va);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(t_addr + 32, vb);
                consume(va, 0);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(t_addr + 64, va);
                consume(vb, 1);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(t_addr + 96, vb);
                consume(va, 2);
                tmem_ld_wait();
                // this warp's share of the accumulator is in registers: hand the buffer back early
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(bar_acc_empty + buf, 0);
                consume(vb, 3);
                // log the chain minima that cannot be ruled out yet
                {
                    const float rnew = fminf(fminf(rmin, fminf(m1[0], m1[1])), fminf(m1[2], m1[3]));
                    const float t = rnew + margin;
                    if (t < rmin) n = 0;            // the new minimum beats everything logged so far by more than the margin
                    rmin = rnew;
                    int slot[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        slot[q] = n;
                        n += m1[q] <= t ? 1 : 0;
                    }
                    if (n <= SC_LOG) {              // the common case: predicated appends
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            VQ_ASSERT(!(m1[q] <= t) || (slot[q] >= 0 && slot[q] < SC_LOG));
                            VQ_ASSERT(i1f[q] >= 0.0f && i1f[q] < 128.0f);
                            if (m1[q] <= t)
                                my_log[slot[q] * 256] = make_float2(m1[q], __int_as_float(k0 + half * 128 + static_cast<int>(i1f[q])));
                        }
                    } else {                        // log full (rare): drop what the minimum has ruled out, then append
                        int m = 0;
                        for (int j = 0; j < slot[0]; ++j) {
                            const float2 o = my_log[j * 256];
                            if (o.x <= t) my_log[(m++) * 256] = o;
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (m1[q] <= t) {
                                if (m < SC_LOG) {
                                    my_log[(m++) * 256] = make_float2(m1[q], __int_as_float(k0 + half * 128 + static_cast<int>(i1f[q])));
                                } else if (!spill(m1[q], k0 + half * 128 + static_cast<int>(i1f[q]), 0)) {
                                    lost = (lost <= t) ? fminf(lost, m1[q]) : m1[q];    // nowhere to keep it: can be anywhere
                                    lost_loc = -2;
                                }
                            }
                        }
                        n = m;
                        VQ_ASSERT(n >= 0 && n <= SC_LOG);
                    }
                    const float w = fminf(fminf(m2[0], m2[1]), fminf(m2[2], m2[3]));
                    if (w <= t) {                   // a chain's runner-up is hidden behind its best: remember the 32 columns
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (m2[q] <= t) {
                                const int loc = ct * 8 + half * 4 + q;
                                if (lost_loc == -1 || !(lost <= t)) {   // nothing (still relevant) recorded yet
                                    lost = m2[q];
                                    lost_loc = loc;
                                } else if (lost_loc == -2 || !spill(m2[q], loc, 1)) {   // a second place to look: spill it
                                    lost = fminf(lost, m2[q]);
                                    lost_loc = -2;
                                }
                            }
                        }
                    }
                }
                if (et == 0) VQ_TR(0, 3 * ctg + 2);
            }
            // merge the two column halves and publish the row's candidates
            if (half == 1) {
                mrg[row] = make_float4(rmin, __int_as_float(n | (spilled ? 256 : 0)), lost, __int_as_float(lost_loc));
            } else {
                mbar_wait(bar_idx_free + (it & 1), ((it >> 1) & 1) ^ 1);   // the workers are done with this handoff slot
            }
            named_bar_sync(2, 256);
            if (half == 0) {
                const float4 ms = mrg[row];
                const float smin = fminf(rmin, ms.x);
                const float thr = smin + margin;
                int* cand = h_cand + ((it & 1) * TC_ROWS + row) * 4;
                int* locs = h_loc + ((it & 1) * TC_ROWS + row) * 2;
                int nc = 0, nloc = 0;
                bool full = false;
                const int n1 = __float_as_int(ms.y) & 255;
                if (__float_as_int(ms.y) & 256) spilled = true;
#pragma unroll
                for (int j = 0; j < SC_LOG; ++j) {
                    if (j < n) {
                        const float2 e = my_log[j * 256];
                        if (e.x <= thr) {
                            if (nc < 4) cand[nc++] = __float_as_int(e.y);
                            else if (!spill(e.x, __float_as_int(e.y), 0)) full = true;
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < SC_LOG; ++j) {
                    if (j < n1) {
                        const float2 e = my_log[j * 256 + 128];
                        if (e.x <= thr) {
                            if (nc < 4) cand[nc++] = __float_as_int(e.y);
                            else if (!spill(e.x, __float_as_int(e.y), 0)) full = true;
                        }
                    }
                }
                if (lost <= thr) {
                    if (lost_loc >= 0) locs[nloc++] = lost_loc; else full = true;
                }
                if (ms.z <= thr) {
                    const int l1 = __float_as_int(ms.w);
                    if (l1 >= 0) locs[nloc++] = l1; else full = true;
                }
                if (nc == 0) full = true;
                VQ_ASSERT(nc >= 0 && nc <= 4 && nloc >= 0 && nloc <= 2);
                h_nc[(it & 1) * TC_ROWS + row] = full ? SC_NC_OVERFLOW : (nc | (nloc << 4) | (spilled ? 0x40 : 0));
                h_an[(it & 1) * TC_ROWS + row] = a_n;
                h_thr[(it & 1) * TC_ROWS + row] = thr;
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_idx_ready + (it & 1));
                if (et == 0) VQ_TR(3, 4 * it + 3);
            }
        }
    } else if (warp == 12 && have_oh) {
        // ===== one-hot zero filler (1 thread): TMA bulk copies from one shared zero row, from kernel start =====
        if (lane == 0) {
            const uint32_t row_bytes = static_cast<uint32_t>(K) * 4u;
            const uint32_t chunk = row_bytes < TC2_ZERO_BYTES ? row_bytes : TC2_ZERO_BYTES;
            const uint64_t pol = l2_policy_evict_first();
            VQ_TR(4, 0);
            for (int w = pair; w < n_items; w += n_pairs) {
                const long long row0 = static_cast<long long>(2 * w + static_cast<int>(cta_rank)) * TC_ROWS;
                const long long left = N - row0;
                const int rows_here = left <= 0 ? 0 : (left < TC_ROWS ? static_cast<int>(left) : TC_ROWS);
                uint8_t* obase = reinterpret_cast<uint8_t*>(fr.onehot + row0 * K);
                const size_t total = static_cast<size_t>(rows_here) * row_bytes;
                if (fr.onehot_evict_first) {
                    for (size_t off = 0; off < total; off += chunk) bulk_store_s2g_hint(obase + off, zero_row, chunk, pol);
                } else {
                    for (size_t off = 0; off < total; off += chunk) bulk_store_s2g(obase + off, zero_row, chunk);
                }
                bulk_commit();
            }
            VQ_TR(4, 1);
            bulk_wait_all();
            VQ_TR(4, 2);
            fence_proxy_async_all();
            st_release_cta(zeros_done, 1);
        }
    } else {
        // ===== row workers (3 or 4 warps): exact refine, then vector_quantizer.py:40-56 =====
        // Per item the work is a chain of short, latency-bound phases, so without the one-hot filler the four warps
        // run as TWO independent groups, one per handoff slot (even / odd items), which overlaps those latencies.
        constexpr int DV = NSLAB * 8;
        const int NW_all = have_oh ? 96 : 128;
        const int wt_all = have_oh ? threadIdx.x - 416 : threadIdx.x - 384;
        const int groups = have_oh ? 1 : 2;
        const int NW = NW_all / groups;                      // threads of this group
        const int grp = wt_all / NW, wt = wt_all - grp * NW; // group, thread within the group
        const int gbar = 4 + 2 * grp;                        // the group's named barrier
        const int pmax = SC_PMAX / groups;
        unsigned long long* const g_key = s_key + grp * TC_ROWS;
        int* const g_idx = s_idx + grp * TC_ROWS;
        int* const g_ovf = s_ovf + grp * TC_ROWS;
        int* const g_rng = s_rng + grp * TC_ROWS;
        int* const g_pair = pair_rc + grp * pmax;
        int* const g_pcount = pair_count + grp;
        int* const g_ocount = ovf_count + grp;
        unsigned long long* const g_okey = ovf_key + grp;
        float sse = 0.0f;
        // Usage counts.  Reds on the same address queue up in L2 (~0.1 - 0.2 us apiece: N / K of them per code), which at
        // K = 512, N = 1M is longer than the rest of the kernel.  Without the dense one-hot the zero row is idle: it holds
        // a per-CTA histogram (shared-memory integer atomics), flushed with one red per used code when the CTA is done.
        int* const s_hist = reinterpret_cast<int*>(zero_row);
        const bool smem_hist = !have_oh && K <= TC2_ZERO_BYTES / 4;
        for (int w = pair + grp * n_pairs, it = grp; w < n_items; w += groups * n_pairs, it += groups) {
            const long long row0 = static_cast<long long>(2 * w + static_cast<int>(cta_rank)) * TC_ROWS;
            const long long left = N - row0;
            const int rows_here = left <= 0 ? 0 : (left < TC_ROWS ? static_cast<int>(left) : TC_ROWS);
            mbar_wait(bar_idx_ready + (it & 1), (it >> 1) & 1);
            if (wt == 0) VQ_TR(5, 4 * it);
            const int* cand = h_cand + (it & 1) * TC_ROWS * 4;
            const int* ncs = h_nc + (it & 1) * TC_ROWS;
            const float* ans = h_an + (it & 1) * TC_ROWS;
            // -- refine: exact fp32 distances (oracle order) of every (row, code) pair that is still undecided -------
            const int* locs = h_loc + (it & 1) * TC_ROWS * 2;
            auto push_pairs = [&](int r, int code, int nloc_entries, int loc0, int loc1) -> bool {
                // appends (r, code) if code >= 0 and the 32 codes of every listed chain-instance; false if the list is full
                const int cnt = (code >= 0 ? 1 : 0) + 32 * nloc_entries;
                const int base = atomicAdd(g_pcount, cnt);
                if (base + cnt > pmax) {
                    for (int p = base; p < pmax; ++p) g_pair[p] = -1;
                    return false;
                }
                int p = base;
                VQ_ASSERT(base >= 0 && base + cnt <= pmax && r >= 0 && r < TC_ROWS && code < K);
                if (code >= 0) g_pair[p++] = (r << 20) | code;
                for (int l = 0; l < nloc_entries; ++l) {
                    const int loc = l == 0 ? loc0 : loc1;
                    const int kb = (loc >> 3) * TC2_CODES + ((loc >> 2) & 1) * 128 + (loc & 3);
                    VQ_ASSERT(loc >= 0 && kb + 4 * 31 < K);
                    for (int j = 0; j < 32; ++j) g_pair[p++] = (r << 20) | (kb + 4 * j);
                }
                return true;
            };
            for (int r = wt; r < rows_here; r += NW) {       // 1. enumerate the pairs
                const int st = ncs[r];
                int state = 0;
                if (st == SC_NC_OVERFLOW) {
                    state = 2;
                } else {
                    const int nc = st & 0xf, nloc = (st >> 4) & 3;
                    if (nc == 1 && nloc == 0 && !(st & 0x40)) {
                        g_idx[r] = cand[r * 4];              // a single candidate IS the argmin: nothing to compute
                    } else {
                        state = 1;
                        g_key[r] = ~0ull;
                        for (int j = 0; j < nc && state == 1; ++j)
                            if (!push_pairs(r, cand[r * 4 + j], 0, 0, 0)) state = 2;
                        if (state == 1 && nloc > 0 && !push_pairs(r, -1, nloc, locs[r * 2], locs[r * 2 + 1])) state = 2;   // list full: rescan the row
                    }
                }
                if (state == 2) {
                    const int o = atomicAdd(g_ocount, 1);
                    VQ_ASSERT(o >= 0 && o < TC_ROWS);
                    g_ovf[o] = r;
                }
                g_rng[r] = state;
            }
            named_bar_sync(gbar, NW);
            const int n_spill = min(spill_cnt[it & 3], SC_SPILL);
            if (n_spill > 0) {                               // 1b. rows whose candidates overflowed into global memory
                const int4* sp = fr.spill + (static_cast<size_t>(blockIdx.x) * 4 + (it & 3)) * SC_SPILL;
                const float* thrs = h_thr + (it & 1) * TC_ROWS;
                for (int e = wt; e < n_spill; e += NW) {
                    const int4 en = __ldcg(sp + e);
                    const int r = en.x;
                    if (r >= rows_here || ncs[r] == SC_NC_OVERFLOW || !(__int_as_float(en.y) <= thrs[r])) continue;
                    const bool ok = en.w == 0 ? push_pairs(r, en.z, 0, 0, 0) : push_pairs(r, -1, 1, en.z, 0);
                    if (!ok && atomicExch(g_rng + r, 2) != 2) g_ovf[atomicAdd(g_ocount, 1)] = r;
                }
                named_bar_sync(gbar, NW);
            }
            {                                                // 2. one exact distance per thread and pass
                const int np = min(*g_pcount, pmax);
                for (int p = wt; p < np; p += NW) {
                    const int rc = g_pair[p];
                    if (rc >= 0) {
                        const int r = rc >> 20, k = rc & 0xfffff;
                        VQ_ASSERT(r >= 0 && r < rows_here && k < K);
                        const float c = dot_chain_exact<D>(fr.z + (row0 + r) * D, fr.E + static_cast<size_t>(k) * D);
                        atomicMin(g_key + r, pack_key(fmaf(-2.0f, c, ans[r] + __ldg(e_norm2 + k)), k));
                    }
                }
            }
            named_bar_sync(gbar, NW);
            for (int r = wt; r < rows_here; r += NW)         // 3. per row: smallest distance, first index on ties
                if (g_rng[r] == 1) g_idx[r] = static_cast<int>(g_key[r] & 0xffffffffu);
            named_bar_sync(gbar, NW);
            if (wt == 0) VQ_TR(5, 4 * it + 1);
            // -- rare: rows whose candidate set could not be bounded -> exact scan of the whole codebook ---------
            const int n_ovf = *g_ocount;
#ifdef VQ_TRACE
            if (wt == 0 && fr.trace != nullptr) {
                fr.trace[(blockIdx.x * 8 + 6) * 64 + 34] += n_ovf;
                fr.trace[(blockIdx.x * 8 + 6) * 64 + 35] += *g_pcount;
            }
#endif
            for (int o = 0; o < n_ovf; ++o) {
                const int r = g_ovf[o];
                if (wt == 0) *g_okey = ~0ull;
                named_bar_sync(gbar, NW);
                const float4* z4r = reinterpret_cast<const float4*>(fr.z + (row0 + r) * D);
                const float a = ans[r];
                unsigned long long key = ~0ull;
                // 4 codes per thread at a time, interleaved: their loads overlap, each code keeps its own d-ordered chain
                for (int kb = wt; kb < K; kb += 4 * NW) {
                    float acc[4];
                    const float4* e4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[j] = 0.0f;
                        const int k = kb + j * NW;
                        e4[j] = reinterpret_cast<const float4*>(fr.E + static_cast<size_t>(k < K ? k : 0) * D);
                    }
#pragma unroll 2
                    for (int i = 0; i < D / 4; ++i) {
                        const float4 zv = __ldg(z4r + i);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 ev = __ldg(e4[j] + i);
                            acc[j] = fmaf(zv.x, ev.x, acc[j]);
                            acc[j] = fmaf(zv.y, ev.y, acc[j]);
                            acc[j] = fmaf(zv.z, ev.z, acc[j]);
                            acc[j] = fmaf(zv.w, ev.w, acc[j]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = kb + j * NW;
                        if (k < K) {
                            const float dist = fmaf(-2.0f, acc[j], a + __ldg(e_norm2 + k));
                            const unsigned long long kk = pack_key(dist, k);
                            key = kk < key ? kk : key;
                        }
                    }
                }
                atomicMin(g_okey, key);
                named_bar_sync(gbar, NW);
                if (wt == 0) g_idx[r] = static_cast<int>(*g_okey & 0xffffffffu);
                named_bar_sync(gbar, NW);
            }
            if (wt == 0) {
                *g_ocount = 0;
                *g_pcount = 0;
                spill_cnt[it & 3] = 0;
            }
            named_bar_sync(gbar, NW);
            if (wt == 0) VQ_TR(5, 4 * it + 2);
            // -- indices, usage histogram -----------------------------------------------------------------------
            for (int r = wt; r < rows_here; r += NW) {
                const int code = g_idx[r];
                VQ_ASSERT(code >= 0 && code < K);
                idx_out[row0 + r] = code;
                if (!fr.rows_later) {
                    if (smem_hist) atomicAdd(s_hist + code, 1);
                    else atomicAdd(fr.hist + code, 1.0f);
                }
            }
            // -- q_out = fl(z + fl(E[idx] - z)), sse += (E[idx] - z)^2 ---------------------------------------------
            if (quant) {
                const float4* z4 = reinterpret_cast<const float4*>(fr.z + row0 * D);
                float4* q4 = reinterpret_cast<float4*>(fr.q_out + row0 * D);
                const int n_el = rows_here * DV;
                constexpr int UB = 8;
#pragma unroll 1
                for (int ub = 0; wt + ub * NW < n_el; ub += UB) {
                    float4 zv[UB], ev[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        const int e = wt + (ub + u) * NW;
                        if (e < n_el) {
                            zv[u] = __ldg(z4 + e);
                            ev[u] = __ldg(reinterpret_cast<const float4*>(fr.E + static_cast<size_t>(g_idx[e / DV]) * D) + (e % DV));
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
            }
            named_bar_sync(gbar, NW);                       // everyone is done with g_idx and the handoff slot
            if (lane == 0) mbar_arrive(bar_idx_free + (it & 1));
            if (wt == 0) VQ_TR(5, 4 * it + 3);
        }
        if (have_oh) {
            if (wt_all == 0) VQ_TR(6, 0);
            patch_onehot_ones(fr.onehot, idx_out, N, K, pair, n_items, n_pairs, wt_all, NW_all,
                              [&](int w) { return static_cast<long long>(2 * w + static_cast<int>(cta_rank)) * TC_ROWS; }, zeros_done);
        }
        if (wt_all == 0) VQ_TR(6, 2);
        // per-CTA SSE partial, then last-CTA-done reduction in a fixed order (+ loss / perplexity)
        if (smem_hist && !fr.rows_later) {
            named_bar_sync(7, NW_all);                      // both worker groups are through their items
            for (int k = wt_all; k < K; k += NW_all) {
                const int c = s_hist[k];
                if (c != 0) atomicAdd(fr.hist + k, static_cast<float>(c));     // integers below 2^24: exact in any order
            }
        }
        // (rows_later: the rows kernel behind this one owns the completion counter and the statistics)
        if (!fr.rows_later) publish_and_finalize(fr, sse, N, K, D, wt_all, NW_all, lane, have_oh ? warp - 13 : warp - 12, have_oh ? 3 : 4, red, 7);
    }

    if (threadIdx.x == 0) VQ_TR(7, 1);
    tc_fence_before();
    cluster_sync_all();
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
#ifdef VQ_TRACE
    if (threadIdx.x == 64 && fr.trace != nullptr) fr.trace[(blockIdx.x * 8 + 7) * 64 + 6] = static_cast<long long>(global_timer_ns());
#endif
}

}  // namespace b200vq
