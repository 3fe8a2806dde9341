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
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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

}  // namespace b200vq
