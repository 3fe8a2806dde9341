// kernels_dp.cuh -- data-parallel exchange of the packed step buffer [dE (K*D) | usage histogram (K) | sse (1)]
// over NVLink peer memory (the one collective of the path; the reference has none -- DDP
// would all-reduce `_embedding.weight.grad` after backward).
//
// Every rank owns two symmetric RECEIVE buffers (alternating between calls).  Data travels as 16-byte lines
// {d0, seq, d1, seq}: payload and the call's sequence number in the same store, so a line whose two flag words
// carry `seq` is complete (16-byte stores may tear only at 8-byte granularity) -- no barrier, no second message.
//   one step  (world < 8): every rank stores line i into slot [rank] of EVERY rank's buffer (one multimem.st when
//             the NVLS multicast mapping exists), then sums its own `world` slots in rank order.
//   two steps (world >= 8): rank j owns slice j; (1) line i goes to slot [rank] of region A of its owner,
//             (2) the owner sums its slots in rank order and stores the result into region B of every rank,
//             (3) everybody collects region B.  2 x payload instead of world x payload lands in each rank.
// Sums run in rank order, so the result is bit-identical on every rank.
//
// The sequence number lives in DEVICE memory (state[0] = calls completed): the kernels derive buffer parity and
// `seq` from it and the last CTA of the finishing kernel advances it, so a captured CUDA graph can be replayed
// and no host state has to stay in step with the device.  Every wait is bounded: after `spin_limit` polls a
// thread raises state[1] and the kernel drains without waiting any further (results are then garbage, the host
// reads the error word with vq_dp_status).
//
// The kernel is launched with the programmatic-launch attribute right behind the backward: its CTAs are resident and
// past their prologue when the backward's last store lands, and the next step's first kernel is launched while it polls.
#pragma once
#include "common.cuh"

namespace b200vq {

constexpr int DP_MAX_RANKS = 16;
constexpr int DP_THREADS = 256;

struct DpCtxDev {
    float* recv[2][DP_MAX_RANKS];   // [parity][rank]: that rank's receive buffer as mapped into this process
    float* mc[2];                   // NVLS multicast mapping of the same buffers, or nullptr
    unsigned int* state;            // [0] calls completed, [1] error word, [2] finishing-CTA counter
    long long L;                    // 16-byte lines per payload = ceil(n / 2)
    long long S;                    // two steps: lines per owner slice = ceil(L / world)
    long long n;                    // payload floats
    int world, rank;
    int two_step;
    unsigned int spin_limit;
};

struct DpCall {
    int which;                      // receive-buffer parity of this call
    unsigned int seq;               // per-buffer sequence number (never 0)
};

__device__ __forceinline__ DpCall dp_begin(const DpCtxDev& c) {
    const unsigned int s = __ldcg(c.state) + 1u;      // L2: the previous call's last CTA advanced it with an atomic
    DpCall k;
    k.which = static_cast<int>((s - 1u) & 1u);
    k.seq = (s + 1u) >> 1;
    return k;
}

__device__ __forceinline__ void st_volatile_v4(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// one store, delivered by the NVSwitch to the same offset of EVERY rank's buffer (NVLS multicast mapping)
__device__ __forceinline__ void multimem_st_v4(void* mc_addr, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}

// store a line at line offset `off` of every rank's receive buffer
__device__ __forceinline__ void dp_store_all(const DpCtxDev& c, int which, long long off, uint4 line) {
    if (c.mc[which] != nullptr) {
        multimem_st_v4(reinterpret_cast<uint4*>(c.mc[which]) + off, line);
    } else {
        for (int p = 1; p < c.world; ++p) {             // not to ourselves: the reducer takes our own lines from the local payload
            const int dst = (c.rank + p) % c.world;     // spread the ranks' first targets over the links
            st_volatile_v4(reinterpret_cast<uint4*>(c.recv[which][dst]) + off, line);
        }
    }
}

// first hop of line i of the local payload
__device__ __forceinline__ void dp_push_line(const DpCtxDev& c, const DpCall& k, long long i, float d0, float d1) {
    const uint4 line = make_uint4(__float_as_uint(d0), k.seq, __float_as_uint(d1), k.seq);
    if (c.two_step) {
        const long long j = i / c.S, l = i - j * c.S;
        if (j != c.rank) st_volatile_v4(reinterpret_cast<uint4*>(c.recv[k.which][j]) + static_cast<long long>(c.rank) * c.S + l, line);
    } else {
        dp_store_all(c, k.which, static_cast<long long>(c.rank) * c.L + i, line);
    }
}

// bounded wait for a line: false (and the error word raised) when it did not arrive within spin_limit polls
struct DpWaiter {
    const DpCtxDev& c;
    int* s_abort;                    // shared: some thread of this CTA gave up
    __device__ __forceinline__ bool wait(const uint4* p, unsigned int seq, uint4& v) const {
        unsigned int polls = 0;
        while (v.y != seq || v.w != seq) {
            if (++polls > c.spin_limit || *reinterpret_cast<volatile int*>(s_abort) != 0) {
                if (*reinterpret_cast<volatile int*>(s_abort) == 0) {
                    *reinterpret_cast<volatile int*>(s_abort) = 1;
                    atomicOr(c.state + 1, 1u);
                }
                return false;
            }
            // back off once the line is clearly not about to land: thousands of threads polling at full rate cost the kernels
            // the exchange overlaps with (vq_dp_allreduce_start) a visible share of the L2 request bandwidth
            if (polls > 16u) __nanosleep(polls > 64u ? 400u : 100u);
            v = ld_volatile_v4(p);
        }
        return true;
    }
};

// one step: sum our `world` slots in rank order as the lines arrive
__device__ __forceinline__ void dp_collect_one_step(const DpCtxDev& c, const DpCall& k, const DpWaiter& w, const float* __restrict__ payload,
                                                    float* __restrict__ out, long long first, long long stride) {
    const uint4* mine = reinterpret_cast<const uint4*>(c.recv[k.which][c.rank]);
    for (long long i = first; i < c.L; i += stride) {
        uint4 v[DP_MAX_RANKS];
#pragma unroll
        for (int p = 0; p < DP_MAX_RANKS; ++p)
            if (p < c.world && p != c.rank) v[p] = ld_volatile_v4(mine + static_cast<long long>(p) * c.L + i);
        // our own contribution never travels: it is read where it lies, at its place in the rank order
        const float own0 = __ldcg(payload + 2 * i), own1 = (2 * i + 1 < c.n) ? __ldcg(payload + 2 * i + 1) : 0.0f;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int p = 0; p < DP_MAX_RANKS; ++p) {
            if (p < c.world) {
                if (p == c.rank) {
                    a0 += own0;
                    a1 += own1;
                } else {
                    w.wait(mine + static_cast<long long>(p) * c.L + i, k.seq, v[p]);
                    a0 += __uint_as_float(v[p].x);
                    a1 += __uint_as_float(v[p].z);
                }
            }
        }
        out[2 * i] = a0;
        if (2 * i + 1 < c.n) out[2 * i + 1] = a1;
    }
}

// two steps, (2): reduce our own slice in rank order and broadcast it into region B of every rank
__device__ __forceinline__ void dp_reduce_bcast(const DpCtxDev& c, const DpCall& k, const DpWaiter& w, const float* __restrict__ payload,
                                                float* __restrict__ out, long long first, long long stride) {
    const uint4* mine = reinterpret_cast<const uint4*>(c.recv[k.which][c.rank]);
    const long long own = min(c.S, c.L - static_cast<long long>(c.rank) * c.S);
    const long long b_off = static_cast<long long>(c.world) * c.S + static_cast<long long>(c.rank) * c.S;
    for (long long l = first; l < own; l += stride) {
        const long long gi = static_cast<long long>(c.rank) * c.S + l;        // this line in the payload: our own share stays local
        const float own0 = __ldcg(payload + 2 * gi), own1 = (2 * gi + 1 < c.n) ? __ldcg(payload + 2 * gi + 1) : 0.0f;
        float a0 = 0.f, a1 = 0.f;
        for (int p0 = 0; p0 < c.world; p0 += 8) {           // 8 slots requested together, summed in rank order
            uint4 v[8];
#pragma unroll
            for (int p = 0; p < 8; ++p)
                if (p0 + p < c.world && p0 + p != c.rank) v[p] = ld_volatile_v4(mine + static_cast<long long>(p0 + p) * c.S + l);
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                if (p0 + p < c.world) {
                    if (p0 + p == c.rank) {
                        a0 += own0;
                        a1 += own1;
                    } else {
                        w.wait(mine + static_cast<long long>(p0 + p) * c.S + l, k.seq, v[p]);
                        a0 += __uint_as_float(v[p].x);
                        a1 += __uint_as_float(v[p].z);
                    }
                }
            }
        }
        dp_store_all(c, k.which, b_off + l, make_uint4(__float_as_uint(a0), k.seq, __float_as_uint(a1), k.seq));
        out[2 * gi] = a0;                                   // our own slice needs no second hop
        if (2 * gi + 1 < c.n) out[2 * gi + 1] = a1;
    }
}

// two steps, (3): region B is owner-major with S lines per owner, so reduced line i sits at offset i
__device__ __forceinline__ void dp_gather(const DpCtxDev& c, const DpCall& k, const DpWaiter& w, float* __restrict__ out, long long first,
                                          long long stride) {
    const uint4* gathered = reinterpret_cast<const uint4*>(c.recv[k.which][c.rank]) + static_cast<long long>(c.world) * c.S;
    const long long own_lo = static_cast<long long>(c.rank) * c.S, own_hi = own_lo + c.S;   // reduced by ourselves: already in `out`
    for (long long i0 = first; i0 < c.L; i0 += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * stride;
            if (i < c.L && (i < own_lo || i >= own_hi)) v[u] = ld_volatile_v4(gathered + i);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * stride;
            if (i < c.L && (i < own_lo || i >= own_hi)) {
                w.wait(gathered + i, k.seq, v[u]);
                out[2 * i] = __uint_as_float(v[u].x);
                if (2 * i + 1 < c.n) out[2 * i + 1] = __uint_as_float(v[u].z);
            }
        }
    }
}

// body shared by the production kernel (one rank per launch) and the single-GPU emulation (one rank per blockIdx.y)
__device__ __forceinline__ void dp_allreduce_body(const DpCtxDev& c, const DpCall& k, const float* __restrict__ payload, float* __restrict__ out,
                                                  int* s_abort, long long first, long long stride) {
    const DpWaiter w{c, s_abort};
    for (long long i = first; i < c.L; i += stride) {
        const float d0 = __ldcg(payload + 2 * i);
        const float d1 = (2 * i + 1 < c.n) ? __ldcg(payload + 2 * i + 1) : 0.0f;
        dp_push_line(c, k, i, d0, d1);
    }
    if (c.two_step) {
        dp_reduce_bcast(c, k, w, payload, out, first, stride);
        dp_gather(c, k, w, out, first, stride);
    } else {
        dp_collect_one_step(c, k, w, payload, out, first, stride);
    }
}

// the last CTA to finish advances the call counter (the next call -- or graph replay -- uses the other buffer)
__device__ __forceinline__ void dp_end(const DpCtxDev& c, unsigned int n_ctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(c.state + 2, 1u) == n_ctas - 1) {
            c.state[2] = 0u;
            c.state[3] = 0u;     // (backward_dp_kernel's completion tickets)
            __threadfence();
            atomicAdd(c.state, 1u);
        }
    }
}

// complete sum all-reduce of `payload` (written by the previous kernel of the stream)
__global__ void __launch_bounds__(DP_THREADS) dp_allreduce_kernel(const __grid_constant__ DpCtxDev c, const float* __restrict__ payload,
                                                                  float* __restrict__ out) {
    __shared__ int s_abort;
    pdl_launch_dependents();     // whatever follows orders itself behind this kernel's completion
    if (threadIdx.x == 0) s_abort = 0;
    __syncthreads();
    pdl_wait_prior_grids();      // the local contribution is complete -- and so is a previous exchange (its counter update)
    const DpCall k = dp_begin(c);
    dp_allreduce_body(c, k, payload, out, &s_abort, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x,
                      static_cast<long long>(gridDim.x) * blockDim.x);
    dp_end(c, gridDim.x);
}

// single-GPU emulation of `world` ranks (tests): blockIdx.y is the rank, launched cooperatively so that all ranks'
// CTAs are resident at once -- the ranks wait for each other's lines exactly as they do across GPUs.
__global__ void __launch_bounds__(DP_THREADS) dp_emulate_kernel(const DpCtxDev* __restrict__ ctxs, const float* const* __restrict__ payloads,
                                                                float* const* __restrict__ outs) {
    __shared__ int s_abort;
    __shared__ DpCtxDev c;
    if (threadIdx.x == 0) {
        c = ctxs[blockIdx.y];
        s_abort = 0;
    }
    __syncthreads();
    const DpCall k = dp_begin(c);
    dp_allreduce_body(c, k, payloads[blockIdx.y], outs[blockIdx.y], &s_abort, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x,
                      static_cast<long long>(gridDim.x) * blockDim.x);
    dp_end(c, gridDim.x);
}

// ---------------------------------------------------------------------------------------------------------------
// The data-parallel backward: backward_kernel's flat pass (dz, scatter-add of dE into the front of the packed step
// buffer) with the exchange in its TAIL -- no second launch.  A kernel of its own costs ~7 us on the step's critical path
// even when nothing has to cross NVLink (kernel boundary, the serial round trips for sequence number, payload, fence,
// counter); here every CTA takes a completion ticket when its reds are out, and the last DP_TAIL_CTAS CTAs to finish
// become the exchange: they wait (resident, so nobody is blocked) until all tickets are drawn -- the gradient is
// complete in L2 then -- and push / collect / sum their share of the lines exactly as dp_allreduce_kernel does.
// ---------------------------------------------------------------------------------------------------------------
constexpr unsigned int DP_TAIL_CTAS = 128;

template <bool HAS_GQ>
__global__ void __launch_bounds__(256) backward_dp_kernel(const float* __restrict__ g_q, const float* __restrict__ g_loss,
                                                          const float* __restrict__ z, const float* __restrict__ E,
                                                          const int* __restrict__ idx, long long N, float denom_dz, float denom_dE, int D,
                                                          float beta, float* __restrict__ dz, float* __restrict__ packed,
                                                          const unsigned int* ready, const __grid_constant__ DpCtxDev c,
                                                          float* __restrict__ out) {
    __shared__ int s_ok, s_abort;
    __shared__ unsigned int s_ticket;
    pdl_launch_dependents();
    if (ready != nullptr) {          // right behind the fused forward: start on its ready word (see backward_kernel)
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
    const int DV = D >> 2;
    const long long n_el = N * DV;
    // a few hundred long-lived CTAs (4 independent 16-byte elements in flight per thread) instead of one element per
    // thread: the completion ticket below costs every CTA a fence (its reds must have been performed) and an atomic on
    // ONE address -- with 3 216 one-shot CTAs that alone took 23 us
    const long long gstride = static_cast<long long>(gridDim.x) * 256;
    for (long long e0 = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; e0 < n_el; e0 += 4 * gstride) {
        float4 zv[4], ev[4], gv[4];
        int code[4], cc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long e = e0 + u * gstride;
            if (e < n_el) {
                const long long r = e / DV;
                cc[u] = static_cast<int>(e - r * DV);
                code[u] = __ldg(idx + r);
                zv[u] = __ldcs(reinterpret_cast<const float4*>(z) + e);
                ev[u] = __ldg(reinterpret_cast<const float4*>(E + static_cast<size_t>(code[u]) * D) + cc[u]);
                gv[u] = HAS_GQ ? __ldcs(reinterpret_cast<const float4*>(g_q) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long e = e0 + u * gstride;
            if (e < n_el) {
                float4 df, o, a;
                df.x = ev[u].x - zv[u].x; df.y = ev[u].y - zv[u].y; df.z = ev[u].z - zv[u].z; df.w = ev[u].w - zv[u].w;
                o.x = fmaf(-cz, df.x, gv[u].x); o.y = fmaf(-cz, df.y, gv[u].y);
                o.z = fmaf(-cz, df.z, gv[u].z); o.w = fmaf(-cz, df.w, gv[u].w);
                __stcs(reinterpret_cast<float4*>(dz) + e, o);
                a.x = ce * df.x; a.y = ce * df.y; a.z = ce * df.z; a.w = ce * df.w;
                atomicAdd(reinterpret_cast<float4*>(packed + static_cast<size_t>(code[u]) * D) + cc[u], a);
            }
        }
    }
    // ---- tail: completion ticket; the last DP_TAIL_CTAS finishers are the exchange ----
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                    // this CTA's reds are performed before the ticket is drawn
        s_ticket = atomicAdd(c.state + 3, 1u);
        s_abort = 0;
    }
    __syncthreads();
    const unsigned int total = gridDim.x;
    const unsigned int helpers = total < DP_TAIL_CTAS ? total : DP_TAIL_CTAS;
    if (s_ticket + helpers >= total) {
        const unsigned int h = s_ticket - (total - helpers);
        if (threadIdx.x == 0) {                              // everybody else is running or done: a bounded, deadlock-free wait
            unsigned int polls = 0;
            while (ld_acquire_gpu_u32(c.state + 3) < total) {
                if (++polls > c.spin_limit) {
                    atomicOr(c.state + 1, 2u);
                    break;
                }
            }
        }
        __syncthreads();
        const DpCall k = dp_begin(c);
        dp_allreduce_body(c, k, packed, out, &s_abort, static_cast<long long>(h) * 256 + threadIdx.x, static_cast<long long>(helpers) * 256);
        dp_end(c, helpers);
    }
    // order this kernel behind the forward's completion before it exits: "previous kernel complete" stays transitive
    if (ready != nullptr) pdl_wait_prior_grids();
}

}  // namespace b200vq
