/*
 * b200vq.h -- C ABI of the B200-native VectorQuantizer hot path (libb200vq.so).
 *
 * The reference has no FFI: the hot path sits behind a Python nn.Module,
 *   /root/reference/src/acoustic_locating_vq_vae/vq_vae/vector_quantizer.py:8-58
 * called from convolutional_vq_vae.py:98,105.  This header is the boundary a binding for that
 * module calls (the ctypes binding is acoustic_locating_vq-vae_b200/_lib.py; INTEGRATION.md shows
 * the reference-side stub).  Plain pointers and sizes only; no torch types; no C++ exceptions
 * cross the boundary.  Every function returns 0 on success or a VQ_ERR_* code, and
 * vq_last_error() then holds a message for the calling thread.
 *
 * Conventions
 *   - All matrices are fp32, row-major, contiguous.  z is (N, D): N rows of D consecutive floats,
 *     i.e. the reference's `inputs.view(-1, D)` (vector_quantizer.py:32) -- no permute.
 *   - E is the codebook (K, D) == `_embedding.weight` (vector_quantizer.py:15).
 *   - Device pointers unless the name says host.  All work is enqueued on `stream`
 *     (a cudaStream_t); nothing synchronises the host except the *_host entry points.
 *   - The library never keeps caller memory and never allocates device memory, except inside a
 *     vq_host_ctx (which owns its staging buffers).
 */
#ifndef B200VQ_H_
#define B200VQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQ_ABI_VERSION 18

/* error codes */
#define VQ_OK            0
#define VQ_ERR_ARG       1   /* bad shape / null pointer / misaligned pointer */
#define VQ_ERR_CUDA      2   /* a CUDA runtime or driver call failed */
#define VQ_ERR_NO_DEVICE 3   /* no sm_100 device: there is no CPU fallback */
#define VQ_ERR_WORKSPACE 4   /* workspace too small */

/* flags for vq_forward / vq_backward */
#define VQ_FLAG_ONEHOT     (1 << 0)  /* forward: also write the dense (N,K) one-hot (vector_quantizer.py:39-40) */
#define VQ_FLAG_TRAIN_VQ   (1 << 1)  /* backward: produce dE (vector_quantizer.py:47-50 `_train_vq`) */
#define VQ_FLAG_EXACT      (1 << 2)  /* forward: every distance on CUDA cores in the oracle's fp32 FMA-chain order instead of
                                        the tcgen05 path (default there: one TF32 screening pass + exact fp32 refine of the
                                        candidates, also bit-exact vs oracle/vq_oracle.c; 3xTF32 for the other shapes) */
#define VQ_FLAG_DEFER_STATS (1 << 3) /* forward: leave loss/perplexity to vq_finalize_stats (data parallel) */
#define VQ_FLAG_NO_QUANT   (1 << 4)  /* forward: indices/hist only; q_out, sse, loss are not produced */
#define VQ_FLAG_TC_1CTA    (1 << 6)  /* forward: single-CTA tensor kernel (M=128,N=128) even when the CTA-pair kernel applies */
#define VQ_FLAG_NO_FUSE    (1 << 7)  /* forward: keep argmin and the row epilogue as two kernels (default: fused into one) */
#define VQ_FLAG_STATE_READY (1 << 8) /* forward: vq_prepare_step already reset hist and the workspace counter for this call */
#define VQ_FLAG_NO_SCREEN  (1 << 9)  /* forward: 3xTF32 tensor kernels, never the screen (1xTF32) + exact-refine kernel */
#define VQ_FLAG_SCREEN     (1 << 10) /* forward: force the screen + exact-refine kernel wherever its shape constraints allow */
#define VQ_FLAG_ZERO_DE    (1 << 5)  /* backward: dE = gradient instead of dE += gradient (a memset on `stream`, or plain stores on
                                        the bucket path, which writes every element exactly once) */
#define VQ_FLAG_BWD_FLAT    (1 << 12) /* backward: force the flat kernel (one 16-byte red.global.add per element of dE) */
#define VQ_FLAG_BWD_PRIVATE (1 << 14) /* backward: force the shared-memory-private kernel where the shape allows */

typedef void* vq_stream_t;   /* cudaStream_t */

/* -- introspection ------------------------------------------------------------------------- */
int         vq_abi_version(void);
const char* vq_last_error(void);
/* 0 when the current CUDA device is sm_100 (B200); VQ_ERR_NO_DEVICE otherwise. */
int         vq_device_check(void);
/* Which forward path vq_forward would take: 1 = tcgen05 tensor path, 0 = exact CUDA-core path. */
int         vq_forward_uses_tensor_path(int64_t n_rows, int K, int D, int flags);
/* Kernel launches issued by this library since load (all streams); bench.py's gpu_launches. */
int64_t     vq_launch_count(void);

/* Per-kernel device timing for the roofline report: while enabled, every launch is bracketed by CUDA
 * events on its own stream.  vq_profile_read (after a device synchronise) returns the summed elapsed
 * time and launch count of kernel `kid`: 0 prepare, 1 tensor-core argmin, 2 exact argmin, 3 rows,
 * 4 backward, 5 finalize, 6 one-hot. */
void        vq_profile_enable(int on);
int         vq_profile_read(int kid, double* total_ms, int64_t* count);

/* Debug hook: in -DVQ_TRACE builds the fused forward kernel writes clock64 stamps (148 CTAs x 8 roles x 64
 * int64) into this device buffer; a no-op in production builds.  See tools/trace_fused.py. */
void        vq_debug_set_trace(long long* device_buffer);

/* -- codebook preparation (once per optimizer step: Adam changes E) ------------------------- */
/* e_norm2[k] = |E_k|^2 (vector_quantizer.py:35) as an fp32 FMA chain over d;
 * E_hi = tf32_rna(E), E_lo = tf32_rna(E - E_hi): the split operands of the 3xTF32 contraction.
 * E_hi / E_lo may be NULL when only the exact path will be used. */
int vq_prepare_codebook(const float* E, int K, int D,
                        float* e_norm2, float* E_hi, float* E_lo, vq_stream_t stream);

/* vq_prepare_codebook plus the per-step state reset in the same launch: zeroes hist (K), the completion
 * counter inside `workspace`, and -- when not NULL -- the dE accumulator (K*D).  Pass VQ_FLAG_STATE_READY to the
 * vq_forward that follows (and no VQ_FLAG_ZERO_DE to vq_backward) to skip their own resets. */
int vq_prepare_step(const float* E, int K, int D, float* e_norm2, float* E_hi, float* E_lo,
                    float* hist, void* workspace, size_t workspace_bytes, float* dE, vq_stream_t stream);

/* -- forward: vector_quantizer.py:29-58 ------------------------------------------------------ */
size_t vq_workspace_bytes(int64_t n_rows, int K, int D, int flags);

/* Outputs:
 *   q_out  (N,D)  fl(z + fl(E[idx]-z))            vector_quantizer.py:43,54
 *   idx    (N)    int32 code index, first minimum vector_quantizer.py:38
 *   onehot (N,K)  fp32 one-hot or NULL            vector_quantizer.py:39-40   (VQ_FLAG_ONEHOT)
 *   hist   (K)    fp32 usage counts               vector_quantizer.py:55 (mean(enc,0) * N)
 *   sse    (1)    sum (E[idx]-z)^2                vector_quantizer.py:46-50 numerator
 *   loss   (1)    (1+beta) * sse / (N*D)          vector_quantizer.py:52
 *   perplexity(1) exp(-sum p log(p+1e-10))        vector_quantizer.py:56
 * loss/perplexity are untouched under VQ_FLAG_DEFER_STATS. */
int vq_forward(const float* z, const float* E, const float* e_norm2,
               const float* E_hi, const float* E_lo,
               int64_t n_rows, int K, int D, float beta, int flags,
               float* q_out, int32_t* idx, float* onehot, float* hist, float* sse,
               float* loss, float* perplexity,
               void* workspace, size_t workspace_bytes, vq_stream_t stream);

/* One entry for "prepare + forward" (the call a training step makes: Adam has just changed E): vq_prepare_step followed
 * by vq_forward, chained by programmatic dependent launch.  The scratch buffers e_norm2 (K), E_hi (K,D), E_lo (K,D)
 * belong to the caller; E_lo may be NULL for shapes the screen + refine kernel takes (K % 256 == 0, D in
 * {32,64,96,128,192,256}), both for shapes that run the exact path.  dE_zero: NULL, or the (K,D) accumulator the
 * backward will add into -- zeroed by the prepare launch, so that vq_backward needs no VQ_FLAG_ZERO_DE memset.
 * Outputs and flags as for vq_forward (VQ_FLAG_STATE_READY is implied). */
int vq_step_forward(const float* z, const float* E, int64_t n_rows, int K, int D, float beta, int flags,
                    float* e_norm2, float* E_hi, float* E_lo, float* dE_zero,
                    float* q_out, int32_t* idx, float* onehot, float* hist, float* sse,
                    float* loss, float* perplexity,
                    void* workspace, size_t workspace_bytes, vq_stream_t stream);

/* loss / perplexity from (all-reduced) usage counts and squared error over n_rows_global rows. */
int vq_finalize_stats(const float* hist, const float* sse, int64_t n_rows_global, int K, int D,
                      float beta, float* loss, float* perplexity, vq_stream_t stream);

/* (N,K) one-hot from indices (vector_quantizer.py:39-40), for callers that ask for it late. */
int vq_onehot(const int32_t* idx, int64_t n_rows, int K, float* onehot, vq_stream_t stream);

/* -- backward: autograd of vector_quantizer.py:46-54 ----------------------------------------- */
/*   dz[n,:]       = g_q[n,:] + g_loss*beta*2*(z - E[idx])/(n_rows_dz*D)     g_q may be NULL (zero)
 *   dE[idx[n],:] += g_loss*2*(E[idx] - z)/(n_rows_dE*D)   when VQ_FLAG_TRAIN_VQ and dE != NULL
 * dE is accumulated into: zero it first or pass VQ_FLAG_ZERO_DE (under data parallelism the caller
 * all-reduces it afterwards).  g_loss is a device scalar (NULL means 1).  The straight-through output sends
 * no gradient to E.  dz == NULL with VQ_FLAG_TRAIN_VQ computes the codebook gradient only. */
int vq_backward(const float* g_q, const float* g_loss, const float* z, const float* E,
                const int32_t* idx, int64_t n_rows, int64_t n_rows_dz, int64_t n_rows_dE,
                int K, int D, float beta, int flags, float* dz, float* dE, vq_stream_t stream);
/* vq_backward launched RIGHT BEHIND vq_step_forward on the same stream and workspace (nothing in between; `forward_flags`
 * = the flags that forward got): on the screen + refine path the kernel starts as soon as the forward's last CTA has
 * raised the workspace's ready word -- everything the backward reads is complete then -- and overlaps the forward's
 * serial statistics tail (loss / perplexity, ~5 us), which it does not need.  It orders itself behind the forward's
 * completion before it exits, and falls back to the full dependency when the word does not show up.  Elsewhere (other
 * shapes, VQ_FLAG_ZERO_DE, dz == NULL, the private path) it is vq_backward. */
int vq_step_backward(const float* g_q, const float* g_loss, const float* z, const float* E,
                     const int32_t* idx, int64_t n_rows, int64_t n_rows_dz, int64_t n_rows_dE,
                     int K, int D, float beta, int flags, float* dz, float* dE,
                     void* workspace, size_t workspace_bytes, int forward_flags, vq_stream_t stream);
/* Which kernel the backward takes for dE (16-byte aligned pointers assumed): 0 flat (one red.global.add per element),
 * 2 private (per-CTA copy of dE in shared memory, flushed once; N >> K and K*D small; dz may be NULL: codebook gradient
 * only), 3 replicated -- vq_step_backward only, which has scratch (the dead tail of the forward's workspace): the flat
 * kernel with its reds spread over R zeroed copies of dE (R * K * D/4 ~ 256 k addresses, R <= 32) and a small launch
 * that folds them into dE; taken when K * D/4 < 128 k addresses and N >= max(256 k, 256 K), where the flat scatter is
 * bound by reds queueing on the same L2 addresses.  vq_backward (no scratch) takes 2 or 0 there. */
int vq_backward_path(int64_t n_rows, int K, int D, int flags);

/* -- the consumer of `encodings` as an index gather ------------------------------------ */
/* LocationModule.fc_1 (location_model.py:10,21) applied to flatten(one_hot(B,T,K)) (train_location.py:74-75):
 *   y[b,:] = bias + sum_t Wt[t*K + idx[b,t], :]        Wt = fc_1.weight transposed, (T*K, O) row-major, O % 4 == 0
 * and the dense gradient of its weight, dWt[t*K + idx[b,t], :] += g[b,:] (dWt zeroed by the caller). */
int vq_gather_sum_rows(const int32_t* idx, const float* Wt, const float* bias_or_null, float* y,
                       int B, int T, int K, int O, vq_stream_t stream);
int vq_scatter_add_rows(const int32_t* idx, const float* g, float* dWt, int B, int T, int K, int O, vq_stream_t stream);

/* -- the time-mean variant in front of the quantizer (first half) ------------------------------ */
/* convolutional_vq_vae.py:96-97 (`encoder_average_pooling`): z = mean over time of the `_pre_vq_conv` output, x (B, D, T)
 * -> z (B, D, 1), which the quantizer then sees as B rows.  z[row] = sum_t x[row, t] / T with rows = B*D, in a fixed
 * summation order; launched with the programmatic-launch attribute so that the forward behind it chains without a gap.
 * The backward spreads dz[row] / T over the T positions. */
int vq_time_mean(const float* x, int64_t rows, int T, float* z, vq_stream_t stream);
int vq_time_mean_backward(const float* dz, int64_t rows, int T, float* dx, vq_stream_t stream);

/* -- Jitter on the quantizer's output (modules/jitter.py:47-70) ------------------------- */
/* In place on q (rows = B*D, T): column t becomes the ORIGINAL column src[t] (src[t] in {t-1, t, t+1}; the host
 * draws src with the reference's np.random calls).  Backward zeroes the gradient of the replaced columns. */
int vq_jitter_apply(float* q, const int32_t* src, int64_t rows, int T, vq_stream_t stream);
int vq_jitter_backward(float* g, const int32_t* src, int64_t rows, int T, vq_stream_t stream);

/* -- data parallel: sum all-reduce of the packed step buffer over NVLink peer memory -------------------------------- */
/* Rows shard across ranks with no data-path collective; per step ONE packed buffer
 *   [ dE (K*D) | usage histogram (K) | squared error (1) ]
 * is summed over the ranks, in rank order (bit-identical on every rank).  Transport: every rank owns two symmetric
 * RECEIVE buffers (alternating between calls) of vq_dp_recv_lines(world, n_floats) 16-byte lines, zero-initialised
 * and mapped into every peer (torch symmetric memory / CUDA IPC); data travels as lines {d0, seq, d1, seq}, so a line
 * whose flag words carry the call's sequence number is complete -- no barrier.  Below 8 ranks every rank stores its
 * lines into every rank (one multimem.st per line with the NVLS multicast mappings), from 8 ranks on reduce-scatter +
 * all-gather through the same buffers (B200VQ_AR_ALGO=1|2 forces either).  The sequence number lives in device
 * memory and is advanced by the kernel, so the calls can be captured in a CUDA graph and replayed.  Waits are bounded
 * (spin_limit polls, 0 = default of a few seconds): on expiry the kernel raises the error word and drains.
 * The reference has no counterpart (single process); under DDP it would be the all-reduce of `_embedding.weight.grad`. */
typedef struct vq_dp_ctx vq_dp_ctx;
int64_t vq_dp_recv_lines(int world, int64_t n_floats);
/* recv0 / recv1: host arrays of `world` device pointers, entry p = rank p's receive buffer (parity 0 / 1) as mapped
 * into THIS process; multicast0 / multicast1: NVLS multicast mappings of the same buffers or NULL. */
int  vq_dp_create(const void* const* recv0, const void* const* recv1, void* multicast0, void* multicast1,
                  int world, int rank, int64_t n_floats, uint32_t spin_limit, vq_dp_ctx** out);
void vq_dp_destroy(vq_dp_ctx* ctx);
/* out[i] = sum over ranks of payload[i], i < n_floats (payload: written by earlier work on `stream`). */
int  vq_dp_allreduce(vq_dp_ctx* ctx, const float* payload, float* out, vq_stream_t stream);
/* The data-parallel backward: vq_step_backward and the exchange of the packed step buffer as ONE launch.  Every CTA of the
 * backward takes a completion ticket when its share of the codebook gradient is out; the last 128 CTAs to finish become
 * the exchange (push / collect / sum, as vq_dp_allreduce) the moment the gradient is complete -- no second kernel, whose
 * launch boundary and serial round trips cost ~7 us per step even with nothing to send.  `packed` = [dE (K*D) | hist (K) |
 * sse] with the forward's statistics already in place and the dE slot zeroed (vq_step_forward's dE_zero); out = the summed
 * buffer, complete when the call's kernel is.  Falls back to vq_step_backward + vq_dp_allreduce for shapes the 16-byte
 * flat pass does not cover (and under B200VQ_DP_TAIL=0). */
int  vq_step_backward_dp(const float* g_q, const float* g_loss, const float* z, const float* E, const int32_t* idx,
                         int64_t n_rows, int64_t n_rows_dz, int64_t n_rows_dE, int K, int D, float beta, int flags,
                         float* dz, float* packed, vq_dp_ctx* ctx, float* out,
                         void* workspace, size_t workspace_bytes, int forward_flags, vq_stream_t stream);
/* Overlapped form: the exchange is ordered behind everything enqueued on `stream` so far, but runs on a stream of the
 * context's own, so `stream` carries on at once -- the next step's codebook preparation and forward overlap the NVLink
 * transfer and absorb the ranks' skew (in training the encoder's backward does, as under DDP's bucketed all-reduce).
 * vq_dp_wait makes `stream` wait for every exchange started so far except the keep_in_flight most recent ones
 * (exchanges complete in the order they were started): keep_in_flight = S - 1 (< 32) at the top of a step whose
 * payload / out buffers rotate over S sets, 0 before the results are read, before vq_dp_destroy and before a stream
 * capture ends.  Fork and join are event edges, so a stream capture records them; a captured graph of S steps over S
 * buffer sets needs no join but the final one, which leaves every programmatic-launch edge of the step chain intact. */
int  vq_dp_allreduce_start(vq_dp_ctx* ctx, const float* payload, float* out, vq_stream_t stream);
int  vq_dp_wait(vq_dp_ctx* ctx, int keep_in_flight, vq_stream_t stream);
/* Synchronises `stream`; calls completed and the error word (bit 0: a wait expired). */
int  vq_dp_status(vq_dp_ctx* ctx, uint32_t* calls_done, uint32_t* error_word, vq_stream_t stream);
/* Test hook: `world` emulated ranks on ONE GPU in a single cooperative launch (blockIdx.y = rank), `rounds` calls back
 * to back.  payloads / outs: host arrays of `world` device pointers.  error_word: OR of the ranks' error words,
 * bit 8 = a call counter that did not advance once per round. */
int  vq_dp_emulate(int world, int two_step, int64_t n_floats, const float* const* payloads, float* const* outs,
                   int rounds, uint32_t spin_limit, uint32_t* error_word, vq_stream_t stream);

/* -- host-buffer entry points (what a non-torch caller binds; used for the end-to-end figure) - */
typedef struct vq_host_ctx vq_host_ctx;

/* Owns device staging for up to max_rows rows, two copy/compute lanes and pinned result slots. */
int  vq_host_ctx_create(int64_t max_rows, int K, int D, vq_host_ctx** out);
void vq_host_ctx_destroy(vq_host_ctx* ctx);
/* Upload the codebook (host pointer) and prepare it. */
int  vq_host_set_codebook(vq_host_ctx* ctx, const float* E_host);
/* One forward+backward step on host buffers: copies z (and g_q if not NULL; NULL means ones, the
 * `(loss + quantized.sum()).backward()` workload) to the device, runs prepare/forward/backward,
 * and copies back loss, perplexity and -- when the pointers are not NULL -- idx (N), q_out (N,D),
 * dz (N,D), dE (K,D).  Host pointers should be pinned for full PCIe speed.  Asynchronous: the
 * step is enqueued on lane (step_no % 2); results are valid after vq_host_wait(ctx, lane). */
int  vq_host_step_async(vq_host_ctx* ctx, int lane, const float* z_host, const float* gq_host,
                        int64_t n_rows, int64_t n_rows_dE /* 0: n_rows; data parallel: global rows */,
                        float beta, int flags,
                        float* loss_host, float* perplexity_host, int32_t* idx_host,
                        float* q_host, float* dz_host, float* dE_host);
int  vq_host_wait(vq_host_ctx* ctx, int lane);
/* Device-side stopwatch spanning both lanes (CUDA events on the lanes' own streams). */
int  vq_host_timer_start(vq_host_ctx* ctx);
int  vq_host_timer_stop_ms(vq_host_ctx* ctx, float* ms);
/* Device-side handles of a lane, for callers that must all-reduce dE (K*D floats) across ranks on the
 * lane's stream before reading it back: the stream, the dE buffer and the hist buffer (K floats). */
int  vq_host_lane_buffers(vq_host_ctx* ctx, int lane, void** stream, float** dE, float** hist);

#ifdef __cplusplus
}
#endif
#endif /* B200VQ_H_ */
