// b200vq.cu -- C ABI (include/b200vq.h) of the B200-native VectorQuantizer hot path.
// Host launchers only; kernels live in kernels_simt.cuh / kernels_tc.cuh.  No torch, no CPU fallback.
#include "../../include/b200vq.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "kernels_simt.cuh"
#include "kernels_tc.cuh"
#include "kernels_screen.cuh"
#include "kernels_bwd.cuh"
#include "kernels_dp.cuh"

using namespace b200vq;

namespace {

thread_local char g_err[512] = "";
long long* g_trace_buf = nullptr;   // VQ_TRACE builds: device buffer handed to the fused kernel
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) return fail(VQ_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

#define LAUNCH_CHECK(name)                                                                         \
    do {                                                                                           \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess) return fail(VQ_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                        \
    } while (0)

// ---- optional per-kernel timing (bench.py's roofline leg): CUDA events on the launching stream ------
enum KernelId { KID_PREP = 0, KID_ARGMIN_TC, KID_ARGMIN_SIMT, KID_ROWS, KID_BACKWARD, KID_FINALIZE, KID_ONEHOT, KID_ALLREDUCE, KID_COUNT };
struct ProfSlot { cudaEvent_t a, b; int kid; };
bool g_prof_on = false;
std::vector<ProfSlot> g_prof_slots;
size_t g_prof_used = 0;
std::mutex g_prof_mu;

struct ProfScope {
    cudaStream_t st;
    ProfSlot* slot = nullptr;
    ProfScope(int kid, cudaStream_t s) : st(s) {
        if (!g_prof_on) return;
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (g_prof_used == g_prof_slots.size()) {
            ProfSlot ns{};
            if (cudaEventCreate(&ns.a) != cudaSuccess || cudaEventCreate(&ns.b) != cudaSuccess) return;
            g_prof_slots.push_back(ns);
        }
        slot = &g_prof_slots[g_prof_used++];
        slot->kid = kid;
        cudaEventRecord(slot->a, st);
    }
    ~ProfScope() {
        if (slot != nullptr) cudaEventRecord(slot->b, st);
    }
};

// launch with the programmatic-stream-serialization attribute (PDL): the kernel may start while its predecessor in
// the stream is finishing; it orders itself with griddepcontrol.wait
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cudaFuncAttributeMaxDynamicSharedMemorySize applies to the CURRENT device only: remember per kernel instantiation
// (one static DeviceOnce per call site) which devices have been configured.  Thread safe.
struct DeviceOnce {
    std::atomic<unsigned long long> done{0};     // bit d = device d configured (devices >= 64: set it every time)
    template <typename Kernel>
    cudaError_t max_smem(Kernel* kernel, int bytes) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 64 && (done.load(std::memory_order_acquire) >> dev) & 1ull) return cudaSuccess;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
        return e;
    }
};

// SM count of the current device (grids of the persistent kernels are sized from it, not from a constant)
int sm_count() {
    static thread_local int cached_dev = -1, cached = kNumSMs;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMs;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        cached_dev = dev;
    }
    return cached;
}

// ---- device gate: this library only runs on sm_100 ------------------------------------------------
int check_device() {
    static thread_local int cached_dev = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(VQ_ERR_NO_DEVICE, "no CUDA device: %s (b200vq has no CPU fallback)", cudaGetErrorString(e));
    if (dev == cached_dev) return VQ_OK;
    int major = 0, minor = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) return fail(VQ_ERR_NO_DEVICE, "device %d is sm_%d%d; b200vq kernels are built for sm_100a only", dev, major, minor);
    cached_dev = dev;
    return VQ_OK;
}

// ---- TMA descriptors ---------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// (rows, D) fp32 row-major matrix; box = 32 floats (128 B) x 128 rows, 128-byte swizzle.
// A descriptor depends on (address, rows, D) only and encoding one costs about a microsecond of host time, so the
// last few are kept per thread (training loops present the same codebook and the same few activation buffers).
struct TmapSlot { const float* base; long long rows; int D; unsigned long long stamp; CUtensorMap map; };
int make_tmap(CUtensorMap* out, const float* base, long long rows, int D) {
    constexpr int NSLOT = 16;
    static thread_local TmapSlot cache[NSLOT] = {};
    static thread_local unsigned long long tick = 0;
    int victim = 0;
    for (int i = 0; i < NSLOT; ++i) {
        if (cache[i].stamp != 0 && cache[i].base == base && cache[i].rows == rows && cache[i].D == D) {
            cache[i].stamp = ++tick;
            *out = cache[i].map;
            return VQ_OK;
        }
        if (cache[i].stamp < cache[victim].stamp) victim = i;
    }
    auto fn = get_encode_fn();
    if (fn == nullptr) return fail(VQ_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(D) * sizeof(float)};
    cuuint32_t box[2] = {TC_SLAB_FLOATS, TC_ROWS};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VQ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    cache[victim] = TmapSlot{base, rows, D, ++tick, *out};
    return VQ_OK;
}

// ---- path selection ------------------------------------------------------------------------------------
bool tensor_path_ok(long long N, int K, int D, int flags, const float* z = nullptr, const float* ehi = nullptr,
                    const float* elo = nullptr, bool check_ptrs = false) {
    if (flags & VQ_FLAG_EXACT) return false;
    if (N < 1 || N >= (1ll << 31) - TC_ROWS) return false;
    if (D % TC_SLAB_FLOATS != 0 || D < 32 || D > 256) return false;
    if (K % TC_CODES != 0 || K < TC_CODES) return false;
    // D in (128, 256] (and D = 160, 224) only fits the screen+refine kernel (tf32(z) alone in shared memory)
    const bool screen_shape = (K % TC2_CODES == 0) && K <= (1 << 20) && (D <= 128 || D == 192 || D == 256) && !(flags & VQ_FLAG_NO_SCREEN) &&
                              !(flags & VQ_FLAG_NO_FUSE) && !(flags & VQ_FLAG_TC_1CTA);
    if (D > 128 && !screen_shape) return false;
    if (check_ptrs && (ehi == nullptr || elo == nullptr || !aligned16(z) || !aligned16(ehi) || !aligned16(elo))) return false;
    return true;
}

// shapes / alignments the screen + refine kernel takes (E_hi is not needed in self-prepared mode)
bool screen_path_ok(long long N, int K, int D, int flags, const float* z, const float* E, const float* q_out, const float* onehot) {
    if (flags & (VQ_FLAG_EXACT | VQ_FLAG_NO_SCREEN | VQ_FLAG_NO_FUSE | VQ_FLAG_TC_1CTA)) return false;
    if (N < 1 || N >= (1ll << 31) - TC_ROWS) return false;
    if (!(D == 32 || D == 64 || D == 96 || D == 128 || D == 192 || D == 256)) return false;
    if (K % TC2_CODES != 0 || K < TC2_CODES || K > (1 << 20)) return false;   // (row, code) pairs pack the code into 20 bits
    const bool quant = (flags & VQ_FLAG_NO_QUANT) == 0, want_onehot = (flags & VQ_FLAG_ONEHOT) != 0;
    return aligned16(z) && aligned16(E) && (!quant || aligned16(q_out)) && (!want_onehot || aligned16(onehot));
}

// Screen + refine with the row epilogue as a SECOND kernel (quantize_rows_kernel behind an indices-only screen kernel):
// an experiment, OFF unless B200VQ_SPLIT_ROWS=1.  The idea was that with few code tiles per item (small K) the four worker
// warps set the pace; the per-role trace (tools/trace_fused.py, N = 1M, K = 512, D = 64) shows they do not -- a worker group
// is done with an item 4.5 - 6.5 us after it gets it and the two groups alternate, while the epilogue publishes an item
// every 5.2 - 6.2 us (2 x 1.45 us of scanning + ~1.1 us of per-item merge / handoff / barriers).  Measured: 393 vs 404 us
// at (512, 64), slower everywhere else (z is read twice).
bool screen_split_rows(long long N, int K, int D, int flags) {
    (void)N; (void)K; (void)D;
    if ((flags & (VQ_FLAG_ONEHOT | VQ_FLAG_NO_QUANT)) != 0) return false;
    static const bool forced = [] { const char* e = getenv("B200VQ_SPLIT_ROWS"); return e != nullptr && atoi(e) != 0; }();
    return forced;
}

// number of codebook splits per row tile: fill the 148 SMs when there are few row tiles.
// cost(s) ~ waves(s) * (code tiles per CTA + fixed per-CTA overhead of ~1 tile)
int choose_splits(long long row_tiles, int code_tiles, int max_splits, int slots = kNumSMs, double overhead = 1.0) {
    int best_s = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= code_tiles && s <= max_splits; ++s) {
        if (code_tiles % s != 0) continue;
        const long long ctas = row_tiles * s;
        const long long waves = (ctas + slots - 1) / slots;
        const double cost = static_cast<double>(waves) * (static_cast<double>(code_tiles / s) + overhead);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best_s = s;
        }
    }
    return best_s;
}

struct WsLayout {
    size_t partials_off, counter_off, keys_off, total;
    int rows_grid;
};

// rows per CTA group of the streaming kernels: 256 threads, one 16-byte (or 4-byte) element each
int rows_per_group(int D, bool vec) {
    const int dv = vec ? D / 4 : D;
    int R = (vec ? 1024 : 256) / (dv < 1 ? 1 : dv);     // vector path: four elements in flight per thread
    if (R < 1) R = 1;
    if (R > ROWS_MAX_R) R = ROWS_MAX_R;
    return R;
}

constexpr int kMaxRowsGrid = kNumSMs * 16;   // persistent-ish: at most 16 CTAs of 256 threads per SM

int rows_grid_for(long long N, int R) {
    long long g = (N + R - 1) / R;
    if (g > kMaxRowsGrid) g = kMaxRowsGrid;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

WsLayout ws_layout(long long N) {
    WsLayout w;
    w.rows_grid = 0;
    w.partials_off = 0;
    w.counter_off = static_cast<size_t>(kMaxRowsGrid) * sizeof(double);
    w.keys_off = w.counter_off + 256;
    // the per-row key buffer (unfused paths) and the screen kernel's spill lists (fused path) share the tail
    const size_t keys_bytes = static_cast<size_t>(N) * sizeof(unsigned long long);
    const size_t spill_bytes = static_cast<size_t>(kNumSMs) * 4 * SC_SPILL * sizeof(int4);
    w.total = w.keys_off + (keys_bytes > spill_bytes ? keys_bytes : spill_bytes);
    return w;
}

template <int NSLAB, int NSTAGE>
int launch_tc(const CUtensorMap& tz, const CUtensorMap& thi, const CUtensorMap& tlo, const float* e_norm2, long long N,
              int K, int codes_per_split, int splits, int* idx, unsigned long long* keys, float* hist,
              unsigned int* counter, cudaStream_t st) {
    constexpr int smem = tc_smem_bytes(NSLAB, NSTAGE);
    static DeviceOnce once;
    CUDA_TRY(once.max_smem(argmin_tc_kernel<NSLAB, NSTAGE>, smem));
    dim3 grid(static_cast<unsigned>((N + TC_ROWS - 1) / TC_ROWS), static_cast<unsigned>(splits));
    ProfScope prof(KID_ARGMIN_TC, st);
    argmin_tc_kernel<NSLAB, NSTAGE><<<grid, TC_THREADS, smem, st>>>(tz, thi, tlo, e_norm2, N, K, codes_per_split, idx,
                                                                    keys, hist, counter);
    LAUNCH_CHECK("argmin_tc_kernel");
    return VQ_OK;
}

template <int NSLAB, int NSTAGE, int ZBUF>
int launch_tc2(const CUtensorMap& tz, const CUtensorMap& thi, const CUtensorMap& tlo, const float* e_norm2, long long N,
               int K, int codes_per_split, int splits, int* idx, unsigned long long* keys, float* hist,
               unsigned int* counter, const FusedRowArgs* fused, bool state_ready, cudaStream_t st) {
    constexpr int smem = tc2_smem_bytes(NSLAB, NSTAGE, ZBUF);
    static_assert(smem <= 232448, "CTA-pair kernel exceeds 227 KB of shared memory");
    static DeviceOnce once_plain, once_fused;
    CUDA_TRY(once_plain.max_smem(argmin_tc2_kernel<NSLAB, NSTAGE, ZBUF, false>, smem));
    CUDA_TRY(once_fused.max_smem(argmin_tc2_kernel<NSLAB, NSTAGE, ZBUF, true>, smem));
    const long long row_tiles = (N + TC_ROWS - 1) / TC_ROWS;
    const long long n_items = ((row_tiles + 1) / 2) * splits;          // (row-tile pair, codebook split)
    const int pairs = static_cast<int>(n_items < kNumSMs / 2 ? n_items : kNumSMs / 2);
    if (fused != nullptr) {
        if (!state_ready) {
            zero_state_kernel<<<1, 256, 0, st>>>(hist, K, counter);
            LAUNCH_CHECK("zero_state_kernel");
        }
        ProfScope prof(KID_ARGMIN_TC, st);
        cudaError_t e = launch_pdl(argmin_tc2_kernel<NSLAB, NSTAGE, ZBUF, true>, dim3(2 * pairs), dim3(TC2_THREADS_FUSED), smem, st,
                                   tz, thi, tlo, e_norm2, N, K, codes_per_split, splits, static_cast<int>(n_items), idx, keys, hist,
                                   counter, *fused);
        if (e != cudaSuccess) return fail(VQ_ERR_CUDA, "launch of argmin_tc2_kernel<fused> failed: %s", cudaGetErrorString(e));
        LAUNCH_CHECK("argmin_tc2_kernel<fused>");
    } else {
        FusedRowArgs none{};
        ProfScope prof(KID_ARGMIN_TC, st);
        argmin_tc2_kernel<NSLAB, NSTAGE, ZBUF, false><<<2 * pairs, TC2_THREADS, smem, st>>>(
            tz, thi, tlo, e_norm2, N, K, codes_per_split, splits, static_cast<int>(n_items), idx, keys, hist, counter, none);
        LAUNCH_CHECK("argmin_tc2_kernel");
    }
    return VQ_OK;
}

template <int NSLAB, int NSTAGE, int ZBUF>
int launch_screen(const CUtensorMap& tz, const CUtensorMap& thi, const float* e_norm2, long long N, int K, int* idx,
                  float* hist, unsigned int* counter, const FusedRowArgs& fused, bool state_ready, cudaStream_t st) {
    constexpr int smem = sc_smem_bytes(NSLAB, NSTAGE, ZBUF);
    static_assert(smem <= 232448, "screen kernel exceeds 227 KB of shared memory");
    static DeviceOnce once;
    CUDA_TRY(once.max_smem(vq_screen_kernel<NSLAB, NSTAGE, ZBUF>, smem));
    const long long row_tiles = (N + TC_ROWS - 1) / TC_ROWS;
    const long long n_items = (row_tiles + 1) / 2;
    const int pairs = static_cast<int>(n_items < kNumSMs / 2 ? n_items : kNumSMs / 2);
    if (!state_ready) {
        zero_state_kernel<<<1, 256, 0, st>>>(hist, K, counter);
        LAUNCH_CHECK("zero_state_kernel");
    }
    ProfScope prof(KID_ARGMIN_TC, st);
    cudaError_t e = launch_pdl(vq_screen_kernel<NSLAB, NSTAGE, ZBUF>, dim3(2 * pairs), dim3(SC_THREADS), smem, st, tz, thi, e_norm2, N,
                               K, static_cast<int>(n_items), idx, fused);
    if (e != cudaSuccess) return fail(VQ_ERR_CUDA, "launch of vq_screen_kernel failed: %s", cudaGetErrorString(e));
    LAUNCH_CHECK("vq_screen_kernel");
    return VQ_OK;
}

bool screen_enabled() {
    static const bool on = [] { const char* e = getenv("B200VQ_SCREEN"); return !(e != nullptr && e[0] == '0'); }();
    return on;
}

}  // namespace

// =========================================================================================================
// C ABI
// =========================================================================================================
extern "C" {

int vq_abi_version(void) { return VQ_ABI_VERSION; }
const char* vq_last_error(void) { return g_err; }
int vq_device_check(void) { return check_device(); }
int64_t vq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void vq_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    if (on) g_prof_used = 0;
}

// total elapsed ms and launch count recorded for kernel `kid` since vq_profile_enable(1); call after a
// device synchronise.  kid: 0 prep, 1 argmin_tc, 2 argmin_simt, 3 rows, 4 backward, 5 finalize, 6 onehot.
int vq_profile_read(int kid, double* total_ms, int64_t* count) {
    if (total_ms == nullptr || count == nullptr || kid < 0 || kid >= KID_COUNT) return fail(VQ_ERR_ARG, "vq_profile_read: bad argument");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double t = 0.0;
    int64_t n = 0;
    for (size_t i = 0; i < g_prof_used; ++i) {
        if (g_prof_slots[i].kid != kid) continue;
        float ms = 0.f;
        cudaError_t e = cudaEventElapsedTime(&ms, g_prof_slots[i].a, g_prof_slots[i].b);
        if (e != cudaSuccess) return fail(VQ_ERR_CUDA, "vq_profile_read: %s (synchronise the device first)", cudaGetErrorString(e));
        t += ms;
        ++n;
    }
    *total_ms = t;
    *count = n;
    return VQ_OK;
}

// Debug builds (-DVQ_TRACE) only: device buffer of 148*8*64 int64 that receives per-role clock64 stamps.
void vq_debug_set_trace(long long* device_buffer) { g_trace_buf = device_buffer; }

int vq_forward_uses_tensor_path(int64_t n_rows, int K, int D, int flags) {
    return tensor_path_ok(n_rows, K, D, flags) ? 1 : 0;
}

size_t vq_workspace_bytes(int64_t n_rows, int K, int D, int flags) {
    (void)K; (void)D; (void)flags;
    if (n_rows < 0) n_rows = 0;
    return ws_layout(n_rows).total;
}

static int prepare_impl(const float* E, int K, int D, float* e_norm2, float* E_hi, float* E_lo, float* hist,
                        unsigned int* counter, float* dE, cudaStream_t st) {
    if (int rc = check_device()) return rc;
    if (E == nullptr || e_norm2 == nullptr || K < 1 || D < 1) return fail(VQ_ERR_ARG, "vq_prepare: bad argument (K=%d D=%d)", K, D);
    if (E_hi == nullptr && E_lo != nullptr) return fail(VQ_ERR_ARG, "vq_prepare: E_lo without E_hi");
    ProfScope prof(KID_PREP, st);
    const cudaError_t e = launch_pdl(prep_codebook_kernel, dim3((K + 7) / 8), dim3(256), 0, st, E, K, D, e_norm2, E_hi, E_lo, hist, counter, dE);
    if (e != cudaSuccess) return fail(VQ_ERR_CUDA, "launch of prep_codebook_kernel failed: %s", cudaGetErrorString(e));
    LAUNCH_CHECK("prep_codebook_kernel");
    return VQ_OK;
}

int vq_prepare_codebook(const float* E, int K, int D, float* e_norm2, float* E_hi, float* E_lo, vq_stream_t stream) {
    return prepare_impl(E, K, D, e_norm2, E_hi, E_lo, nullptr, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int vq_prepare_step(const float* E, int K, int D, float* e_norm2, float* E_hi, float* E_lo, float* hist, void* workspace,
                    size_t workspace_bytes, float* dE, vq_stream_t stream) {
    if (hist == nullptr || workspace == nullptr) return fail(VQ_ERR_ARG, "vq_prepare_step: hist / workspace is NULL");
    const WsLayout w = ws_layout(0);
    if (workspace_bytes < w.total) return fail(VQ_ERR_WORKSPACE, "vq_prepare_step: workspace %zu B < %zu B", workspace_bytes, w.total);
    unsigned int* counter = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) + w.counter_off);
    return prepare_impl(E, K, D, e_norm2, E_hi, E_lo, hist, counter, dE, static_cast<cudaStream_t>(stream));
}

int vq_forward(const float* z, const float* E, const float* e_norm2, const float* E_hi, const float* E_lo,
               int64_t n_rows, int K, int D, float beta, int flags, float* q_out, int32_t* idx, float* onehot,
               float* hist, float* sse, float* loss, float* perplexity, void* workspace, size_t workspace_bytes,
               vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    const long long N = n_rows;
    const bool want_onehot = (flags & VQ_FLAG_ONEHOT) != 0;
    const bool quant = (flags & VQ_FLAG_NO_QUANT) == 0;
    const bool defer = (flags & VQ_FLAG_DEFER_STATS) != 0;
    if (K < 1 || D < 1 || N < 0) return fail(VQ_ERR_ARG, "vq_forward: bad shape N=%lld K=%d D=%d", N, K, D);
    if (z == nullptr || E == nullptr || e_norm2 == nullptr || idx == nullptr || hist == nullptr || sse == nullptr)
        return fail(VQ_ERR_ARG, "vq_forward: null pointer among z/E/e_norm2/idx/hist/sse");
    if (quant && q_out == nullptr) return fail(VQ_ERR_ARG, "vq_forward: q_out is NULL without VQ_FLAG_NO_QUANT");
    if (want_onehot && onehot == nullptr) return fail(VQ_ERR_ARG, "vq_forward: VQ_FLAG_ONEHOT with onehot == NULL");
    if (!defer && (perplexity == nullptr || (quant && loss == nullptr))) return fail(VQ_ERR_ARG, "vq_forward: loss/perplexity NULL without VQ_FLAG_DEFER_STATS");
    const WsLayout w = ws_layout(N);
    if (workspace == nullptr || workspace_bytes < w.total)
        return fail(VQ_ERR_WORKSPACE, "vq_forward: workspace %zu B < required %zu B", workspace_bytes, w.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    double* partials = reinterpret_cast<double*>(ws + w.partials_off);
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws + w.counter_off);
    unsigned long long* keys_buf = reinterpret_cast<unsigned long long*>(ws + w.keys_off);

    if (N == 0) {   // empty batch: statistics of nothing
        CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(float) * K, st));
        CUDA_TRY(cudaMemsetAsync(sse, 0, sizeof(float), st));
        return VQ_OK;
    }

    // ---- 1. argmin ----------------------------------------------------------------------------------------
    unsigned long long* keys = nullptr;
    // ---- screen + refine (one TF32 pass, exact fp32 refine of the candidates): the default fused forward ----
    // Indices are bit-identical to the fp32 oracle, and it is the faster kernel everywhere on B200 (RIR-256 with the
    // dense one-hot: 69.0 vs 69.6 us per step; 1.5x at D = 64 and 2x at D = 128 without the one-hot); D > 128 only
    // fits this kernel.  VQ_FLAG_NO_SCREEN / B200VQ_SCREEN=0 select the 3xTF32 kernel, VQ_FLAG_SCREEN forces this one.
    const bool screen = screen_path_ok(N, K, D, flags, z, E, q_out, onehot) && E_hi != nullptr && aligned16(E_hi) &&
                        ((flags & VQ_FLAG_SCREEN) || screen_enabled());
    bool rows_done = false;
    if (screen) {
        const bool split = screen_split_rows(N, K, D, flags);
        FusedRowArgs fr{};
        fr.z = z; fr.E = E; fr.q_out = (quant && !split) ? q_out : nullptr; fr.onehot = want_onehot ? onehot : nullptr; fr.hist = hist;
        fr.partials = partials; fr.counter = counter; fr.sse_out = sse; fr.loss = loss; fr.perplexity = perplexity;
        fr.beta = beta; fr.finalize = defer ? 0 : 1; fr.trace = g_trace_buf;
        fr.rows_later = split ? 1 : 0;
        fr.spill = reinterpret_cast<int4*>(keys_buf);   // 16-byte aligned: keys_off is a multiple of 256
        static const int evict_first = [] { const char* e = getenv("B200VQ_ONEHOT_EVICT_FIRST"); return e != nullptr && e[0] == '0' ? 0 : 1; }();
        fr.onehot_evict_first = evict_first;
        CUtensorMap tz, thi;
        if (int rc = make_tmap(&tz, z, N, D)) return rc;
        if (int rc = make_tmap(&thi, E_hi, K, D)) return rc;
        const bool ready = (flags & VQ_FLAG_STATE_READY) != 0;
        int rc = -1;
        switch (D / TC_SLAB_FLOATS) {
            case 1: rc = launch_screen<1, 8, 2>(tz, thi, e_norm2, N, K, idx, hist, counter, fr, ready, st); break;
            case 2: rc = launch_screen<2, 8, 2>(tz, thi, e_norm2, N, K, idx, hist, counter, fr, ready, st); break;
            case 3: rc = launch_screen<3, 6, 2>(tz, thi, e_norm2, N, K, idx, hist, counter, fr, ready, st); break;
            case 4: rc = launch_screen<4, 4, 2>(tz, thi, e_norm2, N, K, idx, hist, counter, fr, ready, st); break;
            case 6: rc = launch_screen<6, 6, 1>(tz, thi, e_norm2, N, K, idx, hist, counter, fr, ready, st); break;
            case 8: rc = launch_screen<8, 4, 1>(tz, thi, e_norm2, N, K, idx, hist, counter, fr, ready, st); break;
            default: break;
        }
        if (rc >= 0) {
            if (rc != VQ_OK || !split) return rc;
            rows_done = true;           // indices are in place: fall through to the rows kernel
        }
    }
    if (rows_done) {
        // nothing: the screen kernel produced idx
    } else if (tensor_path_ok(N, K, D, flags, z, E_hi, E_lo, true)) {
        const long long row_tiles = (N + TC_ROWS - 1) / TC_ROWS;
        const bool pair = (K % TC2_CODES == 0) && !(flags & VQ_FLAG_TC_1CTA);
        // CTA pairs: a wave is 74 pairs, each covering 256 rows x 256 codes per tile
        const int code_tiles = pair ? K / TC2_CODES : K / TC_CODES;
        int splits = pair ? choose_splits((row_tiles + 1) / 2, code_tiles, 65535, kNumSMs / 2, 0.35)
                          : choose_splits(row_tiles, code_tiles, 65535, kNumSMs);
        const bool can_fuse = pair && quant && !(flags & VQ_FLAG_NO_FUSE) && aligned16(q_out) && aligned16(E) &&
                              (!want_onehot || aligned16(onehot));
        if (can_fuse && splits > 1) {
            // One fused launch (no key merge, no separate rows kernel) is worth about two code tiles of time:
            // split the codebook only when that still wins, i.e. when very few row tiles would leave SMs idle.
            const long long rtp = (row_tiles + 1) / 2;
            const int slots = kNumSMs / 2;
            const double fused_cost = static_cast<double>((rtp + slots - 1) / slots) * (code_tiles + 0.35);
            const double split_cost = static_cast<double>((rtp * splits + slots - 1) / slots) * (code_tiles / splits + 0.35);
            if (fused_cost <= split_cost + 2.0) splits = 1;
        }
        const int cps = K / splits;
        if (splits > 1) {
            keys = keys_buf;
            CUDA_TRY(cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * N, st));
        }
        CUtensorMap tz, thi, tlo;
        if (int rc = make_tmap(&tz, z, N, D)) return rc;
        if (int rc = make_tmap(&thi, E_hi, K, D)) return rc;
        if (int rc = make_tmap(&tlo, E_lo, K, D)) return rc;
        int rc = VQ_OK;
        const int nslab = D / TC_SLAB_FLOATS;
        if (pair) {
            // fused forward: the persistent kernel's writer warps also do vector_quantizer.py:39-56 (one launch)
            FusedRowArgs fr{};
            const bool fuse = can_fuse && splits == 1;
            if (fuse) {
                fr.z = z; fr.E = E; fr.q_out = q_out; fr.onehot = want_onehot ? onehot : nullptr; fr.hist = hist;
                fr.partials = partials; fr.counter = counter; fr.sse_out = sse; fr.loss = loss; fr.perplexity = perplexity;
                fr.beta = beta; fr.finalize = defer ? 0 : 1;
                fr.trace = g_trace_buf;
                static const int evict_first = [] { const char* e = getenv("B200VQ_ONEHOT_EVICT_FIRST"); return e != nullptr && e[0] == '0' ? 0 : 1; }();   // default on
                fr.onehot_evict_first = evict_first;
            }
            const FusedRowArgs* frp = fuse ? &fr : nullptr;
            switch (nslab) {
                case 1: rc = launch_tc2<1, 8, 2>(tz, thi, tlo, e_norm2, N, K, cps, splits, idx, keys, hist, counter, frp, (flags & VQ_FLAG_STATE_READY) != 0, st); break;
                case 2: rc = launch_tc2<2, 5, 2>(tz, thi, tlo, e_norm2, N, K, cps, splits, idx, keys, hist, counter, frp, (flags & VQ_FLAG_STATE_READY) != 0, st); break;
                case 3: rc = launch_tc2<3, 6, 1>(tz, thi, tlo, e_norm2, N, K, cps, splits, idx, keys, hist, counter, frp, (flags & VQ_FLAG_STATE_READY) != 0, st); break;
                case 4: rc = launch_tc2<4, 5, 1>(tz, thi, tlo, e_norm2, N, K, cps, splits, idx, keys, hist, counter, frp, (flags & VQ_FLAG_STATE_READY) != 0, st); break;
                default: return fail(VQ_ERR_ARG, "vq_forward: unsupported D=%d on the tensor path", D);
            }
            if (rc) return rc;
            if (fuse) return VQ_OK;     // rows work already done inside the kernel
        } else {
            switch (nslab) {
                case 1: rc = launch_tc<1, 8>(tz, thi, tlo, e_norm2, N, K, cps, splits, idx, keys, hist, counter, st); break;
                case 2: rc = launch_tc<2, 8>(tz, thi, tlo, e_norm2, N, K, cps, splits, idx, keys, hist, counter, st); break;
                case 3: rc = launch_tc<3, 6>(tz, thi, tlo, e_norm2, N, K, cps, splits, idx, keys, hist, counter, st); break;
                case 4: rc = launch_tc<4, 6>(tz, thi, tlo, e_norm2, N, K, cps, splits, idx, keys, hist, counter, st); break;
                default: return fail(VQ_ERR_ARG, "vq_forward: unsupported D=%d on the tensor path", D);
            }
        }
        if (rc) return rc;
    } else {
        const long long row_tiles = (N + S_TM - 1) / S_TM;
        if (row_tiles > 2147483647ll) return fail(VQ_ERR_ARG, "vq_forward: N=%lld too large", N);
        const int code_tiles = (K + S_TN - 1) / S_TN;
        int splits = 1;
        if (K % S_TN == 0) splits = choose_splits(row_tiles, code_tiles, 65535);
        const int cps = (K + splits - 1) / splits;
        if (splits > 1) {
            keys = keys_buf;
            CUDA_TRY(cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * N, st));
        }
        dim3 grid(static_cast<unsigned>(row_tiles), static_cast<unsigned>(splits));
        ProfScope prof(KID_ARGMIN_SIMT, st);
        argmin_simt_kernel<<<grid, 256, 0, st>>>(z, E, e_norm2, N, K, D, cps, idx, keys, hist, counter);
        LAUNCH_CHECK("argmin_simt_kernel");
    }

    // ---- 2. rows: gather / straight-through value / SSE / histogram / one-hot / loss / perplexity ----------
    const bool vec = (D % 4 == 0) && aligned16(z) && aligned16(E) && (!quant || aligned16(q_out));
    const int oh_vec_ok = want_onehot && (K % 4 == 0) && aligned16(onehot);
    const int fin = defer ? 0 : 1;
    const int R = rows_per_group(D, vec);
    int rgrid = rows_grid_for(N, R);
    // K <= 2048: the kernel keeps the usage counts per CTA in shared memory and flushes them once -- fewer, longer-lived
    // CTAs mean fewer reds per code (four resident CTAs per SM still keep > 100 KB of loads in flight)
    if (K <= 2048 && rgrid > kNumSMs * 4) rgrid = kNumSMs * 4;
    ProfScope prof_rows(KID_ROWS, st);
#define ROWS_ARGS z, E, idx, keys, N, K, D, beta, q_out, idx, onehot, hist, partials, counter, sse, loss, perplexity, fin, R, oh_vec_ok
#define ROWS_LAUNCH(OH, QU)                                                                      \
    do {                                                                                         \
        if (vec) quantize_rows_kernel<OH, QU, 4><<<rgrid, 256, 0, st>>>(ROWS_ARGS);              \
        else quantize_rows_kernel<OH, QU, 1><<<rgrid, 256, 0, st>>>(ROWS_ARGS);                  \
    } while (0)
    if (want_onehot && quant) ROWS_LAUNCH(true, true);
    else if (want_onehot) ROWS_LAUNCH(true, false);
    else if (quant) ROWS_LAUNCH(false, true);
    else ROWS_LAUNCH(false, false);
#undef ROWS_LAUNCH
#undef ROWS_ARGS
    LAUNCH_CHECK("quantize_rows_kernel");
    return VQ_OK;
}

// prepare + forward behind ONE entry (the call a training step makes: Adam has just changed E): the prepare launch
// (codebook norms, tf32 split, reset of hist / completion counter and -- when given -- of the dE accumulator the backward
// will add into) and the fused forward, chained by programmatic dependent launch.
int vq_step_forward(const float* z, const float* E, int64_t n_rows, int K, int D, float beta, int flags, float* e_norm2,
                    float* E_hi, float* E_lo, float* dE_zero, float* q_out, int32_t* idx, float* onehot, float* hist, float* sse,
                    float* loss, float* perplexity, void* workspace, size_t workspace_bytes, vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    if (z == nullptr || E == nullptr || e_norm2 == nullptr || hist == nullptr || K < 1 || D < 1 || n_rows < 0)
        return fail(VQ_ERR_ARG, "vq_step_forward: bad argument");
    const int fl = flags & ~VQ_FLAG_STATE_READY;
    const bool tensor = tensor_path_ok(n_rows, K, D, fl);
    const bool screen = n_rows > 0 && screen_path_ok(n_rows, K, D, fl, z, E, q_out, onehot) && ((fl & VQ_FLAG_SCREEN) || screen_enabled());
    if (tensor && (E_hi == nullptr || (!screen && E_lo == nullptr)))
        return fail(VQ_ERR_ARG, "vq_step_forward: E_hi (and, outside the screen + refine shapes, E_lo) scratch is NULL but this shape runs on the tensor path");
    if (int rc = vq_prepare_step(E, K, D, e_norm2, tensor ? E_hi : nullptr, (tensor && !screen) ? E_lo : nullptr, hist, workspace, workspace_bytes,
                                 dE_zero, stream))
        return rc;
    return vq_forward(z, E, e_norm2, E_hi, E_lo, n_rows, K, D, beta, fl | VQ_FLAG_STATE_READY, q_out, idx, onehot, hist, sse, loss, perplexity,
                      workspace, workspace_bytes, stream);
}

int vq_finalize_stats(const float* hist, const float* sse, int64_t n_rows_global, int K, int D, float beta,
                      float* loss, float* perplexity, vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    if (hist == nullptr || perplexity == nullptr || K < 1 || D < 1 || n_rows_global < 1)
        return fail(VQ_ERR_ARG, "vq_finalize_stats: bad argument");
    ProfScope prof(KID_FINALIZE, static_cast<cudaStream_t>(stream));
    finalize_stats_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(hist, sse, n_rows_global, K, D, beta, loss, perplexity);
    LAUNCH_CHECK("finalize_stats_kernel");
    return VQ_OK;
}

int vq_onehot(const int32_t* idx, int64_t n_rows, int K, float* onehot, vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    if (idx == nullptr || onehot == nullptr || K < 1 || n_rows < 0) return fail(VQ_ERR_ARG, "vq_onehot: bad argument");
    if (n_rows == 0) return VQ_OK;
    const int vec_ok = (K % 4 == 0) && aligned16(onehot);
    ProfScope prof(KID_ONEHOT, static_cast<cudaStream_t>(stream));
    onehot_kernel<<<rows_grid_for(n_rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(idx, n_rows, K, onehot, vec_ok);
    LAUNCH_CHECK("onehot_kernel");
    return VQ_OK;
}

}  // extern "C"

namespace {

// ---- backward strategy ------------------------------------------------------------
enum BwdPath { BWD_FLAT = 0, BWD_PRIVATE = 2, BWD_REPLICATED = 3 };

// Flat kernel with its reds spread over R zeroed copies of dE (folded into dE by a second small launch): R such that
// the copies together hold ~128 k 16-byte addresses, where the scatter stops being bound by same-address queueing in L2.
// Needs scratch (vq_step_backward: the tail of the forward's workspace, dead once the forward is complete) and enough
// rows per code for the queueing to matter.  1 = not applicable.
constexpr long long kReplAddrTarget = 131072;
int repl_factor(long long N, int K, int D, int flags, bool vec) {
    if (!vec || (flags & (VQ_FLAG_BWD_FLAT | VQ_FLAG_BWD_PRIVATE))) return 1;
    const long long addr = static_cast<long long>(K) * (D / 4);
    if (addr >= kReplAddrTarget || N < 262144 || N < 256ll * K) return 1;
    long long r = 2 * kReplAddrTarget / addr;            // >= 2; e.g. K = 1024, D = 256: 64 k addresses, 1024 reds each -> 4 copies
    return static_cast<int>(r > 32 ? 32 : r);
}

// columns per lane of the private kernel (0: does not apply) -- the CTA's share of dE must fit shared memory
int private_nc(int K, int D) {
    if (D % 64 == 0 && pv_smem_bytes(K, 2) <= 227 * 1024) return 2;
    if (D % 32 == 0 && pv_smem_bytes(K, 1) <= 227 * 1024) return 1;
    return 0;
}

int choose_bwd_path(long long N, int K, int D, int flags, bool vec) {
    if (flags & VQ_FLAG_BWD_FLAT) return BWD_FLAT;
    const bool private_ok = vec && private_nc(K, D) != 0 && N < (1ll << 40);
    if (flags & VQ_FLAG_BWD_PRIVATE) return private_ok ? BWD_PRIVATE : BWD_FLAT;
    // the private tables are flushed with (row chunks) * K * D atomics, and the kernel only beats the flat one where
    // the flat one is bound by its atomics (few addresses): measured 196 vs 252 us at K = 512, D = 64, N = 1M
    if (private_ok && N >= 256ll * K && static_cast<long long>(K) * D <= 32768) return BWD_PRIVATE;
    return BWD_FLAT;
}

template <bool HAS_GQ, int NC>
int run_private(const float* g_q, const float* g_loss, const float* z, const float* E, const int* idx, long long N, float denom_dz,
                float denom_dE, int K, int D, float beta, float* dz, float* dE, cudaStream_t st) {
    const int smem = pv_smem_bytes(K, NC);
    static DeviceOnce once;                               // the table size depends on K: opt in to the maximum once
    CUDA_TRY(once.max_smem(backward_private_kernel<HAS_GQ, NC>, 227 * 1024));
    const int slices = D / (32 * NC);
    int chunks = sm_count() / slices;                     // one CTA per SM (the table fills its shared memory)
    if (chunks < 1) chunks = 1;
    const long long max_chunks = (N + PV_WIN - 1) / PV_WIN;
    if (chunks > max_chunks) chunks = static_cast<int>(max_chunks);
    const long long rows_per_cta = (N + chunks - 1) / chunks;
    ProfScope prof(KID_BACKWARD, st);
    const cudaError_t e = launch_pdl(backward_private_kernel<HAS_GQ, NC>, dim3(chunks, slices), dim3(PV_THREADS), smem, st, g_q, g_loss, z,
                                     E, idx, N, rows_per_cta, denom_dz, denom_dE, K, D, beta, dz, dE);
    if (e != cudaSuccess) return fail(VQ_ERR_CUDA, "launch of backward_private_kernel failed: %s", cudaGetErrorString(e));
    LAUNCH_CHECK("backward_private_kernel");
    return VQ_OK;
}

}  // namespace

extern "C" {

int vq_backward_path(int64_t n_rows, int K, int D, int flags) {
    if (repl_factor(n_rows, K, D, flags, D % 4 == 0) > 1) return BWD_REPLICATED;     // vq_step_backward (has the workspace)
    return choose_bwd_path(n_rows, K, D, flags, D % 4 == 0);
}

static int backward_impl(const float* g_q, const float* g_loss, const float* z, const float* E, const int32_t* idx,
                         int64_t n_rows, int64_t n_rows_dz, int64_t n_rows_dE, int K, int D, float beta, int flags, float* dz,
                         float* dE, const unsigned int* ready, float* repl, size_t repl_bytes, vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    const long long N = n_rows;
    const bool train = (flags & VQ_FLAG_TRAIN_VQ) != 0 && dE != nullptr;
    if (z == nullptr || E == nullptr || idx == nullptr || (dz == nullptr && !train) || K < 1 || D < 1 || N < 0 || n_rows_dz < 1 ||
        n_rows_dE < 1)
        return fail(VQ_ERR_ARG, "vq_backward: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool zero_dE = train && (flags & VQ_FLAG_ZERO_DE) != 0;
    if (N == 0) {
        if (zero_dE) CUDA_TRY(cudaMemsetAsync(dE, 0, sizeof(float) * static_cast<size_t>(K) * D, st));
        return VQ_OK;
    }
    const float denom_dz = static_cast<float>(static_cast<double>(n_rows_dz) * static_cast<double>(D));
    const float denom_dE = static_cast<float>(static_cast<double>(n_rows_dE) * static_cast<double>(D));
    const bool vec = (D % 4 == 0) && aligned16(z) && aligned16(E) && (dz == nullptr || aligned16(dz)) && (g_q == nullptr || aligned16(g_q)) &&
                     (!train || aligned16(dE));
    int n_repl = (train && dz != nullptr && repl != nullptr && aligned16(repl)) ? repl_factor(N, K, D, flags, vec) : 1;
    if (n_repl > 1 && static_cast<size_t>(n_repl) * K * D * sizeof(float) > repl_bytes) n_repl = 1;
    const int path = !train ? BWD_FLAT : (n_repl > 1 ? BWD_REPLICATED : choose_bwd_path(N, K, D, flags, vec));
    if (zero_dE) CUDA_TRY(cudaMemsetAsync(dE, 0, sizeof(float) * static_cast<size_t>(K) * D, st));
    if (n_repl > 1) CUDA_TRY(cudaMemsetAsync(repl, 0, sizeof(float) * static_cast<size_t>(n_repl) * K * D, st));
    if (path == BWD_PRIVATE) {
        const int nc = private_nc(K, D);
        if (g_q != nullptr)
            return nc == 2 ? run_private<true, 2>(g_q, g_loss, z, E, idx, N, denom_dz, denom_dE, K, D, beta, dz, dE, st)
                           : run_private<true, 1>(g_q, g_loss, z, E, idx, N, denom_dz, denom_dE, K, D, beta, dz, dE, st);
        return nc == 2 ? run_private<false, 2>(g_q, g_loss, z, E, idx, N, denom_dz, denom_dE, K, D, beta, dz, dE, st)
                       : run_private<false, 1>(g_q, g_loss, z, E, idx, N, denom_dz, denom_dE, K, D, beta, dz, dE, st);
    }
    if (dz == nullptr) {   // codebook gradient only
        const long long n_e = N * (vec ? D / 4 : D);
        long long gb = (n_e + 255) / 256;
        if (gb > kNumSMs * 32) gb = kNumSMs * 32;
        ProfScope prof(KID_BACKWARD, st);
        const cudaError_t e = vec ? launch_pdl(backward_dE_kernel<4>, dim3(static_cast<unsigned>(gb)), dim3(256), 0, st, g_loss, z, E, idx, N, denom_dE, D, dE, static_cast<const unsigned int*>(nullptr))
                                  : launch_pdl(backward_dE_kernel<1>, dim3(static_cast<unsigned>(gb)), dim3(256), 0, st, g_loss, z, E, idx, N, denom_dE, D, dE, static_cast<const unsigned int*>(nullptr));
        if (e != cudaSuccess) return fail(VQ_ERR_CUDA, "launch of backward_dE_kernel failed: %s", cudaGetErrorString(e));
        LAUNCH_CHECK("backward_dE_kernel");
        return VQ_OK;
    }
    const long long n_el = N * (vec ? D / 4 : D);
    long long g = (n_el + 255) / 256;
    if (g > kNumSMs * 32) g = kNumSMs * 32;
    const int grid = static_cast<int>(g);
    ProfScope prof(KID_BACKWARD, st);
    const long long repl_stride = static_cast<long long>(K) * D;
#define BWD_ARGS g_q, g_loss, z, E, idx, N, denom_dz, denom_dE, D, beta, dz, dE, ready, repl, n_repl, repl_stride
#define BWD_LAUNCH(TR, GQ)                                                                                        \
    do {                                                                                                          \
        cudaError_t e__;                                                                                          \
        if (vec) e__ = launch_pdl(backward_kernel<TR, GQ, 4>, dim3(grid), dim3(256), 0, st, BWD_ARGS);            \
        else e__ = launch_pdl(backward_kernel<TR, GQ, 1>, dim3(grid), dim3(256), 0, st, BWD_ARGS);                \
        if (e__ != cudaSuccess) return fail(VQ_ERR_CUDA, "launch of backward_kernel failed: %s", cudaGetErrorString(e__)); \
    } while (0)
    if (train && g_q) BWD_LAUNCH(true, true);
    else if (train) BWD_LAUNCH(true, false);
    else if (g_q) BWD_LAUNCH(false, true);
    else BWD_LAUNCH(false, false);
#undef BWD_LAUNCH
#undef BWD_ARGS
    LAUNCH_CHECK("backward_kernel");
    if (n_repl > 1) {
        const long long n4 = repl_stride / 4;
        long long gr = (n4 + 255) / 256;
        if (gr > kNumSMs * 8) gr = kNumSMs * 8;
        reduce_replicas_kernel<<<static_cast<unsigned>(gr), 256, 0, st>>>(repl, n_repl, n4, dE);
        LAUNCH_CHECK("reduce_replicas_kernel");
    }
    return VQ_OK;
}

int vq_backward(const float* g_q, const float* g_loss, const float* z, const float* E, const int32_t* idx,
                int64_t n_rows, int64_t n_rows_dz, int64_t n_rows_dE, int K, int D, float beta, int flags, float* dz,
                float* dE, vq_stream_t stream) {
    return backward_impl(g_q, g_loss, z, E, idx, n_rows, n_rows_dz, n_rows_dE, K, D, beta, flags, dz, dE, nullptr, nullptr, 0, stream);
}

// The backward launched RIGHT BEHIND vq_step_forward on the same stream and workspace (nothing in between): on the
// screen + refine path the flat kernel starts as soon as the forward's last CTA has raised the workspace's ready word --
// indices and everything else it reads are complete then -- and overlaps the forward's serial statistics tail.
int vq_step_backward(const float* g_q, const float* g_loss, const float* z, const float* E, const int32_t* idx,
                     int64_t n_rows, int64_t n_rows_dz, int64_t n_rows_dE, int K, int D, float beta, int flags, float* dz,
                     float* dE, void* workspace, size_t workspace_bytes, int forward_flags, vq_stream_t stream) {
    const unsigned int* ready = nullptr;
    const int ff = forward_flags & ~VQ_FLAG_STATE_READY;
    const bool train = (flags & VQ_FLAG_TRAIN_VQ) != 0 && dE != nullptr;
    const bool vec = D % 4 == 0;
    // the tail of the forward's workspace (per-row keys / spill lists) is dead once the forward is complete: scratch for
    // the replicated scatter (the kernel orders itself behind the forward before it touches it)
    float* repl = nullptr;
    size_t repl_bytes = 0;
    if (workspace != nullptr && n_rows > 0 && train && repl_factor(n_rows, K, D, flags, vec) > 1) {
        const WsLayout w = ws_layout(n_rows);
        if (workspace_bytes >= w.total) {
            repl = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + w.keys_off);
            repl_bytes = w.total - w.keys_off;
        }
    }
    if (repl == nullptr && workspace != nullptr && n_rows > 0 && dz != nullptr && !(flags & VQ_FLAG_ZERO_DE) &&
        screen_path_ok(n_rows, K, D, ff, z, E, nullptr, nullptr) && ((ff & VQ_FLAG_SCREEN) || screen_enabled()) &&
        !screen_split_rows(n_rows, K, D, ff) &&       // (the rows kernel behind the screen kernel never raises the ready word)
        (!train || choose_bwd_path(n_rows, K, D, flags, vec) == BWD_FLAT)) {
        const WsLayout w = ws_layout(n_rows);
        if (workspace_bytes < w.total) return fail(VQ_ERR_WORKSPACE, "vq_step_backward: workspace %zu B < %zu B", workspace_bytes, w.total);
        ready = reinterpret_cast<const unsigned int*>(static_cast<const uint8_t*>(workspace) + w.counter_off) + 2;
    }
    return backward_impl(g_q, g_loss, z, E, idx, n_rows, n_rows_dz, n_rows_dE, K, D, beta, flags, dz, dE, ready, repl, repl_bytes, stream);
}

// =========================================================================================================
// data parallel: sum all-reduce of the packed step buffer [dE | usage histogram | squared error] over NVLink peer memory
// =========================================================================================================
constexpr int DP_EVENTS = 32;
struct vq_dp_ctx {
    DpCtxDev dev;
    int device;
    // overlapped form (vq_dp_allreduce_start / vq_dp_wait): a stream of the context's own and a small ring of events
    cudaStream_t side;
    cudaEvent_t fork_ev[DP_EVENTS], done_ev[DP_EVENTS];
    unsigned long long started, joined;      // exchanges started on `side` / exchanges the caller's stream already waits for
};

int vq_dp_create(const void* const* recv0, const void* const* recv1, void* multicast0, void* multicast1, int world, int rank,
                 int64_t n_floats, uint32_t spin_limit, vq_dp_ctx** out) {
    if (int rc = check_device()) return rc;
    if (out == nullptr || recv0 == nullptr || recv1 == nullptr || world < 1 || world > DP_MAX_RANKS || rank < 0 || rank >= world || n_floats < 1)
        return fail(VQ_ERR_ARG, "vq_dp_create: bad argument (world=%d rank=%d n=%lld)", world, rank, (long long)n_floats);
    vq_dp_ctx* c = new (std::nothrow) vq_dp_ctx();
    if (c == nullptr) return fail(VQ_ERR_ARG, "vq_dp_create: out of host memory");
    memset(c, 0, sizeof(*c));
    for (int p = 0; p < world; ++p) {
        if (recv0[p] == nullptr || recv1[p] == nullptr || !aligned16(recv0[p]) || !aligned16(recv1[p])) {
            delete c;
            return fail(VQ_ERR_ARG, "vq_dp_create: receive buffer %d is NULL or misaligned", p);
        }
        c->dev.recv[0][p] = static_cast<float*>(const_cast<void*>(recv0[p]));
        c->dev.recv[1][p] = static_cast<float*>(const_cast<void*>(recv1[p]));
    }
    const bool mc = multicast0 != nullptr && multicast1 != nullptr;
    c->dev.mc[0] = mc ? static_cast<float*>(multicast0) : nullptr;
    c->dev.mc[1] = mc ? static_cast<float*>(multicast1) : nullptr;
    c->dev.world = world; c->dev.rank = rank;
    c->dev.n = n_floats;
    c->dev.L = (n_floats + 1) / 2;
    c->dev.S = (c->dev.L + world - 1) / world;
    // one step (every payload to every rank) below 8 ranks; reduce-scatter + all-gather from 8 ranks on, where the
    // world x payload that one step lands in every rank costs more than a second NVLink latency (11.9 vs 14.8 us at 8 GPUs).
    // B200VQ_AR_ALGO=1|2 forces either; both fit receive buffers of world x (L + 2) lines.
    static const int forced = [] { const char* e = getenv("B200VQ_AR_ALGO"); return e == nullptr ? 0 : atoi(e); }();
    c->dev.two_step = (world >= 2 && 2 * c->dev.S <= c->dev.L + 2 && (forced == 2 || (forced != 1 && world >= 8))) ? 1 : 0;
    c->dev.spin_limit = spin_limit != 0 ? spin_limit : (1u << 23);     // a few seconds of polling
    cudaGetDevice(&c->device);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->dev.state), 64);
    if (e == cudaSuccess) e = cudaMemset(c->dev.state, 0, 64);
    if (e != cudaSuccess) {
        delete c;
        return fail(VQ_ERR_CUDA, "vq_dp_create: %s", cudaGetErrorString(e));
    }
    // The overlapped exchange's own stream.  Lowest priority by default: the fused forward needs whole SMs (all registers
    // and shared memory), so an exchange CTA that slips onto an SM first holds the forward's CTA there up for as long as it
    // polls; at low priority the exchange yields to whatever the step's stream has pending and runs next to small CTAs
    // (the backward's, or in training the encoder's).  B200VQ_DP_SIDE_PRIORITY=high reverses it.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    static const bool side_high = [] { const char* v = getenv("B200VQ_DP_SIDE_PRIORITY"); return v != nullptr && v[0] == 'h'; }();
    e = cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, side_high ? prio_hi : prio_lo);
    for (int i = 0; i < DP_EVENTS && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&c->fork_ev[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done_ev[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        vq_dp_destroy(c);
        return fail(VQ_ERR_CUDA, "vq_dp_create: %s", cudaGetErrorString(e));
    }
    *out = c;
    return VQ_OK;
}

void vq_dp_destroy(vq_dp_ctx* c) {
    if (c == nullptr) return;
    if (c->side != nullptr) {
        cudaStreamSynchronize(c->side);
        cudaStreamDestroy(c->side);
    }
    for (int i = 0; i < DP_EVENTS; ++i) {
        if (c->fork_ev[i] != nullptr) cudaEventDestroy(c->fork_ev[i]);
        if (c->done_ev[i] != nullptr) cudaEventDestroy(c->done_ev[i]);
    }
    cudaFree(c->dev.state);
    delete c;
}

int64_t vq_dp_recv_lines(int world, int64_t n_floats) { return static_cast<int64_t>(world) * ((n_floats + 1) / 2 + 2); }

// synchronises `stream`; calls completed so far and the error word (bit 0: a wait ran into spin_limit)
int vq_dp_status(vq_dp_ctx* c, uint32_t* calls_done, uint32_t* error_word, vq_stream_t stream) {
    if (c == nullptr) return fail(VQ_ERR_ARG, "vq_dp_status: null");
    unsigned int h[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(h, c->dev.state, sizeof(h), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    if (calls_done) *calls_done = h[0];
    if (error_word) *error_word = h[1];
    return VQ_OK;
}

// out[i] = sum over ranks (rank order: bit-identical everywhere) of payload[i]
int vq_dp_allreduce(vq_dp_ctx* c, const float* payload, float* out, vq_stream_t stream) {
    if (c == nullptr || payload == nullptr || out == nullptr) return fail(VQ_ERR_ARG, "vq_dp_allreduce: null");
    if (!aligned16(out)) return fail(VQ_ERR_ARG, "vq_dp_allreduce: out is misaligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long blocks = (c->dev.L + DP_THREADS - 1) / DP_THREADS;
    const int cap = sm_count();
    if (blocks > cap) blocks = cap;                          // no CTA waits for another CTA of this grid: no residency requirement
    ProfScope prof(KID_ALLREDUCE, st);
    const cudaError_t e = launch_pdl(dp_allreduce_kernel, dim3(static_cast<unsigned>(blocks)), dim3(DP_THREADS), 0, st, c->dev, payload, out);
    if (e != cudaSuccess) return fail(VQ_ERR_CUDA, "launch of dp_allreduce_kernel failed: %s", cudaGetErrorString(e));
    LAUNCH_CHECK("dp_allreduce_kernel");
    return VQ_OK;
}

// Data-parallel backward with the exchange in the backward kernel's TAIL (backward_dp_kernel, kernels_dp.cuh): one launch
// produces dz, the codebook gradient and the summed packed buffer.  Shapes the flat 16-byte path does not cover take the
// fused backward followed by the exchange kernel.
int vq_step_backward_dp(const float* g_q, const float* g_loss, const float* z, const float* E, const int32_t* idx, int64_t n_rows,
                        int64_t n_rows_dz, int64_t n_rows_dE, int K, int D, float beta, int flags, float* dz, float* packed,
                        vq_dp_ctx* ctx, float* out, void* workspace, size_t workspace_bytes, int forward_flags, vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    if (ctx == nullptr || packed == nullptr || out == nullptr || z == nullptr || E == nullptr || idx == nullptr || K < 1 || D < 1 || n_rows < 0 ||
        n_rows_dz < 1 || n_rows_dE < 1)
        return fail(VQ_ERR_ARG, "vq_step_backward_dp: bad argument");
    if (ctx->dev.n < static_cast<long long>(K) * D) return fail(VQ_ERR_ARG, "vq_step_backward_dp: the context's payload is smaller than K * D");
    const long long N = n_rows;
    const bool vec = (D % 4 == 0) && aligned16(z) && aligned16(E) && aligned16(packed) && aligned16(out) && (dz == nullptr || aligned16(dz)) &&
                     (g_q == nullptr || aligned16(g_q));
    static const bool tail_off = [] { const char* e = getenv("B200VQ_DP_TAIL"); return e != nullptr && e[0] == '0'; }();
    if (tail_off || !(flags & VQ_FLAG_TRAIN_VQ) || dz == nullptr || !vec || N == 0 || (flags & VQ_FLAG_ZERO_DE) ||
        repl_factor(N, K, D, flags, vec) > 1) {
        if (int rc = vq_step_backward(g_q, g_loss, z, E, idx, n_rows, n_rows_dz, n_rows_dE, K, D, beta, flags, dz, packed, workspace,
                                      workspace_bytes, forward_flags, stream))
            return rc;
        return vq_dp_allreduce(ctx, packed, out, stream);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ff = forward_flags & ~VQ_FLAG_STATE_READY;
    const unsigned int* ready = nullptr;
    if (workspace != nullptr && screen_path_ok(n_rows, K, D, ff, z, E, nullptr, nullptr) && ((ff & VQ_FLAG_SCREEN) || screen_enabled()) &&
        !screen_split_rows(n_rows, K, D, ff)) {
        const WsLayout w = ws_layout(n_rows);
        if (workspace_bytes < w.total) return fail(VQ_ERR_WORKSPACE, "vq_step_backward_dp: workspace %zu B < %zu B", workspace_bytes, w.total);
        ready = reinterpret_cast<const unsigned int*>(static_cast<const uint8_t*>(workspace) + w.counter_off) + 2;
    }
    const float denom_dz = static_cast<float>(static_cast<double>(n_rows_dz) * static_cast<double>(D));
    const float denom_dE = static_cast<float>(static_cast<double>(n_rows_dE) * static_cast<double>(D));
    const long long n_el = N * (D / 4);
    long long g = (n_el + 1023) / 1024;                    // four elements per thread and pass
    static const int ctas_per_sm = [] { const char* e = getenv("B200VQ_DP_TAIL_CTAS_PER_SM"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
    if (g > kNumSMs * ctas_per_sm) g = kNumSMs * ctas_per_sm;
    ProfScope prof(KID_BACKWARD, st);
    const cudaError_t e =
        g_q != nullptr ? launch_pdl(backward_dp_kernel<true>, dim3(static_cast<unsigned>(g)), dim3(256), 0, st, g_q, g_loss, z, E, idx, N, denom_dz, denom_dE, D,
                                    beta, dz, packed, ready, ctx->dev, out)
                       : launch_pdl(backward_dp_kernel<false>, dim3(static_cast<unsigned>(g)), dim3(256), 0, st, g_q, g_loss, z, E, idx, N, denom_dz, denom_dE, D,
                                    beta, dz, packed, ready, ctx->dev, out);
    if (e != cudaSuccess) return fail(VQ_ERR_CUDA, "launch of backward_dp_kernel failed: %s", cudaGetErrorString(e));
    LAUNCH_CHECK("backward_dp_kernel");
    return VQ_OK;
}

// Overlapped form.  The exchange is ordered behind everything enqueued on `stream` so far but runs on the context's own
// stream, so `stream` carries on at once: the next step's codebook preparation and forward overlap the NVLink transfer
// and absorb the ranks' skew (in training: the encoder's backward does, as under DDP).  Plain launch (full dependencies:
// the kernel's griddepcontrol instructions are no-ops then) -- the fork and join are event edges, which a stream capture
// turns into graph edges.
int vq_dp_allreduce_start(vq_dp_ctx* c, const float* payload, float* out, vq_stream_t stream) {
    if (c == nullptr || payload == nullptr || out == nullptr) return fail(VQ_ERR_ARG, "vq_dp_allreduce_start: null");
    if (!aligned16(out)) return fail(VQ_ERR_ARG, "vq_dp_allreduce_start: out is misaligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int slot = static_cast<int>(c->started % DP_EVENTS);
    CUDA_TRY(cudaEventRecord(c->fork_ev[slot], st));
    CUDA_TRY(cudaStreamWaitEvent(c->side, c->fork_ev[slot], 0));
    long long blocks = (c->dev.L + DP_THREADS - 1) / DP_THREADS;
    const int cap = sm_count();
    if (blocks > cap) blocks = cap;
    {
        ProfScope prof(KID_ALLREDUCE, c->side);
        dp_allreduce_kernel<<<static_cast<unsigned>(blocks), DP_THREADS, 0, c->side>>>(c->dev, payload, out);
        LAUNCH_CHECK("dp_allreduce_kernel");
    }
    CUDA_TRY(cudaEventRecord(c->done_ev[slot], c->side));
    c->started++;
    return VQ_OK;
}

// `stream` waits for every exchange started so far except the `keep_in_flight` most recent ones (exchanges complete in
// the order they were started).  keep_in_flight = 0 before the results are read, before the context is destroyed and
// before a stream capture ends (a forked stream must rejoin); 1 inside a double-buffered step loop.
int vq_dp_wait(vq_dp_ctx* c, int keep_in_flight, vq_stream_t stream) {
    if (c == nullptr || keep_in_flight < 0 || keep_in_flight >= DP_EVENTS) return fail(VQ_ERR_ARG, "vq_dp_wait: bad argument (keep_in_flight < %d)", DP_EVENTS);
    const unsigned long long keep = static_cast<unsigned long long>(keep_in_flight);
    if (c->started <= keep || c->started - keep <= c->joined) return VQ_OK;
    const unsigned long long upto = c->started - keep;                  // exchanges [joined, upto) must be complete
    CUDA_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->done_ev[(upto - 1) % DP_EVENTS], 0));
    c->joined = upto;
    return VQ_OK;
}

// Single-GPU emulation of `world` ranks for the tests: ONE cooperative launch in which blockIdx.y plays the rank, all
// ranks' receive buffers living on this GPU (plain stores instead of NVLink; no multicast).  payloads / outs: host arrays
// of `world` device pointers.  Runs `rounds` calls back to back (buffer parity, sequence numbers).
int vq_dp_emulate(int world, int two_step, int64_t n_floats, const float* const* payloads, float* const* outs, int rounds,
                  uint32_t spin_limit, uint32_t* error_word, vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    if (world < 1 || world > DP_MAX_RANKS || n_floats < 1 || payloads == nullptr || outs == nullptr || rounds < 1)
        return fail(VQ_ERR_ARG, "vq_dp_emulate: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long L = (n_floats + 1) / 2, S = (L + world - 1) / world;
    if (two_step && 2 * S > L + 2) return fail(VQ_ERR_ARG, "vq_dp_emulate: payload too small for two steps");
    const size_t buf_bytes = static_cast<size_t>(world) * (L + 2) * 16;
    uint8_t* pool = nullptr;
    const size_t pool_bytes = 2 * world * buf_bytes + world * 64 + world * sizeof(DpCtxDev) + 2 * world * sizeof(void*);
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&pool), pool_bytes));
    int rc = VQ_OK;
    do {
        if (cudaMemsetAsync(pool, 0, pool_bytes, st) != cudaSuccess) { rc = fail(VQ_ERR_CUDA, "vq_dp_emulate: memset failed"); break; }
        uint8_t* states = pool + 2 * world * buf_bytes;
        DpCtxDev* d_ctx = reinterpret_cast<DpCtxDev*>(states + world * 64);
        const float** d_pay = reinterpret_cast<const float**>(d_ctx + world);
        float** d_out = const_cast<float**>(d_pay + world);
        std::vector<DpCtxDev> h(world);
        for (int r = 0; r < world; ++r) {
            memset(&h[r], 0, sizeof(DpCtxDev));
            for (int par = 0; par < 2; ++par)
                for (int p = 0; p < world; ++p) h[r].recv[par][p] = reinterpret_cast<float*>(pool + (static_cast<size_t>(par) * world + p) * buf_bytes);
            h[r].state = reinterpret_cast<unsigned int*>(states + r * 64);
            h[r].L = L; h[r].S = S; h[r].n = n_floats; h[r].world = world; h[r].rank = r; h[r].two_step = two_step ? 1 : 0;
            h[r].spin_limit = spin_limit != 0 ? spin_limit : (1u << 22);
        }
        if (cudaMemcpyAsync(d_ctx, h.data(), world * sizeof(DpCtxDev), cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(d_pay, payloads, world * sizeof(void*), cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(d_out, outs, world * sizeof(void*), cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { rc = fail(VQ_ERR_CUDA, "vq_dp_emulate: setup copies failed"); break; }
        long long blocks = (L + DP_THREADS - 1) / DP_THREADS;
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_emulate_kernel, DP_THREADS, 0);
        const long long cap = static_cast<long long>(per_sm < 1 ? 1 : per_sm) * sm_count() / world;   // all ranks resident at once
        if (blocks > cap) blocks = cap;
        if (blocks < 1) { rc = fail(VQ_ERR_ARG, "vq_dp_emulate: world too large for a cooperative launch"); break; }
        const DpCtxDev* a0 = d_ctx; const float* const* a1 = d_pay; float* const* a2 = d_out;
        void* args[] = {&a0, &a1, &a2};
        for (int i = 0; i < rounds && rc == VQ_OK; ++i) {
            const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(dp_emulate_kernel),
                                                              dim3(static_cast<unsigned>(blocks), static_cast<unsigned>(world)), dim3(DP_THREADS), args, 0, st);
            if (e != cudaSuccess) rc = fail(VQ_ERR_CUDA, "vq_dp_emulate: cooperative launch failed: %s", cudaGetErrorString(e));
            else g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        if (rc != VQ_OK) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = fail(VQ_ERR_CUDA, "vq_dp_emulate: kernel failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        unsigned int err = 0;
        for (int r = 0; r < world; ++r) {
            unsigned int hs[2];
            if (cudaMemcpy(hs, states + r * 64, sizeof(hs), cudaMemcpyDeviceToHost) != cudaSuccess) { rc = fail(VQ_ERR_CUDA, "vq_dp_emulate: readback failed"); break; }
            err |= hs[1];
            if (hs[0] != static_cast<unsigned int>(rounds)) err |= 0x100u;      // the call counter did not advance once per round
        }
        if (error_word) *error_word = err;
    } while (false);
    cudaFree(pool);
    return rc;
}

// fc_1(one_hot) as an index gather (location_model.py:10,21; train_location.py:74-75).
int vq_gather_sum_rows(const int32_t* idx, const float* Wt, const float* bias, float* y, int B, int T, int K, int O,
                       vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    if (idx == nullptr || Wt == nullptr || y == nullptr || B < 0 || T < 1 || K < 1 || O < 4 || O % 4 != 0 || !aligned16(Wt) ||
        !aligned16(y) || (bias != nullptr && !aligned16(bias)))
        return fail(VQ_ERR_ARG, "vq_gather_sum_rows: bad argument (B=%d T=%d K=%d O=%d; O must be a multiple of 4, pointers 16-byte aligned)", B, T, K, O);
    if (B == 0) return VQ_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(static_cast<unsigned>(B), static_cast<unsigned>((O / 4 + 255) / 256));
    gather_sum_rows_kernel<<<grid, 256, sizeof(int) * T, st>>>(idx, Wt, bias, y, T, K, O);
    LAUNCH_CHECK("gather_sum_rows_kernel");
    return VQ_OK;
}

int vq_scatter_add_rows(const int32_t* idx, const float* g, float* dWt, int B, int T, int K, int O, vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    if (idx == nullptr || g == nullptr || dWt == nullptr || B < 0 || T < 1 || K < 1 || O < 4 || O % 4 != 0 || !aligned16(g) ||
        !aligned16(dWt) || T > 65535)
        return fail(VQ_ERR_ARG, "vq_scatter_add_rows: bad argument (B=%d T=%d K=%d O=%d)", B, T, K, O);
    if (B == 0) return VQ_OK;
    dim3 grid(static_cast<unsigned>(B), static_cast<unsigned>(T));
    scatter_add_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(idx, g, dWt, T, K, O);
    LAUNCH_CHECK("scatter_add_rows_kernel");
    return VQ_OK;
}

// SURVEY 8(f) rank 2, first half: the time-mean variant in front of the quantizer (convolutional_vq_vae.py:96-97).
int vq_time_mean(const float* x, int64_t rows, int T, float* z, vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    if (x == nullptr || z == nullptr || rows < 0 || T < 1) return fail(VQ_ERR_ARG, "vq_time_mean: bad argument");
    if (rows == 0) return VQ_OK;
    long long blocks = (rows + 7) / 8;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const cudaError_t e = launch_pdl(time_mean_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, x, static_cast<long long>(rows), T, z);
    if (e != cudaSuccess) return fail(VQ_ERR_CUDA, "launch of time_mean_kernel failed: %s", cudaGetErrorString(e));
    LAUNCH_CHECK("time_mean_kernel");
    return VQ_OK;
}

int vq_time_mean_backward(const float* dz, int64_t rows, int T, float* dx, vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    if (dz == nullptr || dx == nullptr || rows < 0 || T < 1) return fail(VQ_ERR_ARG, "vq_time_mean_backward: bad argument");
    if (rows == 0) return VQ_OK;
    long long blocks = (rows * T + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    time_mean_backward_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(dz, static_cast<long long>(rows), T, dx);
    LAUNCH_CHECK("time_mean_backward_kernel");
    return VQ_OK;
}

// Jitter (modules/jitter.py:47-70) -- in-place gather along time of a (rows, T) tensor.
int vq_jitter_apply(float* q, const int32_t* src, int64_t rows, int T, vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    if (q == nullptr || src == nullptr || rows < 0 || T < 1 || T > 12288 || rows > 2147483647ll)
        return fail(VQ_ERR_ARG, "vq_jitter_apply: bad argument (rows=%lld T=%d; T <= 12288)", (long long)rows, T);
    if (rows == 0) return VQ_OK;
    jitter_gather_kernel<<<static_cast<unsigned>(rows), 256, sizeof(float) * T, static_cast<cudaStream_t>(stream)>>>(q, src, T);
    LAUNCH_CHECK("jitter_gather_kernel");
    return VQ_OK;
}

int vq_jitter_backward(float* g, const int32_t* src, int64_t rows, int T, vq_stream_t stream) {
    if (int rc = check_device()) return rc;
    if (g == nullptr || src == nullptr || rows < 0 || T < 1) return fail(VQ_ERR_ARG, "vq_jitter_backward: bad argument");
    if (rows == 0) return VQ_OK;
    long long blocks = (rows * T + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    jitter_backward_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, src, rows, T);
    LAUNCH_CHECK("jitter_backward_kernel");
    return VQ_OK;
}

// =========================================================================================================
// host-buffer context: two lanes (stream + staging) so the H2D copy of step i+1 overlaps the kernels of step i
// =========================================================================================================
struct vq_host_ctx {
    long long max_rows;
    int K, D;
    float *E, *e_norm2, *E_hi, *E_lo;
    struct Lane {
        cudaStream_t st;
        float *z, *gq, *q, *dz, *dE, *hist, *scal;   // scal: [sse, loss, perplexity, g_loss]
        int32_t* idx;
        void* ws;
        size_t ws_bytes;
        bool gq_is_ones;
        cudaEvent_t ev_stop;
    } lane[2];
    cudaEvent_t ev_start;
};

static void host_ctx_free(vq_host_ctx* c) {
    if (c == nullptr) return;
    cudaFree(c->E); cudaFree(c->e_norm2); cudaFree(c->E_hi); cudaFree(c->E_lo);
    for (auto& l : c->lane) {
        cudaFree(l.z); cudaFree(l.gq); cudaFree(l.q); cudaFree(l.dz); cudaFree(l.dE); cudaFree(l.hist);
        cudaFree(l.scal); cudaFree(l.idx); cudaFree(l.ws);
        if (l.ev_stop) cudaEventDestroy(l.ev_stop);
        if (l.st) cudaStreamDestroy(l.st);
    }
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    delete c;
}

int vq_host_ctx_create(int64_t max_rows, int K, int D, vq_host_ctx** out) {
    if (int rc = check_device()) return rc;
    if (out == nullptr || max_rows < 1 || K < 1 || D < 1) return fail(VQ_ERR_ARG, "vq_host_ctx_create: bad argument");
    vq_host_ctx* c = new (std::nothrow) vq_host_ctx();
    if (c == nullptr) return fail(VQ_ERR_ARG, "vq_host_ctx_create: out of host memory");
    memset(c, 0, sizeof(*c));
    c->max_rows = max_rows; c->K = K; c->D = D;
    const size_t kd = sizeof(float) * K * D, nd = sizeof(float) * max_rows * D;
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A(reinterpret_cast<void**>(&c->E), kd); A(reinterpret_cast<void**>(&c->e_norm2), sizeof(float) * K);
    A(reinterpret_cast<void**>(&c->E_hi), kd); A(reinterpret_cast<void**>(&c->E_lo), kd);
    for (auto& l : c->lane) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreate(&l.ev_stop);
        A(reinterpret_cast<void**>(&l.z), nd); A(reinterpret_cast<void**>(&l.gq), nd);
        A(reinterpret_cast<void**>(&l.q), nd); A(reinterpret_cast<void**>(&l.dz), nd);
        A(reinterpret_cast<void**>(&l.dE), kd); A(reinterpret_cast<void**>(&l.hist), sizeof(float) * K);
        A(reinterpret_cast<void**>(&l.scal), sizeof(float) * 4);
        A(reinterpret_cast<void**>(&l.idx), sizeof(int32_t) * max_rows);
        l.ws_bytes = vq_workspace_bytes(max_rows, K, D, 0);
        A(&l.ws, l.ws_bytes);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_start);
    if (e != cudaSuccess) {
        host_ctx_free(c);
        return fail(VQ_ERR_CUDA, "vq_host_ctx_create: %s", cudaGetErrorString(e));
    }
    *out = c;
    return VQ_OK;
}

void vq_host_ctx_destroy(vq_host_ctx* ctx) { host_ctx_free(ctx); }

int vq_host_set_codebook(vq_host_ctx* c, const float* E_host) {
    if (c == nullptr || E_host == nullptr) return fail(VQ_ERR_ARG, "vq_host_set_codebook: null");
    cudaStream_t st = c->lane[0].st;
    CUDA_TRY(cudaMemcpyAsync(c->E, E_host, sizeof(float) * c->K * c->D, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));      // every step prepares the codebook itself (vq_step_forward)
    return VQ_OK;
}

int vq_host_step_async(vq_host_ctx* c, int lane, const float* z_host, const float* gq_host, int64_t n_rows,
                       int64_t n_rows_dE, float beta, int flags, float* loss_host, float* perplexity_host, int32_t* idx_host, float* q_host,
                       float* dz_host, float* dE_host) {
    if (c == nullptr || z_host == nullptr || lane < 0 || lane > 1) return fail(VQ_ERR_ARG, "vq_host_step_async: bad argument");
    if (n_rows < 1 || n_rows > c->max_rows) return fail(VQ_ERR_ARG, "vq_host_step_async: n_rows=%lld outside [1, %lld]", (long long)n_rows, c->max_rows);
    auto& l = c->lane[lane];
    const size_t nd = sizeof(float) * n_rows * c->D, kd = sizeof(float) * c->K * c->D;
    const bool train = (flags & VQ_FLAG_TRAIN_VQ) != 0;
    CUDA_TRY(cudaMemcpyAsync(l.z, z_host, nd, cudaMemcpyHostToDevice, l.st));
    if (gq_host != nullptr) {
        CUDA_TRY(cudaMemcpyAsync(l.gq, gq_host, nd, cudaMemcpyHostToDevice, l.st));
        l.gq_is_ones = false;
    } else if (!l.gq_is_ones) {
        fill_kernel<<<kNumSMs * 4, 256, 0, l.st>>>(l.gq, 1.0f, static_cast<long long>(c->max_rows) * c->D);
        LAUNCH_CHECK("fill_kernel");
        l.gq_is_ones = true;
    }
    // prepare + forward in one entry (the prepare launch also zeroes the lane's dE accumulator).  Both lanes share the codebook
    // scratch: a lane's prepare rewrites e_norm2 / E_hi / E_lo with the values the other lane's kernels may be reading -- identical bits.
    const int keep = VQ_FLAG_EXACT | VQ_FLAG_NO_SCREEN | VQ_FLAG_SCREEN | VQ_FLAG_NO_FUSE | VQ_FLAG_TC_1CTA;
    if (int rc = vq_step_forward(l.z, c->E, n_rows, c->K, c->D, beta, flags & keep, c->e_norm2, c->E_hi, c->E_lo, train ? l.dE : nullptr, l.q, l.idx, nullptr, l.hist,
                                 l.scal + 0, l.scal + 1, l.scal + 2, l.ws, l.ws_bytes, l.st))
        return rc;
    // gq_host == NULL: the lane's g_q buffer still holds the ones written at context creation, i.e. the
    // `(loss + quantized.sum()).backward()` workload; the kernel reads it like any upstream gradient.
    if (int rc = vq_step_backward(l.gq, nullptr, l.z, c->E, l.idx, n_rows, n_rows, n_rows_dE > 0 ? n_rows_dE : n_rows, c->K, c->D, beta,
                                  flags & (VQ_FLAG_TRAIN_VQ | VQ_FLAG_BWD_FLAT | VQ_FLAG_BWD_PRIVATE), l.dz, train ? l.dE : nullptr, l.ws, l.ws_bytes,
                                  flags & keep, l.st))
        return rc;
    if (loss_host) CUDA_TRY(cudaMemcpyAsync(loss_host, l.scal + 1, sizeof(float), cudaMemcpyDeviceToHost, l.st));
    if (perplexity_host) CUDA_TRY(cudaMemcpyAsync(perplexity_host, l.scal + 2, sizeof(float), cudaMemcpyDeviceToHost, l.st));
    if (idx_host) CUDA_TRY(cudaMemcpyAsync(idx_host, l.idx, sizeof(int32_t) * n_rows, cudaMemcpyDeviceToHost, l.st));
    if (q_host) CUDA_TRY(cudaMemcpyAsync(q_host, l.q, nd, cudaMemcpyDeviceToHost, l.st));
    if (dz_host) CUDA_TRY(cudaMemcpyAsync(dz_host, l.dz, nd, cudaMemcpyDeviceToHost, l.st));
    if (dE_host && train) CUDA_TRY(cudaMemcpyAsync(dE_host, l.dE, kd, cudaMemcpyDeviceToHost, l.st));
    return VQ_OK;
}

int vq_host_lane_buffers(vq_host_ctx* c, int lane, void** stream, float** dE, float** hist_sse) {
    if (c == nullptr || lane < 0 || lane > 1) return fail(VQ_ERR_ARG, "vq_host_lane_buffers: bad argument");
    if (stream) *stream = c->lane[lane].st;
    if (dE) *dE = c->lane[lane].dE;
    if (hist_sse) *hist_sse = c->lane[lane].hist;
    return VQ_OK;
}

// Device-side stopwatch over both lanes: start = both lanes idle, then an event on lane 0 that lane 1 waits for;
// stop = an event at the tail of each lane, elapsed = the later of the two.
int vq_host_timer_start(vq_host_ctx* c) {
    if (c == nullptr) return fail(VQ_ERR_ARG, "vq_host_timer_start: null");
    CUDA_TRY(cudaStreamSynchronize(c->lane[0].st));
    CUDA_TRY(cudaStreamSynchronize(c->lane[1].st));
    CUDA_TRY(cudaEventRecord(c->ev_start, c->lane[0].st));
    CUDA_TRY(cudaStreamWaitEvent(c->lane[1].st, c->ev_start, 0));
    return VQ_OK;
}

int vq_host_timer_stop_ms(vq_host_ctx* c, float* ms) {
    if (c == nullptr || ms == nullptr) return fail(VQ_ERR_ARG, "vq_host_timer_stop_ms: null");
    float best = 0.f;
    for (auto& l : c->lane) {
        CUDA_TRY(cudaEventRecord(l.ev_stop, l.st));
        CUDA_TRY(cudaEventSynchronize(l.ev_stop));
        float t = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&t, c->ev_start, l.ev_stop));
        if (t > best) best = t;
    }
    *ms = best;
    return VQ_OK;
}

int vq_host_wait(vq_host_ctx* c, int lane) {
    if (c == nullptr || lane < 0 || lane > 1) return fail(VQ_ERR_ARG, "vq_host_wait: bad argument");
    CUDA_TRY(cudaStreamSynchronize(c->lane[lane].st));
    return VQ_OK;
}

}  // extern "C"
