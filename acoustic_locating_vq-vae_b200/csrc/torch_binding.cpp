// torch_binding.cpp -- the host side of the drop-in nn.Module in C++: one autograd node around the C ABI.
//
// The reference's hot path sits behind a Python nn.Module (vector_quantizer.py:8-58).  Driving libb200vq.so from Python
// (ctypes, six allocations, a Python autograd.Function) costs ~190 us of host time per forward + backward -- 2.7x the
// 70 us the kernels need -- so the eager module was host-bound.  This file is the same logic as
// quantizer.py's _VQFunction, compiled: allocations through ATen, raw pointers into include/b200vq.h, a
// torch::autograd::Function for the backward.  PyTorch still only owns buffers, streams and the autograd graph; every
// number is computed by the CUDA kernels of libb200vq.so.  No fallback: errors of the C ABI become c10::Error.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "../../include/b200vq.h"

namespace {

inline void check(int rc, const char* what) {
    TORCH_CHECK(rc == VQ_OK, "b200vq error ", rc, " in ", what, ": ", vq_last_error());
}
inline float* fptr(const at::Tensor& t) { return t.defined() ? t.data_ptr<float>() : nullptr; }

// Persistent per-module scratch (held by the Python module, passed in): e_norm2 (K), E_hi (K,D), E_lo (K,D), workspace (bytes).
struct VQFunction : public torch::autograd::Function<VQFunction> {
    // returns {loss, quantized, perplexity, encodings (empty when not asked for), indices, stats, reduced}
    //   stats    [dE slot (K*D, only when `pack`) | usage histogram (K) | squared error]: the forward writes its statistics
    //            straight behind the slot the backward fills, so data parallel all-reduces the buffer as it stands
    //   reduced  (pack only) K*D + K + 1 floats the backward's all-reduce fills: [:K*D] becomes the codebook gradient
    static torch::autograd::variable_list forward(torch::autograd::AutogradContext* ctx, const at::Tensor& inputs, const at::Tensor& weight,
                                                  double beta, int64_t flags, bool want_onehot, bool train_vq, int64_t world,
                                                  int64_t dp_ctx /* vq_dp_ctx* or 0 */, bool pack, at::Tensor e_norm2, at::Tensor e_hi,
                                                  at::Tensor e_lo, at::Tensor ws) {
        const int64_t K = weight.size(0), D = weight.size(1);
        const int64_t N = inputs.numel() / D;
        const c10::cuda::CUDAGuard guard(inputs.device());
        auto stream = at::cuda::getCurrentCUDAStream().stream();
        const auto opts = inputs.options();
        at::Tensor w = weight.detach();
        if (!w.is_contiguous()) w = w.contiguous();
        at::Tensor q_out = at::empty_like(inputs);
        at::Tensor idx = at::empty({N}, opts.dtype(at::kInt));
        at::Tensor loss = at::empty({}, opts), perplexity = at::empty({}, opts);
        const int64_t off = pack ? K * D : 0;
        // the codebook gradient's accumulator is zeroed by the forward's prepare launch (no memset in front of the backward):
        // the front of the packed buffer under data parallelism, a buffer of its own otherwise
        const bool will_train = train_vq && weight.requires_grad();     // (grad mode is off inside Function::forward)
        at::Tensor stats = at::empty({off + K + 1}, opts);
        at::Tensor dE_acc = (will_train && !pack) ? at::empty({K, D}, opts) : at::Tensor();
        float* dE_zero = pack ? stats.data_ptr<float>() : fptr(dE_acc);
        at::Tensor reduced = at::empty({pack ? K * D + K + 1 : 0}, opts);
        at::Tensor onehot = at::empty({want_onehot ? N : 0, want_onehot ? K : 0}, opts);
        float* sp = stats.data_ptr<float>() + off;
        check(vq_step_forward(inputs.data_ptr<float>(), w.data_ptr<float>(), N, static_cast<int>(K), static_cast<int>(D), static_cast<float>(beta),
                              static_cast<int>(flags) | (want_onehot ? VQ_FLAG_ONEHOT : 0), fptr(e_norm2), fptr(e_hi), fptr(e_lo), dE_zero,
                              q_out.data_ptr<float>(), idx.data_ptr<int>(), want_onehot ? onehot.data_ptr<float>() : nullptr, sp, sp + K,
                              loss.data_ptr<float>(), perplexity.data_ptr<float>(), ws.data_ptr(), static_cast<size_t>(ws.numel()), stream),
              "vq_step_forward");
        ctx->save_for_backward({inputs, weight, idx});
        ctx->saved_data["stats"] = stats;
        ctx->saved_data["reduced"] = reduced;
        ctx->saved_data["dE_acc"] = dE_acc;
        ctx->saved_data["dE_fresh"] = true;       // the accumulator is zero until the first backward has used it
        ctx->saved_data["beta"] = beta;
        ctx->saved_data["train_vq"] = train_vq;
        ctx->saved_data["world"] = world;
        ctx->saved_data["dp_ctx"] = dp_ctx;
        ctx->saved_data["pack"] = pack;
        ctx->mark_non_differentiable({perplexity, idx, onehot, stats, reduced});
        // only loss and quantized carry gradients: without this the engine hands the backward ZERO tensors for every other
        // output -- a 211 MB fill for the dense one-hot's "gradient" alone (29 us per step on the bench workload)
        ctx->set_materialize_grads(false);
        return {loss, q_out, perplexity, onehot, idx, stats, reduced};
    }

    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx, torch::autograd::variable_list grads) {
        const auto saved = ctx->get_saved_variables();
        const at::Tensor &inputs = saved[0], &weight = saved[1], &idx = saved[2];
        const int64_t K = weight.size(0), D = weight.size(1), N = idx.size(0);
        const double beta = ctx->saved_data["beta"].toDouble();
        const bool train_vq = ctx->saved_data["train_vq"].toBool(), pack = ctx->saved_data["pack"].toBool();
        const int64_t world = ctx->saved_data["world"].toInt();
        vq_dp_ctx* dp = reinterpret_cast<vq_dp_ctx*>(ctx->saved_data["dp_ctx"].toInt());
        at::Tensor packed = ctx->saved_data["stats"].toTensor();
        at::Tensor reduced = ctx->saved_data["reduced"].toTensor();
        const c10::cuda::CUDAGuard guard(inputs.device());
        auto stream = at::cuda::getCurrentCUDAStream().stream();
        const auto opts = inputs.options();
        at::Tensor g_loss = grads[0], g_q = grads[1];
        if (!g_loss.defined()) g_loss = at::zeros({}, opts);
        else if (g_loss.scalar_type() != at::kFloat) g_loss = g_loss.to(at::kFloat);
        if (g_q.defined()) {
            g_q = g_q.contiguous();
            if (g_q.scalar_type() != at::kFloat) g_q = g_q.to(at::kFloat);
        }
        const bool need_dz = ctx->needs_input_grad(0), need_dE = ctx->needs_input_grad(1) && train_vq;
        at::Tensor w = weight.detach();
        if (!w.is_contiguous()) w = w.contiguous();
        at::Tensor dz, dE;
        if (need_dz) dz = at::empty_like(inputs);
        const int64_t n_dz = N > 0 ? N : 1, n_dE = n_dz * world;
        torch::autograd::variable_list out(13);
        if (!need_dE) {
            if (need_dz)
                check(vq_backward(fptr(g_q), g_loss.data_ptr<float>(), inputs.data_ptr<float>(), w.data_ptr<float>(), idx.data_ptr<int>(), N, n_dz,
                                  n_dE, static_cast<int>(K), static_cast<int>(D), static_cast<float>(beta), 0, fptr(dz), nullptr, stream),
                      "vq_backward");
            out[0] = dz;
            return out;
        }
        // a second backward through the same graph (retain_graph) finds the accumulator used: zero it with a memset then
        const bool fresh = ctx->saved_data["dE_fresh"].toBool();
        ctx->saved_data["dE_fresh"] = false;
        if (world == 1 || !pack) {
            dE = ctx->saved_data["dE_acc"].toTensor();
            const bool zeroed = fresh && dE.defined();
            if (!zeroed) dE = at::empty({K, D}, opts);
            check(vq_backward(fptr(g_q), g_loss.data_ptr<float>(), inputs.data_ptr<float>(), w.data_ptr<float>(), idx.data_ptr<int>(), N, n_dz, n_dE,
                              static_cast<int>(K), static_cast<int>(D), static_cast<float>(beta), VQ_FLAG_TRAIN_VQ | (zeroed ? 0 : VQ_FLAG_ZERO_DE), fptr(dz),
                              dE.data_ptr<float>(), stream),
                  "vq_backward");
            out[0] = dz;
            out[1] = dE;      // world > 1 without a packed buffer (codebook frozen at forward time): the Python side all-reduces
            return out;
        }
        // data parallel: dE lands in front of the statistics the forward left in the packed buffer; ONE all-reduce
        check(vq_backward(fptr(g_q), g_loss.data_ptr<float>(), inputs.data_ptr<float>(), w.data_ptr<float>(), idx.data_ptr<int>(), N, n_dz, n_dE,
                          static_cast<int>(K), static_cast<int>(D), static_cast<float>(beta), VQ_FLAG_TRAIN_VQ | (fresh ? 0 : VQ_FLAG_ZERO_DE), fptr(dz),
                          packed.data_ptr<float>(), stream),
              "vq_backward");
        out[0] = dz;
        if (dp != nullptr) {
            // `reduced` was allocated by this step's forward (the module reads the global statistics from its tail)
            check(vq_dp_allreduce(dp, packed.data_ptr<float>(), reduced.data_ptr<float>(), stream), "vq_dp_allreduce");
            out[1] = reduced.narrow(0, 0, K * D).view({K, D});
        } else {
            out[1] = packed.narrow(0, 0, K * D).view({K, D});             // no NVLink exchange: the Python side all-reduces (NCCL)
        }
        return out;
    }
};

std::vector<at::Tensor> vq_apply(const at::Tensor& inputs, const at::Tensor& weight, double beta, int64_t flags, bool want_onehot, bool train_vq,
                                 int64_t world, int64_t dp_ctx, bool pack, at::Tensor e_norm2, at::Tensor e_hi, at::Tensor e_lo, at::Tensor ws) {
    return VQFunction::apply(inputs, weight, beta, flags, want_onehot, train_vq, world, dp_ctx, pack, e_norm2, e_hi, e_lo, ws);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "b200vq: C++ autograd node of the drop-in VectorQuantizer (calls libb200vq.so through include/b200vq.h)";
    m.def("vq_apply", &vq_apply, "forward (+ autograd backward) of the VectorQuantizer through the C ABI");
    m.def("workspace_bytes", [](int64_t n, int64_t K, int64_t D) { return static_cast<int64_t>(vq_workspace_bytes(n, static_cast<int>(K), static_cast<int>(D), 0)); });
    m.def("abi_version", []() { return vq_abi_version(); });
}
