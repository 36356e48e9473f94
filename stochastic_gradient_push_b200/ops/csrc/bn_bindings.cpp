// torch bindings for the fused NHWC BatchNorm kernels (bn_kernels.cu)
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace py = pybind11;

extern "C" {
int bn_supported(long long M, int C);
int bn_partial_rows(long long M, int C);
int bn_partial_rows_bwd(long long M, int C);
cudaError_t bn_launch_stats(int dtype, const void* x, float* partial, long long M, int C, int G, cudaStream_t st);
cudaError_t bn_launch_stats_finalize(const float* partial, int G, long long M, int C, const float* gamma,
                                     const float* beta, float* rmean, float* rvar, long long* nbt,
                                     float momentum, float eps, float* mean, float* invstd,
                                     float* scale, float* shift, cudaStream_t st);
cudaError_t bn_launch_eval_coeff(int C, const float* gamma, const float* beta, const float* rmean,
                                 const float* rvar, float eps, float* scale, float* shift, cudaStream_t st);
cudaError_t bn_launch_apply(int dtype, int relu, int add, const void* x, const void* res,
                            const float* scale, const float* shift, void* y, long long M, int C, cudaStream_t st);
cudaError_t bn_launch_bwd_reduce(int dtype, int mode, const void* dy, const void* x, const void* y,
                                 const float* scale, const float* shift, const float* mean,
                                 const float* invstd, float* partial, void* dz, long long M, int C,
                                 int G, cudaStream_t st);
cudaError_t bn_launch_bwd_finalize(const float* partial, int G, long long M, int C, const float* scale,
                                   const float* mean, const float* invstd, float* ggamma, float* gbeta,
                                   float* c2, float* c3, cudaStream_t st);
cudaError_t bn_launch_bwd_dx(int dtype, int mode, const void* dz, const void* x, const float* scale,
                             const float* shift, const float* c2, const float* c3, void* dx,
                             long long M, int C, cudaStream_t st);
}

extern std::atomic<long long> g_sgp_kernel_launches;       // bindings.cpp

#define BN_CHECK(expr)                                                                   \
    do {                                                                                 \
        if (std::strstr(#expr, "_launch_") != nullptr) ++g_sgp_kernel_launches;          \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

static int dtype_code(const torch::Tensor& t)
{
    if (t.scalar_type() == torch::kBFloat16) return 0;
    if (t.scalar_type() == torch::kFloat32) return 1;
    throw std::runtime_error("fused BN supports bf16 / fp32 activations");
}

// x must be NHWC-dense: a 4-D channels_last tensor or a 2-D [M, C] matrix
static void nhwc_dims(const torch::Tensor& x, long long& M, int& C)
{
    if (x.dim() == 4) {
        TORCH_CHECK(x.is_contiguous(at::MemoryFormat::ChannelsLast), "fused BN needs channels_last input");
        C = (int)x.size(1);
        M = x.numel() / C;
    } else {
        TORCH_CHECK(x.dim() == 2 && x.is_contiguous());
        C = (int)x.size(1);
        M = x.size(0);
    }
}

static bool bn_can_fuse(const torch::Tensor& x)
{
    if (!x.is_cuda()) return false;
    if (x.scalar_type() != torch::kBFloat16 && x.scalar_type() != torch::kFloat32) return false;
    if (x.dim() == 4) { if (!x.is_contiguous(at::MemoryFormat::ChannelsLast)) return false; }
    else if (!(x.dim() == 2 && x.is_contiguous())) return false;
    const int C = (int)x.size(1);
    return bn_supported(x.numel() / C, C) != 0;
}

// returns (y, mean, invstd, scale, shift)
static std::vector<torch::Tensor> bn_forward(torch::Tensor x, c10::optional<torch::Tensor> residual,
                                             torch::Tensor gamma, torch::Tensor beta,
                                             c10::optional<torch::Tensor> running_mean,
                                             c10::optional<torch::Tensor> running_var,
                                             c10::optional<torch::Tensor> num_batches_tracked,
                                             bool training, double momentum, double eps, bool relu)
{
    long long M; int C;
    nhwc_dims(x, M, C);
    TORCH_CHECK(bn_supported(M, C), "unsupported channel count ", C);
    TORCH_CHECK(gamma.scalar_type() == torch::kFloat32 && beta.scalar_type() == torch::kFloat32);
    const int dt = dtype_code(x);
    c10::cuda::CUDAGuard guard(x.get_device());
    auto st = at::cuda::getCurrentCUDAStream();
    auto fopt = torch::TensorOptions().dtype(torch::kFloat32).device(x.device());
    auto y = torch::empty_like(x);
    auto coef = torch::empty({4, C}, fopt);      // mean, invstd, scale, shift
    float* mean = coef.data_ptr<float>();
    float* invstd = mean + C; float* scale = mean + 2 * C; float* shift = mean + 3 * C;
    const bool has_res = residual.has_value() && residual->defined();
    if (has_res) TORCH_CHECK(residual->sizes() == x.sizes() && residual->scalar_type() == x.scalar_type()
                             && residual->strides() == x.strides());
    if (training) {
        const int G = bn_partial_rows(M, C);
        auto partial = torch::empty({G + 1, 2, C}, fopt);   // last row carries the shift K
        BN_CHECK(bn_launch_stats(dt, x.data_ptr(), partial.data_ptr<float>(), M, C, G, st));
        float* rm = nullptr; float* rv = nullptr; long long* nbt = nullptr;
        if (running_mean.has_value() && running_mean->defined()) {
            rm = running_mean->data_ptr<float>();
            rv = running_var->data_ptr<float>();
        }
        if (num_batches_tracked.has_value() && num_batches_tracked->defined())
            nbt = reinterpret_cast<long long*>(num_batches_tracked->data_ptr<int64_t>());
        BN_CHECK(bn_launch_stats_finalize(partial.data_ptr<float>(), G, M, C, gamma.data_ptr<float>(),
                                          beta.data_ptr<float>(), rm, rv, nbt, (float)momentum,
                                          (float)eps, mean, invstd, scale, shift, st));
    } else {
        TORCH_CHECK(running_mean.has_value() && running_var.has_value());
        BN_CHECK(bn_launch_eval_coeff(C, gamma.data_ptr<float>(), beta.data_ptr<float>(),
                                      running_mean->data_ptr<float>(), running_var->data_ptr<float>(),
                                      (float)eps, scale, shift, st));
    }
    BN_CHECK(bn_launch_apply(dt, relu ? 1 : 0, has_res ? 1 : 0, x.data_ptr(),
                             has_res ? residual->data_ptr() : nullptr, scale, shift, y.data_ptr(),
                             M, C, st));
    return {y, coef};
}

// returns (dx, dres or undefined, grad_gamma, grad_beta)
static std::vector<torch::Tensor> bn_backward(torch::Tensor dy, torch::Tensor x,
                                              c10::optional<torch::Tensor> y, torch::Tensor coef,
                                              bool relu, bool add)
{
    long long M; int C;
    nhwc_dims(x, M, C);
    const int dt = dtype_code(x);
    TORCH_CHECK(dy.scalar_type() == x.scalar_type());
    if (dy.strides() != x.strides()) dy = dy.contiguous(x.dim() == 4 ? at::MemoryFormat::ChannelsLast
                                                                     : at::MemoryFormat::Contiguous);
    c10::cuda::CUDAGuard guard(x.get_device());
    auto st = at::cuda::getCurrentCUDAStream();
    auto fopt = torch::TensorOptions().dtype(torch::kFloat32).device(x.device());
    const float* mean = coef.data_ptr<float>();
    const float* invstd = mean + C; const float* scale = mean + 2 * C; const float* shift = mean + 3 * C;
    const int mode = add ? 2 : (relu ? 1 : 0);
    TORCH_CHECK(!add || relu, "residual add without ReLU is plain autograd (not fused)");
    if (mode == 2) TORCH_CHECK(y.has_value() && y->defined() && y->strides() == x.strides());
    const int G = bn_partial_rows_bwd(M, C);
    auto partial = torch::empty({G, 2, C}, fopt);
    auto grads = torch::empty({4, C}, fopt);    // grad_gamma, grad_beta, c2, c3
    float* gg = grads.data_ptr<float>();
    torch::Tensor dz;
    if (mode == 2) dz = torch::empty_like(x);
    BN_CHECK(bn_launch_bwd_reduce(dt, mode, dy.data_ptr(), x.data_ptr(),
                                  mode == 2 ? y->data_ptr() : nullptr, scale, shift, mean, invstd,
                                  partial.data_ptr<float>(), mode == 2 ? dz.data_ptr() : nullptr,
                                  M, C, G, st));
    BN_CHECK(bn_launch_bwd_finalize(partial.data_ptr<float>(), G, M, C, scale, mean, invstd,
                                    gg, gg + C, gg + 2 * C, gg + 3 * C, st));
    auto dx = torch::empty_like(x);
    BN_CHECK(bn_launch_bwd_dx(dt, mode == 1 ? 1 : 0, mode == 2 ? dz.data_ptr() : dy.data_ptr(),
                              x.data_ptr(), scale, shift, gg + 2 * C, gg + 3 * C, dx.data_ptr(),
                              M, C, st));
    return {dx, dz, grads.select(0, 0), grads.select(0, 1)};
}

// ---------------------------------------------------------------------------
// NHWC max-pool
// ---------------------------------------------------------------------------
extern "C" {
cudaError_t pool_launch_fwd(int dtype, const void* x, void* y, uint8_t* code, int N, int H, int W, int C,
                            int OH, int OW, int k, int s, int p, cudaStream_t st);
cudaError_t pool_launch_bwd(int dtype, const void* dy, const uint8_t* code, void* dx, int N, int H, int W,
                            int C, int OH, int OW, int k, int s, int p, cudaStream_t st);
}

static bool pool_can_fuse(const torch::Tensor& x, int k, int s, int p)
{
    return x.is_cuda() && x.dim() == 4 && x.is_contiguous(at::MemoryFormat::ChannelsLast)
        && (x.scalar_type() == torch::kBFloat16 || x.scalar_type() == torch::kFloat32)
        && x.size(1) % 8 == 0 && k >= 1 && k <= 11 && s >= 1 && p >= 0 && 2 * p <= k;
}

// returns (y, code)
static std::vector<torch::Tensor> maxpool_forward(torch::Tensor x, int k, int s, int p)
{
    TORCH_CHECK(pool_can_fuse(x, k, s, p));
    const int N = (int)x.size(0), C = (int)x.size(1), H = (int)x.size(2), W = (int)x.size(3);
    const int OH = (H + 2 * p - k) / s + 1, OW = (W + 2 * p - k) / s + 1;
    c10::cuda::CUDAGuard guard(x.get_device());
    auto y = torch::empty({N, C, OH, OW}, x.options().memory_format(at::MemoryFormat::ChannelsLast));
    auto code = torch::empty({N, OH, OW, C}, x.options().dtype(torch::kUInt8));
    BN_CHECK(pool_launch_fwd(dtype_code(x), x.data_ptr(), y.data_ptr(), code.data_ptr<uint8_t>(),
                             N, H, W, C, OH, OW, k, s, p, at::cuda::getCurrentCUDAStream()));
    return {y, code};
}

static torch::Tensor maxpool_backward(torch::Tensor dy, torch::Tensor code, int H, int W, int k, int s, int p)
{
    const int N = (int)dy.size(0), C = (int)dy.size(1), OH = (int)dy.size(2), OW = (int)dy.size(3);
    if (!dy.is_contiguous(at::MemoryFormat::ChannelsLast)) dy = dy.contiguous(at::MemoryFormat::ChannelsLast);
    c10::cuda::CUDAGuard guard(dy.get_device());
    auto dx = torch::empty({N, C, H, W}, dy.options().memory_format(at::MemoryFormat::ChannelsLast));
    BN_CHECK(pool_launch_bwd(dtype_code(dy), dy.data_ptr(), code.data_ptr<uint8_t>(), dx.data_ptr(),
                             N, H, W, C, OH, OW, k, s, p, at::cuda::getCurrentCUDAStream()));
    return dx;
}

// ---------------------------------------------------------------------------
// ResNet stem convolution (3 -> 64, 7x7 / 2, pad 3), NHWC bf16
// ---------------------------------------------------------------------------
extern "C" {
cudaError_t stem_launch_fwd(const void* x, const void* w, void* y, int N, int H, int W, int OH, int OW,
                            cudaStream_t st);
cudaError_t stem_launch_wgrad(const void* x, const void* dy, float* dw_acc, int N, int H, int W,
                              int OH, int OW, cudaStream_t st);
cudaError_t stem_launch_fwd_f32(const void* x, const void* w, void* y, int N, int H, int W, int OH, int OW,
                                cudaStream_t st);
cudaError_t stem_launch_wgrad_f32(const void* x, const void* dy, float* dw_acc, int N, int H, int W,
                                  int OH, int OW, cudaStream_t st);
}

static bool stem_can_fuse(const torch::Tensor& x, const torch::Tensor& w)
{
    // bf16 (mma.sync bf16) or fp32 operands (mma.sync TF32), same dtype for both
    return x.is_cuda() && x.dim() == 4 && x.size(1) == 3
        && (x.scalar_type() == torch::kBFloat16 || x.scalar_type() == torch::kFloat32)
        && x.is_contiguous(at::MemoryFormat::ChannelsLast)
        && w.dim() == 4 && w.size(0) == 64 && w.size(1) == 3 && w.size(2) == 7 && w.size(3) == 7
        && w.scalar_type() == x.scalar_type() && w.is_contiguous(at::MemoryFormat::ChannelsLast)
        && x.size(2) >= 7 && x.size(3) >= 7;
}

static torch::Tensor stem_forward(torch::Tensor x, torch::Tensor w)
{
    TORCH_CHECK(stem_can_fuse(x, w), "stem_forward: unsupported tensors");
    const int N = (int)x.size(0), H = (int)x.size(2), W = (int)x.size(3);
    const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
    c10::cuda::CUDAGuard guard(x.get_device());
    auto y = torch::empty({N, 64, OH, OW}, x.options().memory_format(at::MemoryFormat::ChannelsLast));
    if (x.scalar_type() == torch::kFloat32)
        BN_CHECK(stem_launch_fwd_f32(x.data_ptr(), w.data_ptr(), y.data_ptr(), N, H, W, OH, OW,
                                     at::cuda::getCurrentCUDAStream()));
    else
        BN_CHECK(stem_launch_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), N, H, W, OH, OW,
                                 at::cuda::getCurrentCUDAStream()));
    return y;
}

// returns dW as an fp32 [64, 3, 7, 7] tensor in channels_last layout
static torch::Tensor stem_wgrad(torch::Tensor x, torch::Tensor dy)
{
    TORCH_CHECK((x.scalar_type() == torch::kBFloat16 || x.scalar_type() == torch::kFloat32) &&
                x.is_contiguous(at::MemoryFormat::ChannelsLast));
    TORCH_CHECK(dy.scalar_type() == x.scalar_type() && dy.size(1) == 64);
    if (!dy.is_contiguous(at::MemoryFormat::ChannelsLast)) dy = dy.contiguous(at::MemoryFormat::ChannelsLast);
    const int N = (int)x.size(0), H = (int)x.size(2), W = (int)x.size(3);
    const int OH = (int)dy.size(2), OW = (int)dy.size(3);
    c10::cuda::CUDAGuard guard(x.get_device());
    auto acc = torch::zeros({64, 7, 7, 3}, x.options().dtype(torch::kFloat32));
    if (x.scalar_type() == torch::kFloat32)
        BN_CHECK(stem_launch_wgrad_f32(x.data_ptr(), dy.data_ptr(), acc.data_ptr<float>(), N, H, W, OH, OW,
                                       at::cuda::getCurrentCUDAStream()));
    else
        BN_CHECK(stem_launch_wgrad(x.data_ptr(), dy.data_ptr(), acc.data_ptr<float>(), N, H, W, OH, OW,
                                   at::cuda::getCurrentCUDAStream()));
    return acc.permute({0, 3, 1, 2});
}

// ---------------------------------------------------------------------------
// 1x1 convolution as a tcgen05 GEMM, BatchNorm statistics fused into its epilogue
// (conv1x1_kernels.cu)
// ---------------------------------------------------------------------------
extern "C" {
int c1_supported(long long M, int N, int K, int f32);
int c1_partial_rows(long long M, int N, int K, int num_sms, int f32);
void c1_describe_plan(long long M, int N, int K, int num_sms, int residual, int f32, int* out);
cudaError_t c1_launch_gemm(const void* x, const void* w, void* y, long long M, int N, int K, float* partial,
                           const void* residual, int num_sms, int f32, cudaStream_t st);
cudaError_t c1_launch_stats_finalize(const float* partial, int R, int C, const float* gamma, const float* beta,
                                     float* rmean, float* rvar, long long* nbt, float momentum, float eps,
                                     float* mean, float* invstd, float* scale, float* shift, cudaStream_t st);
}

// x: NHWC bf16 / fp32 activation (4-D channels_last) or [M, K] matrix; w: [N, K, 1, 1] or [N, K], rows
// contiguous, same dtype as x (fp32 operands are multiplied as TF32, fp32 accumulate)
static bool conv1x1_can_fuse(const torch::Tensor& x, const torch::Tensor& w)
{
    if (!x.is_cuda() || !w.is_cuda() || x.scalar_type() != w.scalar_type() ||
        (x.scalar_type() != torch::kBFloat16 && x.scalar_type() != torch::kFloat32))
        return false;
    if (x.dim() == 4) { if (!x.is_contiguous(at::MemoryFormat::ChannelsLast)) return false; }
    else if (!(x.dim() == 2 && x.is_contiguous())) return false;
    if (w.dim() == 4) { if (w.size(2) != 1 || w.size(3) != 1) return false; }
    else if (w.dim() != 2) return false;
    const int K = (int)x.size(1), N = (int)w.size(0);
    if (w.size(1) != K || w.stride(1) != 1 || w.stride(0) != K) return false;
    if ((reinterpret_cast<uintptr_t>(x.data_ptr()) | reinterpret_cast<uintptr_t>(w.data_ptr())) & 15) return false;
    return c1_supported(x.numel() / K, N, K, x.scalar_type() == torch::kFloat32) != 0;
}

static torch::Tensor conv1x1_alloc_out(const torch::Tensor& x, int N)
{
    if (x.dim() == 4)
        return torch::empty({x.size(0), N, x.size(2), x.size(3)},
                            x.options().memory_format(at::MemoryFormat::ChannelsLast));
    return torch::empty({x.size(0), N}, x.options());
}

// host-side launch plan of the GEMM for a shape (no GPU needed)
static py::dict conv1x1_plan(long long M, int N, int K, int num_sms, bool residual, bool f32)
{
    TORCH_CHECK(c1_supported(M, N, K, f32), "unsupported GEMM shape");
    int v[7];
    c1_describe_plan(M, N, K, num_sms, residual ? 1 : 0, f32 ? 1 : 0, v);
    py::dict d;
    d["block_n"] = v[0]; d["grid"] = v[1]; d["ctas_per_n"] = v[2]; d["stages"] = v[3];
    d["resident_w"] = v[4] != 0; d["store_slabs"] = v[5]; d["smem_bytes"] = v[6];
    d["partial_rows"] = c1_partial_rows(M, N, K, num_sms, f32 ? 1 : 0);
    return d;
}

static int sm_count() { return at::cuda::getCurrentDeviceProperties()->multiProcessorCount; }

// y = x . w^T [+ residual] (residual: same shape / layout as y, added in fp32 in the epilogue).
// with_stats: also run the statistics epilogue into a scratch buffer (benchmarking the GEMM alone)
static torch::Tensor conv1x1_forward(torch::Tensor x, torch::Tensor w, bool with_stats,
                                     c10::optional<torch::Tensor> residual)
{
    TORCH_CHECK(conv1x1_can_fuse(x, w), "conv1x1_forward: unsupported tensors");
    const int K = (int)x.size(1), N = (int)w.size(0);
    const long long M = x.numel() / K;
    c10::cuda::CUDAGuard guard(x.get_device());
    auto y = conv1x1_alloc_out(x, N);
    const bool has_res = residual.has_value() && residual->defined();
    TORCH_CHECK(!(has_res && with_stats));
    const int f32 = x.scalar_type() == torch::kFloat32 ? 1 : 0;
    if (has_res)
        TORCH_CHECK(residual->is_cuda() && residual->scalar_type() == x.scalar_type() &&
                    residual->sizes() == y.sizes() && residual->strides() == y.strides() &&
                    (reinterpret_cast<uintptr_t>(residual->data_ptr()) & 15) == 0,
                    "conv1x1_forward: residual must match the output's shape and layout");
    torch::Tensor partial;
    if (with_stats)
        partial = torch::empty({c1_partial_rows(M, N, K, sm_count(), f32), 3, N},
                               torch::TensorOptions().dtype(torch::kFloat32).device(x.device()));
    BN_CHECK(c1_launch_gemm(x.data_ptr(), w.data_ptr(), y.data_ptr(), M, N, K,
                            with_stats ? partial.data_ptr<float>() : nullptr,
                            has_res ? residual->data_ptr() : nullptr, sm_count(), f32,
                            at::cuda::getCurrentCUDAStream()));
    return y;
}

// training-mode  out = act(bn(conv1x1(x, w)) [+ residual]);  returns (conv output, out, coef)
static std::vector<torch::Tensor> conv1x1_bn_forward(torch::Tensor x, torch::Tensor w,
                                                     c10::optional<torch::Tensor> residual,
                                                     torch::Tensor gamma, torch::Tensor beta,
                                                     c10::optional<torch::Tensor> running_mean,
                                                     c10::optional<torch::Tensor> running_var,
                                                     c10::optional<torch::Tensor> num_batches_tracked,
                                                     double momentum, double eps, bool relu)
{
    TORCH_CHECK(conv1x1_can_fuse(x, w), "conv1x1_bn_forward: unsupported tensors");
    const int K = (int)x.size(1), N = (int)w.size(0);
    const long long M = x.numel() / K;
    TORCH_CHECK(bn_supported(M, N), "unsupported channel count ", N);
    TORCH_CHECK(gamma.scalar_type() == torch::kFloat32 && beta.scalar_type() == torch::kFloat32);
    c10::cuda::CUDAGuard guard(x.get_device());
    auto st = at::cuda::getCurrentCUDAStream();
    auto fopt = torch::TensorOptions().dtype(torch::kFloat32).device(x.device());
    auto yraw = conv1x1_alloc_out(x, N);
    const bool has_res = residual.has_value() && residual->defined();
    if (has_res) TORCH_CHECK(residual->sizes() == yraw.sizes() && residual->scalar_type() == yraw.scalar_type()
                             && residual->strides() == yraw.strides());
    const int sms = sm_count();
    const int f32 = x.scalar_type() == torch::kFloat32 ? 1 : 0;
    const int R = c1_partial_rows(M, N, K, sms, f32);
    auto partial = torch::empty({R, 3, N}, fopt);
    BN_CHECK(c1_launch_gemm(x.data_ptr(), w.data_ptr(), yraw.data_ptr(), M, N, K, partial.data_ptr<float>(),
                            nullptr, sms, f32, st));
    auto coef = torch::empty({4, N}, fopt);      // mean, invstd, scale, shift
    float* mean = coef.data_ptr<float>();
    float* rm = nullptr; float* rv = nullptr; long long* nbt = nullptr;
    if (running_mean.has_value() && running_mean->defined()) {
        rm = running_mean->data_ptr<float>();
        rv = running_var->data_ptr<float>();
    }
    if (num_batches_tracked.has_value() && num_batches_tracked->defined())
        nbt = reinterpret_cast<long long*>(num_batches_tracked->data_ptr<int64_t>());
    BN_CHECK(c1_launch_stats_finalize(partial.data_ptr<float>(), R, N, gamma.data_ptr<float>(),
                                      beta.data_ptr<float>(), rm, rv, nbt, (float)momentum, (float)eps,
                                      mean, mean + N, mean + 2 * N, mean + 3 * N, st));
    auto out = torch::empty_like(yraw);
    BN_CHECK(bn_launch_apply(f32, relu ? 1 : 0, has_res ? 1 : 0, yraw.data_ptr(),
                             has_res ? residual->data_ptr() : nullptr, mean + 2 * N, mean + 3 * N,
                             out.data_ptr(), M, N, st));
    return {yraw, out, coef};
}

// ---------------------------------------------------------------------------
// fused softmax cross-entropy + top-1 / top-5 accuracy (csrc/loss_kernels.cu)
// ---------------------------------------------------------------------------
extern "C" {
cudaError_t xent_launch_fwd(int dtype, const void* logits, const long long* target, float* lse, float* out3, int B,
                            int C, cudaStream_t st);
cudaError_t xent_launch_bwd(int dtype, const void* logits, const long long* target, const float* lse,
                            const float* grad_out, void* dlogits, int B, int C, cudaStream_t st);
}

static bool xent_can_fuse(const torch::Tensor& logits, const torch::Tensor& target)
{
    return logits.is_cuda() && target.is_cuda() && logits.dim() == 2 && logits.is_contiguous() &&
           (logits.scalar_type() == torch::kFloat32 || logits.scalar_type() == torch::kBFloat16) &&
           target.dim() == 1 && target.scalar_type() == torch::kInt64 && target.is_contiguous() &&
           target.size(0) == logits.size(0) && logits.size(0) > 0 && logits.size(1) > 0;
}

// returns (metrics[3] = mean loss, prec@1 %, prec@5 %; lse[B])
static std::vector<torch::Tensor> xent_forward(torch::Tensor logits, torch::Tensor target)
{
    TORCH_CHECK(xent_can_fuse(logits, target), "xent_forward: unsupported tensors");
    c10::cuda::CUDAGuard guard(logits.get_device());
    auto st = at::cuda::getCurrentCUDAStream();
    auto fopt = torch::TensorOptions().dtype(torch::kFloat32).device(logits.device());
    const int B = (int)logits.size(0), C = (int)logits.size(1);
    auto out3 = torch::empty({3}, fopt);
    auto lse = torch::empty({B}, fopt);
    BN_CHECK(cudaMemsetAsync(out3.data_ptr(), 0, 3 * sizeof(float), st));
    BN_CHECK(xent_launch_fwd(dtype_code(logits), logits.data_ptr(),
                             reinterpret_cast<const long long*>(target.data_ptr<int64_t>()),
                             lse.data_ptr<float>(), out3.data_ptr<float>(), B, C, st));
    return {out3, lse};
}

static torch::Tensor xent_backward(torch::Tensor logits, torch::Tensor target, torch::Tensor lse,
                                   torch::Tensor grad_out)
{
    TORCH_CHECK(xent_can_fuse(logits, target) && lse.is_cuda() && lse.scalar_type() == torch::kFloat32 &&
                grad_out.is_cuda() && grad_out.scalar_type() == torch::kFloat32 && grad_out.numel() == 1);
    c10::cuda::CUDAGuard guard(logits.get_device());
    auto dl = torch::empty_like(logits);
    BN_CHECK(xent_launch_bwd(dtype_code(logits), logits.data_ptr(),
                             reinterpret_cast<const long long*>(target.data_ptr<int64_t>()), lse.data_ptr<float>(),
                             grad_out.data_ptr<float>(), dl.data_ptr(), (int)logits.size(0), (int)logits.size(1),
                             at::cuda::getCurrentCUDAStream()));
    return dl;
}

void bind_bn(py::module& mod)
{
    mod.def("xent_can_fuse", &xent_can_fuse);
    mod.def("xent_forward", &xent_forward);
    mod.def("xent_backward", &xent_backward);
    mod.def("conv1x1_can_fuse", &conv1x1_can_fuse);
    mod.def("conv1x1_plan", &conv1x1_plan, py::arg("M"), py::arg("N"), py::arg("K"), py::arg("num_sms") = 148,
            py::arg("residual") = false, py::arg("f32") = false);
    mod.def("conv1x1_forward", &conv1x1_forward, py::arg("x"), py::arg("w"), py::arg("with_stats") = false,
            py::arg("residual") = py::none());
    mod.def("conv1x1_bn_forward", &conv1x1_bn_forward);
    mod.def("stem_can_fuse", &stem_can_fuse);
    mod.def("stem_forward", &stem_forward);
    mod.def("stem_wgrad", &stem_wgrad);
    mod.def("pool_can_fuse", &pool_can_fuse);
    mod.def("maxpool_forward", &maxpool_forward);
    mod.def("maxpool_backward", &maxpool_backward);
    mod.def("bn_can_fuse", &bn_can_fuse);
    mod.def("bn_forward", &bn_forward);
    mod.def("bn_backward", &bn_backward);
}
