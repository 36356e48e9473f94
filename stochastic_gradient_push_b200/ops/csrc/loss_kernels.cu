// Fused softmax cross-entropy + top-1 / top-5 accuracy (forward) and its gradient (backward).
//
// The reference computes, every iteration, nn.CrossEntropyLoss (log_softmax + nll_loss kernels),
// then `accuracy(output, target, topk=(1, 5))` (topk + transpose + eq + 2 x sum + mul, ~10
// launches) and reads the three scalars back with three `.item()` host syncs
// (/root/reference/gossip_sgd.py:372-373, 394-399, 192-198).  Here one launch produces all three
// numbers on the device (mean loss, prec@1 %, prec@5 %) plus the per-row log-sum-exp the backward
// kernel needs; the training step copies the 12 bytes to a pinned ring without a host sync.
//
//   forward : one warp per row:  m = max_j z_j ; lse = m + log sum_j exp(z_j - m)
//             loss_i = lse - z_t ; rank_i = #{ j : z_j > z_t }   (top-k hit <=> rank < k)
//             out[0] += loss_i / B ; out[1] += 100/B [rank == 0] ; out[2] += 100/B [rank < 5]
//   backward: dz_ij = g * (exp(z_ij - lse_i) - [j == t_i]) / B
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int XE_THREADS = 256;                 // 8 rows per CTA

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

template <typename T>
__global__ void __launch_bounds__(XE_THREADS)
xent_fwd_kernel(const T* __restrict__ logits, const long long* __restrict__ target, float* __restrict__ lse_out,
                float* __restrict__ out3, int B, int C)
{
    __shared__ float s_acc[3];
    if (threadIdx.x < 3) s_acc[threadIdx.x] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (XE_THREADS / 32) + (threadIdx.x >> 5);
    if (row < B) {
        const T* z = logits + (size_t)row * C;
        long long t = target[row];
        const bool valid = t >= 0 && t < C;
        if (!valid) t = 0;
        const float zt = to_f<T>(z[t]);
        // online softmax: one pass for max + sum, rank counted in the same pass
        float m = -INFINITY, s = 0.f, above = 0.f;
        for (int j = lane; j < C; j += 32) {
            const float v = to_f<T>(z[j]);
            const float mn = fmaxf(m, v);
            s = s * __expf(m - mn) + __expf(v - mn);
            m = mn;
            above += (v > zt) ? 1.f : 0.f;
        }
        const float mw = warp_max(m);
        s = warp_sum(s * __expf(m - mw));
        above = warp_sum(above);
        const float lse = mw + __logf(s);
        if (lane == 0) {
            lse_out[row] = lse;
            if (valid) {
                const float inv_b = 1.f / (float)B;
                atomicAdd(&s_acc[0], (lse - zt) * inv_b);
                if (above < 0.5f) atomicAdd(&s_acc[1], 100.f * inv_b);
                if (above < 4.5f) atomicAdd(&s_acc[2], 100.f * inv_b);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(out3 + threadIdx.x, s_acc[threadIdx.x]);
}

template <typename T>
__global__ void __launch_bounds__(XE_THREADS)
xent_bwd_kernel(const T* __restrict__ logits, const long long* __restrict__ target, const float* __restrict__ lse,
                const float* __restrict__ grad_out, T* __restrict__ dlogits, int B, int C)
{
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (XE_THREADS / 32) + (threadIdx.x >> 5);
    if (row >= B) return;
    const float g = grad_out[0] / (float)B;
    const T* z = logits + (size_t)row * C;
    T* dz = dlogits + (size_t)row * C;
    const long long t = target[row];
    const bool valid = t >= 0 && t < C;
    const float l = lse[row];
    for (int j = lane; j < C; j += 32) {
        float p = __expf(to_f<T>(z[j]) - l);
        if (j == t) p -= 1.f;
        dz[j] = from_f<T>(valid ? g * p : 0.f);
    }
}

}  // namespace

extern "C" {

// dtype: 0 = bf16 logits, 1 = fp32 logits; out3 must be zeroed by the caller (same stream)
cudaError_t xent_launch_fwd(int dtype, const void* logits, const long long* target, float* lse, float* out3, int B,
                            int C, cudaStream_t st)
{
    const int grid = (B + XE_THREADS / 32 - 1) / (XE_THREADS / 32);
    if (dtype == 0)
        xent_fwd_kernel<__nv_bfloat16><<<grid, XE_THREADS, 0, st>>>((const __nv_bfloat16*)logits, target, lse, out3, B, C);
    else
        xent_fwd_kernel<float><<<grid, XE_THREADS, 0, st>>>((const float*)logits, target, lse, out3, B, C);
    return cudaGetLastError();
}

cudaError_t xent_launch_bwd(int dtype, const void* logits, const long long* target, const float* lse,
                            const float* grad_out, void* dlogits, int B, int C, cudaStream_t st)
{
    const int grid = (B + XE_THREADS / 32 - 1) / (XE_THREADS / 32);
    if (dtype == 0)
        xent_bwd_kernel<__nv_bfloat16><<<grid, XE_THREADS, 0, st>>>((const __nv_bfloat16*)logits, target, lse, grad_out,
                                                                   (__nv_bfloat16*)dlogits, B, C);
    else
        xent_bwd_kernel<float><<<grid, XE_THREADS, 0, st>>>((const float*)logits, target, lse, grad_out,
                                                           (float*)dlogits, B, C);
    return cudaGetLastError();
}

}  // extern "C"
