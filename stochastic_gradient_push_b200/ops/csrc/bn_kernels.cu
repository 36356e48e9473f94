// Fused NHWC BatchNorm (+residual add) (+ReLU), training forward and backward,
// for sm_100a.  An ncu launch list of the ResNet-50 step (profiles/) showed the
// framework's native channels-last BatchNorm + separate ReLU / add kernels
// taking ~70 % of the step at batch 256 while the tensor-core convolutions took
// ~15 %: these kernels are memory-bound, so the fix is fewer passes over HBM.
//
//   forward :  stats (1 read of x)  ->  finalize (C threads)  ->  apply
//              y = relu(x*scale + shift [+ res])   (1 read of x [+res], 1 write)
//   backward:  reduce (dy, x [,y]) -> finalize -> dx = c1*dz - c2*x + c3
//              the ReLU mask is recomputed from x*scale+shift (no mask tensor, no
//              extra read); with a residual the mask comes from the saved output
//              and dz is written once because it IS the residual branch's grad.
//
// Layout: x is [M, C] row-major (M = N*H*W, NHWC), C % 8 == 0, C <= 2048.
// One thread owns 8 consecutive channels (16 B of bf16 / 32 B of fp32) and
// strides over rows, so per-channel constants live in registers and every
// global access is a full 16-byte vector.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "vec_io.cuh"

#define BN_THREADS 256
#define BN_VEC 8
#define BN_UNROLL 4
// rows in flight per thread and tensor.  A thread's vector is 8 channels = 16 B of bf16 but 32 B
// of fp32 (8 registers per load), so fp32 uses half the unroll: the same bytes in flight and the
// same register footprint as bf16.  (With the bf16 unroll the fp32 instantiations spilled 250-460
// bytes per thread under their 64 / 80-register caps and ran at 58-71 % of the HBM roofline
// instead of 78-94 %: profiles/launches_r2_fp32_bs256_before_tf32_gemm_summary.json.)
template <typename T> struct Unroll {
    static constexpr int N = sizeof(T) == 4 ? 2 : 4;        // stats / apply / dx
    static constexpr int RED = sizeof(T) == 4 ? 1 : 2;      // backward reduce (up to 3 tensors)
};

namespace {

// thread -> (channel vector, row lane) mapping shared by all kernels
struct Map {
    int tpr;        // threads per row  = C / 8
    int rpi;        // rows per CTA iteration = 256 / tpr
    int cv;         // this thread's channel-vector index
    int rl;         // this thread's row lane
    bool active;
    __device__ __forceinline__ Map(int C) {
        tpr = C / BN_VEC;
        rpi = BN_THREADS / tpr;
        cv = threadIdx.x % tpr;
        rl = threadIdx.x / tpr;
        active = rl < rpi;
    }
};

// block reduce of 2 x 8 per-thread accumulators across the row lanes that share
// a channel vector, then one partial row per CTA: partial[blockIdx][2][C]
__device__ __forceinline__ void reduce_store_partials(const Map& mp, const F8& a, const F8& b,
                                                      float* __restrict__ partial, int C)
{
    __shared__ float sm[BN_THREADS * 16];
    float* mine = sm + threadIdx.x * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) { mine[i] = mp.active ? a.v[i] : 0.f; mine[8 + i] = mp.active ? b.v[i] : 0.f; }
    __syncthreads();
    if (threadIdx.x < mp.tpr) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        for (int r = 0; r < mp.rpi; ++r) {
            const float* o = sm + (r * mp.tpr + threadIdx.x) * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] += o[i];
        }
        float* out = partial + (size_t)blockIdx.x * 2 * C + threadIdx.x * BN_VEC;
#pragma unroll
        for (int i = 0; i < 8; ++i) { out[i] = acc[i]; out[C + i] = acc[8 + i]; }
    }
}

}  // namespace

// ---------------------------------------------------------------------------
// forward: statistics
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BN_THREADS, 4)
bn_stats_kernel(const T* __restrict__ x, float* __restrict__ partial, long long M, int C)
{
    const Map mp(C);
    F8 s, q;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.v[i] = 0.f; q.v[i] = 0.f; }
    if (mp.active) {
        const long long stride = (long long)gridDim.x * mp.rpi;
        long long r = (long long)blockIdx.x * mp.rpi + mp.rl;
        const T* base = x + mp.cv * BN_VEC;
        // shifted sums: accumulate (x - K) with K = x[row 0] so that
        // var = E[(x-K)^2] - E[x-K]^2 does not cancel when |mean| >> std
        const F8 k = Io<T>::load(base);
        if (blockIdx.x == 0 && mp.rl == 0) {
            float* krow = partial + (size_t)gridDim.x * 2 * C + mp.cv * BN_VEC;
#pragma unroll
            for (int i = 0; i < 8; ++i) krow[i] = k.v[i];
        }
        constexpr int UN = Unroll<T>::N;
        for (; r + (UN - 1) * stride < M; r += UN * stride) {
            typename Io<T>::raw_t raw[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) raw[u] = Io<T>::load_raw(base + (r + u * stride) * C);
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const F8 d = Io<T>::decode(raw[u]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float c = d.v[i] - k.v[i];
                    s.v[i] += c; q.v[i] = fmaf(c, c, q.v[i]);
                }
            }
        }
        for (; r < M; r += stride) {
            const F8 d = Io<T>::load(base + r * C);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float c = d.v[i] - k.v[i];
                s.v[i] += c; q.v[i] = fmaf(c, c, q.v[i]);
            }
        }
    }
    reduce_store_partials(mp, s, q, partial, C);
}

// Finalize kernels: a CTA owns 32 channels; its 1024 threads are 32 partial-row
// lanes x 32 channels, so the G (<= 592) partial rows are summed with ~G/128
// dependent steps per thread instead of G (a one-thread-per-channel loop over G
// L2-latency-bound loads took ~78 us per launch -- longer than the streaming
// kernels it finalises).
#define BN_FIN_THREADS 1024
#define BN_FIN_ILP 10          // rows per lane in flight (G <= 592 -> 19 rows per lane: two batches)

__device__ __forceinline__ void sum_partials(const float* __restrict__ partial, int G, int C,
                                             float& s_out, float& q_out, bool& owner, int& ch)
{
    __shared__ float sm_s[32][33];
    __shared__ float sm_q[32][33];
    const int cl = threadIdx.x & 31;
    const int lane = threadIdx.x >> 5;
    ch = blockIdx.x * 32 + cl;
    // All of a lane's rows are requested before any is consumed: the partials sit in L2, each
    // dependent load costs ~0.7 us, and a two-deep loop made these kernels 8-13 us each --
    // 106 launches, ~5 % of the training step (profiles/launches_r1_bs256_final.csv).
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    if (ch < C) {
        for (int g0 = lane; g0 < G; g0 += 32 * BN_FIN_ILP) {
            float sv[BN_FIN_ILP], qv[BN_FIN_ILP];
#pragma unroll
            for (int u = 0; u < BN_FIN_ILP; ++u) {
                const int g = g0 + 32 * u;
                const bool ok = g < G;
                sv[u] = ok ? partial[(size_t)g * 2 * C + ch] : 0.f;
                qv[u] = ok ? partial[(size_t)g * 2 * C + C + ch] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < BN_FIN_ILP; u += 2) {
                s0 += sv[u]; q0 += qv[u];
                s1 += sv[u + 1]; q1 += qv[u + 1];
            }
        }
    }
    sm_s[lane][cl] = s0 + s1;
    sm_q[lane][cl] = q0 + q1;
    __syncthreads();
    owner = (lane == 0) && (ch < C);
    float s = 0.f, q = 0.f;
    if (owner) {
#pragma unroll
        for (int l = 0; l < 32; ++l) { s += sm_s[l][cl]; q += sm_q[l][cl]; }
    }
    s_out = s;
    q_out = q;
}

__global__ void __launch_bounds__(BN_FIN_THREADS)
bn_stats_finalize_kernel(const float* __restrict__ partial, int G, long long M, int C,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* running_mean, float* running_var, long long* nbt,
                         float momentum, float eps,
                         float* __restrict__ mean, float* __restrict__ invstd,
                         float* __restrict__ scale, float* __restrict__ shift)
{
    float s, q; bool owner; int c;
    sum_partials(partial, G, C, s, q, owner, c);
    if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
    if (!owner) return;
    const float k = partial[(size_t)G * 2 * C + c];          // the shift K = x[row 0]
    const float inv_m = 1.f / (float)M;
    const float ds = s * inv_m;
    const float mu = k + ds;
    const float var = fmaxf(fmaf(-ds, ds, q * inv_m), 0.f);
    const float is = rsqrtf(var + eps);
    mean[c] = mu;
    invstd[c] = is;
    const float sc = gamma[c] * is;
    scale[c] = sc;
    shift[c] = fmaf(-mu, sc, beta[c]);
    if (running_mean != nullptr) {
        const float unbiased = (M > 1) ? var * ((float)M / (float)(M - 1)) : var;
        running_mean[c] = fmaf(momentum, mu - running_mean[c], running_mean[c]);
        running_var[c] = fmaf(momentum, unbiased - running_var[c], running_var[c]);
    }
}

// eval mode: scale/shift from the running statistics
__global__ void bn_eval_coeff_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ running_mean,
                                     const float* __restrict__ running_var, float eps,
                                     float* __restrict__ scale, float* __restrict__ shift)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = gamma[c] * rsqrtf(running_var[c] + eps);
    scale[c] = sc;
    shift[c] = fmaf(-running_mean[c], sc, beta[c]);
}

// ---------------------------------------------------------------------------
// forward: apply   y = act(x*scale + shift [+ res])
// ---------------------------------------------------------------------------
template <typename T, bool RELU, bool ADD>
__global__ void __launch_bounds__(BN_THREADS, (sizeof(T) == 4 && ADD) ? 3 : 4)     // fp32 + residual: 80 registers
bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ scale,
                const float* __restrict__ shift, T* __restrict__ y, long long M, int C)
{
    const Map mp(C);
    if (!mp.active) return;
    const F8 sc = load_c8(scale + mp.cv * BN_VEC);
    const F8 sh = load_c8(shift + mp.cv * BN_VEC);
    const long long stride = (long long)gridDim.x * mp.rpi;
    const long long off = mp.cv * BN_VEC;
    constexpr int UN = Unroll<T>::N;
    for (long long r = (long long)blockIdx.x * mp.rpi + mp.rl; r < M; r += UN * stride) {
        typename Io<T>::raw_t dr[UN], ar[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long rr = r + u * stride;
            if (rr < M) {
                dr[u] = Io<T>::load_raw(x + rr * C + off);
                if (ADD) ar[u] = Io<T>::load_raw(res + rr * C + off);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long rr = r + u * stride;
            if (rr < M) {
                const F8 d = Io<T>::decode(dr[u]);
                F8 a;
                if (ADD) a = Io<T>::decode(ar[u]);
                F8 o;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float v = fmaf(d.v[i], sc.v[i], sh.v[i]);
                    if (ADD) v += a.v[i];
                    if (RELU) v = fmaxf(v, 0.f);
                    o.v[i] = v;
                }
                Io<T>::store(y + rr * C + off, o);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// backward: reduce   partial[.][0] = sum dz ; partial[.][1] = sum dz * xhat
// MODE 0: dz = dy            MODE 1: dz = dy * (x*scale+shift > 0)
// MODE 2: dz = dy * (y > 0), dz is also written (gradient of the residual branch)
// ---------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(BN_THREADS, 3)
bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ mean, const float* __restrict__ invstd,
                     float* __restrict__ partial, T* __restrict__ dz_out, long long M, int C)
{
    const Map mp(C);
    F8 s, q;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.v[i] = 0.f; q.v[i] = 0.f; }
    if (mp.active) {
        const long long off = mp.cv * BN_VEC;
        const F8 mu = load_c8(mean + off);
        const F8 is = load_c8(invstd + off);
        F8 sc, sh;
        if (MODE == 1) { sc = load_c8(scale + off); sh = load_c8(shift + off); }
        const long long stride = (long long)gridDim.x * mp.rpi;
        constexpr int UN = Unroll<T>::RED;
        for (long long r = (long long)blockIdx.x * mp.rpi + mp.rl; r < M; r += UN * stride) {
            typename Io<T>::raw_t gr[UN], xr[UN], yr[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const long long rr = r + u * stride;
                if (rr < M) {
                    gr[u] = Io<T>::load_raw(dy + rr * C + off);
                    xr[u] = Io<T>::load_raw(x + rr * C + off);
                    if (MODE == 2) yr[u] = Io<T>::load_raw(y + rr * C + off);
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const long long rr = r + u * stride;
                if (rr < M) {
                    const F8 g = Io<T>::decode(gr[u]);
                    const F8 d = Io<T>::decode(xr[u]);
                    F8 o;
                    if (MODE == 2) o = Io<T>::decode(yr[u]);
                    F8 dz;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float gz = g.v[i];
                        if (MODE == 1) gz = (fmaf(d.v[i], sc.v[i], sh.v[i]) > 0.f) ? gz : 0.f;
                        if (MODE == 2) gz = (o.v[i] > 0.f) ? gz : 0.f;
                        dz.v[i] = gz;
                        s.v[i] += gz;
                        q.v[i] = fmaf(gz, (d.v[i] - mu.v[i]) * is.v[i], q.v[i]);
                    }
                    if (MODE == 2) Io<T>::store(dz_out + rr * C + off, dz);
                }
            }
        }
    }
    reduce_store_partials(mp, s, q, partial, C);
}

// per-channel: grad_gamma, grad_beta and the dx coefficients
//   dx = scale*(dz - sdz/M - xhat*sdzx/M) = c1*dz - c2*x + c3
__global__ void __launch_bounds__(BN_FIN_THREADS)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int G, long long M, int C,
                       const float* __restrict__ scale, const float* __restrict__ mean,
                       const float* __restrict__ invstd,
                       float* __restrict__ grad_gamma, float* __restrict__ grad_beta,
                       float* __restrict__ c2, float* __restrict__ c3)
{
    float s, q; bool owner; int c;
    sum_partials(partial, G, C, s, q, owner, c);
    if (!owner) return;
    grad_beta[c] = s;
    grad_gamma[c] = q;
    const float inv_m = 1.f / (float)M;
    const float k1 = s * inv_m, k2 = q * inv_m;
    const float a = invstd[c] * k2;          // xhat*k2 = x*a - mean*a
    const float sc = scale[c];
    c2[c] = sc * a;
    c3[c] = sc * (mean[c] * a - k1);
}

// MODE 0/2: dz given directly (dy, or the dz written by the reduce kernel)
// MODE 1  : dz = dy * (x*scale+shift > 0) recomputed
template <typename T, int MODE>
__global__ void __launch_bounds__(BN_THREADS, 3)
bn_bwd_dx_kernel(const T* __restrict__ dz_in, const T* __restrict__ x, const float* __restrict__ scale,
                 const float* __restrict__ shift, const float* __restrict__ c2,
                 const float* __restrict__ c3, T* __restrict__ dx, long long M, int C)
{
    const Map mp(C);
    if (!mp.active) return;
    const long long off = mp.cv * BN_VEC;
    const F8 k1 = load_c8(scale + off);
    const F8 k2 = load_c8(c2 + off);
    const F8 k3 = load_c8(c3 + off);
    F8 sh;
    if (MODE == 1) sh = load_c8(shift + off);
    const long long stride = (long long)gridDim.x * mp.rpi;
    constexpr int UN = Unroll<T>::N;
    for (long long r = (long long)blockIdx.x * mp.rpi + mp.rl; r < M; r += UN * stride) {
        typename Io<T>::raw_t gr[UN], xr[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long rr = r + u * stride;
            if (rr < M) {
                gr[u] = Io<T>::load_raw(dz_in + rr * C + off);
                xr[u] = Io<T>::load_raw(x + rr * C + off);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long rr = r + u * stride;
            if (rr < M) {
                const F8 g = Io<T>::decode(gr[u]);
                const F8 d = Io<T>::decode(xr[u]);
                F8 o;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float gz = g.v[i];
                    if (MODE == 1) gz = (fmaf(d.v[i], k1.v[i], sh.v[i]) > 0.f) ? gz : 0.f;
                    o.v[i] = fmaf(k1.v[i], gz, fmaf(-k2.v[i], d.v[i], k3.v[i]));
                }
                Io<T>::store(dx + rr * C + off, o);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// launchers (C linkage; dtype: 0 = bf16, 1 = fp32)
// ---------------------------------------------------------------------------
static inline int bn_grid(long long M, int C, int per_sm)
{
    const int tpr = C / BN_VEC;
    const int rpi = BN_THREADS / tpr;
    long long iters = (M + rpi - 1) / rpi;
    long long g = (iters + BN_UNROLL - 1) / BN_UNROLL;
    const long long cap = 148LL * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" {

int bn_supported(long long M, int C) { return (C % BN_VEC == 0) && (C / BN_VEC <= BN_THREADS) && M > 0; }

int bn_partial_rows(long long M, int C) { return bn_grid(M, C, 4); }

// the backward reduce kernel holds 3 CTAs per SM (80 registers): a grid of 4 per SM
// would run as 1.33 waves with an idle tail (ncu: 4.4 TB/s vs 5.9 TB/s for its siblings)
int bn_partial_rows_bwd(long long M, int C) { return bn_grid(M, C, 3); }

cudaError_t bn_launch_stats(int dtype, const void* x, float* partial, long long M, int C, int G,
                            cudaStream_t st)
{
    if (dtype == 0) bn_stats_kernel<__nv_bfloat16><<<G, BN_THREADS, 0, st>>>((const __nv_bfloat16*)x, partial, M, C);
    else            bn_stats_kernel<float><<<G, BN_THREADS, 0, st>>>((const float*)x, partial, M, C);
    return cudaGetLastError();
}

cudaError_t bn_launch_stats_finalize(const float* partial, int G, long long M, int C, const float* gamma,
                                     const float* beta, float* rmean, float* rvar, long long* nbt,
                                     float momentum, float eps, float* mean, float* invstd,
                                     float* scale, float* shift, cudaStream_t st)
{
    bn_stats_finalize_kernel<<<(C + 31) / 32, BN_FIN_THREADS, 0, st>>>(partial, G, M, C, gamma, beta, rmean, rvar,
                                                              nbt, momentum, eps, mean, invstd, scale, shift);
    return cudaGetLastError();
}

cudaError_t bn_launch_eval_coeff(int C, const float* gamma, const float* beta, const float* rmean,
                                 const float* rvar, float eps, float* scale, float* shift, cudaStream_t st)
{
    bn_eval_coeff_kernel<<<(C + 127) / 128, 128, 0, st>>>(C, gamma, beta, rmean, rvar, eps, scale, shift);
    return cudaGetLastError();
}

#define BN_APPLY(T, R, A) bn_apply_kernel<T, R, A><<<G, BN_THREADS, 0, st>>>( \
    (const T*)x, (const T*)res, scale, shift, (T*)y, M, C)

cudaError_t bn_launch_apply(int dtype, int relu, int add, const void* x, const void* res,
                            const float* scale, const float* shift, void* y, long long M, int C,
                            cudaStream_t st)
{
    const int G = bn_grid(M, C, (dtype == 1 && add) ? 6 : 8);   // fp32 + residual holds 3 CTAs/SM: 2 waves
    if (dtype == 0) {
        if (relu && add) BN_APPLY(__nv_bfloat16, true, true);
        else if (relu)   BN_APPLY(__nv_bfloat16, true, false);
        else if (add)    BN_APPLY(__nv_bfloat16, false, true);
        else             BN_APPLY(__nv_bfloat16, false, false);
    } else {
        if (relu && add) BN_APPLY(float, true, true);
        else if (relu)   BN_APPLY(float, true, false);
        else if (add)    BN_APPLY(float, false, true);
        else             BN_APPLY(float, false, false);
    }
    return cudaGetLastError();
}

#define BN_RED(T, MODE) bn_bwd_reduce_kernel<T, MODE><<<G, BN_THREADS, 0, st>>>( \
    (const T*)dy, (const T*)x, (const T*)y, scale, shift, mean, invstd, partial, (T*)dz, M, C)

cudaError_t bn_launch_bwd_reduce(int dtype, int mode, const void* dy, const void* x, const void* y,
                                 const float* scale, const float* shift, const float* mean,
                                 const float* invstd, float* partial, void* dz, long long M, int C,
                                 int G, cudaStream_t st)
{
    if (dtype == 0) {
        if (mode == 0) BN_RED(__nv_bfloat16, 0); else if (mode == 1) BN_RED(__nv_bfloat16, 1); else BN_RED(__nv_bfloat16, 2);
    } else {
        if (mode == 0) BN_RED(float, 0); else if (mode == 1) BN_RED(float, 1); else BN_RED(float, 2);
    }
    return cudaGetLastError();
}

cudaError_t bn_launch_bwd_finalize(const float* partial, int G, long long M, int C, const float* scale,
                                   const float* mean, const float* invstd, float* ggamma, float* gbeta,
                                   float* c2, float* c3, cudaStream_t st)
{
    bn_bwd_finalize_kernel<<<(C + 31) / 32, BN_FIN_THREADS, 0, st>>>(partial, G, M, C, scale, mean, invstd,
                                                            ggamma, gbeta, c2, c3);
    return cudaGetLastError();
}

#define BN_DX(T, MODE) bn_bwd_dx_kernel<T, MODE><<<G, BN_THREADS, 0, st>>>( \
    (const T*)dz, (const T*)x, scale, shift, c2, c3, (T*)dx, M, C)

cudaError_t bn_launch_bwd_dx(int dtype, int mode, const void* dz, const void* x, const float* scale,
                             const float* shift, const float* c2, const float* c3, void* dx,
                             long long M, int C, cudaStream_t st)
{
    const int G = bn_grid(M, C, 6);      // 3 CTAs/SM resident -> exactly 2 waves
    if (dtype == 0) { if (mode == 1) BN_DX(__nv_bfloat16, 1); else BN_DX(__nv_bfloat16, 0); }
    else            { if (mode == 1) BN_DX(float, 1); else BN_DX(float, 0); }
    return cudaGetLastError();
}

}  // extern "C"
