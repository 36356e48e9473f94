// 1x1 convolution (NHWC) as a tcgen05 GEMM with the BatchNorm statistics fused
// into the epilogue -- sm_100a only.
//
//   Y[M, N] = X[M, K] . W[N, K]^T        M = batch*H*W, K = C_in, N = C_out, bf16 in / fp32 acc
//
// In a bottleneck ResNet two of the three convolutions of every block are 1x1 and
// every one of them feeds a training-mode BatchNorm, whose first pass re-reads the
// whole activation just to get per-channel mean / variance.  At batch 256 these
// GEMMs are HBM-bound (K is 64..2048 and M is up to 802,816), so the kernel is
// organised around the epilogue, not the MMA:
//
//   warp 0   TMA producer   X / W tiles -> 128B-swizzled smem ring (cp.async.bulk.tensor,
//                           mbarrier complete_tx); X is streamed evict_first, W evict_last
//   warp 1   MMA issuer     one thread: tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16,
//                           accumulators in TMEM, double-buffered (2 x BN columns) so the
//                           MMAs of tile i+1 run under the epilogue of tile i
//   warp 2-5 epilogue       tcgen05.ld (32 lanes x 32 columns) -> bf16 -> swizzled smem ->
//                           TMA store; then every thread owns a pair of channels of the
//                           staged tile and accumulates n / shifted sum / shifted sum of
//                           squares of the ROUNDED outputs (what BatchNorm will read)
//
// The grid is persistent: CTA (nb, j) owns output-channel block nb for its whole life
// and walks the row tiles j, j + ctas_per_n, ...; its statistics therefore stay in
// registers until the end, when each thread writes one partial row
// (n, K, sum(y-K), sum((y-K)^2)); c1_stats_finalize_kernel merges the <= 592 partial rows
// with the pairwise (Chan) update and emits mean / invstd / scale / shift and the
// running-statistics update, exactly what bn_stats_finalize_kernel does for the
// stand-alone statistics pass.
//
// Reference call site: torchvision Bottleneck conv1/conv3 + BatchNorm2d inside
// /root/reference/gossip_sgd.py (models.resnet50()).

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BM = 128;                 // rows per tile == TMEM lanes == UMMA_M
constexpr int BK = 64;                  // bf16 per k-block == one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;    // 16 KB
constexpr int SUB_BYTES = BM * 128;     // one 64-column output sub-tile, 16 KB
constexpr int kThreads = 192;           // producer warp, MMA warp, 4 epilogue warps
constexpr int kEpiThreads = 128;

constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;

template <int BN> struct Cfg {
    static constexpr int STAGES = BN == 256 ? 3 : (BN == 128 ? 5 : 8);
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int O_BYTES = (BN / 64) * SUB_BYTES;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;         // 128 / 256 / 512: powers of two
    static constexpr int SMEM = 1024 + STAGES * (A_BYTES + B_BYTES) + O_BYTES + 256;
};

// ---------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A protocol bug must not hang the GPU: after ~2 s of spinning the kernel traps.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    if (mbar_try(bar, parity)) return;
    unsigned long long t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        if (mbar_try(bar, parity)) return;
        if ((spins & 0xFFFu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1,
                                            uint64_t hint)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T, issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every MMA issued so far has retired (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

// K-major operand tile, rows of 64 bf16 (128 B) in the TMA 128B-swizzle layout, 8-row
// groups 1024 B apart.  (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

// 32 TMEM lanes (this warp's quadrant) x 32 consecutive fp32 columns -> 32 registers / thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(uint32_t lo_f32, uint32_t hi_f32)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi_f32)), "f"(__uint_as_float(lo_f32)));
    return r;
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// ---------------------------------------------------------------------------
// the GEMM
// ---------------------------------------------------------------------------
template <int BN, bool STATS>
__global__ void __launch_bounds__(kThreads, 1)
c1_gemm_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_y, int M, int N, int K, float* __restrict__ partial)
{
    using C = Cfg<BN>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = sA + STAGES * A_BYTES;
    uint8_t* sO = sB + STAGES * C::B_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sO + C::O_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;        // accumulator stage ready for the epilogue
    uint64_t* tempty = tfull + 2;            // accumulator stage drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;       // warp-uniform
    const int lane = threadIdx.x & 31;

    const int n_blocks = N / BN;
    const int ctas_per_n = gridDim.x / n_blocks;
    const int nb = blockIdx.x % n_blocks;
    const int j = blockIdx.x / n_blocks;
    const int m_tiles = (M + BM - 1) / BM;
    const int k_blocks = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_y);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int mt = j; mt < m_tiles; mt += ctas_per_n) {
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], A_BYTES + C::B_BYTES);
                    tma_load_2d(&tm_x, &full[stage], sA + stage * A_BYTES, kb * BK, mt * BM, L2_EVICT_FIRST);
                    tma_load_2d(&tm_w, &full[stage], sB + stage * C::B_BYTES, kb * BK, nb * BN, L2_EVICT_LAST);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // cute::UMMA::InstrDescriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
            // A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                       ((uint32_t)(BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int mt = j; mt < m_tiles; mt += ctas_per_n) {
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = umma_desc_sw128(smem_u32(sA + stage * A_BYTES));
                    const uint64_t b_desc = umma_desc_sw128(smem_u32(sB + stage * C::B_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)      // +32 B per UMMA_K inside the swizzle row
                        umma_bf16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                  (uint32_t)((kb | k) != 0));
                    umma_commit(&empty[stage]);                 // smem slot free once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[as]);
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;                       // TMEM lane quadrant this warp may read
        const int et = threadIdx.x - 64;              // 0..127
        const int row = q * 32 + lane;                // tile row owned for the TMEM -> smem copy
        constexpr int P = BN / 2;                     // bf16 pairs per tile row
        constexpr int G = kEpiThreads / P;            // row groups for the statistics (1, 2, 4)
        constexpr int RG = BM / G;                    // rows per group
        const int w = et % P;
        const int rg = et / P;
        const uint32_t stat_base = smem_u32(sO) + (uint32_t)((w >> 5) * SUB_BYTES + (w & 3) * 4);
        const int jchunk = (w & 31) >> 2;
        float cnt = 0.f, k0 = 0.f, k1 = 0.f, s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        bool have_k = false;
        int as = 0;
        uint32_t aphase = 0;
        for (int mt = j; mt < m_tiles; mt += ctas_per_n) {
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging free again
            epi_barrier();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(t_row + (uint32_t)(c * 32), v);
                tmem_ld_wait();
                const uint32_t dst = smem_u32(sO) + (uint32_t)((c >> 1) * SUB_BYTES + row * 128);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t chunk = (uint32_t)(((c & 1) * 4 + i) ^ (row & 7));
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + chunk * 16),
                                 "r"(pack_bf16(v[8 * i + 0], v[8 * i + 1])), "r"(pack_bf16(v[8 * i + 2], v[8 * i + 3])),
                                 "r"(pack_bf16(v[8 * i + 4], v[8 * i + 5])), "r"(pack_bf16(v[8 * i + 6], v[8 * i + 7]))
                                 : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);                    // MMA may overwrite this stage
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> TMA reads
            epi_barrier();
            if (et == 0) {
#pragma unroll
                for (int sub = 0; sub < BN / 64; ++sub)
                    tma_store_2d(&tm_y, sO + sub * SUB_BYTES, nb * BN + sub * 64, mt * BM);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (STATS) {
                const int valid = min(BM, M - mt * BM);
                const int r_end = min(rg * RG + RG, valid);
                int r = rg * RG;
                if (r < r_end && !have_k) {
                    uint32_t u;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(stat_base + r * 128 + ((jchunk ^ (r & 7)) << 4)));
                    k0 = __uint_as_float(u << 16);
                    k1 = __uint_as_float(u & 0xFFFF0000u);
                    have_k = true;
                }
                for (; r + 8 <= r_end; r += 8) {
                    uint32_t u[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)      // r is a multiple of 8 here: (r + i) & 7 == i
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u[i]) : "r"(stat_base + (r + i) * 128 + ((jchunk ^ i) << 4)));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float d0 = __uint_as_float(u[i] << 16) - k0;
                        const float d1 = __uint_as_float(u[i] & 0xFFFF0000u) - k1;
                        s0 += d0; q0 = fmaf(d0, d0, q0);
                        s1 += d1; q1 = fmaf(d1, d1, q1);
                    }
                    cnt += 8.f;
                }
                for (; r < r_end; ++r) {
                    uint32_t u;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(stat_base + r * 128 + ((jchunk ^ (r & 7)) << 4)));
                    const float d0 = __uint_as_float(u << 16) - k0;
                    const float d1 = __uint_as_float(u & 0xFFFF0000u) - k1;
                    s0 += d0; q0 = fmaf(d0, d0, q0);
                    s1 += d1; q1 = fmaf(d1, d1, q1);
                    cnt += 1.f;
                }
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
        if (STATS) {
            // partial[row][field][N], fields = n, K, sum(y-K), sum((y-K)^2); row = j*G + rg
            float* p = partial + ((size_t)(j * G + rg) * 4) * N + nb * BN + 2 * w;
            *reinterpret_cast<float2*>(p) = make_float2(cnt, cnt);
            *reinterpret_cast<float2*>(p + N) = make_float2(k0, k1);
            *reinterpret_cast<float2*>(p + 2 * (size_t)N) = make_float2(s0, s1);
            *reinterpret_cast<float2*>(p + 3 * (size_t)N) = make_float2(q0, q1);
        }
        if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------
// statistics finalize: merge R partial rows per channel (Chan et al. pairwise update),
// then the same outputs as bn_stats_finalize_kernel.  1024 threads = 32 row lanes x 32 channels.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
c1_stats_finalize_kernel(const float* __restrict__ partial, int R, int C, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* running_mean, float* running_var, long long* nbt,
                         float momentum, float eps, float* __restrict__ mean, float* __restrict__ invstd,
                         float* __restrict__ scale, float* __restrict__ shift)
{
    __shared__ float sm_a[32][33];
    __shared__ float sm_b[32][33];
    __shared__ float sm_mu[32];
    const int cl = threadIdx.x & 31;
    const int lane = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + cl;
    const bool live = ch < C;
    if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += 1;

    // pass 1: total count and mean
    float n_sum = 0.f, m_sum = 0.f;
    if (live)
        for (int r = lane; r < R; r += 32) {
            const float* p = partial + (size_t)r * 4 * C + ch;
            const float n = p[0];
            if (n > 0.f) { n_sum += n; m_sum += fmaf(n, p[C], p[2 * (size_t)C]); }   // n*K + S1
        }
    sm_a[lane][cl] = n_sum;
    sm_b[lane][cl] = m_sum;
    __syncthreads();
    if (lane == 0) {
        float n = 0.f, m = 0.f;
        for (int l = 0; l < 32; ++l) { n += sm_a[l][cl]; m += sm_b[l][cl]; }
        sm_mu[cl] = (n > 0.f) ? m / n : 0.f;
        sm_a[0][cl] = n;
    }
    __syncthreads();
    const float mu = sm_mu[cl];
    const float n_tot = sm_a[0][cl];
    __syncthreads();

    // pass 2: M2 = sum_p [ S2_p - S1_p^2/n_p + n_p (mean_p - mu)^2 ]
    float m2 = 0.f;
    if (live)
        for (int r = lane; r < R; r += 32) {
            const float* p = partial + (size_t)r * 4 * C + ch;
            const float n = p[0];
            if (n > 0.f) {
                const float s1 = p[2 * (size_t)C];
                const float mp = p[C] + s1 / n;
                const float d = mp - mu;
                m2 += fmaxf(fmaf(-s1, s1 / n, p[3 * (size_t)C]), 0.f) + n * d * d;
            }
        }
    sm_b[lane][cl] = m2;
    __syncthreads();
    if (lane != 0 || !live) return;
    float tot = 0.f;
    for (int l = 0; l < 32; ++l) tot += sm_b[l][cl];
    const float var = (n_tot > 0.f) ? tot / n_tot : 0.f;
    const float is = rsqrtf(var + eps);
    mean[ch] = mu;
    invstd[ch] = is;
    const float sc = gamma[ch] * is;
    scale[ch] = sc;
    shift[ch] = fmaf(-mu, sc, beta[ch]);
    if (running_mean != nullptr) {
        const float unbiased = (n_tot > 1.f) ? var * (n_tot / (n_tot - 1.f)) : var;
        running_mean[ch] = fmaf(momentum, mu - running_mean[ch], running_mean[ch]);
        running_var[ch] = fmaf(momentum, unbiased - running_var[ch], running_var[ch]);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major [rows, cols] bf16 matrix, box = [box_rows, 64 cols], 128B swizzle
bool make_map(CUtensorMap* tm, const void* base, long long rows, int cols, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int pick_bn(int N) { return (N % 256 == 0) ? 256 : ((N % 128 == 0) ? 128 : 64); }

void grid_shape(long long M, int N, int num_sms, int& bn, int& ctas_per_n, int& grid)
{
    bn = pick_bn(N);
    const int n_blocks = N / bn;
    const long long m_tiles = (M + BM - 1) / BM;
    long long per = num_sms / n_blocks;
    if (per < 1) per = 1;
    if (per > m_tiles) per = m_tiles;
    ctas_per_n = (int)per;
    grid = n_blocks * ctas_per_n;
}

template <int BN>
cudaError_t launch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, int M, int N, int K,
                   float* partial, int grid, cudaStream_t st)
{
    auto kern = partial ? c1_gemm_kernel<BN, true> : c1_gemm_kernel<BN, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM);
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads, Cfg<BN>::SMEM, st>>>(tx, tw, ty, M, N, K, partial);
    return cudaGetLastError();
}

}  // namespace

extern "C" {

// shapes the kernel covers: bf16, K % 8 == 0 (16-byte row pitch), N % 64 == 0, M < 2^31
int c1_supported(long long M, int N, int K)
{
    return M >= 1 && M < (1ll << 31) && N >= 64 && N % 64 == 0 && K >= 8 && K % 8 == 0;
}

// number of partial-statistics rows c1_launch_gemm writes for this shape
int c1_partial_rows(long long M, int N, int num_sms)
{
    int bn, per, grid;
    grid_shape(M, N, num_sms, bn, per, grid);
    return per * (kEpiThreads / (bn / 2));
}

// y[M,N] = x[M,K] . w[N,K]^T (bf16); partial (nullable) = [c1_partial_rows][4][N] fp32
cudaError_t c1_launch_gemm(const void* x, const void* w, void* y, long long M, int N, int K, float* partial,
                           int num_sms, cudaStream_t st)
{
    if (!c1_supported(M, N, K)) return cudaErrorInvalidValue;
    int bn, per, grid;
    grid_shape(M, N, num_sms, bn, per, grid);
    CUtensorMap tx, tw, ty;
    if (!make_map(&tx, x, M, K, BM) || !make_map(&tw, w, N, K, bn) || !make_map(&ty, y, M, N, BM))
        return cudaErrorInvalidValue;
    switch (bn) {
    case 256: return launch<256>(tx, tw, ty, (int)M, N, K, partial, grid, st);
    case 128: return launch<128>(tx, tw, ty, (int)M, N, K, partial, grid, st);
    default: return launch<64>(tx, tw, ty, (int)M, N, K, partial, grid, st);
    }
}

cudaError_t c1_launch_stats_finalize(const float* partial, int R, int C, const float* gamma, const float* beta,
                                     float* rmean, float* rvar, long long* nbt, float momentum, float eps,
                                     float* mean, float* invstd, float* scale, float* shift, cudaStream_t st)
{
    c1_stats_finalize_kernel<<<(C + 31) / 32, 1024, 0, st>>>(partial, R, C, gamma, beta, rmean, rvar, nbt, momentum,
                                                            eps, mean, invstd, scale, shift);
    return cudaGetLastError();
}

}  // extern "C"
