// 1x1 convolution (NHWC) as a tcgen05 GEMM with the BatchNorm statistics (forward) or the
// skip-branch gradient accumulation (backward) fused into the epilogue -- sm_100a only.
//
//   Y[M, N] = X[M, K] . W[N, K]^T        M = batch*H*W, K = C_in, N = C_out, bf16 in / fp32 acc
//
// In a bottleneck ResNet two of the three convolutions of every block are 1x1 and
// every one of them feeds a training-mode BatchNorm, whose first pass re-reads the
// whole activation just to get per-channel mean / variance.  At batch 256 these
// GEMMs are HBM-bound (K is 64..2048 and M is up to 802,816), so the kernel is
// organised around the epilogue, not the MMA:
//
//   warp 0   TMA producer   X tiles -> 128B-swizzled smem ring (cp.async.bulk.tensor, mbarrier
//                           complete_tx).  The CTA's W block [BN, K] is loaded ONCE and stays
//                           resident whenever it fits next to a >= 3-deep X ring (K*BN <= 64K
//                           elements); otherwise W tiles travel through the ring with X.
//   warp 1   MMA issuer     one thread: tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16,
//                           accumulators in TMEM, double-buffered (2 x BN columns) so the
//                           MMAs of tile i+1 run under the epilogue of tile i
//   warp 2-9 epilogue       two sets of four warps, set h drains accumulator stage h (every other
//                           tile).  Each warp owns 32 tile rows (its TMEM lane quadrant) end to end:
//                           tcgen05.ld (32 lanes x 64 columns) -> bf16 -> its own 4 KB swizzled
//                           slab -> its own TMA store (box 64 x 32), then the statistics of
//                           exactly those rows are read back from the slab (lane = channel pair):
//                           n / shifted sum / shifted sum of squares of the ROUNDED outputs
//                           (what BatchNorm will read).  No CTA-wide barrier in the loop; slabs
//                           recycle through the issuing lane's bulk-group counter.
//
// The grid is persistent: CTA (nb, j) owns output-channel block nb for its whole life
// and walks the row tiles j, j + ctas_per_n, ...; its statistics therefore stay in
// registers until the end: each epilogue warp keeps (n, K, sum(y-K), sum((y-K)^2)) with its
// own shift K (its first output row), the eight warps are merged in shared memory into one
// (n, mean, M2) row per CTA, and c1_stats_finalize_kernel merges the <= 148 rows with the
// pairwise (Chan) update and emits mean / invstd / scale / shift and the running-statistics
// update, exactly what bn_stats_finalize_kernel does for the stand-alone statistics pass.
//
// Three epilogue modes share the producer / MMA pipeline:
//   MODE 0  Y = X.W^T
//   MODE 1  + the BatchNorm statistics above (forward of conv + BN: no separate statistics pass)
//   MODE 2  Y = X.W^T + R, R's 4 KB slab prefetched by TMA one sub-tile ahead (own mbarrier per
//           slab), summed in fp32 before the single bf16 rounding.  Used as the dgrad of the first
//           convolution of a residual block with R = the skip-branch gradient, which replaces
//           autograd's read-read-write accumulation pass over the block's widest tensor.
//
// Measured on B200 (benchmarks/conv1x1_bench.py, profiles/conv1x1_bench_*.log).
//
// Reference call site: torchvision Bottleneck conv1/conv3 + BatchNorm2d inside
// /root/reference/gossip_sgd.py (models.resnet50()).

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128;                 // rows per tile == TMEM lanes == UMMA_M
constexpr int ROW_BYTES = 128;          // one k-block row == one 128-byte swizzle row: 64 bf16 or 32 fp32
constexpr int A_BYTES = BM * ROW_BYTES; // 16 KB
constexpr int SLAB_BYTES = 32 * 128;    // one epilogue warp's store slab: 32 rows x 128 B (64 bf16 / 32 fp32), 4 KB
// Element-type dependent tile geometry.  F32 = fp32 activations / weights multiplied on the tensor
// cores as TF32 (tcgen05.mma.kind::tf32 reads fp32 words from shared memory and uses their top 19
// bits: the same arithmetic as the reference's cuDNN TF32 convolutions), fp32 accumulate, fp32 out.
template <bool F32> struct Elt {
    static constexpr int BYTES = F32 ? 4 : 2;
    static constexpr int BK = ROW_BYTES / BYTES;      // elements per k-block: 32 / 64
    static constexpr int UMMA_K = 32 / BYTES;         // 8 (tf32) / 16 (bf16): 32 bytes of K per instruction
    static constexpr int COLS = ROW_BYTES / BYTES;    // output columns per store slab: 32 / 64
};
constexpr int kThreads = 320;           // producer warp, MMA warp, 2 x 4 epilogue warps
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100
constexpr int SMEM_FIXED = 1024 + 512;  // alignment slack + barriers / TMEM slot

// cute::TMA::CacheHintSm90 encodings
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;

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
__device__ __forceinline__ void tma_load_2d_s(const CUtensorMap* tm, uint32_t bar_smem, uint32_t dst_smem, int c0,
                                              int c1, uint64_t hint)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst_smem),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_smem), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src_smem, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(src_smem), "r"(c0), "r"(c1)
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
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
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

// ---------------------------------------------------------------------------
// the GEMM
// ---------------------------------------------------------------------------
struct C1Plan {            // host-computed shared-memory plan (bytes are multiples of 1024)
    int stages;            // depth of the X (or X+W) ring, 2..MAX_STAGES
    int resident;          // 1: this CTA's whole W block [BN, K] is loaded once and stays in smem
    int nbuf;              // store slabs per epilogue warp (2 or 1)
    int x_hint_first;      // 1: X tiles are read by one CTA only -> L2 evict_first
};

// MODE 0: Y = X.W^T          MODE 1: + BatchNorm statistics of Y          MODE 2: Y = X.W^T + R
template <int BN, int MODE, bool F32>
__global__ void __launch_bounds__(kThreads, 1)
c1_gemm_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_r, int M, int N,
               int K, float* __restrict__ partial, const C1Plan plan)
{
    constexpr bool STATS = MODE == 1;
    constexpr bool RES = MODE == 2;
    constexpr int BK = Elt<F32>::BK;
    constexpr int UMMA_K = Elt<F32>::UMMA_K;
    constexpr int COLS = Elt<F32>::COLS;
    constexpr int B_BYTES = BN * ROW_BYTES;
    constexpr int NS = BN / COLS;             // store-slab wide sub-tiles per tile (64 bf16 / 32 fp32 columns)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5;        // warp-uniform
    const int lane = threadIdx.x & 31;
    const int n_blocks = N / BN;
    const int ctas_per_n = gridDim.x / n_blocks;
    const int nb = blockIdx.x % n_blocks;
    const int j = blockIdx.x / n_blocks;
    const int m_tiles = (M + BM - 1) / BM;
    const int k_blocks = (K + BK - 1) / BK;
    const int stages = plan.stages;
    const bool resident = plan.resident != 0;
    const int nbuf = plan.nbuf;

    // [ resident W : k_blocks x B_BYTES ][ ring : stages x (X tile [+ W tile]) ][ store slabs ][ barriers ]
    uint8_t* sW = smem;
    uint8_t* ring = sW + (resident ? k_blocks * B_BYTES : 0);
    const int stage_bytes = A_BYTES + (resident ? 0 : B_BYTES);
    uint8_t* sO = ring + stages * stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sO + 8 * nbuf * SLAB_BYTES);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* tfull = empty + MAX_STAGES;    // accumulator stage ready for the epilogue
    uint64_t* tempty = tfull + 2;            // accumulator stage drained
    uint64_t* wfull = tempty + 2;            // resident W landed
    uint64_t* rbar0 = wfull + 1;             // MODE 2: residual slab landed, [8 warps][2 slots]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar0 + 16);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_y);
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        mbar_init(wfull, 1);
        if (RES) { tma_prefetch_desc(&tm_r); for (int i = 0; i < 16; ++i) mbar_init(&rbar0[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)(2 * BN))
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
            if (resident) {
                mbar_expect_tx(wfull, (uint32_t)(k_blocks * B_BYTES));
                for (int kb = 0; kb < k_blocks; ++kb)
                    tma_load_2d(&tm_w, wfull, sW + kb * B_BYTES, kb * BK, nb * BN, L2_EVICT_LAST);
            }
            const uint64_t x_hint = plan.x_hint_first ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
            int stage = 0;
            uint32_t phase = 0;
            for (int mt = j; mt < m_tiles; mt += ctas_per_n) {
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
                    uint8_t* dst = ring + stage * stage_bytes;
                    tma_load_2d(&tm_x, &full[stage], dst, kb * BK, mt * BM, x_hint);
                    if (!resident) tma_load_2d(&tm_w, &full[stage], dst + A_BYTES, kb * BK, nb * BN, L2_EVICT_LAST);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // cute::UMMA::InstrDescriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
            // A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
            // (tf32: A/B format code 2)
            constexpr uint32_t fmt = F32 ? 2u : 1u;
            constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) |
                                       ((uint32_t)(BM >> 4) << 24);
            if (resident) { mbar_wait(wfull, 0); tc_fence_after(); }
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
                    const uint8_t* a_ptr = ring + stage * stage_bytes;
                    const uint8_t* b_ptr = resident ? sW + kb * B_BYTES : a_ptr + A_BYTES;
                    const uint64_t a_desc = umma_desc_sw128(smem_u32(a_ptr));
                    const uint64_t b_desc = umma_desc_sw128(smem_u32(b_ptr));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {    // +32 B per UMMA_K inside the swizzle row
                        if (F32)
                            umma_tf32(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                      (uint32_t)((kb | k) != 0));
                        else
                            umma_bf16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                      (uint32_t)((kb | k) != 0));
                    }
                    umma_commit(&empty[stage]);                 // smem slot free once these MMAs retire
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[as]);
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else {
        // ===================== epilogue =====================
        // Two sets of four warps; set h drains accumulator stage h, i.e. every other tile, so each
        // SM sub-partition has two epilogue warps to interleave (a lone warp per scheduler exposes
        // every TMEM / shared-memory latency) and one stage drains while the other fills.
        // Within a set, warp q owns tile rows [32q, 32q+32) (its TMEM lane quadrant) end to end:
        // TMEM -> bf16 -> its own 4 KB swizzled slab -> its own TMA store (box 64 x 32) ->
        // statistics of exactly those rows read back from the slab.  No CTA-wide barrier; slabs
        // are recycled through the issuing lane's bulk-group counter.
        const int set = (warp - 2) >> 2;
        const int q = warp & 3;
        const uint32_t slab0 = smem_u32(sO) + (uint32_t)((set * 4 + q) * nbuf * SLAB_BYTES);
        const uint32_t wr_off = (uint32_t)(lane * 128);                    // this lane's row in the slab
        const uint32_t rd_off = (uint32_t)((lane & 3) * 4);                // this lane's bf16 pair in a row
        const int rd_chunk = lane >> 2;
        int slot = 0;
        float cnt = 0.f;
        float k0[NS], k1[NS], s0[NS], s1[NS], q0[NS], q1[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) { k0[s] = k1[s] = s0[s] = s1[s] = q0[s] = q1[s] = 0.f; }
        bool have_k = false;
        const int as = set;
        uint32_t aphase = 0;
        // MODE 2: the residual slab of sub-tile i+1 is fetched (TMA, own mbarrier per slab) while
        // sub-tile i is converted; the sum is formed in fp32 in place and stored from the same slab.
        const uint32_t rbar = smem_u32(rbar0 + (set * 4 + q) * 2);
        uint32_t rphase = 0;                                     // bit b: parity of slab b's barrier
        const int tile_step = 2 * ctas_per_n;
        if (RES && lane == 0) {
            const int mt0 = j + set * ctas_per_n;
            if (mt0 < m_tiles && M - mt0 * BM - q * 32 > 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rbar), "r"(SLAB_BYTES)
                             : "memory");
                tma_load_2d_s(&tm_r, rbar, slab0, nb * BN, mt0 * BM + q * 32, L2_EVICT_FIRST);
            }
        }
        for (int mt = j + set * ctas_per_n; mt < m_tiles; mt += tile_step) {
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const int nrows = min(32, M - mt * BM - q * 32);       // rows of this slab that exist (<= 0: none)
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const uint32_t slab = slab0 + (uint32_t)(slot * SLAB_BYTES);
                if (!RES) {
                    if (lane == 0) {                               // the store that last used this slab has read it
                        if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
                }
                uint32_t v[64];
                tmem_ld32(t_row + (uint32_t)(s * COLS), v);
                if (!F32) tmem_ld32(t_row + (uint32_t)(s * COLS + 32), v + 32);
                tmem_ld_wait();
                if (s == NS - 1) {                                 // accumulator fully read: hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                }
                if (RES && nrows > 0) {
                    const uint32_t bar = rbar + (uint32_t)(slot * 8);
                    const uint32_t par = (rphase >> slot) & 1u;
                    while (true) {
                        uint32_t ok;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                            "selp.u32 %0, 1, 0, p;\n\t}"
                            : "=r"(ok) : "r"(bar), "r"(par) : "memory");
                        if (ok) break;
                    }
                    rphase ^= 1u << slot;
                }
                if (!RES || nrows > 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t addr = slab + wr_off + (uint32_t)((i ^ (lane & 7)) * 16);
                        if (RES) {
                            uint32_t r0, r1, r2, r3;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
                            const uint32_t rr[4] = {r0, r1, r2, r3};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (F32) {
                                    v[4 * i + e] = __float_as_uint(__uint_as_float(v[4 * i + e]) +
                                                                   __uint_as_float(rr[e]));
                                } else {
                                    v[8 * i + 2 * e] = __float_as_uint(__uint_as_float(v[8 * i + 2 * e]) +
                                                                       __uint_as_float(rr[e] << 16));
                                    v[8 * i + 2 * e + 1] = __float_as_uint(__uint_as_float(v[8 * i + 2 * e + 1]) +
                                                                           __uint_as_float(rr[e] & 0xFFFF0000u));
                                }
                            }
                        }
                        if (F32)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * i + 0]),
                                         "r"(v[4 * i + 1]), "r"(v[4 * i + 2]), "r"(v[4 * i + 3])
                                         : "memory");
                        else
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                                         "r"(pack_bf16(v[8 * i + 0], v[8 * i + 1])), "r"(pack_bf16(v[8 * i + 2], v[8 * i + 3])),
                                         "r"(pack_bf16(v[8 * i + 4], v[8 * i + 5])), "r"(pack_bf16(v[8 * i + 6], v[8 * i + 7]))
                                         : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> TMA reads
                __syncwarp();
                if (lane == 0) {
                    if (nrows > 0) tma_store_2d(&tm_y, slab, nb * BN + s * COLS, mt * BM + q * 32);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (RES) {                                     // fetch the next residual slab
                        const int nmt = (s + 1 < NS) ? mt : mt + tile_step;
                        const int ns = (s + 1 < NS) ? s + 1 : 0;
                        if (nmt < m_tiles && M - nmt * BM - q * 32 > 0) {
                            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // other slab is free
                            const uint32_t nbar = rbar + (uint32_t)((slot ^ 1) * 8);
                            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nbar),
                                         "r"(SLAB_BYTES)
                                         : "memory");
                            tma_load_2d_s(&tm_r, nbar, slab0 + (uint32_t)((slot ^ 1) * SLAB_BYTES), nb * BN + ns * COLS,
                                          nmt * BM + q * 32, L2_EVICT_FIRST);
                        }
                    }
                }
                if (STATS && nrows > 0) {
                    const uint32_t rd = slab + rd_off;
                    if (!have_k) {
                        uint32_t u;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(rd + (uint32_t)(rd_chunk << 4)));
                        k0[s] = F32 ? __uint_as_float(u) : __uint_as_float(u << 16);
                        k1[s] = __uint_as_float(u & 0xFFFF0000u);
                    }
                    if (nrows == 32) {
#pragma unroll
                        for (int r8 = 0; r8 < 32; r8 += 8) {
                            uint32_t u[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u[i])
                                             : "r"(rd + (uint32_t)((r8 + i) * 128 + ((rd_chunk ^ i) << 4))));
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float d0 = (F32 ? __uint_as_float(u[i]) : __uint_as_float(u[i] << 16)) - k0[s];
                                s0[s] += d0; q0[s] = fmaf(d0, d0, q0[s]);
                                if (!F32) {          // fp32: one channel per lane; bf16: a channel pair
                                    const float d1 = __uint_as_float(u[i] & 0xFFFF0000u) - k1[s];
                                    s1[s] += d1; q1[s] = fmaf(d1, d1, q1[s]);
                                }
                            }
                        }
                    } else {
                        for (int r = 0; r < nrows; ++r) {
                            uint32_t u;
                            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u)
                                         : "r"(rd + (uint32_t)(r * 128 + ((rd_chunk ^ (r & 7)) << 4))));
                            const float d0 = (F32 ? __uint_as_float(u) : __uint_as_float(u << 16)) - k0[s];
                            s0[s] += d0; q0[s] = fmaf(d0, d0, q0[s]);
                            if (!F32) {
                                const float d1 = __uint_as_float(u & 0xFFFF0000u) - k1[s];
                                s1[s] += d1; q1[s] = fmaf(d1, d1, q1[s]);
                            }
                        }
                    }
                }
                slot = (slot + 1) & (nbuf - 1);      // nbuf is 1 or 2
            }
            if (nrows > 0) { cnt += (float)nrows; have_k = true; }
            aphase ^= 1;
        }
        if (STATS) {
            // Merge the eight warps' partials (each with its own shift) into ONE row per CTA,
            // partial[j][field][N] with fields = n, mean, M2 (Chan et al.), so the finalize kernel
            // reads <= 148 rows.  The store slabs double as scratch once their stores have drained.
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            float4* sc = reinterpret_cast<float4*>(sO);
            const int e = set * 4 + q;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (F32) {
                    sc[e * BN + s * COLS + lane] = make_float4(cnt, k0[s], s0[s], q0[s]);
                } else {
                    sc[e * BN + s * COLS + 2 * lane] = make_float4(cnt, k0[s], s0[s], q0[s]);
                    sc[e * BN + s * COLS + 2 * lane + 1] = make_float4(cnt, k1[s], s1[s], q1[s]);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int t = (int)threadIdx.x - 64;
            if (t < BN) {
                float n_tot = 0.f, m_sum = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v = sc[i * BN + t];
                    n_tot += v.x;
                    m_sum += fmaf(v.x, v.y, v.z);                  // n*K + sum(y-K)
                }
                const float mu = n_tot > 0.f ? m_sum / n_tot : 0.f;
                float m2 = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v = sc[i * BN + t];
                    if (v.x > 0.f) {
                        const float ds = v.z / v.x;                // mean_i - K_i
                        const float d = v.y + ds - mu;
                        m2 += fmaxf(fmaf(-v.z, ds, v.w), 0.f) + v.x * d * d;
                    }
                }
                float* p = partial + (size_t)j * 3 * N + nb * BN + t;
                p[0] = n_tot;
                p[N] = mu;
                p[2 * (size_t)N] = m2;
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)(2 * BN))
                     : "memory");
    }
}

// ---------------------------------------------------------------------------
// statistics finalize: merge the R per-CTA rows (n, mean, M2) per channel (Chan et al.),
// then the same outputs as bn_stats_finalize_kernel.  1024 threads = 32 row lanes x 32 channels.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
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

    // every row of this lane is loaded up front (R <= 148 -> at most 5 rows per lane per batch), so
    // the kernel pays one L2 round trip instead of one per row
    constexpr int ILP = 5;
    float n_sum = 0.f, m_sum = 0.f;
    float nv[ILP], mv[ILP], qv[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
        const int r = lane + 32 * u;
        const bool ok = live && r < R;
        const int at = r * 3 * C + ch;                 // R * 3 * C < 2^20: 32-bit indexing
        nv[u] = ok ? partial[at] : 0.f;
        mv[u] = ok ? partial[at + C] : 0.f;
        qv[u] = ok ? partial[at + 2 * C] : 0.f;
        n_sum += nv[u];
        m_sum = fmaf(nv[u], mv[u], m_sum);
    }
    if (live)
        for (int r = lane + 32 * ILP; r < R; r += 32) {            // (only on parts with > 160 SMs)
            const float* p = partial + (size_t)r * 3 * C + ch;
            n_sum += p[0];
            m_sum = fmaf(p[0], p[C], m_sum);
        }
    sm_a[lane][cl] = n_sum;
    sm_b[lane][cl] = m_sum;
    __syncthreads();
    if (lane == 0) {
        float n = 0.f, m = 0.f;
#pragma unroll 4
        for (int l = 0; l < 32; ++l) { n += sm_a[l][cl]; m += sm_b[l][cl]; }
        sm_mu[cl] = (n > 0.f) ? m / n : 0.f;
        sm_a[0][cl] = n;
    }
    __syncthreads();
    const float mu = sm_mu[cl];
    const float n_tot = sm_a[0][cl];
    __syncthreads();

    // M2 = sum_p [ M2_p + n_p (mean_p - mu)^2 ]
    float m2 = 0.f;
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
        const float d = mv[u] - mu;
        m2 += qv[u] + nv[u] * d * d;
    }
    if (live)
        for (int r = lane + 32 * ILP; r < R; r += 32) {
            const float* p = partial + (size_t)r * 3 * C + ch;
            const float d = p[C] - mu;
            m2 += p[2 * (size_t)C] + p[0] * d * d;
        }
    sm_b[lane][cl] = m2;
    __syncthreads();
    if (lane != 0 || !live) return;
    float tot = 0.f;
#pragma unroll 4
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

// row-major [rows, cols] bf16 / fp32 matrix, box = [box_rows, 128 bytes of columns], 128B swizzle.  wide_promotion:
// let L2 fetch 256 B per request -- right for operand tiles whose rows are consumed whole, wrong
// for the 128 B-per-row output / residual slabs (ncu: +21 % DRAM reads on the residual stream).
bool make_map(CUtensorMap* tm, const void* base, long long rows, int cols, int box_rows, bool wide_promotion,
              bool f32)
{
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * (f32 ? 4 : 2)};
    const cuuint32_t box[2] = {f32 ? 32u : 64u, (cuuint32_t)box_rows};     // 128 bytes wide
    const cuuint32_t estr[2] = {1u, 1u};
    return fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
              const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              wide_promotion ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Tile width: the widest BLOCK_N dividing N unless a narrower one shortens the critical path
// (waves of row tiles per CTA x BLOCK_N) by more than 10 % -- that only happens when M is small.
bool c1_f32_prefer_resident()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SGP_B200_C1_F32_RESIDENT");
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

int bn_override()        // SGP_B200_C1_BN=64|128|256 forces a tile width (benchmarking)
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SGP_B200_C1_BN");
        v = e ? atoi(e) : 0;
    }
    return v;
}

// can the CTA's whole W block stay resident next to a >= 3-deep X ring and one slab per epilogue warp?
bool fits_resident(int bn, int K, bool f32)
{
    const int bk = f32 ? Elt<true>::BK : Elt<false>::BK;
    const int k_blocks = (K + bk - 1) / bk;
    return SMEM_LIMIT - SMEM_FIXED - 8 * SLAB_BYTES - k_blocks * bn * ROW_BYTES >= 3 * A_BYTES;
}

int pick_bn(long long M, int N, int K, int num_sms, bool f32)
{
    const long long m_tiles = (M + BM - 1) / BM;
    int best = 0;
    long long best_cost = 0;
    const int force = bn_override();
    if (force >= 64 && force <= 256 && N % force == 0) return force;
    // fp32 operands are twice as wide: prefer the widest tile whose W block stays resident (a
    // non-resident 256-wide tile re-streams 3x the bytes of X per row tile through L2)
    int bn_max = 256;
    if (f32 && c1_f32_prefer_resident())
        for (int bn = 256; bn >= 128; bn >>= 1)      // (a 64-wide tile re-reads X once per 64 channels)
            if (N % bn == 0 && fits_resident(bn, K, true)) { bn_max = bn; break; }
    for (int bn = bn_max; bn >= 64; bn >>= 1) {
        if (N % bn) continue;
        long long per = num_sms / (N / bn);
        if (per < 1) per = 1;
        const long long cost = ((m_tiles + per - 1) / per) * bn;
        if (best == 0 || cost * 10 < best_cost * 9) { best = bn; best_cost = cost; }
    }
    return best;
}

void grid_shape(long long M, int N, int K, int num_sms, bool f32, int& bn, int& ctas_per_n, int& grid)
{
    bn = pick_bn(M, N, K, num_sms, f32);
    const int n_blocks = N / bn;
    const long long m_tiles = (M + BM - 1) / BM;
    long long per = num_sms / n_blocks;
    if (per < 1) per = 1;
    if (per > m_tiles) per = m_tiles;
    ctas_per_n = (int)per;
    grid = n_blocks * ctas_per_n;
}

// Shared-memory plan.  W stays resident when the CTA's [BN, K] block plus a >= 3-deep X ring
// fits (then every row tile costs one 16 KB X load instead of X + W); the store slabs shrink
// from 2 to 1 per warp if that is what makes it fit.
C1Plan make_plan(int bn, int K, int n_blocks, bool residual, bool f32, int& smem_bytes)
{
    const int bk = f32 ? Elt<true>::BK : Elt<false>::BK;
    const int k_blocks = (K + bk - 1) / bk;
    const int b_bytes = bn * ROW_BYTES;
    C1Plan p;
    p.x_hint_first = n_blocks == 1;
    p.resident = 0;
    p.nbuf = 2;
    for (int nbuf = 2; nbuf >= (residual ? 2 : 1) && !p.resident; --nbuf) {   // MODE 2 double-buffers its slabs
        const int ring = SMEM_LIMIT - SMEM_FIXED - 8 * nbuf * SLAB_BYTES - k_blocks * b_bytes;
        if (ring >= 3 * A_BYTES) { p.resident = 1; p.nbuf = nbuf; }
    }
    const int stage_bytes = A_BYTES + (p.resident ? 0 : b_bytes);
    const int ring = SMEM_LIMIT - SMEM_FIXED - 8 * p.nbuf * SLAB_BYTES - (p.resident ? k_blocks * b_bytes : 0);
    p.stages = ring / stage_bytes;
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    smem_bytes = SMEM_FIXED + (p.resident ? k_blocks * b_bytes : 0) + p.stages * stage_bytes +
                 8 * p.nbuf * SLAB_BYTES;
    return p;
}

template <int BN, bool F32>
cudaError_t launch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const CUtensorMap& tr,
                   int M, int N, int K, float* partial, int mode, int grid, cudaStream_t st)
{
    auto kern = mode == 1 ? c1_gemm_kernel<BN, 1, F32>
                          : (mode == 2 ? c1_gemm_kernel<BN, 2, F32> : c1_gemm_kernel<BN, 0, F32>);
    static bool configured[3][64] = {};            // function attributes are per device (and per instantiation)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[mode][dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[mode][dev] = true;
    }
    int smem = 0;
    const C1Plan plan = make_plan(BN, K, N / BN, mode == 2, F32, smem);
    if (plan.stages < 2) return cudaErrorInvalidValue;
    kern<<<grid, kThreads, smem, st>>>(tx, tw, ty, tr, M, N, K, partial, plan);
    return cudaGetLastError();
}

}  // namespace

extern "C" {

// shapes the kernel covers: 16-byte row pitch (K % 8 == 0 for bf16, K % 4 == 0 for fp32), N % 64 == 0,
// M < 2^31
int c1_supported(long long M, int N, int K, int f32)
{
    const int kq = f32 ? 4 : 8;
    return M >= 1 && M < (1ll << 31) && N >= 64 && N % 64 == 0 && K >= 8 && K % kq == 0;
}

// launch plan for a shape (host logic only; exercised by the CPU tests):
// out = {BLOCK_N, grid, ctas_per_n, stages, resident, nbuf, dynamic smem bytes}
void c1_describe_plan(long long M, int N, int K, int num_sms, int residual, int f32, int* out)
{
    int bn, per, grid, smem = 0;
    grid_shape(M, N, K, num_sms, f32 != 0, bn, per, grid);
    const C1Plan p = make_plan(bn, K, N / bn, residual != 0, f32 != 0, smem);
    out[0] = bn; out[1] = grid; out[2] = per; out[3] = p.stages; out[4] = p.resident; out[5] = p.nbuf;
    out[6] = smem;
}

// number of partial-statistics rows c1_launch_gemm writes for this shape
int c1_partial_rows(long long M, int N, int K, int num_sms, int f32)
{
    int bn, per, grid;
    grid_shape(M, N, K, num_sms, f32 != 0, bn, per, grid);
    return per;              // one merged partial row per CTA
}

// y[M,N] = x[M,K] . w[N,K]^T [+ residual[M,N]]; all operands bf16 (f32 == 0) or all fp32 with TF32
// tensor-core math (f32 != 0); partial (nullable) = [c1_partial_rows][3][N] fp32 statistics of y.
// partial and residual are mutually exclusive.
cudaError_t c1_launch_gemm(const void* x, const void* w, void* y, long long M, int N, int K, float* partial,
                           const void* residual, int num_sms, int f32, cudaStream_t st)
{
    if (!c1_supported(M, N, K, f32) || (partial && residual)) return cudaErrorInvalidValue;
    const int mode = partial ? 1 : (residual ? 2 : 0);
    const bool f = f32 != 0;
    int bn, per, grid;
    grid_shape(M, N, K, num_sms, f, bn, per, grid);
    CUtensorMap tx, tw, ty, tr;
    if (!make_map(&tx, x, M, K, BM, true, f) || !make_map(&tw, w, N, K, bn, true, f) ||
        !make_map(&ty, y, M, N, 32, false, f) || !make_map(&tr, residual ? residual : y, M, N, 32, false, f))
        return cudaErrorInvalidValue;
    if (f) {
        switch (bn) {
        case 256: return launch<256, true>(tx, tw, ty, tr, (int)M, N, K, partial, mode, grid, st);
        case 128: return launch<128, true>(tx, tw, ty, tr, (int)M, N, K, partial, mode, grid, st);
        default: return launch<64, true>(tx, tw, ty, tr, (int)M, N, K, partial, mode, grid, st);
        }
    }
    switch (bn) {
    case 256: return launch<256, false>(tx, tw, ty, tr, (int)M, N, K, partial, mode, grid, st);
    case 128: return launch<128, false>(tx, tw, ty, tr, (int)M, N, K, partial, mode, grid, st);
    default: return launch<64, false>(tx, tw, ty, tr, (int)M, N, K, partial, mode, grid, st);
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
