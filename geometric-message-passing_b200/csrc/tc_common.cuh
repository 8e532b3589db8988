// Raw PTX wrappers for the Blackwell tensor-core path (tcgen05 / TMEM / mbarrier), sm_100a.
// No CUTLASS: descriptors are built by hand; csrc/schnet_tc.cu's gmp_umma_selftest pins them on hardware.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace gmp {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#ifdef GMP_MBAR_WATCHDOG
// debug build (GMP_NVCC_EXTRA=-DGMP_MBAR_WATCHDOG): a wait that spins for ~a second reports who is stuck on what and traps.
// The spin itself stays inside the asm block so that the timing is as close to the normal build as possible.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .u32 c;\n\t"
        "mov.u32 c, 0;\n\t"
        "WD_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "@p bra WD_DONE;\n\t"
        "add.u32 c, c, 1;\n\t"
        "setp.lt.u32 q, c, 0x400000;\n\t"
        "@q bra WD_LOOP;\n\t"
        "mov.u32 %0, 0;\n\t"
        "bra WD_END;\n\t"
        "WD_DONE:\n\t"
        "mov.u32 %0, 1;\n\t"
        "WD_END:\n\t"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!ok) {
        printf("mbar_wait stuck: block %d thread %d (warp %d) barrier smem+0x%x parity %u\n", (int)blockIdx.x, (int)threadIdx.x,
               (int)(threadIdx.x >> 5), smem_u32(bar), parity);
        __trap();
    }
}
#else
// try_wait suspends the thread until the phase completes or a time limit expires.  Without the optional suspend-time
// hint the limit is a short system default: a waiting warp then comes back every few hundred cycles to re-issue
// try_wait + two branches, and in a warp-specialised kernel where most of the 27 warps wait most of the time that spin
// took 37 % of all issued instructions (profiles/r02_summary.md: BRA + SYNCS.PHASECHK + YIELD in schnet_fwd_tc2_kernel).
// With the hint (10 ms, the value CUTLASS's ClusterBarrier::wait passes) the warp sleeps until the barrier wakes it.
constexpr uint32_t kMbarSuspendHintNs = 0x989680u;
#if defined(GMP_MBAR_SLEEP_NS) && GMP_MBAR_SLEEP_NS > 0
// Variant with a fixed sleep between probes (profiles/r02_summary.md, "mbarrier wait loops"): NANOSLEEP.SYNCS -- what the
// suspend-time hint compiles to -- is cut short by the traffic on the CTA's other barriers, so a waiting warp still probes
// every ~20 cycles; a plain nanosleep is not.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    while (!done) {
        __nanosleep(GMP_MBAR_SLEEP_NS);
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
#else
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendHintNs) : "memory");
}
#endif

#endif

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive + announce `bytes` of asynchronous (bulk-copy) traffic that will complete on this barrier
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine (size: multiple of 16 B, both addresses 16-byte aligned);
// completion is signalled as `bytes` transaction counts on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// named barriers (ids 1..15; id 0 is __syncthreads): producer/consumer hand-off between warp groups
__device__ __forceinline__ void bar_sync_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive_named(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ---- proxies / fences ---------------------------------------------------------------------------
// generic-proxy shared-memory writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane of its warp quarter)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// split form for software pipelining: issue the load, later wait for it (tcgen05.wait::ld waits for every load this thread
// has issued; the registers are tied to the wait statement so that the compiler cannot read them early)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// ---- UMMA descriptors -----------------------------------------------------------------------------
// Shared-memory operand, K-major, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major) |
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups) | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Shared-memory operand, MN-major, 128-byte swizzle: the tile is stored as [K rows][64 bf16 along M/N] slabs
// (row = 128 B, 8-row groups 1024 B apart = stride byte offset; slabs of 64 M/N elements `lbo` bytes apart =
// leading byte offset).  The same bytes read through umma_desc_k128 are the transposed operand.
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, both operands K-major, M x N tile.
//   [4,6) D format = 1 (f32) | [7,10) A format = 1 (bf16) | [10,13) B format = 1 (bf16) | bit 15/16 A/B major = 0 (K) |
//   [17,23) N >> 3 | [24,29) M >> 4
__device__ __forceinline__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn = false, bool b_mn = false) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// one lane of a converged warp (the same one every time): the warp keeps uniform control flow, so descriptors and
// addresses stay in uniform registers instead of being broadcast out of a divergent `lane == 0` branch per instruction
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(pred));
    return pred != 0;
}

// all previously issued MMAs of this thread arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- K-major SWIZZLE_128B tile addressing (bf16) ---------------------------------------------------
// byte offset of the 16-byte chunk holding k in [8*chunk, 8*chunk+8) of row r inside one [rows][64] slab
__device__ __forceinline__ uint32_t sw128_chunk_off(int r, int chunk) { return (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4)); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// K = 16 * nk MMAs over one or more 64-wide K slabs (slab stride in bytes); issued by one thread
__device__ __forceinline__ void umma_tile(uint32_t d_tmem, uint32_t a_base, uint32_t a_slab, uint32_t b_base, uint32_t b_slab,
                                          int K, uint32_t idesc, bool accumulate_first = false) {
    for (int k = 0; k < K; k += 16) {
        const uint32_t ao = a_base + (k >> 6) * a_slab + ((k & 63) >> 4) * 32;
        const uint32_t bo = b_base + (k >> 6) * b_slab + ((k & 63) >> 4) * 32;
        umma_bf16(d_tmem, umma_desc_k128(ao), umma_desc_k128(bo), idesc, (k > 0 || accumulate_first) ? 1u : 0u);
    }
}

// D[M x N] (+)= sum over K = 16*nk tile rows of A[k][m] * B[k][n]: both operands stored [K rows][M or N] (MN-major),
// slabs of 64 columns `a_lbo` / `b_lbo` bytes apart; a K step of 16 rows advances the start address by 2048 B.
__device__ __forceinline__ void umma_tile_mn(uint32_t d_tmem, uint32_t a_base, uint32_t a_lbo, uint32_t b_base, uint32_t b_lbo,
                                             int K, uint32_t idesc, bool accumulate_first) {
    for (int k = 0; k < K; k += 16)
        umma_bf16(d_tmem, umma_desc_mn128(a_base + (k >> 3) * 1024, a_lbo), umma_desc_mn128(b_base + (k >> 3) * 1024, b_lbo), idesc,
                  (k > 0 || accumulate_first) ? 1u : 0u);
}

}  // namespace tc
}  // namespace gmp
