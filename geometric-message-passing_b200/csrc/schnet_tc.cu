// SchNet CFConv forward on the 5th-generation tensor cores (GMP_BF16_TC):
//   per 128-edge tile   rbf (bf16, smem) --tcgen05.mma--> TMEM --ssp--> h1 (bf16, smem) --tcgen05.mma--> TMEM
//                       --(+b2)*C * gathered x1 row--> fp32 smem tile --column threads--> segmented sum.
// bf16 operands, fp32 accumulation in TMEM, everything else fp32: 1e-2 relative against the fp32 reference.
// Same work decomposition as the strict kernel (schnet.cu): whole CSR rows per CTA, persistent grid.
// Also gmp_umma_selftest: one 128 x 128 x K product through exactly the descriptors / swizzle used here.
#include <cuda_pipeline.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gmp {

using namespace tc;

constexpr int kTcTile = 128;
constexpr int kTcRange = 2048;  // edges per work range (16 tiles)

// ------------------------------------------------------------------------------------------------
// self test: out[128][128] = bf16(A[128][K]) * bf16(B[128][K])^T, K in {64, 128}
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ out, int K) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    uint8_t* sA = sm;               // K/64 slabs of [128][64] bf16
    uint8_t* sB = sm + 32768;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;
    // row t of A and B -> swizzled K-major bf16
    for (int kb = 0; kb < K / 64; ++kb)
        for (int ch = 0; ch < 8; ++ch) {
            uint32_t pa[4], pb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = kb * 64 + ch * 8 + 2 * j;
                pa[j] = pack_bf16(A[t * K + k], A[t * K + k + 1]);
                pb[j] = pack_bf16(B[t * K + k], B[t * K + k + 1]);
            }
            *reinterpret_cast<uint4*>(sA + kb * 16384 + sw128_chunk_off(t, ch)) = make_uint4(pa[0], pa[1], pa[2], pa[3]);
            *reinterpret_cast<uint4*>(sB + kb * 16384 + sw128_chunk_off(t, ch)) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
        }
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<128>(&tmem_base);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    if (t == 0) {
        umma_tile(tm, smem_u32(sA), 16384, smem_u32(sB), 16384, K, umma_idesc_bf16(128, 128));
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) out[t * 128 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tm);
}

// self test, MN-major operands: out[128][N] = sum_k bf16(A[k][m]) * bf16(B[k][n]),  A [128 k][128 m], B [128 k][N], N in {64,128}
__global__ void __launch_bounds__(128, 1) umma_selftest_mn_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                  float* __restrict__ out, int N) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    uint8_t* sA = sm;               // 2 slabs [128 k][64 m]
    uint8_t* sB = sm + 32768;       // N/64 slabs [128 k][64 n]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;  // thread t stores tile row k = t
    for (int sl = 0; sl < 2; ++sl)
        for (int ch = 0; ch < 8; ++ch) {
            uint32_t pa[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) pa[j] = pack_bf16(A[t * 128 + sl * 64 + ch * 8 + 2 * j], A[t * 128 + sl * 64 + ch * 8 + 2 * j + 1]);
            *reinterpret_cast<uint4*>(sA + sl * 16384 + sw128_chunk_off(t, ch)) = make_uint4(pa[0], pa[1], pa[2], pa[3]);
        }
    for (int sl = 0; sl < N / 64; ++sl)
        for (int ch = 0; ch < 8; ++ch) {
            uint32_t pb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) pb[j] = pack_bf16(B[t * N + sl * 64 + ch * 8 + 2 * j], B[t * N + sl * 64 + ch * 8 + 2 * j + 1]);
            *reinterpret_cast<uint4*>(sB + sl * 16384 + sw128_chunk_off(t, ch)) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
        }
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<128>(&tmem_base);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    if (t == 0) {
        umma_tile_mn(tm, smem_u32(sA), 16384, smem_u32(sB), 16384, 128, umma_idesc_bf16(128, N, true, true), false);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) out[t * N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tm);
}

// ------------------------------------------------------------------------------------------------
// fused forward
// ------------------------------------------------------------------------------------------------
struct TcArgs {
    const int32_t *rowptr, *col, *perm;
    int64_t n, E;
    const float *ew, *ea, *x1;
    const float *w1, *b1, *w2, *b2, *goff;
    int G;
    float cutoff, gcoeff;
    int nranges;
};

// shared-memory map (bytes from the 1024-aligned base), F = 128
constexpr int kLdX = 132;                       // fp32 row stride of the X / message tile (conflict-free float4 rows)
constexpr int kXBytes = kTcTile * kLdX * 4;     // 66 KB
constexpr int oW1b = 0;                         // [128][64] bf16  16 KB
constexpr int oW2b = oW1b + 16384;              // 2 x [128][64]   32 KB
constexpr int oA2 = oW2b + 32768;               // 2 x [128][64]   32 KB   (h1; slab 0 doubles as A1 = rbf tile)
constexpr int oA1 = oA2;                        // rbf tile is dead once GEMM1 has completed
constexpr int oXt = oA2 + 32768;                // 2 x [128][132] fp32 (double-buffered gather / message tile)
constexpr int oBias = oXt + 2 * kXBytes;        // b1[128], b2[128], goff[64]
constexpr int oScal = oBias + (128 + 128 + 64) * 4;  // 2 x { d[128], C[128] }
constexpr int oInts = oScal + 2 * 2 * 128 * 4;  // 2 x { src[128], eid[128], rowid[128] }
constexpr int kIntsBytes = 3 * 128 * 4;
constexpr int oBar = oInts + 2 * kIntsBytes;    // 2 mbarriers + tmem ptr
constexpr int kTcSmem = oBar + 64 + 1024;       // + alignment slack

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// softplus(x) - ln2 = ln2 * log2(1 + 2^(x log2 e)) - ln2 ; the clamp keeps 2^y finite (x <= 87 is exact to fp32)
__device__ __forceinline__ float ssp_fast(float x) {
    const float e = ex2_approx(fminf(x * 1.4426950408889634f, 126.f));
    return fmaf(0.6931471805599453f, lg2_approx(1.0f + e), -0.6931471805599453f);
}

// Software pipeline per CTA (tiles i of one work range; buffers indexed by i & 1):
//   [barrier: A1(i) ready]  GEMM1 | publish scalars(i+1) | epilogue 1 -> A2 | GEMM2 | issue gather(i+1), wait gather(i) |
//   epilogue 2 (message tile in place) | warps 0-3: column walkers(i)   ||   warps 4-7: rbf(i+1) -> A1
template <bool HAS_ATTR>
__global__ void __launch_bounds__(256, 1) schnet_fwd_tc_kernel(TcArgs a, float* __restrict__ agg) {
    constexpr int F = 128;
    extern __shared__ __align__(16) uint8_t smraw[];
    // align to 1024 B by offsetting the shared array itself (keeps the shared address space for LDS/STS)
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    float* b1s = reinterpret_cast<float*>(sm + oBias);
    float* b2s = b1s + 128;
    float* goff = b2s + 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + oBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + oBar + 32);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    auto Xbuf = [&](int b) { return reinterpret_cast<float*>(sm + oXt + b * kXBytes); };
    auto dsb = [&](int b) { return reinterpret_cast<float*>(sm + oScal + b * 1024); };          // d[128], C[128]
    auto isb = [&](int b) { return reinterpret_cast<int*>(sm + oInts + b * kIntsBytes); };      // src, eid, rowid

    // ---- one-time setup: weights -> bf16 swizzled K-major, barriers, TMEM
    for (int x = t; x < 128 * 8; x += 256) {  // W1 [f][g], g padded to 64
        const int f = x >> 3, ch = x & 7;
        uint32_t p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g0 = ch * 8 + 2 * j;
            p[j] = pack_bf16(g0 < a.G ? __ldg(a.w1 + f * a.G + g0) : 0.f, g0 + 1 < a.G ? __ldg(a.w1 + f * a.G + g0 + 1) : 0.f);
        }
        *reinterpret_cast<uint4*>(sm + oW1b + sw128_chunk_off(f, ch)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
    for (int x = t; x < 128 * 16; x += 256) {  // W2 [f'][f], two K slabs
        const int f = x >> 4, ch16 = x & 15, kb = ch16 >> 3, ch = ch16 & 7;
        const float4 lo = ldg4(a.w2 + f * F + kb * 64 + ch * 8), hi = ldg4(a.w2 + f * F + kb * 64 + ch * 8 + 4);
        *reinterpret_cast<uint4*>(sm + oW2b + kb * 16384 + sw128_chunk_off(f, ch)) =
            make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
    }
    if (t < 128) {
        b1s[t] = __ldg(a.b1 + t);
        b2s[t] = __ldg(a.b2 + t);
    }
    // Gaussian centres; padding columns get a far-away centre so that their basis value underflows to exactly 0
    if (t < 64) goff[t] = (!HAS_ATTR && t < a.G) ? __ldg(a.goff + t) : 1.0e18f;
    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<256>(tmem_ptr);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr;
    const uint32_t tmD1 = tm, tmD2 = tm + 128;
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;  // TMEM lanes this warp may read
    const int row = (warp & 3) * 32 + lane;                        // tile row (edge) owned in the epilogues
    const int chalf = warp >> 2;                                   // column half [64*chalf, 64*chalf + 64)
    const float cw = 3.14159265358979323846f / a.cutoff;
    const float c2 = a.gcoeff * 1.4426950408889634f;               // exp(c u^2) = 2^(c2 u^2)
    uint32_t phase = 0;

    struct Pre { int src, eid, rowid; float d; };
    // edge scalars of one tile: loads (registers) and publication (smem) are split so the latency overlaps compute
    auto load_scalars = [&](int64_t e0, int cnt, int r0, int r1) {
        Pre p{0, 0, r1, 0.f};
        if (t < cnt) {
            const int64_t k = e0 + t;
            p.eid = a.perm ? __ldg(a.perm + k) : (int)k;
            p.src = __ldg(a.col + k);
            p.d = __ldg(a.ew + p.eid);
        }
        {   // CSR row of the edge (last row whose first edge is <= k); padding slots repeat the last valid edge's row
            const int64_t k = e0 + min(t, cnt - 1);
            int lo = r0, hi = r1;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if ((int64_t)__ldg(a.rowptr + mid) <= k) lo = mid; else hi = mid;
            }
            p.rowid = lo;
        }
        return p;
    };
    auto publish_scalars = [&](const Pre& p, int cnt, int b) {
        float* d = dsb(b);
        int* is = isb(b);
        d[t] = t < cnt ? p.d : 1.0e18f;  // padded rows: every Gaussian underflows to 0
        d[128 + t] = t < cnt ? 0.5f * (__cosf(p.d * cw) + 1.0f) : 0.f;
        is[t] = p.src;
        is[128 + t] = p.eid;
        is[256 + t] = p.rowid;
    };
    auto issue_gather = [&](int cnt, int b) {
        float* X = Xbuf(b);
        const int* srcs = isb(b);
        for (int x = t; x < kTcTile * 32; x += 256) {
            const int r = x >> 5, c = x & 31;
            if (r < cnt) __pipeline_memcpy_async(X + r * kLdX + 4 * c, a.x1 + (int64_t)srcs[r] * F + 4 * c, 16);
            else *reinterpret_cast<float4*>(X + r * kLdX + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __pipeline_commit();
    };
    // radial-basis tile of buffer b -> A1 (bf16, swizzled), computed by `nthr` threads with local index `tl`
    auto rbf_tile = [&](int b, int cnt, int tl, int nthr) {
        const float* ds = dsb(b);
        const int* eids = isb(b) + 128;
        for (int x = tl; x < kTcTile * 8; x += nthr) {  // (row, 16-byte chunk)
            const int r = x >> 3, ch = x & 7;
            float v[8];
            if (HAS_ATTR) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int g = ch * 8 + q;
                    v[q] = (r < cnt && g < a.G) ? __ldg(a.ea + (int64_t)eids[r] * a.G + g) : 0.f;
                }
            } else {
                const float d = ds[r];
                const float4 o0 = *reinterpret_cast<const float4*>(goff + ch * 8), o1 = *reinterpret_cast<const float4*>(goff + ch * 8 + 4);
                const float off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float u = d - off[q];
                    v[q] = ex2_approx(c2 * u * u);
                }
            }
            *reinterpret_cast<uint4*>(sm + oA1 + sw128_chunk_off(r, ch)) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
        fence_proxy_async();
    };

    for (int rg = blockIdx.x; rg < a.nranges; rg += gridDim.x) {
        const int r0 = lower_bound_row(a.rowptr, (int)a.n, (int64_t)rg * kTcRange);
        const int r1 = (rg + 1 == a.nranges) ? (int)a.n : lower_bound_row(a.rowptr, (int)a.n, (int64_t)(rg + 1) * kTcRange);
        if (r0 >= r1) continue;
        const int64_t eb = __ldg(a.rowptr + r0), ee = __ldg(a.rowptr + r1);
        int cur = r0;      // walker state (threads < 128): row being accumulated and its running sum
        float acc = 0.f;
        int buf = 0;
        __syncthreads();   // previous range fully drained (shared buffers reusable)
        if (eb < ee) {     // prologue: scalars, gather and radial basis of the first tile
            const int cnt0 = (int)min((int64_t)kTcTile, ee - eb);
            if (t < kTcTile) publish_scalars(load_scalars(eb, cnt0, r0, r1), cnt0, 0);
            __syncthreads();
            issue_gather(cnt0, 0);
            rbf_tile(0, cnt0, t, 256);
        }
        for (int64_t e0 = eb; e0 < ee; e0 += kTcTile, buf ^= 1) {
            const int cnt = (int)min((int64_t)kTcTile, ee - e0);
            const bool has_next = e0 + kTcTile < ee;
            const int cnt_next = has_next ? (int)min((int64_t)kTcTile, ee - e0 - kTcTile) : 0;
            Pre nxt{0, 0, r1, 0.f};
            if (has_next && t < kTcTile) nxt = load_scalars(e0 + kTcTile, cnt_next, r0, r1);
            const float* ds = dsb(buf);
            float* Xt = Xbuf(buf);
            __syncthreads();  // A1(i) complete (written during the previous walker phase / prologue)
            // ---- GEMM1  D1 = rbf W1^T
            if (warp == 0) {  // warp-uniform: one elected lane issues, operands stay in uniform registers
                tc_fence_after();
                if (elect_one()) {
                    umma_tile(tmD1, smem_u32(sm + oA1), 16384, smem_u32(sm + oW1b), 16384, 64, idesc);
                    umma_commit(&bars[0]);
                }
                __syncwarp();
            }
            if (has_next && t < kTcTile) publish_scalars(nxt, cnt_next, buf ^ 1);
            // ---- epilogue 1: h1 = ssp(D1 + b1) -> A2 (bf16, swizzled; slab 0 overwrites the rbf tile GEMM1 has consumed)
            mbar_wait(&bars[0], phase);
            tc_fence_after();
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                float v[32];
                const int c0 = chalf * 64 + part * 32;
                tmem_ld32(tmD1 + lane_base + c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ssp_fast(v[j] + b1s[c0 + j]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ch = part * 4 + q;  // chunk inside K slab `chalf`
                    *reinterpret_cast<uint4*>(sm + oA2 + chalf * 16384 + sw128_chunk_off(row, ch)) =
                        make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                                   pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncthreads();  // also publishes the next tile's scalars
            // ---- GEMM2  D2 = h1 W2^T
            if (warp == 0) {  // warp-uniform: one elected lane issues, operands stay in uniform registers
                tc_fence_after();
                if (elect_one()) {
                    umma_tile(tmD2, smem_u32(sm + oA2), 16384, smem_u32(sm + oW2b), 16384, 128, idesc);
                    umma_commit(&bars[1]);
                }
                __syncwarp();
            }
            // ---- prefetch: gather of the next tile into the other X buffer
            if (has_next) {
                issue_gather(cnt_next, buf ^ 1);
                __pipeline_wait_prior(1);  // everything but the gather just issued: this tile's rows have landed
            } else {
                __pipeline_wait_prior(0);
            }
            __syncthreads();
            // ---- epilogue 2: message = (D2 + b2) * C * x1[src]  (in place over the gathered tile)
            mbar_wait(&bars[1], phase);
            tc_fence_after();
            {
                const float C = ds[128 + row];
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    float v[32];
                    const int c0 = chalf * 64 + part * 32;
                    tmem_ld32(tmD2 + lane_base + c0, v);
                    float* xr = Xt + row * kLdX + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 x = *reinterpret_cast<float4*>(xr + j);
                        x.x *= (v[j] + b2s[c0 + j]) * C;
                        x.y *= (v[j + 1] + b2s[c0 + j + 1]) * C;
                        x.z *= (v[j + 2] + b2s[c0 + j + 2]) * C;
                        x.w *= (v[j + 3] + b2s[c0 + j + 3]) * C;
                        *reinterpret_cast<float4*>(xr + j) = x;
                    }
                }
            }
            tc_fence_before();
            __syncthreads();
            phase ^= 1;
            if (t < F) {
                // ---- warps 0-3: segmented sum, thread = feature column; row ids come from shared memory.
                // Groups of 8 edges: when the whole group still belongs to the open row (the common case) it is a
                // branch-free tree add; otherwise edge by edge.  Padding rows hold zeros and repeat the last row id.
                const int* rid = isb(buf) + 256;
                for (int k0 = 0; k0 < cnt; k0 += 8) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = Xt[(k0 + j) * kLdX + t];
                    const int4 ra = *reinterpret_cast<const int4*>(rid + k0), rb = *reinterpret_cast<const int4*>(rid + k0 + 4);
                    if (rb.w == cur) {
                        acc += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
                    } else {
                        const int r[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (r[j] != cur) {  // row change: store the finished row, zero-fill rows without edges
                                agg[(int64_t)cur * F + t] = acc;
                                for (int z = cur + 1; z < r[j]; ++z) agg[(int64_t)z * F + t] = 0.f;
                                acc = 0.f;
                                cur = r[j];
                            }
                            acc += v[j];
                        }
                    }
                }
            } else if (has_next) {
                // ---- warps 4-7: radial basis of the next tile -> A1 (GEMM2 has completed: A2 slab 0 is free)
                rbf_tile(buf ^ 1, cnt_next, t - 128, 128);
            }
        }
        if (t < F) {
            while (cur < r1) {
                agg[(int64_t)cur * F + t] = acc;
                acc = 0.f;
                ++cur;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tm);
}

// ------------------------------------------------------------------------------------------------
// filter-side backward on tcgen05 (weight gradients only; d edge_weight / d edge_attr use the fp32 kernel)
//
// Per 128-edge tile (e = tile row = TMEM lane; all shared operand tiles are [e][feature] bf16, 128B-swizzled, so the
// same bytes serve as K-major A of the "per-edge" GEMMs and as MN-major operands of the "sum over edges" GEMMs):
//   GEMM1  D1  = R W1^T                      (pre1; R = radial basis with a ones column at g = 63)
//   P      = x1[src] * g[dst] * C            (dpre2, SIMT)            H = ssp(D1 + b1)   (h1)
//   GEMM3  D3  = P W2                        (dh1)
//   acc    dW2 += P^T H ,  db2|. += P^T R    (column 63 of R is 1 -> db2)            [TMEM accumulators, whole kernel]
//   Q      = D3 * sigmoid(pre1)              (dpre1, written over P)
//   acc    dW1|db1 += Q^T R                  (column 63 -> db1)
// TMEM: D1 128 | D3 128 | dW2 128 | dW1 64 | db2 64 = 512 columns.
// ------------------------------------------------------------------------------------------------
constexpr int bW1b = 0;                          // [128 f][64 g]        16 KB
constexpr int bW2T = bW1b + 16384;               // 2 x [128 f][64 f']   32 KB  (W2 transposed)
constexpr int bR = bW2T + 32768;                 // 2 x [128 e][64 g]    32 KB  (double buffered)
constexpr int bP = bR + 32768;                   // 2 x [128 e][64 f']   32 KB  (dpre2, then dpre1)
constexpr int bH = bP + 32768;                   // 2 x [128 e][64 f]    32 KB  (h1)
constexpr int bX = bH + 32768;                   // [128][132] fp32      66 KB
constexpr int bBias = bX + kXBytes;              // b1[128], goff[64]
constexpr int bScal = bBias + (128 + 64) * 4;    // 2 x { d[128], C[128] }
constexpr int bInts = bScal + 2 * 1024;          // 2 x { src[128], eid[128], rowid[128] }
constexpr int bBar = bInts + 2 * kIntsBytes;     // 3 mbarriers + tmem ptr
constexpr int kTcBwdSmem = bBar + 64 + 1024;

template <bool HAS_ATTR>
__global__ void __launch_bounds__(256, 1)
schnet_bwd_tc_kernel(TcArgs a, const float* __restrict__ g_agg, float* __restrict__ parts) {
    constexpr int F = 128;
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    float* b1s = reinterpret_cast<float*>(sm + bBias);
    float* goff = b1s + 128;
    float* Xt = reinterpret_cast<float*>(sm + bX);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + bBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + bBar + 32);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    auto dsb = [&](int b) { return reinterpret_cast<float*>(sm + bScal + b * 1024); };
    auto isb = [&](int b) { return reinterpret_cast<int*>(sm + bInts + b * kIntsBytes); };

    for (int x = t; x < 128 * 8; x += 256) {  // W1 [f][g] K-major, g padded to 64 with zeros
        const int f = x >> 3, ch = x & 7;
        uint32_t p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g0 = ch * 8 + 2 * j;
            p[j] = pack_bf16(g0 < a.G ? __ldg(a.w1 + f * a.G + g0) : 0.f, g0 + 1 < a.G ? __ldg(a.w1 + f * a.G + g0 + 1) : 0.f);
        }
        *reinterpret_cast<uint4*>(sm + bW1b + sw128_chunk_off(f, ch)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
    for (int x = t; x < 128 * 16; x += 256) {  // W2T [f][f'] = W2[f'][f], K = f' in two slabs
        const int f = x >> 4, ch16 = x & 15, kb = ch16 >> 3, ch = ch16 & 7;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __ldg(a.w2 + (int64_t)(kb * 64 + ch * 8 + q) * F + f);
        *reinterpret_cast<uint4*>(sm + bW2T + kb * 16384 + sw128_chunk_off(f, ch)) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
    if (t < 128) b1s[t] = __ldg(a.b1 + t);
    if (t < 64) goff[t] = (!HAS_ATTR && t < a.G) ? __ldg(a.goff + t) : 1.0e18f;
    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_ptr);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr;
    const uint32_t tmD1 = tm, tmD3 = tm + 128, tmW2 = tm + 256, tmW1 = tm + 384, tmB2 = tm + 448;
    const uint32_t id_kk = umma_idesc_bf16(128, 128), id_mn128 = umma_idesc_bf16(128, 128, true, true),
                   id_mn64 = umma_idesc_bf16(128, 64, true, true);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int row = (warp & 3) * 32 + lane;
    const int chalf = warp >> 2;
    const float cw = 3.14159265358979323846f / a.cutoff;
    const float c2 = a.gcoeff * 1.4426950408889634f;
    uint32_t it = 0;  // tiles processed by this CTA (mbarrier phase = it & 1)

    struct Pre { int src, eid, rowid; float d; };
    auto load_scalars = [&](int64_t e0, int cnt, int r0, int r1) {
        Pre p{0, 0, r0, 0.f};
        if (t < cnt) {
            const int64_t k = e0 + t;
            p.eid = a.perm ? __ldg(a.perm + k) : (int)k;
            p.src = __ldg(a.col + k);
            p.d = __ldg(a.ew + p.eid);
            int lo = r0, hi = r1;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if ((int64_t)__ldg(a.rowptr + mid) <= k) lo = mid; else hi = mid;
            }
            p.rowid = lo;
        }
        return p;
    };
    auto publish_scalars = [&](const Pre& p, int cnt, int b) {
        float* d = dsb(b);
        int* is = isb(b);
        d[t] = t < cnt ? p.d : 1.0e18f;
        d[128 + t] = t < cnt ? 0.5f * (__cosf(p.d * cw) + 1.0f) : 0.f;
        is[t] = p.src;
        is[128 + t] = p.eid;
        is[256 + t] = p.rowid;
    };
    auto issue_gather = [&](int cnt, int b) {
        const int* srcs = isb(b);
        for (int x = t; x < kTcTile * 32; x += 256) {
            const int r = x >> 5, c = x & 31;
            if (r < cnt) __pipeline_memcpy_async(Xt + r * kLdX + 4 * c, a.x1 + (int64_t)srcs[r] * F + 4 * c, 16);
            else *reinterpret_cast<float4*>(Xt + r * kLdX + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __pipeline_commit();
    };
    auto rbf_tile = [&](int b, int cnt) {  // -> R[b], with a ones column at g = 63
        const float* ds = dsb(b);
        const int* eids = isb(b) + 128;
        for (int x = t; x < kTcTile * 8; x += 256) {
            const int r = x >> 3, ch = x & 7;
            float v[8];
            if (HAS_ATTR) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int g = ch * 8 + q;
                    v[q] = (r < cnt && g < a.G) ? __ldg(a.ea + (int64_t)eids[r] * a.G + g) : 0.f;
                }
            } else {
                const float d = ds[r];
                const float4 o0 = *reinterpret_cast<const float4*>(goff + ch * 8), o1 = *reinterpret_cast<const float4*>(goff + ch * 8 + 4);
                const float off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float u = d - off[q];
                    v[q] = ex2_approx(c2 * u * u);
                }
            }
            if (ch == 7) v[7] = 1.0f;
            *reinterpret_cast<uint4*>(sm + bR + b * 16384 + sw128_chunk_off(r, ch)) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
        fence_proxy_async();
    };

    for (int rg = blockIdx.x; rg < a.nranges; rg += gridDim.x) {
        const int r0 = lower_bound_row(a.rowptr, (int)a.n, (int64_t)rg * kTcRange);
        const int r1 = (rg + 1 == a.nranges) ? (int)a.n : lower_bound_row(a.rowptr, (int)a.n, (int64_t)(rg + 1) * kTcRange);
        if (r0 >= r1) continue;
        const int64_t eb = __ldg(a.rowptr + r0), ee = __ldg(a.rowptr + r1);
        if (eb >= ee) continue;
        int buf = 0;
        if (it > 0) {  // the previous range's last dW1 MMA still reads P and R[.]: drain it before reusing the buffers
            mbar_wait(&bars[2], (it - 1) & 1);
            tc_fence_after();
        }
        __syncthreads();
        {
            const int cnt0 = (int)min((int64_t)kTcTile, ee - eb);
            if (t < kTcTile) publish_scalars(load_scalars(eb, cnt0, r0, r1), cnt0, 0);
            __syncthreads();
            issue_gather(cnt0, 0);
            rbf_tile(0, cnt0);
        }
        bool drained = true;  // no dW1 MMA outstanding on P / R[buf ^ 1]
        for (int64_t e0 = eb; e0 < ee; e0 += kTcTile, buf ^= 1, ++it) {
            const int cnt = (int)min((int64_t)kTcTile, ee - e0);
            const bool has_next = e0 + kTcTile < ee;
            const int cnt_next = has_next ? (int)min((int64_t)kTcTile, ee - e0 - kTcTile) : 0;
            const uint32_t ph = it & 1;
            Pre nxt{0, 0, r0, 0.f};
            if (has_next && t < kTcTile) nxt = load_scalars(e0 + kTcTile, cnt_next, r0, r1);
            if (!drained) {  // dW1 MMA of the previous tile (reads P as Q and R[buf ^ 1])
                mbar_wait(&bars[2], (it - 1) & 1);
                tc_fence_after();
            }
            __pipeline_wait_prior(0);  // x1 rows of this tile
            __syncthreads();           // [A] R[buf], X, scalars(buf) complete
            const uint32_t Rb = smem_u32(sm + bR + buf * 16384), Pb = smem_u32(sm + bP), Hb = smem_u32(sm + bH);
            if (warp == 0) {  // warp-uniform: one elected lane issues, operands stay in uniform registers
                tc_fence_after();
                if (elect_one()) {
                    umma_tile(tmD1, Rb, 16384, smem_u32(sm + bW1b), 16384, 64, id_kk);
                    umma_commit(&bars[0]);
                }
                __syncwarp();
            }
            // ---- P = x1[src] * g[dst] * C   (bf16, [e][f'])
            {
                const float C = dsb(buf)[128 + row];
                const float* grow = g_agg + (int64_t)isb(buf)[256 + row] * F;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const int c = chalf * 64 + ch * 8;
                    const float4 x0 = *reinterpret_cast<const float4*>(Xt + row * kLdX + c), x1v = *reinterpret_cast<const float4*>(Xt + row * kLdX + c + 4);
                    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
                    if (row < cnt) {
                        g0 = ldg4(grow + c);
                        g1 = ldg4(grow + c + 4);
                    }
                    *reinterpret_cast<uint4*>(sm + bP + chalf * 16384 + sw128_chunk_off(row, ch)) =
                        make_uint4(pack_bf16(x0.x * g0.x * C, x0.y * g0.y * C), pack_bf16(x0.z * g0.z * C, x0.w * g0.w * C),
                                   pack_bf16(x1v.x * g1.x * C, x1v.y * g1.y * C), pack_bf16(x1v.z * g1.z * C, x1v.w * g1.w * C));
                }
            }
            if (has_next && t < kTcTile) publish_scalars(nxt, cnt_next, buf ^ 1);
            // ---- H = ssp(D1 + b1)   (bf16, [e][f])
            mbar_wait(&bars[0], ph);
            tc_fence_after();
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                float v[32];
                const int c0 = chalf * 64 + part * 32;
                tmem_ld32(tmD1 + lane_base + c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ssp_fast(v[j] + b1s[c0 + j]);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(sm + bH + chalf * 16384 + sw128_chunk_off(row, part * 4 + q)) =
                        make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                                   pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
            }
            tc_fence_before();
            fence_proxy_async();
            __syncthreads();  // [B] P, H complete; X consumed; next scalars published
            if (warp == 0) {  // warp-uniform: one elected lane issues, operands stay in uniform registers
                tc_fence_after();
                if (elect_one()) {
                    umma_tile(tmD3, Pb, 16384, smem_u32(sm + bW2T), 16384, 128, id_kk);       // dh1 = P W2
                    umma_tile_mn(tmW2, Pb, 16384, Hb, 16384, 128, id_mn128, it > 0);            // dW2 += P^T H
                    umma_tile_mn(tmB2, Pb, 16384, Rb, 16384, 128, id_mn64, it > 0);             // [. | db2] += P^T R
                    umma_commit(&bars[1]);
                }
                __syncwarp();
            }
            if (has_next) issue_gather(cnt_next, buf ^ 1);
            // ---- Q = D3 * sigmoid(pre1), sigmoid from h1: 1 - exp(-softplus) = 1 - 0.5 * 2^(-h1 log2 e)
            mbar_wait(&bars[1], ph);
            tc_fence_after();
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                float v[32];
                const int c0 = chalf * 64 + part * 32;
                tmem_ld32(tmD3 + lane_base + c0, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t off = chalf * 16384 + sw128_chunk_off(row, part * 4 + q);
                    const uint4 hp = *reinterpret_cast<const uint4*>(sm + bH + off);
                    const uint32_t hw[4] = {hp.x, hp.y, hp.z, hp.w};
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __nv_bfloat162 hb = *reinterpret_cast<const __nv_bfloat162*>(&hw[j]);
                        const float2 hf = __bfloat1622float2(hb);
                        const float s0 = fmaf(-0.5f, ex2_approx(-1.4426950408889634f * hf.x), 1.0f);
                        const float s1 = fmaf(-0.5f, ex2_approx(-1.4426950408889634f * hf.y), 1.0f);
                        o[j] = pack_bf16(v[8 * q + 2 * j] * s0, v[8 * q + 2 * j + 1] * s1);
                    }
                    *reinterpret_cast<uint4*>(sm + bP + off) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncthreads();  // [C] Q complete
            if (warp == 0) {  // warp-uniform: one elected lane issues, operands stay in uniform registers
                tc_fence_after();
                if (elect_one()) {
                    umma_tile_mn(tmW1, Pb, 16384, Rb, 16384, 128, id_mn64, it > 0);             // [dW1 | db1] += Q^T R
                    umma_commit(&bars[2]);
                }
                __syncwarp();
            }
            drained = false;
            if (has_next) rbf_tile(buf ^ 1, cnt_next);  // R[buf ^ 1] was last read by the tile before this one (drained above)
        }
    }
    // ---- drain, then write this CTA's partial gradients: [dW1 F x 64 | db1 F | dW2 F x F | db2 F]
    float* my = parts + (int64_t)blockIdx.x * (F * 64 + F + F * F + F);
    if (it > 0) {
        mbar_wait(&bars[2], (it - 1) & 1);
        tc_fence_after();
        for (int c0 = chalf * 64; c0 < chalf * 64 + 64; c0 += 32) {  // dW2[f' = row][f]
            float v[32];
            tmem_ld32(tmW2 + lane_base + c0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(my + F * 64 + F + row * F + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        {   // dW1[f = row][g] (32 columns per warp half); column 63 = db1
            float v[32];
            const int c0 = chalf * 32;
            tmem_ld32(tmW1 + lane_base + c0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(my + row * 64 + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (chalf == 1) my[F * 64 + row] = v[31];
        }
        if (chalf == 1) {  // db2[f' = row] = column 63 of the P^T R accumulator
            float v[32];
            tmem_ld32(tmB2 + lane_base + 32, v);
            my[F * 64 + F + F * F + row] = v[31];
        }
    } else {
        for (int x = t; x < F * 64 + F + F * F + F; x += 256) my[x] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

int schnet_bwd_tc_launch(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t E,
                         const float* ew, const float* ea, const float* x1, const gmp_schnet_filter* f, const float* g_agg,
                         float* parts, int nparts, cudaStream_t stream) {
    TcArgs a;
    a.rowptr = rowptr; a.col = col; a.perm = perm; a.n = n; a.E = E; a.ew = ew; a.ea = ea; a.x1 = x1;
    a.w1 = f->w1; a.b1 = f->b1; a.w2 = f->w2; a.b2 = f->b2; a.goff = f->gauss_offset;
    a.G = f->num_gaussians; a.cutoff = f->cutoff; a.gcoeff = f->gauss_coeff;
    a.nranges = (int)(E > 0 ? ceil_div(E, kTcRange) : 1);
    const int grid = nparts;  // the caller sized the partial buffer with gmp_schnet_bwd_num_parts
    if (ea) {
        GMP_CUDA(cudaFuncSetAttribute(schnet_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcBwdSmem));
        schnet_bwd_tc_kernel<true><<<grid, 256, kTcBwdSmem, stream>>>(a, g_agg, parts);
    } else {
        GMP_CUDA(cudaFuncSetAttribute(schnet_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcBwdSmem));
        schnet_bwd_tc_kernel<false><<<grid, 256, kTcBwdSmem, stream>>>(a, g_agg, parts);
    }
    return check_launch("schnet_bwd_tc_kernel");
}

int schnet_fwd_tc_launch(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t E,
                         const float* ew, const float* ea, const float* x1, const gmp_schnet_filter* f, float* agg,
                         cudaStream_t stream) {
    GMP_REQUIRE(f->num_filters == 128, "schnet (bf16 tensor-core path): num_filters must be 128 (got %d)", f->num_filters);
    TcArgs a;
    a.rowptr = rowptr; a.col = col; a.perm = perm; a.n = n; a.E = E; a.ew = ew; a.ea = ea; a.x1 = x1;
    a.w1 = f->w1; a.b1 = f->b1; a.w2 = f->w2; a.b2 = f->b2; a.goff = f->gauss_offset;
    a.G = f->num_gaussians; a.cutoff = f->cutoff; a.gcoeff = f->gauss_coeff;
    a.nranges = (int)(E > 0 ? ceil_div(E, kTcRange) : 1);
    const int grid = a.nranges < num_sms() ? a.nranges : num_sms();
    if (ea) {
        GMP_CUDA(cudaFuncSetAttribute(schnet_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
        schnet_fwd_tc_kernel<true><<<grid, 256, kTcSmem, stream>>>(a, agg);
    } else {
        GMP_CUDA(cudaFuncSetAttribute(schnet_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
        schnet_fwd_tc_kernel<false><<<grid, 256, kTcSmem, stream>>>(a, agg);
    }
    return check_launch("schnet_fwd_tc_kernel");
}

}  // namespace gmp

using namespace gmp;

extern "C" int gmp_umma_selftest(const float* A, const float* B, float* out, int32_t K, gmp_stream_t stream) {
    GMP_REQUIRE(A && B && out && (K == 64 || K == 128), "umma_selftest: K must be 64 or 128");
    const int smem = 65536 + 1024;
    GMP_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_selftest_kernel<<<1, 128, smem, stream>>>(A, B, out, K);
    return check_launch("umma_selftest_kernel");
}

extern "C" int gmp_umma_selftest_mn(const float* A, const float* B, float* out, int32_t N, gmp_stream_t stream) {
    GMP_REQUIRE(A && B && out && (N == 64 || N == 128), "umma_selftest_mn: N must be 64 or 128");
    const int smem = 65536 + 1024;
    GMP_CUDA(cudaFuncSetAttribute(umma_selftest_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_selftest_mn_kernel<<<1, 128, smem, stream>>>(A, B, out, N);
    return check_launch("umma_selftest_mn_kernel");
}
