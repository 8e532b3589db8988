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
    extern __shared__ __align__(1024) uint8_t smraw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~uintptr_t(1023));
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
constexpr int oW1b = 0;                         // [128][64] bf16  16 KB
constexpr int oW2b = oW1b + 16384;              // 2 x [128][64]   32 KB
constexpr int oA1 = oW2b + 32768;               // [128][64]       16 KB
constexpr int oA2 = oA1 + 16384;                // 2 x [128][64]   32 KB
constexpr int oXt = oA2 + 32768;                // [128][132] fp32 66 KB
constexpr int oBias = oXt + kTcTile * kLdX * 4; // b1[128], b2[128], goff[64]
constexpr int oScal = oBias + (128 + 128 + 64) * 4;  // d[128], C[128]
constexpr int oInts = oScal + 2 * 128 * 4;      // src[128], eid[128]
constexpr int oBar = oInts + 2 * 128 * 4;       // 2 mbarriers + tmem ptr
constexpr int kTcSmem = oBar + 64 + 1024;       // + alignment slack

__device__ __forceinline__ float ssp_fast(float x) {
    // softplus(x) - ln2 = ln2 * (log2(1 + 2^(x log2 e)) - 1); ex2/lg2 approximations (rel. err ~1e-7 << bf16)
    const float e = exp2f(x * 1.4426950408889634f);
    const float sp = x > 15.f ? x : 0.6931471805599453f * __log2f(1.0f + e);
    return sp - 0.6931471805599453f;
}

template <bool HAS_ATTR>
__global__ void __launch_bounds__(256, 1) schnet_fwd_tc_kernel(TcArgs a, float* __restrict__ agg) {
    constexpr int F = 128;
    extern __shared__ __align__(1024) uint8_t smraw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~uintptr_t(1023));
    float* Xt = reinterpret_cast<float*>(sm + oXt);
    float* b1s = reinterpret_cast<float*>(sm + oBias);
    float* b2s = b1s + 128;
    float* goff = b2s + 128;
    float* ds = reinterpret_cast<float*>(sm + oScal);
    float* Cs = ds + 128;
    int* srcs = reinterpret_cast<int*>(sm + oInts);
    int* eids = srcs + 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + oBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + oBar + 32);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

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
    if (t < 64) goff[t] = (!HAS_ATTR && t < a.G) ? __ldg(a.goff + t) : 0.f;
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
    uint32_t phase = 0;

    for (int rg = blockIdx.x; rg < a.nranges; rg += gridDim.x) {
        const int r0 = lower_bound_row(a.rowptr, (int)a.n, (int64_t)rg * kTcRange);
        const int r1 = (rg + 1 == a.nranges) ? (int)a.n : lower_bound_row(a.rowptr, (int)a.n, (int64_t)(rg + 1) * kTcRange);
        if (r0 >= r1) continue;
        const int64_t eb = __ldg(a.rowptr + r0), ee = __ldg(a.rowptr + r1);
        int cur = r0;
        int64_t row_end = __ldg(a.rowptr + r0 + 1);
        float acc = 0.f;
        for (int64_t e0 = eb; e0 < ee; e0 += kTcTile) {
            const int cnt = (int)min((int64_t)kTcTile, ee - e0);
            // ---- S1: per-edge scalars
            if (t < kTcTile) {
                float d = 0.f, C = 0.f;
                int src = 0, eid = 0;
                if (t < cnt) {
                    const int64_t k = e0 + t;
                    eid = a.perm ? __ldg(a.perm + k) : (int)k;
                    src = __ldg(a.col + k);
                    d = __ldg(a.ew + eid);
                    C = 0.5f * (__cosf(d * 3.14159265358979323846f / a.cutoff) + 1.0f);
                }
                ds[t] = d; Cs[t] = C; srcs[t] = src; eids[t] = eid;
            }
            __syncthreads();
            // ---- S2: async gather of the x1 rows
            for (int x = t; x < kTcTile * 32; x += 256) {
                const int r = x >> 5, c = x & 31;
                if (r < cnt) __pipeline_memcpy_async(Xt + r * kLdX + 4 * c, a.x1 + (int64_t)srcs[r] * F + 4 * c, 16);
                else *reinterpret_cast<float4*>(Xt + r * kLdX + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __pipeline_commit();
            // ---- S3: radial basis tile -> A1 (bf16, swizzled).  thread = (row t/2, 4 chunks)
            {
                const int r = t >> 1, h = t & 1;
                const float d = ds[r];
                const bool valid = r < cnt;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = h * 4 + j;
                    float v[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int g = ch * 8 + q;
                        float val = 0.f;
                        if (valid && g < a.G) {
                            if (HAS_ATTR) {
                                val = __ldg(a.ea + (int64_t)eids[r] * a.G + g);
                            } else {
                                const float u = d - goff[g];
                                val = __expf(a.gcoeff * u * u);
                            }
                        }
                        v[q] = val;
                    }
                    *reinterpret_cast<uint4*>(sm + oA1 + sw128_chunk_off(r, ch)) =
                        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                }
            }
            fence_proxy_async();
            __syncthreads();
            // ---- S4: GEMM1  D1 = rbf W1^T
            if (t == 0) {
                tc_fence_after();
                umma_tile(tmD1, smem_u32(sm + oA1), 16384, smem_u32(sm + oW1b), 16384, 64, idesc);
                umma_commit(&bars[0]);
            }
            // ---- S5: epilogue 1: h1 = ssp(D1 + b1) -> A2 (bf16, swizzled)
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
            __syncthreads();
            // ---- S6: GEMM2  D2 = h1 W2^T
            if (t == 0) {
                tc_fence_after();
                umma_tile(tmD2, smem_u32(sm + oA2), 16384, smem_u32(sm + oW2b), 16384, 128, idesc);
                umma_commit(&bars[1]);
            }
            // ---- S7: epilogue 2: message = (D2 + b2) * C * x1[src]  (in place over the gathered tile)
            __pipeline_wait_prior(0);
            __syncthreads();
            mbar_wait(&bars[1], phase);
            tc_fence_after();
            {
                const float C = Cs[row];
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
            // ---- S8: segmented sum by column threads
            if (t < F) {
                for (int k = 0; k < cnt; ++k) {
                    const int64_t e = e0 + k;
                    while (e >= row_end) {
                        agg[(int64_t)cur * F + t] = acc;
                        acc = 0.f;
                        ++cur;
                        row_end = __ldg(a.rowptr + cur + 1);
                    }
                    acc += Xt[k * kLdX + t];
                }
            }
            __syncthreads();
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
