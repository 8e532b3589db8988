// Parameter gradients of a node-side nn.Linear (y = x W^T + b) in the GMP_BF16_TC mode:
//     dW[out, in] = sum_n g[n, out] * x[n, in],     db[out] = sum_n g[n, out]
// for out = 128, in in {64, 128}: a reduction over all N nodes with a tiny output, which cuBLAS runs on a handful of
// CTAs (the K = N dimension is not split) and ATen's column sum at a fraction of the memory bandwidth.  Here the rows
// are split over all SMs; each CTA streams its 128-row tiles of g and x once (fp32 -> bf16 on the fly, written as
// MN-major UMMA operand images), accumulates g^T x in tensor memory over its whole row range, adds the column sums of
// g from registers, and writes one partial; gmp_reduce_partials_f32 sums the partials in a fixed order.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gmp {

using namespace tc;

constexpr int kLT = 64;         // rows per tile (the K extent of one MMA group); small tiles so that several CTAs share an SM and
                                // one CTA's load -> convert -> barrier bubble is covered by the others' loads
constexpr int kLThreads = 256;
constexpr int kLCtasPerSm = 3;  // 64 KB of shared memory and 128 tensor-memory columns each
constexpr int kLSlab = kLT * 128;   // one slab: [kLT rows][64 columns] bf16
constexpr int kLImg = 2 * kLSlab;   // one operand image: 2 slabs
constexpr int oLBar = 4 * kLImg;  // 2 buffers x (g image, x image), then 2 mbarriers + tmem pointer
constexpr int kLinSmem = oLBar + 64 + 1024;

__global__ void __launch_bounds__(kLThreads, kLCtasPerSm)
linear_wgrad_tc_kernel(const float* __restrict__ g, const float* __restrict__ x, int64_t n, int in_dim, float* __restrict__ parts) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + oLBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + oLBar + 16);
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<128>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr;
    const int64_t ntiles = (n + kLT - 1) / kLT;
    const int64_t t0 = (ntiles * blockIdx.x) / gridDim.x, t1 = (ntiles * (blockIdx.x + 1)) / gridDim.x;
    const int ch = t & 15, r0 = t >> 4;          // this thread's 8-column chunk and first row (rows r0 + 16 i)
    const int xch = in_dim >> 3;                 // chunks per x row (8 or 16)
    float bsum[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) bsum[u] = 0.f;
    const uint32_t idesc = umma_idesc_bf16(128, in_dim, true, true);
    uint32_t it = 0;
    for (int64_t tile = t0; tile < t1; ++tile, ++it) {
        const uint32_t buf = it & 1u;
        if (it >= 2) mbar_wait(&bars[buf], ((it >> 1) - 1) & 1u);  // the MMAs that read this buffer two tiles ago are done
        uint8_t* gi = sm + buf * 2 * kLImg;
        uint8_t* xi = gi + kLImg;
        const int64_t row0 = tile * kLT;
#pragma unroll
        for (int i = 0; i < kLT / 16; ++i) {
            const int r = r0 + 16 * i;
            const int64_t row = row0 + r;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (row < n) {
                a = ldg4(g + row * 128 + ch * 8);
                b = ldg4(g + row * 128 + ch * 8 + 4);
            }
            bsum[0] += a.x; bsum[1] += a.y; bsum[2] += a.z; bsum[3] += a.w;
            bsum[4] += b.x; bsum[5] += b.y; bsum[6] += b.z; bsum[7] += b.w;
            *reinterpret_cast<uint4*>(gi + (ch >> 3) * kLSlab + sw128_chunk_off(r, ch & 7)) =
                make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
            if (ch < xch) {
                float4 c = make_float4(0.f, 0.f, 0.f, 0.f), d = c;
                if (row < n) {
                    c = ldg4(x + row * in_dim + ch * 8);
                    d = ldg4(x + row * in_dim + ch * 8 + 4);
                }
                *reinterpret_cast<uint4*>(xi + (ch >> 3) * kLSlab + sw128_chunk_off(r, ch & 7)) =
                    make_uint4(pack_bf16(c.x, c.y), pack_bf16(c.z, c.w), pack_bf16(d.x, d.y), pack_bf16(d.z, d.w));
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                umma_tile_mn(tm, smem_u32(gi), kLSlab, smem_u32(xi), kLSlab, kLT, idesc, it > 0);
                umma_commit(&bars[buf]);
            }
            __syncwarp();
        }
    }
    // wait for the last MMAs on both buffers
    if (it >= 1) mbar_wait(&bars[(it - 1) & 1u], ((it - 1) >> 1) & 1u);
    if (it >= 2) mbar_wait(&bars[it & 1u], ((it - 2) >> 1) & 1u);
    tc_fence_after();
    const int64_t plen = (int64_t)128 * in_dim + 128;
    float* my = parts + (int64_t)blockIdx.x * plen;
    if (warp < 4) {  // lane = output feature, columns = input features
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int c0 = 0; c0 < in_dim; c0 += 32) {
            float v[32];
            tmem_ld32(tm + lane_base + c0, v);
            float* dst = my + (int64_t)t * in_dim + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) = it > 0 ? make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    // column sums of g: 16 row groups per chunk
    float* red = reinterpret_cast<float*>(sm);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 8; ++u) red[r0 * 128 + ch * 8 + u] = bsum[u];
    __syncthreads();
    if (t < 128) {
        float s = 0.f;
        for (int k = 0; k < 16; ++k) s += red[k * 128 + t];
        my[(int64_t)128 * in_dim + t] = s;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tm);
}


// ------------------------------------------------------------------------------------------------------------------------
// in = 128: the same reduction as a bulk-copy pipeline (round 2).  Rows of g and x are contiguous in memory, so a 32-row tile
// of each is ONE 16 KB cp.async.bulk; a producer warp keeps kW3Stages tiles in flight per CTA (two CTAs per SM) and the eight
// consumer warps never wait on a global load: they convert a landed tile from its fp32 staging buffer into the bf16 MN-major
// operand images (conflict-free: consecutive lanes read consecutive 16-byte groups), add the column sums of g, free the
// staging slot and hand the images to the tensor core.  The first kernel above kept its loads in registers (8 per thread in
// flight) and ran at 50 % of the HBM rate (profiles/r02_summary.md).
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kW3T = 32;                       // rows per tile
constexpr int kW3Stages = 2;                   // staging ring depth per CTA (x 2 CTAs per SM)
constexpr int kW3Consumers = 256;
constexpr int kW3Threads = kW3Consumers + 32;  // + the producer warp
constexpr int kW3Tile = kW3T * 128 * 4;        // one fp32 tile: 16 KB
constexpr int kW3Slab = kW3T * 128;            // one image slab: [32 rows][64 bf16]
constexpr int kW3Img = 2 * kW3Slab;            // one operand image: 8 KB
constexpr int o5Stage = 0;                                    // [stage][g | x] fp32 tiles
constexpr int o5Img = o5Stage + kW3Stages * 2 * kW3Tile;      // [buffer][g | x] bf16 images
constexpr int o5Bar = o5Img + 2 * 2 * kW3Img;                 // full[stages], empty[stages], imgfree[2], tmem pointer
constexpr int kW3Smem = o5Bar + (2 * kW3Stages + 2) * 8 + 16 + 1024;

__global__ void __launch_bounds__(kW3Threads, 2)
linear_wgrad_tc3_kernel(const float* __restrict__ g, const float* __restrict__ x, int64_t n, float* __restrict__ parts) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + o5Bar);
    uint64_t* empty = full + kW3Stages;
    uint64_t* imgfree = empty + kW3Stages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(imgfree + 2);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t == 0) {
        for (int i = 0; i < kW3Stages; ++i) {
            mbar_init(&full[i], 1);                 // the producer's expect_tx arrival + the bytes
            mbar_init(&empty[i], kW3Consumers / 32);   // one elected arrival per consumer warp
        }
        mbar_init(&imgfree[0], 1);
        mbar_init(&imgfree[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<128>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr;
    const int64_t ntiles = (n + kW3T - 1) / kW3T;
    const int64_t t0 = (ntiles * blockIdx.x) / gridDim.x, t1 = (ntiles * (blockIdx.x + 1)) / gridDim.x;
    const uint32_t nt = (uint32_t)(t1 - t0);

    if (warp == kW3Consumers / 32) {
        // ===================== producer: one lane issues two bulk copies per tile =====================
        if (lane == 0) {
            for (uint32_t it = 0; it < nt; ++it) {
                const uint32_t s = it % kW3Stages, use = it / kW3Stages;
                if (use > 0) mbar_wait(&empty[s], (use - 1) & 1u);
                const int64_t row0 = (t0 + it) * kW3T;
                const uint32_t rows = (uint32_t)min((int64_t)kW3T, n - row0);
                const uint32_t bytes = rows * 512u;
                mbar_expect_tx(&full[s], 2 * bytes);
                bulk_g2s(sm + o5Stage + (s * 2 + 0) * kW3Tile, g + row0 * 128, bytes, &full[s]);
                bulk_g2s(sm + o5Stage + (s * 2 + 1) * kW3Tile, x + row0 * 128, bytes, &full[s]);
            }
        }
    } else {
        // ===================== consumers: thread = float4 column c4 of rows (w, w + 8, w + 16, w + 24) =====================
        const int c4 = t & 31, w = t >> 5;
        float bsum[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t idesc = umma_idesc_bf16(128, 128, true, true);
        for (uint32_t it = 0; it < nt; ++it) {
            const uint32_t s = it % kW3Stages, use = it / kW3Stages, b = it & 1u;
            const int64_t row0 = (t0 + it) * kW3T;
            const int rows = (int)min((int64_t)kW3T, n - row0);
            mbar_wait(&full[s], use & 1u);
            if (it >= 2) mbar_wait(&imgfree[b], ((it >> 1) - 1) & 1u);   // the MMAs that read image buffer b two tiles ago are done
            const float4* gs = reinterpret_cast<const float4*>(sm + o5Stage + (s * 2 + 0) * kW3Tile);
            const float4* xs = reinterpret_cast<const float4*>(sm + o5Stage + (s * 2 + 1) * kW3Tile);
            uint8_t* gi = sm + o5Img + (b * 2 + 0) * kW3Img;
            uint8_t* xi = sm + o5Img + (b * 2 + 1) * kW3Img;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = w + 8 * i;
                float4 a = gs[r * 32 + c4], c = xs[r * 32 + c4];
                if (r >= rows) { a = make_float4(0.f, 0.f, 0.f, 0.f); c = a; }   // (stale bytes of an earlier tile beyond the copy)
                bsum[0] += a.x; bsum[1] += a.y; bsum[2] += a.z; bsum[3] += a.w;
                // columns 4 c4 .. 4 c4 + 3 = half of the 16-byte chunk (c4 >> 1) of slab (c4 >> 4)
                const uint32_t off = (uint32_t)((c4 >> 4) * kW3Slab) + sw128_chunk_off(r, (c4 >> 1) & 7) + (c4 & 1) * 8;
                *reinterpret_cast<uint2*>(gi + off) = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
                *reinterpret_cast<uint2*>(xi + off) = make_uint2(pack_bf16(c.x, c.y), pack_bf16(c.z, c.w));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);      // this warp has read its part of the staging slot
            tc_fence_before();
            bar_sync_named(1, kW3Consumers);
            if (warp == 0) {
                tc_fence_after();
                if (elect_one()) {
                    umma_tile_mn(tm, smem_u32(gi), kW3Slab, smem_u32(xi), kW3Slab, kW3T, idesc, it > 0);
                    umma_commit(&imgfree[b]);
                }
                __syncwarp();
            }
        }
        // the last MMAs on both image buffers
        if (nt >= 1) mbar_wait(&imgfree[(nt - 1) & 1u], ((nt - 1) >> 1) & 1u);
        if (nt >= 2) mbar_wait(&imgfree[nt & 1u], ((nt - 2) >> 1) & 1u);
        tc_fence_after();
        const int64_t plen = (int64_t)128 * 128 + 128;
        float* my = parts + (int64_t)blockIdx.x * plen;
        if (warp < 4) {  // lane = output feature, columns = input features
            const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
            for (int c0 = 0; c0 < 128; c0 += 32) {
                float v[32];
                tmem_ld32(tm + lane_base + c0, v);
                float* dst = my + (int64_t)t * 128 + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + j) = nt > 0 ? make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        // column sums of g: 8 row groups (warps) per column; the staging area is free by now
        float* red = reinterpret_cast<float*>(sm + o5Stage);
        bar_sync_named(1, kW3Consumers);
#pragma unroll
        for (int u = 0; u < 4; ++u) red[w * 128 + c4 * 4 + u] = bsum[u];
        bar_sync_named(1, kW3Consumers);
        if (t < 128) {
            float sacc = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) sacc += red[k * 128 + t];
            my[(int64_t)128 * 128 + t] = sacc;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tm);
}

}  // namespace gmp

using namespace gmp;

extern "C" {

int32_t gmp_linear_wgrad_num_parts(int64_t n) {
    const int64_t nt = ceil_div(n, kLT);
    const int64_t cap = 2 * (int64_t)num_sms();   // two partials per SM: more would only lengthen the reduction
    return (int32_t)(nt < cap ? (nt < 1 ? 1 : nt) : cap);
}

int gmp_linear_wgrad_tc(const float* g, const float* x, int64_t n, int32_t out_dim, int32_t in_dim, float* parts, gmp_stream_t stream) {
    GMP_REQUIRE(g && x && parts, "linear_wgrad_tc: NULL pointer");
    GMP_REQUIRE(out_dim == 128 && (in_dim == 64 || in_dim == 128), "linear_wgrad_tc: built for out = 128, in in {64, 128} (got %d, %d)",
                out_dim, in_dim);
    GMP_REQUIRE(n >= 1, "linear_wgrad_tc: empty input");
    static const bool first_kernel = getenv("GMP_WGRAD_V1") != nullptr;   // A/B timing
    if (in_dim == 128 && !first_kernel) {
        GMP_REQUIRE(((uintptr_t)g & 15) == 0 && ((uintptr_t)x & 15) == 0, "linear_wgrad_tc: 16-byte aligned operands");
        GMP_CUDA(cudaFuncSetAttribute(linear_wgrad_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kW3Smem));
        linear_wgrad_tc3_kernel<<<gmp_linear_wgrad_num_parts(n), kW3Threads, kW3Smem, stream>>>(g, x, n, parts);
        return check_launch("linear_wgrad_tc3_kernel");
    }
    GMP_CUDA(cudaFuncSetAttribute(linear_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLinSmem));
    linear_wgrad_tc_kernel<<<gmp_linear_wgrad_num_parts(n), kLThreads, kLinSmem, stream>>>(g, x, n, in_dim, parts);
    return check_launch("linear_wgrad_tc_kernel");
}

}  // extern "C"
