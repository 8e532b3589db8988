// Parameter gradients of a node-side nn.Linear (y = x W^T + b) in the GMP_BF16_TC mode:
//     dW[out, in] = sum_n g[n, out] * x[n, in],     db[out] = sum_n g[n, out]
// for out = 128, in in {64, 128}: a reduction over all N nodes with a tiny output, which cuBLAS runs on a handful of
// CTAs (the K = N dimension is not split) and ATen's column sum at a fraction of the memory bandwidth.  Here the rows
// are split over all SMs; each CTA streams its 128-row tiles of g and x once (fp32 -> bf16 on the fly, written as
// MN-major UMMA operand images), accumulates g^T x in tensor memory over its whole row range, adds the column sums of
// g from registers, and writes one partial; gmp_reduce_partials_f32 sums the partials in a fixed order.
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
    GMP_CUDA(cudaFuncSetAttribute(linear_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLinSmem));
    linear_wgrad_tc_kernel<<<gmp_linear_wgrad_num_parts(n), kLThreads, kLinSmem, stream>>>(g, x, n, in_dim, parts);
    return check_launch("linear_wgrad_tc_kernel");
}

}  // extern "C"
