// Halo rows of the destination-partitioned graph pulled straight out of the owners' memory: every
// rank stages the boundary rows its neighbours need in a symmetric (peer-mapped) buffer, and after a device-side barrier
// each rank's own kernel reads the peers' rows over NVLink (plain global loads through the peer mapping) into its local
// [left halo | owned | right halo] array -- one launch instead of ncclSend / ncclRecv pairs plus a concatenation.  The reverse
// direction (halo gradients back to their owners) is the same kernel with add = 1, one launch per peer in rank order, so the
// sums are deterministic.  No torch types: segments are raw (src, dst, bytes) triples.
#include "common.cuh"

namespace gmp {
namespace {

constexpr int kHaloMaxSegs = 16;
struct HaloSegs {
    const void* src[kHaloMaxSegs];
    void* dst[kHaloMaxSegs];
    int64_t bytes[kHaloMaxSegs];
    int n, add, vec16;
};

__global__ void __launch_bounds__(256) halo_pull_kernel(const HaloSegs s) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    for (int k = 0; k < s.n; ++k) {
        if (s.vec16) {
            const uint4* src = reinterpret_cast<const uint4*>(s.src[k]);
            uint4* dst = reinterpret_cast<uint4*>(s.dst[k]);
            const int64_t cnt = s.bytes[k] >> 4;
            for (int64_t i = tid; i < cnt; i += stride) {
                const uint4 v = __ldcg(src + i);      // L2 only: the peer rewrites these lines every exchange
                if (s.add) {
                    float4 d = *reinterpret_cast<float4*>(dst + i);
                    d.x += __uint_as_float(v.x); d.y += __uint_as_float(v.y); d.z += __uint_as_float(v.z); d.w += __uint_as_float(v.w);
                    *reinterpret_cast<float4*>(dst + i) = d;
                } else {
                    dst[i] = v;
                }
            }
        } else {
            const float* src = reinterpret_cast<const float*>(s.src[k]);
            float* dst = reinterpret_cast<float*>(s.dst[k]);
            const int64_t cnt = s.bytes[k] >> 2;
            for (int64_t i = tid; i < cnt; i += stride) {
                const float v = __ldcg(src + i);
                dst[i] = s.add ? dst[i] + v : v;
            }
        }
    }
}

}  // namespace
}  // namespace gmp

using namespace gmp;

extern "C" {

int gmp_halo_pull(const void* const* src, void* const* dst, const int64_t* bytes, int32_t num_segments, int32_t add_f32, gmp_stream_t stream) {
    GMP_REQUIRE(num_segments >= 0 && num_segments <= kHaloMaxSegs, "halo_pull: at most %d segments per launch (got %d)", kHaloMaxSegs, num_segments);
    if (num_segments == 0) return GMP_OK;
    GMP_REQUIRE(src && dst && bytes, "halo_pull: NULL arrays");
    HaloSegs s;
    s.n = 0;
    s.add = add_f32 ? 1 : 0;
    s.vec16 = 1;
    int64_t total = 0;
    for (int k = 0; k < num_segments; ++k) {
        if (bytes[k] <= 0) continue;
        GMP_REQUIRE(src[k] && dst[k] && bytes[k] % 4 == 0, "halo_pull: segment %d: NULL pointer or a size that is not a multiple of 4", k);
        s.src[s.n] = src[k];
        s.dst[s.n] = dst[k];
        s.bytes[s.n] = bytes[k];
        if (((uintptr_t)src[k] | (uintptr_t)dst[k] | (uintptr_t)bytes[k]) & 15) s.vec16 = 0;
        total += bytes[k];
        ++s.n;
    }
    if (s.n == 0) return GMP_OK;
    const int64_t work = total / (s.vec16 ? 16 : 4);
    int64_t blocks = ceil_div(work, 256 * 4);
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    halo_pull_kernel<<<(unsigned)blocks, 256, 0, stream>>>(s);
    return check_launch("halo_pull_kernel");
}

}  // extern "C"
