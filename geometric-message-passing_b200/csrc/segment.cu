// Deterministic, atomics-free segmented reductions over a CSR (torch_scatter replacements) and the
// K0 "gather x multiply -> segmented sum" kernel.  HBM-bound: one warp per destination row, float4
// lanes across the feature dimension, 4 neighbour rows in flight per lane.
#include <cuda_bf16.h>
#include "common.cuh"

namespace gmp {

// out[r,:] = sum_k x[col[k],:] * w[perm ? perm[k] : k, :] with the per-edge factor w stored as bf16 rows (F = 128): one warp per
// row, one lane per four columns, eight edges in flight.  This is dL/dx1 of the CFConv once the forward pass has kept its
// filter values (gmp_schnet_cfconv_fwd_tc2_keep): HBM sees 256 B per edge, the gathered fp32 rows come from L2.
template <bool XBF16>
__global__ void __launch_bounds__(256)
segsum_wbf16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                    const void* __restrict__ xv_, const __nv_bfloat16* __restrict__ w, float* __restrict__ out, int64_t n) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k0 = b; k0 < e; k0 += 8) {
        float4 xv[XBF16 ? 1 : 8];
        uint2 xb[XBF16 ? 8 : 1];
        uint2 wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u;
            if (XBF16) xb[u % (XBF16 ? 8 : 1)] = make_uint2(0u, 0u); else xv[u % (XBF16 ? 1 : 8)] = make_float4(0.f, 0.f, 0.f, 0.f);
            wv[u] = make_uint2(0u, 0u);
            if (k < e) {
                const int64_t srow = __ldg(col + k), wrow = perm ? (int64_t)__ldg(perm + k) : (int64_t)k;
                if (XBF16) xb[u % (XBF16 ? 8 : 1)] = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(xv_) + srow * 128 + lane * 4));
                else xv[u % (XBF16 ? 1 : 8)] = ldg4(static_cast<const float*>(xv_) + srow * 128 + lane * 4);
                wv[u] = __ldg(reinterpret_cast<const uint2*>(w + wrow * 128 + lane * 4));
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {   // fixed order of the additions: deterministic
            float4 xx;
            if (XBF16) {
                const uint2 q = xb[u % (XBF16 ? 8 : 1)];
                xx = make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                                 __uint_as_float(q.y & 0xffff0000u));
            } else {
                xx = xv[u % (XBF16 ? 1 : 8)];
            }
            acc.x = fmaf(xx.x, __uint_as_float(wv[u].x << 16), acc.x);
            acc.y = fmaf(xx.y, __uint_as_float(wv[u].x & 0xffff0000u), acc.y);
            acc.z = fmaf(xx.z, __uint_as_float(wv[u].y << 16), acc.z);
            acc.w = fmaf(xx.w, __uint_as_float(wv[u].y & 0xffff0000u), acc.w);
        }
    }
    *reinterpret_cast<float4*>(out + row * 128 + lane * 4) = acc;
}

template <bool HAS_W, bool GATHER>
__global__ void __launch_bounds__(256)
segsum_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
              const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out, int64_t n, int F,
              int mean) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
    const float scale = (mean && e > b) ? 1.0f / (float)(e - b) : 1.0f;
    for (int f = lane * 4; f < F; f += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int k = b;
        for (; k + 4 <= e; k += 4) {
            float4 xv[4], wv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t srow = GATHER ? (int64_t)__ldg(col + k + u) : (perm ? (int64_t)__ldg(perm + k + u) : (int64_t)(k + u));
                xv[u] = ldg4(x + srow * F + f);
                if (HAS_W) {
                    const int64_t wrow = perm ? (int64_t)__ldg(perm + k + u) : (int64_t)(k + u);
                    wv[u] = ldg4(w + wrow * F + f);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (HAS_W) {
                    acc.x = fmaf(xv[u].x, wv[u].x, acc.x); acc.y = fmaf(xv[u].y, wv[u].y, acc.y);
                    acc.z = fmaf(xv[u].z, wv[u].z, acc.z); acc.w = fmaf(xv[u].w, wv[u].w, acc.w);
                } else {
                    acc.x += xv[u].x; acc.y += xv[u].y; acc.z += xv[u].z; acc.w += xv[u].w;
                }
            }
        }
        for (; k < e; ++k) {
            const int64_t srow = GATHER ? (int64_t)__ldg(col + k) : (perm ? (int64_t)__ldg(perm + k) : (int64_t)k);
            const float4 xv = ldg4(x + srow * F + f);
            if (HAS_W) {
                const int64_t wrow = perm ? (int64_t)__ldg(perm + k) : (int64_t)k;
                const float4 wv = ldg4(w + wrow * F + f);
                acc.x = fmaf(xv.x, wv.x, acc.x); acc.y = fmaf(xv.y, wv.y, acc.y);
                acc.z = fmaf(xv.z, wv.z, acc.z); acc.w = fmaf(xv.w, wv.w, acc.w);
            } else {
                acc.x += xv.x; acc.y += xv.y; acc.z += xv.z; acc.w += xv.w;
            }
        }
        acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
        *reinterpret_cast<float4*>(out + row * F + f) = acc;
    }
}

// narrow features (F not a multiple of 4, e.g. the EGNN coordinate update F = 3): one thread per (row, f)
__global__ void segsum_narrow_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                                     const float* __restrict__ src, float* __restrict__ out, int64_t n, int F, int mean) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * F) return;
    const int64_t row = t / F;
    const int f = (int)(t - row * F);
    const int b = rowptr[row], e = rowptr[row + 1];
    float acc = 0.f;
    for (int k = b; k < e; ++k) acc += src[(int64_t)(perm ? perm[k] : k) * F + f];
    if (mean && e > b) acc /= (float)(e - b);
    out[t] = acc;
}

__global__ void gather_rows_kernel(const int32_t* __restrict__ idx, const float* __restrict__ x, float* __restrict__ out,
                                   int64_t rows, int F4) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * F4) return;
    const int64_t r = t / F4;
    const int f = (int)(t - r * F4);
    reinterpret_cast<float4*>(out)[t] = __ldg(reinterpret_cast<const float4*>(x) + (int64_t)idx[r] * F4 + f);
}

__global__ void gather_rows_scalar_kernel(const int32_t* __restrict__ idx, const float* __restrict__ x,
                                          float* __restrict__ out, int64_t rows, int F) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * F) return;
    const int64_t r = t / F;
    out[t] = x[(int64_t)idx[r] * F + (t - r * F)];
}

// out[i] = sum_p part[p][i]: 32 outputs x 8 groups of partials per block (the outputs are few -- one weight matrix -- and the
// partials many, so the partial index is split over threads too); every group sums its contiguous range of partials in order,
// the eight group sums are added in order: one fixed association, deterministic.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ part, int nparts, int64_t len, float* __restrict__ out) {
    __shared__ float red[8][32];
    const int ix = threadIdx.x & 31, gy = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + ix;
    const int p0 = (int)(((int64_t)nparts * gy) / 8), p1 = (int)(((int64_t)nparts * (gy + 1)) / 8);
    float acc = 0.f;
    if (i < len) {
        int p = p0;
        for (; p + 4 <= p1; p += 4) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldcg(part + (int64_t)(p + u) * len + i);
            acc += (v[0] + v[1]) + (v[2] + v[3]);
        }
        for (; p < p1; ++p) acc += __ldcg(part + (int64_t)p * len + i);
    }
    red[gy][ix] = acc;
    __syncthreads();
    if (gy == 0 && i < len) {
        float s = red[0][ix];
#pragma unroll
        for (int g = 1; g < 8; ++g) s += red[g][ix];
        out[i] = s;
    }
}

// several independent reductions in one launch (blockIdx.y = problem): a backward pass collects the partial buffers of all its
// parameter gradients and reduces them together (SchNet: 25 launches -> 1); same association per problem as the kernel above
constexpr int kReduceBatchMax = 32;
struct ReduceBatch {
    const float* part[kReduceBatchMax];
    float* out[kReduceBatchMax];
    int64_t len[kReduceBatchMax];
    int nparts[kReduceBatchMax];
};
__global__ void __launch_bounds__(256) reduce_partials_batch_kernel(const ReduceBatch b) {
    __shared__ float red[8][32];
    const int m = blockIdx.y;
    const float* __restrict__ part = b.part[m];
    const int64_t len = b.len[m];
    const int nparts = b.nparts[m];
    if ((int64_t)blockIdx.x * 32 >= len) return;
    const int ix = threadIdx.x & 31, gy = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + ix;
    const int p0 = (int)(((int64_t)nparts * gy) / 8), p1 = (int)(((int64_t)nparts * (gy + 1)) / 8);
    float acc = 0.f;
    if (i < len) {
        int p = p0;
        for (; p + 4 <= p1; p += 4) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldcg(part + (int64_t)(p + u) * len + i);
            acc += (v[0] + v[1]) + (v[2] + v[3]);
        }
        for (; p < p1; ++p) acc += __ldcg(part + (int64_t)p * len + i);
    }
    red[gy][ix] = acc;
    __syncthreads();
    if (gy == 0 && i < len) {
        float s_ = red[0][ix];
#pragma unroll
        for (int g = 1; g < 8; ++g) s_ += red[g][ix];
        b.out[m][i] = s_;
    }
}

__global__ void edge_length_kernel(const float* __restrict__ pos, const int64_t* __restrict__ src,
                                   const int64_t* __restrict__ dst, int64_t E, float* __restrict__ dist) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t a = src[e], b = dst[e];
    const float dx = pos[3 * a] - pos[3 * b], dy = pos[3 * a + 1] - pos[3 * b + 1], dz = pos[3 * a + 2] - pos[3 * b + 2];
    dist[e] = sqrtf(dx * dx + dy * dy + dz * dz);
}

// dpos[i] = sum_{e: src=i} g u_e - sum_{e: dst=i} g u_e ; one thread per node, both CSRs walked in order
__global__ void edge_length_bwd_kernel(const float* __restrict__ pos, const int64_t* __restrict__ src,
                                       const int64_t* __restrict__ dst, const float* __restrict__ g,
                                       const int32_t* __restrict__ rp_s, const int32_t* __restrict__ pm_s,
                                       const int32_t* __restrict__ rp_d, const int32_t* __restrict__ pm_d, int64_t n,
                                       float* __restrict__ dpos) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
        const int32_t* rp = pass ? rp_d : rp_s;
        const int32_t* pm = pass ? pm_d : pm_s;
        const float sgn = pass ? -1.f : 1.f;
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int64_t e = pm ? pm[k] : k;
            const int64_t a = src[e], b = dst[e];
            const float dx = pos[3 * a] - pos[3 * b], dy = pos[3 * a + 1] - pos[3 * b + 1], dz = pos[3 * a + 2] - pos[3 * b + 2];
            const float d = sqrtf(dx * dx + dy * dy + dz * dz);
            // torch.norm backward: grad * x / norm, with the subgradient 0 at norm == 0
            const float s = d > 0.f ? sgn * g[e] / d : 0.f;
            ax = fmaf(s, dx, ax); ay = fmaf(s, dy, ay); az = fmaf(s, dz, az);
        }
    }
    dpos[3 * i] = ax; dpos[3 * i + 1] = ay; dpos[3 * i + 2] = az;
}

// TFN / MACE edge prologue (models/tfn.py:171-175): relative vector, length, real spherical harmonics
// (e3nn 'component' normalisation, unit-normalised direction, l <= 2) and Bessel x polynomial-cutoff radial basis.
__global__ void edge_geometry_kernel(const float* __restrict__ pos, const int64_t* __restrict__ src,
                                     const int64_t* __restrict__ dst, int64_t E, int lmax, float r_max, int nb, float p,
                                     float* __restrict__ sh, float* __restrict__ rbf) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t a = src[e], b = dst[e];
    const float vx = pos[3 * a] - pos[3 * b], vy = pos[3 * a + 1] - pos[3 * b + 1], vz = pos[3 * a + 2] - pos[3 * b + 2];
    const float len = sqrtf(vx * vx + vy * vy + vz * vz);
    const float inv = 1.0f / fmaxf(len, 1e-12f);  // F.normalize
    const float x = vx * inv, y = vy * inv, z = vz * inv;
    const int S = (lmax + 1) * (lmax + 1);
    float* o = sh + e * S;
    o[0] = 1.0f;
    if (lmax >= 1) {
        const float s3 = 1.7320508075688772f;
        o[1] = s3 * x; o[2] = s3 * y; o[3] = s3 * z;
    }
    if (lmax >= 2) {
        const float s3 = 1.7320508075688772f, s5 = 2.23606797749979f;
        const float x2z2 = x * x + z * z;
        o[4] = s5 * (s3 * x * z);
        o[5] = s5 * (s3 * x * y);
        o[6] = s5 * (y * y - 0.5f * x2z2);
        o[7] = s5 * (s3 * y * z);
        o[8] = s5 * ((s3 / 2.0f) * (z * z - x * x));
    }
    const float u = len / r_max;
    const float env = 1.0f - ((p + 1.0f) * (p + 2.0f) / 2.0f) * powf(u, p) + p * (p + 2.0f) * powf(u, p + 1.0f) -
                      (p * (p + 1.0f) / 2.0f) * powf(u, p + 2.0f);
    const float cut = len < r_max ? env : 0.0f;
    const float pref = sqrtf(2.0f / r_max);
    for (int k = 0; k < nb; ++k) {
        const float w = (3.14159265358979323846f / r_max) * (float)(k + 1);
        rbf[e * nb + k] = pref * (sinf(w * len) / len) * cut;
    }
}

}  // namespace gmp

using namespace gmp;

// out[r, :] = sum_{k in row r} bf16 rows src[perm ? perm[k] : k, :]   (fp32 accumulation, F = 128: warp per row, 4 columns per lane)
__global__ void __launch_bounds__(256) segsum_bf16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                                                          const uint2* __restrict__ src, float* __restrict__ out, int64_t n) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int64_t b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int64_t k = b;
    for (; k + 4 <= e; k += 4) {  // four rows in flight
        uint2 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t idx = perm ? (int64_t)__ldg(perm + k + j) : k + j;
            v[j] = __ldg(src + idx * 32 + lane);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a0 += __uint_as_float(v[j].x << 16); a1 += __uint_as_float(v[j].x & 0xffff0000u);
            a2 += __uint_as_float(v[j].y << 16); a3 += __uint_as_float(v[j].y & 0xffff0000u);
        }
    }
    for (; k < e; ++k) {
        const int64_t idx = perm ? (int64_t)__ldg(perm + k) : k;
        const uint2 v = __ldg(src + idx * 32 + lane);
        a0 += __uint_as_float(v.x << 16); a1 += __uint_as_float(v.x & 0xffff0000u);
        a2 += __uint_as_float(v.y << 16); a3 += __uint_as_float(v.y & 0xffff0000u);
    }
    *reinterpret_cast<float4*>(out + row * 128 + 4 * lane) = make_float4(a0, a1, a2, a3);
}

extern "C" {

int gmp_segment_sum_bf16_f32(const int32_t* rowptr, const int32_t* perm, const void* src_bf16, float* out, int64_t n, int32_t F,
                             gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && out && n >= 0, "segment_sum_bf16: bad arguments");
    GMP_REQUIRE(F == 128, "segment_sum_bf16: built for rows of 128 bf16 values (got %d)", F);
    if (n == 0) return GMP_OK;
    segsum_bf16_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, stream>>>(rowptr, perm, (const uint2*)src_bf16, out, n);
    return check_launch("segsum_bf16_kernel");
}

int gmp_segment_reduce_f32(const int32_t* rowptr, const int32_t* perm, const float* src, float* out, int64_t n,
                           int32_t F, int32_t mean, gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && out && n >= 0 && F > 0, "segment_reduce: bad arguments");
    if (n == 0) return GMP_OK;
    if (F % 4 == 0) {
        segsum_kernel<false, false><<<(unsigned)ceil_div(n * 32, 256), 256, 0, stream>>>(rowptr, nullptr, perm, src, nullptr, out, n, F, mean);
    } else {
        segsum_narrow_kernel<<<(unsigned)ceil_div(n * F, 256), 256, 0, stream>>>(rowptr, perm, src, out, n, F, mean);
    }
    return check_launch("segment_reduce");
}

int gmp_gather_mul_segsum_f32(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const float* x,
                              const float* w, float* out, int64_t n, int32_t F, gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && col && x && out && n >= 0 && F > 0 && F % 4 == 0, "gather_mul_segsum: bad arguments (F % 4 == 0)");
    if (n == 0) return GMP_OK;
    const unsigned grid = (unsigned)ceil_div(n * 32, 256);
    if (w) segsum_kernel<true, true><<<grid, 256, 0, stream>>>(rowptr, col, perm, x, w, out, n, F, 0);
    else segsum_kernel<false, true><<<grid, 256, 0, stream>>>(rowptr, col, perm, x, nullptr, out, n, F, 0);
    return check_launch("gather_mul_segsum");
}

int gmp_gather_mul_segsum_wbf16(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const void* x, int32_t x_is_bf16,
                                const void* w_bf16, float* out, int64_t n, int32_t F, gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && x && out && n >= 0, "gather_mul_segsum_wbf16: bad arguments");
    GMP_REQUIRE(F == 128, "gather_mul_segsum_wbf16: built for 128 columns (got %d)", F);
    if (n == 0) return GMP_OK;
    if (!col || !w_bf16) {   // only an edgeless graph has no column / factor arrays: every row is an empty sum
        GMP_CUDA(cudaMemsetAsync(out, 0, (size_t)n * 128 * sizeof(float), stream));
        return GMP_OK;
    }
    const unsigned grid = (unsigned)ceil_div(n * 32, 256);
    if (x_is_bf16) segsum_wbf16_kernel<true><<<grid, 256, 0, stream>>>(rowptr, col, perm, x, (const __nv_bfloat16*)w_bf16, out, n);
    else segsum_wbf16_kernel<false><<<grid, 256, 0, stream>>>(rowptr, col, perm, x, (const __nv_bfloat16*)w_bf16, out, n);
    return check_launch("segsum_wbf16_kernel");
}

int gmp_gather_rows_f32(const int32_t* idx, const float* x, float* out, int64_t num_rows, int32_t F, gmp_stream_t stream) {
    GMP_REQUIRE(num_rows == 0 || (idx && x && out && F > 0), "gather_rows: bad arguments");
    if (num_rows == 0) return GMP_OK;
    if (F % 4 == 0) gather_rows_kernel<<<(unsigned)ceil_div(num_rows * (F / 4), 256), 256, 0, stream>>>(idx, x, out, num_rows, F / 4);
    else gather_rows_scalar_kernel<<<(unsigned)ceil_div(num_rows * F, 256), 256, 0, stream>>>(idx, x, out, num_rows, F);
    return check_launch("gather_rows");
}

int gmp_reduce_partials_f32(const float* part, int32_t nparts, int64_t len, float* out, gmp_stream_t stream) {
    GMP_REQUIRE(part && out && nparts >= 1 && len >= 0, "reduce_partials: bad arguments");
    if (len == 0) return GMP_OK;
    reduce_partials_kernel<<<(unsigned)ceil_div(len, 32), 256, 0, stream>>>(part, nparts, len, out);
    return check_launch("reduce_partials");
}

int gmp_reduce_partials_batch_f32(const float* const* parts, const int32_t* nparts, const int64_t* lens, float* const* outs, int32_t count,
                                  gmp_stream_t stream) {
    GMP_REQUIRE(count >= 0 && (count == 0 || (parts && nparts && lens && outs)), "reduce_partials_batch: bad arguments");
    for (int base = 0; base < count; base += kReduceBatchMax) {
        ReduceBatch b;
        const int m = count - base < kReduceBatchMax ? count - base : kReduceBatchMax;
        int64_t maxlen = 0;
        for (int i = 0; i < m; ++i) {
            GMP_REQUIRE(parts[base + i] && outs[base + i] && nparts[base + i] >= 1 && lens[base + i] >= 0, "reduce_partials_batch: problem %d", base + i);
            b.part[i] = parts[base + i]; b.out[i] = outs[base + i]; b.len[i] = lens[base + i]; b.nparts[i] = nparts[base + i];
            if (lens[base + i] > maxlen) maxlen = lens[base + i];
        }
        if (maxlen == 0) continue;
        reduce_partials_batch_kernel<<<dim3((unsigned)ceil_div(maxlen, 32), (unsigned)m), 256, 0, stream>>>(b);
        if (int rc = check_launch("reduce_partials_batch")) return rc;
    }
    return GMP_OK;
}

int gmp_edge_length_fwd(const float* pos, const int64_t* src, const int64_t* dst, int64_t num_edges, float* dist,
                        gmp_stream_t stream) {
    GMP_REQUIRE(num_edges == 0 || (pos && src && dst && dist), "edge_length_fwd: bad arguments");
    if (num_edges == 0) return GMP_OK;
    edge_length_kernel<<<(unsigned)ceil_div(num_edges, 256), 256, 0, stream>>>(pos, src, dst, num_edges, dist);
    return check_launch("edge_length_kernel");
}

int gmp_edge_geometry_fwd(const float* pos, const int64_t* src, const int64_t* dst, int64_t num_edges, int32_t max_ell,
                          float r_max, int32_t num_bessel, float poly_p, float* edge_sh, float* edge_feat,
                          gmp_stream_t stream) {
    GMP_REQUIRE(num_edges == 0 || (pos && src && dst && edge_sh && edge_feat), "edge_geometry_fwd: bad arguments");
    GMP_REQUIRE(max_ell >= 0 && max_ell <= 2, "edge_geometry_fwd: max_ell must be <= 2 (got %d)", max_ell);
    GMP_REQUIRE(num_bessel >= 1 && r_max > 0.f, "edge_geometry_fwd: bad radial basis");
    if (num_edges == 0) return GMP_OK;
    edge_geometry_kernel<<<(unsigned)ceil_div(num_edges, 256), 256, 0, stream>>>(pos, src, dst, num_edges, max_ell, r_max,
                                                                               num_bessel, poly_p, edge_sh, edge_feat);
    return check_launch("edge_geometry_kernel");
}

int gmp_edge_length_bwd(const float* pos, const int64_t* src, const int64_t* dst, const float* g_dist,
                        const int32_t* rowptr_s, const int32_t* perm_s, const int32_t* rowptr_d, const int32_t* perm_d,
                        int64_t n, float* dpos, gmp_stream_t stream) {
    GMP_REQUIRE(pos && src && dst && g_dist && rowptr_s && rowptr_d && dpos, "edge_length_bwd: bad arguments");
    if (n == 0) return GMP_OK;
    edge_length_bwd_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(pos, src, dst, g_dist, rowptr_s, perm_s, rowptr_d, perm_d, n, dpos);
    return check_launch("edge_length_bwd_kernel");
}

}  // extern "C"
