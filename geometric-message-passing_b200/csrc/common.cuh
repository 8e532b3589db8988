// Shared device/host helpers for the gmp_b200 CUDA core (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gmp_b200.h"

namespace gmp {

// ---------------------------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns 0 or a negative gmp_status; the message of the
// last failure is kept per thread for gmp_last_error().
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define GMP_REQUIRE(cond, ...)                \
    do {                                      \
        if (!(cond)) {                        \
            ::gmp::set_error(__VA_ARGS__);    \
            return GMP_ERR_INVALID_ARGUMENT;  \
        }                                     \
    } while (0)

#define GMP_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t _e = (call);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ::gmp::set_error("%s failed: %s", #call, cudaGetErrorString(_e));           \
            return GMP_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// gathered bf16 row j of a destination-partitioned operand (include/gmp_b200.h, gmp_peer_rows): own rows local, halo rows in
// the neighbouring ranks' memory.  n_own == 0: not partitioned, every row is local.
struct PeerRows {
    const uint8_t* left;
    const uint8_t* right;
    int64_t n_left, n_own;
};
__device__ __forceinline__ const uint8_t* peer_row(const void* own, const PeerRows& p, int64_t j, int row_bytes) {
    if (p.n_own == 0) return reinterpret_cast<const uint8_t*>(own) + j * row_bytes;
    if (j < p.n_left) return p.left + j * row_bytes;
    j -= p.n_left;
    if (j < p.n_own) return reinterpret_cast<const uint8_t*>(own) + j * row_bytes;
    return p.right + (j - p.n_own) * row_bytes;
}
inline PeerRows make_peer_rows(const gmp_peer_rows* p) {
    PeerRows r{nullptr, nullptr, 0, 0};
    if (p) { r.left = (const uint8_t*)p->left; r.right = (const uint8_t*)p->right; r.n_left = p->n_left; r.n_own = p->n_own; }
    return r;
}

// first row r in [0, n] with rowptr[r] >= target  (rowptr non-decreasing, rowptr[n] = E)
__device__ __forceinline__ int lower_bound_row(const int32_t* __restrict__ rowptr, int n, int64_t target) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(rowptr + mid) < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// shifted softplus and its derivative, computed the way ATen's softplus does (threshold 20)
__device__ __forceinline__ float ssp(float x) {
    const float sp = (x > 20.f) ? x : log1pf(expf(x));
    return sp - 0.6931471805599453f;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
// 64-row SIMT tile GEMMs for 256-thread CTAs.  All operands live in shared memory, fp32.
//   thread (tx = tid & 15, ty = tid >> 4) owns rows ty*RPT + i (i < RPT; tile = 16*RPT rows) and
//   CPT = NCOL/16 columns.
//   K must be a multiple of 4 and every row start 16-byte aligned.
// ---------------------------------------------------------------------------------------------
template <int NCOL, int RPT = 4>
struct Frag {
    static constexpr int CPT = NCOL / 16;
    static constexpr int ROWS = 16 * RPT;  // rows of the tile
    float v[RPT][CPT];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
            for (int j = 0; j < CPT; ++j) v[i][j] = 0.f;
    }
};

// column owned by slot j:  NT layout -> tx + 16 j ;  NN layout -> 4*tx + (j&3) + 64*(j>>2)

// C[r][c] += sum_k A[r][k] * B[c][k]          (B is [NCOL][K], K contiguous: nn.Linear weight layout)
template <int NCOL, int RPT>
__device__ __forceinline__ void gemm_nt(Frag<NCOL, RPT>& acc, const float* __restrict__ A, int lda,
                                        const float* __restrict__ B, int ldb, int K) {
    constexpr int CPT = Frag<NCOL, RPT>::CPT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* a0 = A + (ty * RPT) * lda;
    for (int k = 0; k < K; k += 4) {
        float4 a[RPT], b[CPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * lda + k);
#pragma unroll
        for (int j = 0; j < CPT; ++j) b[j] = *reinterpret_cast<const float4*>(B + (tx + 16 * j) * ldb + k);
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                acc.v[i][j] = fmaf(a[i].x, b[j].x, acc.v[i][j]);
                acc.v[i][j] = fmaf(a[i].y, b[j].y, acc.v[i][j]);
                acc.v[i][j] = fmaf(a[i].z, b[j].z, acc.v[i][j]);
                acc.v[i][j] = fmaf(a[i].w, b[j].w, acc.v[i][j]);
            }
    }
}

// C[r][c] += sum_k A[r][k] * B[k][c]          (B is [K][NCOL], NCOL contiguous)
template <int NCOL, int RPT>
__device__ __forceinline__ void gemm_nn(Frag<NCOL, RPT>& acc, const float* __restrict__ A, int lda,
                                        const float* __restrict__ B, int ldb, int K) {
    constexpr int CPT = Frag<NCOL, RPT>::CPT;
    static_assert(CPT % 4 == 0, "gemm_nn needs NCOL multiple of 64");
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* a0 = A + (ty * RPT) * lda;
    for (int k = 0; k < K; k += 4) {
        float4 a[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * lda + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float4 b[CPT / 4];
#pragma unroll
            for (int q = 0; q < CPT / 4; ++q)
                b[q] = *reinterpret_cast<const float4*>(B + (k + kk) * ldb + 4 * tx + 64 * q);
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
                for (int q = 0; q < CPT / 4; ++q) {
                    acc.v[i][4 * q + 0] = fmaf(av, b[q].x, acc.v[i][4 * q + 0]);
                    acc.v[i][4 * q + 1] = fmaf(av, b[q].y, acc.v[i][4 * q + 1]);
                    acc.v[i][4 * q + 2] = fmaf(av, b[q].z, acc.v[i][4 * q + 2]);
                    acc.v[i][4 * q + 3] = fmaf(av, b[q].w, acc.v[i][4 * q + 3]);
                }
            }
        }
    }
}

template <int NCOL, int RPT>
__device__ __forceinline__ void store_nt(const Frag<NCOL, RPT>& acc, float* __restrict__ C, int ldc) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < Frag<NCOL, RPT>::CPT; ++j) C[(ty * RPT + i) * ldc + tx + 16 * j] = acc.v[i][j];
}

template <int NCOL, int RPT>
__device__ __forceinline__ void store_nn(const Frag<NCOL, RPT>& acc, float* __restrict__ C, int ldc) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int q = 0; q < Frag<NCOL, RPT>::CPT / 4; ++q)
            *reinterpret_cast<float4*>(C + (ty * RPT + i) * ldc + 4 * tx + 64 * q) =
                make_float4(acc.v[i][4 * q], acc.v[i][4 * q + 1], acc.v[i][4 * q + 2], acc.v[i][4 * q + 3]);
}

// Weight-gradient tile:  W[m][n] += sum_{r < rows} P[r][m] * Q[r][n]   with M x N = (16*MT) x (16*NT) outputs,
// thread (tx, ty) owns m = ty + 16 i, n = tx + 16 j.  P, Q in shared memory.
template <int MT, int NT>
__device__ __forceinline__ void wgrad_tn(float (&acc)[MT][NT], const float* __restrict__ P, int ldp,
                                         const float* __restrict__ Q, int ldq, int rows) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int r = 0; r < rows; ++r) {
        float p[MT], q[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i) p[i] = P[r * ldp + ty + 16 * i];
#pragma unroll
        for (int j = 0; j < NT; ++j) q[j] = Q[r * ldq + tx + 16 * j];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(p[i], q[j], acc[i][j]);
    }
}

}  // namespace gmp
