// MACE symmetric contraction (models/mace_modules/symmetric_contraction.py:169-185, Eq. 10-11 of the MACE
// paper), node side, fp32.  For every node b and channel c the reference evaluates, through three chained
// einsums, a polynomial of degree <= 3 in the D = 9 components x[b,c,:]:
//     out[b,c,K] = sum_nu sum_eta w_nu[eta,c] sum_{i1..i_nu} U_nu[K, i1..i_nu, eta] x_i1 ... x_i_nu .
// Here the host folds U and w into per-channel coefficients over the symmetric monomial basis
// (coef[c][K][m], m over the 9 + 45 + 165 monomials; a tiny differentiable einsum), and the kernel evaluates
//     out[b,c,K] = sum_m coef[c][K][m] * x_{i1(m)} x_{i2(m)} x_{i3(m)}          (x_D := 1 pads lower degrees)
// with the channel's coefficients broadcast from shared memory.  The backward produces dx and dcoef
// (per-block partial sums over nodes, reduced deterministically); dw follows from dcoef by autograd.
#include "common.cuh"

namespace gmp {

constexpr int kSCNodes = 128;  // nodes per tile
constexpr int kSCMaxK = 16;    // output components per channel (1 + 3 + 5 = 9 at l_max = 2)
constexpr int kSCMaxD = 15;    // input components per channel (+1 slot for the constant 1)

struct SCArgs {
    const float* x;         // [N, C, D]
    const float* coef;      // [C, K, M]
    const int32_t* mono;    // [M, 3] component indices in [0, D]  (D = the constant 1)
    const int32_t* out_map; // [K, 3]: (block offset in components, block dim d, local k) -> out[b][C*off + c*d + k]
    int64_t N;
    int32_t C, D, K, M, out_len;
};

__global__ void __launch_bounds__(kSCNodes) symcontract_fwd_kernel(SCArgs a, float* __restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    float* coef = sm;                               // [K][M]
    float* xs = coef + a.K * a.M;                   // [128][D+1]   (stride D+1)
    int* mono = reinterpret_cast<int*>(xs + kSCNodes * (kSCMaxD + 1));  // [M][3]
    const int c = blockIdx.y, t = threadIdx.x;
    const int DS = (a.D + 1) | 1;  // odd stride: conflict-free per-thread rows
    for (int i = t; i < a.K * a.M; i += blockDim.x) coef[i] = __ldg(a.coef + (int64_t)c * a.K * a.M + i);
    for (int i = t; i < a.M * 3; i += blockDim.x) mono[i] = __ldg(a.mono + i);
    for (int64_t tile = blockIdx.x; tile * kSCNodes < a.N; tile += gridDim.x) {
        const int64_t b = tile * kSCNodes + t;
        __syncthreads();
        if (b < a.N) {
            const float* xr = a.x + (b * a.C + c) * a.D;
            for (int i = 0; i < a.D; ++i) xs[t * DS + i] = __ldg(xr + i);
        } else {
            for (int i = 0; i < a.D; ++i) xs[t * DS + i] = 0.f;
        }
        xs[t * DS + a.D] = 1.0f;
        __syncthreads();
        float acc[kSCMaxK];
#pragma unroll
        for (int k = 0; k < kSCMaxK; ++k) acc[k] = 0.f;
        const float* xme = xs + t * DS;
        for (int m = 0; m < a.M; ++m) {
            const float v = xme[mono[3 * m]] * xme[mono[3 * m + 1]] * xme[mono[3 * m + 2]];
#pragma unroll
            for (int k = 0; k < kSCMaxK; ++k)
                if (k < a.K) acc[k] = fmaf(coef[k * a.M + m], v, acc[k]);
        }
        if (b < a.N) {
#pragma unroll
            for (int k = 0; k < kSCMaxK; ++k)
                if (k < a.K) {
                    const int off = __ldg(a.out_map + 3 * k), d = __ldg(a.out_map + 3 * k + 1), kl = __ldg(a.out_map + 3 * k + 2);
                    out[b * a.out_len + (int64_t)a.C * off + c * d + kl] = acc[k];
                }
        }
    }
}

// grid = (node chunks, C).  Per tile: thread t < 128 owns node t (dx, monomial row); then the 256 threads own
// monomial columns and accumulate dcoef[K][m] over the tile's nodes.
__global__ void __launch_bounds__(256) symcontract_bwd_kernel(SCArgs a, const float* __restrict__ g, float* __restrict__ dx,
                                                              float* __restrict__ dcoef_parts) {
    extern __shared__ __align__(16) float sm[];
    const int DS = (a.D + 1) | 1;
    const int LM = (a.M + 2) | 1;                             // odd monomial-tile row stride >= M + 1
    float* coef = sm;                                          // [K][M]
    float* Ms = coef + a.K * a.M;                              // [128][LM]
    float* xs = Ms + kSCNodes * LM;                            // [128][DS]
    float* dxs = xs + kSCNodes * (kSCMaxD + 1);                // [128][DS]
    float* gs = dxs + kSCNodes * (kSCMaxD + 1);                // [128][K]
    int* mono = reinterpret_cast<int*>(gs + kSCNodes * kSCMaxK);
    const int c = blockIdx.y, t = threadIdx.x;
    for (int i = t; i < a.K * a.M; i += blockDim.x) coef[i] = __ldg(a.coef + (int64_t)c * a.K * a.M + i);
    for (int i = t; i < a.M * 3; i += blockDim.x) mono[i] = __ldg(a.mono + i);
    float dacc[kSCMaxK];   // dcoef[k][m = t] partial (threads < M)
#pragma unroll
    for (int k = 0; k < kSCMaxK; ++k) dacc[k] = 0.f;
    float dacc2[kSCMaxK];  // second monomial column for M > 256 is not needed (M <= 219); kept zero
    (void)dacc2;

    for (int64_t tile = blockIdx.x; tile * kSCNodes < a.N; tile += gridDim.x) {
        __syncthreads();
        if (t < kSCNodes) {
            const int64_t b = tile * kSCNodes + t;
            if (b < a.N) {
                const float* xr = a.x + (b * a.C + c) * a.D;
                for (int i = 0; i < a.D; ++i) xs[t * DS + i] = __ldg(xr + i);
                for (int k = 0; k < a.K; ++k) {
                    const int off = __ldg(a.out_map + 3 * k), d = __ldg(a.out_map + 3 * k + 1), kl = __ldg(a.out_map + 3 * k + 2);
                    gs[t * kSCMaxK + k] = __ldg(g + b * a.out_len + (int64_t)a.C * off + c * d + kl);
                }
            } else {
                for (int i = 0; i < a.D; ++i) xs[t * DS + i] = 0.f;
                for (int k = 0; k < a.K; ++k) gs[t * kSCMaxK + k] = 0.f;
            }
            xs[t * DS + a.D] = 1.0f;
            for (int i = 0; i <= a.D; ++i) dxs[t * DS + i] = 0.f;
            // monomials, q_m = sum_K g_K coef[K][m], dx via the product rule
            const float* xme = xs + t * DS;
            float* dme = dxs + t * DS;
            float gk[kSCMaxK];
#pragma unroll
            for (int k = 0; k < kSCMaxK; ++k) gk[k] = k < a.K ? gs[t * kSCMaxK + k] : 0.f;
            for (int m = 0; m < a.M; ++m) {
                const int i1 = mono[3 * m], i2 = mono[3 * m + 1], i3 = mono[3 * m + 2];
                const float x1 = xme[i1], x2 = xme[i2], x3 = xme[i3];
                Ms[t * LM + m] = x1 * x2 * x3;
                float q = 0.f;
#pragma unroll
                for (int k = 0; k < kSCMaxK; ++k)
                    if (k < a.K) q = fmaf(gk[k], coef[k * a.M + m], q);
                dme[i1] = fmaf(q, x2 * x3, dme[i1]);
                dme[i2] = fmaf(q, x1 * x3, dme[i2]);
                dme[i3] = fmaf(q, x1 * x2, dme[i3]);
            }
            if (b < a.N) {
                float* dr = dx + (b * a.C + c) * a.D;
                for (int i = 0; i < a.D; ++i) dr[i] = dme[i];
            }
        }
        __syncthreads();
        if (t < a.M) {
            const int64_t nb = min((int64_t)kSCNodes, a.N - tile * kSCNodes);
            for (int n = 0; n < nb; ++n) {
                const float mv = Ms[n * LM + t];
#pragma unroll
                for (int k = 0; k < kSCMaxK; ++k)
                    if (k < a.K) dacc[k] = fmaf(gs[n * kSCMaxK + k], mv, dacc[k]);
            }
        }
    }
    if (t < a.M) {
        float* my = dcoef_parts + ((int64_t)blockIdx.x * a.C + c) * a.K * a.M;
#pragma unroll
        for (int k = 0; k < kSCMaxK; ++k)
            if (k < a.K) my[k * a.M + t] = dacc[k];
    }
}

static int sc_check(int64_t N, int32_t C, int32_t D, int32_t K, int32_t M) {
    GMP_REQUIRE(N >= 0 && C >= 1 && C <= 65535, "symcontract: bad N / C");
    GMP_REQUIRE(D >= 1 && D <= kSCMaxD && K >= 1 && K <= kSCMaxK && M >= 1 && M <= 256,
                "symcontract: need D <= %d, K <= %d, M <= 256 (got D=%d K=%d M=%d)", kSCMaxD, kSCMaxK, D, K, M);
    return GMP_OK;
}

}  // namespace gmp

using namespace gmp;

extern "C" {

int32_t gmp_symcontract_bwd_num_parts(int64_t num_nodes) {
    const int64_t tiles = ceil_div(num_nodes > 0 ? num_nodes : 1, kSCNodes);
    return (int32_t)(tiles < 8 ? tiles : 8);
}

int gmp_symcontract_fwd(const float* x, const float* coef, const int32_t* mono, const int32_t* out_map, int64_t num_nodes,
                        int32_t C, int32_t D, int32_t K, int32_t M, float* out, int32_t out_len, gmp_stream_t stream) {
    if (int rc = sc_check(num_nodes, C, D, K, M)) return rc;
    GMP_REQUIRE(coef && mono && out_map && (num_nodes == 0 || (x && out)), "symcontract_fwd: NULL pointer");
    if (num_nodes == 0) return GMP_OK;
    SCArgs a{x, coef, mono, out_map, num_nodes, C, D, K, M, out_len};
    const size_t smem = ((size_t)K * M + kSCNodes * (kSCMaxD + 1) + 3 * M) * sizeof(float);
    GMP_CUDA(cudaFuncSetAttribute(symcontract_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = ceil_div(num_nodes, kSCNodes);
    dim3 grid((unsigned)(tiles < 64 ? tiles : 64), (unsigned)C);
    symcontract_fwd_kernel<<<grid, kSCNodes, smem, stream>>>(a, out);
    return check_launch("symcontract_fwd_kernel");
}

int gmp_symcontract_bwd(const float* x, const float* coef, const int32_t* mono, const int32_t* out_map, int64_t num_nodes,
                        int32_t C, int32_t D, int32_t K, int32_t M, const float* g_out, int32_t out_len, float* dx,
                        float* dcoef_parts, gmp_stream_t stream) {
    if (int rc = sc_check(num_nodes, C, D, K, M)) return rc;
    GMP_REQUIRE(coef && mono && out_map && dcoef_parts && (num_nodes == 0 || (x && g_out && dx)), "symcontract_bwd: NULL pointer");
    SCArgs a{x, coef, mono, out_map, num_nodes, C, D, K, M, out_len};
    const size_t smem = ((size_t)K * M + (size_t)kSCNodes * (M + 3) + 2 * kSCNodes * (kSCMaxD + 1) + kSCNodes * kSCMaxK + 3 * M) * sizeof(float);
    GMP_CUDA(cudaFuncSetAttribute(symcontract_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)gmp_symcontract_bwd_num_parts(num_nodes), (unsigned)C);
    symcontract_bwd_kernel<<<grid, 256, smem, stream>>>(a, g_out, dx, dcoef_parts);
    return check_launch("symcontract_bwd_kernel");
}

}  // extern "C"
