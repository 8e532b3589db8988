// EGNN edge path, fused (fp32 strict): gather -> geometry -> 3-stage edge MLP with LayerNorm -> segmented
// sum (features) / mean (coordinates).  Reference: models/layers/egnn_layer.py:62-80.
//
// The first Linear of mlp_msg acts on cat[h_i, h_j, dist]; it is split as
//     pre1_e = P[i] + Q[j] + dist_e * wd,   P = h W0[:, :d]^T + b0,  Q = h W0[:, d:2d]^T,  wd = W0[:, 2d]
// so the K = 2d+1 edge GEMM becomes two node GEMMs (done by the caller) and the kernel starts from P, Q.
// Per edge: a1 = act(LN1(pre1)); m = act(LN2(a1 W1^T + b1)); a3 = act(LN3(m W2^T + b2)); s = a3.w3 + b3;
//           message m, coordinate shift (pos_i - pos_j) * s.
//
// Work decomposition as in schnet.cu (ranges of whole CSR rows, persistent grid, 32-edge tiles so that the
// three normalised activations needed by the backward stay in shared memory next to both weight matrices).
// The backward recomputes the forward per tile and runs twice: over the dst-sorted CSR (dP, dpos_i, weight
// gradients) and over the src-sorted CSR (dQ, dpos_j); no per-edge tensor is stored, no atomics.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace gmp {

constexpr int kET = 32;                 // edges per tile
constexpr int kERangeMax = 1024;        // edges per work range on large graphs; small graphs get shorter ranges so that
                                        // the work still spreads over the SMs (the reference's own workloads are a few
                                        // hundred edges: one 1024-edge range would run them on a single SM)
constexpr int kRPT = 2;                 // rows per thread in the tile GEMMs (16 * 2 = 32 rows)

struct EgnnArgs {
    const int32_t *rowptr, *col, *deg_rowptr;
    int64_t n, E;
    const float *P, *Q, *pos;
    const float *wd, *g1, *be1, *w1, *b1, *g2, *be2, *w2, *b2, *g3, *be3, *w3, *b3;
    int act, aggr_mean, nranges, erange;
    float eps;
};

template <int F>
struct EgnnSmem {
    static constexpr int LD = F + 4;
    static constexpr int oW1 = 0;                  // [F][LD]
    static constexpr int oW2 = oW1 + F * LD;       // [F][LD]
    static constexpr int oVec = oW2 + F * LD;      // wd g1 be1 b1 g2 be2 b2 g3 be3 w3  (10 x F)
    static constexpr int oT0 = oVec + 10 * F;      // xhat1            [32][LD]
    static constexpr int oT1 = oT0 + kET * LD;     // a1 -> m -> work  [32][LD]
    static constexpr int oT2 = oT1 + kET * LD;     // xhat2            [32][LD]
    static constexpr int oT3 = oT2 + kET * LD;     // xhat3 -> work    [32][LD]
    static constexpr int oSc = oT3 + kET * LD;     // per-edge scalars: 13 x 32 (enum S_*)
    static constexpr int oI = oSc + 13 * kET;      // ints: i[32], j[32]
    static constexpr int oRed = oI + 2 * kET;      // end-of-kernel cross-warp reduction: 8 x F
    static constexpr int total = oRed + 8 * F;
};
enum { V_WD = 0, V_G1, V_BE1, V_B1, V_G2, V_BE2, V_B2, V_G3, V_BE3, V_W3 };
enum { S_DX = 0, S_DY, S_DZ, S_DIST, S_S, S_R1, S_R2, S_R3, S_GS, S_SCALE, S_GX, S_GY, S_GZ, S_COUNT };

template <int ACT> __device__ __forceinline__ float actf(float y) {
    if (ACT == 0) return fmaxf(y, 0.f);
    return y * sigmoidf_(y);
}
template <int ACT> __device__ __forceinline__ float dactf(float y) {
    if (ACT == 0) return y > 0.f ? 1.f : 0.f;
    const float s = sigmoidf_(y);
    return s * (1.f + y * (1.f - s));
}

template <int F>
__device__ __forceinline__ void egnn_load_weights(float* sm, const EgnnArgs& a) {
    using S = EgnnSmem<F>;
    for (int i = threadIdx.x; i < F * S::LD; i += blockDim.x) {
        const int r = i / S::LD, c = i - r * S::LD;
        sm[S::oW1 + i] = c < F ? __ldg(a.w1 + r * F + c) : 0.f;
        sm[S::oW2 + i] = c < F ? __ldg(a.w2 + r * F + c) : 0.f;
    }
    const float* vecs[10] = {a.wd, a.g1, a.be1, a.b1, a.g2, a.be2, a.b2, a.g3, a.be3, a.w3};
    for (int i = threadIdx.x; i < 10 * F; i += blockDim.x) sm[S::oVec + i] = __ldg(vecs[i / F] + (i % F));
}

// LayerNorm of one row held as CPL values per lane (columns lane + 32 c): returns rstd, leaves xhat in v
template <int F>
__device__ __forceinline__ float ln_row(float (&v)[F / 32], float eps) {
    constexpr int CPL = F / 32;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) s += v[c];
    const float mean = warp_sum(s) * (1.f / F);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        v[c] -= mean;
        q = fmaf(v[c], v[c], q);
    }
    const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / F) + eps);
#pragma unroll
    for (int c = 0; c < CPL; ++c) v[c] *= rstd;
    return rstd;
}

// Forward of one tile.  On exit: T0 = xhat1, T2 = xhat2, T1 = m, T3 = xhat3, scalars s / rstd1..3 / geometry.
// SRC_PASS: CSR rows are the source nodes j and col holds the destination i.
template <int F, int ACT, bool SRC_PASS>
__device__ __forceinline__ void egnn_tile_forward(float* sm, const EgnnArgs& a, int r0, int r1, int64_t e0, int cnt) {
    using S = EgnnSmem<F>;
    constexpr int LD = S::LD, CPL = F / 32;
    int* smi = reinterpret_cast<int*>(sm + S::oI);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < kET) {
        const int t = threadIdx.x;
        int i = 0, j = 0;
        float dx = 0.f, dy = 0.f, dz = 0.f, dist = 0.f;
        if (t < cnt) {
            const int64_t e = e0 + t;
            int lo = r0, hi = r1;  // CSR row of edge e
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if ((int64_t)__ldg(a.rowptr + mid) <= e) lo = mid; else hi = mid;
            }
            const int other = __ldg(a.col + e);
            i = SRC_PASS ? other : lo;
            j = SRC_PASS ? lo : other;
            dx = __ldg(a.pos + 3 * (int64_t)i) - __ldg(a.pos + 3 * (int64_t)j);
            dy = __ldg(a.pos + 3 * (int64_t)i + 1) - __ldg(a.pos + 3 * (int64_t)j + 1);
            dz = __ldg(a.pos + 3 * (int64_t)i + 2) - __ldg(a.pos + 3 * (int64_t)j + 2);
            dist = sqrtf(dx * dx + dy * dy + dz * dz);
        }
        smi[t] = i;
        smi[kET + t] = j;
        sm[S::oSc + S_DX * kET + t] = dx;
        sm[S::oSc + S_DY * kET + t] = dy;
        sm[S::oSc + S_DZ * kET + t] = dz;
        sm[S::oSc + S_DIST * kET + t] = dist;
    }
    __syncthreads();
    {   // gather P[i] -> T0, Q[j] -> T1
        constexpr int CH = F / 4;
        for (int x = threadIdx.x; x < kET * CH; x += blockDim.x) {
            const int t = x / CH, c = x - t * CH;
            if (t < cnt) {
                __pipeline_memcpy_async(sm + S::oT0 + t * LD + 4 * c, a.P + (int64_t)smi[t] * F + 4 * c, 16);
                __pipeline_memcpy_async(sm + S::oT1 + t * LD + 4 * c, a.Q + (int64_t)smi[kET + t] * F + 4 * c, 16);
            } else {
                *reinterpret_cast<float4*>(sm + S::oT0 + t * LD + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(sm + S::oT1 + t * LD + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __pipeline_commit();
        __pipeline_wait_prior(0);
    }
    __syncthreads();
    // stage 1: pre1 = P[i] + Q[j] + dist*wd -> LN1 -> xhat1 (T0), a1 (T1)
    for (int r = warp * 4; r < warp * 4 + 4; ++r) {
        float v[CPL];
        const float dist = sm[S::oSc + S_DIST * kET + r];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int col = lane + 32 * c;
            v[c] = sm[S::oT0 + r * LD + col] + sm[S::oT1 + r * LD + col] + dist * sm[S::oVec + V_WD * F + col];
        }
        const float rstd = ln_row<F>(v, a.eps);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int col = lane + 32 * c;
            sm[S::oT0 + r * LD + col] = v[c];
            sm[S::oT1 + r * LD + col] = actf<ACT>(fmaf(v[c], sm[S::oVec + V_G1 * F + col], sm[S::oVec + V_BE1 * F + col]));
        }
        if (lane == 0) sm[S::oSc + S_R1 * kET + r] = rstd;
    }
    __syncthreads();
    const int tx = threadIdx.x & 15;
    {   // stage 2: pre2 = a1 W1^T + b1 -> T2
        Frag<F, kRPT> f;
        f.zero();
        gemm_nt<F, kRPT>(f, sm + S::oT1, LD, sm + S::oW1, LD, F);
#pragma unroll
        for (int i = 0; i < kRPT; ++i)
#pragma unroll
            for (int j = 0; j < Frag<F, kRPT>::CPT; ++j) f.v[i][j] += sm[S::oVec + V_B1 * F + tx + 16 * j];
        store_nt<F, kRPT>(f, sm + S::oT2, LD);
    }
    __syncthreads();
    for (int r = warp * 4; r < warp * 4 + 4; ++r) {  // LN2 -> xhat2 (T2), m (T1)
        float v[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) v[c] = sm[S::oT2 + r * LD + lane + 32 * c];
        const float rstd = ln_row<F>(v, a.eps);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int col = lane + 32 * c;
            sm[S::oT2 + r * LD + col] = v[c];
            sm[S::oT1 + r * LD + col] = actf<ACT>(fmaf(v[c], sm[S::oVec + V_G2 * F + col], sm[S::oVec + V_BE2 * F + col]));
        }
        if (lane == 0) sm[S::oSc + S_R2 * kET + r] = rstd;
    }
    __syncthreads();
    {   // stage 3: pre3 = m W2^T + b2 -> T3
        Frag<F, kRPT> f;
        f.zero();
        gemm_nt<F, kRPT>(f, sm + S::oT1, LD, sm + S::oW2, LD, F);
#pragma unroll
        for (int i = 0; i < kRPT; ++i)
#pragma unroll
            for (int j = 0; j < Frag<F, kRPT>::CPT; ++j) f.v[i][j] += sm[S::oVec + V_B2 * F + tx + 16 * j];
        store_nt<F, kRPT>(f, sm + S::oT3, LD);
    }
    __syncthreads();
    const float b3 = __ldg(a.b3);
    for (int r = warp * 4; r < warp * 4 + 4; ++r) {  // LN3 -> xhat3 (T3); s = act(LN3) . w3 + b3
        float v[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) v[c] = sm[S::oT3 + r * LD + lane + 32 * c];
        const float rstd = ln_row<F>(v, a.eps);
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int col = lane + 32 * c;
            sm[S::oT3 + r * LD + col] = v[c];
            const float a3 = actf<ACT>(fmaf(v[c], sm[S::oVec + V_G3 * F + col], sm[S::oVec + V_BE3 * F + col]));
            dot = fmaf(a3, sm[S::oVec + V_W3 * F + col], dot);
        }
        dot = warp_sum(dot);
        if (lane == 0) {
            sm[S::oSc + S_R3 * kET + r] = rstd;
            sm[S::oSc + S_S * kET + r] = dot + b3;
        }
    }
    __syncthreads();
}

template <int F, int ACT>
__global__ void __launch_bounds__(256, 1) egnn_fwd_kernel(EgnnArgs a, float* __restrict__ msg_aggr, float* __restrict__ pos_aggr) {
    using S = EgnnSmem<F>;
    constexpr int LD = S::LD;
    extern __shared__ __align__(16) float sm[];
    egnn_load_weights<F>(sm, a);
    __syncthreads();
    for (int rg = blockIdx.x; rg < a.nranges; rg += gridDim.x) {
        const int r0 = lower_bound_row(a.rowptr, (int)a.n, (int64_t)rg * a.erange);
        const int r1 = (rg + 1 == a.nranges) ? (int)a.n : lower_bound_row(a.rowptr, (int)a.n, (int64_t)(rg + 1) * a.erange);
        if (r0 >= r1) continue;
        const int64_t eb = __ldg(a.rowptr + r0), ee = __ldg(a.rowptr + r1);
        int cur = r0;
        int64_t row_beg = eb, row_end = __ldg(a.rowptr + r0 + 1);
        float acc = 0.f;
        const int role = threadIdx.x < F ? 0 : (threadIdx.x < F + 3 ? 1 : 2);  // 0: feature column, 1: coordinate, 2: idle
        const int comp = threadIdx.x - F;
        for (int64_t e0 = eb; e0 < ee; e0 += kET) {
            const int cnt = (int)min((int64_t)kET, ee - e0);
            __syncthreads();
            egnn_tile_forward<F, ACT, false>(sm, a, r0, r1, e0, cnt);
            if (role != 2) {
                for (int t = 0; t < cnt; ++t) {
                    const int64_t e = e0 + t;
                    while (e >= row_end) {
                        const float deg = (float)(row_end - row_beg);
                        if (role == 0) msg_aggr[(int64_t)cur * F + threadIdx.x] = (a.aggr_mean && deg > 0.f) ? acc / deg : acc;
                        else pos_aggr[(int64_t)cur * 3 + comp] = deg > 0.f ? acc / deg : 0.f;
                        acc = 0.f;
                        ++cur;
                        row_beg = row_end;
                        row_end = __ldg(a.rowptr + cur + 1);
                    }
                    if (role == 0) acc += sm[S::oT1 + t * LD + threadIdx.x];
                    else acc += sm[S::oSc + (S_DX + comp) * kET + t] * sm[S::oSc + S_S * kET + t];
                }
            }
        }
        if (role != 2) {
            while (cur < r1) {
                const float deg = (float)(row_end - row_beg);
                if (role == 0) msg_aggr[(int64_t)cur * F + threadIdx.x] = (a.aggr_mean && deg > 0.f) ? acc / deg : acc;
                else pos_aggr[(int64_t)cur * 3 + comp] = deg > 0.f ? acc / deg : 0.f;
                acc = 0.f;
                ++cur;
                row_beg = row_end;
                if (cur < r1) row_end = __ldg(a.rowptr + cur + 1);
            }
        }
    }
}

// LayerNorm backward of one row: dy (already multiplied by act') -> d(pre);   accumulates dgamma / dbeta
template <int F>
__device__ __forceinline__ void ln_row_bwd(float (&dy)[F / 32], const float (&xh)[F / 32], const float* gamma, float rstd,
                                           int lane, float (&dg)[F / 32], float (&db)[F / 32]) {
    constexpr int CPL = F / 32;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        dg[c] = fmaf(dy[c], xh[c], dg[c]);
        db[c] += dy[c];
        dy[c] *= gamma[lane + 32 * c];
        s1 += dy[c];
        s2 = fmaf(dy[c], xh[c], s2);
    }
    s1 = warp_sum(s1) * (1.f / F);
    s2 = warp_sum(s2) * (1.f / F);
#pragma unroll
    for (int c = 0; c < CPL; ++c) dy[c] = rstd * (dy[c] - s1 - xh[c] * s2);
}

template <int F, int ACT, bool SRC_PASS>
__global__ void __launch_bounds__(256, 1)
egnn_bwd_kernel(EgnnArgs a, const float* __restrict__ g_msg, const float* __restrict__ g_pos, float* __restrict__ dnode,
                float* __restrict__ dpos, float* __restrict__ parts) {
    using S = EgnnSmem<F>;
    constexpr int LD = S::LD, CPL = F / 32, MT = F / 16;
    extern __shared__ __align__(16) float sm[];
    int* smi = reinterpret_cast<int*>(sm + S::oI);
    egnn_load_weights<F>(sm, a);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* vec = sm + S::oVec;

    float dW1[MT][MT], dW2[MT][MT];
    float dg1[CPL], dbe1[CPL], dg2[CPL], dbe2[CPL], dg3[CPL], dbe3[CPL], db1[CPL], db2[CPL], dw3[CPL], dwd[CPL];
    float db3 = 0.f;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < MT; ++j) dW1[i][j] = dW2[i][j] = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) dg1[c] = dbe1[c] = dg2[c] = dbe2[c] = dg3[c] = dbe3[c] = db1[c] = db2[c] = dw3[c] = dwd[c] = 0.f;

    for (int rg = blockIdx.x; rg < a.nranges; rg += gridDim.x) {
        const int r0 = lower_bound_row(a.rowptr, (int)a.n, (int64_t)rg * a.erange);
        const int r1 = (rg + 1 == a.nranges) ? (int)a.n : lower_bound_row(a.rowptr, (int)a.n, (int64_t)(rg + 1) * a.erange);
        if (r0 >= r1) continue;
        const int64_t eb = __ldg(a.rowptr + r0), ee = __ldg(a.rowptr + r1);
        int cur = r0;
        int64_t row_end = __ldg(a.rowptr + r0 + 1);
        float acc = 0.f;
        const int role = threadIdx.x < F ? 0 : (threadIdx.x < F + 3 ? 1 : 2);
        const int comp = threadIdx.x - F;
        for (int64_t e0 = eb; e0 < ee; e0 += kET) {
            const int cnt = (int)min((int64_t)kET, ee - e0);
            __syncthreads();
            egnn_tile_forward<F, ACT, SRC_PASS>(sm, a, r0, r1, e0, cnt);
            // upstream gradients per edge: coordinate part (mean over the destination's in-edges)
            if (threadIdx.x < kET) {
                const int t = threadIdx.x;
                float gx = 0.f, gy = 0.f, gz = 0.f, scale = 0.f;
                if (t < cnt) {
                    const int i = smi[t];
                    const float deg = (float)(__ldg(a.deg_rowptr + i + 1) - __ldg(a.deg_rowptr + i));
                    const float inv = deg > 0.f ? 1.f / deg : 0.f;
                    gx = __ldg(g_pos + 3 * (int64_t)i) * inv;
                    gy = __ldg(g_pos + 3 * (int64_t)i + 1) * inv;
                    gz = __ldg(g_pos + 3 * (int64_t)i + 2) * inv;
                    scale = a.aggr_mean ? inv : 1.f;
                }
                sm[S::oSc + S_GS * kET + t] = gx * sm[S::oSc + S_DX * kET + t] + gy * sm[S::oSc + S_DY * kET + t] + gz * sm[S::oSc + S_DZ * kET + t];
                sm[S::oSc + S_SCALE * kET + t] = scale;
                // d(delta) first part: g * s  (kept in registers of this thread until the end of the tile)
                sm[S::oSc + S_GX * kET + t] = gx;
                sm[S::oSc + S_GY * kET + t] = gy;
                sm[S::oSc + S_GZ * kET + t] = gz;
            }
            __syncthreads();
            // ---- stage 3 backward (rows by warp): ds -> dpre3 (into T3)
            for (int r = warp * 4; r < warp * 4 + 4; ++r) {
                float xh[CPL], dy[CPL];
                const float ds = sm[S::oSc + S_GS * kET + r];
                const float rstd = sm[S::oSc + S_R3 * kET + r];
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int col = lane + 32 * c;
                    xh[c] = sm[S::oT3 + r * LD + col];
                    const float y = fmaf(xh[c], vec[V_G3 * F + col], vec[V_BE3 * F + col]);
                    if (!SRC_PASS) dw3[c] = fmaf(ds, actf<ACT>(y), dw3[c]);
                    dy[c] = ds * vec[V_W3 * F + col] * dactf<ACT>(y);
                }
                if (!SRC_PASS && lane == 0 && r < cnt) db3 += ds;
                ln_row_bwd<F>(dy, xh, vec + V_G3 * F, rstd, lane, dg3, dbe3);
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const float v = r < cnt ? dy[c] : 0.f;
                    sm[S::oT3 + r * LD + lane + 32 * c] = v;
                    db2[c] += v;
                }
            }
            __syncthreads();
            if (!SRC_PASS) wgrad_tn<MT, MT>(dW2, sm + S::oT3, LD, sm + S::oT1, LD, cnt);  // dW2 += dpre3^T m
            {   // dm = dpre3 W2 (+ upstream) -> T1
                Frag<F, kRPT> f;
                f.zero();
                gemm_nn<F, kRPT>(f, sm + S::oT3, LD, sm + S::oW2, LD, F);
                __syncthreads();
                store_nn<F, kRPT>(f, sm + S::oT1, LD);
            }
            __syncthreads();
            // ---- stage 2 backward: dm (+ g_msg[i]) -> dpre2 (into T1); a1 recomputed into T3
            for (int r = warp * 4; r < warp * 4 + 4; ++r) {
                float xh[CPL], dy[CPL];
                const float rstd = sm[S::oSc + S_R2 * kET + r];
                const float scale = sm[S::oSc + S_SCALE * kET + r];
                const int i = smi[r];
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int col = lane + 32 * c;
                    xh[c] = sm[S::oT2 + r * LD + col];
                    const float y = fmaf(xh[c], vec[V_G2 * F + col], vec[V_BE2 * F + col]);
                    const float up = r < cnt ? __ldg(g_msg + (int64_t)i * F + col) * scale : 0.f;
                    dy[c] = (sm[S::oT1 + r * LD + col] + up) * dactf<ACT>(y);
                }
                ln_row_bwd<F>(dy, xh, vec + V_G2 * F, rstd, lane, dg2, dbe2);
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int col = lane + 32 * c;
                    const float v = r < cnt ? dy[c] : 0.f;
                    sm[S::oT1 + r * LD + col] = v;
                    db1[c] += v;
                    const float x1 = sm[S::oT0 + r * LD + col];
                    sm[S::oT3 + r * LD + col] = actf<ACT>(fmaf(x1, vec[V_G1 * F + col], vec[V_BE1 * F + col]));
                }
            }
            __syncthreads();
            if (!SRC_PASS) wgrad_tn<MT, MT>(dW1, sm + S::oT1, LD, sm + S::oT3, LD, cnt);  // dW1 += dpre2^T a1
            {   // da1 = dpre2 W1 -> T3
                Frag<F, kRPT> f;
                f.zero();
                gemm_nn<F, kRPT>(f, sm + S::oT1, LD, sm + S::oW1, LD, F);
                __syncthreads();
                store_nn<F, kRPT>(f, sm + S::oT3, LD);
            }
            __syncthreads();
            // ---- stage 1 backward: da1 -> dpre1 (into T3); distance / coordinate gradient
            for (int r = warp * 4; r < warp * 4 + 4; ++r) {
                float xh[CPL], dy[CPL];
                const float rstd = sm[S::oSc + S_R1 * kET + r];
                const float dist = sm[S::oSc + S_DIST * kET + r];
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int col = lane + 32 * c;
                    xh[c] = sm[S::oT0 + r * LD + col];
                    const float y = fmaf(xh[c], vec[V_G1 * F + col], vec[V_BE1 * F + col]);
                    dy[c] = sm[S::oT3 + r * LD + col] * dactf<ACT>(y);
                }
                ln_row_bwd<F>(dy, xh, vec + V_G1 * F, rstd, lane, dg1, dbe1);
                float dd = 0.f;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int col = lane + 32 * c;
                    const float v = r < cnt ? dy[c] : 0.f;
                    sm[S::oT3 + r * LD + col] = v;
                    dwd[c] = fmaf(dist, v, dwd[c]);
                    dd = fmaf(v, vec[V_WD * F + col], dd);
                }
                dd = warp_sum(dd);
                if (lane == 0) {
                    // d(delta) = g_pos_mean * s + d(dist) * delta / dist   (subgradient 0 at dist = 0, as torch.norm)
                    const float s = sm[S::oSc + S_S * kET + r];
                    const float k = dist > 0.f ? dd / dist : 0.f;
                    const float gx = sm[S::oSc + S_GX * kET + r], gy = sm[S::oSc + S_GY * kET + r], gz = sm[S::oSc + S_GZ * kET + r];
                    sm[S::oSc + S_GX * kET + r] = fmaf(k, sm[S::oSc + S_DX * kET + r], gx * s);
                    sm[S::oSc + S_GY * kET + r] = fmaf(k, sm[S::oSc + S_DY * kET + r], gy * s);
                    sm[S::oSc + S_GZ * kET + r] = fmaf(k, sm[S::oSc + S_DZ * kET + r], gz * s);
                }
            }
            __syncthreads();
            // ---- segmented sums over the CSR rows: d(pre1) -> dP (dst pass) / dQ (src pass); d(delta) -> dpos
            if (role != 2) {
                for (int t = 0; t < cnt; ++t) {
                    const int64_t e = e0 + t;
                    while (e >= row_end) {
                        if (role == 0) dnode[(int64_t)cur * F + threadIdx.x] = acc;
                        else dpos[(int64_t)cur * 3 + comp] = SRC_PASS ? -acc : acc;
                        acc = 0.f;
                        ++cur;
                        row_end = __ldg(a.rowptr + cur + 1);
                    }
                    if (role == 0) acc += sm[S::oT3 + t * LD + threadIdx.x];
                    else acc += sm[S::oSc + (S_GX + comp) * kET + t];
                }
            }
        }
        if (role != 2) {
            while (cur < r1) {
                if (role == 0) dnode[(int64_t)cur * F + threadIdx.x] = acc;
                else dpos[(int64_t)cur * 3 + comp] = SRC_PASS ? -acc : acc;
                acc = 0.f;
                ++cur;
            }
        }
    }
    if (SRC_PASS) return;
    // ---- per-CTA partial parameter gradients
    const int64_t plen = 2 * F * F + 10 * F + 4;
    float* my = parts + (int64_t)blockIdx.x * plen;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < MT; ++j) {
            my[(ty + 16 * i) * F + tx + 16 * j] = dW1[i][j];
            my[F * F + (ty + 16 * i) * F + tx + 16 * j] = dW2[i][j];
        }
    // column partials live per warp: reduce the 8 warps through shared memory, vector by vector
    float* red = sm + S::oRed;
    float* vout = my + 2 * F * F;
    auto reduce_vec = [&](const float (&v)[CPL], int slot) {
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CPL; ++c) red[warp * F + lane + 32 * c] = v[c];
        __syncthreads();
        if (threadIdx.x < F) {
            float s = 0.f;
            for (int w = 0; w < 8; ++w) s += red[w * F + threadIdx.x];
            vout[slot * F + threadIdx.x] = s;
        }
    };
    reduce_vec(db1, 0); reduce_vec(db2, 1); reduce_vec(dg1, 2); reduce_vec(dbe1, 3); reduce_vec(dg2, 4);
    reduce_vec(dbe2, 5); reduce_vec(dg3, 6); reduce_vec(dbe3, 7); reduce_vec(dw3, 8); reduce_vec(dwd, 9);
    __syncthreads();
    if (lane == 0) red[warp] = db3;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        vout[10 * F] = s;
        vout[10 * F + 1] = vout[10 * F + 2] = vout[10 * F + 3] = 0.f;
    }
}

static int egnn_check(const gmp_egnn_edge_params* p, int64_t n, int64_t E) {
    GMP_REQUIRE(p, "egnn: params is NULL");
    GMP_REQUIRE(p->d == 64 || p->d == 128, "egnn: emb_dim must be 64 or 128 (got %d)", p->d);
    GMP_REQUIRE(p->act == 0 || p->act == 1, "egnn: act must be 0 (relu) or 1 (swish)");
    GMP_REQUIRE(p->wd && p->ln1_g && p->ln1_b && p->w1 && p->b1 && p->ln2_g && p->ln2_b && p->w2 && p->b2 && p->ln3_g &&
                p->ln3_b && p->w3 && p->b3, "egnn: NULL parameter pointer");
    GMP_REQUIRE(n >= 0 && E >= 0 && n < (1ll << 31) && E < (1ll << 31), "egnn: sizes out of range");
    return GMP_OK;
}

// edges per work range: aim at two ranges per SM, between one 32-edge tile and kERangeMax, in whole tiles
static int egnn_range(int64_t E) {
    int64_t r = ceil_div(E > 0 ? E : 1, 2 * (int64_t)num_sms());
    r = ceil_div(r, 32) * 32;
    return (int)(r < 32 ? 32 : (r > kERangeMax ? kERangeMax : r));
}

static EgnnArgs egnn_args(const int32_t* rowptr, const int32_t* col, const int32_t* deg_rowptr, int64_t n, int64_t E,
                          const float* P, const float* Q, const float* pos, const gmp_egnn_edge_params* p) {
    EgnnArgs a;
    a.rowptr = rowptr; a.col = col; a.deg_rowptr = deg_rowptr; a.n = n; a.E = E; a.P = P; a.Q = Q; a.pos = pos;
    a.wd = p->wd; a.g1 = p->ln1_g; a.be1 = p->ln1_b; a.w1 = p->w1; a.b1 = p->b1; a.g2 = p->ln2_g; a.be2 = p->ln2_b;
    a.w2 = p->w2; a.b2 = p->b2; a.g3 = p->ln3_g; a.be3 = p->ln3_b; a.w3 = p->w3; a.b3 = p->b3;
    a.act = p->act; a.aggr_mean = p->aggr_mean; a.eps = p->ln_eps;
    a.erange = egnn_range(E);
    a.nranges = (int)(E > 0 ? ceil_div(E, a.erange) : 1);
    return a;
}

template <int F, int ACT>
static int egnn_launch_fwd(const EgnnArgs& a, float* m, float* p, cudaStream_t s) {
    const size_t smem = EgnnSmem<F>::total * sizeof(float);
    const int grid = a.nranges < num_sms() ? a.nranges : num_sms();
    GMP_CUDA(cudaFuncSetAttribute(egnn_fwd_kernel<F, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    egnn_fwd_kernel<F, ACT><<<grid, 256, smem, s>>>(a, m, p);
    return check_launch("egnn_fwd_kernel");
}

template <int F, int ACT, bool SRC>
static int egnn_launch_bwd(const EgnnArgs& a, const float* gm, const float* gp, float* dn, float* dp, float* parts, cudaStream_t s) {
    const size_t smem = EgnnSmem<F>::total * sizeof(float);
    const int grid = a.nranges < num_sms() ? a.nranges : num_sms();
    GMP_CUDA(cudaFuncSetAttribute(egnn_bwd_kernel<F, ACT, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    egnn_bwd_kernel<F, ACT, SRC><<<grid, 256, smem, s>>>(a, gm, gp, dn, dp, parts);
    return check_launch("egnn_bwd_kernel");
}

}  // namespace gmp

using namespace gmp;

extern "C" {

int32_t gmp_egnn_bwd_num_parts(int64_t num_edges) {
    const int64_t nr = num_edges > 0 ? ceil_div(num_edges, egnn_range(num_edges)) : 1;
    return (int32_t)(nr < num_sms() ? nr : num_sms());
}

int64_t gmp_egnn_bwd_part_len(int32_t d) { return 2ll * d * d + 10ll * d + 4; }

int gmp_egnn_edge_fwd(const int32_t* rowptr, const int32_t* col, int64_t n, int64_t num_edges, const float* P,
                      const float* Q, const float* pos, const gmp_egnn_edge_params* prm, float* msg_aggr,
                      float* pos_aggr, int32_t precision, gmp_stream_t stream) {
    if (int rc = egnn_check(prm, n, num_edges)) return rc;
    GMP_REQUIRE(rowptr && msg_aggr && pos_aggr && pos && (num_edges == 0 || (col && P && Q)), "egnn_edge_fwd: NULL pointer");
    if (precision != GMP_FP32_STRICT) {
        set_error("egnn_edge_fwd: precision mode %d is not built into this library", precision);
        return GMP_ERR_UNSUPPORTED;
    }
    if (n == 0) return GMP_OK;
    const EgnnArgs a = egnn_args(rowptr, col, rowptr, n, num_edges, P, Q, pos, prm);
    if (prm->d == 128) return prm->act ? egnn_launch_fwd<128, 1>(a, msg_aggr, pos_aggr, stream) : egnn_launch_fwd<128, 0>(a, msg_aggr, pos_aggr, stream);
    return prm->act ? egnn_launch_fwd<64, 1>(a, msg_aggr, pos_aggr, stream) : egnn_launch_fwd<64, 0>(a, msg_aggr, pos_aggr, stream);
}

int gmp_egnn_edge_bwd(const int32_t* rowptr, const int32_t* col, const int32_t* dst_rowptr, int64_t n, int64_t num_edges,
                      const float* P, const float* Q, const float* pos, const gmp_egnn_edge_params* prm,
                      const float* g_msg, const float* g_pos, int32_t src_pass, float* d_node, float* d_pos,
                      float* wgrad_parts, int32_t precision, gmp_stream_t stream) {
    if (int rc = egnn_check(prm, n, num_edges)) return rc;
    GMP_REQUIRE(rowptr && dst_rowptr && g_msg && g_pos && d_node && d_pos && pos && (num_edges == 0 || (col && P && Q)),
                "egnn_edge_bwd: NULL pointer");
    GMP_REQUIRE(src_pass || wgrad_parts, "egnn_edge_bwd: the dst pass needs wgrad_parts");
    if (precision != GMP_FP32_STRICT) {
        set_error("egnn_edge_bwd: precision mode %d is not built into this library", precision);
        return GMP_ERR_UNSUPPORTED;
    }
    if (n == 0) return GMP_OK;
    const EgnnArgs a = egnn_args(rowptr, col, dst_rowptr, n, num_edges, P, Q, pos, prm);
#define GMP_EGNN_BWD(F_, A_)                                                                                      \
    (src_pass ? egnn_launch_bwd<F_, A_, true>(a, g_msg, g_pos, d_node, d_pos, wgrad_parts, stream)               \
              : egnn_launch_bwd<F_, A_, false>(a, g_msg, g_pos, d_node, d_pos, wgrad_parts, stream))
    if (prm->d == 128) return prm->act ? GMP_EGNN_BWD(128, 1) : GMP_EGNN_BWD(128, 0);
    return prm->act ? GMP_EGNN_BWD(64, 1) : GMP_EGNN_BWD(64, 0);
#undef GMP_EGNN_BWD
}

}  // extern "C"
