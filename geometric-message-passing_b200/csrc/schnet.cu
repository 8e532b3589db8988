// SchNet continuous-filter convolution, fused: per-edge filter MLP -> x1[src] * W_e -> segmented sum.
// fp32 strict path (CUDA-core FFMA).  The per-edge filter W_e [E,F] never touches HBM.
//
// Work decomposition: the dst-sorted edge list is cut into ranges of up to kEdgesPerRangeMax edges at row
// granularity (a CTA owns whole destination rows, so no cross-CTA reduction exists); a persistent
// grid of <= #SM CTAs walks ranges b, b+grid, ...  Inside a range, 64-edge tiles go through
//   rbf tile -> GEMM1 (+b1, ssp) -> GEMM2 (+b2, *C) -> * gathered x1 rows -> column-thread segmented sum.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace gmp {

constexpr int kTile = 64;             // edges per tile
constexpr int kEdgesPerRangeMax = 1024;  // edges per work range (16 tiles) on large graphs; small graphs: two ranges per SM, >= 1 tile
constexpr int kGP = 64;               // Gaussian dimension padded (G <= 64)
constexpr int kLdR = kGP + 4;

struct SchnetArgs {
    const int32_t *rowptr, *col, *perm;
    int64_t n, E;
    const float *ew, *ea, *x1;
    const float *w1, *b1, *w2, *b2, *goff;
    int G;
    float cutoff, gcoeff;
    int nranges, erange;
};

template <int F>
struct SchnetSmem {
    static constexpr int LD = F + 4;
    // layout (floats)
    static constexpr int oW1 = 0;                    // [F][kLdR]
    static constexpr int oW2 = oW1 + F * kLdR;       // [F][LD]
    static constexpr int oB1 = oW2 + F * LD;         // [F]
    static constexpr int oB2 = oB1 + F;              // [F]
    static constexpr int oR = oB2 + F;               // [64][kLdR]
    static constexpr int oH = oR + kTile * kLdR;     // [64][LD]
    static constexpr int oX = oH + kTile * LD;       // [64][LD]
    static constexpr int oG = oX + kTile * LD;       // [64][LD]   (backward only)
    static constexpr int oSc = oG + kTile * LD;      // per-edge scalars: d[64], C[64], dC[64], gchain[64]
    static constexpr int oI = oSc + 4 * kTile;       // per-edge ints: src[64], eid[64], row[64]
    static constexpr int total_floats = oI + 3 * kTile;  // one layout for forward and backward (~216 KB at F=128)
};

template <int F>
__device__ __forceinline__ void load_filter(float* sm, const SchnetArgs& a) {
    using S = SchnetSmem<F>;
    for (int i = threadIdx.x; i < F * kLdR; i += blockDim.x) {
        const int f = i / kLdR, g = i - f * kLdR;
        sm[S::oW1 + i] = (g < a.G) ? __ldg(a.w1 + f * a.G + g) : 0.f;
    }
    for (int i = threadIdx.x; i < F * S::LD; i += blockDim.x) {
        const int f = i / S::LD, k = i - f * S::LD;
        sm[S::oW2 + i] = (k < F) ? __ldg(a.w2 + f * F + k) : 0.f;
    }
    for (int i = threadIdx.x; i < F; i += blockDim.x) {
        sm[S::oB1 + i] = __ldg(a.b1 + i);
        sm[S::oB2 + i] = __ldg(a.b2 + i);
    }
}

// per-edge scalars + rbf tile for edges [e0, e0 + cnt)
template <int F, bool HAS_ATTR>
__device__ __forceinline__ void stage_edges(float* sm, int* smi, const SchnetArgs& a, int64_t e0, int cnt) {
    using S = SchnetSmem<F>;
    if (threadIdx.x < kTile) {
        const int t = threadIdx.x;
        float d = 0.f, C = 0.f;
        int src = 0, eid = 0;
        if (t < cnt) {
            const int64_t k = e0 + t;
            eid = a.perm ? __ldg(a.perm + k) : (int)k;
            src = __ldg(a.col + k);
            d = __ldg(a.ew + eid);
            C = 0.5f * (cosf(d * 3.14159265358979323846f / a.cutoff) + 1.0f);
        }
        sm[S::oSc + t] = d;
        sm[S::oSc + kTile + t] = C;
        smi[t] = src;
        smi[kTile + t] = eid;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kTile * kGP; i += blockDim.x) {
        const int t = i / kGP, g = i - t * kGP;
        float v = 0.f;
        if (t < cnt && g < a.G) {
            if (HAS_ATTR) {
                v = __ldg(a.ea + (int64_t)smi[kTile + t] * a.G + g);
            } else {
                const float u = sm[S::oSc + t] - __ldg(a.goff + g);
                v = expf(a.gcoeff * u * u);
            }
        }
        sm[S::oR + t * kLdR + g] = v;
    }
}

// async gather of `cnt` rows of F floats: dst[t][:] = base[idx[t]][:]
template <int F>
__device__ __forceinline__ void gather_rows_async(float* dst, int ld, const float* __restrict__ base, const int* idx, int cnt) {
    constexpr int CH = F / 4;
    for (int i = threadIdx.x; i < kTile * CH; i += blockDim.x) {
        const int t = i / CH, c = i - t * CH;
        if (t < cnt) {
            __pipeline_memcpy_async(dst + t * ld + 4 * c, base + (int64_t)idx[t] * F + 4 * c, 16);
        } else {
            *reinterpret_cast<float4*>(dst + t * ld + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __pipeline_commit();
}

template <int F, bool HAS_ATTR>
__global__ void __launch_bounds__(256, 1) schnet_fwd_kernel(SchnetArgs a, float* __restrict__ agg) {
    using S = SchnetSmem<F>;
    constexpr int LD = S::LD;
    extern __shared__ __align__(16) float sm[];
    int* smi = reinterpret_cast<int*>(sm + S::oI);
    load_filter<F>(sm, a);
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    for (int rg = blockIdx.x; rg < a.nranges; rg += gridDim.x) {
        const int r0 = lower_bound_row(a.rowptr, (int)a.n, (int64_t)rg * a.erange);
        const int r1 = (rg + 1 == a.nranges) ? (int)a.n : lower_bound_row(a.rowptr, (int)a.n, (int64_t)(rg + 1) * a.erange);
        if (r0 >= r1) continue;
        const int64_t eb = __ldg(a.rowptr + r0), ee = __ldg(a.rowptr + r1);
        // column-thread reduction state
        int cur = r0;
        int64_t row_end = __ldg(a.rowptr + r0 + 1);
        float acc = 0.f;
        for (int64_t e0 = eb; e0 < ee; e0 += kTile) {
            const int cnt = (int)min((int64_t)kTile, ee - e0);
            __syncthreads();  // previous tile fully consumed
            stage_edges<F, HAS_ATTR>(sm, smi, a, e0, cnt);
            gather_rows_async<F>(sm + S::oX, LD, a.x1, smi, cnt);
            __syncthreads();
            // GEMM1: pre1 = R W1^T + b1 ; h1 = ssp(pre1)
            {
                Frag<F> f1;
                f1.zero();
                gemm_nt<F>(f1, sm + S::oR, kLdR, sm + S::oW1, kLdR, kGP);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < Frag<F>::CPT; ++j) f1.v[i][j] = ssp(f1.v[i][j] + sm[S::oB1 + tx + 16 * j]);
                store_nt<F>(f1, sm + S::oH, LD);
            }
            __syncthreads();
            // GEMM2: W_e = (h1 W2^T + b2) * C ; message = W_e * x1[src]
            {
                Frag<F> f2;
                f2.zero();
                gemm_nt<F>(f2, sm + S::oH, LD, sm + S::oW2, LD, F);
                __pipeline_wait_prior(0);
                __syncthreads();  // gathered x1 rows visible to all threads
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = ty * 4 + i;
                    const float C = sm[S::oSc + kTile + r];
#pragma unroll
                    for (int j = 0; j < Frag<F>::CPT; ++j) {
                        const int c = tx + 16 * j;
                        float* px = sm + S::oX + r * LD + c;
                        *px = (f2.v[i][j] + sm[S::oB2 + c]) * C * (*px);
                    }
                }
            }
            __syncthreads();
            // segmented sum: thread f walks the tile's edges in order
            if (threadIdx.x < F) {
                const int f = threadIdx.x;
                for (int t = 0; t < cnt; ++t) {
                    const int64_t e = e0 + t;
                    while (e >= row_end) {
                        agg[(int64_t)cur * F + f] = acc;
                        acc = 0.f;
                        ++cur;
                        row_end = __ldg(a.rowptr + cur + 1);
                    }
                    acc += sm[S::oX + t * LD + f];
                }
            }
        }
        if (threadIdx.x < F) {
            while (cur < r1) {
                agg[(int64_t)cur * F + threadIdx.x] = acc;
                acc = 0.f;
                ++cur;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward over the dst-sorted CSR: weight gradients (per-CTA partials) and edge-input gradients
// ------------------------------------------------------------------------------------------------
template <int F, bool HAS_ATTR, bool NEED_DEDGE>
__global__ void __launch_bounds__(256, 1)
schnet_bwd_kernel(SchnetArgs a, const float* __restrict__ g_agg, float* __restrict__ parts,
                  float* __restrict__ d_ew, float* __restrict__ d_ea) {
    using S = SchnetSmem<F>;
    constexpr int LD = S::LD;
    constexpr int CPT = Frag<F>::CPT;
    constexpr int MT = F / 16;
    extern __shared__ __align__(16) float sm[];
    int* smi = reinterpret_cast<int*>(sm + S::oI);
    load_filter<F>(sm, a);
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    float dW2[MT][MT];   // [f' = ty+16i][f = tx+16j]
    float dW1[MT][4];    // [f  = ty+16i][g = tx+16j]
#pragma unroll
    for (int i = 0; i < MT; ++i) {
#pragma unroll
        for (int j = 0; j < MT; ++j) dW2[i][j] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) dW1[i][j] = 0.f;
    }
    float db1 = 0.f, db2 = 0.f;  // threads < F own one bias column each

    for (int rg = blockIdx.x; rg < a.nranges; rg += gridDim.x) {
        const int r0 = lower_bound_row(a.rowptr, (int)a.n, (int64_t)rg * a.erange);
        const int r1 = (rg + 1 == a.nranges) ? (int)a.n : lower_bound_row(a.rowptr, (int)a.n, (int64_t)(rg + 1) * a.erange);
        if (r0 >= r1) continue;
        const int64_t eb = __ldg(a.rowptr + r0), ee = __ldg(a.rowptr + r1);
        for (int64_t e0 = eb; e0 < ee; e0 += kTile) {
            const int cnt = (int)min((int64_t)kTile, ee - e0);
            __syncthreads();
            stage_edges<F, HAS_ATTR>(sm, smi, a, e0, cnt);
            if (threadIdx.x < kTile) {
                // destination row of each edge: last row in [r0, r1) whose start is <= e
                const int64_t e = e0 + threadIdx.x;
                int lo = r0, hi = r1;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if ((int64_t)__ldg(a.rowptr + mid) <= e) lo = mid; else hi = mid;
                }
                smi[2 * kTile + threadIdx.x] = lo;
            }
            __syncthreads();
            gather_rows_async<F>(sm + S::oX, LD, a.x1, smi, cnt);
            gather_rows_async<F>(sm + S::oG, LD, g_agg, smi + 2 * kTile, cnt);
            // GEMM1 -> h1 (smem), sigmoid(pre1) (registers, NT layout)
            Frag<F> sig;
            {
                Frag<F> f1;
                f1.zero();
                gemm_nt<F>(f1, sm + S::oR, kLdR, sm + S::oW1, kLdR, kGP);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        const float pre = f1.v[i][j] + sm[S::oB1 + tx + 16 * j];
                        sig.v[i][j] = sigmoidf_(pre);
                        f1.v[i][j] = ssp(pre);
                    }
                store_nt<F>(f1, sm + S::oH, LD);
            }
            __pipeline_wait_prior(0);
            __syncthreads();
            // dWf = x1[src] * g[dst];  dC = <dWf, nn_out>;  dpre2 = dWf * C   (written over the X tile)
            {
                Frag<F> f2;
                f2.zero();
                if (NEED_DEDGE) gemm_nt<F>(f2, sm + S::oH, LD, sm + S::oW2, LD, F);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = ty * 4 + i;
                    const float C = sm[S::oSc + kTile + r];
                    float dc = 0.f;
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        const int c = tx + 16 * j;
                        float* px = sm + S::oX + r * LD + c;
                        const float dwf = (*px) * sm[S::oG + r * LD + c];
                        if (NEED_DEDGE) dc = fmaf(dwf, f2.v[i][j] + sm[S::oB2 + c], dc);
                        *px = dwf * C;
                    }
                    if (NEED_DEDGE) {
                        // reduce over the 16 tx lanes that share this row (half-warp)
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) dc += __shfl_xor_sync(0xffffffffu, dc, o);
                        if (tx == 0) sm[S::oSc + 2 * kTile + r] = dc;
                    }
                }
            }
            __syncthreads();
            // weight gradient of the second Linear and its bias
            wgrad_tn<MT, MT>(dW2, sm + S::oX, LD, sm + S::oH, LD, cnt);
            if (threadIdx.x < F) {
                float s = 0.f;
                for (int t = 0; t < cnt; ++t) s += sm[S::oX + t * LD + threadIdx.x];
                db2 += s;
            }
            // dh1 = dpre2 W2  -> G tile (NN layout), then dpre1 = dh1 * sigmoid(pre1) at the NT positions
            {
                Frag<F> f3;
                f3.zero();
                gemm_nn<F>(f3, sm + S::oX, LD, sm + S::oW2, LD, F);
                __syncthreads();  // all reads of G (dWf) and X done before G is overwritten
                store_nn<F>(f3, sm + S::oG, LD);
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    float* p = sm + S::oG + (ty * 4 + i) * LD + tx + 16 * j;
                    *p = (*p) * sig.v[i][j];
                }
            __syncthreads();
            // weight gradient of the first Linear and its bias
            wgrad_tn<MT, 4>(dW1, sm + S::oG, LD, sm + S::oR, kLdR, cnt);
            if (threadIdx.x < F) {
                float s = 0.f;
                for (int t = 0; t < cnt; ++t) s += sm[S::oG + t * LD + threadIdx.x];
                db1 += s;
            }
            if (NEED_DEDGE) {
                // dR[e][g] = sum_f dpre1[e][f] W1[f][g]   (NN with B = W1 [F][kLdR]) -> H tile (free now)
                __syncthreads();
                Frag<64> f4;
                f4.zero();
                gemm_nn<64>(f4, sm + S::oG, LD, sm + S::oW1, kLdR, F);
                store_nn<64>(f4, sm + S::oH, LD);
                __syncthreads();
                if (HAS_ATTR) {
                    if (d_ea) {
                        for (int i = threadIdx.x; i < cnt * a.G; i += blockDim.x) {
                            const int t = i / a.G, g = i - t * a.G;
                            d_ea[(int64_t)smi[kTile + t] * a.G + g] = sm[S::oH + t * LD + g];
                        }
                    }
                }
                if (d_ew && threadIdx.x < cnt) {
                    const int t = threadIdx.x;
                    const float d = sm[S::oSc + t];
                    const float w = 3.14159265358979323846f / a.cutoff;
                    float gd = sm[S::oSc + 2 * kTile + t] * (-0.5f * w * sinf(d * w));
                    if (!HAS_ATTR) {
                        // chain through the recomputed Gaussian expansion: d rbf_g / dd = 2 coeff (d - off_g) rbf_g
                        float s = 0.f;
                        for (int g = 0; g < a.G; ++g) {
                            const float u = d - __ldg(a.goff + g);
                            s = fmaf(sm[S::oH + t * LD + g] * sm[S::oR + t * kLdR + g], 2.f * a.gcoeff * u, s);
                        }
                        gd += s;
                    }
                    d_ew[smi[kTile + t]] = gd;
                }
            }
        }
    }
    // per-CTA partial gradients: [dW1 F x 64 | db1 F | dW2 F x F | db2 F]
    float* my = parts + (int64_t)blockIdx.x * (F * kGP + F + F * F + F);
#pragma unroll
    for (int i = 0; i < MT; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) my[(ty + 16 * i) * kGP + tx + 16 * j] = dW1[i][j];
#pragma unroll
        for (int j = 0; j < MT; ++j) my[F * kGP + F + (ty + 16 * i) * F + tx + 16 * j] = dW2[i][j];
    }
    if (threadIdx.x < F) {
        my[F * kGP + threadIdx.x] = db1;
        my[F * kGP + F + F * F + threadIdx.x] = db2;
    }
}

static int schnet_check(const gmp_schnet_filter* f, int64_t n, int64_t E) {
    GMP_REQUIRE(f, "schnet: filter descriptor is NULL");
    GMP_REQUIRE(f->num_filters == 64 || f->num_filters == 128, "schnet: num_filters must be 64 or 128 (got %d)", f->num_filters);
    GMP_REQUIRE(f->num_gaussians >= 1 && f->num_gaussians <= kGP, "schnet: num_gaussians must be in [1, %d]", kGP);
    GMP_REQUIRE(f->w1 && f->b1 && f->w2 && f->b2, "schnet: NULL filter weights");
    GMP_REQUIRE(n >= 0 && E >= 0 && n < (1ll << 31) && E < (1ll << 31), "schnet: sizes out of range");
    return GMP_OK;
}

static int schnet_range(int64_t E) {
    int64_t r = ceil_div(E > 0 ? E : 1, 2 * (int64_t)num_sms());
    r = ceil_div(r, kTile) * kTile;
    return (int)(r < kTile ? kTile : (r > kEdgesPerRangeMax ? kEdgesPerRangeMax : r));
}

static SchnetArgs make_args(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t E,
                            const float* ew, const float* ea, const float* x1, const gmp_schnet_filter* f) {
    SchnetArgs a;
    a.rowptr = rowptr; a.col = col; a.perm = perm; a.n = n; a.E = E;
    a.ew = ew; a.ea = ea; a.x1 = x1;
    a.w1 = f->w1; a.b1 = f->b1; a.w2 = f->w2; a.b2 = f->b2; a.goff = f->gauss_offset;
    a.G = f->num_gaussians; a.cutoff = f->cutoff; a.gcoeff = f->gauss_coeff;
    a.erange = schnet_range(E);
    a.nranges = (int)(E > 0 ? ceil_div(E, a.erange) : 1);
    return a;
}

template <int F>
static int launch_fwd(const SchnetArgs& a, bool has_attr, float* agg, cudaStream_t s) {
    const size_t smem = SchnetSmem<F>::total_floats * sizeof(float);
    const int grid = a.nranges < num_sms() ? a.nranges : num_sms();
    if (has_attr) {
        GMP_CUDA(cudaFuncSetAttribute(schnet_fwd_kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        schnet_fwd_kernel<F, true><<<grid, 256, smem, s>>>(a, agg);
    } else {
        GMP_CUDA(cudaFuncSetAttribute(schnet_fwd_kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        schnet_fwd_kernel<F, false><<<grid, 256, smem, s>>>(a, agg);
    }
    return check_launch("schnet_fwd_kernel");
}

template <int F, bool HAS_ATTR, bool NEED>
static int launch_bwd_t(const SchnetArgs& a, const float* g, float* parts, float* d_ew, float* d_ea, cudaStream_t s) {
    const size_t smem = SchnetSmem<F>::total_floats * sizeof(float);
    const int grid = a.nranges < num_sms() ? a.nranges : num_sms();
    GMP_CUDA(cudaFuncSetAttribute(schnet_bwd_kernel<F, HAS_ATTR, NEED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    schnet_bwd_kernel<F, HAS_ATTR, NEED><<<grid, 256, smem, s>>>(a, g, parts, d_ew, d_ea);
    return check_launch("schnet_bwd_kernel");
}

template <int F>
static int launch_bwd(const SchnetArgs& a, bool has_attr, bool need, const float* g, float* parts, float* d_ew,
                      float* d_ea, cudaStream_t s) {
    if (has_attr) return need ? launch_bwd_t<F, true, true>(a, g, parts, d_ew, d_ea, s) : launch_bwd_t<F, true, false>(a, g, parts, d_ew, d_ea, s);
    return need ? launch_bwd_t<F, false, true>(a, g, parts, d_ew, d_ea, s) : launch_bwd_t<F, false, false>(a, g, parts, d_ew, d_ea, s);
}

int schnet_fwd_tc_launch(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t E,
                         const float* ew, const float* ea, const float* x1, const gmp_schnet_filter* f, float* agg,
                         cudaStream_t stream);  // schnet_tc.cu
int schnet_bwd_tc_launch(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t E,
                         const float* ew, const float* ea, const float* x1, const gmp_schnet_filter* f, const float* g_agg,
                         float* parts, int nparts, cudaStream_t stream);  // schnet_tc.cu

}  // namespace gmp

using namespace gmp;

extern "C" {

int32_t gmp_schnet_bwd_num_parts(int64_t num_edges) {
    const int64_t nr = num_edges > 0 ? ceil_div(num_edges, schnet_range(num_edges)) : 1;
    return (int32_t)(nr < num_sms() ? nr : num_sms());
}

int64_t gmp_schnet_bwd_part_len(int32_t num_gaussians, int32_t num_filters) {
    (void)num_gaussians;
    return (int64_t)num_filters * kGP + num_filters + (int64_t)num_filters * num_filters + num_filters;
}

int gmp_schnet_cfconv_fwd(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t num_edges,
                          const float* edge_weight, const float* edge_attr, const float* x1,
                          const gmp_schnet_filter* filt, float* agg, int32_t precision, gmp_stream_t stream) {
    if (int rc = schnet_check(filt, n, num_edges)) return rc;
    GMP_REQUIRE(rowptr && agg && (num_edges == 0 || (col && edge_weight && x1)), "schnet_cfconv_fwd: NULL pointer");
    GMP_REQUIRE(edge_attr || filt->gauss_offset, "schnet_cfconv_fwd: need edge_attr or gauss_offset");
    GMP_REQUIRE(precision == GMP_FP32_STRICT || precision == GMP_BF16_TC, "schnet_cfconv_fwd: unknown precision mode %d", precision);
    if (n == 0) return GMP_OK;
    if (precision == GMP_BF16_TC) return schnet_fwd_tc_launch(rowptr, col, perm, n, num_edges, edge_weight, edge_attr, x1, filt, agg, stream);
    const SchnetArgs a = make_args(rowptr, col, perm, n, num_edges, edge_weight, edge_attr, x1, filt);
    return filt->num_filters == 128 ? launch_fwd<128>(a, edge_attr != nullptr, agg, stream)
                                    : launch_fwd<64>(a, edge_attr != nullptr, agg, stream);
}

int gmp_schnet_cfconv_bwd(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t num_edges,
                          const float* edge_weight, const float* edge_attr, const float* x1,
                          const gmp_schnet_filter* filt, const float* g_agg, float* wgrad_parts, float* d_edge_weight,
                          float* d_edge_attr, int32_t precision, gmp_stream_t stream) {
    if (int rc = schnet_check(filt, n, num_edges)) return rc;
    GMP_REQUIRE(rowptr && g_agg && wgrad_parts && (num_edges == 0 || (col && edge_weight && x1)), "schnet_cfconv_bwd: NULL pointer");
    GMP_REQUIRE(edge_attr || filt->gauss_offset, "schnet_cfconv_bwd: need edge_attr or gauss_offset");
    GMP_REQUIRE(!d_edge_attr || edge_attr, "schnet_cfconv_bwd: d_edge_attr requested without edge_attr");
    GMP_REQUIRE(precision == GMP_FP32_STRICT || precision == GMP_BF16_TC, "schnet_cfconv_bwd: unknown precision mode %d", precision);
    const SchnetArgs a = make_args(rowptr, col, perm, n, num_edges, edge_weight, edge_attr, x1, filt);
    const bool need = d_edge_weight != nullptr || d_edge_attr != nullptr;
    // tensor-core variant: weight gradients only, F = 128, G <= 63 (column 63 of the basis tile carries the bias sums);
    // gradients w.r.t. the edge inputs run the fp32 kernel in either mode
    if (precision == GMP_BF16_TC && !need && filt->num_filters == 128 && filt->num_gaussians <= 63 && num_edges > 0)
        return schnet_bwd_tc_launch(rowptr, col, perm, n, num_edges, edge_weight, edge_attr, x1, filt, g_agg, wgrad_parts,
                                    gmp_schnet_bwd_num_parts(num_edges), stream);
    return filt->num_filters == 128 ? launch_bwd<128>(a, edge_attr != nullptr, need, g_agg, wgrad_parts, d_edge_weight, d_edge_attr, stream)
                                    : launch_bwd<64>(a, edge_attr != nullptr, need, g_agg, wgrad_parts, d_edge_weight, d_edge_attr, stream);
}

}  // extern "C"
