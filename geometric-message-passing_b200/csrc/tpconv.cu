// TFN / MACE tensor-product convolution (models/layers/tfn_layer.py:82-87), fused, fp32 strict:
//     w_e   = fc(edge_feat_e)                    [weight_numel]   (Linear(R,H) -> ReLU -> Linear(H,numel))
//     tp_e  = FullyConnectedTensorProduct(x[gather_e], sh_e; w_e)  ('uvw' paths, e3nn normalisation)
//     out_n = sum_{e in row n} tp_e
// The per-edge weight vector (272 KB / 704 KB per edge at the BASELINE configs) is never materialised:
// a CTA keeps a 128-row slice of fc's second Linear in shared memory (weight-stationary), streams the
// edges of its row chunk through it in 32-edge tiles (hidden layer recomputed per tile), and contracts the
// 128 generated weights of each edge with the gathered features on the spot.
//
// One templated "contract" kernel serves the forward and the feature gradient, because
//     out_e[w,k] = sum_u T_e[u,w] * ( sum_i x[u,i]  Z_e[i,k] )        (forward;  kept = w, summed = u)
//     dx_e [u,i] = sum_w T_e[u,w] * ( sum_k g[w,k]  Z_e[i,k] )        (backward; kept = u, summed = w)
// are the same contraction with the roles of the two multiplicity indices exchanged
// (Z_e[i,k] = coeff * sum_j sh_e[j] C[i,j,k]).  A "pass" = (path, chunk of the summed index); passes that write
// the same result block run one after the other inside one CTA, so accumulation across paths is a plain
// read-modify-write of rows this CTA owns: deterministic, no atomics.
// The weight gradient (dW2, db2, dW1, db1) has its own kernel: a CTA owns 64 rows of W2 and streams all edges.
#include "common.cuh"

namespace gmp {

constexpr int kTT = 32;          // edges per tile
constexpr int kTRPT = 2;
constexpr int kSliceCols = 128;  // generated weights per edge per pass
constexpr int kMaxRes = 80;      // max (kept slice) x (2l+1) results per edge per pass
constexpr int kMaxR = 16;        // max edge_feat dim
constexpr int kMaxS = 16;        // max edge_sh dim (l <= 3)

// one pass of the contract kernel (host-built, lives in device memory)
struct TpPass {
    int32_t w_off;      // first row of the path's block in W2 / b2
    int32_t stride_a;   // W2 row = w_off + a * stride_a + b * stride_b   (a = summed index, b = kept index)
    int32_t stride_b;
    int32_t a0, MC;     // chunk of the summed index: a in [a0, a0 + MC)
    int32_t v_off, DA;  // gathered tensor: V[node][v_off + a*DA + iA]
    int32_t sh_off, DS; // spherical-harmonic block of this path
    int32_t cg_off;     // Zc[iA][j][kB] (DA x DS x DB floats, coefficient folded in) in the CG table
    int32_t pad0, pad1;
};

// result blocks (output irreps for the forward, input irreps for the feature gradient)
struct TpBlock {
    int32_t r_off;      // float offset of the block in a result row
    int32_t MB, DB;     // kept multiplicity and irrep dimension of the block
    int32_t WS;         // kept-slice width: a work unit owns b in [b0, b0 + WS)
    int32_t pass_begin, pass_end;  // passes of this block
    int32_t unit_begin; // first work unit of this block (units are (block, slice) pairs)
    int32_t pad;
};

struct TpArgs {
    const int32_t *rowptr, *col, *perm;   // CSR over the result nodes; col = gather node; perm = caller's edge id
    int64_t n, E;
    const float *V;       // gathered tensor [n, v_len]
    int32_t v_len, r_len; // row lengths of V and of the result
    const float *sh, *feat;  // [E,S], [E,R] in the caller's edge order
    int32_t S, R, H;
    const float *w1, *b1, *w2, *b2;
    const TpPass* passes;
    const TpBlock* blocks;
    const float* cg;
    int32_t nblocks, nunits, nchunks;
};

struct TpSmem {
    int ldh, oW2, oB2, oHid, oT, oO, oZ, oSh, oFt, oW1, oB1, oI, total;
    __host__ __device__ TpSmem(int H, int R) {
        ldh = H + 4;
        oW2 = 0;                           // [128][ldh]
        oB2 = oW2 + kSliceCols * ldh;      // [128]
        oHid = oB2 + kSliceCols;           // [32][ldh]
        oT = oHid + kTT * ldh;             // [32][132]
        oO = oT + kTT * (kSliceCols + 4);  // [32][80]
        oZ = oO + kTT * kMaxRes;           // [32][25]
        oSh = oZ + kTT * 25;               // [32][16]
        oFt = oSh + kTT * kMaxS;           // [32][16]
        oW1 = oFt + kTT * kMaxR;           // [H][R]
        oB1 = oW1 + H * R;                 // [H]
        oI = oB1 + H;                      // ints: eid[32], gnode[32]
        total = oI + 2 * kTT;
        total = (total + 3) & ~3;
    }
};

// hidden layer of fc for the tile: HID[e][m] = relu(b1[m] + sum_a W1[m][a] feat[e][a])
__device__ __forceinline__ void tp_hidden_tile(float* sm, const TpSmem& L, int H, int R, int cnt) {
    for (int m = threadIdx.x; m < H; m += blockDim.x) {
        float w[kMaxR];
#pragma unroll
        for (int a = 0; a < kMaxR; ++a) w[a] = a < R ? sm[L.oW1 + m * R + a] : 0.f;
        const float b = sm[L.oB1 + m];
        for (int t = 0; t < kTT; ++t) {
            float acc = b;
#pragma unroll
            for (int a = 0; a < kMaxR; ++a)
                if (a < R) acc = fmaf(w[a], sm[L.oFt + t * kMaxR + a], acc);
            sm[L.oHid + t * L.ldh + m] = t < cnt ? fmaxf(acc, 0.f) : 0.f;
        }
    }
}

__device__ __forceinline__ void tp_stage_edges(float* sm, int* smi, const TpSmem& L, const TpArgs& a, int64_t e0, int cnt) {
    if (threadIdx.x < kTT) {
        const int t = threadIdx.x;
        int eid = 0, g = 0;
        if (t < cnt) {
            const int64_t k = e0 + t;
            eid = a.perm ? __ldg(a.perm + k) : (int)k;
            g = __ldg(a.col + k);
        }
        smi[t] = eid;
        smi[kTT + t] = g;
    }
    __syncthreads();
    for (int x = threadIdx.x; x < kTT * kMaxS; x += blockDim.x) {
        const int t = x / kMaxS, j = x - t * kMaxS;
        sm[L.oSh + x] = (t < cnt && j < a.S) ? __ldg(a.sh + (int64_t)smi[t] * a.S + j) : 0.f;
    }
    for (int x = threadIdx.x; x < kTT * kMaxR; x += blockDim.x) {
        const int t = x / kMaxR, r = x - t * kMaxR;
        sm[L.oFt + x] = (t < cnt && r < a.R) ? __ldg(a.feat + (int64_t)smi[t] * a.R + r) : 0.f;
    }
}

__global__ void __launch_bounds__(256, 1) tp_contract_kernel(TpArgs a, float* __restrict__ res) {
    extern __shared__ __align__(16) float sm[];
    const TpSmem L(a.H, a.R);
    int* smi = reinterpret_cast<int*>(sm + L.oI);
    const int H = a.H, R = a.R;
    for (int x = threadIdx.x; x < H * R; x += blockDim.x) sm[L.oW1 + x] = __ldg(a.w1 + x);
    for (int x = threadIdx.x; x < H; x += blockDim.x) sm[L.oB1 + x] = __ldg(a.b1 + x);

    const int tx = threadIdx.x & 15;
    const int eloc = threadIdx.x >> 3, q = threadIdx.x & 7;  // 8 threads per edge in the contraction phase

    for (int item = blockIdx.x; item < a.nunits * a.nchunks; item += gridDim.x) {
        const int unit = item / a.nchunks, chunk = item - unit * a.nchunks;
        // block of this unit
        int bi = 0;
        while (bi + 1 < a.nblocks && a.blocks[bi + 1].unit_begin <= unit) ++bi;
        const TpBlock blk = a.blocks[bi];
        const int b0 = (unit - blk.unit_begin) * blk.WS;
        const int nb = min(blk.WS, blk.MB - b0);
        const int DB = blk.DB;
        // row chunk (whole rows)
        const int r0 = lower_bound_row(a.rowptr, (int)a.n, (a.E * chunk) / a.nchunks);
        const int r1 = (chunk + 1 == a.nchunks) ? (int)a.n : lower_bound_row(a.rowptr, (int)a.n, (a.E * (chunk + 1)) / a.nchunks);
        if (r0 >= r1) continue;
        const int64_t eb = __ldg(a.rowptr + r0), ee = __ldg(a.rowptr + r1);
        if (blk.pass_begin == blk.pass_end) {  // block without any path: zeros
            for (int64_t x = threadIdx.x; x < (int64_t)(r1 - r0) * nb * DB; x += blockDim.x) {
                const int64_t row = r0 + x / (nb * DB);
                const int c = (int)(x % (nb * DB));
                res[row * a.r_len + blk.r_off + b0 * DB + c] = 0.f;
            }
            continue;
        }
        for (int ps = blk.pass_begin; ps < blk.pass_end; ++ps) {
            const TpPass P = a.passes[ps];
            const bool first = ps == blk.pass_begin;
            const int WS = blk.WS;
            __syncthreads();
            // ---- weight slice: column c = a_loc * WS + b  <->  W2 row w_off + (a0+a_loc)*stride_a + (b0+b)*stride_b
            for (int x = threadIdx.x; x < kSliceCols * (H / 4); x += blockDim.x) {
                const int c = x / (H / 4), k4 = x - c * (H / 4);
                const int al = c / WS, b = c - al * WS;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (al < P.MC && b < nb) {
                    const int64_t row = P.w_off + (int64_t)(P.a0 + al) * P.stride_a + (int64_t)(b0 + b) * P.stride_b;
                    v = ldg4(a.w2 + row * H + 4 * k4);
                }
                *reinterpret_cast<float4*>(sm + L.oW2 + c * L.ldh + 4 * k4) = v;
            }
            for (int c = threadIdx.x; c < kSliceCols; c += blockDim.x) {
                const int al = c / WS, b = c - al * WS;
                sm[L.oB2 + c] = (al < P.MC && b < nb)
                                    ? __ldg(a.b2 + P.w_off + (int64_t)(P.a0 + al) * P.stride_a + (int64_t)(b0 + b) * P.stride_b) : 0.f;
            }
            const int DA = P.DA, DS = P.DS;
            // walker state (threads < nb*DB own one result column each)
            int cur = r0;
            int64_t row_end = __ldg(a.rowptr + r0 + 1);
            float acc = 0.f;
            const bool walker = threadIdx.x < nb * DB;
            for (int64_t e0 = eb; e0 < ee; e0 += kTT) {
                const int cnt = (int)min((int64_t)kTT, ee - e0);
                __syncthreads();
                tp_stage_edges(sm, smi, L, a, e0, cnt);
                __syncthreads();
                tp_hidden_tile(sm, L, H, R, cnt);
                // per-edge geometric factor Z[iA][kB] = sum_j sh[j] Zc[iA][j][kB]
                for (int z = q; z < DA * DB; z += 8) {
                    const int iA = z / DB, kB = z - iA * DB;
                    float s = 0.f;
                    for (int j = 0; j < DS; ++j) s = fmaf(sm[L.oSh + eloc * kMaxS + P.sh_off + j], __ldg(a.cg + P.cg_off + (iA * DS + j) * DB + kB), s);
                    sm[L.oZ + eloc * 25 + z] = s;
                }
                __syncthreads();
                {   // T = HID W2s^T + b2s
                    Frag<kSliceCols, kTRPT> f;
                    f.zero();
                    gemm_nt<kSliceCols, kTRPT>(f, sm + L.oHid, L.ldh, sm + L.oW2, L.ldh, H);
#pragma unroll
                    for (int i = 0; i < kTRPT; ++i)
#pragma unroll
                        for (int j = 0; j < Frag<kSliceCols, kTRPT>::CPT; ++j) f.v[i][j] += sm[L.oB2 + tx + 16 * j];
                    store_nt<kSliceCols, kTRPT>(f, sm + L.oT, kSliceCols + 4);
                }
                __syncthreads();
                {   // contraction: 8 threads per edge, thread q takes a_loc = q, q+8, ...
                    float r[16][5];
#pragma unroll
                    for (int b = 0; b < 16; ++b)
#pragma unroll
                        for (int k = 0; k < 5; ++k) r[b][k] = 0.f;
                    if (eloc < cnt) {
                        const float* vrow = a.V + (int64_t)smi[kTT + eloc] * a.v_len + P.v_off;
                        const float* Z = sm + L.oZ + eloc * 25;
                        const float* T = sm + L.oT + eloc * (kSliceCols + 4);
                        for (int al = q; al < P.MC; al += 8) {
                            float y[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                            for (int iA = 0; iA < DA; ++iA) {
                                const float v = __ldg(vrow + (P.a0 + al) * DA + iA);
#pragma unroll
                                for (int k = 0; k < 5; ++k)
                                    if (k < DB) y[k] = fmaf(v, Z[iA * DB + k], y[k]);
                            }
#pragma unroll
                            for (int b = 0; b < 16; ++b)
                                if (b < nb) {
                                    const float t = T[al * WS + b];
#pragma unroll
                                    for (int k = 0; k < 5; ++k)
                                        if (k < DB) r[b][k] = fmaf(t, y[k], r[b][k]);
                                }
                        }
                    }
#pragma unroll
                    for (int b = 0; b < 16; ++b)
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            if (b < nb && k < DB) {  // uniform across the warp: nb, DB are per-unit constants
                                float v = r[b][k];
                                v += __shfl_xor_sync(0xffffffffu, v, 1);
                                v += __shfl_xor_sync(0xffffffffu, v, 2);
                                v += __shfl_xor_sync(0xffffffffu, v, 4);
                                if (q == 0) sm[L.oO + eloc * kMaxRes + b * DB + k] = v;
                            }
                        }
                }
                __syncthreads();
                if (walker) {
                    const int c = threadIdx.x;
                    for (int t = 0; t < cnt; ++t) {
                        const int64_t e = e0 + t;
                        while (e >= row_end) {
                            float* p = res + (int64_t)cur * a.r_len + blk.r_off + b0 * DB + c;
                            *p = first ? acc : *p + acc;
                            acc = 0.f;
                            ++cur;
                            row_end = __ldg(a.rowptr + cur + 1);
                        }
                        acc += sm[L.oO + t * kMaxRes + c];
                    }
                }
            }
            if (walker) {
                while (cur < r1) {
                    float* p = res + (int64_t)cur * a.r_len + blk.r_off + b0 * DB + threadIdx.x;
                    *p = first ? acc : *p + acc;
                    acc = 0.f;
                    ++cur;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: a CTA owns 64 consecutive rows of one path's W2 block and streams all edges
// ------------------------------------------------------------------------------------------------
struct TpWUnit {
    int32_t w_row0;     // first global W2 row of the unit
    int32_t nrows;      // <= 64
    int32_t rel0;       // row index inside the path block (row = u * mul_out + w)
    int32_t mul_out;
    int32_t in_off, DA;   // x[gather][in_off + u*DA + i]
    int32_t out_off, DB;  // g[row node][out_off + w*DB + k]
    int32_t sh_off, DS;
    int32_t cg_off;       // Zc[i][j][k] forward arrangement, coefficient folded in
    int32_t pad;
};

struct TpWArgs {
    const int32_t *rowptr, *col, *perm;  // CSR over the aggregation nodes (rows), col = gather node
    int64_t n, E;
    const float *x, *g;   // x [n, x_len] gathered at col; g [n, g_len] read at the row node
    int32_t x_len, g_len;
    const float *sh, *feat;
    int32_t S, R, H;
    const float *w1, *b1, *w2;
    const TpWUnit* units;
    const float* cg;
    int32_t nunits;
};

template <int H>
struct TpWSmem {
    static constexpr int ldh = H + 4;
    static constexpr int oW2 = 0;                    // [64][ldh]
    static constexpr int oHid = oW2 + 64 * ldh;      // [32][ldh]
    static constexpr int oDh = oHid + kTT * ldh;     // [32][ldh]
    static constexpr int oDT = oDh + kTT * ldh;      // [32][68]
    static constexpr int oZ = oDT + kTT * 68;        // [32][25]
    static constexpr int oSh = oZ + kTT * 25;        // [32][16]
    static constexpr int oFt = oSh + kTT * kMaxS;    // [32][16]
    static constexpr int oW1 = oFt + kTT * kMaxR;    // [H][kMaxR]
    static constexpr int oB1 = oW1 + H * kMaxR;      // [H]
    static constexpr int oI = oB1 + H;               // eid[32], gnode[32], rownode[32]
    static constexpr int total = oI + 3 * kTT;
};

template <int H>
__global__ void __launch_bounds__(256, 1)
tp_wgrad_kernel(TpWArgs a, float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ parts) {
    using S = TpWSmem<H>;
    constexpr int ldh = S::ldh;
    constexpr int NT = H / 16;
    extern __shared__ __align__(16) float sm[];
    int* smi = reinterpret_cast<int*>(sm + S::oI);
    const int R = a.R;
    for (int x = threadIdx.x; x < H * kMaxR; x += blockDim.x) {
        const int m = x / kMaxR, r = x - m * kMaxR;
        sm[S::oW1 + x] = r < R ? __ldg(a.w1 + m * R + r) : 0.f;
    }
    for (int x = threadIdx.x; x < H; x += blockDim.x) sm[S::oB1 + x] = __ldg(a.b1 + x);
    const int eloc = threadIdx.x >> 3, q = threadIdx.x & 7;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    for (int ui = blockIdx.x; ui < a.nunits; ui += gridDim.x) {
        const TpWUnit U = a.units[ui];
        __syncthreads();
        for (int x = threadIdx.x; x < 64 * (H / 4); x += blockDim.x) {
            const int c = x / (H / 4), k4 = x - c * (H / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < U.nrows) v = ldg4(a.w2 + (int64_t)(U.w_row0 + c) * H + 4 * k4);
            *reinterpret_cast<float4*>(sm + S::oW2 + c * ldh + 4 * k4) = v;
        }
        float accW[4][NT];   // dW2[c = ty+16i][m = tx+16j]
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) accW[i][j] = 0.f;
        float accb2 = 0.f;            // threads < 64: db2[c]
        float accW1[kMaxR], accb1 = 0.f;  // threads < H: dW1[m][:], db1[m]
#pragma unroll
        for (int r = 0; r < kMaxR; ++r) accW1[r] = 0.f;
        const int DA = U.DA, DB = U.DB, DS = U.DS;

        for (int64_t e0 = 0; e0 < a.E; e0 += kTT) {
            const int cnt = (int)min((int64_t)kTT, a.E - e0);
            __syncthreads();
            if (threadIdx.x < kTT) {
                const int t = threadIdx.x;
                int eid = 0, gn = 0, rn = 0;
                if (t < cnt) {
                    const int64_t k = e0 + t;
                    eid = a.perm ? __ldg(a.perm + k) : (int)k;
                    gn = __ldg(a.col + k);
                    int lo = 0, hi = (int)a.n;  // row node of edge k
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if ((int64_t)__ldg(a.rowptr + mid) <= k) lo = mid; else hi = mid;
                    }
                    rn = lo;
                }
                smi[t] = eid; smi[kTT + t] = gn; smi[2 * kTT + t] = rn;
            }
            __syncthreads();
            for (int x = threadIdx.x; x < kTT * kMaxS; x += blockDim.x) {
                const int t = x / kMaxS, j = x - t * kMaxS;
                sm[S::oSh + x] = (t < cnt && j < a.S) ? __ldg(a.sh + (int64_t)smi[t] * a.S + j) : 0.f;
            }
            for (int x = threadIdx.x; x < kTT * kMaxR; x += blockDim.x) {
                const int t = x / kMaxR, r = x - t * kMaxR;
                sm[S::oFt + x] = (t < cnt && r < R) ? __ldg(a.feat + (int64_t)smi[t] * R + r) : 0.f;
            }
            __syncthreads();
            // hidden tile
            for (int m = threadIdx.x; m < H; m += blockDim.x) {
                const float b = sm[S::oB1 + m];
                for (int t = 0; t < kTT; ++t) {
                    float acc = b;
#pragma unroll
                    for (int r = 0; r < kMaxR; ++r) acc = fmaf(sm[S::oW1 + m * kMaxR + r], sm[S::oFt + t * kMaxR + r], acc);
                    sm[S::oHid + t * ldh + m] = t < cnt ? fmaxf(acc, 0.f) : 0.f;
                }
            }
            for (int z = q; z < DA * DB; z += 8) {
                const int iA = z / DB, kB = z - iA * DB;
                float s = 0.f;
                for (int j = 0; j < DS; ++j) s = fmaf(sm[S::oSh + eloc * kMaxS + U.sh_off + j], __ldg(a.cg + U.cg_off + (iA * DS + j) * DB + kB), s);
                sm[S::oZ + eloc * 25 + z] = s;
            }
            __syncthreads();
            // dT[e][c] = sum_k g[row_e][out_off + w*DB + k] * ( sum_i x[gather_e][in_off + u*DA + i] Z[i][k] )
            for (int c = q; c < 64; c += 8) {
                float v = 0.f;
                if (eloc < cnt && c < U.nrows) {
                    const int rel = U.rel0 + c, u = rel / U.mul_out, w = rel - u * U.mul_out;
                    const float* xr = a.x + (int64_t)smi[kTT + eloc] * a.x_len + U.in_off + u * DA;
                    const float* gr = a.g + (int64_t)smi[2 * kTT + eloc] * a.g_len + U.out_off + w * DB;
                    const float* Z = sm + S::oZ + eloc * 25;
                    for (int k = 0; k < DB; ++k) {
                        float y = 0.f;
                        for (int i = 0; i < DA; ++i) y = fmaf(__ldg(xr + i), Z[i * DB + k], y);
                        v = fmaf(__ldg(gr + k), y, v);
                    }
                }
                sm[S::oDT + eloc * 68 + c] = v;
            }
            __syncthreads();
            wgrad_tn<4, NT>(accW, sm + S::oDT, 68, sm + S::oHid, ldh, cnt);
            if (threadIdx.x < 64) {
                float s = 0.f;
                for (int t = 0; t < cnt; ++t) s += sm[S::oDT + t * 68 + threadIdx.x];
                accb2 += s;
            }
            {   // dhid = dT W2s  -> Dh tile
                Frag<H, kTRPT> f;
                f.zero();
                gemm_nn<H, kTRPT>(f, sm + S::oDT, 68, sm + S::oW2, ldh, 64);
                store_nn<H, kTRPT>(f, sm + S::oDh, ldh);
            }
            __syncthreads();
            for (int m = threadIdx.x; m < H; m += blockDim.x) {
                for (int t = 0; t < cnt; ++t) {
                    const float d = sm[S::oHid + t * ldh + m] > 0.f ? sm[S::oDh + t * ldh + m] : 0.f;
                    accb1 += d;
#pragma unroll
                    for (int r = 0; r < kMaxR; ++r) accW1[r] = fmaf(d, sm[S::oFt + t * kMaxR + r], accW1[r]);
                }
            }
        }
        // results of this unit
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int c = ty + 16 * i;
                if (c < U.nrows) dW2[(int64_t)(U.w_row0 + c) * H + tx + 16 * j] = accW[i][j];
            }
        if (threadIdx.x < 64 && threadIdx.x < U.nrows) db2[U.w_row0 + threadIdx.x] = accb2;
        float* my = parts + (int64_t)ui * (H * kMaxR + H);
        for (int m = threadIdx.x; m < H; m += blockDim.x) {  // (one m per thread when H <= 256)
#pragma unroll
            for (int r = 0; r < kMaxR; ++r) my[m * kMaxR + r] = accW1[r];
            my[H * kMaxR + m] = accb1;
        }
    }
}

}  // namespace gmp

using namespace gmp;

extern "C" {

int64_t gmp_tp_contract_smem_bytes(int32_t H, int32_t R) { return (int64_t)TpSmem(H, R).total * sizeof(float); }

int gmp_tp_contract(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t num_edges,
                    const float* V, int32_t v_len, float* res, int32_t r_len, const float* edge_sh, int32_t S,
                    const float* edge_feat, int32_t R, const float* w1, const float* b1, const float* w2, const float* b2,
                    int32_t H, const void* passes, const void* blocks, int32_t nblocks, int32_t nunits, const float* cg,
                    int32_t precision, gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && res && passes && blocks && cg && w1 && b1 && w2 && b2, "tp_contract: NULL pointer");
    GMP_REQUIRE(num_edges == 0 || (col && V && edge_sh && edge_feat), "tp_contract: NULL edge/feature pointer");
    GMP_REQUIRE(H >= 4 && H <= 256 && H % 4 == 0, "tp_contract: mlp_dim must be a multiple of 4 in [4, 256] (got %d)", H);
    GMP_REQUIRE(R >= 1 && R <= kMaxR && S >= 1 && S <= kMaxS, "tp_contract: edge_feats_dim <= %d and sh dim <= %d", kMaxR, kMaxS);
    GMP_REQUIRE(n >= 0 && n < (1ll << 31) && num_edges >= 0 && num_edges < (1ll << 31), "tp_contract: sizes out of range");
    if (precision != GMP_FP32_STRICT) {
        set_error("tp_contract: precision mode %d is not built into this library", precision);
        return GMP_ERR_UNSUPPORTED;
    }
    if (n == 0 || nunits == 0) return GMP_OK;
    TpArgs a;
    a.rowptr = rowptr; a.col = col; a.perm = perm; a.n = n; a.E = num_edges; a.V = V; a.v_len = v_len; a.r_len = r_len;
    a.sh = edge_sh; a.feat = edge_feat; a.S = S; a.R = R; a.H = H; a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2;
    a.passes = (const TpPass*)passes; a.blocks = (const TpBlock*)blocks; a.cg = cg; a.nblocks = nblocks; a.nunits = nunits;
    // enough row chunks to give every SM a few work items, at least ~4096 edges per item
    int64_t want = ceil_div(4ll * num_sms(), nunits);
    int64_t cap = num_edges / 4096 + 1;
    a.nchunks = (int)(want < cap ? want : cap);
    if (a.nchunks < 1) a.nchunks = 1;
    const size_t smem = (size_t)TpSmem(H, R).total * sizeof(float);
    GMP_CUDA(cudaFuncSetAttribute(tp_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t items = (int64_t)nunits * a.nchunks;
    const int grid = (int)(items < 4ll * num_sms() ? items : 4ll * num_sms());
    tp_contract_kernel<<<grid, 256, smem, stream>>>(a, res);
    return check_launch("tp_contract_kernel");
}

int64_t gmp_tp_wgrad_part_len(int32_t H) { return (int64_t)H * kMaxR + H; }

int gmp_tp_wgrad(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t num_edges,
                 const float* x, int32_t x_len, const float* g, int32_t g_len, const float* edge_sh, int32_t S,
                 const float* edge_feat, int32_t R, const float* w1, const float* b1, const float* w2, int32_t H,
                 const void* units, int32_t nunits, const float* cg, float* dW2, float* db2, float* w1_parts,
                 int32_t precision, gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && units && cg && w1 && b1 && w2 && dW2 && db2 && w1_parts && g, "tp_wgrad: NULL pointer");
    GMP_REQUIRE(num_edges == 0 || (col && x && edge_sh && edge_feat), "tp_wgrad: NULL edge/feature pointer");
    GMP_REQUIRE(H == 64 || H == 128 || H == 256, "tp_wgrad: mlp_dim must be 64, 128 or 256 (got %d)", H);
    GMP_REQUIRE(R >= 1 && R <= kMaxR && S >= 1 && S <= kMaxS, "tp_wgrad: edge_feats_dim <= %d and sh dim <= %d", kMaxR, kMaxS);
    if (precision != GMP_FP32_STRICT) {
        set_error("tp_wgrad: precision mode %d is not built into this library", precision);
        return GMP_ERR_UNSUPPORTED;
    }
    if (nunits == 0) return GMP_OK;
    TpWArgs a;
    a.rowptr = rowptr; a.col = col; a.perm = perm; a.n = n; a.E = num_edges; a.x = x; a.g = g; a.x_len = x_len; a.g_len = g_len;
    a.sh = edge_sh; a.feat = edge_feat; a.S = S; a.R = R; a.H = H; a.w1 = w1; a.b1 = b1; a.w2 = w2;
    a.units = (const TpWUnit*)units; a.cg = cg; a.nunits = nunits;
    const int grid = nunits;
#define GMP_TPW(H_)                                                                                                   \
    {                                                                                                                 \
        const size_t smem = (size_t)TpWSmem<H_>::total * sizeof(float);                                               \
        GMP_CUDA(cudaFuncSetAttribute(tp_wgrad_kernel<H_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        tp_wgrad_kernel<H_><<<grid, 256, smem, stream>>>(a, dW2, db2, w1_parts);                                     \
    }
    if (H == 256) GMP_TPW(256) else if (H == 128) GMP_TPW(128) else GMP_TPW(64)
#undef GMP_TPW
    return check_launch("tp_wgrad_kernel");
}

}  // extern "C"
