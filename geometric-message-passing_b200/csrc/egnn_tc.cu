// EGNN edge path on the 5th-generation tensor cores (GMP_BF16_TC), models/layers/egnn_layer.py:62-80, emb_dim = 128.
//
// Same mathematics and work decomposition as egnn.cu (pre1 = P[i] + Q[j] + dist * wd, three LayerNorm'd stages, whole CSR
// rows per work range, two recompute passes backward), with 128-edge tiles and both 128 x 128 edge GEMMs -- and in the
// backward their data- and weight-gradient GEMMs -- on tcgen05.mma with fp32 accumulators in tensor memory:
//   a thread owns (edge row e = TMEM lane, column quarter hf): it reads its 32 pre-activations straight out of tensor
//   memory, LayerNorm statistics are completed by exchanging two partial sums with the threads that own the other quarters,
//   and the next operand (bf16, 128B-swizzled K-major image, row = edge) is written in place for the next MMA.
// The gathered operand (Q[j] in the dst pass, P[i] in the src pass) is read as bf16 rows with cp.async (256 B per edge);
// the row-constant operand is read in fp32.  bf16 operands / fp32 accumulation: 1e-2 relative (tests/test_gpu_tc.py).
#include <cuda_pipeline.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gmp {

using namespace tc;

constexpr int kGF = 128;         // emb_dim
constexpr int kGT = 128;         // edges per tile
constexpr int kGRangeMax = 2048; // edges per work range (whole rows) on large graphs; smaller graphs get shorter ranges (two per SM,
                                 // at least one 128-edge tile) so that they still spread over the machine
constexpr int kGQ = 4;          // column quarters: a thread owns (edge row, 32 columns); 16 warps per CTA
constexpr int kGCW = 32;        // columns per thread
constexpr int kGThreads = 512;
constexpr int kGLd = 272;        // byte stride of a bf16 row in the gather tile (conflict-free 16-byte chunks per row)

struct EgnnTcArgs {
    const int32_t *rowptr, *col, *rowid, *deg_rowptr, *perm;   // perm: caller's edge id of sorted edge k (NULL = identity)
    int64_t n, E;
    const float* rowt;              // [n,128] fp32, indexed by the CSR row: P (dst pass) / Q (src pass)
    const __nv_bfloat16* gath;      // [n,128] bf16, indexed by col:       Q (dst pass) / P (src pass)
    PeerRows peer;                  // fused dst pass on a partitioned graph: halo rows of Q in the neighbours' memory
    const float* pos;
    const float *wd, *g1, *be1, *w1, *b1, *g2, *be2, *w2, *b2, *g3, *be3, *w3, *b3;
    int aggr_mean, nranges, grange;
    float eps;
    // fused single-pass backward: per-edge d(pre1) (bf16 rows) and d(delta) (float4), indexed by the caller's edge id
    __nv_bfloat16* dpre_out;
    float* ddelta_out;
};

// shared-memory map (bytes from the 1024-aligned base)
constexpr int oGW1 = 0;                        // W1 image  [128 out][2 slabs of 64 in] bf16
constexpr int oGW2 = 32768;                    // W2 image
constexpr int oGA1 = 65536;                    // a1  (A operand of GEMM 1; B operand of the dW1 GEMM)
constexpr int oGM = 98304;                     // m   (A operand of GEMM 2; walker input; B operand of the dW2 GEMM)
constexpr int oGDT = 131072;                   // dpre3 / dpre2 (backward A operands)
constexpr int oGG = 163840;                    // gathered rows -> xhat1 -> dpre1, bf16 [128][136]
constexpr int oGOnes = oGG + kGT * kGLd;       // 4 KB of bf16 ones (column sums through the tensor core)
constexpr int oGVec = oGOnes + 4096;           // wd g1 be1 b1 g2 be2 b2 g3 be3 w3
constexpr int oGSc = oGVec + 10 * kGF * 4;     // per-edge scalars [16][128]
constexpr int oGInt = oGSc + 16 * kGT * 4;     // rrow[128] gcol[128] inode[128] eid[128]
constexpr int oGRed = oGInt + 4 * kGT * 4;     // partial-sum exchange: 2 buffers x [4 quarters][128] float2
constexpr int oGBar = oGRed + 2 * kGQ * kGT * 8; // mbarrier + tmem pointer
constexpr int kEgnnTcSmem = oGBar + 64 + 1024;

enum { TV_WD = 0, TV_G1, TV_BE1, TV_B1, TV_G2, TV_BE2, TV_B2, TV_G3, TV_BE3, TV_W3 };
enum { TS_DX = 0, TS_DY, TS_DZ, TS_DIST, TS_S, TS_R1, TS_M2, TS_R2, TS_M3, TS_R3, TS_GS, TS_SCALE, TS_GX, TS_GY, TS_GZ };

template <int ACT> __device__ __forceinline__ float tact(float y) {
    if (ACT == 0) return fmaxf(y, 0.f);
    return y / (1.f + __expf(-y));
}
template <int ACT> __device__ __forceinline__ float tdact(float y) {
    if (ACT == 0) return y > 0.f ? 1.f : 0.f;
    const float s = 1.f / (1.f + __expf(-y));
    return s * (1.f + y * (1.f - s));
}

struct GCtx {
    uint8_t* sm;
    float *vec, *sc;
    int* ints;
    uint64_t* bar;
    uint32_t tm, lane_base, ph, xb;  // TMEM base, this warp's lane quarter, mbarrier phase, exchange buffer parity
    int t, warp, e, hf;
};

// complete a per-row partial (two floats per quarter) with the threads that own the other column quarters
__device__ __forceinline__ float2 gx_exchange(GCtx& c, float a, float b) {
    float2* red = reinterpret_cast<float2*>(c.sm + oGRed) + (c.xb & 1u) * kGQ * kGT;
    red[c.hf * kGT + c.e] = make_float2(a, b);
    bar_sync_named(1 + (c.warp & 3), 32 * kGQ);  // only the four warps that share these 32 rows (w, w+4, w+8, w+12) meet here
    float2 tot = make_float2(0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kGQ; ++p) {  // fixed order: every quarter computes the same bits
        const float2 o = red[p * kGT + c.e];
        tot.x += o.x;
        tot.y += o.y;
    }
    ++c.xb;
    return tot;
}

__device__ __forceinline__ void unpack8(const uint4 u, float (&f)[8]) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
// eight consecutive parameter values (32-byte aligned) with two 16-byte shared-memory loads
struct V8 { float v[8]; };
__device__ __forceinline__ V8 ldv8(const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    return V8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// issue one 128 x N x 128 product (K-major operands) and wait for it
__device__ __forceinline__ void gx_mma_wait(GCtx& c) {
    mbar_wait(c.bar, c.ph);
    c.ph ^= 1u;
    tc_fence_after();
}

// this thread's 32 accumulator columns: [32*hf, 32*hf + 32) of the 128-column block at `tcol`
__device__ __forceinline__ void gx_ld64(const GCtx& c, uint32_t tcol, float (&v)[kGCW]) {
    tmem_ld32(c.tm + c.lane_base + tcol + kGCW * c.hf, v);
}

// LayerNorm statistics of the row from this quarter's 32 values (single pass, completed across the quarters)
__device__ __forceinline__ void gx_stats(GCtx& c, const float (&x)[kGCW], float eps, float& mean, float& rstd) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < kGCW; ++k) { s += x[k]; q = fmaf(x[k], x[k], q); }
    const float2 tot = gx_exchange(c, s, q);
    mean = tot.x * (1.f / kGF);
    const float var = fmaxf(tot.y * (1.f / kGF) - mean * mean, 0.f);
    rstd = 1.f / sqrtf(var + eps);
}

// Forward of one tile.  On exit: G = xhat1 (bf16), A1 = a1, M = m, D1 / D2 (tensor memory columns 0 / 128) = the raw
// outputs of the two GEMMs (biases not yet added), scalars: geometry, s, LayerNorm statistics.
// SRC: CSR rows are the source nodes j and col holds the destination i.
template <int ACT, bool SRC>
__device__ __forceinline__ void egnn_tc_tile_forward(GCtx& c, const EgnnTcArgs& a, int64_t e0, int cnt) {
    uint8_t* sm = c.sm;
    float* sc = c.sc;
    const float* vec = c.vec;
    const int t = c.t, e = c.e, hf = c.hf;
    if (t < kGT) {
        const int tt = min(t, cnt - 1);
        const int64_t k = e0 + tt;
        const int rr = __ldg(a.rowid + k), gc = __ldg(a.col + k);
        const int i = SRC ? gc : rr, j = SRC ? rr : gc;
        float dx = 0.f, dy = 0.f, dz = 0.f, dist = 0.f;
        if (t < cnt) {
            dx = __ldg(a.pos + 3 * (int64_t)i) - __ldg(a.pos + 3 * (int64_t)j);
            dy = __ldg(a.pos + 3 * (int64_t)i + 1) - __ldg(a.pos + 3 * (int64_t)j + 1);
            dz = __ldg(a.pos + 3 * (int64_t)i + 2) - __ldg(a.pos + 3 * (int64_t)j + 2);
            dist = sqrtf(dx * dx + dy * dy + dz * dz);
        }
        c.ints[t] = rr;
        c.ints[kGT + t] = gc;
        c.ints[2 * kGT + t] = i;
        c.ints[3 * kGT + t] = a.perm ? __ldg(a.perm + k) : (int)k;
        sc[TS_DX * kGT + t] = dx; sc[TS_DY * kGT + t] = dy; sc[TS_DZ * kGT + t] = dz; sc[TS_DIST * kGT + t] = dist;
    }
    __syncthreads();
    // ---- gather: bf16 rows of the col-side operand
    for (int x = t; x < kGT * 16; x += kGThreads) {
        const int r = x >> 4, ch = x & 15;
        uint8_t* dst = sm + oGG + r * kGLd + ch * 16;
        if (r < cnt) __pipeline_memcpy_async(dst, peer_row(a.gath, a.peer, c.ints[kGT + r], kGF * 2) + ch * 16, 16);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
    // ---- stage 1: pre1 = gathered + row operand + dist * wd -> LN1 -> xhat1 (G, in place), a1 (A1)
    {
        float x[kGCW];
        const float dist = sc[TS_DIST * kGT + e];
        const float4* rp = reinterpret_cast<const float4*>(a.rowt + (int64_t)c.ints[e] * kGF + kGCW * hf);
        const float4* wd4 = reinterpret_cast<const float4*>(vec + TV_WD * kGF + kGCW * hf);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float g[8];
            unpack8(*reinterpret_cast<const uint4*>(sm + oGG + e * kGLd + (4 * hf + q) * 16), g);
            const float4 r0 = __ldg(rp + 2 * q), r1 = __ldg(rp + 2 * q + 1), w0 = wd4[2 * q], w1 = wd4[2 * q + 1];
            x[8 * q + 0] = g[0] + r0.x + dist * w0.x; x[8 * q + 1] = g[1] + r0.y + dist * w0.y;
            x[8 * q + 2] = g[2] + r0.z + dist * w0.z; x[8 * q + 3] = g[3] + r0.w + dist * w0.w;
            x[8 * q + 4] = g[4] + r1.x + dist * w1.x; x[8 * q + 5] = g[5] + r1.y + dist * w1.y;
            x[8 * q + 6] = g[6] + r1.z + dist * w1.z; x[8 * q + 7] = g[7] + r1.w + dist * w1.w;
        }
        float mean, rstd;
        gx_stats(c, x, a.eps, mean, rstd);
        if (hf == 0) sc[TS_R1 * kGT + e] = rstd;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const V8 pG1 = ldv8(vec + TV_G1 * kGF + kGCW * hf + 8 * q);
            const V8 pE1 = ldv8(vec + TV_BE1 * kGF + kGCW * hf + 8 * q);
            float xh[8], a1[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int col = kGCW * hf + 8 * q + u;
                xh[u] = (x[8 * q + u] - mean) * rstd;
                a1[u] = tact<ACT>(fmaf(xh[u], pG1.v[u], pE1.v[u]));
            }
            *reinterpret_cast<uint4*>(sm + oGG + e * kGLd + (4 * hf + q) * 16) = pack8(xh);
            *reinterpret_cast<uint4*>(sm + oGA1 + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(a1);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    if (c.warp == 0) {  // GEMM 1: D1 = a1 W1^T
        tc_fence_after();
        if (elect_one()) {
            umma_tile(c.tm, smem_u32(sm + oGA1), 16384, smem_u32(sm + oGW1), 16384, 128, idesc);
            umma_commit(c.bar);
        }
        __syncwarp();
    }
    gx_mma_wait(c);
    // ---- stage 2: pre2 = D1 + b1 -> LN2 -> m (M)
    {
        float x[kGCW];
        gx_ld64(c, 0, x);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const V8 pb = ldv8(vec + TV_B1 * kGF + kGCW * hf + 8 * q);
#pragma unroll
            for (int u = 0; u < 8; ++u) x[8 * q + u] += pb.v[u];
        }
        float mean, rstd;
        gx_stats(c, x, a.eps, mean, rstd);
        if (hf == 0) { sc[TS_M2 * kGT + e] = mean; sc[TS_R2 * kGT + e] = rstd; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const V8 pG2 = ldv8(vec + TV_G2 * kGF + kGCW * hf + 8 * q);
            const V8 pE2 = ldv8(vec + TV_BE2 * kGF + kGCW * hf + 8 * q);
            float m[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int col = kGCW * hf + 8 * q + u;
                m[u] = tact<ACT>(fmaf((x[8 * q + u] - mean) * rstd, pG2.v[u], pE2.v[u]));
            }
            *reinterpret_cast<uint4*>(sm + oGM + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(m);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (c.warp == 0) {  // GEMM 2: D2 = m W2^T
        tc_fence_after();
        if (elect_one()) {
            umma_tile(c.tm + 128, smem_u32(sm + oGM), 16384, smem_u32(sm + oGW2), 16384, 128, idesc);
            umma_commit(c.bar);
        }
        __syncwarp();
    }
    gx_mma_wait(c);
    // ---- stage 3: pre3 = D2 + b2 -> LN3 -> s = act(.) . w3 + b3
    {
        float x[kGCW];
        gx_ld64(c, 128, x);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const V8 pb = ldv8(vec + TV_B2 * kGF + kGCW * hf + 8 * q);
#pragma unroll
            for (int u = 0; u < 8; ++u) x[8 * q + u] += pb.v[u];
        }
        float mean, rstd;
        gx_stats(c, x, a.eps, mean, rstd);
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const V8 pg = ldv8(vec + TV_G3 * kGF + kGCW * hf + 8 * q), pe = ldv8(vec + TV_BE3 * kGF + kGCW * hf + 8 * q),
                     pw = ldv8(vec + TV_W3 * kGF + kGCW * hf + 8 * q);
#pragma unroll
            for (int u = 0; u < 8; ++u) dot = fmaf(tact<ACT>(fmaf((x[8 * q + u] - mean) * rstd, pg.v[u], pe.v[u])), pw.v[u], dot);
        }
        const float2 tot = gx_exchange(c, dot, 0.f);
        if (hf == 0) {
            sc[TS_M3 * kGT + e] = mean;
            sc[TS_R3 * kGT + e] = rstd;
            sc[TS_S * kGT + e] = tot.x + __ldg(a.b3);
        }
    }
    tc_fence_before();
    __syncthreads();
}

// weights -> bf16 swizzled K-major images, vectors, barrier, tensor memory
template <int NCOLS>
__device__ __forceinline__ void egnn_tc_setup(GCtx& c, const EgnnTcArgs& a, uint8_t* sm) {
    const int t = threadIdx.x;
    for (int x = t; x < kGF * 16; x += kGThreads) {
        const int f = x >> 4, ch16 = x & 15, kb = ch16 >> 3, ch = ch16 & 7;
        {
            const float4 lo = ldg4(a.w1 + f * kGF + kb * 64 + ch * 8), hi = ldg4(a.w1 + f * kGF + kb * 64 + ch * 8 + 4);
            *reinterpret_cast<uint4*>(sm + oGW1 + kb * 16384 + sw128_chunk_off(f, ch)) =
                make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
        }
        {
            const float4 lo = ldg4(a.w2 + f * kGF + kb * 64 + ch * 8), hi = ldg4(a.w2 + f * kGF + kb * 64 + ch * 8 + 4);
            *reinterpret_cast<uint4*>(sm + oGW2 + kb * 16384 + sw128_chunk_off(f, ch)) =
                make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
        }
    }
    float* vec = reinterpret_cast<float*>(sm + oGVec);
    const float* vecs[10] = {a.wd, a.g1, a.be1, a.b1, a.g2, a.be2, a.b2, a.g3, a.be3, a.w3};
    for (int i = t; i < 10 * kGF; i += kGThreads) vec[i] = __ldg(vecs[i / kGF] + (i % kGF));
    for (int i = t; i < 2048; i += kGThreads) reinterpret_cast<uint16_t*>(sm + oGOnes)[i] = 0x3f80;  // bf16 1.0
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + oGBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + oGBar + 16);
    if (t == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if ((t >> 5) == 0) tmem_alloc<NCOLS>(tmem_ptr);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.sm = sm;
    c.vec = vec;
    c.sc = reinterpret_cast<float*>(sm + oGSc);
    c.ints = reinterpret_cast<int*>(sm + oGInt);
    c.bar = bar;
    c.tm = *tmem_ptr;
    c.t = t;
    c.warp = t >> 5;
    c.e = (c.warp & 3) * 32 + (t & 31);
    c.hf = c.warp >> 2;
    c.lane_base = (uint32_t)((c.warp & 3) * 32) << 16;
    c.ph = 0;
    c.xb = 0;
}

// bf16 element (row r, column col) of a swizzled operand image
__device__ __forceinline__ float img_at(const uint8_t* img, int r, int col) {
    const uint16_t h = *reinterpret_cast<const uint16_t*>(img + (col >> 6) * 16384 + sw128_chunk_off(r, (col & 63) >> 3) + (col & 7) * 2);
    return __uint_as_float((uint32_t)h << 16);
}

// Segmented sums of one value column over the tile's edge slots, in groups of 8: a plain tree add while the group stays in
// the open row (the common case), edge by edge across a row change.  rid = CSR row of every slot; get(k) = value of slot k;
// flush(row, sum) stores a finished row (rows without edges are flushed with 0).
template <class Get, class Flush>
__device__ __forceinline__ void gx_walk(const int* __restrict__ rid, int cnt, int& cur, float& acc, Get get, Flush flush) {
    for (int k0 = 0; k0 < cnt; k0 += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (k0 + j < cnt) ? get(k0 + j) : 0.f;
        if (rid[min(k0 + 7, cnt - 1)] == cur) {
            acc += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (k0 + j < cnt) {
                    const int r = rid[k0 + j];
                    if (r != cur) {
                        flush(cur, acc);
                        for (int z = cur + 1; z < r; ++z) flush(z, 0.f);
                        acc = 0.f;
                        cur = r;
                    }
                    acc += v[j];
                }
            }
        }
    }
}

template <int ACT>
__global__ void __launch_bounds__(kGThreads, 1) egnn_fwd_tc_kernel(EgnnTcArgs a, float* __restrict__ msg_aggr, float* __restrict__ pos_aggr) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    GCtx c;
    egnn_tc_setup<256>(c, a, sm);
    const int t = c.t;
    for (int rg = blockIdx.x; rg < a.nranges; rg += gridDim.x) {
        const int r0 = lower_bound_row(a.rowptr, (int)a.n, (int64_t)rg * a.grange);
        const int r1 = (rg + 1 == a.nranges) ? (int)a.n : lower_bound_row(a.rowptr, (int)a.n, (int64_t)(rg + 1) * a.grange);
        if (r0 >= r1) continue;
        const int64_t eb = __ldg(a.rowptr + r0), ee = __ldg(a.rowptr + r1);
        int cur = r0;
        float acc = 0.f;
        const int role = t < kGF ? 0 : (t < kGF + 3 ? 1 : 2);  // 0: feature column, 1: coordinate, 2: idle
        const int comp = t - kGF;
        auto flush = [&](int row, float v) {
            const float deg = (float)(__ldg(a.rowptr + row + 1) - __ldg(a.rowptr + row));
            if (role == 0) msg_aggr[(int64_t)row * kGF + t] = (a.aggr_mean && deg > 0.f) ? v / deg : v;
            else pos_aggr[(int64_t)row * 3 + comp] = deg > 0.f ? v / deg : 0.f;
        };
        for (int64_t e0 = eb; e0 < ee; e0 += kGT) {
            const int cnt = (int)min((int64_t)kGT, ee - e0);
            egnn_tc_tile_forward<ACT, false>(c, a, e0, cnt);
            if (role == 0) {
                // element (slot k, column t) of the swizzled m image: chunk index XOR (k & 7), 128 B per slot
                const uint8_t* base = sm + oGM + (t >> 6) * 16384 + (t & 7) * 2;
                const int ch = (t & 63) >> 3;
                gx_walk(c.ints, cnt, cur, acc,
                        [&](int k) { return __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(base + k * 128 + ((ch ^ (k & 7)) << 4))) << 16); },
                        flush);
            } else if (role == 1) {
                gx_walk(c.ints, cnt, cur, acc, [&](int k) { return c.sc[(TS_DX + comp) * kGT + k] * c.sc[TS_S * kGT + k]; }, flush);
            }
            __syncthreads();  // the tile buffers are free again
        }
        if (role != 2) {
            flush(cur, acc);
            for (int z = cur + 1; z < r1; ++z) flush(z, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (c.warp == 0) tmem_dealloc<256>(c.tm);
}

// Column sums over the 32 rows of a warp for 8 columns held per thread: a halving butterfly (lane bits 0-2 select the
// column, 7 shuffles) plus two more steps across the four lane groups.  Returns, on every lane, the sum of column
// (lane & 7) of this chunk; `v` is destroyed.
__device__ __forceinline__ float gx_bfly8(float (&v)[8], int lane) {
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    float r = v[0];
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 16);
    return r;
}
// chunk q (8 columns) of a thread's 32: lane l keeps the sum for column 8 * (l >> 3) + (l & 7) = l of its quarter
#define GX_VACC(slot, arr, q) { const float r_ = gx_bfly8(arr, lane); if ((lane >> 3) == (q)) vacc[slot] += r_; }

// ------------------------------------------------------------------------------------------------
// backward: recompute the tile, then LN3 / LN2 / LN1 backward with the data-gradient GEMMs on the tensor cores.
//   dst pass (SRC = false): d_node = dL/dP, d_pos = the pos_i part, and the weight gradients dW1 / dW2 accumulated in
//                           tensor memory over the whole kernel (dpre^T a1, dpre^T m: MN-major operands = the same images)
//   src pass (SRC = true) : d_node = dL/dQ, d_pos = the pos_j part, and the ten vector gradients as column sums through
//                           the tensor core (per-edge tile^T x ones, N = 16), also accumulated in tensor memory
// Tensor memory: D1 @0 (later da1), D2 @128 (later dm), @256 dW1 | vectors, @384 dW2.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gx_issue_begin(GCtx& c) {
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
}

// D[128 x 128] = A (K-major image, rows = edges) x W (image [out][in] read as [K = out][N = in])
__device__ __forceinline__ void gx_dgrad(const GCtx& c, uint32_t dcol, int a_off, int w_off) {
    const uint32_t idesc = umma_idesc_bf16(128, 128, false, true);
    const uint32_t ab = smem_u32(c.sm + a_off), wb = smem_u32(c.sm + w_off);
#pragma unroll
    for (int k16 = 0; k16 < 8; ++k16)
        umma_bf16(c.tm + dcol, umma_desc_k128(ab + (k16 >> 2) * 16384 + (k16 & 3) * 32), umma_desc_mn128(wb + k16 * 2048, 16384), idesc,
                  k16 ? 1u : 0u);
}
// D[128 out x 128 in] (+)= A^T B over the 128 edges of the tile (both images rows = edges)
__device__ __forceinline__ void gx_wgrad(const GCtx& c, uint32_t dcol, int a_off, int b_off, bool accumulate) {
    umma_tile_mn(c.tm + dcol, smem_u32(c.sm + a_off), 16384, smem_u32(c.sm + b_off), 16384, 128, umma_idesc_bf16(128, 128, true, true), accumulate);
}
// column sums of a per-edge tile: D[128 columns x 16] (+)= tile^T x ones
__device__ __forceinline__ void gx_colsum(const GCtx& c, int vidx, int a_off, bool accumulate) {
    const uint32_t idesc = umma_idesc_bf16(128, 16, true, false);
    const uint32_t ab = smem_u32(c.sm + a_off);
    const uint64_t ones = umma_desc_k128(smem_u32(c.sm + oGOnes));
#pragma unroll
    for (int k16 = 0; k16 < 8; ++k16)
        umma_bf16(c.tm + 256 + 16 * vidx, umma_desc_mn128(ab + k16 * 2048, 16384), ones, idesc, (k16 || accumulate) ? 1u : 0u);
}

// slots of the vector gradients (order of the partial buffer, as in egnn.cu)
enum { VG_DB1 = 0, VG_DB2, VG_DG1, VG_DBE1, VG_DG2, VG_DBE2, VG_DG3, VG_DBE3, VG_DW3, VG_DWD };

// FUSED (dst pass only): the ten vector gradients are taken here as well (warp-shuffle column sums accumulated in
// registers), and d(pre1) / d(delta) are written per edge so that dL/dQ and the pos_j part follow from two segmented
// sums over the src-sorted CSR instead of a second recompute pass.
template <int ACT, bool SRC, bool FUSED>
__global__ void __launch_bounds__(kGThreads, 1)
egnn_bwd_tc_kernel(EgnnTcArgs a, const float* __restrict__ g_msg, const float* __restrict__ g_pos, float* __restrict__ dnode,
                   float* __restrict__ dpos, float* __restrict__ parts) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    GCtx c;
    egnn_tc_setup<512>(c, a, sm);
    const int t = c.t, e = c.e, hf = c.hf;
    float* sc = c.sc;
    const float* vec = c.vec;
    bool acc_w = false;   // the tensor-memory accumulators hold a previous tile's contribution
    float db3 = 0.f;
    const int lane = t & 31;
    float vacc[10];
#pragma unroll
    for (int v = 0; v < 10; ++v) vacc[v] = 0.f;

    for (int rg = blockIdx.x; rg < a.nranges; rg += gridDim.x) {
        const int r0 = lower_bound_row(a.rowptr, (int)a.n, (int64_t)rg * a.grange);
        const int r1 = (rg + 1 == a.nranges) ? (int)a.n : lower_bound_row(a.rowptr, (int)a.n, (int64_t)(rg + 1) * a.grange);
        if (r0 >= r1) continue;
        const int64_t eb = __ldg(a.rowptr + r0), ee = __ldg(a.rowptr + r1);
        int cur = r0;
        float acc = 0.f;
        const int role = t < kGF ? 0 : (t < kGF + 3 ? 1 : 2);
        const int comp = t - kGF;
        auto flush = [&](int row, float v) {
            if (role == 0) dnode[(int64_t)row * kGF + t] = v;
            else dpos[(int64_t)row * 3 + comp] = SRC ? -v : v;
        };
        for (int64_t e0 = eb; e0 < ee; e0 += kGT) {
            const int cnt = (int)min((int64_t)kGT, ee - e0);
            egnn_tc_tile_forward<ACT, SRC>(c, a, e0, cnt);
            const bool live = e < cnt;
            // ---- upstream gradients per edge: coordinate part (mean over the destination's in-edges)
            if (t < kGT) {
                float gx = 0.f, gy = 0.f, gz = 0.f, scale = 0.f;
                if (t < cnt) {
                    const int i = c.ints[2 * kGT + t];
                    const float deg = (float)(__ldg(a.deg_rowptr + i + 1) - __ldg(a.deg_rowptr + i));
                    const float inv = deg > 0.f ? 1.f / deg : 0.f;
                    gx = __ldg(g_pos + 3 * (int64_t)i) * inv;
                    gy = __ldg(g_pos + 3 * (int64_t)i + 1) * inv;
                    gz = __ldg(g_pos + 3 * (int64_t)i + 2) * inv;
                    scale = a.aggr_mean ? inv : 1.f;
                }
                sc[TS_GS * kGT + t] = gx * sc[TS_DX * kGT + t] + gy * sc[TS_DY * kGT + t] + gz * sc[TS_DZ * kGT + t];
                sc[TS_SCALE * kGT + t] = scale;
                sc[TS_GX * kGT + t] = gx; sc[TS_GY * kGT + t] = gy; sc[TS_GZ * kGT + t] = gz;
            }
            __syncthreads();
            const float ds = sc[TS_GS * kGT + e];
            if ((SRC || FUSED) && hf == 0 && live) db3 += ds;
            // ================= stage 3 backward =================
            if (SRC) {  // first the two tiles that do not need the row statistics of dy: ds * a3 (dw3) and dy (dbe3)
                float x[kGCW];
                gx_ld64(c, 128, x);
                const float mean = sc[TS_M3 * kGT + e], rstd = sc[TS_R3 * kGT + e];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const V8 pB2 = ldv8(vec + TV_B2 * kGF + kGCW * hf + 8 * q);
                    const V8 pG3 = ldv8(vec + TV_G3 * kGF + kGCW * hf + 8 * q);
                    const V8 pE3 = ldv8(vec + TV_BE3 * kGF + kGCW * hf + 8 * q);
                    const V8 pW3 = ldv8(vec + TV_W3 * kGF + kGCW * hf + 8 * q);
                    float ta[8], tb[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int col = kGCW * hf + 8 * q + u;
                        const float y = fmaf((x[8 * q + u] + pB2.v[u] - mean) * rstd, pG3.v[u], pE3.v[u]);
                        ta[u] = live ? ds * tact<ACT>(y) : 0.f;
                        tb[u] = live ? ds * pW3.v[u] * tdact<ACT>(y) : 0.f;
                    }
                    *reinterpret_cast<uint4*>(sm + oGA1 + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(ta);
                    *reinterpret_cast<uint4*>(sm + oGM + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(tb);
                }
                gx_issue_begin(c);
                if (c.warp == 0) {
                    tc_fence_after();
                    if (elect_one()) {
                        gx_colsum(c, VG_DW3, oGA1, acc_w);
                        gx_colsum(c, VG_DBE3, oGM, acc_w);
                        umma_commit(c.bar);
                    }
                    __syncwarp();
                }
                gx_mma_wait(c);
            }
            {
                float x[kGCW], d[kGCW];
                gx_ld64(c, 128, x);
                const float mean = sc[TS_M3 * kGT + e], rstd = sc[TS_R3 * kGT + e];
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const V8 pB2 = ldv8(vec + TV_B2 * kGF + kGCW * hf + 8 * q);
                    const V8 pG3 = ldv8(vec + TV_G3 * kGF + kGCW * hf + 8 * q);
                    const V8 pE3 = ldv8(vec + TV_BE3 * kGF + kGCW * hf + 8 * q);
                    const V8 pW3 = ldv8(vec + TV_W3 * kGF + kGCW * hf + 8 * q);
                    float tg[8], ta[8], tb[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = 8 * q + u, col = kGCW * hf + k;
                        const float xh = (x[k] + pB2.v[u] - mean) * rstd;
                        const float y = fmaf(xh, pG3.v[u], pE3.v[u]);
                        const float dy = ds * pW3.v[u] * tdact<ACT>(y);
                        tg[u] = live ? dy * xh : 0.f;
                        if (FUSED) { ta[u] = live ? ds * tact<ACT>(y) : 0.f; tb[u] = live ? dy : 0.f; }
                        x[k] = xh;
                        d[k] = dy * pG3.v[u];
                        s1 += d[k];
                        s2 = fmaf(d[k], xh, s2);
                    }
                    if (SRC) *reinterpret_cast<uint4*>(sm + oGA1 + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(tg);
                    if (FUSED) { GX_VACC(VG_DW3, ta, q) GX_VACC(VG_DBE3, tb, q) GX_VACC(VG_DG3, tg, q) }
                }
                const float2 tot = gx_exchange(c, s1, s2);
                s1 = tot.x * (1.f / kGF);
                s2 = tot.y * (1.f / kGF);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float o[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) o[u] = live ? rstd * (d[8 * q + u] - s1 - x[8 * q + u] * s2) : 0.f;
                    *reinterpret_cast<uint4*>(sm + oGDT + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(o);
                    if (FUSED) GX_VACC(VG_DB2, o, q)
                }
            }
            gx_issue_begin(c);
            if (c.warp == 0) {  // dm = dpre3 W2 (over D2); dW2 += dpre3^T m  |  column sums dg3, db2
                tc_fence_after();
                if (elect_one()) {
                    gx_dgrad(c, 128, oGDT, oGW2);
                    if (SRC) {
                        gx_colsum(c, VG_DG3, oGA1, acc_w);
                        gx_colsum(c, VG_DB2, oGDT, acc_w);
                    } else {
                        gx_wgrad(c, 384, oGDT, oGM, acc_w);
                    }
                    umma_commit(c.bar);
                }
                __syncwarp();
            }
            gx_mma_wait(c);
            // ================= stage 2 backward =================
            {
                float x[kGCW], d[kGCW];
                gx_ld64(c, 0, x);      // pre2 - b1
                gx_ld64(c, 128, d);    // dm
                const float mean = sc[TS_M2 * kGT + e], rstd = sc[TS_R2 * kGT + e], scale = sc[TS_SCALE * kGT + e];
                const float4* gm = reinterpret_cast<const float4*>(g_msg + (int64_t)c.ints[2 * kGT + e] * kGF + kGCW * hf);
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const V8 pB1 = ldv8(vec + TV_B1 * kGF + kGCW * hf + 8 * q);
                    const V8 pG2 = ldv8(vec + TV_G2 * kGF + kGCW * hf + 8 * q);
                    const V8 pE2 = ldv8(vec + TV_BE2 * kGF + kGCW * hf + 8 * q);
                    float tg[8], tb[8], up[8];
                    if (live) {
                        const float4 u0 = __ldg(gm + 2 * q), u1 = __ldg(gm + 2 * q + 1);
                        up[0] = u0.x; up[1] = u0.y; up[2] = u0.z; up[3] = u0.w; up[4] = u1.x; up[5] = u1.y; up[6] = u1.z; up[7] = u1.w;
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u) up[u] = 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = 8 * q + u, col = kGCW * hf + k;
                        const float xh = (x[k] + pB1.v[u] - mean) * rstd;
                        const float y = fmaf(xh, pG2.v[u], pE2.v[u]);
                        const float dy = live ? (d[k] + up[u] * scale) * tdact<ACT>(y) : 0.f;
                        tg[u] = dy * xh;
                        tb[u] = dy;
                        x[k] = xh;
                        d[k] = dy * pG2.v[u];
                        s1 += d[k];
                        s2 = fmaf(d[k], xh, s2);
                    }
                    if (SRC) {
                        *reinterpret_cast<uint4*>(sm + oGA1 + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(tg);
                        *reinterpret_cast<uint4*>(sm + oGM + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(tb);
                    }
                    if (FUSED) { GX_VACC(VG_DG2, tg, q) GX_VACC(VG_DBE2, tb, q) }
                }
                const float2 tot = gx_exchange(c, s1, s2);
                s1 = tot.x * (1.f / kGF);
                s2 = tot.y * (1.f / kGF);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float o[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) o[u] = live ? rstd * (d[8 * q + u] - s1 - x[8 * q + u] * s2) : 0.f;
                    *reinterpret_cast<uint4*>(sm + oGDT + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(o);
                    if (FUSED) GX_VACC(VG_DB1, o, q)
                }
            }
            gx_issue_begin(c);
            if (c.warp == 0) {  // da1 = dpre2 W1 (over D1); dW1 += dpre2^T a1  |  column sums dg2, dbe2, db1
                tc_fence_after();
                if (elect_one()) {
                    gx_dgrad(c, 0, oGDT, oGW1);
                    if (SRC) {
                        gx_colsum(c, VG_DG2, oGA1, acc_w);
                        gx_colsum(c, VG_DBE2, oGM, acc_w);
                        gx_colsum(c, VG_DB1, oGDT, acc_w);
                    } else {
                        gx_wgrad(c, 256, oGDT, oGA1, acc_w);
                    }
                    umma_commit(c.bar);
                }
                __syncwarp();
            }
            gx_mma_wait(c);
            // ================= stage 1 backward =================
            {
                float x[kGCW], d[kGCW];
                gx_ld64(c, 0, d);  // da1
                const float rstd = sc[TS_R1 * kGT + e], dist = sc[TS_DIST * kGT + e];
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const V8 pG1 = ldv8(vec + TV_G1 * kGF + kGCW * hf + 8 * q);
                    const V8 pE1 = ldv8(vec + TV_BE1 * kGF + kGCW * hf + 8 * q);
                    float xh[8], tg[8], tb[8];
                    unpack8(*reinterpret_cast<const uint4*>(sm + oGG + e * kGLd + (4 * hf + q) * 16), xh);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = 8 * q + u, col = kGCW * hf + k;
                        const float y = fmaf(xh[u], pG1.v[u], pE1.v[u]);
                        const float dy = live ? d[k] * tdact<ACT>(y) : 0.f;
                        tg[u] = dy * xh[u];
                        tb[u] = dy;
                        x[k] = xh[u];
                        d[k] = dy * pG1.v[u];
                        s1 += d[k];
                        s2 = fmaf(d[k], xh[u], s2);
                    }
                    if (SRC) {
                        *reinterpret_cast<uint4*>(sm + oGA1 + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(tg);
                        *reinterpret_cast<uint4*>(sm + oGM + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(tb);
                    }
                    if (FUSED) { GX_VACC(VG_DG1, tg, q) GX_VACC(VG_DBE1, tb, q) }
                }
                const float2 tot = gx_exchange(c, s1, s2);
                s1 = tot.x * (1.f / kGF);
                s2 = tot.y * (1.f / kGF);
                float dd = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const V8 pWD = ldv8(vec + TV_WD * kGF + kGCW * hf + 8 * q);
                    float o[8], od[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        o[u] = live ? rstd * (d[8 * q + u] - s1 - x[8 * q + u] * s2) : 0.f;
                        od[u] = dist * o[u];
                        dd = fmaf(o[u], pWD.v[u], dd);
                    }
                    const uint4 po = pack8(o);
                    *reinterpret_cast<uint4*>(sm + oGG + e * kGLd + (4 * hf + q) * 16) = po;   // dpre1, for the column walkers
                    if (SRC) *reinterpret_cast<uint4*>(sm + oGDT + (hf >> 1) * 16384 + sw128_chunk_off(e, (hf & 1) * 4 + q)) = pack8(od);
                    if (FUSED) {
                        if (live) *reinterpret_cast<uint4*>(a.dpre_out + (int64_t)c.ints[3 * kGT + e] * kGF + kGCW * hf + 8 * q) = po;
                        GX_VACC(VG_DWD, od, q)
                    }
                }
                const float2 dt = gx_exchange(c, dd, 0.f);
                if (hf == 0) {
                    // d(delta) = g_pos_mean * s + d(dist) * delta / dist   (subgradient 0 at dist = 0, as torch.norm)
                    const float s = sc[TS_S * kGT + e];
                    const float k = dist > 0.f ? dt.x / dist : 0.f;
                    const float gx = sc[TS_GX * kGT + e], gy = sc[TS_GY * kGT + e], gz = sc[TS_GZ * kGT + e];
                    sc[TS_GX * kGT + e] = live ? fmaf(k, sc[TS_DX * kGT + e], gx * s) : 0.f;
                    sc[TS_GY * kGT + e] = live ? fmaf(k, sc[TS_DY * kGT + e], gy * s) : 0.f;
                    sc[TS_GZ * kGT + e] = live ? fmaf(k, sc[TS_DZ * kGT + e], gz * s) : 0.f;
                    if (FUSED && live)
                        *reinterpret_cast<float4*>(a.ddelta_out + 4 * (int64_t)c.ints[3 * kGT + e]) =
                            make_float4(sc[TS_GX * kGT + e], sc[TS_GY * kGT + e], sc[TS_GZ * kGT + e], 0.f);
                }
            }
            gx_issue_begin(c);
            if (SRC && c.warp == 0) {  // column sums dg1, dbe1, dwd (run while the walkers below reduce)
                tc_fence_after();
                if (elect_one()) {
                    gx_colsum(c, VG_DG1, oGA1, acc_w);
                    gx_colsum(c, VG_DBE1, oGM, acc_w);
                    gx_colsum(c, VG_DWD, oGDT, acc_w);
                    umma_commit(c.bar);
                }
                __syncwarp();
            }
            acc_w = true;
            // ---- segmented sums over the CSR rows: d(pre1) -> dP (dst pass) / dQ (src pass); d(delta) -> dpos
            if (role == 0) {
                gx_walk(c.ints, cnt, cur, acc,
                        [&](int k) { return __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(sm + oGG + k * kGLd + t * 2)) << 16); }, flush);
            } else if (role == 1) {
                gx_walk(c.ints, cnt, cur, acc, [&](int k) { return sc[(TS_GX + comp) * kGT + k]; }, flush);
            }
            if (SRC) gx_mma_wait(c);
            __syncthreads();
        }
        if (role != 2) {
            flush(cur, acc);
            for (int z = cur + 1; z < r1; ++z) flush(z, 0.f);
        }
    }
    // ---- per-CTA partial parameter gradients out of tensor memory (zeros when this CTA saw no tile)
    const int64_t plen = 2 * kGF * kGF + 10 * kGF + 4;
    float* my = parts + (int64_t)blockIdx.x * plen;
    tc_fence_after();
    if (FUSED) {   // vector gradients: registers -> sum over the four row quarters that share a column quarter
        float* vout = my + 2 * kGF * kGF;
        float* red = reinterpret_cast<float*>(sm + oGRed);
#pragma unroll
        for (int v = 0; v < 10; ++v) {
            __syncthreads();
            red[(c.warp & 3) * kGF + kGCW * hf + lane] = vacc[v];
            __syncthreads();
            if (t < kGF) vout[v * kGF + t] = (red[t] + red[kGF + t]) + (red[2 * kGF + t] + red[3 * kGF + t]);
        }
        __syncthreads();
        if (hf == 0) red[e] = db3;
        __syncthreads();
        if (t == 0) {
            float sacc = 0.f;
            for (int k = 0; k < kGT; ++k) sacc += red[k];
            vout[10 * kGF] = sacc;
            vout[10 * kGF + 1] = vout[10 * kGF + 2] = vout[10 * kGF + 3] = 0.f;
        }
    }
    if (!SRC) {
        for (int w = 0; w < 2; ++w) {  // lane = output feature (row of W), columns = input features
            float v[kGCW];
            gx_ld64(c, 256 + 128 * w, v);
            float* dst = my + (int64_t)w * kGF * kGF + e * kGF + kGCW * hf;
#pragma unroll
            for (int j = 0; j < kGCW; j += 4)
                *reinterpret_cast<float4*>(dst + j) = acc_w ? make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        float* vout = my + 2 * kGF * kGF;
        for (int p = hf; p < 5; p += kGQ) {  // lane = feature column; vectors 2p and 2p+1 sit in accumulator columns 0 and 16
            float v[32];
            tmem_ld32(c.tm + c.lane_base + 256 + 32 * p, v);
            vout[(2 * p) * kGF + e] = acc_w ? v[0] : 0.f;
            vout[(2 * p + 1) * kGF + e] = acc_w ? v[16] : 0.f;
        }
        float* red = reinterpret_cast<float*>(sm + oGRed);
        __syncthreads();
        if (hf == 0) red[e] = db3;
        __syncthreads();
        if (t == 0) {
            float s = 0.f;
            for (int k = 0; k < kGT; ++k) s += red[k];
            vout[10 * kGF] = s;
            vout[10 * kGF + 1] = vout[10 * kGF + 2] = vout[10 * kGF + 3] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (c.warp == 0) tmem_dealloc<512>(c.tm);
}

// edges per work range: two ranges per SM, in whole 128-edge tiles, between one tile and kGRangeMax
static int egnn_tc_range(int64_t E) {
    int64_t r = ceil_div(E > 0 ? E : 1, 2 * (int64_t)num_sms());
    r = ceil_div(r, 128) * 128;
    return (int)(r < 128 ? 128 : (r > kGRangeMax ? kGRangeMax : r));
}

static EgnnTcArgs egnn_tc_args(const int32_t* rowptr, const int32_t* col, const int32_t* rowid, const int32_t* deg_rowptr, int64_t n,
                               int64_t E, const float* rowt, const void* gath, const float* pos, const gmp_egnn_edge_params* p) {
    EgnnTcArgs a;
    a.rowptr = rowptr; a.col = col; a.rowid = rowid; a.deg_rowptr = deg_rowptr; a.perm = nullptr; a.n = n; a.E = E; a.rowt = rowt;
    a.dpre_out = nullptr; a.ddelta_out = nullptr;
    a.peer = make_peer_rows(nullptr);
    a.gath = (const __nv_bfloat16*)gath; a.pos = pos;
    a.wd = p->wd; a.g1 = p->ln1_g; a.be1 = p->ln1_b; a.w1 = p->w1; a.b1 = p->b1; a.g2 = p->ln2_g; a.be2 = p->ln2_b;
    a.w2 = p->w2; a.b2 = p->b2; a.g3 = p->ln3_g; a.be3 = p->ln3_b; a.w3 = p->w3; a.b3 = p->b3;
    a.aggr_mean = p->aggr_mean; a.eps = p->ln_eps;
    a.grange = egnn_tc_range(E);
    a.nranges = (int)(E > 0 ? ceil_div(E, a.grange) : 1);
    return a;
}

static int egnn_tc_check(const gmp_egnn_edge_params* p, int64_t n, int64_t E) {
    GMP_REQUIRE(p, "egnn_tc: params is NULL");
    GMP_REQUIRE(p->d == 128, "egnn_tc: the tensor-core path is built for emb_dim = 128 (got %d)", p->d);
    GMP_REQUIRE(p->act == 0 || p->act == 1, "egnn_tc: act must be 0 (relu) or 1 (swish)");
    GMP_REQUIRE(p->wd && p->ln1_g && p->ln1_b && p->w1 && p->b1 && p->ln2_g && p->ln2_b && p->w2 && p->b2 && p->ln3_g &&
                p->ln3_b && p->w3 && p->b3, "egnn_tc: NULL parameter pointer");
    GMP_REQUIRE(n >= 0 && E >= 0 && n < (1ll << 31) && E < (1ll << 31), "egnn_tc: sizes out of range");
    return GMP_OK;
}

}  // namespace gmp

using namespace gmp;

extern "C" {

int gmp_egnn_tc_edge_fwd(const int32_t* rowptr, const int32_t* col, const int32_t* rowid, int64_t n, int64_t num_edges, const float* P,
                         const void* Q_bf16, const float* pos, const gmp_egnn_edge_params* prm, float* msg_aggr, float* pos_aggr,
                         gmp_stream_t stream) {
    if (int rc = egnn_tc_check(prm, n, num_edges)) return rc;
    GMP_REQUIRE(rowptr && msg_aggr && pos_aggr && pos && (num_edges == 0 || (col && rowid && P && Q_bf16)), "egnn_tc_edge_fwd: NULL pointer");
    if (n == 0) return GMP_OK;
    const EgnnTcArgs a = egnn_tc_args(rowptr, col, rowid, rowptr, n, num_edges, P, Q_bf16, pos, prm);
    const int grid = a.nranges < num_sms() ? a.nranges : num_sms();
    if (prm->act) {
        GMP_CUDA(cudaFuncSetAttribute(egnn_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEgnnTcSmem));
        egnn_fwd_tc_kernel<1><<<grid, kGThreads, kEgnnTcSmem, stream>>>(a, msg_aggr, pos_aggr);
    } else {
        GMP_CUDA(cudaFuncSetAttribute(egnn_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEgnnTcSmem));
        egnn_fwd_tc_kernel<0><<<grid, kGThreads, kEgnnTcSmem, stream>>>(a, msg_aggr, pos_aggr);
    }
    return check_launch("egnn_fwd_tc_kernel");
}

int32_t gmp_egnn_tc_bwd_num_parts(int64_t num_edges) {
    const int64_t nr = num_edges > 0 ? ceil_div(num_edges, egnn_tc_range(num_edges)) : 1;
    return (int32_t)(nr < num_sms() ? nr : num_sms());
}

int gmp_egnn_tc_edge_bwd(const int32_t* rowptr, const int32_t* col, const int32_t* rowid, const int32_t* dst_rowptr, int64_t n,
                         int64_t num_edges, const float* row_operand, const void* col_operand_bf16, const float* pos,
                         const gmp_egnn_edge_params* prm, const float* g_msg, const float* g_pos, int32_t src_pass, float* d_node,
                         float* d_pos, float* wgrad_parts, gmp_stream_t stream) {
    if (int rc = egnn_tc_check(prm, n, num_edges)) return rc;
    GMP_REQUIRE(rowptr && dst_rowptr && g_msg && g_pos && d_node && d_pos && pos && wgrad_parts &&
                (num_edges == 0 || (col && rowid && row_operand && col_operand_bf16)), "egnn_tc_edge_bwd: NULL pointer");
    if (n == 0) return GMP_OK;
    const EgnnTcArgs a = egnn_tc_args(rowptr, col, rowid, dst_rowptr, n, num_edges, row_operand, col_operand_bf16, pos, prm);
    const int grid = a.nranges < num_sms() ? a.nranges : num_sms();
#define GMP_EGNN_TC_BWD(A_, S_)                                                                                              \
    {                                                                                                                        \
        GMP_CUDA(cudaFuncSetAttribute(egnn_bwd_tc_kernel<A_, S_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEgnnTcSmem)); \
        egnn_bwd_tc_kernel<A_, S_, false><<<grid, kGThreads, kEgnnTcSmem, stream>>>(a, g_msg, g_pos, d_node, d_pos, wgrad_parts);  \
    }
    if (prm->act) { if (src_pass) GMP_EGNN_TC_BWD(1, true) else GMP_EGNN_TC_BWD(1, false) }
    else { if (src_pass) GMP_EGNN_TC_BWD(0, true) else GMP_EGNN_TC_BWD(0, false) }
#undef GMP_EGNN_TC_BWD
    return check_launch("egnn_bwd_tc_kernel");
}

// Single-pass backward: the dst pass with everything (dL/dP, pos_i part, weight AND vector gradients) plus per-edge d(pre1)
// (bf16 [E,128]) and d(delta) (float4 [E]) in the caller's edge order; dL/dQ and the pos_j part are then two segmented sums
// over the src-sorted CSR (gmp_segment_sum_bf16_f32 / gmp_segment_reduce_f32), no second recompute pass.
int gmp_egnn_tc_edge_bwd_fused(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid, int64_t n,
                               int64_t num_edges, const float* P, const void* Q_bf16, const float* pos, const gmp_egnn_edge_params* prm,
                               const float* g_msg, const float* g_pos, float* dP, float* dpos_i, float* wgrad_parts, void* dpre1_bf16,
                               float* ddelta, const gmp_peer_rows* peer, gmp_stream_t stream) {
    if (int rc = egnn_tc_check(prm, n, num_edges)) return rc;
    GMP_REQUIRE(!peer || (peer->n_left >= 0 && peer->n_own > 0 && peer->n_left + peer->n_own <= n && (peer->n_left == 0 || peer->left) &&
                          (peer->n_left + peer->n_own == n || peer->right)), "egnn_tc_edge_bwd_fused: inconsistent peer rows");
    GMP_REQUIRE(rowptr && g_msg && g_pos && dP && dpos_i && pos && wgrad_parts &&
                (num_edges == 0 || (col && rowid && P && Q_bf16 && dpre1_bf16 && ddelta)), "egnn_tc_edge_bwd_fused: NULL pointer");
    if (n == 0) return GMP_OK;
    EgnnTcArgs a = egnn_tc_args(rowptr, col, rowid, rowptr, n, num_edges, P, Q_bf16, pos, prm);
    a.perm = perm; a.dpre_out = (__nv_bfloat16*)dpre1_bf16; a.ddelta_out = ddelta;
    a.peer = make_peer_rows(peer);
    const int grid = a.nranges < num_sms() ? a.nranges : num_sms();
    if (prm->act) {
        GMP_CUDA(cudaFuncSetAttribute(egnn_bwd_tc_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEgnnTcSmem));
        egnn_bwd_tc_kernel<1, false, true><<<grid, kGThreads, kEgnnTcSmem, stream>>>(a, g_msg, g_pos, dP, dpos_i, wgrad_parts);
    } else {
        GMP_CUDA(cudaFuncSetAttribute(egnn_bwd_tc_kernel<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEgnnTcSmem));
        egnn_bwd_tc_kernel<0, false, true><<<grid, kGThreads, kEgnnTcSmem, stream>>>(a, g_msg, g_pos, dP, dpos_i, wgrad_parts);
    }
    return check_launch("egnn_bwd_tc_kernel<fused>");
}

}  // extern "C"
