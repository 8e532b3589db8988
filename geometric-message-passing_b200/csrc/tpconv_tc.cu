// TFN / MACE tensor-product convolution on the 5th-generation tensor cores (GMP_BF16_TC), models/layers/tfn_layer.py:82-87.
//
//   T_e[a,b] = sum_h hid_e[h] * W2[(a,b),h]          bf16 x bf16 -> fp32, tcgen05.mma, accumulators in TMEM
//   res_e[b,k] = sum_a T_e[a,b] * Y_e[a,k]           fp32 FFMA by the epilogue warps straight out of TMEM
//   res[n]     = sum_{e in CSR row n} res_e          fp32, walker warps, fixed order, no atomics
//
// (a = summed multiplicity index, b = kept multiplicity index: forward a = u, b = w; feature gradient a = w, b = u.)
// The [E, weight_numel] tensor of the reference (272 KB / 704 KB per edge at the BASELINE configs) lives only as
// 128 x 256 fp32 tiles in tensor memory.  The CTA is EDGE-stationary: it keeps one 128-edge tile of the hidden layer
// (bf16, 64 KB, UMMA K-major 128B-swizzle image prepared by tp_pack_hid_kernel) in shared memory and streams every
// 256-column slice of W2 ("N-tile", images prepared by tp_pack_w2_kernel, L2-resident) through a 4-slot ring with
// cp.async.bulk; per (tile, path, range of a) the geometric factor Y (or the gathered row itself when that is the
// smaller side) is built once into a thread-private shared-memory scratch and reused by every N-tile of the path.
// Warp roles (384 threads): 0-7 epilogue (thread = edge = TMEM lane; warps 0-3 take columns 0-127, 4-7 columns
// 128-255 of each N-tile), 8-9 walker (segmented row sums + read-modify-write of rows this CTA owns), 10 bulk-copy
// producer, 11 MMA issuer.  fc's second bias never enters the MMA: sum_e sum_a b2[a,b] Y_e[a,k] is linear in
// YS[n] = sum_e Y_e, which tp_ysum_kernel aggregates per node (fp32) and a node-level GEMM finishes on the host side.
#include "common.cuh"
#include "tc_common.cuh"

namespace gmp {

using namespace tc;

constexpr int kET = 128;        // edges per tile
constexpr int kStage = 16384;   // bytes of one [128 rows][64 bf16] swizzled slab
constexpr int kNB = 4;          // W2 ring slots
constexpr int kFBytes = 49152;  // factor scratch
constexpr int kOLd = 41;        // fp32 row stride of the per-edge result tile (odd: conflict-free)
constexpr int kTcThreads = 384;  // 12 warps = 3 per scheduler: 16384 / 96 = 170 registers per thread
constexpr int kEpiThreads = 256, kWalkThreads = 64;
constexpr int kWalkBatch = 6;   // (segment, column) pairs a walker thread keeps in flight

// One (path, range of the summed index) group: the factor built for it serves nslices * nsub N-tiles.
struct TcYGroup {
    int32_t v_off, DA, DB, DS;        // gathered block offset / irrep dims of the gathered, kept and sh sides
    int32_t sh_off, cg_off, A0, AR;   // sh block, CG table (Zc[iA][j][kB], coefficient folded in), a in [A0, A0+AR)
    int32_t r_off, MB, nslices, nsub; // result block offset, kept multiplicity, kept slices, N-tiles per slice
    int32_t nt_begin, pad0, pad1, pad2;  // first N-tile of the group in the w2 image
};
// One N-tile = 256 generated weights per edge: column c = a_loc * WS + b  (WS = 32 when DB == 1 <= DA, else 8)
struct TcNTile {
    int32_t w_off, stride_a, stride_b, a_begin;  // W2 row = w_off + (a_begin + a_loc) * stride_a + (b0 + b) * stride_b
    int32_t a_end, b0, b_end, WS;
};
// gathered blocks for tp_ysum_kernel
struct TcYPath {
    int32_t v_off, DA, DB, MA;     // YS[n][y_off + a*DB + k] = sum_e sum_i V[col_e][v_off + a*DA + i] Z_e[i][k]
    int32_t y_off, z_off, pad0, pad1;
};
struct TcZEntry { int32_t sh_off, DS, cg_base, cg_stride; };  // Z[z] = sum_j sh[sh_off + j] * cg[cg_base + j*cg_stride]

// ------------------------------------------------------------------------------------------------
// operand images
// ------------------------------------------------------------------------------------------------
// hid = relu(W1 feat + b1) for the 128 edges of a tile (CSR order), bf16, as KS = H/64 swizzled K-major slabs
__global__ void __launch_bounds__(256) tp_pack_hid_kernel(const int32_t* __restrict__ perm, int64_t E, const float* __restrict__ feat,
                                                          int R, const float* __restrict__ w1, const float* __restrict__ b1, int H,
                                                          uint8_t* __restrict__ img) {
    extern __shared__ __align__(16) float smf[];
    float* w1s = smf;            // [H][R]
    float* b1s = w1s + H * R;    // [H]
    float* fs = b1s + H;         // [128][R]
    const int t = threadIdx.x;
    const int64_t e0 = (int64_t)blockIdx.x * kET;
    for (int x = t; x < H * R; x += 256) {  // transposed: w1s[c][h], so that a thread's 8 consecutive h are two float4
        const int h = x / R, c = x - h * R;
        w1s[c * H + h] = __ldg(w1 + x);
    }
    for (int x = t; x < H; x += 256) b1s[x] = __ldg(b1 + x);
    for (int x = t; x < kET * R; x += 256) {
        const int r = x / R, c = x - r * R;
        const int64_t k = e0 + r;
        float v = 0.f;
        if (k < E) v = __ldg(feat + (int64_t)(perm ? __ldg(perm + k) : (int)k) * R + c);
        fs[x] = v;
    }
    __syncthreads();
    const int CH = H / 8;  // 16-byte chunks per row
    uint8_t* dst = img + (int64_t)blockIdx.x * (H / 64) * kStage;
    for (int x = t; x < kET * CH; x += 256) {
        const int r = x / CH, cg = x - r * CH;
        const bool live = e0 + r < E;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = b1s[cg * 8 + q];
        for (int c = 0; c < R; ++c) {
            const float f = fs[r * R + c];
            const float4 wa = *reinterpret_cast<const float4*>(w1s + c * H + cg * 8), wb = *reinterpret_cast<const float4*>(w1s + c * H + cg * 8 + 4);
            v[0] = fmaf(wa.x, f, v[0]); v[1] = fmaf(wa.y, f, v[1]); v[2] = fmaf(wa.z, f, v[2]); v[3] = fmaf(wa.w, f, v[3]);
            v[4] = fmaf(wb.x, f, v[4]); v[5] = fmaf(wb.y, f, v[5]); v[6] = fmaf(wb.z, f, v[6]); v[7] = fmaf(wb.w, f, v[7]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = live ? fmaxf(v[q], 0.f) : 0.f;
        *reinterpret_cast<uint4*>(dst + (cg >> 3) * kStage + sw128_chunk_off(r, cg & 7)) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
}

// W2 slices in N-tile order: stage s = half * KS + ks holds rows (columns of T) [128*half, 128*half+128) x K slab ks
__global__ void __launch_bounds__(256) tp_pack_w2_kernel(const float* __restrict__ w2, int H, const TcNTile* __restrict__ tiles,
                                                         uint8_t* __restrict__ img) {
    const int KS = H / 64;
    const int nt = blockIdx.x / (2 * KS), s = blockIdx.x - nt * 2 * KS;
    const int half = s / KS, ks = s - half * KS;
    const TcNTile T = tiles[nt];
    uint8_t* dst = img + (int64_t)blockIdx.x * kStage;
    for (int x = threadIdx.x; x < 128 * 8; x += 256) {
        const int rr = x >> 3, ch = x & 7;
        const int c = half * 128 + rr;
        const int al = c / T.WS, b = c - al * T.WS;
        const int a = T.a_begin + al, bb = T.b0 + b;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (a < T.a_end && bb < T.b_end) {
            const float* src = w2 + ((int64_t)T.w_off + (int64_t)a * T.stride_a + (int64_t)bb * T.stride_b) * H + ks * 64 + ch * 8;
            const float4 lo = ldg4(src), hi = ldg4(src + 4);
            o = make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
        }
        *reinterpret_cast<uint4*>(dst + sw128_chunk_off(rr, ch)) = o;
    }
}

// ------------------------------------------------------------------------------------------------
// YS[n][y_off_p + a*DB + k] = sum_{e in row n} sum_i V[col_e][v_off_p + a*DA + i] * Z^p_e[i][k]      (fp32)
// ------------------------------------------------------------------------------------------------
constexpr int kYsPairs = 12;  // (path, a) pairs per thread
constexpr int kYsMaxZ = 256;
constexpr int kYsEB = 4;      // edges per synchronisation

struct YsArgs {
    const int32_t *rowptr, *col, *perm;
    int64_t n;
    const float* V;
    int32_t v_len;
    const float* sh;
    int32_t S;
    const TcYPath* paths;
    int32_t npaths, npairs;
    const TcZEntry* zent;
    int32_t nz;
    const float* cg;
    float* YS;
    int32_t y_len;
};

__global__ void __launch_bounds__(128) tp_ysum_kernel(YsArgs a) {
    __shared__ float Zs[kYsEB][kYsMaxZ];
    __shared__ const float* vrows[kYsEB];
    __shared__ TcYPath ps[32];
    const int t = threadIdx.x;
    if (t < a.npaths) ps[t] = a.paths[t];
    __syncthreads();
    // this thread's pairs: flat index f = t + 128*j over the concatenation of the paths' a ranges
    int pp[kYsPairs], pa[kYsPairs];
#pragma unroll
    for (int j = 0; j < kYsPairs; ++j) {
        int f = t + 128 * j, p = -1, aa = 0;
        if (f < a.npairs) {
            p = 0;
            while (f >= ps[p].MA) { f -= ps[p].MA; ++p; }
            aa = f;
        }
        pp[j] = p;
        pa[j] = aa;
    }
    for (int64_t row = blockIdx.x; row < a.n; row += gridDim.x) {
        float acc[kYsPairs][5];
#pragma unroll
        for (int j = 0; j < kYsPairs; ++j)
#pragma unroll
            for (int k = 0; k < 5; ++k) acc[j][k] = 0.f;
        const int64_t eb = __ldg(a.rowptr + row), ee = __ldg(a.rowptr + row + 1);
        for (int64_t e0 = eb; e0 < ee; e0 += kYsEB) {
            const int cnt = (int)min((int64_t)kYsEB, ee - e0);
            __syncthreads();
            for (int z = t; z < a.nz * cnt; z += 128) {
                const int q = z / a.nz, zz = z - q * a.nz;
                const int64_t eid = a.perm ? __ldg(a.perm + e0 + q) : e0 + q;
                const TcZEntry ze = a.zent[zz];
                float sacc = 0.f;
                for (int j = 0; j < ze.DS; ++j) sacc = fmaf(__ldg(a.sh + eid * a.S + ze.sh_off + j), __ldg(a.cg + ze.cg_base + j * ze.cg_stride), sacc);
                Zs[q][zz] = sacc;
            }
            if (t < cnt) vrows[t] = a.V + (int64_t)__ldg(a.col + e0 + t) * a.v_len;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < kYsPairs; ++j) {
                if (pp[j] < 0) continue;
                const TcYPath P = ps[pp[j]];
                for (int q = 0; q < cnt; ++q) {
                    const float* x = vrows[q] + P.v_off + pa[j] * P.DA;
                    const float* Z = Zs[q] + P.z_off;
                    for (int i = 0; i < P.DA; ++i) {
                        const float xv = __ldg(x + i);
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                            if (k < P.DB) acc[j][k] = fmaf(xv, Z[i * P.DB + k], acc[j][k]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kYsPairs; ++j) {
            if (pp[j] < 0) continue;
            const TcYPath P = ps[pp[j]];
            float* y = a.YS + row * a.y_len + P.y_off + pa[j] * P.DB;
#pragma unroll
            for (int k = 0; k < 5; ++k)
                if (k < P.DB) y[k] = acc[j][k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
struct TcTpArgs {
    const int32_t *rowptr, *col, *perm;
    int64_t n, E;
    const float* V;
    int32_t v_len;
    float* res;
    int32_t r_len;
    float* head;           // [nchunks][r_len]: partial sums of a chunk's leading edges whose row started in an earlier chunk
    const float* sh;
    int32_t S;
    const uint8_t* hid_img;
    const uint8_t* w2_img;
    const TcYGroup* yg;
    int32_t nyg, NT, KS;
    const float* cg;
    int64_t ntiles;
};

// shared-memory map (bytes from the 1024-aligned base)
constexpr int oA = 0;                            // hid tile, up to 4 slabs
constexpr int oB = 65536;                        // W2 ring
constexpr int oF = oB + kNB * kStage;            // factor scratch: F[(aidx*W + word)*256 + g*128 + e]
constexpr int oO = oF + kFBytes;                 // 2 planes x [128][41] fp32
constexpr int oSeg = oO + 2 * kET * kOLd * 4;    // srow[128] | seg_start[132] | seg_row[128] | wcount[4]
constexpr int oBar2 = oSeg + (128 + 132 + 128 + 4) * 4;
// barriers: 0 A_full, 1 A_empty, 2..5 B_full, 6..9 B_empty, 10..13 T_full[buf*2+half], 14..15 T_empty[buf]
constexpr int kTcTpSmem = oBar2 + 16 * 8 + 16 + 1024;

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ---- thread-block-cluster helpers (CTA pairs sharing the W2 stream: each CTA fetches half of every stage and multicasts it)
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bulk copy global -> the same shared offset in every CTA of `mask`; each destination's mbarrier receives the byte count
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// all MMAs issued so far by this thread arrive on the barrier at this shared offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

struct EpiCtx {
    const TcTpArgs* a;
    uint8_t* sm;
    uint64_t* bars;
    uint32_t tm, lane_base;
    int e, g;          // tile row (= TMEM lane) and column half
    bool valid;
    int64_t eid;
    int gnode;
    uint32_t gi;       // N-tile counter (all tiles)
};

// Z[i][k] = sum_j sh_e[sh_off + j] * Zc[i][j][k]   (Zc = CG block with the path coefficient folded in)
template <int DA, int DB>
__device__ __forceinline__ void tc_compute_z(const float* __restrict__ sh_row, const float* __restrict__ cg, int DS, bool valid,
                                             float (&Z)[DA][DB]) {
    float sb[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) sb[j] = (valid && j < DS) ? __ldg(sh_row + j) : 0.f;
#pragma unroll
    for (int i = 0; i < DA; ++i)
#pragma unroll
        for (int k = 0; k < DB; ++k) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j)
                if (j < DS) s = fmaf(sb[j], __ldg(cg + (i * DS + j) * DB + k), s);
            Z[i][k] = s;
        }
}

// factor of one a from the gathered components x[0..DA): Y[k] = sum_i x[i] Z[i][k] when DA >= DB, else x itself
template <int DA, int DB>
__device__ __forceinline__ void tc_factor(const float* x, const float (&Z)[DA][DB], float (&y)[(DA < DB) ? DA : DB]) {
    constexpr bool XM = DA < DB;
    constexpr int M = XM ? DA : DB;
    if constexpr (XM) {
#pragma unroll
        for (int i = 0; i < M; ++i) y[i] = x[i];
    } else {
#pragma unroll
        for (int k = 0; k < M; ++k) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < DA; ++i) s = fmaf(x[i], Z[i][k], s);
            y[k] = s;
        }
    }
}

// 4 consecutive a's of the gathered row -> xs[4*DA]  (a0 relative to the y-group's A0; zeros beyond AR / for padding edges)
template <int DA>
__device__ __forceinline__ void tc_load_x4(const float* __restrict__ xrow, int A0, int AR, int al0, bool valid, bool vec_ok,
                                           float (&xs)[4 * DA]) {
    if (valid && vec_ok && al0 + 4 <= AR) {
        const float4* p = reinterpret_cast<const float4*>(xrow + (int64_t)(A0 + al0) * DA);
#pragma unroll
        for (int v = 0; v < DA; ++v) {
            const float4 f = __ldg(p + v);
            xs[4 * v] = f.x; xs[4 * v + 1] = f.y; xs[4 * v + 2] = f.z; xs[4 * v + 3] = f.w;
        }
    } else {
#pragma unroll
        for (int v = 0; v < 4 * DA; ++v) xs[v] = (valid && al0 + v / DA < AR) ? __ldg(xrow + (int64_t)(A0 + al0) * DA + v) : 0.f;
    }
}

template <int M>
__device__ __forceinline__ void tc_store_factor(uint32_t* dst, const float (&y)[M]) {
    if constexpr (M == 1) dst[0] = __float_as_uint(y[0]);
    if constexpr (M == 3) { dst[0] = pack_bf16(y[0], y[1]); dst[256] = pack_bf16(y[2], 0.f); }
    if constexpr (M == 5) { dst[0] = pack_bf16(y[0], y[1]); dst[256] = pack_bf16(y[2], y[3]); dst[512] = pack_bf16(y[4], 0.f); }
}
template <int M>
__device__ __forceinline__ void tc_load_factor(const uint32_t* src, float (&y)[M]) {
    if constexpr (M == 1) y[0] = __uint_as_float(src[0]);
    if constexpr (M == 3) { const uint32_t w0 = src[0], w1 = src[256]; y[0] = bf_lo(w0); y[1] = bf_hi(w0); y[2] = bf_lo(w1); }
    if constexpr (M == 5) {
        const uint32_t w0 = src[0], w1 = src[256], w2 = src[512];
        y[0] = bf_lo(w0); y[1] = bf_hi(w0); y[2] = bf_lo(w1); y[3] = bf_hi(w1); y[4] = bf_lo(w2);
    }
}

// factor scratch of one y-group for this thread's a's (sub-chunk q, local a in [g*HA, g*HA + HA)):
// F[((q*HA + a_local) * W + word) * 256]  (F already offset by g*128 + e)
template <int DA, int DB>
__device__ __forceinline__ void tc_build_factor(uint32_t* F, const float* __restrict__ xrow, int v_len, const TcYGroup& G, int g,
                                                bool valid, const float (&Z)[DA][DB]) {
    constexpr bool XM = DA < DB;
    constexpr int M = XM ? DA : DB;
    constexpr int WS = (!XM && DB == 1) ? 32 : 8;
    constexpr int MC = 256 / WS, HA = MC / 2;
    constexpr int W = (M + 1) / 2;
    const bool vec_ok = ((v_len | G.v_off) & 3) == 0;
    for (int q = 0; q < G.nsub; ++q)
        for (int a4 = 0; a4 < HA; a4 += 4) {
            float xs[4 * DA];
            tc_load_x4<DA>(xrow, G.A0, G.AR, q * MC + g * HA + a4, valid, vec_ok, xs);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float y[M];
                tc_factor<DA, DB>(xs + j * DA, Z, y);
                tc_store_factor<M>(F + ((q * HA + a4 + j) * W) * 256, y);
            }
        }
}

// One y-group for compile-time irrep dimensions DA (gathered) / DB (kept).
//   DA >= DB: the factor is Y[a][k] = sum_i x[a][i] Z[i][k]  (M = DB values per a), results come out directly;
//   DA <  DB: the factor is x[a][i] itself (M = DA values), r[b][i] = sum_a T x, and out[b][k] = sum_i r[b][i] Z[i][k] at the flush.
template <int DA, int DB>
__device__ __forceinline__ void tc_ygroup(EpiCtx& c, const TcYGroup& G) {
    constexpr bool XM = DA < DB;
    constexpr int M = XM ? DA : DB;
    constexpr int WS = (!XM && DB == 1) ? 32 : 8;
    constexpr int MC = 256 / WS, HA = MC / 2;
    constexpr int W = (M + 1) / 2;
    const TcTpArgs& a = *c.a;
    uint32_t* F = reinterpret_cast<uint32_t*>(c.sm + oF) + c.g * 128 + c.e;
    float Z[DA][DB];
    tc_compute_z<DA, DB>(a.sh + c.eid * a.S + G.sh_off, a.cg + G.cg_off, G.DS, c.valid, Z);
    tc_build_factor<DA, DB>(F, a.V + (int64_t)c.gnode * a.v_len + G.v_off, a.v_len, G, c.g, c.valid, Z);
    // ---- N-tiles: kept slices x sub-chunks of the summed index
    for (int sl = 0; sl < G.nslices; ++sl) {
        float r[WS * M];
#pragma unroll
        for (int x = 0; x < WS * M; ++x) r[x] = 0.f;
        for (int q = 0; q < G.nsub; ++q) {
            const uint32_t buf = c.gi & 1u;
            mbar_wait(&c.bars[10 + buf * 2 + c.g], (c.gi >> 1) & 1u);
            tc_fence_after();
            const uint32_t tb = c.tm + c.lane_base + buf * 256 + c.g * 128;
            // the load of block j+1 is in flight while block j is consumed
            uint32_t va[32], vb[32];
            tmem_ld32_issue(tb, va);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t(&v)[32] = (j & 1) ? vb : va;
                uint32_t(&vn)[32] = (j & 1) ? va : vb;
                tmem_ld_wait(v);
                if (j < 3) tmem_ld32_issue(tb + 32 * (j + 1), vn);
                if constexpr (WS == 32) {
                    const float y = __uint_as_float(F[(q * HA + j) * 256]);
#pragma unroll
                    for (int b = 0; b < 32; ++b) r[b] = fmaf(__uint_as_float(v[b]), y, r[b]);
                } else {
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa) {
                        const uint32_t* src = F + ((q * HA + 4 * j + aa) * W) * 256;
                        float y[M];
                        tc_load_factor<M>(src, y);
#pragma unroll
                        for (int b = 0; b < 8; ++b)
#pragma unroll
                            for (int k = 0; k < M; ++k) r[b * M + k] = fmaf(__uint_as_float(v[aa * 8 + b]), y[k], r[b * M + k]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&c.bars[14 + buf]);
            ++c.gi;
        }
        // ---- flush: per-edge results of this kept slice -> O plane g
        bar_sync_named(2, kEpiThreads + kWalkThreads);  // O free
        float* O = reinterpret_cast<float*>(c.sm + oO) + (c.g * kET + c.e) * kOLd;
        if constexpr (!XM) {
#pragma unroll
            for (int x = 0; x < WS * M; ++x) O[x] = r[x];
        } else {
#pragma unroll
            for (int b = 0; b < 8; ++b)
#pragma unroll
                for (int k = 0; k < DB; ++k) {
                    float s = 0.f;
#pragma unroll
                    for (int i = 0; i < M; ++i) s = fmaf(r[b * M + i], Z[i][k], s);
                    O[b * DB + k] = s;
                }
        }
        bar_arrive_named(1, kEpiThreads + kWalkThreads);  // O full
    }
}

// CL2: launched as clusters of two CTAs that walk the N-tiles in lockstep (same rotation, same number of tiles); every W2
// stage is fetched once per pair -- each CTA requests its half and multicasts it into both rings -- which halves the
// L2 -> SM stream, the largest one of this kernel (128 KB per N-tile per CTA otherwise).
template <bool CL2>
__global__ void __launch_bounds__(kTcThreads, 1) tp_contract_tc_kernel(TcTpArgs a) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + oBar2);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + oBar2 + 16 * 8);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int KS = a.KS, NST = 2 * KS;

    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        for (int s = 0; s < kNB; ++s) { mbar_init(&bars[2 + s], 1); mbar_init(&bars[6 + s], CL2 ? 2 : 1); }  // a slot is free when both CTAs' MMAs read it
        for (int s = 0; s < 4; ++s) mbar_init(&bars[10 + s], 1);
        mbar_init(&bars[14], kEpiThreads);
        mbar_init(&bars[15], kEpiThreads);
        fence_mbar_init();
    }
    if (warp == 11) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    if (CL2) cluster_sync_all();   // the peer's barriers are initialised before anything can arrive on them
    tc_fence_after();
    const uint32_t tm = *tmem_ptr;

    // every CTA processes the same number of tile slots (slots beyond the last tile are empty: all edges invalid)
    const int64_t tpc = (a.ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = tpc * blockIdx.x, t1 = t0 + tpc;
    // every CTA (pair) walks the y-groups cyclically from its own starting point, so that at any moment the CTAs stream
    // different parts of the (L2-resident) w2 image instead of all hammering the same lines
    const int y_start = CL2 ? (int)(((int64_t)a.nyg * (blockIdx.x >> 1)) / (gridDim.x >> 1)) : (int)(((int64_t)a.nyg * blockIdx.x) / gridDim.x);

    if (warp < 8) {
        // ================= epilogue: thread = edge of the tile = TMEM lane =================
        EpiCtx c;
        c.a = &a; c.sm = sm; c.bars = bars; c.tm = tm;
        c.lane_base = (uint32_t)((warp & 3) * 32) << 16;
        c.e = (warp & 3) * 32 + lane;
        c.g = warp >> 2;
        c.gi = 0;
        for (int64_t tile = t0; tile < t1; ++tile) {
            const int64_t k = tile * kET + c.e;
            c.valid = k < a.E;
            c.eid = 0; c.gnode = 0;
            if (c.valid) {
                c.eid = a.perm ? __ldg(a.perm + k) : k;
                c.gnode = __ldg(a.col + k);
            }
            for (int yy = 0; yy < a.nyg; ++yy) {
                const TcYGroup G = a.yg[(y_start + yy) % a.nyg];
                switch (G.DA * 8 + G.DB) {
                    case 1 * 8 + 1: tc_ygroup<1, 1>(c, G); break;
                    case 3 * 8 + 1: tc_ygroup<3, 1>(c, G); break;
                    case 5 * 8 + 1: tc_ygroup<5, 1>(c, G); break;
                    case 3 * 8 + 3: tc_ygroup<3, 3>(c, G); break;
                    case 5 * 8 + 3: tc_ygroup<5, 3>(c, G); break;
                    case 5 * 8 + 5: tc_ygroup<5, 5>(c, G); break;
                    case 1 * 8 + 3: tc_ygroup<1, 3>(c, G); break;
                    case 1 * 8 + 5: tc_ygroup<1, 5>(c, G); break;
                    default: tc_ygroup<3, 5>(c, G); break;
                }
            }
        }
    } else if (warp < 10) {
        // ================= walker: segmented sums of the per-edge results, rows owned by this CTA =================
        const int wt = t - 256;
        int* srow = reinterpret_cast<int*>(sm + oSeg);
        int* seg_start = srow + 128;
        int* seg_row = seg_start + 132;
        int* wcount = seg_row + 128;
        const float* O0 = reinterpret_cast<const float*>(sm + oO);
        const float* O1 = O0 + kET * kOLd;
        const int64_t chunk_e0 = t0 * kET;
        if (t0 < t1) bar_arrive_named(2, kEpiThreads + kWalkThreads);  // O starts free
        for (int64_t tile = t0; tile < t1; ++tile) {
            {   // row of every edge of the tile (padding slots repeat the last edge's row), then the segment table;
                // thread wt looks after edges wt and wt + 64: block blk = pass * 2 + (warp - 8) covers 32 consecutive edges
                int myrow[2];
#pragma unroll
                for (int ps = 0; ps < 2; ++ps) {
                    int64_t k = tile * kET + ps * 64 + wt;
                    if (k > a.E - 1) k = a.E - 1;
                    int lo = 0, hi = (int)a.n;  // last row with rowptr[row] <= k
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if ((int64_t)__ldg(a.rowptr + mid) <= k) lo = mid; else hi = mid;
                    }
                    myrow[ps] = lo;
                    srow[ps * 64 + wt] = lo;
                }
                bar_sync_named(3, kWalkThreads);
                bool flag[2];
                unsigned bal[2];
#pragma unroll
                for (int ps = 0; ps < 2; ++ps) {
                    const int e = ps * 64 + wt;
                    flag[ps] = e == 0 || srow[e - 1] != myrow[ps];
                    bal[ps] = __ballot_sync(0xffffffffu, flag[ps]);
                    if (lane == 0) wcount[ps * 2 + (warp - 8)] = __popc(bal[ps]);
                }
                bar_sync_named(3, kWalkThreads);
                int total = 0;
#pragma unroll
                for (int w = 0; w < 4; ++w) total += wcount[w];
#pragma unroll
                for (int ps = 0; ps < 2; ++ps) {
                    const int blk = ps * 2 + (warp - 8);
                    int base = 0;
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                        if (w < blk) base += wcount[w];
                    if (flag[ps]) {
                        const int idx = base + __popc(bal[ps] & ((1u << lane) - 1u));
                        seg_start[idx] = ps * 64 + wt;
                        seg_row[idx] = myrow[ps];
                    }
                }
                if (wt == 0) { seg_start[total] = kET; seg_start[131] = total; }
                bar_sync_named(3, kWalkThreads);
            }
            const int nseg = tile < a.ntiles ? seg_start[131] : 0;   // empty slot: nothing to store
            const bool head0 = (int64_t)__ldg(a.rowptr + seg_row[0]) < chunk_e0;
            for (int yy = 0; yy < a.nyg; ++yy) {
                const TcYGroup G = a.yg[(y_start + yy) % a.nyg];
                const int WS = (G.DA >= G.DB && G.DB == 1) ? 32 : 8;
                for (int sl = 0; sl < G.nslices; ++sl) {
                    const int nb = min(WS, G.MB - sl * WS);
                    const int nv = nb * G.DB;
                    const int cbase = G.r_off + sl * WS * G.DB;
                    bar_sync_named(1, kEpiThreads + kWalkThreads);  // O full
                    // lane = result column, walker warp w takes segments w, w+2, ...  The old values are requested first (the L2
                    // round trip overlaps the shared-memory sums) and O is handed back to the epilogue before the last stores.
                    // No extra barrier between flushes: a thread's stores of flush f precede its own arrival at the O-full barrier
                    // of flush f+1, which every walker thread passes before it loads for f+1.
                    const int wl = wt & 31, ww = wt >> 5;
                    bool released = false;
                    for (int c0 = 0; c0 < nv; c0 += 32) {
                        const int cc = c0 + wl;
                        const bool colok = cc < nv;
                        for (int s0 = ww; s0 < nseg; s0 += 2 * kWalkBatch) {
                            float* dst[kWalkBatch];
                            float old[kWalkBatch], sum[kWalkBatch];
#pragma unroll
                            for (int j = 0; j < kWalkBatch; ++j) {
                                const int sg = s0 + 2 * j;
                                dst[j] = nullptr;
                                old[j] = 0.f;
                                if (colok && sg < nseg) {
                                    dst[j] = (sg == 0 && head0) ? a.head + (int64_t)blockIdx.x * a.r_len + cbase + cc
                                                                : a.res + (int64_t)seg_row[sg] * a.r_len + cbase + cc;
                                    old[j] = __ldcg(dst[j]);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < kWalkBatch; ++j) {
                                const int sg = s0 + 2 * j;
                                float acc = 0.f;
                                if (colok && sg < nseg) {
                                    const int b = seg_start[sg], e = seg_start[sg + 1];
                                    for (int x = b; x < e; ++x) acc += O0[x * kOLd + cc] + O1[x * kOLd + cc];
                                }
                                sum[j] = acc;
                            }
                            if (c0 + 32 >= nv && s0 + 2 * kWalkBatch >= nseg) {   // last reads of O for this flush (uniform per warp)
                                bar_arrive_named(2, kEpiThreads + kWalkThreads);  // O free
                                released = true;
                            }
#pragma unroll
                            for (int j = 0; j < kWalkBatch; ++j)
                                if (dst[j]) *dst[j] = old[j] + sum[j];
                        }
                    }
                    if (!released) bar_arrive_named(2, kEpiThreads + kWalkThreads);  // nothing to do for this warp
                }
            }
            bar_sync_named(3, kWalkThreads);  // seg table reusable
        }
    } else if (warp == 10) {
        // ================= producer: bulk copies of the hidden tile and the W2 stages =================
        if (lane == 0) {
            uint32_t si = 0, ti = 0;
            const uint32_t rank = CL2 ? cluster_rank() : 0u;
            for (int64_t tile = t0; tile < t1; ++tile, ++ti) {
                const int64_t tsrc = tile < a.ntiles ? tile : a.ntiles - 1;   // empty slots reread the last tile (their edges are invalid)
                mbar_wait(&bars[1], (ti & 1u) ^ 1u);
                mbar_expect_tx(&bars[0], (uint32_t)(KS * kStage));
                for (int ks = 0; ks < KS; ++ks)
                    bulk_g2s(sm + oA + ks * kStage, a.hid_img + (tsrc * KS + ks) * (int64_t)kStage, kStage, &bars[0]);
                for (int yy = 0; yy < a.nyg; ++yy) {
                    const TcYGroup G = a.yg[(y_start + yy) % a.nyg];
                    const uint8_t* src = a.w2_img + (int64_t)G.nt_begin * NST * kStage;
                    const int nstage = G.nslices * G.nsub * NST;
                    for (int s = 0; s < nstage; ++s, ++si) {
                        const uint32_t slot = si % kNB, ph = (si / kNB) & 1u;
                        mbar_wait(&bars[6 + slot], ph ^ 1u);
                        mbar_expect_tx(&bars[2 + slot], kStage);
                        if (CL2)   // my half of the stage, into both CTAs' rings
                            bulk_g2s_mc(sm + oB + slot * kStage + rank * (kStage / 2), src + (int64_t)s * kStage + rank * (kStage / 2), kStage / 2,
                                        &bars[2 + slot], (uint16_t)3);
                        else
                            bulk_g2s(sm + oB + slot * kStage, src + (int64_t)s * kStage, kStage, &bars[2 + slot]);
                    }
                }
            }
        }
    } else {
        // ================= MMA issuer: the whole warp runs the loop, one elected lane issues =================
        const uint32_t idesc = umma_idesc_bf16(128, 128);
        const uint64_t adesc0 = umma_desc_k128(smem_u32(sm + oA)), bdesc0 = umma_desc_k128(smem_u32(sm + oB));
        uint32_t si = 0, ti = 0, gi = 0;
        for (int64_t tile = t0; tile < t1; ++tile, ++ti) {
            mbar_wait(&bars[0], ti & 1u);
            for (int nt = 0; nt < a.NT; ++nt, ++gi) {
                const uint32_t buf = gi & 1u;
                mbar_wait(&bars[14 + buf], ((gi >> 1) & 1u) ^ 1u);
                if (KS == 4) {
                    // 8 stages = two revolutions of the 4-slot ring: slot and phase of every stage are compile-time constants
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const int half = s >> 2, ks = s & 3;
                        mbar_wait(&bars[2 + ks], (uint32_t)half);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t d = tm + buf * 256 + half * 128;
                            const uint64_t ad = adesc0 + (uint64_t)(ks * (kStage >> 4)), bd = bdesc0 + (uint64_t)(ks * (kStage >> 4));
                            umma_bf16(d, ad, bd, idesc, ks ? 1u : 0u);
                            umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
                            umma_bf16(d, ad + 4, bd + 4, idesc, 1u);
                            umma_bf16(d, ad + 6, bd + 6, idesc, 1u);
                            if (CL2) umma_commit_mc(&bars[6 + ks], (uint16_t)3); else umma_commit(&bars[6 + ks]);
                            if (ks == 3) umma_commit(&bars[10 + buf * 2 + half]);
                        }
                        __syncwarp();
                    }
                    si += 8;
                    continue;
                }
                for (int half = 0; half < 2; ++half) {
                    const uint32_t d = tm + buf * 256 + half * 128;
                    for (int ks = 0; ks < KS; ++ks, ++si) {
                        const uint32_t slot = si % kNB, ph = (si / kNB) & 1u;
                        mbar_wait(&bars[2 + slot], ph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t ad = adesc0 + (uint64_t)(ks * (kStage >> 4)), bd = bdesc0 + (uint64_t)(slot * (kStage >> 4));
                            umma_bf16(d, ad, bd, idesc, ks ? 1u : 0u);
                            umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
                            umma_bf16(d, ad + 4, bd + 4, idesc, 1u);
                            umma_bf16(d, ad + 6, bd + 6, idesc, 1u);
                            if (CL2) umma_commit_mc(&bars[6 + slot], (uint16_t)3); else umma_commit(&bars[6 + slot]);
                            if (ks == KS - 1) umma_commit(&bars[10 + buf * 2 + half]);
                        }
                        __syncwarp();
                    }
                }
            }
            if (elect_one()) umma_commit(&bars[1]);
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL2) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it or arrive on its barriers
    if (warp == 11) tmem_dealloc<512>(tm);
}

// ------------------------------------------------------------------------------------------------
// weight-gradient side.  dT_e[a,b] = sum_k g[row_e][b,k] * Y_e[a,k] is generated on the CUDA cores, written as a bf16
// UMMA operand image [128 edges][256 columns] (row = edge, 128B swizzle: K-major for dhid = dT W2, the same bytes
// are the MN-major operand of dW2 = dT^T hid), and both products accumulate in tensor memory:
//   tp_dhid_tc_kernel : edge-stationary, D[128 e][H]    += dT[e, N-tile] * W2[N-tile, :]  over every N-tile
//   tp_dw2_tc_kernel  : weight-stationary, D[256 c][H]  += dT[:, c]^T * hid[:, :]         over every edge tile
// ------------------------------------------------------------------------------------------------
// WS elements of one a:  dT[b] = sum_m r[b*M+m] * y[m]  ->  16-byte chunks of row e starting at column c0
template <int M, int WS>
__device__ __forceinline__ void tc_dt_store(uint8_t* A, int e, int c0, const float (&r)[WS * M], const float (&y)[M]) {
#pragma unroll
    for (int ch = 0; ch < WS / 8; ++ch) {
        float d[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float s = 0.f;
#pragma unroll
            for (int m = 0; m < M; ++m) s = fmaf(r[(ch * 8 + q) * M + m], y[m], s);
            d[q] = s;
        }
        const int c = c0 + ch * 8;
        *reinterpret_cast<uint4*>(A + (c >> 6) * kStage + sw128_chunk_off(e, (c & 63) >> 3)) =
            make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
    }
}

// r[b][m] of a kept slice from the row node's gradient: g values (DA >= DB, m = k) or their contraction with Z (m = i)
template <int DA, int DB, int WS>
__device__ __forceinline__ void tc_load_gside(const float* __restrict__ grow, int nb, bool valid, const float (&Z)[DA][DB],
                                              float (&r)[WS * ((DA < DB) ? DA : DB)]) {
    constexpr bool XM = DA < DB;
    constexpr int M = XM ? DA : DB;
#pragma unroll
    for (int b = 0; b < WS; ++b) {
        float gv[DB];
#pragma unroll
        for (int k = 0; k < DB; ++k) gv[k] = (valid && b < nb) ? __ldg(grow + b * DB + k) : 0.f;
        if constexpr (XM) {
#pragma unroll
            for (int i = 0; i < M; ++i) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < DB; ++k) s = fmaf(gv[k], Z[i][k], s);
                r[b * M + i] = s;
            }
        } else {
#pragma unroll
            for (int k = 0; k < M; ++k) r[b * M + k] = gv[k];
        }
    }
}

struct TcDhArgs {
    const int32_t *rowptr, *col, *perm, *rowid;
    int64_t n, E;
    const float* x;
    int32_t x_len;
    const float* g;
    int32_t g_len;
    const float* sh;
    int32_t S;
    const float* feat;
    int32_t R;
    const float *w1, *b1;
    const uint8_t* w2_img;
    const TcYGroup* yg;
    int32_t nyg, NT, KS;
    const float* cg;
    float* dpre;       // [E, H] in the caller's edge order: dL/d(pre-activation of fc's hidden layer)
    int64_t ntiles;
};

constexpr int oDhA = 0;                       // dT tile: 4 slabs
constexpr int oDhB = 65536;                   // W2 ring
constexpr int oDhF = oDhB + kNB * kStage;     // factor scratch
constexpr int oDhW = oDhF + kFBytes;          // w1 [H][R] | b1 [H]
constexpr int oDhBar = oDhW + (256 * 16 + 256) * 4;
// barriers: 0..3 B_full, 4..7 B_empty, 8..9 A_full[half], 10..11 A_empty[half], 12 D_full, 13 D_empty
constexpr int kTcDhSmem = oDhBar + 14 * 8 + 16 + 1024;
constexpr int kGenThreads = 320;  // warps 0-7 generate, 8 bulk-copy producer, 9 MMA issuer

struct DhCtx {
    const TcDhArgs* a;
    uint8_t* sm;
    uint64_t* bars;
    int e, g;
    bool valid;
    int64_t eid;
    int gnode, rowid;
    uint32_t gi;
};

template <int DA, int DB>
__device__ __forceinline__ void dh_ygroup(DhCtx& c, const TcYGroup& G) {
    constexpr bool XM = DA < DB;
    constexpr int M = XM ? DA : DB;
    constexpr int WS = (!XM && DB == 1) ? 32 : 8;
    constexpr int MC = 256 / WS, HA = MC / 2;
    constexpr int W = (M + 1) / 2;
    const TcDhArgs& a = *c.a;
    uint32_t* F = reinterpret_cast<uint32_t*>(c.sm + oDhF) + c.g * 128 + c.e;
    uint8_t* A = c.sm + oDhA;
    float Z[DA][DB];
    tc_compute_z<DA, DB>(a.sh + c.eid * a.S + G.sh_off, a.cg + G.cg_off, G.DS, c.valid, Z);
    tc_build_factor<DA, DB>(F, a.x + (int64_t)c.gnode * a.x_len + G.v_off, a.x_len, G, c.g, c.valid, Z);
    for (int sl = 0; sl < G.nslices; ++sl) {
        float r[WS * M];
        tc_load_gside<DA, DB, WS>(a.g + (int64_t)c.rowid * a.g_len + G.r_off + sl * WS * DB, min(WS, G.MB - sl * WS), c.valid, Z, r);
        for (int q = 0; q < G.nsub; ++q) {
            mbar_wait(&c.bars[10 + c.g], (c.gi & 1u) ^ 1u);  // the MMAs that read this half of the previous dT tile are done
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if constexpr (WS == 32) {
                    float y[1];
                    tc_load_factor<1>(F + (q * HA + j) * 256, y);
                    tc_dt_store<1, 32>(A, c.e, c.g * 128 + j * 32, r, y);
                } else {
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa) {
                        float y[M];
                        tc_load_factor<M>(F + ((q * HA + 4 * j + aa) * W) * 256, y);
                        tc_dt_store<M, 8>(A, c.e, c.g * 128 + (4 * j + aa) * 8, r, y);
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&c.bars[8 + c.g]);
            ++c.gi;
        }
    }
}

#define GMP_TC_DISPATCH(FN, ...)                                  \
    switch (G.DA * 8 + G.DB) {                                    \
        case 1 * 8 + 1: FN<1, 1>(__VA_ARGS__); break;             \
        case 3 * 8 + 1: FN<3, 1>(__VA_ARGS__); break;             \
        case 5 * 8 + 1: FN<5, 1>(__VA_ARGS__); break;             \
        case 3 * 8 + 3: FN<3, 3>(__VA_ARGS__); break;             \
        case 5 * 8 + 3: FN<5, 3>(__VA_ARGS__); break;             \
        case 5 * 8 + 5: FN<5, 5>(__VA_ARGS__); break;             \
        case 1 * 8 + 3: FN<1, 3>(__VA_ARGS__); break;             \
        case 1 * 8 + 5: FN<1, 5>(__VA_ARGS__); break;             \
        default: FN<3, 5>(__VA_ARGS__); break;                    \
    }

// last CSR row whose first edge is <= k
__device__ __forceinline__ int row_of_edge(const int32_t* __restrict__ rowptr, int n, int64_t k) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(rowptr + mid) <= k) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kGenThreads, 1) tp_dhid_tc_kernel(TcDhArgs a) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + oDhBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + oDhBar + 14 * 8);
    float* w1s = reinterpret_cast<float*>(sm + oDhW);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int KS = a.KS, NST = 2 * KS, H = 64 * KS, R = a.R;
    float* b1s = w1s + H * R;
    for (int x = t; x < H * R; x += kGenThreads) w1s[x] = __ldg(a.w1 + x);
    for (int x = t; x < H; x += kGenThreads) b1s[x] = __ldg(a.b1 + x);
    if (t == 0) {
        for (int s = 0; s < kNB; ++s) { mbar_init(&bars[s], 1); mbar_init(&bars[4 + s], 1); }
        mbar_init(&bars[8], 128); mbar_init(&bars[9], 128);
        mbar_init(&bars[10], 1); mbar_init(&bars[11], 1);
        mbar_init(&bars[12], 1); mbar_init(&bars[13], kEpiThreads);
        fence_mbar_init();
    }
    if (warp == 9) tmem_alloc<256>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr;
    const int64_t t0 = (a.ntiles * blockIdx.x) / gridDim.x, t1 = (a.ntiles * (blockIdx.x + 1)) / gridDim.x;
    const int y_start = (int)(((int64_t)a.nyg * blockIdx.x) / gridDim.x);  // see tp_contract_tc_kernel

    if (warp < 8) {
        DhCtx c;
        c.a = &a; c.sm = sm; c.bars = bars;
        c.e = (warp & 3) * 32 + lane;
        c.g = warp >> 2;
        c.gi = 0;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t ti = 0;
        for (int64_t tile = t0; tile < t1; ++tile, ++ti) {
            const int64_t k = tile * kET + c.e;
            c.valid = k < a.E;
            c.eid = 0; c.gnode = 0; c.rowid = 0;
            if (c.valid) {
                c.eid = a.perm ? __ldg(a.perm + k) : k;
                c.gnode = __ldg(a.col + k);
                c.rowid = __ldg(a.rowid + k);
            }
            for (int yy = 0; yy < a.nyg; ++yy) {
                const TcYGroup G = a.yg[(y_start + yy) % a.nyg];
                GMP_TC_DISPATCH(dh_ygroup, c, G)
            }
            // ---- dhid of the tile: relu mask (pre-activation recomputed in fp32), store in the caller's edge order
            mbar_wait(&bars[12], ti & 1u);
            tc_fence_after();
            float f[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) f[q] = (c.valid && q < R) ? __ldg(a.feat + c.eid * R + q) : 0.f;
            const int hh = H / 2;
            for (int cc = 0; cc < hh; cc += 32) {
                float v[32];
                const int h0 = c.g * hh + cc;
                tmem_ld32(tm + lane_base + h0, v);
                if (c.valid) {
                    float* dst = a.dpre + c.eid * H + h0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float o[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int h = h0 + j + u;
                            float pre = b1s[h];
#pragma unroll
                            for (int q = 0; q < 16; ++q)
                                if (q < R) pre = fmaf(w1s[h * R + q], f[q], pre);
                            o[u] = pre > 0.f ? v[j + u] : 0.f;
                        }
                        *reinterpret_cast<float4*>(dst + j) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars[13]);
        }
    } else if (warp == 8) {
        if (lane == 0) {
            uint32_t si = 0;
            for (int64_t tile = t0; tile < t1; ++tile) {
                for (int yy = 0; yy < a.nyg; ++yy) {
                    const TcYGroup G = a.yg[(y_start + yy) % a.nyg];
                    const uint8_t* src = a.w2_img + (int64_t)G.nt_begin * NST * kStage;
                    const int nstage = G.nslices * G.nsub * NST;
                    for (int s = 0; s < nstage; ++s, ++si) {
                        const uint32_t slot = si % kNB, ph = (si / kNB) & 1u;
                        mbar_wait(&bars[4 + slot], ph ^ 1u);
                        mbar_expect_tx(&bars[slot], kStage);
                        bulk_g2s(sm + oDhB + slot * kStage, src + (int64_t)s * kStage, kStage, &bars[slot]);
                    }
                }
            }
        }
    } else {
        const uint32_t idesc = umma_idesc_bf16(128, 64, false, true);
        const uint64_t adesc0 = umma_desc_k128(smem_u32(sm + oDhA)), bdesc0 = umma_desc_mn128(smem_u32(sm + oDhB), kStage);
        uint32_t si = 0, ti = 0, gi = 0;
        for (int64_t tile = t0; tile < t1; ++tile, ++ti) {
            mbar_wait(&bars[13], (ti & 1u) ^ 1u);  // the previous tile's dhid has been read out of tensor memory
            for (int nt = 0; nt < a.NT; ++nt, ++gi) {
                if (KS == 4) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const int half = s >> 2, ks = s & 3;
                        if (ks == 0) mbar_wait(&bars[8 + half], gi & 1u);
                        mbar_wait(&bars[ks], (uint32_t)half);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t ad = adesc0 + (uint64_t)(2 * half * (kStage >> 4)), bd = bdesc0 + (uint64_t)(ks * (kStage >> 4));
                            const uint32_t d = tm + ks * 64;
#pragma unroll
                            for (int k16 = 0; k16 < 8; ++k16)
                                umma_bf16(d, ad + (uint64_t)((k16 >> 2) * (kStage >> 4) + (k16 & 3) * 2), bd + (uint64_t)(k16 * (2048 >> 4)), idesc,
                                          (nt | half | k16) ? 1u : 0u);
                            umma_commit(&bars[4 + ks]);
                            if (ks == 3) umma_commit(&bars[10 + half]);
                        }
                        __syncwarp();
                    }
                    si += 8;
                    continue;
                }
                for (int half = 0; half < 2; ++half) {
                    mbar_wait(&bars[8 + half], gi & 1u);
                    for (int ks = 0; ks < KS; ++ks, ++si) {
                        const uint32_t slot = si % kNB, ph = (si / kNB) & 1u;
                        mbar_wait(&bars[slot], ph);
                        tc_fence_after();
                        if (elect_one()) {
                            // A: dT slabs 2*half, 2*half+1 (K = 128 columns of this half); B: stage rows = K, 64 h columns
                            const uint64_t ad = adesc0 + (uint64_t)(2 * half * (kStage >> 4)), bd = bdesc0 + (uint64_t)(slot * (kStage >> 4));
                            const uint32_t d = tm + ks * 64;
#pragma unroll
                            for (int k16 = 0; k16 < 8; ++k16)
                                umma_bf16(d, ad + (uint64_t)((k16 >> 2) * (kStage >> 4) + (k16 & 3) * 2), bd + (uint64_t)(k16 * (2048 >> 4)), idesc,
                                          (nt | half | k16) ? 1u : 0u);
                            umma_commit(&bars[4 + slot]);
                            if (ks == KS - 1) umma_commit(&bars[10 + half]);
                        }
                        __syncwarp();
                    }
                }
            }
            if (elect_one()) umma_commit(&bars[12]);
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc<256>(tm);
}

// one N-tile of W2 rows with everything the generator needs
struct TcWTile {
    int32_t w_off, stride_a, stride_b, a_begin, a_end, b0, b_end, WS;
    int32_t v_off, DA, DB, DS, sh_off, cg_off, r_off, pad;
};

struct TcDwArgs {
    const int32_t *rowptr, *col, *perm, *rowid;
    int64_t n, E;
    const float* x;
    int32_t x_len;
    const float* g;
    int32_t g_len;
    const float* sh;
    int32_t S;
    const uint8_t* hid_img;
    const TcWTile* wt;
    int32_t KS;
    const float* cg;
    float* dW2;        // [ngroups][numel][H]: one partial per edge group (summed by the caller when ngroups > 1)
    int64_t ntiles, numel;
    int32_t NT, ngroups;
};

constexpr int oDwH = 0;                // hid tiles, 2 x 64 KB
constexpr int oDwA = 131072;           // dT tile
constexpr int oDwBar = oDwA + 65536;   // 0,1 hid_full | 2,3 hid_empty | 4,5 A_full[half] | 6,7 A_empty[half] | 8 D_full
constexpr int kTcDwSmem = oDwBar + 9 * 8 + 16 + 1024;

template <int DA, int DB>
__device__ __forceinline__ void dw_generate(const TcDwArgs& a, const TcWTile& T, uint8_t* sm, uint64_t* bars, int e, int g) {
    constexpr bool XM = DA < DB;
    constexpr int M = XM ? DA : DB;
    constexpr int WS = (!XM && DB == 1) ? 32 : 8;
    constexpr int MC = 256 / WS, HA = MC / 2;
    uint8_t* A = sm + oDwA;
    const bool vec_ok = ((a.x_len | T.v_off) & 3) == 0;
    const int AR = T.a_end - T.a_begin, nb = min(WS, T.b_end - T.b0);
    uint32_t ti = 0;
    // this CTA's edge group, walked cyclically from a CTA-specific starting tile (spreads the L2 traffic of the hid image)
    const int nt = blockIdx.x / a.ngroups, grp = blockIdx.x - nt * a.ngroups;
    const int64_t g0 = (a.ntiles * grp) / a.ngroups, g1 = (a.ntiles * (grp + 1)) / a.ngroups, glen = g1 - g0;
    const int64_t rot = (glen * nt) / a.NT;
    // edge scalars are fetched one tile ahead
    int64_t eid_n = 0;
    int gnode_n = 0, rowid_n = 0;
    auto fetch = [&](int64_t tt) {
        const int64_t k = (g0 + (rot + tt) % glen) * kET + e;
        eid_n = 0; gnode_n = 0; rowid_n = 0;
        if (tt < glen && k < a.E) {
            eid_n = a.perm ? __ldg(a.perm + k) : k;
            gnode_n = __ldg(a.col + k);
            rowid_n = __ldg(a.rowid + k);
        }
        return tt < glen && k < a.E;
    };
    bool valid_n = glen > 0 ? fetch(0) : false;
    for (int64_t tt = 0; tt < glen; ++tt, ++ti) {
        const bool valid = valid_n;
        const int64_t eid = eid_n;
        const int gnode = gnode_n, rowid = rowid_n;
        valid_n = fetch(tt + 1);
        float Z[DA][DB];
        tc_compute_z<DA, DB>(a.sh + eid * a.S + T.sh_off, a.cg + T.cg_off, T.DS, valid, Z);
        float r[WS * M];
        tc_load_gside<DA, DB, WS>(a.g + (int64_t)rowid * a.g_len + T.r_off + T.b0 * DB, nb, valid, Z, r);
        const float* xrow = a.x + (int64_t)gnode * a.x_len + T.v_off;
        float xs[4 * DA];
        tc_load_x4<DA>(xrow, T.a_begin, AR, g * HA, valid, vec_ok, xs);  // in flight while waiting for the buffer
        mbar_wait(&bars[6 + g], (ti & 1u) ^ 1u);
#pragma unroll
        for (int a4 = 0; a4 < HA; a4 += 4) {
            float xn[4 * DA];
            if (a4 + 4 < HA) tc_load_x4<DA>(xrow, T.a_begin, AR, g * HA + a4 + 4, valid, vec_ok, xn);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float y[M];
                tc_factor<DA, DB>(xs + j * DA, Z, y);
                tc_dt_store<M, WS>(A, e, g * 128 + (a4 + j) * WS, r, y);
            }
            if (a4 + 4 < HA) {
#pragma unroll
                for (int v = 0; v < 4 * DA; ++v) xs[v] = xn[v];
            }
        }
        fence_proxy_async();
        mbar_arrive(&bars[4 + g]);
    }
}

__global__ void __launch_bounds__(kGenThreads, 1) tp_dw2_tc_kernel(TcDwArgs a) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + oDwBar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + oDwBar + 9 * 8);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int KS = a.KS, H = 64 * KS;
    const TcWTile T = a.wt[blockIdx.x / a.ngroups];
    if (t == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
        mbar_init(&bars[4], 128); mbar_init(&bars[5], 128); mbar_init(&bars[6], 1); mbar_init(&bars[7], 1);
        mbar_init(&bars[8], 1);
        fence_mbar_init();
    }
    if (warp == 9) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr;

    if (warp < 8) {
        const int e = (warp & 3) * 32 + lane, g = warp >> 2;
        {
            const TcWTile& G = T;
            GMP_TC_DISPATCH(dw_generate, a, T, sm, bars, e, g)
        }
        // ---- D half g, lane = column c of the N-tile = one row of W2
        mbar_wait(&bars[8], 0);
        tc_fence_after();
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int c = g * 128 + e;
        const int al = c / T.WS, b = c - al * T.WS;
        const int aa = T.a_begin + al, bb = T.b0 + b;
        const bool live = aa < T.a_end && bb < T.b_end;
        const int grp = blockIdx.x % a.ngroups;
        const bool any = (a.ntiles * (grp + 1)) / a.ngroups > (a.ntiles * grp) / a.ngroups;  // an empty group wrote nothing to D
        float* dst = a.dW2 + ((int64_t)(blockIdx.x % a.ngroups) * a.numel + (int64_t)T.w_off + (int64_t)aa * T.stride_a + (int64_t)bb * T.stride_b) * H;
        for (int cc = 0; cc < H; cc += 32) {
            float v[32];
            tmem_ld32(tm + lane_base + g * H + cc, v);
            if (live) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + cc + j) = any ? make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    } else if (warp == 8) {
        if (lane == 0) {
            uint32_t ti = 0;
            const int nt = blockIdx.x / a.ngroups, grp = blockIdx.x - nt * a.ngroups;
            const int64_t g0 = (a.ntiles * grp) / a.ngroups, g1 = (a.ntiles * (grp + 1)) / a.ngroups, glen = g1 - g0;
            const int64_t rot = (glen * nt) / a.NT;
            for (int64_t tt = 0; tt < glen; ++tt, ++ti) {
                const int64_t tile = g0 + (rot + tt) % glen;
                const uint32_t buf = ti & 1u;
                mbar_wait(&bars[2 + buf], ((ti >> 1) & 1u) ^ 1u);
                mbar_expect_tx(&bars[buf], (uint32_t)(KS * kStage));
                for (int ks = 0; ks < KS; ++ks)
                    bulk_g2s(sm + oDwH + buf * 65536 + ks * kStage, a.hid_img + (tile * KS + ks) * (int64_t)kStage, kStage, &bars[buf]);
            }
        }
    } else {
        const uint32_t idesc = umma_idesc_bf16(128, H, true, true);
        const uint64_t adesc0 = umma_desc_mn128(smem_u32(sm + oDwA), kStage), bdesc0 = umma_desc_mn128(smem_u32(sm + oDwH), kStage);
        uint32_t ti = 0;
        const int grp = blockIdx.x % a.ngroups;
        const int64_t glen = (a.ntiles * (grp + 1)) / a.ngroups - (a.ntiles * grp) / a.ngroups;
        for (int64_t tt = 0; tt < glen; ++tt, ++ti) {
            const uint32_t buf = ti & 1u;
            mbar_wait(&bars[buf], (ti >> 1) & 1u);
            for (int half = 0; half < 2; ++half) {
                mbar_wait(&bars[4 + half], ti & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ad = adesc0 + (uint64_t)(2 * half * (kStage >> 4)), bd = bdesc0 + (uint64_t)(buf * (65536 >> 4));
#pragma unroll
                    for (int k16 = 0; k16 < 8; ++k16)
                        umma_bf16(tm + half * H, ad + (uint64_t)(k16 * (2048 >> 4)), bd + (uint64_t)(k16 * (2048 >> 4)), idesc, (ti | k16) ? 1u : 0u);
                    umma_commit(&bars[6 + half]);
                    if (half == 1) umma_commit(&bars[2 + buf]);
                }
                __syncwarp();
            }
        }
        if (elect_one()) umma_commit(&bars[8]);
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc<512>(tm);
}

// rows that straddle a chunk boundary: add the later chunks' head partials in chunk order
__global__ void __launch_bounds__(128) tp_tc_fixup_kernel(const int32_t* __restrict__ rowptr, int64_t n, int64_t E, int64_t ntiles, int nchunks,
                                                          const float* __restrict__ head, float* __restrict__ res, int r_len) {
    __shared__ int hrow[1024];  // row that chunk ch's head belongs to, or -1
    for (int ch = threadIdx.x; ch < nchunks; ch += 128) {
        const int64_t e0 = ((ntiles + nchunks - 1) / nchunks) * ch * kET;
        int row = -1;
        if (ch > 0 && e0 < E) {
            const int lo = row_of_edge(rowptr, (int)n, e0);
            if ((int64_t)__ldg(rowptr + lo) < e0) row = lo;
        }
        hrow[ch] = row;
    }
    __syncthreads();
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= r_len) return;
    for (int ch = 1; ch < nchunks; ++ch)
        if (hrow[ch] >= 0) res[(int64_t)hrow[ch] * r_len + c] += head[(int64_t)ch * r_len + c];
}

}  // namespace gmp

using namespace gmp;

extern "C" {

int32_t gmp_tp_tc_num_chunks(int64_t num_edges) {
    const int64_t nt = ceil_div(num_edges, kET);
    return (int32_t)(nt < num_sms() ? (nt < 1 ? 1 : nt) : num_sms());
}

int64_t gmp_tp_tc_hid_bytes(int64_t num_edges, int32_t H) { return ceil_div(num_edges, kET) * (H / 64) * (int64_t)kStage; }
int64_t gmp_tp_tc_w2_bytes(int32_t ntiles, int32_t H) { return (int64_t)ntiles * 2 * (H / 64) * (int64_t)kStage; }

int gmp_tp_tc_pack_hid(const int32_t* perm, int64_t num_edges, const float* edge_feat, int32_t R, const float* w1, const float* b1,
                       int32_t H, void* hid_img, gmp_stream_t stream) {
    if (num_edges == 0) return GMP_OK;
    GMP_REQUIRE(edge_feat && w1 && b1 && hid_img, "tp_tc_pack_hid: NULL pointer");
    GMP_REQUIRE(H >= 64 && H <= 256 && H % 64 == 0, "tp_tc_pack_hid: mlp_dim must be 64, 128, 192 or 256 (got %d)", H);
    GMP_REQUIRE(R >= 1 && R <= 16, "tp_tc_pack_hid: edge_feats_dim in [1, 16] (got %d)", R);
    if (num_edges == 0) return GMP_OK;
    const size_t smem = (size_t)(H * R + H + kET * R) * sizeof(float);
    tp_pack_hid_kernel<<<(unsigned)ceil_div(num_edges, kET), 256, smem, stream>>>(perm, num_edges, edge_feat, R, w1, b1, H, (uint8_t*)hid_img);
    return check_launch("tp_pack_hid_kernel");
}

int gmp_tp_tc_pack_w2(const float* w2, int32_t H, const void* ntile_table, int32_t ntiles, void* w2_img, gmp_stream_t stream) {
    GMP_REQUIRE(w2 && ntile_table && w2_img, "tp_tc_pack_w2: NULL pointer");
    GMP_REQUIRE(H >= 64 && H <= 256 && H % 64 == 0, "tp_tc_pack_w2: mlp_dim must be 64, 128, 192 or 256 (got %d)", H);
    if (ntiles == 0) return GMP_OK;
    tp_pack_w2_kernel<<<(unsigned)(ntiles * 2 * (H / 64)), 256, 0, stream>>>(w2, H, (const TcNTile*)ntile_table, (uint8_t*)w2_img);
    return check_launch("tp_pack_w2_kernel");
}

int gmp_tp_ysum(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t num_edges, const float* V,
                int32_t v_len, const float* edge_sh, int32_t S, const void* ypaths, int32_t npaths, int32_t npairs,
                const void* zentries, int32_t nz, const float* cg, float* YS, int32_t y_len, gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && ypaths && zentries && cg && YS, "tp_ysum: NULL pointer");
    GMP_REQUIRE(num_edges == 0 || (col && V && edge_sh), "tp_ysum: NULL edge/feature pointer");
    GMP_REQUIRE(npaths <= 32 && nz <= kYsMaxZ && npairs <= 128 * kYsPairs,
                "tp_ysum: at most 32 paths, %d geometric factors and %d (path, multiplicity) pairs", kYsMaxZ, 128 * kYsPairs);
    if (n == 0) return GMP_OK;
    YsArgs a;
    a.rowptr = rowptr; a.col = col; a.perm = perm; a.n = n; a.V = V; a.v_len = v_len; a.sh = edge_sh; a.S = S;
    a.paths = (const TcYPath*)ypaths; a.npaths = npaths; a.npairs = npairs; a.zent = (const TcZEntry*)zentries; a.nz = nz;
    a.cg = cg; a.YS = YS; a.y_len = y_len;
    const int64_t grid = n < 16ll * num_sms() ? n : 16ll * num_sms();
    tp_ysum_kernel<<<(unsigned)grid, 128, 0, stream>>>(a);
    return check_launch("tp_ysum_kernel");
}

int gmp_tp_tc_contract(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n, int64_t num_edges, const float* V,
                       int32_t v_len, float* res, int32_t r_len, float* head, const float* edge_sh, int32_t S, const void* hid_img,
                       const void* w2_img, const void* ygroups, int32_t nyg, int32_t ntiles_n, int32_t H, const float* cg,
                       gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && res && head && ygroups && cg, "tp_tc_contract: NULL pointer");
    GMP_REQUIRE(num_edges == 0 || (col && V && edge_sh && hid_img && w2_img), "tp_tc_contract: NULL edge/feature pointer");
    GMP_REQUIRE(H >= 64 && H <= 256 && H % 64 == 0, "tp_tc_contract: mlp_dim must be 64, 128, 192 or 256 (got %d)", H);
    GMP_REQUIRE(n >= 0 && n < (1ll << 31) && num_edges >= 0 && num_edges < (1ll << 31), "tp_tc_contract: sizes out of range");
    GMP_CUDA(cudaMemsetAsync(res, 0, (size_t)n * r_len * sizeof(float), stream));
    if (n == 0 || num_edges == 0 || nyg == 0) return GMP_OK;
    const int nchunks = gmp_tp_tc_num_chunks(num_edges);
    int nchunks_launched = nchunks;
    GMP_CUDA(cudaMemsetAsync(head, 0, (size_t)nchunks * r_len * sizeof(float), stream));
    TcTpArgs a;
    a.rowptr = rowptr; a.col = col; a.perm = perm; a.n = n; a.E = num_edges; a.V = V; a.v_len = v_len; a.res = res; a.r_len = r_len;
    a.head = head; a.sh = edge_sh; a.S = S; a.hid_img = (const uint8_t*)hid_img; a.w2_img = (const uint8_t*)w2_img;
    a.yg = (const TcYGroup*)ygroups; a.nyg = nyg; a.NT = ntiles_n; a.KS = H / 64; a.cg = cg; a.ntiles = ceil_div(num_edges, kET);
    if (nchunks >= 2) {
        // CTA pairs (thread-block clusters of 2) share the W2 stream through multicast
        const int grid = nchunks & ~1;
        GMP_CUDA(cudaFuncSetAttribute(tp_contract_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcTpSmem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kTcThreads);
        cfg.dynamicSmemBytes = kTcTpSmem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        GMP_CUDA(cudaLaunchKernelEx(&cfg, tp_contract_tc_kernel<true>, a));
        nchunks_launched = grid;
    } else {
        GMP_CUDA(cudaFuncSetAttribute(tp_contract_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcTpSmem));
        tp_contract_tc_kernel<false><<<nchunks, kTcThreads, kTcTpSmem, stream>>>(a);
        nchunks_launched = nchunks;
    }
    int rc = check_launch("tp_contract_tc_kernel");
    if (rc != GMP_OK) return rc;
    if (nchunks_launched > 1) {
        tp_tc_fixup_kernel<<<(unsigned)ceil_div(r_len, 128), 128, 0, stream>>>(rowptr, n, num_edges, a.ntiles, nchunks_launched, head, res, r_len);
        rc = check_launch("tp_tc_fixup_kernel");
    }
    return rc;
}

int gmp_tp_tc_dhid(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid, int64_t n, int64_t num_edges, const float* x,
                   int32_t x_len, const float* g, int32_t g_len, const float* edge_sh, int32_t S, const float* edge_feat, int32_t R,
                   const float* w1, const float* b1, const void* w2_img, const void* ygroups, int32_t nyg, int32_t ntiles_n, int32_t H,
                   const float* cg, float* dpre, gmp_stream_t stream) {
    if (num_edges == 0) return GMP_OK;
    GMP_REQUIRE(rowptr && ygroups && cg && w1 && b1 && dpre, "tp_tc_dhid: NULL pointer");
    GMP_REQUIRE(num_edges == 0 || (col && rowid && x && g && edge_sh && edge_feat && w2_img), "tp_tc_dhid: NULL edge/feature pointer");
    GMP_REQUIRE(H >= 64 && H <= 256 && H % 64 == 0, "tp_tc_dhid: mlp_dim must be 64, 128, 192 or 256 (got %d)", H);
    GMP_REQUIRE(R >= 1 && R <= 16, "tp_tc_dhid: edge_feats_dim in [1, 16] (got %d)", R);
    if (n == 0 || num_edges == 0) return GMP_OK;
    if (nyg == 0 || ntiles_n == 0) {
        GMP_CUDA(cudaMemsetAsync(dpre, 0, (size_t)num_edges * H * sizeof(float), stream));
        return GMP_OK;
    }
    TcDhArgs a;
    a.rowptr = rowptr; a.col = col; a.perm = perm; a.rowid = rowid; a.n = n; a.E = num_edges; a.x = x; a.x_len = x_len; a.g = g; a.g_len = g_len;
    a.sh = edge_sh; a.S = S; a.feat = edge_feat; a.R = R; a.w1 = w1; a.b1 = b1; a.w2_img = (const uint8_t*)w2_img;
    a.yg = (const TcYGroup*)ygroups; a.nyg = nyg; a.NT = ntiles_n; a.KS = H / 64; a.cg = cg; a.dpre = dpre;
    a.ntiles = ceil_div(num_edges, kET);
    const int grid = gmp_tp_tc_num_chunks(num_edges);
    GMP_CUDA(cudaFuncSetAttribute(tp_dhid_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcDhSmem));
    tp_dhid_tc_kernel<<<grid, kGenThreads, kTcDhSmem, stream>>>(a);
    return check_launch("tp_dhid_tc_kernel");
}

int gmp_tp_tc_dw2(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid, int64_t n, int64_t num_edges,
                  const float* x, int32_t x_len, const float* g, int32_t g_len, const float* edge_sh, int32_t S, const void* hid_img,
                  const void* wtile_table, int32_t ntiles_n, int32_t H, const float* cg, int64_t numel, int32_t ngroups, float* dW2,
                  gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && wtile_table && cg && dW2, "tp_tc_dw2: NULL pointer");
    GMP_REQUIRE(num_edges == 0 || (col && rowid && x && g && edge_sh && hid_img), "tp_tc_dw2: NULL edge/feature pointer");
    GMP_REQUIRE(ngroups >= 1 && ngroups <= 64, "tp_tc_dw2: ngroups in [1, 64] (got %d)", ngroups);
    GMP_REQUIRE(H >= 64 && H <= 256 && H % 64 == 0, "tp_tc_dw2: mlp_dim must be 64, 128, 192 or 256 (got %d)", H);
    if (ntiles_n == 0) return GMP_OK;
    TcDwArgs a;
    a.rowptr = rowptr; a.col = col; a.perm = perm; a.rowid = rowid; a.n = n; a.E = num_edges; a.x = x; a.x_len = x_len; a.g = g; a.g_len = g_len;
    a.sh = edge_sh; a.S = S; a.hid_img = (const uint8_t*)hid_img; a.wt = (const TcWTile*)wtile_table; a.KS = H / 64; a.cg = cg;
    a.dW2 = dW2; a.ntiles = ceil_div(num_edges, kET); a.numel = numel; a.NT = ntiles_n; a.ngroups = ngroups;
    GMP_CUDA(cudaFuncSetAttribute(tp_dw2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcDwSmem));
    tp_dw2_tc_kernel<<<ntiles_n * ngroups, kGenThreads, kTcDwSmem, stream>>>(a);
    return check_launch("tp_dw2_tc_kernel");
}

}  // extern "C"
