// EGNN edge path, second tcgen05 design (GMP_BF16_TC, emb_dim = 128): models/layers/egnn_layer.py:62-80.
//
// Same mathematics as egnn_tc.cu (pre1 = P[i] + Q[j] + dist * wd, three LayerNorm'd stages, both 128 x 128 edge GEMMs on
// tcgen05.mma with fp32 accumulators in tensor memory), different decomposition:
//
//   * a thread owns one EDGE ROW of a 128-edge tile and walks its 128 columns in four 32-column chunks, twice per stage
//     (pass A: LayerNorm sums, pass B: normalise / activate / pack).  The row statistics are thread-local, so a tile needs
//     no partial-sum exchange and no CTA-wide barrier: the only synchronisation inside a tile is the mbarrier hand-off with
//     the warp that issues its MMAs.  (egnn_tc.cu: thread = (row, column quarter), ~10 full-CTA barriers per tile, one tile
//     in flight per SM, 40 k cycles per tile -- profiles/r02_summary.md.)
//   * three independent STREAMS per CTA, each = four compute warps + one MMA-issuing warp, its own operand buffers, its own
//     128 tensor-memory columns and its own contiguous range of the sorted edge list (a "chunk"): while one stream waits for
//     a product or a gather, the other two compute.  The weight images are shared.
//   * the aggregation over destination rows is a third MMA per tile, D3[c][s] = sum_e m[e][c] * S[e][s] with a one-hot
//     row-membership tile S (as in schnet_tc2.cu); rows that continue into the next tile are carried in registers, rows that
//     straddle a chunk boundary go through a head buffer + fix-up kernel.  Every output element is written by one thread, in
//     tile order: deterministic, no atomics.  The coordinate update (3 columns) is reduced in fp32 by 96 threads.
#include <cuda_pipeline.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gmp {

using namespace tc;

constexpr int kE2Streams = 3;
constexpr int kE2Threads = (4 * kE2Streams + kE2Streams) * 32;   // 12 compute warps + 3 MMA warps = 480
constexpr int kE2Head = 132;                                      // floats per chunk in the head buffer: 128 message columns + 3 + pad

// shared-memory map (bytes from the 1024-aligned base)
constexpr int o4W1 = 0;                                   // W1 image [128 out][2 slabs of 64 in] bf16, 32 KB
constexpr int o4W2 = 32768;
constexpr int o4A = 65536;                                // per stream: operand image (Q row -> a1 -> m), 32 KB
constexpr int o4S = o4A + kE2Streams * 32768;             // per stream: one-hot tile, 16 KB (rows of 128 B, first 64 B used)
constexpr int o4Vec = o4S + kE2Streams * 16384;           // wd g1 be1 b1 g2 be2 b2 g3 be3 w3 (fp32 [10][128])
constexpr int o4PP = o4Vec + 10 * 128 * 4;                // per stream: per-edge coordinate contributions fp32 [128][3]
constexpr int o4Meta = o4PP + kE2Streams * 1536;          // per stream: StreamMeta
constexpr int kE2MetaBytes = (128 + 132 + 4 + 12) * 4;
constexpr int o4Bar = o4Meta + kE2Streams * kE2MetaBytes; // per stream: ready, done
constexpr int kE2Smem = o4Bar + kE2Streams * 16 + 16 + 1024;

enum { V2_WD = 0, V2_G1, V2_BE1, V2_B1, V2_G2, V2_BE2, V2_B2, V2_G3, V2_BE3, V2_W3 };

struct Egnn2Args {
    const int32_t *rowptr, *col, *rowid;
    int64_t n, E;
    const float* P;                  // [n,128] fp32, indexed by the CSR row (destination i): h_i half of the first Linear (+ bias)
    const __nv_bfloat16* Q;          // [n,128] bf16, indexed by col (source j); with `peer`: the owned rows only
    PeerRows peer;                   // halo rows of Q in the neighbouring ranks' memory (n_own = 0: none)
    const float* pos;
    const float *wd, *g1, *be1, *w1, *b1, *g2, *be2, *w2, *b2, *g3, *be3, *w3, *b3;
    int aggr_mean;
    float eps;
    float *msg_aggr, *pos_aggr;      // [n,128], [n,3]: zeroed by the caller (rows without edges stay zero)
    float* head;                     // [chunks][kE2Head], zeroed
};

struct StreamMeta {
    int seg_row[128];
    int seg_start[132];
    int wcnt[4];
    int cmd;
    int cpos_row[3], cpos_head[3];   // open row of the coordinate sums, one copy per component (= per warp: no cross-warp hand-off)
    float cpos[3];
};

template <int ACT> __device__ __forceinline__ float act2(float y) {
    if (ACT == 0) return fmaxf(y, 0.f);
    return y / (1.f + __expf(-y));
}

__device__ __forceinline__ void unpack8_(const uint4 u, float (&f)[8]) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8_(const float* f) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
// eight consecutive parameter values (32-byte aligned) with two 16-byte shared-memory loads
struct P8 { float v[8]; };
__device__ __forceinline__ P8 ldp8(const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    return P8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// byte offset of the 16-byte chunk c16 (0..15: columns 8 c16 .. 8 c16 + 7) of row r in a two-slab K-major operand image
__device__ __forceinline__ uint32_t img_chunk(int r, int c16) { return (uint32_t)((c16 >> 3) * 16384) + sw128_chunk_off(r, c16 & 7); }

template <int ACT>
__global__ void __launch_bounds__(kE2Threads, 1) egnn_fwd_tc2_kernel(Egnn2Args a) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    float* vec = reinterpret_cast<float*>(sm + o4Vec);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + o4Bar + kE2Streams * 16);

    // ---- setup: weight images, vectors, barriers, tensor memory
    for (int x = t; x < 128 * 16; x += kE2Threads) {
        const int f = x >> 4, ch16 = x & 15, kb = ch16 >> 3, ch = ch16 & 7;
        {
            const float4 lo = ldg4(a.w1 + f * 128 + kb * 64 + ch * 8), hi = ldg4(a.w1 + f * 128 + kb * 64 + ch * 8 + 4);
            *reinterpret_cast<uint4*>(sm + o4W1 + kb * 16384 + sw128_chunk_off(f, ch)) =
                make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
        }
        {
            const float4 lo = ldg4(a.w2 + f * 128 + kb * 64 + ch * 8), hi = ldg4(a.w2 + f * 128 + kb * 64 + ch * 8 + 4);
            *reinterpret_cast<uint4*>(sm + o4W2 + kb * 16384 + sw128_chunk_off(f, ch)) =
                make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
        }
    }
    {
        const float* vecs[10] = {a.wd, a.g1, a.be1, a.b1, a.g2, a.be2, a.b2, a.g3, a.be3, a.w3};
        for (int i = t; i < 10 * 128; i += kE2Threads) vec[i] = __ldg(vecs[i >> 7] + (i & 127));
    }
    if (t == 0) {
        for (int s = 0; s < kE2Streams; ++s) {
            uint64_t* bars = reinterpret_cast<uint64_t*>(sm + o4Bar + s * 16);
            mbar_init(&bars[0], 128);   // ready: every compute thread of the stream arrives
            mbar_init(&bars[1], 1);     // done: tcgen05.commit
        }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_ptr);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr;

    const int stream = warp < 4 * kE2Streams ? (warp >> 2) : (warp - 4 * kE2Streams);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + o4Bar + stream * 16);
    StreamMeta& M = *reinterpret_cast<StreamMeta*>(sm + o4Meta + stream * kE2MetaBytes);
    uint8_t* A = sm + o4A + stream * 32768;
    uint8_t* S = sm + o4S + stream * 16384;
    const uint32_t tmD = tm + 128u * stream;
    const int chunk = blockIdx.x * kE2Streams + stream, nchunks = gridDim.x * kE2Streams;
    const int64_t e_begin = (a.E * chunk) / nchunks, e_end = (a.E * (chunk + 1)) / nchunks;

    if (warp >= 4 * kE2Streams) {
        // ===================== MMA issue (one warp per stream): wait for the operands, issue what the command word says ==========
        const uint32_t id_kk = umma_idesc_bf16(128, 128), id_seg = umma_idesc_bf16(128, 32, true, true);
        const uint32_t ab = smem_u32(A), sb = smem_u32(S);
        uint32_t ph = 0;
        for (;;) {
            mbar_wait(&bars[0], ph);
            ph ^= 1u;
            const int cmd = *reinterpret_cast<volatile int*>(&M.cmd);
            if (cmd == 0) break;
            tc_fence_after();
            if (elect_one()) {
                if (cmd == 1) umma_tile(tmD, ab, 16384, smem_u32(sm + o4W1), 16384, 128, id_kk);
                else if (cmd == 2) umma_tile(tmD, ab, 16384, smem_u32(sm + o4W2), 16384, 128, id_kk);
                else {
#pragma unroll
                    for (int k16 = 0; k16 < 8; ++k16)
                        umma_bf16(tmD, umma_desc_mn128(ab + k16 * 2048, 16384), umma_desc_mn128(sb + k16 * 2048, 16384), id_seg, k16 ? 1u : 0u);
                }
                umma_commit(&bars[1]);
            }
            __syncwarp();
        }
    } else {
        // ===================== compute: thread = edge slot r of the stream's current tile ======================================
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t ph_done = 0;
        auto issue = [&](int cmd) {   // operands written (generic proxy) -> visible to the tensor core -> hand over, wait for the result
            fence_proxy_async();
            tc_fence_before();
            if (r == 0) M.cmd = cmd;
            mbar_arrive(&bars[0]);
            mbar_wait(&bars[1], ph_done);
            ph_done ^= 1u;
            tc_fence_after();
        };
        // open row carried across tiles: message column r (every thread), coordinates (smem, lanes of warps 0-2)
        float carry = 0.f;
        int carry_row = -1;
        bool carry_head = false;
        if (r < 3) { M.cpos_row[r] = -1; M.cpos_head[r] = 0; }   // read again only after the first tile's barriers
        auto inv_deg = [&](int row) {
            const float d = (float)(__ldg(a.rowptr + row + 1) - __ldg(a.rowptr + row));
            return d > 0.f ? 1.f / d : 0.f;
        };
        auto store_msg = [&](int row, bool head, float val) {
            if (a.aggr_mean) val *= inv_deg(row);
            float* dst = head ? a.head + (int64_t)chunk * kE2Head + r : a.msg_aggr + (int64_t)row * 128 + r;
            *dst = val;
        };
        auto store_pos = [&](int row, bool head, int comp, float val) {
            val *= inv_deg(row);
            float* dst = head ? a.head + (int64_t)chunk * kE2Head + 128 + comp : a.pos_aggr + (int64_t)row * 3 + comp;
            *dst = val;
        };
        const float eps = a.eps;
        const float b3 = __ldg(a.b3);
        float* ppbuf = reinterpret_cast<float*>(sm + o4PP + stream * 1536);

        for (int64_t e0 = e_begin; e0 < e_end; e0 += 128) {
            const int cnt = (int)min((int64_t)128, e_end - e0);
            const bool valid = r < cnt;
            const int64_t k = e0 + min(r, cnt - 1);
            const int rid = __ldg(a.rowid + k), src = __ldg(a.col + k);
            const int ridp = (r > 0 && valid) ? __ldg(a.rowid + k - 1) : -1;
            const bool flag = valid && (r == 0 || ridp != rid);
            // gather this edge's Q row (bf16, 256 B) into its own row of the operand image: 16 asynchronous 16-byte copies
            {
                const __nv_bfloat16* qrow = reinterpret_cast<const __nv_bfloat16*>(peer_row(a.Q, a.peer, src, 256));
#pragma unroll
                for (int c16 = 0; c16 < 16; ++c16) cp_async16(A + img_chunk(r, c16), qrow + c16 * 8);
            }
            float dx = 0.f, dy = 0.f, dz = 0.f, dist = 0.f;
            if (valid) {
                dx = __ldg(a.pos + 3 * (int64_t)rid) - __ldg(a.pos + 3 * (int64_t)src);
                dy = __ldg(a.pos + 3 * (int64_t)rid + 1) - __ldg(a.pos + 3 * (int64_t)src + 1);
                dz = __ldg(a.pos + 3 * (int64_t)rid + 2) - __ldg(a.pos + 3 * (int64_t)src + 2);
                dist = sqrtf(dx * dx + dy * dy + dz * dz);
            }
            // segmentation of the tile by destination row (the one barrier among the stream's four warps per tile)
            const unsigned bal = __ballot_sync(0xffffffffu, flag);
            if (lane == 0) M.wcnt[q] = __popc(bal);
            bar_sync_named(1 + stream, 128);
            int base = 0, nseg = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int c = M.wcnt[w];
                if (w < q) base += c;
                nseg += c;
            }
            const int seg = base + __popc(bal & ((2u << lane) - 1u)) - 1;   // 0-based segment of this slot (valid slots)
            if (flag) { M.seg_row[seg] = rid; M.seg_start[seg] = r; }
            if (r == cnt - 1) M.seg_start[nseg] = cnt;

            // ---------------- stage 1: pre1 = Q[j] + P[i] + dist * wd -> LN1 -> act -> a1 (in place over the Q row) ----------------
            cp_async_wait_all();
            const float4* prow = reinterpret_cast<const float4*>(a.P + (int64_t)rid * 128);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int c16 = 0; c16 < 16; ++c16) {
                float g[8];
                unpack8_(*reinterpret_cast<const uint4*>(A + img_chunk(r, c16)), g);
                const float4 p0 = __ldg(prow + 2 * c16), p1 = __ldg(prow + 2 * c16 + 1);
                const P8 ww = ldp8(vec + V2_WD * 128 + 8 * c16);
                const float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float x = fmaf(dist, ww.v[u], g[u] + pp[u]);
                    s1 += x;
                    s2 = fmaf(x, x, s2);
                }
            }
            {
                const float mean = s1 * (1.f / 128.f);
                const float rstd = 1.f / sqrtf(fmaxf(s2 * (1.f / 128.f) - mean * mean, 0.f) + eps);
#pragma unroll
                for (int c16 = 0; c16 < 16; ++c16) {
                    uint4* pa = reinterpret_cast<uint4*>(A + img_chunk(r, c16));
                    float g[8], o[8];
                    unpack8_(*pa, g);
                    const float4 p0 = __ldg(prow + 2 * c16), p1 = __ldg(prow + 2 * c16 + 1);
                    const P8 ww = ldp8(vec + V2_WD * 128 + 8 * c16), g1 = ldp8(vec + V2_G1 * 128 + 8 * c16), e1 = ldp8(vec + V2_BE1 * 128 + 8 * c16);
                    const float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float x = fmaf(dist, ww.v[u], g[u] + pp[u]);
                        o[u] = valid ? act2<ACT>(fmaf((x - mean) * rstd, g1.v[u], e1.v[u])) : 0.f;
                    }
                    *pa = pack8_(o);
                }
            }
            issue(1);   // D = a1 W1^T
            // ---------------- stage 2: pre2 = D + b1 -> LN2 -> act -> m (over a1: GEMM 1 is complete) ----------------
            s1 = 0.f; s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v[32];
                tmem_ld32(tmD + lane_base + 32 * j, v);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const P8 bb = ldp8(vec + V2_B1 * 128 + 32 * j + 8 * c);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float x = v[8 * c + u] + bb.v[u];
                        s1 += x;
                        s2 = fmaf(x, x, s2);
                    }
                }
            }
            {
                const float mean = s1 * (1.f / 128.f);
                const float rstd = 1.f / sqrtf(fmaxf(s2 * (1.f / 128.f) - mean * mean, 0.f) + eps);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v[32];
                    tmem_ld32(tmD + lane_base + 32 * j, v);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int col = 32 * j + 8 * c;
                        const P8 bb = ldp8(vec + V2_B1 * 128 + col), gg = ldp8(vec + V2_G2 * 128 + col), ee = ldp8(vec + V2_BE2 * 128 + col);
                        float o[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) o[u] = valid ? act2<ACT>(fmaf((v[8 * c + u] + bb.v[u] - mean) * rstd, gg.v[u], ee.v[u])) : 0.f;
                        *reinterpret_cast<uint4*>(A + img_chunk(r, 4 * j + c)) = pack8_(o);
                    }
                }
            }
            issue(2);   // D = m W2^T
            // ---------------- stage 3: pre3 = D + b2 -> LN3 -> act -> s = . w3 + b3; coordinate contribution delta * s ----------------
            s1 = 0.f; s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v[32];
                tmem_ld32(tmD + lane_base + 32 * j, v);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const P8 bb = ldp8(vec + V2_B2 * 128 + 32 * j + 8 * c);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float x = v[8 * c + u] + bb.v[u];
                        s1 += x;
                        s2 = fmaf(x, x, s2);
                    }
                }
            }
            float sdot = 0.f;
            {
                const float mean = s1 * (1.f / 128.f);
                const float rstd = 1.f / sqrtf(fmaxf(s2 * (1.f / 128.f) - mean * mean, 0.f) + eps);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v[32];
                    tmem_ld32(tmD + lane_base + 32 * j, v);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int col = 32 * j + 8 * c;
                        const P8 bb = ldp8(vec + V2_B2 * 128 + col), gg = ldp8(vec + V2_G3 * 128 + col), ee = ldp8(vec + V2_BE3 * 128 + col),
                                 w3 = ldp8(vec + V2_W3 * 128 + col);
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            sdot = fmaf(act2<ACT>(fmaf((v[8 * c + u] + bb.v[u] - mean) * rstd, gg.v[u], ee.v[u])), w3.v[u], sdot);
                    }
                }
            }
            {
                const float sc = valid ? sdot + b3 : 0.f;
                ppbuf[3 * r] = dx * sc; ppbuf[3 * r + 1] = dy * sc; ppbuf[3 * r + 2] = dz * sc;
            }
            // ---------------- aggregation: one-hot MMA per window of 32 segments, carried read-out ----------------
            for (int lo = 0; lo < nseg; lo += 32) {
                const int sl = valid ? seg - lo : -1;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t o[4] = {0u, 0u, 0u, 0u};
                    if (sl >= 0 && sl < 32 && (sl >> 3) == c) o[(sl & 7) >> 1] = (sl & 1) ? 0x3f800000u : 0x00003f80u;
                    *reinterpret_cast<uint4*>(S + sw128_chunk_off(r, c)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                issue(3);   // D[c][s] = sum_e m[e][c] S[e][s]   (the hand-off also orders the pp / seg_row / seg_start writes before the reads below)
                float v[32];
                tmem_ld32(tmD + lane_base, v);
                const int hi = min(nseg, lo + 32);
                const bool head0 = (e0 == e_begin) && ((int64_t)__ldg(a.rowptr + M.seg_row[0]) < e_begin);   // the chunk starts inside a row
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int s = lo + j;
                    if (s < hi) {
                        const int row = M.seg_row[s];
                        float val = v[j];
                        bool head = false;
                        if (s == 0) {
                            head = head0;
                            if (carry_row == row) { val += carry; head = carry_head; }
                            else if (carry_row >= 0) store_msg(carry_row, carry_head, carry);
                        }
                        if (s == nseg - 1) { carry = val; carry_row = row; carry_head = head; }
                        else store_msg(row, head, val);
                    }
                }
                // coordinates: warp q = component (q < 3), lane = segment of the window; fp32 sums over the segment's edges
                if (q < 3) {
                    const int s = lo + lane;
                    const bool on = s < hi;
                    float val = 0.f;
                    int row = -1;
                    bool head = false;
                    if (on) {
                        row = M.seg_row[s];
                        const int b = M.seg_start[s], e = M.seg_start[s + 1];
                        for (int x = b; x < e; ++x) val += ppbuf[3 * x + q];
                        if (s == 0) {
                            head = head0;
                            const int crow = M.cpos_row[q];
                            if (crow == row) { val += M.cpos[q]; head = M.cpos_head[q] != 0; }
                            else if (crow >= 0) store_pos(crow, M.cpos_head[q] != 0, q, M.cpos[q]);
                        }
                    }
                    __syncwarp();
                    if (on) {
                        if (s == nseg - 1) { M.cpos[q] = val; M.cpos_row[q] = row; M.cpos_head[q] = head ? 1 : 0; }
                        else store_pos(row, head, q, val);
                    }
                    __syncwarp();
                }
            }
        }
        // the rows left open by the last tile
        if (carry_row >= 0) store_msg(carry_row, carry_head, carry);
        if (q < 3) {
            __syncwarp();
            if (lane == 0 && M.cpos_row[q] >= 0) store_pos(M.cpos_row[q], M.cpos_head[q] != 0, q, M.cpos[q]);
        }
        // tell the MMA warp to leave
        if (r == 0) M.cmd = 0;
        mbar_arrive(&bars[0]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

// rows that straddle a chunk boundary: add the later chunks' head partials in chunk order (one block per chunk; the block of
// the FIRST chunk whose head belongs to a row adds every consecutive head of that row)
__device__ __forceinline__ int e2_head_row(const int32_t* __restrict__ rowptr, int64_t n, int64_t E, int nchunks, int ch) {
    if (ch <= 0 || ch >= nchunks) return -1;
    const int64_t e0 = (E * ch) / nchunks, e1 = (E * (ch + 1)) / nchunks;
    if (e0 >= e1) return -1;
    int lo = 0, hi = (int)n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(rowptr + mid) <= e0) lo = mid; else hi = mid;
    }
    return (int64_t)__ldg(rowptr + lo) < e0 ? lo : -1;
}

__global__ void __launch_bounds__(160) egnn_tc2_fixup_kernel(const int32_t* __restrict__ rowptr, int64_t n, int64_t E, int nchunks,
                                                             const float* __restrict__ head, float* __restrict__ msg, float* __restrict__ pos) {
    const int ch = blockIdx.x + 1;
    const int row = e2_head_row(rowptr, n, E, nchunks, ch);
    if (row < 0 || e2_head_row(rowptr, n, E, nchunks, ch - 1) == row) return;
    const int c = threadIdx.x;
    if (c >= 131) return;
    float* dst = c < 128 ? msg + (int64_t)row * 128 + c : pos + (int64_t)row * 3 + (c - 128);
    float v = *dst;
    for (int c2 = ch; c2 < nchunks && (c2 == ch || e2_head_row(rowptr, n, E, nchunks, c2) == row); ++c2) v += head[(int64_t)c2 * kE2Head + c];
    *dst = v;
}

static int egnn2_grid(int64_t E) {
    const int64_t tiles = ceil_div(E > 0 ? E : 1, 128);
    const int64_t g = ceil_div(tiles, kE2Streams);
    return (int)(g < num_sms() ? (g < 1 ? 1 : g) : num_sms());
}

}  // namespace gmp

using namespace gmp;

extern "C" {

int32_t gmp_egnn_tc2_num_chunks(int64_t num_edges) { return egnn2_grid(num_edges) * kE2Streams; }

int gmp_egnn_tc2_edge_fwd(const int32_t* rowptr, const int32_t* col, const int32_t* rowid, int64_t n, int64_t num_edges, const float* P,
                          const void* Q_bf16, const float* pos, const gmp_egnn_edge_params* p, float* msg_aggr, float* pos_aggr,
                          float* head, const gmp_peer_rows* peer, gmp_stream_t stream) {
    GMP_REQUIRE(p, "egnn_tc2: params is NULL");
    GMP_REQUIRE(!peer || (peer->n_left >= 0 && peer->n_own > 0 && peer->n_left + peer->n_own <= n && (peer->n_left == 0 || peer->left) &&
                          (peer->n_left + peer->n_own == n || peer->right)), "egnn_tc2: inconsistent peer rows");
    GMP_REQUIRE(p->d == 128, "egnn_tc2: the tensor-core path is built for emb_dim = 128 (got %d)", p->d);
    GMP_REQUIRE(p->act == 0 || p->act == 1, "egnn_tc2: act must be 0 (relu) or 1 (swish)");
    GMP_REQUIRE(p->wd && p->ln1_g && p->ln1_b && p->w1 && p->b1 && p->ln2_g && p->ln2_b && p->w2 && p->b2 && p->ln3_g && p->ln3_b &&
                p->w3 && p->b3, "egnn_tc2: NULL parameter pointer");
    GMP_REQUIRE(n >= 0 && num_edges >= 0 && n < (1ll << 31) && num_edges < (1ll << 31), "egnn_tc2: sizes out of range");
    GMP_REQUIRE(rowptr && msg_aggr && pos_aggr && pos && head && (num_edges == 0 || (col && rowid && P && Q_bf16)), "egnn_tc2_edge_fwd: NULL pointer");
    if (n == 0) return GMP_OK;
    GMP_CUDA(cudaMemsetAsync(msg_aggr, 0, (size_t)n * 128 * sizeof(float), stream));
    GMP_CUDA(cudaMemsetAsync(pos_aggr, 0, (size_t)n * 3 * sizeof(float), stream));
    if (num_edges == 0) return GMP_OK;
    const int grid = egnn2_grid(num_edges), nchunks = grid * kE2Streams;
    GMP_CUDA(cudaMemsetAsync(head, 0, (size_t)nchunks * kE2Head * sizeof(float), stream));
    Egnn2Args a;
    a.rowptr = rowptr; a.col = col; a.rowid = rowid; a.n = n; a.E = num_edges; a.P = P; a.Q = (const __nv_bfloat16*)Q_bf16; a.pos = pos;
    a.peer = make_peer_rows(peer);
    a.wd = p->wd; a.g1 = p->ln1_g; a.be1 = p->ln1_b; a.w1 = p->w1; a.b1 = p->b1; a.g2 = p->ln2_g; a.be2 = p->ln2_b;
    a.w2 = p->w2; a.b2 = p->b2; a.g3 = p->ln3_g; a.be3 = p->ln3_b; a.w3 = p->w3; a.b3 = p->b3;
    a.aggr_mean = p->aggr_mean; a.eps = p->ln_eps; a.msg_aggr = msg_aggr; a.pos_aggr = pos_aggr; a.head = head;
    if (p->act) {
        GMP_CUDA(cudaFuncSetAttribute(egnn_fwd_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kE2Smem));
        egnn_fwd_tc2_kernel<1><<<grid, kE2Threads, kE2Smem, stream>>>(a);
    } else {
        GMP_CUDA(cudaFuncSetAttribute(egnn_fwd_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kE2Smem));
        egnn_fwd_tc2_kernel<0><<<grid, kE2Threads, kE2Smem, stream>>>(a);
    }
    int rc = check_launch("egnn_fwd_tc2_kernel");
    if (rc != GMP_OK) return rc;
    if (nchunks > 1) {
        egnn_tc2_fixup_kernel<<<nchunks - 1, 160, 0, stream>>>(rowptr, n, num_edges, nchunks, head, msg_aggr, pos_aggr);
        rc = check_launch("egnn_tc2_fixup_kernel");
    }
    return rc;
}

}  // extern "C"
