// SchNet CFConv forward, pipelined (GMP_BF16_TC, F = 128, lazily recomputed Gaussian basis): the warp-specialised
// successor of schnet_fwd_tc_kernel.  Per tile of <= 128 destination-sorted edges (and <= 32 distinct rows):
//
//   warps 0-3   meta : edge scalars, cosine cutoff, row segments; cp.async gather of the bf16 x1[src] rows straight into
//                      the swizzled operand image; Gaussian basis -> A1 (bf16)
//   warps 20-22 MMA  : G1 = rbf W1^T -> D1;  G2 = h1 W2^T -> D2;  G3 = msg^T S -> D3   (tcgen05, accumulators in TMEM;
//                      one issuing warp per product so that no product waits behind another one's operands)
//   warps 4-11  epi1 : D1 (+ b1, folded into a spare basis column) -> shifted softplus -> h1 (bf16) -> A2
//   warps 12-19 epi2 : (D2 + b2) * C * x1[src] -> msg (bf16, in place over the gathered rows) and the one-hot
//                      row-membership tile S; then, one tile behind, D3[c, seg] -> agg[row(seg), c] += ...
//                      (both epilogues as two groups of four warps: thread = (edge row, 64 columns))
//
// The segmented sum over the destination rows is the third MMA: D3[c][s] = sum_e msg[e][c] * S[e][s] with S[e][s] = 1
// when edge e belongs to the s-th row of the tile (both operands are read as MN-major images whose rows are edges).
// Every (row, column) of agg is only ever touched by the one thread that owns the column in the one CTA that owns the
// edge range, in tile order: deterministic, no atomics; agg is zeroed first and rows that straddle a CTA boundary go
// through the head buffer + tp_tc_fixup_kernel-style fix-up (gmp_schnet_cfconv_fwd_tc2 does both).
// Stages: gathered rows / msg, meta blocks and D3 three deep; D1 two deep; A1, A2 (released per K slab), S, D2 single.
#include <cuda_pipeline.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gmp {

using namespace tc;

constexpr int kS2Threads = 864;   // 27 warps: meta 0-7, epi1 8-15, epi2 16-23 (two groups each: columns 0-63 / 64-127), MMA 24-26
constexpr int kS2MaxSeg = 32;
constexpr int kS2Stages = 3;      // gathered-row / msg stages, D3 accumulators
constexpr int kS2MetaStages = 6;  // meta blocks: small, so the scalar side of a tile can be published well ahead of its row stage

struct Tc2Args {
    const int32_t *rowptr, *col, *perm, *rowid;
    int64_t n, E;
    const float* ew;                 // [E] distances, caller's edge order
    const __nv_bfloat16* x1;         // [n,128] bf16
    const float *w1, *b1, *w2, *b2, *goff;
    int G;
    float cutoff, gcoeff;
    float* agg;                      // [n,128], zeroed
    float* head;                     // [gridDim.x,128], zeroed
    int* dbg;                        // debug builds (GMP_TC2_PROGRESS): host-mapped progress words, [gridDim.x][32 warps]
    __nv_bfloat16* wout;             // optional [E,128]: the filter value W(e) * C(e) of every edge ...
    const int32_t* wrow;             // ... at row wrow[edge id] (NULL: the caller's edge id itself)
};

// shared-memory map (bytes from the 1024-aligned base)
constexpr int o2W1 = 0;                     // [128][64]  bf16 image, 16 KB
constexpr int o2W2 = 16384;                 // 2 slabs [128][64], 32 KB
constexpr int o2A1 = 49152;                 // 16 KB
constexpr int o2A2 = o2A1 + 16384;          // 32 KB
constexpr int o2X = o2A2 + 32768;           // 3 stages x 32 KB: gathered rows -> msg
constexpr int o2S = o2X + kS2Stages * 32768;  // 16 KB (rows of 128 B, first 64 B used: 32 segments)
constexpr int o2Vec = o2S + 16384;          // b1[128] b2[128] goff[64]
constexpr int o2Meta = o2Vec + 320 * 4;     // 6 x { C[128] f32, d[128] f32, seg[128] i32, eid[128] i32, seg_row[32] i32, cnt, nseg, head0, pad }
constexpr int kMetaBytes = (128 + 128 + 128 + 128 + 32 + 4) * 4;
constexpr int o2Tmp = o2Meta + kS2MetaStages * kMetaBytes;   // meta scratch per group: wcount[4], first-overflow[4]; then the end-of-stream tile numbers
constexpr int o2Bar = o2Tmp + 128;
// barriers
enum { B_A1F = 0, B_A1E = 1, B_XF = 2, B_D1F = 5, B_D1E = 7, B_A2F = 9, B_A2E = 10, B_D2F = 12, B_D2E = 13, B_MSGF = 14,
       B_D3F = 17, B_D3E = 20, B_DONE = 23, B_MF = 29, B_COUNT = 35 };
constexpr int kTc2Smem = o2Bar + B_COUNT * 8 + 16 + 1024;

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// softplus(x) - ln 2 = max(x, 0) + log1p(exp(-|x|)) - ln 2: one MUFU (ex2) and u * (degree-3 polynomial) for log1p(u) on
// [0, 1] (minimax fit, |error| <= 7.2e-5: far inside the bf16 rounding of the result, which is all the next GEMM sees)
__device__ __forceinline__ float ssp2(float x) {
    const float u = ex2a(-fabsf(x) * 1.4426950408889634f);
    float p = fmaf(u, -0.058759499f, 0.22568959f);
    p = fmaf(u, p, -0.47130314f);
    p = fmaf(u, p, 0.99744922f);
    return fmaf(u, p, fmaxf(x, 0.f) - 0.6931471805599453f);
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#ifdef GMP_TC2_PROGRESS
// last wait site each warp entered: readable from the host while the kernel hangs
#define TC2_MARK(site, tcv) do { if (lane == 0 && a.dbg) { *((volatile int*)a.dbg + blockIdx.x * 32 + warp) = ((site) << 20) | (int)((tcv) & 0xfffffu); __threadfence_system(); } } while (0)
static int* g_tc2_dbg = nullptr;
#else
#define TC2_MARK(site, tcv) do { } while (0)
#endif

struct Meta {
    float C[128];
    float d[128];     // edge length (1e18 for the padding slots: every Gaussian underflows to 0)
    int seg[128];
    int eid[128];     // row of the optional filter output that belongs to the slot
    int seg_row[32];
    int cnt, nseg, head0, pad;
};

__global__ void __launch_bounds__(kS2Threads, 1) schnet_fwd_tc2_kernel(Tc2Args a) {  // 864 threads: at most 72 registers
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    float* b1s = reinterpret_cast<float*>(sm + o2Vec);
    float* b2s = b1s + 128;
    float* goff = b2s + 128;
    Meta* meta = reinterpret_cast<Meta*>(sm + o2Meta);
    int* tmp = reinterpret_cast<int*>(sm + o2Tmp);
    volatile int* end_g1 = tmp + 16;   // tile number of the end-of-stream marker, passed G1 -> epi1 -> G2
    volatile int* end_e1 = tmp + 17;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + o2Bar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + o2Bar + B_COUNT * 8);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    TC2_MARK(199, 0);

    // ---- setup: weights -> bf16 images, vectors, barriers, tensor memory
    for (int x = t; x < 128 * 8; x += kS2Threads) {  // W1 [f][g], g padded to 64
        const int f = x >> 3, ch = x & 7;
        uint32_t p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g0 = ch * 8 + 2 * j;
            // column G (the first padding column, when there is one) carries b1: the basis tile holds a 1 there
            const float lo = g0 < a.G ? __ldg(a.w1 + f * a.G + g0) : (g0 == a.G ? __ldg(a.b1 + f) : 0.f);
            const float hi = g0 + 1 < a.G ? __ldg(a.w1 + f * a.G + g0 + 1) : (g0 + 1 == a.G ? __ldg(a.b1 + f) : 0.f);
            p[j] = pack_bf16(lo, hi);
        }
        *reinterpret_cast<uint4*>(sm + o2W1 + sw128_chunk_off(f, ch)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
    for (int x = t; x < 128 * 16; x += kS2Threads) {  // W2 [f'][f], two K slabs
        const int f = x >> 4, ch16 = x & 15, kb = ch16 >> 3, ch = ch16 & 7;
        const float4 lo = ldg4(a.w2 + f * 128 + kb * 64 + ch * 8), hi = ldg4(a.w2 + f * 128 + kb * 64 + ch * 8 + 4);
        *reinterpret_cast<uint4*>(sm + o2W2 + kb * 16384 + sw128_chunk_off(f, ch)) =
            make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
    }
    if (t < 128) { b1s[t] = a.G < 64 ? 0.f : __ldg(a.b1 + t); b2s[t] = __ldg(a.b2 + t); }
    if (t < 128) {
        // b2 rides in GEMM 2 as one more K16 step: A slab = a column of ones, B slab = b2 in its first K column.  Both live in
        // the half of the one-hot tile's 128-byte rows that the 32 segments never use (logical 16-byte chunks 4-5 and 6-7).
        // (two K columns: b2 = hi + lo in bf16, so the bias keeps ~16 mantissa bits)
        const float bias = __ldg(a.b2 + t);
        const float hi = __bfloat162float(__float2bfloat16_rn(bias));
        *reinterpret_cast<uint4*>(sm + o2S + sw128_chunk_off(t, 4)) = make_uint4(0x3f803f80u, 0u, 0u, 0u);   // bf16 1.0, 1.0
        *reinterpret_cast<uint4*>(sm + o2S + sw128_chunk_off(t, 5)) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(sm + o2S + sw128_chunk_off(t, 6)) = make_uint4(pack_bf16(hi, bias - hi), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(sm + o2S + sw128_chunk_off(t, 7)) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (t < 64) goff[t] = t < a.G ? __ldg(a.goff + t) : 1.0e18f;  // padding columns: the Gaussian underflows to exactly 0
    if (t == 0) {
        *end_g1 = -1;
        *end_e1 = -1;
        for (int i = 0; i < B_COUNT; ++i) mbar_init(&bars[i], 1);
        for (int i = 0; i < kS2Stages; ++i) mbar_init(&bars[B_D3E + i], 128);
        for (int i = 0; i < kS2MetaStages; ++i) { mbar_init(&bars[B_DONE + i], 128); mbar_init(&bars[B_MF + i], 128); }
        const int c256[] = {B_A1F, B_D1E, B_D1E + 1, B_A2F, B_D2E, B_MSGF, B_MSGF + 1, B_MSGF + 2};   // both groups arrive
        for (int i = 0; i < 8; ++i) mbar_init(&bars[c256[i]], 256);
        for (int i = 0; i < kS2Stages; ++i) mbar_init(&bars[B_XF + i], 512);  // 256 cp.async completions + 256 plain arrives
        fence_mbar_init();
    }
    if (warp == 24) tmem_alloc<512>(tmem_ptr);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr;
    const uint32_t tmD1 = tm, tmD2 = tm + 256, tmD3 = tm + 384;  // D1[p] = tmD1 + 128 p, D3[st] = tmD3 + 32 st
    TC2_MARK(200, 0);

    const int64_t e_begin = (a.E * blockIdx.x) / gridDim.x, e_end = (a.E * (blockIdx.x + 1)) / gridDim.x;

    if (warp < 8) {
        // ===================== meta: thread = (edge slot, half).  Both halves derive the tile's segmentation (cheap, and
        // it saves a hand-off); half 0 publishes the meta block; each half gathers one K slab of the row and expands 32 of the
        // 64 basis columns -- the meta work is the longest serial chain per tile, so it gets two threads per edge
        const int e = t & 127, mh = t >> 7, mw = warp & 3;
        int* mtmp = tmp + 8 * mh;
        const float cw = 3.14159265358979323846f / a.cutoff;
        const float c2 = a.gcoeff * 1.4426950408889634f;
        int64_t e_cur = e_begin;
        // Scalars of a tile (CSR row, edge id, source, length, first row pointer) are fetched one tile ahead, speculating
        // that the current tile is not cut short: two dependent L2 round trips per tile that would otherwise sit on the
        // meta warps' critical path.  Level 1 = indexed by the sorted position, level 2 = indexed by what level 1 returned.
        struct Pre { int rid, ridp, eid, src, rp0, wpos; float d; };
        auto fetch1 = [&](int64_t e0, Pre& P) {
            const int64_t rem = e_end - e0;
            P.rid = 0; P.ridp = -1; P.eid = 0; P.src = 0;
            if (rem > 0) {
                const int c = (int)min((int64_t)128, rem);
                const int64_t k = e0 + min(e, c - 1);
                P.rid = __ldg(a.rowid + k);
                if (e > 0 && e < c) P.ridp = __ldg(a.rowid + k - 1);
                P.eid = a.perm ? __ldg(a.perm + k) : (int)k;
                P.src = __ldg(a.col + k);
            }
        };
        auto fetch2 = [&](int64_t e0, Pre& P) {
            P.d = 1.0e18f; P.rp0 = 0; P.wpos = P.eid;
            if (e_end - e0 > 0) {
                P.d = __ldg(a.ew + P.eid);
                if (a.wrow) P.wpos = __ldg(a.wrow + P.eid);
                if (e == 0) P.rp0 = __ldg(a.rowptr + P.rid);
            }
        };
        Pre cur, nxt;
        fetch1(e_begin, cur);
        fetch2(e_begin, cur);
        for (uint32_t tc = 0;; ++tc) {
            const uint32_t st = tc % kS2Stages, par = (tc / kS2Stages) & 1u;
            const int64_t remain = e_end - e_cur;
            fetch1(e_cur + 128, nxt);
            int cnt = (int)min((int64_t)128, remain);
            const int rid = cur.rid, src = cur.src;
            int seg = 0;
            bool flag = false, valid = false;
            float d = 1.0e18f, C = 0.f;
            if (remain > 0) {
                flag = e < cnt && cur.ridp != rid;
                const unsigned bal = __ballot_sync(0xffffffffu, flag);
                if (lane == 0) mtmp[mw] = __popc(bal);
                bar_sync_named(1 + mh, 128);
                int base = 0;
#pragma unroll
                for (int w = 0; w < 4; ++w)
                    if (w < mw) base += mtmp[w];
                seg = base + __popc(bal & ((2u << lane) - 1u)) - 1;   // 0-based segment of this slot
                // at most 32 rows per tile: cut the tile in front of the 33rd row
                const unsigned over = __ballot_sync(0xffffffffu, e < cnt && seg >= kS2MaxSeg);
                if (lane == 0) mtmp[4 + mw] = over ? mw * 32 + (__ffs(over) - 1) : 128;
                bar_sync_named(1 + mh, 128);
                cnt = min(cnt, min(min(mtmp[4], mtmp[5]), min(mtmp[6], mtmp[7])));
                valid = e < cnt;
                if (valid) {
                    d = cur.d;
                    C = 0.5f * (__cosf(d * cw) + 1.0f);
                }
            }
            const int head0 = (remain > 0 && e == 0) ? ((int64_t)cur.rp0 < e_begin ? 1 : 0) : 0;
            // (the scratch words are not rewritten before every thread of the group has passed the first barrier of the
            // next tile, which is after its last read here)
            const uint32_t ms = tc % kS2MetaStages;
            TC2_MARK(1, tc);
            mbar_wait(&bars[B_DONE + ms], ((tc / kS2MetaStages) & 1u) ^ 1u);   // tile tc-6 read out: its meta block is free
            // The meta block is published BEFORE the wait for the row stage: the scalar side of the tile (basis -> G1 -> softplus
            // -> G2) then runs while the stage is still held by tile tc-3, i.e. a row stage is occupied from the gather to G3
            // only, not through the first half of the chain (epilogue 1 spent 15 % of its samples waiting for this block).
            Meta& M = meta[ms];
            if (remain <= 0) {  // end-of-stream marker: cnt = 0
                if (t == 0) { M.cnt = 0; M.nseg = 0; }
                if (mh == 0) mbar_arrive(&bars[B_MF + ms]);
                TC2_MARK(2, tc);
                mbar_wait(&bars[B_D3F + st], par ^ 1u);
                cp_async_arrive(&bars[B_XF + st]);
                mbar_arrive(&bars[B_XF + st]);
                break;
            }
            if (mh == 0) {
                M.C[e] = C;
                M.d[e] = d;
                M.seg[e] = valid ? seg : -1;
                M.eid[e] = cur.wpos;
                if (flag && seg < kS2MaxSeg && valid) M.seg_row[seg] = rid;
                if (e == cnt - 1) { M.cnt = cnt; M.nseg = seg + 1; }
                if (e == 0) M.head0 = head0;
                mbar_arrive(&bars[B_MF + ms]);   // meta block published: epilogue 1 expands the basis from d
            }
            TC2_MARK(2, tc);
            mbar_wait(&bars[B_D3F + st], par ^ 1u);    // G3 of tile tc-3 has read the msg image: the row stage is free
            // gather x1[src] (bf16 rows of 256 B) into the swizzled row image, half a warp per row: one copy instruction then
            // touches 4 cache lines instead of 32 (a lane-per-row gather costs the load/store unit one cycle per line, and this
            // kernel is short of exactly those).  Meta half mh takes rows [16 mh, 16 mh + 16) of its warp's 32 edge slots.
            {
                uint8_t* xs = sm + o2X + st * 32768;
                const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                const int cchunk = lane & 15, sub = lane >> 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rl = mh * 16 + 2 * i + sub;                     // row within the warp's 32 slots
                    const int srow = __shfl_sync(0xffffffffu, src, rl);
                    uint8_t* dst = xs + (cchunk >> 3) * 16384 + sw128_chunk_off(mw * 32 + rl, cchunk & 7);
                    if ((vmask >> rl) & 1u) __pipeline_memcpy_async(dst, a.x1 + (int64_t)srow * 128 + cchunk * 8, 16);
                    else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            cp_async_arrive(&bars[B_XF + st]);   // fires when this thread's copies have landed
            mbar_arrive(&bars[B_XF + st]);       // releases the meta block and the zero rows (plain stores)
            fetch2(e_cur + 128, nxt);
            if (cnt != 128 && e_cur + cnt < e_end) {  // the tile was cut at 32 rows: the speculative fetch missed
                fetch1(e_cur + cnt, nxt);
                fetch2(e_cur + cnt, nxt);
            }
            e_cur += cnt;
            cur = nxt;
        }
    } else if (warp < 16) {
        // ===================== epilogue 1: h1 = ssp(D1 + b1) -> A2; group g owns columns [64 g, 64 g + 64) = K slab g =====================
        const int e = (warp & 3) * 32 + lane, g = (warp - 8) >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const bool add_b1 = a.G >= 64;   // otherwise b1 came through the spare basis column
        const float c2 = a.gcoeff * 1.4426950408889634f;
        // Gaussian basis of tile tq -> A1 (this group's 32 of the 64 columns).  Called when G1 of tile tq-1 is known to be
        // complete (its D1F has been waited for), so the single A1 buffer is free; G1 of tile tq then runs on the tensor core
        // while these threads apply the softplus to tile tq-1.
        auto basis = [&](uint32_t tq) {
            const uint32_t sq = tq % kS2MetaStages;
            TC2_MARK(3, tq);
            mbar_wait(&bars[B_MF + sq], (tq / kS2MetaStages) & 1u);
            const Meta& M = meta[sq];
            if (M.cnt != 0) {
                const float d = M.d[e];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int ch = 4 * g + c4;
                    float v[8];
                    const float4 g0 = *reinterpret_cast<const float4*>(goff + ch * 8), g1 = *reinterpret_cast<const float4*>(goff + ch * 8 + 4);
                    const float go[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};   // two 16-byte broadcast loads, not eight predicated ones
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float u = d - go[q];
                        v[q] = (ch * 8 + q == a.G) ? 1.0f : ex2a(c2 * u * u);   // column G = 1: picks up b1 from the W1 image
                    }
                    *reinterpret_cast<uint4*>(sm + o2A1 + sw128_chunk_off(e, ch)) =
                        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                }
                fence_proxy_async();
            }
            mbar_arrive(&bars[B_A1F]);   // also for the end-of-stream marker: G1 passes it on
        };
        basis(0);
        for (uint32_t tc = 0;; ++tc) {
            const uint32_t p = tc & 1u, par = (tc >> 1) & 1u;
            TC2_MARK(4, tc);
            mbar_wait(&bars[B_D1F + p], par);
            if (*end_g1 == (int)tc) {  // end of stream: pass it on to the G2 warp
                if (t == 256) *end_e1 = (int)tc;
                // same back-pressure as a real tile: G2 must have consumed the previous phase of A2F before this arrival
                // completes the next one -- a waiter that falls two phases behind sees the parity it is waiting for as
                // "not yet complete" and never wakes (seen as a rare hang on one-tile chunks)
                mbar_wait(&bars[B_A2E + g], (tc & 1u) ^ 1u);
                mbar_arrive(&bars[B_A2F]);
                break;
            }
            basis(tc + 1);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float v[32];
                tmem_ld32(tmD1 + p * 128 + lane_base + 64 * g + 32 * j, v);
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = ssp2(add_b1 ? v[q] + b1s[64 * g + 32 * j + q] : v[q]);
                TC2_MARK(5, tc);
                if (j == 0) mbar_wait(&bars[B_A2E + g], (tc & 1u) ^ 1u);  // G2 of the previous tile has read this K slab
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(sm + o2A2 + g * 16384 + sw128_chunk_off(e, j * 4 + q)) =
                        make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                                   pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
            }
            tc_fence_before();
            mbar_arrive(&bars[B_D1E + p]);
            fence_proxy_async();
            mbar_arrive(&bars[B_A2F]);
        }
    } else if (warp < 24) {
        // ===================== epilogue 2: msg (group g: columns [64 g, 64 g + 64)); group 0 also writes the one-hot tile,
        // group 1 reads out the previous tile (lane = feature column) =====================
        const int e = (warp & 3) * 32 + lane, g = (warp - 16) >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        // Read-out of D3 (lane = feature column e, accumulator column = segment).  A row can continue across tiles, so the
        // LAST segment of a tile is not stored but carried in a register (value, row, head flag) into the next read-out,
        // where it is either completed by that tile's first segment (same row) or stored; every other segment is a
        // finished row and goes out with one plain store.  Nothing is read back from global memory (the earlier
        // read-modify-write cost four dependent L2 round trips per tile on the critical path, profiles/r02_summary.md);
        // every agg element is still written by exactly one thread, in tile order: deterministic, no atomics.
        float carry = 0.f;
        int carry_row = -1;
        bool carry_head = false;
        auto store_row = [&](int row, bool head, float val) {
            float* dst = head ? a.head + (int64_t)blockIdx.x * 128 + e : a.agg + (int64_t)row * 128 + e;
            *dst = val;
        };
        auto readout = [&](uint32_t tq) {
            const uint32_t st = tq % kS2Stages, ms = tq % kS2MetaStages;
            const Meta& M = meta[ms];
            float v[32];
            tmem_ld32(tmD3 + st * 32 + lane_base, v);
            tc_fence_before();
            const int nseg = M.nseg;
            if (nseg > 0) {
                const int row0 = M.seg_row[0];
                float first = v[0];
                bool first_head = M.head0 != 0;
                if (carry_row == row0) { first += carry; first_head = carry_head; }
                else if (carry_row >= 0) store_row(carry_row, carry_head, carry);
                if (nseg == 1) {
                    carry = first; carry_row = row0; carry_head = first_head;
                } else {
                    store_row(row0, first_head, first);
                    float last = 0.f;
#pragma unroll
                    for (int s = 1; s < kS2MaxSeg; ++s) {
                        if (s < nseg - 1) a.agg[(int64_t)M.seg_row[s] * 128 + e] = v[s];
                        if (s == nseg - 1) last = v[s];
                    }
                    carry = last; carry_row = M.seg_row[nseg - 1]; carry_head = false;
                }
            }
            mbar_arrive(&bars[B_D3E + st]);
            mbar_arrive(&bars[B_DONE + ms]);
        };
        uint32_t tc = 0;
        for (;; ++tc) {
            const uint32_t st = tc % kS2Stages, par = (tc / kS2Stages) & 1u;
            TC2_MARK(6, tc);
            mbar_wait(&bars[B_XF + st], par);    // gathered rows (cp.async) + meta block (plain stores)
            const Meta& M = meta[tc % kS2MetaStages];
            const int cnt = M.cnt;
            if (cnt == 0) {  // end of stream: read out the last tile, wake the G3 warp
                if (tc > 0) {
                    TC2_MARK(7, tc);
                    mbar_wait(&bars[B_D3F + ((tc - 1) % kS2Stages)], ((tc - 1) / kS2Stages) & 1u);
                    tc_fence_after();
                    if (g == 1) readout(tc - 1);
                }
                if (g == 1 && carry_row >= 0) store_row(carry_row, carry_head, carry);   // the row left open by the last tile
                mbar_arrive(&bars[B_MSGF + st]);
                break;
            }
            TC2_MARK(8, tc);
            mbar_wait(&bars[B_D2F], tc & 1u);
            tc_fence_after();
            const float C = M.C[e];
            uint8_t* xs = sm + o2X + st * 32768;
            // optional: keep the filter value of this edge (bf16, 128 B per thread) for the backward pass, which then gets
            // dL/dx1 from a plain gather-multiply-reduce instead of running this whole kernel again on the transposed CSR
            uint4* wrow = (a.wout != nullptr && e < cnt) ? reinterpret_cast<uint4*>(a.wout + (int64_t)M.eid[e] * 128 + 64 * g) : nullptr;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float v[32];
                uint32_t fo[8];
                tmem_ld32(tmD2 + lane_base + 64 * g + 32 * j, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4* px = reinterpret_cast<uint4*>(xs + g * 16384 + sw128_chunk_off(e, j * 4 + q));
                    const uint4 xr = *px;
                    const uint32_t w[4] = {xr.x, xr.y, xr.z, xr.w};
                    uint32_t o[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float f0 = v[8 * q + 2 * u] * C, f1 = v[8 * q + 2 * u + 1] * C;   // b2 came through the GEMM
                        fo[4 * (q & 1) + u] = pack_bf16(f0, f1);
                        o[u] = pack_bf16(f0 * __uint_as_float(w[u] << 16), f1 * __uint_as_float(w[u] & 0xffff0000u));   // C = 0 and zero rows for the padding slots
                    }
                    *px = make_uint4(o[0], o[1], o[2], o[3]);
                    // 32 B per lane and instruction (a whole sector): every lane of the warp writes a different row, so the
                    // cost of these stores is the number of instructions, not the bytes
                    if ((q & 1) && wrow)
                        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(wrow + j * 4 + q - 1), "r"(fo[0]), "r"(fo[1]),
                                     "r"(fo[2]), "r"(fo[3]), "r"(fo[4]), "r"(fo[5]), "r"(fo[6]), "r"(fo[7]) : "memory");
                }
            }
            tc_fence_before();
            mbar_arrive(&bars[B_D2E]);   // G2 of the next tile may start: nothing below is on its path
            // G3 of the previous tile is done (it has had a whole message phase to finish): the single one-hot tile can be
            // rewritten and the previous accumulator read out
            if (tc > 0) {
                TC2_MARK(9, tc);
                mbar_wait(&bars[B_D3F + ((tc - 1) % kS2Stages)], ((tc - 1) / kS2Stages) & 1u);
                tc_fence_after();
            }
            if (g == 0) {
                const int seg = M.seg[e];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t o[4] = {0u, 0u, 0u, 0u};
                    if (seg >= 0 && (seg >> 3) == q) o[(seg & 7) >> 1] = (seg & 1) ? 0x3f800000u : 0x00003f80u;
                    *reinterpret_cast<uint4*>(sm + o2S + sw128_chunk_off(e, q)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            fence_proxy_async();
            mbar_arrive(&bars[B_MSGF + st]);
            if (g == 1 && tc > 0) readout(tc - 1);   // global read-modify-write latency: off the G2 / G3 critical path
        }
    } else if (warp == 24) {
        // ===================== G1 = rbf W1^T (whole warp, one elected lane issues) =====================
        const uint32_t id1 = umma_idesc_bf16(128, 128);
        const uint32_t w1b = smem_u32(sm + o2W1), a1b = smem_u32(sm + o2A1);
        for (uint32_t tc = 0;; ++tc) {
            const uint32_t p = tc & 1u, par = (tc >> 1) & 1u, st = tc % kS2Stages;
            TC2_MARK(10, tc);
            mbar_wait(&bars[B_A1F], tc & 1u);
            if (meta[tc % kS2MetaStages].cnt == 0) {  // end of stream: tell epilogue 1
                mbar_wait(&bars[B_D1E + p], par ^ 1u);   // epilogue 1 has consumed the previous phase of D1F[p] (see the note there)
                if (elect_one()) {
                    *end_g1 = (int)tc;
                    __threadfence_block();
                    mbar_arrive(&bars[B_D1F + p]);
                }
                __syncwarp();
                break;
            }
            TC2_MARK(11, tc);
            mbar_wait(&bars[B_D1E + p], par ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                umma_tile(tmD1 + p * 128, a1b, 16384, w1b, 16384, 64, id1);
                umma_commit(&bars[B_D1F + p]);
            }
            __syncwarp();
        }
    } else if (warp == 25) {
        // ===================== G2 = h1 W2^T =====================
        const uint32_t id1 = umma_idesc_bf16(128, 128);
        const uint32_t w2b = smem_u32(sm + o2W2), a2b = smem_u32(sm + o2A2), sbias = smem_u32(sm + o2S);
        for (uint32_t tc = 0;; ++tc) {
            TC2_MARK(12, tc);
            mbar_wait(&bars[B_A2F], tc & 1u);
            if (*end_e1 == (int)tc) break;
            TC2_MARK(13, tc);
            mbar_wait(&bars[B_D2E], (tc & 1u) ^ 1u);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16)
                        umma_bf16(tmD2, umma_desc_k128(a2b + ks * 16384 + k16 * 32), umma_desc_k128(w2b + ks * 16384 + k16 * 32), id1,
                                  (ks | k16) ? 1u : 0u);
                    umma_commit(&bars[B_A2E + ks]);
                }
                umma_bf16(tmD2, umma_desc_k128(sbias + 64), umma_desc_k128(sbias + 96), id1, 1u);   // + 1 * b2^T
                umma_commit(&bars[B_D2F]);
            }
            __syncwarp();
        }
    } else {
        // ===================== G3 = msg^T S (the segmented row sum) =====================
        const uint32_t id3 = umma_idesc_bf16(128, 32, true, true);
        const uint32_t sb = smem_u32(sm + o2S);
        for (uint32_t tc = 0;; ++tc) {
            const uint32_t st = tc % kS2Stages, par = (tc / kS2Stages) & 1u;
            TC2_MARK(14, tc);
            mbar_wait(&bars[B_MSGF + st], par);
            if (meta[tc % kS2MetaStages].cnt == 0) break;
            TC2_MARK(15, tc);
            mbar_wait(&bars[B_D3E + st], par ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t mb = smem_u32(sm + o2X + st * 32768);
#pragma unroll
                for (int k16 = 0; k16 < 8; ++k16)
                    umma_bf16(tmD3 + st * 32, umma_desc_mn128(mb + k16 * 2048, 16384), umma_desc_mn128(sb + k16 * 2048, 16384), id3, k16 ? 1u : 0u);
                umma_commit(&bars[B_D3F + st]);
            }
            __syncwarp();
        }
    }
    TC2_MARK(255, 0);
    tc_fence_before();
    __syncthreads();
    TC2_MARK(256, 0);
    if (warp == 24) tmem_dealloc<512>(tm);
}

// rows that straddle a CTA boundary: add the later CTAs' head partials in CTA order.  One block per chunk; the block of
// the FIRST chunk whose head belongs to a row adds every consecutive head of that row, so rows are handled in parallel
// and the order of the additions into one row is fixed.
__device__ __forceinline__ int tc2_head_row(const int32_t* __restrict__ rowptr, int64_t n, int64_t E, int nchunks, int ch) {
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

__global__ void __launch_bounds__(128) schnet_tc2_fixup_kernel(const int32_t* __restrict__ rowptr, int64_t n, int64_t E, int nchunks,
                                                               const float* __restrict__ head, float* __restrict__ agg) {
    const int ch = blockIdx.x + 1;
    const int row = tc2_head_row(rowptr, n, E, nchunks, ch);
    if (row < 0 || tc2_head_row(rowptr, n, E, nchunks, ch - 1) == row) return;  // no head, or an earlier block owns this row
    const int c = threadIdx.x;
    float v = agg[(int64_t)row * 128 + c];
    for (int c2 = ch; c2 < nchunks && (c2 == ch || tc2_head_row(rowptr, n, E, nchunks, c2) == row); ++c2) v += head[(int64_t)c2 * 128 + c];
    agg[(int64_t)row * 128 + c] = v;
}

// ------------------------------------------------------------------------------------------------
// filter-side backward, pipelined (weight gradients of the filter MLP; GMP_BF16_TC, F = 128, G <= 63, lazy basis).
// Same mathematics as schnet_bwd_tc_kernel:
//   G1  D1 = R W1^T (b1 rides in column 63 of the basis tile)       H = ssp(D1)                       (epiH)
//   P   = x1[src] * g[dst] * C  (over the gathered bf16 rows)                                          (epiP)
//   G3  D3 = P W2;  dW2 += P^T H;  db2 += P^T 1                                                        (accumulators in TMEM)
//   Q   = D3 * sigmoid(pre1), sigmoid from h1, written over P                                          (epiQ)
//   G4  [dW1 | db1] += Q^T R
// Warps: meta 0-3 | epiP 4-7 | epiH 8-15 | epiQ 16-23 | MMA issue 24 (G1), 25 (G3), 26 (G4).  Stages R, rows/P/Q, H two deep;
// D1, D3 single.  TMEM: D1 @0 | D3 @128 | dW2 @256 | dW1|db1 @384 (64) | db2 @448 (16).
// Tried in round 2 and measured slower (profiles/r02_summary.md): g read as bf16 rows in two batches of eight loads (0.372 ->
// 0.382 ms), epiP on eight warps with epiH on four (0.407 ms).  The kernel is bound by the latency of the per-stage chain
// gather -> P -> G3 -> Q -> G4 with two stages in flight (212 KB of shared memory), not by any one role.
// ------------------------------------------------------------------------------------------------
constexpr int kB2Threads = 864;   // 27 warps
constexpr int o3W1 = 0;                    // [128 f][64 g], column 63 = b1          16 KB
constexpr int o3W2T = 16384;               // 2 slabs [128 f][64 f'] (W2 transposed) 32 KB
constexpr int o3R = o3W2T + 32768;         // 2 x 16 KB basis tiles (ones column at g = 63)
constexpr int o3X = o3R + 2 * 16384;       // 2 x 32 KB gathered rows -> P -> Q
constexpr int o3H = o3X + 2 * 32768;       // 2 x 32 KB h1
constexpr int o3Ones = o3H + 2 * 32768;    // 4 KB of bf16 ones
constexpr int o3Vec = o3Ones + 4096;       // goff[64]
constexpr int o3Meta = o3Vec + 256;        // 2 x { C[128] f32, grow[128] i32 }
constexpr int o3Bar = o3Meta + 2 * 1024;
enum { C_RF = 0, C_XF = 2, C_PF = 4, C_HF = 6, C_QF = 8, C_DONE = 10, C_D1F = 12, C_D1E = 13, C_D3F = 14, C_D3E = 15, C_COUNT = 16 };
constexpr int kTc2BwdSmem = o3Bar + C_COUNT * 8 + 16 + 1024;

struct Tc2BwdArgs {
    Tc2Args f;
    const float* g_agg;   // [n,128] dL/dagg, read at the destination (CSR row) node
    float* parts;         // [gridDim.x][128*64 + 128 + 128*128 + 128]
};

__global__ void __launch_bounds__(896, 1) schnet_bwd_tc2_kernel(Tc2BwdArgs b) {  // 896: caps the registers at 72
    const Tc2Args& a = b.f;
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    float* goff = reinterpret_cast<float*>(sm + o3Vec);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + o3Bar);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm + o3Bar + C_COUNT * 8);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    auto metaC = [&](uint32_t p) { return reinterpret_cast<float*>(sm + o3Meta + p * 1024); };
    auto metaRow = [&](uint32_t p) { return reinterpret_cast<int*>(sm + o3Meta + p * 1024 + 512); };

    for (int x = t; x < 128 * 8; x += kB2Threads) {  // W1 [f][g] K-major, g padded to 64; column 63 = b1
        const int f = x >> 3, ch = x & 7;
        uint32_t p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g0 = ch * 8 + 2 * j;
            const float lo = g0 < a.G ? __ldg(a.w1 + f * a.G + g0) : 0.f;
            const float hi = g0 + 1 < a.G ? __ldg(a.w1 + f * a.G + g0 + 1) : (g0 + 1 == 63 ? __ldg(a.b1 + f) : 0.f);
            p[j] = pack_bf16(lo, hi);
        }
        *reinterpret_cast<uint4*>(sm + o3W1 + sw128_chunk_off(f, ch)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
    for (int x = t; x < 128 * 16; x += kB2Threads) {  // W2T [f][f'] = W2[f'][f], K = f' in two slabs
        const int f = x >> 4, ch16 = x & 15, kb = ch16 >> 3, ch = ch16 & 7;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __ldg(a.w2 + (int64_t)(kb * 64 + ch * 8 + q) * 128 + f);
        *reinterpret_cast<uint4*>(sm + o3W2T + kb * 16384 + sw128_chunk_off(f, ch)) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
    for (int i = t; i < 2048; i += kB2Threads) reinterpret_cast<uint16_t*>(sm + o3Ones)[i] = 0x3f80;
    if (t < 64) goff[t] = t < a.G ? __ldg(a.goff + t) : 1.0e18f;
    if (t == 0) {
        for (int i = 0; i < C_COUNT; ++i) mbar_init(&bars[i], 1);
        mbar_init(&bars[C_RF], 128); mbar_init(&bars[C_RF + 1], 128);
        mbar_init(&bars[C_XF], 256); mbar_init(&bars[C_XF + 1], 256);
        mbar_init(&bars[C_PF], 128); mbar_init(&bars[C_PF + 1], 128);
        mbar_init(&bars[C_HF], 256); mbar_init(&bars[C_HF + 1], 256);
        mbar_init(&bars[C_QF], 256); mbar_init(&bars[C_QF + 1], 256);
        mbar_init(&bars[C_D1E], 256);
        mbar_init(&bars[C_D3E], 256);
        fence_mbar_init();
    }
    if (warp == 24) tmem_alloc<512>(tmem_ptr);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr;
    const uint32_t tmD1 = tm, tmD3 = tm + 128, tmW2 = tm + 256, tmW1 = tm + 384, tmB2 = tm + 448;

    const int64_t e_begin = (a.E * blockIdx.x) / gridDim.x, e_end = (a.E * (blockIdx.x + 1)) / gridDim.x;
    const uint32_t ntile = (uint32_t)((e_end - e_begin + 127) / 128);

    if (warp < 4) {
        // ===================== meta =====================
        const int e = t;
        const float cw = 3.14159265358979323846f / a.cutoff;
        const float c2 = a.gcoeff * 1.4426950408889634f;
        // scalars are fetched one tile ahead (two dependent L2 round trips: position -> edge id -> length)
        struct Pre { int eid, src, grow; float d; };
        auto fetch1 = [&](uint32_t tq, Pre& P) {
            const int64_t k = e_begin + (int64_t)tq * 128 + e;
            P.eid = 0; P.src = 0; P.grow = 0;
            if (k < e_end) {
                P.eid = a.perm ? __ldg(a.perm + k) : (int)k;
                P.src = __ldg(a.col + k);
                P.grow = __ldg(a.rowid + k);
            }
        };
        auto fetch2 = [&](uint32_t tq, Pre& P) {
            P.d = (e_begin + (int64_t)tq * 128 + e < e_end) ? __ldg(a.ew + P.eid) : 1.0e18f;
        };
        Pre cur, nxt;
        fetch1(0, cur);
        fetch2(0, cur);
        for (uint32_t tc = 0; tc < ntile; ++tc) {
            const uint32_t p = tc & 1u, par = (tc >> 1) & 1u;
            const int64_t k = e_begin + (int64_t)tc * 128 + e;
            const bool valid = k < e_end;
            fetch1(tc + 1, nxt);
            const int src = cur.src, grow = cur.grow;
            const float d = cur.d;
            const float C = valid ? 0.5f * (__cosf(d * cw) + 1.0f) : 0.f;
            mbar_wait(&bars[C_DONE + p], par ^ 1u);   // G4 of tile tc-2 done: every buffer of stage p is free
            metaC(p)[e] = C;
            metaRow(p)[e] = grow;
            {   // half a warp per gathered row (see the forward kernel): 4 cache lines per copy instruction instead of 32
                uint8_t* xs = sm + o3X + p * 32768;
                const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                const int cchunk = lane & 15, sub = lane >> 4;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int rl = 2 * i + sub;
                    const int srow = __shfl_sync(0xffffffffu, src, rl);
                    uint8_t* dst = xs + (cchunk >> 3) * 16384 + sw128_chunk_off(warp * 32 + rl, cchunk & 7);
                    if ((vmask >> rl) & 1u) __pipeline_memcpy_async(dst, a.x1 + (int64_t)srow * 128 + cchunk * 8, 16);
                    else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            cp_async_arrive(&bars[C_XF + p]);
            mbar_arrive(&bars[C_XF + p]);
            fetch2(tc + 1, nxt);   // in flight during the basis phase
            uint8_t* rs = sm + o3R + p * 16384;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                float v[8];
                const float4 g0 = *reinterpret_cast<const float4*>(goff + ch * 8), g1 = *reinterpret_cast<const float4*>(goff + ch * 8 + 4);
                const float go[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float u = d - go[q];
                    v[q] = ex2a(c2 * u * u);
                }
                if (ch == 7) v[7] = 1.0f;   // ones column: bias of GEMM 1, and the column sums db1 in the dW1 accumulator
                *reinterpret_cast<uint4*>(rs + sw128_chunk_off(e, ch)) =
                    make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            }
            fence_proxy_async();
            mbar_arrive(&bars[C_RF + p]);
            cur = nxt;
        }
    } else if (warp < 8) {
        // ===================== epiP: P = x1[src] * g[dst] * C over the gathered rows =====================
        const int e = t - 128;
        for (uint32_t tc = 0; tc < ntile; ++tc) {
            const uint32_t p = tc & 1u, par = (tc >> 1) & 1u;
            mbar_wait(&bars[C_XF + p], par);
            const float C = metaC(p)[e];
            const float4* gr = reinterpret_cast<const float4*>(b.g_agg + (int64_t)metaRow(p)[e] * 128);
            uint8_t* xs = sm + o3X + p * 32768;
#pragma unroll
            for (int ch = 0; ch < 16; ++ch) {
                uint4* px = reinterpret_cast<uint4*>(xs + (ch >> 3) * 16384 + sw128_chunk_off(e, ch & 7));
                const uint4 xr = *px;
                const float4 g0 = __ldg(gr + 2 * ch), g1 = __ldg(gr + 2 * ch + 1);   // C = 0 for the padding slots
                const uint32_t w[4] = {xr.x, xr.y, xr.z, xr.w};
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                uint32_t o[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    o[u] = pack_bf16(__uint_as_float(w[u] << 16) * gg[2 * u] * C, __uint_as_float(w[u] & 0xffff0000u) * gg[2 * u + 1] * C);
                *px = make_uint4(o[0], o[1], o[2], o[3]);
            }
            fence_proxy_async();
            mbar_arrive(&bars[C_PF + p]);
        }
    } else if (warp < 16) {
        // ===================== epiH: h1 = ssp(D1) -> H; group g owns columns [64 g, 64 g + 64) =====================
        const int e = (warp & 3) * 32 + lane, g = (warp - 8) >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (uint32_t tc = 0; tc < ntile; ++tc) {
            const uint32_t p = tc & 1u;
            mbar_wait(&bars[C_D1F], tc & 1u);
            tc_fence_after();
            uint8_t* hs = sm + o3H + p * 32768 + g * 16384;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float v[32];
                tmem_ld32(tmD1 + lane_base + 64 * g + 32 * j, v);
                if (j == 1) {   // both loads of this thread are done: G1 of the next tile may overwrite D1
                    tc_fence_before();
                    mbar_arrive(&bars[C_D1E]);
                }
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = ssp2(v[q]);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(hs + sw128_chunk_off(e, j * 4 + q)) =
                        make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                                   pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
            }
            fence_proxy_async();
            mbar_arrive(&bars[C_HF + p]);
        }
    } else if (warp < 24) {
        // ===================== epiQ: Q = D3 * sigmoid(pre1) over P; at the end, the accumulators =====================
        const int e = (warp & 3) * 32 + lane, g = (warp - 16) >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (uint32_t tc = 0; tc < ntile; ++tc) {
            const uint32_t p = tc & 1u;
            mbar_wait(&bars[C_D3F], tc & 1u);   // G3 done: D3 ready, P and H no longer read by the tensor core
            tc_fence_after();
            uint8_t* hs = sm + o3H + p * 32768 + g * 16384;
            uint8_t* qs = sm + o3X + p * 32768 + g * 16384;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float v[32];
                tmem_ld32(tmD3 + lane_base + 64 * g + 32 * j, v);
                if (j == 1) {
                    tc_fence_before();
                    mbar_arrive(&bars[C_D3E]);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t off = sw128_chunk_off(e, j * 4 + q);
                    const uint4 hp = *reinterpret_cast<const uint4*>(hs + off);
                    const uint32_t hw[4] = {hp.x, hp.y, hp.z, hp.w};
                    uint32_t o[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        // sigmoid(pre1) = 1 - exp(-softplus(pre1)) = 1 - 0.5 * 2^(-h1 log2 e)   (h1 = softplus - ln 2)
                        const float s0 = fmaf(-0.5f, ex2a(-1.4426950408889634f * __uint_as_float(hw[u] << 16)), 1.0f);
                        const float s1 = fmaf(-0.5f, ex2a(-1.4426950408889634f * __uint_as_float(hw[u] & 0xffff0000u)), 1.0f);
                        o[u] = pack_bf16(v[8 * q + 2 * u] * s0, v[8 * q + 2 * u + 1] * s1);
                    }
                    *reinterpret_cast<uint4*>(qs + off) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            fence_proxy_async();
            mbar_arrive(&bars[C_QF + p]);
        }
        // ---- this CTA's partial gradients: [dW1 128 x 64 | db1 128 | dW2 128 x 128 | db2 128]
        float* my = b.parts + (int64_t)blockIdx.x * (128 * 64 + 128 + 128 * 128 + 128);
        if (ntile > 0) {
            const uint32_t last = ntile - 1;
            mbar_wait(&bars[C_DONE + (last & 1u)], (last >> 1) & 1u);   // the last G4; everything earlier completed before it
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {  // dW2[f' = e][f]
                float v[32];
                tmem_ld32(tmW2 + lane_base + 64 * g + 32 * j, v);
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                    *reinterpret_cast<float4*>(my + 128 * 64 + 128 + e * 128 + 64 * g + 32 * j + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
            }
            {   // dW1[f = e][g] (32 columns per group); column 63 = db1
                float v[32];
                tmem_ld32(tmW1 + lane_base + 32 * g, v);
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                    *reinterpret_cast<float4*>(my + e * 64 + 32 * g + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
                if (g == 1) my[128 * 64 + e] = v[31];
            }
            if (g == 0) {  // db2[f' = e] = column 0 of the P^T 1 accumulator
                float v[32];
                tmem_ld32(tmB2 + lane_base, v);
                my[128 * 64 + 128 + 128 * 128 + e] = v[0];
            }
            tc_fence_before();
        } else {
            for (int x = t - 512; x < 128 * 64 + 128 + 128 * 128 + 128; x += 256) my[x] = 0.f;
        }
    } else if (warp == 24) {
        // ===================== G1: D1 = R W1^T =====================
        const uint32_t id = umma_idesc_bf16(128, 128);
        const uint32_t w1b = smem_u32(sm + o3W1);
        for (uint32_t tc = 0; tc < ntile; ++tc) {
            const uint32_t p = tc & 1u, par = (tc >> 1) & 1u;
            mbar_wait(&bars[C_RF + p], par);
            mbar_wait(&bars[C_D1E], (tc & 1u) ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                umma_tile(tmD1, smem_u32(sm + o3R + p * 16384), 16384, w1b, 16384, 64, id);
                umma_commit(&bars[C_D1F]);
            }
            __syncwarp();
        }
    } else if (warp == 25) {
        // ===================== G3: D3 = P W2; dW2 += P^T H; db2 += P^T 1 =====================
        const uint32_t id_kk = umma_idesc_bf16(128, 128), id_mn = umma_idesc_bf16(128, 128, true, true), id_16 = umma_idesc_bf16(128, 16, true, false);
        const uint32_t w2t = smem_u32(sm + o3W2T);
        const uint64_t ones = umma_desc_k128(smem_u32(sm + o3Ones));
        for (uint32_t tc = 0; tc < ntile; ++tc) {
            const uint32_t p = tc & 1u, par = (tc >> 1) & 1u;
            mbar_wait(&bars[C_PF + p], par);
            mbar_wait(&bars[C_HF + p], par);
            mbar_wait(&bars[C_D3E], (tc & 1u) ^ 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t pb = smem_u32(sm + o3X + p * 32768), hb = smem_u32(sm + o3H + p * 32768);
                umma_tile(tmD3, pb, 16384, w2t, 16384, 128, id_kk);
                umma_tile_mn(tmW2, pb, 16384, hb, 16384, 128, id_mn, tc > 0);
#pragma unroll
                for (int k16 = 0; k16 < 8; ++k16)
                    umma_bf16(tmB2, umma_desc_mn128(pb + k16 * 2048, 16384), ones, id_16, (k16 || tc > 0) ? 1u : 0u);
                umma_commit(&bars[C_D3F]);
            }
            __syncwarp();
        }
    } else {
        // ===================== G4: [dW1 | db1] += Q^T R =====================
        const uint32_t id_64 = umma_idesc_bf16(128, 64, true, true);
        for (uint32_t tc = 0; tc < ntile; ++tc) {
            const uint32_t p = tc & 1u, par = (tc >> 1) & 1u;
            mbar_wait(&bars[C_QF + p], par);
            tc_fence_after();
            if (elect_one()) {
                umma_tile_mn(tmW1, smem_u32(sm + o3X + p * 32768), 16384, smem_u32(sm + o3R + p * 16384), 16384, 128, id_64, tc > 0);
                umma_commit(&bars[C_DONE + p]);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 24) tmem_dealloc<512>(tm);
}

}  // namespace gmp

using namespace gmp;

extern "C" {

#ifdef GMP_TC2_PROGRESS
void gmp_debug_tc2_progress(int* host_mapped) { g_tc2_dbg = host_mapped; }
#endif

int32_t gmp_schnet_tc2_num_chunks(int64_t num_edges) {
    const int64_t nt = ceil_div(num_edges, 128);
    return (int32_t)(nt < num_sms() ? (nt < 1 ? 1 : nt) : num_sms());
}

int gmp_schnet_cfconv_fwd_tc2_keep(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid, int64_t n,
                                   int64_t num_edges, const float* edge_weight, const void* x1_bf16, const gmp_schnet_filter* f,
                                   float* agg, float* head, void* filter_out_bf16, const int32_t* filter_row, gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && f && agg && head, "schnet_cfconv_fwd_tc2: NULL pointer");
    GMP_REQUIRE(num_edges == 0 || (col && rowid && edge_weight && x1_bf16), "schnet_cfconv_fwd_tc2: NULL edge/feature pointer");
    GMP_REQUIRE(f->num_filters == 128 && f->num_gaussians >= 1 && f->num_gaussians <= 64 && f->gauss_offset,
                "schnet_cfconv_fwd_tc2: built for 128 filters, <= 64 lazily expanded Gaussians");
    GMP_REQUIRE(n >= 0 && n < (1ll << 31) && num_edges >= 0 && num_edges < (1ll << 31), "schnet_cfconv_fwd_tc2: sizes out of range");
    GMP_CUDA(cudaMemsetAsync(agg, 0, (size_t)n * 128 * sizeof(float), stream));
    if (n == 0 || num_edges == 0) return GMP_OK;
    const int nchunks = gmp_schnet_tc2_num_chunks(num_edges);
    GMP_CUDA(cudaMemsetAsync(head, 0, (size_t)nchunks * 128 * sizeof(float), stream));
    Tc2Args a;
    a.rowptr = rowptr; a.col = col; a.perm = perm; a.rowid = rowid; a.n = n; a.E = num_edges; a.ew = edge_weight;
    a.x1 = (const __nv_bfloat16*)x1_bf16; a.w1 = f->w1; a.b1 = f->b1; a.w2 = f->w2; a.b2 = f->b2; a.goff = f->gauss_offset;
    a.G = f->num_gaussians; a.cutoff = f->cutoff; a.gcoeff = f->gauss_coeff; a.agg = agg; a.head = head; a.wout = (__nv_bfloat16*)filter_out_bf16; a.wrow = filter_row;
#ifdef GMP_TC2_PROGRESS
    a.dbg = g_tc2_dbg;
#else
    a.dbg = nullptr;
#endif
    GMP_CUDA(cudaFuncSetAttribute(schnet_fwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTc2Smem));
    schnet_fwd_tc2_kernel<<<nchunks, kS2Threads, kTc2Smem, stream>>>(a);
    int rc = check_launch("schnet_fwd_tc2_kernel");
    if (rc != GMP_OK) return rc;
    if (nchunks > 1) {
        schnet_tc2_fixup_kernel<<<nchunks - 1, 128, 0, stream>>>(rowptr, n, num_edges, nchunks, head, agg);
        rc = check_launch("schnet_tc2_fixup_kernel");
    }
    return rc;
}

int gmp_schnet_cfconv_fwd_tc2(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid, int64_t n,
                              int64_t num_edges, const float* edge_weight, const void* x1_bf16, const gmp_schnet_filter* f, float* agg,
                              float* head, gmp_stream_t stream) {
    return gmp_schnet_cfconv_fwd_tc2_keep(rowptr, col, perm, rowid, n, num_edges, edge_weight, x1_bf16, f, agg, head, nullptr, nullptr, stream);
}

int gmp_schnet_cfconv_bwd_tc2(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const int32_t* rowid, int64_t n,
                              int64_t num_edges, const float* edge_weight, const void* x1_bf16, const gmp_schnet_filter* f,
                              const float* g_agg, float* wgrad_parts, int32_t nparts, gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && f && g_agg && wgrad_parts && col && rowid && edge_weight && x1_bf16, "schnet_cfconv_bwd_tc2: NULL pointer");
    GMP_REQUIRE(f->num_filters == 128 && f->num_gaussians >= 1 && f->num_gaussians <= 63 && f->gauss_offset,
                "schnet_cfconv_bwd_tc2: built for 128 filters, <= 63 lazily expanded Gaussians");
    GMP_REQUIRE(num_edges > 0 && num_edges < (1ll << 31) && n < (1ll << 31), "schnet_cfconv_bwd_tc2: sizes out of range");
    GMP_REQUIRE(nparts >= 1 && nparts <= 1024, "schnet_cfconv_bwd_tc2: nparts in [1, 1024]");
    Tc2BwdArgs b;
    Tc2Args& a = b.f;
    a.rowptr = rowptr; a.col = col; a.perm = perm; a.rowid = rowid; a.n = n; a.E = num_edges; a.ew = edge_weight;
    a.x1 = (const __nv_bfloat16*)x1_bf16; a.w1 = f->w1; a.b1 = f->b1; a.w2 = f->w2; a.b2 = f->b2; a.goff = f->gauss_offset;
    a.G = f->num_gaussians; a.cutoff = f->cutoff; a.gcoeff = f->gauss_coeff; a.agg = nullptr; a.head = nullptr; a.wout = nullptr; a.wrow = nullptr;
    a.dbg = nullptr;
    b.g_agg = g_agg; b.parts = wgrad_parts;
    GMP_CUDA(cudaFuncSetAttribute(schnet_bwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTc2BwdSmem));
    schnet_bwd_tc2_kernel<<<nparts, kB2Threads, kTc2BwdSmem, stream>>>(b);
    return check_launch("schnet_bwd_tc2_kernel");
}

}  // extern "C"
