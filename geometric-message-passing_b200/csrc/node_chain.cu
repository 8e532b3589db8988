// Node-side dense chains on tcgen05 (GMP_BF16_TC mode): up to three chained 128-wide nn.Linear layers with their
// elementwise neighbours fused into the epilogues, one kernel per chain, the intermediate rows never leaving the SM.
//   SchNet interaction, forward   agg -> lin2 + b -> ssp -> [Y] -> lin + b -> + h -> [h'] -> next lin1 -> [x1' (bf16)]
//                                 (PyG CFConv.lin2 / InteractionBlock.act, .lin, the residual of models/schnet.py:72 and the
//                                 next block's CFConv.lin1)
//   SchNet interaction, backward  dx1' -> lin1^T -> + G -> [G'] -> lin^T -> * ssp'(Y) -> [dT] -> lin2^T -> [dAgg (fp32 + bf16)]
//   EGNN node update              cat[h, agg] -> Linear(2d -> d) -> LayerNorm -> act -> Linear -> LayerNorm -> act (+ h)
//                                 (models/layers/egnn_layer.py:41-48, 82-86; the 2d-wide input is two K = 128 sources
//                                 accumulated into one tile)
//   o3.Linear blocks, P/Q projections, single layers: one stage.
//
// Work decomposition: a CTA runs NS independent "streams" (one warpgroup of 128 threads each, thread = tile row = TMEM
// lane) over 128-row tiles; the streams share the resident bf16 weight images and overlap each other's load / MMA /
// epilogue phases.  Per tile and stage: A image (bf16, K-major, 128-byte swizzle) x weight image -> 128 x 128 fp32
// accumulator in tensor memory -> epilogue straight out of TMEM in 32-column chunks (bias, LayerNorm over the row --
// thread-local, the thread owns its whole row --, activation, multiply by an auxiliary row, residual add), global outputs
// through a per-warp shared-memory transpose so that every global access is a full 128-byte row segment, and the bf16 copy
// of the result written in place as the next stage's A image.
// HBM-bound: 512 B read per row and source, 512 B (fp32) / 256 B (bf16) written per row and output.
#include "common.cuh"
#include "tc_common.cuh"

namespace gmp {

using namespace tc;

namespace {

constexpr int kNcSlab = 128 * 128;            // one [128 rows][64 bf16] slab
constexpr int kNcImg = 2 * kNcSlab;           // one 128 x 128 bf16 image (32 KB)
constexpr int kNcStgRow = 36;                 // floats per staged row (32 + 4: conflict-free float4 rows)
constexpr int kNcStgWarp = 32 * kNcStgRow * 4;   // bytes per warp
constexpr int kNcMaxStages = 3;
constexpr int kNcMaxImgs = 4;                 // stage 0 may take two K = 128 sources

struct NodeChainArgs {
    const float* a0;
    const float* a1;
    int64_t n;
    int32_t nstage, nsrc;
    gmp_node_stage st[kNcMaxStages];
};

// MUFU-based transcendentals (ex2 / lg2 / rcp approximations, ~1e-6 relative): this kernel only exists in the bf16-operand mode,
// whose next GEMM rounds these values to 8 mantissa bits anyway; the precise expf / log1pf forms cost 60 us per 16 M elements.
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_fast(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// softplus(x) - ln 2 = max(x, 0) + ln(1 + exp(-|x|)) - ln 2
__device__ __forceinline__ float ssp_fast(float x) {
    const float u = ex2_fast(-fabsf(x) * 1.4426950408889634f);
    return fmaf(lg2_fast(1.f + u), 0.6931471805599453f, fmaxf(x, 0.f) - 0.6931471805599453f);
}

// 128 rows x 128 fp32 columns of `src` starting at row0 -> bf16 K-major swizzled image; 128 threads (r = 0..127)
__device__ __forceinline__ void load_tile_image(const float* __restrict__ src, int64_t row0, int64_t n, uint8_t* img, int r) {
    const int ch = r & 15, rr0 = r >> 4;
    // two batches of 8 rows per thread: 16 independent 16-byte loads in flight per thread (the tile load is latency-bound)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        float4 lo[8], hi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t row = row0 + rr0 + 8 * (8 * b + i);
            lo[i] = hi[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < n) {
                lo[i] = ldg4(src + row * 128 + ch * 8);
                hi[i] = ldg4(src + row * 128 + ch * 8 + 4);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rr = rr0 + 8 * (8 * b + i);
            *reinterpret_cast<uint4*>(img + (ch >> 3) * kNcSlab + sw128_chunk_off(rr, ch & 7)) =
                make_uint4(pack_bf16(lo[i].x, lo[i].y), pack_bf16(lo[i].z, lo[i].w), pack_bf16(hi[i].x, hi[i].y), pack_bf16(hi[i].z, hi[i].w));
        }
    }
}

// one quarter (32 of the 128 rows: 4 per thread) of the next tile's image, split into issue (registers) and commit, so that
// the loads are in flight while the caller works on something else
__device__ __forceinline__ void tile_quarter_issue(const float* __restrict__ src, int64_t row0, int64_t n, int r, int q, float4 (&lo)[4], float4 (&hi)[4]) {
    const int ch = r & 15, rr0 = r >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = row0 + rr0 + 8 * (4 * q + i);
        lo[i] = hi[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < n) {
            lo[i] = ldg4(src + row * 128 + ch * 8);
            hi[i] = ldg4(src + row * 128 + ch * 8 + 4);
        }
    }
}
__device__ __forceinline__ void tile_quarter_commit(uint8_t* img, int r, int q, const float4 (&lo)[4], const float4 (&hi)[4]) {
    const int ch = r & 15, rr0 = r >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rr = rr0 + 8 * (4 * q + i);
        *reinterpret_cast<uint4*>(img + (ch >> 3) * kNcSlab + sw128_chunk_off(rr, ch & 7)) =
            make_uint4(pack_bf16(lo[i].x, lo[i].y), pack_bf16(lo[i].z, lo[i].w), pack_bf16(hi[i].x, hi[i].y), pack_bf16(hi[i].z, hi[i].w));
    }
}

// the same block read in two halves: issue the coalesced loads early (registers), hand them to the owning threads later
__device__ __forceinline__ void aux_issue(const float* __restrict__ src, int64_t rbase, int64_t n, int c0, int lane, float4 (&p)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t row = rbase + i * 4 + (lane >> 3);
        p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < n) p[i] = ldg4(src + row * 128 + c0 + (lane & 7) * 4);
    }
}
__device__ __forceinline__ void aux_commit(const float4 (&p)[8], float* stg, int lane, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(stg + (i * 4 + (lane >> 3)) * kNcStgRow + (lane & 7) * 4) = p[i];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 q = *reinterpret_cast<const float4*>(stg + lane * kNcStgRow + j * 4);
        v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    }
    __syncwarp();
}

// coalesced read of a [32 rows][32 cols] block of a [n,128] fp32 array into the calling thread's row (lane = row)
__device__ __forceinline__ void staged_load(const float* __restrict__ src, int64_t rbase, int64_t n, int c0, float* stg, int lane, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3);
        const int64_t row = rbase + rr;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < n) q = ldg4(src + row * 128 + c0 + (lane & 7) * 4);
        *reinterpret_cast<float4*>(stg + rr * kNcStgRow + (lane & 7) * 4) = q;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 q = *reinterpret_cast<const float4*>(stg + lane * kNcStgRow + j * 4);
        v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    }
    __syncwarp();
}

// the transposed way: every thread's 32 values -> full 128-byte row segments of the fp32 and / or bf16 output
__device__ __forceinline__ void staged_store(const float (&v)[32], float* __restrict__ o32, __nv_bfloat16* __restrict__ o16, int64_t rbase,
                                             int64_t n, int c0, float* stg, int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * kNcStgRow + j * 4) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3);
        const int64_t row = rbase + rr;
        const float4 q = *reinterpret_cast<const float4*>(stg + rr * kNcStgRow + (lane & 7) * 4);
        if (row < n) {
            if (o32) *reinterpret_cast<float4*>(o32 + row * 128 + c0 + (lane & 7) * 4) = q;
            if (o16) *reinterpret_cast<uint2*>(o16 + row * 128 + c0 + (lane & 7) * 4) = make_uint2(pack_bf16(q.x, q.y), pack_bf16(q.z, q.w));
        }
    }
    __syncwarp();
}

template <int NS>
__global__ void __launch_bounds__(NS * 128, 1) node_chain_kernel(const NodeChainArgs a, const int nimg) {
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* sm = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
    // layout: weight images | NS A images | NS x 4 staging blocks | vectors | barriers
    uint8_t* wimg = sm;
    uint8_t* aimg0 = wimg + nimg * kNcImg;
    uint8_t* stg0 = aimg0 + NS * kNcImg;
    float* vec = reinterpret_cast<float*>(stg0 + NS * 4 * kNcStgWarp);        // [stage][bias | gamma | beta][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(vec + kNcMaxStages * 3 * 128);   // [0] weights, [1 + s] stream s
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 1 + NS);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    constexpr int kTmemCols = NS <= 2 ? 256 : 512;
    if (t == 0) {
        for (int i = 0; i < 1 + NS; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_ptr);
    for (int i = t; i < a.nstage * 3 * 128; i += NS * 128) {
        const int s = i / 384, k = (i % 384) >> 7, c = i & 127;
        const float* p = k == 0 ? a.st[s].bias : (k == 1 ? a.st[s].ln_g : a.st[s].ln_b);
        vec[i] = p ? __ldg(p + c) : (k == 1 ? 1.f : 0.f);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr;
    if (t == 0) {   // resident weight images: stage 0 has nsrc of them, the others one each
        mbar_expect_tx(&bars[0], (uint32_t)(nimg * kNcImg));
        int k = 0;
        for (int s = 0; s < a.nstage; ++s) {
            const int cnt = s == 0 ? a.nsrc : 1;
            for (int j = 0; j < cnt * 2; ++j)
                bulk_g2s(wimg + k * kNcImg + j * kNcSlab, reinterpret_cast<const uint8_t*>(a.st[s].w_img) + j * kNcSlab, kNcSlab, &bars[0]);
            k += cnt;
        }
    }
    mbar_wait(&bars[0], 0);

    const int wg = t >> 7, r = t & 127, wq = (t >> 5) & 3;
    uint8_t* aimg = aimg0 + wg * kNcImg;
    float* stg = reinterpret_cast<float*>(stg0 + (wg * 4 + wq) * kNcStgWarp);
    uint64_t* bar = &bars[1 + wg];
    const uint32_t tcol = tm + wg * 128;                              // accumulator columns of this stream
    const uint32_t tacc = tcol + ((uint32_t)(wq * 32) << 16);       // + this warp's lane quarter
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const int64_t ntiles = (a.n + 127) / 128;
    uint32_t phase = 0;
    bool have_img = false;   // the image of this tile's first source was already built under the previous tile's last epilogue
    for (int64_t tile = (int64_t)blockIdx.x * NS + wg; tile < ntiles; tile += (int64_t)gridDim.x * NS) {
        const int64_t row0 = tile * 128;
        const int64_t next_row0 = (tile + (int64_t)gridDim.x * NS) * 128;   // >= n: no next tile
        const int64_t rbase = row0 + wq * 32;   // first row of this warp's 32 x 32 staging blocks
        int wi = 0;                              // weight image index of the current stage
        for (int s = 0; s < a.nstage; ++s) {
            const gmp_node_stage& S = a.st[s];
            const int nsrc = s == 0 ? a.nsrc : 1;
            for (int src = 0; src < nsrc; ++src) {
                if (s == 0) {
                    if (src > 0) {   // the MMAs over the first source must have read the image before it is overwritten
                        mbar_wait(bar, phase);
                        phase ^= 1u;
                    }
                    if (src > 0 || !have_img) load_tile_image(src == 0 ? a.a0 : a.a1, row0, a.n, aimg, r);
                }
                fence_proxy_async();
                tc_fence_before();
                bar_sync_named(1 + wg, 128);
                if (wq == 0) {
                    tc_fence_after();
                    if (elect_one()) {
                        umma_tile(tcol, smem_u32(aimg), kNcSlab, smem_u32(wimg + (wi + src) * kNcImg), kNcSlab, 128, idesc, src > 0);
                        umma_commit(bar);
                    }
                    __syncwarp();
                }
            }
            wi += nsrc;
            if (s + 1 == a.nstage) have_img = next_row0 < a.n;
            // the auxiliary rows of the first chunk are requested before the wait for the MMAs, those of chunk c + 1 while
            // chunk c is processed (one of the two when a stage has both a multiplier and a residual)
            const float* aux = S.mul_aux ? S.mul_aux : S.add_res;
            float4 pf[8];
            if (aux) aux_issue(aux, rbase, a.n, 0, lane, pf);
            mbar_wait(bar, phase);
            phase ^= 1u;
            tc_fence_after();

            // ---- epilogue of stage s ----
            const float* bias = vec + s * 384;
            const float* gam = bias + 128;
            const float* bet = bias + 256;
            const bool has_ln = S.ln_g != nullptr;
            float mean = 0.f, rstd = 1.f;
            if (has_ln) {   // two passes over the row in tensor memory: mean, then the centred second moment
                float sum = 0.f;
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    float v[32];
                    tmem_ld32(tacc + c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum += v[j] + bias[c0 + j];
                }
                mean = sum * (1.f / 128.f);
                float sq = 0.f;
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    float v[32];
                    tmem_ld32(tacc + c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float d = v[j] + bias[c0 + j] - mean;
                        sq = fmaf(d, d, sq);
                    }
                }
                rstd = rsqrtf(sq * (1.f / 128.f) + S.ln_eps);
            }
            // last stage: its MMAs were the last readers of the A image, so the next tile's image is built here, a quarter per
            // chunk, the loads in flight under the chunk's epilogue work (the read stream never pauses for the epilogues)
            const bool pre_next = (s + 1 == a.nstage) && next_row0 < a.n;
            for (int c0 = 0; c0 < 128; c0 += 32) {
                float4 nlo[4], nhi[4];
                if (pre_next) tile_quarter_issue(a.a0, next_row0, a.n, r, c0 >> 5, nlo, nhi);
                float v[32];
                tmem_ld32(tacc + c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += bias[c0 + j];
                if (S.out_pre) staged_store(v, S.out_pre, nullptr, rbase, a.n, c0, stg, lane);
                if (has_ln) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaf((v[j] - mean) * rstd, gam[c0 + j], bet[c0 + j]);
                }
                // (the activation id is uniform: one branch per chunk, not one per element)
                if (S.act == GMP_NODE_ACT_SSP) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = ssp_fast(v[j]);
                } else if (S.act == GMP_NODE_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                } else if (S.act == GMP_NODE_ACT_SILU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __fdividef(v[j], 1.f + ex2_fast(-1.4426950408889634f * v[j]));
                }
                if (aux) {
                    float x[32];
                    aux_commit(pf, stg, lane, x);
                    if (c0 + 32 < 128) aux_issue(aux, rbase, a.n, c0 + 32, lane, pf);
                    if (S.mul_aux) {
                        if (S.mul_mode == GMP_NODE_MUL_DSSP) {   // d ssp / d pre from the saved ssp output y: 1 - exp(-(y + ln 2)) = 1 - 2^(-y log2 e) / 2
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= fmaf(-0.5f, ex2_fast(-1.4426950408889634f * x[j]), 1.f);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= x[j];
                        }
                        if (S.add_res) staged_load(S.add_res, rbase, a.n, c0, stg, lane, x);
                    }
                    if (S.add_res) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += x[j];
                    }
                }
                if (S.out_f32 || S.out_bf16) staged_store(v, S.out_f32, reinterpret_cast<__nv_bfloat16*>(S.out_bf16), rbase, a.n, c0, stg, lane);
                if (pre_next) tile_quarter_commit(aimg, r, c0 >> 5, nlo, nhi);
                if (s + 1 < a.nstage) {   // this stage's result is the next stage's A operand
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<uint4*>(aimg + (c0 >> 6) * kNcSlab + sw128_chunk_off(r, ((c0 & 63) >> 3) + q)) =
                            make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                                       pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(tm);
}

// W [out][in] fp32 row-major -> bf16 K-major swizzled images of the B operand, one 128 x 128 image per 128 columns of K.
//   transpose = 0: B[r = out][k = in]  (y = x W^T);   transpose = 1: B[r = in][k = out]  (dx = g W)
__global__ void node_pack_w_kernel(const float* __restrict__ w, int out_dim, int in_dim, int transpose, uint8_t* __restrict__ img) {
    const int R = transpose ? in_dim : out_dim, K = transpose ? out_dim : in_dim;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // one 8-element chunk each
    if (idx >= R * (K >> 3)) return;
    const int r = idx / (K >> 3), kc = idx % (K >> 3), k0 = kc * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = transpose ? __ldg(w + (int64_t)(k0 + j) * in_dim + r) : __ldg(w + (int64_t)r * in_dim + k0 + j);
    // image (k0 / 128) holds k in [128 i, 128 i + 128): two slabs of 64
    uint8_t* dst = img + (k0 >> 7) * kNcImg + ((k0 & 127) >> 6) * kNcSlab + sw128_chunk_off(r, (k0 & 63) >> 3);
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}


// many 128 x 128 weights in one launch (a model packs every operand image of a step at once: 36 launches -> 1 for SchNet)
constexpr int kPackBatchMax = 48;
struct PackBatch {
    const float* w[kPackBatchMax];
    uint8_t* img[kPackBatchMax];
    int transpose[kPackBatchMax];
};
__global__ void __launch_bounds__(256) node_pack_w_batch_kernel(const PackBatch b) {
    const int m = blockIdx.x >> 3;                                   // 8 blocks of 256 chunks per matrix
    const int idx = (blockIdx.x & 7) * 256 + threadIdx.x;            // one 8-element chunk each: 128 rows x 16 chunks
    const int r = idx >> 4, k0 = (idx & 15) * 8;
    const float* w = b.w[m];
    const int tr = b.transpose[m];
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = tr ? __ldg(w + (k0 + j) * 128 + r) : __ldg(w + r * 128 + k0 + j);
    uint8_t* dst = b.img[m] + ((k0 & 127) >> 6) * kNcSlab + sw128_chunk_off(r, (k0 & 63) >> 3);
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// Backward of  a = act(LayerNorm(pre) * gamma + beta)  over 128-wide rows (EGNN mlp_upd, models/layers/egnn_layer.py:41-48):
//   d_pre = rstd * (gamma dy - mean(gamma dy) - xhat mean(gamma dy xhat)),   dy = g * act'(y)
// plus the per-CTA partial sums of d gamma = sum dy xhat and d beta = sum dy, and optionally the recomputed activations
// (the weight gradient of the following Linear needs them; they are not kept by the forward pass).
// One warp per row, a lane owns 4 consecutive columns: fully coalesced 512-byte rows, shuffle reductions, fp32 throughout.
constexpr int kLnbThreads = 256;
__global__ void __launch_bounds__(kLnbThreads) ln_act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ pre,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                                 int act, int64_t n, float* __restrict__ d_pre, float* __restrict__ act_out,
                                                                 float* __restrict__ parts) {
    __shared__ float red[kLnbThreads / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 gm = ldg4(gamma + lane * 4), bt = ldg4(beta + lane * 4);
    const float gmv[4] = {gm.x, gm.y, gm.z, gm.w}, btv[4] = {bt.x, bt.y, bt.z, bt.w};
    float dgam[4] = {0.f, 0.f, 0.f, 0.f}, dbet[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t wstride = (int64_t)gridDim.x * (kLnbThreads / 32);
    for (int64_t row = (int64_t)blockIdx.x * (kLnbThreads / 32) + warp; row < n; row += wstride) {
        const float4 xq = ldg4(pre + row * 128 + lane * 4), gq = ldg4(g + row * 128 + lane * 4);
        const float x[4] = {xq.x, xq.y, xq.z, xq.w}, gi[4] = {gq.x, gq.y, gq.z, gq.w};
        const float mean = warp_sum(x[0] + x[1] + x[2] + x[3]) * (1.f / 128.f);
        float sq = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) sq = fmaf(x[k] - mean, x[k] - mean, sq);
        const float rstd = rsqrtf(warp_sum(sq) * (1.f / 128.f) + eps);
        float xh[4], gdy[4], av[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            xh[k] = (x[k] - mean) * rstd;
            const float y = fmaf(xh[k], gmv[k], btv[k]);
            float da;
            if (act == GMP_NODE_ACT_RELU) {
                da = y > 0.f ? 1.f : 0.f;
                av[k] = fmaxf(y, 0.f);
            } else if (act == GMP_NODE_ACT_SILU) {
                const float sg = sigmoidf_(y);
                da = sg * fmaf(y, 1.f - sg, 1.f);
                av[k] = y * sg;
            } else {
                da = 1.f;
                av[k] = y;
            }
            const float dy = gi[k] * da;
            dgam[k] = fmaf(dy, xh[k], dgam[k]);
            dbet[k] += dy;
            gdy[k] = dy * gmv[k];
            s1 += gdy[k];
            s2 = fmaf(gdy[k], xh[k], s2);
        }
        s1 = warp_sum(s1) * (1.f / 128.f);
        s2 = warp_sum(s2) * (1.f / 128.f);
        float4 o;
        o.x = rstd * (gdy[0] - s1 - xh[0] * s2);
        o.y = rstd * (gdy[1] - s1 - xh[1] * s2);
        o.z = rstd * (gdy[2] - s1 - xh[2] * s2);
        o.w = rstd * (gdy[3] - s1 - xh[3] * s2);
        *reinterpret_cast<float4*>(d_pre + row * 128 + lane * 4) = o;
        if (act_out) *reinterpret_cast<float4*>(act_out + row * 128 + lane * 4) = make_float4(av[0], av[1], av[2], av[3]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        red[warp][lane * 4 + k] = dgam[k];
        red[warp][128 + lane * 4 + k] = dbet[k];
    }
    __syncthreads();
    {   // fixed-order sum over the CTA's warps: thread t owns element t of [d gamma | d beta]
        float sacc = 0.f;
#pragma unroll
        for (int w = 0; w < kLnbThreads / 32; ++w) sacc += red[w][threadIdx.x];
        parts[(int64_t)blockIdx.x * 256 + threadIdx.x] = sacc;
    }
}

template <int NS>
int launch_chain(const NodeChainArgs& a, int nimg, gmp_stream_t stream) {
    const int smem = nimg * kNcImg + NS * kNcImg + NS * 4 * kNcStgWarp + kNcMaxStages * 3 * 128 * 4 + 64 + 1024;
    GMP_REQUIRE(smem <= 232448, "node_chain: %d bytes of shared memory", smem);
    GMP_CUDA(cudaFuncSetAttribute(node_chain_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t ntiles = ceil_div(a.n, 128);
    const int64_t want = ceil_div(ntiles, NS);
    const unsigned grid = (unsigned)(want < num_sms() ? want : num_sms());
    node_chain_kernel<NS><<<grid, NS * 128, smem, stream>>>(a, nimg);
    return check_launch("node_chain_kernel");
}

}  // namespace
}  // namespace gmp

using namespace gmp;

extern "C" {

int64_t gmp_node_w_image_bytes(int32_t out_dim, int32_t in_dim, int32_t transpose) {
    const int R = transpose ? in_dim : out_dim, K = transpose ? out_dim : in_dim;
    if (R != 128 || K < 128 || K % 128 != 0) return -1;
    return (int64_t)(K / 128) * kNcImg;
}

int gmp_node_pack_w(const float* w, int32_t out_dim, int32_t in_dim, int32_t transpose, void* img, gmp_stream_t stream) {
    GMP_REQUIRE(w && img, "node_pack_w: NULL pointer");
    GMP_REQUIRE(gmp_node_w_image_bytes(out_dim, in_dim, transpose) > 0,
                "node_pack_w: the operand must have 128 rows and a multiple of 128 columns (out %d, in %d, transpose %d)", out_dim, in_dim, transpose);
    const int chunks = out_dim * in_dim / 8;
    node_pack_w_kernel<<<(unsigned)ceil_div(chunks, 256), 256, 0, stream>>>(w, out_dim, in_dim, transpose, reinterpret_cast<uint8_t*>(img));
    return check_launch("node_pack_w_kernel");
}

int gmp_node_pack_w_batch(const float* const* w, const int32_t* transpose, void* const* img, int32_t count, gmp_stream_t stream) {
    GMP_REQUIRE(count >= 0 && (count == 0 || (w && transpose && img)), "node_pack_w_batch: bad arguments");
    for (int base = 0; base < count; base += kPackBatchMax) {
        PackBatch b;
        const int m = count - base < kPackBatchMax ? count - base : kPackBatchMax;
        for (int i = 0; i < m; ++i) {
            GMP_REQUIRE(w[base + i] && img[base + i], "node_pack_w_batch: NULL pointer at %d", base + i);
            b.w[i] = w[base + i];
            b.img[i] = reinterpret_cast<uint8_t*>(img[base + i]);
            b.transpose[i] = transpose[base + i];
        }
        node_pack_w_batch_kernel<<<(unsigned)(m * 8), 256, 0, stream>>>(b);
        if (int rc = check_launch("node_pack_w_batch_kernel")) return rc;
    }
    return GMP_OK;
}

int gmp_node_chain_tc(const float* a0, const float* a1, int64_t n, int32_t nstage, const gmp_node_stage* stages, gmp_stream_t stream) {
    GMP_REQUIRE(a0 && stages && n >= 0 && nstage >= 1 && nstage <= kNcMaxStages, "node_chain_tc: bad arguments (1 <= nstage <= 3)");
    if (n == 0) return GMP_OK;
    NodeChainArgs a;
    a.a0 = a0;
    a.a1 = a1;
    a.n = n;
    a.nstage = nstage;
    a.nsrc = a1 ? 2 : 1;
    for (int s = 0; s < nstage; ++s) {
        a.st[s] = stages[s];
        GMP_REQUIRE(stages[s].w_img, "node_chain_tc: stage %d has no weight image", s);
        GMP_REQUIRE((stages[s].ln_g == nullptr) == (stages[s].ln_b == nullptr), "node_chain_tc: LayerNorm needs both gamma and beta");
        GMP_REQUIRE(stages[s].act >= GMP_NODE_ACT_NONE && stages[s].act <= GMP_NODE_ACT_SILU, "node_chain_tc: unknown activation %d", stages[s].act);
    }
    const int nimg = nstage - 1 + a.nsrc;
    // streams per CTA: as many as the shared memory left by the resident weight images holds (each: 32 KB image + 18 KB staging)
    if (nimg <= 1) return launch_chain<3>(a, nimg, stream);
    if (nimg <= 3) return launch_chain<2>(a, nimg, stream);
    return launch_chain<1>(a, nimg, stream);
}

int32_t gmp_ln_act_bwd_num_parts(int64_t n) {
    const int64_t want = ceil_div(n, kLnbThreads / 32);
    const int64_t cap = 4 * (int64_t)num_sms();
    return (int32_t)(want < 1 ? 1 : (want < cap ? want : cap));
}

int gmp_ln_act_bwd(const float* g_out, const float* pre, const float* gamma, const float* beta, float eps, int32_t act, int64_t n,
                   float* d_pre, float* act_out, float* parts, gmp_stream_t stream) {
    GMP_REQUIRE(g_out && pre && gamma && beta && d_pre && parts && n >= 1, "ln_act_bwd: bad arguments");
    GMP_REQUIRE(act == GMP_NODE_ACT_NONE || act == GMP_NODE_ACT_RELU || act == GMP_NODE_ACT_SILU, "ln_act_bwd: activation %d", act);
    ln_act_bwd_kernel<<<gmp_ln_act_bwd_num_parts(n), kLnbThreads, 0, stream>>>(g_out, pre, gamma, beta, eps, act, n, d_pre, act_out, parts);
    return check_launch("ln_act_bwd_kernel");
}

}  // extern "C"
