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

__device__ __forceinline__ float act_apply(float v, int act) {
    switch (act) {
        case GMP_NODE_ACT_SSP: return ssp(v);
        case GMP_NODE_ACT_RELU: return fmaxf(v, 0.f);
        case GMP_NODE_ACT_SILU: return v * sigmoidf_(v);
        default: return v;
    }
}

// 128 rows x 128 fp32 columns of `src` starting at row0 -> bf16 K-major swizzled image; 128 threads (r = 0..127)
__device__ __forceinline__ void load_tile_image(const float* __restrict__ src, int64_t row0, int64_t n, uint8_t* img, int r) {
    const int ch = r & 15, rr0 = r >> 4;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float4 lo[4], hi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t row = row0 + rr0 + 8 * (4 * b + i);
            lo[i] = hi[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < n) {
                lo[i] = ldg4(src + row * 128 + ch * 8);
                hi[i] = ldg4(src + row * 128 + ch * 8 + 4);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = rr0 + 8 * (4 * b + i);
            *reinterpret_cast<uint4*>(img + (ch >> 3) * kNcSlab + sw128_chunk_off(rr, ch & 7)) =
                make_uint4(pack_bf16(lo[i].x, lo[i].y), pack_bf16(lo[i].z, lo[i].w), pack_bf16(hi[i].x, hi[i].y), pack_bf16(hi[i].z, hi[i].w));
        }
    }
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
    for (int64_t tile = (int64_t)blockIdx.x * NS + wg; tile < ntiles; tile += (int64_t)gridDim.x * NS) {
        const int64_t row0 = tile * 128;
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
                    load_tile_image(src == 0 ? a.a0 : a.a1, row0, a.n, aimg, r);
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
            for (int c0 = 0; c0 < 128; c0 += 32) {
                float v[32];
                tmem_ld32(tacc + c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += bias[c0 + j];
                if (S.out_pre) staged_store(v, S.out_pre, nullptr, rbase, a.n, c0, stg, lane);
                if (has_ln) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaf((v[j] - mean) * rstd, gam[c0 + j], bet[c0 + j]);
                }
                if (S.act != GMP_NODE_ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = act_apply(v[j], S.act);
                }
                if (S.mul_aux) {
                    float x[32];
                    staged_load(S.mul_aux, rbase, a.n, c0, stg, lane, x);
                    if (S.mul_mode == GMP_NODE_MUL_DSSP) {   // d ssp / d pre from the saved ssp output y: 1 - exp(-(y + ln 2))
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= -expm1f(-(x[j] + 0.6931471805599453f));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= x[j];
                    }
                }
                if (S.add_res) {
                    float x[32];
                    staged_load(S.add_res, rbase, a.n, c0, stg, lane, x);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += x[j];
                }
                if (S.out_f32 || S.out_bf16) staged_store(v, S.out_f32, reinterpret_cast<__nv_bfloat16*>(S.out_bf16), rbase, a.n, c0, stg, lane);
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

}  // extern "C"
