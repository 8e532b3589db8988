// e3nn nn.Gate / nn.Activation as used by the TFN layer (models/layers/tfn_layer.py:45-63, 89-90), forward and backward as
// one elementwise kernel each instead of ~8 ATen launches (slice, silu, mul, sigmoid, mul, index_select, mul, cat):
//   input  [n, ns + ng + nv] = [scalars | gates | gated]      output [n, ns + nv] = [silu(s) c_s | gated * sigmoid(gate) c_g]
// Every gate multiplies the 2l + 1 components of one gated irrep copy: expand[j] = gate of gated element j, and the elements
// of gate u are the contiguous range [gstart[u], gstart[u] + gdim[u]) of the gated block.
#include "common.cuh"

namespace gmp {
namespace {

__global__ void __launch_bounds__(256) gate_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ expand, int64_t n, int ns,
                                                       int ng, int nv, float c_s, float c_g, float* __restrict__ out) {
    const int wo = ns + nv, wi = ns + ng + nv;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * wo) return;
    const int64_t row = idx / wo;
    const int c = (int)(idx - row * wo);
    const float* xr = x + row * wi;
    float v;
    if (c < ns) {
        const float s = __ldg(xr + c);
        v = s * sigmoidf_(s) * c_s;
    } else {
        const int j = c - ns;
        v = __ldg(xr + ns + ng + j) * sigmoidf_(__ldg(xr + ns + __ldg(expand + j))) * c_g;
    }
    out[idx] = v;
}

__global__ void __launch_bounds__(256) gate_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const int32_t* __restrict__ expand,
                                                       const int32_t* __restrict__ gstart, const int32_t* __restrict__ gdim, int64_t n, int ns,
                                                       int ng, int nv, float c_s, float c_g, float* __restrict__ dx) {
    const int wo = ns + nv, wi = ns + ng + nv;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * wi) return;
    const int64_t row = idx / wi;
    const int c = (int)(idx - row * wi);
    const float* xr = x + row * wi;
    const float* gr = g + row * wo;
    float v;
    if (c < ns) {
        const float s = __ldg(xr + c), sg = sigmoidf_(s);
        v = __ldg(gr + c) * c_s * sg * fmaf(s, 1.f - sg, 1.f);
    } else if (c < ns + ng) {
        const int u = c - ns, b = __ldg(gstart + u), d = __ldg(gdim + u);
        float acc = 0.f;
        for (int k = 0; k < d; ++k) acc = fmaf(__ldg(gr + ns + b + k), __ldg(xr + ns + ng + b + k), acc);
        const float sg = sigmoidf_(__ldg(xr + c));
        v = c_g * sg * (1.f - sg) * acc;
    } else {
        const int j = c - ns - ng;
        v = __ldg(gr + ns + j) * sigmoidf_(__ldg(xr + ns + __ldg(expand + j))) * c_g;
    }
    dx[idx] = v;
}

}  // namespace
}  // namespace gmp

using namespace gmp;

extern "C" {

int gmp_gate_fwd(const float* x, const int32_t* expand, int64_t n, int32_t num_scalars, int32_t num_gates, int32_t num_gated,
                 float c_silu, float c_sigmoid, float* out, gmp_stream_t stream) {
    GMP_REQUIRE(n >= 0 && num_scalars >= 0 && num_gates >= 0 && num_gated >= 0 && num_scalars + num_gated > 0, "gate_fwd: bad sizes");
    if (n == 0) return GMP_OK;
    GMP_REQUIRE(x && out && (num_gated == 0 || expand), "gate_fwd: NULL pointer");
    const int64_t total = n * (num_scalars + num_gated);
    gate_fwd_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(x, expand, n, num_scalars, num_gates, num_gated, c_silu, c_sigmoid, out);
    return check_launch("gate_fwd_kernel");
}

int gmp_gate_bwd(const float* x, const float* g_out, const int32_t* expand, const int32_t* gate_start, const int32_t* gate_dim, int64_t n,
                 int32_t num_scalars, int32_t num_gates, int32_t num_gated, float c_silu, float c_sigmoid, float* dx, gmp_stream_t stream) {
    GMP_REQUIRE(n >= 0 && num_scalars >= 0 && num_gates >= 0 && num_gated >= 0 && num_scalars + num_gated > 0, "gate_bwd: bad sizes");
    if (n == 0) return GMP_OK;
    GMP_REQUIRE(x && g_out && dx && (num_gates == 0 || (expand && gate_start && gate_dim)), "gate_bwd: NULL pointer");
    const int64_t total = n * (num_scalars + num_gates + num_gated);
    gate_bwd_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(x, g_out, expand, gate_start, gate_dim, n, num_scalars, num_gates,
                                                                         num_gated, c_silu, c_sigmoid, dx);
    return check_launch("gate_bwd_kernel");
}

}  // extern "C"
