// Graph construction: radius_graph (torch_cluster CUDA canonical order), cell lists, CSR build.
// Integer / byte path: results are bit-exact by construction (no floating-point reductions).
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"

namespace gmp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return GMP_ERR_CUDA;
    }
    return GMP_OK;
}

// squared distance exactly as torch_cluster's radius kernel accumulates it (dist += d*d, FMA-contracted)
__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    float d = __fmul_rn(dx, dx);
    d = __fmaf_rn(dy, dy, d);
    d = __fmaf_rn(dz, dz, d);
    return d;
}

template <bool FILL>
__global__ void radius_brute_kernel(const float* __restrict__ pos, const int64_t* __restrict__ gptr, int64_t ngraphs,
                                    int64_t n, float r2, int cap, int loop, int32_t* __restrict__ deg,
                                    const int64_t* __restrict__ rowptr, int64_t* __restrict__ esrc,
                                    int64_t* __restrict__ edst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // example of node i: last g with gptr[g] <= i
    int64_t lo = 0, hi = ngraphs;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (gptr[mid] <= i) lo = mid; else hi = mid;
    }
    const int64_t j0 = gptr[lo], j1 = gptr[lo + 1];
    const float qx = pos[3 * i], qy = pos[3 * i + 1], qz = pos[3 * i + 2];
    int count = 0, written = 0;
    int64_t base = FILL ? rowptr[i] : 0;
    for (int64_t j = j0; j < j1; ++j) {
        const float d = sqdist(pos[3 * j], pos[3 * j + 1], pos[3 * j + 2], qx, qy, qz);
        if (d < r2) {
            if (loop || j != i) {
                if (FILL) {
                    esrc[base + written] = j;
                    edst[base + written] = i;
                }
                ++written;
            }
            ++count;
        }
        if (count >= cap) break;
    }
    if (!FILL) deg[i] = written;
}

// ----------------------------------------------------------------------------------------------
// uniform cell list
// ----------------------------------------------------------------------------------------------
struct Cells {
    float ox, oy, oz, inv;
    int nx, ny, nz;
};

__device__ __forceinline__ int3 cell_coord(const Cells& c, float x, float y, float z) {
    int cx = (int)floorf((x - c.ox) * c.inv), cy = (int)floorf((y - c.oy) * c.inv), cz = (int)floorf((z - c.oz) * c.inv);
    cx = min(max(cx, 0), c.nx - 1);
    cy = min(max(cy, 0), c.ny - 1);
    cz = min(max(cz, 0), c.nz - 1);
    return make_int3(cx, cy, cz);
}

__global__ void cell_assign_kernel(const float* __restrict__ pos, int64_t n, Cells c, int32_t* __restrict__ cell_of,
                                   int32_t* __restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int3 cc = cell_coord(c, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    const int id = (cc.x * c.ny + cc.y) * c.nz + cc.z;
    cell_of[i] = id;
    atomicAdd(counts + id, 1);  // integer histogram: the result does not depend on arrival order
}

__global__ void scan_small_kernel(int32_t* __restrict__ a, int64_t n) {
    // in-place exclusive scan of a[0..n] (n+1 entries, a[n] receives the total); single block
    __shared__ int32_t carry;
    __shared__ int32_t buf[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base <= n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        int32_t v = (i < n) ? a[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int32_t t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        const int32_t incl = buf[threadIdx.x];
        if (i <= n) a[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
}

__global__ void cell_fill_kernel(const int32_t* __restrict__ cell_of, int64_t n, const int32_t* __restrict__ cell_start,
                                 int32_t* __restrict__ cursor, int32_t* __restrict__ cell_nodes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = cell_of[i];
    const int slot = atomicAdd(cursor + id, 1);
    cell_nodes[cell_start[id] + slot] = (int32_t)i;  // order inside a cell is irrelevant: hits are sorted below
}

constexpr int kMaxCap = 160;

template <bool FILL>
__global__ void radius_cells_kernel(const float* __restrict__ pos, int64_t n, float r2, Cells c,
                                    const int32_t* __restrict__ cell_start, const int32_t* __restrict__ cell_nodes,
                                    int cap, int loop, int32_t* __restrict__ deg, const int64_t* __restrict__ rowptr,
                                    int64_t* __restrict__ esrc, int64_t* __restrict__ edst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float qx = pos[3 * i], qy = pos[3 * i + 1], qz = pos[3 * i + 2];
    const int3 cc = cell_coord(c, qx, qy, qz);
    int32_t buf[kMaxCap];  // the `cap` smallest hit indices, ascending
    int cnt = 0;
    for (int dx = -1; dx <= 1; ++dx) {
        const int x = cc.x + dx;
        if (x < 0 || x >= c.nx) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = cc.y + dy;
            if (y < 0 || y >= c.ny) continue;
            for (int dz = -1; dz <= 1; ++dz) {
                const int z = cc.z + dz;
                if (z < 0 || z >= c.nz) continue;
                const int id = (x * c.ny + y) * c.nz + z;
                for (int k = cell_start[id]; k < cell_start[id + 1]; ++k) {
                    const int j = cell_nodes[k];
                    const float d = sqdist(pos[3 * (int64_t)j], pos[3 * (int64_t)j + 1], pos[3 * (int64_t)j + 2], qx, qy, qz);
                    if (!(d < r2)) continue;
                    if (cnt == cap && j > buf[cnt - 1]) continue;
                    int p = (cnt < cap) ? cnt : cap - 1;  // slot that will be shifted out / appended
                    while (p > 0 && buf[p - 1] > j) {
                        buf[p] = buf[p - 1];
                        --p;
                    }
                    buf[p] = j;
                    if (cnt < cap) ++cnt;
                }
            }
        }
    }
    int written = 0;
    const int64_t base = FILL ? rowptr[i] : 0;
    for (int k = 0; k < cnt; ++k) {
        if (!loop && buf[k] == (int32_t)i) continue;
        if (FILL) {
            esrc[base + written] = buf[k];
            edst[base + written] = i;
        }
        ++written;
    }
    if (!FILL) deg[i] = written;
}

// ----------------------------------------------------------------------------------------------
// exclusive scan int32 -> int64 (three small kernels)
// ----------------------------------------------------------------------------------------------
__global__ void scan_block_sums_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ sums) {
    __shared__ int64_t red[32];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    int64_t v = (i < n) ? in[i] : 0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = red[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = v;
    }
}

__global__ void scan_sums_kernel(int64_t* __restrict__ sums, int64_t nb) {
    // exclusive scan of sums[0..nb) in place, total in sums[nb]; single thread block, sequential chunks
    __shared__ int64_t carry;
    __shared__ int64_t buf[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base <= nb; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = (i < nb) ? sums[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int64_t t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        const int64_t incl = buf[threadIdx.x];
        if (i <= nb) sums[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
}

__global__ void scan_apply_kernel(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ sums,
                                  int64_t* __restrict__ out) {
    __shared__ int64_t buf[1024];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int64_t v = (i < n) ? in[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int64_t t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
        __syncthreads();
        buf[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) out[i] = sums[blockIdx.x] + buf[threadIdx.x] - v;
    if (i == n - 1) out[n] = sums[blockIdx.x] + buf[threadIdx.x];
}

// ----------------------------------------------------------------------------------------------
// CSR: stable counting sort by an int64 key
// ----------------------------------------------------------------------------------------------
__global__ void csr_count_kernel(const int64_t* __restrict__ index, int64_t E, int64_t n, int32_t* __restrict__ counts) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t r = index[e];
    if (r >= 0 && r < n) atomicAdd(counts + r, 1);
}

__global__ void csr_bucket_kernel(const int64_t* __restrict__ index, int64_t E, int64_t n,
                                  const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                                  int32_t* __restrict__ tmp) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t r = index[e];
    if (r < 0 || r >= n) return;
    const int slot = atomicAdd(cursor + r, 1);
    tmp[rowptr[r] + slot] = (int32_t)e;
}

// Rows up to kLongRow edges: one warp per row rank-sorts the (unique) edge ids of the row ascending -> stable order.
// Longer rows (pooling one large graph puts every node into one row) are left to csr_sort_long_rows_kernel: the rank
// sort is quadratic in the row length.
constexpr int kLongRow = 1024;

__global__ void csr_rank_kernel(const int32_t* __restrict__ rowptr, int64_t n, const int32_t* __restrict__ tmp,
                                int32_t* __restrict__ perm) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int b = rowptr[row], d = rowptr[row + 1] - b;
    if (d > kLongRow) return;
    for (int k = lane; k < d; k += 32) {
        const int32_t mine = tmp[b + k];
        int rank = 0;
        for (int m = 0; m < d; ++m) rank += (tmp[b + m] < mine);
        perm[b + rank] = mine;
    }
}

// Long rows: a block-wide LSD radix sort (8 bits per pass) of the row's edge ids, ping-ponging between the bucketed
// array `tmp` and `perm`, so the result is the ascending (= stable) order in linear time.  A pass streams the row in
// tiles of 1024 keys: thread = key (tile order = current order, which is what makes the pass stable); the rank of a key
// among the equal digits of its tile comes from __match_any_sync inside the warp plus a scan of the per-warp counts.
struct LongRowSmem {
    int wcount[32][257];   // per-warp digit counts of the current tile (257: digit 256 = "no key", and bank spread)
    int base[256];         // running output offset of every digit within the row
    int hist[256];
    int rows[1024];
    int n_rows;
};

__device__ void sort_long_row(LongRowSmem& S, int b, int d, int32_t* tmp, int32_t* perm, int passes) {
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    int32_t *src = tmp + b, *dst = perm + b;
    for (int ps = 0; ps < passes; ++ps) {
        const int shift = 8 * ps;
        if (t < 256) S.hist[t] = 0;
        __syncthreads();
        for (int k = t; k < d; k += 1024) atomicAdd(&S.hist[(src[k] >> shift) & 255], 1);   // integer histogram
        __syncthreads();
        if (t == 0) {
            int acc = 0;
            for (int q = 0; q < 256; ++q) { S.base[q] = acc; acc += S.hist[q]; }
        }
        __syncthreads();
        for (int k0 = 0; k0 < d; k0 += 1024) {
            const int k = k0 + t;
            const bool ok = k < d;
            const int32_t key = ok ? src[k] : 0;
            const int dg = ok ? ((key >> shift) & 255) : 256;
            for (int q = lane; q < 257; q += 32) S.wcount[warp][q] = 0;
            __syncwarp();
            const unsigned peers = __match_any_sync(0xffffffffu, dg);
            const int rank_w = __popc(peers & ((1u << lane) - 1u));
            if (rank_w == 0) S.wcount[warp][dg] = __popc(peers);
            __syncthreads();
            if (ok) {
                int off = 0;
                for (int w = 0; w < warp; ++w) off += S.wcount[w][dg];
                dst[S.base[dg] + off + rank_w] = key;
            }
            __syncthreads();
            if (t < 256) {
                int tot = 0;
                for (int w = 0; w < 32; ++w) tot += S.wcount[w][t];
                S.base[t] += tot;
            }
            __syncthreads();
        }
        int32_t* sw = src; src = dst; dst = sw;
    }
    if (src != perm + b)
        for (int k = t; k < d; k += 1024) perm[b + k] = src[k];
    __syncthreads();
}

// each block owns a contiguous range of rows, finds the long ones 1024 rows at a time and sorts them one after the other
__global__ void __launch_bounds__(1024) csr_sort_long_rows_kernel(const int32_t* __restrict__ rowptr, int64_t n, int32_t* tmp,
                                                                  int32_t* perm, int passes) {
    __shared__ LongRowSmem S;
    const int t = threadIdx.x;
    const int64_t rows_per_block = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r_lo = rows_per_block * blockIdx.x, r_hi = min(n, r_lo + rows_per_block);
    for (int64_t r0 = r_lo; r0 < r_hi; r0 += 1024) {
        if (t == 0) S.n_rows = 0;
        __syncthreads();
        const int64_t mine = r0 + t;
        if (mine < r_hi && rowptr[mine + 1] - rowptr[mine] > kLongRow) S.rows[atomicAdd(&S.n_rows, 1)] = t;
        __syncthreads();
        const int nl = S.n_rows;
        for (int i = 0; i < nl; ++i) {
            const int64_t row = r0 + S.rows[i];
            sort_long_row(S, rowptr[row], rowptr[row + 1] - rowptr[row], tmp, perm, passes);
        }
        __syncthreads();
    }
}

__global__ void gather_i64_i32_kernel(const int64_t* __restrict__ src, const int32_t* __restrict__ perm, int64_t E,
                                      int32_t* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) out[e] = (int32_t)src[perm ? perm[e] : e];
}

__global__ void is_sorted_kernel(const int64_t* __restrict__ index, int64_t E, int32_t* __restrict__ flag) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e + 1 < E && index[e] > index[e + 1]) *flag = 0;
}

// coalesce (torch_geometric.utils.to_undirected / coalesce): after the lexicographic sort, keep the first of every run of
// equal (row, col) pairs
__global__ void mark_unique_pairs_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                         const int32_t* __restrict__ perm, int64_t E, int32_t* __restrict__ keep) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= E) return;
    const int64_t a = perm ? perm[k] : k;
    int flag = 1;
    if (k > 0) {
        const int64_t b = perm ? perm[k - 1] : k - 1;
        flag = (row[a] != row[b]) || (col[a] != col[b]);
    }
    keep[k] = flag;
}

__global__ void compact_pairs_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                     const int32_t* __restrict__ perm, const int32_t* __restrict__ keep,
                                     const int64_t* __restrict__ pos, int64_t E, int64_t* __restrict__ out_row,
                                     int64_t* __restrict__ out_col) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= E || !keep[k]) return;
    const int64_t a = perm ? perm[k] : k;
    out_row[pos[k]] = row[a];
    out_col[pos[k]] = col[a];
}

static Cells make_cells(float cell, const float* origin, const int32_t* dims) {
    Cells c;
    c.ox = origin[0]; c.oy = origin[1]; c.oz = origin[2];
    c.inv = 1.0f / cell;
    c.nx = dims[0]; c.ny = dims[1]; c.nz = dims[2];
    return c;
}

}  // namespace gmp

using namespace gmp;

extern "C" {

int gmp_version(void) { return 100; }
const char* gmp_last_error(void) { return gmp::g_err; }

int gmp_radius_graph_count(const float* pos, const int64_t* graph_ptr, int64_t num_graphs, int64_t n, float r,
                           int32_t max_num_neighbors, int32_t loop, int32_t* deg, gmp_stream_t stream) {
    GMP_REQUIRE(pos && graph_ptr && deg && num_graphs >= 1 && n >= 0, "radius_graph_count: bad arguments");
    if (n == 0) return GMP_OK;
    const int cap = max_num_neighbors + (loop ? 0 : 1);
    radius_brute_kernel<false><<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(
        pos, graph_ptr, num_graphs, n, r * r, cap, loop, deg, nullptr, nullptr, nullptr);
    return check_launch("radius_brute_kernel<count>");
}

int gmp_radius_graph_fill(const float* pos, const int64_t* graph_ptr, int64_t num_graphs, int64_t n, float r,
                          int32_t max_num_neighbors, int32_t loop, const int64_t* rowptr, int64_t* edge_src,
                          int64_t* edge_dst, gmp_stream_t stream) {
    GMP_REQUIRE(pos && graph_ptr && rowptr && edge_src && edge_dst && num_graphs >= 1, "radius_graph_fill: bad arguments");
    if (n == 0) return GMP_OK;
    const int cap = max_num_neighbors + (loop ? 0 : 1);
    radius_brute_kernel<true><<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(
        pos, graph_ptr, num_graphs, n, r * r, cap, loop, nullptr, rowptr, edge_src, edge_dst);
    return check_launch("radius_brute_kernel<fill>");
}

int gmp_cells_build(const float* pos, int64_t n, float cell, const float* origin_host, const int32_t* dims_host,
                    int32_t* cell_of, int32_t* cell_start, int32_t* cell_nodes, int32_t* cursor_ws,
                    gmp_stream_t stream) {
    GMP_REQUIRE(pos && origin_host && dims_host && cell_of && cell_start && cell_nodes && cursor_ws && cell > 0,
                "cells_build: bad arguments");
    const int64_t ncells = (int64_t)dims_host[0] * dims_host[1] * dims_host[2];
    GMP_REQUIRE(ncells > 0 && ncells < (1ll << 31) && n < (1ll << 31), "cells_build: grid too large");
    const Cells c = make_cells(cell, origin_host, dims_host);
    GMP_CUDA(cudaMemsetAsync(cell_start, 0, (ncells + 1) * sizeof(int32_t), stream));
    GMP_CUDA(cudaMemsetAsync(cursor_ws, 0, ncells * sizeof(int32_t), stream));
    if (n == 0) return GMP_OK;
    cell_assign_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(pos, n, c, cell_of, cell_start);
    scan_small_kernel<<<1, 1024, 0, stream>>>(cell_start, ncells);
    cell_fill_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(cell_of, n, cell_start, cursor_ws, cell_nodes);
    return check_launch("cells_build");
}

int gmp_radius_cells_count(const float* pos, int64_t n, float r, float cell, const float* origin_host,
                           const int32_t* dims_host, const int32_t* cell_start, const int32_t* cell_nodes,
                           int32_t max_num_neighbors, int32_t loop, int32_t* deg, gmp_stream_t stream) {
    GMP_REQUIRE(pos && cell_start && cell_nodes && deg && cell >= r, "radius_cells_count: bad arguments (cell >= r required)");
    const int cap = max_num_neighbors + (loop ? 0 : 1);
    GMP_REQUIRE(cap <= kMaxCap, "radius_cells: max_num_neighbors + 1 must be <= %d", kMaxCap);
    if (n == 0) return GMP_OK;
    const Cells c = make_cells(cell, origin_host, dims_host);
    radius_cells_kernel<false><<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(
        pos, n, r * r, c, cell_start, cell_nodes, cap, loop, deg, nullptr, nullptr, nullptr);
    return check_launch("radius_cells_kernel<count>");
}

int gmp_radius_cells_fill(const float* pos, int64_t n, float r, float cell, const float* origin_host,
                          const int32_t* dims_host, const int32_t* cell_start, const int32_t* cell_nodes,
                          int32_t max_num_neighbors, int32_t loop, const int64_t* rowptr, int64_t* edge_src,
                          int64_t* edge_dst, gmp_stream_t stream) {
    GMP_REQUIRE(pos && cell_start && cell_nodes && rowptr && edge_src && edge_dst && cell >= r,
                "radius_cells_fill: bad arguments");
    const int cap = max_num_neighbors + (loop ? 0 : 1);
    GMP_REQUIRE(cap <= kMaxCap, "radius_cells: max_num_neighbors + 1 must be <= %d", kMaxCap);
    if (n == 0) return GMP_OK;
    const Cells c = make_cells(cell, origin_host, dims_host);
    radius_cells_kernel<true><<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(
        pos, n, r * r, c, cell_start, cell_nodes, cap, loop, nullptr, rowptr, edge_src, edge_dst);
    return check_launch("radius_cells_kernel<fill>");
}

int gmp_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, int64_t* ws, gmp_stream_t stream) {
    GMP_REQUIRE(out && ws && n >= 0 && (in || n == 0), "exclusive_scan: bad arguments");
    if (n == 0) {
        GMP_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), stream));
        return GMP_OK;
    }
    const int64_t nb = ceil_div(n, 1024);
    scan_block_sums_kernel<<<(unsigned)nb, 1024, 0, stream>>>(in, n, ws);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(ws, nb);
    scan_apply_kernel<<<(unsigned)nb, 1024, 0, stream>>>(in, n, ws, out);
    return check_launch("exclusive_scan");
}

int gmp_csr_count(const int64_t* index, int64_t num_edges, int64_t n, int32_t* counts, gmp_stream_t stream) {
    GMP_REQUIRE(counts && n >= 0 && num_edges >= 0 && (index || num_edges == 0), "csr_count: bad arguments");
    GMP_CUDA(cudaMemsetAsync(counts, 0, n * sizeof(int32_t), stream));
    if (num_edges == 0) return GMP_OK;
    csr_count_kernel<<<(unsigned)ceil_div(num_edges, 256), 256, 0, stream>>>(index, num_edges, n, counts);
    return check_launch("csr_count_kernel");
}

int gmp_csr_fill(const int64_t* index, int64_t num_edges, int64_t n, const int32_t* rowptr, int32_t* cursor_ws,
                 int32_t* tmp_ws, int32_t* perm, gmp_stream_t stream) {
    GMP_REQUIRE(rowptr && cursor_ws && (num_edges == 0 || (index && tmp_ws && perm)), "csr_fill: bad arguments");
    GMP_REQUIRE(num_edges < (1ll << 31), "csr_fill: more than 2^31-1 edges");
    if (num_edges == 0 || n == 0) return GMP_OK;
    GMP_CUDA(cudaMemsetAsync(cursor_ws, 0, n * sizeof(int32_t), stream));
    csr_bucket_kernel<<<(unsigned)ceil_div(num_edges, 256), 256, 0, stream>>>(index, num_edges, n, rowptr, cursor_ws, tmp_ws);
    csr_rank_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, stream>>>(rowptr, n, tmp_ws, perm);
    if (num_edges > kLongRow) {   // rows longer than kLongRow: linear-time radix sort of their edge ids (edge ids < num_edges)
        int passes = 1;
        while (passes < 4 && (num_edges >> (8 * passes)) != 0) ++passes;
        const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 1024), 148 * 2);
        csr_sort_long_rows_kernel<<<blocks, 1024, 0, stream>>>(rowptr, n, tmp_ws, perm, passes);
    }
    return check_launch("csr_fill");
}

int gmp_gather_i64_to_i32(const int64_t* src, const int32_t* perm, int64_t num_edges, int32_t* out, gmp_stream_t stream) {
    GMP_REQUIRE(num_edges == 0 || (src && out), "gather_i64_to_i32: bad arguments");
    if (num_edges == 0) return GMP_OK;
    gather_i64_i32_kernel<<<(unsigned)ceil_div(num_edges, 256), 256, 0, stream>>>(src, perm, num_edges, out);
    return check_launch("gather_i64_i32_kernel");
}

int gmp_mark_unique_pairs(const int64_t* row, const int64_t* col, const int32_t* perm, int64_t num_edges, int32_t* keep,
                          gmp_stream_t stream) {
    GMP_REQUIRE(num_edges == 0 || (row && col && keep), "mark_unique_pairs: bad arguments");
    if (num_edges == 0) return GMP_OK;
    mark_unique_pairs_kernel<<<(unsigned)ceil_div(num_edges, 256), 256, 0, stream>>>(row, col, perm, num_edges, keep);
    return check_launch("mark_unique_pairs_kernel");
}

int gmp_compact_pairs(const int64_t* row, const int64_t* col, const int32_t* perm, const int32_t* keep, const int64_t* pos,
                      int64_t num_edges, int64_t* out_row, int64_t* out_col, gmp_stream_t stream) {
    GMP_REQUIRE(num_edges == 0 || (row && col && keep && pos && out_row && out_col), "compact_pairs: bad arguments");
    if (num_edges == 0) return GMP_OK;
    compact_pairs_kernel<<<(unsigned)ceil_div(num_edges, 256), 256, 0, stream>>>(row, col, perm, keep, pos, num_edges, out_row, out_col);
    return check_launch("compact_pairs_kernel");
}

int gmp_index_is_sorted(const int64_t* index, int64_t num_edges, int32_t* flag, gmp_stream_t stream) {
    GMP_REQUIRE(flag, "index_is_sorted: flag is NULL");
    const int32_t one = 1;
    GMP_CUDA(cudaMemcpyAsync(flag, &one, sizeof(one), cudaMemcpyHostToDevice, stream));
    if (num_edges < 2) return GMP_OK;
    is_sorted_kernel<<<(unsigned)ceil_div(num_edges, 256), 256, 0, stream>>>(index, num_edges, flag);
    return check_launch("is_sorted_kernel");
}

}  // extern "C"
