// ballgrid.cu -- K2 through a uniform cell grid: query_ball_point (/root/reference/models/pointnet2_utils.py:87-107)
// with the SAME result as the index-order radius scan of ballquery.cu (and hence as the reference's mask + sort), at
// ~100 distance evaluations per query instead of N.
//
//   build : one CTA per cloud.  Bounding box -> cells of edge h >= 1.01 r -> counting sort of the cloud by cell
//           (histogram and scan in shared memory); the sorted copy holds {x, y, z, |p|^2} + the original index.
//   query : one warp per query.  The 3 x 3 x 3 cells around the query are 9 contiguous runs of the sorted cloud (cells are
//           numbered x-fastest); lanes evaluate the reference's fp32 expanded-form distance for the candidates and set the
//           bits of the accepted ones in an N-bit map in shared memory; a popc prefix scan of the map then emits the first
//           nsample set bits -- "the first nsample in-radius indices in index order, padded with the first" (:102-106),
//           whatever order the cells were visited in.
//
// Why the 27 cells are enough, bit for bit: a point passes the reference's test !(d > r^2) with
// d = ((-2 mm) + |q|^2) + |p|^2 evaluated in fp32, whose absolute rounding error is below 2^-18 R^2
// (R^2 = the largest |p|^2 of the cloud, a bound with a factor ~8 to spare), so every accepted point has TRUE distance
// <= sqrt(r^2 + 2^-18 R^2).  The build kernel checks r^2 + 2^-18 R^2 <= (1.005 r)^2 from the bounding box; then any accepted
// point is within 1.005 r < h (1 - 4e-3) per axis and -- the cell coordinate floor((x - min) / h) being monotone in x --
// at most one cell away.  Clouds that fail the check (huge un-centred coordinates, where the reference's own test is
// mostly rounding noise) or contain NaN (which the reference counts as inside) take the index-order scan instead, inside
// the same launch.
#include <math_constants.h>

#include "common.cuh"

namespace pn2 {

constexpr int kBgBuildThreads = 1024;
constexpr int kBgMaxCells = 10240;      // histogram + scan in shared memory (40 KB)
constexpr int kBgQueryWarps = 8;

struct BgHeader {          // per cloud, at the start of its workspace slice
    float minx, miny, minz, inv_h;
    int gx, gy, gz, use_grid;
};

__host__ __device__ inline size_t bg_cloud_bytes(int N) {
    // header (32 B) | cell_start [kBgMaxCells + 1] int32 | sorted float4 [N] | sorted index [N] int32, 16-byte aligned parts
    size_t o = 32;
    o += ((size_t)(kBgMaxCells + 1) * 4 + 15) & ~(size_t)15;
    o += (size_t)N * 16;
    o += ((size_t)N * 4 + 15) & ~(size_t)15;
    return o;
}

__device__ __forceinline__ int bg_cell1(float x, float lo, float inv_h, int g) {
    const int c = (int)floorf((x - lo) * inv_h);
    return min(g - 1, max(0, c));
}

__global__ void __launch_bounds__(kBgBuildThreads)
bg_build_kernel(const float *__restrict__ xyz, int64_t sB, int64_t sN, int64_t sC, int N, float radius, float r2,
                uint8_t *__restrict__ ws) {
    __shared__ int hist[kBgMaxCells];
    __shared__ float red[6][32];
    __shared__ int s_flag[32];
    __shared__ int warp_tot[32];
    __shared__ BgHeader hd;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *pts = xyz + (int64_t)b * sB;
    uint8_t *base = ws + (size_t)b * bg_cloud_bytes(N);
    BgHeader *g_hd = reinterpret_cast<BgHeader *>(base);
    int *cell_start = reinterpret_cast<int *>(base + 32);
    float4 *sorted = reinterpret_cast<float4 *>(base + 32 + (((size_t)(kBgMaxCells + 1) * 4 + 15) & ~(size_t)15));
    int *sidx = reinterpret_cast<int *>(reinterpret_cast<uint8_t *>(sorted) + (size_t)N * 16);

    // ---- bounding box, NaN / Inf check ----
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    int bad = 0;
    for (int i = tid; i < N; i += kBgBuildThreads) {
        const float *p = pts + (int64_t)i * sN;
        const float v[3] = {p[0], p[sC], p[2 * sC]};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (!(fabsf(v[d]) <= 3.0e38f)) bad = 1;          // NaN or Inf
            lo[d] = fminf(lo[d], v[d]);
            hi[d] = fmaxf(hi[d], v[d]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { red[d][warp] = lo[d]; red[3 + d][warp] = hi[d]; }
        s_flag[warp] = bad;
    }
    __syncthreads();
    if (tid == 0) {
        float l[3], h[3];
        int any_bad = 0;
        for (int d = 0; d < 3; ++d) { l[d] = 3.0e38f; h[d] = -3.0e38f; }
        for (int w = 0; w < kBgBuildThreads / 32; ++w) {
            for (int d = 0; d < 3; ++d) { l[d] = fminf(l[d], red[d][w]); h[d] = fmaxf(h[d], red[3 + d][w]); }
            any_bad |= s_flag[w];
        }
        // R^2 bound from the box, the rounding-error budget of the fp32 expanded-form distance against the cell margin
        double R2 = 0.0;
        for (int d = 0; d < 3; ++d) { const double m = fmax(fabs((double)l[d]), fabs((double)h[d])); R2 += m * m; }
        // radius <= 0 (3-NN search, no radius given): cells of about two point spacings, the spacing estimated from the box both
        // as a volume and as a surface (facade clouds are surfaces inside a tall box) -- any cell size is correct (the search
        // falls back to the full scan for a query whose third neighbour is not provably inside its 27 cells), this only sets
        // how many candidates a query looks at
        double rad = (double)radius, r2d = (double)r2;
        if (!(radius > 0.0f) && !any_bad && N > 0) {
            double e[3];
            for (int d = 0; d < 3; ++d) e[d] = (double)h[d] - (double)l[d];
            const double vol = e[0] * e[1] * e[2];
            const double area = fmax(e[0] * e[1], fmax(e[1] * e[2], e[0] * e[2]));
            rad = fmax(2.0 * cbrt(vol / (double)N), 2.5 * sqrt(area / (double)N)) / 1.01;
            r2d = rad * rad;
        }
        const double rr = rad * rad;
        const bool margin_ok = !any_bad && rad > 0.0 && (r2d + ldexp(R2, -18) <= rr * 1.010025) && (r2d <= rr * 1.0001);
        double cell = 1.01 * rad;
        int gx = 1, gy = 1, gz = 1;
        if (margin_ok) {
            for (int it = 0; it < 64; ++it) {
                const double ex = ((double)h[0] - l[0]) / cell, ey = ((double)h[1] - l[1]) / cell, ez = ((double)h[2] - l[2]) / cell;
                if (ex < 4096.0 && ey < 4096.0 && ez < 4096.0) {
                    gx = (int)ex + 1; gy = (int)ey + 1; gz = (int)ez + 1;
                    if ((long long)gx * gy * gz <= kBgMaxCells) break;
                }
                cell *= 1.5;        // coarser cells are still correct (the 27-neighbourhood only grows)
                gx = gy = gz = 1;
            }
        }
        hd.minx = l[0]; hd.miny = l[1]; hd.minz = l[2];
        hd.inv_h = (float)(1.0 / cell);
        hd.gx = gx; hd.gy = gy; hd.gz = gz;
        hd.use_grid = (margin_ok && (long long)gx * gy * gz <= kBgMaxCells) ? 1 : 0;
        *g_hd = hd;
    }
    __syncthreads();
    if (!hd.use_grid) return;
    const int ncell = hd.gx * hd.gy * hd.gz;
    for (int c = tid; c < ncell; c += kBgBuildThreads) hist[c] = 0;
    __syncthreads();
    // ---- histogram ----
    for (int i = tid; i < N; i += kBgBuildThreads) {
        const float *p = pts + (int64_t)i * sN;
        const int c = (bg_cell1(p[2 * sC], hd.minz, hd.inv_h, hd.gz) * hd.gy + bg_cell1(p[sC], hd.miny, hd.inv_h, hd.gy)) * hd.gx +
                      bg_cell1(p[0], hd.minx, hd.inv_h, hd.gx);
        atomicAdd(&hist[c], 1);
    }
    __syncthreads();
    // ---- exclusive scan over ncell entries: each thread owns a contiguous run ----
    const int per = (ncell + kBgBuildThreads - 1) / kBgBuildThreads;
    const int c0 = tid * per, c1 = min(ncell, c0 + per);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += hist[c];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = warp_tot[lane], inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        warp_tot[lane] = inc - v;        // exclusive prefix of the warp totals
    }
    __syncthreads();
    int run = warp_tot[warp] + incl - sum;
    for (int c = c0; c < c1; ++c) {
        const int n = hist[c];
        hist[c] = run;                    // becomes the scatter cursor
        cell_start[c] = run;
        run += n;
    }
    if (tid == 0) cell_start[ncell] = N;
    __syncthreads();
    // ---- scatter (order inside a cell is irrelevant: the query selects through a bitmap) ----
    for (int i = tid; i < N; i += kBgBuildThreads) {
        const float *p = pts + (int64_t)i * sN;
        const float x = p[0], y = p[sC], z = p[2 * sC];
        const int c = (bg_cell1(z, hd.minz, hd.inv_h, hd.gz) * hd.gy + bg_cell1(y, hd.miny, hd.inv_h, hd.gy)) * hd.gx +
                      bg_cell1(x, hd.minx, hd.inv_h, hd.gx);
        const int pos = atomicAdd(&hist[c], 1);
        sorted[pos] = make_float4(x, y, z, sq_norm3(x, y, z));
        sidx[pos] = i;
    }
}

__global__ void __launch_bounds__(kBgQueryWarps * 32)
bg_query_kernel(const float *__restrict__ xyz, int64_t sB, int64_t sN, int64_t sC, const float *__restrict__ new_xyz,
                int64_t qB, int64_t qN, int64_t qC, int N, int S, float r2, int nsample, const uint8_t *__restrict__ ws,
                int64_t *__restrict__ out_idx, int32_t *__restrict__ out_cnt) {
    extern __shared__ uint32_t bitmaps[];          // [warps][words]
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * kBgQueryWarps + warp;
    if (s >= S) return;
    const int words = (N + 31) >> 5;
    uint32_t *bm = bitmaps + (size_t)warp * words;
    const uint8_t *base = ws + (size_t)b * bg_cloud_bytes(N);
    const BgHeader hd = *reinterpret_cast<const BgHeader *>(base);
    const int *cell_start = reinterpret_cast<const int *>(base + 32);
    const float4 *sorted = reinterpret_cast<const float4 *>(base + 32 + (((size_t)(kBgMaxCells + 1) * 4 + 15) & ~(size_t)15));
    const int *sidx = reinterpret_cast<const int *>(reinterpret_cast<const uint8_t *>(sorted) + (size_t)N * 16);
    const float *qp = new_xyz + (int64_t)b * qB + (int64_t)s * qN;
    const float qx = qp[0], qy = qp[qC], qz = qp[2 * qC];
    const float qn = sq_norm3(qx, qy, qz);
    int64_t *o = out_idx + ((int64_t)b * S + s) * nsample;
    const unsigned lt_mask = (1u << lane) - 1u;

    if (!hd.use_grid) {
        // ---- the index-order scan of ballquery.cu for this query (rare clouds: see the header) ----
        const float *pts = xyz + (int64_t)b * sB;
        int cnt = 0, first = N;
        for (int j0 = 0; j0 < N && cnt < nsample; j0 += 32) {
            const int j = j0 + lane;
            bool hit = false;
            if (j < N) {
                const float *p = pts + (int64_t)j * sN;
                const float x = p[0], y = p[sC], z = p[2 * sC];
                hit = !(expanded_sqdist(qx, qy, qz, qn, x, y, z, sq_norm3(x, y, z)) > r2);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                if (cnt == 0) first = j0 + __ffs(m) - 1;
                const int pos = cnt + __popc(m & lt_mask);
                if (hit && pos < nsample) o[pos] = (int64_t)j;
                cnt += __popc(m);
            }
        }
        const int c = min(cnt, nsample);
        for (int k = c + lane; k < nsample; k += 32) o[k] = (int64_t)first;
        if (out_cnt && lane == 0) out_cnt[(int64_t)b * S + s] = c;
        return;
    }

    for (int w = lane; w < words; w += 32) bm[w] = 0u;
    __syncwarp();
    const int cx = bg_cell1(qx, hd.minx, hd.inv_h, hd.gx), cy = bg_cell1(qy, hd.miny, hd.inv_h, hd.gy),
              cz = bg_cell1(qz, hd.minz, hd.inv_h, hd.gz);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, hd.gx - 1);
    // the 9 (dz, dy) rows of the neighbourhood are 9 contiguous runs of the sorted cloud: lanes 0-8 fetch their run's
    // bounds at once, the runs are concatenated by a prefix sum, and lanes stride over the union -- three dependent
    // global-load latencies per query instead of three per run
    int beg = 0, len = 0;
    if (lane < 9) {
        const int z = cz + lane / 3 - 1, y = cy + lane % 3 - 1;
        if (z >= 0 && z < hd.gz && y >= 0 && y < hd.gy) {
            const int row = (z * hd.gy + y) * hd.gx;
            beg = cell_start[row + x0];
            len = cell_start[row + x1 + 1] - beg;
        }
    }
    int incl_len = len;
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl_len, off);
        if (lane >= off) incl_len += v;
    }
    const int tot = __shfl_sync(0xffffffffu, incl_len, 8);
    const int excl_len = incl_len - len;
    for (int base_i = 0; base_i < tot; base_i += 32) {
        const int i = base_i + lane;
        int src = -1;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int e = __shfl_sync(0xffffffffu, excl_len, r), b0 = __shfl_sync(0xffffffffu, beg, r), l = __shfl_sync(0xffffffffu, len, r);
            if (i >= e && i < e + l) src = b0 + (i - e);
        }
        if (src >= 0) {
            const float4 p = sorted[src];
            const int idx = sidx[src];
            if (!(expanded_sqdist(qx, qy, qz, qn, p.x, p.y, p.z, p.w) > r2)) atomicOr(&bm[idx >> 5], 1u << (idx & 31));
        }
    }
    __syncwarp();
    // ---- first nsample set bits in ascending order: each lane owns a contiguous run of words ----
    const int wpl = (words + 31) >> 5;
    const int w0 = lane * wpl, w1 = min(words, w0 + wpl);
    int mine = 0;
    for (int w = w0; w < w1; ++w) mine += __popc(bm[w]);
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int rank = incl - mine;
    int first_local = N;
    for (int w = w0; w < w1 && rank < nsample; ++w) {
        uint32_t m = bm[w];
        while (m && rank < nsample) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int idx = (w << 5) + bit;
            if (first_local == N) first_local = idx;
            o[rank++] = (int64_t)idx;
        }
    }
    // the first hit = the lowest set bit: the first lane that holds any
    const unsigned have = __ballot_sync(0xffffffffu, mine > 0);
    int first = N;
    if (have) {
        const int src = __ffs(have) - 1;
        // (that lane's first_local is set: its rank starts at 0 < nsample)
        first = __shfl_sync(0xffffffffu, first_local, src);
    }
    const int c = min(total, nsample);
    __syncwarp();
    for (int k = c + lane; k < nsample; k += 32) o[k] = (int64_t)first;
    if (out_cnt && lane == 0) out_cnt[(int64_t)b * S + s] = c;
}

// ---- 3-NN through the same grid (pointnet2_utils.py:296-302) -----------------------------------------------------
// One thread per fine point.  The coarse cloud is the gridded one; the candidates are the points of the 27 cells around
// the query's (clamped) cell, the list keeps the three smallest (distance, index) pairs -- what three_nn_kernel's
// index-order scan with its strict '<' keeps, whatever order the candidates come in.  The list is final when no point
// OUTSIDE the 27 cells can enter it: such a point is at least D away from the query along one axis (D = the query's
// distance to the nearest face of its block that has cells beyond it), its fp32 expanded-form distance is therefore
// >= D^2 - err with err < 2^-18 R^2 (R^2 bounds |q|^2 and every |p|^2), and the test is
// d_third < (0.995 D)^2 - 2^-17 R^2 (the 0.995 absorbs the fp32 rounding of the face coordinates: R / h < 60 by the build
// kernel's margin check).  Queries that fail it (third neighbour farther than the block, fewer than three candidates,
// NaN) and clouds without a grid take the full index-order scan -- same arithmetic, same result as three_nn_kernel.
constexpr int kNn3Threads = 128;

__device__ __forceinline__ void nn3_insert(float d, int j, float &d0, float &d1, float &d2, int &i0, int &i1, int &i2) {
    if (d < d2 || (d == d2 && j < i2)) {
        if (d < d1 || (d == d1 && j < i1)) {
            d2 = d1; i2 = i1;
            if (d < d0 || (d == d0 && j < i0)) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
            else { d1 = d; i1 = j; }
        } else { d2 = d; i2 = j; }
    }
}

__global__ void __launch_bounds__(kNn3Threads)
nn3_grid_kernel(const float *__restrict__ xyz1, int64_t aB, int64_t aN, int64_t aC, const float *__restrict__ xyz2, int64_t cB,
                int64_t cN, int64_t cC, int N, int S, const uint8_t *__restrict__ ws, int64_t *__restrict__ idx3,
                float *__restrict__ w3, unsigned *__restrict__ fallbacks) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * kNn3Threads + threadIdx.x;
    if (n >= N) return;
    const uint8_t *base = ws + (size_t)b * bg_cloud_bytes(S);
    const BgHeader hd = *reinterpret_cast<const BgHeader *>(base);
    const int *cell_start = reinterpret_cast<const int *>(base + 32);
    const float4 *sorted = reinterpret_cast<const float4 *>(base + 32 + (((size_t)(kBgMaxCells + 1) * 4 + 15) & ~(size_t)15));
    const int *sidx = reinterpret_cast<const int *>(reinterpret_cast<const uint8_t *>(sorted) + (size_t)S * 16);
    const float *qp = xyz1 + (int64_t)b * aB + (int64_t)n * aN;
    const float qx = qp[0], qy = qp[aC], qz = qp[2 * aC];
    const float qn = sq_norm3(qx, qy, qz);
    float d0 = CUDART_INF_F, d1 = CUDART_INF_F, d2 = CUDART_INF_F;
    int i0 = 0, i1 = 0, i2 = 0;
    bool final_list = false;
    if (hd.use_grid) {
        const float h = 1.0f / hd.inv_h;
        const int cx = bg_cell1(qx, hd.minx, hd.inv_h, hd.gx), cy = bg_cell1(qy, hd.miny, hd.inv_h, hd.gy),
                  cz = bg_cell1(qz, hd.minz, hd.inv_h, hd.gz);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, hd.gx - 1);
        for (int dz = -1; dz <= 1; ++dz) {
            const int z = cz + dz;
            if (z < 0 || z >= hd.gz) continue;
            for (int dy = -1; dy <= 1; ++dy) {
                const int y = cy + dy;
                if (y < 0 || y >= hd.gy) continue;
                const int row = (z * hd.gy + y) * hd.gx;
                const int beg = cell_start[row + x0], end = cell_start[row + x1 + 1];
                for (int src = beg; src < end; ++src) {
                    const float4 p = sorted[src];
                    nn3_insert(expanded_sqdist(qx, qy, qz, qn, p.x, p.y, p.z, p.w), sidx[src], d0, d1, d2, i0, i1, i2);
                }
            }
        }
        // distance to the nearest face of the block with cells beyond it, per axis (see the header comment)
        float D = CUDART_INF_F;
        const float q[3] = {qx, qy, qz}, lo[3] = {hd.minx, hd.miny, hd.minz};
        const int c[3] = {cx, cy, cz}, g[3] = {hd.gx, hd.gy, hd.gz};
        float R2 = qn;
        float m2 = 0.0f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (c[a] >= 2) D = fminf(D, q[a] - (lo[a] + (float)(c[a] - 1) * h));
            if (c[a] <= g[a] - 3) D = fminf(D, (lo[a] + (float)(c[a] + 2) * h) - q[a]);
            const float m = fmaxf(fabsf(lo[a]), fabsf(lo[a] + (float)g[a] * h));
            m2 += m * m;
        }
        R2 = fmaxf(R2, m2 * 1.0001f);
        const float lim = D == CUDART_INF_F ? CUDART_INF_F : (0.995f * D) * (0.995f * D) - R2 * 7.62939453125e-6f;   // 2^-17
        final_list = D > 0.0f && d2 < lim;       // (d2 == inf or NaN, a negative D: not final)
    }
    if (!final_list) {
        if (fallbacks) atomicAdd(fallbacks, 1u);
        d0 = d1 = d2 = CUDART_INF_F;
        i0 = i1 = i2 = 0;
        const float *coarse = xyz2 + (int64_t)b * cB;
        for (int j = 0; j < S; ++j) {
            const float *p = coarse + (int64_t)j * cN;
            const float x = p[0], y = p[cC], z = p[2 * cC];
            const float d = expanded_sqdist(qx, qy, qz, qn, x, y, z, sq_norm3(x, y, z));
            if (d < d2) {
                if (d < d1) {
                    d2 = d1; i2 = i1;
                    if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
                    else { d1 = d; i1 = j; }
                } else { d2 = d; i2 = j; }
            }
        }
    }
    const int K3 = S < 3 ? S : 3;
    const float r0 = __fdiv_rn(1.0f, __fadd_rn(d0, 1e-8f));
    const float r1 = K3 > 1 ? __fdiv_rn(1.0f, __fadd_rn(d1, 1e-8f)) : 0.0f;
    const float r2 = K3 > 2 ? __fdiv_rn(1.0f, __fadd_rn(d2, 1e-8f)) : 0.0f;
    float norm = r0;
    if (K3 > 1) norm = __fadd_rn(norm, r1);
    if (K3 > 2) norm = __fadd_rn(norm, r2);
    const int64_t o = ((int64_t)b * N + n) * 3;
    idx3[o + 0] = i0;
    idx3[o + 1] = K3 > 1 ? i1 : 0;
    idx3[o + 2] = K3 > 2 ? i2 : 0;
    w3[o + 0] = __fdiv_rn(r0, norm);
    w3[o + 1] = K3 > 1 ? __fdiv_rn(r1, norm) : 0.0f;
    w3[o + 2] = K3 > 2 ? __fdiv_rn(r2, norm) : 0.0f;
}

}  // namespace pn2

using namespace pn2;

extern "C" size_t pn2_ball_grid_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return (size_t)B * bg_cloud_bytes(N);
}

extern "C" int pn2_query_ball_point_grid(const float *xyz, int64_t sB, int64_t sN, int64_t sC, const float *new_xyz,
                                         int64_t qB, int64_t qN, int64_t qC, int B, int N, int S, float radius, float r2,
                                         int nsample, int64_t *out_idx, int32_t *out_cnt, void *workspace, size_t workspace_bytes,
                                         void *stream) {
    PN2_REQUIRE(xyz && new_xyz && out_idx && workspace, "ball query (grid): null pointer");
    PN2_REQUIRE(B >= 0 && N > 0 && S >= 0 && nsample > 0, "ball query (grid): bad sizes B=%d N=%d S=%d nsample=%d", B, N, S, nsample);
    PN2_REQUIRE(B <= 65535, "ball query (grid): B=%d > 65535", B);
    PN2_REQUIRE(workspace_bytes >= pn2_ball_grid_workspace_bytes(B, N), "ball query (grid): workspace too small");
    PN2_REQUIRE(((uintptr_t)workspace & 15) == 0, "ball query (grid): workspace must be 16-byte aligned");
    if (B == 0 || S == 0) return PN2_OK;
    const size_t smem = (size_t)kBgQueryWarps * (size_t)((N + 31) / 32) * 4;
    PN2_REQUIRE(smem <= 200 * 1024, "ball query (grid): N=%d too large for the per-query bitmaps (use pn2_query_ball_point)", N);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(bg_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            set_error("ball query (grid): shared-memory opt-in failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return PN2_ERR_CUDA;
        }
        attr_done = true;
    }
    bg_build_kernel<<<B, kBgBuildThreads, 0, (cudaStream_t)stream>>>(xyz, sB, sN, sC, N, radius, r2, (uint8_t *)workspace);
    count_launch();
    int rc = check_launch("ball_grid_build");
    if (rc != PN2_OK) return rc;
    dim3 grid((S + kBgQueryWarps - 1) / kBgQueryWarps, B);
    bg_query_kernel<<<grid, kBgQueryWarps * 32, smem, (cudaStream_t)stream>>>(xyz, sB, sN, sC, new_xyz, qB, qN, qC, N, S, r2, nsample,
                                                                              (const uint8_t *)workspace, out_idx, out_cnt);
    count_launch();
    return check_launch("ball_grid_query");
}

extern "C" int pn2_three_nn_grid(const float *xyz1, int64_t aB, int64_t aN, int64_t aC, const float *xyz2, int64_t cB, int64_t cN,
                                 int64_t cC, int B, int N, int S, int64_t *idx3, float *w3, void *workspace, size_t workspace_bytes,
                                 unsigned *fallback_count, void *stream) {
    PN2_REQUIRE(xyz1 && xyz2 && idx3 && w3 && workspace, "three_nn (grid): null pointer");
    PN2_REQUIRE(B >= 0 && N >= 0 && S >= 1, "three_nn (grid): bad sizes B=%d N=%d S=%d", B, N, S);
    PN2_REQUIRE(B <= 65535, "three_nn (grid): B too large");
    PN2_REQUIRE(workspace_bytes >= pn2_ball_grid_workspace_bytes(B, S), "three_nn (grid): workspace too small");
    PN2_REQUIRE(((uintptr_t)workspace & 15) == 0, "three_nn (grid): workspace must be 16-byte aligned");
    if (B == 0 || N == 0) return PN2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    bg_build_kernel<<<B, kBgBuildThreads, 0, st>>>(xyz2, cB, cN, cC, S, 0.0f, 0.0f, (uint8_t *)workspace);      // radius 0: cells from the density
    count_launch();
    int rc = check_launch("three_nn (grid) build");
    if (rc != PN2_OK) return rc;
    dim3 grid((N + kNn3Threads - 1) / kNn3Threads, B);
    nn3_grid_kernel<<<grid, kNn3Threads, 0, st>>>(xyz1, aB, aN, aC, xyz2, cB, cN, cC, N, S, (const uint8_t *)workspace, idx3, w3,
                                                  fallback_count);
    count_launch();
    return check_launch("three_nn (grid)");
}
