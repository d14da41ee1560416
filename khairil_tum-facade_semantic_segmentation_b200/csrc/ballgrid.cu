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
        const double rr = (double)radius * (double)radius;
        const bool margin_ok = !any_bad && radius > 0.0f && ((double)r2 + ldexp(R2, -18) <= rr * 1.010025) && ((double)r2 <= rr * 1.0001);
        double cell = 1.01 * (double)radius;
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
