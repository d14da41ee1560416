// ballquery.cu -- K2: radius scan (query_ball_point, /root/reference/models/pointnet2_utils.py:87-107)
// and the a1 square_distance matrix for API parity (:19-40).
//
// The reference materialises a [B,S,N] int64 index tensor, masks it by the [B,S,N] distance
// matrix and SORTS it along N.  Here one warp owns Q queries and walks the cloud in index
// order: points are staged once per CTA in shared memory as float4 {x,y,z,|p|^2} (one 16-byte
// load per lane and step), every lane tests its point against the warp's Q queries held in
// registers, and ballot/popc prefix sums append the hits in ascending index order -- which is
// exactly "first nsample of the sorted in-radius indices".  A query stops scanning as soon as
// it has nsample hits; the slots it could not fill repeat its first hit (:104-106).
//
// Bit-exactness: the in-radius test is !(d > r2) on d = ((-2*mm)+|q|^2)+|p|^2 with
// mm = fma(qz,pz, fma(qy,py, qx*px)) -- the rounding order of the reference's matmul-based
// square_distance, NOT (dx^2+dy^2+dz^2).
#include "common.cuh"

namespace pn2 {

constexpr int kBqThreads = 256;
constexpr int kBqQ = 4;        // queries per warp
constexpr int kBqTile = 2048;  // points staged per pass (32 KB)

__global__ void __launch_bounds__(kBqThreads)
ball_query_kernel(const float *__restrict__ xyz, int64_t sB, int64_t sN, int64_t sC,
                  const float *__restrict__ new_xyz, int64_t qB, int64_t qN, int64_t qC, int N, int S,
                  float r2, int nsample, int64_t *__restrict__ out_idx, int32_t *__restrict__ out_cnt) {
    __shared__ float4 tile[kBqTile];
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = (blockIdx.x * (kBqThreads / 32) + warp) * kBqQ;
    const float *pts = xyz + (int64_t)b * sB;
    const float *qs = new_xyz + (int64_t)b * qB;

    float qx[kBqQ], qy[kBqQ], qz[kBqQ], qn[kBqQ];
    int cnt[kBqQ], first[kBqQ];
#pragma unroll
    for (int q = 0; q < kBqQ; ++q) {
        int s = q0 + q;
        if (s < S) {
            const float *p = qs + (int64_t)s * qN;
            qx[q] = p[0]; qy[q] = p[qC]; qz[q] = p[2 * qC];
            qn[q] = sq_norm3(qx[q], qy[q], qz[q]);
            cnt[q] = 0;
        } else {
            qx[q] = qy[q] = qz[q] = qn[q] = 0.0f;
            cnt[q] = nsample;   // nothing to do
        }
        first[q] = N;
    }
    const unsigned lt_mask = (1u << lane) - 1u;

    for (int t0 = 0; t0 < N; t0 += kBqTile) {
        const int tn = min(kBqTile, N - t0);
        __syncthreads();   // previous tile fully consumed
        for (int i = tid; i < tn; i += kBqThreads) {
            const float *p = pts + (int64_t)(t0 + i) * sN;
            float x = p[0], y = p[sC], z = p[2 * sC];
            tile[i] = make_float4(x, y, z, sq_norm3(x, y, z));
        }
        __syncthreads();
        bool warp_done = true;
#pragma unroll
        for (int q = 0; q < kBqQ; ++q) warp_done = warp_done && (cnt[q] >= nsample);
        if (!warp_done) {
            for (int j0 = 0; j0 < tn; j0 += 32) {
                const int j = j0 + lane;
                float4 p = tile[min(j, tn - 1)];
                bool all_done = true;
#pragma unroll
                for (int q = 0; q < kBqQ; ++q) {
                    if (cnt[q] < nsample) {   // warp-uniform
                        float d = expanded_sqdist(qx[q], qy[q], qz[q], qn[q], p.x, p.y, p.z, p.w);
                        bool hit = (j < tn) && !(d > r2);
                        unsigned m = __ballot_sync(0xffffffffu, hit);
                        if (m) {
                            if (cnt[q] == 0) first[q] = t0 + j0 + __ffs(m) - 1;
                            int pos = cnt[q] + __popc(m & lt_mask);
                            if (hit && pos < nsample)
                                out_idx[((int64_t)b * S + (q0 + q)) * nsample + pos] = (int64_t)(t0 + j);
                            cnt[q] += __popc(m);
                        }
                        all_done = all_done && (cnt[q] >= nsample);
                    }
                }
                if (all_done) break;
            }
        }
        // stop staging tiles once every query of the CTA is full
        bool done = true;
#pragma unroll
        for (int q = 0; q < kBqQ; ++q) done = done && (cnt[q] >= nsample);
        if (__syncthreads_and(done)) break;
    }

    // pad with the first hit (or N when the ball is empty, as the reference's sort leaves it)
#pragma unroll
    for (int q = 0; q < kBqQ; ++q) {
        int s = q0 + q;
        if (s >= S) continue;
        int c = min(cnt[q], nsample);
        int64_t *o = out_idx + ((int64_t)b * S + s) * nsample;
        for (int k = c + lane; k < nsample; k += 32) o[k] = (int64_t)first[q];
        if (out_cnt && lane == 0) out_cnt[(int64_t)b * S + s] = c;
    }
}

__global__ void square_distance_kernel(const float *__restrict__ src, const float *__restrict__ dst,
                                       int N, int M, float *__restrict__ out, int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int64_t j = e % M, i = (e / M) % N, b = e / ((int64_t)M * N);
        const float *s = src + (b * N + i) * 3;
        const float *d = dst + (b * M + j) * 3;
        float sx = s[0], sy = s[1], sz = s[2], dx = d[0], dy = d[1], dz = d[2];
        out[e] = expanded_sqdist(sx, sy, sz, sq_norm3(sx, sy, sz), dx, dy, dz, sq_norm3(dx, dy, dz));
    }
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_query_ball_point(const float *xyz, int64_t sB, int64_t sN, int64_t sC,
                                    const float *new_xyz, int64_t qB, int64_t qN, int64_t qC, int B,
                                    int N, int S, float r2, int nsample, int64_t *out_idx,
                                    int32_t *out_cnt, void *stream) {
    PN2_REQUIRE(xyz && new_xyz && out_idx, "ball query: null pointer");
    PN2_REQUIRE(B >= 0 && N > 0 && S >= 0 && nsample > 0, "ball query: bad sizes B=%d N=%d S=%d nsample=%d", B, N, S, nsample);
    PN2_REQUIRE(B <= 65535, "ball query: B=%d > 65535", B);
    if (B == 0 || S == 0) return PN2_OK;
    const int qpb = (kBqThreads / 32) * kBqQ;
    dim3 grid((S + qpb - 1) / qpb, B);
    ball_query_kernel<<<grid, kBqThreads, 0, (cudaStream_t)stream>>>(xyz, sB, sN, sC, new_xyz, qB, qN, qC, N,
                                                                      S, r2, nsample, out_idx, out_cnt);
    count_launch();
    return check_launch("ball_query");
}

extern "C" int pn2_square_distance(const float *src, const float *dst, int B, int N, int M, float *out,
                                   void *stream) {
    PN2_REQUIRE(src && dst && out, "square_distance: null pointer");
    PN2_REQUIRE(B >= 0 && N >= 0 && M >= 0, "square_distance: bad sizes");
    int64_t total = (int64_t)B * N * M;
    if (total == 0) return PN2_OK;
    square_distance_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, N, M, out, total);
    count_launch();
    return check_launch("square_distance");
}
