// threenn.cu -- K4: feature propagation head of PointNetFeaturePropagation.forward
// (/root/reference/models/pointnet2_utils.py:296-307).
//
// The reference builds the [B,N,S] distance matrix, fully SORTS it along S and keeps three
// columns.  Here one thread owns one fine point and keeps a 3-entry insertion list ordered by
// (distance, index) while the CTA streams the coarse cloud through shared memory as float4
// {x,y,z,|p|^2} (all lanes read the same address per step: a broadcast, no bank conflicts).
// A strict '<' on insertion keeps the lower index first among equal distances, i.e. the order
// of a stable ascending sort.  Distances use the reference's matmul-form rounding order
// (they may be slightly negative for coincident points; the weights reproduce that).
//
//   recip_k = 1/(d_k + 1e-8);  w_k = recip_k / ((recip_0 + recip_1) + recip_2)       (:300-302)
//   interpolated = (p2[i0]*w0 + p2[i1]*w1) + p2[i2]*w2, products rounded separately   (:303)
//   rows = [points1 | interpolated]                                                    (:307)
#include <math_constants.h>

#include "common.cuh"

namespace pn2 {

constexpr int kNnThreads = 128;
constexpr int kNnTile = 1024;

__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(const float *__restrict__ xyz1, int64_t aB, int64_t aN, int64_t aC,
                const float *__restrict__ xyz2, int64_t cB, int64_t cN, int64_t cC, int N, int S,
                int64_t *__restrict__ idx3, float *__restrict__ w3) {
    __shared__ float4 tile[kNnTile];
    const int b = blockIdx.y;
    const int n = blockIdx.x * kNnThreads + threadIdx.x;
    const bool active = n < N;
    float fx = 0.f, fy = 0.f, fz = 0.f, fn = 0.f;
    if (active) {
        const float *p = xyz1 + (int64_t)b * aB + (int64_t)n * aN;
        fx = p[0]; fy = p[aC]; fz = p[2 * aC];
        fn = sq_norm3(fx, fy, fz);
    }
    float d0 = CUDART_INF_F, d1 = CUDART_INF_F, d2 = CUDART_INF_F;
    int i0 = 0, i1 = 0, i2 = 0;
    const float *coarse = xyz2 + (int64_t)b * cB;
    for (int t0 = 0; t0 < S; t0 += kNnTile) {
        const int tn = min(kNnTile, S - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += kNnThreads) {
            const float *p = coarse + (int64_t)(t0 + i) * cN;
            float x = p[0], y = p[cC], z = p[2 * cC];
            tile[i] = make_float4(x, y, z, sq_norm3(x, y, z));
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int j = 0; j < tn; ++j) {
                float4 c = tile[j];
                float d = expanded_sqdist(fx, fy, fz, fn, c.x, c.y, c.z, c.w);
                if (d < d2) {
                    int jj = t0 + j;
                    if (d < d1) {
                        d2 = d1; i2 = i1;
                        if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = jj; }
                        else        { d1 = d;  i1 = jj; }
                    } else { d2 = d; i2 = jj; }
                }
            }
        }
    }
    if (!active) return;
    const int K3 = S < 3 ? S : 3;
    float r0 = __fdiv_rn(1.0f, __fadd_rn(d0, 1e-8f));
    float r1 = K3 > 1 ? __fdiv_rn(1.0f, __fadd_rn(d1, 1e-8f)) : 0.0f;
    float r2 = K3 > 2 ? __fdiv_rn(1.0f, __fadd_rn(d2, 1e-8f)) : 0.0f;
    float norm = r0;
    if (K3 > 1) norm = __fadd_rn(norm, r1);
    if (K3 > 2) norm = __fadd_rn(norm, r2);
    int64_t o = ((int64_t)b * N + n) * 3;
    idx3[o + 0] = i0;
    idx3[o + 1] = K3 > 1 ? i1 : 0;
    idx3[o + 2] = K3 > 2 ? i2 : 0;
    w3[o + 0] = __fdiv_rn(r0, norm);
    w3[o + 1] = K3 > 1 ? __fdiv_rn(r1, norm) : 0.0f;
    w3[o + 2] = K3 > 2 ? __fdiv_rn(r2, norm) : 0.0f;
}

// one warp per fine point, lanes over channels
template <typename T>
__global__ void __launch_bounds__(256)
interp_concat_kernel(const float *__restrict__ points1, int64_t pB, int64_t pN, int64_t pD,
                     const float *__restrict__ points2, int64_t qB, int64_t qN, int64_t qD,
                     const int64_t *__restrict__ idx3, const float *__restrict__ w3, int N, int S, int D1,
                     int D2, T *__restrict__ rows, int ld, int64_t M) {
    const int lane = threadIdx.x & 31;
    const int K3 = S < 3 ? S : 3;
    for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M;
         m += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t b = m / N, n = m % N;
        T *row = rows + m * ld;
        if (D1 > 0) {
            const float *p1 = points1 + b * pB + n * pN;
            for (int c = lane; c < D1; c += 32) st_act<T>(row + c, p1[c * pD]);
        }
        const int64_t j0 = idx3[m * 3 + 0], j1 = idx3[m * 3 + 1], j2 = idx3[m * 3 + 2];
        const float w0 = w3[m * 3 + 0], w1 = w3[m * 3 + 1], w2 = w3[m * 3 + 2];
        const float *a0 = points2 + b * qB + j0 * qN;
        const float *a1 = points2 + b * qB + j1 * qN;
        const float *a2 = points2 + b * qB + j2 * qN;
        for (int c = lane; c < D2; c += 32) {
            float v = __fmul_rn(a0[c * qD], w0);
            if (K3 > 1) v = __fadd_rn(v, __fmul_rn(a1[c * qD], w1));
            if (K3 > 2) v = __fadd_rn(v, __fmul_rn(a2[c * qD], w2));
            st_act<T>(row + D1 + c, v);
        }
        for (int c = D1 + D2 + lane; c < ld; c += 32) st_act<T>(row + c, 0.0f);
    }
}

// bf16 rows, unit channel strides, D1 % 4 == D2 % 4 == 0: a lane owns FOUR channels (float4 loads of the three neighbours,
// one 8-byte store of four bf16) -- a 128-wide level is one pass of the warp instead of four with 2-byte stores
__global__ void __launch_bounds__(256)
interp_concat_vec4_kernel(const float *__restrict__ points1, int64_t pB, int64_t pN, const float *__restrict__ points2,
                          int64_t qB, int64_t qN, const int64_t *__restrict__ idx3, const float *__restrict__ w3, int N, int S,
                          int D1, int D2, __nv_bfloat16 *__restrict__ rows, int ld, int64_t M) {
    const int lane = threadIdx.x & 31;
    const int K3 = S < 3 ? S : 3;
    for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M;
         m += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t b = m / N, n = m % N;
        __nv_bfloat16 *row = rows + m * ld;
        const int64_t j0 = idx3[m * 3 + 0], j1 = K3 > 1 ? idx3[m * 3 + 1] : 0, j2 = K3 > 2 ? idx3[m * 3 + 2] : 0;
        const float w0 = w3[m * 3 + 0], w1 = K3 > 1 ? w3[m * 3 + 1] : 0.0f, w2 = K3 > 2 ? w3[m * 3 + 2] : 0.0f;
        const float *a0 = points2 + b * qB + j0 * qN, *a1 = points2 + b * qB + j1 * qN, *a2 = points2 + b * qB + j2 * qN;
        for (int c = 4 * lane; c < D2; c += 128) {
            const float4 x0 = *reinterpret_cast<const float4 *>(a0 + c);
            float4 v = make_float4(__fmul_rn(x0.x, w0), __fmul_rn(x0.y, w0), __fmul_rn(x0.z, w0), __fmul_rn(x0.w, w0));
            if (K3 > 1) {
                const float4 x1 = *reinterpret_cast<const float4 *>(a1 + c);
                v.x = __fadd_rn(v.x, __fmul_rn(x1.x, w1)); v.y = __fadd_rn(v.y, __fmul_rn(x1.y, w1));
                v.z = __fadd_rn(v.z, __fmul_rn(x1.z, w1)); v.w = __fadd_rn(v.w, __fmul_rn(x1.w, w1));
            }
            if (K3 > 2) {
                const float4 x2 = *reinterpret_cast<const float4 *>(a2 + c);
                v.x = __fadd_rn(v.x, __fmul_rn(x2.x, w2)); v.y = __fadd_rn(v.y, __fmul_rn(x2.y, w2));
                v.z = __fadd_rn(v.z, __fmul_rn(x2.z, w2)); v.w = __fadd_rn(v.w, __fmul_rn(x2.w, w2));
            }
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t *>(&lo);
            o.y = *reinterpret_cast<uint32_t *>(&hi);
            *reinterpret_cast<uint2 *>(row + D1 + c) = o;
        }
        if (D1 > 0) {
            const float *p1 = points1 + b * pB + n * pN;
            for (int c = 4 * lane; c < D1; c += 128) {
                const float4 x = *reinterpret_cast<const float4 *>(p1 + c);
                __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
                uint2 o;
                o.x = *reinterpret_cast<uint32_t *>(&lo);
                o.y = *reinterpret_cast<uint32_t *>(&hi);
                *reinterpret_cast<uint2 *>(row + c) = o;
            }
        }
        for (int c = D1 + D2 + lane; c < ld; c += 32) row[c] = __float2bfloat16_rn(0.0f);
    }
}

// bf16 gradient rows, D2 % 4 == 0, 16-byte aligned targets: one red.global.add.v4.f32 per four channels
__global__ void __launch_bounds__(256)
interp_bwd_vec4_kernel(const __nv_bfloat16 *__restrict__ drows, int ld, const int64_t *__restrict__ idx3,
                       const float *__restrict__ w3, int N, int S, int D1, int D2, float *__restrict__ dpoints2, int64_t M) {
    const int lane = threadIdx.x & 31;
    const int K3 = S < 3 ? S : 3;
    for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M;
         m += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t b = m / N;
        const __nv_bfloat16 *row = drows + m * ld + D1;
        for (int c = 4 * lane; c < D2; c += 128) {
            const uint2 g = *reinterpret_cast<const uint2 *>(row + c);
            const float2 g01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&g.x));
            const float2 g23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&g.y));
            for (int k = 0; k < K3; ++k) {
                const float w = w3[m * 3 + k];
                float *dst = dpoints2 + (b * S + idx3[m * 3 + k]) * D2 + c;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(w * g01.x), "f"(w * g01.y), "f"(w * g23.x),
                             "f"(w * g23.y)
                             : "memory");
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
interp_bwd_kernel(const T *__restrict__ drows, int ld, const int64_t *__restrict__ idx3,
                  const float *__restrict__ w3, int N, int S, int D1, int D2, float *__restrict__ dpoints2,
                  int64_t M) {
    const int lane = threadIdx.x & 31;
    const int K3 = S < 3 ? S : 3;
    for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M;
         m += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t b = m / N;
        const T *row = drows + m * ld + D1;
        for (int k = 0; k < K3; ++k) {
            const float w = w3[m * 3 + k];
            float *dst = dpoints2 + (b * S + idx3[m * 3 + k]) * D2;
            for (int c = lane; c < D2; c += 32) atomicAdd(dst + c, w * ld_act<T>(row + c));
        }
    }
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_three_nn(const float *xyz1, int64_t aB, int64_t aN, int64_t aC, const float *xyz2,
                            int64_t cB, int64_t cN, int64_t cC, int B, int N, int S, int64_t *idx3,
                            float *w3, void *stream) {
    PN2_REQUIRE(xyz1 && xyz2 && idx3 && w3, "three_nn: null pointer");
    PN2_REQUIRE(B >= 0 && N >= 0 && S >= 1, "three_nn: bad sizes B=%d N=%d S=%d", B, N, S);
    PN2_REQUIRE(B <= 65535, "three_nn: B too large");
    if (B == 0 || N == 0) return PN2_OK;
    dim3 grid((N + kNnThreads - 1) / kNnThreads, B);
    three_nn_kernel<<<grid, kNnThreads, 0, (cudaStream_t)stream>>>(xyz1, aB, aN, aC, xyz2, cB, cN, cC, N, S, idx3, w3);
    count_launch();
    return check_launch("three_nn");
}

extern "C" int pn2_interp_concat(const float *points1, int64_t pB, int64_t pN, int64_t pD,
                                 const float *points2, int64_t qB, int64_t qN, int64_t qD,
                                 const int64_t *idx3, const float *w3, int B, int N, int S, int D1, int D2,
                                 void *rows, int ld, int dtype, void *stream) {
    PN2_REQUIRE(points2 && idx3 && w3 && rows, "interp_concat: null pointer");
    PN2_REQUIRE(D1 == 0 || points1, "interp_concat: points1 NULL but D1=%d", D1);
    PN2_REQUIRE(ld >= D1 + D2 && valid_dtype(dtype), "interp_concat: bad ld/dtype");
    int64_t M = (int64_t)B * N;
    if (M == 0) return PN2_OK;
    int grid = grid_for(M, 8);
    if (dtype == PN2_BF16 && qD == 1 && (D1 == 0 || pD == 1) && D1 % 4 == 0 && D2 % 4 == 0 && ld % 4 == 0 && qN % 4 == 0 && qB % 4 == 0 &&
        (D1 == 0 || (pN % 4 == 0 && pB % 4 == 0 && ((uintptr_t)points1 & 15) == 0)) && ((uintptr_t)points2 & 15) == 0 &&
        ((uintptr_t)rows & 7) == 0) {
        interp_concat_vec4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points1, pB, pN, points2, qB, qN, idx3, w3, N, S, D1, D2,
                                                                         (__nv_bfloat16 *)rows, ld, M);
        count_launch();
        return check_launch("interp_concat_vec4");
    }
    PN2_DISPATCH_DTYPE(dtype, T, (interp_concat_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
        points1, pB, pN, pD, points2, qB, qN, qD, idx3, w3, N, S, D1, D2, (T *)rows, ld, M)));
    count_launch();
    return check_launch("interp_concat");
}

extern "C" int pn2_interp_bwd(const void *drows, int ld, int dtype, const int64_t *idx3, const float *w3,
                              int B, int N, int S, int D1, int D2, float *dpoints2, void *stream) {
    PN2_REQUIRE(drows && idx3 && w3 && dpoints2, "interp_bwd: null pointer");
    PN2_REQUIRE(ld >= D1 + D2 && valid_dtype(dtype), "interp_bwd: bad ld/dtype");
    int64_t M = (int64_t)B * N;
    if (M == 0 || D2 == 0) return PN2_OK;
    int grid = grid_for(M, 8);
    if (dtype == PN2_BF16 && D1 % 4 == 0 && D2 % 4 == 0 && ld % 4 == 0 && ((uintptr_t)drows & 7) == 0 && ((uintptr_t)dpoints2 & 15) == 0) {
        interp_bwd_vec4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)drows, ld, idx3, w3, N, S, D1, D2, dpoints2, M);
        count_launch();
        return check_launch("interp_bwd_vec4");
    }
    PN2_DISPATCH_DTYPE(dtype, T, (interp_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const T *)drows, ld, idx3, w3, N, S, D1, D2, dpoints2, M)));
    count_launch();
    return check_launch("interp_bwd");
}
