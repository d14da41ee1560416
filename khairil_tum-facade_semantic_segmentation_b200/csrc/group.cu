// group.cu -- gathers of the set-abstraction path:
//   index_points                 /root/reference/models/pointnet2_utils.py:43-60
//   sample_and_group (gather)    :127-132   rows = [xyz[idx]-new_xyz | feats[idx]]
// and their backward (index_put_ accumulate == scatter-add), plus the layout helpers that
// turn channel-major API tensors into point-major rows.  All HBM/L2-bound element moves:
// consecutive threads walk consecutive channels of one row so global accesses coalesce.
#include "common.cuh"

namespace pn2 {

__global__ void index_points_kernel(const float *__restrict__ points, int64_t sB, int64_t sN, int64_t sC,
                                    int N, int C, const int64_t *__restrict__ idx, int64_t J,
                                    float *__restrict__ out, int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(e % C);
        int64_t r = e / C;          // (b, j)
        int64_t b = r / J;
        int64_t i = idx[r];
        out[e] = (i >= 0 && i < N) ? points[b * sB + i * sN + c * sC] : 0.0f;
    }
}

__global__ void index_points_bwd_kernel(const float *__restrict__ dout, const int64_t *__restrict__ idx,
                                        int N, int C, int64_t J, float *__restrict__ dpoints, int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(e % C);
        int64_t r = e / C;
        int64_t b = r / J;
        int64_t i = idx[r];
        if (i >= 0 && i < N) atomicAdd(dpoints + (b * N + i) * C + c, dout[e]);
    }
}

template <typename T>
__global__ void group_points_kernel(const float *__restrict__ xyz, int64_t sB, int64_t sN, int64_t sC,
                                    const float *__restrict__ new_xyz, const float *__restrict__ feats,
                                    int64_t fB, int64_t fN, int64_t fD, const int64_t *__restrict__ idx,
                                    int N, int S, int nsample, int D, T *__restrict__ rows, int ld,
                                    int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(e % ld);
        int64_t m = e / ld;                       // (b, s, k)
        int64_t g = m / nsample;                  // (b, s)
        int64_t b = g / S;
        int64_t i = idx[m];
        float v = 0.0f;
        if (i >= 0 && i < N) {
            if (c < 3) v = __fsub_rn(xyz[b * sB + i * sN + c * sC], new_xyz[g * 3 + c]);
            else if (c < 3 + D) v = feats[b * fB + i * fN + (c - 3) * fD];
        }
        st_act<T>(rows + e, v);
    }
}

// bf16 rows, ld % 8 == 0: one thread per 16-byte chunk of a row (8 columns)
__global__ void __launch_bounds__(256)
group_points_vec8_kernel(const float *__restrict__ xyz, int64_t sB, int64_t sN, int64_t sC,
                         const float *__restrict__ new_xyz, const float *__restrict__ feats, int64_t fB, int64_t fN,
                         int64_t fD, const int64_t *__restrict__ idx, int N, int S, int nsample, int D,
                         __nv_bfloat16 *__restrict__ rows, int ld, int64_t total_chunks) {
    const int cpr = ld >> 3;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_chunks;
         q += (int64_t)gridDim.x * blockDim.x) {
        int64_t m, g, b;
        int c8, k, sidx;
        fast_divmod(q, cpr, m, c8);
        fast_divmod(m, nsample, g, k);
        fast_divmod(g, S, b, sidx);
        const int c0 = c8 << 3;
        const int64_t i = idx[m];
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.0f;
        if (i >= 0 && i < N) {
            const float *px = xyz + b * sB + i * sN;
            const float *pf = feats + b * fB + i * fN;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = c0 + e;
                if (c < 3) v[e] = __fsub_rn(px[c * sC], new_xyz[g * 3 + c]);
                else if (c < 3 + D) v[e] = pf[(int64_t)(c - 3) * fD];
            }
        }
        uint4 o;
        __nv_bfloat162 h;
        h = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t *>(&h);
        h = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t *>(&h);
        h = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t *>(&h);
        h = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t *>(&h);
        *reinterpret_cast<uint4 *>(rows + m * ld + c0) = o;
    }
}

template <typename T>
__global__ void group_points_bwd_kernel(const T *__restrict__ drows, int ld, const int64_t *__restrict__ idx,
                                        int N, int S, int nsample, int D, float *__restrict__ dfeats,
                                        int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(e % D);
        int64_t m = e / D;
        int64_t b = m / ((int64_t)S * nsample);
        int64_t i = idx[m];
        if (i >= 0 && i < N) atomicAdd(dfeats + (b * N + i) * D + c, ld_act<T>(drows + m * ld + 3 + c));
    }
}

// D % 4 == 0: one thread adds 4 consecutive channels of one row with a single vector reduction (red.global.add.v4.f32,
// sm_90+): a quarter of the L2 atomic operations of the scalar kernel, and no 64-bit divisions in the loop
template <typename T>
__global__ void __launch_bounds__(256)
group_points_bwd_vec4_kernel(const T *__restrict__ drows, int ld, const int64_t *__restrict__ idx, int N, int S,
                             int nsample, int D, float *__restrict__ dfeats, int64_t total4) {
    const int qpr = D >> 2;
    const int per_cloud = S * nsample;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += (int64_t)gridDim.x * blockDim.x) {
        int64_t m, b;
        int c4, r;
        fast_divmod(q, qpr, m, c4);
        fast_divmod(m, per_cloud, b, r);
        const int64_t i = idx[m];
        if (i < 0 || i >= N) continue;
        const T *src = drows + m * ld + 3 + 4 * c4;
        const float v0 = ld_act<T>(src), v1 = ld_act<T>(src + 1), v2 = ld_act<T>(src + 2), v3 = ld_act<T>(src + 3);
        float *dst = dfeats + (b * N + i) * D + 4 * c4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
    }
}

// 32x32 tiled transpose: src[b, r, c] at src + b*sB + r*sR + c*sC (fast along r) -> dst rows
__global__ void to_rows_transpose_kernel(const float *__restrict__ src, int64_t sB, int64_t sR, int64_t sC,
                                         int64_t R, int C, float *__restrict__ dst, int64_t dB, int ldd) {
    __shared__ float t[32][33];
    const int b = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    const float *s = src + (int64_t)b * sB;
    float *d = dst + (int64_t)b * dB;
    for (int i = ty; i < 32; i += 8) {
        int64_t r = r0 + tx;
        int c = c0 + i;
        t[i][tx] = (r < R && c < C) ? s[r * sR + c * sC] : 0.0f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int64_t r = r0 + i;
        int c = c0 + tx;
        if (r < R && c < C) d[r * ldd + c] = t[tx][i];
    }
}

__global__ void to_rows_copy_kernel(const float *__restrict__ src, int64_t sB, int64_t sR, int64_t sC,
                                    int64_t R, int C, float *__restrict__ dst, int64_t dB, int ldd,
                                    int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(e % C);
        int64_t r = (e / C) % R;
        int64_t b = e / ((int64_t)C * R);
        dst[b * dB + r * ldd + c] = src[b * sB + r * sR + c * sC];
    }
}

template <typename T>
__global__ void rows_to_f32_kernel(const T *__restrict__ rows, int ld, int c0, int C, float *__restrict__ dst,
                                   int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(e % C);
        int64_t m = e / C;
        dst[e] = ld_act<T>(rows + m * ld + c0 + c);
    }
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_index_points(const float *points, int64_t sB, int64_t sN, int64_t sC, int B, int N,
                                int C, const int64_t *idx, int64_t J, float *out, void *stream) {
    PN2_REQUIRE(points && idx && out, "index_points: null pointer");
    PN2_REQUIRE(B >= 0 && N > 0 && C > 0 && J >= 0, "index_points: bad sizes");
    int64_t total = (int64_t)B * J * C;
    if (total == 0) return PN2_OK;
    index_points_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(points, sB, sN, sC, N, C, idx, J, out, total);
    count_launch();
    return check_launch("index_points");
}

extern "C" int pn2_index_points_bwd(const float *dout, const int64_t *idx, int B, int N, int C, int64_t J,
                                    float *dpoints, void *stream) {
    PN2_REQUIRE(dout && idx && dpoints, "index_points_bwd: null pointer");
    int64_t total = (int64_t)B * J * C;
    if (total == 0) return PN2_OK;
    index_points_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dout, idx, N, C, J, dpoints, total);
    count_launch();
    return check_launch("index_points_bwd");
}

extern "C" int pn2_group_points(const float *xyz, int64_t sB, int64_t sN, int64_t sC, const float *new_xyz,
                                const float *feats, int64_t fB, int64_t fN, int64_t fD, const int64_t *idx,
                                int B, int N, int S, int nsample, int D, void *rows, int ld, int dtype,
                                void *stream) {
    PN2_REQUIRE(xyz && new_xyz && idx && rows, "group_points: null pointer");
    PN2_REQUIRE(D == 0 || feats, "group_points: feats is NULL but D=%d", D);
    PN2_REQUIRE(ld >= 3 + D, "group_points: ld=%d < 3+D=%d", ld, 3 + D);
    PN2_REQUIRE(valid_dtype(dtype), "group_points: bad dtype %d", dtype);
    int64_t total = (int64_t)B * S * nsample * ld;
    if (total == 0) return PN2_OK;
    if (dtype == PN2_BF16 && ld % 8 == 0) {
        const int64_t chunks = total / 8;
        group_points_vec8_kernel<<<grid_for(chunks, 256), 256, 0, (cudaStream_t)stream>>>(
            xyz, sB, sN, sC, new_xyz, feats, fB, fN, fD, idx, N, S, nsample, D, (__nv_bfloat16 *)rows, ld, chunks);
        count_launch();
        return check_launch("group_points_vec8");
    }
    PN2_DISPATCH_DTYPE(dtype, T, (group_points_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        xyz, sB, sN, sC, new_xyz, feats, fB, fN, fD, idx, N, S, nsample, D, (T *)rows, ld, total)));
    count_launch();
    return check_launch("group_points");
}

extern "C" int pn2_group_points_bwd(const void *drows, int ld, int dtype, const int64_t *idx, int B, int N,
                                    int S, int nsample, int D, float *dfeats, void *stream) {
    PN2_REQUIRE(drows && idx && dfeats, "group_points_bwd: null pointer");
    PN2_REQUIRE(valid_dtype(dtype), "group_points_bwd: bad dtype %d", dtype);
    int64_t total = (int64_t)B * S * nsample * D;
    if (total == 0) return PN2_OK;
    if (D % 4 == 0 && ((uintptr_t)dfeats & 15) == 0) {
        PN2_DISPATCH_DTYPE(dtype, T, (group_points_bwd_vec4_kernel<T><<<grid_for(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(
            (const T *)drows, ld, idx, N, S, nsample, D, dfeats, total / 4)));
        count_launch();
        return check_launch("group_points_bwd_vec4");
    }
    PN2_DISPATCH_DTYPE(dtype, T, (group_points_bwd_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const T *)drows, ld, idx, N, S, nsample, D, dfeats, total)));
    count_launch();
    return check_launch("group_points_bwd");
}

extern "C" int pn2_to_rows(const float *src, int64_t sB, int64_t sR, int64_t sC, int B, int64_t R, int C,
                           float *dst, int64_t dB, int ldd, void *stream) {
    PN2_REQUIRE(src && dst, "to_rows: null pointer");
    PN2_REQUIRE(B >= 0 && R >= 0 && C > 0 && ldd >= C, "to_rows: bad sizes");
    if (B == 0 || R == 0) return PN2_OK;
    if (sC == 1 || C < 4) {
        int64_t total = (int64_t)B * R * C;
        to_rows_copy_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, sB, sR, sC, R, C, dst, dB, ldd, total);
    } else {
        PN2_REQUIRE(B <= 65535, "to_rows: B too large");
        dim3 grid((unsigned)((R + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B);
        to_rows_transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, sB, sR, sC, R, C, dst, dB, ldd);
    }
    count_launch();
    return check_launch("to_rows");
}

extern "C" int pn2_rows_to_f32(const void *rows, int ld, int dtype, int64_t M, int c0, int C, float *dst,
                               void *stream) {
    PN2_REQUIRE(rows && dst, "rows_to_f32: null pointer");
    PN2_REQUIRE(valid_dtype(dtype) && c0 >= 0 && c0 + C <= ld, "rows_to_f32: bad arguments");
    int64_t total = M * C;
    if (total == 0) return PN2_OK;
    PN2_DISPATCH_DTYPE(dtype, T, (rows_to_f32_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const T *)rows, ld, c0, C, dst, total)));
    count_launch();
    return check_launch("rows_to_f32");
}
