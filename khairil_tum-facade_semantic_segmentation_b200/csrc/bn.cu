// bn.cu -- BatchNorm (+ReLU, + max over nsample) around the MLP layers, forward and backward.
// Reference: bn(conv(x)) -> relu -> torch.max(.., 2) in PointNetSetAbstraction.forward
// (/root/reference/models/pointnet2_utils.py:196-200) and relu(bn(conv(x))) in
// PointNetFeaturePropagation.forward (:311-314); modules are nn.BatchNorm2d/1d, i.e. batch
// statistics over every row in train mode (padding duplicates included) with biased variance
// for the normalisation and unbiased variance for running_var.
//
// All of these are streaming (HBM-bound) passes over [M, C] rows; threads walk channels so
// accesses coalesce; every cross-thread reduction has a fixed order (deterministic).
#include "common.cuh"

namespace pn2 {


int linear_num_partials(int64_t M);

// "Last block finalizes": every block has added its partial sums into the fp64 accumulator (L2 atomics); the block
// that draws the last ticket turns the accumulator into dbeta / dgamma, zeroes it and resets the ticket, which saves
// the separate finalize launch (and its dependency gap) behind every reduction.  ticket == nullptr: plain reduction.
__device__ __forceinline__ void bn_bwd_last_block_finalize(double *accum, int C, unsigned *ticket, float *dgamma,
                                                           float *dbeta) {
    __shared__ int s_last;
    if (!ticket) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int r = 0; r < kStatReplicas; ++r) {   // fixed order
            double *a = accum + (size_t)r * 2 * C;
            s1 += __ldcg(a + c);
            s2 += __ldcg(a + C + c);
            a[c] = 0.0;
            a[C + c] = 0.0;
        }
        dbeta[c] = (float)s1;
        dgamma[c] = (float)s2;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

// ------------------------------------------------------------------ statistics -> scale/shift
__global__ void bn_train_finalize_kernel(double *__restrict__ accum, int64_t M, int N,
                                         const float *__restrict__ gamma, const float *__restrict__ beta,
                                         const float *__restrict__ conv_bias, float eps, float momentum,
                                         float *__restrict__ running_mean, float *__restrict__ running_var,
                                         float *__restrict__ scale, float *__restrict__ shift,
                                         float *__restrict__ save_mean, float *__restrict__ save_invstd,
                                         long long *__restrict__ num_batches_tracked) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
    if (c >= N) return;
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int r = 0; r < kStatReplicas; ++r) {   // fixed order; self-cleaning: ready for the next producer
        double *a = accum + (size_t)r * 2 * N;
        s1 += a[c];
        s2 += a[N + c];
        a[c] = 0.0;
        a[N + c] = 0.0;
    }
    bn_finalize_channel(s1, s2, M, c, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, scale, shift,
                        save_mean, save_invstd);
}

__global__ void bn_eval_fold_kernel(const float *__restrict__ gamma, const float *__restrict__ beta,
                                    const float *__restrict__ running_mean, const float *__restrict__ running_var,
                                    float eps, int N, float *__restrict__ scale, float *__restrict__ shift) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    float invstd = 1.0f / sqrtf(running_var[c] + eps);
    float sc = (gamma ? gamma[c] : 1.0f) * invstd;
    scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.0f) - running_mean[c] * sc;
}

// ------------------------------------------------------------------ forward tails
template <typename T>
__global__ void bn_relu_max_kernel(const T *__restrict__ Z, int ldz, const float *__restrict__ scale,
                                   const float *__restrict__ shift, int64_t G, int nsample, int C,
                                   float *__restrict__ out, int32_t *__restrict__ arg, T *__restrict__ zmax, int ldzm) {
    const int64_t total = G * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(e % C);
        int64_t g = e / C;
        const float sc = scale[c], sh = shift[c];
        const T *z = Z + g * nsample * (int64_t)ldz + c;
        float best = -1.0f, bz = 0.0f;
        int bk = 0;
        for (int k = 0; k < nsample; ++k) {
            const float zz = ld_act<T>(z + (int64_t)k * ldz);
            float a = fmaxf(fmaf(zz, sc, sh), 0.0f);
            if (a > best) { best = a; bk = k; bz = zz; }
        }
        out[e] = best;
        if (arg) arg[e] = bk;
        if (zmax) st_act<T>(zmax + g * ldzm + c, bz);
    }
}

// bf16 rows, C % 8 == 0: one thread owns 8 channels of one group and streams its nsample rows with 16-byte loads, U rows
// in flight (the scalar kernel above issues one 2-byte load per element: 1.9 TB/s on sa1's 134 MB; this one is HBM-bound)
__global__ void __launch_bounds__(256)
bn_relu_max_vec8_kernel(const __nv_bfloat16 *__restrict__ Z, int ldz, const float *__restrict__ scale,
                        const float *__restrict__ shift, int64_t G, int nsample, int C, float *__restrict__ out,
                        int32_t *__restrict__ arg, __nv_bfloat16 *__restrict__ zmax, int ldzm) {
    const int cpr = C >> 3;
    const int64_t total = G * cpr;
    constexpr int U = 8;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t g;
        int c8;
        fast_divmod(t, cpr, g, c8);
        const int c0 = c8 << 3;
        float sc[8], sh[8], best[8], bz[8];
        int bk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            sc[e] = scale[c0 + e];
            sh[e] = shift[c0 + e];
            best[e] = -1.0f;
            bz[e] = 0.0f;
            bk[e] = 0;
        }
        const __nv_bfloat16 *z = Z + g * nsample * (int64_t)ldz + c0;
        for (int k0 = 0; k0 < nsample; k0 += U) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (k0 + u < nsample) v[u] = *reinterpret_cast<const uint4 *>(z + (int64_t)(k0 + u) * ldz);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (k0 + u >= nsample) break;
                const uint32_t *w = reinterpret_cast<const uint32_t *>(&v[u]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w[i]));
                    const float a0 = fmaxf(fmaf(f.x, sc[2 * i], sh[2 * i]), 0.0f);
                    const float a1 = fmaxf(fmaf(f.y, sc[2 * i + 1], sh[2 * i + 1]), 0.0f);
                    if (a0 > best[2 * i]) { best[2 * i] = a0; bk[2 * i] = k0 + u; bz[2 * i] = f.x; }
                    if (a1 > best[2 * i + 1]) { best[2 * i + 1] = a1; bk[2 * i + 1] = k0 + u; bz[2 * i + 1] = f.y; }
                }
            }
        }
        float *o = out + g * C + c0;
        *reinterpret_cast<float4 *>(o) = make_float4(best[0], best[1], best[2], best[3]);
        *reinterpret_cast<float4 *>(o + 4) = make_float4(best[4], best[5], best[6], best[7]);
        if (arg) {
            int32_t *a = arg + g * C + c0;
            *reinterpret_cast<int4 *>(a) = make_int4(bk[0], bk[1], bk[2], bk[3]);
            *reinterpret_cast<int4 *>(a + 4) = make_int4(bk[4], bk[5], bk[6], bk[7]);
        }
        if (zmax) {      // the pre-BatchNorm value that won (exactly the stored bf16): the pooled backward reduces over [G, C] rows
            uint4 zo;
            uint32_t *zw = reinterpret_cast<uint32_t *>(&zo);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h = __floats2bfloat162_rn(bz[2 * i], bz[2 * i + 1]);
                zw[i] = *reinterpret_cast<uint32_t *>(&h);
            }
            *reinterpret_cast<uint4 *>(zmax + g * ldzm + c0) = zo;
        }
    }
}

template <typename T>
__global__ void bn_relu_kernel(const T *__restrict__ Z, int ldz, const float *__restrict__ scale,
                               const float *__restrict__ shift, int64_t M, int C, float *__restrict__ out) {
    const int64_t total = M * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int c;
        int64_t m;
        fast_divmod(e, C, m, c);
        out[e] = fmaxf(fmaf(ld_act<T>(Z + m * ldz + c), scale[c], shift[c]), 0.0f);
    }
}

// ------------------------------------------------------------------ backward: reductions
// Block layout: 256 threads = 8 row-lanes x 32 column-lanes; a block owns a contiguous slab
// of rows and writes one partial [2][C].  POOL: "rows" are groups, and the only contributing
// row of (group, channel) is its arg-max sample.
template <typename TA, typename TZ, bool POOL>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const TA *__restrict__ dA, int ldda, const int32_t *__restrict__ arg,
                     const TZ *__restrict__ Z, int ldz, const float *__restrict__ scale,
                     const float *__restrict__ shift, const float *__restrict__ save_mean,
                     const float *__restrict__ save_invstd, int64_t R, int nsample, int C,
                     double *__restrict__ accum, unsigned *ticket, float *dgamma, float *dbeta) {
    extern __shared__ float red[];   // [8][2][C]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t rows_per_block = (R + gridDim.x - 1) / gridDim.x;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(R, r_begin + rows_per_block);
    for (int c = lane; c < C; c += 32) {
        const float sc = scale[c], sh = shift[c];
        const float mu = save_mean ? save_mean[c] : 0.0f, is = save_invstd ? save_invstd[c] : 0.0f;
        float s1 = 0.0f, s2 = 0.0f;
        for (int64_t r = r_begin + w; r < r_end; r += 8) {
            int64_t zrow = POOL ? r * nsample + arg[r * C + c] : r;
            float z = ld_act<TZ>(Z + zrow * ldz + c);
            float g = POOL ? ((const float *)dA)[r * C + c] : ld_act<TA>(dA + r * ldda + c);
            if (!(fmaf(z, sc, sh) > 0.0f)) g = 0.0f;
            s1 += g;
            s2 = fmaf(g, (z - mu) * is, s2);
        }
        red[(w * 2 + 0) * C + c] = s1;
        red[(w * 2 + 1) * C + c] = s2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += 256) {
        int which = i / C, c = i % C;
        float s = 0.0f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) s += red[(ww * 2 + which) * C + c];
        atomicAdd(accum + (size_t)(blockIdx.x % kStatReplicas) * 2 * C + i, (double)s);
    }
    bn_bwd_last_block_finalize(accum, C, ticket, dgamma, dbeta);
}

// bf16 Z, C % 8 == 0 and 256 % (C/8) == 0: one thread owns 8 channels (16-byte loads) of every
// (256 / (C/8))-th row, four rows in flight per thread, 16 fp32 accumulators in registers.
// MODE 0: dense bf16 dA, 1: dense fp32 dA, 2: pooled fp32 dOut + arg-max map ("rows" are groups).
template <int MODE>
__global__ void __launch_bounds__(256, 3)
bn_bwd_reduce_vec8_kernel(const void *__restrict__ dA_, int ldda, const int32_t *__restrict__ arg,
                          const __nv_bfloat16 *__restrict__ Z, int ldz, const float *__restrict__ scale,
                          const float *__restrict__ shift, const float *__restrict__ save_mean,
                          const float *__restrict__ save_invstd, int64_t R, int nsample, int C,
                          double *__restrict__ accum, unsigned *ticket, float *dgamma, float *dbeta) {
    constexpr int kStride = 264;
    __shared__ float red[16 * kStride];
    const int cpr = C >> 3;
    const int c8 = threadIdx.x % cpr, rl = threadIdx.x / cpr, rstep = 256 / cpr;
    const int c0 = c8 << 3;
    float sc[8], sh[8], mu[8], is[8], s1[8], s2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        sc[e] = scale[c0 + e];
        sh[e] = shift[c0 + e];
        mu[e] = save_mean ? save_mean[c0 + e] : 0.0f;
        is[e] = save_invstd ? save_invstd[c0 + e] : 0.0f;
        s1[e] = s2[e] = 0.0f;
    }
    const int64_t rows_per_block = (R + gridDim.x - 1) / gridDim.x;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(R, r_begin + rows_per_block);
    constexpr int U = MODE == 2 ? 2 : 4;
    for (int64_t r = r_begin + rl; r < r_end; r += (int64_t)rstep * U) {
        float z[U][8], g[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t rr = r + (int64_t)u * rstep;
            const bool ok = rr < r_end;
            if (MODE == 2) {
                int4 a0 = make_int4(0, 0, 0, 0), a1 = a0;
                float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0;
                if (ok) {
                    a0 = *reinterpret_cast<const int4 *>(arg + rr * C + c0);
                    a1 = *reinterpret_cast<const int4 *>(arg + rr * C + c0 + 4);
                    d0 = *reinterpret_cast<const float4 *>((const float *)dA_ + rr * C + c0);
                    d1 = *reinterpret_cast<const float4 *>((const float *)dA_ + rr * C + c0 + 4);
                }
                const int ai[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float di[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    g[u][e] = di[e];
                    z[u][e] = ok ? __bfloat162float(Z[(rr * nsample + ai[e]) * ldz + c0 + e]) : 0.0f;
                }
            } else {
                uint4 zr = make_uint4(0u, 0u, 0u, 0u);
                if (ok) zr = *reinterpret_cast<const uint4 *>(Z + rr * ldz + c0);
                const uint32_t *zw = reinterpret_cast<const uint32_t *>(&zr);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&zw[i]));
                    z[u][2 * i] = f.x;
                    z[u][2 * i + 1] = f.y;
                }
                if (MODE == 0) {
                    uint4 gr = make_uint4(0u, 0u, 0u, 0u);
                    if (ok) gr = *reinterpret_cast<const uint4 *>((const __nv_bfloat16 *)dA_ + rr * ldda + c0);
                    const uint32_t *gw = reinterpret_cast<const uint32_t *>(&gr);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&gw[i]));
                        g[u][2 * i] = f.x;
                        g[u][2 * i + 1] = f.y;
                    }
                } else {
                    float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0;
                    if (ok) {
                        d0 = *reinterpret_cast<const float4 *>((const float *)dA_ + rr * ldda + c0);
                        d1 = *reinterpret_cast<const float4 *>((const float *)dA_ + rr * ldda + c0 + 4);
                    }
                    g[u][0] = d0.x; g[u][1] = d0.y; g[u][2] = d0.z; g[u][3] = d0.w;
                    g[u][4] = d1.x; g[u][5] = d1.y; g[u][6] = d1.z; g[u][7] = d1.w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float gg = (fmaf(z[u][e], sc[e], sh[e]) > 0.0f) ? g[u][e] : 0.0f;
                s1[e] += gg;
                s2[e] = fmaf(gg, (z[u][e] - mu[e]) * is[e], s2[e]);
            }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        red[e * kStride + threadIdx.x] = s1[e];
        red[(8 + e) * kStride + threadIdx.x] = s2[e];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += 256) {
        const int which = i / C, c = i - which * C;
        const float *p = red + (which * 8 + (c & 7)) * kStride + (c >> 3);
        float s = 0.0f;
        for (int q = 0; q < rstep; ++q) s += p[q * cpr];     // fixed order
        atomicAdd(accum + (size_t)(blockIdx.x % kStatReplicas) * 2 * C + i, (double)s);
    }
    bn_bwd_last_block_finalize(accum, C, ticket, dgamma, dbeta);
}

__global__ void bn_bwd_finalize_kernel(double *__restrict__ accum, int C, float *__restrict__ dgamma,
                                       float *__restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int r = 0; r < kStatReplicas; ++r) {
        double *a = accum + (size_t)r * 2 * C;
        s1 += a[c];
        s2 += a[C + c];
        a[c] = 0.0;
        a[C + c] = 0.0;
    }
    dbeta[c] = (float)s1;
    dgamma[c] = (float)s2;
}

// ------------------------------------------------------------------ backward: dz
template <typename TA, typename TZ, typename TD, bool POOL>
__global__ void bn_bwd_dz_kernel(const TA *dA, int ldda, const int32_t *__restrict__ arg,
                                 const TZ *__restrict__ Z, int ldz, const float *__restrict__ scale,
                                 const float *__restrict__ shift, const float *__restrict__ save_mean,
                                 const float *__restrict__ save_invstd, const float *__restrict__ dgamma,
                                 const float *__restrict__ dbeta, int64_t M, int nsample, int C, float inv_m,
                                 TD *dZ, int lddz) {
    const int64_t total = M * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(e % C);
        int64_t m = e / C;
        float z = ld_act<TZ>(Z + m * ldz + c);
        float g;
        if (POOL) {
            int64_t grp = m / nsample;
            int k = (int)(m % nsample);
            g = (arg[grp * C + c] == k) ? ((const float *)dA)[grp * C + c] : 0.0f;
        } else {
            g = ld_act<TA>(dA + m * ldda + c);
        }
        const float sc = scale[c];
        if (!(fmaf(z, sc, shift[c]) > 0.0f)) g = 0.0f;
        float d;
        if (save_mean) {   // train: dz = gamma*invstd*(g - dbeta/M - zhat*dgamma/M), scale == gamma*invstd
            float zhat = (z - save_mean[c]) * save_invstd[c];
            d = sc * (g - dbeta[c] * inv_m - zhat * dgamma[c] * inv_m);
        } else {
            d = sc * g;
        }
        st_act<TD>(dZ + m * lddz + c, d);
    }
}

// bf16 rows, C % 8 == 0, all leading dimensions % 8 == 0: one thread per 16-byte chunk (8 channels).
// dz = sc*g + a*z + b with a = -sc*dgamma*invstd/M, b = -sc*dbeta/M - a*mean (train) or a = b = 0 (frozen
// statistics); the four per-channel coefficient vectors are staged in shared memory once per CTA.
template <bool POOL>
__global__ void __launch_bounds__(256)
bn_bwd_dz_vec8_kernel(const __nv_bfloat16 *dA, int ldda, const float *__restrict__ dOut,
                      const int32_t *__restrict__ arg, const __nv_bfloat16 *__restrict__ Z, int ldz,
                      const float *__restrict__ scale, const float *__restrict__ shift,
                      const float *__restrict__ save_mean, const float *__restrict__ save_invstd,
                      const float *__restrict__ dgamma, const float *__restrict__ dbeta, int64_t M, int nsample,
                      int C, float inv_m, __nv_bfloat16 *dZ, int lddz) {
    extern __shared__ float coef[];   // [4][C]: sc, sh, a, b
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sc = scale[c];
        float a = 0.0f, b = 0.0f;
        if (save_mean) {
            a = -sc * dgamma[c] * save_invstd[c] * inv_m;
            b = -sc * dbeta[c] * inv_m - a * save_mean[c];
        }
        coef[c] = sc;
        coef[C + c] = shift[c];
        coef[2 * C + c] = a;
        coef[3 * C + c] = b;
    }
    __syncthreads();
    const int cpr = C >> 3;
    const int64_t total = M * cpr;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        int64_t m;
        int c8;
        fast_divmod(q, cpr, m, c8);
        const int c0 = c8 << 3;
        const uint4 zr = *reinterpret_cast<const uint4 *>(Z + m * ldz + c0);
        float g[8];
        if (POOL) {
            int64_t grp;
            int k;
            fast_divmod(m, nsample, grp, k);
            const int4 a0 = *reinterpret_cast<const int4 *>(arg + grp * C + c0);
            const int4 a1 = *reinterpret_cast<const int4 *>(arg + grp * C + c0 + 4);
            const float4 d0 = *reinterpret_cast<const float4 *>(dOut + grp * C + c0);
            const float4 d1 = *reinterpret_cast<const float4 *>(dOut + grp * C + c0 + 4);
            g[0] = a0.x == k ? d0.x : 0.f; g[1] = a0.y == k ? d0.y : 0.f; g[2] = a0.z == k ? d0.z : 0.f; g[3] = a0.w == k ? d0.w : 0.f;
            g[4] = a1.x == k ? d1.x : 0.f; g[5] = a1.y == k ? d1.y : 0.f; g[6] = a1.z == k ? d1.z : 0.f; g[7] = a1.w == k ? d1.w : 0.f;
        } else {
            const uint4 gr = *reinterpret_cast<const uint4 *>(dA + m * ldda + c0);
            const uint32_t *gw = reinterpret_cast<const uint32_t *>(&gr);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&gw[i]);
                float2 f = __bfloat1622float2(h);
                g[2 * i] = f.x;
                g[2 * i + 1] = f.y;
            }
        }
        const uint32_t *zw = reinterpret_cast<const uint32_t *>(&zr);
        uint4 out;
        uint32_t *ow = reinterpret_cast<uint32_t *>(&out);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&zw[i]);
            float2 z = __bfloat1622float2(h);
            float d[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = c0 + 2 * i + j;
                const float zz = j ? z.y : z.x;
                const float sc = coef[c];
                const float gg = (fmaf(zz, sc, coef[C + c]) > 0.0f) ? g[2 * i + j] : 0.0f;
                d[j] = fmaf(sc, gg, fmaf(coef[2 * C + c], zz, coef[3 * C + c]));
            }
            __nv_bfloat162 o = __floats2bfloat162_rn(d[0], d[1]);
            ow[i] = *reinterpret_cast<uint32_t *>(&o);
        }
        *reinterpret_cast<uint4 *>(dZ + m * lddz + c0) = out;
    }
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_bn_train_finalize(double *stat_accum, int64_t M, int N,
                                     const float *gamma, const float *beta, const float *conv_bias, float eps,
                                     float momentum, float *running_mean, float *running_var, float *scale,
                                     float *shift, float *save_mean, float *save_invstd,
                                     int64_t *num_batches_tracked, void *stream) {
    PN2_REQUIRE(stat_accum && scale && shift, "bn_train_finalize: null pointer");
    PN2_REQUIRE(M > 0 && N > 0, "bn_train_finalize: bad sizes");
    bn_train_finalize_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        stat_accum, M, N, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, scale,
        shift, save_mean, save_invstd, (long long *)num_batches_tracked);
    count_launch();
    return check_launch("bn_train_finalize");
}

extern "C" int pn2_bn_eval_fold(const float *gamma, const float *beta, const float *running_mean,
                                const float *running_var, float eps, int N, float *scale, float *shift,
                                void *stream) {
    PN2_REQUIRE(running_mean && running_var && scale && shift, "bn_eval_fold: null pointer");
    bn_eval_fold_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var, eps, N,
                                                                           scale, shift);
    count_launch();
    return check_launch("bn_eval_fold");
}

static int bn_relu_max_impl(const void *Z, int ldz, int z_dtype, const float *scale, const float *shift, int64_t G, int nsample,
                            int C, float *out, int32_t *arg, void *zmax, int ldzm, void *stream) {
    int64_t total = G * C;
    if (total == 0) return PN2_OK;
    if (z_dtype == PN2_BF16 && C % 8 == 0 && ldz % 8 == 0 && ((uintptr_t)Z & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
        (!arg || ((uintptr_t)arg & 15) == 0) && (!zmax || (((uintptr_t)zmax & 15) == 0 && ldzm % 8 == 0))) {
        bn_relu_max_vec8_kernel<<<grid_for(total / 8, 256), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16 *)Z, ldz, scale, shift, G, nsample, C, out, arg, (__nv_bfloat16 *)zmax, ldzm);
        count_launch();
        return check_launch("bn_relu_max_vec8");
    }
    PN2_DISPATCH_DTYPE(z_dtype, T, (bn_relu_max_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const T *)Z, ldz, scale, shift, G, nsample, C, out, arg, (T *)zmax, ldzm)));
    count_launch();
    return check_launch("bn_relu_max");
}

extern "C" int pn2_bn_relu_max(const void *Z, int ldz, int z_dtype, const float *scale, const float *shift,
                               int64_t G, int nsample, int C, float *out, int32_t *arg, void *stream) {
    PN2_REQUIRE(Z && scale && shift && out, "bn_relu_max: null pointer");
    PN2_REQUIRE(valid_dtype(z_dtype) && nsample >= 1 && C >= 1 && ldz >= C, "bn_relu_max: bad arguments");
    return bn_relu_max_impl(Z, ldz, z_dtype, scale, shift, G, nsample, C, out, arg, nullptr, 0, stream);
}

extern "C" int pn2_bn_relu_max_keep(const void *Z, int ldz, int z_dtype, const float *scale, const float *shift,
                                    int64_t G, int nsample, int C, float *out, int32_t *arg, void *zmax, int ldzm,
                                    void *stream) {
    PN2_REQUIRE(Z && scale && shift && out && zmax, "bn_relu_max_keep: null pointer");
    PN2_REQUIRE(valid_dtype(z_dtype) && nsample >= 1 && C >= 1 && ldz >= C && ldzm >= C, "bn_relu_max_keep: bad arguments");
    return bn_relu_max_impl(Z, ldz, z_dtype, scale, shift, G, nsample, C, out, arg, zmax, ldzm, stream);
}

extern "C" int pn2_bn_relu(const void *Z, int ldz, int z_dtype, const float *scale, const float *shift,
                           int64_t M, int C, float *out, void *stream) {
    PN2_REQUIRE(Z && scale && shift && out, "bn_relu: null pointer");
    PN2_REQUIRE(valid_dtype(z_dtype) && C >= 1 && ldz >= C, "bn_relu: bad arguments");
    int64_t total = M * C;
    if (total == 0) return PN2_OK;
    PN2_DISPATCH_DTYPE(z_dtype, T, (bn_relu_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const T *)Z, ldz, scale, shift, M, C, out)));
    count_launch();
    return check_launch("bn_relu");
}

template <bool POOL>
static int reduce_dispatch(const void *dA, int ldda, int da_dtype, const int32_t *arg, const void *Z, int ldz,
                           int z_dtype, const float *scale, const float *shift, const float *save_mean,
                           const float *save_invstd, int64_t R, int nsample, int C, double *accum,
                           int n_partials, cudaStream_t st, unsigned *ticket = nullptr, float *dgamma = nullptr,
                           float *dbeta = nullptr) {
    if (z_dtype == PN2_BF16 && C % 8 == 0 && 256 % (C / 8) == 0 && ldz % 8 == 0 &&
        (POOL || (da_dtype == PN2_BF16 ? ldda % 8 == 0 : ldda % 4 == 0))) {
        const int rstep = 256 / (C / 8);
        const int per_pass = rstep * (POOL ? 2 : 4);
        int64_t want = (R + per_pass - 1) / per_pass;
        // three co-resident CTAs per SM (register-limited, __launch_bounds__(256, 3)): every loop trip exposes one
        // full memory latency, so the bytes in flight come from co-resident warps (2/SM measured 2.3 TB/s on sa1)
        const int grid = (int)(want < 1 ? 1 : (want > 3 * kNumSMs ? 3 * kNumSMs : want));
        if (POOL)
            bn_bwd_reduce_vec8_kernel<2><<<grid, 256, 0, st>>>(dA, ldda, arg, (const __nv_bfloat16 *)Z, ldz, scale, shift,
                                                               save_mean, save_invstd, R, nsample, C, accum, ticket, dgamma, dbeta);
        else if (da_dtype == PN2_BF16)
            bn_bwd_reduce_vec8_kernel<0><<<grid, 256, 0, st>>>(dA, ldda, arg, (const __nv_bfloat16 *)Z, ldz, scale, shift,
                                                               save_mean, save_invstd, R, nsample, C, accum, ticket, dgamma, dbeta);
        else
            bn_bwd_reduce_vec8_kernel<1><<<grid, 256, 0, st>>>(dA, ldda, arg, (const __nv_bfloat16 *)Z, ldz, scale, shift,
                                                               save_mean, save_invstd, R, nsample, C, accum, ticket, dgamma, dbeta);
        count_launch();
        return check_launch("bn_bwd_reduce_vec8");
    }
    size_t smem = sizeof(float) * 16 * (size_t)C;
#define PN2_LAUNCH_RED(TA, TZ)                                                                        \
    bn_bwd_reduce_kernel<TA, TZ, POOL><<<n_partials, 256, smem, st>>>((const TA *)dA, ldda, arg, (const TZ *)Z, ldz, \
                                                                      scale, shift, save_mean, save_invstd, R, nsample, C, accum, ticket, dgamma, dbeta)
    if (da_dtype == PN2_F32 && z_dtype == PN2_F32) PN2_LAUNCH_RED(float, float);
    else if (da_dtype == PN2_F32) PN2_LAUNCH_RED(float, __nv_bfloat16);
    else if (z_dtype == PN2_F32) PN2_LAUNCH_RED(__nv_bfloat16, float);
    else PN2_LAUNCH_RED(__nv_bfloat16, __nv_bfloat16);
#undef PN2_LAUNCH_RED
    count_launch();
    return check_launch("bn_bwd_reduce");
}

extern "C" int pn2_bn_relu_bwd_reduce(const void *dA, int ldda, int da_dtype, const void *Z, int ldz, int z_dtype,
                                      const float *scale, const float *shift, const float *save_mean,
                                      const float *save_invstd, int64_t M, int C, double *accum, void *stream) {
    PN2_REQUIRE(dA && Z && scale && shift && accum, "bn_relu_bwd_reduce: null pointer");
    PN2_REQUIRE(valid_dtype(da_dtype) && valid_dtype(z_dtype) && C >= 1 && C <= 768, "bn_relu_bwd_reduce: bad arguments (C <= 768)");
    if (M == 0) return PN2_OK;
    return reduce_dispatch<false>(dA, ldda, da_dtype, nullptr, Z, ldz, z_dtype, scale, shift, save_mean, save_invstd, M,
                                  1, C, accum, linear_num_partials(M), (cudaStream_t)stream);
}

extern "C" int pn2_pool_bn_relu_bwd_reduce(const float *dOut, const int32_t *arg, const void *Z, int ldz,
                                           int z_dtype, const float *scale, const float *shift,
                                           const float *save_mean, const float *save_invstd, int64_t G,
                                           int nsample, int C, double *accum, void *stream) {
    PN2_REQUIRE(dOut && arg && Z && scale && shift && accum, "pool_bn_relu_bwd_reduce: null pointer");
    PN2_REQUIRE(valid_dtype(z_dtype) && C >= 1 && C <= 768, "pool_bn_relu_bwd_reduce: bad arguments (C <= 768)");
    if (G == 0) return PN2_OK;
    return reduce_dispatch<true>(dOut, C, PN2_F32, arg, Z, ldz, z_dtype, scale, shift, save_mean, save_invstd, G, nsample,
                                 C, accum, linear_num_partials(G * nsample), (cudaStream_t)stream);
}

extern "C" int pn2_bn_relu_bwd_reduce_finalize(const void *dA, int ldda, int da_dtype, const void *Z, int ldz,
                                               int z_dtype, const float *scale, const float *shift,
                                               const float *save_mean, const float *save_invstd, int64_t M, int C,
                                               double *accum, unsigned *ticket, float *dgamma, float *dbeta,
                                               void *stream) {
    PN2_REQUIRE(dA && Z && scale && shift && accum && ticket && dgamma && dbeta, "bn_relu_bwd_reduce_finalize: null pointer");
    PN2_REQUIRE(valid_dtype(da_dtype) && valid_dtype(z_dtype) && C >= 1 && C <= 768 && M >= 1,
                "bn_relu_bwd_reduce_finalize: bad arguments (C <= 768, M >= 1)");
    return reduce_dispatch<false>(dA, ldda, da_dtype, nullptr, Z, ldz, z_dtype, scale, shift, save_mean, save_invstd, M,
                                  1, C, accum, linear_num_partials(M), (cudaStream_t)stream, ticket, dgamma, dbeta);
}

extern "C" int pn2_pool_bn_relu_bwd_reduce_finalize(const float *dOut, const int32_t *arg, const void *Z, int ldz,
                                                    int z_dtype, const float *scale, const float *shift,
                                                    const float *save_mean, const float *save_invstd, int64_t G,
                                                    int nsample, int C, double *accum, unsigned *ticket,
                                                    float *dgamma, float *dbeta, void *stream) {
    PN2_REQUIRE(dOut && arg && Z && scale && shift && accum && ticket && dgamma && dbeta,
                "pool_bn_relu_bwd_reduce_finalize: null pointer");
    PN2_REQUIRE(valid_dtype(z_dtype) && C >= 1 && C <= 768 && G >= 1, "pool_bn_relu_bwd_reduce_finalize: bad arguments (C <= 768, G >= 1)");
    return reduce_dispatch<true>(dOut, C, PN2_F32, arg, Z, ldz, z_dtype, scale, shift, save_mean, save_invstd, G, nsample,
                                 C, accum, linear_num_partials(G * nsample), (cudaStream_t)stream, ticket, dgamma, dbeta);
}

extern "C" int pn2_bn_bwd_finalize(double *accum, int C, float *dgamma, float *dbeta, void *stream) {
    PN2_REQUIRE(accum && dgamma && dbeta && C > 0, "bn_bwd_finalize: bad arguments");
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(accum, C, dgamma, dbeta);
    count_launch();
    return check_launch("bn_bwd_finalize");
}

// Dense bf16 dA / Z / dZ with C % 8 == 0 and 256 % (C/8) == 0 (every layer width of the network): a thread owns 8 channels --
// its 32 coefficients (sc, sh, a, b) stay in REGISTERS for the whole kernel -- of every (256 / (C/8))-th row of the block's
// contiguous row range, U rows (2U 16-byte loads) in flight.  bn_bwd_dz_vec8_kernel re-reads the four coefficients of every
// element from shared memory (~26 LDS per 16 bytes of dZ: half of the issue slots on the 1 M-row layers, ncu: 51 % issue,
// mio_throttle) and keeps one row in flight per thread.  dZ may alias dA (a thread reads its chunks before it writes them).
template <int U>
__global__ void __launch_bounds__(256, 3)
bn_bwd_dz_rows_kernel(const __nv_bfloat16 *dA, int ldda, const __nv_bfloat16 *__restrict__ Z, int ldz,
                      const float *__restrict__ scale, const float *__restrict__ shift,
                      const float *__restrict__ save_mean, const float *__restrict__ save_invstd,
                      const float *__restrict__ dgamma, const float *__restrict__ dbeta, int64_t M, int C, float inv_m,
                      __nv_bfloat16 *dZ, int lddz) {
    const int cpr = C >> 3;
    const int c8 = threadIdx.x % cpr, rl = threadIdx.x / cpr, rstep = 256 / cpr;
    const int c0 = c8 << 3;
    float sc[8], sh[8], ca[8], cb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        sc[e] = scale[c0 + e];
        sh[e] = shift[c0 + e];
        ca[e] = cb[e] = 0.0f;
        if (save_mean) {
            ca[e] = -sc[e] * dgamma[c0 + e] * save_invstd[c0 + e] * inv_m;
            cb[e] = -sc[e] * dbeta[c0 + e] * inv_m - ca[e] * save_mean[c0 + e];
        }
    }
    const int64_t rows_per_block = (M + gridDim.x - 1) / gridDim.x;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(M, r_begin + rows_per_block);
    for (int64_t r = r_begin + rl; r < r_end; r += (int64_t)rstep * U) {
        uint4 zr[U], gr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t rr = r + (int64_t)u * rstep;
            if (rr < r_end) {
                zr[u] = *reinterpret_cast<const uint4 *>(Z + rr * ldz + c0);
                gr[u] = *reinterpret_cast<const uint4 *>(dA + rr * ldda + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t rr = r + (int64_t)u * rstep;
            if (rr >= r_end) break;
            const uint32_t *zw = reinterpret_cast<const uint32_t *>(&zr[u]);
            const uint32_t *gw = reinterpret_cast<const uint32_t *>(&gr[u]);
            uint4 out;
            uint32_t *ow = reinterpret_cast<uint32_t *>(&out);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 z = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&zw[i]));
                const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&gw[i]));
                const float g0 = (fmaf(z.x, sc[2 * i], sh[2 * i]) > 0.0f) ? g.x : 0.0f;
                const float g1 = (fmaf(z.y, sc[2 * i + 1], sh[2 * i + 1]) > 0.0f) ? g.y : 0.0f;
                const float d0 = fmaf(sc[2 * i], g0, fmaf(ca[2 * i], z.x, cb[2 * i]));
                const float d1 = fmaf(sc[2 * i + 1], g1, fmaf(ca[2 * i + 1], z.y, cb[2 * i + 1]));
                const __nv_bfloat162 o = __floats2bfloat162_rn(d0, d1);
                ow[i] = *reinterpret_cast<const uint32_t *>(&o);
            }
            *reinterpret_cast<uint4 *>(dZ + rr * lddz + c0) = out;
        }
    }
}

// Pooled variant of bn_bwd_dz_vec8_kernel with the group's gradient and arg-max map loaded ONCE per thread for KU samples:
// the per-sample form re-reads 64 bytes of (dOut, arg) from L2 for every 16 bytes of Z, which made it L2-bound (1.8 TB/s of
// HBM traffic on sa1); here the ratio is 64 : 16*KU and the KU row loads are in flight together.
template <int KU>
__global__ void __launch_bounds__(256)
pool_bwd_dz_vec8_kernel(const float *__restrict__ dOut, const int32_t *__restrict__ arg, const __nv_bfloat16 *__restrict__ Z,
                        int ldz, const float *__restrict__ scale, const float *__restrict__ shift,
                        const float *__restrict__ save_mean, const float *__restrict__ save_invstd,
                        const float *__restrict__ dgamma, const float *__restrict__ dbeta, int64_t G, int nsample, int C,
                        float inv_m, __nv_bfloat16 *__restrict__ dZ, int lddz) {
    extern __shared__ float coef[];   // [4][C]: sc, sh, a, b
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sc = scale[c];
        float a = 0.0f, b = 0.0f;
        if (save_mean) {
            a = -sc * dgamma[c] * save_invstd[c] * inv_m;
            b = -sc * dbeta[c] * inv_m - a * save_mean[c];
        }
        coef[c] = sc;
        coef[C + c] = shift[c];
        coef[2 * C + c] = a;
        coef[3 * C + c] = b;
    }
    __syncthreads();
    const int cpr = C >> 3, kchunks = (nsample + KU - 1) / KU;
    const int64_t total = G * kchunks * cpr;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        int64_t t, grp;
        int c8, kc;
        fast_divmod(q, cpr, t, c8);
        fast_divmod(t, kchunks, grp, kc);
        const int c0 = c8 << 3, k0 = kc * KU;
        const int4 a0 = *reinterpret_cast<const int4 *>(arg + grp * C + c0);
        const int4 a1 = *reinterpret_cast<const int4 *>(arg + grp * C + c0 + 4);
        const float4 d0 = *reinterpret_cast<const float4 *>(dOut + grp * C + c0);
        const float4 d1 = *reinterpret_cast<const float4 *>(dOut + grp * C + c0 + 4);
        const int ai[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float di[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const int64_t m0 = grp * nsample + k0;
        uint4 zr[KU];
#pragma unroll
        for (int u = 0; u < KU; ++u)
            if (k0 + u < nsample) zr[u] = *reinterpret_cast<const uint4 *>(Z + (m0 + u) * ldz + c0);
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            if (k0 + u >= nsample) break;
            const uint32_t *zw = reinterpret_cast<const uint32_t *>(&zr[u]);
            uint4 out;
            uint32_t *ow = reinterpret_cast<uint32_t *>(&out);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 z = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&zw[i]));
                float d[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int e = 2 * i + j, c = c0 + e;
                    const float zz = j ? z.y : z.x;
                    const float sc = coef[c];
                    const float g = ai[e] == k0 + u ? di[e] : 0.0f;
                    const float gg = (fmaf(zz, sc, coef[C + c]) > 0.0f) ? g : 0.0f;
                    d[j] = fmaf(sc, gg, fmaf(coef[2 * C + c], zz, coef[3 * C + c]));
                }
                __nv_bfloat162 o = __floats2bfloat162_rn(d[0], d[1]);
                ow[i] = *reinterpret_cast<uint32_t *>(&o);
            }
            *reinterpret_cast<uint4 *>(dZ + (m0 + u) * lddz + c0) = out;
        }
    }
}

template <bool POOL>
static int dz_dispatch(const void *dA, int ldda, int da_dtype, const int32_t *arg, const void *Z, int ldz,
                       int z_dtype, const float *scale, const float *shift, const float *save_mean,
                       const float *save_invstd, const float *dgamma, const float *dbeta, int64_t M, int nsample,
                       int C, void *dZ, int lddz, int dz_dtype, cudaStream_t st) {
    const float inv_m = 1.0f / (float)M;
    if (z_dtype == PN2_BF16 && dz_dtype == PN2_BF16 && (POOL || da_dtype == PN2_BF16) && C % 8 == 0 && ldz % 8 == 0 &&
        lddz % 8 == 0 && (POOL || ldda % 8 == 0) && C <= 2048) {
        if (POOL && M % nsample == 0) {
            constexpr int KU = 8;
            const int64_t G = M / nsample;
            const int pgrid = grid_for(G * ((nsample + KU - 1) / KU) * (C / 8), 256, kNumSMs * 4);
            pool_bwd_dz_vec8_kernel<KU><<<pgrid, 256, sizeof(float) * 4 * C, st>>>(
                (const float *)dA, arg, (const __nv_bfloat16 *)Z, ldz, scale, shift, save_mean, save_invstd, dgamma, dbeta, G,
                nsample, C, inv_m, (__nv_bfloat16 *)dZ, lddz);
            count_launch();
            return check_launch("pool_bwd_dz_vec8");
        }
        if (!POOL && 256 % (C / 8) == 0) {
            constexpr int U = 3;
            const int per_pass = (256 / (C / 8)) * U;
            const int64_t want = (M + per_pass - 1) / per_pass;
            const int rgrid = (int)(want < 1 ? 1 : (want > 3 * kNumSMs ? 3 * kNumSMs : want));
            bn_bwd_dz_rows_kernel<U><<<rgrid, 256, 0, st>>>((const __nv_bfloat16 *)dA, ldda, (const __nv_bfloat16 *)Z, ldz, scale,
                                                            shift, save_mean, save_invstd, dgamma, dbeta, M, C, inv_m,
                                                            (__nv_bfloat16 *)dZ, lddz);
            count_launch();
            return check_launch("bn_bwd_dz_rows");
        }
        const int vgrid = grid_for(M * (C / 8), 256, kNumSMs * 8);
        bn_bwd_dz_vec8_kernel<POOL><<<vgrid, 256, sizeof(float) * 4 * C, st>>>(
            POOL ? nullptr : (const __nv_bfloat16 *)dA, ldda, POOL ? (const float *)dA : nullptr, arg,
            (const __nv_bfloat16 *)Z, ldz, scale, shift, save_mean, save_invstd, dgamma, dbeta, M, nsample, C, inv_m,
            (__nv_bfloat16 *)dZ, lddz);
        count_launch();
        return check_launch("bn_bwd_dz_vec8");
    }
    const int grid = grid_for(M * C, 256);
#define PN2_LAUNCH_DZ(TA, TZ, TD)                                                                                   \
    bn_bwd_dz_kernel<TA, TZ, TD, POOL><<<grid, 256, 0, st>>>((const TA *)dA, ldda, arg, (const TZ *)Z, ldz, scale, shift, \
                                                             save_mean, save_invstd, dgamma, dbeta, M, nsample, C, inv_m, \
                                                             (TD *)dZ, lddz)
    const bool af = da_dtype == PN2_F32, zf = z_dtype == PN2_F32, df = dz_dtype == PN2_F32;
    if (af && zf && df) PN2_LAUNCH_DZ(float, float, float);
    else if (af && zf) PN2_LAUNCH_DZ(float, float, __nv_bfloat16);
    else if (af && df) PN2_LAUNCH_DZ(float, __nv_bfloat16, float);
    else if (af) PN2_LAUNCH_DZ(float, __nv_bfloat16, __nv_bfloat16);
    else if (zf && df) PN2_LAUNCH_DZ(__nv_bfloat16, float, float);
    else if (zf) PN2_LAUNCH_DZ(__nv_bfloat16, float, __nv_bfloat16);
    else if (df) PN2_LAUNCH_DZ(__nv_bfloat16, __nv_bfloat16, float);
    else PN2_LAUNCH_DZ(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16);
#undef PN2_LAUNCH_DZ
    count_launch();
    return check_launch("bn_bwd_dz");
}

extern "C" int pn2_bn_relu_bwd_dz(const void *dA, int ldda, int da_dtype, const void *Z, int ldz, int z_dtype,
                                  const float *scale, const float *shift, const float *save_mean,
                                  const float *save_invstd, const float *dgamma, const float *dbeta, int64_t M,
                                  int C, void *dZ, int lddz, int dz_dtype, void *stream) {
    PN2_REQUIRE(dA && Z && scale && shift && dZ, "bn_relu_bwd_dz: null pointer");
    PN2_REQUIRE(!save_mean || (save_invstd && dgamma && dbeta), "bn_relu_bwd_dz: train mode needs invstd/dgamma/dbeta");
    PN2_REQUIRE(valid_dtype(da_dtype) && valid_dtype(z_dtype) && valid_dtype(dz_dtype), "bn_relu_bwd_dz: bad dtype");
    if (M == 0) return PN2_OK;
    return dz_dispatch<false>(dA, ldda, da_dtype, nullptr, Z, ldz, z_dtype, scale, shift, save_mean, save_invstd, dgamma,
                              dbeta, M, 1, C, dZ, lddz, dz_dtype, (cudaStream_t)stream);
}

extern "C" int pn2_pool_bn_relu_bwd_dz(const float *dOut, const int32_t *arg, const void *Z, int ldz, int z_dtype,
                                       const float *scale, const float *shift, const float *save_mean,
                                       const float *save_invstd, const float *dgamma, const float *dbeta,
                                       int64_t G, int nsample, int C, void *dZ, int lddz, int dz_dtype,
                                       void *stream) {
    PN2_REQUIRE(dOut && arg && Z && scale && shift && dZ, "pool_bn_relu_bwd_dz: null pointer");
    PN2_REQUIRE(!save_mean || (save_invstd && dgamma && dbeta), "pool_bn_relu_bwd_dz: train mode needs invstd/dgamma/dbeta");
    PN2_REQUIRE(valid_dtype(z_dtype) && valid_dtype(dz_dtype), "pool_bn_relu_bwd_dz: bad dtype");
    if (G == 0) return PN2_OK;
    return dz_dispatch<true>(dOut, C, PN2_F32, arg, Z, ldz, z_dtype, scale, shift, save_mean, save_invstd, dgamma, dbeta,
                             G * nsample, nsample, C, dZ, lddz, dz_dtype, (cudaStream_t)stream);
}
