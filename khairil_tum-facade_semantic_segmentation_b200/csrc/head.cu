// head.cu -- the tail of the segmentation head on point-major rows:
//
//   x = drop1(relu(bn1(conv1(l0_points))));  x = conv2(x);  x = log_softmax(x, dim=1)
//   (/root/reference/models/pointnet2_sem_seg.py:36-39)
//
// conv1 + bn1 run as one more layer of the fp1 MLP on the tensor-core layer kernels (linear_tc.cu); this file
// is what follows the pre-BatchNorm product Z[M,C] (bf16 rows) of that layer:
//   forward : a = relu(z*scale+shift) -> dropout -> logits = a.W2^T + b2 (C -> NC <= 32 classes) -> log_softmax,
//             one thread per point; W2 sits in shared memory ([k][class], broadcast reads), the NC accumulators in
//             registers; optionally stores the dropped activation (bf16 rows) that conv2's weight gradient needs.
//   backward: dlogits = dlogp - exp(logp)*sum(dlogp) (log_softmax), d(a) = (dlogits.W2) * mask/(1-p), sum of
//             dlogits per class (conv2's bias gradient) through fp64 atomics; dlogits is also written as bf16 rows
//             so that conv2's weight gradient is one call of the tensor-core wgrad kernel.
// Dropout: keep-mask bit of element (m, k) = hash(seed, m*C + k) >= p (counter-based, recomputed in backward from
// the same seed; the seed is read from DEVICE memory so a captured CUDA graph draws fresh masks on every replay).
// The reference's nn.Dropout consumes the CUDA Philox stream instead: parity with it is statistical (keep
// probability 1-p, scaling 1/(1-p)), and exact for p = 0 and in eval mode.
#include "common.cuh"

namespace pn2 {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {   // murmur3 finalizer
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}
// 8 keep bits for elements e0 .. e0+7 (e0 % 8 == 0) of the [M, C] activation: one hash per 4 elements, 8 bits each
__device__ __forceinline__ uint32_t keep_bits8(uint32_t seed, uint64_t e0, uint32_t thresh8) {
    uint32_t bits = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint64_t q = (e0 >> 2) + h;
        const uint32_t r = mix32(mix32((uint32_t)q ^ seed) + (uint32_t)(q >> 32) * 0x9e3779b9u + 0x7f4a7c15u);
#pragma unroll
        for (int i = 0; i < 4; ++i) bits |= (((r >> (8 * i)) & 0xffu) >= thresh8 ? 1u : 0u) << (4 * h + i);
    }
    return bits;
}

constexpr int kHeadThreads = 128;

// W2s: [C][NCP] fp32 in shared memory (zero padded classes), b2s [NCP]
template <int NCP>
__global__ void __launch_bounds__(kHeadThreads)
head_tail_fwd_kernel(const __nv_bfloat16 *__restrict__ Z, int ldz, const float *__restrict__ scale,
                     const float *__restrict__ shift, const float *__restrict__ W2, const float *__restrict__ b2,
                     int64_t M, int C, int NC, const int64_t *__restrict__ seed_ptr, uint32_t thresh8, float keep_scale,
                     float *__restrict__ logp, __nv_bfloat16 *__restrict__ act_out, int ldo) {
    extern __shared__ float smem[];
    float *w2s = smem, *b2s = smem + (size_t)C * NCP, *scs = b2s + NCP, *shs = scs + C;
    for (int i = threadIdx.x; i < C * NCP; i += kHeadThreads) {
        const int k = i / NCP, j = i - k * NCP;
        w2s[i] = j < NC ? W2[(size_t)j * C + k] : 0.0f;
    }
    for (int i = threadIdx.x; i < NCP; i += kHeadThreads) b2s[i] = (i < NC && b2) ? b2[i] : 0.0f;
    for (int i = threadIdx.x; i < C; i += kHeadThreads) {
        scs[i] = scale[i];
        shs[i] = shift[i];
    }
    __syncthreads();
    const uint32_t seed = seed_ptr ? (uint32_t)(*seed_ptr) : 0u;
    for (int64_t m = (int64_t)blockIdx.x * kHeadThreads + threadIdx.x; m < M; m += (int64_t)gridDim.x * kHeadThreads) {
        float acc[NCP];
#pragma unroll
        for (int j = 0; j < NCP; ++j) acc[j] = b2s[j];
        const __nv_bfloat16 *zrow = Z + m * ldz;
        for (int k0 = 0; k0 < C; k0 += 8) {
            const uint4 zr = *reinterpret_cast<const uint4 *>(zrow + k0);
            const uint32_t *zw = reinterpret_cast<const uint32_t *>(&zr);
            const uint32_t keep = thresh8 ? keep_bits8(seed, (uint64_t)m * C + k0, thresh8) : 0xffu;
            float a[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&zw[i]));
                a[2 * i] = fmaxf(fmaf(f.x, scs[k0 + 2 * i], shs[k0 + 2 * i]), 0.0f);
                a[2 * i + 1] = fmaxf(fmaf(f.y, scs[k0 + 2 * i + 1], shs[k0 + 2 * i + 1]), 0.0f);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = ((keep >> e) & 1u) ? a[e] * keep_scale : 0.0f;
            if (act_out) {
                uint4 o;
                __nv_bfloat162 h;
                h = __floats2bfloat162_rn(a[0], a[1]); o.x = *reinterpret_cast<uint32_t *>(&h);
                h = __floats2bfloat162_rn(a[2], a[3]); o.y = *reinterpret_cast<uint32_t *>(&h);
                h = __floats2bfloat162_rn(a[4], a[5]); o.z = *reinterpret_cast<uint32_t *>(&h);
                h = __floats2bfloat162_rn(a[6], a[7]); o.w = *reinterpret_cast<uint32_t *>(&h);
                *reinterpret_cast<uint4 *>(act_out + m * ldo + k0) = o;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4 *wr = reinterpret_cast<const float4 *>(w2s + (size_t)(k0 + e) * NCP);
#pragma unroll
                for (int j4 = 0; j4 < NCP / 4; ++j4) {
                    const float4 w = wr[j4];
                    acc[4 * j4 + 0] = fmaf(a[e], w.x, acc[4 * j4 + 0]);
                    acc[4 * j4 + 1] = fmaf(a[e], w.y, acc[4 * j4 + 1]);
                    acc[4 * j4 + 2] = fmaf(a[e], w.z, acc[4 * j4 + 2]);
                    acc[4 * j4 + 3] = fmaf(a[e], w.w, acc[4 * j4 + 3]);
                }
            }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < NCP; ++j)
            if (j < NC) mx = fmaxf(mx, acc[j]);
        float se = 0.0f;
#pragma unroll
        for (int j = 0; j < NCP; ++j)
            if (j < NC) se += __expf(acc[j] - mx);
        const float lse = mx + __logf(se);
        float *out = logp + m * NC;
#pragma unroll
        for (int j = 0; j < NCP; ++j)
            if (j < NC) out[j] = acc[j] - lse;
    }
}

// W2s: [NCP][C] fp32 in shared memory; class sums: warp shuffle -> shared -> one fp64 atomic per class and block
template <int NCP>
__global__ void __launch_bounds__(kHeadThreads)
head_tail_bwd_kernel(const float *__restrict__ dlogp, const float *__restrict__ logp, const float *__restrict__ W2,
                     int64_t M, int C, int NC, const int64_t *__restrict__ seed_ptr, uint32_t thresh8, float keep_scale,
                     __nv_bfloat16 *__restrict__ dA, int ldda, __nv_bfloat16 *__restrict__ dlogits_rows, int lddl,
                     double *__restrict__ db2_accum) {
    extern __shared__ float smem[];
    float *w2s = smem;                                   // [C][NCP]: row k holds W2[:, k]
    float *cls = smem + (size_t)C * NCP;                 // [4 warps][NCP]
    for (int i = threadIdx.x; i < C * NCP; i += kHeadThreads) {
        const int k = i / NCP, j = i - k * NCP;
        w2s[i] = j < NC ? W2[(size_t)j * C + k] : 0.0f;
    }
    __syncthreads();
    const uint32_t seed = seed_ptr ? (uint32_t)(*seed_ptr) : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float csum[NCP];
#pragma unroll
    for (int j = 0; j < NCP; ++j) csum[j] = 0.0f;
    for (int64_t m0 = (int64_t)blockIdx.x * kHeadThreads; m0 < M; m0 += (int64_t)gridDim.x * kHeadThreads) {
        const int64_t m = m0 + threadIdx.x;
        if (m >= M) continue;
        float dl[NCP];
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < NCP; ++j) {
            dl[j] = j < NC ? dlogp[m * NC + j] : 0.0f;
            s += dl[j];
        }
#pragma unroll
        for (int j = 0; j < NCP; ++j) {
            dl[j] = j < NC ? dl[j] - __expf(logp[m * NC + j]) * s : 0.0f;
            csum[j] += dl[j];
        }
        if (dlogits_rows) {
            __nv_bfloat16 *o = dlogits_rows + m * lddl;
#pragma unroll
            for (int j = 0; j < NCP; j += 2) {
                if (j < lddl) *reinterpret_cast<__nv_bfloat162 *>(o + j) = __floats2bfloat162_rn(dl[j], dl[j + 1]);
            }
            for (int j = NCP; j < lddl; j += 2) *reinterpret_cast<__nv_bfloat162 *>(o + j) = __floats2bfloat162_rn(0.f, 0.f);
        }
        __nv_bfloat16 *drow = dA + m * ldda;
        for (int k0 = 0; k0 < C; k0 += 8) {
            const uint32_t keep = thresh8 ? keep_bits8(seed, (uint64_t)m * C + k0, thresh8) : 0xffu;
            float d[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4 *wr = reinterpret_cast<const float4 *>(w2s + (size_t)(k0 + e) * NCP);
                float t = 0.0f;
#pragma unroll
                for (int j4 = 0; j4 < NCP / 4; ++j4) {
                    const float4 w = wr[j4];
                    t = fmaf(dl[4 * j4 + 0], w.x, t);
                    t = fmaf(dl[4 * j4 + 1], w.y, t);
                    t = fmaf(dl[4 * j4 + 2], w.z, t);
                    t = fmaf(dl[4 * j4 + 3], w.w, t);
                }
                d[e] = ((keep >> e) & 1u) ? t * keep_scale : 0.0f;
            }
            uint4 o;
            __nv_bfloat162 h;
            h = __floats2bfloat162_rn(d[0], d[1]); o.x = *reinterpret_cast<uint32_t *>(&h);
            h = __floats2bfloat162_rn(d[2], d[3]); o.y = *reinterpret_cast<uint32_t *>(&h);
            h = __floats2bfloat162_rn(d[4], d[5]); o.z = *reinterpret_cast<uint32_t *>(&h);
            h = __floats2bfloat162_rn(d[6], d[7]); o.w = *reinterpret_cast<uint32_t *>(&h);
            *reinterpret_cast<uint4 *>(drow + k0) = o;
        }
    }
#pragma unroll
    for (int j = 0; j < NCP; ++j) {
        float v = csum[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) cls[warp * NCP + j] = v;
    }
    __syncthreads();
    if (threadIdx.x < NC && db2_accum) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kHeadThreads / 32; ++w) t += cls[w * NCP + threadIdx.x];
        atomicAdd(db2_accum + threadIdx.x, (double)t);
    }
}

__global__ void head_db2_finalize_kernel(double *accum, int NC, float *db2) {
    const int j = threadIdx.x;
    if (j < NC) {
        db2[j] = (float)accum[j];
        accum[j] = 0.0;
    }
}

static uint32_t drop_threshold(float p) {
    if (!(p > 0.0f)) return 0u;
    int t = (int)(p * 256.0f + 0.5f);
    return (uint32_t)(t < 1 ? 1 : (t > 255 ? 255 : t));
}

}  // namespace pn2

using namespace pn2;

#define PN2_HEAD_DISPATCH(NCP_RT, ...)                     \
    switch (NCP_RT) {                                      \
        case 4: { constexpr int NCP = 4; __VA_ARGS__; } break;   \
        case 8: { constexpr int NCP = 8; __VA_ARGS__; } break;   \
        case 12: { constexpr int NCP = 12; __VA_ARGS__; } break; \
        case 16: { constexpr int NCP = 16; __VA_ARGS__; } break; \
        case 20: { constexpr int NCP = 20; __VA_ARGS__; } break; \
        case 24: { constexpr int NCP = 24; __VA_ARGS__; } break; \
        case 28: { constexpr int NCP = 28; __VA_ARGS__; } break; \
        default: { constexpr int NCP = 32; __VA_ARGS__; } break; \
    }

extern "C" int pn2_head_tail_fwd(const void *Z, int ldz, const float *scale, const float *shift, const float *W2,
                                 const float *b2, int64_t M, int C, int NC, float drop_p, const int64_t *seed,
                                 float *logp, void *act_out, int ldo, void *stream) {
    PN2_REQUIRE(Z && scale && shift && W2 && logp, "head_tail_fwd: null pointer");
    PN2_REQUIRE(M >= 0 && C >= 8 && C % 8 == 0 && C <= 256 && ldz >= C && ldz % 8 == 0 && NC >= 1 && NC <= 32,
                "head_tail_fwd: bad sizes C=%d (multiple of 8, <= 256) NC=%d (<= 32) ldz=%d", C, NC, ldz);
    PN2_REQUIRE(drop_p >= 0.0f && drop_p < 1.0f && (drop_p == 0.0f || seed), "head_tail_fwd: dropout needs 0 <= p < 1 and a seed");
    PN2_REQUIRE(!act_out || (ldo >= C && ldo % 8 == 0), "head_tail_fwd: bad ldo");
    if (M == 0) return PN2_OK;
    const int ncp = (NC + 3) / 4 * 4;
    const uint32_t thr = drop_threshold(drop_p);
    const float ks = thr ? 256.0f / (float)(256 - (int)thr) : 1.0f;
    const size_t smem = sizeof(float) * ((size_t)C * ncp + ncp + 2 * C);
    const int grid = grid_for(M, kHeadThreads, kNumSMs * 8);
    PN2_HEAD_DISPATCH(ncp, (head_tail_fwd_kernel<NCP><<<grid, kHeadThreads, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)Z, ldz, scale, shift, W2, b2, M, C, NC, seed, thr, ks, logp, (__nv_bfloat16 *)act_out, ldo)));
    count_launch();
    return check_launch("head_tail_fwd");
}

extern "C" int pn2_head_tail_bwd(const float *dlogp, const float *logp, const float *W2, int64_t M, int C, int NC,
                                 float drop_p, const int64_t *seed, void *dA, int ldda, void *dlogits_rows, int lddl,
                                 double *db2_accum, float *db2, void *stream) {
    PN2_REQUIRE(dlogp && logp && W2 && dA && db2_accum && db2, "head_tail_bwd: null pointer");
    PN2_REQUIRE(M >= 0 && C >= 8 && C % 8 == 0 && C <= 256 && ldda >= C && ldda % 8 == 0 && NC >= 1 && NC <= 32,
                "head_tail_bwd: bad sizes");
    PN2_REQUIRE(!dlogits_rows || (lddl >= (NC + 3) / 4 * 4 && lddl % 8 == 0), "head_tail_bwd: bad lddl");
    PN2_REQUIRE(drop_p >= 0.0f && drop_p < 1.0f && (drop_p == 0.0f || seed), "head_tail_bwd: dropout needs 0 <= p < 1 and a seed");
    if (M == 0) return PN2_OK;
    const int ncp = (NC + 3) / 4 * 4;
    const uint32_t thr = drop_threshold(drop_p);
    const float ks = thr ? 256.0f / (float)(256 - (int)thr) : 1.0f;
    const size_t smem = sizeof(float) * ((size_t)C * ncp + 4 * ncp);
    const int grid = grid_for(M, kHeadThreads, kNumSMs * 8);
    PN2_HEAD_DISPATCH(ncp, (head_tail_bwd_kernel<NCP><<<grid, kHeadThreads, smem, (cudaStream_t)stream>>>(
        dlogp, logp, W2, M, C, NC, seed, thr, ks, (__nv_bfloat16 *)dA, ldda, (__nv_bfloat16 *)dlogits_rows, lddl, db2_accum)));
    count_launch();
    head_db2_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(db2_accum, NC, db2);
    count_launch();
    return check_launch("head_tail_bwd");
}
