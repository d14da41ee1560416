// head.cu -- the tail of the segmentation head on point-major rows, optionally with the loss:
//
//   x = drop1(relu(bn1(conv1(l0_points))));  x = conv2(x);  x = log_softmax(x, dim=1)
//   (/root/reference/models/pointnet2_sem_seg.py:36-39)
//   loss = F.nll_loss(pred, target, weight=weight)                               (pointnet2_sem_seg.py:47-48)
//
// conv1 + bn1 run as one more layer of the fp1 MLP on the tensor-core layer kernels (linear_tc.cu); this file
// is what follows the pre-BatchNorm product Z[M,C] (bf16 rows) of that layer:
//   forward : a = relu(z*scale+shift) -> dropout -> logits = a.W2^T + b2 (C -> NC <= 32 classes) -> log_softmax
//             (-> weighted NLL sums).  One warp owns 16 rows: every lane reads 16 B of two rows per 32-channel chunk,
//             applies BatchNorm / ReLU / dropout in registers and feeds the values to mma.sync m16n8k16 as A fragments
//             (the k order inside a chunk is permuted so that a lane's 8 consecutive channels ARE its fragment
//             elements; the W2 fragments in shared memory are permuted the same way).  a and W2 are split into
//             bf16 hi + lo parts and three products (hi.hi + lo.hi + hi.lo) are accumulated in fp32, so the
//             log-probabilities carry fp32-level error (2^-16 relative per term) although the multiplier is bf16.
//             Optionally stores the dropped activation (bf16 rows) that conv2's weight gradient needs.
//             The first version did this contraction with fp32 FMAs and W2 in shared memory: 640 broadcast LDS.128
//             per point made it shared-memory-bound (73 us for 131 072 points at 6 % of DRAM bandwidth).
//   backward: dlogits = dlogp - exp(logp)*sum(dlogp) (log_softmax) -- or, with the fused loss,
//             dlogits = (dloss * w[t] / sum w) * (exp(logp) - onehot(t)) straight from the targets, no [M,NC] gradient
//             tensor -- then d(a) = (dlogits.W2) * mask/(1-p) with the same split-bf16 mma (K = classes, N = channels),
//             sum of dlogits per class (conv2's bias gradient) through fp64 atomics; dlogits is also written as bf16
//             rows so that conv2's weight gradient is one call of the tensor-core wgrad kernel.
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
// the hash word that decides elements 4*qi .. 4*qi+3 of the [M, C] activation (8 bits each)
__device__ __forceinline__ uint32_t keep_word(uint32_t seed, uint64_t qi) {
    return mix32(mix32((uint32_t)qi ^ seed) + (uint32_t)(qi >> 32) * 0x9e3779b9u + 0x7f4a7c15u);
}
// 8 keep bits for elements e0 .. e0+7 (e0 % 8 == 0)
__device__ __forceinline__ uint32_t keep_bits8(uint32_t seed, uint64_t e0, uint32_t thresh8) {
    uint32_t bits = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t r = keep_word(seed, (e0 >> 2) + h);
#pragma unroll
        for (int i = 0; i < 4; ++i) bits |= (((r >> (8 * i)) & 0xffu) >= thresh8 ? 1u : 0u) << (4 * h + i);
    }
    return bits;
}
// 2 keep bits for elements e0 + 2q, e0 + 2q + 1 (e0 % 8 == 0, q < 4): the same bits keep_bits8 hands out
__device__ __forceinline__ uint32_t keep_bits2(uint32_t seed, uint64_t e0, int q, uint32_t thresh8) {
    const uint32_t r = keep_word(seed, (e0 >> 2) + (uint64_t)(q >> 1));
    const int sh = 16 * (q & 1);
    return (((r >> sh) & 0xffu) >= thresh8 ? 1u : 0u) | (((r >> (sh + 8)) & 0xffu) >= thresh8 ? 2u : 0u);
}

// D += A(16x16, row) * B(16x8, col), bf16 operands, fp32 accumulate.  Fragment layout (g = lane/4, q = lane%4):
//   a[0] = A[g][2q,2q+1]  a[1] = A[g+8][2q,2q+1]  a[2] = A[g][2q+8,2q+9]  a[3] = A[g+8][2q+8,2q+9]
//   b.x = B[2q,2q+1][g]   b.y = B[2q+8,2q+9][g]
//   c[0], c[1] = D[g][2q,2q+1]                    c[2], c[3] = D[g+8][2q,2q+1]
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
// (x, y) -> packed bf16 pair hi (round to nearest) and the bf16 pair of the remainders
__device__ __forceinline__ void split2(float x, float y, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
// hi.hi + lo.hi + hi.lo
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], const uint2 bhi,
                                     const uint2 blo) {
    mma_bf16(c, alo, bhi);
    mma_bf16(c, ahi, blo);
    mma_bf16(c, ahi, bhi);
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

constexpr int kHeadThreads = 128;
constexpr int kHeadWarps = kHeadThreads / 32;

// relu(bn(z)) -> dropout for the 8 channels k0 .. k0+7 of one row; H = bf16 pairs (the stored activation), L = remainders
__device__ __forceinline__ void head_activate8(const uint4 zr, const float *scs, const float *shs, int k0, uint32_t keep,
                                               float keep_scale, uint32_t (&H)[4], uint32_t (&L)[4]) {
    const uint32_t *zw = reinterpret_cast<const uint32_t *>(&zr);
    const float4 s0 = *reinterpret_cast<const float4 *>(scs + k0), s1 = *reinterpret_cast<const float4 *>(scs + k0 + 4);
    const float4 h0 = *reinterpret_cast<const float4 *>(shs + k0), h1 = *reinterpret_cast<const float4 *>(shs + k0 + 4);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&zw[i]));
        float a0 = fmaxf(fmaf(f.x, sc[2 * i], sh[2 * i]), 0.0f);
        float a1 = fmaxf(fmaf(f.y, sc[2 * i + 1], sh[2 * i + 1]), 0.0f);
        a0 = ((keep >> (2 * i)) & 1u) ? a0 * keep_scale : 0.0f;
        a1 = ((keep >> (2 * i + 1)) & 1u) ? a1 * keep_scale : 0.0f;
        split2(a0, a1, H[i], L[i]);
    }
}

// NT: class tiles of 8 (NC <= 8*NT).  Shared memory: W2 fragments hi/lo [C/16][NT][32] uint2, scale[C], shift[C], b2[8*NT].
// target != NULL: also accumulates sum w[t]*logp[t] and sum w[t] into loss_accum[0..1] (fp64 atomics, one pair per block).
template <int NT>
__global__ void __launch_bounds__(kHeadThreads)
head_tail_fwd_kernel(const __nv_bfloat16 *__restrict__ Z, int ldz, const float *__restrict__ scale,
                     const float *__restrict__ shift, const float *__restrict__ W2, const float *__restrict__ b2,
                     int64_t M, int C, int NC, const int64_t *__restrict__ seed_ptr, uint32_t thresh8, float keep_scale,
                     float *__restrict__ logp, __nv_bfloat16 *__restrict__ act_out, int ldo,
                     const int64_t *__restrict__ target, const float *__restrict__ class_weight,
                     double *__restrict__ loss_accum, int64_t *__restrict__ labels) {
    extern __shared__ __align__(16) unsigned char head_smem[];
    const int KS = C >> 4, nfrag = KS * NT * 32;
    uint2 *bhi = reinterpret_cast<uint2 *>(head_smem), *blo = bhi + nfrag;
    float *scs = reinterpret_cast<float *>(blo + nfrag), *shs = scs + C, *b2s = shs + C, *red = b2s + 8 * NT;
    for (int i = threadIdx.x; i < nfrag; i += kHeadThreads) {
        const int ln = i & 31, st = i >> 5, t = st % NT, s = st / NT;
        const int n = 8 * t + (ln >> 2), kk = 32 * (s >> 1) + 8 * (ln & 3) + 4 * (s & 1);
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < NC) w = *reinterpret_cast<const float4 *>(W2 + (size_t)n * C + kk);
        uint2 h, l;
        split2(w.x, w.y, h.x, l.x);
        split2(w.z, w.w, h.y, l.y);
        bhi[i] = h;
        blo[i] = l;
    }
    for (int i = threadIdx.x; i < C; i += kHeadThreads) {
        scs[i] = scale[i];
        shs[i] = shift[i];
    }
    for (int i = threadIdx.x; i < 8 * NT; i += kHeadThreads) b2s[i] = (i < NC && b2) ? b2[i] : 0.0f;
    __syncthreads();
    const uint32_t seed = seed_ptr ? (uint32_t)(*seed_ptr) : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const int nchunk = C >> 5;
    float num = 0.0f, den = 0.0f;
    const int64_t stride = (int64_t)gridDim.x * kHeadWarps * 16;
    for (int64_t base = ((int64_t)blockIdx.x * kHeadWarps + warp) * 16; base < M; base += stride) {
        const int64_t r0 = base + g, r1 = r0 + 8;
        const bool v0 = r0 < M, v1 = r1 < M;
        const __nv_bfloat16 *z0p = Z + (v0 ? r0 : M - 1) * ldz + 8 * q, *z1p = Z + (v1 ? r1 : M - 1) * ldz + 8 * q;
        float acc[NT][4];
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.0f;
        uint4 z0 = *reinterpret_cast<const uint4 *>(z0p), z1 = *reinterpret_cast<const uint4 *>(z1p);
        for (int c = 0; c < nchunk; ++c) {
            const uint4 c0 = z0, c1 = z1;
            if (c + 1 < nchunk) {          // next chunk in flight while this one is used
                z0 = *reinterpret_cast<const uint4 *>(z0p + 32 * (c + 1));
                z1 = *reinterpret_cast<const uint4 *>(z1p + 32 * (c + 1));
            }
            const int k0 = 32 * c + 8 * q;
            uint32_t keep0 = 0xffu, keep1 = 0xffu;
            if (thresh8) {
                keep0 = keep_bits8(seed, (uint64_t)r0 * C + k0, thresh8);
                keep1 = keep_bits8(seed, (uint64_t)r1 * C + k0, thresh8);
            }
            uint32_t H0[4], L0[4], H1[4], L1[4];
            head_activate8(c0, scs, shs, k0, keep0, keep_scale, H0, L0);
            head_activate8(c1, scs, shs, k0, keep1, keep_scale, H1, L1);
            if (act_out) {
                if (v0) *reinterpret_cast<uint4 *>(act_out + r0 * ldo + k0) = make_uint4(H0[0], H0[1], H0[2], H0[3]);
                if (v1) *reinterpret_cast<uint4 *>(act_out + r1 * ldo + k0) = make_uint4(H1[0], H1[1], H1[2], H1[3]);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t ahi[4] = {H0[2 * h], H1[2 * h], H0[2 * h + 1], H1[2 * h + 1]};
                const uint32_t alo[4] = {L0[2 * h], L1[2 * h], L0[2 * h + 1], L1[2 * h + 1]};
                const int fb = ((2 * c + h) * NT) * 32 + lane;
#pragma unroll
                for (int t = 0; t < NT; ++t) mma3(acc[t], ahi, alo, bhi[fb + 32 * t], blo[fb + 32 * t]);
            }
        }
        // log_softmax over the row: a row's classes sit in the 4 lanes of a quad
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = 8 * t + 2 * q + j;
                acc[t][j] += b2s[col];
                acc[t][2 + j] += b2s[col];
                if (col < NC) {
                    mx0 = fmaxf(mx0, acc[t][j]);
                    mx1 = fmaxf(mx1, acc[t][2 + j]);
                }
            }
        }
        mx0 = quad_max(mx0);
        mx1 = quad_max(mx1);
        float se0 = 0.0f, se1 = 0.0f;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (8 * t + 2 * q + j < NC) {
                    se0 += __expf(acc[t][j] - mx0);
                    se1 += __expf(acc[t][2 + j] - mx1);
                }
            }
        }
        const float lse0 = mx0 + __logf(quad_sum(se0)), lse1 = mx1 + __logf(quad_sum(se1));
        int t0 = -1, t1 = -1;
        float w0 = 0.0f, w1 = 0.0f;
        if (target) {
            if (v0) {
                const int64_t tt = target[r0];
                if (tt >= 0 && tt < NC) {
                    t0 = (int)tt;
                    w0 = class_weight ? class_weight[t0] : 1.0f;
                }
            }
            if (v1) {
                const int64_t tt = target[r1];
                if (tt >= 0 && tt < NC) {
                    t1 = (int)tt;
                    w1 = class_weight ? class_weight[t1] : 1.0f;
                }
            }
            if (q == 0) den += w0 + w1;
        }
        float bv0 = -INFINITY, bv1 = -INFINITY;       // arg-max of the STORED log-probabilities, first maximum (torch.argmax)
        int bc0 = 0, bc1 = 0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = 8 * t + 2 * q + j;
                if (col < NC) {
                    const float l0 = acc[t][j] - lse0, l1 = acc[t][2 + j] - lse1;
                    if (v0) logp[r0 * NC + col] = l0;
                    if (v1) logp[r1 * NC + col] = l1;
                    if (col == t0) num = fmaf(w0, l0, num);
                    if (col == t1) num = fmaf(w1, l1, num);
                    if (l0 > bv0) { bv0 = l0; bc0 = col; }
                    if (l1 > bv1) { bv1 = l1; bc1 = col; }
                }
            }
        }
        if (labels) {          // uniform
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const float ov0 = __shfl_xor_sync(0xffffffffu, bv0, o), ov1 = __shfl_xor_sync(0xffffffffu, bv1, o);
                const int oc0 = __shfl_xor_sync(0xffffffffu, bc0, o), oc1 = __shfl_xor_sync(0xffffffffu, bc1, o);
                if (ov0 > bv0 || (ov0 == bv0 && oc0 < bc0)) { bv0 = ov0; bc0 = oc0; }
                if (ov1 > bv1 || (ov1 == bv1 && oc1 < bc1)) { bv1 = ov1; bc1 = oc1; }
            }
            if (q == 0 && v0) labels[r0] = bc0;
            if (q == 0 && v1) labels[r1] = bc1;
        }
    }
    if (target) {       // uniform across the block
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            num += __shfl_xor_sync(0xffffffffu, num, o);
            den += __shfl_xor_sync(0xffffffffu, den, o);
        }
        if (lane == 0) {
            red[2 * warp] = num;
            red[2 * warp + 1] = den;
        }
        __syncthreads();
        if (threadIdx.x < 2) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kHeadWarps; ++w) s += (double)red[2 * w + threadIdx.x];
            atomicAdd(loss_accum + threadIdx.x, s);
        }
    }
}

// loss = -(sum w[t] logp[t]) / (sum w[t]);  out = {loss, sum w[t]};  the accumulator is left zero
__global__ void head_loss_finalize_kernel(double *accum, float *out) {
    if (threadIdx.x == 0) {
        const double num = accum[0], den = accum[1];
        out[0] = (float)(-num / den);
        out[1] = (float)den;
        accum[0] = 0.0;
        accum[1] = 0.0;
    }
}

// KSB: class k-steps of 16 (NC <= 16*KSB).  Shared memory: W2 fragments hi/lo [KSB][C/8][32] uint2, class sums [warps][16*KSB].
// dlogp != NULL: dense incoming gradient.  Otherwise the fused loss: targets, class weights, loss_out[1] = sum of the
// target weights, *dloss (NULL = 1) the gradient of the loss.
template <int KSB>
__global__ void __launch_bounds__(kHeadThreads)
head_tail_bwd_kernel(const float *__restrict__ dlogp, const float *__restrict__ logp, const float *__restrict__ W2,
                     int64_t M, int C, int NC, const int64_t *__restrict__ seed_ptr, uint32_t thresh8, float keep_scale,
                     __nv_bfloat16 *__restrict__ dA, int ldda, __nv_bfloat16 *__restrict__ dlogits_rows, int lddl,
                     double *__restrict__ db2_accum, const int64_t *__restrict__ target,
                     const float *__restrict__ class_weight, const float *__restrict__ loss_out,
                     const float *__restrict__ dloss) {
    extern __shared__ __align__(16) unsigned char head_smem[];
    const int NTC = C >> 3, nfrag = KSB * NTC * 32;
    uint2 *bhi = reinterpret_cast<uint2 *>(head_smem), *blo = bhi + nfrag;
    float *cls = reinterpret_cast<float *>(blo + nfrag);            // [warps][16*KSB]
    for (int i = threadIdx.x; i < nfrag; i += kHeadThreads) {
        const int ln = i & 31, st = i >> 5, t = st % NTC, s = st / NTC;
        const int n = 8 * t + (ln >> 2), k0 = 16 * s + 2 * (ln & 3);
        float w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + (j & 1) + 8 * (j >> 1);
            w[j] = k < NC ? W2[(size_t)k * C + n] : 0.0f;
        }
        uint2 h, l;
        split2(w[0], w[1], h.x, l.x);
        split2(w[2], w[3], h.y, l.y);
        bhi[i] = h;
        blo[i] = l;
    }
    __syncthreads();
    const uint32_t seed = seed_ptr ? (uint32_t)(*seed_ptr) : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    const float gscale = dlogp ? 0.0f : (dloss ? *dloss : 1.0f) / loss_out[1];
    float csum[KSB][4];
#pragma unroll
    for (int s = 0; s < KSB; ++s) csum[s][0] = csum[s][1] = csum[s][2] = csum[s][3] = 0.0f;
    const int64_t stride = (int64_t)gridDim.x * kHeadWarps * 16;
    for (int64_t base = ((int64_t)blockIdx.x * kHeadWarps + warp) * 16; base < M; base += stride) {
        const int64_t r0 = base + g, r1 = r0 + 8;
        const bool v0 = r0 < M, v1 = r1 < M;
        // slot j of k-step s: class 16s + 2q + (j&1) + 8*(j>>1)
        float p0[KSB][4], p1[KSB][4];          // exp(logp), later dlogits
        float d0[KSB][4], d1[KSB][4];
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int s = 0; s < KSB; ++s) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = 16 * s + 2 * q + (j & 1) + 8 * (j >> 1);
                const bool in = col < NC;
                p0[s][j] = (in && v0) ? __expf(logp[r0 * NC + col]) : 0.0f;
                p1[s][j] = (in && v1) ? __expf(logp[r1 * NC + col]) : 0.0f;
                d0[s][j] = (dlogp && in && v0) ? dlogp[r0 * NC + col] : 0.0f;
                d1[s][j] = (dlogp && in && v1) ? dlogp[r1 * NC + col] : 0.0f;
                s0 += d0[s][j];
                s1 += d1[s][j];
            }
        }
        int t0 = -1, t1 = -1;
        float coef0 = 0.0f, coef1 = 0.0f;
        if (dlogp) {
            s0 = quad_sum(s0);
            s1 = quad_sum(s1);
        } else {
            if (v0) {
                const int64_t tt = target[r0];
                if (tt >= 0 && tt < NC) {
                    t0 = (int)tt;
                    coef0 = (class_weight ? class_weight[t0] : 1.0f) * gscale;
                }
            }
            if (v1) {
                const int64_t tt = target[r1];
                if (tt >= 0 && tt < NC) {
                    t1 = (int)tt;
                    coef1 = (class_weight ? class_weight[t1] : 1.0f) * gscale;
                }
            }
        }
        uint32_t ahi[KSB][4], alo[KSB][4];
#pragma unroll
        for (int s = 0; s < KSB; ++s) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = 16 * s + 2 * q + (j & 1) + 8 * (j >> 1);
                if (dlogp) {
                    p0[s][j] = d0[s][j] - p0[s][j] * s0;
                    p1[s][j] = d1[s][j] - p1[s][j] * s1;
                } else {
                    p0[s][j] = (t0 >= 0) ? coef0 * (p0[s][j] - (col == t0 ? 1.0f : 0.0f)) : 0.0f;
                    p1[s][j] = (t1 >= 0) ? coef1 * (p1[s][j] - (col == t1 ? 1.0f : 0.0f)) : 0.0f;
                }
                if (col >= NC) p0[s][j] = p1[s][j] = 0.0f;
                csum[s][j] += p0[s][j] + p1[s][j];
            }
            split2(p0[s][0], p0[s][1], ahi[s][0], alo[s][0]);
            split2(p1[s][0], p1[s][1], ahi[s][1], alo[s][1]);
            split2(p0[s][2], p0[s][3], ahi[s][2], alo[s][2]);
            split2(p1[s][2], p1[s][3], ahi[s][3], alo[s][3]);
            if (dlogits_rows) {
                const int ca = 16 * s + 2 * q, cb = ca + 8;
                if (v0 && ca < lddl) *reinterpret_cast<uint32_t *>(dlogits_rows + r0 * lddl + ca) = ahi[s][0];
                if (v1 && ca < lddl) *reinterpret_cast<uint32_t *>(dlogits_rows + r1 * lddl + ca) = ahi[s][1];
                if (v0 && cb < lddl) *reinterpret_cast<uint32_t *>(dlogits_rows + r0 * lddl + cb) = ahi[s][2];
                if (v1 && cb < lddl) *reinterpret_cast<uint32_t *>(dlogits_rows + r1 * lddl + cb) = ahi[s][3];
            }
        }
        if (dlogits_rows && v0)
            for (int c = 16 * KSB + 2 * q; c < lddl; c += 8) *reinterpret_cast<uint32_t *>(dlogits_rows + r0 * lddl + c) = 0u;
        if (dlogits_rows && v1)
            for (int c = 16 * KSB + 2 * q; c < lddl; c += 8) *reinterpret_cast<uint32_t *>(dlogits_rows + r1 * lddl + c) = 0u;
        __nv_bfloat16 *o0 = dA + r0 * ldda + 2 * q, *o1 = dA + r1 * ldda + 2 * q;
#pragma unroll 4
        for (int t = 0; t < NTC; ++t) {
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int s = 0; s < KSB; ++s) {
                const int fb = (s * NTC + t) * 32 + lane;
                mma3(acc, ahi[s], alo[s], bhi[fb], blo[fb]);
            }
            uint32_t k0 = 3u, k1 = 3u;
            if (thresh8) {
                k0 = keep_bits2(seed, (uint64_t)r0 * C + 8 * t, q, thresh8);
                k1 = keep_bits2(seed, (uint64_t)r1 * C + 8 * t, q, thresh8);
            }
            const __nv_bfloat162 e0 = __floats2bfloat162_rn((k0 & 1u) ? acc[0] * keep_scale : 0.0f, (k0 & 2u) ? acc[1] * keep_scale : 0.0f);
            const __nv_bfloat162 e1 = __floats2bfloat162_rn((k1 & 1u) ? acc[2] * keep_scale : 0.0f, (k1 & 2u) ? acc[3] * keep_scale : 0.0f);
            if (v0) *reinterpret_cast<__nv_bfloat162 *>(o0 + 8 * t) = e0;
            if (v1) *reinterpret_cast<__nv_bfloat162 *>(o1 + 8 * t) = e1;
        }
    }
    // class sums: lanes with the same q hold the same classes -> reduce over g, then over the block's warps
#pragma unroll
    for (int s = 0; s < KSB; ++s) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = csum[s][j];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) cls[warp * 16 * KSB + 16 * s + 2 * q + (j & 1) + 8 * (j >> 1)] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x < NC && db2_accum) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kHeadWarps; ++w) t += cls[w * 16 * KSB + threadIdx.x];
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

// persistent blocks: as many as are resident at once (registers / shared memory decide), each warp strides over row groups
template <typename Kernel>
static int head_grid(Kernel kernel, size_t smem, int64_t M) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kHeadThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int64_t groups = (M + 15) / 16, blocks = (groups + kHeadWarps - 1) / kHeadWarps;
    const int64_t cap = (int64_t)kNumSMs * per_sm;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

static int head_fwd_launch(const void *Z, int ldz, const float *scale, const float *shift, const float *W2, const float *b2,
                           int64_t M, int C, int NC, float drop_p, const int64_t *seed, float *logp, void *act_out, int ldo,
                           const int64_t *target, const float *class_weight, double *loss_accum, int64_t *labels,
                           cudaStream_t stream) {
    const int nt = (NC + 7) / 8;
    const uint32_t thr = drop_threshold(drop_p);
    const float ks = thr ? 256.0f / (float)(256 - (int)thr) : 1.0f;
    const size_t smem = 2 * sizeof(uint2) * (size_t)(C / 16) * nt * 32 + sizeof(float) * (2 * (size_t)C + 8 * nt + 2 * kHeadWarps);
#define PN2_HEAD_FWD(NT)                                                                                                \
    head_tail_fwd_kernel<NT><<<head_grid(head_tail_fwd_kernel<NT>, smem, M), kHeadThreads, smem, stream>>>((const __nv_bfloat16 *)Z, ldz, scale, shift, W2, b2, M, C, \
                                                                   NC, seed, thr, ks, logp, (__nv_bfloat16 *)act_out, ldo,  \
                                                                   target, class_weight, loss_accum, labels)
    switch (nt) {
        case 1: PN2_HEAD_FWD(1); break;
        case 2: PN2_HEAD_FWD(2); break;
        case 3: PN2_HEAD_FWD(3); break;
        default: PN2_HEAD_FWD(4); break;
    }
#undef PN2_HEAD_FWD
    count_launch();
    return PN2_OK;
}

static int head_bwd_launch(const float *dlogp, const float *logp, const float *W2, int64_t M, int C, int NC, float drop_p,
                           const int64_t *seed, void *dA, int ldda, void *dlogits_rows, int lddl, double *db2_accum,
                           float *db2, const int64_t *target, const float *class_weight, const float *loss_out,
                           const float *dloss, cudaStream_t stream) {
    const int ksb = (NC + 15) / 16;
    const uint32_t thr = drop_threshold(drop_p);
    const float ks = thr ? 256.0f / (float)(256 - (int)thr) : 1.0f;
    const size_t smem = 2 * sizeof(uint2) * (size_t)ksb * (C / 8) * 32 + sizeof(float) * (size_t)kHeadWarps * 16 * ksb;
    if (ksb == 1)
        head_tail_bwd_kernel<1><<<head_grid(head_tail_bwd_kernel<1>, smem, M), kHeadThreads, smem, stream>>>(dlogp, logp, W2, M, C, NC, seed, thr, ks, (__nv_bfloat16 *)dA,
                                                                      ldda, (__nv_bfloat16 *)dlogits_rows, lddl, db2_accum, target,
                                                                      class_weight, loss_out, dloss);
    else
        head_tail_bwd_kernel<2><<<head_grid(head_tail_bwd_kernel<2>, smem, M), kHeadThreads, smem, stream>>>(dlogp, logp, W2, M, C, NC, seed, thr, ks, (__nv_bfloat16 *)dA,
                                                                      ldda, (__nv_bfloat16 *)dlogits_rows, lddl, db2_accum, target,
                                                                      class_weight, loss_out, dloss);
    count_launch();
    head_db2_finalize_kernel<<<1, 32, 0, stream>>>(db2_accum, NC, db2);
    count_launch();
    return PN2_OK;
}

}  // namespace pn2

using namespace pn2;

#define PN2_HEAD_FWD_CHECKS(what)                                                                                         \
    PN2_REQUIRE(Z && scale && shift && W2 && logp, what ": null pointer");                                                \
    PN2_REQUIRE(M >= 0 && C >= 32 && C % 32 == 0 && C <= 256 && ldz >= C && ldz % 8 == 0 && NC >= 1 && NC <= 32,         \
                what ": bad sizes C=%d (multiple of 32, <= 256) NC=%d (<= 32) ldz=%d", C, NC, ldz);                      \
    PN2_REQUIRE(drop_p >= 0.0f && drop_p < 1.0f && (drop_p == 0.0f || seed), what ": dropout needs 0 <= p < 1 and a seed"); \
    PN2_REQUIRE(!act_out || (ldo >= C && ldo % 8 == 0), what ": bad ldo")

#define PN2_HEAD_BWD_CHECKS(what)                                                                                         \
    PN2_REQUIRE(logp && W2 && dA && db2_accum && db2, what ": null pointer");                                             \
    PN2_REQUIRE(M >= 0 && C >= 32 && C % 32 == 0 && C <= 256 && ldda >= C && ldda % 8 == 0 && NC >= 1 && NC <= 32,       \
                what ": bad sizes");                                                                                      \
    PN2_REQUIRE(!dlogits_rows || (lddl >= (NC + 3) / 4 * 4 && lddl % 8 == 0), what ": bad lddl");                         \
    PN2_REQUIRE(drop_p >= 0.0f && drop_p < 1.0f && (drop_p == 0.0f || seed), what ": dropout needs 0 <= p < 1 and a seed")

extern "C" int pn2_head_tail_fwd(const void *Z, int ldz, const float *scale, const float *shift, const float *W2,
                                 const float *b2, int64_t M, int C, int NC, float drop_p, const int64_t *seed,
                                 float *logp, void *act_out, int ldo, int64_t *labels, void *stream) {
    PN2_HEAD_FWD_CHECKS("head_tail_fwd");
    if (M == 0) return PN2_OK;
    head_fwd_launch(Z, ldz, scale, shift, W2, b2, M, C, NC, drop_p, seed, logp, act_out, ldo, nullptr, nullptr, nullptr,
                    labels, (cudaStream_t)stream);
    return check_launch("head_tail_fwd");
}

extern "C" int pn2_head_tail_bwd(const float *dlogp, const float *logp, const float *W2, int64_t M, int C, int NC,
                                 float drop_p, const int64_t *seed, void *dA, int ldda, void *dlogits_rows, int lddl,
                                 double *db2_accum, float *db2, void *stream) {
    PN2_REQUIRE(dlogp, "head_tail_bwd: null pointer");
    PN2_HEAD_BWD_CHECKS("head_tail_bwd");
    if (M == 0) return PN2_OK;
    head_bwd_launch(dlogp, logp, W2, M, C, NC, drop_p, seed, dA, ldda, dlogits_rows, lddl, db2_accum, db2, nullptr, nullptr,
                    nullptr, nullptr, (cudaStream_t)stream);
    return check_launch("head_tail_bwd");
}

extern "C" int pn2_head_tail_loss_fwd(const void *Z, int ldz, const float *scale, const float *shift, const float *W2,
                                      const float *b2, int64_t M, int C, int NC, float drop_p, const int64_t *seed,
                                      const int64_t *target, const float *class_weight, float *logp, void *act_out, int ldo,
                                      double *loss_accum, float *loss_out, void *stream) {
    PN2_HEAD_FWD_CHECKS("head_tail_loss_fwd");
    PN2_REQUIRE(target && loss_accum && loss_out, "head_tail_loss_fwd: null pointer");
    if (M > 0)
        head_fwd_launch(Z, ldz, scale, shift, W2, b2, M, C, NC, drop_p, seed, logp, act_out, ldo, target, class_weight,
                        loss_accum, nullptr, (cudaStream_t)stream);
    head_loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss_accum, loss_out);
    count_launch();
    return check_launch("head_tail_loss_fwd");
}

extern "C" int pn2_head_tail_loss_bwd(const float *logp, const int64_t *target, const float *class_weight,
                                      const float *loss_out, const float *dloss, const float *W2, int64_t M, int C, int NC,
                                      float drop_p, const int64_t *seed, void *dA, int ldda, void *dlogits_rows, int lddl,
                                      double *db2_accum, float *db2, void *stream) {
    PN2_REQUIRE(target && loss_out, "head_tail_loss_bwd: null pointer");
    PN2_HEAD_BWD_CHECKS("head_tail_loss_bwd");
    if (M == 0) return PN2_OK;
    head_bwd_launch(nullptr, logp, W2, M, C, NC, drop_p, seed, dA, ldda, dlogits_rows, lddl, db2_accum, db2, target,
                    class_weight, loss_out, dloss, (cudaStream_t)stream);
    return check_launch("head_tail_loss_bwd");
}
