// linear_simt.cu -- the 1x1-conv MLP layers on fp32 rows, on the fp32 FMA pipes.
//
// This is the full-precision path (PN2_F32 rows): it matches the reference's CPU fp32
// convolutions (/root/reference/models/pointnet2_utils.py:196-198, :311-314) to summation
// order, and is what the tight-tolerance parity tests run.  The bf16 path (linear_tc.cu)
// runs the same contractions on the tcgen05 tensor cores.
//
//   forward      Z[M,N]  = act(X)[M,K] . W[N,K]^T (+bias)      act = prev layer's BN+ReLU on load
//                + per-CTA column sums of Z, Z^2 for train-mode BatchNorm
//   data grad    dX[M,K] = dZ[M,N] . W[N,K]                    (same kernel, W read transposed)
//   weight grad  dW[N,K] = sum_m dZ[m,n] . act(X)[m,k]         split over M, deterministic reduce
//
// All reductions are performed in a fixed order (no atomics), so results are run-to-run
// deterministic.
#include "common.cuh"

namespace pn2 {

constexpr int kLinBM = 128, kLinBK = 16, kLinThreads = 256, kLinTM = 8;
constexpr int kMaxPartials = 4 * kNumSMs;   // up to four resident CTAs per SM for the row-streaming kernels

static int num_partials(int64_t M) {
    int64_t t = (M + kLinBM - 1) / kLinBM;
    if (t < 1) t = 1;
    return (int)(t < kMaxPartials ? t : kMaxPartials);
}

// BN = columns per tile (64 or 32); thread tile = kLinTM x TN with TN = BN/16
template <typename TX, typename TZ, int BN>
__global__ void __launch_bounds__(kLinThreads)
linear_nt_kernel(const TX *__restrict__ X, int ldx, const float *__restrict__ in_scale,
                 const float *__restrict__ in_shift, const float *__restrict__ W, int64_t w_sn, int64_t w_sk,
                 const float *__restrict__ bias, int64_t M, int K, int N, TZ *__restrict__ Z, int ldz,
                 double *__restrict__ stat_accum) {
    constexpr int TN = BN / 16;
    __shared__ __align__(16) float As[kLinBK][kLinBM + 4];
    __shared__ __align__(16) float Bs[kLinBK][BN + 4];
    __shared__ float red[2][16][BN];
    extern __shared__ float tot[];   // [2][N] running column sums of this CTA (stats only)

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m_tiles = (M + kLinBM - 1) / kLinBM;
    const int n_tiles = (N + BN - 1) / BN;
    if (stat_accum)
        for (int i = tid; i < 2 * N; i += kLinThreads) tot[i] = 0.0f;

    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        const int64_t m0 = mt * kLinBM;
        for (int nt = 0; nt < n_tiles; ++nt) {
            const int n0 = nt * BN;
            float acc[kLinTM][TN];
#pragma unroll
            for (int i = 0; i < kLinTM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

            for (int k0 = 0; k0 < K; k0 += kLinBK) {
                __syncthreads();
#pragma unroll
                for (int i = 0; i < (kLinBM * kLinBK) / kLinThreads; ++i) {
                    int e = tid + i * kLinThreads;
                    int k = e % kLinBK, r = e / kLinBK;
                    int64_t m = m0 + r;
                    float v = 0.0f;
                    if (m < M && k0 + k < K) {
                        v = ld_act<TX>(X + m * ldx + k0 + k);
                        if (in_scale) v = fmaxf(fmaf(v, in_scale[k0 + k], in_shift[k0 + k]), 0.0f);
                    }
                    As[k][r] = v;
                }
#pragma unroll
                for (int i = 0; i < (BN * kLinBK) / kLinThreads; ++i) {
                    int e = tid + i * kLinThreads;
                    int k = e % kLinBK, n = e / kLinBK;
                    float v = 0.0f;
                    if (n0 + n < N && k0 + k < K) v = W[(int64_t)(n0 + n) * w_sn + (int64_t)(k0 + k) * w_sk];
                    Bs[k][n] = v;
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < kLinBK; ++k) {
                    float a[kLinTM], bb[TN];
                    const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * kLinTM]);
                    const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][ty * kLinTM + 4]);
                    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                    a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
                    for (int j = 0; j < TN; ++j) bb[j] = Bs[k][tx * TN + j];
#pragma unroll
                    for (int i = 0; i < kLinTM; ++i)
#pragma unroll
                        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
                }
            }

            // ---- epilogue: store (+bias) and column statistics of the bias-free product ----
            float s1[TN], s2[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) s1[j] = s2[j] = 0.0f;
#pragma unroll
            for (int i = 0; i < kLinTM; ++i) {
                int64_t m = m0 + ty * kLinTM + i;
                if (m >= M) continue;
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    int n = n0 + tx * TN + j;
                    if (n >= N) continue;
                    float v = acc[i][j];
                    s1[j] += v;
                    s2[j] = fmaf(v, v, s2[j]);
                    if (bias) v += bias[n];
                    st_act<TZ>(Z + m * ldz + n, v);
                }
            }
            if (stat_accum) {
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    red[0][ty][tx * TN + j] = s1[j];
                    red[1][ty][tx * TN + j] = s2[j];
                }
                __syncthreads();
                if (tid < 2 * BN) {
                    int which = tid / BN, c = tid % BN;
                    if (n0 + c < N) {
                        float s = 0.0f;
#pragma unroll
                        for (int r = 0; r < 16; ++r) s += red[which][r][c];
                        tot[which * N + n0 + c] += s;
                    }
                }
            }
        }
    }
    if (stat_accum) {   // one fp64 atomic per column and CTA: order-independent to ~1e-16, i.e. deterministic in fp32
        __syncthreads();
        for (int i = tid; i < 2 * N; i += kLinThreads)
            atomicAdd(stat_accum + (size_t)(blockIdx.x % kStatReplicas) * 2 * N + i, (double)tot[i]);
    }
}

// dW tile 64(n) x 64(k); grid (k_tiles, n_tiles, splits); scratch[split][N][K]
template <typename TD, typename TX>
__global__ void __launch_bounds__(256)
linear_wgrad_kernel(const TD *__restrict__ dZ, int lddz, const TX *__restrict__ X, int ldx,
                    const float *__restrict__ in_scale, const float *__restrict__ in_shift, int64_t M, int K,
                    int N, int64_t rows_per_split, float *__restrict__ scratch) {
    __shared__ __align__(16) float Ds[16][64 + 4];
    __shared__ __align__(16) float Xs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r_end = min(M, r_begin + rows_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int64_t r0 = r_begin; r0 < r_end; r0 += 16) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = tid + i * 256;
            int c = e & 63, r = e >> 6;
            int64_t m = r0 + r;
            float dv = 0.0f, xv = 0.0f;
            if (m < r_end) {
                if (n0 + c < N) dv = ld_act<TD>(dZ + m * lddz + n0 + c);
                if (k0 + c < K) {
                    xv = ld_act<TX>(X + m * ldx + k0 + c);
                    if (in_scale) xv = fmaxf(fmaf(xv, in_scale[k0 + c], in_shift[k0 + c]), 0.0f);
                }
            }
            Ds[r][c] = dv;
            Xs[r][c] = xv;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float4 a = *reinterpret_cast<const float4 *>(&Ds[r][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Xs[r][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
    float *out = scratch + (int64_t)blockIdx.z * N * K;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k < K) out[(int64_t)n * K + k] = acc[i][j];
        }
    }
}

__global__ void wgrad_reduce_kernel(const float *__restrict__ scratch, int splits, int64_t NK, float *__restrict__ dW) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < NK; e += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.0f;
        for (int z = 0; z < splits; ++z) s += scratch[(int64_t)z * NK + e];
        dW[e] = s;
    }
}

static void wgrad_plan(int64_t M, int K, int N, int &splits, int64_t &rows_per_split) {
    int tiles = ((K + 63) / 64) * ((N + 63) / 64);
    int64_t want = (M + 1023) / 1024;                 // >= 1024 rows per split
    int64_t cap = (4 * kNumSMs + tiles - 1) / tiles;  // about 4 CTAs per SM in total
    int64_t s = want < cap ? want : cap;
    if (s < 1) s = 1;
    rows_per_split = ((M + s - 1) / s + 15) / 16 * 16;
    if (rows_per_split < 16) rows_per_split = 16;
    splits = (int)((M + rows_per_split - 1) / rows_per_split);
    if (splits < 1) splits = 1;
}

template <typename TX, typename TZ>
static int launch_nt(const void *X, int ldx, const float *in_scale, const float *in_shift, const float *W,
                     int64_t w_sn, int64_t w_sk, const float *bias, int64_t M, int K, int N, void *Z, int ldz,
                     double *stat_accum, cudaStream_t st) {
    int grid = num_partials(M);
    size_t dyn = stat_accum ? sizeof(float) * 2 * (size_t)N : 0;
    if (N <= 32)
        linear_nt_kernel<TX, TZ, 32><<<grid, kLinThreads, dyn, st>>>((const TX *)X, ldx, in_scale, in_shift, W, w_sn, w_sk,
                                                                     bias, M, K, N, (TZ *)Z, ldz, stat_accum);
    else
        linear_nt_kernel<TX, TZ, 64><<<grid, kLinThreads, dyn, st>>>((const TX *)X, ldx, in_scale, in_shift, W, w_sn, w_sk,
                                                                     bias, M, K, N, (TZ *)Z, ldz, stat_accum);
    count_launch();
    return check_launch("linear_nt");
}

int simt_linear_nt(const void *X, int ldx, int x_dtype, const float *in_scale, const float *in_shift,
                   const float *W, int64_t w_sn, int64_t w_sk, const float *bias, int64_t M, int K, int N,
                   void *Z, int ldz, int z_dtype, double *stat_accum, cudaStream_t st) {
    if (x_dtype == PN2_F32 && z_dtype == PN2_F32)
        return launch_nt<float, float>(X, ldx, in_scale, in_shift, W, w_sn, w_sk, bias, M, K, N, Z, ldz, stat_accum, st);
    if (x_dtype == PN2_F32 && z_dtype == PN2_BF16)
        return launch_nt<float, __nv_bfloat16>(X, ldx, in_scale, in_shift, W, w_sn, w_sk, bias, M, K, N, Z, ldz, stat_accum, st);
    if (x_dtype == PN2_BF16 && z_dtype == PN2_F32)
        return launch_nt<__nv_bfloat16, float>(X, ldx, in_scale, in_shift, W, w_sn, w_sk, bias, M, K, N, Z, ldz, stat_accum, st);
    return launch_nt<__nv_bfloat16, __nv_bfloat16>(X, ldx, in_scale, in_shift, W, w_sn, w_sk, bias, M, K, N, Z, ldz, stat_accum, st);
}

template <typename TD, typename TX>
static int launch_wgrad(const void *dZ, int lddz, const void *X, int ldx, const float *in_scale,
                        const float *in_shift, int64_t M, int K, int N, float *dW, void *scratch,
                        cudaStream_t st) {
    int splits;
    int64_t rps;
    wgrad_plan(M, K, N, splits, rps);
    dim3 grid((K + 63) / 64, (N + 63) / 64, splits);
    linear_wgrad_kernel<TD, TX><<<grid, 256, 0, st>>>((const TD *)dZ, lddz, (const TX *)X, ldx, in_scale, in_shift, M, K, N,
                                                      rps, (float *)scratch);
    count_launch();
    int rc = check_launch("linear_wgrad");
    if (rc != PN2_OK) return rc;
    int64_t NK = (int64_t)N * K;
    wgrad_reduce_kernel<<<grid_for(NK, 256), 256, 0, st>>>((const float *)scratch, splits, NK, dW);
    count_launch();
    return check_launch("wgrad_reduce");
}

int simt_linear_wgrad(const void *dZ, int lddz, int dz_dtype, const void *X, int ldx, int x_dtype,
                      const float *in_scale, const float *in_shift, int64_t M, int K, int N, float *dW,
                      void *scratch, cudaStream_t st) {
    if (dz_dtype == PN2_F32 && x_dtype == PN2_F32)
        return launch_wgrad<float, float>(dZ, lddz, X, ldx, in_scale, in_shift, M, K, N, dW, scratch, st);
    if (dz_dtype == PN2_F32 && x_dtype == PN2_BF16)
        return launch_wgrad<float, __nv_bfloat16>(dZ, lddz, X, ldx, in_scale, in_shift, M, K, N, dW, scratch, st);
    if (dz_dtype == PN2_BF16 && x_dtype == PN2_F32)
        return launch_wgrad<__nv_bfloat16, float>(dZ, lddz, X, ldx, in_scale, in_shift, M, K, N, dW, scratch, st);
    return launch_wgrad<__nv_bfloat16, __nv_bfloat16>(dZ, lddz, X, ldx, in_scale, in_shift, M, K, N, dW, scratch, st);
}

size_t simt_wgrad_scratch_bytes(int64_t M, int K, int N) {
    int splits;
    int64_t rps;
    wgrad_plan(M, K, N, splits, rps);
    return sizeof(float) * (size_t)splits * (size_t)N * (size_t)K;
}

int linear_num_partials(int64_t M) { return num_partials(M); }

}  // namespace pn2
