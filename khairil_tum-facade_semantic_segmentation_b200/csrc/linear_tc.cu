// linear_tc.cu -- K3: the 1x1-conv MLP layers on bf16 rows as tcgen05 (5th-gen tensor core) GEMMs.
//
//   Z[M,N] = act(X)[M,K] . W[N,K]^T (+bias)        forward  (pointnet2_utils.py:196-198, :311-314)
//   dX[M,K] = dZ[M,N] . W[N,K]                     data gradient (same kernel, W packed transposed)
//
// Shape of the problem: M is huge (up to 2^20 rows = clouds x centroids x samples), K and N are
// tiny (12..768), so every layer is a stream over rows: HBM-bound, not tensor-bound.  The
// kernel is therefore organised around the row stream:
//   * one CTA = 128 threads owns 128-row tiles (UMMA M = 128, cta_group::1), persistent over tiles,
//     several CTAs per SM so that one CTA's loads overlap another's MMA/epilogue;
//   * A operand: rows are read with coalesced 16-byte loads, the previous layer's BatchNorm+ReLU
//     is applied in registers (scale/shift/relu, fp32), the result is rounded to bf16 and written
//     into shared memory in the canonical K-major 128-byte-swizzle UMMA layout;
//   * B operand: the weight is pre-packed once per call (fp32 -> bf16, already swizzled, one
//     image per 64-wide K chunk) and fetched with cp.async.bulk (TMA) onto an mbarrier; when K
//     fits one chunk it stays resident in shared memory for the CTA's whole life;
//   * D accumulates in tensor memory (fp32, N columns); one elected thread issues tcgen05.mma and
//     commits to an mbarrier; chunks are double buffered so loads overlap the MMA;
//   * epilogue: tcgen05.ld -> (+bias) -> bf16 -> staging tile in shared memory -> coalesced 16-byte
//     stores; the train-mode BatchNorm column sums (sum z, sum z^2 of the STORED values) are
//     accumulated from the staging tile with lanes walking columns (bank-conflict free) and
//     kept in registers across tiles; added once per CTA into fp64 accumulators at the end.
#include "common.cuh"
#include "tc_common.cuh"

namespace pn2 {

using namespace tc;

constexpr int kTcThreads = 128;
constexpr int kTcBM = 128;
constexpr int kTcBK = 64;                       // bf16 elements per 128-byte swizzle row
constexpr int kTcAStage = kTcBM * kTcBK * 2;    // 16 KB

// ---- weight packing: fp32 W (any strides) -> bf16 chunk images in UMMA K-major SW128 layout ----
// image kc: n_pad rows x 128 bytes; element (n, kc*64 + c*8 + e) at sw128_offset(n, c) + 2e.
__global__ void pack_weight_kernel(const float *__restrict__ W, int64_t w_sn, int64_t w_sk, int N, int K, int n_pad,
                                   int KC, uint8_t *__restrict__ img) {
    const int total = KC * n_pad * 8;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const int c = q & 7, n = (q >> 3) % n_pad, kc = (q >> 3) / n_pad;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = kc * kTcBK + c * 8 + e;
            v[e] = (n < N && k < K) ? W[(int64_t)n * w_sn + (int64_t)k * w_sk] : 0.0f;
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]);
        o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]);
        o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4 *>(img + (size_t)kc * n_pad * 128 + sw128_offset(n, c)) = o;
    }
}

struct TcLinearArgs {
    const __nv_bfloat16 *X;
    int ldx;
    const float *in_scale, *in_shift;
    const uint8_t *Wimg;
    const float *bias;
    int64_t M;
    int K, N, n_pad, n_store, KC;
    __nv_bfloat16 *Z;
    int ldz;
    double *stat_accum;     // [2][stat_ld] fp64 accumulators, this call adds into columns stat_off .. stat_off+N
    int stat_ld, stat_off;
    int w_resident;
};

template <bool STATS>
__global__ void __launch_bounds__(kTcThreads) linear_tc_kernel(const TcLinearArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_mma[2], bar_w[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_part[4][2][256];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned (SW128 atoms)
    const uint32_t b_bytes = (uint32_t)a.n_pad * 128u;
    uint8_t *const A_st[2] = {smem, smem + kTcAStage};
    // B stage 1 first so that a resident W (stage 0) sits behind everything the epilogue staging may use
    uint8_t *const B_st[2] = {smem + 2 * kTcAStage + b_bytes, smem + 2 * kTcAStage};
    uint8_t *const staging = smem;
    const int st_stride = a.n_store * 2 + 16;

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < a.n_pad) tmem_cols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    if (tid == 0) {
        mbar_init(&bar_mma[0], 1);
        mbar_init(&bar_mma[1], 1);
        mbar_init(&bar_w[0], 1);
        mbar_init(&bar_w[1], 1);
        mbar_init_fence();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_bf16(kTcBM, a.n_pad, 0, 0);

    uint32_t par_mma[2] = {0, 0}, par_w[2] = {0, 0};
    if (a.w_resident) {
        if (tid == 0) {
            mbar_expect_tx(&bar_w[0], b_bytes);
            bulk_g2s(B_st[0], a.Wimg, b_bytes, &bar_w[0]);
        }
        mbar_wait(&bar_w[0], 0);
        par_w[0] ^= 1;
    }

    float sum1[8], sum2[8];   // STATS: this lane's column pairs p = lane + 32 j  (columns 2p, 2p+1)
#pragma unroll
    for (int i = 0; i < 8; ++i) sum1[i] = sum2[i] = 0.0f;

    const int64_t m_tiles = (a.M + kTcBM - 1) / kTcBM;
    const int c16 = tid & 7;          // this thread's 16-byte column chunk inside a K chunk
    const int r_base = tid >> 3;      // rows r_base + 16 i

    for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
        const int64_t m0 = tile * kTcBM;
        for (int kc = 0; kc < a.KC; ++kc) {
            const int s = kc & 1;
            if (kc >= 2) {   // the MMA of chunk kc-2 must have drained this stage
                mbar_wait(&bar_mma[s], par_mma[s]);
                par_mma[s] ^= 1;
            }
            // ---- B chunk: TMA bulk copy of the pre-swizzled image ----
            if (!a.w_resident && tid == 0) {
                mbar_expect_tx(&bar_w[s], b_bytes);
                bulk_g2s(B_st[s], a.Wimg + (size_t)kc * b_bytes, b_bytes, &bar_w[s]);
            }
            // ---- A chunk: coalesced 16-byte loads, BN+ReLU of the previous layer, swizzled store ----
            const int k = kc * kTcBK + c16 * 8;
            float sc[8], sh[8];
            if (a.in_scale) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const bool ok = k + e < a.K;
                    sc[e] = ok ? a.in_scale[k + e] : 0.0f;
                    sh[e] = ok ? a.in_shift[k + e] : 0.0f;
                }
            }
            uint4 raw[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t m = m0 + r_base + 16 * i;
                raw[i] = make_uint4(0u, 0u, 0u, 0u);
                if (m < a.M && k < a.ldx) raw[i] = *reinterpret_cast<const uint4 *>(a.X + m * a.ldx + k);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint4 v = raw[i];
                if (a.in_scale) {
                    const bool row_ok = (m0 + r_base + 16 * i) < a.M;
                    uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
                    for (int e2 = 0; e2 < 4; ++e2) {
                        float2 f = unpack_bf16x2(w[e2]);
                        f.x = fmaxf(fmaf(f.x, sc[2 * e2], sh[2 * e2]), 0.0f);
                        f.y = fmaxf(fmaf(f.y, sc[2 * e2 + 1], sh[2 * e2 + 1]), 0.0f);
                        w[e2] = row_ok ? pack_bf16x2(f.x, f.y) : 0u;
                    }
                }
                *reinterpret_cast<uint4 *>(A_st[s] + sw128_offset(r_base + 16 * i, c16)) = v;
            }
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                if (!a.w_resident) mbar_wait(&bar_w[s], par_w[s]);
                fence_after_sync();
                const int k_left = a.K - kc * kTcBK;
                const int nk = k_left >= kTcBK ? 4 : (k_left + 15) / 16;
                const uint32_t a_base = smem_addr(A_st[s]);
                const uint32_t b_base = smem_addr(a.w_resident ? B_st[0] : B_st[s]);
                for (int j = 0; j < nk; ++j)
                    umma_bf16(tmem, make_desc(a_base + 32 * j, 0, 1024), make_desc(b_base + 32 * j, 0, 1024), idesc,
                              (uint32_t)((kc | j) != 0));
                umma_commit(&bar_mma[s]);
            }
            if (!a.w_resident) par_w[s] ^= 1;
        }
        // ---- wait for the accumulator (consume the outstanding commits in order) ----
        if (a.KC >= 2) {
            const int s2 = (a.KC - 2) & 1;
            mbar_wait(&bar_mma[s2], par_mma[s2]);
            par_mma[s2] ^= 1;
        }
        {
            const int s1 = (a.KC - 1) & 1;
            mbar_wait(&bar_mma[s1], par_mma[s1]);
            par_mma[s1] ^= 1;
        }
        fence_after_sync();

        // ---- epilogue 1: TMEM -> registers -> (+bias) -> bf16 -> staging row `tid` ----
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        uint8_t *my_row = staging + (size_t)tid * st_stride;
        for (int c0 = 0; c0 < a.n_store; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + c0, v);
            if (a.bias) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i < a.N) v[i] += a.bias[c0 + i];
            }
            uint4 lo, hi;
            lo.x = pack_bf16x2(v[0], v[1]);   lo.y = pack_bf16x2(v[2], v[3]);
            lo.z = pack_bf16x2(v[4], v[5]);   lo.w = pack_bf16x2(v[6], v[7]);
            hi.x = pack_bf16x2(v[8], v[9]);   hi.y = pack_bf16x2(v[10], v[11]);
            hi.z = pack_bf16x2(v[12], v[13]); hi.w = pack_bf16x2(v[14], v[15]);
            *reinterpret_cast<uint4 *>(my_row + c0 * 2) = lo;
            if (c0 + 8 < a.n_store) *reinterpret_cast<uint4 *>(my_row + c0 * 2 + 16) = hi;
        }
        fence_before_sync();
        __syncthreads();

        // ---- epilogue 2: coalesced store + column statistics of the stored values ----
        const int cpr = a.n_store >> 3;
        for (int q = tid; q < kTcBM * cpr; q += kTcThreads) {
            const int r = q / cpr, c = q - r * cpr;
            const int64_t m = m0 + r;
            if (m < a.M)
                *reinterpret_cast<uint4 *>(a.Z + m * a.ldz + c * 8) =
                    *reinterpret_cast<const uint4 *>(staging + (size_t)r * st_stride + c * 16);
        }
        if (STATS) {
            const int rows = (int)min((int64_t)32, a.M - (m0 + warp * 32));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int p = lane + 32 * j;
                if (2 * p < a.N) {
                    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
                    for (int r = 0; r < rows; ++r) {
                        const float2 f = unpack_bf16x2(
                            *reinterpret_cast<const uint32_t *>(staging + (size_t)(warp * 32 + r) * st_stride + p * 4));
                        s1a += f.x; s1b += f.y;
                        s2a = fmaf(f.x, f.x, s2a); s2b = fmaf(f.y, f.y, s2b);
                    }
                    sum1[2 * j] += s1a; sum1[2 * j + 1] += s1b;
                    sum2[2 * j] += s2a; sum2[2 * j + 1] += s2b;
                }
            }
        }
        fence_proxy_async();   // generic-proxy accesses of the staging tile before the next TMA write into it
        __syncthreads();       // staging consumed before the next tile's operands overwrite it
    }

    if (STATS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = lane + 32 * j;
            if (2 * p < 256) {
                s_part[warp][0][2 * p] = sum1[2 * j];     s_part[warp][0][2 * p + 1] = sum1[2 * j + 1];
                s_part[warp][1][2 * p] = sum2[2 * j];     s_part[warp][1][2 * p + 1] = sum2[2 * j + 1];
            }
        }
        __syncthreads();
        for (int c = tid; c < a.N; c += kTcThreads) {
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) { t1 += s_part[w][0][c]; t2 += s_part[w][1][c]; }
            double *acc = a.stat_accum + (size_t)(blockIdx.x % kStatReplicas) * 2 * a.stat_ld + a.stat_off;
            atomicAdd(acc + c, (double)t1);
            atomicAdd(acc + a.stat_ld + c, (double)t2);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

int linear_num_partials(int64_t M);

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

size_t tc_wpack_bytes(int K, int N) {
    // images for every 256-column block of N, each [KC][n_pad][128 B]
    size_t KC = (size_t)(K + kTcBK - 1) / kTcBK, total = 0;
    for (int n0 = 0; n0 < N; n0 += 256) total += KC * (size_t)round_up(N - n0 < 256 ? N - n0 : 256, 16) * 128;
    return total;
}

// Z[M,N] = act(X) . W^T with W element (n,k) at W[n*w_sn + k*w_sk]
int tc_linear_nt(const void *X, int ldx, const float *in_scale, const float *in_shift, const float *W, int64_t w_sn,
                 int64_t w_sk, const float *bias, int64_t M, int K, int N, void *Z, int ldz, double *stat_accum,
                 void *wpack, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, linear_tc_kernel<true>);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - (int)fa.sharedSizeBytes);
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, linear_tc_kernel<false>);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(linear_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - (int)fa.sharedSizeBytes);
        if (e != cudaSuccess) {
            set_error("linear_tc: shared-memory opt-in failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return PN2_ERR_CUDA;
        }
        attr_done = true;
    }
    const int KC = (K + kTcBK - 1) / kTcBK;
    uint8_t *img = (uint8_t *)wpack;
    for (int n0 = 0; n0 < N; n0 += 256) {
        const int nb = N - n0 < 256 ? N - n0 : 256;
        const int n_pad = round_up(nb, 16);
        int n_store = round_up(nb, 8);
        if (n0 + n_store > ldz) n_store = ldz - n0;      // ldz is a multiple of 8 on this path
        const size_t img_bytes = (size_t)KC * n_pad * 128;
        const int total = KC * n_pad * 8;
        pack_weight_kernel<<<(total + 255) / 256, 256, 0, st>>>(W + (int64_t)n0 * w_sn, w_sn, w_sk, nb, K, n_pad, KC, img);
        count_launch();
        TcLinearArgs a;
        a.X = (const __nv_bfloat16 *)X;
        a.ldx = ldx;
        a.in_scale = in_scale;
        a.in_shift = in_shift;
        a.Wimg = img;
        a.bias = bias ? bias + n0 : nullptr;
        a.M = M; a.K = K; a.N = nb; a.n_pad = n_pad; a.n_store = n_store; a.KC = KC;
        a.Z = (__nv_bfloat16 *)Z + n0;
        a.ldz = ldz;
        a.stat_accum = stat_accum;
        a.stat_ld = N;
        a.stat_off = n0;
        const size_t staging = (size_t)kTcBM * (n_store * 2 + 16);
        a.w_resident = (KC == 1 && staging <= (size_t)2 * kTcAStage + (size_t)n_pad * 128) ? 1 : 0;
        const size_t dyn = 1024 + 2 * kTcAStage + 2 * (size_t)n_pad * 128;
        const int grid = linear_num_partials(M);
        if (stat_accum)
            linear_tc_kernel<true><<<grid, kTcThreads, dyn, st>>>(a);
        else
            linear_tc_kernel<false><<<grid, kTcThreads, dyn, st>>>(a);
        count_launch();
        int rc = check_launch("linear_tc");
        if (rc != PN2_OK) return rc;
        img += img_bytes;
    }
    return PN2_OK;
}

}  // namespace pn2

// =================================================================================================
// Weight gradient  dW[N,K] = sum_m dZ[m,n] * act(X)[m,k]   on the tensor cores.
//
// The reduction runs over ROWS, so both operands are "MN-major" in UMMA terms: a stage holds R
// consecutive rows of dZ and of act(X) copied straight from their row-major storage (16-byte
// chunks, 128-byte swizzle); a 64-column slab of a row is one 128-byte swizzle row, slabs are
// LBO = R*128 bytes apart, groups of 8 rows SBO = 1024 bytes apart, and each tcgen05.mma consumes
// 16 rows (advance the descriptor start by 16*128 bytes).  A CTA owns one (128 x <=256) block of
// dW and a contiguous range of rows, accumulates it in tensor memory over all its stages, and
// writes one fp32 partial block; wgrad_reduce_kernel sums the partials in a fixed order.
// =================================================================================================
namespace pn2 {

struct TcWgradArgs {
    const __nv_bfloat16 *dZ;
    int lddz;
    const __nv_bfloat16 *X;
    int ldx;
    const float *in_scale, *in_shift;
    int64_t M, rows_per_split;
    int K, N, K_ld;    // K_ld: row stride of the fp32 partials (K rounded up to 4 -> 16-byte stores)
    int nblk_k;        // blockIdx.y = (dW row block of 128) * nblk_k + (dW column block of 256)
    int a_slabs, b_slabs, R;   // ring-slot geometry of the largest block
    float *scratch;    // [splits][N][K_ld]
};

constexpr int kWgStages = 4;           // ring slots
constexpr int kWgAhead = kWgStages - 2;   // copies run this many iterations ahead; one MMA may still be draining

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid) {
    // 16-byte asynchronous global->shared copy (LDGSTS); src-size 0 writes zeros without touching memory
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}

__global__ void __launch_bounds__(kTcThreads) wgrad_tc_kernel(const TcWgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_mma[kWgStages];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_scale[256], s_shift[256];

    const int tid = threadIdx.x, warp = tid >> 5;
    const int n0 = ((int)blockIdx.y / a.nblk_k) * 128, k0 = ((int)blockIdx.y % a.nblk_k) * 256;
    const int nb = min(128, a.N - n0), kb = min(256, a.K - k0);
    const int kb_pad = (kb + 15) & ~15;
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const uint32_t slab_bytes = (uint32_t)a.R * 128u;
    const uint32_t stage_bytes = slab_bytes * (uint32_t)(a.a_slabs + a.b_slabs);
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < kb_pad) tmem_cols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < kWgStages; ++i) mbar_init(&bar_mma[i], 1);
        mbar_init_fence();
    }
    if (a.in_scale)
        for (int i = tid; i < kb; i += kTcThreads) {
            s_scale[i] = a.in_scale[k0 + i];
            s_shift[i] = a.in_shift[k0 + i];
        }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_bf16(128, kb_pad, 1, 1);

    const int64_t r_begin = (int64_t)blockIdx.x * a.rows_per_split;
    const int64_t r_end = min(a.M, r_begin + a.rows_per_split);
    const int n_it = r_end > r_begin ? (int)((r_end - r_begin + a.R - 1) / a.R) : 0;
    // valid 16-byte chunks per row of each operand (the rest of a 64-column slab is never read back
    // into a stored output element, so it is left untouched)
    const int a_cpr = (min(nb, a.lddz - n0) + 7) >> 3, b_cpr = (min(kb, a.ldx - k0) + 7) >> 3;
    uint32_t par[kWgStages];
    for (int i = 0; i < kWgStages; ++i) par[i] = 0;

    // issue the asynchronous copies of iteration `it` into its ring slot
    auto issue = [&](int it) {
        const int s = it % kWgStages;
        const int64_t r0 = r_begin + (int64_t)it * a.R;
        const int rows = (int)min((int64_t)a.R, r_end - r0);
        const int rows16 = (rows + 15) & ~15;
        const uint32_t stA = smem_addr(smem + (size_t)s * stage_bytes);
        const uint32_t stB = stA + (uint32_t)a.a_slabs * slab_bytes;
        for (int q = tid; q < rows16 * a_cpr; q += kTcThreads) {
            const int r = q / a_cpr, cc = q - r * a_cpr;
            const bool ok = r < rows;
            cp_async16(stA + (uint32_t)(cc >> 3) * slab_bytes + sw128_offset(r, cc & 7),
                       a.dZ + (ok ? (r0 + r) * a.lddz + n0 + cc * 8 : 0), ok);
        }
        for (int q = tid; q < rows16 * b_cpr; q += kTcThreads) {
            const int r = q / b_cpr, cc = q - r * b_cpr;
            const bool ok = r < rows;
            cp_async16(stB + (uint32_t)(cc >> 3) * slab_bytes + sw128_offset(r, cc & 7),
                       a.X + (ok ? (r0 + r) * a.ldx + k0 + cc * 8 : 0), ok);
        }
    };

    for (int p = 0; p < kWgAhead; ++p) {
        if (p < n_it) issue(p);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int it = 0; it < n_it; ++it) {
        const int s = it % kWgStages;
        // refill the slot that iteration it-2 used (its MMA was issued a whole iteration ago) with it+kWgAhead
        const int nxt = it + kWgAhead;
        if (nxt < n_it) {
            if (nxt >= kWgStages) {
                const int sp = nxt % kWgStages;
                mbar_wait(&bar_mma[sp], par[sp]);
                par[sp] ^= 1;
            }
            issue(nxt);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(kWgAhead) : "memory");   // iteration `it` has landed

        const int64_t r0 = r_begin + (int64_t)it * a.R;
        const int rows = (int)min((int64_t)a.R, r_end - r0);
        const int rows16 = (rows + 15) & ~15;
        uint8_t *stA = smem + (size_t)s * stage_bytes;
        uint8_t *stB = stA + (size_t)a.a_slabs * slab_bytes;
        if (a.in_scale) {   // act(X) = relu(bn(.)) of the previous layer, in place on the chunks this thread copied
            for (int q = tid; q < rows * b_cpr; q += kTcThreads) {
                const int r = q / b_cpr, cc = q - r * b_cpr;
                uint4 *ptr = reinterpret_cast<uint4 *>(stB + (size_t)(cc >> 3) * slab_bytes + sw128_offset(r, cc & 7));
                uint4 v = *ptr;
                uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                    float2 f = unpack_bf16x2(w[e2]);
                    const int i0 = cc * 8 + 2 * e2;
                    f.x = i0 < kb ? fmaxf(fmaf(f.x, s_scale[i0], s_shift[i0]), 0.0f) : 0.0f;
                    f.y = i0 + 1 < kb ? fmaxf(fmaf(f.y, s_scale[i0 + 1], s_shift[i0 + 1]), 0.0f) : 0.0f;
                    w[e2] = pack_bf16x2(f.x, f.y);
                }
                *ptr = v;
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            const uint32_t a_base = smem_addr(stA), b_base = smem_addr(stB);
            for (int j = 0; j < rows16 / 16; ++j)
                umma_bf16(tmem, make_desc(a_base + 2048 * j, slab_bytes, 1024), make_desc(b_base + 2048 * j, slab_bytes, 1024),
                          idesc, (uint32_t)((it | j) != 0));
            umma_commit(&bar_mma[s]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // drain the commits nobody waited for yet, oldest first: iteration j was consumed in the loop
    // iff a refill of its slot followed, i.e. iff j + kWgStages < n_it
    for (int j = max(0, n_it - kWgStages); j < n_it; ++j) {
        const int sj = j % kWgStages;
        mbar_wait(&bar_mma[sj], par[sj]);
        par[sj] ^= 1;
    }
    fence_after_sync();
    // ---- epilogue: this thread's dW row n0 + tid, fp32 partial, 16-byte stores ----
    const int n = n0 + tid;
    float *out = a.scratch + ((size_t)blockIdx.x * a.N + n) * a.K_ld + k0;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < kb_pad; c0 += 16) {
        float v[16];
        if (n_it > 0) tmem_ld16(taddr + c0, v);
        else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.0f;
        }
        if (tid < nb) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
                if (k0 + c0 + i < a.K_ld)
                    *reinterpret_cast<float4 *>(out + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// blockDim (32, 8): 32 consecutive elements per CTA, the 8 warps stride over the splits, then a
// fixed-order combine through shared memory (deterministic)
__global__ void wgrad_reduce_kernel2(const float *__restrict__ scratch, int splits, int N, int K, int K_ld,
                                     float *__restrict__ dW) {
    __shared__ float red[8][32];
    const int64_t e = (int64_t)blockIdx.x * 32 + threadIdx.x, NK = (int64_t)N * K;
    float s = 0.0f;
    if (e < NK) {
        const int64_t src = (e / K) * K_ld + (e % K), plane = (int64_t)N * K_ld;
        for (int z = threadIdx.y; z < splits; z += 8) s += scratch[(int64_t)z * plane + src];
    }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && e < NK) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        dW[e] = t;
    }
}

struct WgradPlan {
    int splits, nblk_n, nblk_k, a_slabs, b_slabs, R, K_ld;
    int64_t rows_per_split;
    size_t dyn_smem;
};

static WgradPlan tc_wgrad_plan(int64_t M, int K, int N) {
    WgradPlan p;
    p.nblk_n = (N + 127) / 128;
    p.nblk_k = (K + 255) / 256;
    p.K_ld = round_up(K, 4);
    p.a_slabs = ((N < 128 ? N : 128) + 63) / 64;
    p.b_slabs = ((K < 256 ? K : 256) + 63) / 64;
    const int per_row = 128 * (p.a_slabs + p.b_slabs);
    int R = (26 * 1024 / per_row) / 16 * 16;          // <= 26 KB per ring slot: four slots, two CTAs per SM
    if (R > 256) R = 256;
    if (R < 16) R = 16;
    p.R = R;
    p.dyn_smem = 1024 + kWgStages * (size_t)R * per_row;
    // splits: fill the machine (two CTAs per SM), but keep >= 2 ring iterations per CTA, and balance the
    // per-CTA iteration chain (~1 us each) against writing + re-reading the fp32 partials
    const int blocks = p.nblk_n * p.nblk_k;
    int64_t s = (2 * kNumSMs + blocks - 1) / blocks;
    const int64_t max_rows = (M + 2 * R - 1) / (2 * R);
    const double chain = (double)M / R * 1e-6, part = 2.0 * N * p.K_ld * 4.0 / 3.0e12;
    int64_t s_bal = (int64_t)(sqrt(chain / part) + 0.5);
    if (s > max_rows) s = max_rows;
    if (s > s_bal) s = s_bal;
    if (s < 1) s = 1;
    p.rows_per_split = ((M + s - 1) / s + 15) / 16 * 16;
    p.splits = (int)((M + p.rows_per_split - 1) / p.rows_per_split);
    return p;
}

size_t tc_wgrad_scratch_bytes(int64_t M, int K, int N) {
    const WgradPlan p = tc_wgrad_plan(M, K, N);
    return sizeof(float) * (size_t)p.splits * (size_t)N * (size_t)p.K_ld + 16;
}

int tc_linear_wgrad(const void *dZ, int lddz, const void *X, int ldx, const float *in_scale, const float *in_shift,
                    int64_t M, int K, int N, float *dW, void *scratch, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, wgrad_tc_kernel);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - (int)fa.sharedSizeBytes);
        if (e != cudaSuccess) {
            set_error("wgrad_tc: shared-memory opt-in failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return PN2_ERR_CUDA;
        }
        attr_done = true;
    }
    const WgradPlan p = tc_wgrad_plan(M, K, N);
    TcWgradArgs a;
    a.dZ = (const __nv_bfloat16 *)dZ; a.lddz = lddz;
    a.X = (const __nv_bfloat16 *)X;   a.ldx = ldx;
    a.in_scale = in_scale; a.in_shift = in_shift;
    a.M = M; a.rows_per_split = p.rows_per_split; a.K = K; a.N = N; a.K_ld = p.K_ld;
    a.nblk_k = p.nblk_k; a.a_slabs = p.a_slabs; a.b_slabs = p.b_slabs; a.R = p.R;
    a.scratch = (float *)(((uintptr_t)scratch + 15) & ~(uintptr_t)15);
    wgrad_tc_kernel<<<dim3((unsigned)p.splits, (unsigned)(p.nblk_n * p.nblk_k)), kTcThreads, p.dyn_smem, st>>>(a);
    count_launch();
    int rc = check_launch("wgrad_tc");
    if (rc != PN2_OK) return rc;
    const int64_t NK = (int64_t)N * K;
    wgrad_reduce_kernel2<<<(unsigned)((NK + 31) / 32), dim3(32, 8), 0, st>>>(a.scratch, p.splits, N, K, p.K_ld, dW);
    count_launch();
    return check_launch("wgrad_reduce");
}

}  // namespace pn2
