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
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pn2 {

using namespace tc;

constexpr int kTcBM = 128;
constexpr int kTcBK = 64;                       // bf16 elements per 128-byte swizzle row
constexpr int kTcAStage = kTcBM * kTcBK * 2;    // 16 KB

// ---- weight packing: fp32 W (any strides) -> bf16 chunk images in UMMA K-major SW128 layout ----
// image kc: n_pad rows x 128 bytes; element (n, kc*64 + c*8 + e) at sw128_offset(n, c) + 2e.
// k_rot: image column k holds W column (k + k_rot) % K (sa_fused.cu gathers [feats | xyz] while the conv sees [xyz | feats]).
__global__ void pack_weight_kernel(const float *__restrict__ W, int64_t w_sn, int64_t w_sk, int N, int K, int n_pad,
                                   int KC, int k_rot, uint8_t *__restrict__ img) {
    const int total = KC * n_pad * 8;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const int c = q & 7, n = (q >> 3) % n_pad, kc = (q >> 3) / n_pad;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = kc * kTcBK + c * 8 + e;
            v[e] = (n < N && k < K) ? W[(int64_t)n * w_sn + (int64_t)((k + k_rot) % K) * w_sk] : 0.0f;
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]);
        o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]);
        o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4 *>(img + (size_t)kc * n_pad * 128 + sw128_offset(n, c)) = o;
    }
}

// One launch packs up to kPackJobs weights (every layer of an MLP, both orientations): job j owns blocks
// [first_block[j], first_block[j+1]); its image is the concatenation over 256-row blocks of N of [KC][n_pad][128 B].
constexpr int kPackJobs = 48;      // the whole network (23 layers x 2 orientations) in one launch
struct PackJobs {
    int n;
    const float *W[kPackJobs];
    int64_t w_sn[kPackJobs], w_sk[kPackJobs];
    int N[kPackJobs], K[kPackJobs];
    uint8_t *img[kPackJobs];
    int first_block[kPackJobs + 1];
};
__global__ void pack_weights_multi_kernel(const __grid_constant__ PackJobs a) {
    int j = 0;
    while (j + 1 < a.n && (int)blockIdx.x >= a.first_block[j + 1]) ++j;
    const int N = a.N[j], K = a.K[j], KC = (K + kTcBK - 1) / kTcBK;
    const int q = ((int)blockIdx.x - a.first_block[j]) * blockDim.x + threadIdx.x;     // 16-byte unit of the job's image
    const int full = KC * 256 * 8;                                                  // units of one full 256-row block
    const int blk = q / full, rem = q - blk * full;
    const int n0 = blk * 256;
    if (n0 >= N) return;
    const int nb = N - n0 < 256 ? N - n0 : 256, n_pad = (nb + 15) & ~15;
    if (rem >= KC * n_pad * 8) return;
    const int c = rem & 7, n = (rem >> 3) % n_pad, kc = (rem >> 3) / n_pad;
    const float *W = a.W[j] + (int64_t)n0 * a.w_sn[j];
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = kc * kTcBK + c * 8 + e;
        v[e] = (n < nb && k < K) ? W[(int64_t)n * a.w_sn[j] + (int64_t)k * a.w_sk[j]] : 0.0f;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4 *>(a.img[j] + (size_t)blk * KC * 256 * 128 + (size_t)kc * n_pad * 128 + sw128_offset(n, c)) = o;
}

constexpr int kTcMaxNBlocks = 4;
struct TcLinearArgs {
    CUtensorMap tm_x, tm_z[kTcMaxNBlocks];   // X [M, ldx] (boxes 64 x 128 rows); Z column block y: [M, n_store(y)] (row stride ldz)
    const float *in_scale, *in_shift;
    const uint8_t *Wimg;
    const float *bias;
    int64_t M;
    int K, N, ldz, KC;       // N: all output columns of this launch (blocks of 256 along blockIdx.y)
    double *stat_accum;     // [replicas][2][stat_ld] fp64 accumulators, this call adds into columns stat_off .. stat_off+N
    int stat_ld, stat_off;
    int w_resident, stages;
    BnFinalize fin;         // fin.ticket != null: the CTA that draws the last ticket turns the sums into scale/shift
};

constexpr int kLinTcThreads = 160;   // warps 0-3: A transform, MMA issue (thread 0), epilogue; warp 4: TMA producer

// Column statistics (sum z, sum z^2 over this warp's 32 rows) of the bf16 tile staged in SW128 box layout, the values exactly
// as they are stored.  Lane <-> column pair p (slab p >> 5, 16-byte chunk (p & 31) >> 2, byte (p & 3) * 4 inside it); row
// 8g + ri of the warp's 4 KB sits at g * 1024 + ri * 128 + ((chunk ^ ri) << 4): with ri unrolled the eight offsets are
// registers and g an immediate, so an element pair costs LDS + 2 unpack + 4 FP instructions (the rolled loop with the offset
// recomputed per row was 35 % of the instructions this kernel issued on the sa1 layers).  N <= 32: only 16 lanes own a
// column pair, so the two half-warps split the rows (groups 0-1 / 2-3); their sums are added once at the end of the CTA.
template <int GROUPS>
__device__ __forceinline__ void tile_pair_stats(const uint8_t *base, int ch, int inb, float &s1a, float &s1b, float &s2a, float &s2b) {
    uint32_t off[8];
#pragma unroll
    for (int ri = 0; ri < 8; ++ri) off[ri] = (uint32_t)(ri * 128 + ((ch ^ ri) << 4) + inb);
#pragma unroll
    for (int g = 0; g < GROUPS; ++g)
#pragma unroll
        for (int ri = 0; ri < 8; ++ri) {
            const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(base + g * 1024 + off[ri]));
            s1a += f.x; s1b += f.y;
            s2a = fmaf(f.x, f.x, s2a); s2b = fmaf(f.y, f.y, s2b);
        }
}

__device__ __forceinline__ void tile_col_stats(const uint8_t *staging, int warp, int lane, int N, int rows, float *sum1, float *sum2) {
    if (rows == 32 && N <= 32) {
        const int p = lane & 15;
        if (2 * p < N)
            tile_pair_stats<2>(staging + (size_t)(warp * 4 + 2 * (lane >> 4)) * 1024, p >> 2, (p & 3) * 4, sum1[0], sum1[1], sum2[0], sum2[1]);
        return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = lane + 32 * j;
        if (2 * p < N) {
            const uint8_t *slab = staging + (size_t)(p >> 5) * kTcAStage;
            const int ch = (p & 31) >> 2, inb = (p & 3) * 4;
            if (rows == 32) {
                tile_pair_stats<4>(slab + (size_t)warp * 4096, ch, inb, sum1[2 * j], sum1[2 * j + 1], sum2[2 * j], sum2[2 * j + 1]);
            } else {                              // the ragged last tile
                float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
                for (int r = 0; r < rows; ++r) {
                    const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(slab + sw128_offset(warp * 32 + r, ch) + inb));
                    s1a += f.x; s1b += f.y;
                    s2a = fmaf(f.x, f.x, s2a); s2b = fmaf(f.y, f.y, s2b);
                }
                sum1[2 * j] += s1a; sum1[2 * j + 1] += s1b;
                sum2[2 * j] += s2a; sum2[2 * j + 1] += s2b;
            }
        }
    }
}

// N <= 32: fold the upper half-warp's partial sums (same column pairs, other rows) into the lower one before they are published
__device__ __forceinline__ void fold_split_stats(int lane, int N, float *sum1, float *sum2) {
    if (N <= 32) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float o1 = __shfl_down_sync(0xffffffffu, sum1[i], 16), o2 = __shfl_down_sync(0xffffffffu, sum2[i], 16);
            sum1[i] = lane < 16 ? sum1[i] + o1 : 0.0f;
            sum2[i] = lane < 16 ? sum2[i] + o2 : 0.0f;
        }
    }
}

// Shared-memory map (1024-byte aligned): [W image: KC chunks if resident] [staging: ceil(n_store/64) slabs of
// 128 rows x 128 B] [ring: `stages` slots of {A chunk 16 KB, + W chunk n_pad*128 B when W is streamed}]
template <bool STATS>
__global__ void __launch_bounds__(kLinTcThreads) linear_tc_kernel(const __grid_constant__ TcLinearArgs a) {
    extern __shared__ uint8_t smem_raw[];
    // this CTA's column block: a call with N > 256 output columns runs its <= kTcMaxNBlocks blocks of 256 as blockIdx.y of ONE
    // launch (they used to be serial launches: fp4.1's data gradient, 3 blocks of 16 tiles, took 46 us)
    const int yb = blockIdx.y;
    const int bN = min(256, a.N - 256 * yb);
    const int b_n_pad = (bN + 15) & ~15;
    const int b_n_store = min((bN + 7) & ~7, a.ldz - 256 * yb);
    const uint8_t *const bWimg = a.Wimg + (size_t)yb * a.KC * 256 * 128;
    const float *const b_bias = a.bias ? a.bias + 256 * yb : nullptr;
    const int b_stat_off = a.stat_off + 256 * yb;
    unsigned *const b_ticket = a.fin.ticket ? a.fin.ticket + yb : nullptr;
    const CUtensorMap *const b_tm_z = &a.tm_z[yb];
    __shared__ __align__(8) uint64_t bar_full[8], bar_empty[8], bar_w, bar_acc;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_last;

    const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned (SW128 atoms)
    const uint32_t b_bytes = (uint32_t)b_n_pad * 128u;
    const int z_slabs = (b_n_store + 63) >> 6;
    uint8_t *const w_res = smem;
    uint8_t *const staging = smem + (a.w_resident ? (size_t)a.KC * b_bytes : 0);
    uint8_t *const ring = staging + (size_t)z_slabs * kTcAStage;
    const uint32_t slot_bytes = kTcAStage + (a.w_resident ? 0u : b_bytes);

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < b_n_pad) tmem_cols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
        }
        mbar_init(&bar_w, 1);
        mbar_init(&bar_acc, 1);
        mbar_init_fence();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const int64_t m_tiles = (a.M + kTcBM - 1) / kTcBM;

    if (warp == 4) {
        // ---- producer: one thread streams the A chunks (and W chunks unless resident) with TMA ----
        if (lane == 0) {
            tma_prefetch_desc(&a.tm_x);
            if (a.w_resident) {
                mbar_expect_tx(&bar_w, (uint32_t)a.KC * b_bytes);
                for (int kc = 0; kc < a.KC; ++kc) bulk_g2s(w_res + (size_t)kc * b_bytes, bWimg + (size_t)kc * b_bytes, b_bytes, &bar_w);
            }
            uint32_t n = 0;
            for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x)
                for (int kc = 0; kc < a.KC; ++kc, ++n) {
                    const uint32_t s = n % (uint32_t)a.stages, use = n / (uint32_t)a.stages;
                    if (use > 0) mbar_wait(&bar_empty[s], (use - 1) & 1u);
                    uint8_t *slot = ring + (size_t)s * slot_bytes;
                    mbar_expect_tx(&bar_full[s], slot_bytes);
                    tma_load_2d(slot, &a.tm_x, kc * kTcBK, (int)(tile * kTcBM), &bar_full[s]);
                    if (!a.w_resident) bulk_g2s(slot + kTcAStage, bWimg + (size_t)kc * b_bytes, b_bytes, &bar_full[s]);
                }
        }
        __syncwarp();
    } else {
        const uint32_t idesc = make_idesc_bf16(kTcBM, b_n_pad, 0, 0);
        float sum1[8], sum2[8];   // STATS: this lane's column pairs p = lane + 32 j  (columns 2p, 2p+1)
#pragma unroll
        for (int i = 0; i < 8; ++i) sum1[i] = sum2[i] = 0.0f;
        // transform ownership: physical 16-byte unit q = tid + 128 i of a chunk: row r = q >> 3 (r & 7 is the same
        // for every i), logical column chunk cc = (q & 7) ^ (r & 7)
        const int t_cc = (tid & 7) ^ ((tid >> 3) & 7);
        if (a.w_resident && warp == 0) mbar_wait(&bar_w, 0);
        uint32_t n = 0, acc_par = 0;
        for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
            const int64_t m0 = tile * kTcBM;
            for (int kc = 0; kc < a.KC; ++kc, ++n) {
                const uint32_t s = n % (uint32_t)a.stages, par = (n / (uint32_t)a.stages) & 1u;
                uint8_t *slot = ring + (size_t)s * slot_bytes;
                if (a.in_scale) {
                    // previous layer's BatchNorm + ReLU, in place on the landed chunk
                    const int k = kc * kTcBK + t_cc * 8;
                    float sc[8], sh[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const bool ok = k + e < a.K;
                        sc[e] = ok ? a.in_scale[k + e] : 0.0f;
                        sh[e] = ok ? a.in_shift[k + e] : 0.0f;
                    }
                    mbar_wait(&bar_full[s], par);
                    if (k < a.K) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            uint4 *ptr = reinterpret_cast<uint4 *>(slot + (size_t)(tid + 128 * i) * 16);
                            uint4 v = *ptr;
                            uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2) {
                                float2 f = unpack_bf16x2(w[e2]);
                                f.x = fmaxf(fmaf(f.x, sc[2 * e2], sh[2 * e2]), 0.0f);
                                f.y = fmaxf(fmaf(f.y, sc[2 * e2 + 1], sh[2 * e2 + 1]), 0.0f);
                                w[e2] = pack_bf16x2(f.x, f.y);
                            }
                            *ptr = v;
                        }
                    }
                    fence_proxy_async();
                    named_bar_sync(1, 128);
                } else if (warp == 0) {
                    mbar_wait(&bar_full[s], par);
                }
                // warp 0, converged, one elected lane issues: descriptors stay on the uniform datapath (tc_common.cuh: elect_one_sync)
                if (warp == 0) {
                    fence_after_sync();
                    const int k_left = a.K - kc * kTcBK;
                    const int nk = k_left >= kTcBK ? 4 : (k_left + 15) / 16;
                    const uint64_t ad = make_desc(smem_addr(slot), 0, 1024);
                    const uint64_t bd = make_desc(smem_addr(a.w_resident ? w_res + (size_t)kc * b_bytes : slot + kTcAStage), 0, 1024);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < nk) umma_bf16(tmem, ad + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc, (uint32_t)((kc | j) != 0));
                        umma_commit(&bar_empty[s]);
                        if (kc == a.KC - 1) umma_commit(&bar_acc);
                    }
                    __syncwarp();
                }
            }
            mbar_wait(&bar_acc, acc_par);
            acc_par ^= 1u;
            fence_after_sync();

            // ---- epilogue 1: TMEM -> registers -> (+bias) -> bf16 -> staging row `tid` (SW128 box layout) ----
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < b_n_store; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
                if (b_bias) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < bN) v[i] += b_bias[c0 + i];
                }
                uint4 lo, hi;
                lo.x = pack_bf16x2(v[0], v[1]);   lo.y = pack_bf16x2(v[2], v[3]);
                lo.z = pack_bf16x2(v[4], v[5]);   lo.w = pack_bf16x2(v[6], v[7]);
                hi.x = pack_bf16x2(v[8], v[9]);   hi.y = pack_bf16x2(v[10], v[11]);
                hi.z = pack_bf16x2(v[12], v[13]); hi.w = pack_bf16x2(v[14], v[15]);
                uint8_t *slab = staging + (size_t)(c0 >> 6) * kTcAStage;
                const int ch = (c0 & 63) >> 3;
                *reinterpret_cast<uint4 *>(slab + sw128_offset(tid, ch)) = lo;
                if (c0 + 8 < b_n_store) *reinterpret_cast<uint4 *>(slab + sw128_offset(tid, ch + 1)) = hi;
            }
            fence_before_sync();     // TMEM reads done before the next tile's MMA (issued after the barrier below)
            fence_proxy_async();     // staging writes -> visible to the TMA store
            named_bar_sync(1, 128);

            // ---- epilogue 2: TMA store of the tile (rows >= M and columns >= n_store are clipped) + column statistics ----
            if (tid == 0)
                for (int j = 0; j < z_slabs; ++j) tma_store_2d(b_tm_z, 64 * j, (int)m0, staging + (size_t)j * kTcAStage);
            if (STATS) {
                const int rows = (int)max((int64_t)0, min((int64_t)32, a.M - (m0 + warp * 32)));
                tile_col_stats(staging, warp, lane, bN, rows, sum1, sum2);
            }
            if (tid == 0) tma_store_wait_read();
            named_bar_sync(1, 128);   // staging free again (store has read it, statistics have read it)
        }
        if (tid == 0) tma_store_wait_all();

        if (STATS) {
            fold_split_stats(lane, bN, sum1, sum2);
            float(*s_part)[2][256] = reinterpret_cast<float(*)[2][256]>(staging);   // 8 KB of the (now idle) staging tile
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int p = lane + 32 * j;
                s_part[warp][0][2 * p] = sum1[2 * j];     s_part[warp][0][2 * p + 1] = sum1[2 * j + 1];
                s_part[warp][1][2 * p] = sum2[2 * j];     s_part[warp][1][2 * p + 1] = sum2[2 * j + 1];
            }
            named_bar_sync(1, 128);
            for (int c = tid; c < bN; c += 128) {
                float t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) { t1 += s_part[w][0][c]; t2 += s_part[w][1][c]; }
                double *acc = a.stat_accum + (size_t)(blockIdx.x % kStatReplicas) * 2 * a.stat_ld + b_stat_off;
                atomicAdd(acc + c, (double)t1);
                atomicAdd(acc + a.stat_ld + c, (double)t2);
            }
            if (b_ticket) {
                // ---- "last CTA finalizes": mean / variance -> scale, shift, running statistics; accumulator and ticket
                //      are left zeroed for the next layer (saves the finalize launch between two layers) ----
                __threadfence();
                named_bar_sync(1, 128);
                if (tid == 0) s_last = atomicAdd(b_ticket, 1u) == gridDim.x - 1 ? 1 : 0;
                named_bar_sync(1, 128);
                if (s_last) {
                    __threadfence();
                    for (int c = tid; c < bN; c += 128) {
                        double s1 = 0.0, s2 = 0.0;
#pragma unroll
                        for (int r = 0; r < kStatReplicas; ++r) {   // fixed order
                            double *acc = a.stat_accum + (size_t)r * 2 * a.stat_ld + b_stat_off;
                            s1 += __ldcg(acc + c);
                            s2 += __ldcg(acc + a.stat_ld + c);
                            acc[c] = 0.0;
                            acc[a.stat_ld + c] = 0.0;
                        }
                        bn_finalize_channel(s1, s2, a.M, b_stat_off + c, a.fin.gamma, a.fin.beta, a.fin.conv_bias, a.fin.eps,
                                            a.fin.momentum_dev ? __ldg(a.fin.momentum_dev) : a.fin.momentum, a.fin.running_mean, a.fin.running_var, a.fin.scale, a.fin.shift,
                                            a.fin.save_mean, a.fin.save_invstd);
                    }
                    if (tid == 0) {
                        *b_ticket = 0u;
                        if (a.fin.num_batches_tracked && b_stat_off == 0) *a.fin.num_batches_tracked += 1;
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// -------------------------------------------------------------------------------------------------
// Version 2 of the layer kernel: the same tile arithmetic, software-pipelined across tiles inside ONE persistent CTA.
//   warps 0-3  epilogue group : TMEM -> (+bias) -> bf16 -> staging -> TMA store (+ column statistics)
//   warps 4-7  transform group: previous layer's BatchNorm+ReLU in place on the landed A chunk; thread 128 issues the MMAs
//   warp  8    producer       : TMA loads of the A chunks (and W chunks when streamed) into a deep ring
// The accumulator is double buffered in tensor memory (2 x n_pad columns), so the MMAs of tile t+1 (and the loads of
// tiles t+2...) run while the epilogue group drains tile t: the ring, not the number of co-resident CTAs, hides the HBM
// latency (v1 measured 0.8-2.1 TB/s on the 128 x 128 layers with the tensor pipe 5 % busy).
// -------------------------------------------------------------------------------------------------
constexpr int kLinTc2Threads = 288;
constexpr int kTc2MaxStages = 12;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

template <bool STATS>
__global__ void __launch_bounds__(kLinTc2Threads) linear_tc2_kernel(const __grid_constant__ TcLinearArgs a) {
    extern __shared__ uint8_t smem_raw[];
    // this CTA's column block: a call with N > 256 output columns runs its <= kTcMaxNBlocks blocks of 256 as blockIdx.y of ONE
    // launch (they used to be serial launches: fp4.1's data gradient, 3 blocks of 16 tiles, took 46 us)
    const int yb = blockIdx.y;
    const int bN = min(256, a.N - 256 * yb);
    const int b_n_pad = (bN + 15) & ~15;
    const int b_n_store = min((bN + 7) & ~7, a.ldz - 256 * yb);
    const uint8_t *const bWimg = a.Wimg + (size_t)yb * a.KC * 256 * 128;
    const float *const b_bias = a.bias ? a.bias + 256 * yb : nullptr;
    const int b_stat_off = a.stat_off + 256 * yb;
    unsigned *const b_ticket = a.fin.ticket ? a.fin.ticket + yb : nullptr;
    const CUtensorMap *const b_tm_z = &a.tm_z[yb];
    __shared__ __align__(8) uint64_t bar_full[kTc2MaxStages], bar_empty[kTc2MaxStages], bar_w, acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_last;

    const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const uint32_t b_bytes = (uint32_t)b_n_pad * 128u;
    const int z_slabs = (b_n_store + 63) >> 6;
    uint8_t *const w_res = smem;
    uint8_t *const staging = smem + (a.w_resident ? (size_t)a.KC * b_bytes : 0);
    uint8_t *const ring = staging + (size_t)z_slabs * kTcAStage;
    const uint32_t slot_bytes = kTcAStage + (a.w_resident ? 0u : b_bytes);

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * b_n_pad) tmem_cols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
        }
        mbar_init(&bar_w, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 1);
        }
        mbar_init_fence();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const int64_t m_tiles = (a.M + kTcBM - 1) / kTcBM;

    if (warp == 8) {
        // ---- producer ----
        if (lane == 0) {
            tma_prefetch_desc(&a.tm_x);
            if (a.w_resident) {
                mbar_expect_tx(&bar_w, (uint32_t)a.KC * b_bytes);
                for (int kc = 0; kc < a.KC; ++kc) bulk_g2s(w_res + (size_t)kc * b_bytes, bWimg + (size_t)kc * b_bytes, b_bytes, &bar_w);
            }
            uint32_t n = 0;
            for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x)
                for (int kc = 0; kc < a.KC; ++kc, ++n) {
                    const uint32_t s = n % (uint32_t)a.stages, use = n / (uint32_t)a.stages;
                    if (use > 0) mbar_wait(&bar_empty[s], (use - 1) & 1u);
                    uint8_t *slot = ring + (size_t)s * slot_bytes;
                    mbar_expect_tx(&bar_full[s], slot_bytes);
                    tma_load_2d(slot, &a.tm_x, kc * kTcBK, (int)(tile * kTcBM), &bar_full[s]);
                    if (!a.w_resident) bulk_g2s(slot + kTcAStage, bWimg + (size_t)kc * b_bytes, b_bytes, &bar_full[s]);
                }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ---- transform group + MMA issue (thread 128) ----
        const int ttid = tid - 128;
        const uint32_t idesc = make_idesc_bf16(kTcBM, b_n_pad, 0, 0);
        const int t_cc = (ttid & 7) ^ ((ttid >> 3) & 7);
        if (a.w_resident && warp == 4) mbar_wait(&bar_w, 0);
        uint32_t n = 0, t = 0;
        for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++t) {
            const uint32_t buf = t & 1u, buse = t >> 1;
            for (int kc = 0; kc < a.KC; ++kc, ++n) {
                const uint32_t s = n % (uint32_t)a.stages, par = (n / (uint32_t)a.stages) & 1u;
                uint8_t *slot = ring + (size_t)s * slot_bytes;
                if (a.in_scale) {
                    const int k = kc * kTcBK + t_cc * 8;
                    float sc[8], sh[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const bool ok = k + e < a.K;
                        sc[e] = ok ? a.in_scale[k + e] : 0.0f;
                        sh[e] = ok ? a.in_shift[k + e] : 0.0f;
                    }
                    mbar_wait(&bar_full[s], par);
                    if (k < a.K) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            uint4 *ptr = reinterpret_cast<uint4 *>(slot + (size_t)(ttid + 128 * i) * 16);
                            uint4 v = *ptr;
                            uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2) {
                                float2 f = unpack_bf16x2(w[e2]);
                                f.x = fmaxf(fmaf(f.x, sc[2 * e2], sh[2 * e2]), 0.0f);
                                f.y = fmaxf(fmaf(f.y, sc[2 * e2 + 1], sh[2 * e2 + 1]), 0.0f);
                                w[e2] = pack_bf16x2(f.x, f.y);
                            }
                            *ptr = v;
                        }
                    }
                    fence_proxy_async();
                    named_bar_sync(2, 128);
                } else if (warp == 4) {
                    mbar_wait(&bar_full[s], par);
                }
                if (warp == 4) {        // converged; one elected lane issues (tc_common.cuh: elect_one_sync)
                    if (kc == 0 && buse > 0) mbar_wait(&acc_empty[buf], (buse - 1) & 1u);   // the epilogue drained this buffer
                    fence_after_sync();
                    const int k_left = a.K - kc * kTcBK;
                    const int nk = k_left >= kTcBK ? 4 : (k_left + 15) / 16;
                    const uint64_t ad = make_desc(smem_addr(slot), 0, 1024);
                    const uint64_t bd = make_desc(smem_addr(a.w_resident ? w_res + (size_t)kc * b_bytes : slot + kTcAStage), 0, 1024);
                    const uint32_t d_acc = tmem + buf * (uint32_t)b_n_pad;
                    if (elect_one_sync()) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < nk) umma_bf16(d_acc, ad + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc, (uint32_t)((kc | j) != 0));
                        umma_commit(&bar_empty[s]);
                        if (kc == a.KC - 1) umma_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ---- epilogue group ----
        float sum1[8], sum2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) sum1[i] = sum2[i] = 0.0f;
        uint32_t t = 0;
        for (int64_t tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++t) {
            const int64_t m0 = tile * kTcBM;
            const uint32_t buf = t & 1u;
            mbar_wait(&acc_full[buf], (t >> 1) & 1u);
            fence_after_sync();
            const uint32_t taddr = tmem + buf * (uint32_t)b_n_pad + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < b_n_store; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
                if (b_bias) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < bN) v[i] += b_bias[c0 + i];
                }
                uint4 lo, hi;
                lo.x = pack_bf16x2(v[0], v[1]);   lo.y = pack_bf16x2(v[2], v[3]);
                lo.z = pack_bf16x2(v[4], v[5]);   lo.w = pack_bf16x2(v[6], v[7]);
                hi.x = pack_bf16x2(v[8], v[9]);   hi.y = pack_bf16x2(v[10], v[11]);
                hi.z = pack_bf16x2(v[12], v[13]); hi.w = pack_bf16x2(v[14], v[15]);
                uint8_t *slab = staging + (size_t)(c0 >> 6) * kTcAStage;
                const int ch = (c0 & 63) >> 3;
                *reinterpret_cast<uint4 *>(slab + sw128_offset(tid, ch)) = lo;
                if (c0 + 8 < b_n_store) *reinterpret_cast<uint4 *>(slab + sw128_offset(tid, ch + 1)) = hi;
            }
            fence_before_sync();     // this thread's TMEM reads are complete
            fence_proxy_async();     // staging writes -> visible to the TMA store
            named_bar_sync(1, 128);
            if (tid == 0) {
                mbar_arrive(&acc_empty[buf]);    // the MMAs of tile t+2 may overwrite this accumulator buffer
                for (int j = 0; j < z_slabs; ++j) tma_store_2d(b_tm_z, 64 * j, (int)m0, staging + (size_t)j * kTcAStage);
            }
            if (STATS) {
                const int rows = (int)max((int64_t)0, min((int64_t)32, a.M - (m0 + warp * 32)));
                tile_col_stats(staging, warp, lane, bN, rows, sum1, sum2);
            }
            if (tid == 0) tma_store_wait_read();
            named_bar_sync(1, 128);   // staging free again
        }
        if (tid == 0) tma_store_wait_all();

        if (STATS) {
            fold_split_stats(lane, bN, sum1, sum2);
            float(*s_part)[2][256] = reinterpret_cast<float(*)[2][256]>(staging);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int p = lane + 32 * j;
                s_part[warp][0][2 * p] = sum1[2 * j];     s_part[warp][0][2 * p + 1] = sum1[2 * j + 1];
                s_part[warp][1][2 * p] = sum2[2 * j];     s_part[warp][1][2 * p + 1] = sum2[2 * j + 1];
            }
            named_bar_sync(1, 128);
            for (int c = tid; c < bN; c += 128) {
                float t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) { t1 += s_part[w][0][c]; t2 += s_part[w][1][c]; }
                double *acc = a.stat_accum + (size_t)(blockIdx.x % kStatReplicas) * 2 * a.stat_ld + b_stat_off;
                atomicAdd(acc + c, (double)t1);
                atomicAdd(acc + a.stat_ld + c, (double)t2);
            }
            if (b_ticket) {
                __threadfence();
                named_bar_sync(1, 128);
                if (tid == 0) s_last = atomicAdd(b_ticket, 1u) == gridDim.x - 1 ? 1 : 0;
                named_bar_sync(1, 128);
                if (s_last) {
                    __threadfence();
                    for (int c = tid; c < bN; c += 128) {
                        double s1 = 0.0, s2 = 0.0;
#pragma unroll
                        for (int r = 0; r < kStatReplicas; ++r) {
                            double *acc = a.stat_accum + (size_t)r * 2 * a.stat_ld + b_stat_off;
                            s1 += __ldcg(acc + c);
                            s2 += __ldcg(acc + a.stat_ld + c);
                            acc[c] = 0.0;
                            acc[a.stat_ld + c] = 0.0;
                        }
                        bn_finalize_channel(s1, s2, a.M, b_stat_off + c, a.fin.gamma, a.fin.beta, a.fin.conv_bias, a.fin.eps,
                                            a.fin.momentum_dev ? __ldg(a.fin.momentum_dev) : a.fin.momentum, a.fin.running_mean, a.fin.running_var, a.fin.scale, a.fin.shift,
                                            a.fin.save_mean, a.fin.save_invstd);
                    }
                    if (tid == 0) {
                        *b_ticket = 0u;
                        if (a.fin.num_batches_tracked && b_stat_off == 0) *a.fin.num_batches_tracked += 1;
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

int tc_pack_weights(int n, const float *const *W, const int *K, const int *N, const int *transposed, void *const *wpack,
                    cudaStream_t st) {
    // job i: the image tc_linear_nt streams for Z = X.W^T (transposed[i] == 0: rows = N out channels, columns = K) or
    // for dX = dZ.W (transposed[i] != 0: rows = K, columns = N), W row-major [N, K]
    for (int i0 = 0; i0 < n; i0 += kPackJobs) {
        PackJobs a;
        a.n = n - i0 < kPackJobs ? n - i0 : kPackJobs;
        int blocks = 0;
        for (int j = 0; j < a.n; ++j) {
            const int i = i0 + j;
            const int rows = transposed[i] ? K[i] : N[i], cols = transposed[i] ? N[i] : K[i];
            a.W[j] = W[i];
            a.w_sn[j] = transposed[i] ? 1 : K[i];
            a.w_sk[j] = transposed[i] ? K[i] : 1;
            a.N[j] = rows;
            a.K[j] = cols;
            a.img[j] = (uint8_t *)wpack[i];
            a.first_block[j] = blocks;
            const int KC = (cols + kTcBK - 1) / kTcBK;
            const int units = ((rows + 255) / 256) * KC * 256 * 8;
            blocks += (units + 255) / 256;
        }
        a.first_block[a.n] = blocks;
        pack_weights_multi_kernel<<<blocks, 256, 0, st>>>(a);
        count_launch();
    }
    return check_launch("pack_weights");
}

size_t tc_wpack_bytes(int K, int N) {
    // images for every 256-column block of N, each [KC][n_pad][128 B]
    size_t KC = (size_t)(K + kTcBK - 1) / kTcBK, total = 0;
    for (int n0 = 0; n0 < N; n0 += 256) total += KC * (size_t)round_up(N - n0 < 256 ? N - n0 : 256, 16) * 128;
    return total;
}

// Z[M,N] = act(X) . W^T with W element (n,k) at W[n*w_sn + k*w_sk]
int tc_linear_nt(const void *X, int ldx, const float *in_scale, const float *in_shift, const float *W, int64_t w_sn,
                 int64_t w_sk, const float *bias, int64_t M, int K, int N, void *Z, int ldz, double *stat_accum,
                 void *wpack, cudaStream_t st, bool packed, const BnFinalize *fin) {
    static bool attr_done = false;
    static int static_smem = 0, static_smem2 = 0;
    // PN2_TC2: 0 = v1 kernel (co-resident CTAs), 1 = v2 pipelined kernel with one CTA per SM, 2 = v2 with up to two,
    // 3 (default) = per layer: v1 when K is one 64-wide chunk (the sa1/sa2 layers: a tile is one load + <= 4 MMAs, the
    // co-resident CTAs win by 5-15 %), v2 (<= 2 CTAs) when K spans several chunks (fp1 128 x 128: 46 -> 34 us, sa3.3 38 -> 32 us)
    static int tc2_mode = -1;
    if (tc2_mode < 0) {
        const char *e = getenv("PN2_TC2");
        tc2_mode = e ? atoi(e) : 3;
        if (tc2_mode < 0 || tc2_mode > 3) tc2_mode = 3;
    }
    if (!attr_done) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, linear_tc_kernel<true>);
        static_smem = (int)fa.sharedSizeBytes;
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - (int)fa.sharedSizeBytes);
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, linear_tc_kernel<false>);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(linear_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - (int)fa.sharedSizeBytes);
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, linear_tc2_kernel<true>);
        static_smem2 = (int)fa.sharedSizeBytes;
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(linear_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - (int)fa.sharedSizeBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(linear_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
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
    const int64_t m_tiles = (M + kTcBM - 1) / kTcBM;
    static int max_ctas = 0;
    if (!max_ctas) {
        const char *e = getenv("PN2_TC_MAX_CTAS");
        max_ctas = e ? atoi(e) : 4;
        if (max_ctas < 1 || max_ctas > 8) max_ctas = 4;
    }
    // output columns in blocks of 256 (one accumulator tile); up to kTcMaxNBlocks blocks share ONE launch as blockIdx.y
    for (int g0 = 0; g0 < N; g0 += 256 * kTcMaxNBlocks) {
        const int gN = N - g0 < 256 * kTcMaxNBlocks ? N - g0 : 256 * kTcMaxNBlocks;
        const int ng = (gN + 255) / 256;
        TcLinearArgs a;
        memset(&a.fin, 0, sizeof(a.fin));
        if (fin && stat_accum) a.fin = *fin;
        bool maps_ok = make_rows_tensor_map(&a.tm_x, X, M, ldx, ldx, kTcBM);
        size_t group_img_bytes = 0;
        for (int y = 0; y < ng; ++y) {
            const int n0 = g0 + 256 * y;
            const int nb = N - n0 < 256 ? N - n0 : 256;
            const int n_pad_y = round_up(nb, 16);
            int n_store_y = round_up(nb, 8);
            if (n0 + n_store_y > ldz) n_store_y = ldz - n0;      // ldz is a multiple of 8 on this path
            if (!packed) {
                const int total = KC * n_pad_y * 8;
                pack_weight_kernel<<<(total + 255) / 256, 256, 0, st>>>(W + (int64_t)n0 * w_sn, w_sn, w_sk, nb, K, n_pad_y, KC, 0,
                                                                        img + group_img_bytes);
                count_launch();
            }
            group_img_bytes += (size_t)KC * n_pad_y * 128;
            maps_ok = maps_ok && make_rows_tensor_map(&a.tm_z[y], (const __nv_bfloat16 *)Z + n0, M, n_store_y, ldz, kTcBM);
        }
        for (int y = ng; y < kTcMaxNBlocks; ++y) a.tm_z[y] = a.tm_z[0];
        if (!maps_ok) {
            set_error("linear_tc: cuTensorMapEncodeTiled failed (M=%lld ldx=%d ldz=%d)", (long long)M, ldx, ldz);
            return PN2_ERR_CUDA;
        }
        // the plan below is made for the group's first (widest) block; narrower last blocks use less of the same carve-up
        const int nb = gN < 256 ? gN : 256;
        const int n_pad = round_up(nb, 16);
        int n_store = round_up(nb, 8);
        if (g0 + n_store > ldz) n_store = ldz - g0;
        const size_t img_bytes = (size_t)KC * n_pad * 128;
        a.in_scale = in_scale;
        a.in_shift = in_shift;
        a.Wimg = img;
        a.bias = bias ? bias + g0 : nullptr;
        a.M = M; a.K = K; a.N = gN; a.ldz = ldz - g0; a.KC = KC;
        a.stat_accum = stat_accum;
        a.stat_ld = N;
        a.stat_off = g0;
        // shared-memory plan: resident W when its image is <= 64 KB; as many ring slots as fit the per-CTA budget
        // of 3, 2 or 1 CTAs per SM (whichever is the densest that still leaves >= 3 slots, at most 8)
        a.w_resident = img_bytes <= 64 * 1024 ? 1 : 0;
        const size_t fixed = 1024 + (a.w_resident ? img_bytes : 0) + (size_t)((n_store + 63) / 64) * kTcAStage;
        const size_t slot = kTcAStage + (a.w_resident ? 0 : (size_t)n_pad * 128);
        // CTAs per SM: these layers stream rows and every tile is a serial chain (TMA -> transform -> MMA -> epilogue ->
        // store), so co-resident CTAs are what hides the latency.  Take the densest packing (<= max_ctas, tensor memory
        // allowing: n * columns <= 512) that still leaves a ring of >= 2 slots, except that 1-2 CTAs/SM prefer >= 3 slots.
        // Measured: the 128 x 128 layers missed a third slot at 2/SM by 160 bytes, ran 1/SM at 0.8 TB/s, tensor pipe 5 % busy.
        const int v2_ctas = tc2_mode == 3 ? (KC >= 2 ? 2 : 0) : tc2_mode;
        bool launched = false;
        if (v2_ctas > 0) {
            // v2: the ring hides the latency -- one CTA per SM with every byte of shared memory that is left as ring slots
            // (or two CTAs when each still gets >= 4 slots and 2 x 2 accumulator buffers fit tensor memory)
            uint32_t cols2 = 32;
            while ((int)cols2 < 2 * n_pad) cols2 <<= 1;
            int stages2 = 0, per_sm2 = 1;
            for (int n = v2_ctas; n >= 1 && stages2 == 0; --n) {
                if (n * (int)cols2 > 512) continue;
                const size_t budget = (size_t)(233472 / n) - 1024 - (size_t)static_smem2;
                if (budget < fixed + slot * (size_t)(n > 1 ? 4 : 2)) continue;
                stages2 = (int)((budget - fixed) / slot);
                per_sm2 = n;
            }
            if (stages2 >= 2) {
                if (stages2 > kTc2MaxStages) stages2 = kTc2MaxStages;
                a.stages = stages2;
                const size_t dyn2 = fixed + slot * stages2;
                int64_t grid2 = ((int64_t)per_sm2 * sm_budget() + ng - 1) / ng;
                if (grid2 > m_tiles) grid2 = m_tiles;
                const dim3 grid((unsigned)grid2, (unsigned)ng);
                if (stat_accum)
                    linear_tc2_kernel<true><<<grid, kLinTc2Threads, dyn2, st>>>(a);
                else
                    linear_tc2_kernel<false><<<grid, kLinTc2Threads, dyn2, st>>>(a);
                count_launch();
                int rc2 = check_launch("linear_tc2");
                if (rc2 != PN2_OK) return rc2;
                launched = true;
            }
        }
        if (!launched) {
            uint32_t tmem_cols = 32;
            while ((int)tmem_cols < n_pad) tmem_cols <<= 1;
            int stages = 0, per_sm = 1;
            for (int pass = 0; pass < 2 && stages == 0; ++pass)
                for (int n = max_ctas; n >= 1 && stages == 0; --n) {
                    if (n * (int)tmem_cols > 512) continue;
                    const size_t budget = (size_t)(233472 / n) - 1024 - (size_t)static_smem;
                    const int min_slots = (pass == 0 && n <= 2) ? 3 : 2;
                    if (budget < fixed + slot * (size_t)min_slots) continue;
                    stages = (int)((budget - fixed) / slot);
                    per_sm = n;
                }
            if (stages < 2) {
                set_error("linear_tc: layer K=%d N=%d does not fit shared memory", K, nb);
                return PN2_ERR_UNSUPPORTED;
            }
            if (stages > 8) stages = 8;
            a.stages = stages;
            const size_t dyn = fixed + slot * stages;
            int64_t gx = ((int64_t)per_sm * sm_budget() + ng - 1) / ng;
            if (gx > m_tiles) gx = m_tiles;
            const dim3 grid((unsigned)gx, (unsigned)ng);
            if (stat_accum)
                linear_tc_kernel<true><<<grid, kLinTcThreads, dyn, st>>>(a);
            else
                linear_tc_kernel<false><<<grid, kLinTcThreads, dyn, st>>>(a);
            count_launch();
            int rc = check_launch("linear_tc");
            if (rc != PN2_OK) return rc;
        }
        img += group_img_bytes;
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
    CUtensorMap tm_dz, tm_x;   // [M, lddz] / [M, ldx] bf16, boxes of 64 columns x R rows, 128-byte swizzle
    const float *in_scale, *in_shift;
    int64_t M, rows_per_split;   // rows_per_split % R == 0: only the global tail is a partial (zero-filled) box
    int K, N, K_ld;    // K_ld: row stride of the fp32 partials (K rounded up to 4 -> 16-byte stores)
    int lddz, ldx;
    int nblk_k;        // blockIdx.y = (dW row block of 128) * nblk_k + (dW column block of 256)
    int a_slabs, b_slabs, R;   // ring-slot geometry of the largest block
    float *scratch;    // [splits][N][K_ld]
    float *dW_accum;   // non-null: no partials -- every CTA adds its block straight into dW[N][K] (red.global.add.f32)
    int accum_vec4;    // dW is 16-byte aligned and K % 4 == 0: red.global.add.v4.f32
};

__device__ __forceinline__ void red_add_f32(float *p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4_f32(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int kWgStages = 4;           // ring slots
constexpr int kWgThreads = 160;        // warps 0-3: activation transform + epilogue (thread 0 issues the MMAs); warp 4: TMA producer

__global__ void __launch_bounds__(kWgThreads) wgrad_tc_kernel(const __grid_constant__ TcWgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kWgStages], bar_empty[kWgStages], bar_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_scale[256], s_shift[256];

    const int tid = threadIdx.x, warp = uniform_warp_idx();
    const int n0 = ((int)blockIdx.y / a.nblk_k) * 128, k0 = ((int)blockIdx.y % a.nblk_k) * 256;
    const int nb = min(128, a.N - n0), kb = min(256, a.K - k0);
    const int kb_pad = (kb + 15) & ~15;
    const int a_sl = (nb + 63) >> 6, b_sl = (kb + 63) >> 6;      // slabs this block really needs
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const uint32_t slab_bytes = (uint32_t)a.R * 128u;
    const uint32_t stage_bytes = slab_bytes * (uint32_t)(a.a_slabs + a.b_slabs);
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < kb_pad) tmem_cols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < kWgStages; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
        }
        mbar_init(&bar_done, 1);
        mbar_init_fence();
    }
    if (a.in_scale)
        for (int i = tid; i < kb; i += kWgThreads) {
            s_scale[i] = a.in_scale[k0 + i];
            s_shift[i] = a.in_shift[k0 + i];
        }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    const int64_t r_begin = (int64_t)blockIdx.x * a.rows_per_split;
    const int64_t r_end = min(a.M, r_begin + a.rows_per_split);
    const int n_it = r_end > r_begin ? (int)((r_end - r_begin + a.R - 1) / a.R) : 0;

    if (warp == 4) {
        // ---- producer: one thread streams (dZ, X) row boxes into the ring with tensor-map TMA ----
        if ((tid & 31) == 0) {
            tma_prefetch_desc(&a.tm_dz);
            tma_prefetch_desc(&a.tm_x);
            const uint32_t tx = slab_bytes * (uint32_t)(a_sl + b_sl);
            for (int it = 0; it < n_it; ++it) {
                const int s = it % kWgStages;
                if (it >= kWgStages) mbar_wait(&bar_empty[s], (uint32_t)((it / kWgStages - 1) & 1));
                const int r0 = (int)(r_begin + (int64_t)it * a.R);
                uint8_t *stA = smem + (size_t)s * stage_bytes;
                uint8_t *stB = stA + (size_t)a.a_slabs * slab_bytes;
                mbar_expect_tx(&bar_full[s], tx);
                for (int j = 0; j < a_sl; ++j) tma_load_2d(stA + (size_t)j * slab_bytes, &a.tm_dz, n0 + 64 * j, r0, &bar_full[s]);
                for (int j = 0; j < b_sl; ++j) tma_load_2d(stB + (size_t)j * slab_bytes, &a.tm_x, k0 + 64 * j, r0, &bar_full[s]);
            }
        }
        __syncwarp();
    } else {
        const uint32_t idesc = make_idesc_bf16(128, kb_pad, 1, 1);
        for (int it = 0; it < n_it; ++it) {
            const int s = it % kWgStages;
            const uint32_t par = (uint32_t)((it / kWgStages) & 1);
            const int64_t r0 = r_begin + (int64_t)it * a.R;
            const int rows = (int)min((int64_t)a.R, r_end - r0);
            const int rows16 = (rows + 15) & ~15;
            uint8_t *stA = smem + (size_t)s * stage_bytes;
            uint8_t *stB = stA + (size_t)a.a_slabs * slab_bytes;
            if (a.in_scale) {
                // act(X) = relu(bn(.)) of the previous layer, in place; threads walk PHYSICAL 16-byte units (conflict-free).
                // Unit q = tid + 128 i sits in row r = q >> 3 at logical column chunk cc = (q & 7) ^ (r & 7); both q & 7 and
                // r & 7 depend on tid only, so a thread transforms the SAME 8 columns of every row it touches: their
                // scale/shift live in registers per slab (columns past kb get scale = shift = 0 -> stay 0), two units in flight
                mbar_wait(&bar_full[s], par);
                const int cc = (tid & 7) ^ ((tid >> 3) & 7);
                const int n_units = rows16 * 8;
                for (int j = 0; j < b_sl; ++j) {
                    const int i0 = j * 64 + cc * 8;
                    if (i0 >= kb) continue;
                    float sc[8], sh[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const bool ok = i0 + e < kb;
                        sc[e] = ok ? s_scale[i0 + e] : 0.0f;
                        sh[e] = ok ? s_shift[i0 + e] : 0.0f;
                    }
                    uint8_t *slab = stB + (size_t)j * slab_bytes;
#pragma unroll 2
                    for (int q = tid; q < n_units; q += 128) {
                        uint4 *ptr = reinterpret_cast<uint4 *>(slab + (size_t)q * 16);
                        uint4 v = *ptr;
                        uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            float2 f = unpack_bf16x2(w[e2]);
                            f.x = fmaxf(fmaf(f.x, sc[2 * e2], sh[2 * e2]), 0.0f);
                            f.y = fmaxf(fmaf(f.y, sc[2 * e2 + 1], sh[2 * e2 + 1]), 0.0f);
                            w[e2] = pack_bf16x2(f.x, f.y);
                        }
                        *ptr = v;
                    }
                }
                fence_proxy_async();
                named_bar_sync(1, 128);
            } else if (warp == 0) {
                mbar_wait(&bar_full[s], par);
            }
            if (warp == 0) {            // converged; one elected lane issues (tc_common.cuh: elect_one_sync)
                fence_after_sync();
                const uint64_t ad = make_desc(smem_addr(stA), slab_bytes, 1024), bd = make_desc(smem_addr(stB), slab_bytes, 1024);
                const int nj = rows16 / 16;
                if (elect_one_sync()) {
                    for (int j = 0; j < nj; ++j)
                        umma_bf16(tmem, ad + (uint64_t)(128 * j), bd + (uint64_t)(128 * j), idesc, (uint32_t)((it | j) != 0));
                    umma_commit(&bar_empty[s]);
                    if (it == n_it - 1) umma_commit(&bar_done);
                }
                __syncwarp();
            }
        }
        if (n_it > 0) mbar_wait(&bar_done, 0);
        fence_after_sync();
        // ---- epilogue: this thread's dW row n0 + tid, fp32 partial, 16-byte stores ----
        const int n = n0 + tid;
        float *out = a.scratch + ((size_t)blockIdx.x * a.N + n) * a.K_ld + k0;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        if (a.dW_accum) {
            // accumulate mode: dW was zeroed by its owner (the gradient sink, once per step); fp32 reductions at L2
            float *acc = a.dW_accum + (size_t)n * a.K + k0;
            if (n_it > 0)
                for (int c0 = 0; c0 < kb_pad; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);
                    if (tid < nb) {
                        if (a.accum_vec4) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4)
                                if (c0 + i < kb) red_add_v4_f32(acc + c0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < kb) red_add_f32(acc + c0 + i, v[i]);
                        }
                    }
                }
        } else
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

// fixed-order sum of `splits` fp32 partial blocks [splits][N][K_ld] into dW[N][K] (also used by bwd_fused.cu)
int tc_wgrad_reduce(const float *scratch, int splits, int N, int K, int K_ld, float *dW, cudaStream_t st) {
    const int64_t NK = (int64_t)N * K;
    wgrad_reduce_kernel2<<<(unsigned)((NK + 31) / 32), dim3(32, 8), 0, st>>>(scratch, splits, N, K, K_ld, dW);
    count_launch();
    return check_launch("wgrad_reduce");
}

struct WgradPlan {
    int splits, nblk_n, nblk_k, a_slabs, b_slabs, R, K_ld;
    int64_t rows_per_split;
    size_t dyn_smem;
};

static WgradPlan tc_wgrad_plan(int64_t M, int K, int N, bool accumulate = false) {
    WgradPlan p;
    p.nblk_n = (N + 127) / 128;
    p.nblk_k = (K + 255) / 256;
    p.K_ld = round_up(K, 4);
    p.a_slabs = ((N < 128 ? N : 128) + 63) / 64;
    p.b_slabs = ((K < 256 ? K : 256) + 63) / 64;
    const int per_row = 128 * (p.a_slabs + p.b_slabs);
    // ring-slot size and CTAs per SM (tuning knobs PN2_WG_SLOT_KB / PN2_WG_CTAS).  Measured on B200 (profiles/microbench_mlp.py):
    // the 1 M-row sa1 layers want two co-resident CTAs with 26 KB slots (60 us vs 82 us with one CTA of 40 KB slots); every
    // smaller layer is faster with ONE CTA per SM and 40 KB slots -- half the splits, so half the partial blocks to add into
    // dW, and longer TMA boxes (fp2.2 45 -> 33 us, sa4.2 35 -> 27 us, fp1.1 34 -> 27 us)
    static int env_slot_kb = -1, env_ctas = -1;
    if (env_slot_kb < 0) {
        const char *e = getenv("PN2_WG_SLOT_KB");
        env_slot_kb = e ? atoi(e) : 0;
        if (env_slot_kb < 4 || env_slot_kb > 52) env_slot_kb = 0;
        e = getenv("PN2_WG_CTAS");
        env_ctas = e ? atoi(e) : 0;
        if (env_ctas < 1 || env_ctas > 8) env_ctas = 0;
    }
    const bool big = M >= (int64_t)1 << 19;
    const int slot_kb = env_slot_kb ? env_slot_kb : (big ? 26 : 40);
    const int ctas = env_ctas ? env_ctas : (big ? 2 : 1);
    int R = (slot_kb * 1024 / per_row) / 16 * 16;
    if (R > 256) R = 256;
    if (R < 16) R = 16;
    p.R = R;
    p.dyn_smem = 1024 + kWgStages * (size_t)R * per_row;
    // splits: fill the machine (two CTAs per SM), but keep >= 2 ring iterations per CTA, and balance the
    // per-CTA iteration chain (~1 us each) against writing + re-reading the fp32 partials
    const int blocks = p.nblk_n * p.nblk_k;
    int64_t s = (ctas * kNumSMs + blocks - 1) / blocks;
    const int64_t max_rows = (M + 2 * R - 1) / (2 * R);
    // partial cost per split: write + re-read of an fp32 block (reduce kernel), or -- accumulate mode -- one pass of
    // L2 reductions, which also lets the small layers use every SM
    const double chain = (double)M / R * 1e-6, part = (accumulate ? 0.5 : 2.0) * N * p.K_ld * 4.0 / 3.0e12;
    int64_t s_bal = (int64_t)(sqrt(chain / part) + 0.5);
    if (s > max_rows) s = max_rows;
    if (s > s_bal) s = s_bal;
    if (s < 1) s = 1;
    p.rows_per_split = ((M + s - 1) / s + R - 1) / R * R;
    p.splits = (int)((M + p.rows_per_split - 1) / p.rows_per_split);
    return p;
}

size_t tc_wgrad_scratch_bytes(int64_t M, int K, int N) {
    const WgradPlan p = tc_wgrad_plan(M, K, N);
    return sizeof(float) * (size_t)p.splits * (size_t)N * (size_t)p.K_ld + 16;
}

// scratch == nullptr: accumulate mode, dW += dZ^T act(X) (dW zeroed by the caller); else dW = ... through fp32 partials
int tc_linear_wgrad(const void *dZ, int lddz, const void *X, int ldx, const float *in_scale, const float *in_shift,
                    int64_t M, int K, int N, float *dW, void *scratch, cudaStream_t st) {
    const bool accumulate = scratch == nullptr;
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
    const WgradPlan p = tc_wgrad_plan(M, K, N, accumulate);
    TcWgradArgs a;
    if (!make_rows_tensor_map(&a.tm_dz, dZ, M, lddz, lddz, p.R) || !make_rows_tensor_map(&a.tm_x, X, M, ldx, ldx, p.R)) {
        set_error("wgrad_tc: cuTensorMapEncodeTiled failed (M=%lld lddz=%d ldx=%d R=%d)", (long long)M, lddz, ldx, p.R);
        return PN2_ERR_CUDA;
    }
    a.lddz = lddz; a.ldx = ldx;
    a.in_scale = in_scale; a.in_shift = in_shift;
    a.M = M; a.rows_per_split = p.rows_per_split; a.K = K; a.N = N; a.K_ld = p.K_ld;
    a.nblk_k = p.nblk_k; a.a_slabs = p.a_slabs; a.b_slabs = p.b_slabs; a.R = p.R;
    a.scratch = (float *)(((uintptr_t)scratch + 15) & ~(uintptr_t)15);
    a.dW_accum = accumulate ? dW : nullptr;
    a.accum_vec4 = (accumulate && ((uintptr_t)dW & 15) == 0 && K % 4 == 0) ? 1 : 0;
    wgrad_tc_kernel<<<dim3((unsigned)p.splits, (unsigned)(p.nblk_n * p.nblk_k)), kWgThreads, p.dyn_smem, st>>>(a);
    count_launch();
    int rc = check_launch("wgrad_tc");
    if (rc != PN2_OK || accumulate) return rc;
    const int64_t NK = (int64_t)N * K;
    wgrad_reduce_kernel2<<<(unsigned)((NK + 31) / 32), dim3(32, 8), 0, st>>>(a.scratch, p.splits, N, K, p.K_ld, dW);
    count_launch();
    return check_launch("wgrad_reduce");
}

}  // namespace pn2
