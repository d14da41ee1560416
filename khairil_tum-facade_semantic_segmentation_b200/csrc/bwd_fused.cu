// bwd_fused.cu -- one backward step of an MLP layer on bf16 rows as ONE tcgen05 kernel.
//
// Reference semantics: autograd through relu(bn(conv(x))) of PointNetSetAbstraction / PointNetFeaturePropagation
// (/root/reference/models/pointnet2_utils.py:196-198, :311-314).  For layer l with input X (the previous layer's
// pre-BatchNorm product Z_{l-1}, or the MLP's input rows) the layer-by-layer path ran five launches over [M, C] rows:
//
//     reduce(l-1)  dgamma_{l-1}, dbeta_{l-1} = sum_m mask.dA_{l-1}.zhat, sum_m mask.dA_{l-1}            (reads dA, Z)
//     dz(l)        dZ_l = gamma.invstd.(mask.dA_l - dbeta/M - zhat.dgamma/M)                              (reads dA, Z; writes dZ)
//     dgrad(l)     dA_{l-1} = dZ_l . W_l                                                                 (reads dZ; writes dA)
//     wgrad(l)     dW_l = dZ_l^T . relu(bn(Z_{l-1}))                                                     (reads dZ, Z_{l-1})
//
// Here a CTA owns 128-row tiles and does all of it while the tile is in shared / tensor memory:
//   * TMA loads the (dA_l, Z_l, X) tiles; dZ_l is formed IN PLACE on the dA tile (never stored);
//   * the X tile becomes relu(bn(.)) (the weight gradient's operand) and zhat (the statistics' operand);
//   * tcgen05.mma: dA_{l-1} = dZ.W (K-major A, resident W image), dW += dZ^T.act (both operands MN-major straight from
//     the row tiles, accumulated in tensor memory over ALL tiles of the CTA, added to dW with L2 reductions at the end);
//   * epilogue: dA_{l-1} leaves tensor memory once, is masked by layer l-1's ReLU, rounded to bf16, staged and stored by
//     TMA; the staged tile is also the MN-major A operand of two more MMAs that accumulate
//         S2[k,k'] += sum_m dA'[m,k] zhat[m,k']   (its diagonal is dgamma_{l-1})   and   S1[k,.] += sum_m dA'[m,k] . 1
//     in tensor memory -- the column statistics cost no issue slots on the FMA pipes;
//   * the CTA that draws the last ticket turns the fp64 accumulators into dgamma_{l-1} / dbeta_{l-1}.
// HBM traffic per layer: dA_l + Z_l + X read, dA_{l-1} written (4 passes over [M, C] instead of 10), one launch instead of 4-5.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pn2 {

using namespace tc;

int tc_wgrad_reduce(const float *scratch, int splits, int N, int K, int K_ld, float *dW, cudaStream_t st);   // linear_tc.cu

constexpr int kBfBM = 128;
constexpr int kBfSlab = kBfBM * 128;          // one 64-column slab of a 128-row tile: 16 KB (SW128 box layout)
constexpr int kBfEpi = 128;                   // warps 0-3 : epilogue group (tensor-memory lane quarters), statistics MMAs
constexpr int kBfXf = 256;                    // warps 4-11: transform group
constexpr int kBfThreads = kBfEpi + kBfXf + 64;   // warp 12: TMA producer, warp 13: MMA issue (every product of the kernel)
constexpr int kBfMaxC = 128;                  // widest layer (N) / input (K_ld) this kernel takes
constexpr int kBfMaxStages = 4;

struct BwdFusedArgs {
    CUtensorMap tm_da, tm_z, tm_x, tm_dx;
    const float *scale, *shift, *mean, *invstd, *dgamma, *dbeta;   // layer l; mean == null: frozen statistics
    const float *p_scale, *p_shift, *p_mean, *p_invstd;            // layer l-1 (X = Z_{l-1}); null: X is the plain MLP input
    const uint8_t *Wimg;         // transposed image of W_l: [nS chunks][K_pad rows][128 B] (tc_pack_weights, transposed)
    int64_t M;
    int K, N, K_pad, k_store;    // K_pad = round_up(K, 16); k_store = columns of dX written (round_up(K, 8) <= lddx)
    int nS, kS;                  // 64-column slabs of the dA / Z tiles and of the X tile
    int da_mode;                 // 0: dA dense, mask applied here; 1: dA already masked by its producer; 2: pooled (dOut, arg); 3: dZ_l given
    const float *dOut;           // mode 2: [G, N] fp32 gradient of the max-pooled output, arg [G, N] int32 winners (32 samples per group)
    const int32_t *arg;
    uint32_t o_pool;             // per stage: [4][N] fp32 dOut rows then [4][N] int32 arg rows of the tile's four groups
    int want_dx, want_dw, want_stats, load_x;
    int ones_col;                // >= 0: the statistics' ones live in columns [ones_col, ones_col + 16) of the zhat tile; < 0: own tile
    int s2_cols;                 // N extent of the S2 product
    float *dW;                   // accumulate target [N, K] (L2 reductions) or null when scratch is used
    float *scratch;              // [gridDim.x][N][K_ld4] fp32 partial blocks (K % 4 != 0) or null
    int dw_vec4, K_ld4;
    double *stat_accum;          // [replicas][2][K] fp64
    unsigned *ticket;
    float *dgamma_prev, *dbeta_prev;
    float inv_m;
    int tmem_cols, acc_bufs, off_dw, off_s2, off_s1;
    int defer_store_wait;
    int combined;                // one product [dZ | dA']^T . [act | zhat | ones] gives the weight gradient AND the statistics
    int stages, alias_act;       // alias_act: act lives in the (dead) Z tile -> the transform group syncs between T1 and T2
    // shared-memory byte offsets from the 1024-aligned base; per-stage buffers are stage_stride apart
    uint32_t o_w, o_da, o_z, o_x, o_act, o_stage, o_ones, o_coef, stage_stride, bytes_a, bytes_b;
    int coef_ld;
    int coef_global;             // 1: every per-column coefficient is formed from the layers' arrays on the fly (no shared tables);
                                 // 2: only layer l-1's (the T2 phase) -- layer l's three tables (sc, cb, cc) stay in shared memory
    long long *dbg_buf;
    int dbg;                     // PN2_BWD_DBG bit mask (profiling experiments): 1 skip wgrad MMAs, 2 skip statistics MMAs, 4 skip T1/T2
};

__device__ __forceinline__ void bf_red_add_f32(float *p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void bf_red_add_v4_f32(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// un-swizzled K-major descriptor with zero strides: every core matrix of the operand is the 128 bytes at saddr
__device__ __forceinline__ uint64_t make_desc_ones(uint32_t saddr) { return (1ull << 46) | (uint64_t)((saddr >> 4) & 0x3fff); }
__device__ __forceinline__ void bf_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// PN2_BWD_DBG & 32: CTA 0 stamps clock64 at the phase boundaries of its first 24 tiles into `scratch` ([24][16] int64)
#define BF_STAMP(slot)                                                                                   \
    do {                                                                                                 \
        if ((a.dbg & 32) && blockIdx.x == 0 && t < 24 && a.dbg_buf)                                      \
            a.dbg_buf[t * 16 + (slot)] = clock64();                                                      \
    } while (0)

// Which 16-byte unit of a 64-column slab a transform thread owns in iteration i, for a tile whose rows hold `nch` valid chunks.
//   nch == 8 (or anything else): physical unit q = tid + 256 i: row q >> 3, logical chunk (q & 7) ^ (row & 7) -- constant per
//            thread, every thread busy, conflict free (4 iterations per slab);
//   nch == 4: only half of every 128-byte row is data.  A warp takes one 8-row swizzle atom per iteration; its four 8-thread
//            phases take the row pairs (p, p + 4): logical chunks 0-3 of row p sit in the physical chunks {0..3} ^ p and those
//            of row p + 4 in the complementary half, so a phase touches all 32 banks once (2 iterations per slab).
// eight consecutive per-column coefficients (c0 % 8 == 0, tables 32-byte aligned): two LDS.128 instead of eight LDS
__device__ __forceinline__ void ld_coef8(const float *tab, int c0, float (&v)[8]) {
    const float4 lo = *reinterpret_cast<const float4 *>(tab + c0), hi = *reinterpret_cast<const float4 *>(tab + c0 + 4);
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

__device__ __forceinline__ void ldg_coef8(const float *tab, int c0, float (&v)[8]) {
    const float4 lo = __ldg(reinterpret_cast<const float4 *>(tab + c0)), hi = __ldg(reinterpret_cast<const float4 *>(tab + c0 + 4));
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

struct UnitMap {
    int chunk;        // logical 16-byte chunk (8 columns) inside the slab, constant per thread
    int row0, rstep;  // rows row0 + rstep * i
    int iters;
};
__device__ __forceinline__ UnitMap unit_map(int tid, int nch) {
    UnitMap m;
    if (nch == 4) {
        const int w = tid >> 5, l = tid & 31;
        m.chunk = l & 3;
        m.row0 = 8 * w + (l >> 3) + 4 * ((l >> 2) & 1);
        m.rstep = 64;
        m.iters = 2;
    } else {
        m.chunk = (tid & 7) ^ ((tid >> 3) & 7);
        m.row0 = tid >> 3;
        m.rstep = 32;
        m.iters = 4;
    }
    return m;
}

// Three warp-specialised roles pipelined over the CTA's tiles (one CTA per SM):
//   producer  : TMA loads of tile t+1 (t+2) while tile t is worked on -- group A = (dA, Z) is released as soon as the data /
//               weight-gradient MMAs have read it, group B = (X, act, staging) when the statistics MMAs have;
//   transform : dZ in place, X -> (act, zhat), then ONE thread issues the tile's data- and weight-gradient MMAs;
//   epilogue  : dA_{l-1} out of tensor memory -> ReLU mask -> bf16 -> staging -> TMA store, then ONE thread issues the
//               statistics MMAs on the staged tile.  The data-gradient accumulator is double buffered when it fits.
__global__ void __launch_bounds__(kBfThreads, 1) bwd_fused_kernel(const __grid_constant__ BwdFusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_a[kBfMaxStages], full_b[kBfMaxStages], empty[kBfMaxStages], act_rdy[kBfMaxStages],
        acc_full[2], acc_empty[2], e_rdy[kBfMaxStages], bar_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_last;

    const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    uint8_t *const s_w = smem + a.o_w, *const s_ones = smem + a.o_ones;
    // per-column coefficients: layer l  dz = sc.g + cb.z + cc (mask test sc.z + sh > 0), layer l-1  act = relu(psc.z + psh)
    float *const coef = reinterpret_cast<float *>(smem + a.o_coef);
    const int cl = a.coef_ld;
    // (coef_global 2 keeps three tables only: sc, cb, cc at 0, 1, 2 -- the shift table is not needed by da_mode 1)
    float *const c_sc = coef, *const c_sh = coef + (a.coef_global == 2 ? 0 : cl), *const c_b = coef + (a.coef_global == 2 ? 1 : 2) * cl,
                 *const c_c = coef + (a.coef_global == 2 ? 2 : 3) * cl;
    float *const p_sc = coef + 4 * cl, *const p_sh = coef + 5 * cl;
    const bool prev = a.p_scale != nullptr;
    const int S = a.stages, AB = a.acc_bufs;

    if (warp == 0) tmem_alloc(&tmem_base_s, (uint32_t)a.tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < kBfMaxStages; ++i) {
            mbar_init(&full_a[i], 1); mbar_init(&full_b[i], 1); mbar_init(&empty[i], 1); mbar_init(&act_rdy[i], 1); mbar_init(&e_rdy[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 1); }
        mbar_init(&bar_done, 1);
        mbar_init_fence();
    }
    for (int c = a.coef_global == 1 ? cl : tid; c < cl; c += kBfThreads) {
        float sc = 0.f, sh = 0.f, cb = 0.f, cc = 0.f;
        if (c < a.N && a.da_mode != 3) {
            sc = a.scale[c];
            sh = a.shift[c];
            if (a.mean) {      // train: dz = sc.(g - dbeta/M - zhat.dgamma/M) = sc.g + cb.z + cc
                cb = -sc * a.dgamma[c] * a.invstd[c] * a.inv_m;
                cc = -sc * a.dbeta[c] * a.inv_m - cb * a.mean[c];
            }
        }
        if (a.coef_global != 2) c_sh[c] = sh;
        c_sc[c] = sc; c_b[c] = cb; c_c[c] = cc;
        if (a.coef_global) continue;
        float ps = 0.f, ph = 0.f;
        if (prev && c < a.K) { ps = a.p_scale[c]; ph = a.p_shift[c]; }
        p_sc[c] = ps; p_sh[c] = ph;
    }
    // the all-ones K-major B operand of the S1 product: ONE 128-byte core matrix (8 n-rows x 8 k) in the un-swizzled layout,
    // read for every (n group, k step) through a descriptor whose strides are 0 -- every element is 1.0, so 128 bytes stand in
    // for the [16 x 128] tile (it used to be a 4 KB swizzled tile: the bytes that kept the 128-wide layers at one stage)
    if (a.want_stats && a.ones_col < 0 && tid < 32) reinterpret_cast<uint32_t *>(s_ones)[tid] = 0x3f803f80u;
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const int64_t m_tiles = (a.M + kBfBM - 1) / kBfBM;
    const uint32_t n_my = (uint32_t)(((m_tiles - 1 - (int64_t)blockIdx.x) / (int64_t)gridDim.x) + 1);   // tiles of this CTA (>= 1)
    const uint32_t w_chunk_bytes = (uint32_t)a.K_pad * 128u;
    // MN-major operands of a one-slab tile: the M = 128 product reads "the next slab" through LBO; 0 makes it the same slab
    // again (rows 64-127 of those products are never used)
    const uint32_t lbo_n = a.nS > 1 ? (uint32_t)kBfSlab : 0u, lbo_k = a.kS > 1 ? (uint32_t)kBfSlab : 0u;

    if (warp == (kBfEpi + kBfXf) / 32) {
        // ================= producer =================
        if (lane == 0) {
            tma_prefetch_desc(&a.tm_da);
            tma_prefetch_desc(&a.tm_x);
            // (stage index / use count kept incrementally: S is a run-time value, t % S and t / S are ~100-cycle divisions)
            for (uint32_t t = 0, s = 0, u = 0; t < n_my; ++t, s = (s + 1 == (uint32_t)S ? 0u : s + 1), u += (s == 0)) {
                const int m0 = (int)(((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * kBfBM);
                uint8_t *st = smem + (size_t)s * a.stage_stride;
                BF_STAMP(0);
                if (u > 0) mbar_wait(&empty[s], (u - 1) & 1u);     // the whole stage is released at once (its buffers are re-used
                uint32_t bytes = a.bytes_a;                        // inside a tile: staging over dZ, act over Z)
                if (a.da_mode == 2) bytes += 2u * (uint32_t)min((int64_t)4, a.M / 32 - (int64_t)m0 / 32) * (uint32_t)a.N * 4u;
                BF_STAMP(1);
                if (t == 0 && a.want_dx) bytes += (uint32_t)a.nS * w_chunk_bytes;
                mbar_expect_tx(&full_a[s], bytes);
                if (t == 0 && a.want_dx)
                    for (int j = 0; j < a.nS; ++j) bulk_g2s(s_w + (size_t)j * w_chunk_bytes, a.Wimg + (size_t)j * w_chunk_bytes, w_chunk_bytes, &full_a[s]);
                if (a.da_mode == 2) {
                    // the four groups of the tile: their dOut / arg rows are contiguous in global memory
                    const int64_t g0 = (int64_t)m0 / 32;
                    const int64_t G = a.M / 32;
                    const uint32_t ng = (uint32_t)min((int64_t)4, G - g0);
                    bulk_g2s(st + a.o_pool, a.dOut + g0 * a.N, ng * (uint32_t)a.N * 4u, &full_a[s]);
                    bulk_g2s(st + a.o_pool + 4u * (uint32_t)a.N * 4u, a.arg + g0 * a.N, ng * (uint32_t)a.N * 4u, &full_a[s]);
                } else
                for (int j = 0; j < a.nS; ++j) tma_load_2d(st + a.o_da + (size_t)j * kBfSlab, &a.tm_da, 64 * j, m0, &full_a[s]);
                if (a.da_mode != 3)
                    for (int j = 0; j < a.nS; ++j) tma_load_2d(st + a.o_z + (size_t)j * kBfSlab, &a.tm_z, 64 * j, m0, &full_a[s]);
                if (a.load_x) {
                    mbar_expect_tx(&full_b[s], a.bytes_b);
                    for (int j = 0; j < a.kS; ++j) tma_load_2d(st + a.o_x + (size_t)j * kBfSlab, &a.tm_x, 64 * j, m0, &full_b[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp >= kBfEpi / 32 && warp < (kBfEpi + kBfXf) / 32) {
        // ================= transform group =================
        const int ttid = tid - kBfEpi;
        for (uint32_t t = 0, s = 0, u = 0; t < n_my; ++t, s = (s + 1 == (uint32_t)S ? 0u : s + 1), u += (s == 0)) {
            const int64_t m0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * kBfBM;
            const int rows_valid = (int)min((int64_t)kBfBM, a.M - m0);
            uint8_t *st = smem + (size_t)s * a.stage_stride;
            uint8_t *const s_da = st + a.o_da, *const s_z = st + a.o_z, *const s_x = st + a.o_x, *const s_act = st + a.o_act;
            if (ttid == 0) BF_STAMP(2);
            mbar_wait(&full_a[s], u & 1u);
            if (ttid == 0) BF_STAMP(3);
            // ---- T1: dZ_l in place on the dA tile ----
            if (a.da_mode != 3 && !(a.dbg & 4)) {
                for (int j = 0; j < a.nS; ++j) {
                    const int nch = min(8, (a.N - 64 * j) >> 3);
                    const UnitMap um = unit_map(ttid, nch);
                    const int c0 = 64 * j + 8 * um.chunk;
                    if (um.chunk >= nch) continue;
                    float sc[8], sh[8], cb[8], cc[8];
                    if (a.coef_global == 1) {
                        // (same expressions as the shared tables of the prologue)
                        ldg_coef8(a.scale, c0, sc);
                        if (a.da_mode == 0 || a.da_mode == 2) ldg_coef8(a.shift, c0, sh);
                        if (a.mean) {
                            float dg[8], is[8], db[8], mu[8];
                            ldg_coef8(a.dgamma, c0, dg); ldg_coef8(a.invstd, c0, is); ldg_coef8(a.dbeta, c0, db); ldg_coef8(a.mean, c0, mu);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                cb[e] = -sc[e] * dg[e] * is[e] * a.inv_m;
                                cc[e] = -sc[e] * db[e] * a.inv_m - cb[e] * mu[e];
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) cb[e] = cc[e] = 0.0f;
                        }
                    } else {
                        ld_coef8(c_sc, c0, sc); ld_coef8(c_b, c0, cb); ld_coef8(c_c, c0, cc);
                        if (a.da_mode == 0 || a.da_mode == 2) ld_coef8(c_sh, c0, sh);
                    }
                    // (all of the thread's units are loaded before any arithmetic: 2 x <= 4 LDS.128 in flight instead of a
                    //  load -> compute -> store chain per unit; the loop bodies are fully unrolled and predicated on um.iters)
                    uint4 g4[4], z4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i < um.iters) {
                            const int r = um.row0 + um.rstep * i;
                            const uint32_t off = (uint32_t)j * kBfSlab + sw128_offset(r, um.chunk);
                            z4[i] = *reinterpret_cast<const uint4 *>(s_z + off);
                            if (a.da_mode == 2) {
                                // the pooled gradient reaches sample k of group (r >> 5) only where k won the max (:200)
                                const float *dp = reinterpret_cast<const float *>(st + a.o_pool) + (r >> 5) * a.N + c0;
                                const int *ap = reinterpret_cast<const int *>(st + a.o_pool + 4u * (uint32_t)a.N * 4u) + (r >> 5) * a.N + c0;
                                const float4 d0 = *reinterpret_cast<const float4 *>(dp), d1 = *reinterpret_cast<const float4 *>(dp + 4);
                                const int4 a0 = *reinterpret_cast<const int4 *>(ap), a1 = *reinterpret_cast<const int4 *>(ap + 4);
                                const int k = r & 31;
                                g4[i].x = pack_bf16x2(a0.x == k ? d0.x : 0.f, a0.y == k ? d0.y : 0.f);
                                g4[i].y = pack_bf16x2(a0.z == k ? d0.z : 0.f, a0.w == k ? d0.w : 0.f);
                                g4[i].z = pack_bf16x2(a1.x == k ? d1.x : 0.f, a1.y == k ? d1.y : 0.f);
                                g4[i].w = pack_bf16x2(a1.z == k ? d1.z : 0.f, a1.w == k ? d1.w : 0.f);
                            } else {
                                g4[i] = *reinterpret_cast<const uint4 *>(s_da + off);
                            }
                        }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i < um.iters) {
                            const int r = um.row0 + um.rstep * i;
                            uint32_t *gw = reinterpret_cast<uint32_t *>(&g4[i]);
                            const uint32_t *zw = reinterpret_cast<const uint32_t *>(&z4[i]);
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2) {
                                float2 g = unpack_bf16x2(gw[e2]);
                                const float2 z = unpack_bf16x2(zw[e2]);
                                if (a.da_mode == 0 || a.da_mode == 2) {
                                    if (!(fmaf(z.x, sc[2 * e2], sh[2 * e2]) > 0.0f)) g.x = 0.0f;
                                    if (!(fmaf(z.y, sc[2 * e2 + 1], sh[2 * e2 + 1]) > 0.0f)) g.y = 0.0f;
                                }
                                const float dx = fmaf(sc[2 * e2], g.x, fmaf(cb[2 * e2], z.x, cc[2 * e2]));
                                const float dy = fmaf(sc[2 * e2 + 1], g.y, fmaf(cb[2 * e2 + 1], z.y, cc[2 * e2 + 1]));
                                gw[e2] = pack_bf16x2(dx, dy);
                            }
                            if (r >= rows_valid) g4[i] = make_uint4(0u, 0u, 0u, 0u);      // rows past M (zero-filled loads) must stay zero
                            *reinterpret_cast<uint4 *>(s_da + (uint32_t)j * kBfSlab + sw128_offset(r, um.chunk)) = g4[i];
                        }
                }
            }
            if (ttid == 0) BF_STAMP(4);
            if (a.load_x) mbar_wait(&full_b[s], u & 1u);
            if (a.alias_act) named_bar_sync(1, kBfXf);       // every thread is done reading Z before act overwrites it
            // ---- T2: X tile -> act (weight-gradient operand).  The statistics product reads the X tile AS LOADED: S2 = sum dA'.z, and
            //      the finalize forms dgamma = invstd . (S2 - mean . S1) -- sum dA'.zhat without a zhat tile (it used to be computed and
            //      written back in place here: 8 FMAs, 4 packs and a 16-byte store per unit of the pipeline's longest stage; the
            //      operand is now the stored bf16 z itself instead of a second rounding of it) ----
            if (prev && !(a.dbg & 4)) {
                for (int j = 0; j < a.kS; ++j) {
                    const int nch = min(8, (a.K - 64 * j) >> 3);
                    const UnitMap um = unit_map(ttid, nch);
                    const int c0 = 64 * j + 8 * um.chunk;
                    if (um.chunk >= nch) continue;
                    float sc[8], sh[8];
                    if (a.coef_global) {
                        ldg_coef8(a.p_scale, c0, sc); ldg_coef8(a.p_shift, c0, sh);
                    } else {
                        ld_coef8(p_sc, c0, sc); ld_coef8(p_sh, c0, sh);
                    }
                    uint4 x4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i < um.iters)
                            x4[i] = *reinterpret_cast<const uint4 *>(s_x + (uint32_t)j * kBfSlab + sw128_offset(um.row0 + um.rstep * i, um.chunk));
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i < um.iters) {
                            const uint32_t off = (uint32_t)j * kBfSlab + sw128_offset(um.row0 + um.rstep * i, um.chunk);
                            uint4 a4;
                            const uint32_t *zw = reinterpret_cast<const uint32_t *>(&x4[i]);
                            uint32_t *aw = reinterpret_cast<uint32_t *>(&a4);
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2) {
                                const float2 z = unpack_bf16x2(zw[e2]);
                                aw[e2] = pack_bf16x2(fmaxf(fmaf(z.x, sc[2 * e2], sh[2 * e2]), 0.0f), fmaxf(fmaf(z.y, sc[2 * e2 + 1], sh[2 * e2 + 1]), 0.0f));
                            }
                            *reinterpret_cast<uint4 *>(s_act + off) = a4;
                        }
                }
                if (a.ones_col >= 0) {
                    // the statistics' ones: 16 columns of the zhat tile past its data (the TMA load zero-filled them): 256 units
                    const int r = ttid >> 1, ch = ((a.ones_col & 63) >> 3) + (ttid & 1);
                    *reinterpret_cast<uint4 *>(s_x + (size_t)(a.ones_col >> 6) * kBfSlab + sw128_offset(r, ch)) =
                        make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
                }
            }
            if (ttid == 0) BF_STAMP(5);
            fence_proxy_async();
            named_bar_sync(1, kBfXf);
            if (ttid == 0) BF_STAMP(6);
            // the tile's operands are in place: the MMA warp issues its products, the epilogue group may read act[s]
            if (warp == kBfEpi / 32 && elect_one_sync()) bf_mbar_arrive(&act_rdy[s]);
        }
    } else if (warp == (kBfEpi + kBfXf) / 32 + 1) {
        // ================= MMA warp: one elected lane issues every product =================
        // (in a warp of its own the issue -- ~400 cycles for the data/weight gradient, ~800 for the statistics, waits on the
        //  accumulator buffers included -- is off the transform and epilogue groups' critical paths)
        const uint32_t idesc_dx = make_idesc_bf16(kBfBM, a.K_pad, 0, 0);      // dA = dZ . W      (A, B K-major)
        const uint32_t idesc_dw = make_idesc_bf16(kBfBM, a.K_pad, 1, 1);      // dW = dZ^T . act  (A, B MN-major)
        const uint32_t idesc_s2 = make_idesc_bf16(kBfBM, a.s2_cols, 1, 1);    // S2 = dA'^T . [zhat | ones]
        const uint32_t idesc_s1 = make_idesc_bf16(kBfBM, 16, 1, 0);           // S1 = dA'^T . ones (B K-major, own tile)
        const uint32_t idesc_c = make_idesc_bf16(kBfBM, 64 + a.s2_cols, 1, 1);
        // Two kinds of work per tile, taken in whichever order their inputs arrive (a blocking tile-by-tile loop made the data
        // gradient of tile t + 1 wait for the epilogue of tile t): "front" = data + weight gradient once the transform group
        // is done, "back" = the statistics (+ combined weight gradient) once the epilogue staged dA'.
        auto front = [&](const uint32_t t, const uint32_t s, const uint32_t b) {
            uint8_t *st = smem + (size_t)s * a.stage_stride;
            uint8_t *const s_da = st + a.o_da, *const s_x = st + a.o_x, *const s_act = st + a.o_act;
            fence_after_sync();
            if (elect_one_sync()) {
                BF_STAMP(7);
                // (descriptors: the start address sits in the low 14 bits in 16-byte units -- stepping through a tile is an
                //  integer add on a descriptor built once; building each one from scratch cost ~170 cycles per MMA of dependent
                //  64-bit arithmetic in this single thread)
                if (a.want_dx) {
                    const uint32_t d_acc = tmem + b * (uint32_t)a.K_pad;
                    for (int j = 0; j < a.nS; ++j) {
                        const int n_left = a.N - 64 * j;
                        const int nk = n_left >= 64 ? 4 : (n_left + 15) / 16;
                        const uint64_t ad = make_desc(smem_addr(s_da + (size_t)j * kBfSlab), 0, 1024);
                        const uint64_t bd = make_desc(smem_addr(s_w + (size_t)j * w_chunk_bytes), 0, 1024);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            if (kk < nk) umma_bf16(d_acc, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc_dx, (uint32_t)((j | kk) != 0));
                    }
                }
                if (a.want_dw && !a.combined && !(a.dbg & 1)) {
                    const uint64_t ad = make_desc(smem_addr(s_da), lbo_n, 1024), bd = make_desc(smem_addr(prev ? s_act : s_x), lbo_k, 1024);
                    const uint32_t d_w = tmem + (uint32_t)a.off_dw;
                    umma_bf16(d_w, ad, bd, idesc_dw, (uint32_t)(t != 0));
#pragma unroll
                    for (int r = 1; r < kBfBM / 16; ++r) umma_bf16(d_w, ad + (uint64_t)(128 * r), bd + (uint64_t)(128 * r), idesc_dw, 1u);
                }
                if (a.want_dx) umma_commit(&acc_full[b]);
                else umma_commit(&empty[s]);                      // no epilogue: the stage may be refilled
                BF_STAMP(8);
            }
            __syncwarp();
        };
        auto back = [&](const uint32_t t, const uint32_t s) {
            uint8_t *st = smem + (size_t)s * a.stage_stride;
            uint8_t *const s_da = st + a.o_da, *const s_x = st + a.o_x, *const s_act = st + a.o_act, *const s_stage = st + a.o_stage;
            fence_after_sync();
            if (elect_one_sync()) {
                if (a.combined) {
                    // A = [dZ slab | staged dA' slab] (M = 128 through LBO), B = [act slab | zhat (+ ones) slab] (N through LBO):
                    // D[0:64, 0:64] = dZ^T.act = dW, D[64:128, 64:] = dA'^T.[zhat | ones] = the statistics; 8 MMAs instead of 16
                    const uint64_t ad = make_desc(smem_addr(s_da), a.o_stage - a.o_da, 1024);
                    const uint64_t bd = make_desc(smem_addr(s_act), a.o_x - a.o_act, 1024);
                    const uint32_t d_c = tmem + (uint32_t)a.off_dw;
#pragma unroll
                    for (int r = 0; r < kBfBM / 16; ++r)
                        umma_bf16(d_c, ad + (uint64_t)(128 * r), bd + (uint64_t)(128 * r), idesc_c, r ? 1u : (uint32_t)(t != 0));
                    if (a.ones_col < 0) {
                        const uint64_t as = make_desc(smem_addr(s_stage), 0, 1024), b1 = make_desc_ones(smem_addr(s_ones));
                        const uint32_t d1 = tmem + (uint32_t)a.off_s1;
#pragma unroll
                        for (int r = 0; r < kBfBM / 16; ++r)
                            umma_bf16(d1, as + (uint64_t)(128 * r), b1, idesc_s1, r ? 1u : (uint32_t)(t != 0));
                    }
                } else if (a.want_stats && !(a.dbg & 2)) {
                    const uint64_t ad = make_desc(smem_addr(s_stage), lbo_k, 1024), b2 = make_desc(smem_addr(s_x), lbo_k, 1024);
                    const uint64_t b1 = make_desc_ones(smem_addr(s_ones));
                    const uint32_t d2 = tmem + (uint32_t)a.off_s2, d1 = tmem + (uint32_t)a.off_s1;
                    const bool own_ones = a.ones_col < 0;
#pragma unroll
                    for (int r = 0; r < kBfBM / 16; ++r) {
                        const uint32_t acc = r ? 1u : (uint32_t)(t != 0);
                        umma_bf16(d2, ad + (uint64_t)(128 * r), b2 + (uint64_t)(128 * r), idesc_s2, acc);
                        if (own_ones) umma_bf16(d1, ad + (uint64_t)(128 * r), b1, idesc_s1, acc);
                    }
                }
                BF_STAMP(13);
                umma_commit(&empty[s]);       // every MMA that reads the stage is done (and the store has read its staging) -> refill
            }
            __syncwarp();
        };
        const uint32_t n_back = a.want_dx ? n_my : 0u;
        uint32_t sd = 0, ud = 0, ss = 0, us = 0;      // stage / use count of the next front (td) and back (ts) tile
        for (uint32_t td = 0, ts = 0; td < n_my || ts < n_back;) {
            if (td < n_my) {
                const uint32_t b = AB == 2 ? (td & 1u) : 0u, ub = AB == 2 ? (td >> 1) : td;
                bool ok = mbar_test(&act_rdy[sd], ud & 1u);
                if (ok && a.want_dx && td >= (uint32_t)AB) ok = mbar_test(&acc_empty[b], (ub - 1) & 1u);
                if (__all_sync(0xffffffffu, ok)) {
                    front(td, sd, b);
                    ++td;
                    if (++sd == (uint32_t)S) { sd = 0; ++ud; }
                }
            }
            if (ts < td && ts < n_back) {
                if (__all_sync(0xffffffffu, mbar_test(&e_rdy[ss], us & 1u))) {
                    back(ts, ss);
                    ++ts;
                    if (++ss == (uint32_t)S) { ss = 0; ++us; }
                }
            }
        }
        if (elect_one_sync()) umma_commit(&bar_done);     // every product of this CTA has landed in tensor memory
        __syncwarp();
    } else {
        // ================= epilogue group =================
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        if (a.want_dx) {
            for (uint32_t t = 0, s = 0, u = 0; t < n_my; ++t, s = (s + 1 == (uint32_t)S ? 0u : s + 1), u += (s == 0)) {
                const uint32_t b = AB == 2 ? (t & 1u) : 0u, ub = AB == 2 ? (t >> 1) : t;      // t % AB, t / AB (AB is 1 or 2)
                const int64_t m0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * kBfBM;
                uint8_t *st = smem + (size_t)s * a.stage_stride;
                uint8_t *const s_act = st + a.o_act, *const s_stage = st + a.o_stage;
                if (tid == 0) BF_STAMP(9);
                mbar_wait(&acc_full[b], ub & 1u);
                mbar_wait(&act_rdy[s], u & 1u);
                if (tid == 0) BF_STAMP(10);
                fence_after_sync();
                const uint32_t taddr = tmem + b * (uint32_t)a.K_pad + lane_addr;
                for (int c0 = 0; c0 < a.k_store; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);
                    const int ch = (c0 & 63) >> 3;
                    const uint32_t o_lo = (uint32_t)(c0 >> 6) * kBfSlab + sw128_offset(tid, ch);
                    const uint32_t o_hi = (uint32_t)(c0 >> 6) * kBfSlab + sw128_offset(tid, ch + 1);
                    const bool two = c0 + 8 < a.k_store;
                    if (prev) {
                        const uint4 m_lo = *reinterpret_cast<const uint4 *>(s_act + o_lo);
                        uint4 m_hi = make_uint4(0u, 0u, 0u, 0u);
                        if (two) m_hi = *reinterpret_cast<const uint4 *>(s_act + o_hi);
                        const uint32_t mw[8] = {m_lo.x, m_lo.y, m_lo.z, m_lo.w, m_hi.x, m_hi.y, m_hi.z, m_hi.w};
#pragma unroll
                        for (int e2 = 0; e2 < 8; ++e2) {
                            // act is relu(.) >= 0 in bf16: positive <=> its 15 magnitude bits are non-zero
                            if (!(mw[e2] & 0x00007fffu)) v[2 * e2] = 0.0f;
                            if (!(mw[e2] & 0x7fff0000u)) v[2 * e2 + 1] = 0.0f;
                        }
                    }
                    uint4 lo, hi;
                    lo.x = pack_bf16x2(v[0], v[1]);   lo.y = pack_bf16x2(v[2], v[3]);
                    lo.z = pack_bf16x2(v[4], v[5]);   lo.w = pack_bf16x2(v[6], v[7]);
                    hi.x = pack_bf16x2(v[8], v[9]);   hi.y = pack_bf16x2(v[10], v[11]);
                    hi.z = pack_bf16x2(v[12], v[13]); hi.w = pack_bf16x2(v[14], v[15]);
                    *reinterpret_cast<uint4 *>(s_stage + o_lo) = lo;
                    if (two) *reinterpret_cast<uint4 *>(s_stage + o_hi) = hi;
                }
                fence_before_sync();      // this thread's tensor-memory reads are complete
                fence_proxy_async();      // staging writes -> visible to the TMA store / the statistics MMAs
                if (tid == 0) BF_STAMP(11);
                // staging in a buffer of its own (not over dZ) and >= 2 stages: the previous tile's store had a whole tile to read
                // its buffer -- make sure of it here, before anyone passes into the next epilogue, instead of stalling behind
                // every store
                if (a.defer_store_wait && warp == 0 && elect_one_sync()) tma_store_wait_read();
                named_bar_sync(2, kBfEpi);
                if (warp == 0) {
                  if (elect_one_sync()) {
                    BF_STAMP(12);
                    bf_mbar_arrive(&acc_empty[b]);       // the data-gradient MMAs of tile t + AB may overwrite this buffer
                    for (int j = 0; j < a.kS; ++j)
                        if (64 * j < a.k_store) tma_store_2d(&a.tm_dx, 64 * j, (int)m0, s_stage + (size_t)j * kBfSlab);
                    if (!a.defer_store_wait) tma_store_wait_read();   // (staging sits in this stage's dA buffer: the refill must wait)
                    BF_STAMP(14);
                    bf_mbar_arrive(&e_rdy[s]);           // MMA warp: the statistics products, then the stage is released
                  }
                  __syncwarp();
                }
            }
            if (warp == 0 && elect_one_sync()) tma_store_wait_all();
        }
        // ---- end of the CTA: drain the accumulators that lived in tensor memory across all its tiles ----
        mbar_wait(&bar_done, 0);
        fence_after_sync();
        const uint32_t taddr = tmem + lane_addr;
        if (a.want_dw && warp * 32 < a.N && !(a.dbg & 8)) {      // (tensor-memory loads are warp-collective: whole warps take part, stores are per row)
            float *acc = a.dW ? a.dW + (size_t)tid * a.K : nullptr;
            float *part = a.scratch ? a.scratch + ((size_t)blockIdx.x * a.N + tid) * a.K_ld4 : nullptr;
            const bool mine = tid < a.N;
            for (int c0 = 0; c0 < a.K_pad; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)a.off_dw + c0, v);
                if (!mine) continue;
                if (part) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        if (c0 + i < a.K_ld4) *reinterpret_cast<float4 *>(part + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else if (a.dw_vec4) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        if (c0 + i < a.K) bf_red_add_v4_f32(acc + c0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < a.K) bf_red_add_f32(acc + c0 + i, v[i]);
                }
            }
        }
        if (a.want_stats && a.combined) {
            // statistics block of the combined product: row 64 + k, columns 64 + k (S2's diagonal) and 64 + ones_col (S1)
            if (warp >= 2 && (warp - 2) * 32 < a.K) {
                const int k = tid - 64;
                float v0[16], v1[16], w[16];
                const uint32_t c_stat = (uint32_t)a.off_dw + 64u;
                tmem_ld16(taddr + c_stat + (uint32_t)((warp - 2) * 32), v0);
                tmem_ld16(taddr + c_stat + (uint32_t)((warp - 2) * 32 + 16), v1);
                tmem_ld16(taddr + (a.ones_col >= 0 ? c_stat + (uint32_t)a.ones_col : (uint32_t)a.off_s1), w);
                float s2 = 0.0f;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if ((lane & 15) == i) s2 = lane < 16 ? v0[i] : v1[i];
                if (k < a.K) {
                    double *acc = a.stat_accum + (size_t)(blockIdx.x % kStatReplicas) * 2 * a.K;
                    atomicAdd(acc + k, (double)w[0]);
                    atomicAdd(acc + a.K + k, (double)s2);
                }
            }
        } else if (a.want_stats) {
            // S2's diagonal (row k, column k) and S1 (row k, the ones column): lanes 0-15 need chunk 2 * warp, lanes 16-31 the next
            if (warp * 32 < a.K) {
                float v0[16], v1[16], w[16];
                tmem_ld16(taddr + (uint32_t)a.off_s2 + (uint32_t)(warp * 32), v0);
                tmem_ld16(taddr + (uint32_t)a.off_s2 + (uint32_t)(warp * 32 + 16), v1);
                tmem_ld16(taddr + (uint32_t)(a.ones_col >= 0 ? a.off_s2 + a.ones_col : a.off_s1), w);
                float s2 = 0.0f;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if ((lane & 15) == i) s2 = lane < 16 ? v0[i] : v1[i];
                if (tid < a.K) {
                    double *acc = a.stat_accum + (size_t)(blockIdx.x % kStatReplicas) * 2 * a.K;
                    atomicAdd(acc + tid, (double)w[0]);
                    atomicAdd(acc + a.K + tid, (double)s2);
                }
            }
        }
        if (a.want_stats) {
            // "last CTA finalizes" (as bn.cu: bn_bwd_last_block_finalize, among the epilogue threads)
            __threadfence();
            named_bar_sync(2, kBfEpi);
            if (tid == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1 ? 1 : 0;
            named_bar_sync(2, kBfEpi);
            if (s_last) {
                __threadfence();
                for (int c = tid; c < a.K; c += kBfEpi) {
                    double t1 = 0.0, t2 = 0.0;
#pragma unroll
                    for (int r = 0; r < kStatReplicas; ++r) {
                        double *acc = a.stat_accum + (size_t)r * 2 * a.K;
                        t1 += __ldcg(acc + c);
                        t2 += __ldcg(acc + a.K + c);
                        acc[c] = 0.0;
                        acc[a.K + c] = 0.0;
                    }
                    a.dbeta_prev[c] = (float)t1;
                    a.dgamma_prev[c] = (float)((t2 - (double)a.p_mean[c] * t1) * (double)a.p_invstd[c]);      // t2 = sum dA'.z (see T2)
                }
                if (tid == 0) *a.ticket = 0u;
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, (uint32_t)a.tmem_cols);
}

static int bf_round_up(int v, int m) { return (v + m - 1) / m * m; }

struct BwdFusedPlan {
    bool ok;
    int nS, kS, K_pad, k_store, tmem_cols, acc_bufs, off_dw, off_s2, off_s1, ones_col, s2_cols, coef_ld, stages, load_x, alias_act, combined;
    int coef_global;
    uint32_t o_w, o_da, o_z, o_x, o_act, o_stage, o_ones, o_coef, o_pool, stage_stride, bytes_a, bytes_b;
    size_t dyn_smem;
};

static BwdFusedPlan bwd_fused_plan(int K, int N, int ldx, int lddx, int da_mode, bool prev, bool want_dx, bool want_dw,
                                   bool allow_coef_global = true) {
    BwdFusedPlan p;
    memset(&p, 0, sizeof(p));
    if (N < 8 || N > kBfMaxC || N % 8 != 0 || K < 1 || ldx > kBfMaxC || ldx % 8 != 0 || bf_round_up(K, 16) > kBfMaxC) return p;
    if (!want_dx && !want_dw) return p;
    if (prev && (!want_dx || K % 8 != 0)) return p;       // statistics of layer l-1 come out of the data-gradient epilogue
    p.nS = (N + 63) / 64;
    p.kS = (ldx + 63) / 64;
    p.K_pad = bf_round_up(K, 16);
    p.k_store = bf_round_up(K, 8);
    if (want_dx && p.k_store > lddx) p.k_store = lddx;
    p.load_x = (want_dw || prev) ? 1 : 0;
    // the statistics' ones: inside the zhat tile when its last slab has 16 spare columns, else a 4 KB tile of their own
    p.ones_col = -1;
    p.s2_cols = p.kS * 64;
    if (prev && p.K_pad + 16 <= p.kS * 64) { p.ones_col = p.K_pad; p.s2_cols = p.K_pad + 16; }
    else if (prev) p.s2_cols = p.K_pad;
    // one stage: [dA | Z | X | act | staging]; inside a tile dZ (the dA buffer) is dead once the data / weight-gradient MMAs
    // are done -- the staged dA_{l-1} tile takes its place -- and Z is dead once dZ is formed -- act takes its place
    uint32_t o = 0;
    p.o_da = o; o += (uint32_t)p.nS * kBfSlab;
    p.o_z = o;
    if (da_mode != 3) o += (uint32_t)p.nS * kBfSlab;
    p.bytes_a = da_mode == 2 ? (uint32_t)p.nS * kBfSlab : o;      // (pooled: only Z is a tile load; dOut / arg rows are added per tile)
    p.o_x = o;
    if (p.load_x) o += (uint32_t)p.kS * kBfSlab;
    p.bytes_b = p.load_x ? (uint32_t)p.kS * kBfSlab : 0u;
    // narrow layers: ONE product gives the weight gradient and the statistics (the tensor pipe takes ~170 cycles per
    // tcgen05.mma of these tiny shapes whatever it holds, so the MMA COUNT is what matters); it needs dZ alive next to the staged
    // tile and act next to zhat: [dZ][Z -> act][X -> zhat][staging]
    p.combined = (prev && want_dw && want_dx && p.nS == 1 && p.kS == 1) ? 1 : 0;
    p.alias_act = (prev && da_mode != 3 && p.kS <= p.nS) ? 1 : 0;
    if (p.combined && !p.alias_act) {            // (dZ given: no Z tile -- act gets its own slab BEFORE X so that LBO(act -> zhat) > 0)
        p.o_act = p.o_x;
        p.o_x = o;
        o += (uint32_t)p.kS * kBfSlab;
    } else {
        p.o_act = p.alias_act ? p.o_z : o;
        if (prev && !p.alias_act) o += (uint32_t)p.kS * kBfSlab;
    }
    if (want_dx && p.kS <= p.nS && !p.combined) p.o_stage = p.o_da;
    else { p.o_stage = o; if (want_dx) o += (uint32_t)p.kS * kBfSlab; }
    p.o_pool = o;
    if (da_mode == 2) o += (uint32_t)bf_round_up(8 * N * 4, 1024);
    p.stage_stride = o;
    const uint32_t w_bytes = want_dx ? (uint32_t)bf_round_up(p.nS * p.K_pad * 128, 1024) : 0u;
    const uint32_t ones_bytes = (prev && p.ones_col < 0) ? 128u : 0u;       // one core matrix (see the kernel's prologue)
    p.coef_ld = bf_round_up(N > p.kS * 64 ? N : p.kS * 64, 64);
    uint32_t coef_bytes = (uint32_t)(6 * p.coef_ld * sizeof(float));      // sc, sh, cb, cc of layer l; scale, shift of layer l-1
    // dynamic + static shared memory <= 227 KB: the kernel's static part is 208 bytes (barriers), 256 reserved; the 1 KB
    // alignment slack is added below
    const uint32_t budget = 227 * 1024 - 256;
    auto stages_with = [&](uint32_t cb) {
        int st = kBfMaxStages;
        while (st > 1 && (size_t)st * p.stage_stride + w_bytes + ones_bytes + cb + 1024 > budget) --st;
        return st;
    };
    p.stages = stages_with(coef_bytes);
    // The 128-wide layers: two 96 KB stages + the 32 KB weight image leave < 2 KB.  With 4 KB of coefficient tables (and a
    // 4 KB ones tile) they ran ONE stage -- load -> transform -> MMA -> epilogue -> store strictly in turn, 15.5 k cycles per
    // tile.  Then: layer l-1's coefficients (T2) straight from its arrays (__ldg, L1-resident), and layer l's three derived
    // tables (1.5 KB) in shared memory when the shift is not needed (da_mode 1), else from the arrays as well.
    static const bool coef_global_on = !(getenv("PN2_BWD_COEF_SMEM") && atoi(getenv("PN2_BWD_COEF_SMEM")) == 1);
    if (allow_coef_global && coef_global_on && p.stages < 2) {
        const uint32_t three = (uint32_t)(3 * p.coef_ld * sizeof(float));
        if ((da_mode == 1 || da_mode == 3) && stages_with(three) >= 2) { p.coef_global = 2; coef_bytes = three; }
        else if (stages_with(0) >= 2) { p.coef_global = 1; coef_bytes = 0; }
        p.stages = stages_with(coef_bytes);
    }
    if ((size_t)p.stage_stride + w_bytes + ones_bytes + coef_bytes + 1024 > budget) return p;
    o = (uint32_t)p.stages * p.stage_stride;
    p.o_w = o; o += w_bytes;
    p.o_ones = o; o += ones_bytes;
    p.o_coef = o; o += coef_bytes;
    p.dyn_smem = (size_t)o + 1024;
    // tensor memory: [dA_{l-1}: acc_bufs x K_pad][dW: K_pad][S2: s2_cols][S1: 16 when the ones have their own tile]
    const int rest = p.combined ? 64 + p.s2_cols + (p.ones_col < 0 ? 16 : 0)
                                : (want_dw ? p.K_pad : 0) + (prev ? p.s2_cols + (p.ones_col < 0 ? 16 : 0) : 0);
    p.acc_bufs = (want_dx && 2 * p.K_pad + rest <= 512) ? 2 : 1;
    int cols = want_dx ? p.acc_bufs * p.K_pad : 0;
    p.off_dw = cols;
    if (p.combined) cols += 64 + p.s2_cols;
    else if (want_dw) cols += p.K_pad;
    p.off_s2 = cols;
    if (prev && !p.combined) cols += p.s2_cols;
    p.off_s1 = cols;
    if (prev && p.ones_col < 0) cols += 16;
    int alloc = 32;
    while (alloc < cols) alloc <<= 1;
    if (alloc > 512) return p;
    p.tmem_cols = alloc;
    p.ok = true;
    return p;
}

static bool bwd_fused_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("PN2_FUSED_BWD");
        on = (e && atoi(e) == 0) ? 0 : 1;
        const char *t = getenv("PN2_DISABLE_TC");
        if (t && atoi(t) != 0) on = 0;
    }
    return on != 0;
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_mlp_bwd_layer_supported(int64_t M, int K, int N, int ldx, int lddx, int da_mode, int has_prev, int want_dx,
                                           int want_dw) {
    if (!bwd_fused_enabled() || M < 1) return 0;
    return bwd_fused_plan(K, N, ldx, lddx, da_mode, has_prev != 0, want_dx != 0, want_dw != 0).ok ? 1 : 0;
}

extern "C" size_t pn2_mlp_bwd_layer_scratch_bytes(int64_t M, int K, int N) {
    // fp32 partial blocks of the weight gradient when K % 4 != 0 (rows of dW are not 16-byte aligned: no vector reductions)
    if (K % 4 == 0) return 0;
    const int64_t m_tiles = (M + kBfBM - 1) / kBfBM;
    const int64_t grid = m_tiles < kNumSMs ? m_tiles : kNumSMs;
    return (size_t)grid * (size_t)N * (size_t)bf_round_up(K, 4) * sizeof(float) + 16;
}

extern "C" int pn2_mlp_bwd_layer(const pn2_bwd_layer *L, void *stream) {
    PN2_REQUIRE(L, "mlp_bwd_layer: null argument block");
    PN2_REQUIRE(L->M >= 1 && L->K >= 1 && L->N >= 1, "mlp_bwd_layer: bad sizes");
    PN2_REQUIRE(L->da_mode >= 0 && L->da_mode <= 3, "mlp_bwd_layer: da_mode must be 0, 1, 2 or 3");
    PN2_REQUIRE(L->da_mode == 2 ? (L->dOut && L->arg && L->nsample == 32 && L->M % 32 == 0) : L->dA != nullptr,
                "mlp_bwd_layer: dA (or, pooled: dOut, arg, nsample == 32, M %% 32 == 0) is required");
    PN2_REQUIRE(L->da_mode == 3 || (L->Z && L->scale && L->shift), "mlp_bwd_layer: layer l's Z / scale / shift are required");
    PN2_REQUIRE(L->da_mode == 3 || !L->mean || (L->invstd && L->dgamma && L->dbeta), "mlp_bwd_layer: train-mode BatchNorm needs invstd, dgamma, dbeta");
    const bool prev = L->prev_scale != nullptr, want_dx = L->dX != nullptr, want_dw = L->dW != nullptr;
    PN2_REQUIRE(!prev || (L->prev_shift && L->prev_mean && L->prev_invstd && L->stat_accum && L->ticket && L->dgamma_prev && L->dbeta_prev),
                "mlp_bwd_layer: the statistics of layer l-1 need shift, mean, invstd, accumulator, ticket, dgamma_prev, dbeta_prev");
    PN2_REQUIRE(!(want_dw || prev) || L->X, "mlp_bwd_layer: X is required");
    PN2_REQUIRE(!want_dx || L->wpack_t, "mlp_bwd_layer: the data gradient needs the transposed weight image");
    PN2_REQUIRE((L->da_mode == 2 || L->ldda >= L->N) && (L->da_mode == 3 || L->ldz >= L->N) && (!L->X || L->ldx >= L->K) && (!want_dx || L->lddx >= L->K),
                "mlp_bwd_layer: leading dimensions");
    PN2_REQUIRE((L->da_mode == 2 || L->ldda % 8 == 0) && (L->da_mode == 3 || L->ldz % 8 == 0) && (!L->X || L->ldx % 8 == 0) && (!want_dx || L->lddx % 8 == 0),
                "mlp_bwd_layer: bf16 rows need a 16-byte row pitch");
    // (coefficients straight from the arrays need 16-byte aligned arrays: float4 loads)
    auto al16 = [](const void *q) { return ((uintptr_t)q & 15) == 0; };
    const bool arrays_aligned = al16(L->scale) && al16(L->shift) && al16(L->mean) && al16(L->invstd) && al16(L->dgamma) && al16(L->dbeta) &&
                                al16(L->prev_scale) && al16(L->prev_shift) && al16(L->prev_mean) && al16(L->prev_invstd);
    const BwdFusedPlan p = bwd_fused_plan(L->K, L->N, L->X ? L->ldx : bf_round_up(L->K, 8), want_dx ? L->lddx : 0, L->da_mode, prev, want_dx, want_dw,
                                          arrays_aligned);
    if (!p.ok || !bwd_fused_enabled()) {
        set_error("mlp_bwd_layer: layer K=%d N=%d is not supported by the fused kernel (see pn2_mlp_bwd_layer_supported)", L->K, L->N);
        return PN2_ERR_UNSUPPORTED;
    }
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256);
        if (e != cudaSuccess) {
            set_error("mlp_bwd_layer: shared-memory opt-in failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return PN2_ERR_CUDA;
        }
        attr_done = true;
    }
    BwdFusedArgs a;
    memset(&a, 0, sizeof(a));
    bool ok = true;
    if (L->da_mode != 2) ok = make_rows_tensor_map(&a.tm_da, L->dA, L->M, bf_round_up(L->N, 8) <= L->ldda ? bf_round_up(L->N, 8) : L->ldda, L->ldda, kBfBM);
    if (L->da_mode != 3) ok = ok && make_rows_tensor_map(&a.tm_z, L->Z, L->M, bf_round_up(L->N, 8) <= L->ldz ? bf_round_up(L->N, 8) : L->ldz, L->ldz, kBfBM);
    else a.tm_z = a.tm_da;
    if (L->da_mode == 2) a.tm_da = a.tm_z;
    if (want_dw || prev) ok = ok && make_rows_tensor_map(&a.tm_x, L->X, L->M, L->ldx, L->ldx, kBfBM);
    else a.tm_x = a.tm_da;
    if (want_dx) ok = ok && make_rows_tensor_map(&a.tm_dx, L->dX, L->M, p.k_store, L->lddx, kBfBM);
    else a.tm_dx = a.tm_da;
    if (!ok) {
        set_error("mlp_bwd_layer: cuTensorMapEncodeTiled failed (M=%lld N=%d K=%d)", (long long)L->M, L->N, L->K);
        return PN2_ERR_CUDA;
    }
    a.scale = L->scale; a.shift = L->shift; a.mean = L->mean; a.invstd = L->invstd; a.dgamma = L->dgamma; a.dbeta = L->dbeta;
    a.p_scale = L->prev_scale; a.p_shift = L->prev_shift; a.p_mean = L->prev_mean; a.p_invstd = L->prev_invstd;
    a.Wimg = (const uint8_t *)L->wpack_t;
    a.M = L->M; a.K = L->K; a.N = L->N; a.K_pad = p.K_pad; a.k_store = p.k_store; a.nS = p.nS; a.kS = p.kS;
    a.da_mode = L->da_mode; a.want_dx = want_dx; a.want_dw = want_dw; a.want_stats = prev;
    const int64_t m_tiles = (L->M + kBfBM - 1) / kBfBM;
    int64_t grid = sm_budget();                              // one CTA per SM, pipelined over its tiles (pn2_set_sm_budget: fewer)
    if (grid > m_tiles) grid = m_tiles;
    a.K_ld4 = bf_round_up(L->K, 4);
    if (want_dw) {
        if (L->K % 4 == 0 && ((uintptr_t)L->dW & 15) == 0) { a.dW = L->dW; a.dw_vec4 = 1; }
        else if (L->scratch) a.scratch = (float *)(((uintptr_t)L->scratch + 15) & ~(uintptr_t)15);
        else a.dW = L->dW;                                   // scalar L2 reductions
    }
    a.stat_accum = L->stat_accum; a.ticket = L->ticket; a.dgamma_prev = L->dgamma_prev; a.dbeta_prev = L->dbeta_prev;
    a.inv_m = 1.0f / (float)L->M;
    a.tmem_cols = p.tmem_cols; a.off_dw = p.off_dw; a.off_s2 = p.off_s2; a.off_s1 = p.off_s1;
    a.ones_col = p.ones_col; a.s2_cols = p.s2_cols; a.o_coef = p.o_coef; a.coef_ld = p.coef_ld;
    a.o_w = p.o_w; a.o_da = p.o_da; a.o_z = p.o_z; a.o_x = p.o_x; a.o_act = p.o_act; a.o_stage = p.o_stage; a.o_ones = p.o_ones;
    a.stage_stride = p.stage_stride; a.bytes_a = p.bytes_a; a.bytes_b = p.bytes_b; a.stages = p.stages; a.acc_bufs = p.acc_bufs;
    a.load_x = p.load_x; a.alias_act = p.alias_act; a.combined = p.combined; a.coef_global = p.coef_global;
    a.o_pool = p.o_pool; a.dOut = L->dOut; a.arg = L->arg;
    a.defer_store_wait = (want_dx && p.o_stage != p.o_da && p.stages >= 2) ? 1 : 0;
    { const char *e = getenv("PN2_BWD_DBG"); a.dbg = e ? atoi(e) : 0; }
    if ((a.dbg & 32) && L->scratch && L->K % 4 == 0) a.dbg_buf = (long long *)L->scratch;
    bwd_fused_kernel<<<(unsigned)grid, kBfThreads, p.dyn_smem, (cudaStream_t)stream>>>(a);
    count_launch();
    int rc = check_launch("mlp_bwd_layer");
    if (rc != PN2_OK || !a.scratch) return rc;
    return tc_wgrad_reduce(a.scratch, (int)grid, L->N, L->K, a.K_ld4, L->dW, (cudaStream_t)stream);
}
