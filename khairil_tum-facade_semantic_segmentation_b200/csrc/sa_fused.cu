// sa_fused.cu -- K3d: a whole set-abstraction level in ONE kernel for inference (eval-mode BatchNorm):
//
//   ball-query gather -> [ relu(bn(conv1x1)) ] x L on the tensor cores -> max over nsample
//   (/root/reference/models/pointnet2_utils.py:127-132 gather + concat, :196-198 MLP, :200 max)
//
// Rows are (cloud, centroid, sample) with nsample = 32, so a 128-row UMMA tile is four groups and
// the 32 TMEM lanes a warp owns in the epilogue are exactly one group: the max over nsample is a
// reduction across the lanes of one warp.  Activations never leave the SM:
//   * gather: the 128 threads build the tile's input rows [feats[idx] | xyz[idx] - new_xyz | 0] straight
//     into shared memory as bf16 in the K-major 128-byte-swizzle UMMA layout (features FIRST so their
//     16-byte chunks stay aligned with the fp32 source rows; the first layer's weight image is packed
//     with the same column rotation);
//   * layer l: thread 0 issues tcgen05.mma over the K chunks, D in tensor memory (fp32); weights are
//     bf16 pre-swizzled chunk images streamed by a producer warp with cp.async.bulk (TMA) through an
//     mbarrier ring -- or loaded once and kept resident when all of a level's images fit;
//   * epilogue l < L-1: tcgen05.ld -> BatchNorm (eval scale/shift, conv bias folded) -> ReLU -> bf16 ->
//     back into the SAME shared-memory tile as the next layer's A operand;
//   * last epilogue: BatchNorm -> ReLU -> REDUX.MAX over the warp's 32 lanes per channel (values are >= 0,
//     so the unsigned order of the bit patterns is the float order) -> one coalesced fp32 store per 32
//     channels into out[B, S, C].
// HBM traffic per level: the index tensor, the gathered source rows (L2-resident: every point is read
// ~nsample/4 times), the weights once per CTA, and the pooled output -- no grouped tensor, no activations.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pn2 {

using namespace tc;

__global__ void pack_weight_kernel(const float *__restrict__ W, int64_t w_sn, int64_t w_sk, int N, int K, int n_pad,
                                   int KC, int k_rot, uint8_t *__restrict__ img);

constexpr int kSaMaxLayers = 4;
constexpr int kSaMaxChunks = 24;
constexpr int kSaTile = 128;
constexpr int kSaSlab = kSaTile * 128;      // one 64-column slab of the activation tile: 16 KB
constexpr int kSaThreads = 160;             // warps 0-3: gather, MMA issue (thread 0), epilogues; warp 4: weight producer

struct SaChunk {
    uint32_t off;        // byte offset of the chunk image inside Wimg
    uint32_t bytes;      // n_pad * 128
    uint32_t idesc;      // instruction descriptor (M = 128, N = n_pad)
    uint16_t tmem_col;   // accumulator column of this chunk's N block
    uint8_t layer, kc, nk, pad;
};

struct SaFusedArgs {
    const float *xyz;
    int64_t sB, sN, sC;
    const float *new_xyz;       // [B, S, 3] contiguous
    const float *feats;         // [B, N, D] rows contiguous (element stride 1), may be null (D = 0)
    int64_t fB, fN;
    const int64_t *idx;         // [B, S, 32]
    int N, S, D;
    int64_t G;                  // B * S groups
    int L;
    int width[kSaMaxLayers];    // output channels per layer
    int chunk_begin[kSaMaxLayers + 1];
    const float *scale[kSaMaxLayers], *shift[kSaMaxLayers], *bias[kSaMaxLayers];
    const uint8_t *Wimg;
    int n_chunks, a_slabs, stages, slot_bytes, resident, tmem_cols, ss_floats;
    SaChunk chunks[kSaMaxChunks];
    float *out;                 // [G, width[L-1]]
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory map (1024-byte aligned): [activation tile: a_slabs x 16 KB] [weight ring: stages x slot_bytes]
// [scale | shift of every layer: 2 x ss_floats fp32] [gather row offsets: 128 int64]
template <int kGU>
__global__ void __launch_bounds__(kSaThreads) sa_fused_eval_kernel(const __grid_constant__ SaFusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kSaMaxChunks], bar_empty[kSaMaxChunks], bar_acc;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int warp_u = uniform_warp_idx();
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    uint8_t *const act = smem;
    uint8_t *const ring = act + (size_t)a.a_slabs * kSaSlab;
    float *const s_scale = reinterpret_cast<float *>(ring + (size_t)a.stages * a.slot_bytes);
    float *const s_shift = s_scale + a.ss_floats;
    int64_t *const s_off = reinterpret_cast<int64_t *>(s_shift + a.ss_floats);     // feature-row offsets of the tile's 128 rows

    if (warp == 0) tmem_alloc(&tmem_base_s, (uint32_t)a.tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
        }
        mbar_init(&bar_acc, 1);
        mbar_init_fence();
    }
    {   // eval-mode BatchNorm of every layer, conv bias folded: y = z*scale + (shift + scale*bias)
        int base = 0;
        for (int l = 0; l < a.L; ++l) {
            for (int c = tid; c < a.width[l]; c += kSaThreads) {
                const float sc = a.scale[l][c];
                s_scale[base + c] = sc;
                s_shift[base + c] = a.shift[l][c] + (a.bias[l] ? sc * a.bias[l][c] : 0.0f);
            }
            base += (a.width[l] + 31) & ~31;
        }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const int64_t n_tiles = (a.G * 32 + kSaTile - 1) / kSaTile;

    if (warp == 4) {
        // ---- producer: stream (or load once) the weight chunk images ----
        if (lane == 0) {
            uint32_t n = 0;
            bool first = true;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first = false) {
                if (a.resident && !first) break;
                for (int c = 0; c < a.n_chunks; ++c, ++n) {
                    const uint32_t s = a.resident ? (uint32_t)c : n % (uint32_t)a.stages;
                    const uint32_t use = a.resident ? 0u : n / (uint32_t)a.stages;
                    if (use > 0) mbar_wait(&bar_empty[s], (use - 1) & 1u);
                    mbar_expect_tx(&bar_full[s], a.chunks[c].bytes);
                    bulk_g2s(ring + (size_t)s * a.slot_bytes, a.Wimg + a.chunks[c].off, a.chunks[c].bytes, &bar_full[s]);
                }
            }
        }
        __syncwarp();
    } else {
        const int K0 = a.D + 3;
        const int cpr = ((K0 + 15) >> 4) << 1;          // 16-byte chunks per gathered row, zero-padded to 16 columns
        const int cpf = a.D >> 3;                       // chunks that hold features only; the rest is the row's tail
        const int cpf_shift = (cpf > 0 && (cpf & (cpf - 1)) == 0) ? 31 - __clz(cpf) : -1;
        const bool vec_ok = (a.D & 3) == 0 && (a.fN & 3) == 0 && (a.fB & 3) == 0 &&
                            (reinterpret_cast<uintptr_t>(a.feats) & 15) == 0;
        uint32_t n = 0, acc_par = 0;
        // the source index of this thread's row, fetched one tile ahead (takes one global-memory latency off every tile)
        int64_t v_next = -1;
        if (((int64_t)blockIdx.x * kSaTile + tid) >> 5 < a.G) v_next = __ldg(a.idx + (int64_t)blockIdx.x * kSaTile + tid);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            // ---- gather, phase 1: thread r resolves row r (source index, cloud, feature-row offset) once and
            //      writes the row's TAIL chunks itself: [last features (D % 8) | xyz[idx] - new_xyz | zeros] ----
            {
                const int64_t m = tile * kSaTile + tid;
                const int64_t g = m >> 5;
                int64_t off = -1;
                float tail[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) tail[e] = 0.0f;
                const int64_t v = v_next;
                {
                    const int64_t m_next = m + (int64_t)gridDim.x * kSaTile;
                    v_next = (m_next >> 5) < a.G ? __ldg(a.idx + m_next) : -1;
                }
                if (g < a.G) {
                    if (v >= 0 && v < a.N) {
                        const int64_t b = g / a.S;
                        off = b * a.fB + v * a.fN;
                        const float *px = a.xyz + b * a.sB + v * a.sN;
                        const float *pc = a.new_xyz + g * 3;
                        const int nf = a.D & 7;        // features that share the tail's first chunk
                        const float dx = __fsub_rn(__ldg(px), __ldg(pc));
                        const float dy = __fsub_rn(__ldg(px + a.sC), __ldg(pc + 1));
                        const float dz = __fsub_rn(__ldg(px + 2 * a.sC), __ldg(pc + 2));
#pragma unroll
                        for (int j = 0; j < 10; ++j) {     // nf <= 7: the tail's non-zero part ends before column 10
                            float val = 0.0f;
                            if (j < nf) val = __ldg(a.feats + off + (cpf << 3) + j);
                            else if (j == nf) val = dx;
                            else if (j == nf + 1) val = dy;
                            else if (j == nf + 2) val = dz;
                            tail[j] = val;
                        }
                    }
                }
                s_off[tid] = off;
                for (int c = cpf, t = 0; c < cpr; ++c, ++t) {
                    uint4 o = make_uint4(0u, 0u, 0u, 0u);
                    if (t == 0) {
                        o.x = pack_bf16x2(tail[0], tail[1]); o.y = pack_bf16x2(tail[2], tail[3]);
                        o.z = pack_bf16x2(tail[4], tail[5]); o.w = pack_bf16x2(tail[6], tail[7]);
                    } else if (t == 1) {
                        o.x = pack_bf16x2(tail[8], tail[9]); o.y = pack_bf16x2(tail[10], tail[11]);
                        o.z = pack_bf16x2(tail[12], tail[13]); o.w = pack_bf16x2(tail[14], tail[15]);
                    }
                    *reinterpret_cast<uint4 *>(act + (size_t)(c >> 3) * kSaSlab + sw128_offset(tid, c & 7)) = o;
                }
            }
            named_bar_sync(1, 128);
            // ---- phase 2: the pure-feature chunks, 8 consecutive lanes per row (coalesced 256-byte segments); kGU
            //      chunks per thread and pass with all their global loads issued before the first conversion, so a
            //      thread keeps 2*kGU 16-byte requests in flight (the gather is L2-latency bound otherwise) ----
            const int n_tasks = kSaTile * cpf;
            for (int q0 = tid; q0 < n_tasks; q0 += 128 * kGU) {
                float4 lo[kGU], hi[kGU];
                int rr[kGU], cc[kGU];
#pragma unroll
                for (int u = 0; u < kGU; ++u) {
                    const int q = q0 + 128 * u;
                    const int r = cpf_shift >= 0 ? q >> cpf_shift : q / cpf;
                    const int c = q - r * cpf;
                    rr[u] = r;
                    cc[u] = c;
                    lo[u] = hi[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (q >= n_tasks) continue;
                    const int64_t off = s_off[r];
                    if (off < 0) continue;
                    const float *src = a.feats + off + (c << 3);
                    if (vec_ok) {
                        lo[u] = __ldg(reinterpret_cast<const float4 *>(src));
                        hi[u] = __ldg(reinterpret_cast<const float4 *>(src) + 1);
                    } else {
                        lo[u] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
                        hi[u] = make_float4(__ldg(src + 4), __ldg(src + 5), __ldg(src + 6), __ldg(src + 7));
                    }
                }
#pragma unroll
                for (int u = 0; u < kGU; ++u) {
                    if (q0 + 128 * u >= n_tasks) continue;
                    uint4 o;
                    o.x = pack_bf16x2(lo[u].x, lo[u].y); o.y = pack_bf16x2(lo[u].z, lo[u].w);
                    o.z = pack_bf16x2(hi[u].x, hi[u].y); o.w = pack_bf16x2(hi[u].z, hi[u].w);
                    *reinterpret_cast<uint4 *>(act + (size_t)(cc[u] >> 3) * kSaSlab + sw128_offset(rr[u], cc[u] & 7)) = o;
                }
            }
            fence_before_sync();      // the previous tile's TMEM reads are done before this tile's first MMA
            fence_proxy_async();      // generic-proxy writes of the tile -> visible to the tensor core
            named_bar_sync(1, 128);

            int ss_base = 0;
            for (int l = 0; l < a.L; ++l) {
                if (warp_u == 0) {      // warp 0, converged; one elected lane issues (tc_common.cuh: elect_one_sync)
                    fence_after_sync();
                    for (int c = a.chunk_begin[l]; c < a.chunk_begin[l + 1]; ++c, ++n) {
                        const SaChunk &ch = a.chunks[c];
                        const uint32_t s = a.resident ? (uint32_t)c : n % (uint32_t)a.stages;
                        const uint32_t par = a.resident ? 0u : (n / (uint32_t)a.stages) & 1u;
                        mbar_wait(&bar_full[s], par);
                        fence_after_sync();
                        const uint64_t ad = make_desc(smem_addr(act + (size_t)ch.kc * kSaSlab), 0, 1024);
                        const uint64_t bd = make_desc(smem_addr(ring + (size_t)s * a.slot_bytes), 0, 1024);
                        const int nk = ch.nk;
                        const uint32_t d_acc = tmem + ch.tmem_col, idesc = ch.idesc, first = (uint32_t)(ch.kc != 0);
                        if (elect_one_sync()) {
                            for (int j = 0; j < nk; ++j) umma_bf16(d_acc, ad + (uint64_t)(2 * j), bd + (uint64_t)(2 * j), idesc, first | (uint32_t)(j != 0));
                            if (!a.resident) umma_commit(&bar_empty[s]);
                        }
                        __syncwarp();
                    }
                    if (elect_one_sync()) umma_commit(&bar_acc);
                    __syncwarp();
                }
                mbar_wait(&bar_acc, acc_par);
                acc_par ^= 1u;
                fence_after_sync();
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
                const int Nl = a.width[l];
                if (l + 1 < a.L) {
                    // ---- BatchNorm + ReLU -> bf16 -> the next layer's A operand, row `tid` of the same tile ----
                    for (int c0 = 0; c0 < Nl; c0 += 16) {
                        float v[16];
                        tmem_ld16(taddr + c0, v);
#pragma unroll
                        for (int i4 = 0; i4 < 16; i4 += 4) {
                            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + ss_base + c0 + i4);
                            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + ss_base + c0 + i4);
                            v[i4 + 0] = fmaxf(fmaf(v[i4 + 0], sc.x, sh.x), 0.0f);
                            v[i4 + 1] = fmaxf(fmaf(v[i4 + 1], sc.y, sh.y), 0.0f);
                            v[i4 + 2] = fmaxf(fmaf(v[i4 + 2], sc.z, sh.z), 0.0f);
                            v[i4 + 3] = fmaxf(fmaf(v[i4 + 3], sc.w, sh.w), 0.0f);
                        }
                        uint4 lo, hi;
                        lo.x = pack_bf16x2(v[0], v[1]);   lo.y = pack_bf16x2(v[2], v[3]);
                        lo.z = pack_bf16x2(v[4], v[5]);   lo.w = pack_bf16x2(v[6], v[7]);
                        hi.x = pack_bf16x2(v[8], v[9]);   hi.y = pack_bf16x2(v[10], v[11]);
                        hi.z = pack_bf16x2(v[12], v[13]); hi.w = pack_bf16x2(v[14], v[15]);
                        uint8_t *slab = act + (size_t)(c0 >> 6) * kSaSlab;
                        const int ch = (c0 & 63) >> 3;
                        *reinterpret_cast<uint4 *>(slab + sw128_offset(tid, ch)) = lo;
                        *reinterpret_cast<uint4 *>(slab + sw128_offset(tid, ch + 1)) = hi;
                    }
                    fence_before_sync();
                    fence_proxy_async();
                    named_bar_sync(1, 128);
                } else {
                    // ---- BatchNorm + ReLU + max over the warp's 32 samples -> out[g, :] ----
                    const int64_t g = tile * 4 + warp;
                    for (int c0 = 0; c0 < Nl; c0 += 32) {
                        float v[32];
                        tmem_ld32(taddr + c0, v);
                        unsigned keep = 0u;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float y = fmaxf(fmaf(v[i], s_scale[ss_base + c0 + i], s_shift[ss_base + c0 + i]), 0.0f);
                            const unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(y));
                            if (lane == i) keep = mx;
                        }
                        if (g < a.G && c0 + lane < Nl) a.out[g * Nl + c0 + lane] = __uint_as_float(keep);
                    }
                }
                ss_base += (Nl + 31) & ~31;
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

static int sa_round_up(int v, int m) { return (v + m - 1) / m * m; }

struct SaPlan {
    SaFusedArgs a;
    size_t wimg_bytes, dyn_smem;
    int per_sm;
    bool ok;
    const char *why;
};

// Chunk table, shared-memory plan and occupancy for one level; everything but the pointers.
static SaPlan sa_plan(int D, int L, const int *widths) {
    SaPlan p;
    memset(&p, 0, sizeof(p));
    p.ok = false;
    SaFusedArgs &a = p.a;
    if (L < 1 || L > kSaMaxLayers) { p.why = "1..4 layers"; return p; }
    a.L = L;
    a.D = D;
    int K = D + 3, n_chunks = 0, max_cols = 32, a_slabs = 1, slot = 0, ss = 0;
    size_t off = 0;
    for (int l = 0; l < L; ++l) {
        const int N = widths[l];
        if (N < 1 || N > 512) { p.why = "layer width must be 1..512"; return p; }
        if (l + 1 < L && (N & 15)) { p.why = "hidden widths must be multiples of 16"; return p; }
        a.width[l] = N;
        a.chunk_begin[l] = n_chunks;
        const int KC = (K + 63) / 64;
        if (KC > a_slabs) a_slabs = KC;
        int cols = 0;
        for (int n0 = 0; n0 < N; n0 += 256) {
            const int nb = N - n0 < 256 ? N - n0 : 256, n_pad = sa_round_up(nb, 16);
            for (int kc = 0; kc < KC; ++kc) {
                if (n_chunks >= kSaMaxChunks) { p.why = "too many weight chunks"; return p; }
                SaChunk &c = a.chunks[n_chunks++];
                c.off = (uint32_t)off;
                c.bytes = (uint32_t)n_pad * 128u;
                c.idesc = make_idesc_bf16(kSaTile, n_pad, 0, 0);
                c.tmem_col = (uint16_t)n0;
                c.layer = (uint8_t)l;
                c.kc = (uint8_t)kc;
                const int k_left = K - kc * 64;
                c.nk = (uint8_t)(k_left >= 64 ? 4 : (k_left + 15) / 16);
                off += c.bytes;
                if ((int)c.bytes > slot) slot = (int)c.bytes;
            }
            cols = n0 + n_pad;
        }
        if (sa_round_up(N, 32) > cols) cols = sa_round_up(N, 32);     // the pooling epilogue reads 32 columns at a time
        if (cols > max_cols) max_cols = cols;
        ss += sa_round_up(N, 32);
        K = N;
    }
    a.chunk_begin[L] = n_chunks;
    a.n_chunks = n_chunks;
    a.a_slabs = a_slabs;
    a.slot_bytes = slot;
    a.ss_floats = ss;
    int tc = 32;
    while (tc < max_cols) tc <<= 1;
    if (tc > 512) { p.why = "accumulator exceeds tensor memory"; return p; }
    a.tmem_cols = tc;
    p.wimg_bytes = off;
    const size_t fixed = 1024 + (size_t)a_slabs * kSaSlab + (size_t)ss * 8 + 128 * 8;
    const size_t budget = 227 * 1024 - 2048;       // static shared memory (barriers) stays far below 2 KB
    // resident weights when the whole level fits next to the activation tile (then as many CTAs per SM as fit,
    // tensor memory allowing); otherwise a ring of at least two chunk slots, one CTA per SM
    if (fixed + (size_t)n_chunks * slot <= budget) {
        a.resident = 1;
        a.stages = n_chunks;
    } else {
        a.resident = 0;
        a.stages = (int)((budget - fixed) / slot);
        if (a.stages > kSaMaxChunks) a.stages = kSaMaxChunks;
        if (a.stages < 2) { p.why = "level does not fit shared memory"; return p; }
    }
    p.dyn_smem = fixed + (size_t)a.stages * slot;
    int per_sm = (int)((227 * 1024) / (p.dyn_smem + 2048));
    if (per_sm > 512 / tc) per_sm = 512 / tc;
    if (per_sm > 6) per_sm = 6;
    if (per_sm < 1) per_sm = 1;
    p.per_sm = per_sm;
    p.ok = true;
    return p;
}

}  // namespace pn2

using namespace pn2;

extern "C" size_t pn2_sa_fused_eval_workspace_bytes(int D, int L, const int *widths_host) {
    if (!widths_host) return 0;
    const SaPlan p = sa_plan(D, L, widths_host);
    return p.ok ? p.wimg_bytes + 1024 : 0;
}

extern "C" int pn2_sa_fused_eval(const float *xyz, int64_t sB, int64_t sN, int64_t sC, const float *new_xyz,
                                 const float *feats, int64_t fB, int64_t fN, const int64_t *idx, int B, int N, int S,
                                 int nsample, int D, int L, const int *widths_host, const float *const *W_host,
                                 const float *const *bias_host, const float *const *scale_host,
                                 const float *const *shift_host, float *out, void *workspace, void *stream) {
    PN2_REQUIRE(xyz && new_xyz && idx && out && workspace && widths_host && scale_host && shift_host && bias_host,
                "sa_fused_eval: null pointer");
    const bool prepacked = W_host == nullptr;     // the workspace still holds the images a previous call packed from the same weights
    PN2_REQUIRE(B >= 0 && N > 0 && S >= 0 && D >= 0 && (D == 0 || feats), "sa_fused_eval: bad sizes B=%d N=%d S=%d D=%d", B, N, S, D);
    if (nsample != 32) {
        set_error("sa_fused_eval: nsample=%d (the fused kernel pools over exactly one warp of 32 samples)", nsample);
        return PN2_ERR_UNSUPPORTED;
    }
    SaPlan p = sa_plan(D, L, widths_host);
    if (!p.ok) {
        set_error("sa_fused_eval: unsupported level (%s)", p.why);
        return PN2_ERR_UNSUPPORTED;
    }
    if (B == 0 || S == 0) return PN2_OK;
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, sa_fused_eval_kernel<4>);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(sa_fused_eval_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - (int)fa.sharedSizeBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(sa_fused_eval_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024 - (int)fa.sharedSizeBytes);
        if (e != cudaSuccess) {
            set_error("sa_fused_eval: shared-memory opt-in failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return PN2_ERR_CUDA;
        }
        attr_done = true;
    }
    SaFusedArgs &a = p.a;
    uint8_t *img = (uint8_t *)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    // weight images: layer 0 with its input columns rotated by 3 (the gather writes features first, xyz last)
    int K = D + 3;
    for (int l = 0; l < L; ++l) {
        PN2_REQUIRE((prepacked || W_host[l]) && scale_host[l] && shift_host[l], "sa_fused_eval: null layer pointer");
        const int Nl = widths_host[l], KC = (K + 63) / 64;
        int c = a.chunk_begin[l];
        for (int n0 = 0; n0 < Nl && !prepacked; n0 += 256) {
            const int nb = Nl - n0 < 256 ? Nl - n0 : 256, n_pad = sa_round_up(nb, 16);
            const int total = KC * n_pad * 8;
            pack_weight_kernel<<<(total + 255) / 256, 256, 0, st>>>(W_host[l] + (int64_t)n0 * K, K, 1, nb, K, n_pad, KC,
                                                                     l == 0 && D > 0 ? 3 : 0, img + a.chunks[c].off);
            count_launch();
            c += KC;
        }
        a.scale[l] = scale_host[l];
        a.shift[l] = shift_host[l];
        a.bias[l] = bias_host[l];
        K = Nl;
    }
    int rc = check_launch("sa_fused_eval: pack_weight");
    if (rc != PN2_OK) return rc;
    a.xyz = xyz; a.sB = sB; a.sN = sN; a.sC = sC;
    a.new_xyz = new_xyz;
    a.feats = feats; a.fB = fB; a.fN = fN;
    a.idx = idx;
    a.N = N; a.S = S;
    a.G = (int64_t)B * S;
    a.Wimg = img;
    a.out = out;
    const int64_t n_tiles = (a.G * 32 + kSaTile - 1) / kSaTile;
    int64_t grid = (int64_t)p.per_sm * sm_budget();
    if (grid > n_tiles) grid = n_tiles;
    // rows in flight per thread during the gather: 8 when shared memory allows at most two CTAs per SM (registers are
    // plentiful then and nothing else hides the L2 latency), 4 otherwise
    if (p.per_sm <= 2)
        sa_fused_eval_kernel<8><<<(unsigned)grid, kSaThreads, p.dyn_smem, st>>>(a);
    else
        sa_fused_eval_kernel<4><<<(unsigned)grid, kSaThreads, p.dyn_smem, st>>>(a);
    count_launch();
    return check_launch("sa_fused_eval");
}
