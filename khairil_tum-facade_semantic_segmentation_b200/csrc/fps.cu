// fps.cu -- K1: farthest point sampling as a persistent one-CTA-(or one-cluster-)per-cloud
// kernel.  Replaces the Python loop of /root/reference/models/pointnet2_utils.py:63-84.
//
// Every point of the cloud lives in registers of exactly one thread (x, y, z and its running
// minimum distance) for the whole kernel, two points per 64-bit register pair so that the
// distance arithmetic runs on the packed fp32x2 pipe (FADD2 / FFMA2: one instruction, two
// points).  One iteration is
//     packed distance update + FMNMX (min) + FMNMX3 (running max)          [P/2 x 11 instr]
//     -> warp REDUX.MAX of the value -> the (rare) lanes holding it find their lowest k
//     -> REDUX.MIN of the index -> one 8-byte slot per warp in shared memory
//     -> ONE __syncthreads -> every warp reduces the <=32 slots itself (2 x REDUX)
//     -> the winner's coordinates are one broadcast LDS.128 from the cloud's copy in shared memory
// and, for clouds too large for one SM (N > 8192, e.g. the 65536-point microbenchmark), one
// more exchange of CTA winners (with coordinates) through distributed shared memory with an
// mbarrier per parity.
//
// Bit-exactness contract: dist = (dx*dx + dy*dy) + dz*dz with separately
// rounded sub/mul/add (no FMA contraction), distance = min(dist, distance) starting from
// float32(1e10), next centroid = LOWEST index among the maxima (torch.max semantics).
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with explicit rounding
// modifiers, so each square is written as fma.rn.f32x2(d, d, -0.0) with the -0.0 pair passed
// as a KERNEL ARGUMENT (opaque to the compiler): d*d + (-0.0) is exactly the rounded product
// (a square is never -0), and an fma feeding an add cannot be contracted further.
#include <stdlib.h>

#include "common.cuh"

namespace pn2 {

typedef unsigned long long u64;

struct __align__(32) FpsSlot {
    unsigned d;    // float bits of the candidate's min-distance (>= +0, so uint order == float order)
    unsigned idx;  // global point index, 0xffffffff for empty slots
    float x, y, z;
    unsigned pad[3];
};

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}

// ---- packed fp32x2 arithmetic, round-to-nearest, never contracted (see file header) ----
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 square2(u64 a, u64 neg_zero2) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(r) : "l"(a), "l"(neg_zero2));
    return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// max d, ties -> min idx, over the lanes of a warp (every lane gets the result)
__device__ __forceinline__ void warp_argmax(unsigned d, unsigned idx, unsigned &wd, unsigned &wi) {
    wd = __reduce_max_sync(0xffffffffu, d);
    wi = __reduce_min_sync(0xffffffffu, (d == wd) ? idx : 0xffffffffu);
}

// Dynamic shared memory: float4 {x, y, z, 0} of this CTA's points, indexed by (global index - rank * chunk).
template <int P, bool CLUSTER, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
fps_kernel(const float *__restrict__ xyz, int64_t sB, int64_t sN, int64_t sC, int N, int npoint,
           const int64_t *__restrict__ start_idx, int64_t *__restrict__ out_idx,
           float *__restrict__ out_xyz, u64 neg_zero2) {
    static_assert(P % 2 == 0, "points are held in pairs");
    extern __shared__ float4 s_pts[];
    __shared__ uint2 warp_slots[2][32];
    __shared__ FpsSlot cta_slots[2][16];
    __shared__ __align__(8) unsigned long long mbar[2];

    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const unsigned CL = CLUSTER ? cluster_nctarank() : 1u;
    const unsigned rank = CLUSTER ? cluster_ctarank() : 0u;
    const int cloud = blockIdx.x / CL;
    const int chunk = (N + (int)CL - 1) / (int)CL;
    const int first = (int)rank * chunk;
    const float *base = xyz + (int64_t)cloud * sB;

    if (CLUSTER) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[0])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[1])), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        cluster_sync_all();
    }

    // thread-local point k is CTA-local point tid + k*T: ascending k == ascending index
    u64 X[P / 2], Y[P / 2], Z[P / 2];
    float md[P];
#pragma unroll
    for (int j = 0; j < P / 2; ++j) {
        float x[2], y[2], z[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 2 * j + h;
            const int loc = tid + k * T;
            const int gi = first + loc;
            const bool valid = (loc < chunk) && (gi < N);
            if (valid) {
                const float *p = base + (int64_t)gi * sN;
                x[h] = p[0];
                y[h] = p[sC];
                z[h] = p[2 * sC];
                md[k] = 1e10f;
                s_pts[loc] = make_float4(x[h], y[h], z[h], 0.0f);
            } else {   // padding: distance pinned at 0 and never selected (real points win ties by index)
                x[h] = y[h] = z[h] = 0.0f;
                md[k] = 0.0f;
            }
        }
        X[j] = pack2(x[0], x[1]);
        Y[j] = pack2(y[0], y[1]);
        Z[j] = pack2(z[0], z[1]);
    }

    unsigned cur = (unsigned)start_idx[cloud];
    float cx, cy, cz;
    {
        const float *p = base + (int64_t)cur * sN;
        cx = p[0];
        cy = p[sC];
        cz = p[2 * sC];
    }
    __syncthreads();   // s_pts complete

    for (int it = 0; it < npoint; ++it) {
        if (tid == 0 && rank == 0) {
            int64_t o = (int64_t)cloud * npoint + it;
            out_idx[o] = (int64_t)cur;
            if (out_xyz) {
                out_xyz[o * 3 + 0] = cx;
                out_xyz[o * 3 + 1] = cy;
                out_xyz[o * 3 + 2] = cz;
            }
        }
        if (it == npoint - 1) break;
        const int par = it & 1;

        // ---- distance update + running maximum ----
        const u64 CX = pack2(cx, cx), CY = pack2(cy, cy), CZ = pack2(cz, cz);
        float best = 0.0f;
#pragma unroll
        for (int j = 0; j < P / 2; ++j) {
            const u64 dx = sub2(X[j], CX), dy = sub2(Y[j], CY), dz = sub2(Z[j], CZ);
            const u64 d = add2(add2(square2(dx, neg_zero2), square2(dy, neg_zero2)), square2(dz, neg_zero2));
            float d0, d1;
            unpack2(d, d0, d1);
            md[2 * j] = fminf(d0, md[2 * j]);
            md[2 * j + 1] = fminf(d1, md[2 * j + 1]);
            best = max3(best, md[2 * j], md[2 * j + 1]);
        }

        // ---- warp level: value first, then the lowest index among the lanes (and k) that hold it.  The lane's
        //      own lowest k (of its own maximum) does not depend on the warp result, so it overlaps the REDUX ----
        int bk = 0;
#pragma unroll
        for (int k = P - 1; k >= 0; --k) bk = (md[k] == best) ? k : bk;
        const unsigned wd = __reduce_max_sync(0xffffffffu, __float_as_uint(best));
        const unsigned loc = (unsigned)(tid + bk * T);
        // padding slots (md == 0, beyond `chunk`) can only match when every remaining distance is 0: then the
        // lowest REAL index must win, so they are excluded explicitly (a lower k of the same thread is never padding
        // while a higher one is real)
        const bool mine = __float_as_uint(best) == wd && (int)loc < chunk && first + (int)loc < N;
        const unsigned wi = __reduce_min_sync(0xffffffffu, mine ? (unsigned)first + loc : 0xffffffffu);

        unsigned gd, gi;
        if (W == 1) {
            gd = wd;
            gi = wi;
        } else {
            if (lane == 0) warp_slots[par][warp] = make_uint2(wd, wi);
            __syncthreads();
            // ---- CTA level: every warp reduces the W slots itself (no second barrier) ----
            uint2 s = make_uint2(0u, 0xffffffffu);
            if (lane < W) s = warp_slots[par][lane];
            warp_argmax(s.x, s.y, gd, gi);
        }
        float wx, wy, wz;
        {
            const float4 c = s_pts[gi - (unsigned)first];   // broadcast read; gi is always one of THIS CTA's real points
            wx = c.x;
            wy = c.y;
            wz = c.z;
        }

        if (CLUSTER) {
            // ---- cluster level: lane r of warp 0 posts this CTA's winner into CTA r with asynchronous remote
            //      stores that complete on the RECEIVER's mbarrier (complete_tx); the receiver arms the barrier with
            //      the expected byte count and waits with CTA scope -- no cluster-scope acquire, which would cost an
            //      L1 invalidate (CCTL.IVALL) per thread and iteration ----
            if (tid == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[par])),
                             "r"(CL * 20u)
                             : "memory");
            if (warp == 0 && lane < (int)CL) {
                const uint32_t rs = map_to_cta(smem_u32(&cta_slots[par][rank]), (unsigned)lane);
                const uint32_t rb = map_to_cta(smem_u32(&mbar[par]), (unsigned)lane);
                asm volatile(
                    "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(rs),
                    "r"(gd), "r"(gi), "r"(__float_as_uint(wx)), "r"(__float_as_uint(wy)), "r"(rb)
                    : "memory");
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(rs + 16),
                             "r"(__float_as_uint(wz)), "r"(rb)
                             : "memory");
            }
            const unsigned phase = (unsigned)(it >> 1) & 1u;
            const uint32_t lb = smem_u32(&mbar[par]);
            unsigned done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(lb), "r"(phase)
                    : "memory");
            }
            FpsSlot s;
            s.d = 0; s.idx = 0xffffffffu; s.x = s.y = s.z = 0.0f;
            if (lane < (int)CL) s = cta_slots[par][lane];
            warp_argmax(s.d, s.idx, gd, gi);
            const int src = __ffs(__ballot_sync(0xffffffffu, s.idx == gi)) - 1;
            wx = __shfl_sync(0xffffffffu, s.x, src);
            wy = __shfl_sync(0xffffffffu, s.y, src);
            wz = __shfl_sync(0xffffffffu, s.z, src);
        }
        cur = gi;
        cx = wx;
        cy = wy;
        cz = wz;
    }
    if (CLUSTER) cluster_sync_all();   // nobody may exit while a peer can still write into it
}

template <int P, bool CLUSTER, int MAXT>
static int launch_fps(const float *xyz, int64_t sB, int64_t sN, int64_t sC, int B, int N, int npoint,
                      const int64_t *start, int64_t *out_idx, float *out_xyz, int T, int CL,
                      cudaStream_t st) {
    auto kern = fps_kernel<P, CLUSTER, MAXT>;
    static bool attr_done = false;      // per instantiation
    if (!attr_done) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kern);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes);
        if (e != cudaSuccess) {
            set_error("fps: shared-memory opt-in failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return PN2_ERR_CUDA;
        }
        attr_done = true;
    }
    const int chunk = (N + CL - 1) / CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CL));
    cfg.blockDim = dim3((unsigned)T);
    cfg.dynamicSmemBytes = (size_t)chunk * sizeof(float4);     // the CTA's copy of its points (winner look-up)
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (CLUSTER) {
        if (CL > 8) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) {
                set_error("fps: non-portable cluster size refused: %s", cudaGetErrorString(e));
                return PN2_ERR_CUDA;
            }
        }
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    const u64 neg_zero2 = 0x8000000080000000ull;   // (-0.0f, -0.0f): see the file header
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xyz, sB, sN, sC, N, npoint, start, out_idx, out_xyz, neg_zero2);
    count_launch();
    if (e != cudaSuccess) {
        set_error("fps launch failed: %s", cudaGetErrorString(e));
        return PN2_ERR_CUDA;
    }
    return PN2_OK;
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_farthest_point_sample(const float *xyz, int64_t sB, int64_t sN, int64_t sC, int B,
                                         int N, int npoint, const int64_t *start_idx,
                                         int64_t *out_idx, float *out_xyz, void *stream) {
    PN2_REQUIRE(B >= 0 && N > 0 && npoint >= 0, "fps: bad sizes B=%d N=%d npoint=%d", B, N, npoint);
    if (B == 0 || npoint == 0) return PN2_OK;
    PN2_REQUIRE(xyz && start_idx && out_idx, "fps: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // points per thread P, threads T, CTAs per cloud CL (see file header).  Every warp pays ~100 instructions per
    // iteration for the reductions whatever P is, so few fat warps beat many thin ones: measured on B200 (packed
    // math), 4096 -> 1024 x 32 clouds: P=16/T=256 0.377 ms, P=8/T=512 0.408 ms, P=4/T=1024 0.479 ms.
    int P, T, CL = 1;
    auto r32 = [](int v) { return (v + 31) / 32 * 32; };
    if (N <= 64) { P = 2; T = 32; }
    else if (N < 512) { P = 4; T = r32((N + 3) / 4); }
    else if (N < 2048) { P = 8; T = r32((N + 7) / 8); }
    else if (N <= 8192) { P = 16; T = r32((N + 15) / 16); }
    else {
        P = 32; T = 256;      // 65536 x 18 clouds, 8-CTA clusters: P=32 1.63 us/iteration, P=16 1.75, P=8 2.10
        CL = 2;
        while (CL <= 16 && (N + CL - 1) / CL > 8192) CL *= 2;
        {   // tuning knob: PN2_FPS_CL = 2|4|8|16 forces a (larger) cluster, i.e. fewer points per CTA
            const char *e = getenv("PN2_FPS_CL");
            const int want = e ? atoi(e) : 0;
            if ((want == 2 || want == 4 || want == 8 || want == 16) && want >= CL) CL = want;
        }
        if (CL > 16) {
            set_error("fps: N=%d exceeds the register-resident capacity of a 16-CTA cluster (131072)", N);
            return PN2_ERR_UNSUPPORTED;
        }
    }
    // tuning knob (profiles/fps_sweep.py): PN2_FPS_P = 4|8|16|32 trades warps per CTA for points per thread
    {
        const char *e = getenv("PN2_FPS_P");
        const int want = e ? atoi(e) : 0;
        const int per_cta = (N + CL - 1) / CL;
        if (N > 64 && (want == 4 || want == 8 || want == 16 || want == 32) && (per_cta + want - 1) / want <= (want == 32 ? 256 : want == 16 ? 512 : 1024)) {
            P = want;
            T = CL > 1 ? r32((per_cta + P - 1) / P) : r32((N + P - 1) / P);
        }
    }
#define PN2_FPS_ARGS xyz, sB, sN, sC, B, N, npoint, start_idx, out_idx, out_xyz, T
    if (CL > 1) {
        switch (P) {
            case 8: return launch_fps<8, true, 1024>(PN2_FPS_ARGS, CL, st);
            case 16: return launch_fps<16, true, 512>(PN2_FPS_ARGS, CL, st);
            default: return launch_fps<32, true, 256>(PN2_FPS_ARGS, CL, st);
        }
    }
    switch (P) {
        case 2: return launch_fps<2, false, 1024>(PN2_FPS_ARGS, 1, st);
        case 4: return launch_fps<4, false, 1024>(PN2_FPS_ARGS, 1, st);
        case 8: return launch_fps<8, false, 1024>(PN2_FPS_ARGS, 1, st);
        case 32: return launch_fps<32, false, 256>(PN2_FPS_ARGS, 1, st);
        default: return launch_fps<16, false, 512>(PN2_FPS_ARGS, 1, st);
    }
#undef PN2_FPS_ARGS
}
