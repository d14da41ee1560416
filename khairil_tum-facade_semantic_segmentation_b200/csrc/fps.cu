// fps.cu -- K1: farthest point sampling as a persistent one-CTA-(or one-cluster-)per-cloud
// kernel.  Replaces the Python loop of /root/reference/models/pointnet2_utils.py:63-84.
//
// Every point of the cloud lives in registers of exactly one thread (x, y, z and its running
// minimum distance) for the whole kernel; one iteration is
//     distance update  ->  per-thread arg-max  ->  warp arg-max (2x REDUX)  ->  one slot per
//     warp in shared memory  ->  ONE __syncthreads  ->  every warp reduces the <=32 slots itself
// and, for clouds too large for one SM (N > 8192, e.g. the 65536-point microbenchmark), one
// more exchange of CTA winners through distributed shared memory with an mbarrier per parity.
// The slot carries the winner's coordinates, so the next centroid never has to be re-read
// from global or shared memory.
//
// Bit-exactness contract: dist = (dx*dx + dy*dy) + dz*dz with separately
// rounded sub/mul/add (no FMA), distance = dist < distance ? dist : distance starting from
// float32(1e10), next centroid = LOWEST index among the maxima (torch.max semantics).
#include <stdlib.h>

#include "common.cuh"

namespace pn2 {

struct __align__(32) FpsSlot {
    unsigned d;    // float bits of the candidate's min-distance (>= +0, so uint order == float order)
    unsigned idx;  // global point index, 0x7fffffff for padding lanes
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

// Reduce (d, idx) over a warp: max d, ties -> min idx.  Returns the lane holding the winner.
__device__ __forceinline__ int warp_argmax(unsigned d, unsigned idx, unsigned &wd, unsigned &wi) {
    wd = __reduce_max_sync(0xffffffffu, d);
    unsigned cand = (d == wd) ? idx : 0xffffffffu;
    wi = __reduce_min_sync(0xffffffffu, cand);
    return __ffs(__ballot_sync(0xffffffffu, cand == wi)) - 1;
}

template <int P, bool CLUSTER, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
fps_kernel(const float *__restrict__ xyz, int64_t sB, int64_t sN, int64_t sC, int N, int npoint,
           const int64_t *__restrict__ start_idx, int64_t *__restrict__ out_idx,
           float *__restrict__ out_xyz) {
    __shared__ FpsSlot warp_slots[2][32];
    __shared__ FpsSlot cta_slots[2][16];
    __shared__ __align__(8) unsigned long long mbar[2];

    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const unsigned CL = CLUSTER ? cluster_nctarank() : 1u;
    const unsigned rank = CLUSTER ? cluster_ctarank() : 0u;
    const int cloud = blockIdx.x / CL;
    const int chunk = (N + (int)CL - 1) / (int)CL;
    const float *base = xyz + (int64_t)cloud * sB;

    if (CLUSTER) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[0])), "r"(CL));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[1])), "r"(CL));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        cluster_sync_all();
    }

    float px[P], py[P], pz[P], md[P];
    unsigned gidx[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        int j = tid + k * T;
        int gi = (int)rank * chunk + j;
        bool valid = (j < chunk) && (gi < N);
        if (valid) {
            const float *p = base + (int64_t)gi * sN;
            px[k] = p[0];
            py[k] = p[sC];
            pz[k] = p[2 * sC];
            md[k] = 1e10f;
            gidx[k] = (unsigned)gi;
        } else {   // padding: distance pinned at 0 and the largest index -> never beats a real point
            px[k] = py[k] = pz[k] = 0.0f;
            md[k] = 0.0f;
            gidx[k] = 0x7fffffffu;
        }
    }

    unsigned cur = (unsigned)start_idx[cloud];
    float cx, cy, cz;
    {
        const float *p = base + (int64_t)cur * sN;
        cx = p[0];
        cy = p[sC];
        cz = p[2 * sC];
    }

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

        // ---- distance update + per-thread arg-max (ascending k == ascending index) ----
        float best = -1.0f;
        int bk = 0;
#pragma unroll
        for (int k = 0; k < P; ++k) {
            float dx = __fsub_rn(px[k], cx), dy = __fsub_rn(py[k], cy), dz = __fsub_rn(pz[k], cz);
            float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            float m = (d < md[k]) ? d : md[k];
            md[k] = m;
            if (m > best) {
                best = m;
                bk = k;
            }
        }
        unsigned bi = gidx[0];
        float bx = px[0], by = py[0], bz = pz[0];
#pragma unroll
        for (int k = 1; k < P; ++k)
            if (bk == k) {
                bi = gidx[k];
                bx = px[k];
                by = py[k];
                bz = pz[k];
            }

        // ---- warp level ----
        unsigned wd, wi;
        int src = warp_argmax(__float_as_uint(best), bi, wd, wi);
        float wx, wy, wz;
        if (W == 1) {
            wx = __shfl_sync(0xffffffffu, bx, src);
            wy = __shfl_sync(0xffffffffu, by, src);
            wz = __shfl_sync(0xffffffffu, bz, src);
        } else {
            if (lane == src) {
                FpsSlot s;
                s.d = wd; s.idx = wi; s.x = bx; s.y = by; s.z = bz;
                s.pad[0] = s.pad[1] = s.pad[2] = 0;
                warp_slots[par][warp] = s;
            }
            __syncthreads();
            // ---- CTA level: every warp reduces the W slots itself (no second barrier) ----
            FpsSlot s;
            s.d = 0; s.idx = 0xffffffffu; s.x = s.y = s.z = 0.0f;
            if (lane < W) s = warp_slots[par][lane];
            src = warp_argmax(s.d, s.idx, wd, wi);
            wx = __shfl_sync(0xffffffffu, s.x, src);
            wy = __shfl_sync(0xffffffffu, s.y, src);
            wz = __shfl_sync(0xffffffffu, s.z, src);
        }

        if (CLUSTER) {
            // ---- cluster level: lane r of warp 0 posts this CTA's winner into CTA r ----
            if (warp == 0 && lane < (int)CL) {
                uint32_t rs = map_to_cta(smem_u32(&cta_slots[par][rank]), (unsigned)lane);
                asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rs), "r"(wd),
                             "r"(wi), "r"(__float_as_uint(wx)), "r"(__float_as_uint(wy))
                             : "memory");
                asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(rs + 16), "r"(__float_as_uint(wz))
                             : "memory");
                uint32_t rb = map_to_cta(smem_u32(&mbar[par]), (unsigned)lane);
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rb)
                             : "memory");
            }
            const unsigned phase = (unsigned)(it >> 1) & 1u;
            const uint32_t lb = smem_u32(&mbar[par]);
            unsigned done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(lb), "r"(phase)
                    : "memory");
            }
            FpsSlot s;
            s.d = 0; s.idx = 0xffffffffu; s.x = s.y = s.z = 0.0f;
            if (lane < (int)CL) s = cta_slots[par][lane];
            src = warp_argmax(s.d, s.idx, wd, wi);
            wx = __shfl_sync(0xffffffffu, s.x, src);
            wy = __shfl_sync(0xffffffffu, s.y, src);
            wz = __shfl_sync(0xffffffffu, s.z, src);
        }
        cur = wi;
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
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CL));
    cfg.blockDim = dim3((unsigned)T);
    cfg.dynamicSmemBytes = 0;
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xyz, sB, sN, sC, N, npoint, start, out_idx, out_xyz);
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
    // points per thread P, threads T, CTAs per cloud CL (see file header)
    int P, T, CL = 1;
    auto r32 = [](int v) { return (v + 31) / 32 * 32; };
    if (N <= 32) { P = 1; T = 32; }
    else if (N <= 64) { P = 2; T = 32; }
    else if (N < 512) { P = 4; T = r32((N + 3) / 4); }
    else if (N < 2048) { P = 8; T = r32((N + 7) / 8); }
    else if (N <= 4096) { P = 16; T = r32((N + 15) / 16); }   // measured on B200: 0.62 ms vs 0.94 ms (P=4) for 4096->1024
    else if (N <= 8192) { P = 8; T = 1024; }
    else {
        P = 8; T = 1024;
        CL = 2;
        while (CL <= 16 && (N + CL - 1) / CL > 8192) CL *= 2;
        if (CL > 16) {
            set_error("fps: N=%d exceeds the register-resident capacity of a 16-CTA cluster (131072)", N);
            return PN2_ERR_UNSUPPORTED;
        }
    }
    // tuning knob (profiles/microbench.py): PN2_FPS_P=8|16 trades warps per CTA for points per thread
    if (CL == 1 && N > 64) {
        const char *e = getenv("PN2_FPS_P");
        int want = e ? atoi(e) : 0;
        if ((want == 8 && N <= 8192) || (want == 16 && N <= 4096) || (want == 4 && N <= 4096)) {
            P = want;
            T = r32((N + P - 1) / P);
        }
    }
#define PN2_FPS_ARGS xyz, sB, sN, sC, B, N, npoint, start_idx, out_idx, out_xyz, T
    if (CL > 1) return launch_fps<8, true, 1024>(PN2_FPS_ARGS, CL, st);
    switch (P) {
        case 1: return launch_fps<1, false, 1024>(PN2_FPS_ARGS, 1, st);
        case 2: return launch_fps<2, false, 1024>(PN2_FPS_ARGS, 1, st);
        case 4: return launch_fps<4, false, 1024>(PN2_FPS_ARGS, 1, st);
        case 16: return launch_fps<16, false, 256>(PN2_FPS_ARGS, 1, st);
        default: return launch_fps<8, false, 1024>(PN2_FPS_ARGS, 1, st);
    }
#undef PN2_FPS_ARGS
}
