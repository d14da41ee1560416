// common.cuh -- shared helpers of libpn2b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pn2b200.h"

namespace pn2 {

// ---- error / launch bookkeeping (the only global state of the library) --------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int check_launch(const char *what);   // cudaGetLastError -> PN2_OK / PN2_ERR_CUDA

#define PN2_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            pn2::set_error(__VA_ARGS__);  \
            return PN2_ERR_ARG;           \
        }                                 \
    } while (0)

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs
constexpr int kStatReplicas = PN2_STAT_REPLICAS;   // copies of every fp64 column-sum accumulator (atomic contention)

// ---- exact fp32 arithmetic of the reference (never contracted by the compiler) -
// sum(v**2, -1) of pointnet2_utils.py:38-39 == (x*x + y*y) + z*z, separately rounded
__device__ __forceinline__ float sq_norm3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
// square_distance of pointnet2_utils.py:37-39:
//   mm = fma(az,bz, fma(ay,by, ax*bx));  d = ((-2*mm) + |src|^2) + |dst|^2
__device__ __forceinline__ float expanded_sqdist(float ax, float ay, float az, float an, float bx,
                                                 float by, float bz, float bn) {
    float mm = __fmul_rn(ax, bx);
    mm = __fmaf_rn(ay, by, mm);
    mm = __fmaf_rn(az, bz, mm);
    return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, mm), an), bn);
}

// ---- row element access for fp32 / bf16 activations -----------------------------
template <typename T> __device__ __forceinline__ float ld_act(const T *p);
template <> __device__ __forceinline__ float ld_act<float>(const float *p) { return *p; }
template <> __device__ __forceinline__ float ld_act<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_act(T *p, float v);
template <> __device__ __forceinline__ void st_act<float>(float *p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_act<__nv_bfloat16>(__nv_bfloat16 *p, float v) {
    *p = __float2bfloat16_rn(v);
}

inline bool valid_dtype(int d) { return d == PN2_F32 || d == PN2_BF16; }

// dispatch a callable templated on the activation storage type
#define PN2_DISPATCH_DTYPE(dtype, T, ...)                 \
    do {                                                  \
        if ((dtype) == PN2_F32) {                         \
            using T = float;                              \
            __VA_ARGS__;                                  \
        } else {                                          \
            using T = __nv_bfloat16;                      \
            __VA_ARGS__;                                  \
        }                                                 \
    } while (0)

inline int grid_for(int64_t work_items, int threads, int max_blocks = kNumSMs * 16) {
    int64_t b = (work_items + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

}  // namespace pn2
