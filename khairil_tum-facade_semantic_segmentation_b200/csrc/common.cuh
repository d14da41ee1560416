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
int sm_budget();               // SMs the persistent kernels size their grids for (pn2_set_sm_budget), default kNumSMs
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

// q = quo * d + rem for a small positive divisor: a shift when d is a power of two (channel chunks per row and nsample
// almost always are), 32-bit division when q fits, 64-bit division (~100 instructions) only as the last resort --
// the streaming kernels below do 2-3 of these per 16-byte element
__device__ __forceinline__ void fast_divmod(int64_t q, int d, int64_t &quo, int &rem) {
    if ((d & (d - 1)) == 0) {
        const int sh = 31 - __clz(d);
        quo = q >> sh;
        rem = (int)(q & (d - 1));
    } else if (q <= 0xffffffffLL) {
        const uint32_t u = (uint32_t)q, v = u / (uint32_t)d;
        quo = v;
        rem = (int)(u - v * (uint32_t)d);
    } else {
        quo = q / d;
        rem = (int)(q - quo * d);
    }
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

// ---- train-mode BatchNorm statistics -> scale/shift for ONE channel (shared by bn_train_finalize_kernel and the
// fused "last CTA finalizes" tail of the tensor-core layer kernel).  s1, s2: sum z, sum z^2 over M rows (bias excluded).
struct BnFinalize {           // device pointers; mirrors pn2_bn_finalize of include/pn2b200.h
    unsigned *ticket;
    const float *gamma, *beta, *conv_bias;
    float eps, momentum;
    float *running_mean, *running_var, *scale, *shift, *save_mean, *save_invstd;
    long long *num_batches_tracked;
    const float *momentum_dev;     // when non-null: the momentum, read at run time (graph replays follow a schedule)
};
__device__ __forceinline__ void bn_finalize_channel(double s1, double s2, int64_t M, int c, const float *gamma,
                                                    const float *beta, const float *conv_bias, float eps, float momentum,
                                                    float *running_mean, float *running_var, float *scale, float *shift,
                                                    float *save_mean, float *save_invstd) {
    double mean = s1 / (double)M;
    double var = s2 / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
    float sc = g * invstd;
    scale[c] = sc;
    shift[c] = b - (float)mean * sc;
    if (save_mean) save_mean[c] = (float)mean;
    if (save_invstd) save_invstd[c] = invstd;
    if (running_mean) {
        float full_mean = (float)mean + (conv_bias ? conv_bias[c] : 0.0f);
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * full_mean;
    }
    if (running_var) {
        double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
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
