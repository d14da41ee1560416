// api.cu -- library-level entry points of libpn2b200.so and the dtype dispatch of the MLP layer
// calls (fp32 rows -> FMA-pipe kernels in linear_simt.cu, bf16 rows -> tcgen05 kernels in
// linear_tc.cu).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace pn2 {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return PN2_ERR_CUDA;
    }
    return PN2_OK;
}

int simt_linear_nt(const void *X, int ldx, int x_dtype, const float *in_scale, const float *in_shift,
                   const float *W, int64_t w_sn, int64_t w_sk, const float *bias, int64_t M, int K, int N,
                   void *Z, int ldz, int z_dtype, double *stat_accum, cudaStream_t st);
int simt_linear_wgrad(const void *dZ, int lddz, int dz_dtype, const void *X, int ldx, int x_dtype,
                      const float *in_scale, const float *in_shift, int64_t M, int K, int N, float *dW,
                      void *scratch, cudaStream_t st);
size_t simt_wgrad_scratch_bytes(int64_t M, int K, int N);
int linear_num_partials(int64_t M);
size_t tc_wpack_bytes(int K, int N);
int tc_linear_nt(const void *X, int ldx, const float *in_scale, const float *in_shift, const float *W, int64_t w_sn,
                 int64_t w_sk, const float *bias, int64_t M, int K, int N, void *Z, int ldz, double *stat_accum,
                 void *wpack, cudaStream_t st);

size_t tc_wgrad_scratch_bytes(int64_t M, int K, int N);
int tc_linear_wgrad(const void *dZ, int lddz, const void *X, int ldx, const float *in_scale, const float *in_shift,
                    int64_t M, int K, int N, float *dW, void *scratch, cudaStream_t st);

static bool tc_enabled() {
    static int state = -1;
    if (state < 0) {
        const char *e = getenv("PN2_DISABLE_TC");
        state = (e && e[0] == '1') ? 0 : 1;
    }
    return state == 1;
}
static bool tc_eligible(int a_dtype, int lda, int b_dtype, int ldb, const void *wpack) {
    return tc_enabled() && wpack && a_dtype == PN2_BF16 && b_dtype == PN2_BF16 && lda % 8 == 0 && ldb % 8 == 0;
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_version(void) { return PN2_VERSION; }
extern "C" const char *pn2_last_error(void) { return g_err; }
extern "C" unsigned long long pn2_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int pn2_linear_fwd(const void *X, int ldx, int x_dtype, const float *in_scale, const float *in_shift,
                              const float *W, const float *bias, int64_t M, int K, int N, void *Z, int ldz,
                              int z_dtype, double *stat_accum, void *wpack, void *stream) {
    PN2_REQUIRE(X && W && Z, "linear_fwd: null pointer");
    PN2_REQUIRE(M >= 0 && K >= 1 && N >= 1 && ldx >= K && ldz >= N, "linear_fwd: bad sizes M=%lld K=%d N=%d ldx=%d ldz=%d",
                (long long)M, K, N, ldx, ldz);
    PN2_REQUIRE(valid_dtype(x_dtype) && valid_dtype(z_dtype), "linear_fwd: bad dtype");
    PN2_REQUIRE(!in_scale == !in_shift, "linear_fwd: in_scale and in_shift go together");
    PN2_REQUIRE(!stat_accum || N <= 4096, "linear_fwd: N=%d too wide for the statistics epilogue", N);
    if (M == 0) return PN2_OK;
    if (tc_eligible(x_dtype, ldx, z_dtype, ldz, wpack) && N <= 4096)
        return tc_linear_nt(X, ldx, in_scale, in_shift, W, K, 1, bias, M, K, N, Z, ldz, stat_accum, wpack,
                            (cudaStream_t)stream);
    return simt_linear_nt(X, ldx, x_dtype, in_scale, in_shift, W, K, 1, bias, M, K, N, Z, ldz, z_dtype, stat_accum,
                          (cudaStream_t)stream);
}

extern "C" int pn2_linear_bwd_data(const void *dZ, int lddz, int dz_dtype, const float *W, int64_t M, int K, int N,
                                   void *dX, int lddx, int dx_dtype, void *wpack, void *stream) {
    PN2_REQUIRE(dZ && W && dX, "linear_bwd_data: null pointer");
    PN2_REQUIRE(M >= 0 && K >= 1 && N >= 1 && lddz >= N && lddx >= K, "linear_bwd_data: bad sizes");
    PN2_REQUIRE(valid_dtype(dz_dtype) && valid_dtype(dx_dtype), "linear_bwd_data: bad dtype");
    if (M == 0) return PN2_OK;
    // dX[m,k] = sum_n dZ[m,n] W[n,k]: the NT kernel with the roles of W's two strides swapped
    if (tc_eligible(dz_dtype, lddz, dx_dtype, lddx, wpack))
        return tc_linear_nt(dZ, lddz, nullptr, nullptr, W, 1, K, nullptr, M, N, K, dX, lddx, nullptr, wpack,
                            (cudaStream_t)stream);
    return simt_linear_nt(dZ, lddz, dz_dtype, nullptr, nullptr, W, 1, K, nullptr, M, N, K, dX, lddx, dx_dtype, nullptr,
                          (cudaStream_t)stream);
}

extern "C" size_t pn2_linear_wpack_bytes(int K, int N) { return tc_wpack_bytes(K, N); }

extern "C" size_t pn2_linear_wgrad_scratch_bytes(int64_t M, int K, int N) {
    size_t a = simt_wgrad_scratch_bytes(M, K, N), b = tc_wgrad_scratch_bytes(M, K, N);
    return a > b ? a : b;
}

extern "C" int pn2_linear_bwd_weight(const void *dZ, int lddz, int dz_dtype, const void *X, int ldx, int x_dtype,
                                     const float *in_scale, const float *in_shift, int64_t M, int K, int N,
                                     float *dW, void *scratch, void *stream) {
    PN2_REQUIRE(dZ && X && dW && scratch, "linear_bwd_weight: null pointer");
    PN2_REQUIRE(M >= 1 && K >= 1 && N >= 1 && lddz >= N && ldx >= K, "linear_bwd_weight: bad sizes");
    PN2_REQUIRE(valid_dtype(dz_dtype) && valid_dtype(x_dtype), "linear_bwd_weight: bad dtype");
    PN2_REQUIRE(!in_scale == !in_shift, "linear_bwd_weight: in_scale and in_shift go together");
    if (tc_eligible(dz_dtype, lddz, x_dtype, ldx, scratch))
        return tc_linear_wgrad(dZ, lddz, X, ldx, in_scale, in_shift, M, K, N, dW, scratch, (cudaStream_t)stream);
    return simt_linear_wgrad(dZ, lddz, dz_dtype, X, ldx, x_dtype, in_scale, in_shift, M, K, N, dW, scratch,
                             (cudaStream_t)stream);
}
