// api.cu -- library-level entry points of libpn2b200.so and the dtype dispatch of the MLP layer
// calls (fp32 rows -> FMA-pipe kernels in linear_simt.cu, bf16 rows -> tcgen05 kernels in
// linear_tc.cu).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

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
static std::atomic<int> g_sm_budget{0};
int sm_budget() {
    const int b = g_sm_budget.load(std::memory_order_relaxed);
    return b > 0 && b < kNumSMs ? b : kNumSMs;
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
                 void *wpack, cudaStream_t st, bool packed = false, const BnFinalize *fin = nullptr);
int tc_pack_weights(int n, const float *const *W, const int *K, const int *N, const int *transposed, void *const *wpack,
                    cudaStream_t st);

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
extern "C" int pn2_set_sm_budget(int sms) { return g_sm_budget.exchange(sms < 0 ? 0 : sms, std::memory_order_relaxed); }

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

extern "C" int pn2_pack_weights(int n, const float *const *W_host, const int *K_host, const int *N_host,
                                const int *transposed_host, void *const *wpack_host, void *stream) {
    PN2_REQUIRE(n >= 0 && (n == 0 || (W_host && K_host && N_host && transposed_host && wpack_host)), "pack_weights: null pointer");
    for (int i = 0; i < n; ++i)
        PN2_REQUIRE(W_host[i] && wpack_host[i] && K_host[i] >= 1 && N_host[i] >= 1, "pack_weights: bad job %d", i);
    if (n == 0) return PN2_OK;
    return tc_pack_weights(n, W_host, K_host, N_host, transposed_host, wpack_host, (cudaStream_t)stream);
}

static_assert(sizeof(pn2_bn_finalize) == sizeof(BnFinalize), "pn2_bn_finalize must mirror pn2::BnFinalize");

extern "C" int pn2_linear_bwd_weight_accum(const void *dZ, int lddz, int dz_dtype, const void *X, int ldx, int x_dtype,
                                           const float *in_scale, const float *in_shift, int64_t M, int K, int N,
                                           float *dW, void *stream) {
    PN2_REQUIRE(dZ && X && dW, "linear_bwd_weight_accum: null pointer");
    PN2_REQUIRE(M >= 1 && K >= 1 && N >= 1 && lddz >= N && ldx >= K, "linear_bwd_weight_accum: bad sizes");
    PN2_REQUIRE(!in_scale == !in_shift, "linear_bwd_weight_accum: in_scale and in_shift go together");
    PN2_REQUIRE(tc_eligible(dz_dtype, lddz, x_dtype, ldx, dW), "linear_bwd_weight_accum: bf16 rows with 16-byte row pitch only");
    return tc_linear_wgrad(dZ, lddz, X, ldx, in_scale, in_shift, M, K, N, dW, nullptr, (cudaStream_t)stream);
}

extern "C" int pn2_linear_fwd_prepacked(const void *X, int ldx, int x_dtype, const float *in_scale,
                                        const float *in_shift, const float *W, const float *bias, int64_t M, int K, int N,
                                        void *Z, int ldz, int z_dtype, double *stat_accum, const void *wpack,
                                        const pn2_bn_finalize *fin_host, void *stream) {
    PN2_REQUIRE(X && W && Z, "linear_fwd_prepacked: null pointer");
    PN2_REQUIRE(M >= 0 && K >= 1 && N >= 1 && ldx >= K && ldz >= N, "linear_fwd_prepacked: bad sizes M=%lld K=%d N=%d ldx=%d ldz=%d",
                (long long)M, K, N, ldx, ldz);
    PN2_REQUIRE(valid_dtype(x_dtype) && valid_dtype(z_dtype), "linear_fwd_prepacked: bad dtype");
    PN2_REQUIRE(!in_scale == !in_shift, "linear_fwd_prepacked: in_scale and in_shift go together");
    PN2_REQUIRE(!stat_accum || N <= 4096, "linear_fwd_prepacked: N=%d too wide for the statistics epilogue", N);
    PN2_REQUIRE(!fin_host || (stat_accum && fin_host->ticket && fin_host->scale && fin_host->shift && M > 0),
                "linear_fwd_prepacked: the fused finalize needs stat_accum, ticket, scale, shift and M > 0");
    if (M == 0) return PN2_OK;
    BnFinalize fin;
    if (fin_host) memcpy(&fin, fin_host, sizeof(fin));
    if (tc_eligible(x_dtype, ldx, z_dtype, ldz, wpack) && N <= 4096)
        return tc_linear_nt(X, ldx, in_scale, in_shift, W, K, 1, bias, M, K, N, Z, ldz, stat_accum, (void *)wpack,
                            (cudaStream_t)stream, true, fin_host ? &fin : nullptr);
    // rows this build does not run on the tensor cores: the FMA-pipe kernel, then the separate finalize launch
    int rc = simt_linear_nt(X, ldx, x_dtype, in_scale, in_shift, W, K, 1, bias, M, K, N, Z, ldz, z_dtype, stat_accum,
                            (cudaStream_t)stream);
    if (rc != PN2_OK || !fin_host) return rc;
    PN2_REQUIRE(!fin.momentum_dev, "linear_fwd_prepacked: momentum_dev needs the fused (tensor-core) finalize");
    return pn2_bn_train_finalize(stat_accum, M, N, fin.gamma, fin.beta, fin.conv_bias, fin.eps, fin.momentum,
                                 fin.running_mean, fin.running_var, fin.scale, fin.shift, fin.save_mean, fin.save_invstd,
                                 (int64_t *)fin.num_batches_tracked, stream);
}

extern "C" int pn2_linear_bwd_data_prepacked(const void *dZ, int lddz, int dz_dtype, const float *W, int64_t M, int K,
                                             int N, void *dX, int lddx, int dx_dtype, const void *wpack, void *stream) {
    PN2_REQUIRE(dZ && W && dX, "linear_bwd_data_prepacked: null pointer");
    PN2_REQUIRE(M >= 0 && K >= 1 && N >= 1 && lddz >= N && lddx >= K, "linear_bwd_data_prepacked: bad sizes");
    PN2_REQUIRE(valid_dtype(dz_dtype) && valid_dtype(dx_dtype), "linear_bwd_data_prepacked: bad dtype");
    if (M == 0) return PN2_OK;
    if (tc_eligible(dz_dtype, lddz, dx_dtype, lddx, wpack))
        return tc_linear_nt(dZ, lddz, nullptr, nullptr, W, 1, K, nullptr, M, N, K, dX, lddx, nullptr, (void *)wpack,
                            (cudaStream_t)stream, true, nullptr);
    return simt_linear_nt(dZ, lddz, dz_dtype, nullptr, nullptr, W, 1, K, nullptr, M, N, K, dX, lddx, dx_dtype, nullptr,
                          (cudaStream_t)stream);
}

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
