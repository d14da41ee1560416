/*
 * pn2b200.h -- C ABI of libpn2b200.so: the B200 (sm_100a) implementation of the
 * PointNet++ SSG set-abstraction / feature-propagation hot path of
 * KhairilAriffinYahya/Khairil_TUM-Facade_Semantic_Segmentation.
 *
 * Every entry point replaces a function (or a step of a module's forward /
 * backward) of the reference file models/pointnet2_utils.py; the line numbers
 * quoted below are into that file.  The reference has no FFI of its own (it is
 * pure PyTorch); the binding a maintainer adds is the ctypes stub shown in
 * INTEGRATION.md (this repo's khairil_tum-facade_semantic_segmentation_b200/_lib.py).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *    name ends in _host; `stream` is a cudaStream_t passed as void*.
 *  - all calls are asynchronous on `stream`, allocate nothing, keep no mutable
 *    global state besides the last-error string and a launch counter, and may be
 *    used from one thread per GPU.
 *  - return value: PN2_OK (0) or a negative PN2_ERR_* code; pn2_last_error()
 *    gives the text.  Nothing throws across the ABI.
 *  - "rows" = point-major activations: a [M, C] matrix with leading dimension ld
 *    (elements), row m = (cloud b, point/centroid s[, sample k]) flattened.
 *  - dtype codes: PN2_F32 / PN2_BF16 select the storage type of activation rows
 *    (accumulation and all statistics are always fp32).
 *  - index tensors are int64 at this boundary exactly as in the reference.
 */
#ifndef PN2B200_H
#define PN2B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN2_VERSION 100

enum { PN2_OK = 0, PN2_ERR_ARG = -1, PN2_ERR_CUDA = -2, PN2_ERR_UNSUPPORTED = -3 };
enum { PN2_F32 = 0, PN2_BF16 = 1 };
/* fp64 column-sum accumulators (stat_accum / accum below) hold PN2_STAT_REPLICAS copies of
 * [2][C] doubles: CTA i adds into copy i % PN2_STAT_REPLICAS (spreads same-address atomics),
 * the *_finalize entry points add the copies in a fixed order and zero them. */
#ifndef PN2_STAT_REPLICAS
#define PN2_STAT_REPLICAS 4
#endif

/* ---- library ---------------------------------------------------------------- */
int pn2_version(void);
const char *pn2_last_error(void);
/* kernels launched by this library since load (bench.py's "gpu_launches") */
unsigned long long pn2_launch_count(void);
/* SM budget of the persistent kernels (linear layers, fused set abstraction, fused backward layer): their grids are sized
 * for `sms` SMs instead of all 148 until changed again; 0 restores the default.  Returns the previous value.  A caller that
 * knows other work pins SMs at the same time (the pipelined trainer / predictor: the FPS chain of the next batch holds one
 * CTA per cloud) sizes the grid to the SMs that are really free, so that every CTA is resident at once instead of a
 * second partial wave.  Process-global; takes effect at launch (so at capture time for CUDA graphs). */
int pn2_set_sm_budget(int sms);

/* ---- a1 square_distance (:19-40) ---------------------------------------------
 * out[b,i,j] = ((-2*<src_i,dst_j>) + |src_i|^2) + |dst_j|^2 in the reference's
 * fp32 rounding order.  src [B,N,3], dst [B,M,3], out [B,N,M], all contiguous.
 * Provided for API parity; the network path never materialises this matrix. */
int pn2_square_distance(const float *src, const float *dst, int B, int N, int M, float *out,
                        void *stream);

/* ---- a2 index_points (:43-60) ------------------------------------------------
 * out[b,j,:] = points[b, idx[b,j], :].  points has element strides (sB,sN,sC),
 * idx is [B,J] int64 (J = product of the trailing index dims), out is [B,J,C]
 * contiguous fp32.  Indices outside [0,N) produce zeros. */
int pn2_index_points(const float *points, int64_t sB, int64_t sN, int64_t sC, int B, int N, int C,
                     const int64_t *idx, int64_t J, float *out, void *stream);
/* backward of a2: dpoints[b, idx[b,j], :] += dout[b,j,:]  (dpoints [B,N,C] contiguous, pre-zeroed) */
int pn2_index_points_bwd(const float *dout, const int64_t *idx, int B, int N, int C, int64_t J,
                         float *dpoints, void *stream);

/* ---- a3 farthest_point_sample (:63-84) ---------------------------------------
 * xyz [B,N,3] with element strides (sB,sN,sC); start_idx [B] int64 = the
 * torch.randint draw of :75 (made by the caller on the CPU generator);
 * out_idx [B,npoint] int64; out_xyz [B,npoint,3] fp32 (may be NULL) receives
 * index_points(xyz, out_idx) (:125).  Bit-exact: dist = (dx*dx+dy*dy)+dz*dz with
 * separately rounded operations, running minimum from 1e10f, first arg-max. */
int pn2_farthest_point_sample(const float *xyz, int64_t sB, int64_t sN, int64_t sC, int B, int N,
                              int npoint, const int64_t *start_idx, int64_t *out_idx,
                              float *out_xyz, void *stream);

/* ---- a4 query_ball_point (:87-107) -------------------------------------------
 * xyz [B,N,3] strided, new_xyz [B,S,3] strided; r2 = (float)((double)radius*radius);
 * out_idx [B,S,nsample] int64: ascending indices of the first nsample points with
 * !(d > r2) (d in the a1 rounding order), remaining slots = first hit, all slots
 * = N if there is no hit; out_cnt [B,S] int32 (may be NULL) = number of hits kept. */
int pn2_query_ball_point(const float *xyz, int64_t sB, int64_t sN, int64_t sC,
                         const float *new_xyz, int64_t qB, int64_t qN, int64_t qC, int B, int N,
                         int S, float r2, int nsample, int64_t *out_idx, int32_t *out_cnt,
                         void *stream);

/* The same result through a uniform cell grid (csrc/ballgrid.cu): the cloud is counting-sorted into cells of edge >= 1.01 r,
 * a query evaluates only the points of its 3 x 3 x 3 cells (~100 instead of N) and selects through an N-bit map, which
 * reproduces "first nsample in index order, padded with the first" exactly.  `radius` is the reference's Python float,
 * r2 = float32(radius ** 2) as above.  Clouds whose coordinates are so large that the fp32 rounding error of the
 * reference's distance could reach the cell margin, or that hold NaN / Inf, are scanned in index order inside the same
 * launch, so the output is always that of pn2_query_ball_point.  workspace: pn2_ball_grid_workspace_bytes(B, N) bytes,
 * 16-byte aligned; N <= 409600 (per-query bitmaps in shared memory). */
size_t pn2_ball_grid_workspace_bytes(int B, int N);
int pn2_query_ball_point_grid(const float *xyz, int64_t sB, int64_t sN, int64_t sC,
                              const float *new_xyz, int64_t qB, int64_t qN, int64_t qC, int B, int N,
                              int S, float radius, float r2, int nsample, int64_t *out_idx, int32_t *out_cnt,
                              void *workspace, size_t workspace_bytes, void *stream);

/* ---- a5 sample_and_group, gather half (:127-132) -----------------------------
 * rows[(b,s,k), 0:3]     = xyz[b, idx[b,s,k], :] - new_xyz[b,s,:]
 * rows[(b,s,k), 3:3+D]   = feats[b, idx[b,s,k], :]        (feats may be NULL, D = 0)
 * rows[(b,s,k), 3+D:ld]  = 0
 * feats [B,N,D] with element strides (fB,fN,fD); rows has dtype `dtype`, M = B*S*nsample. */
int pn2_group_points(const float *xyz, int64_t sB, int64_t sN, int64_t sC, const float *new_xyz,
                     const float *feats, int64_t fB, int64_t fN, int64_t fD, const int64_t *idx,
                     int B, int N, int S, int nsample, int D, void *rows, int ld, int dtype,
                     void *stream);
/* backward: dfeats[b, idx, :] += drows[(b,s,k), 3:3+D]  (dfeats [B,N,D] fp32 contiguous, pre-zeroed) */
int pn2_group_points_bwd(const void *drows, int ld, int dtype, const int64_t *idx, int B, int N,
                         int S, int nsample, int D, float *dfeats, void *stream);

/* ---- a7/a8 the 1x1-conv MLP on rows -------------------------------------------
 * One layer:  Z[M,N] = act(X)[M,K] * W[N,K]^T (+ bias)
 *   act(x)[m,k] = in_scale ? relu(x[m,k]*in_scale[k] + in_shift[k]) : x[m,k]
 *   (the previous layer's BatchNorm + ReLU, :198 / :314, applied while loading).
 * W is the Conv2d/Conv1d weight [N,K(,1,1)] fp32 contiguous.  If stat_accum is
 * non-NULL the kernel also adds the column sums of Z and Z^2 (bias excluded)
 * into it for the train-mode batch statistics: layout [PN2_STAT_REPLICAS][2][N] fp64, one fp64
 * atomicAdd per column and CTA (order-independent to ~1e-16 relative, so the
 * fp32 statistics derived from it are run-to-run stable).  The accumulator must
 * be ZERO on entry; pn2_bn_train_finalize consumes it and zeroes it again.
 * x_dtype/z_dtype: storage of X and Z rows.  With PN2_BF16 rows the product runs
 * on the tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM); with
 * PN2_F32 rows on the fp32 FMA pipes. */
/* wpack: device scratch of pn2_linear_wpack_bytes(K, N) bytes for the bf16, pre-swizzled copy
 * of W that the tensor-core path streams with TMA (bf16 rows with ld % 8 == 0; may be NULL,
 * which selects the FMA-pipe kernel, as does PN2_DISABLE_TC=1 in the environment). */
size_t pn2_linear_wpack_bytes(int K, int N);
int pn2_linear_fwd(const void *X, int ldx, int x_dtype, const float *in_scale,
                   const float *in_shift, const float *W, const float *bias, int64_t M, int K,
                   int N, void *Z, int ldz, int z_dtype, double *stat_accum, void *wpack,
                   void *stream);
/* One launch packs the bf16 images of n weights (every layer of an MLP, both orientations) ahead of the layer calls:
 * image i is what pn2_linear_fwd streams for W_i [N_i, K_i] (transposed_host[i] == 0, pn2_linear_wpack_bytes(K_i, N_i)
 * bytes) or what pn2_linear_bwd_data streams (transposed_host[i] != 0, pn2_linear_wpack_bytes(N_i, K_i) bytes).
 * All five arrays are HOST arrays of n entries (W_host / wpack_host hold device pointers). */
int pn2_pack_weights(int n, const float *const *W_host, const int *K_host, const int *N_host,
                     const int *transposed_host, void *const *wpack_host, void *stream);
/* Train-mode BatchNorm finalize fused into the layer kernel: the CTA that draws the last ticket does what
 * pn2_bn_train_finalize does (same fields) and leaves stat_accum and *ticket zeroed.  ticket: a zero-initialised
 * uint32 in device memory, reusable by consecutive calls on one stream.  momentum_dev (may be NULL): a float in DEVICE
 * memory read at run time instead of `momentum`, so that a captured CUDA graph follows the reference's per-epoch
 * BatchNorm-momentum schedule (localfunctions.py:191-195) without re-capture. */
typedef struct pn2_bn_finalize {
    uint32_t *ticket;
    const float *gamma, *beta, *conv_bias;
    float eps, momentum;
    float *running_mean, *running_var, *scale, *shift, *save_mean, *save_invstd;
    int64_t *num_batches_tracked;
    const float *momentum_dev;
} pn2_bn_finalize;
/* pn2_linear_fwd with wpack ALREADY holding the image (pn2_pack_weights) and, when fin_host (a HOST struct of device
 * pointers) is non-NULL, the BatchNorm finalize of the layer's statistics fused in (stat_accum required). */
int pn2_linear_fwd_prepacked(const void *X, int ldx, int x_dtype, const float *in_scale,
                             const float *in_shift, const float *W, const float *bias, int64_t M, int K,
                             int N, void *Z, int ldz, int z_dtype, double *stat_accum, const void *wpack,
                             const pn2_bn_finalize *fin_host, void *stream);
/* pn2_linear_bwd_data with wpack already holding the transposed image (pn2_pack_weights, transposed = 1) */
int pn2_linear_bwd_data_prepacked(const void *dZ, int lddz, int dz_dtype, const float *W, int64_t M, int K,
                                  int N, void *dX, int lddx, int dx_dtype, const void *wpack, void *stream);
/* dX[M,K] = dZ[M,N] * W[N,K]   (no activation handling; see pn2_bn_relu_bwd_*);
 * wpack: scratch of pn2_linear_wpack_bytes(N, K) bytes (note the swapped roles) or NULL */
int pn2_linear_bwd_data(const void *dZ, int lddz, int dz_dtype, const float *W, int64_t M, int K,
                        int N, void *dX, int lddx, int dx_dtype, void *wpack, void *stream);
/* dW[N,K] = sum_m dZ[m,n] * act(X)[m,k];  scratch holds pn2_linear_wgrad_scratch_bytes() bytes */
size_t pn2_linear_wgrad_scratch_bytes(int64_t M, int K, int N);
int pn2_linear_bwd_weight(const void *dZ, int lddz, int dz_dtype, const void *X, int ldx,
                          int x_dtype, const float *in_scale, const float *in_shift, int64_t M,
                          int K, int N, float *dW, void *scratch, void *stream);
/* dW[N,K] += sum_m dZ[m,n] * act(X)[m,k]  (bf16 rows only): no scratch, no reduce launch -- every CTA adds its fp32
 * block into dW with L2 reductions (red.global.add), so the caller zeroes dW first (the trainer's flat gradient
 * buffer is zeroed once per step) and the summation order, hence the last fp32 bit, varies run to run -- like the
 * reference's own cuDNN weight gradients (pointnet2_utils.py:198 backward); pn2_linear_bwd_weight is the fixed-order one */
int pn2_linear_bwd_weight_accum(const void *dZ, int lddz, int dz_dtype, const void *X, int ldx,
                                int x_dtype, const float *in_scale, const float *in_shift, int64_t M,
                                int K, int N, float *dW, void *stream);

/* ---- one backward step of an MLP layer, fused (bf16 rows, tcgen05) -----------------------------------------------
 * Autograd through relu(bn_l(conv_l(act_{l-1}))) (:196-198, :311-314) for layer l in ONE launch instead of the
 * reduce / dz / data-gradient / weight-gradient launches above:
 *     dZ_l      = BatchNorm+ReLU backward of layer l applied to dA_l while the tile is in shared memory (never stored)
 *     dX        = dZ_l . W_l, masked by layer l-1's ReLU when prev_* are given (so its consumer passes da_mode 1)
 *     dW       += dZ_l^T . act(X),  act = relu(prev_scale . X + prev_shift) or X itself (the MLP's input rows)
 *     dgamma_prev, dbeta_prev of layer l-1's BatchNorm from dX (when prev_* are given; "last CTA finalizes")
 * da_mode: 0 = dA dense, layer l's ReLU mask applied here; 1 = dA already masked by the call that produced it;
 *          2 = pooled: the gradient is dOut [G, N] fp32 of the max over nsample = 32 with the arg-max map `arg` [G, N] int32
 *              of pn2_bn_relu_max (row (g, k) receives dOut[g, c] where arg[g, c] == k; dA unused, M = 32 G);
 *          3 = `dA` IS dZ_l (Z / scale / ... unused).  mean == NULL: frozen statistics (dZ = scale . mask . dA).
 * dW is ADDED to (zero it first; L2 reductions, order varies run to run); K % 4 != 0 needs `scratch`
 * (pn2_mlp_bwd_layer_scratch_bytes) for a fixed-order partial sum instead.  dX / dW may be NULL (not wanted).
 * HOST struct of device pointers; pn2_mlp_bwd_layer_supported tells whether the kernel takes a layer
 * (N, ldx <= 128, row pitches % 8 == 0, shared / tensor memory) -- otherwise use the per-step entry points. */
typedef struct pn2_bwd_layer {
    const void *dA; int ldda; int da_mode;
    const float *dOut; const int32_t *arg; int nsample;
    const void *Z; int ldz;
    const float *scale, *shift, *mean, *invstd, *dgamma, *dbeta;
    const void *wpack_t;
    const void *X; int ldx;
    const float *prev_scale, *prev_shift, *prev_mean, *prev_invstd;
    void *dX; int lddx;
    float *dW; void *scratch;
    double *stat_accum; uint32_t *ticket; float *dgamma_prev, *dbeta_prev;
    int64_t M; int K, N;
} pn2_bwd_layer;
int pn2_mlp_bwd_layer_supported(int64_t M, int K, int N, int ldx, int lddx, int da_mode, int has_prev, int want_dx, int want_dw);
size_t pn2_mlp_bwd_layer_scratch_bytes(int64_t M, int K, int N);
int pn2_mlp_bwd_layer(const pn2_bwd_layer *layer_host, void *stream);

/* ---- BatchNorm (train statistics / eval fold) ---------------------------------
 * Train (:198 with module.training): turns the [PN2_STAT_REPLICAS][2][N] fp64 sums of pn2_linear_fwd
 * into mean / biased variance over M rows (and zeroes the accumulator), writes
 *   scale = gamma*invstd, shift = beta - mean*scale, save_mean, save_invstd
 * and updates running_mean/var in place with `momentum` (running_mean includes
 * the conv bias that the GEMM left out; running_var uses the unbiased variance) and, when
 * num_batches_tracked is non-NULL, increments that int64 counter (nn.BatchNorm's buffer). */
int pn2_bn_train_finalize(double *stat_accum, int64_t M, int N,
                          const float *gamma, const float *beta, const float *conv_bias,
                          float eps, float momentum, float *running_mean, float *running_var,
                          float *scale, float *shift, float *save_mean, float *save_invstd,
                          int64_t *num_batches_tracked, void *stream);
/* Eval: scale = gamma/sqrt(running_var+eps), shift = beta - running_mean*scale (the GEMM adds the bias). */
int pn2_bn_eval_fold(const float *gamma, const float *beta, const float *running_mean,
                     const float *running_var, float eps, int N, float *scale, float *shift,
                     void *stream);

/* ---- a7 tail: BN + ReLU + max over nsample (:198-200) --------------------------
 * out[g,c] = max_k relu(Z[(g,k),c]*scale[c]+shift[c]); arg[g,c] = first maximal k (int32).
 * G groups of nsample consecutive rows; out fp32 [G,C] contiguous. */
int pn2_bn_relu_max(const void *Z, int ldz, int z_dtype, const float *scale, const float *shift,
                    int64_t G, int nsample, int C, float *out, int32_t *arg, void *stream);
/* ... and zmax[g,c] = Z[(g, arg[g,c]), c] (rows of Z's dtype, leading dimension ldzm): with it the pooled backward's
 * BatchNorm gradients are a plain [G, C] column reduction -- pn2_bn_relu_bwd_reduce(dOut fp32, zmax) -- instead of a gather
 * of one element per (group, channel) out of the [G * nsample, C] rows (5 % of HBM bandwidth, latency bound). */
int pn2_bn_relu_max_keep(const void *Z, int ldz, int z_dtype, const float *scale, const float *shift,
                         int64_t G, int nsample, int C, float *out, int32_t *arg, void *zmax, int ldzm,
                         void *stream);
/* a8 tail: out[m,c] = relu(Z[m,c]*scale[c]+shift[c])  (fp32 [M,C] contiguous) */
int pn2_bn_relu(const void *Z, int ldz, int z_dtype, const float *scale, const float *shift,
                int64_t M, int C, float *out, void *stream);

/* ---- a7 inference: one whole set-abstraction level in one kernel (:127-132, :196-200) ------------
 * out[b,s,:] = max_k MLP([xyz[b,idx[b,s,k]] - new_xyz[b,s] | feats[b,idx[b,s,k]]]) with L layers of
 * relu(bn_eval(conv1x1)) -- the ball-query gather, the L tensor-core GEMMs (bf16 operands, fp32
 * accumulation in tensor memory), eval-mode BatchNorm, ReLU and the max over nsample fused; grouped
 * rows and activations never reach HBM.  Requirements: nsample == 32, L <= 4, every hidden width a
 * multiple of 16, every width <= 512 (PN2_ERR_UNSUPPORTED otherwise: use the per-layer entry points).
 * The *_host arrays are HOST arrays of L entries: widths (out channels), and DEVICE pointers to
 * W_l [N_l, K_l] fp32 (K_0 = 3 + D with the reference's [xyz | feats] column order, K_l = N_{l-1}),
 * conv bias (entries may be NULL), and the eval-mode scale / shift of pn2_bn_eval_fold.
 * xyz [B,N,3] strided; new_xyz [B,S,3] contiguous; feats [B,N,D] with row strides (fB,fN), element
 * stride 1 (NULL when D = 0); idx [B,S,32] int64 (out-of-range entries gather zeros);
 * out [B,S,N_{L-1}] fp32.  workspace: pn2_sa_fused_eval_workspace_bytes() bytes (0 = unsupported).  W_host == NULL:
 * the workspace still holds the weight images packed by a previous call for the same weights (inference with frozen
 * parameters: no pack launches). */
size_t pn2_sa_fused_eval_workspace_bytes(int D, int L, const int *widths_host);
int pn2_sa_fused_eval(const float *xyz, int64_t sB, int64_t sN, int64_t sC, const float *new_xyz,
                      const float *feats, int64_t fB, int64_t fN, const int64_t *idx, int B, int N,
                      int S, int nsample, int D, int L, const int *widths_host,
                      const float *const *W_host, const float *const *bias_host,
                      const float *const *scale_host, const float *const *shift_host, float *out,
                      void *workspace, void *stream);

/* ---- backward of BN(train)+ReLU -----------------------------------------------
 * With g = dA * [bn(z) > 0]:  dbeta = sum g, dgamma = sum g*zhat,
 *   dz = gamma*invstd * (g - dbeta/M - zhat*dgamma/M)          (train)
 *   dz = scale * g                                              (eval: pass save_mean = NULL)
 * Two passes: *_reduce adds (sum g, sum g*zhat) into a zeroed [PN2_STAT_REPLICAS][2][C] fp64 accumulator (same
 * contract as stat_accum above); pn2_bn_bwd_finalize turns it into dbeta/dgamma and zeroes it;
 * *_dz writes dZ (dZ may alias dA: the update is element-wise).  The "pool" variants take the pooled gradient dOut[G,C] and the
 * arg-max map of pn2_bn_relu_max instead of a dense dA (g is non-zero only on the
 * arg-max row of each (group, channel)). */
int pn2_bn_relu_bwd_reduce(const void *dA, int ldda, int da_dtype, const void *Z, int ldz,
                           int z_dtype, const float *scale, const float *shift,
                           const float *save_mean, const float *save_invstd, int64_t M, int C,
                           double *accum, void *stream);
int pn2_pool_bn_relu_bwd_reduce(const float *dOut, const int32_t *arg, const void *Z, int ldz,
                                int z_dtype, const float *scale, const float *shift,
                                const float *save_mean, const float *save_invstd, int64_t G,
                                int nsample, int C, double *accum, void *stream);
int pn2_bn_bwd_finalize(double *accum, int C, float *dgamma, float *dbeta, void *stream);
/* The two reductions with pn2_bn_bwd_finalize fused in: the block that draws the last ticket (a zero-initialised
 * uint32 in device memory, left zero again) writes dgamma / dbeta and zeroes the accumulator. */
int pn2_bn_relu_bwd_reduce_finalize(const void *dA, int ldda, int da_dtype, const void *Z, int ldz,
                                    int z_dtype, const float *scale, const float *shift,
                                    const float *save_mean, const float *save_invstd, int64_t M, int C,
                                    double *accum, uint32_t *ticket, float *dgamma, float *dbeta,
                                    void *stream);
int pn2_pool_bn_relu_bwd_reduce_finalize(const float *dOut, const int32_t *arg, const void *Z, int ldz,
                                         int z_dtype, const float *scale, const float *shift,
                                         const float *save_mean, const float *save_invstd, int64_t G,
                                         int nsample, int C, double *accum, uint32_t *ticket,
                                         float *dgamma, float *dbeta, void *stream);
int pn2_bn_relu_bwd_dz(const void *dA, int ldda, int da_dtype, const void *Z, int ldz, int z_dtype,
                       const float *scale, const float *shift, const float *save_mean,
                       const float *save_invstd, const float *dgamma, const float *dbeta, int64_t M,
                       int C, void *dZ, int lddz, int dz_dtype, void *stream);
int pn2_pool_bn_relu_bwd_dz(const float *dOut, const int32_t *arg, const void *Z, int ldz,
                            int z_dtype, const float *scale, const float *shift,
                            const float *save_mean, const float *save_invstd, const float *dgamma,
                            const float *dbeta, int64_t G, int nsample, int C, void *dZ, int lddz,
                            int dz_dtype, void *stream);

/* ---- a8 head: three nearest neighbours + inverse-distance interpolation (:296-307)
 * xyz1 [B,N,3] strided (fine), xyz2 [B,S,3] strided (coarse); idx3 [B,N,3] int64 and
 * w3 [B,N,3] fp32 follow a stable ascending sort of the a1-order distances;
 * K3 = min(3,S) valid columns (the rest are idx 0 / weight 0). */
int pn2_three_nn(const float *xyz1, int64_t aB, int64_t aN, int64_t aC, const float *xyz2,
                 int64_t cB, int64_t cN, int64_t cC, int B, int N, int S, int64_t *idx3, float *w3,
                 void *stream);
/* The same result through a cell grid over the coarse cloud (csrc/ballgrid.cu: the ball query's build kernel with cells
 * sized from the cloud's density, one thread per fine point over the 27 cells around it).  A query whose third neighbour
 * is not provably inside its 27 cells -- and every query of a cloud the grid refuses (NaN/Inf, coordinates so large that
 * fp32 rounding of the a1-order distance reaches the cell size) -- takes the full index-order scan inside the same launch;
 * fallback_count (may be NULL) is incremented once per such query.  workspace: pn2_ball_grid_workspace_bytes(B, S) bytes,
 * 16-byte aligned. */
int pn2_three_nn_grid(const float *xyz1, int64_t aB, int64_t aN, int64_t aC, const float *xyz2,
                      int64_t cB, int64_t cN, int64_t cC, int B, int N, int S, int64_t *idx3, float *w3,
                      void *workspace, size_t workspace_bytes, unsigned *fallback_count, void *stream);
/* rows[(b,n), 0:D1]      = points1[b,n,:]                 (may be NULL, D1 = 0)
 * rows[(b,n), D1:D1+D2]  = (p2[i0]*w0 + p2[i1]*w1) + p2[i2]*w2   (products rounded, :303)
 * rows[(b,n), D1+D2:ld]  = 0.   points1 [B,N,D1] strides (pB,pN,pD); points2 [B,S,D2] strides (qB,qN,qD). */
int pn2_interp_concat(const float *points1, int64_t pB, int64_t pN, int64_t pD,
                      const float *points2, int64_t qB, int64_t qN, int64_t qD, const int64_t *idx3,
                      const float *w3, int B, int N, int S, int D1, int D2, void *rows, int ld,
                      int dtype, void *stream);
/* dpoints2[b, idx3[b,n,k], :] += w3[b,n,k] * drows[(b,n), D1:D1+D2]   (dpoints2 [B,S,D2] fp32, pre-zeroed) */
int pn2_interp_bwd(const void *drows, int ld, int dtype, const int64_t *idx3, const float *w3, int B,
                   int N, int S, int D1, int D2, float *dpoints2, void *stream);

/* ---- segmentation head tail on rows (models/pointnet2_sem_seg.py:36-39) -----------
 * conv1 + bn1 run as one more pn2_linear_fwd layer; these two calls are what follows its pre-BatchNorm
 * product Z [M, C] (bf16 rows, C % 32 == 0, C <= 256):
 *   forward : a = dropout(relu(Z*scale + shift)); logp[M, NC] = log_softmax(a.W2^T + b2)   (NC <= 32 classes,
 *             W2 [NC, C] fp32, b2 may be NULL); act_out (bf16 rows [M, ldo], may be NULL) receives a, which
 *             conv2's weight gradient needs; labels (int64 [M], may be NULL) receives argmax_c logp[m, c] (first
 *             maximum, as torch.argmax -- localfunctions.py:400 takes it with numpy on the host).
 *   backward: dlogits = dlogp - exp(logp)*rowsum(dlogp);  dA[M, C] (bf16) = (dlogits.W2) * keep/(1-p), i.e. the
 *             gradient w.r.t. relu(bn1(.)) that pn2_bn_relu_bwd_* take; db2[NC] = column sums of dlogits (through
 *             the zeroed fp64 accumulator db2_accum[>= 32], left zero); dlogits_rows (bf16 [M, lddl], lddl % 8 == 0,
 *             may be NULL) receives dlogits for pn2_linear_bwd_weight (conv2's weight gradient).
 * Dropout keeps element (m, k) when hash(*seed, m*C + k) >= p (p quantised to 1/256; kept values scaled by
 * 1/(1-p)); *seed is an int64 in DEVICE memory read at run time; drop_p = 0 (eval mode) ignores it.
 * C % 32 == 0 (the contraction runs on mma.sync m16n8k16 with bf16 hi+lo split operands, fp32-level accuracy).
 *
 * pn2_head_tail_loss_fwd / _bwd: the same tail FUSED with the loss of pointnet2_sem_seg.py:47-48,
 * F.nll_loss(pred, target, weight) = -(sum_m w[t_m] logp[m, t_m]) / (sum_m w[t_m]):
 *   forward : additionally reads target[M] (int64; values outside [0, NC), e.g. F.nll_loss's ignore_index -100,
 *             contribute nothing) and class_weight[NC] (NULL = ones); loss_accum[2] is a zeroed fp64 scratch (left
 *             zero); loss_out[0] = the loss, loss_out[1] = sum of the target weights (backward reads it).
 *   backward: dlogits = (*dloss) * w[t]/loss_out[1] * (exp(logp) - onehot(t)) straight from the targets -- the dense
 *             [M, NC] gradient tensor of pn2_head_tail_bwd does not exist; dloss (DEVICE, NULL = 1) is dL/dloss. */
int pn2_head_tail_fwd(const void *Z, int ldz, const float *scale, const float *shift, const float *W2,
                      const float *b2, int64_t M, int C, int NC, float drop_p, const int64_t *seed,
                      float *logp, void *act_out, int ldo, int64_t *labels, void *stream);
int pn2_head_tail_bwd(const float *dlogp, const float *logp, const float *W2, int64_t M, int C, int NC,
                      float drop_p, const int64_t *seed, void *dA, int ldda, void *dlogits_rows, int lddl,
                      double *db2_accum, float *db2, void *stream);
int pn2_head_tail_loss_fwd(const void *Z, int ldz, const float *scale, const float *shift, const float *W2,
                           const float *b2, int64_t M, int C, int NC, float drop_p, const int64_t *seed,
                           const int64_t *target, const float *class_weight, float *logp, void *act_out, int ldo,
                           double *loss_accum, float *loss_out, void *stream);
int pn2_head_tail_loss_bwd(const float *logp, const int64_t *target, const float *class_weight, const float *loss_out,
                           const float *dloss, const float *W2, int64_t M, int C, int NC, float drop_p,
                           const int64_t *seed, void *dA, int ldda, void *dlogits_rows, int lddl, double *db2_accum,
                           float *db2, void *stream);

/* ---- optimizer step of the training loop (sem_seg_training.py:576-582: torch.optim.Adam with L2 weight decay) ----
 * One launch over the flat gradient buffer.  params[T]: device table of the parameter tensors' fp32 pointers;
 * tensor_off[T] / tensor_n[T]: each tensor's first element in the three flat buffers (a multiple of 4) and its size;
 * chunks[n_chunks]: (tensor, first element) pairs as int32x2, each covering <= chunk_elems (a multiple of 4)
 * elements -- one CTA per chunk.  hyper[5] = {lr, beta1, beta2, eps, weight_decay} (fp64, DEVICE memory, read at run
 * time); *step: fp32 count of finished steps, incremented by the call; *ticket: a zeroed uint32, left zero.
 * Arithmetic of torch.optim.Adam (amsgrad / maximize off):  g += wd*p;  m += (1-b1)(g-m);  v = b2 v + (1-b2) g^2;
 * p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps). */
int pn2_adam_step(void *const *params, const int64_t *tensor_off, const int64_t *tensor_n, const void *chunks,
                  int n_chunks, int chunk_elems, const float *grad_flat, float *exp_avg_flat, float *exp_avg_sq_flat,
                  const double *hyper, float *step, void *ticket, void *stream);

/* ---- layout helpers ------------------------------------------------------------
 * dst[r, c] (fp32, leading dim ldd) = src[r*sR + c*sC] for r < R, c < C: turns a
 * channel-major [C,R] slab into point-major rows (batched over B with strides). */
int pn2_to_rows(const float *src, int64_t sB, int64_t sR, int64_t sC, int B, int64_t R, int C,
                float *dst, int64_t dB, int ldd, void *stream);
/* dst fp32 [M,C] contiguous = rows (any dtype) [M, ld] columns c0..c0+C */
int pn2_rows_to_f32(const void *rows, int ld, int dtype, int64_t M, int c0, int C, float *dst,
                    void *stream);

/* ---- test-time vote accumulation and final arg-max -------------------------
 * replaces /root/reference/localfunctions.py:336-343 (add_vote, a Python loop over B x N pairs) and :405
 * (np.argmax(vote_label_pool, 1)).  For every pair i < count with weight[i] != 0 and not inf (weight NULL = all
 * pairs; fp32, or fp64 when weight_is_f64): votes[point_idx[i], pred_label[i]] += 1.  votes is the [P, NC] int32
 * pool (the reference keeps float64 counts -- same integers), accumulated across calls; pairs whose index or label is
 * out of range (an IndexError in the reference) are skipped and counted into *skipped (device, may be NULL). */
int pn2_add_vote(const int64_t *point_idx, const int64_t *pred_label, const void *weight, int weight_is_f64,
                 int64_t count, int64_t P, int NC, int32_t *votes, unsigned long long *skipped, void *stream);
/* labels[p] = first class with the largest count (np.argmax); labels is int64 [P], or uint8 [P] when labels_are_u8 */
int pn2_vote_argmax(const int32_t *votes, int64_t P, int NC, void *labels, int labels_are_u8, void *stream);

/* ---- z-rotation augmentation of the training batch, in place in HBM ------------------------
 * replaces provider.rotate_point_cloud_z (/root/reference/provider.py:66-84) applied to points[:, :, :3] at
 * localfunctions.py:205: cloud b is rotated about z by the angle whose (cos, sin) is cos_sin[2b], cos_sin[2b+1]
 * (float64, device), products and sums in float64 in the reference's np.dot order, result rounded to float32.
 * points: the xyz channels of a [B, N, C] fp32 batch, element strides (sB, sN, sC). */
int pn2_rotate_z(float *points, int64_t sB, int64_t sN, int64_t sC, const double *cos_sin, int B, int N, void *stream);

/* ---- sliding-block slicer of the test-time dataset -----------------------------------------
 * replaces TestCustomDataset.__getitem__ (/root/reference/sem_seg_testing.py:182-254).  The grid of cells (index_y outer,
 * index_x inner, :193-194) is described per column / row by lo = s - padding, hi = e + padding (the reference's s_x/e_x
 * expressions, :196-201, evaluated on the host in float64); cell = iy * gx + ix.
 * pn2_slice_cells: for every point p (float64 xyz, element strides sP / sC) and every cell with lo_x <= x <= hi_x and
 *   lo_y <= y <= hi_y (:202): k = counts[cell]++ and, in the fill pass (cell_offset / slot_point / slot_cell non-NULL),
 *   slot_point[cell_offset[cell] + k] = p.  Count pass first, then zero `counts` and run the fill pass.
 * pn2_slice_pad: padding slot i (pad_cell[i], rank pad_rank[i] among the cell's padding slots) takes the pad_rank-th
 *   member when the cell needs no more padding than it has members (np.random.choice(..., replace=False), :207-208; the
 *   caller has randomly permuted the members), else member rnd[i] % n (replace=True).
 * pn2_slice_rows: rows[s] = [x - cx, y - cy, z, x/max_x, y/max_y, z/max_z, extra_e / extra_div_e ...] of slot s's point in
 *   float64, rounded once to float32 (:216-241, localfunctions.py:394); label and labelweights[label] (:223-224).
 *   gx > 0: slot_cell is a grid cell, centre (cx[cell % gx], cy[cell / gx]); gx == 0: one centre per cell, (cx[cell], cy[cell]). */
int pn2_slice_cells(const double *points, int64_t sP, int64_t sC, int64_t P, const double *lo_x, const double *hi_x,
                    int gx, const double *lo_y, const double *hi_y, int gy, double min_x, double min_y, double stride,
                    double block_size, double padding, int32_t *counts, const int64_t *cell_offset,
                    int64_t *slot_point, int32_t *slot_cell, void *stream);
int pn2_slice_pad(const int32_t *counts, const int64_t *cell_offset, const int32_t *pad_cell, const int64_t *pad_rank,
                  const int64_t *rnd, int64_t n_pad, int block_points, int64_t *slot_point, int32_t *slot_cell,
                  void *stream);
int pn2_slice_rows(const double *points, int64_t sP, int64_t sC, const int64_t *labels, const double *extra, int64_t eE,
                   int64_t eP, const double *extra_div, int E, const float *labelweights, const int64_t *slot_point,
                   const int32_t *slot_cell, const double *cx, const double *cy, int gx, double max_x, double max_y,
                   double max_z, int64_t S, float *rows, int64_t *out_label, float *out_weight, void *stream);

/* ---- SURVEY 8(f) n3, training half: the random crops of TrainCustomDataset.__getitem__ ---------------------
 * (/root/reference/sem_seg_training.py:200-259).  boxes_host: HOST array [K][4] of float64 {lo_x, hi_x, lo_y, hi_y} =
 * centre -+ block_size / 2 (:208-209), K <= 64; every point with lo_x <= x <= hi_x and lo_y <= y <= hi_y (:210-213) is
 * a member of box k.  Count pass (crop_offset / slot_* NULL): counts[k] += members; fill pass (counts zeroed again):
 * slot_point[crop_offset[k] + i] = point, slot_crop[...] = crop_id0 + k (counts / crop_offset point at box 0 of this call).
 * Rows of the selected points: pn2_slice_rows with gx = 0
 * (cx / cy indexed by crop). */
int pn2_crop_members(const double *points, int64_t sP, int64_t sC, int64_t P, const double *boxes_host, int K,
                     int crop_id0, int32_t *counts, const int64_t *crop_offset, int64_t *slot_point,
                     int32_t *slot_crop, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PN2B200_H */
