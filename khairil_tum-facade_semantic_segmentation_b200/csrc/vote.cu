// vote.cu -- SURVEY.md 8(f) n1: the test-time vote accumulation of /root/reference/localfunctions.py:336-343 (add_vote)
// and the final arg-max (:405), on the device.
//
// The reference walks B x N (block slot, point) pairs in a Python double loop and bumps vote_label_pool[point, label]
// for every pair whose sample weight is neither 0 nor inf (~35 M iterations for a 10 M-point facade, minutes of host
// time per scene); the predicted labels make a device->host round trip per batch for it.  Here one thread owns one pair
// and does one integer atomic into the [P, NC] pool that stays in HBM; counts are exact, so the pool and the labels are
// bit-identical to the reference's whatever the order of the additions.
#include "common.cuh"

namespace pn2 {

template <typename TW>
__global__ void __launch_bounds__(256)
add_vote_kernel(const int64_t *__restrict__ point_idx, const int64_t *__restrict__ pred_label, const TW *__restrict__ weight,
                int64_t count, int64_t P, int NC, int32_t *__restrict__ votes, unsigned long long *__restrict__ skipped) {
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        if (weight) {
            const TW w = weight[i];
            if (!(w != (TW)0) || isinf(w)) continue;        // :341  `weight != 0 and not np.isinf(weight)` (NaN counts)
        }
        const int64_t p = point_idx[i], c = pred_label[i];
        if (p < 0 || p >= P || c < 0 || c >= NC) {          // the reference raises IndexError here (or wraps a negative index)
            ++bad;
            continue;
        }
        atomicAdd(votes + p * NC + c, 1);
    }
    if (skipped && bad) atomicAdd(skipped, bad);
}

// labels[p] = first class with the largest count (np.argmax, :405); NC <= 32 classes per lane-group would be overkill:
// a row is NC * 4 bytes (72 for the TUM-Facade label set), one thread reads it
template <typename TL>
__global__ void __launch_bounds__(256)
vote_argmax_kernel(const int32_t *__restrict__ votes, int64_t P, int NC, TL *__restrict__ labels) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t *row = votes + p * NC;
        int32_t best = row[0];
        int bc = 0;
        for (int c = 1; c < NC; ++c) {
            const int32_t v = row[c];
            if (v > best) { best = v; bc = c; }
        }
        labels[p] = (TL)bc;
    }
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_add_vote(const int64_t *point_idx, const int64_t *pred_label, const void *weight, int weight_is_f64,
                            int64_t count, int64_t P, int NC, int32_t *votes, unsigned long long *skipped, void *stream) {
    PN2_REQUIRE(count >= 0 && P >= 0 && NC >= 1, "add_vote: bad sizes count=%lld P=%lld NC=%d", (long long)count, (long long)P, NC);
    if (count == 0) return PN2_OK;
    PN2_REQUIRE(point_idx && pred_label && votes, "add_vote: null pointer");
    const int grid = grid_for(count, 256);
    if (weight_is_f64)
        add_vote_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(point_idx, pred_label, (const double *)weight, count, P,
                                                                       NC, votes, skipped);
    else
        add_vote_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(point_idx, pred_label, (const float *)weight, count, P,
                                                                      NC, votes, skipped);
    count_launch();
    return check_launch("add_vote");
}

extern "C" int pn2_vote_argmax(const int32_t *votes, int64_t P, int NC, void *labels, int labels_are_u8, void *stream) {
    PN2_REQUIRE(P >= 0 && NC >= 1 && (!labels_are_u8 || NC <= 256), "vote_argmax: bad sizes P=%lld NC=%d", (long long)P, NC);
    if (P == 0) return PN2_OK;
    PN2_REQUIRE(votes && labels, "vote_argmax: null pointer");
    const int grid = grid_for(P, 256);
    if (labels_are_u8)
        vote_argmax_kernel<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(votes, P, NC, (uint8_t *)labels);
    else
        vote_argmax_kernel<int64_t><<<grid, 256, 0, (cudaStream_t)stream>>>(votes, P, NC, (int64_t *)labels);
    count_launch();
    return check_launch("vote_argmax");
}

// ---- the training loop's z-rotation augmentation (provider.rotate_point_cloud_z, /root/reference/provider.py:66-84,
// called on points[:, :, :3] at localfunctions.py:205) on the batch that is already in HBM.  The reference multiplies the
// float32 coordinates with a float64 rotation matrix (np.dot -> float64) and stores the result into a float32 array:
//   x' = fp32(x*c + y*(-s) + z*0),  y' = fp32(x*s + y*c + z*0),  z' = fp32(x*0 + y*0 + z*1)
// evaluated here in fp64 in that order (no FMA contraction), per cloud b with (c, s) = cs[b].
namespace pn2 {
__global__ void __launch_bounds__(256)
rotate_z_kernel(float *__restrict__ points, int64_t sB, int64_t sN, int64_t sC, const double *__restrict__ cs, int B, int N) {
    const int64_t total = (int64_t)B * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t b;
        int n;
        fast_divmod(i, N, b, n);
        float *p = points + b * sB + (int64_t)n * sN;
        const double c = cs[2 * b], s = cs[2 * b + 1];
        const double x = (double)p[0], y = (double)p[sC], z = (double)p[2 * sC];
        const double xr = __dadd_rn(__dadd_rn(__dmul_rn(x, c), __dmul_rn(y, -s)), __dmul_rn(z, 0.0));
        const double yr = __dadd_rn(__dadd_rn(__dmul_rn(x, s), __dmul_rn(y, c)), __dmul_rn(z, 0.0));
        const double zr = __dadd_rn(__dadd_rn(__dmul_rn(x, 0.0), __dmul_rn(y, 0.0)), z);
        p[0] = (float)xr;
        p[sC] = (float)yr;
        p[2 * sC] = (float)zr;
    }
}
}  // namespace pn2

extern "C" int pn2_rotate_z(float *points, int64_t sB, int64_t sN, int64_t sC, const double *cos_sin, int B, int N, void *stream) {
    PN2_REQUIRE(B >= 0 && N >= 0, "rotate_z: bad sizes B=%d N=%d", B, N);
    if (B == 0 || N == 0) return PN2_OK;
    PN2_REQUIRE(points && cos_sin, "rotate_z: null pointer");
    rotate_z_kernel<<<grid_for((int64_t)B * N, 256), 256, 0, (cudaStream_t)stream>>>(points, sB, sN, sC, cos_sin, B, N);
    count_launch();
    return check_launch("rotate_z");
}
