// slicer.cu -- SURVEY.md 8(f) n3: the sliding-block slicer of the test-time dataset
// (/root/reference/sem_seg_testing.py:182-254, TestCustomDataset.__getitem__) on the device.
//
// The reference walks the (index_y, index_x) grid of block_size x block_size cells at `stride` spacing and, PER CELL, scans
// the whole scene with np.where (O(cells x points): minutes for a 10 M-point facade), pads the cell's point list to a
// multiple of block_points with np.random.choice, shuffles it and builds the feature rows in float64.  Here:
//   slice_cells_kernel<COUNT>  one thread per point finds the (few) cells that contain it with EXACTLY the reference's
//                              float64 comparisons (x >= s_x - padding, x <= e_x + padding, ...; the per-column / per-row
//                              bounds are computed on the host with the reference's own expressions) and counts them;
//   slice_cells_kernel<FILL>   the same walk appends the point to its cells' member lists;
//   slice_pad_kernel           fills a cell's padding slots from its (already randomly permuted) members: the first `extra`
//                              of the permutation when extra <= n (np.random.choice(..., replace=False)), random members
//                              otherwise (replace=True) -- the reference's rule at :207;
//   slice_rows_kernel          builds the rows [x - cx, y - cy, z, x/max_x, y/max_y, z/max_z, extras (/255 for colours)]
//                              in float64 in the reference's operation order and rounds once to float32 (what
//                              torch.Tensor(batch_data) does at localfunctions.py:394), plus labels and sample weights.
// Random permutations / sorting / prefix sums are host-side torch calls (ops.slice_scene); which points pad a cell and the
// order inside a cell are random in the reference too (numpy's global generator), so parity is: identical cells in identical
// order, identical member sets and block counts, bit-identical rows for every (cell, point) pair.
#include "common.cuh"

namespace pn2 {

struct SliceGrid {
    const double *lo_x, *hi_x, *lo_y, *hi_y;   // [gx] / [gy]: s - padding and e + padding of every column / row of cells
    int gx, gy;
    double min_x, min_y, stride, reach;        // reach = block_size + padding: a cell starting more than this below x misses it
};

// first column (row) that can contain coordinate v: everything before it ends (even un-clamped, with 2 cells of slack for
// rounding) below v; columns are then walked upwards while their lower bound is <= v (lower bounds never decrease)
__device__ __forceinline__ int first_candidate(double v, double vmin, double stride, double reach, int g) {
    const double t = floor((v - vmin - reach) / stride) - 2.0;
    return t < 0.0 ? 0 : (t >= (double)g ? g : (int)t);
}

template <bool FILL>
__global__ void __launch_bounds__(256)
slice_cells_kernel(const double *__restrict__ pts, int64_t sP, int64_t sC, int64_t P, SliceGrid g, int32_t *__restrict__ counts,
                   const int64_t *__restrict__ cell_offset, int64_t *__restrict__ slot_point, int32_t *__restrict__ slot_cell) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        const double x = pts[p * sP], y = pts[p * sP + sC];
        for (int iy = first_candidate(y, g.min_y, g.stride, g.reach, g.gy); iy < g.gy && g.lo_y[iy] <= y; ++iy) {
            if (!(y <= g.hi_y[iy])) continue;
            for (int ix = first_candidate(x, g.min_x, g.stride, g.reach, g.gx); ix < g.gx && g.lo_x[ix] <= x; ++ix) {
                if (!(x <= g.hi_x[ix])) continue;
                const int cell = iy * g.gx + ix;                      // the reference's loop order: index_y outer, index_x inner
                const int k = atomicAdd(counts + cell, 1);
                if (FILL) {
                    slot_point[cell_offset[cell] + k] = p;
                    slot_cell[cell_offset[cell] + k] = cell;
                }
            }
        }
    }
}

// one thread per padding slot: slot j >= n of a cell with n members and `padded` slots
__global__ void __launch_bounds__(256)
slice_pad_kernel(const int32_t *__restrict__ counts, const int64_t *__restrict__ cell_offset, const int32_t *__restrict__ pad_cell,
                 const int64_t *__restrict__ pad_rank, const int64_t *__restrict__ rnd, int64_t n_pad, int block_points,
                 int64_t *__restrict__ slot_point, int32_t *__restrict__ slot_cell) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * blockDim.x) {
        const int cell = pad_cell[i];
        const int64_t n = counts[cell], j = pad_rank[i];              // j-th padding slot of the cell, 0-based
        const int64_t padded = (n + block_points - 1) / block_points * block_points, extra = padded - n;
        const int64_t base = cell_offset[cell];
        const int64_t src = extra <= n ? j : (int64_t)((uint64_t)rnd[i] % (uint64_t)n);   // :207 replace = extra > n
        slot_point[base + n + j] = slot_point[base + src];
        slot_cell[base + n + j] = cell;
    }
}

__global__ void __launch_bounds__(256)
slice_rows_kernel(const double *__restrict__ pts, int64_t sP, int64_t sC, const int64_t *__restrict__ labels,
                  const double *__restrict__ extra, int64_t eE, int64_t eP, const double *__restrict__ extra_div, int E,
                  const float *__restrict__ labelweights, const int64_t *__restrict__ slot_point,
                  const int32_t *__restrict__ slot_cell, const double *__restrict__ cx, const double *__restrict__ cy, int gx,
                  double max_x, double max_y, double max_z, int64_t S, float *__restrict__ rows, int64_t *__restrict__ out_label,
                  float *__restrict__ out_weight) {
    const int C = 6 + E;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = slot_point[s];
        const int cell = slot_cell[s];
        const double x = pts[p * sP], y = pts[p * sP + sC], z = pts[p * sP + 2 * sC];
        float *r = rows + s * C;
        // gx > 0: grid of cells (centre column / row arrays); gx == 0: one centre per cell (the training crops)
        r[0] = (float)__dsub_rn(x, gx > 0 ? cx[cell % gx] : cx[cell]);    // :221  data_batch[:, 0] - (s_x + block_size / 2.0)
        r[1] = (float)__dsub_rn(y, gx > 0 ? cy[cell / gx] : cy[cell]);
        r[2] = (float)z;
        r[3] = (float)__ddiv_rn(x, max_x);                             // :217-219 normalised by the scene maximum
        r[4] = (float)__ddiv_rn(y, max_y);
        r[5] = (float)__ddiv_rn(z, max_z);
        for (int e = 0; e < E; ++e) {
            const double v = extra[e * eE + p * eP];
            r[6 + e] = (float)(extra_div[e] != 1.0 ? __ddiv_rn(v, extra_div[e]) : v);   // :236-237 colours / 255
        }
        const int64_t lab = labels ? labels[p] : 0;
        out_label[s] = lab;
        if (out_weight) out_weight[s] = labelweights ? labelweights[lab] : 1.0f;       // :224 batch_weight = labelweights[label]
    }
}

// ---- the training crops (/root/reference/sem_seg_training.py:206-215, TrainCustomDataset.__getitem__): K candidate boxes
// [lo_x, hi_x] x [lo_y, hi_y] (centre -+ block_size / 2 in float64), every point tested against every box with the reference's
// comparisons; count pass / fill pass like slice_cells_kernel.  A box of a dense room holds 1e5+ points, so the per-box
// counter is bumped once per WARP (ballot + popc), not once per point.
constexpr int kMaxCrops = 64;
struct CropBoxes {
    double lo_x[kMaxCrops], hi_x[kMaxCrops], lo_y[kMaxCrops], hi_y[kMaxCrops];
    int K;
};

template <bool FILL>
__global__ void __launch_bounds__(256)
crop_members_kernel(const double *__restrict__ pts, int64_t sP, int64_t sC, int64_t P, const __grid_constant__ CropBoxes bx,
                    int32_t *__restrict__ counts, const int64_t *__restrict__ crop_offset, int64_t *__restrict__ slot_point,
                    int32_t *__restrict__ slot_crop, int crop_id0) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < P; base += stride) {
        const int64_t p = base + lane;
        const bool live = p < P;
        const double x = live ? pts[p * sP] : 0.0, y = live ? pts[p * sP + sC] : 0.0;
        for (int k = 0; k < bx.K; ++k) {
            const bool in = live && x >= bx.lo_x[k] && x <= bx.hi_x[k] && y >= bx.lo_y[k] && y <= bx.hi_y[k];
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (m == 0u) continue;
            const int leader = __ffs(m) - 1;
            int first = 0;
            if (lane == leader) first = atomicAdd(counts + k, __popc(m));
            first = __shfl_sync(0xffffffffu, first, leader);
            if (FILL && in) {
                const int64_t slot = crop_offset[k] + first + __popc(m & ((1u << lane) - 1u));
                slot_point[slot] = p;
                slot_crop[slot] = crop_id0 + k;
            }
        }
    }
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_crop_members(const double *points, int64_t sP, int64_t sC, int64_t P, const double *boxes_host, int K,
                                int crop_id0, int32_t *counts, const int64_t *crop_offset, int64_t *slot_point,
                                int32_t *slot_crop, void *stream) {
    PN2_REQUIRE(P >= 0 && K >= 0 && K <= kMaxCrops, "crop_members: bad sizes P=%lld K=%d (<= %d boxes per call)", (long long)P, K, kMaxCrops);
    if (P == 0 || K == 0) return PN2_OK;
    PN2_REQUIRE(points && boxes_host && counts, "crop_members: null pointer");
    PN2_REQUIRE(!crop_offset == !slot_point && !slot_point == !slot_crop, "crop_members: the fill pass needs crop_offset, slot_point and slot_crop");
    CropBoxes bx;
    bx.K = K;
    for (int k = 0; k < K; ++k) {                 // boxes_host: [K][4] = lo_x, hi_x, lo_y, hi_y (HOST memory, float64)
        bx.lo_x[k] = boxes_host[4 * k];
        bx.hi_x[k] = boxes_host[4 * k + 1];
        bx.lo_y[k] = boxes_host[4 * k + 2];
        bx.hi_y[k] = boxes_host[4 * k + 3];
    }
    const int grid = grid_for(P, 256);
    if (slot_point)
        crop_members_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(points, sP, sC, P, bx, counts, crop_offset, slot_point, slot_crop, crop_id0);
    else
        crop_members_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(points, sP, sC, P, bx, counts, nullptr, nullptr, nullptr, crop_id0);
    count_launch();
    return check_launch("crop_members");
}

static SliceGrid make_grid(const double *lo_x, const double *hi_x, int gx, const double *lo_y, const double *hi_y, int gy,
                           double min_x, double min_y, double stride, double block_size, double padding) {
    SliceGrid g;
    g.lo_x = lo_x; g.hi_x = hi_x; g.lo_y = lo_y; g.hi_y = hi_y;
    g.gx = gx; g.gy = gy; g.min_x = min_x; g.min_y = min_y; g.stride = stride; g.reach = block_size + padding;
    return g;
}

extern "C" int pn2_slice_cells(const double *points, int64_t sP, int64_t sC, int64_t P, const double *lo_x, const double *hi_x,
                               int gx, const double *lo_y, const double *hi_y, int gy, double min_x, double min_y, double stride,
                               double block_size, double padding, int32_t *counts, const int64_t *cell_offset,
                               int64_t *slot_point, int32_t *slot_cell, void *stream) {
    PN2_REQUIRE(P >= 0 && gx >= 0 && gy >= 0 && stride > 0.0, "slice_cells: bad sizes P=%lld gx=%d gy=%d", (long long)P, gx, gy);
    if (P == 0 || gx == 0 || gy == 0) return PN2_OK;
    PN2_REQUIRE(points && lo_x && hi_x && lo_y && hi_y && counts, "slice_cells: null pointer");
    PN2_REQUIRE(!cell_offset == !slot_point && !slot_point == !slot_cell, "slice_cells: the fill pass needs cell_offset, slot_point and slot_cell");
    const SliceGrid g = make_grid(lo_x, hi_x, gx, lo_y, hi_y, gy, min_x, min_y, stride, block_size, padding);
    const int grid = grid_for(P, 256);
    if (slot_point)
        slice_cells_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(points, sP, sC, P, g, counts, cell_offset, slot_point, slot_cell);
    else
        slice_cells_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(points, sP, sC, P, g, counts, nullptr, nullptr, nullptr);
    count_launch();
    return check_launch("slice_cells");
}

extern "C" int pn2_slice_pad(const int32_t *counts, const int64_t *cell_offset, const int32_t *pad_cell, const int64_t *pad_rank,
                             const int64_t *rnd, int64_t n_pad, int block_points, int64_t *slot_point, int32_t *slot_cell,
                             void *stream) {
    PN2_REQUIRE(n_pad >= 0 && block_points >= 1, "slice_pad: bad sizes");
    if (n_pad == 0) return PN2_OK;
    PN2_REQUIRE(counts && cell_offset && pad_cell && pad_rank && rnd && slot_point && slot_cell, "slice_pad: null pointer");
    slice_pad_kernel<<<grid_for(n_pad, 256), 256, 0, (cudaStream_t)stream>>>(counts, cell_offset, pad_cell, pad_rank, rnd, n_pad,
                                                                            block_points, slot_point, slot_cell);
    count_launch();
    return check_launch("slice_pad");
}

extern "C" int pn2_slice_rows(const double *points, int64_t sP, int64_t sC, const int64_t *labels, const double *extra, int64_t eE,
                              int64_t eP, const double *extra_div, int E, const float *labelweights, const int64_t *slot_point,
                              const int32_t *slot_cell, const double *cx, const double *cy, int gx, double max_x, double max_y,
                              double max_z, int64_t S, float *rows, int64_t *out_label, float *out_weight, void *stream) {
    PN2_REQUIRE(S >= 0 && E >= 0 && gx >= 0, "slice_rows: bad sizes");
    if (S == 0) return PN2_OK;
    PN2_REQUIRE(points && slot_point && slot_cell && cx && cy && rows && out_label, "slice_rows: null pointer");
    PN2_REQUIRE(E == 0 || (extra && extra_div), "slice_rows: extra features without data");
    slice_rows_kernel<<<grid_for(S, 256), 256, 0, (cudaStream_t)stream>>>(points, sP, sC, labels, extra, eE, eP, extra_div, E,
                                                                         labelweights, slot_point, slot_cell, cx, cy, gx, max_x,
                                                                         max_y, max_z, S, rows, out_label, out_weight);
    count_launch();
    return check_launch("slice_rows");
}
