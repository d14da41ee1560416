// optim.cu -- the optimizer step of the training loop (/root/reference/sem_seg_training.py:576-582: torch.optim.Adam,
// betas (0.9, 0.999), eps 1e-8, L2 weight decay) as ONE launch over the flat gradient buffer.
//
// Every parameter gradient already lives in one flat fp32 buffer (trainer.FlatGradients: the backward kernels write
// into it, NCCL all-reduces it); the two Adam moments are flat buffers with the same offsets.  The parameters stay
// separate allocations owned by the nn.Conv / nn.BatchNorm modules, so the kernel walks a chunk table
// (tensor, first element) and reads each tensor's pointer / flat offset / size from a device table.
// torch.optim.Adam(fused=True) takes 6 multi-tensor launches (~110 us) for the 969 k parameters of the network;
// this is 27 MB of traffic in one launch.
//
// Arithmetic (torch/optim/adam.py single-tensor form, amsgrad off, maximize off):
//   g += wd * p;  m += (1-b1) * (g - m);  v = b2*v + (1-b2)*g*g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),      t = step count after the increment.
// Hyper-parameters and the step counter are read from DEVICE memory, so a captured CUDA graph follows learning-rate
// changes and counts its own replays.  The last block to finish increments the counter (ticket), after every block
// has read it.
#include "common.cuh"

namespace pn2 {

constexpr int kAdamThreads = 256;

__device__ __forceinline__ void adam_update(float &p, float g, float &m, float &v, float wd, float one_minus_b1, float b2,
                                            float one_minus_b2, float step_size, float bc2_sqrt, float eps) {
    g = fmaf(wd, p, g);
    m = fmaf(one_minus_b1, g - m, m);
    v = fmaf(b2, v, one_minus_b2 * g * g);
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p -= (step_size * m) / denom;
}

// hyper: [lr, beta1, beta2, eps, weight_decay] (fp64);  step: fp32 counter of finished steps
__global__ void __launch_bounds__(kAdamThreads)
adam_flat_kernel(float *const *__restrict__ params, const int64_t *__restrict__ tensor_off, const int64_t *__restrict__ tensor_n,
                 const int2 *__restrict__ chunks, int chunk_elems, const float *__restrict__ grad, float *__restrict__ exp_avg,
                 float *__restrict__ exp_avg_sq, const double *__restrict__ hyper, float *step, unsigned *ticket) {
    __shared__ float sh[6];
    if (threadIdx.x == 0) {
        const double t = (double)(*reinterpret_cast<volatile float *>(step)) + 1.0;
        const double lr = hyper[0], b1 = hyper[1], b2 = hyper[2];
        const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
        sh[0] = (float)(lr / bc1);
        sh[1] = (float)sqrt(bc2);
        sh[2] = (float)(1.0 - b1);
        sh[3] = (float)b2;
        sh[4] = (float)(1.0 - b2);
        sh[5] = (float)hyper[3];
    }
    __syncthreads();
    const float step_size = sh[0], bc2_sqrt = sh[1], omb1 = sh[2], b2 = sh[3], omb2 = sh[4], eps = sh[5];
    const float wd = (float)hyper[4];
    const int2 ch = chunks[blockIdx.x];
    const int64_t n = tensor_n[ch.x], first = (int64_t)ch.y;
    const int len = (int)((n - first < chunk_elems) ? (n - first) : chunk_elems);
    float *p = params[ch.x] + first;
    const int64_t o = tensor_off[ch.x] + first;          // flat offsets are multiples of 4 elements, chunk starts too
    const float *g = grad + o;
    float *m = exp_avg + o, *v = exp_avg_sq + o;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);
    const int nv = vec ? (len >> 2) : 0;
    for (int i = threadIdx.x; i < nv; i += kAdamThreads) {
        float4 pp = reinterpret_cast<float4 *>(p)[i], mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        const float4 gg = reinterpret_cast<const float4 *>(g)[i];
        adam_update(pp.x, gg.x, mm.x, vv.x, wd, omb1, b2, omb2, step_size, bc2_sqrt, eps);
        adam_update(pp.y, gg.y, mm.y, vv.y, wd, omb1, b2, omb2, step_size, bc2_sqrt, eps);
        adam_update(pp.z, gg.z, mm.z, vv.z, wd, omb1, b2, omb2, step_size, bc2_sqrt, eps);
        adam_update(pp.w, gg.w, mm.w, vv.w, wd, omb1, b2, omb2, step_size, bc2_sqrt, eps);
        reinterpret_cast<float4 *>(p)[i] = pp;
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
    }
    for (int i = 4 * nv + threadIdx.x; i < len; i += kAdamThreads) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_update(pp, g[i], mm, vv, wd, omb1, b2, omb2, step_size, bc2_sqrt, eps);
        p[i] = pp;
        m[i] = mm;
        v[i] = vv;
    }
    // every block has read *step before it takes a ticket; the last one publishes t and re-arms the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned k = atomicAdd(ticket, 1u);
        if (k == gridDim.x - 1) {
            *reinterpret_cast<volatile float *>(step) = *reinterpret_cast<volatile float *>(step) + 1.0f;
            *ticket = 0u;
        }
    }
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_adam_step(void *const *params, const int64_t *tensor_off, const int64_t *tensor_n, const void *chunks,
                             int n_chunks, int chunk_elems, const float *grad_flat, float *exp_avg_flat,
                             float *exp_avg_sq_flat, const double *hyper, float *step, void *ticket, void *stream) {
    PN2_REQUIRE(params && tensor_off && tensor_n && chunks && grad_flat && exp_avg_flat && exp_avg_sq_flat && hyper && step &&
                    ticket, "adam_step: null pointer");
    PN2_REQUIRE(n_chunks >= 0 && chunk_elems >= 4 && chunk_elems % 4 == 0, "adam_step: bad chunking (%d chunks of %d)",
                n_chunks, chunk_elems);
    if (n_chunks == 0) return PN2_OK;
    adam_flat_kernel<<<n_chunks, kAdamThreads, 0, (cudaStream_t)stream>>>(
        (float *const *)params, tensor_off, tensor_n, (const int2 *)chunks, chunk_elems, grad_flat, exp_avg_flat,
        exp_avg_sq_flat, hyper, step, (unsigned *)ticket);
    count_launch();
    return check_launch("adam_step");
}
