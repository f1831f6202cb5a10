// Fused Adam over one flat fp32 parameter buffer (torch.optim.Adam semantics: L2 weight decay folded into the gradient,
// bias-corrected moments; reference train.py:299-301 uses betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6).
// HBM-bound: reads p, g, m, v and writes p, m, v = 28 B per parameter.
#include "common.cuh"
#include "kernels.h"

namespace dx {

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                            float grad_scale, const StepState* dyn, const float* __restrict__ clip) {
    if (dyn) { lr = dyn->lr; bc1 = dyn->bc1; bc2_sqrt = dyn->bc2_sqrt; }   // graph replay: per-step scalars from device memory
    if (clip) grad_scale *= clip[1];   // clip_grad_norm_ coefficient computed on the device by grad_norm_clip()
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        const float gi = g[i] * grad_scale + wd * pi;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

int adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
              float weight_decay, int step, float grad_scale, const StepState* dyn, const float* clip, cudaStream_t st) {
    DX_REQUIRE(step >= 1, "adam_step: step must be >= 1");
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    adam_kernel<<<grid_1d(n), 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                            bc1, bc2_sqrt, grad_scale, dyn, clip);
    return check_launch("adam_step");
}

// torch.nn.utils.clip_grad_norm_ (reference train.py:399) over the flat gradient buffer, entirely on the device:
// out[0] = ||grad_scale * g||_2, out[1] = min(1, max_norm / (out[0] + 1e-6)) (the factor clip_grad_norm_ multiplies the gradients by;
// 1 when max_norm is inf), out[2] = scratch sum of squares.  The Adam kernel folds out[1] into its gradient scale.
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out) {
    float acc = 0.f;
    const size_t n4 = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = g4[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[n4 * 4 + threadIdx.x]; acc += v * v; }
    __shared__ float part[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(out + 2, t);
    }
}
__global__ void grad_clip_finalize_kernel(float* out, float grad_scale, float max_norm) {
    const float norm = sqrtf(out[2]) * fabsf(grad_scale);
    out[0] = norm;
    out[1] = fminf(1.f, max_norm / (norm + 1e-6f));   // inf / x = inf -> 1
}

int grad_norm_clip(const float* g, size_t n, float grad_scale, float max_norm, float* out, cudaStream_t st) {
    DX_REQUIRE(((uintptr_t)g & 15) == 0, "grad_norm_clip: gradient buffer must be 16-byte aligned");
    DX_CUDA(cudaMemsetAsync(out, 0, 3 * sizeof(float), st));
    grad_sumsq_kernel<<<grid_1d(n / 4 + 1, 256, 148 * 4), 256, 0, st>>>(g, n, out);
    int rc = check_launch("grad_sumsq");
    if (rc) return rc;
    grad_clip_finalize_kernel<<<1, 1, 0, st>>>(out, grad_scale, max_norm);
    return check_launch("grad_clip_finalize");
}

}  // namespace dx
