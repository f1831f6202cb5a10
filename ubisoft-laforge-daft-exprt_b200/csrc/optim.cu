// Fused Adam over one flat fp32 parameter buffer (torch.optim.Adam semantics: L2 weight decay folded into the gradient,
// bias-corrected moments; reference train.py:299-301 uses betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6).
// HBM-bound: reads p, g, m, v and writes p, m, v = 28 B per parameter.
#include "common.cuh"
#include "kernels.h"

namespace dx {

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                            float grad_scale, const StepState* dyn) {
    if (dyn) { lr = dyn->lr; bc1 = dyn->bc1; bc2_sqrt = dyn->bc2_sqrt; }   // graph replay: per-step scalars from device memory
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
              float weight_decay, int step, float grad_scale, const StepState* dyn, cudaStream_t st) {
    DX_REQUIRE(step >= 1, "adam_step: step must be >= 1");
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    adam_kernel<<<grid_1d(n), 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                            bc1, bc2_sqrt, grad_scale, dyn);
    return check_launch("adam_step");
}

}  // namespace dx
