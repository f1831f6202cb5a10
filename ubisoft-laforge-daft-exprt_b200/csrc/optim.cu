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


// ---- fused gradient exchange + optimiser over NVLink peer memory -------------------------------------------------------------------
// Replaces `NCCL all-reduce of the flat gradient bucket -> Adam over the whole bucket` (reference: DDP all-reduce + torch.optim.Adam,
// train.py:293,391,401) by ONE kernel per rank over symmetric memory (every rank's flat gradient / parameter buffers mapped into every
// other rank's address space, torch.distributed._symmetric_memory):
//     reduce-scatter : this rank sums ITS shard [begin, begin + n) of the gradient over all ranks — one `multimem.ld_reduce` per 16 bytes
//                      when the buffers have an NVSwitch multicast mapping (the switch adds the N copies), else N peer loads;
//     Adam           : on the shard only (moments exist per shard: ZeRO-1 style);
//     all-gather     : the updated parameters are written to every rank — one `multimem.st` per 16 bytes, else N peer stores.
// NVLink traffic per rank = shard in + shard out (x N without multicast), no intermediate buffers, no second launch.
// The caller brackets the launch with cross-GPU barriers (all gradients written before / all parameters visible after).
__device__ __forceinline__ float4 ld_peer(const float* p) {   // remote HBM: system-scope, never from a stale L1 line
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer(float* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct PeerPtrs { unsigned long long p[16]; };

__global__ void __launch_bounds__(256) fused_reduce_adam_kernel(const float* g_mc, PeerPtrs g_peers, float* p_mc, PeerPtrs p_peers,
                                                                const float* p_local, float* m, float* v, size_t begin4, size_t n4, int world,
                                                                float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                                float grad_scale, const StepState* dyn) {
    if (dyn) { lr = dyn->lr; bc1 = dyn->bc1; bc2_sqrt = dyn->bc2_sqrt; }
    const float step = lr / bc1;
    constexpr int U = 4;   // independent 16-byte NVLink loads in flight per thread (remote latency, not issue rate, is the limit)
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < n4; i0 += U * stride) {
        float4 g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * stride;
            g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n4) {
                const size_t e = (begin4 + i) * 4;
                if (g_mc) {
                    g[u] = mc_ld_reduce(g_mc + e);
                } else {
                    for (int r = 0; r < world; ++r) {
                        const float4 t = ld_peer(reinterpret_cast<const float*>(g_peers.p[r]) + e);
                        g[u].x += t.x; g[u].y += t.y; g[u].z += t.z; g[u].w += t.w;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * stride;
            if (i >= n4) break;
            const size_t e = (begin4 + i) * 4;
            const float4 pv = *reinterpret_cast<const float4*>(p_local + e);
            const float4 mv = *reinterpret_cast<const float4*>(m + e), vv = *reinterpret_cast<const float4*>(v + e);
            float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {g[u].x, g[u].y, g[u].z, g[u].w}, ma[4] = {mv.x, mv.y, mv.z, mv.w},
                  va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gi = ga[k] * grad_scale + wd * pa[k];
                ma[k] = b1 * ma[k] + (1.f - b1) * gi;
                va[k] = b2 * va[k] + (1.f - b2) * gi * gi;
                pa[k] = pa[k] - step * (ma[k] / (sqrtf(va[k]) / bc2_sqrt + eps));
            }
            *reinterpret_cast<float4*>(m + e) = make_float4(ma[0], ma[1], ma[2], ma[3]);
            *reinterpret_cast<float4*>(v + e) = make_float4(va[0], va[1], va[2], va[3]);
            const float4 pn = make_float4(pa[0], pa[1], pa[2], pa[3]);
            if (p_mc) {
                mc_st(p_mc + e, pn);
            } else {
                for (int r = 0; r < world; ++r) st_peer(reinterpret_cast<float*>(p_peers.p[r]) + e, pn);
            }
        }
    }
}

int fused_reduce_adam(const float* g_mc, const unsigned long long* g_peers, float* p_mc, const unsigned long long* p_peers, const float* p_local,
                      float* m, float* v, size_t begin, size_t n, int world, float lr, float beta1, float beta2, float eps, float weight_decay,
                      int step, float grad_scale, const StepState* dyn, cudaStream_t st) {
    DX_REQUIRE(step >= 1 && world >= 1 && world <= 16, "fused_reduce_adam: step %d world %d (1..16 ranks)", step, world);
    DX_REQUIRE(begin % 4 == 0 && n % 4 == 0, "fused_reduce_adam: the shard must start and end on 16-byte boundaries (begin %zu n %zu)", begin, n);
    DX_REQUIRE((g_mc || g_peers) && (p_mc || p_peers) && p_local && m && v, "fused_reduce_adam: missing buffers");
    PeerPtrs gp = {}, pp = {};
    for (int r = 0; r < world; ++r) { gp.p[r] = g_peers ? g_peers[r] : 0ull; pp.p[r] = p_peers ? p_peers[r] : 0ull; }
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    if (n == 0) return DX_OK;
    fused_reduce_adam_kernel<<<grid_1d(n / 16 + 1, 256, 148 * 8), 256, 0, st>>>(g_mc, gp, p_mc, pp, p_local, m, v, begin / 4, n / 4, world, lr, beta1, beta2,
                                                                         eps, weight_decay, bc1, bc2_sqrt, grad_scale, dyn);
    return check_launch("fused_reduce_adam");
}

}  // namespace dx
