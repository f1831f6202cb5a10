// Shared device/host helpers for the daft_exprt_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define DX_OK 0
#define DX_ERR_ARG -1
#define DX_ERR_CUDA -2
#define DX_ERR_UNSUPPORTED -3

namespace dx {

void set_last_error(const char* fmt, ...);
int check_launch(const char* what);

#define DX_REQUIRE(cond, ...)                          \
    do {                                               \
        if (!(cond)) {                                 \
            dx::set_last_error(__VA_ARGS__);           \
            return DX_ERR_ARG;                         \
        }                                              \
    } while (0)

#define DX_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            dx::set_last_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return DX_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
// 1-D grid for a grid-stride loop: ceil(n / per) blocks, at least 1, capped at a few waves of the 148 SMs
static inline int grid_1d(size_t n, int per = 256, int cap = 148 * 16) {
    size_t g = (n + per - 1) / per;
    if (g < 1) g = 1;
    if (g > (size_t)cap) g = cap;
    return (int)g;
}

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// round-to-nearest fp32 -> tf32 (kept in an fp32 container); tcgen05 kind::tf32 ignores the low 13 mantissa bits,
// so operands that only feed tensor-core GEMMs are pre-rounded to avoid a truncation bias.
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// Per-step scalars that a CUDA-graph replay must read from DEVICE memory instead of baked kernel arguments (see
// dx_set_step_state in include/daft_exprt_b200.h).  32 bytes, written by the host before every replay.
struct StepState {
    unsigned long long seed_epoch;   // mixed into every dropout seed: fresh masks per replay, same masks in forward and backward
    float w_adv;                     // adversarial speaker-loss weight of this iteration (loss.py:30-38)
    float lr;                        // Adam learning rate
    float bc1, bc2_sqrt;             // Adam bias corrections 1 - beta1^t, sqrt(1 - beta2^t)
    float pad[2];
};
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long dyn_seed(unsigned long long seed, const StepState* d) {
    return d ? seed + d->seed_epoch * 0x9E3779B97F4A7C15ull : seed;
}
#endif

// Stateless counter-based dropout: the same (seed, index) gives the same decision in forward and backward, so masks are
// never stored.  32-bit murmur3-style finaliser over (index, seed): ~10 integer instructions per element (the attention
// kernels evaluate it once per score).  keep <=> hash >= p * 2^32.
__device__ __forceinline__ uint32_t hash_u32(unsigned long long seed, unsigned long long idx) {
    uint32_t x = (uint32_t)idx * 0x9E3779B1u + (uint32_t)seed;
    x ^= (uint32_t)(idx >> 32) * 0x85EBCA77u + (uint32_t)(seed >> 32);
    x ^= x >> 16; x *= 0x85EBCA6Bu;
    x ^= x >> 13; x *= 0xC2B2AE35u;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long idx) {
    return (float)(hash_u32(seed, idx) >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float dropout_scale(unsigned long long seed, unsigned long long idx, float p, float inv_keep) {
    return hash_uniform(seed, idx) >= p ? inv_keep : 0.f;
}
// Two-level variant for the tensor-core attention kernels (one decision per score: the per-score ALU work is what bounds them):
// a fully mixed key per query row (hash_u32, once per row) xor a Weyl term of the key index, then ONE xorshift-multiply round;
// the unsigned compare is decided by the high bits of the product, which depend on every bit of the input.
// keep <=> x >= p * 2^32.  (tests: kept fraction, per-row binomial variance, forward/backward mask agreement.)
__device__ __forceinline__ uint32_t drop_col_term(uint32_t col) { return col * 0x9E3779B1u; }
__device__ __forceinline__ bool drop_keep(uint32_t row_key, uint32_t col_term, uint32_t thresh) {
    uint32_t x = row_key ^ col_term;
    x ^= x >> 16; x *= 0x85EBCA6Bu;
    return x >= thresh;
}
__device__ __forceinline__ uint32_t drop_threshold(float p) { return __float2uint_rz(fminf(p, 0.99999994f) * 4294967296.f); }
#endif

}  // namespace dx
