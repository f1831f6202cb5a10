// Fused residual + LayerNorm + FiLM + padding-mask kernels (forward / backward), one warp per row.
//
//   v = dropout_in(a) + res ;  y = mask( gamma_film * dropout_out(LN(v) * w + b) + beta_film )
//
// Covers the three LayerNorm sites of the hot path:
//   * attention epilogue        LN(dropout(out_proj) + x), zero padded rows        reference model.py:189-191, :259
//   * conv feed-forward epilogue gamma * LN(dropout(conv2) + x) + beta, zero padded model.py:226-235, :262
//   * pre-net / predictor        dropout(LN(relu(conv)))  (ReLU is done in the GEMM epilogue)  model.py:341-363, :528-543
// HBM-bound: algorithmic bytes per row = (2 reads + 2 writes) * D * 4.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace dx {

// 4 consecutive fp32 -> bf16 hi (= bf16(x)) and lo (= bf16(x - hi)), 8 bytes each
__device__ __forceinline__ void store_planes4(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&o)[4]) {
    uint32_t h[2], l[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[k]) : "f"(o[2 * k + 1]), "f"(o[2 * k]));
        const float xr = o[2 * k] - __uint_as_float(h[k] << 16), yr = o[2 * k + 1] - __uint_as_float(h[k] & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l[k]) : "f"(yr), "f"(xr));
    }
    *reinterpret_cast<uint2*>(hi) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(lo) = make_uint2(l[0], l[1]);
}

template <int VPT>  // values per lane, D = 32 * VPT
__global__ void __launch_bounds__(256) ln_fwd_kernel(LnArgs p) {
    const unsigned long long seed_in = dyn_seed(p.seed_in, p.dyn), seed_out = dyn_seed(p.seed_out, p.dyn);
    (void)seed_in; (void)seed_out;
    constexpr int D = 32 * VPT, NV = VPT / 4;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int R = p.B * p.S;
    if (row >= R) return;
    const int b = row / p.S, s = row - b * p.S;
    const bool masked = p.lens && s >= (int)p.lens[b];
    float* yrow = p.y + (size_t)row * D;
    float* hrow = p.xhat + (size_t)row * D;
    // optional: y also as bf16 hi|lo operand planes [2][B*S][D] for the GEMM that consumes it (saves a split pass)
    __nv_bfloat16* phi = p.y_planes ? (__nv_bfloat16*)p.y_planes + (size_t)row * D : nullptr;
    __nv_bfloat16* plo = phi ? phi + (size_t)R * D : nullptr;
    if (masked) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = (j * 32 + lane) * 4;
            *reinterpret_cast<float4*>(yrow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(hrow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (phi) {
                *reinterpret_cast<uint2*>(phi + c) = make_uint2(0u, 0u);
                *reinterpret_cast<uint2*>(plo + c) = make_uint2(0u, 0u);
            }
        }
        if (lane == 0) p.rstd[row] = 0.f;
        return;
    }
    float v[VPT];
    const float* arow = p.a + (size_t)row * D;
    const float inv_keep_in = p.p_in > 0.f ? 1.f / (1.f - p.p_in) : 1.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        const float4 t = *reinterpret_cast<const float4*>(arow + c);
        v[j * 4 + 0] = t.x; v[j * 4 + 1] = t.y; v[j * 4 + 2] = t.z; v[j * 4 + 3] = t.w;
        if (p.p_in > 0.f) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[j * 4 + e] *= dropout_scale(seed_in, (unsigned long long)row * D + c + e, p.p_in, inv_keep_in);
        }
        if (p.res) {
            const float4 r = *reinterpret_cast<const float4*>(p.res + (size_t)row * D + c);
            v[j * 4 + 0] += r.x; v[j * 4 + 1] += r.y; v[j * 4 + 2] += r.z; v[j * 4 + 3] += r.w;
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) sum += v[i];
    const float mean = warp_sum(sum) * (1.f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) { const float d = v[i] - mean; sq += d * d; }
    const float rstd = rsqrtf(warp_sum(sq) * (1.f / D) + 1e-5f);
    if (lane == 0) p.rstd[row] = rstd;
    const float inv_keep_out = p.p_out > 0.f ? 1.f / (1.f - p.p_out) : 1.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        const float4 w = *reinterpret_cast<const float4*>(p.ln_w + c);
        const float4 bb = *reinterpret_cast<const float4*>(p.ln_b + c);
        const float wv[4] = {w.x, w.y, w.z, w.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
        float h[4], o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            h[e] = (v[j * 4 + e] - mean) * rstd;
            o[e] = h[e] * wv[e] + bv[e];
        }
        if (p.p_out > 0.f) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] *= dropout_scale(seed_out, (unsigned long long)row * D + c + e, p.p_out, inv_keep_out);
        }
        if (p.film) {
            const float4 g = *reinterpret_cast<const float4*>(p.film + (size_t)b * p.film_stride + c);
            const float4 be = *reinterpret_cast<const float4*>(p.film + (size_t)b * p.film_stride + D + c);
            o[0] = g.x * o[0] + be.x; o[1] = g.y * o[1] + be.y; o[2] = g.z * o[2] + be.z; o[3] = g.w * o[3] + be.w;
        }
        *reinterpret_cast<float4*>(hrow + c) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(yrow + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (phi) store_planes4(phi + c, plo + c, o);
    }
}

// Backward.  grad wrt v:  g = dy * mask * drop_out * film_gamma * w ;  dv = rstd * (g - mean(g) - xhat * mean(g * xhat));
// column reductions: dln_w[c] = sum_r q * xhat, dln_b[c] = sum_r q with q = dy * mask * drop_out * film_gamma;
// dfilm_gamma[b][c] = sum_s e * drop_out(xhat * w + b), dfilm_beta[b][c] = sum_s e with e = dy * mask.

// Fused backward for D <= 256: grad wrt v / a AND the column reductions (parameters, FiLM) AND (optionally) the
// bf16 hi|lo operand planes + column sums (= bias gradient of the producing GEMM) of the gradient that leaves through `a`.
// Block = 8 warps over `rpb` consecutive rows of ONE utterance (grid (ceil(S/rpb), B)); a warp walks rows w, w+8, ...; each
// lane keeps its columns' partial sums in registers; one shared-memory pass + fp32 atomics per block at the end.
template <int VPT>
__global__ void __launch_bounds__(256) ln_bwd_fused_kernel(LnArgs p, int rpb) {
    const unsigned long long seed_in = dyn_seed(p.seed_in, p.dyn), seed_out = dyn_seed(p.seed_out, p.dyn);
    (void)seed_in; (void)seed_out;
    constexpr int D = 32 * VPT, NV = VPT / 4;
    __shared__ float red[8][D];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int R = p.B * p.S;
    const int len = p.lens ? min((int)p.lens[b], p.S) : p.S;
    const int s_begin = blockIdx.x * rpb, s_end = min(p.S, s_begin + rpb);
    const float inv_keep_out = p.p_out > 0.f ? 1.f / (1.f - p.p_out) : 1.f;
    const float inv_keep_in = p.p_in > 0.f ? 1.f / (1.f - p.p_in) : 1.f;
    float wv[VPT], bv[VPT], fg[VPT];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        const float4 w = *reinterpret_cast<const float4*>(p.ln_w + c);
        const float4 bb = *reinterpret_cast<const float4*>(p.ln_b + c);
        wv[j * 4] = w.x; wv[j * 4 + 1] = w.y; wv[j * 4 + 2] = w.z; wv[j * 4 + 3] = w.w;
        bv[j * 4] = bb.x; bv[j * 4 + 1] = bb.y; bv[j * 4 + 2] = bb.z; bv[j * 4 + 3] = bb.w;
        if (p.film) {
            const float4 f4 = *reinterpret_cast<const float4*>(p.film + (size_t)b * p.film_stride + c);
            fg[j * 4] = f4.x; fg[j * 4 + 1] = f4.y; fg[j * 4 + 2] = f4.z; fg[j * 4 + 3] = f4.w;
        } else {
            fg[j * 4] = fg[j * 4 + 1] = fg[j * 4 + 2] = fg[j * 4 + 3] = 1.f;
        }
    }
    float aw[VPT], ab[VPT], ag[VPT], abe[VPT], ac[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) aw[i] = ab[i] = ag[i] = abe[i] = ac[i] = 0.f;

    for (int s = s_begin + warp; s < s_end; s += 8) {
        const size_t row = (size_t)b * p.S + s;
        float* dvrow = p.dv + row * D;
        float* darow = p.da ? p.da + row * D : nullptr;
        __nv_bfloat16* phi = p.g_planes ? (__nv_bfloat16*)p.g_planes + row * D : nullptr;
        __nv_bfloat16* plo = phi ? phi + (size_t)R * D : nullptr;
        if (s >= len) {   // padded row: zero gradients, nothing to accumulate
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int c = (j * 32 + lane) * 4;
                *reinterpret_cast<float4*>(dvrow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (darow) *reinterpret_cast<float4*>(darow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (phi) {
                    *reinterpret_cast<uint2*>(phi + c) = make_uint2(0u, 0u);
                    *reinterpret_cast<uint2*>(plo + c) = make_uint2(0u, 0u);
                }
            }
            continue;
        }
        float g[VPT], h[VPT];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = (j * 32 + lane) * 4;
            const float4 d = *reinterpret_cast<const float4*>(p.dy + row * D + c);
            const float4 hh = *reinterpret_cast<const float4*>(p.xhat + row * D + c);
            float dv[4] = {d.x, d.y, d.z, d.w};
            h[j * 4 + 0] = hh.x; h[j * 4 + 1] = hh.y; h[j * 4 + 2] = hh.z; h[j * 4 + 3] = hh.w;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = j * 4 + e;
                abe[i] += dv[e];                                                     // dfilm beta: e = dy * mask
                if (p.p_out > 0.f) dv[e] *= dropout_scale(seed_out, row * D + c + e, p.p_out, inv_keep_out);
                ag[i] += dv[e] * (h[i] * wv[i] + bv[i]);                            // dfilm gamma
                dv[e] *= fg[i];
                aw[i] += dv[e] * h[i];                                               // dln_w
                ab[i] += dv[e];                                                      // dln_b
                g[i] = dv[e] * wv[i];
                s1 += g[i];
                s2 += g[i] * h[i];
            }
        }
        const float c1 = warp_sum(s1) * (1.f / D), c2 = warp_sum(s2) * (1.f / D);
        const float rstd = p.rstd[row];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = (j * 32 + lane) * 4;
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = rstd * (g[j * 4 + e] - c1 - h[j * 4 + e] * c2);
            if (p.relu_src) {
                const float4 rs = *reinterpret_cast<const float4*>(p.relu_src + row * D + c);
                o[0] = rs.x > 0.f ? o[0] : 0.f; o[1] = rs.y > 0.f ? o[1] : 0.f; o[2] = rs.z > 0.f ? o[2] : 0.f; o[3] = rs.w > 0.f ? o[3] : 0.f;
            }
            *reinterpret_cast<float4*>(dvrow + c) = make_float4(o[0], o[1], o[2], o[3]);
            if (darow) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] *= dropout_scale(seed_in, row * D + c + e, p.p_in, inv_keep_in);
                *reinterpret_cast<float4*>(darow + c) = make_float4(o[0], o[1], o[2], o[3]);
            }
            if (phi) store_planes4(phi + c, plo + c, o);          // planes / column sums of da (p_in > 0) or dv
#pragma unroll
            for (int e = 0; e < 4; ++e) ac[j * 4 + e] += o[e];
        }
    }

    // block reduction of the per-lane column sums, one quantity at a time
    auto flush = [&](const float (&acc)[VPT], float* dst) {
        if (dst == nullptr) return;   // (uniform across the block)
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = (j * 32 + lane) * 4;
            *reinterpret_cast<float4*>(&red[warp][c]) = make_float4(acc[j * 4], acc[j * 4 + 1], acc[j * 4 + 2], acc[j * 4 + 3]);
        }
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += 256) {
            float tot = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) tot += red[k][c];
            atomicAdd(dst + c, tot);
        }
    };
    flush(aw, p.dln_w);
    flush(ab, p.dln_b);
    flush(ag, p.dfilm ? p.dfilm + (size_t)b * 2 * D : nullptr);
    flush(abe, p.dfilm ? p.dfilm + (size_t)b * 2 * D + D : nullptr);
    flush(ac, p.g_colsum);
}

// Fused backward for D = 1024 (pre-net): same outputs as ln_bwd_fused_kernel, organised the other way round because a lane
// cannot keep 32 columns x 3 running sums in registers.  Block = 8 warps over `rpb` rows of ONE utterance; WARP w owns the
// column slice [128 w, 128 w + 128) of every row (a lane 4 consecutive columns, so its column sums stay in registers), rows are
// processed four at a time: slice partial sums of the two LayerNorm statistics -> shared memory -> row totals -> outputs.
constexpr int LNW_ROWS = 4;   // rows in flight per warp (sixteen 16-byte loads per lane and phase)
__global__ void __launch_bounds__(256) ln_bwd_fused_wide_kernel(LnArgs p, int rpb) {
    const unsigned long long seed_in = dyn_seed(p.seed_in, p.dyn), seed_out = dyn_seed(p.seed_out, p.dyn);
    (void)seed_in; (void)seed_out;
    constexpr int D = 1024;
    __shared__ float part[2][LNW_ROWS][8];    // [statistic][row][warp slice]
    __shared__ float tot[2][LNW_ROWS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int R = p.B * p.S;
    const int len = p.lens ? min((int)p.lens[b], p.S) : p.S;
    const int s_begin = blockIdx.x * rpb, s_end = min(p.S, s_begin + rpb);
    const int c = warp * 128 + lane * 4;      // this lane's 4 columns
    const float inv_keep_out = p.p_out > 0.f ? 1.f / (1.f - p.p_out) : 1.f;
    const float inv_keep_in = p.p_in > 0.f ? 1.f / (1.f - p.p_in) : 1.f;
    const float4 w4 = *reinterpret_cast<const float4*>(p.ln_w + c);
    const float4 b4 = *reinterpret_cast<const float4*>(p.ln_b + c);
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
    float fg[4] = {1.f, 1.f, 1.f, 1.f};
    if (p.film) {
        const float4 f4 = *reinterpret_cast<const float4*>(p.film + (size_t)b * p.film_stride + c);
        fg[0] = f4.x; fg[1] = f4.y; fg[2] = f4.z; fg[3] = f4.w;
    }
    float aw[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f}, ag[4] = {0.f, 0.f, 0.f, 0.f}, abe[4] = {0.f, 0.f, 0.f, 0.f},
          ac[4] = {0.f, 0.f, 0.f, 0.f};

    for (int s0 = s_begin; s0 < s_end; s0 += LNW_ROWS) {
        float q[LNW_ROWS][4], h[LNW_ROWS][4];   // q = dy * mask * drop_out * film_gamma (gradient wrt the normalised, scaled value)
        // all sixteen loads of the eight rows are issued before the first use (a load inside the per-row branch / shuffle sequence
        // is not hoisted by the compiler and costs one exposed DRAM latency per row); invalid rows read row s0 and are discarded
        float4 dld[LNW_ROWS], hld[LNW_ROWS];
#pragma unroll
        for (int r = 0; r < LNW_ROWS; ++r) {
            const int s = (s0 + r < s_end && s0 + r < len) ? s0 + r : s0;
            const size_t row = (size_t)b * p.S + s;
            dld[r] = __ldg(reinterpret_cast<const float4*>(p.dy + row * D + c));
            hld[r] = __ldg(reinterpret_cast<const float4*>(p.xhat + row * D + c));
        }
#pragma unroll
        for (int r = 0; r < LNW_ROWS; ++r) {
            const int s = s0 + r;
            float s1 = 0.f, s2 = 0.f;
            if (s < s_end && s < len) {
                const size_t row = (size_t)b * p.S + s;
                const float4 d = dld[r];
                const float4 hh = hld[r];
                float dv[4] = {d.x, d.y, d.z, d.w};
                h[r][0] = hh.x; h[r][1] = hh.y; h[r][2] = hh.z; h[r][3] = hh.w;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    abe[e] += dv[e];
                    if (p.p_out > 0.f) dv[e] *= dropout_scale(seed_out, row * D + c + e, p.p_out, inv_keep_out);
                    ag[e] += dv[e] * (h[r][e] * wv[e] + bv[e]);
                    dv[e] *= fg[e];
                    aw[e] += dv[e] * h[r][e];
                    ab[e] += dv[e];
                    q[r][e] = dv[e];
                    const float g = dv[e] * wv[e];
                    s1 += g;
                    s2 += g * h[r][e];
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) { q[r][e] = 0.f; h[r][e] = 0.f; }
            }
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0) { part[0][r][warp] = s1; part[1][r][warp] = s2; }
        }
        __syncthreads();
        if (threadIdx.x < 2 * LNW_ROWS) {
            const int st = threadIdx.x / LNW_ROWS, r = threadIdx.x % LNW_ROWS;
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += part[st][r][k];
            tot[st][r] = t * (1.f / D);
        }
        __syncthreads();
        float4 rld[LNW_ROWS];
        float rstd_ld[LNW_ROWS];
#pragma unroll
        for (int r = 0; r < LNW_ROWS; ++r) {   // same for the ReLU sources and the row scales of the second phase
            const int s = (s0 + r < s_end && s0 + r < len) ? s0 + r : s0;
            const size_t row = (size_t)b * p.S + s;
            rld[r] = p.relu_src ? __ldg(reinterpret_cast<const float4*>(p.relu_src + row * D + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
            rstd_ld[r] = __ldg(p.rstd + row);
        }
#pragma unroll
        for (int r = 0; r < LNW_ROWS; ++r) {
            const int s = s0 + r;
            if (s >= s_end) break;
            const size_t row = (size_t)b * p.S + s;
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            if (s < len) {
                const float c1 = tot[0][r], c2 = tot[1][r], rstd = rstd_ld[r];
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = rstd * (q[r][e] * wv[e] - c1 - h[r][e] * c2);
                if (p.relu_src) {
                    const float4 rs = rld[r];
                    o[0] = rs.x > 0.f ? o[0] : 0.f; o[1] = rs.y > 0.f ? o[1] : 0.f; o[2] = rs.z > 0.f ? o[2] : 0.f; o[3] = rs.w > 0.f ? o[3] : 0.f;
                }
            }
            *reinterpret_cast<float4*>(p.dv + row * D + c) = make_float4(o[0], o[1], o[2], o[3]);
            if (p.da) {
                if (s < len) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] *= dropout_scale(seed_in, row * D + c + e, p.p_in, inv_keep_in);
                }
                *reinterpret_cast<float4*>(p.da + row * D + c) = make_float4(o[0], o[1], o[2], o[3]);
            }
            if (p.g_planes) {
                __nv_bfloat16* phi = (__nv_bfloat16*)p.g_planes + row * D + c;
                store_planes4(phi, phi + (size_t)R * D, o);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) ac[e] += o[e];
        }
        __syncthreads();   // tot[] has been read by everyone before the next eight rows overwrite part[] / tot[]
    }
    atomicAdd(p.dln_w + c + 0, aw[0]); atomicAdd(p.dln_w + c + 1, aw[1]); atomicAdd(p.dln_w + c + 2, aw[2]); atomicAdd(p.dln_w + c + 3, aw[3]);
    atomicAdd(p.dln_b + c + 0, ab[0]); atomicAdd(p.dln_b + c + 1, ab[1]); atomicAdd(p.dln_b + c + 2, ab[2]); atomicAdd(p.dln_b + c + 3, ab[3]);
    if (p.dfilm) {
        float* df = p.dfilm + (size_t)b * 2 * D;
#pragma unroll
        for (int e = 0; e < 4; ++e) { atomicAdd(df + c + e, ag[e]); atomicAdd(df + D + c + e, abe[e]); }
    }
    if (p.g_colsum) {
#pragma unroll
        for (int e = 0; e < 4; ++e) atomicAdd(p.g_colsum + c + e, ac[e]);
    }
}

int ln_fwd(const LnArgs& a, cudaStream_t st) {
    const int R = a.B * a.S;
    const int blocks = ceil_div(R, 8);
    switch (a.D) {
        case 128: ln_fwd_kernel<4><<<blocks, 256, 0, st>>>(a); break;
        case 256: ln_fwd_kernel<8><<<blocks, 256, 0, st>>>(a); break;
        case 1024: ln_fwd_kernel<32><<<blocks, 256, 0, st>>>(a); break;
        default: set_last_error("ln_fwd: unsupported width D=%d (128, 256, 1024)", a.D); return DX_ERR_UNSUPPORTED;
    }
    return check_launch("ln_fwd");
}

int ln_bwd(const LnArgs& a, cudaStream_t st) {
    // one fused launch per call: dv / da + parameter / FiLM / bias-gradient column sums + operand planes
    if (a.D != 128 && a.D != 256 && a.D != 1024) {
        set_last_error("ln_bwd: unsupported width D=%d (128, 256, 1024)", a.D);
        return DX_ERR_UNSUPPORTED;
    }
    DX_CUDA(cudaMemsetAsync(a.dln_w, 0, (size_t)a.D * sizeof(float), st));
    DX_CUDA(cudaMemsetAsync(a.dln_b, 0, (size_t)a.D * sizeof(float), st));
    if (a.dfilm) DX_CUDA(cudaMemsetAsync(a.dfilm, 0, (size_t)a.B * 2 * a.D * sizeof(float), st));
    if (a.g_colsum) DX_CUDA(cudaMemsetAsync(a.g_colsum, 0, (size_t)a.D * sizeof(float), st));
    if (a.D == 1024) {   // pre-net width: warp-per-column-slice organisation
        const int rpb = a.S >= 512 ? 64 : (a.S >= 64 ? 32 : 8);
        dim3 grid(ceil_div(a.S, rpb), a.B);
        ln_bwd_fused_wide_kernel<<<grid, 256, 0, st>>>(a, rpb);
        return check_launch("ln_bwd_fused_wide");
    }
    const int rpb = a.S >= 512 ? 64 : (a.S >= 128 ? 32 : 16);
    dim3 grid(ceil_div(a.S, rpb), a.B);
    if (a.D == 128) ln_bwd_fused_kernel<4><<<grid, 256, 0, st>>>(a, rpb);
    else ln_bwd_fused_kernel<8><<<grid, 256, 0, st>>>(a, rpb);
    return check_launch("ln_bwd_fused");
}

}  // namespace dx
