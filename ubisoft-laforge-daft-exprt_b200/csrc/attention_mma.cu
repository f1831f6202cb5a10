// Tensor-core flash attention (forward + backward) with fp32-grade accuracy.
//
// Every operand of the five products (Q·K^T, P·V, dO·V^T, P^T·dO, dS^T·Q, dS·K) is split into bf16 hi + lo and each product
// is formed as hi*hi + lo*hi + hi*lo with fp32 accumulation (the same "bf16x3" scheme as the tcgen05 GEMMs), so the result
// matches the reference's fp32 nn.MultiheadAttention math path (model.py:182-186) to ~1e-5 while running on the tensor pipe.
// Scores, softmax state and accumulators never leave registers (`mma.sync.m16n8k16`, flash-attention-2 dataflow: the score
// accumulator's C-fragment layout is re-used directly as the A-fragment of the next product).
//
// Operand planes.  `attn_prep_kernel` converts the fp32 projections ONCE per layer into per-head bf16 hi|lo planes, row-major
// R[plane][b][head][s][dh] and transposed Tr[plane][b][head][dh][s] (q pre-scaled by 1/sqrt(dh), rows >= S zero-filled up to a
// multiple of 64).  The attention CTAs then only issue 16-byte `cp.async` copies of ready-made tiles into a double-buffered
// shared-memory ring, overlapped with the MMAs of the current tile: no conversion work and no exposed global latency inside the
// key/query loops (the first version re-split every K/V tile in each of the 16 query-tile CTAs that consumed it).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"

namespace dx {

namespace {

__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (x, y) -> packed bf16 pairs {lo16 = x, hi16 = y}: hi = bf16(v), lo = bf16(v - hi)
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y), "f"(x));
    const float xr = x - __uint_as_float(hi << 16), yr = y - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(yr), "f"(xr));
}
// exp(x) for x <= 0 through ex2 (2^-22 relative): softmax inputs only
__device__ __forceinline__ float fast_exp(float x) { return exp2f(x * 1.4426950408889634f); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------------
// planes: fp32 [B, S, ld] (ncols = NH * dh columns starting at src) -> R / Tr bf16 hi|lo
// grid (Sp / 64, ncols / 64, B), 256 threads, one 64 (s) x 64 (c) tile per block
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_prep_kernel(const float* __restrict__ src, int ld, int B, int S, int Sp, int NH, int dh,
                                                        int scale_cols, float scale, __nv_bfloat16* __restrict__ R,
                                                        __nv_bfloat16* __restrict__ Tr) {
    __shared__ __align__(16) unsigned short th[64][72], tl[64][72];   // [c][s] bf16 bit patterns
    const int b = blockIdx.z, s0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int t = threadIdx.x;
    const size_t plane = (size_t)B * NH * Sp * dh;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = t + 256 * i, r = idx >> 4, cq = (idx & 15) * 4;
        const int s = s0 + r, c = c0 + cq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < S) v = *reinterpret_cast<const float4*>(src + ((size_t)b * S + s) * ld + c);
        if (c < scale_cols) { v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale; }
        uint32_t h0, l0, h1, l1;
        split_pair(v.x, v.y, h0, l0);
        split_pair(v.z, v.w, h1, l1);
        if (R) {
            const int hh = c / dh, d = c - hh * dh;
            const size_t o = (((size_t)b * NH + hh) * Sp + s) * dh + d;
            *reinterpret_cast<uint2*>(R + o) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(R + plane + o) = make_uint2(l0, l1);
        }
        if (Tr) {
            th[cq + 0][r] = (unsigned short)(h0 & 0xffff); th[cq + 1][r] = (unsigned short)(h0 >> 16);
            th[cq + 2][r] = (unsigned short)(h1 & 0xffff); th[cq + 3][r] = (unsigned short)(h1 >> 16);
            tl[cq + 0][r] = (unsigned short)(l0 & 0xffff); tl[cq + 1][r] = (unsigned short)(l0 >> 16);
            tl[cq + 2][r] = (unsigned short)(l1 & 0xffff); tl[cq + 3][r] = (unsigned short)(l1 >> 16);
        }
    }
    if (!Tr) return;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = t + 256 * i, cl = idx >> 3, sq = (idx & 7) * 8;   // 64 c-rows x 8 chunks of 8 bf16 (16 bytes)
        const int c = c0 + cl, hh = c / dh, d = c - hh * dh;
        const size_t o = (((size_t)b * NH + hh) * dh + d) * Sp + s0 + sq;
        *reinterpret_cast<uint4*>(Tr + o) = *reinterpret_cast<const uint4*>(&th[cl][sq]);
        *reinterpret_cast<uint4*>(Tr + plane + o) = *reinterpret_cast<const uint4*>(&tl[cl][sq]);
    }
}

// cp.async a [rows x cols] bf16 tile (cols * 2 bytes multiple of 16) from global (row pitch gp elements) to smem (pitch sp)
template <int ROWS, int COLS>
__device__ __forceinline__ void tile_async(__nv_bfloat16* smem, int sp, const __nv_bfloat16* gmem, size_t gp) {
    constexpr int CH = COLS / 8;   // 16-byte chunks per row
    for (int i = threadIdx.x; i < ROWS * CH; i += 128) {
        const int r = i / CH, c = (i % CH) * 8;
        cp_async16(smem + r * sp + c, gmem + (size_t)r * gp + c);
    }
}

// =====================================================================================================================
// forward: CTA = 4 warps, (64-query tile, head, utterance); warp owns 16 query rows; keys in tiles of 64, double-buffered
// =====================================================================================================================
template <int DH>
struct FwdSmem {
    static constexpr int KP = DH + 8, VP = 64 + 8;
    __nv_bfloat16 K[2][2][64][KP];    // [stage][plane][key][d]
    __nv_bfloat16 Vt[2][2][DH][VP];   // [stage][plane][d][key]
};

template <int DH>
__global__ void __launch_bounds__(128) attn_fwd_mma_kernel(AttnArgs p) {
    using SM = FwdSmem<DH>;
    constexpr int KS = DH / 16, ND = DH / 8;
    extern __shared__ __align__(16) uint8_t fwd_smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(fwd_smem_raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
    const int D = p.H * DH, NH = 3 * p.H, Sp = p.Sp;
    const int len = min((int)p.lens[b], p.S);
    float* ctx = p.ctx + (size_t)b * p.S * D + h * DH;
    float* lse = p.lse + ((size_t)b * p.H + h) * p.S;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;   // this lane's two query rows

    if (q0 >= len) {   // whole tile is padding: zeros (their only consumer masks them, model.py:259)
        for (int idx = tid; idx < 64 * DH; idx += 128) {
            const int r = idx / DH, c = idx % DH;
            if (q0 + r < p.S) ctx[(size_t)(q0 + r) * D + c] = 0.f;
        }
        if (tid < 64 && q0 + tid < p.S) lse[q0 + tid] = 0.f;
        return;
    }
    const size_t plane = (size_t)p.B * NH * Sp * DH;
    const __nv_bfloat16* Qr = p.R + ((size_t)b * NH + h) * Sp * DH;                 // [Sp][DH]
    const __nv_bfloat16* Kr = p.R + ((size_t)b * NH + p.H + h) * Sp * DH;
    const __nv_bfloat16* Vtr = p.Tr + ((size_t)b * NH + 2 * p.H + h) * DH * Sp;     // [DH][Sp]

    auto issue = [&](int stage, int k0) {
        tile_async<64, DH>(&sm.K[stage][0][0][0], SM::KP, Kr + (size_t)k0 * DH, DH);
        tile_async<64, DH>(&sm.K[stage][1][0][0], SM::KP, Kr + plane + (size_t)k0 * DH, DH);
        tile_async<DH, 64>(&sm.Vt[stage][0][0][0], SM::VP, Vtr + k0, Sp);
        tile_async<DH, 64>(&sm.Vt[stage][1][0][0], SM::VP, Vtr + plane + k0, Sp);
        cp_async_commit();
    };
    issue(0, 0);

    // Q fragments (already scaled and split), straight from the planes (rows < Sp always exist; rows >= S are zero)
    uint32_t qh[KS][4], ql[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int c = ks * 16 + 2 * t + 8 * half;
            qh[ks][2 * half] = *reinterpret_cast<const uint32_t*>(Qr + (size_t)r0 * DH + c);
            qh[ks][2 * half + 1] = *reinterpret_cast<const uint32_t*>(Qr + (size_t)r1 * DH + c);
            ql[ks][2 * half] = *reinterpret_cast<const uint32_t*>(Qr + plane + (size_t)r0 * DH + c);
            ql[ks][2 * half + 1] = *reinterpret_cast<const uint32_t*>(Qr + plane + (size_t)r1 * DH + c);
        }
    }

    float o[ND][4];
#pragma unroll
    for (int n = 0; n < ND; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const float inv_keep = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
    const unsigned long long bh = ((unsigned long long)b * p.H + h) * (unsigned long long)p.S;

    int stage = 0;
    for (int k0 = 0; k0 < len; k0 += 64, stage ^= 1) {
        cp_async_wait<0>();
        __syncthreads();                                  // tile k0 landed for everyone; everyone is done with tile k0 - 64
        if (k0 + 64 < len) issue(stage ^ 1, k0 + 64);     // prefetch the next tile while computing this one

        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int key = n * 8 + g, c = ks * 16 + 2 * t;
                const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&sm.K[stage][0][key][c]);
                const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&sm.K[stage][0][key][c + 8]);
                const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(&sm.K[stage][1][key][c]);
                const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(&sm.K[stage][1][key][c + 8]);
                mma_bf16(s[n], qh[ks], bh0, bh1);
                mma_bf16(s[n], ql[ks], bh0, bh1);
                mma_bf16(s[n], qh[ks], bl0, bl1);
            }
        }
        // key padding -> -inf, online softmax (rows g and g+8; a row's 64 scores live in the 4 lanes of a quad)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int key = k0 + n * 8 + 2 * t;
            if (key >= len) { s[n][0] = -INFINITY; s[n][2] = -INFINITY; }
            if (key + 1 >= len) { s[n][1] = -INFINITY; s[n][3] = -INFINITY; }
            mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
            mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every processed tile has >= 1 valid key
        const float c0 = fast_exp(m0 - mn0), c1 = fast_exp(m1 - mn1);
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            s[n][0] = fast_exp(s[n][0] - mn0); s[n][1] = fast_exp(s[n][1] - mn0);
            s[n][2] = fast_exp(s[n][2] - mn1); s[n][3] = fast_exp(s[n][3] - mn1);
            rs0 += s[n][0] + s[n][1];
            rs1 += s[n][2] + s[n][3];
            if (p.dropout_p > 0.f) {   // dropout on the attention weights: applied to what multiplies V, not to the row sum
                const unsigned long long kk = (unsigned long long)(k0 + n * 8 + 2 * t);
                const unsigned long long i0 = (bh + r0) * (unsigned long long)p.S + kk, i1 = (bh + r1) * (unsigned long long)p.S + kk;
                s[n][0] *= dropout_scale(p.seed, i0, p.dropout_p, inv_keep);
                s[n][1] *= dropout_scale(p.seed, i0 + 1, p.dropout_p, inv_keep);
                s[n][2] *= dropout_scale(p.seed, i1, p.dropout_p, inv_keep);
                s[n][3] *= dropout_scale(p.seed, i1 + 1, p.dropout_p, inv_keep);
            }
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
        m0 = mn0; m1 = mn1;
#pragma unroll
        for (int n = 0; n < ND; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }

        // O += P V : the C-fragments of two adjacent score n-tiles form one A-fragment (16 keys)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t ph[4], pl[4];
            split_pair(s[2 * j][0], s[2 * j][1], ph[0], pl[0]);
            split_pair(s[2 * j][2], s[2 * j][3], ph[1], pl[1]);
            split_pair(s[2 * j + 1][0], s[2 * j + 1][1], ph[2], pl[2]);
            split_pair(s[2 * j + 1][2], s[2 * j + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int n = 0; n < ND; ++n) {
                const int d = n * 8 + g, c = j * 16 + 2 * t;
                const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&sm.Vt[stage][0][d][c]);
                const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&sm.Vt[stage][0][d][c + 8]);
                const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(&sm.Vt[stage][1][d][c]);
                const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(&sm.Vt[stage][1][d][c + 8]);
                mma_bf16(o[n], ph, bh0, bh1);
                mma_bf16(o[n], pl, bh0, bh1);
                mma_bf16(o[n], ph, bl0, bl1);
            }
        }
    }

    const bool v0 = r0 < len, v1 = r1 < len;
    const float i0 = v0 ? 1.f / l0 : 0.f, i1 = v1 ? 1.f / l1 : 0.f;
#pragma unroll
    for (int n = 0; n < ND; ++n) {
        const int c = n * 8 + 2 * t;
        if (r0 < p.S) *reinterpret_cast<float2*>(ctx + (size_t)r0 * D + c) = make_float2(v0 ? o[n][0] * i0 : 0.f, v0 ? o[n][1] * i0 : 0.f);
        if (r1 < p.S) *reinterpret_cast<float2*>(ctx + (size_t)r1 * D + c) = make_float2(v1 ? o[n][2] * i1 : 0.f, v1 ? o[n][3] * i1 : 0.f);
    }
    if (t == 0) {
        if (r0 < p.S) lse[r0] = v0 ? m0 + logf(l0) : 0.f;
        if (r1 < p.S) lse[r1] = v1 ? m1 + logf(l1) : 0.f;
    }
}

// =====================================================================================================================
// backward: CTA = 4 warps, (64-key tile, head, utterance); warp w owns keys 16w..16w+15 and accumulates dK, dV for them in
// registers while the CTA walks the valid query tiles:
//   S^T = K Q^T, dP^T = V dO^T            (A = K / V rows of this warp, B = Q / dO tiles in smem)
//   P^T = exp(S^T - lse), dS^T = P^T (dP^T * drop - delta)
//   dV += P^T_drop dO, dK += dS^T Q       (score C-fragments re-used as A-fragments; B = dO^T / Q^T tiles in smem)
//   dQ += scale * dS K                    (dS staged in smem q-major; warp w takes query rows 16w..; fp32 atomics to HBM)
// The next query tile's four operand tiles are fetched with cp.async while the dQ product of the current one runs.
// =====================================================================================================================
template <int DH>
struct BwdSmem {
    static constexpr int RP = DH + 8;   // row pitch of [row][d] tiles
    static constexpr int TP = 64 + 8;   // row pitch of [d][row] and [q][key] tiles
    __nv_bfloat16 K[2][64][RP], V[2][64][RP], Kt[2][DH][TP];
    __nv_bfloat16 Q[2][64][RP], G[2][64][RP], Qt[2][DH][TP], Gt[2][DH][TP];
    __nv_bfloat16 dS[2][64][TP];
    float lse[64], delta[64];
};

// acc[n] (16 x 8 each, n < NT) += A(16 x 16*KSTEPS, rows a_row0.., from smem [row][k]) * B^T (B from smem [n][k])
template <int KSTEPS, int NT, int AP, int BP>
__device__ __forceinline__ void mma_smem_ab(float (*acc)[4], const __nv_bfloat16 (*A)[64][AP], int a_row0,
                                            const __nv_bfloat16 (*Bm)[64][BP], int g, int t) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
        uint32_t ah[4], al[4];
        const int c = ks * 16 + 2 * t;
        ah[0] = *reinterpret_cast<const uint32_t*>(&A[0][a_row0 + g][c]);
        ah[1] = *reinterpret_cast<const uint32_t*>(&A[0][a_row0 + g + 8][c]);
        ah[2] = *reinterpret_cast<const uint32_t*>(&A[0][a_row0 + g][c + 8]);
        ah[3] = *reinterpret_cast<const uint32_t*>(&A[0][a_row0 + g + 8][c + 8]);
        al[0] = *reinterpret_cast<const uint32_t*>(&A[1][a_row0 + g][c]);
        al[1] = *reinterpret_cast<const uint32_t*>(&A[1][a_row0 + g + 8][c]);
        al[2] = *reinterpret_cast<const uint32_t*>(&A[1][a_row0 + g][c + 8]);
        al[3] = *reinterpret_cast<const uint32_t*>(&A[1][a_row0 + g + 8][c + 8]);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&Bm[0][n * 8 + g][c]);
            const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&Bm[0][n * 8 + g][c + 8]);
            const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(&Bm[1][n * 8 + g][c]);
            const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(&Bm[1][n * 8 + g][c + 8]);
            mma_bf16(acc[n], ah, bh0, bh1);
            mma_bf16(acc[n], al, bh0, bh1);
            mma_bf16(acc[n], ah, bl0, bl1);
        }
    }
}

// acc[n] (n < NT) += A (from score C-fragments sc[8][4], 64 columns = 4 k-steps) * B^T (B from smem [n][k], pitch BP)
template <int NT, int BP, int BR>
__device__ __forceinline__ void mma_frag_b(float (*acc)[4], const float (*sc)[4], const __nv_bfloat16 (*Bm)[BR][BP], int g, int t) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t ah[4], al[4];
        split_pair(sc[2 * j][0], sc[2 * j][1], ah[0], al[0]);
        split_pair(sc[2 * j][2], sc[2 * j][3], ah[1], al[1]);
        split_pair(sc[2 * j + 1][0], sc[2 * j + 1][1], ah[2], al[2]);
        split_pair(sc[2 * j + 1][2], sc[2 * j + 1][3], ah[3], al[3]);
        const int c = j * 16 + 2 * t;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&Bm[0][n * 8 + g][c]);
            const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&Bm[0][n * 8 + g][c + 8]);
            const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(&Bm[1][n * 8 + g][c]);
            const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(&Bm[1][n * 8 + g][c + 8]);
            mma_bf16(acc[n], ah, bh0, bh1);
            mma_bf16(acc[n], al, bh0, bh1);
            mma_bf16(acc[n], ah, bl0, bl1);
        }
    }
}

template <int DH>
__global__ void __launch_bounds__(128) attn_bwd_mma_kernel(AttnArgs p) {
    using SM = BwdSmem<DH>;
    constexpr int KS = DH / 16, ND = DH / 8;
    extern __shared__ __align__(16) uint8_t bwd_smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(bwd_smem_raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int k0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
    const int D = p.H * DH, ld = 3 * D, NH = 3 * p.H, Sp = p.Sp;
    const int len = min((int)p.lens[b], p.S);
    if (k0 >= len) return;   // dK = dV = 0 for padded keys (dqkv is zero-initialised)
    float* dbase = p.dqkv + (size_t)b * p.S * ld + h * DH;
    const float* lse = p.lse + ((size_t)b * p.H + h) * p.S;
    const float* delta = p.delta + ((size_t)b * p.H + h) * p.S;
    const float scale = rsqrtf((float)DH);
    const float inv_keep = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
    const unsigned long long bh = ((unsigned long long)b * p.H + h) * (unsigned long long)p.S;

    const size_t plane = (size_t)p.B * NH * Sp * DH, gplane = (size_t)p.B * p.H * Sp * DH;
    const __nv_bfloat16* Qr = p.R + ((size_t)b * NH + h) * Sp * DH;
    const __nv_bfloat16* Kr = p.R + ((size_t)b * NH + p.H + h) * Sp * DH;
    const __nv_bfloat16* Vr = p.R + ((size_t)b * NH + 2 * p.H + h) * Sp * DH;
    const __nv_bfloat16* Qtr = p.Tr + ((size_t)b * NH + h) * DH * Sp;
    const __nv_bfloat16* Ktr = p.Tr + ((size_t)b * NH + p.H + h) * DH * Sp;
    const __nv_bfloat16* Gr = p.GR + ((size_t)b * p.H + h) * Sp * DH;
    const __nv_bfloat16* Gtr = p.GTr + ((size_t)b * p.H + h) * DH * Sp;

    auto issue_q = [&](int q0) {
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            tile_async<64, DH>(&sm.Q[pl][0][0], SM::RP, Qr + pl * plane + (size_t)q0 * DH, DH);
            tile_async<64, DH>(&sm.G[pl][0][0], SM::RP, Gr + pl * gplane + (size_t)q0 * DH, DH);
            tile_async<DH, 64>(&sm.Qt[pl][0][0], SM::TP, Qtr + pl * plane + q0, Sp);
            tile_async<DH, 64>(&sm.Gt[pl][0][0], SM::TP, Gtr + pl * gplane + q0, Sp);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
        tile_async<64, DH>(&sm.K[pl][0][0], SM::RP, Kr + pl * plane + (size_t)k0 * DH, DH);
        tile_async<64, DH>(&sm.V[pl][0][0], SM::RP, Vr + pl * plane + (size_t)k0 * DH, DH);
        tile_async<DH, 64>(&sm.Kt[pl][0][0], SM::TP, Ktr + pl * plane + k0, Sp);
    }
    issue_q(0);

    float dk[ND][4], dv[ND][4];
#pragma unroll
    for (int n = 0; n < ND; ++n) { dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f; dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f; }
    const int key0 = k0 + warp * 16 + g, key1 = key0 + 8;   // this lane's two key rows

    for (int q0 = 0; q0 < len; q0 += 64) {
        if (tid < 64) {
            sm.lse[tid] = (q0 + tid < len) ? lse[q0 + tid] : 0.f;
            sm.delta[tid] = (q0 + tid < len) ? delta[q0 + tid] : 0.f;
        }
        cp_async_wait<0>();
        __syncthreads();   // (A) this query tile's operands (and, first trip, K/V/Kt) landed; lse/delta visible

        float st[8][4], dp[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) { st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f; dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f; }
        mma_smem_ab<KS, 8, SM::RP, SM::RP>(st, sm.K, warp * 16, sm.Q, g, t);   // S^T  [16 keys][64 q]
        mma_smem_ab<KS, 8, SM::RP, SM::RP>(dp, sm.V, warp * 16, sm.G, g, t);   // dP^T [16 keys][64 q]

        // P^T (dropped) in st, dS^T in dp
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ql = n * 8 + 2 * t + (e & 1), q = q0 + ql;
                const int key = (e < 2) ? key0 : key1;
                float pd = 0.f, ds = 0.f;
                if (q < len && key < len) {
                    const float pv = fast_exp(st[n][e] - sm.lse[ql]);
                    float dm = 1.f;
                    if (p.dropout_p > 0.f) dm = dropout_scale(p.seed, (bh + q) * (unsigned long long)p.S + key, p.dropout_p, inv_keep);
                    pd = pv * dm;
                    ds = pv * (dp[n][e] * dm - sm.delta[ql]);
                }
                st[n][e] = pd;
                dp[n][e] = ds;
            }
        }
        mma_frag_b<ND, SM::TP, DH>(dv, st, sm.Gt, g, t);   // dV += P^T dO
        mma_frag_b<ND, SM::TP, DH>(dk, dp, sm.Qt, g, t);   // dK += dS^T Q   (Q carries 1/sqrt(dh))
        // dS, q-major, for the dQ product
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ql = n * 8 + 2 * t + (e & 1);
                const int kl = warp * 16 + g + ((e < 2) ? 0 : 8);
                const __nv_bfloat16 hi = __float2bfloat16_rn(dp[n][e]);
                sm.dS[0][ql][kl] = hi;
                sm.dS[1][ql][kl] = __float2bfloat16_rn(dp[n][e] - __bfloat162float(hi));
            }
        }
        __syncthreads();   // (B) dS complete; nobody reads Q/G/Qt/Gt of this tile any more
        if (q0 + 64 < len) issue_q(q0 + 64);   // overlaps with the dQ product below
        // dQ[16 q rows of this warp][DH] = dS[q][64 keys] K[64 keys][DH]  (B = K^T tile [d][key])
        float dq[ND][4];
#pragma unroll
        for (int n = 0; n < ND; ++n) { dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t ah[4], al[4];
            const int c = ks * 16 + 2 * t, r = warp * 16 + g;
            ah[0] = *reinterpret_cast<const uint32_t*>(&sm.dS[0][r][c]);
            ah[1] = *reinterpret_cast<const uint32_t*>(&sm.dS[0][r + 8][c]);
            ah[2] = *reinterpret_cast<const uint32_t*>(&sm.dS[0][r][c + 8]);
            ah[3] = *reinterpret_cast<const uint32_t*>(&sm.dS[0][r + 8][c + 8]);
            al[0] = *reinterpret_cast<const uint32_t*>(&sm.dS[1][r][c]);
            al[1] = *reinterpret_cast<const uint32_t*>(&sm.dS[1][r + 8][c]);
            al[2] = *reinterpret_cast<const uint32_t*>(&sm.dS[1][r][c + 8]);
            al[3] = *reinterpret_cast<const uint32_t*>(&sm.dS[1][r + 8][c + 8]);
#pragma unroll
            for (int n = 0; n < ND; ++n) {
                const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&sm.Kt[0][n * 8 + g][c]);
                const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&sm.Kt[0][n * 8 + g][c + 8]);
                const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(&sm.Kt[1][n * 8 + g][c]);
                const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(&sm.Kt[1][n * 8 + g][c + 8]);
                mma_bf16(dq[n], ah, bh0, bh1);
                mma_bf16(dq[n], al, bh0, bh1);
                mma_bf16(dq[n], ah, bl0, bl1);
            }
        }
        const int qa = q0 + warp * 16 + g, qb = qa + 8;
#pragma unroll
        for (int n = 0; n < ND; ++n) {
            const int c = n * 8 + 2 * t;
            if (qa < len) { atomicAdd(dbase + (size_t)qa * ld + c, scale * dq[n][0]); atomicAdd(dbase + (size_t)qa * ld + c + 1, scale * dq[n][1]); }
            if (qb < len) { atomicAdd(dbase + (size_t)qb * ld + c, scale * dq[n][2]); atomicAdd(dbase + (size_t)qb * ld + c + 1, scale * dq[n][3]); }
        }
        // the next trip's barrier (A) orders this trip's dS / lse reads before they are overwritten
    }
#pragma unroll
    for (int n = 0; n < ND; ++n) {
        const int c = n * 8 + 2 * t;
        if (key0 < len) {
            *reinterpret_cast<float2*>(dbase + (size_t)key0 * ld + D + c) = make_float2(dk[n][0], dk[n][1]);
            *reinterpret_cast<float2*>(dbase + (size_t)key0 * ld + 2 * D + c) = make_float2(dv[n][0], dv[n][1]);
        }
        if (key1 < len) {
            *reinterpret_cast<float2*>(dbase + (size_t)key1 * ld + D + c) = make_float2(dk[n][2], dk[n][3]);
            *reinterpret_cast<float2*>(dbase + (size_t)key1 * ld + 2 * D + c) = make_float2(dv[n][2], dv[n][3]);
        }
    }
}

template <int DH>
int launch_fwd_mma(const AttnArgs& a, cudaStream_t st) {
    const size_t smem = sizeof(FwdSmem<DH>);
    static bool configured = false;
    if (!configured) {
        DX_CUDA(cudaFuncSetAttribute(attn_fwd_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid(ceil_div(a.S, 64), a.H, a.B);
    attn_fwd_mma_kernel<DH><<<grid, 128, smem, st>>>(a);
    return check_launch("attn_fwd_mma");
}

template <int DH>
int launch_bwd_mma(const AttnArgs& a, cudaStream_t st) {
    const size_t smem = sizeof(BwdSmem<DH>);
    static bool configured = false;
    if (!configured) {
        DX_CUDA(cudaFuncSetAttribute(attn_bwd_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid(ceil_div(a.S, 64), a.H, a.B);
    attn_bwd_mma_kernel<DH><<<grid, 128, smem, st>>>(a);
    return check_launch("attn_bwd_mma");
}

int prep(const float* src, int ld, int B, int S, int Sp, int NH, int dh, int scale_cols, float scale, __nv_bfloat16* R,
         __nv_bfloat16* Tr, cudaStream_t st) {
    dim3 grid(Sp / 64, (NH * dh) / 64, B);
    attn_prep_kernel<<<grid, 256, 0, st>>>(src, ld, B, S, Sp, NH, dh, scale_cols, scale, R, Tr);
    return check_launch("attn_prep");
}

inline int pad64(int s) { return (s + 63) / 64 * 64; }

void bind_planes(AttnArgs& a, void* planes) {
    uint8_t* base = (uint8_t*)(((uintptr_t)planes + 255) & ~(uintptr_t)255);
    const size_t half = (size_t)2 * a.B * 3 * a.H * pad64(a.S) * a.dh * 2;
    a.Sp = pad64(a.S);
    a.R = (const __nv_bfloat16*)base;
    a.Tr = (const __nv_bfloat16*)(base + half);
}

}  // namespace

bool attention_mma_supported(const AttnArgs& a) {
    return (a.dh == 16 || a.dh == 32 || a.dh == 64) && ((a.H * a.dh) % 64 == 0);
}

// qkv planes: R and Tr, 2 planes each, [B][3H][Sp][dh] bf16
size_t attention_planes_bytes(int B, int S, int H, int dh) { return (size_t)4 * B * 3 * H * pad64(S) * dh * 2 + 256; }
// backward scratch: dO planes (R and Tr) + delta [B,H,S] fp32
size_t attention_bwd_scratch_bytes(int B, int S, int H, int dh) {
    return (size_t)4 * B * H * pad64(S) * dh * 2 + (size_t)B * H * S * 4 + 512;
}

int attention_fwd_mma(const AttnArgs& a_in, void* planes, cudaStream_t st) {
    AttnArgs a = a_in;
    DX_REQUIRE(planes != nullptr, "attention_fwd_mma: planes workspace required (dx_attention_planes_bytes)");
    bind_planes(a, planes);
    const int D = a.H * a.dh;
    int rc = prep(a.qkv, 3 * D, a.B, a.S, a.Sp, 3 * a.H, a.dh, D, rsqrtf((float)a.dh), (__nv_bfloat16*)a.R, (__nv_bfloat16*)a.Tr, st);
    if (rc) return rc;
    switch (a.dh) {
        case 16: return launch_fwd_mma<16>(a, st);
        case 32: return launch_fwd_mma<32>(a, st);
        case 64: return launch_fwd_mma<64>(a, st);
        default: set_last_error("attention_mma: unsupported head_dim %d", a.dh); return DX_ERR_UNSUPPORTED;
    }
}

// scratch holds the dO planes followed by delta [B,H,S]
int attention_bwd_mma(const AttnArgs& a_in, void* planes, void* scratch, cudaStream_t st) {
    AttnArgs a = a_in;
    DX_REQUIRE(planes != nullptr && scratch != nullptr, "attention_bwd_mma: planes and scratch required");
    bind_planes(a, planes);
    uint8_t* sb = (uint8_t*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    const size_t ghalf = (size_t)2 * a.B * a.H * a.Sp * a.dh * 2;
    a.GR = (const __nv_bfloat16*)sb;
    a.GTr = (const __nv_bfloat16*)(sb + ghalf);
    a.delta = (float*)(sb + 2 * ghalf);
    const int D = a.H * a.dh;
    int rc = attention_bwd_prepare(a, st);   // delta + zero dqkv
    if (rc) return rc;
    rc = prep(a.dctx, D, a.B, a.S, a.Sp, a.H, a.dh, 0, 1.f, (__nv_bfloat16*)a.GR, (__nv_bfloat16*)a.GTr, st);
    if (rc) return rc;
    switch (a.dh) {
        case 16: return launch_bwd_mma<16>(a, st);
        case 32: return launch_bwd_mma<32>(a, st);
        case 64: return launch_bwd_mma<64>(a, st);
        default: set_last_error("attention_mma: unsupported head_dim %d", a.dh); return DX_ERR_UNSUPPORTED;
    }
}

}  // namespace dx
