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
// exp(x) through ex2.approx (2^-22 relative): softmax inputs only.  exp2_fma(x, m2) = exp(x - m) with m2 = m * log2(e)
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_exp(float x) { return ex2(x * kLog2e); }
__device__ __forceinline__ float exp2_fma(float x, float m2) { return ex2(fmaf(x, kLog2e, -m2)); }

// ldmatrix / stmatrix: four 8x8 bf16 matrices per instruction; lane L passes the address of row L%8 of matrix L/8.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void stsm_x4_trans(void* p, const uint32_t (&r)[4]) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};"
                 ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// Lane -> (row, col) offsets of its ldmatrix.x4 address.
//   A operand (16 rows x 16 k, smem [row][k]):      matrices (rows 0-7,k 0-7) (rows 8-15,k 0-7) (rows 0-7,k 8-15) (rows 8-15,k 8-15)
//   B operand pair (2 n-tiles x 16 k, smem [n][k]): matrices (n 0-7,k 0-7) (n 0-7,k 8-15) (n 8-15,k 0-7) (n 8-15,k 8-15)
struct LaneMap {
    int a_r, a_c, b_r, b_c;
    __device__ __forceinline__ explicit LaneMap(int lane)
        : a_r(lane & 15), a_c((lane >> 4) << 3), b_r((lane & 7) + ((lane >> 4) << 3)), b_c(((lane >> 3) & 1) << 3) {}
};
//   B operand pair from a K-MAJOR tile (smem [k][n], e.g. V[key][d] for P*V): ldmatrix.trans with the A-style lane map,
//   matrices (k 0-7,n 0-7) (k 8-15,n 0-7) (k 0-7,n 8-15) (k 8-15,n 8-15) -> no transposed copies of any operand are needed.
// acc[2np], acc[2np+1] += A(hi|lo) * B^T with B = the n-tile pair at Bm[plane][16 np ..][k0 ..] (bf16x3: hi*hi + lo*hi + hi*lo)
template <int BR, int BP>
__device__ __forceinline__ void mma_pair(float (*acc)[4], int np, const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                         const __nv_bfloat16 (*Bm)[BR][BP], int k0, const LaneMap& lm) {
    uint32_t bh[4], bl[4];
    ldsm_x4(bh, &Bm[0][np * 16 + lm.b_r][k0 + lm.b_c]);
    ldsm_x4(bl, &Bm[1][np * 16 + lm.b_r][k0 + lm.b_c]);
    mma_bf16(acc[2 * np], ah, bh[0], bh[1]);
    mma_bf16(acc[2 * np], al, bh[0], bh[1]);
    mma_bf16(acc[2 * np], ah, bl[0], bl[1]);
    mma_bf16(acc[2 * np + 1], ah, bh[2], bh[3]);
    mma_bf16(acc[2 * np + 1], al, bh[2], bh[3]);
    mma_bf16(acc[2 * np + 1], ah, bl[2], bl[3]);
}
// same with B = T^T, T = the k-major tile T[plane][k0 ..][16 np ..]
template <int BR, int BP>
__device__ __forceinline__ void mma_pair_t(float (*acc)[4], int np, const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           const __nv_bfloat16 (*T)[BR][BP], int k0, const LaneMap& lm) {
    uint32_t bh[4], bl[4];
    ldsm_x4_trans(bh, &T[0][k0 + lm.a_r][np * 16 + lm.a_c]);
    ldsm_x4_trans(bl, &T[1][k0 + lm.a_r][np * 16 + lm.a_c]);
    mma_bf16(acc[2 * np], ah, bh[0], bh[1]);
    mma_bf16(acc[2 * np], al, bh[0], bh[1]);
    mma_bf16(acc[2 * np], ah, bl[0], bl[1]);
    mma_bf16(acc[2 * np + 1], ah, bh[2], bh[3]);
    mma_bf16(acc[2 * np + 1], al, bh[2], bh[3]);
    mma_bf16(acc[2 * np + 1], ah, bl[2], bl[3]);
}

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
                                                        __nv_bfloat16* __restrict__ Tr, const float* __restrict__ dot_with = nullptr,
                                                        float* __restrict__ dot_out = nullptr) {
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
        if (dot_out) {
            // backward: delta[b, head, s] = sum_d dctx * ctx, fused into the pass that already reads dctx (src): the dh / 4 lanes
            // that hold one (row, head) reduce with shuffles, the first of them writes (64-column blocks are head aligned)
            float part = 0.f;
            if (s < S) {
                const float4 o = *reinterpret_cast<const float4*>(dot_with + ((size_t)b * S + s) * ld + c);
                part = v.x * o.x + v.y * o.y + v.z * o.z + v.w * o.w;
            }
            for (int off = dh / 8; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
            if (s < S && (cq % dh) == 0) dot_out[((size_t)b * NH + c / dh) * S + s] = part;
        }
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
    static constexpr int KP = DH + 8;
    __nv_bfloat16 K[2][2][64][KP];    // [stage][plane][key][d]
    __nv_bfloat16 V[2][2][64][KP];    // [stage][plane][key][d]  (B operand of P*V through ldmatrix.trans)
};

template <int DH>
__global__ void __launch_bounds__(128) attn_fwd_mma_kernel(AttnArgs p) {
    const unsigned long long seed = dyn_seed(p.seed, p.dyn);
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

    // optional: ctx as bf16 hi|lo operand planes [2][B*S][D] for the out-projection GEMM
    __nv_bfloat16* cph = p.ctx_planes ? (__nv_bfloat16*)p.ctx_planes + (size_t)b * p.S * D + h * DH : nullptr;
    const size_t cplane = (size_t)p.B * p.S * D;
    if (q0 >= len) {   // whole tile is padding: zeros (their only consumer masks them, model.py:259)
        for (int idx = tid; idx < 64 * DH; idx += 128) {
            const int r = idx / DH, c = idx % DH;
            if (q0 + r < p.S) {
                ctx[(size_t)(q0 + r) * D + c] = 0.f;
                if (cph) {
                    cph[(size_t)(q0 + r) * D + c] = __float2bfloat16_rn(0.f);
                    cph[cplane + (size_t)(q0 + r) * D + c] = __float2bfloat16_rn(0.f);
                }
            }
        }
        if (tid < 64 && q0 + tid < p.S) lse[q0 + tid] = 0.f;
        return;
    }
    const size_t plane = (size_t)p.B * NH * Sp * DH;
    const __nv_bfloat16* Qr = p.R + ((size_t)b * NH + h) * Sp * DH;                 // [Sp][DH]
    const __nv_bfloat16* Kr = p.R + ((size_t)b * NH + p.H + h) * Sp * DH;
    const __nv_bfloat16* Vr = p.R + ((size_t)b * NH + 2 * p.H + h) * Sp * DH;

    auto issue = [&](int stage, int k0) {
        tile_async<64, DH>(&sm.K[stage][0][0][0], SM::KP, Kr + (size_t)k0 * DH, DH);
        tile_async<64, DH>(&sm.K[stage][1][0][0], SM::KP, Kr + plane + (size_t)k0 * DH, DH);
        tile_async<64, DH>(&sm.V[stage][0][0][0], SM::KP, Vr + (size_t)k0 * DH, DH);
        tile_async<64, DH>(&sm.V[stage][1][0][0], SM::KP, Vr + plane + (size_t)k0 * DH, DH);
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
    const bool drop = p.dropout_p > 0.f;
    const float inv_keep = drop ? 1.f / (1.f - p.dropout_p) : 1.f;   // folded into the final normalisation
    const unsigned long long bh = ((unsigned long long)b * p.H + h) * (unsigned long long)p.S;
    const uint32_t thresh = drop_threshold(p.dropout_p);
    const uint32_t rk0 = hash_u32(seed, bh + r0), rk1 = hash_u32(seed, bh + r1);   // per-row dropout keys
    const LaneMap lm(lane);

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
            for (int np = 0; np < 4; ++np) mma_pair<64, SM::KP>(s, np, qh[ks], ql[ks], sm.K[stage], ks * 16, lm);
        }
        // key padding -> -inf (last tile only), online softmax (rows g and g+8; a row's 64 scores live in the 4 lanes of a quad)
        if (k0 + 64 > len) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int key = k0 + n * 8 + 2 * t;
                if (key >= len) { s[n][0] = -INFINITY; s[n][2] = -INFINITY; }
                if (key + 1 >= len) { s[n][1] = -INFINITY; s[n][3] = -INFINITY; }
            }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
            mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every processed tile has >= 1 valid key
        const float c0 = fast_exp(m0 - mn0), c1 = fast_exp(m1 - mn1);
        const float mn0l = mn0 * kLog2e, mn1l = mn1 * kLog2e;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            s[n][0] = exp2_fma(s[n][0], mn0l); s[n][1] = exp2_fma(s[n][1], mn0l);
            s[n][2] = exp2_fma(s[n][2], mn1l); s[n][3] = exp2_fma(s[n][3], mn1l);
            rs0 += s[n][0] + s[n][1];
            rs1 += s[n][2] + s[n][3];
            if (drop) {   // dropout on the attention weights: applied to what multiplies V, not to the row sum
                const uint32_t ka = drop_col_term((uint32_t)(k0 + n * 8 + 2 * t)), kb = drop_col_term((uint32_t)(k0 + n * 8 + 2 * t + 1));
                s[n][0] = drop_keep(rk0, ka, thresh) ? s[n][0] : 0.f;
                s[n][1] = drop_keep(rk0, kb, thresh) ? s[n][1] : 0.f;
                s[n][2] = drop_keep(rk1, ka, thresh) ? s[n][2] : 0.f;
                s[n][3] = drop_keep(rk1, kb, thresh) ? s[n][3] : 0.f;
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
            for (int np = 0; np < ND / 2; ++np) mma_pair_t<64, SM::KP>(o, np, ph, pl, sm.V[stage], j * 16, lm);
        }
    }

    const bool v0 = r0 < len, v1 = r1 < len;
    const float i0 = v0 ? inv_keep / l0 : 0.f, i1 = v1 ? inv_keep / l1 : 0.f;
#pragma unroll
    for (int n = 0; n < ND; ++n) {
        const int c = n * 8 + 2 * t;
        const float a0 = v0 ? o[n][0] * i0 : 0.f, a1 = v0 ? o[n][1] * i0 : 0.f, b0 = v1 ? o[n][2] * i1 : 0.f, b1 = v1 ? o[n][3] * i1 : 0.f;
        if (r0 < p.S) *reinterpret_cast<float2*>(ctx + (size_t)r0 * D + c) = make_float2(a0, a1);
        if (r1 < p.S) *reinterpret_cast<float2*>(ctx + (size_t)r1 * D + c) = make_float2(b0, b1);
        if (cph) {
            uint32_t hi, lo;
            if (r0 < p.S) {
                split_pair(a0, a1, hi, lo);
                *reinterpret_cast<uint32_t*>(cph + (size_t)r0 * D + c) = hi;
                *reinterpret_cast<uint32_t*>(cph + cplane + (size_t)r0 * D + c) = lo;
            }
            if (r1 < p.S) {
                split_pair(b0, b1, hi, lo);
                *reinterpret_cast<uint32_t*>(cph + (size_t)r1 * D + c) = hi;
                *reinterpret_cast<uint32_t*>(cph + cplane + (size_t)r1 * D + c) = lo;
            }
        }
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
    static constexpr int RP = DH + 8;   // row pitch of the [row][d] operand tiles
    static constexpr int TP = 64 + 8;   // row pitch of the [q][key] dS tile
    __nv_bfloat16 K[2][64][RP], V[2][64][RP];   // [plane][key][d]
    __nv_bfloat16 Q[2][64][RP], G[2][64][RP];   // [plane][q][d]   (G = dO)
    __nv_bfloat16 dS[2][64][TP];
    float lse2[64], delta[64];   // lse * log2(e), delta
    uint32_t rk[64];             // per-query-row dropout keys
};

// acc[n] (16 x 8 each, n < NT) += A(16 x 16*KSTEPS, rows a_row0.., from smem [row][k]) * B^T (B from smem [n][k])
template <int KSTEPS, int NT, int AP, int BP>
__device__ __forceinline__ void mma_smem_ab(float (*acc)[4], const __nv_bfloat16 (*A)[64][AP], int a_row0,
                                            const __nv_bfloat16 (*Bm)[64][BP], const LaneMap& lm) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
        uint32_t ah[4], al[4];
        ldsm_x4(ah, &A[0][a_row0 + lm.a_r][ks * 16 + lm.a_c]);
        ldsm_x4(al, &A[1][a_row0 + lm.a_r][ks * 16 + lm.a_c]);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) mma_pair<64, BP>(acc, np, ah, al, Bm, ks * 16, lm);
    }
}

// acc[n] (n < NT) += A (from score C-fragments sc[8][4], 64 columns = 4 k-steps) * B (B = k-major tile in smem [k][n], pitch BP).
// With STAGE, the split fragments are also stored TRANSPOSED (stmatrix.trans) into dst[plane][64][TP] at columns col0.. :
// the C-fragment of a [16 keys][64 q] tile lands q-major, ready to be the A operand of the dQ product.
template <int NT, int BP, bool STAGE, int TP>
__device__ __forceinline__ void mma_frag_b(float (*acc)[4], const float (*sc)[4], const __nv_bfloat16 (*Bm)[64][BP], const LaneMap& lm,
                                           __nv_bfloat16 (*dst)[64][TP], int col0, int lane) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t ah[4], al[4];
        split_pair(sc[2 * j][0], sc[2 * j][1], ah[0], al[0]);
        split_pair(sc[2 * j][2], sc[2 * j][3], ah[1], al[1]);
        split_pair(sc[2 * j + 1][0], sc[2 * j + 1][1], ah[2], al[2]);
        split_pair(sc[2 * j + 1][2], sc[2 * j + 1][3], ah[3], al[3]);
        if (STAGE) {   // matrices: (n = 2j, keys +0) (n = 2j, keys +8) (n = 2j+1, keys +0) (n = 2j+1, keys +8)
            const int row = (2 * j + (lane >> 4)) * 8 + (lane & 7), col = col0 + (((lane >> 3) & 1) << 3);
            stsm_x4_trans(&dst[0][row][col], ah);
            stsm_x4_trans(&dst[1][row][col], al);
        }
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) mma_pair_t<64, BP>(acc, np, ah, al, Bm, j * 16, lm);
    }
}

template <int DH>
__global__ void __launch_bounds__(128, DH == 16 ? 4 : 2) attn_bwd_mma_kernel(AttnArgs p) {
    const unsigned long long seed = dyn_seed(p.seed, p.dyn);
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
    const bool drop = p.dropout_p > 0.f;
    const float inv_keep = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
    const unsigned long long bh = ((unsigned long long)b * p.H + h) * (unsigned long long)p.S;
    const uint32_t thresh = drop_threshold(p.dropout_p);
    const LaneMap lm(lane);

    const size_t plane = (size_t)p.B * NH * Sp * DH, gplane = (size_t)p.B * p.H * Sp * DH;
    const __nv_bfloat16* Qr = p.R + ((size_t)b * NH + h) * Sp * DH;
    const __nv_bfloat16* Kr = p.R + ((size_t)b * NH + p.H + h) * Sp * DH;
    const __nv_bfloat16* Vr = p.R + ((size_t)b * NH + 2 * p.H + h) * Sp * DH;
    const __nv_bfloat16* Gr = p.GR + ((size_t)b * p.H + h) * Sp * DH;

    auto issue_q = [&](int q0) {
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            tile_async<64, DH>(&sm.Q[pl][0][0], SM::RP, Qr + pl * plane + (size_t)q0 * DH, DH);
            tile_async<64, DH>(&sm.G[pl][0][0], SM::RP, Gr + pl * gplane + (size_t)q0 * DH, DH);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
        tile_async<64, DH>(&sm.K[pl][0][0], SM::RP, Kr + pl * plane + (size_t)k0 * DH, DH);
        tile_async<64, DH>(&sm.V[pl][0][0], SM::RP, Vr + pl * plane + (size_t)k0 * DH, DH);
    }
    issue_q(0);

    float dk[ND][4], dv[ND][4];
#pragma unroll
    for (int n = 0; n < ND; ++n) { dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f; dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f; }
    const int key0 = k0 + warp * 16 + g, key1 = key0 + 8;   // this lane's two key rows
    const uint32_t kc0 = drop_col_term((uint32_t)key0), kc1 = drop_col_term((uint32_t)key1);

    for (int q0 = 0; q0 < len; q0 += 64) {
        if (tid < 64) {
            sm.lse2[tid] = (q0 + tid < len) ? lse[q0 + tid] * kLog2e : 0.f;
            sm.delta[tid] = (q0 + tid < len) ? delta[q0 + tid] : 0.f;
            if (drop) sm.rk[tid] = hash_u32(seed, bh + q0 + tid);
        }
        cp_async_wait<0>();
        __syncthreads();   // (A) this query tile's operands (and, first trip, K/V/Kt) landed; lse/delta visible

        float st[8][4], dp[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) { st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f; dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f; }
        mma_smem_ab<KS, 8, SM::RP, SM::RP>(st, sm.K, warp * 16, sm.Q, lm);   // S^T  [16 keys][64 q]
        mma_smem_ab<KS, 8, SM::RP, SM::RP>(dp, sm.V, warp * 16, sm.G, lm);   // dP^T [16 keys][64 q]

        // P^T (dropped, WITHOUT the 1/keep factor: dV is scaled once at the end) in st, dS^T in dp
        const bool edge = (q0 + 64 > len) || (k0 + 64 > len);   // only edge tiles hold padded rows / keys
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int ql = n * 8 + 2 * t;
            const float2 ls = *reinterpret_cast<const float2*>(&sm.lse2[ql]);
            const float2 dl = *reinterpret_cast<const float2*>(&sm.delta[ql]);
            uint2 rk = make_uint2(0u, 0u);
            if (drop) rk = *reinterpret_cast<const uint2*>(&sm.rk[ql]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool odd = e & 1, second = e >= 2;
                float pv = exp2_fma(st[n][e], odd ? ls.y : ls.x);
                if (edge && !(q0 + ql + (e & 1) < len && (second ? key1 : key0) < len)) pv = 0.f;
                const bool keep = !drop || drop_keep(odd ? rk.y : rk.x, second ? kc1 : kc0, thresh);
                st[n][e] = keep ? pv : 0.f;
                dp[n][e] = pv * ((keep ? dp[n][e] * inv_keep : 0.f) - (odd ? dl.y : dl.x));
            }
        }
        mma_frag_b<ND, SM::RP, false, SM::TP>(dv, st, sm.G, lm, sm.dS, 0, lane);          // dV += P^T dO
        mma_frag_b<ND, SM::RP, true, SM::TP>(dk, dp, sm.Q, lm, sm.dS, warp * 16, lane);   // dK += dS^T Q (Q carries 1/sqrt(dh)); dS staged q-major
        __syncthreads();   // (B) dS complete; nobody reads Q/G/Qt/Gt of this tile any more
        if (q0 + 64 < len) issue_q(q0 + 64);   // overlaps with the dQ product below
        // dQ[16 q rows of this warp][DH] = dS[q][64 keys] K[64 keys][DH]  (B = the K tile itself, k-major)
        float dq[ND][4];
#pragma unroll
        for (int n = 0; n < ND; ++n) { dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t ah[4], al[4];
            ldsm_x4(ah, &sm.dS[0][warp * 16 + lm.a_r][ks * 16 + lm.a_c]);
            ldsm_x4(al, &sm.dS[1][warp * 16 + lm.a_r][ks * 16 + lm.a_c]);
#pragma unroll
            for (int np = 0; np < ND / 2; ++np) mma_pair_t<64, SM::RP>(dq, np, ah, al, sm.K, ks * 16, lm);
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
            *reinterpret_cast<float2*>(dbase + (size_t)key0 * ld + 2 * D + c) = make_float2(dv[n][0] * inv_keep, dv[n][1] * inv_keep);
        }
        if (key1 < len) {
            *reinterpret_cast<float2*>(dbase + (size_t)key1 * ld + D + c) = make_float2(dk[n][2], dk[n][3]);
            *reinterpret_cast<float2*>(dbase + (size_t)key1 * ld + 2 * D + c) = make_float2(dv[n][2] * inv_keep, dv[n][3] * inv_keep);
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
         __nv_bfloat16* Tr, cudaStream_t st, const float* dot_with = nullptr, float* dot_out = nullptr) {
    dim3 grid(Sp / 64, (NH * dh) / 64, B);
    attn_prep_kernel<<<grid, 256, 0, st>>>(src, ld, B, S, Sp, NH, dh, scale_cols, scale, R, Tr, dot_with, dot_out);
    return check_launch("attn_prep");
}

// attention backend selection (dx_set_attention_backend; initial values from DX_ATTN_TC / DX_ATTN_BWD_TC, default tcgen05)
int g_attn_fwd_tc = -1, g_attn_bwd_tc = -1;
int env_flag(const char* name) {
    const char* e = getenv(name);
    return (e && atoi(e) == 0) ? 0 : 1;
}
bool attn_bwd_tc_enabled(int dh) {
    (void)dh;
    if (g_attn_bwd_tc < 0) g_attn_bwd_tc = env_flag("DX_ATTN_BWD_TC");
    return g_attn_bwd_tc == 1;
}

inline int pad64(int s) { return (s + 63) / 64 * 64; }

bool attn_tc_enabled() {
    if (g_attn_fwd_tc < 0) g_attn_fwd_tc = env_flag("DX_ATTN_TC");
    return g_attn_fwd_tc == 1;
}

void bind_planes(AttnArgs& a, void* planes) {
    uint8_t* base = (uint8_t*)(((uintptr_t)planes + 255) & ~(uintptr_t)255);
    a.Sp = pad64(a.S);
    a.R = (const __nv_bfloat16*)base;
    a.Tr = nullptr;   // no transposed planes: the kernels transpose on the fly with ldmatrix.trans
}

}  // namespace

void set_attention_backend(int fwd_tc, int bwd_tc) {
    g_attn_fwd_tc = fwd_tc;
    g_attn_bwd_tc = bwd_tc;
}

bool attention_mma_supported(const AttnArgs& a) {
    return (a.dh == 16 || a.dh == 32 || a.dh == 64) && ((a.H * a.dh) % 64 == 0);
}

// qkv planes: hi|lo, [2][B][3H][Sp][dh] bf16
size_t attention_planes_bytes(int B, int S, int H, int dh) { return (size_t)2 * B * 3 * H * pad64(S) * dh * 2 + 256; }
// backward scratch: dO planes (hi|lo) + delta [B,H,S] fp32
size_t attention_bwd_scratch_bytes(int B, int S, int H, int dh) {
    return (size_t)2 * B * H * pad64(S) * dh * 2 + (size_t)B * H * S * 4 + 512;
}

int attention_fwd_mma(const AttnArgs& a_in, void* planes, cudaStream_t st) {
    AttnArgs a = a_in;
    DX_REQUIRE(planes != nullptr, "attention_fwd_mma: planes workspace required (dx_attention_planes_bytes)");
    bind_planes(a, planes);
    const int D = a.H * a.dh;
    // qkv == NULL: the in-projection GEMM already wrote the operand planes itself (dx_inproj_head_planes): no fp32 qkv exists
    int rc = a.qkv ? prep(a.qkv, 3 * D, a.B, a.S, a.Sp, 3 * a.H, a.dh, D, rsqrtf((float)a.dh), (__nv_bfloat16*)a.R, (__nv_bfloat16*)a.Tr, st)
                   : DX_OK;
    if (rc) return rc;
    if (attn_tc_enabled() && attention_fwd_tc_supported(a)) return attention_fwd_tc(a, st);   // tcgen05 / TMEM forward
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
    a.GTr = nullptr;
    a.delta = (float*)(sb + ghalf);
    const int D = a.H * a.dh;
    DX_CUDA(cudaMemsetAsync(a.dqkv, 0, (size_t)a.B * a.S * 3 * D * sizeof(float), st));   // dq is accumulated with atomics
    // dO planes and delta = rowsum(dO * O) per head in ONE pass over dctx
    // dctx == NULL: the out-projection's input-gradient GEMM already wrote the dO planes and delta itself (dx_outproj_dgrad_head_planes)
    int rc = a.dctx ? prep(a.dctx, D, a.B, a.S, a.Sp, a.H, a.dh, 0, 1.f, (__nv_bfloat16*)a.GR, (__nv_bfloat16*)a.GTr, st, a.ctx, a.delta) : DX_OK;
    if (rc) return rc;
    if (attn_bwd_tc_enabled(a.dh) && attention_bwd_tc_supported(a)) return attention_bwd_tc(a, st);   // tcgen05 / TMEM backward
    switch (a.dh) {
        case 16: return launch_bwd_mma<16>(a, st);
        case 32: return launch_bwd_mma<32>(a, st);
        case 64: return launch_bwd_mma<64>(a, st);
        default: set_last_error("attention_mma: unsupported head_dim %d", a.dh); return DX_ERR_UNSUPPORTED;
    }
}

}  // namespace dx
