// Flash-style multi-head self-attention over the packed QKV projection (forward + backward), fp32.
//
// Reference: nn.MultiheadAttention math path as called at model.py:182-186 — q scaled by 1/sqrt(dh), key-padding -> -inf,
// softmax over keys, dropout on the attention weights, P·V.  The S x S score matrix is never written to HBM (the reference
// materialises [B*H, S, S] = 1.02 GB per prosody-encoder layer at S=1000); only ctx [B,S,D] and the
// per-row log-sum-exp [B,H,S] leave the SM.  Padded QUERY rows are skipped and written as zeros: their only consumer is
// LN(...).masked_fill(mask, 0) (model.py:191,259), so neither their values nor their gradients reach any valid output.
//
// Tile: 64 queries x 64 keys per step, 256 threads as a 16 x 16 grid; each thread owns a 4 x 4 score patch
// (rows ty*4+i, key columns tx+16*j) and a 4 x (dh/16) slice of the output accumulator.
#include "common.cuh"
#include "kernels.h"

namespace dx {

constexpr int TQ = 64, TK = 64;

__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// load a [64 x DH] tile of (q|k|v) rows r0.. from qkv (+col offset), rows >= limit -> 0, into smem [64][DH+1]
template <int DH>
__device__ __forceinline__ void load_tile(float (*dst)[DH + 1], const float* __restrict__ base, int ld, int r0, int limit,
                                          float scale) {
    constexpr int V = DH / 4;
    for (int idx = threadIdx.x; idx < 64 * V; idx += 256) {
        const int r = idx / V, c = (idx % V) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < limit) v = *reinterpret_cast<const float4*>(base + (size_t)(r0 + r) * ld + c);
        dst[r][c + 0] = v.x * scale; dst[r][c + 1] = v.y * scale; dst[r][c + 2] = v.z * scale; dst[r][c + 3] = v.w * scale;
    }
}

template <int DH>
__global__ void __launch_bounds__(256) attn_fwd_kernel(AttnArgs p) {
    const unsigned long long seed = dyn_seed(p.seed, p.dyn);
    constexpr int CN = DH / 16;
    extern __shared__ __align__(16) float smem[];
    float (*Qs)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(smem);
    float (*Ks)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(smem + 64 * (DH + 1));
    float (*Vs)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(smem + 2 * 64 * (DH + 1));
    float (*Ps)[TK + 1] = reinterpret_cast<float (*)[TK + 1]>(smem + 3 * 64 * (DH + 1));

    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
    const int D = p.H * DH, ld = 3 * D;
    const int len = min((int)p.lens[b], p.S);
    const float* base = p.qkv + (size_t)b * p.S * ld + h * DH;
    float* ctx = p.ctx + (size_t)b * p.S * D + h * DH;
    float* lse = p.lse + ((size_t)b * p.H + h) * p.S;

    if (q0 >= len) {  // whole tile is padding: zeros
        for (int idx = t; idx < TQ * DH; idx += 256) {
            const int r = idx / DH, c = idx % DH;
            if (q0 + r < p.S) ctx[(size_t)(q0 + r) * D + c] = 0.f;
        }
        if (t < TQ && q0 + t < p.S) lse[q0 + t] = 0.f;
        return;
    }
    const float scale = rsqrtf((float)DH);
    load_tile<DH>(Qs, base, ld, q0, len, scale);

    float m[4], l[4], o[4][CN];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
        for (int c = 0; c < CN; ++c) o[i][c] = 0.f;
    }
    const float inv_keep = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
    const unsigned long long bh = ((unsigned long long)b * p.H + h) * (unsigned long long)p.S;

    for (int k0 = 0; k0 < len; k0 += TK) {
        __syncthreads();  // previous tile fully consumed (also orders the Q load on the first trip)
        load_tile<DH>(Ks, base + D, ld, k0, len, 1.f);
        load_tile<DH>(Vs, base + 2 * D, ld, k0, len, 1.f);
        __syncthreads();
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
        for (int d = 0; d < DH; ++d) {
            float qv[4], kv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = Qs[ty * 4 + i][d];
#pragma unroll
            for (int j = 0; j < 4; ++j) kv[j] = Ks[tx + 16 * j][d];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + tx + 16 * j >= len) s[i][j] = -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
            mx = half_warp_max(mx);
            const float m_new = fmaxf(m[i], mx);  // finite: every processed tile has >= 1 valid key
            const float corr = expf(m[i] - m_new);
            float rs = 0.f;
            const int qrow = q0 + ty * 4 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float pv = expf(s[i][j] - m_new);
                rs += pv;
                if (p.dropout_p > 0.f)
                    pv *= dropout_scale(seed, (bh + qrow) * (unsigned long long)p.S + (k0 + tx + 16 * j), p.dropout_p, inv_keep);
                Ps[ty * 4 + i][tx + 16 * j] = pv;
            }
            rs = half_warp_sum(rs);
            l[i] = l[i] * corr + rs;
            m[i] = m_new;
#pragma unroll
            for (int c = 0; c < CN; ++c) o[i][c] *= corr;
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < TK; ++j) {
            float pv[4], vv[CN];
#pragma unroll
            for (int i = 0; i < 4; ++i) pv[i] = Ps[ty * 4 + i][j];
#pragma unroll
            for (int c = 0; c < CN; ++c) vv[c] = Vs[j][tx + 16 * c];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < CN; ++c) o[i][c] = fmaf(pv[i], vv[c], o[i][c]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ty * 4 + i;
        if (q >= p.S) continue;
        const bool valid = q < len;
        const float inv_l = valid ? 1.f / l[i] : 0.f;
#pragma unroll
        for (int c = 0; c < CN; ++c) ctx[(size_t)q * D + tx + 16 * c] = valid ? o[i][c] * inv_l : 0.f;
        if (tx == 0) lse[q] = valid ? m[i] + logf(l[i]) : 0.f;
    }
}

// delta[b,h,s] = sum_d dctx[b,s,h*dh+d] * ctx[b,s,h*dh+d]
__global__ void attn_delta_kernel(AttnArgs p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = p.B * p.S * p.H;
    if (idx >= total) return;
    const int h = idx % p.H, rs = idx / p.H;  // rs = b*S + s
    const int b = rs / p.S, s = rs - b * p.S;
    const int D = p.H * p.dh;
    const float* o = p.ctx + (size_t)rs * D + h * p.dh;
    const float* d = p.dctx + (size_t)rs * D + h * p.dh;
    float acc = 0.f;
    for (int c = 0; c < p.dh; c += 4) {
        const float4 a = *reinterpret_cast<const float4*>(o + c);
        const float4 g = *reinterpret_cast<const float4*>(d + c);
        acc += a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
    }
    p.delta[((size_t)b * p.H + h) * p.S + s] = acc;
}

// One block per (key tile, head, utterance): dK, dV accumulate in registers over all query tiles; dQ via fp32 atomics.
template <int DH>
__global__ void __launch_bounds__(256) attn_bwd_kernel(AttnArgs p) {
    const unsigned long long seed = dyn_seed(p.seed, p.dyn);
    constexpr int CN = DH / 16;
    extern __shared__ __align__(16) float smem[];
    float (*Ks)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(smem);
    float (*Vs)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(smem + 64 * (DH + 1));
    float (*Qs)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(smem + 2 * 64 * (DH + 1));
    float (*Gs)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(smem + 3 * 64 * (DH + 1));  // dctx tile
    float (*Ps)[TK + 1] = reinterpret_cast<float (*)[TK + 1]>(smem + 4 * 64 * (DH + 1));
    float (*Ss)[TK + 1] = reinterpret_cast<float (*)[TK + 1]>(smem + 4 * 64 * (DH + 1) + 64 * (TK + 1));  // dS tile
    float* Ls = smem + 4 * 64 * (DH + 1) + 2 * 64 * (TK + 1);  // lse[64]
    float* Ds = Ls + 64;                                        // delta[64]

    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int k0 = blockIdx.x * TK, h = blockIdx.y, b = blockIdx.z;
    const int D = p.H * DH, ld = 3 * D;
    const int len = min((int)p.lens[b], p.S);
    if (k0 >= len) return;  // dK = dV = 0 for padded keys (dqkv is zero-initialised)
    const float* base = p.qkv + (size_t)b * p.S * ld + h * DH;
    const float* dctx = p.dctx + (size_t)b * p.S * D + h * DH;
    float* dbase = p.dqkv + (size_t)b * p.S * ld + h * DH;
    const float* lse = p.lse + ((size_t)b * p.H + h) * p.S;
    const float* delta = p.delta + ((size_t)b * p.H + h) * p.S;
    const float scale = rsqrtf((float)DH);
    const float inv_keep = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
    const unsigned long long bh = ((unsigned long long)b * p.H + h) * (unsigned long long)p.S;

    load_tile<DH>(Ks, base + D, ld, k0, len, 1.f);
    load_tile<DH>(Vs, base + 2 * D, ld, k0, len, 1.f);
    float dk[4][CN], dv[4][CN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < CN; ++c) { dk[i][c] = 0.f; dv[i][c] = 0.f; }

    for (int q0 = 0; q0 < len; q0 += TQ) {
        __syncthreads();
        load_tile<DH>(Qs, base, ld, q0, len, scale);
        // dctx tile is [64 x DH] with row stride D
        {
            constexpr int V = DH / 4;
            for (int idx = t; idx < 64 * V; idx += 256) {
                const int r = idx / V, c = (idx % V) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q0 + r < len) v = *reinterpret_cast<const float4*>(dctx + (size_t)(q0 + r) * D + c);
                Gs[r][c + 0] = v.x; Gs[r][c + 1] = v.y; Gs[r][c + 2] = v.z; Gs[r][c + 3] = v.w;
            }
            if (t < 64) {
                Ls[t] = (q0 + t < len) ? lse[q0 + t] : 0.f;
                Ds[t] = (q0 + t < len) ? delta[q0 + t] : 0.f;
            }
        }
        __syncthreads();
        float s[4][4], dp[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[i][j] = 0.f; dp[i][j] = 0.f; }
#pragma unroll 4
        for (int d = 0; d < DH; ++d) {
            float qv[4], gv[4], kv[4], vv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { qv[i] = Qs[ty * 4 + i][d]; gv[i] = Gs[ty * 4 + i][d]; }
#pragma unroll
            for (int j = 0; j < 4; ++j) { kv[j] = Ks[tx + 16 * j][d]; vv[j] = Vs[tx + 16 * j][d]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
                    dp[i][j] = fmaf(gv[i], vv[j], dp[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int qi = ty * 4 + i, q = q0 + qi;
            const float li = Ls[qi], di = Ds[qi];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kj = tx + 16 * j, k = k0 + kj;
                float pv = 0.f, dsv = 0.f, pd = 0.f;
                if (q < len && k < len) {
                    pv = expf(s[i][j] - li);
                    float dm = 1.f;
                    if (p.dropout_p > 0.f) dm = dropout_scale(seed, (bh + q) * (unsigned long long)p.S + k, p.dropout_p, inv_keep);
                    pd = pv * dm;                          // dropped probabilities (what multiplied V)
                    dsv = pv * (dp[i][j] * dm - di);       // dS
                }
                Ps[qi][kj] = pd;
                Ss[qi][kj] = dsv;
            }
        }
        __syncthreads();
        // dV[k][d] += sum_q P[q][k] * dctx[q][d];  dK[k][d] += sum_q dS[q][k] * Qs[q][d]  (Qs already carries 1/sqrt(dh))
#pragma unroll 4
        for (int q = 0; q < TQ; ++q) {
            float pk[4], sk[4], gv[CN], qv[CN];
#pragma unroll
            for (int i = 0; i < 4; ++i) { pk[i] = Ps[q][ty * 4 + i]; sk[i] = Ss[q][ty * 4 + i]; }
#pragma unroll
            for (int c = 0; c < CN; ++c) { gv[c] = Gs[q][tx + 16 * c]; qv[c] = Qs[q][tx + 16 * c]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < CN; ++c) {
                    dv[i][c] = fmaf(pk[i], gv[c], dv[i][c]);
                    dk[i][c] = fmaf(sk[i], qv[c], dk[i][c]);
                }
        }
        // dQ[q][d] += scale * sum_k dS[q][k] * K[k][d]
        float dq[4][CN];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < CN; ++c) dq[i][c] = 0.f;
#pragma unroll 4
        for (int k = 0; k < TK; ++k) {
            float sv[4], kv[CN];
#pragma unroll
            for (int i = 0; i < 4; ++i) sv[i] = Ss[ty * 4 + i][k];
#pragma unroll
            for (int c = 0; c < CN; ++c) kv[c] = Ks[k][tx + 16 * c];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < CN; ++c) dq[i][c] = fmaf(sv[i], kv[c], dq[i][c]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = q0 + ty * 4 + i;
            if (q >= len) continue;
#pragma unroll
            for (int c = 0; c < CN; ++c) atomicAdd(dbase + (size_t)q * ld + tx + 16 * c, scale * dq[i][c]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty * 4 + i;
        if (k >= len) continue;
#pragma unroll
        for (int c = 0; c < CN; ++c) {
            dbase[(size_t)k * ld + D + tx + 16 * c] = dk[i][c];
            dbase[(size_t)k * ld + 2 * D + tx + 16 * c] = dv[i][c];
        }
    }
}

template <int DH>
static int launch_fwd(const AttnArgs& a, cudaStream_t st) {
    const size_t smem = (3 * 64 * (DH + 1) + 64 * (TK + 1)) * sizeof(float);
    DX_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(a.S, TQ), a.H, a.B);
    attn_fwd_kernel<DH><<<grid, 256, smem, st>>>(a);
    return check_launch("attn_fwd");
}

template <int DH>
static int launch_bwd(const AttnArgs& a, cudaStream_t st) {
    const size_t smem = (4 * 64 * (DH + 1) + 2 * 64 * (TK + 1) + 128) * sizeof(float);
    DX_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(a.S, TK), a.H, a.B);
    attn_bwd_kernel<DH><<<grid, 256, smem, st>>>(a);
    return check_launch("attn_bwd");
}

int attention_fwd(const AttnArgs& a, cudaStream_t st) {
    switch (a.dh) {
        case 16: return launch_fwd<16>(a, st);
        case 32: return launch_fwd<32>(a, st);
        case 64: return launch_fwd<64>(a, st);
        default: set_last_error("attention: unsupported head_dim %d (16, 32, 64)", a.dh); return DX_ERR_UNSUPPORTED;
    }
}

int attention_bwd_prepare(const AttnArgs& a, cudaStream_t st) {
    const int total = a.B * a.S * a.H;
    attn_delta_kernel<<<ceil_div(total, 256), 256, 0, st>>>(a);
    int rc = check_launch("attn_delta");
    if (rc) return rc;
    DX_CUDA(cudaMemsetAsync(a.dqkv, 0, (size_t)a.B * a.S * 3 * a.H * a.dh * sizeof(float), st));
    return DX_OK;
}

int attention_bwd(const AttnArgs& a, cudaStream_t st) {
    int rc = attention_bwd_prepare(a, st);
    if (rc) return rc;
    switch (a.dh) {
        case 16: return launch_bwd<16>(a, st);
        case 32: return launch_bwd<32>(a, st);
        case 64: return launch_bwd<64>(a, st);
        default: set_last_error("attention: unsupported head_dim %d (16, 32, 64)", a.dh); return DX_ERR_UNSUPPORTED;
    }
}

}  // namespace dx
