// fp32 CUDA-core GEMM family for channels-last Conv1d / Linear (forward, dgrad, wgrad).
//
// This is the EXACT-fp32 path ("parity mode"): every product and sum is IEEE fp32, so results agree with the
// reference's fp32 math to ~1e-6.  The tcgen05/TMEM path in gemm_tcgen05.cu computes the same contractions on the
// 5th-gen tensor cores (tf32 multiply, fp32 accumulate) and is the default for the large GEMMs.
//
// Conv1d over the padded [B, S, C] layout, reference model.py:86-94 (ConvNorm1D) — stride 1, zero 'same' padding at
// s = -1 and s = S only (no masking between convs: the halo leak of SURVEY.md §0.6 is reproduced by construction).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"

namespace dx {

constexpr int BM = 128, BN = 128, BK = 16, LDS = BM + 4;

// ---------------------------------------------------------------------------------------------------------------------
// y[b, s, n] = epi( alpha * sum_{tap, c} x[b, s + tap - pad, c] * w[tap][n][c] + bias[n] )
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_gemm_simt_kernel(ConvGemmArgs a) {
    __shared__ __align__(16) float As[BK][LDS];
    __shared__ __align__(16) float Bs[BK][LDS];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int s0 = blockIdx.x * BM, n0 = blockIdx.y * BN, b = blockIdx.z;
    const int pad = (a.KW - 1) / 2;
    const float* xb = a.x + (size_t)b * a.S * a.ldx;
    const bool vec = (a.Cin % BK == 0) && (a.ldx % 4 == 0);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < a.KW; ++tap) {
        const float* wt = a.w + (size_t)tap * a.Cout * a.Cin;
        for (int k0 = 0; k0 < a.Cin; k0 += BK) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = t + 256 * i, row = idx >> 2, kq = (idx & 3) * 4;
                const int sr = s0 + row + tap - pad;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int n = n0 + row;
                float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
                if (vec) {
                    if (sr >= 0 && sr < a.S) v = *reinterpret_cast<const float4*>(xb + (size_t)sr * a.ldx + k0 + kq);
                    if (n < a.Cout) u = *reinterpret_cast<const float4*>(wt + (size_t)n * a.Cin + k0 + kq);
                } else {
                    const int k = k0 + kq;
                    if (sr >= 0 && sr < a.S) {
                        const float* src = xb + (size_t)sr * a.ldx + k;
                        if (k + 0 < a.Cin) v.x = src[0];
                        if (k + 1 < a.Cin) v.y = src[1];
                        if (k + 2 < a.Cin) v.z = src[2];
                        if (k + 3 < a.Cin) v.w = src[3];
                    }
                    if (n < a.Cout) {
                        const float* src = wt + (size_t)n * a.Cin + k;
                        if (k + 0 < a.Cin) u.x = src[0];
                        if (k + 1 < a.Cin) u.y = src[1];
                        if (k + 2 < a.Cin) u.z = src[2];
                        if (k + 3 < a.Cin) u.w = src[3];
                    }
                }
                As[kq + 0][row] = v.x; As[kq + 1][row] = v.y; As[kq + 2][row] = v.z; As[kq + 3][row] = v.w;
                Bs[kq + 0][row] = u.x; Bs[kq + 1][row] = u.y; Bs[kq + 2][row] = u.z; Bs[kq + 3][row] = u.w;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int s = s0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (s >= a.S) continue;
        const size_t row = (size_t)b * a.S + s;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= a.Cout) continue;
            float v = a.alpha * acc[i][j];
            if (a.bias) v += a.bias[n];
            if (a.relu) v = fmaxf(v, 0.f);
            if (a.relu_src) v = a.relu_src[row * a.Cout + n] > 0.f ? v : 0.f;
            if (a.add_src) v += a.add_src[row * a.ldy + n];
            if (a.round_tf32) v = round_tf32(v);
            a.y[row * a.ldy + n] = v;
        }
    }
}

int conv_gemm_simt(const ConvGemmArgs& a, cudaStream_t st) {
    dim3 grid(ceil_div(a.S, BM), ceil_div(a.Cout, BN), a.B);
    conv_gemm_simt_kernel<<<grid, 256, 0, st>>>(a);
    return check_launch("conv_gemm_simt");
}

// ---------------------------------------------------------------------------------------------------------------------
// wgrad: part[split][tap][co][ci] = sum_{r in split} dy[r][co] * x[r shifted by tap][ci]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(ConvWgradArgs a, float* part, int rows_per_split) {
    __shared__ __align__(16) float As[BK][LDS];
    __shared__ __align__(16) float Bs[BK][LDS];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int m0 = blockIdx.x * BM;                      // co
    const int ntn = ceil_div(a.Cin, BN);
    const int n0 = (blockIdx.y % ntn) * BN;              // ci
    const int tap = blockIdx.y / ntn;
    const int split = blockIdx.z;
    const int pad = (a.KW - 1) / 2;
    const int R = a.B * a.S;
    const int r_begin = split * rows_per_split, r_end = min(R, r_begin + rows_per_split);
    const bool vec_a = (a.Cout % 4 == 0), vec_b = (a.Cin % 4 == 0) && (a.ldx % 4 == 0);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int r0 = r_begin; r0 < r_end; r0 += BK) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = t + 256 * i, k = idx >> 5, c4 = (idx & 31) * 4;
            const int r = r0 + k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f), u = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < r_end) {
                const int m = m0 + c4;
                const float* src = a.dy + (size_t)r * a.Cout + m;
                if (vec_a && m + 3 < a.Cout) v = *reinterpret_cast<const float4*>(src);
                else {
                    if (m + 0 < a.Cout) v.x = src[0];
                    if (m + 1 < a.Cout) v.y = src[1];
                    if (m + 2 < a.Cout) v.z = src[2];
                    if (m + 3 < a.Cout) v.w = src[3];
                }
                const int bb = r / a.S, s = r - bb * a.S, sx = s + tap - pad;
                if (sx >= 0 && sx < a.S) {
                    const int n = n0 + c4;
                    const float* xs = a.x + ((size_t)bb * a.S + sx) * a.ldx + n;
                    if (vec_b && n + 3 < a.Cin) u = *reinterpret_cast<const float4*>(xs);
                    else {
                        if (n + 0 < a.Cin) u.x = xs[0];
                        if (n + 1 < a.Cin) u.y = xs[1];
                        if (n + 2 < a.Cin) u.z = xs[2];
                        if (n + 3 < a.Cin) u.w = xs[3];
                    }
                }
            }
            *reinterpret_cast<float4*>(&As[k][c4]) = v;
            *reinterpret_cast<float4*>(&Bs[k][c4]) = u;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = part + ((size_t)split * a.KW + tap) * a.Cout * a.Cin;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= a.Cout) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n < a.Cin) out[(size_t)m * a.Cin + n] = acc[i][j];
        }
    }
}

// dw[co][ci][tap] (the parameter's own layout) = alpha * sum_split part[split][tap][co][ci]
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int nsplit, int KW, int Cout,
                                    int Cin, float alpha) {
    const size_t per = (size_t)KW * Cout * Cin;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < nsplit; ++k) s += part[(size_t)k * per + i];
        const int ci = (int)(i % Cin);
        const size_t q = i / Cin;
        const int co = (int)(q % Cout), tap = (int)(q / Cout);
        dw[((size_t)co * Cin + ci) * KW + tap] = alpha * s;
    }
}

// db[c] = alpha * sum_r dy[r][c]; one block = 32 columns x 8 row lanes, grid.y row chunks, atomics across chunks.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, float* __restrict__ db, int R, int C,
                                                     int rows_per_chunk, float alpha) {
    __shared__ float sm[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), lane_r = threadIdx.x >> 5;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(R, r0 + rows_per_chunk);
    float s = 0.f;
    if (c < C)
        for (int r = r0 + lane_r; r < r1; r += 8) s += dy[(size_t)r * C + c];
    sm[lane_r][threadIdx.x & 31] = s;
    __syncthreads();
    if (lane_r == 0 && c < C) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += sm[k][threadIdx.x];
        atomicAdd(db + c, alpha * tot);
    }
}

// db[c] = sum_r (hi[r][c] + lo[r][c]) over bf16 hi|lo planes (the bias gradient when only the planes of a dy exist).
// One block = 64 columns (bf16x2 per thread) x 8 row lanes.
__global__ void __launch_bounds__(256) colsum_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                                            float* __restrict__ db, int R, int C, int rows_per_chunk) {
    __shared__ float2 sm[8][33];
    const int c = blockIdx.x * 64 + 2 * (threadIdx.x & 31), lane_r = threadIdx.x >> 5;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(R, r0 + rows_per_chunk);
    float2 s = make_float2(0.f, 0.f);
    if (c < C) {
        for (int r = r0 + lane_r; r < r1; r += 8) {
            const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(hi + (size_t)r * C + c));
            const float2 l = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(lo + (size_t)r * C + c));
            s.x += h.x + l.x;
            s.y += h.y + l.y;
        }
    }
    sm[lane_r][threadIdx.x & 31] = s;
    __syncthreads();
    if (lane_r == 0 && c < C) {
        float2 tot = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { tot.x += sm[k][threadIdx.x].x; tot.y += sm[k][threadIdx.x].y; }
        atomicAdd(db + c, tot.x);
        atomicAdd(db + c + 1, tot.y);
    }
}

int colsum_planes(const void* planes, float* db, int R, int C, cudaStream_t st) {
    DX_REQUIRE(C % 2 == 0 && (((uintptr_t)planes) & 3) == 0, "colsum_planes: C must be even (C = %d)", C);
    DX_CUDA(cudaMemsetAsync(db, 0, (size_t)C * sizeof(float), st));
    const int chunks = max(1, min(ceil_div(R, 256), 148 * 8 / max(1, ceil_div(C, 64))));
    const int rpc = ceil_div(R, chunks);
    dim3 grid(ceil_div(C, 64), ceil_div(R, rpc));
    const __nv_bfloat16* hi = static_cast<const __nv_bfloat16*>(planes);
    colsum_planes_kernel<<<grid, 256, 0, st>>>(hi, hi + (size_t)R * C, db, R, C, rpc);
    return check_launch("colsum_planes");
}

size_t conv_wgrad_simt_workspace(const ConvWgradArgs& a, int* nsplit_out) {
    const int tiles = ceil_div(a.Cout, BM) * ceil_div(a.Cin, BN) * a.KW;
    const int R = a.B * a.S;
    int nsplit = ceil_div(148 * 3, tiles);
    nsplit = max(1, min(nsplit, ceil_div(R, 8 * BK)));
    if (nsplit_out) *nsplit_out = nsplit;
    return (size_t)nsplit * a.KW * a.Cout * a.Cin * sizeof(float);
}

int colsum(const float* dy, float* db, int R, int C, float alpha, cudaStream_t st) {
    DX_CUDA(cudaMemsetAsync(db, 0, (size_t)C * sizeof(float), st));
    const int chunks = max(1, min(ceil_div(R, 256), 148 * 4 / max(1, ceil_div(C, 32))));
    const int rpc = ceil_div(R, chunks);
    dim3 grid(ceil_div(C, 32), ceil_div(R, rpc));
    colsum_kernel<<<grid, 256, 0, st>>>(dy, db, R, C, rpc, alpha);
    return check_launch("colsum");
}

int conv_wgrad_simt(const ConvWgradArgs& a, cudaStream_t st) {
    int nsplit = 1;
    const size_t need = conv_wgrad_simt_workspace(a, &nsplit);
    DX_REQUIRE(a.workspace && a.workspace_bytes >= need, "conv_wgrad_simt: workspace %zu < %zu bytes", a.workspace_bytes, need);
    const int R = a.B * a.S;
    int rps = ceil_div(ceil_div(R, nsplit), BK) * BK;
    nsplit = ceil_div(R, rps);
    dim3 grid(ceil_div(a.Cout, BM), ceil_div(a.Cin, BN) * a.KW, nsplit);
    conv_wgrad_simt_kernel<<<grid, 256, 0, st>>>(a, (float*)a.workspace, rps);
    int rc = check_launch("conv_wgrad_simt");
    if (rc) return rc;
    const size_t per = (size_t)a.KW * a.Cout * a.Cin;
    const int blocks = grid_1d(per);
    wgrad_reduce_kernel<<<blocks, 256, 0, st>>>((const float*)a.workspace, a.dw, nsplit, a.KW, a.Cout, a.Cin, a.alpha);
    rc = check_launch("wgrad_reduce");
    if (rc) return rc;
    if (a.dbias) return colsum(a.dy, a.dbias, R, a.Cout, a.alpha, st);
    return DX_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// weight repacks: parameter layout [Cout][Cin][KW]  ->  fwd [KW][Cout][Cin]  and  dgrad [KW][Cin][Cout] (taps flipped)
// ---------------------------------------------------------------------------------------------------------------------
// optional fwd_planes / dgrad_planes: bf16 hi plane followed by lo plane (n elements each) in the same packed layouts
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ fwd, float* __restrict__ dgrad,
                                        __nv_bfloat16* __restrict__ fwd_planes, __nv_bfloat16* __restrict__ dgrad_planes,
                                        int Cout, int Cin, int KW, int round) {
    const size_t n = (size_t)Cout * Cin * KW;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int tap = (int)(i % KW);
        const size_t q = i / KW;
        const int ci = (int)(q % Cin), co = (int)(q / Cin);
        float v = w[i];
        if (round) v = round_tf32(v);
        const size_t of = ((size_t)tap * Cout + co) * Cin + ci, od = ((size_t)(KW - 1 - tap) * Cin + ci) * Cout + co;
        if (fwd) fwd[of] = v;
        if (dgrad) dgrad[od] = v;
        if (fwd_planes || dgrad_planes) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v), lo = __float2bfloat16_rn(v - __bfloat162float(hi));
            if (fwd_planes) { fwd_planes[of] = hi; fwd_planes[n + of] = lo; }
            if (dgrad_planes) { dgrad_planes[od] = hi; dgrad_planes[n + od] = lo; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// batched repack: ALL stale weights of the model in one launch (one optimiser step invalidates ~60 packs at once).
// One block = one 32 (co) x 32 (ci) tile of one weight, all taps, transposed through shared memory so that the source reads
// and both packed layouts' writes are coalesced.
// ---------------------------------------------------------------------------------------------------------------------
struct PackDesc {   // 64 bytes, mirrored by ops.py (struct '<5Q4i8x')
    const float* w;
    float* fwd;
    float* dgrad;
    __nv_bfloat16* fwd_planes;
    __nv_bfloat16* dgrad_planes;
    int Cout, Cin, KW, block0;   // block0: index of this weight's first block in the launch
    long long pad;
};
static_assert(sizeof(PackDesc) == 64, "PackDesc is part of the C-ABI (dx_pack_conv_weights_batched)");

__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const PackDesc* __restrict__ descs, int n_desc, int round) {
    __shared__ float tile[4][32][33];
    int lo = 0, hi = n_desc - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const PackDesc d = descs[lo];
    const int KW = d.KW, tiles_ci = (d.Cin + 31) / 32, local = (int)blockIdx.x - d.block0;
    const int co0 = (local / tiles_ci) * 32, ci0 = (local % tiles_ci) * 32;
    const int seg = min(32, d.Cin - ci0) * KW;   // contiguous source floats per output channel
    for (int idx = threadIdx.x; idx < 32 * 32 * KW; idx += 256) {
        const int r = idx / (32 * KW), j = idx - r * (32 * KW);
        if (co0 + r < d.Cout && j < seg) {
            float v = d.w[((size_t)(co0 + r) * d.Cin + ci0) * KW + j];
            if (round) v = round_tf32(v);
            tile[j % KW][r][j / KW] = v;
        }
    }
    __syncthreads();
    const size_t n = (size_t)d.Cout * d.Cin * KW;
    for (int idx = threadIdx.x; idx < KW * 1024; idx += 256) {   // fwd [tap][co][ci]: ci fastest
        const int tap = idx >> 10, r = (idx >> 5) & 31, c = idx & 31;
        const int co = co0 + r, ci = ci0 + c;
        if (co < d.Cout && ci < d.Cin) {
            const float v = tile[tap][r][c];
            const size_t of = ((size_t)tap * d.Cout + co) * d.Cin + ci;
            if (d.fwd) d.fwd[of] = v;
            if (d.fwd_planes) {
                const __nv_bfloat16 h = __float2bfloat16_rn(v);
                d.fwd_planes[of] = h;
                d.fwd_planes[n + of] = __float2bfloat16_rn(v - __bfloat162float(h));
            }
        }
    }
    for (int idx = threadIdx.x; idx < KW * 1024; idx += 256) {   // dgrad [KW-1-tap][ci][co]: co fastest
        const int tap = idx >> 10, c = (idx >> 5) & 31, r = idx & 31;
        const int co = co0 + r, ci = ci0 + c;
        if (co < d.Cout && ci < d.Cin) {
            const float v = tile[tap][r][c];
            const size_t od = ((size_t)(KW - 1 - tap) * d.Cin + ci) * d.Cout + co;
            if (d.dgrad) d.dgrad[od] = v;
            if (d.dgrad_planes) {
                const __nv_bfloat16 h = __float2bfloat16_rn(v);
                d.dgrad_planes[od] = h;
                d.dgrad_planes[n + od] = __float2bfloat16_rn(v - __bfloat162float(h));
            }
        }
    }
}

int pack_conv_weights_batched(const void* descs_device, int n_desc, int total_blocks, int round, cudaStream_t st) {
    DX_REQUIRE(descs_device != nullptr && n_desc > 0 && total_blocks > 0, "pack_conv_weights_batched: empty descriptor table");
    pack_weights_batched_kernel<<<total_blocks, 256, 0, st>>>(static_cast<const PackDesc*>(descs_device), n_desc, round);
    return check_launch("pack_conv_weights_batched");
}

int pack_conv_weight(const float* w, float* fwd, float* dgrad, void* fwd_planes, void* dgrad_planes, int Cout, int Cin, int KW,
                     int round, cudaStream_t st) {
    const size_t n = (size_t)Cout * Cin * KW;
    const int blocks = grid_1d(n);
    pack_conv_weight_kernel<<<blocks, 256, 0, st>>>(w, fwd, dgrad, (__nv_bfloat16*)fwd_planes, (__nv_bfloat16*)dgrad_planes, Cout, Cin,
                                                     KW, round);
    return check_launch("pack_conv_weight");
}

}  // namespace dx
