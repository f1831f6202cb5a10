// Small fused HBM-bound kernels around the GEMMs: embeddings + positional encoding + masks, scalar conv embeddings,
// pooling, FiLM assembly, narrow projections, the masked mel transpose.  All operate on channels-last [B, S, D] fp32.
#include "common.cuh"
#include "kernels.h"

namespace dx {

// ---------------------------------------------------------------------------------------------------------------------
// phoneme embedding + positional encoding + mask      reference model.py:497-504 (PE = rows 0..len-1 of the table)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void embed_pe_fwd_kernel(const long long* __restrict__ sym, const long long* __restrict__ lens,
                                    const float* __restrict__ emb, const float* __restrict__ pe, float* __restrict__ y, int B,
                                    int L, int D, int n_symbols) {
    const int V = D / 4;
    const size_t total = (size_t)B * L * V;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % V) * 4;
        const size_t r = i / V;
        const int b = (int)(r / L), s = (int)(r % L);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < (int)lens[b]) {
            long long id = sym[r];
            id = id < 0 ? 0 : (id >= n_symbols ? n_symbols - 1 : id);
            const float4 e = *reinterpret_cast<const float4*>(emb + (size_t)id * D + c);
            const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)s * D + c);
            o = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
        }
        *reinterpret_cast<float4*>(y + r * D + c) = o;
    }
}

__global__ void embed_pe_bwd_kernel(const long long* __restrict__ sym, const long long* __restrict__ lens,
                                    const float* __restrict__ dy, float* __restrict__ demb, int B, int L, int D, int n_symbols) {
    const size_t total = (size_t)B * L * D;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % D);
        const size_t r = i / D;
        const int b = (int)(r / L), s = (int)(r % L);
        if (s < (int)lens[b]) {
            long long id = sym[r];
            id = id < 0 ? 0 : (id >= n_symbols ? n_symbols - 1 : id);
            atomicAdd(demb + (size_t)id * D + c, dy[i]);
        }
    }
}

int embed_pe_fwd(const long long* symbols, const long long* lens, const float* emb, const float* pe, float* y, int B, int L,
                 int D, int n_symbols, cudaStream_t st) {
    DX_REQUIRE(D % 4 == 0, "embed_pe: D %% 4 != 0");
    const size_t total = (size_t)B * L * (D / 4);
    embed_pe_fwd_kernel<<<grid_1d(total), 256, 0, st>>>(symbols, lens, emb, pe, y, B, L, D, n_symbols);
    return check_launch("embed_pe_fwd");
}

int embed_pe_bwd(const long long* symbols, const long long* lens, const float* dy, float* demb, int B, int L, int D,
                 int n_symbols, cudaStream_t st) {
    DX_CUDA(cudaMemsetAsync(demb, 0, (size_t)n_symbols * D * sizeof(float), st));
    const size_t total = (size_t)B * L * D;
    embed_pe_bwd_kernel<<<grid_1d(total), 256, 0, st>>>(symbols, lens, dy, demb, B, L, D, n_symbols);
    return check_launch("embed_pe_bwd");
}

// ---------------------------------------------------------------------------------------------------------------------
// frame-side input: y = mask * (x + PE + Conv1d(1->D,k3)(energy) + Conv1d(1->D,k3)(pitch))
// reference model.py:400-414 (prosody encoder) and model.py:696-701 (frame decoder: no scalar convs)
// Scalar conv: out[t, c] = sum_tap w[c][0][tap] * u[t + tap - 1] + bias[c]  (zero padding at t=-1 and t=T)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void frame_input_fwd_kernel(const float* __restrict__ x, const float* __restrict__ e, const float* __restrict__ f0,
                                       const float* __restrict__ we, const float* __restrict__ be, const float* __restrict__ wp,
                                       const float* __restrict__ bp, const float* __restrict__ pe,
                                       const long long* __restrict__ lens, float* __restrict__ y, int B, int T, int D) {
    const size_t total = (size_t)B * T * D;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % D);
        const size_t r = i / D;
        const int b = (int)(r / T), t = (int)(r % T);
        float o = 0.f;
        if (t < (int)lens[b]) {
            o = x[i] + pe[(size_t)t * D + c];
            if (e) {
                const float* eb = e + (size_t)b * T;
                const float* fb = f0 + (size_t)b * T;
                const float em = t > 0 ? eb[t - 1] : 0.f, ep = t + 1 < T ? eb[t + 1] : 0.f;
                const float fm = t > 0 ? fb[t - 1] : 0.f, fp = t + 1 < T ? fb[t + 1] : 0.f;
                o += we[c * 3 + 0] * em + we[c * 3 + 1] * eb[t] + we[c * 3 + 2] * ep + be[c];
                o += wp[c * 3 + 0] * fm + wp[c * 3 + 1] * fb[t] + wp[c * 3 + 2] * fp + bp[c];
            }
        }
        y[i] = o;
    }
}

// dx = mask * dy;  dw{e,p}[c][tap] = sum_{b,t valid} dy[b,t,c] * u[b, t+tap-1];  db[c] = sum dy
// Block = 32 channels x 8 row lanes over a chunk of frames of one utterance; atomics across blocks.
__global__ void __launch_bounds__(256) frame_input_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ e,
                                                              const float* __restrict__ f0, const long long* __restrict__ lens,
                                                              float* __restrict__ dx, float* __restrict__ dwe,
                                                              float* __restrict__ dbe, float* __restrict__ dwp,
                                                              float* __restrict__ dbp, int B, int T, int D, int rows_per_chunk) {
    __shared__ float sm[7][8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl, b = blockIdx.z;
    const int len = min((int)lens[b], T);
    const int t0 = blockIdx.y * rows_per_chunk, t1 = min(T, t0 + rows_per_chunk);
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < D) {
        for (int t = t0 + rl; t < t1; t += 8) {
            const size_t i = ((size_t)b * T + t) * D + c;
            const float g = t < len ? dy[i] : 0.f;
            dx[i] = g;
            if (e && t < len) {
                const float* eb = e + (size_t)b * T;
                const float* fb = f0 + (size_t)b * T;
                acc[0] += g * (t > 0 ? eb[t - 1] : 0.f);
                acc[1] += g * eb[t];
                acc[2] += g * (t + 1 < T ? eb[t + 1] : 0.f);
                acc[3] += g * (t > 0 ? fb[t - 1] : 0.f);
                acc[4] += g * fb[t];
                acc[5] += g * (t + 1 < T ? fb[t + 1] : 0.f);
                acc[6] += g;
            }
        }
    }
    if (!e) return;
#pragma unroll
    for (int k = 0; k < 7; ++k) sm[k][rl][cl] = acc[k];
    __syncthreads();
    if (rl < 7 && c < D) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += sm[rl][k][cl];
        if (rl < 3) atomicAdd(dwe + c * 3 + rl, tot);
        else if (rl < 6) atomicAdd(dwp + c * 3 + (rl - 3), tot);
        else { atomicAdd(dbe + c, tot); atomicAdd(dbp + c, tot); }
    }
}

int frame_input_fwd(const float* x, const float* e, const float* f0, const float* we, const float* be, const float* wp,
                    const float* bp, const float* pe, const long long* lens, float* y, int B, int T, int D, cudaStream_t st) {
    const size_t total = (size_t)B * T * D;
    frame_input_fwd_kernel<<<grid_1d(total), 256, 0, st>>>(x, e, f0, we, be, wp, bp, pe, lens, y, B, T, D);
    return check_launch("frame_input_fwd");
}

int frame_input_bwd(const float* dy, const float* e, const float* f0, const long long* lens, float* dx, float* dwe, float* dbe,
                    float* dwp, float* dbp, int B, int T, int D, cudaStream_t st) {
    if (e) {
        DX_CUDA(cudaMemsetAsync(dwe, 0, (size_t)D * 3 * sizeof(float), st));
        DX_CUDA(cudaMemsetAsync(dwp, 0, (size_t)D * 3 * sizeof(float), st));
        DX_CUDA(cudaMemsetAsync(dbe, 0, (size_t)D * sizeof(float), st));
        DX_CUDA(cudaMemsetAsync(dbp, 0, (size_t)D * sizeof(float), st));
    }
    const int col_blocks = ceil_div(D, 32);
    const int chunks = max(1, min(ceil_div(T, 64), ceil_div(148 * 4, col_blocks * B)));
    const int rpc = ceil_div(T, chunks);
    dim3 grid(col_blocks, ceil_div(T, rpc), B);
    frame_input_bwd_kernel<<<grid, 256, 0, st>>>(dy, e, f0, lens, dx, dwe, dbe, dwp, dbp, B, T, D, rpc);
    return check_launch("frame_input_bwd");
}

// ---------------------------------------------------------------------------------------------------------------------
// mean pooling over time: pooled[b, c] = sum_s x[b, s, c] / len[b]        reference model.py:419
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) meanpool_fwd_kernel(const float* __restrict__ x, const long long* __restrict__ lens,
                                                           float* __restrict__ pooled, int B, int S, int D) {
    __shared__ float sm[8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl, b = blockIdx.y;
    float acc = 0.f;
    if (c < D)
        for (int s = rl; s < S; s += 8) acc += x[((size_t)b * S + s) * D + c];
    sm[rl][cl] = acc;
    __syncthreads();
    if (rl == 0 && c < D) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += sm[k][cl];
        pooled[(size_t)b * D + c] = tot / (float)lens[b];
    }
}

__global__ void meanpool_bwd_kernel(const float* __restrict__ dpooled, const long long* __restrict__ lens, float* __restrict__ dx,
                                    int B, int S, int D) {
    const size_t total = (size_t)B * S * D;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % D);
        const int b = (int)(i / ((size_t)S * D));
        dx[i] = dpooled[(size_t)b * D + c] / (float)lens[b];
    }
}

int meanpool_fwd(const float* x, const long long* lens, float* pooled, int B, int S, int D, cudaStream_t st) {
    dim3 grid(ceil_div(D, 32), B);
    meanpool_fwd_kernel<<<grid, 256, 0, st>>>(x, lens, pooled, B, S, D);
    return check_launch("meanpool_fwd");
}

int meanpool_bwd(const float* dpooled, const long long* lens, float* dx, int B, int S, int D, cudaStream_t st) {
    const size_t total = (size_t)B * S * D;
    meanpool_bwd_kernel<<<grid_1d(total), 256, 0, st>>>(dpooled, lens, dx, B, S, D);
    return check_launch("meanpool_bwd");
}

// ---------------------------------------------------------------------------------------------------------------------
// speaker embedding add                                                         reference model.py:423-424
// ---------------------------------------------------------------------------------------------------------------------
__global__ void add_speaker_fwd_kernel(const float* __restrict__ pooled, const long long* __restrict__ spk,
                                       const float* __restrict__ emb, float* __restrict__ h, int B, int D, int n_spk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * D) return;
    const int b = i / D, c = i % D;
    long long id = spk[b];
    id = id < 0 ? 0 : (id >= n_spk ? n_spk - 1 : id);
    h[i] = pooled[i] + emb[(size_t)id * D + c];
}
__global__ void add_speaker_bwd_kernel(const float* __restrict__ dh, const long long* __restrict__ spk, float* __restrict__ demb,
                                       int B, int D, int n_spk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * D) return;
    const int b = i / D, c = i % D;
    long long id = spk[b];
    id = id < 0 ? 0 : (id >= n_spk ? n_spk - 1 : id);
    atomicAdd(demb + (size_t)id * D + c, dh[i]);
}
int add_speaker_fwd(const float* pooled, const long long* spk, const float* spk_emb, float* h, int B, int D, int n_spk,
                    cudaStream_t st) {
    add_speaker_fwd_kernel<<<ceil_div(B * D, 256), 256, 0, st>>>(pooled, spk, spk_emb, h, B, D, n_spk);
    return check_launch("add_speaker_fwd");
}
int add_speaker_bwd(const float* dh, const long long* spk, float* dspk_emb, int B, int D, int n_spk, cudaStream_t st) {
    DX_CUDA(cudaMemsetAsync(dspk_emb, 0, (size_t)n_spk * D * sizeof(float), st));
    add_speaker_bwd_kernel<<<ceil_div(B * D, 256), 256, 0, st>>>(dh, spk, dspk_emb, B, D, n_spk);
    return check_launch("add_speaker_bwd");
}

// ---------------------------------------------------------------------------------------------------------------------
// FiLM assembly                                                               reference model.py:430-461
// raw gammas/betas [B, NF] (NF = sum nb*ch) -> film [B, 2*NF]: per module m, per block k: (gamma[ch] | beta[ch]) with
//   gamma = post[0][blk] * graw + 1,  beta = post[1][blk] * braw     (post == nullptr: gamma = graw + 1, beta = braw)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void film_locate(const FilmLayout& lay, int j, int& blk, int& dst, int& ch_out) {
    int col = 0, b0 = 0;
    for (int m = 0; m < lay.n_modules; ++m) {
        const int n = lay.nb_blocks[m] * lay.channels[m];
        if (j < col + n) {
            const int k = (j - col) / lay.channels[m], c = (j - col) % lay.channels[m];
            blk = b0 + k;
            dst = 2 * col + k * 2 * lay.channels[m] + c;
            ch_out = lay.channels[m];
            return;
        }
        col += n;
        b0 += lay.nb_blocks[m];
    }
    blk = 0; dst = 0; ch_out = 0;
}

__global__ void film_assemble_fwd_kernel(const float* __restrict__ graw, const float* __restrict__ braw,
                                         const float* __restrict__ post, float* __restrict__ film, int B, int NF, int NB,
                                         FilmLayout lay) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * NF) return;
    const int b = i / NF, j = i % NF;
    int blk, dst, ch;
    film_locate(lay, j, blk, dst, ch);
    const float pg = post ? post[blk] : 1.f, pb = post ? post[NB + blk] : 1.f;
    film[(size_t)b * 2 * NF + dst] = pg * graw[i] + 1.f;
    film[(size_t)b * 2 * NF + dst + ch] = pb * braw[i];
}

// dgraw = post_g * dfilm_gamma ; dbraw = post_b * dfilm_beta ; dpost[0][blk] += sum dfilm_gamma * graw ; dpost[1][blk] += ...
__global__ void film_assemble_bwd_kernel(const float* __restrict__ dfilm, const float* __restrict__ graw,
                                         const float* __restrict__ braw, const float* __restrict__ post, float* __restrict__ dgraw,
                                         float* __restrict__ dbraw, float* __restrict__ dpost, int B, int NF, int NB,
                                         FilmLayout lay) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < B * NF;
    int blk = -1, dst = 0, ch = 0;
    float vg = 0.f, vb = 0.f;
    if (live) {
        const int b = i / NF, j = i % NF;
        film_locate(lay, j, blk, dst, ch);
        const float pg = post ? post[blk] : 1.f, pb = post ? post[NB + blk] : 1.f;
        const float dg = dfilm[(size_t)b * 2 * NF + dst], db = dfilm[(size_t)b * 2 * NF + dst + ch];
        dgraw[i] = pg * dg;
        dbraw[i] = pb * db;
        if (dpost) { vg = dg * graw[i]; vb = db * braw[i]; }
    }
    if (dpost) {
        // only 2 * NB (= 18) accumulators: a warp's 32 consecutive channels almost always belong to one block -> one atomic per warp
        const int blk0 = __shfl_sync(0xffffffffu, blk, 0);
        if (__all_sync(0xffffffffu, blk == blk0)) {
            vg = warp_sum(vg);
            vb = warp_sum(vb);
            if ((threadIdx.x & 31) == 0 && blk0 >= 0) { atomicAdd(dpost + blk0, vg); atomicAdd(dpost + NB + blk0, vb); }
        } else if (live) {
            atomicAdd(dpost + blk, vg);
            atomicAdd(dpost + NB + blk, vb);
        }
    }
}

static void film_totals(const FilmLayout& lay, int& NF, int& NB) {
    NF = 0; NB = 0;
    for (int m = 0; m < lay.n_modules; ++m) { NF += lay.nb_blocks[m] * lay.channels[m]; NB += lay.nb_blocks[m]; }
}

int film_assemble_fwd(const float* graw, const float* braw, const float* post, float* film, int B, FilmLayout lay,
                      cudaStream_t st) {
    int NF, NB;
    film_totals(lay, NF, NB);
    film_assemble_fwd_kernel<<<ceil_div(B * NF, 256), 256, 0, st>>>(graw, braw, post, film, B, NF, NB, lay);
    return check_launch("film_assemble_fwd");
}

int film_assemble_bwd(const float* dfilm, const float* graw, const float* braw, const float* post, float* dgraw, float* dbraw,
                      float* dpost, int B, FilmLayout lay, cudaStream_t st) {
    int NF, NB;
    film_totals(lay, NF, NB);
    if (dpost) DX_CUDA(cudaMemsetAsync(dpost, 0, (size_t)2 * NB * sizeof(float), st));
    film_assemble_bwd_kernel<<<ceil_div(B * NF, 256), 256, 0, st>>>(dfilm, graw, braw, post, dgraw, dbraw, dpost, B, NF, NB, lay);
    return check_launch("film_assemble_bwd");
}

// ---------------------------------------------------------------------------------------------------------------------
// narrow projection: out[j][r] = mask * (x[r,:] . w[j,:] + b[j]),  j < NO <= 4, one warp per row.
// reference model.py:566-569 (predictor head, 256 -> 3; input AND output masked).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) narrow_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, const long long* __restrict__ lens,
                                                                float* __restrict__ out, int B, int S, int C, int NO) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int R = B * S;
    if (row >= R) return;
    const int b = row / S, s = row - b * S;
    const bool valid = !lens || s < (int)lens[b];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
        for (int c = lane * 4; c < C; c += 128) {
            const float4 v = *reinterpret_cast<const float4*>(x + (size_t)row * C + c);
            for (int j = 0; j < NO; ++j) {
                const float4 ww = *reinterpret_cast<const float4*>(w + (size_t)j * C + c);
                acc[j] += v.x * ww.x + v.y * ww.y + v.z * ww.z + v.w * ww.w;
            }
        }
    }
    for (int j = 0; j < NO; ++j) {
        const float tot = warp_sum(acc[j]);
        if (lane == 0) out[(size_t)j * R + row] = valid ? tot + bias[j] : 0.f;
    }
}

// dx[r, c] = mask * sum_j dout[j][r] * w[j][c];  dw[j][c] += sum_r mask*dout[j][r]*x[r][c];  db[j] += sum_r mask*dout[j][r]
__global__ void __launch_bounds__(256) narrow_linear_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                                                const float* __restrict__ w, const long long* __restrict__ lens,
                                                                float* __restrict__ dx, float* __restrict__ dw,
                                                                float* __restrict__ db, int B, int S, int C, int NO,
                                                                int rows_per_chunk) {
    __shared__ float sm[4][8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl, b = blockIdx.z;
    const int R = B * S;
    const int len = lens ? min((int)lens[b], S) : S;
    const int s0 = blockIdx.y * rows_per_chunk, s1 = min(S, s0 + rows_per_chunk);
    float aw[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
    float wj[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C)
        for (int j = 0; j < NO; ++j) wj[j] = w[(size_t)j * C + c];
    for (int s = s0 + rl; s < s1; s += 8) {
        const size_t row = (size_t)b * S + s;
        float g = 0.f;
        if (s < len) {
            const float xv = c < C ? x[row * C + c] : 0.f;
            for (int j = 0; j < NO; ++j) {
                const float d = dout[(size_t)j * R + row];
                g += d * wj[j];
                aw[j] += d * xv;
                ab[j] += d;
            }
        }
        if (c < C) dx[row * C + c] = g;
    }
    for (int j = 0; j < NO; ++j) sm[j][rl][cl] = aw[j];
    __syncthreads();
    if (rl < NO && c < C) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += sm[rl][k][cl];
        atomicAdd(dw + (size_t)rl * C + c, tot);
    }
    __syncthreads();
    // bias grads: dout does not depend on the column, so column lane 0 of column-block 0 carries the row-lane partials
    for (int j = 0; j < NO; ++j) sm[j][rl][cl] = ab[j];
    __syncthreads();
    if (blockIdx.x == 0 && rl < NO && cl == 0) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += sm[rl][k][0];
        atomicAdd(db + rl, tot);
    }
}

int narrow_linear_fwd(const float* x, const float* w, const float* b, const long long* lens, float* out, int B, int S, int C,
                      int NO, int mask_input, cudaStream_t st) {
    (void)mask_input;
    DX_REQUIRE(NO >= 1 && NO <= 4 && C % 4 == 0, "narrow_linear: NO=%d (1..4), C=%d (%%4)", NO, C);
    narrow_linear_fwd_kernel<<<ceil_div(B * S, 8), 256, 0, st>>>(x, w, b, lens, out, B, S, C, NO);
    return check_launch("narrow_linear_fwd");
}

int narrow_linear_bwd(const float* dout, const float* x, const float* w, const long long* lens, float* dx, float* dw, float* db,
                      int B, int S, int C, int NO, int mask_input, cudaStream_t st) {
    (void)mask_input;
    DX_REQUIRE(NO >= 1 && NO <= 4, "narrow_linear: NO=%d (1..4)", NO);
    DX_CUDA(cudaMemsetAsync(dw, 0, (size_t)NO * C * sizeof(float), st));
    DX_CUDA(cudaMemsetAsync(db, 0, (size_t)NO * sizeof(float), st));
    const int col_blocks = ceil_div(C, 32);
    const int chunks = max(1, min(ceil_div(S, 32), ceil_div(148 * 4, col_blocks * B)));
    const int rpc = ceil_div(S, chunks);
    dim3 grid(col_blocks, ceil_div(S, rpc), B);
    narrow_linear_bwd_kernel<<<grid, 256, 0, st>>>(dout, x, w, lens, dx, dw, db, B, S, C, NO, rpc);
    return check_launch("narrow_linear_bwd");
}

// ---------------------------------------------------------------------------------------------------------------------
// mel[b, m, t] = mask(t < len[b]) * y[b, t, m]                           reference model.py:707-708
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_transpose_kernel(const float* __restrict__ src, const long long* __restrict__ lens,
                                                             float* __restrict__ dst, int T, int M, int fwd) {
    // fwd: src [B, T, M] -> dst [B, M, T];  bwd: src [B, M, T] -> dst [B, T, M]
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int len = min((int)lens[b], T);
    const int t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (fwd) {
        for (int k = ty; k < 32; k += 8) {
            const int t = t0 + k, m = m0 + tx;
            tile[k][tx] = (t < len && m < M) ? src[((size_t)b * T + t) * M + m] : 0.f;
        }
        __syncthreads();
        for (int k = ty; k < 32; k += 8) {
            const int m = m0 + k, t = t0 + tx;
            if (m < M && t < T) dst[((size_t)b * M + m) * T + t] = tile[tx][k];
        }
    } else {
        for (int k = ty; k < 32; k += 8) {
            const int m = m0 + k, t = t0 + tx;
            tile[k][tx] = (m < M && t < len) ? src[((size_t)b * M + m) * T + t] : 0.f;
        }
        __syncthreads();
        for (int k = ty; k < 32; k += 8) {
            const int t = t0 + k, m = m0 + tx;
            if (t < T && m < M) dst[((size_t)b * T + t) * M + m] = tile[tx][k];
        }
    }
}

int mask_transpose_fwd(const float* y, const long long* lens, float* mel, int B, int T, int M, cudaStream_t st) {
    dim3 grid(ceil_div(T, 32), ceil_div(M, 32), B);
    mask_transpose_kernel<<<grid, 256, 0, st>>>(y, lens, mel, T, M, 1);
    return check_launch("mask_transpose_fwd");
}
int mask_transpose_bwd(const float* dmel, const long long* lens, float* dy, int B, int T, int M, cudaStream_t st) {
    dim3 grid(ceil_div(T, 32), ceil_div(M, 32), B);
    mask_transpose_kernel<<<grid, 256, 0, st>>>(dmel, lens, dy, T, M, 0);
    return check_launch("mask_transpose_bwd");
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}
int relu_bwd(const float* dy, const float* y, float* dx, size_t n, cudaStream_t st) {
    relu_bwd_kernel<<<grid_1d(n), 256, 0, st>>>(dy, y, dx, n);
    return check_launch("relu_bwd");
}
__global__ void scale_copy_kernel(const float* __restrict__ x, float* __restrict__ y, float alpha, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = alpha * x[i];
}
int scale_copy(const float* x, float* y, float alpha, size_t n, cudaStream_t st) {
    scale_copy_kernel<<<grid_1d(n), 256, 0, st>>>(x, y, alpha, n);
    return check_launch("scale_copy");
}

}  // namespace dx
