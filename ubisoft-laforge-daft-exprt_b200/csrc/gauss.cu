// Gaussian upsampling (the reference's duration-driven phoneme -> frame expansion; BASELINE north_star calls it the
// "LengthRegulator"): reference GaussianUpsamplingModule.forward, model.py:608-662.
//
// INTEGER CONTRACT (bit-exact): csum = inclusive prefix sum of durations_int (int64, warp-shuffle scan), total = sum,
// T_max = max(total), centres mu_i = float(d_i)/2 + float(csum_{i-1}).
// FLOAT PART: xp = x + conv3(energy) + conv3(pitch);  sigma = softplus((xp + conv3(dur)) . rw + rb), padded -> 1;
//   p[i,t] = exp(-(t+.5-mu_i)^2 / (2 sigma_i^2) - log sigma_i - log sqrt(2 pi)), padded i -> 0;
//   w = p / (sum_i p + 1e-20);  up[t,:] = sum_i w[i,t] * xp[i,:].
// The reference materialises x[...,None] * w[:, :, None] = [B, L, D, T] (3.3 GB at B=32); here a CTA owns 32 frames of one
// utterance, keeps the [L x 32] weight tile in shared memory and streams xp rows: HBM traffic = algorithmic bytes
// (read xp, write up + weights).
#include "common.cuh"
#include "kernels.h"

namespace dx {

constexpr float kLogSqrt2Pi = 0.91893853320467274178f;

// ---- integer part: one warp per utterance, shuffle-based inclusive scan in chunks of 32 ------------------------------
__global__ void gauss_centres_kernel(const long long* __restrict__ dur, long long* __restrict__ csum, float* __restrict__ mu,
                                     long long* __restrict__ total, int B, int L) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    long long carry = 0;
    for (int i0 = 0; i0 < L; i0 += 32) {
        const int i = i0 + lane;
        const long long d = i < L ? dur[(size_t)b * L + i] : 0;
        long long v = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        const long long incl = carry + v;
        if (i < L) {
            csum[(size_t)b * L + i] = incl;
            mu[(size_t)b * L + i] = (float)d / 2.f + (float)(incl - d);   // model.py:641-643
        }
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
    if (lane == 0) total[b] = carry;
}

// ---- xp, z, sigma: one warp per (b, i) row, D == 128 (4 channels per lane) ----------------------------------------------
__global__ void __launch_bounds__(256) gauss_prep_kernel(GaussArgs p) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int R = p.B * p.L;
    if (row >= R) return;
    const int b = row / p.L, i = row - b * p.L;
    const float* eb = p.energy + (size_t)b * p.L;
    const float* fb = p.pitch + (size_t)b * p.L;
    const float* db = p.dur_f + (size_t)b * p.L;
    const float e0 = i > 0 ? eb[i - 1] : 0.f, e1 = eb[i], e2 = i + 1 < p.L ? eb[i + 1] : 0.f;
    const float f0 = i > 0 ? fb[i - 1] : 0.f, f1 = fb[i], f2 = i + 1 < p.L ? fb[i + 1] : 0.f;
    const float d0 = i > 0 ? db[i - 1] : 0.f, d1 = db[i], d2 = i + 1 < p.L ? db[i + 1] : 0.f;
    float dot = 0.f;
    for (int c0 = lane * 4; c0 < p.D; c0 += 128) {
        const float4 xv = *reinterpret_cast<const float4*>(p.x + (size_t)row * p.D + c0);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c0 + k;
            const float ce = p.we[c * 3] * e0 + p.we[c * 3 + 1] * e1 + p.we[c * 3 + 2] * e2 + p.be[c];
            const float cf = p.wp[c * 3] * f0 + p.wp[c * 3 + 1] * f1 + p.wp[c * 3 + 2] * f2 + p.bp[c];
            const float cd = p.wd[c * 3] * d0 + p.wd[c * 3 + 1] * d1 + p.wd[c * 3 + 2] * d2 + p.bd[c];
            o[k] = xs[k] + ce + cf;
            dot += (o[k] + cd) * p.rw[c];
        }
        *reinterpret_cast<float4*>(p.xp + (size_t)row * p.D + c0) = make_float4(o[0], o[1], o[2], o[3]);
    }
    dot = warp_sum(dot);
    if (lane == 0) {
        const float z = dot + p.rb[0];
        const float sp = z > 20.f ? z : log1pf(expf(z));   // nn.Softplus(beta=1, threshold=20)
        p.z[row] = z;
        p.sigma[row] = i < (int)p.lens[b] ? sp : 1.f;      // model.py:637
    }
}

int gauss_prep(const GaussArgs& a, cudaStream_t st) {
    DX_REQUIRE(a.D % 128 == 0, "gauss_prep: D=%d must be a multiple of 128", a.D);
    gauss_centres_kernel<<<ceil_div(a.B, 4), 128, 0, st>>>(a.dur_i, a.csum, a.mu, a.total, a.B, a.L);
    int rc = check_launch("gauss_centres");
    if (rc) return rc;
    gauss_prep_kernel<<<ceil_div(a.B * a.L, 8), 256, 0, st>>>(a);
    return check_launch("gauss_prep");
}

// ---- forward: CTA = (32 frames, utterance) ---------------------------------------------------------------------------
constexpr int FT = 32;  // frames per CTA

__global__ void __launch_bounds__(256) gauss_upsample_fwd_kernel(GaussArgs p) {
    extern __shared__ __align__(16) float sm[];
    float (*Ws)[FT + 1] = reinterpret_cast<float (*)[FT + 1]>(sm);  // [L][33]
    float* Zs = sm + (size_t)p.L * (FT + 1);                        // [8][32] partial column sums, then [32] totals
    // band of phonemes whose Gaussian is non-zero (in fp32, exactly) on at least one of this CTA's 32 frames: the centres grow
    // monotonically with the phoneme index, so the band is a short contiguous range (~20 of 200) and the weighted sum below only
    // walks it — an exact skip, every term left out is multiplied by w == 0.0f
    int* band = reinterpret_cast<int*>(Zs + 256);                   // [lo, hi]
    const int b = blockIdx.y, t0 = blockIdx.x * FT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int len = min((int)p.lens[b], p.L);
    const float tt = (float)(t0 + lane) + 0.5f;
    if (threadIdx.x == 0) { band[0] = p.L; band[1] = -1; }
    __syncthreads();
    float part = 0.f;
    int my_lo = p.L, my_hi = -1;
    for (int i = warp; i < p.L; i += 8) {
        float pr = 0.f;
        if (i < len) {
            const float mu = p.mu[(size_t)b * p.L + i], sg = p.sigma[(size_t)b * p.L + i];
            const float d = tt - mu;
            pr = expf(-(d * d) / (2.f * sg * sg) - logf(sg) - kLogSqrt2Pi);
        }
        Ws[i][lane] = pr;
        part += pr;
        if (__ballot_sync(0xffffffffu, pr != 0.f)) { my_lo = min(my_lo, i); my_hi = i; }
    }
    if (lane == 0 && my_hi >= 0) { atomicMin(&band[0], my_lo); atomicMax(&band[1], my_hi); }
    Zs[warp * 32 + lane] = part;
    __syncthreads();
    if (warp == 0) {
        float z = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) z += Zs[k * 32 + lane];
        Zs[lane] = 1.f / (z + 1e-20f);
    }
    __syncthreads();
    const float inv = Zs[lane];
    for (int i = warp; i < p.L; i += 8) {
        const float w = Ws[i][lane] * inv;
        Ws[i][lane] = w;
        if (t0 + lane < p.T) p.weights[((size_t)b * p.L + i) * p.T + t0 + lane] = w;
    }
    __syncthreads();
    // up[t0 + f][c] = sum_i w[i][f] * xp[i][c] : warp -> frames 4*warp..+3, lane -> channels 4*lane..+3 (+128 per pass)
    const int i_lo = band[0], i_hi = min(band[1], len - 1);
    for (int c0 = lane * 4; c0 < p.D; c0 += 128) {
        float acc[4][4];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[f][k] = 0.f;
        for (int i = i_lo; i <= i_hi; ++i) {
            const float4 xv = *reinterpret_cast<const float4*>(p.xp + ((size_t)b * p.L + i) * p.D + c0);
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const float w = Ws[i][warp * 4 + f];
                acc[f][0] = fmaf(w, xv.x, acc[f][0]);
                acc[f][1] = fmaf(w, xv.y, acc[f][1]);
                acc[f][2] = fmaf(w, xv.z, acc[f][2]);
                acc[f][3] = fmaf(w, xv.w, acc[f][3]);
            }
        }
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const int t = t0 + warp * 4 + f;
            if (t < p.T) *reinterpret_cast<float4*>(p.up + ((size_t)b * p.T + t) * p.D + c0) = make_float4(acc[f][0], acc[f][1], acc[f][2], acc[f][3]);
        }
    }
}

int gauss_upsample_fwd(const GaussArgs& a, cudaStream_t st) {
    DX_REQUIRE(a.D % 128 == 0, "gauss_upsample: D=%d must be a multiple of 128", a.D);
    const size_t smem = ((size_t)a.L * (FT + 1) + 256 + 4) * sizeof(float);   // weight tile, column partials, band bounds
    DX_REQUIRE(smem <= 200 * 1024, "gauss_upsample: L=%d too large for the shared-memory weight tile", a.L);
    DX_CUDA(cudaFuncSetAttribute(gauss_upsample_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(a.T, FT), a.B);
    gauss_upsample_fwd_kernel<<<grid, 256, smem, st>>>(a);
    return check_launch("gauss_upsample_fwd");
}

// ---- backward --------------------------------------------------------------------------------------------------------
// r[t] = dup[t,:] . up[t,:]  (== sum_j w[j,t] * dw[j,t], the softmax-style correction), one warp per frame row.
__global__ void __launch_bounds__(256) gauss_rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b2,
                                                           float* __restrict__ out, int R, int D) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= R) return;
    float acc = 0.f;
    for (int c = lane * 4; c < D; c += 128) {
        const float4 u = *reinterpret_cast<const float4*>(a + (size_t)row * D + c);
        const float4 v = *reinterpret_cast<const float4*>(b2 + (size_t)row * D + c);
        acc += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc;
}

// One warp per phoneme (b, i), D == 128: lane holds 4 channels of xp[i] and of the dxp accumulator; loops over all frames.
//   dw[i,t] = dup[t,:] . xp[i,:] (+ dweights[i,t]);  dlogp = w[i,t] * (dw - r[t]);
//   dsigma_i = sum_t dlogp * ((t+.5-mu)^2 / sigma^3 - 1/sigma);  dxp[i,:] = sum_t w[i,t] * dup[t,:]
__global__ void __launch_bounds__(256) gauss_upsample_bwd_kernel(GaussArgs p, const float* __restrict__ r,
                                                                 const float* __restrict__ rextra) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int R = p.B * p.L;
    if (row >= R) return;
    const int b = row / p.L, i = row - b * p.L;
    float* dxrow = p.dx + (size_t)row * p.D;
    if (i >= (int)p.lens[b]) {  // padded phoneme: w == 0 everywhere
        *reinterpret_cast<float4*>(dxrow + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane == 0) p.dsigma[row] = 0.f;
        return;
    }
    const float4 xv = *reinterpret_cast<const float4*>(p.xp + (size_t)row * p.D + lane * 4);
    const float mu = p.mu[row], sg = p.sigma[row];
    const float inv_s = 1.f / sg, inv_s3 = inv_s * inv_s * inv_s;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float dsig = 0.f;
    const float* wrow = p.weights + (size_t)row * p.T;
    const float* dwrow = p.dweights ? p.dweights + (size_t)row * p.T : nullptr;
    const float* dupb = p.dup + (size_t)b * p.T * p.D;
    const float* rb = r + (size_t)b * p.T;
    const float* rxb = rextra ? rextra + (size_t)b * p.T : nullptr;
    for (int t0 = 0; t0 < p.T; t0 += 32) {
        const int tl = t0 + lane;
        const float wl = tl < p.T ? wrow[tl] : 0.f;   // coalesced: lane <-> frame
        const unsigned nz = __ballot_sync(0xffffffffu, wl != 0.f);
        if (nz == 0u && !dwrow) continue;             // exact skip: every term below is multiplied by w == 0
        float dlp_l = 0.f;                            // lane-owned dlogp for frame tl
        for (int f = 0; f < 32; ++f) {
            if (!((nz >> f) & 1u)) continue;
            const int t = t0 + f;
            const float w = __shfl_sync(0xffffffffu, wl, f);
            const float4 g = *reinterpret_cast<const float4*>(dupb + (size_t)t * p.D + lane * 4);
            float d = g.x * xv.x + g.y * xv.y + g.z * xv.z + g.w * xv.w;
            d = warp_sum(d);
            acc.x = fmaf(w, g.x, acc.x); acc.y = fmaf(w, g.y, acc.y); acc.z = fmaf(w, g.z, acc.z); acc.w = fmaf(w, g.w, acc.w);
            if (lane == f) dlp_l = d;
        }
        if (tl < p.T) {
            float dwv = dlp_l;
            float rr = rb[tl];
            if (dwrow) { dwv += dwrow[tl]; rr += rxb[tl]; }
            const float dlp = wl * (dwv - rr);
            const float dt = (float)tl + 0.5f - mu;
            dsig += dlp * (dt * dt * inv_s3 - inv_s);
        }
    }
    dsig = warp_sum(dsig);
    *reinterpret_cast<float4*>(dxrow + lane * 4) = acc;
    if (lane == 0) p.dsigma[row] = dsig;
}

// rextra[b,t] = sum_j w[j,t] * dweights[j,t]  (only when a gradient flows into the returned alignments)
__global__ void gauss_wdw_kernel(const float* __restrict__ w, const float* __restrict__ dw, float* __restrict__ out, int B, int L,
                                 int T) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * T) return;
    const int b = idx / T, t = idx % T;
    float acc = 0.f;
    for (int j = 0; j < L; ++j) acc += w[((size_t)b * L + j) * T + t] * dw[((size_t)b * L + j) * T + t];
    out[idx] = acc;
}

// Second stage: dsigma -> dz -> (dxp += dz*rw ; drw, drb ; scalar-conv weight grads).  Block = 32 channels x 8 row lanes
// over a chunk of phonemes of one utterance; cross-block accumulation with atomics.  Writes the final dx in place.
__global__ void __launch_bounds__(256) gauss_prep_bwd_kernel(GaussArgs p, int rows_per_chunk) {
    __shared__ float sm[13][8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl, b = blockIdx.z;
    const int len = min((int)p.lens[b], p.L);
    const int i0 = blockIdx.y * rows_per_chunk, i1 = min(p.L, i0 + rows_per_chunk);
    float acc[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) acc[k] = 0.f;
    const float* eb = p.energy + (size_t)b * p.L;
    const float* fb = p.pitch + (size_t)b * p.L;
    const float* db = p.dur_f + (size_t)b * p.L;
    if (c < p.D) {
        const float rwc = p.rw[c];
        const float wd0 = p.wd[c * 3], wd1 = p.wd[c * 3 + 1], wd2 = p.wd[c * 3 + 2], bdc = p.bd[c];
        for (int i = i0 + rl; i < i1; i += 8) {
            const size_t row = (size_t)b * p.L + i;
            float dz = 0.f;
            if (i < len) {
                const float z = p.z[row];
                const float sgm = z > 20.f ? 1.f : 1.f / (1.f + expf(-z));   // softplus'
                dz = p.dsigma[row] * sgm;
            }
            const float e0 = i > 0 ? eb[i - 1] : 0.f, e1 = eb[i], e2 = i + 1 < p.L ? eb[i + 1] : 0.f;
            const float f0 = i > 0 ? fb[i - 1] : 0.f, f1 = fb[i], f2 = i + 1 < p.L ? fb[i + 1] : 0.f;
            const float d0 = i > 0 ? db[i - 1] : 0.f, d1 = db[i], d2 = i + 1 < p.L ? db[i + 1] : 0.f;
            const float g = p.dx[row * p.D + c] + dz * rwc;   // grad wrt xp (and wrt x)
            p.dx[row * p.D + c] = g;
            const float gd = dz * rwc;                        // grad wrt the duration projection output
            const float rin = p.xp[row * p.D + c] + wd0 * d0 + wd1 * d1 + wd2 * d2 + bdc;
            acc[0] += g * e0; acc[1] += g * e1; acc[2] += g * e2;       // dwe
            acc[3] += g * f0; acc[4] += g * f1; acc[5] += g * f2;       // dwp
            acc[6] += g;                                                 // dbe == dbp
            acc[7] += gd * d0; acc[8] += gd * d1; acc[9] += gd * d2;    // dwd
            acc[10] += gd;                                               // dbd
            acc[11] += dz * rin;                                         // drw
            acc[12] += dz;                                               // drb (same on every column lane)
        }
    }
#pragma unroll
    for (int k = 0; k < 13; ++k) sm[k][rl][cl] = acc[k];
    __syncthreads();
    for (int k = rl; k < 13; k += 8) {
        if (c >= p.D) continue;
        float tot = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) tot += sm[k][q][cl];
        if (k < 3) atomicAdd(p.dwe + c * 3 + k, tot);
        else if (k < 6) atomicAdd(p.dwp + c * 3 + (k - 3), tot);
        else if (k == 6) { atomicAdd(p.dbe + c, tot); atomicAdd(p.dbp + c, tot); }
        else if (k < 10) atomicAdd(p.dwd + c * 3 + (k - 7), tot);
        else if (k == 10) atomicAdd(p.dbd + c, tot);
        else if (k == 11) atomicAdd(p.drw + c, tot);
        else if (c == 0) atomicAdd(p.drb, tot);
    }
}

int gauss_upsample_bwd(const GaussArgs& a, cudaStream_t st) {
    DX_REQUIRE(a.D == 128, "gauss_upsample_bwd: D=%d (only 128)", a.D);
    // scratch layout inside dsigma's tail is not used: r / rextra live after dsigma (caller allocates B*L + 2*B*T floats)
    float* r = a.dsigma + (size_t)a.B * a.L;
    float* rextra = a.dweights ? r + (size_t)a.B * a.T : nullptr;
    gauss_rowdot_kernel<<<ceil_div(a.B * a.T, 8), 256, 0, st>>>(a.dup, a.up, r, a.B * a.T, a.D);
    int rc = check_launch("gauss_rowdot");
    if (rc) return rc;
    if (a.dweights) {
        gauss_wdw_kernel<<<ceil_div(a.B * a.T, 256), 256, 0, st>>>(a.weights, a.dweights, rextra, a.B, a.L, a.T);
        rc = check_launch("gauss_wdw");
        if (rc) return rc;
    }
    gauss_upsample_bwd_kernel<<<ceil_div(a.B * a.L, 8), 256, 0, st>>>(a, r, rextra);
    rc = check_launch("gauss_upsample_bwd");
    if (rc) return rc;
    DX_CUDA(cudaMemsetAsync(a.dwd, 0, (size_t)a.D * 3 * sizeof(float), st));
    DX_CUDA(cudaMemsetAsync(a.dwe, 0, (size_t)a.D * 3 * sizeof(float), st));
    DX_CUDA(cudaMemsetAsync(a.dwp, 0, (size_t)a.D * 3 * sizeof(float), st));
    DX_CUDA(cudaMemsetAsync(a.dbd, 0, (size_t)a.D * sizeof(float), st));
    DX_CUDA(cudaMemsetAsync(a.dbe, 0, (size_t)a.D * sizeof(float), st));
    DX_CUDA(cudaMemsetAsync(a.dbp, 0, (size_t)a.D * sizeof(float), st));
    DX_CUDA(cudaMemsetAsync(a.drw, 0, (size_t)a.D * sizeof(float), st));
    DX_CUDA(cudaMemsetAsync(a.drb, 0, sizeof(float), st));
    const int col_blocks = ceil_div(a.D, 32);
    const int chunks = max(1, min(ceil_div(a.L, 16), ceil_div(148 * 2, col_blocks * a.B)));
    const int rpc = ceil_div(a.L, chunks);
    dim3 grid(col_blocks, ceil_div(a.L, rpc), a.B);
    gauss_prep_bwd_kernel<<<grid, 256, 0, st>>>(a, rpc);
    return check_launch("gauss_prep_bwd");
}

}  // namespace dx
