// DaftExprtLoss as two launches (partial sums, finalize) + one backward launch pair; reference loss.py:30-106.
//   speaker: w_adv * CE(logits, ids)                      post: w_post * ||post||_2
//   dur/energy/pitch: mean_b( sum_i (p - t)^2 / in_len[b] )      mel L1/L2: mean_b( sum_{m,t} |d| or d^2 / (M * out_len[b]) )
// out[8] = {speaker, post_mult, duration, energy, pitch, mel_l1, mel_l2, total}: ONE device buffer, one read-back
// (the reference does 8 .item() syncs, loss.py:102-104 + train.py:382).
#include "common.cuh"
#include "kernels.h"

namespace dx {

// acc[b][0..4] = {sum dur^2, sum energy^2, sum pitch^2, sum |mel|, sum mel^2}
__global__ void __launch_bounds__(256) loss_partial_kernel(LossArgs p, int chunks_mel) {
    __shared__ float sm[8][2];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float a0 = 0.f, a1 = 0.f;
    if (chunk < chunks_mel) {
        const size_t n = (size_t)p.M * p.T;
        const float* mp = p.mel_p + (size_t)b * n;
        const float* mt = p.mel_t + (size_t)b * n;
        for (size_t i = (size_t)chunk * 256 + threadIdx.x; i < n; i += (size_t)chunks_mel * 256) {
            const float d = mp[i] - mt[i];
            a0 += fabsf(d);
            a1 += d * d;
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1);
        if (lane == 0) { sm[warp][0] = a0; sm[warp][1] = a1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float s0 = 0.f, s1 = 0.f;
            for (int k = 0; k < 8; ++k) { s0 += sm[k][0]; s1 += sm[k][1]; }
            atomicAdd(p.acc + b * 8 + 3, s0);
            atomicAdd(p.acc + b * 8 + 4, s1);
        }
    } else {  // the last chunk handles the three prosody MSE sums
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        for (int i = threadIdx.x; i < p.L; i += 256) {
            const size_t k = (size_t)b * p.L + i;
            float d = p.dur_p[k] - p.dur_t[k]; d0 += d * d;
            d = p.energy_p[k] - p.energy_t[k]; d1 += d * d;
            d = p.pitch_p[k] - p.pitch_t[k]; d2 += d * d;
        }
        d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2);
        __shared__ float sp[8][3];
        if (lane == 0) { sp[warp][0] = d0; sp[warp][1] = d1; sp[warp][2] = d2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            for (int k = 0; k < 8; ++k) { s0 += sp[k][0]; s1 += sp[k][1]; s2 += sp[k][2]; }
            p.acc[b * 8 + 0] = s0; p.acc[b * 8 + 1] = s1; p.acc[b * 8 + 2] = s2;
        }
    }
}

__global__ void loss_finalize_kernel(LossArgs p) {
    __shared__ float red[32][6];
    const int t = threadIdx.x;  // 32 threads, loop over utterances
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int b = t; b < p.B; b += 32) {
        const float* lg = p.spk_logits + (size_t)b * p.NS;
        float mx = -INFINITY;
        for (int k = 0; k < p.NS; ++k) mx = fmaxf(mx, lg[k]);
        float se = 0.f;
        for (int k = 0; k < p.NS; ++k) se += expf(lg[k] - mx);
        long long id = p.spk_ids[b];
        id = id < 0 ? 0 : (id >= p.NS ? p.NS - 1 : id);
        v[0] += (mx + logf(se)) - lg[id];
        const float il = (float)p.in_lens[b], ol = (float)p.M * (float)p.out_lens[b];
        v[1] += p.acc[b * 8 + 0] / il;
        v[2] += p.acc[b * 8 + 1] / il;
        v[3] += p.acc[b * 8 + 2] / il;
        v[4] += p.acc[b * 8 + 3] / ol;
        v[5] += p.acc[b * 8 + 4] / ol;
    }
    for (int k = 0; k < 6; ++k) red[t][k] = v[k];
    __syncthreads();
    if (t == 0) {
        float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < 32; ++i)
            for (int k = 0; k < 6; ++k) s[k] += red[i][k];
        const float invB = 1.f / (float)p.B;
        float pn = 0.f;
        if (p.post)
            for (int k = 0; k < p.NP; ++k) pn += p.post[k] * p.post[k];
        p.out[0] = (p.dyn ? p.dyn->w_adv : p.w_adv) * s[0] * invB;
        p.out[1] = p.post ? p.w_post * sqrtf(pn) : 0.f;
        p.out[2] = p.w_dur * s[1] * invB;
        p.out[3] = p.w_energy * s[2] * invB;
        p.out[4] = p.w_pitch * s[3] * invB;
        p.out[5] = p.w_mel * s[4] * invB;
        p.out[6] = p.w_mel * s[5] * invB;
        p.out[7] = p.out[0] + p.out[1] + p.out[2] + p.out[3] + p.out[4] + p.out[5] + p.out[6];
    }
}

int loss_fwd(const LossArgs& a, cudaStream_t st) {
    DX_CUDA(cudaMemsetAsync(a.acc, 0, (size_t)a.B * 8 * sizeof(float), st));
    const size_t n = (size_t)a.M * a.T;
    const int chunks = grid_1d(n, 2048, 64);
    dim3 grid(chunks + 1, a.B);
    loss_partial_kernel<<<grid, 256, 0, st>>>(a, chunks);
    int rc = check_launch("loss_partial");
    if (rc) return rc;
    loss_finalize_kernel<<<1, 32, 0, st>>>(a);
    return check_launch("loss_finalize");
}

// d mel[b,m,t] = g * w_mel / (B * M * out_len[b]) * (sign(d) + 2 d)
__global__ void loss_bwd_mel_kernel(LossArgs p) {
    const float g = p.gout ? p.gout[0] : 1.f;
    const size_t n = (size_t)p.M * p.T, total = (size_t)p.B * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / n);
        const float k = g * p.w_mel / ((float)p.B * (float)p.M * (float)p.out_lens[b]);
        const float d = p.mel_p[i] - p.mel_t[i];
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        p.dmel[i] = k * (sgn + 2.f * d);
    }
}

__global__ void loss_bwd_small_kernel(LossArgs p) {
    const float g = p.gout ? p.gout[0] : 1.f;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nBL = p.B * p.L;
    if (i < nBL) {
        const int b = i / p.L;
        const float k = g * 2.f / ((float)p.B * (float)p.in_lens[b]);
        p.ddur[i] = p.w_dur * k * (p.dur_p[i] - p.dur_t[i]);
        p.denergy[i] = p.w_energy * k * (p.energy_p[i] - p.energy_t[i]);
        p.dpitch[i] = p.w_pitch * k * (p.pitch_p[i] - p.pitch_t[i]);
    }
    if (i < p.B) {  // softmax - onehot
        const float* lg = p.spk_logits + (size_t)i * p.NS;
        float mx = -INFINITY;
        for (int k = 0; k < p.NS; ++k) mx = fmaxf(mx, lg[k]);
        float se = 0.f;
        for (int k = 0; k < p.NS; ++k) se += expf(lg[k] - mx);
        long long id = p.spk_ids[i];
        id = id < 0 ? 0 : (id >= p.NS ? p.NS - 1 : id);
        const float sc = g * (p.dyn ? p.dyn->w_adv : p.w_adv) / (float)p.B;
        for (int k = 0; k < p.NS; ++k) p.dspk_logits[(size_t)i * p.NS + k] = sc * (expf(lg[k] - mx) / se - (k == id ? 1.f : 0.f));
    }
    if (i == 0 && p.post && p.dpost) {
        float pn = 0.f;
        for (int k = 0; k < p.NP; ++k) pn += p.post[k] * p.post[k];
        const float nrm = sqrtf(pn);
        for (int k = 0; k < p.NP; ++k) p.dpost[k] = nrm > 0.f ? g * p.w_post * p.post[k] / nrm : 0.f;
    }
}

int loss_bwd(const LossArgs& a, cudaStream_t st) {
    const size_t total = (size_t)a.B * a.M * a.T;
    loss_bwd_mel_kernel<<<grid_1d(total), 256, 0, st>>>(a);
    int rc = check_launch("loss_bwd_mel");
    if (rc) return rc;
    loss_bwd_small_kernel<<<ceil_div(max(a.B * a.L, 1), 256), 256, 0, st>>>(a);
    return check_launch("loss_bwd_small");
}

}  // namespace dx
