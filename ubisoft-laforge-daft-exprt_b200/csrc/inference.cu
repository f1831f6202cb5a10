// Inference-time controls: float -> integer durations (bit-exact), energy/pitch adjustments.
// reference model.py:789-864 and extract_features.py:69-111.
#include "common.cuh"
#include "kernels.h"

namespace dx {

// One thread per utterance: the reference accumulates interval ends sequentially in Python doubles
// (model.py:800-806), so the order of the fp64 additions is part of the contract.
//   d = float32(pred * factor); d < float32(dur_min) -> 0                    (model.py:793-796, :891)
//   intervals [end_prev, end_prev + d] over non-zero d; total = sum(end - begin) in fp64
//   nb_samples = int(total * sr); nb_frames = 1 + int((nb_samples - n_fft) / hop)     (extract_features.py:76-82)
//   frame centres c_k = n_fft/2 + hop*k; phoneme gets #{k : int(b*sr) < c_k <= int(e*sr)} frames while frames remain
//   (closed form: cnt(e) - cnt(b), cnt(s) = clamp(floor((s - n_fft/2)/hop) + 1, 0, nb_frames)); + edge frames when centered.
__device__ __forceinline__ long long frames_upto(long long s, long long half, long long hop, long long nf) {
    if (s < half) return 0;
    long long c = (s - half) / hop + 1;
    return c < nf ? c : nf;
}

__global__ void int_durations_kernel(const float* __restrict__ pred, const float* __restrict__ factors,
                                     const long long* __restrict__ lens, float* __restrict__ dur_out,
                                     long long* __restrict__ dur_int, long long* __restrict__ totals, int* __restrict__ err,
                                     int B, int L, int sr, int nfft, int hop, int centered) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    (void)lens;
    const double fft_length = (double)nfft / (double)sr;
    const float dur_min = (float)(fft_length / 2.0);
    const float* pr = pred + (size_t)b * L;
    float* out = dur_out + (size_t)b * L;
    long long* di = dur_int + (size_t)b * L;
    // pass 1: thresholded float durations and the fp64 total
    double end_prev = 0.0, total = 0.0;
    int n_nonzero = 0;
    for (int i = 0; i < L; ++i) {
        float d = factors ? __fmul_rn(pr[i], factors[(size_t)b * L + i]) : pr[i];
        if (d < dur_min) d = 0.f;
        out[i] = d;
        di[i] = 0;
        if (d != 0.f) {
            const double e = end_prev + (double)d;
            total += (e - end_prev);
            end_prev = e;
            ++n_nonzero;
        }
    }
    int code = 0;
    const long long nb_samples = (long long)(total * (double)sr);
    const long long nf = 1 + (long long)((double)(nb_samples - nfft) / (double)hop);
    const long long half = (long long)((double)nfft / 2.0);
    long long tot_frames = 0;
    if (n_nonzero == 0 || nf <= 0) {
        code = 1;  // reference raises IndexError (empty int_durations / pop from empty list)
    } else {
        long long curr = 1;
        int first = -1, last = -1, k = 0;
        end_prev = 0.0;
        int i = 0;
        for (; i < L && curr <= nf; ++i) {
            const float d = out[i];
            if (d == 0.f) continue;
            const double bgn = end_prev, e = end_prev + (double)d;
            end_prev = e;
            const long long bi = (long long)(bgn * (double)sr), ei = (long long)(e * (double)sr);
            const long long n = frames_upto(ei, half, hop, nf) - frames_upto(bi, half, hop, nf);
            di[i] = n;
            curr += n;
            if (first < 0) first = i;
            last = i;
            ++k;
        }
        if (curr <= nf) code = 1;  // ran out of phonemes before all frames were assigned: reference pops an empty list
        if (centered && first >= 0) {
            const long long edge = (long long)((double)nfft / 2.0 / (double)hop);
            di[first] += edge;
            if (k < n_nonzero) {  // phonemes remain: the edge frames go to the next non-zero phoneme
                int j = i;
                while (j < L && out[j] == 0.f) ++j;
                if (j < L) di[j] = edge;
                if (k + 1 < n_nonzero) code = 2;  // reference would fail the index assignment (length mismatch)
            } else {
                di[last] += edge;
            }
        }
        for (int j = 0; j < L; ++j) tot_frames += di[j];
    }
    totals[b] = tot_frames;
    err[b] = code;
}

int int_durations(const float* dur_pred, const float* dur_factors, const long long* lens, float* dur_out, long long* dur_int,
                  long long* totals, int* err, int B, int L, int sampling_rate, int filter_length, int hop_length, int centered,
                  cudaStream_t st) {
    int_durations_kernel<<<ceil_div(B, 32), 32, 0, st>>>(dur_pred, dur_factors, lens, dur_out, dur_int, totals, err, B, L,
                                                        sampling_rate, filter_length, hop_length, centered);
    return check_launch("int_durations");
}

// energy *= factor; energy, pitch := 0 where the integer duration is 0          (model.py:895-899)
__global__ void inference_adjust_kernel(float* __restrict__ energy, float* __restrict__ pitch, const float* __restrict__ ef,
                                        const long long* __restrict__ di, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool dead = di[i] == 0;
    energy[i] = dead ? 0.f : __fmul_rn(energy[i], ef[i]);
    if (dead) pitch[i] = 0.f;
}
int inference_adjust(float* energy, float* pitch, const float* energy_factors, const long long* dur_int, int B, int L,
                     cudaStream_t st) {
    inference_adjust_kernel<<<ceil_div(B * L, 256), 256, 0, st>>>(energy, pitch, energy_factors, dur_int, B * L);
    return check_launch("inference_adjust");
}

// pitch shift in Hz (model.py:814-834): p = (log(exp(std*p + mean) + shift) - mean) / std, unvoiced (== 0) stay 0.
// stats[spk] = {mean, std} (float32, as torch casts the Python scalars to the tensor dtype)
__global__ void pitch_shift_kernel(float* __restrict__ pitch, const float* __restrict__ factors, const long long* __restrict__ spk,
                                   const float* __restrict__ stats, int B, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * L) return;
    const int b = i / L;
    const float v = pitch[i];
    if (v == 0.f) return;
    const float mean = stats[spk[b] * 2], sd = stats[spk[b] * 2 + 1];
    float hz = expf(__fadd_rn(__fmul_rn(sd, v), mean));
    hz = __fadd_rn(hz, factors[i]);
    pitch[i] = __fdiv_rn(__fsub_rn(logf(hz), mean), sd);
}
int pitch_shift(float* pitch, const float* factors, const long long* spk, const float* stats, int B, int L, cudaStream_t st) {
    pitch_shift_kernel<<<ceil_div(B * L, 256), 256, 0, st>>>(pitch, factors, spk, stats, B, L);
    return check_launch("pitch_shift");
}

// multiply transform around the voiced mean (model.py:836-864): one warp per utterance.
__global__ void pitch_multiply_kernel(float* __restrict__ pitch, const float* __restrict__ factors, int B, int L) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    float* row = pitch + (size_t)b * L;
    float s = 0.f, n = 0.f;
    for (int i = lane; i < L; i += 32) {
        const float v = row[i];
        if (v != 0.f) { s += v; n += 1.f; }
    }
    s = warp_sum(s); n = warp_sum(n);
    const float mean = s / n;  // n == 0 -> every entry is unvoiced and stays 0
    for (int i = lane; i < L; i += 32) {
        const float v = row[i];
        if (v != 0.f) row[i] = __fadd_rn(v, __fmul_rn(__fsub_rn(v, mean), factors[(size_t)b * L + i]));
    }
}
int pitch_multiply(float* pitch, const float* factors, int B, int L, cudaStream_t st) {
    pitch_multiply_kernel<<<ceil_div(B, 4), 128, 0, st>>>(pitch, factors, B, L);
    return check_launch("pitch_multiply");
}

}  // namespace dx
