// Internal C++ launcher API shared by the .cu files; the public C-ABI over these is in cabi.cu / include/daft_exprt_b200.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dx {

// LayerNorm epilogue of a Cout == 128 GEMM (bf16x3 tensor-core path): v = dropout_in(acc + bias) + res; xhat = (v - mean) * rstd;
// y = mask(film_g * (xhat * ln_w + ln_b) + film_b) -> ConvGemmArgs::y (fp32) and ConvGemmArgs::y_planes (optional bf16 hi|lo planes)
struct LnEpilogueArgs {
    const float* res;       // [B, S, 128] or nullptr
    const float* ln_w;
    const float* ln_b;
    const float* film;      // nullable; gamma at film[b*film_stride + c], beta at film[b*film_stride + 128 + c]
    int film_stride;
    float* xhat;            // [B, S, 128]  (saved for backward)
    float* rstd;            // [B*S]
    float p_in;
    unsigned long long seed_in;
    const StepState* dyn;
};

// y[b, s, n] = epi(alpha * sum_{tap, c} x[b, s + tap - pad, c] * w[tap][n][c] + bias[n]);  w is the packed layout
// [KW][Cout][Cin].  epi: optional ReLU, optional multiply by (relu_src > 0) (ReLU backward fused into a dgrad), optional
// round-to-tf32 of the stored value.
struct ConvGemmArgs {
    const float* x;         // [B, S, ldx]  (first Cin columns used)
    const float* w;         // [KW][Cout][Cin]
    const float* bias;      // [Cout] or nullptr
    const float* relu_src;  // [B, S, Cout] or nullptr
    const float* add_src;   // [B, S, ldy] or nullptr: added after the epilogue (residual / gradient accumulation)
    float* y;               // [B, S, ldy]; nullable when y_planes is given (tensor-core path: no fp32 copy is written)
    void* y_planes;         // optional (tensor-core path, Cout % 32 == 0): the output as bf16 hi|lo planes [2][B*S][Cout], written by the epilogue
    const void* relu_src_hi;  // optional (tensor-core path, Cout % 32 == 0): ReLU mask source as the bf16 hi plane [B*S][Cout] of the forward activation
    float* y_colsum;        // optional (tensor-core path): [Cout] column sums of the stored output over all B*S rows (zeroed + accumulated inside)
    // optional (bf16x3 tensor-core path, the ONLY output then): the output as per-head attention operand planes
    // R[2][B][Cout / head_dim][head_Sp][head_dim] (bf16 hi|lo), columns < head_scale_cols multiplied by head_scale, rows in [S, head_Sp) zero
    const LnEpilogueArgs* ln = nullptr;   // optional: finish with the LayerNorm epilogue above (lens = the row mask)
    void* head_planes = nullptr;
    int head_dim = 0, head_Sp = 0, head_scale_cols = 0;
    float head_scale = 1.f;
    // optional with head_planes: head_dot_out[b][head][s] = sum_d out[b,s,head,d] * head_dot_src[b,s,head*dh + d] (fp32 [B*S][Cout])
    const float* head_dot_src = nullptr;
    float* head_dot_out = nullptr;
    int B, S, Cin, Cout, KW;
    int ldx, ldy;
    float alpha;
    int relu;
    int round_tf32;
    const long long* lens;  // optional [B] valid lengths: output rows >= len[b] + halo are written as zeros by the tensor-core path
    int halo;
    const void* w_planes;   // optional cached bf16 hi|lo planes of w (tensor-core bf16x3 path), else nullptr
    const void* x_planes;   // optional bf16 hi|lo planes of x ([B*S][Cin] each, made by split_activation_planes), else nullptr
    void* workspace;        // tensor-core bf16x3 path: operand planes (see conv_gemm_tc_workspace)
    size_t workspace_bytes;
};

// dw[co][ci][tap] = alpha * sum_{b,s} dy[b,s,co] * x[b, s + tap - pad, ci];  dbias[co] = alpha * sum dy[b,s,co]
struct ConvWgradArgs {
    const float* x;    // [B, S, ldx]
    const float* dy;   // [B, S, Cout] contiguous
    float* dw;         // [Cout][Cin][KW]  (parameter layout)
    float* dbias;      // [Cout] or nullptr
    void* workspace;
    size_t workspace_bytes;
    int B, S, Cin, Cout, KW;
    int ldx;
    float alpha;
    const long long* lens;   // optional [B]: rows >= len[b] + halo carry dy == 0 and are skipped by the tensor-core path
    int halo;
    const void* x_planes;    // optional bf16 hi|lo planes of x  ([B*S][Cin] each)
    const void* dy_planes;   // optional bf16 hi|lo planes of dy ([B*S][Cout] each)
};

int conv_gemm_simt(const ConvGemmArgs& a, cudaStream_t st);
int conv_wgrad_simt(const ConvWgradArgs& a, cudaStream_t st);
size_t conv_wgrad_simt_workspace(const ConvWgradArgs& a, int* nsplit_out);
int colsum(const float* dy, float* db, int R, int C, float alpha, cudaStream_t st);
int wgrad_reduce_flush(cudaStream_t st);
void wgrad_reduce_defer(bool on);
int colsum_planes(const void* planes, float* db, int R, int C, cudaStream_t st);
int pack_conv_weights_batched(const void* descs_device, int n_desc, int total_blocks, int round, cudaStream_t st);
int pack_conv_weight(const float* w, float* fwd, float* dgrad, void* fwd_planes, void* dgrad_planes, int Cout, int Cin, int KW,
                     int round, cudaStream_t st);

// tcgen05 / TMEM / TMA path (gemm_tcgen05.cu)
void set_tc_precision(int tf32);
void set_tc_trace(long long* buf);
void set_tc_passes(int conv, int wgrad);      // BF16X3 passes (1..3) of forward/dgrad GEMMs and of weight-gradient GEMMs
void get_tc_passes(int* conv, int* wgrad);
unsigned long long tc_gemm_launches();   // gemm_tc_kernel launches so far
int split_weight_planes(const float* w, void* planes, size_t n, cudaStream_t st);
int split_activation_planes(const float* x, int ld, void* planes, float* colsum_out, int rows, int C, cudaStream_t st);
size_t conv_gemm_tc_workspace(const ConvGemmArgs& a);
bool conv_gemm_tc_supported(const ConvGemmArgs& a);
int conv_gemm_tc(const ConvGemmArgs& a, cudaStream_t st);
bool conv_wgrad_tc_supported(const ConvWgradArgs& a);
int conv_wgrad_tc(const ConvWgradArgs& a, cudaStream_t st);
size_t conv_wgrad_tc_workspace(const ConvWgradArgs& a);

// attention (attention.cu): qkv [B, S, 3*D] (q | k | v, head h at columns h*dh), ctx [B, S, D], lse [B, H, S]
struct AttnArgs {
    const float* qkv;
    const long long* lens;  // [B] valid keys/queries per utterance
    float* ctx;
    float* lse;
    void* ctx_planes;       // forward, tensor-core path, optional: ctx also as bf16 hi|lo operand planes [2][B*S][H*dh]
    const float* dctx;      // backward only
    float* dqkv;            // backward only, [B, S, 3*D]
    float* delta;           // backward scratch [B, H, S]
    int B, S, H, dh;
    float dropout_p;
    unsigned long long seed;
    const StepState* dyn;   // nullable: graph mode (dx_set_step_state)
    // tensor-core path: per-head bf16 hi|lo operand planes (attention_mma.cu), S padded to Sp (multiple of 64)
    const __nv_bfloat16 *R, *Tr;     // of qkv:  [2][B][3H][Sp][dh] row-major / [2][B][3H][dh][Sp] transposed
    const __nv_bfloat16 *GR, *GTr;   // of dctx: [2][B][H][Sp][dh] / [2][B][H][dh][Sp]
    int Sp;
};
int attention_fwd(const AttnArgs& a, cudaStream_t st);
// tensor-core (mma.sync bf16x3) variants, attention_mma.cu
bool attention_mma_supported(const AttnArgs& a);
int attention_fwd_mma(const AttnArgs& a, void* planes, cudaStream_t st);
int attention_bwd_mma(const AttnArgs& a, void* planes, void* scratch, cudaStream_t st);
size_t attention_planes_bytes(int B, int S, int H, int dh);
size_t attention_bwd_scratch_bytes(int B, int S, int H, int dh);
// tcgen05 / TMEM forward (attention_tc.cu): head_dim 64 / 16, operand planes a.R / a.Sp already bound and filled
bool attention_fwd_tc_supported(const AttnArgs& a);
int attention_fwd_tc(const AttnArgs& a, cudaStream_t st);
void set_attention_backend(int fwd_tc, int bwd_tc);   // 1 = tcgen05, 0 = mma.sync (attention_mma.cu)
// tcgen05 / TMEM backward (attention_bwd_tc.cu): a.R, a.GR (dO planes), a.delta bound and filled, a.dqkv zeroed
bool attention_bwd_tc_supported(const AttnArgs& a);
int attention_bwd_tc(const AttnArgs& a, cudaStream_t st);
// cuTensorMapEncodeTiled wrapper (gemm_tcgen05.cu): 3-D tensor, d0 contiguous, zero out-of-bounds fill; esz 2 (bf16) / 4 (fp32);
// swizzle_bytes in {32, 64, 128}; map = CUtensorMap* (128 bytes, 64-byte aligned)
bool tma_available();
long long* tc_trace_buffer();   // dx_debug_set_trace buffer ([4][256] int64) or nullptr
int make_tma_map_3d(void* map, const void* ptr, int esz, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                    unsigned long long stride1_bytes, unsigned long long stride2_bytes, unsigned b0, unsigned b1, unsigned b2,
                    int swizzle_bytes);
int attention_bwd_prepare(const AttnArgs& a, cudaStream_t st);   // delta + zero dqkv
int attention_bwd(const AttnArgs& a, cudaStream_t st);

// LayerNorm family (norm.cu):  v = dropout_in(a) + res;  y = mask(film_g * dropout_out(LN(v) * w + b) + film_b)
struct LnArgs {
    const float* a;
    const float* res;       // nullable
    const float* ln_w;
    const float* ln_b;
    const float* film;      // nullable; gamma at film[b*film_stride + c], beta at film[b*film_stride + D + c]
    const long long* lens;  // nullable -> no masking
    float* y;
    float* xhat;
    float* rstd;
    int B, S, D, film_stride;
    float p_in, p_out;
    unsigned long long seed_in, seed_out;
    const StepState* dyn;   // nullable: graph mode
    // backward
    const float* dy;
    const float* relu_src;  // nullable [B,S,D]: dv/da are multiplied by (relu_src > 0) (ReLU feeding the LN, pre-net/predictor)
    float* dv;      // grad wrt v (== grad wrt res)
    float* da;      // grad wrt a when p_in > 0 (else nullptr: da == dv)
    float* dln_w;   // [D]   (zeroed + accumulated inside)
    float* dln_b;   // [D]
    float* dfilm;   // [B, 2*D] contiguous (gamma | beta), nullable
    // optional bf16 hi|lo operand planes [2][B*S][D] of the tensor the next GEMM consumes
    void* y_planes;     // forward: planes of y
    void* g_planes;     // backward (D in {128, 256}): planes of da (p_in > 0) or dv
    float* g_colsum;    // backward (D in {128, 256}): [D] column sums of the same gradient (= bias gradient of the GEMM that produced a)
};
int ln_fwd(const LnArgs& a, cudaStream_t st);
int ln_bwd(const LnArgs& a, cudaStream_t st);

// small fused kernels (elementwise.cu)
int embed_pe_fwd(const long long* symbols, const long long* lens, const float* emb, const float* pe, float* y, int B, int L,
                 int D, int n_symbols, cudaStream_t st);
int embed_pe_bwd(const long long* symbols, const long long* lens, const float* dy, float* demb, int B, int L, int D,
                 int n_symbols, cudaStream_t st);
// y = mask * (x + pe + conv3(e; we, be) + conv3(f0; wp, bp));  e/f0 nullable (then only x + pe)
int frame_input_fwd(const float* x, const float* e, const float* f0, const float* we, const float* be, const float* wp,
                    const float* bp, const float* pe, const long long* lens, float* y, int B, int T, int D, cudaStream_t st);
int frame_input_bwd(const float* dy, const float* e, const float* f0, const long long* lens, float* dx, float* dwe, float* dbe,
                    float* dwp, float* dbp, int B, int T, int D, cudaStream_t st);
int meanpool_fwd(const float* x, const long long* lens, float* pooled, int B, int S, int D, cudaStream_t st);
int meanpool_bwd(const float* dpooled, const long long* lens, float* dx, int B, int S, int D, cudaStream_t st);
// h = pooled + spk_emb[spk]
int add_speaker_fwd(const float* pooled, const long long* spk, const float* spk_emb, float* h, int B, int D, int n_spk,
                    cudaStream_t st);
int add_speaker_bwd(const float* dh, const long long* spk, float* dspk_emb, int B, int D, int n_spk, cudaStream_t st);
// FiLM assembly, reference model.py:430-461: raw gammas/betas [B, NF] + post multipliers [2, NB] -> film[B, 2*NF]
// laid out module by module as (gamma | beta) per block; seg_* describe the modules.
struct FilmLayout {
    int n_modules;
    int nb_blocks[4];
    int channels[4];
};
int film_assemble_fwd(const float* graw, const float* braw, const float* post, float* film, int B, FilmLayout lay,
                      cudaStream_t st);
int film_assemble_bwd(const float* dfilm, const float* graw, const float* braw, const float* post, float* dgraw, float* dbraw,
                      float* dpost, int B, FilmLayout lay, cudaStream_t st);
// out[j][r] = mask * (x[r,:] . w[j,:] + b[j]),  j < NO <= 4   (SoA output: NO contiguous [B*S] planes)
int narrow_linear_fwd(const float* x, const float* w, const float* b, const long long* lens, float* out, int B, int S, int C,
                      int NO, int mask_input, cudaStream_t st);
int narrow_linear_bwd(const float* dout, const float* x, const float* w, const long long* lens, float* dx, float* dw, float* db,
                      int B, int S, int C, int NO, int mask_input, cudaStream_t st);
// mel[b, m, t] = mask * y[b, t, m]
int mask_transpose_fwd(const float* y, const long long* lens, float* mel, int B, int T, int M, cudaStream_t st);
int mask_transpose_bwd(const float* dmel, const long long* lens, float* dy, int B, int T, int M, cudaStream_t st);
int relu_bwd(const float* dy, const float* y, float* dx, size_t n, cudaStream_t st);
int scale_copy(const float* x, float* y, float alpha, size_t n, cudaStream_t st);

// Gaussian upsampling (gauss.cu), reference model.py:608-662
struct GaussArgs {
    const float* x;          // [B, L, D] encoder outputs
    const float* dur_f;      // [B, L]
    const long long* dur_i;  // [B, L]
    const float* energy;     // [B, L]
    const float* pitch;      // [B, L]
    const long long* lens;   // [B]
    const float *wd, *bd, *we, *be, *wp, *bp;  // [D,1,3] / [D]
    const float *rw, *rb;                      // [1, D] / [1]
    float* xp;          // [B, L, D]  x + conv(energy) + conv(pitch)
    float* z;           // [B, L] pre-softplus
    float* sigma;       // [B, L]
    float* mu;          // [B, L]
    long long* csum;    // [B, L] inclusive prefix sum of dur_i
    long long* total;   // [B] = sum(dur_i)
    float* up;          // [B, T, D]
    float* weights;     // [B, L, T]
    int B, L, T, D;
    // backward
    const float* dup;       // [B, T, D]
    const float* dweights;  // [B, L, T] or nullptr
    float* dx;              // [B, L, D]
    float* dsigma;          // [B, L] scratch
    float *dwd, *dbd, *dwe, *dbe, *dwp, *dbp, *drw, *drb;
};
int gauss_prep(const GaussArgs& a, cudaStream_t st);       // xp, z, sigma, mu, csum, total
int gauss_upsample_fwd(const GaussArgs& a, cudaStream_t st);
int gauss_upsample_bwd(const GaussArgs& a, cudaStream_t st);

// loss (loss.cu), reference loss.py:30-106.  out[8] = speaker, post_mult, duration, energy, pitch, mel_l1, mel_l2, total
struct LossArgs {
    const float* spk_logits;    // [B, NS]
    const long long* spk_ids;   // [B]
    const float* post;          // [NP] or nullptr
    const float *dur_p, *energy_p, *pitch_p;  // [B, L]
    const float *dur_t, *energy_t, *pitch_t;  // [B, L]
    const float* mel_p;         // [B, M, T]
    const float* mel_t;
    const long long* in_lens;
    const long long* out_lens;
    int B, L, T, M, NS, NP;
    float w_adv, w_post, w_dur, w_energy, w_pitch, w_mel;
    const StepState* dyn;       // nullable: graph mode, w_adv is read from here
    float* acc;                 // scratch [B, 8]
    float* out;                 // [8]
    // backward
    const float* gout;          // device scalar: d(total)
    float *dspk_logits, *dpost, *ddur, *denergy, *dpitch, *dmel;
};
int loss_fwd(const LossArgs& a, cudaStream_t st);
int loss_bwd(const LossArgs& a, cudaStream_t st);

// inference helpers (inference.cu), reference model.py:789-864, extract_features.py:69-111
int int_durations(const float* dur_pred, const float* dur_factors, const long long* lens, float* dur_out, long long* dur_int,
                  long long* totals, int* err, int B, int L, int sampling_rate, int filter_length, int hop_length, int centered,
                  cudaStream_t st);
int inference_adjust(float* energy, float* pitch, const float* energy_factors, const long long* dur_int, int B, int L,
                     cudaStream_t st);
int pitch_shift(float* pitch, const float* factors, const long long* spk, const float* stats, int B, int L, cudaStream_t st);
int pitch_multiply(float* pitch, const float* factors, int B, int L, cudaStream_t st);

// fused Adam step over a flat parameter buffer (optim.cu): torch.optim.Adam semantics (L2 weight decay added to grad)
int adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
              float weight_decay, int step, float grad_scale, const StepState* dyn, const float* clip, cudaStream_t st);
// reduce-scatter (sum over ranks) + Adam on this rank's shard + all-gather of the updated parameters, one kernel over symmetric
// memory: g_mc / p_mc = NVSwitch multicast addresses of the flat gradient / parameter buffers (or nullptr), g_peers / p_peers = HOST
// arrays of the `world` per-rank device addresses of the same buffers (used when there is no multicast mapping), p_local / m / v =
// this rank's own buffers; the shard is elements [begin, begin + n) of the flat layout
int fused_reduce_adam(const float* g_mc, const unsigned long long* g_peers, float* p_mc, const unsigned long long* p_peers, const float* p_local,
                      float* m, float* v, size_t begin, size_t n, int world, float lr, float beta1, float beta2, float eps, float weight_decay,
                      int step, float grad_scale, const StepState* dyn, cudaStream_t st);
// clip_grad_norm_ on the device: out[3] = {norm of grad_scale * g, clip coefficient, scratch}; `clip` above = this buffer
int grad_norm_clip(const float* g, size_t n, float grad_scale, float max_norm, float* out, cudaStream_t st);

}  // namespace dx
