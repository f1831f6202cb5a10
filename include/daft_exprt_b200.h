/* daft_exprt_b200 — C-ABI of the Blackwell-native (sm_100a) Daft-Exprt mel-prediction path.
 *
 * This is the drop-in boundary.  The reference (ubisoft/ubisoft-laforge-daft-exprt) has NO native code and
 * no FFI: every hot-path FLOP is a torch.nn call inside `src/daft_exprt/model.py` / `loss.py`.  Each entry point below
 * therefore cites the reference Python call site(s) it replaces; the host-side binding a maintainer adds is the ctypes stub
 * shown in INTEGRATION.md (and implemented in ubisoft-laforge-daft-exprt_b200/cabi.py).
 *
 * Conventions
 *   - plain C: raw device pointers + sizes, no torch types; all pointers are BORROWED (caller allocates every output and
 *     every workspace; the library never allocates persistent device memory and never frees caller memory);
 *   - activations are channels-last fp32 `[B, S, C]` (row = one phoneme / mel frame); lengths / ids are int64 device arrays,
 *     exactly the tensors `DaftExprt.parse_batch` (model.py:727-753) produces;
 *   - every COMPUTE call is asynchronous on `stream` (a `cudaStream_t` passed as void*; pass torch's current stream), keeps no
 *     state of its own between calls and does no host synchronisation, so different streams may be driven concurrently as long
 *     as the CONFIGURATION below is not changed meanwhile;
 *   - configuration is PROCESS-GLOBAL and not thread-safe (set it once, or between steps, from one thread): the GEMM backend and
 *     pass count (dx_set_gemm_backend / dx_set_gemm_passes), the attention kernel choice (dx_set_attention_backend), the per-step
 *     device block (dx_set_step_state) and the deferred weight-gradient reduction list (dx_wgrad_defer / dx_wgrad_flush: entries
 *     recorded by dx_conv_wgrad calls are flushed by the next dx_wgrad_flush on whatever stream it names).  One model per process
 *     and one stream at a time inside a deferred backward is the supported pattern (the reference runs one model per process too);
 *   - return 0 on success, negative on error (DX_ERR_*); `dx_last_error()` returns a thread-local message; no exceptions
 *     cross the ABI;
 *   - dropout: `p == 0` disables; masks are regenerated from (seed, element index) in backward, never stored.
 */
#ifndef DAFT_EXPRT_B200_H
#define DAFT_EXPRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DX_ABI_VERSION 1
#define DX_GEMM_FP32_CUDA_CORES 0 /* exact fp32 (parity mode) */
#define DX_GEMM_TCGEN05_TF32 1    /* tcgen05.mma kind::tf32 on fp32 tiles: one pass, ~1e-3 per GEMM (forward/dgrad; wgrad runs bf16x3) */
#define DX_GEMM_TCGEN05_BF16X3 2  /* tcgen05.mma kind::f16 on bf16 hi/lo operand planes, 3 passes, fp32-grade results (default; see dx_set_gemm_passes) */

const char* dx_last_error(void);
int dx_abi_version(void);
/* number of CUDA kernels this library has launched so far in this process */
uint64_t dx_launch_count(void);
/* number of tcgen05 tensor-core GEMM kernels (gemm_tc_kernel) among them: lets a caller assert that the tensor-core path,
 * not the exact-fp32 CUDA-core path, is the one running */
uint64_t dx_tc_gemm_launch_count(void);
/* 0 when the current device is sm_100 (B200); negative otherwise */
int dx_device_check(void);
int dx_set_gemm_backend(int backend);
int dx_get_gemm_backend(void);
/* DX_GEMM_TCGEN05_BF16X3 only: tensor-core passes per K-step issued by the GEMMs called from now on — conv_passes for dx_conv_gemm* /
 * dx_inproj_head_planes (forward and input-gradient GEMMs), wgrad_passes for dx_conv_wgrad.  3 (conv default) = hi*hi + lo*hi +
 * hi*lo: fp32-grade, the parity mode.  2 = hi*hi + lo*hi: the second operand (weights; x for wgrad) is rounded to bf16.  1 = hi*hi:
 * plain bf16 operands, fp32 accumulation (the usual mixed-precision training arithmetic).  Reduced modes do not load the planes they
 * skip.  wgrad_passes = 0 (wgrad default) chooses by reduction length: a weight gradient whose sum runs over B*S >= 4096 rows takes
 * ONE pass (the bf16 rounding of the operands is unbiased and independent from row to row; measured against the fp64 oracle the 193
 * gradients are unchanged to three digits, worst tensor 1.05e-3 either way), a shorter one takes three.  conv_passes < 3 and
 * wgrad_passes 1 / 2 are NOT parity modes: measured errors are in profiles/r2_pass_ablation.md. */
int dx_set_gemm_passes(int conv_passes, int wgrad_passes);
/* Attention kernels used with the tensor-core GEMM backends: DX_ATTENTION_TCGEN05 (default: tcgen05/TMEM/TMA forward and
 * backward for head_dim 64 and 16) or DX_ATTENTION_MMA_SYNC (the mma.sync flash kernels; always used for head_dim 32).
 * Initial value from the environment (DX_ATTN_TC=0 / DX_ATTN_BWD_TC=0 select mma.sync for forward / backward). */
#define DX_ATTENTION_MMA_SYNC 0
#define DX_ATTENTION_TCGEN05 1
int dx_set_attention_backend(int forward_backend, int backward_backend);
/* bring-up aid: device buffer of 4*256 int64 receiving clock64() traces of CTA 0 of the tensor-core GEMM (NULL = off) */
int dx_debug_set_trace(void* buf);

/* ---- Conv1d / Linear as channels-last GEMMs --------------------------------------------------------------------------
 * replaces nn.Conv1d inside ConvNorm1D (model.py:82,86-94) and nn.Linear inside LinearNorm (model.py:63,66-72), incl. the
 * in/out projections of nn.MultiheadAttention (model.py:165).  Weights are consumed in a packed, cached layout. */
/* w [Cout][Cin][KW] (the parameter) -> fwd [KW][Cout][Cin] and dgrad [KW][Cin][Cout] with taps flipped (either may be NULL);
 * fwd_planes / dgrad_planes (nullable): bf16 hi|lo planes of the same packed layouts (2 * KW*Cout*Cin bf16 each), fused in the
 * same pass (what dx_split_weight_planes would produce). */
int dx_pack_conv_weight(const float* w, float* fwd, float* dgrad, void* fwd_planes, void* dgrad_planes, int Cout, int Cin, int KW,
                        int round_tf32, void* stream);
/* The same repack for MANY weights in ONE launch (an optimiser step makes every pack stale at once).  descs_device: device
 * array of n_desc 64-byte descriptors { const float* w; float* fwd; float* dgrad; void* fwd_planes; void* dgrad_planes;
 * int Cout, Cin, KW, block0; int64 pad } (KW <= 4), block0 = running sum of ceil(Cout/32)*ceil(Cin/32) over the previous
 * descriptors, total_blocks = that sum over all of them. */
int dx_pack_conv_weights_batched(const void* descs_device, int n_desc, int total_blocks, int round_tf32, void* stream);
/* y[b,s,n] = epi(alpha * sum_{tap,c} x[b, s+tap-(KW-1)/2, c] * w[tap][n][c] + bias[n]); zero padding at s<0, s>=S only.
 * epi: relu, then multiply by (relu_src > 0) when relu_src != NULL (ReLU backward fused into a dgrad), then + add_src
 * (same layout as y; residual / gradient accumulation) when != NULL, then optional tf32 rounding of the stored value.  backend < 0 selects the global default. */
/* w_planes (nullable): cached bf16 hi|lo planes of w_packed made by dx_split_weight_planes (2 * KW*Cout*Cin bf16);
 * workspace: dx_conv_gemm_workspace(...) bytes (operand planes of the bf16x3 tensor-core path; 0 for the fp32 backend). */
/* x_planes / dy_planes (nullable): bf16 hi|lo planes of an activation made once by dx_split_planes (2 * rows*C bf16) and shared by
 * every GEMM that consumes it (forward + wgrad, or dgrad + wgrad); when NULL the call splits into its workspace. */
/* lens (nullable, [B] int64) + halo: padding skip.  The caller asserts that output rows s >= len[b] + halo cannot reach any valid
 * result (they are masked downstream / beyond the receptive field of what follows); the tensor-core path then writes whole
 * 128-row tiles made of such rows as ZEROS without computing them (the fp32 backend computes everything). */
/* Plane hand-over between GEMMs (tensor-core backends, Cout % 32 == 0): y_planes (nullable) receives the output as bf16 hi|lo
 * planes [2][B*S][Cout] straight from the epilogue, so that the next GEMM consumes it as x_planes without a split pass; y may
 * then be NULL (no fp32 copy at all).  x may be NULL when x_planes is given.  relu_src_hi (nullable): the ReLU mask taken from
 * the bf16 hi plane [B*S][Cout] of the forward activation instead of an fp32 relu_src (hi > 0 <=> the fp32 value was > 0).
 * y_colsum (nullable, tensor-core backends): [Cout] column sums of the stored output over all B*S rows, accumulated by the
 * epilogue (when the output is a gradient dh this is the bias gradient of the layer below: no separate reduction pass). */
int dx_conv_gemm(const float* x, const void* x_planes, const float* w_packed, const void* w_planes, const float* bias,
                 const float* relu_src, const void* relu_src_hi, const float* add_src, float* y, void* y_planes, float* y_colsum,
                 void* workspace, size_t workspace_bytes, const int64_t* lens, int halo, int B, int S, int Cin, int Cout, int KW,
                 int ldx, int ldy, float alpha, int relu, int round_tf32, int backend, void* stream);
size_t dx_conv_gemm_workspace(int B, int S, int Cin, int Cout, int KW, int have_x_planes, int have_w_planes, int backend);
/* Deferred split-K reduction (tensor-core backends).  dx_wgrad_defer(1): every following dx_conv_wgrad only RECORDS the
 * reduction of its split-K partials into dw (the caller must then keep each call's workspace alive and must not read dw);
 * dx_wgrad_flush performs all recorded reductions with ONE launch (also issued automatically when 72 are pending).
 * dx_wgrad_defer(0) restores reduce-per-call (the default); both drop anything still pending. */
int dx_wgrad_defer(int on);
int dx_wgrad_flush(void* stream);
/* colsum_out (nullable, [C]): column sums of x accumulated in the same pass (= the bias gradient when x is a dy) */
int dx_split_planes(const float* x, int ld, void* planes, float* colsum_out, int rows, int C, void* stream);
int dx_split_weight_planes(const float* w_packed, void* planes, size_t n, void* stream);
size_t dx_conv_wgrad_workspace(int B, int S, int Cin, int Cout, int KW, int have_x_planes, int have_dy_planes, int backend);
/* dw[co][ci][tap] = alpha * sum_{b,s} dy[b,s,co] * x[b, s+tap-pad, ci]  (parameter layout);  dbias[co] = alpha * sum dy */
/* lens + halo (nullable): the caller asserts dy[b, s, :] == 0 for s >= len[b] + halo; such 64-row K-chunks are skipped. */
int dx_conv_wgrad(const float* x, const void* x_planes, const float* dy, const void* dy_planes, float* dw, float* dbias,
                  void* workspace, size_t workspace_bytes, const int64_t* lens, int halo, int B, int S, int Cin, int Cout, int KW,
                  int ldx, float alpha, int backend, void* stream);
int dx_colsum(const float* dy, float* db, int rows, int C, float alpha, void* stream);
/* the same over bf16 hi|lo planes [2][rows][C] (db[c] = sum_r hi + lo): bias gradient of a dy that only exists as planes */
int dx_colsum_planes(const void* planes, float* db, int rows, int C, void* stream);
int dx_relu_bwd(const float* dy, const float* y, float* dx, size_t n, void* stream);
int dx_scale_copy(const float* x, float* y, float alpha, size_t n, void* stream);

/* ---- multi-head self-attention core (flash style, S x S never materialised) ------------------------------------------
 * replaces nn.MultiheadAttention's scaled-dot-product part as called at model.py:182-186 (key_padding_mask from lengths).
 * qkv [B,S,3*H*dh] = in-projection output (q|k|v); ctx [B,S,H*dh]; lse [B,H,S]; padded query rows are written as 0. */
/* planes: workspace of dx_attention_planes_bytes() written by the forward (per-head bf16 hi|lo operand planes of q|k|v used by
 * the tensor-core kernels) and handed back to the backward; may be NULL with the fp32 backend. */
size_t dx_attention_planes_bytes(int B, int S, int H, int dh);
/* 1 when dx_attention_fwd/bwd will run the tensor-core kernels for this head layout under the current GEMM backend (and
 * therefore WRITE `planes` / `ctx_planes`), 0 when they fall back to the exact-fp32 kernels (which never touch them). */
int dx_attention_uses_planes(int H, int dh);
size_t dx_attention_bwd_scratch_bytes(int B, int S, int H, int dh);
/* ctx_planes (nullable, tensor-core backends): ctx additionally as bf16 hi|lo operand planes [2][B*S][H*dh] for the
 * out-projection GEMM (no split pass over ctx). */
/* GEMM + sub-layer tail in ONE kernel (bf16x3 backend, 128 output channels = one column tile of the tcgen05 GEMM):
 *     v = dropout_{p_in}(conv_KW(x) + bias) + res;   xhat = (v - mean_c v) * rstd;   y = mask(film_g * (xhat * ln_w + ln_b) + film_b)
 * replaces  out_proj -> dropout -> + query -> LayerNorm -> masked_fill  (model.py:186-191,259; KW = 1, film = NULL) and
 *           conv2 -> dropout -> + x -> LayerNorm -> FiLM -> masked_fill (model.py:226-235,262; KW = 3).
 * x_planes [2][B*S][Cin] / w_planes [KW][128][Cin] hi|lo: bf16 operand planes; res [B,S,128] nullable; film nullable (gamma at
 * film[b*film_stride + c], beta at film[b*film_stride + 128 + c]); lens: rows >= lens[b] are written as zeros (y, xhat, rstd).
 * Outputs: y [B,S,128] fp32, y_planes (nullable) [2][B*S][128] bf16 hi|lo for the next GEMM, xhat [B,S,128] and rstd [B*S]
 * (what dx_ln_bwd consumes).  Same dropout mask as dx_ln_fwd for the same seed. */
int dx_conv_gemm_ln(const void* x_planes, const void* w_planes, const float* bias, const float* res, const float* ln_w, const float* ln_b,
                    const float* film, int film_stride, const int64_t* lens, float* y, void* y_planes, float* xhat, float* rstd, int B, int S,
                    int Cin, int KW, float p_in, uint64_t seed_in, void* stream);
/* Fused in-projection -> attention operand planes (bf16x3 backend): qkv = x W_in^T + b_in (model.py:165, nn.MultiheadAttention's
 * in_proj) is never materialised in fp32; the tcgen05 GEMM epilogue writes `planes` (the dx_attention_planes_bytes() workspace)
 * directly: per-head bf16 hi|lo planes, q pre-scaled by 1/sqrt(dh), rows in [S, pad64(S)) zero.  Follow with
 * dx_attention_fwd(qkv = NULL, ..., planes).  x_planes: bf16 hi|lo planes [2][B*S][Cin] of x; w_planes: planes of the packed
 * in_proj weight [1][3*H*dh][Cin]; lens: rows >= lens[b] may be skipped (their keys are masked, their queries unused). */
int dx_inproj_head_planes(const void* x_planes, const void* w_planes, const float* bias, void* head_planes, const int64_t* lens,
                          int B, int S, int Cin, int H, int dh, void* stream);
/* qkv may be NULL when `planes` was filled by dx_inproj_head_planes (tensor-core backends only). */
int dx_attention_fwd(const float* qkv, const int64_t* lens, float* ctx, float* lse, void* planes, void* ctx_planes, int B, int S,
                     int H, int dh, float dropout_p, uint64_t seed, void* stream);
/* Backward counterpart of dx_inproj_head_planes: d(ctx) = d(proj) W_out (the input gradient of nn.MultiheadAttention's out_proj,
 * model.py:165) is never materialised in fp32; the GEMM epilogue writes the per-head bf16 hi|lo planes of dO AND
 * delta[b,h,s] = sum_d dO * O (softmax-backward row term) into `bwd_scratch` (dx_attention_bwd_scratch_bytes()), in the layout
 * dx_attention_bwd builds for itself.  Follow with dx_attention_bwd(qkv = NULL, ..., dctx = NULL, ..., scratch = bwd_scratch).
 * dy_planes: planes [2][B*S][Cin] of d(proj); w_dgrad_planes: planes of the dgrad pack [1][H*dh][Cin] of out_proj.weight; ctx: the
 * forward attention output [B,S,H*dh] fp32. */
int dx_outproj_dgrad_head_planes(const void* dy_planes, const void* w_dgrad_planes, const float* ctx, void* bwd_scratch, const int64_t* lens,
                                 int B, int S, int Cin, int H, int dh, void* stream);
/* scratch: dx_attention_bwd_scratch_bytes() bytes; dqkv [B,S,3*H*dh] is fully written.  qkv may be NULL (tensor-core kernels read
 * `planes` only); dctx may be NULL when `scratch` was filled by dx_outproj_dgrad_head_planes. */
int dx_attention_bwd(const float* qkv, const void* planes, const int64_t* lens, const float* ctx, const float* lse,
                     const float* dctx, float* dqkv, void* scratch, int B, int S, int H, int dh, float dropout_p, uint64_t seed,
                     void* stream);

/* ---- residual + LayerNorm + FiLM + padding mask ----------------------------------------------------------------------
 * v = dropout_in(a) + res;  y = mask(film_gamma * dropout_out(LN(v)*w + b) + film_beta)
 * replaces model.py:189-191,259 (attention epilogue), :226-235,262 (conv-FF epilogue + FiLM), the LayerNorm+Dropout pairs
 * of the pre-net (:347-348,354-355,361-362) and of the prosody predictor (:534-535,541-542,559-566).
 * film: gamma at film[b*film_stride + c], beta at film[b*film_stride + D + c] (NULL = no FiLM); lens NULL = no mask.
 * D in {128, 256, 1024}.  Saves xhat [B,S,D] and rstd [B*S] for backward.
 * y_planes (nullable): y additionally as bf16 hi|lo operand planes [2][B*S][D] for the GEMM that consumes it. */
int dx_ln_fwd(const float* a, const float* res, const float* ln_w, const float* ln_b, const float* film, int film_stride,
              const int64_t* lens, float* y, float* xhat, float* rstd, void* y_planes, int B, int S, int D, float p_in,
              uint64_t seed_in, float p_out, uint64_t seed_out, void* stream);
/* relu_src (nullable, [B,S,D]): dv/da are multiplied by (relu_src > 0) — the ReLU that feeds the LN in the pre-net/predictor.
 * dv = grad wrt v (== grad wrt res); da (nullable) = grad wrt a when p_in > 0; dln_w/dln_b [D]; dfilm [B,2D] nullable.
 * g_planes / g_colsum (nullable): the gradient that leaves through `a` (da when p_in > 0, else dv) as
 * bf16 hi|lo operand planes [2][B*S][D] and its column sums [D] (= bias gradient of the GEMM that produced a). */
int dx_ln_bwd(const float* dy, const float* xhat, const float* rstd, const float* ln_w, const float* ln_b, const float* film,
              int film_stride, const int64_t* lens, const float* relu_src, float* dv, float* da, float* dln_w, float* dln_b, float* dfilm,
              void* g_planes, float* g_colsum, int B, int S, int D, float p_in, uint64_t seed_in, float p_out, uint64_t seed_out,
              void* stream);

/* ---- embeddings, positional encoding, masks ---------------------------------------------------------------------------
 * pe: the reference's sinusoid table (model.py:123-130), rows >= max(S) needed.  replaces model.py:497-504. */
int dx_embed_pe_fwd(const int64_t* symbols, const int64_t* lens, const float* emb, const float* pe, float* y, int B, int L,
                    int D, int n_symbols, void* stream);
int dx_embed_pe_bwd(const int64_t* symbols, const int64_t* lens, const float* dy, float* demb, int B, int L, int D,
                    int n_symbols, void* stream);
/* y = mask * (x + PE + Conv1d(1->D,k3)(energy) + Conv1d(1->D,k3)(pitch)); energy == NULL -> y = mask * (x + PE).
 * replaces model.py:400-414 (prosody encoder input) and model.py:696-701 (frame decoder input). */
int dx_frame_input_fwd(const float* x, const float* energy, const float* pitch, const float* we, const float* be,
                       const float* wp, const float* bp, const float* pe, const int64_t* lens, float* y, int B, int T, int D,
                       void* stream);
int dx_frame_input_bwd(const float* dy, const float* energy, const float* pitch, const int64_t* lens, float* dx, float* dwe,
                       float* dbe, float* dwp, float* dbp, int B, int T, int D, void* stream);
/* pooled[b,:] = sum_s x[b,s,:] / len[b]   (model.py:419) */
int dx_meanpool_fwd(const float* x, const int64_t* lens, float* pooled, int B, int S, int D, void* stream);
int dx_meanpool_bwd(const float* dpooled, const int64_t* lens, float* dx, int B, int S, int D, void* stream);
/* h = pooled + spk_embedding[speaker_ids]   (model.py:423-424) */
int dx_add_speaker_fwd(const float* pooled, const int64_t* spk, const float* spk_emb, float* h, int B, int D, int n_spk,
                       void* stream);
int dx_add_speaker_bwd(const float* dh, const int64_t* spk, float* dspk_emb, int B, int D, int n_spk, void* stream);
/* FiLM split + scalar post-multipliers (model.py:430-461).  graw/braw [B,NF]; post [2,NB] or NULL; film [B, 2*NF] laid out
 * module after module as [nb_blocks][gamma(ch) | beta(ch)].  n_modules <= 4. */
int dx_film_assemble_fwd(const float* graw, const float* braw, const float* post, float* film, int B, int n_modules,
                         const int* nb_blocks, const int* channels, void* stream);
int dx_film_assemble_bwd(const float* dfilm, const float* graw, const float* braw, const float* post, float* dgraw,
                         float* dbraw, float* dpost, int B, int n_modules, const int* nb_blocks, const int* channels,
                         void* stream);
/* out[j][b*S+s] = mask * (x[b,s,:] . w[j,:] + bias[j]), j < NO <= 4  (predictor head, model.py:566-573) */
int dx_narrow_linear_fwd(const float* x, const float* w, const float* bias, const int64_t* lens, float* out, int B, int S,
                         int C, int NO, void* stream);
int dx_narrow_linear_bwd(const float* dout, const float* x, const float* w, const int64_t* lens, float* dx, float* dw,
                         float* db, int B, int S, int C, int NO, void* stream);
/* mel[b,m,t] = mask * y[b,t,m]   (model.py:707-708) */
int dx_mask_transpose_fwd(const float* y, const int64_t* lens, float* mel, int B, int T, int M, void* stream);
int dx_mask_transpose_bwd(const float* dmel, const int64_t* lens, float* dy, int B, int T, int M, void* stream);

/* ---- Gaussian upsampling = duration-driven phoneme -> frame expansion (north_star's "LengthRegulator") ---------------
 * replaces GaussianUpsamplingModule.forward, model.py:608-662.
 * prep: xp = x + conv3(energy) + conv3(pitch); z = (xp + conv3(dur_f)).rw + rb; sigma = softplus(z) (padded -> 1);
 *       csum = inclusive int64 prefix sum of dur_i (BIT-EXACT), total[b] = sum, mu = float(d)/2 + float(csum - d). */
int dx_gauss_prep(const float* x, const float* dur_f, const int64_t* dur_i, const float* energy, const float* pitch,
                  const int64_t* lens, const float* wd, const float* bd, const float* we, const float* be, const float* wp,
                  const float* bp, const float* rw, const float* rb, float* xp, float* z, float* sigma, float* mu,
                  int64_t* csum, int64_t* total, int B, int L, int D, void* stream);
/* up[b,t,:] = sum_i w[i,t] xp[b,i,:];  weights [B,L,T] = the reference's `alignments` output */
int dx_gauss_upsample_fwd(const float* xp, const float* mu, const float* sigma, const int64_t* lens, float* up, float* weights,
                          int B, int L, int T, int D, void* stream);
/* scratch: B*L + 2*B*T floats.  dx [B,L,D] = grad wrt x; parameter grads in the parameters' own layouts. */
int dx_gauss_upsample_bwd(const float* dup, const float* dweights, const float* up, const float* weights, const float* xp,
                          const float* z, const float* mu, const float* sigma, const float* dur_f, const float* energy,
                          const float* pitch, const int64_t* lens, const float* wd, const float* bd, const float* rw, float* dx,
                          float* dwd, float* dbd, float* dwe, float* dbe, float* dwp, float* dbp, float* drw, float* drb,
                          float* scratch, int B, int L, int T, int D, void* stream);

/* ---- loss: replaces DaftExprtLoss.forward, loss.py:30-106 ------------------------------------------------------------
 * out[8] = {speaker, post_mult, duration, energy, pitch, mel_l1, mel_l2, total} (already weighted); acc: scratch [B,8]. */
int dx_loss_fwd(const float* spk_logits, const int64_t* spk_ids, const float* post, const float* dur_p, const float* energy_p,
                const float* pitch_p, const float* dur_t, const float* energy_t, const float* pitch_t, const float* mel_p,
                const float* mel_t, const int64_t* in_lens, const int64_t* out_lens, int B, int L, int T, int M, int NS, int NP,
                float w_adv, float w_post, float w_dur, float w_energy, float w_pitch, float w_mel, float* acc, float* out,
                void* stream);
int dx_loss_bwd(const float* gout, const float* spk_logits, const int64_t* spk_ids, const float* post, const float* dur_p,
                const float* energy_p, const float* pitch_p, const float* dur_t, const float* energy_t, const float* pitch_t,
                const float* mel_p, const float* mel_t, const int64_t* in_lens, const int64_t* out_lens, int B, int L, int T,
                int M, int NS, int NP, float w_adv, float w_post, float w_dur, float w_energy, float w_pitch, float w_mel,
                float* dspk_logits, float* dpost, float* ddur, float* denergy, float* dpitch, float* dmel, void* stream);

/* ---- inference-time controls ------------------------------------------------------------------------------------------
 * replaces DaftExprt.get_int_durations (model.py:789-812) + duration_to_integer (extract_features.py:69-111): BIT-EXACT.
 * dur_out = thresholded float durations, dur_int [B,L], totals[b] = sum, err[b] != 0 where the reference would raise. */
int dx_int_durations(const float* dur_pred, const float* dur_factors, const int64_t* lens, float* dur_out, int64_t* dur_int,
                     int64_t* totals, int* err, int B, int L, int sampling_rate, int filter_length, int hop_length,
                     int centered, void* stream);
/* model.py:895-899 */
int dx_inference_adjust(float* energy, float* pitch, const float* energy_factors, const int64_t* dur_int, int B, int L,
                        void* stream);
/* model.py:814-834; stats [n_speakers][2] = {mean, std} float32 */
int dx_pitch_shift(float* pitch, const float* factors, const int64_t* spk, const float* stats, int B, int L, void* stream);
/* model.py:836-864 */
int dx_pitch_multiply(float* pitch, const float* factors, int B, int L, void* stream);

/* ---- CUDA-graph support: per-step scalars read from DEVICE memory ---------------------------------------------------------
 * A captured training step (forward + loss + backward + Adam, ~600 launches replayed as one graph) must not bake per-step
 * scalars into kernel arguments.  dx_set_step_state(ptr) registers a device block of dx_step_state_bytes() = 32 bytes
 *     { uint64 seed_epoch; float w_adv; float lr; float bc1; float bc2_sqrt; float pad[2]; }
 * that the host rewrites (one 32-byte H2D copy on the same stream) before every replay.  While registered:
 *   - every dropout kernel (dx_ln_fwd/bwd, dx_attention_fwd/bwd) uses seed + seed_epoch * 0x9E3779B97F4A7C15,
 *   - dx_loss_fwd/bwd read the adversarial weight w_adv (loss.py:30-38) from the block instead of their argument,
 *   - dx_adam_step reads lr and the bias corrections 1 - beta1^t, sqrt(1 - beta2^t) from the block instead of lr / step.
 * ptr == NULL restores the by-value behaviour.  Process-global, like the GEMM backend. */
int dx_set_step_state(const void* device_state);
size_t dx_step_state_bytes(void);

/* ---- optimiser: fused Adam over a flat buffer (train.py:299-301,401) ---------------------------------------------------*/
/* clip (nullable): the device buffer dx_grad_norm_clip filled; its coefficient out[1] multiplies grad_scale (train.py:399). */
int dx_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int step, float grad_scale, const float* clip, void* stream);
/* Multi-GPU: gradient exchange FUSED with the optimiser over NVLink / NVSwitch peer memory — replaces DDP's all-reduce + optimizer.step()
 * (train.py:293,391,401) by one kernel per rank: reduce-scatter of this rank's shard of the flat gradient (sum over `world` ranks; ONE
 * multimem.ld_reduce per 16 bytes when grad_multicast != NULL: the NVSwitch adds the copies), Adam on the shard (m, v are only touched
 * inside the shard), all-gather of the updated parameters into every rank's buffer (multimem.st, or `world` peer stores).
 * The flat gradient and parameter buffers must be symmetric-memory allocations mapped on every rank (torch.distributed._symmetric_memory):
 * *_multicast = their multicast addresses (NULL: use the HOST arrays *_peer_ptrs of `world` per-rank device addresses).  shard_begin /
 * shard_n in elements, multiples of 4.  grad_scale = 1 / world for the mean.  The caller issues a cross-GPU barrier on the stream before
 * (every rank's gradients written) and after (every rank's parameters visible) the call.  Honours dx_set_step_state like dx_adam_step. */
int dx_fused_reduce_adam(const void* grad_multicast, const uint64_t* grad_peer_ptrs, void* param_multicast, const uint64_t* param_peer_ptrs,
                         const float* param_local, float* m, float* v, size_t shard_begin, size_t shard_n, int world, float lr, float beta1,
                         float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);
/* torch.nn.utils.clip_grad_norm_(parameters, max_norm) (train.py:399) over the flat gradient buffer, without a host sync:
 * out[0] = || grad_scale * g ||_2 (what the reference logs as grad_norm), out[1] = min(1, max_norm / (out[0] + 1e-6)),
 * out[2] = scratch.  max_norm = INFINITY only measures the norm.  out: 3 floats of device memory. */
int dx_grad_norm_clip(const float* g, size_t n, float grad_scale, float max_norm, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAFT_EXPRT_B200_H */
