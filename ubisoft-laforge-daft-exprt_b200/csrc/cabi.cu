// extern "C" surface of libdaftexprt_b200.so — thin argument marshalling over kernels.h.  See include/daft_exprt_b200.h.
#include <stdarg.h>
#include <string.h>

#include "../../include/daft_exprt_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace dx {

static thread_local char g_err[1024] = "";
static int g_backend = DX_GEMM_TCGEN05_BF16X3;   // default: tcgen05 bf16 hi|lo split, fp32-grade (ops.py mirrors this)
static const StepState* g_step_state = nullptr;   // graph mode: device block of per-step scalars (dx_set_step_state)
static unsigned long long g_launches = 0;   // kernels launched by this library (every launch goes through check_launch)

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return DX_ERR_CUDA;
    }
    return DX_OK;
}

}  // namespace dx

using namespace dx;
#define ST(s) ((cudaStream_t)(s))
typedef const long long* cll;

extern "C" {

const char* dx_last_error(void) { return g_err; }
int dx_abi_version(void) { return DX_ABI_VERSION; }
uint64_t dx_launch_count(void) { return g_launches; }
uint64_t dx_tc_gemm_launch_count(void) { return tc_gemm_launches(); }
int dx_set_step_state(const void* device_state) {
    g_step_state = static_cast<const StepState*>(device_state);
    return DX_OK;
}
size_t dx_step_state_bytes(void) { return sizeof(StepState); }

int dx_device_check(void) {
    int dev = 0;
    DX_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    DX_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_last_error("device %d is sm_%d%d; this library only contains sm_100a code", dev, prop.major, prop.minor);
        return DX_ERR_UNSUPPORTED;
    }
    return DX_OK;
}

int dx_set_gemm_backend(int backend) {
    DX_REQUIRE(backend == DX_GEMM_FP32_CUDA_CORES || backend == DX_GEMM_TCGEN05_TF32 || backend == DX_GEMM_TCGEN05_BF16X3,
               "unknown GEMM backend %d", backend);
    g_backend = backend;
    set_tc_precision(backend == DX_GEMM_TCGEN05_TF32);
    return DX_OK;
}
int dx_get_gemm_backend(void) { return g_backend; }
int dx_set_gemm_passes(int conv_passes, int wgrad_passes) {
    DX_REQUIRE(conv_passes >= 1 && conv_passes <= 3 && wgrad_passes >= 0 && wgrad_passes <= 3,
               "dx_set_gemm_passes: conv 1..3, wgrad 0 (by reduction length) or 1..3 (got %d, %d)", conv_passes, wgrad_passes);
    set_tc_passes(conv_passes, wgrad_passes);
    return DX_OK;
}
int dx_debug_set_trace(void* buf) { set_tc_trace((long long*)buf); return DX_OK; }
int dx_set_attention_backend(int forward_backend, int backward_backend) {
    DX_REQUIRE((forward_backend == 0 || forward_backend == 1) && (backward_backend == 0 || backward_backend == 1),
               "dx_set_attention_backend: backends must be DX_ATTENTION_MMA_SYNC (0) or DX_ATTENTION_TCGEN05 (1)");
    set_attention_backend(forward_backend, backward_backend);
    return DX_OK;
}

int dx_pack_conv_weight(const float* w, float* fwd, float* dgrad, void* fwd_planes, void* dgrad_planes, int Cout, int Cin, int KW,
                        int round_tf32, void* stream) {
    return pack_conv_weight(w, fwd, dgrad, fwd_planes, dgrad_planes, Cout, Cin, KW, round_tf32, ST(stream));
}

int dx_pack_conv_weights_batched(const void* descs_device, int n_desc, int total_blocks, int round_tf32, void* stream) {
    return pack_conv_weights_batched(descs_device, n_desc, total_blocks, round_tf32, ST(stream));
}

static ConvGemmArgs gemm_args(const float* x, const void* x_planes, const float* w_packed, const void* w_planes, const float* bias,
                              const float* relu_src, const float* add_src, float* y, void* ws, size_t wsb, int B, int S, int Cin,
                              int Cout, int KW, int ldx, int ldy, float alpha, int relu, int round_tf32,
                              const int64_t* lens = nullptr, int halo = 0, const void* relu_src_hi = nullptr, void* y_planes = nullptr,
                              float* y_colsum = nullptr) {
    ConvGemmArgs a;
    a.lens = (cll)lens; a.halo = halo; a.relu_src_hi = relu_src_hi; a.y_planes = y_planes; a.y_colsum = y_colsum;
    a.x = x; a.w = w_packed; a.bias = bias; a.relu_src = relu_src; a.add_src = add_src; a.y = y;
    a.B = B; a.S = S; a.Cin = Cin; a.Cout = Cout; a.KW = KW; a.ldx = ldx; a.ldy = ldy;
    a.alpha = alpha; a.relu = relu; a.round_tf32 = round_tf32;
    a.w_planes = w_planes; a.x_planes = x_planes; a.workspace = ws; a.workspace_bytes = wsb;
    return a;
}

int dx_split_weight_planes(const float* w_packed, void* planes, size_t n, void* stream) {
    return split_weight_planes(w_packed, planes, n, ST(stream));
}

int dx_split_planes(const float* x, int ld, void* planes, float* colsum_out, int rows, int C, void* stream) {
    return split_activation_planes(x, ld, planes, colsum_out, rows, C, ST(stream));
}

size_t dx_conv_gemm_workspace(int B, int S, int Cin, int Cout, int KW, int have_x_planes, int have_w_planes, int backend) {
    const int be = backend < 0 ? g_backend : backend;
    if (be == DX_GEMM_FP32_CUDA_CORES) return 0;
    ConvGemmArgs a = gemm_args(nullptr, have_x_planes ? (const void*)16 : nullptr, nullptr, have_w_planes ? (const void*)16 : nullptr,
                               nullptr, nullptr, nullptr, nullptr, nullptr, 0, B, S, Cin, Cout, KW, Cin, Cout, 1.f, 0, 0);
    return conv_gemm_tc_workspace(a);
}

int dx_conv_gemm(const float* x, const void* x_planes, const float* w_packed, const void* w_planes, const float* bias,
                 const float* relu_src, const void* relu_src_hi, const float* add_src, float* y, void* y_planes, float* y_colsum,
                 void* workspace, size_t workspace_bytes, const int64_t* lens, int halo, int B, int S, int Cin, int Cout, int KW,
                 int ldx, int ldy, float alpha, int relu, int round_tf32, int backend, void* stream) {
    ConvGemmArgs a = gemm_args(x, x_planes, w_packed, w_planes, bias, relu_src, add_src, y, workspace, workspace_bytes, B, S, Cin,
                               Cout, KW, ldx, ldy, alpha, relu, round_tf32, lens, halo, relu_src_hi, y_planes, y_colsum);
    DX_REQUIRE(B > 0 && S > 0 && Cin > 0 && Cout > 0 && (KW == 1 || KW == 3), "dx_conv_gemm: bad shape B=%d S=%d Cin=%d Cout=%d KW=%d", B, S, Cin, Cout, KW);
    const int be = backend < 0 ? g_backend : backend;
    if (be != DX_GEMM_FP32_CUDA_CORES && conv_gemm_tc_supported(a)) return conv_gemm_tc(a, ST(stream));
    DX_REQUIRE(!relu_src_hi && !y_planes && !y_colsum && y && x, "dx_conv_gemm: y_planes / y_colsum / relu_src_hi / plane-only operands need the tensor-core path (Cin=%d Cout=%d)", Cin, Cout);
    return conv_gemm_simt(a, ST(stream));
}

int dx_inproj_head_planes(const void* x_planes, const void* w_planes, const float* bias, void* head_planes, const int64_t* lens,
                          int B, int S, int Cin, int H, int dh, void* stream) {
    DX_REQUIRE(g_backend == DX_GEMM_TCGEN05_BF16X3, "dx_inproj_head_planes: needs the bf16x3 tensor-core backend");
    DX_REQUIRE(x_planes && w_planes && head_planes, "dx_inproj_head_planes: x_planes, w_planes and head_planes are required");
    AttnArgs at;
    memset(&at, 0, sizeof(at));
    at.H = H; at.dh = dh;
    DX_REQUIRE(attention_mma_supported(at), "dx_inproj_head_planes: head layout H=%d dh=%d has no tensor-core attention", H, dh);
    const int D = H * dh;
    ConvGemmArgs a = gemm_args(nullptr, x_planes, nullptr, w_planes, bias, nullptr, nullptr, nullptr, nullptr, 0, B, S, Cin, 3 * D, 1, Cin,
                               3 * D, 1.f, 0, 0, lens, 0);
    a.head_planes = (void*)(((uintptr_t)head_planes + 255) & ~(uintptr_t)255);   // same alignment rule as the attention kernels (bind_planes)
    a.head_dim = dh;
    a.head_Sp = (S + 63) / 64 * 64;
    a.head_scale_cols = D;
    a.head_scale = rsqrtf((float)dh);
    DX_REQUIRE(Cin % 8 == 0 && Cin >= 16, "dx_inproj_head_planes: Cin=%d", Cin);
    return conv_gemm_tc(a, ST(stream));
}

int dx_outproj_dgrad_head_planes(const void* dy_planes, const void* w_dgrad_planes, const float* ctx, void* bwd_scratch, const int64_t* lens,
                                 int B, int S, int Cin, int H, int dh, void* stream) {
    DX_REQUIRE(g_backend == DX_GEMM_TCGEN05_BF16X3, "dx_outproj_dgrad_head_planes: needs the bf16x3 tensor-core backend");
    DX_REQUIRE(dy_planes && w_dgrad_planes && ctx && bwd_scratch, "dx_outproj_dgrad_head_planes: dy_planes, w_dgrad_planes, ctx and bwd_scratch are required");
    AttnArgs at;
    memset(&at, 0, sizeof(at));
    at.H = H; at.dh = dh;
    DX_REQUIRE(attention_mma_supported(at) && Cin % 8 == 0 && Cin >= 16, "dx_outproj_dgrad_head_planes: H=%d dh=%d Cin=%d", H, dh, Cin);
    const int D = H * dh, Sp = (S + 63) / 64 * 64;
    ConvGemmArgs a = gemm_args(nullptr, dy_planes, nullptr, w_dgrad_planes, nullptr, nullptr, nullptr, nullptr, nullptr, 0, B, S, Cin, D, 1, Cin, D,
                               1.f, 0, 0, lens, 0);
    // same layout as dx_attention_bwd builds itself: [dO planes hi|lo : 2 * B * H * Sp * dh bf16][delta : B * H * S fp32], 256-byte aligned
    uint8_t* sb = (uint8_t*)(((uintptr_t)bwd_scratch + 255) & ~(uintptr_t)255);
    a.head_planes = sb;
    a.head_dim = dh;
    a.head_Sp = Sp;
    a.head_scale_cols = 0;
    a.head_dot_src = ctx;
    a.head_dot_out = (float*)(sb + (size_t)2 * B * H * Sp * dh * 2);
    return conv_gemm_tc(a, ST(stream));
}

int dx_conv_gemm_ln(const void* x_planes, const void* w_planes, const float* bias, const float* res, const float* ln_w, const float* ln_b,
                    const float* film, int film_stride, const int64_t* lens, float* y, void* y_planes, float* xhat, float* rstd, int B, int S,
                    int Cin, int KW, float p_in, uint64_t seed_in, void* stream) {
    DX_REQUIRE(g_backend == DX_GEMM_TCGEN05_BF16X3, "dx_conv_gemm_ln: needs the bf16x3 tensor-core backend");
    DX_REQUIRE(x_planes && w_planes && y && xhat && rstd, "dx_conv_gemm_ln: x_planes, w_planes, y, xhat and rstd are required");
    DX_REQUIRE(B > 0 && S > 0 && Cin >= 16 && Cin % 8 == 0 && (KW == 1 || KW == 3), "dx_conv_gemm_ln: bad shape B=%d S=%d Cin=%d KW=%d", B, S, Cin, KW);
    ConvGemmArgs a = gemm_args(nullptr, x_planes, nullptr, w_planes, bias, nullptr, nullptr, y, nullptr, 0, B, S, Cin, 128, KW, Cin, 128, 1.f, 0,
                               0, lens, 0, nullptr, y_planes);
    LnEpilogueArgs l;
    l.res = res; l.ln_w = ln_w; l.ln_b = ln_b; l.film = film; l.film_stride = film_stride; l.xhat = xhat; l.rstd = rstd;
    l.p_in = p_in; l.seed_in = seed_in; l.dyn = g_step_state;
    a.ln = &l;
    return conv_gemm_tc(a, ST(stream));
}

static ConvWgradArgs wgrad_args(const float* x, const void* x_planes, const float* dy, const void* dy_planes, float* dw,
                                float* dbias, void* ws, size_t wsb, int B, int S, int Cin, int Cout, int KW, int ldx, float alpha,
                                const int64_t* lens = nullptr, int halo = 0) {
    ConvWgradArgs a;
    a.lens = (cll)lens; a.halo = halo;
    a.x = x; a.dy = dy; a.dw = dw; a.dbias = dbias; a.workspace = ws; a.workspace_bytes = wsb;
    a.B = B; a.S = S; a.Cin = Cin; a.Cout = Cout; a.KW = KW; a.ldx = ldx; a.alpha = alpha;
    a.x_planes = x_planes; a.dy_planes = dy_planes;
    return a;
}

size_t dx_conv_wgrad_workspace(int B, int S, int Cin, int Cout, int KW, int have_x_planes, int have_dy_planes, int backend) {
    ConvWgradArgs a = wgrad_args(nullptr, have_x_planes ? (const void*)16 : nullptr, nullptr, have_dy_planes ? (const void*)16 : nullptr,
                                 nullptr, nullptr, nullptr, 0, B, S, Cin, Cout, KW, Cin, 1.f);
    size_t need = conv_wgrad_simt_workspace(a, nullptr);
    const int be = backend < 0 ? g_backend : backend;
    if (be != DX_GEMM_FP32_CUDA_CORES && conv_wgrad_tc_supported(a)) {
        const size_t t = conv_wgrad_tc_workspace(a);
        if (t > need) need = t;
    }
    return need;
}

int dx_conv_wgrad(const float* x, const void* x_planes, const float* dy, const void* dy_planes, float* dw, float* dbias,
                  void* workspace, size_t workspace_bytes, const int64_t* lens, int halo, int B, int S, int Cin, int Cout, int KW,
                  int ldx, float alpha, int backend, void* stream) {
    ConvWgradArgs a = wgrad_args(x, x_planes, dy, dy_planes, dw, dbias, workspace, workspace_bytes, B, S, Cin, Cout, KW, ldx, alpha,
                                 lens, halo);
    const int be = backend < 0 ? g_backend : backend;
    if (be != DX_GEMM_FP32_CUDA_CORES && conv_wgrad_tc_supported(a)) return conv_wgrad_tc(a, ST(stream));
    return conv_wgrad_simt(a, ST(stream));
}

int dx_colsum(const float* dy, float* db, int rows, int C, float alpha, void* stream) { return colsum(dy, db, rows, C, alpha, ST(stream)); }
int dx_wgrad_defer(int on) { wgrad_reduce_defer(on != 0); return DX_OK; }
int dx_wgrad_flush(void* stream) { return wgrad_reduce_flush(ST(stream)); }
int dx_colsum_planes(const void* planes, float* db, int rows, int C, void* stream) { return colsum_planes(planes, db, rows, C, ST(stream)); }
int dx_relu_bwd(const float* dy, const float* y, float* dx_, size_t n, void* stream) { return relu_bwd(dy, y, dx_, n, ST(stream)); }
int dx_scale_copy(const float* x, float* y, float alpha, size_t n, void* stream) { return scale_copy(x, y, alpha, n, ST(stream)); }

size_t dx_attention_planes_bytes(int B, int S, int H, int dh) { return attention_planes_bytes(B, S, H, dh); }
int dx_attention_uses_planes(int H, int dh) {
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.H = H; a.dh = dh;
    return (g_backend != DX_GEMM_FP32_CUDA_CORES && attention_mma_supported(a)) ? 1 : 0;
}
size_t dx_attention_bwd_scratch_bytes(int B, int S, int H, int dh) { return attention_bwd_scratch_bytes(B, S, H, dh); }

int dx_attention_fwd(const float* qkv, const int64_t* lens, float* ctx, float* lse, void* planes, void* ctx_planes, int B, int S,
                     int H, int dh, float dropout_p, uint64_t seed, void* stream) {
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.ctx_planes = ctx_planes;
    a.qkv = qkv; a.lens = (cll)lens; a.ctx = ctx; a.lse = lse; a.B = B; a.S = S; a.H = H; a.dh = dh;
    a.dropout_p = dropout_p; a.seed = seed;
    a.dyn = g_step_state;
    if (g_backend != DX_GEMM_FP32_CUDA_CORES && attention_mma_supported(a)) return attention_fwd_mma(a, planes, ST(stream));
    DX_REQUIRE(qkv != nullptr, "dx_attention_fwd: qkv == NULL (planes pre-filled by dx_inproj_head_planes) needs the tensor-core kernels");
    DX_REQUIRE(!ctx_planes, "dx_attention_fwd: ctx_planes is only written by the tensor-core kernels (dx_attention_uses_planes(%d, %d) == 0)", H, dh);
    return attention_fwd(a, ST(stream));
}

int dx_attention_bwd(const float* qkv, const void* planes, const int64_t* lens, const float* ctx, const float* lse,
                     const float* dctx, float* dqkv, void* scratch, int B, int S, int H, int dh, float dropout_p, uint64_t seed,
                     void* stream) {
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.qkv = qkv; a.lens = (cll)lens; a.ctx = (float*)ctx; a.lse = (float*)lse; a.dctx = dctx; a.dqkv = dqkv;
    a.B = B; a.S = S; a.H = H; a.dh = dh; a.dropout_p = dropout_p; a.seed = seed;
    a.dyn = g_step_state;
    DX_REQUIRE(scratch != nullptr, "dx_attention_bwd: scratch (dx_attention_bwd_scratch_bytes) required");
    DX_REQUIRE(dctx != nullptr || (g_backend != DX_GEMM_FP32_CUDA_CORES && attention_mma_supported(a)),
               "dx_attention_bwd: dctx == NULL (scratch pre-filled by dx_outproj_dgrad_head_planes) needs the tensor-core kernels");
    if (g_backend != DX_GEMM_FP32_CUDA_CORES && attention_mma_supported(a))
        return attention_bwd_mma(a, (void*)planes, scratch, ST(stream));
    a.delta = (float*)scratch;
    return attention_bwd(a, ST(stream));
}

int dx_ln_fwd(const float* a_, const float* res, const float* ln_w, const float* ln_b, const float* film, int film_stride,
              const int64_t* lens, float* y, float* xhat, float* rstd, void* y_planes, int B, int S, int D, float p_in,
              uint64_t seed_in, float p_out, uint64_t seed_out, void* stream) {
    LnArgs a;
    memset(&a, 0, sizeof(a));
    a.y_planes = y_planes;
    a.a = a_; a.res = res; a.ln_w = ln_w; a.ln_b = ln_b; a.film = film; a.film_stride = film_stride; a.lens = (cll)lens;
    a.y = y; a.xhat = xhat; a.rstd = rstd; a.B = B; a.S = S; a.D = D;
    a.p_in = p_in; a.p_out = p_out; a.seed_in = seed_in; a.seed_out = seed_out;
    a.dyn = g_step_state;
    return ln_fwd(a, ST(stream));
}

int dx_ln_bwd(const float* dy, const float* xhat, const float* rstd, const float* ln_w, const float* ln_b, const float* film,
              int film_stride, const int64_t* lens, const float* relu_src, float* dv, float* da, float* dln_w, float* dln_b, float* dfilm,
              void* g_planes, float* g_colsum, int B, int S, int D, float p_in, uint64_t seed_in, float p_out, uint64_t seed_out,
              void* stream) {
    LnArgs a;
    memset(&a, 0, sizeof(a));
    a.g_planes = g_planes; a.g_colsum = g_colsum;
    a.dy = dy; a.xhat = (float*)xhat; a.rstd = (float*)rstd; a.ln_w = ln_w; a.ln_b = ln_b; a.film = film;
    a.film_stride = film_stride; a.lens = (cll)lens; a.relu_src = relu_src; a.dv = dv; a.da = da; a.dln_w = dln_w; a.dln_b = dln_b; a.dfilm = dfilm;
    a.B = B; a.S = S; a.D = D; a.p_in = p_in; a.p_out = p_out; a.seed_in = seed_in; a.seed_out = seed_out;
    a.dyn = g_step_state;
    DX_REQUIRE(!(p_in > 0.f) || da, "dx_ln_bwd: da must be provided when p_in > 0");
    if (!(p_in > 0.f)) a.da = nullptr;
    return ln_bwd(a, ST(stream));
}

int dx_embed_pe_fwd(const int64_t* symbols, const int64_t* lens, const float* emb, const float* pe, float* y, int B, int L,
                    int D, int n_symbols, void* stream) {
    return embed_pe_fwd((cll)symbols, (cll)lens, emb, pe, y, B, L, D, n_symbols, ST(stream));
}
int dx_embed_pe_bwd(const int64_t* symbols, const int64_t* lens, const float* dy, float* demb, int B, int L, int D,
                    int n_symbols, void* stream) {
    return embed_pe_bwd((cll)symbols, (cll)lens, dy, demb, B, L, D, n_symbols, ST(stream));
}
int dx_frame_input_fwd(const float* x, const float* energy, const float* pitch, const float* we, const float* be,
                       const float* wp, const float* bp, const float* pe, const int64_t* lens, float* y, int B, int T, int D,
                       void* stream) {
    return frame_input_fwd(x, energy, pitch, we, be, wp, bp, pe, (cll)lens, y, B, T, D, ST(stream));
}
int dx_frame_input_bwd(const float* dy, const float* energy, const float* pitch, const int64_t* lens, float* dx_, float* dwe,
                       float* dbe, float* dwp, float* dbp, int B, int T, int D, void* stream) {
    return frame_input_bwd(dy, energy, pitch, (cll)lens, dx_, dwe, dbe, dwp, dbp, B, T, D, ST(stream));
}
int dx_meanpool_fwd(const float* x, const int64_t* lens, float* pooled, int B, int S, int D, void* stream) {
    return meanpool_fwd(x, (cll)lens, pooled, B, S, D, ST(stream));
}
int dx_meanpool_bwd(const float* dpooled, const int64_t* lens, float* dx_, int B, int S, int D, void* stream) {
    return meanpool_bwd(dpooled, (cll)lens, dx_, B, S, D, ST(stream));
}
int dx_add_speaker_fwd(const float* pooled, const int64_t* spk, const float* spk_emb, float* h, int B, int D, int n_spk,
                       void* stream) {
    return add_speaker_fwd(pooled, (cll)spk, spk_emb, h, B, D, n_spk, ST(stream));
}
int dx_add_speaker_bwd(const float* dh, const int64_t* spk, float* dspk_emb, int B, int D, int n_spk, void* stream) {
    return add_speaker_bwd(dh, (cll)spk, dspk_emb, B, D, n_spk, ST(stream));
}

static int film_layout(FilmLayout& lay, int n_modules, const int* nb_blocks, const int* channels) {
    DX_REQUIRE(n_modules >= 1 && n_modules <= 4, "film: n_modules=%d (1..4)", n_modules);
    lay.n_modules = n_modules;
    for (int m = 0; m < n_modules; ++m) { lay.nb_blocks[m] = nb_blocks[m]; lay.channels[m] = channels[m]; }
    return DX_OK;
}
int dx_film_assemble_fwd(const float* graw, const float* braw, const float* post, float* film, int B, int n_modules,
                         const int* nb_blocks, const int* channels, void* stream) {
    FilmLayout lay;
    int rc = film_layout(lay, n_modules, nb_blocks, channels);
    if (rc) return rc;
    return film_assemble_fwd(graw, braw, post, film, B, lay, ST(stream));
}
int dx_film_assemble_bwd(const float* dfilm, const float* graw, const float* braw, const float* post, float* dgraw,
                         float* dbraw, float* dpost, int B, int n_modules, const int* nb_blocks, const int* channels,
                         void* stream) {
    FilmLayout lay;
    int rc = film_layout(lay, n_modules, nb_blocks, channels);
    if (rc) return rc;
    return film_assemble_bwd(dfilm, graw, braw, post, dgraw, dbraw, dpost, B, lay, ST(stream));
}
int dx_narrow_linear_fwd(const float* x, const float* w, const float* bias, const int64_t* lens, float* out, int B, int S,
                         int C, int NO, void* stream) {
    return narrow_linear_fwd(x, w, bias, (cll)lens, out, B, S, C, NO, 1, ST(stream));
}
int dx_narrow_linear_bwd(const float* dout, const float* x, const float* w, const int64_t* lens, float* dx_, float* dw,
                         float* db, int B, int S, int C, int NO, void* stream) {
    return narrow_linear_bwd(dout, x, w, (cll)lens, dx_, dw, db, B, S, C, NO, 1, ST(stream));
}
int dx_mask_transpose_fwd(const float* y, const int64_t* lens, float* mel, int B, int T, int M, void* stream) {
    return mask_transpose_fwd(y, (cll)lens, mel, B, T, M, ST(stream));
}
int dx_mask_transpose_bwd(const float* dmel, const int64_t* lens, float* dy, int B, int T, int M, void* stream) {
    return mask_transpose_bwd(dmel, (cll)lens, dy, B, T, M, ST(stream));
}

int dx_gauss_prep(const float* x, const float* dur_f, const int64_t* dur_i, const float* energy, const float* pitch,
                  const int64_t* lens, const float* wd, const float* bd, const float* we, const float* be, const float* wp,
                  const float* bp, const float* rw, const float* rb, float* xp, float* z, float* sigma, float* mu,
                  int64_t* csum, int64_t* total, int B, int L, int D, void* stream) {
    GaussArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.dur_f = dur_f; a.dur_i = (cll)dur_i; a.energy = energy; a.pitch = pitch; a.lens = (cll)lens;
    a.wd = wd; a.bd = bd; a.we = we; a.be = be; a.wp = wp; a.bp = bp; a.rw = rw; a.rb = rb;
    a.xp = xp; a.z = z; a.sigma = sigma; a.mu = mu; a.csum = (long long*)csum; a.total = (long long*)total;
    a.B = B; a.L = L; a.D = D;
    return gauss_prep(a, ST(stream));
}
int dx_gauss_upsample_fwd(const float* xp, const float* mu, const float* sigma, const int64_t* lens, float* up, float* weights,
                          int B, int L, int T, int D, void* stream) {
    GaussArgs a;
    memset(&a, 0, sizeof(a));
    a.xp = (float*)xp; a.mu = (float*)mu; a.sigma = (float*)sigma; a.lens = (cll)lens; a.up = up; a.weights = weights;
    a.B = B; a.L = L; a.T = T; a.D = D;
    return gauss_upsample_fwd(a, ST(stream));
}
int dx_gauss_upsample_bwd(const float* dup, const float* dweights, const float* up, const float* weights, const float* xp,
                          const float* z, const float* mu, const float* sigma, const float* dur_f, const float* energy,
                          const float* pitch, const int64_t* lens, const float* wd, const float* bd, const float* rw, float* dx_,
                          float* dwd, float* dbd, float* dwe, float* dbe, float* dwp, float* dbp, float* drw, float* drb,
                          float* scratch, int B, int L, int T, int D, void* stream) {
    GaussArgs a;
    memset(&a, 0, sizeof(a));
    a.dup = dup; a.dweights = dweights; a.up = (float*)up; a.weights = (float*)weights; a.xp = (float*)xp; a.z = (float*)z;
    a.mu = (float*)mu; a.sigma = (float*)sigma; a.dur_f = dur_f; a.energy = energy; a.pitch = pitch; a.lens = (cll)lens;
    a.wd = wd; a.bd = bd; a.rw = rw; a.dx = dx_; a.dsigma = scratch;
    a.dwd = dwd; a.dbd = dbd; a.dwe = dwe; a.dbe = dbe; a.dwp = dwp; a.dbp = dbp; a.drw = drw; a.drb = drb;
    a.B = B; a.L = L; a.T = T; a.D = D;
    return gauss_upsample_bwd(a, ST(stream));
}

static LossArgs loss_args(const float* spk_logits, const int64_t* spk_ids, const float* post, const float* dur_p,
                          const float* energy_p, const float* pitch_p, const float* dur_t, const float* energy_t,
                          const float* pitch_t, const float* mel_p, const float* mel_t, const int64_t* in_lens,
                          const int64_t* out_lens, int B, int L, int T, int M, int NS, int NP, float w_adv, float w_post,
                          float w_dur, float w_energy, float w_pitch, float w_mel) {
    LossArgs a;
    memset(&a, 0, sizeof(a));
    a.spk_logits = spk_logits; a.spk_ids = (cll)spk_ids; a.post = post; a.dur_p = dur_p; a.energy_p = energy_p;
    a.pitch_p = pitch_p; a.dur_t = dur_t; a.energy_t = energy_t; a.pitch_t = pitch_t; a.mel_p = mel_p; a.mel_t = mel_t;
    a.in_lens = (cll)in_lens; a.out_lens = (cll)out_lens; a.B = B; a.L = L; a.T = T; a.M = M; a.NS = NS; a.NP = NP;
    a.w_adv = w_adv; a.w_post = w_post; a.w_dur = w_dur; a.w_energy = w_energy; a.w_pitch = w_pitch; a.w_mel = w_mel;
    a.dyn = g_step_state;
    return a;
}
int dx_loss_fwd(const float* spk_logits, const int64_t* spk_ids, const float* post, const float* dur_p, const float* energy_p,
                const float* pitch_p, const float* dur_t, const float* energy_t, const float* pitch_t, const float* mel_p,
                const float* mel_t, const int64_t* in_lens, const int64_t* out_lens, int B, int L, int T, int M, int NS, int NP,
                float w_adv, float w_post, float w_dur, float w_energy, float w_pitch, float w_mel, float* acc, float* out,
                void* stream) {
    LossArgs a = loss_args(spk_logits, spk_ids, post, dur_p, energy_p, pitch_p, dur_t, energy_t, pitch_t, mel_p, mel_t, in_lens,
                           out_lens, B, L, T, M, NS, NP, w_adv, w_post, w_dur, w_energy, w_pitch, w_mel);
    a.acc = acc; a.out = out;
    return loss_fwd(a, ST(stream));
}
int dx_loss_bwd(const float* gout, const float* spk_logits, const int64_t* spk_ids, const float* post, const float* dur_p,
                const float* energy_p, const float* pitch_p, const float* dur_t, const float* energy_t, const float* pitch_t,
                const float* mel_p, const float* mel_t, const int64_t* in_lens, const int64_t* out_lens, int B, int L, int T,
                int M, int NS, int NP, float w_adv, float w_post, float w_dur, float w_energy, float w_pitch, float w_mel,
                float* dspk_logits, float* dpost, float* ddur, float* denergy, float* dpitch, float* dmel, void* stream) {
    LossArgs a = loss_args(spk_logits, spk_ids, post, dur_p, energy_p, pitch_p, dur_t, energy_t, pitch_t, mel_p, mel_t, in_lens,
                           out_lens, B, L, T, M, NS, NP, w_adv, w_post, w_dur, w_energy, w_pitch, w_mel);
    a.gout = gout; a.dspk_logits = dspk_logits; a.dpost = dpost; a.ddur = ddur; a.denergy = denergy; a.dpitch = dpitch;
    a.dmel = dmel;
    return loss_bwd(a, ST(stream));
}

int dx_int_durations(const float* dur_pred, const float* dur_factors, const int64_t* lens, float* dur_out, int64_t* dur_int,
                     int64_t* totals, int* err, int B, int L, int sampling_rate, int filter_length, int hop_length,
                     int centered, void* stream) {
    return int_durations(dur_pred, dur_factors, (cll)lens, dur_out, (long long*)dur_int, (long long*)totals, err, B, L,
                         sampling_rate, filter_length, hop_length, centered, ST(stream));
}
int dx_inference_adjust(float* energy, float* pitch, const float* energy_factors, const int64_t* dur_int, int B, int L,
                        void* stream) {
    return inference_adjust(energy, pitch, energy_factors, (cll)dur_int, B, L, ST(stream));
}
int dx_pitch_shift(float* pitch, const float* factors, const int64_t* spk, const float* stats, int B, int L, void* stream) {
    return pitch_shift(pitch, factors, (cll)spk, stats, B, L, ST(stream));
}
int dx_pitch_multiply(float* pitch, const float* factors, int B, int L, void* stream) {
    return pitch_multiply(pitch, factors, B, L, ST(stream));
}
int dx_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int step, float grad_scale, const float* clip, void* stream) {
    return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, g_step_state, clip, ST(stream));
}
int dx_fused_reduce_adam(const void* grad_multicast, const uint64_t* grad_peer_ptrs, void* param_multicast, const uint64_t* param_peer_ptrs,
                         const float* param_local, float* m, float* v, size_t shard_begin, size_t shard_n, int world, float lr, float beta1,
                         float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
    return fused_reduce_adam((const float*)grad_multicast, (const unsigned long long*)grad_peer_ptrs, (float*)param_multicast,
                             (const unsigned long long*)param_peer_ptrs, param_local, m, v, shard_begin, shard_n, world, lr, beta1, beta2, eps,
                             weight_decay, step, grad_scale, g_step_state, ST(stream));
}
int dx_grad_norm_clip(const float* g, size_t n, float grad_scale, float max_norm, float* out, void* stream) {
    return grad_norm_clip(g, n, grad_scale, max_norm, out, ST(stream));
}

}  // extern "C"
