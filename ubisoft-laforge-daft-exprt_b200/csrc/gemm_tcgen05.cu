// tcgen05 / TMEM / TMA implementation of the channels-last Conv1d / Linear GEMMs (forward, dgrad, wgrad) for sm_100a.
//
//   forward/dgrad:  y[b, s, n] = epi( alpha * sum_{tap < KW} sum_{c < Cin} x[b, s + tap - pad, c] * w[tap][n][c] + bias[n] )
//   wgrad:          part[split][tap][co][ci] = sum_{(b, s) in split} dy[b, s, co] * x[b, s + tap - pad, ci]
//                   (same bf16 row planes as forward/dgrad, consumed as MN-major UMMA operands: K = rows)
//
// im2col-free: the KW taps are KW shifted K-slices of the same activation tensor.  Operand tiles are fetched by 3-D TMA
// tensor maps whose s-coordinate is simply offset by (tap - pad); rows that fall outside [0, S) are zero-filled by the TMA
// unit, which IS the convolution's zero padding (it can never bleed into the neighbouring utterance because the batch
// index is its own tensor dimension).
//
// Precision modes (the reference computes in fp32 and parity is the first gate):
//   BF16X3  every fp32 operand is split as hi + lo (two bf16 planes, |x - hi - lo| <= 2^-17 |x|) and the product is formed as
//           hi*hi + lo*hi + hi*lo by three `tcgen05.mma kind::f16` per K-step with fp32 accumulation in TMEM: fp32-grade
//           results (~1e-5) at 3 bf16 tensor-core passes.  Default.
//   TF32    fp32 tiles are consumed directly by `tcgen05.mma kind::tf32` (10-bit mantissa): one pass at half the bf16 rate,
//           ~1e-3 per GEMM.  Forward/dgrad only.
//
// CTA = 320 threads, persistent over 128 x 128 output tiles, one CTA per SM:
//   warp 0    TMA producer: ring of K-stages (128-byte rows, SWIZZLE_128B), mbarrier complete_tx
//   warp 1    MMA issuer (one thread): UMMA M128 N128, tcgen05.commit frees the stage; two TMEM accumulators (2 x 128
//             columns) so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2-9 epilogue (two per SM sub-partition; 32 lanes x 64 columns each): tcgen05.ld -> registers -> bias / ReLU / ReLU-mask / residual add ->
//             128B-swizzled smem staging -> TMA store (clips partial tiles)
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace dx {

namespace {

enum { PREC_TF32 = 0, PREC_BF16X3 = 1 };
enum { MODE_CONV = 0, MODE_WGRAD = 1, MODE_HALO = 2 };
// EPI_LN (Cout = 128 = ONE column tile, BF16X3): the epilogue finishes the reference's sub-layer tail in place of a separate kernel
//   v = dropout(acc + bias) + residual;  xhat = (v - mean) * rstd;  y = mask(film_g * (xhat * ln_w + ln_b) + film_b)
// (model.py:189-191,259 after the attention out-projection; model.py:226-235,262 after conv2) and writes y (fp32 + bf16 hi|lo
// operand planes for the next GEMM), xhat and rstd (saved for backward).  The two epilogue warps that share a TMEM lane quarter
// (64 columns each) exchange their partial row sums through shared memory.
enum { EPI_STD = 0, EPI_LN = 1 };
// MODE_HALO = MODE_CONV for KW = 3 (BF16X3 only): the three taps are three row-shifted views of ONE activation tile, so the
// tile is fetched once per channel chunk with its halo (rows s0-1 .. s0+128) and each tap's MMAs read it through a UMMA
// descriptor whose start address is advanced by tap rows.  A traffic drops 3x, total operand traffic by a third.

constexpr int TM = 128, TN = 128;
constexpr int TILE_BYTES = 128 * 128;                 // one operand tile: 128 rows x 128 bytes
constexpr int OUT_BYTES = TM * 32 * 4;                // 2 * OUT_BYTES = 16 epilogue warps x 2 KB of store staging
constexpr int TMEM_COLS = 256;
constexpr int NUM_EPI_WARPS = 16;                     // four per SM sub-partition (see the epilogue)
constexpr int NTHREADS = 64 + 32 * NUM_EPI_WARPS;     // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2.. = epilogue

template <int PREC> struct Cfg;
template <> struct Cfg<PREC_TF32>   { static constexpr int TKB = 32, PLANES = 1, NSTAGE = 5; };
template <> struct Cfg<PREC_BF16X3> { static constexpr int TKB = 64, PLANES = 2, NSTAGE = 3; };
template <int PREC> constexpr int stage_bytes() { return 2 * Cfg<PREC>::PLANES * TILE_BYTES; }
constexpr int HALO_ROWS = TM + 2;                      // KW = 3
constexpr int A_HALO_BYTES = 17 * 1024;                // 130 rows x 128 B rounded up to the 1024-byte swizzle repeat
constexpr int HALO_A_STAGES = 2, HALO_B_STAGES = 3;
constexpr int HALO_A_STAGE_BYTES = 2 * A_HALO_BYTES, HALO_B_STAGE_BYTES = 2 * TILE_BYTES;
constexpr int XCH_BYTES = 2 * 128 * 4 * 4;             // EPI_LN: two exchange slots x 128 rows x four column chunks
// k=1 GEMMs with the LayerNorm epilogue (the out-projection: K = 128 = two chunks) run a 2-stage ring to make room for it
template <int PREC, int MODE, int EPI> constexpr int num_stages() { return (EPI == EPI_LN && MODE == MODE_CONV) ? 2 : Cfg<PREC>::NSTAGE; }
template <int PREC, int MODE, int EPI = EPI_STD> constexpr int smem_bytes() {
    return (MODE == MODE_HALO ? HALO_A_STAGES * HALO_A_STAGE_BYTES + HALO_B_STAGES * HALO_B_STAGE_BYTES
                              : num_stages<PREC, MODE, EPI>() * stage_bytes<PREC>()) + 2 * OUT_BYTES + 256 + 1024 +
           (EPI == EPI_LN ? XCH_BYTES : 0);
}

// ---- PTX wrappers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one lane of the (converged) warp; the same lane every time for the same membermask
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile in shared memory, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 B apart (SBO); LBO unused.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// same, start address `rows` rows (128 B each) past a 1024-byte aligned tile.  Measured on B200: the 128B swizzle is applied
// to the absolute shared-memory address bits (exactly like the TMA write that filled the tile), so a row-shifted start
// address needs NO matrix base offset (bits 49-51 stay 0); setting base offset = rows gives wrong products.
__device__ __forceinline__ uint64_t umma_desc_k_sw128_rows(uint32_t tile, uint32_t rows) {
    return umma_desc_k_sw128(tile + rows * 128u);
}
// MN-major operand tile (K = rows): 64-element (128-byte) MN chunks, each chunk = [K rows][128 B] with SWIZZLE_128B;
// LBO = distance between MN chunks (8192 B: one 64-row TMA box), SBO = distance between 8-row K groups (1024 B).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: fp32 accumulate, M x N tile; fmt 1 = bf16 (kind::f16), 2 = tf32 (kind::tf32); mn_major = 1 when
// both operands are MN-major (wgrad), 0 when both are K-major
__host__ __device__ constexpr uint32_t umma_idesc(int fmt, int M, int N, int mn_major = 0) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int PREC>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (PREC == PREC_TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
// completes the tcgen05.ld's above; the 32 registers are tied to the wait so no consumer can be scheduled ahead of it
__device__ __forceinline__ void tmem_ld_wait(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                   "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}

struct TcParams {
    const float* bias;
    const float* relu_src;
    const float* add_src;
    const __nv_bfloat16* relu_src_hi;   // ReLU mask from a bf16 hi plane [B*S][Cout] (Cout % 32 == 0)
    float* colsum;                      // optional [Cout]: column sums of the stored output (fp32 atomics, one per warp and chunk)
    __nv_bfloat16* y_planes;            // optional: output as bf16 hi|lo planes [2][B*S][Cout] (Cout % 32 == 0)
    long long y_plane_elems;            // B*S*Cout
    int skip_y;                         // no fp32 output (planes only)
    // head-plane output (attention in-projection, skip_y = 1): the 32-column chunks go to per-head operand planes
    // R[2][B][hp_NH][hp_Sp][hp_dh] (map_y is the (dh, Sp, 2*B*NH) map over R); columns < hp_scale_cols are multiplied by hp_scale
    // (q pre-scaled by 1/sqrt(dh)); rows >= S are written as zeros up to Sp.  hp_dh == 0: off.
    int hp_dh, hp_NH, hp_scale_cols;
    float hp_scale;
    // optional with head planes (attention backward): hp_dot_out[b][head][s] += sum over the head's columns of out * hp_dot_src
    // (delta = rowsum(dO * O) of the softmax backward; zero-initialised by the launcher), hp_dot_src fp32 [B*S][Cout]
    const float* hp_dot_src;
    float* hp_dot_out;
    // EPI_LN
    const float* ln_res;                // residual [B*S][Cout] or nullptr
    const float* ln_w;
    const float* ln_b;
    const float* film;                  // nullable: gamma at film[b * film_stride + c], beta at film[b * film_stride + Cout + c]
    int film_stride;
    float* ln_rstd;                     // [B*S]
    float ln_p_in;                      // dropout on (acc + bias) before the residual add
    unsigned long long ln_seed_in;
    const StepState* dyn;
    int ln_planes;                      // also write y as bf16 hi|lo planes (map_y3)
    // BF16X3 pass count (precision ablation / reduced-precision modes): 3 = hi*hi + lo*hi + hi*lo (default, fp32-grade);
    // 2 = hi*hi + lo*hi (the B operand = weights / second operand rounded to bf16); 1 = hi*hi (plain bf16).  Planes that a mode
    // does not use are not loaded.
    int passes;
    int B, S, Cin, Cout, KW, ldy;
    int tiles_m_per_b, tiles_n, num_tiles, k_chunks;   // CONV: tiles over (b, s) x n;  WGRAD: tiles_m = co tiles, tiles_n = ci tiles
    int nsplit;                                         // WGRAD: batch ranges
    const long long* lens;   // optional [B]: rows >= len[b] + halo cannot reach a valid output -> tiles / K-chunks skipped
    int halo;
    float alpha;
    int relu, round_tf32;
    long long* trace;   // optional [4][256] clock64 trace of block 0 (DX_TC_TRACE), else nullptr
    int debug;   // bring-up bisection mask (DX_TC_DEBUG): 1 no MMA, 2 no TMA loads, 4 no TMEM loads, 8 no TMA store, 16 no TMEM alloc,
                 // 64 no proxy fence before the TMA store, 256 no bias loads (timing probes: results are wrong under any bit)
};

struct TileCoord {
    int a0, a1, a2;   // first-stage A coordinates (coordinate 0 advances with the K chunk)
    int b0, b1, b2;
    int o0, o1, o2;   // output coordinates of column chunk 0
    int tap, k_begin, k_end;   // K iterations [k_begin, k_end): k -> (outer = k / k_chunks, chunk = k % k_chunks)
};

// (x, y) -> packed bf16 pairs {lo16 = x, hi16 = y}: hi = bf16(v), lo = bf16(v - hi)
__device__ __forceinline__ void split_pair_u32(float x, float y, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y), "f"(x));
    const float xr = x - __uint_as_float(hi << 16), yr = y - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(yr), "f"(xr));
}

template <int MODE>
__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, int tile) {
    TileCoord t;
    if constexpr (MODE != MODE_WGRAD) {
        const int tn = tile % p.tiles_n, tm = tile / p.tiles_n;
        const int b = tm / p.tiles_m_per_b, s0 = (tm % p.tiles_m_per_b) * TM, n0 = tn * TN;
        t.a0 = 0; t.a1 = s0; t.a2 = b;
        t.b0 = 0; t.b1 = n0; t.b2 = 0;
        t.o0 = n0; t.o1 = s0; t.o2 = b;
        t.tap = 0; t.k_begin = 0; t.k_end = p.KW * p.k_chunks;   // outer = tap
    } else {
        // tile = ((split * KW + tap) * tiles_m + tco) * tiles_n + tci
        const int tci = tile % p.tiles_n;
        int r = tile / p.tiles_n;
        const int tco = r % p.tiles_m_per_b;
        r /= p.tiles_m_per_b;
        const int tap = r % p.KW, split = r / p.KW;
        t.a0 = 0; t.a1 = 0; t.a2 = tco * TM;
        t.b0 = 0; t.b1 = 0; t.b2 = tci * TN;
        t.o0 = tci * TN; t.o1 = tco * TM; t.o2 = split * p.KW + tap;
        t.tap = tap;
        const long long total_k = (long long)p.B * p.k_chunks;                 // outer = utterance
        t.k_begin = (int)((split * total_k) / p.nsplit);
        t.k_end = (int)(((split + 1) * total_k) / p.nsplit);
    }
    return t;
}

// ---- padding skip: rows at or beyond len[b] + halo of an utterance never influence a valid output (the caller chooses halo
// from the receptive field of what follows), so CONV output tiles made only of such rows are written as zeros without any
// loads or MMAs, and WGRAD K-chunks made only of such rows (where dy is exactly zero) are not accumulated.
__device__ __forceinline__ int live_rows(const TcParams& p, int b) { return min((int)p.lens[b], p.S) + p.halo; }

template <int MODE>
__device__ __forceinline__ bool k_dead(const TcParams& p, int k) {   // WGRAD only
    if (MODE != MODE_WGRAD || p.lens == nullptr) return false;
    const int b = k / p.k_chunks, kc = k - b * p.k_chunks;
    return kc * 64 >= live_rows(p, b);
}

template <int MODE>
__device__ __forceinline__ bool tile_dead(const TcParams& p, const TileCoord& t) {
    if (p.lens == nullptr) return false;
    if (MODE != MODE_WGRAD) return t.a1 >= live_rows(p, t.a2);
    for (int b = t.k_begin / p.k_chunks; b <= (t.k_end - 1) / p.k_chunks; ++b) {
        const int kc_lo = max(t.k_begin - b * p.k_chunks, 0);
        if (kc_lo * 64 < live_rows(p, b)) return false;
    }
    return true;
}

template <int PREC, int MODE, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi,
                                                         const __grid_constant__ CUtensorMap map_a_lo,
                                                         const __grid_constant__ CUtensorMap map_b_hi,
                                                         const __grid_constant__ CUtensorMap map_b_lo,
                                                         const __grid_constant__ CUtensorMap map_y,
                                                         const __grid_constant__ CUtensorMap map_y2,   // EPI_LN: xhat (fp32)
                                                         const __grid_constant__ CUtensorMap map_y3,   // EPI_LN: y planes (bf16)
                                                         TcParams p) {
    constexpr int NSTAGE = num_stages<PREC, MODE, EPI>(), PLANES = Cfg<PREC>::PLANES, TKB = Cfg<PREC>::TKB;
    static_assert(EPI == EPI_STD || (PREC == PREC_BF16X3 && MODE != MODE_WGRAD), "EPI_LN is a BF16X3 forward epilogue");
    constexpr int STAGE_BYTES = stage_bytes<PREC>();
    static_assert(MODE != MODE_HALO || PREC == PREC_BF16X3, "MODE_HALO is a BF16X3 variant");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t baseB = base + HALO_A_STAGES * HALO_A_STAGE_BYTES;   // MODE_HALO: A ring, then B ring
    const uint32_t sOut = MODE == MODE_HALO ? baseB + HALO_B_STAGES * HALO_B_STAGE_BYTES : base + NSTAGE * STAGE_BYTES;
    const uint32_t bars = sOut + 2 * OUT_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * NSTAGE, tfull0 = bars + 16 * NSTAGE, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;
    const uint32_t afull0 = bars + 96, aempty0 = bars + 112;   // MODE_HALO: full0/empty0 are the B ring (NSTAGE = HALO_B_STAGES = 3)
    static_assert(MODE != MODE_HALO || NSTAGE == HALO_B_STAGES, "B ring shares the full/empty barrier slots");
    const uint32_t xch = bars + 256;   // EPI_LN: [slot][half][row] fp32
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pad = (p.KW - 1) / 2;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < HALO_A_STAGES; ++i) { mbar_init(afull0 + 8 * i, 1); mbar_init(aempty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, NUM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (p.debug & 16) {
            if (lane == 0) *tmem_slot_ptr = 0u;
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (MODE == MODE_HALO && warp == 0) {
        if (lane == 0) {
            int bs = 0, bph = 0, as = 0, aph = 0, gk = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const TileCoord t = tile_coord<MODE>(p, tile);
                if (tile_dead<MODE>(p, t)) continue;
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(aempty0 + 8 * as, aph ^ 1);
                    const uint32_t sa = base + as * HALO_A_STAGE_BYTES, abar = afull0 + 8 * as;
                    if (p.debug & 2) {
                        mbar_arrive(abar);
                    } else {   // rows s0-1 .. s0+128 of this channel chunk, both planes (out-of-range rows arrive as zeros)
                        mbar_expect_tx(abar, (p.passes >= 2 ? 2 : 1) * HALO_ROWS * 128);
                        tma_load_3d(sa, &map_a_hi, abar, kc * TKB, t.a1 - 1, t.a2);
                        if (p.passes >= 2) tma_load_3d(sa + A_HALO_BYTES, &map_a_lo, abar, kc * TKB, t.a1 - 1, t.a2);
                    }
                    if (++as == HALO_A_STAGES) { as = 0; aph ^= 1; }
                    for (int tap = 0; tap < 3; ++tap) {
                        mbar_wait(empty0 + 8 * bs, bph ^ 1);
                        const uint32_t sb = baseB + bs * HALO_B_STAGE_BYTES, bbar = full0 + 8 * bs;
                        if (p.debug & 2) {
                            mbar_arrive(bbar);
                        } else {
                            mbar_expect_tx(bbar, p.passes >= 3 ? HALO_B_STAGE_BYTES : TILE_BYTES);
                            tma_load_3d(sb, &map_b_hi, bbar, kc * TKB, t.b1, tap);
                            if (p.passes >= 3) tma_load_3d(sb + TILE_BYTES, &map_b_lo, bbar, kc * TKB, t.b1, tap);
                        }
                        if (p.trace && blockIdx.x == 0 && gk < 256) p.trace[0 * 256 + gk] = clock64();
                        ++gk;
                        if (++bs == HALO_B_STAGES) { bs = 0; bph ^= 1; }
                    }
                }
            }
        }
    } else if (MODE == MODE_HALO && warp == 1) {
        // The whole warp runs the issue loop (warp-uniform control flow keeps descriptors in uniform registers; a divergent
        // `if (lane == 0)` costs ~12 instructions and a broadcast loop per MMA); one elected lane issues MMAs and commits.
        constexpr uint32_t idesc = umma_idesc(1, TM, TN, 0);
        int bs = 0, bph = 0, as = 0, aph = 0, it = 0, gk = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const TileCoord t = tile_coord<MODE>(p, tile);
            if (tile_dead<MODE>(p, t)) continue;
            const int acc = it & 1, acc_phase = (it >> 1) & 1;
            ++it;
            mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * TN;
            uint32_t started = 0;
            for (int kc = 0; kc < p.k_chunks; ++kc) {
                mbar_wait(afull0 + 8 * as, aph);
                const uint32_t sa = base + as * HALO_A_STAGE_BYTES;
                for (uint32_t tap = 0; tap < 3; ++tap) {
                    mbar_wait(full0 + 8 * bs, bph);
                    tc_fence_after();
                    const uint32_t sb = baseB + bs * HALO_B_STAGE_BYTES;
                    const uint64_t a_hi = umma_desc_k_sw128_rows(sa, tap);
                    const uint64_t a_lo = umma_desc_k_sw128_rows(sa + A_HALO_BYTES, tap);
                    const uint64_t b_hi = umma_desc_k_sw128(sb), b_lo = b_hi + (TILE_BYTES >> 4);
                    if (elect_one()) {
                        if (p.trace && blockIdx.x == 0 && gk < 256) p.trace[1 * 256 + gk] = clock64();
#pragma unroll
                        for (int kk = 0; kk < ((p.debug & 1) ? 0 : 4); ++kk) {
                            umma<PREC>(tmem_d, a_hi + 2 * kk, b_hi + 2 * kk, idesc, started | (uint32_t)kk);
                            if (p.passes >= 2) umma<PREC>(tmem_d, a_lo + 2 * kk, b_hi + 2 * kk, idesc, 1u);
                            if (p.passes >= 3) umma<PREC>(tmem_d, a_hi + 2 * kk, b_lo + 2 * kk, idesc, 1u);
                        }
                        if (p.debug & 1) mbar_arrive(empty0 + 8 * bs);
                        else umma_commit(empty0 + 8 * bs);
                        if (tap == 2) {
                            if (p.debug & 1) mbar_arrive(aempty0 + 8 * as);
                            else umma_commit(aempty0 + 8 * as);      // the halo tile is free once all three taps have retired
                            if (kc == p.k_chunks - 1) {
                                if (p.debug & 1) mbar_arrive(tfull0 + 8 * acc);
                                else umma_commit(tfull0 + 8 * acc);
                            }
                        }
                    }
                    __syncwarp();
                    ++gk;
                    started = 1;
                    if (++bs == HALO_B_STAGES) { bs = 0; bph ^= 1; }
                }
                if (++as == HALO_A_STAGES) { as = 0; aph ^= 1; }
            }
        }
    } else if (warp == 0) {
        if (lane == 0) {
            int stage = 0, phase = 0, gk = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const TileCoord t = tile_coord<MODE>(p, tile);
                if (tile_dead<MODE>(p, t)) continue;
                for (int k = t.k_begin; k < t.k_end; ++k) {
                    if (k_dead<MODE>(p, k)) continue;
                    {
                        const int o = k / p.k_chunks, kc = k - o * p.k_chunks;
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        const uint32_t bar = full0 + 8 * stage;
                        const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + PLANES * TILE_BYTES;
                        if (p.debug & 2) { mbar_arrive(bar); if (++stage == NSTAGE) { stage = 0; phase ^= 1; } continue; }
                        mbar_expect_tx(bar, PLANES == 2 ? (uint32_t)(2 + (p.passes >= 2) + (p.passes >= 3)) * TILE_BYTES : (uint32_t)STAGE_BYTES);
                        int a0, a1, a2, b0, b1, b2;
                        if constexpr (MODE != MODE_WGRAD) {      // o = tap
                            a0 = kc * TKB; a1 = t.a1 + o - pad; a2 = t.a2;
                            b0 = kc * TKB; b1 = t.b1; b2 = o;
                        } else {                                // o = utterance; K = 64 rows of s; operands are (c, s, b) boxes
                            a0 = t.a2; a1 = kc * TKB; a2 = o;
                            b0 = t.b2; b1 = kc * TKB + t.tap - pad; b2 = o;
                        }
                        if constexpr (MODE != MODE_WGRAD) {
                            tma_load_3d(sa, &map_a_hi, bar, a0, a1, a2);
                            tma_load_3d(sb, &map_b_hi, bar, b0, b1, b2);
                            if constexpr (PLANES == 2) {
                                if (p.passes >= 2) tma_load_3d(sa + TILE_BYTES, &map_a_lo, bar, a0, a1, a2);
                                if (p.passes >= 3) tma_load_3d(sb + TILE_BYTES, &map_b_lo, bar, b0, b1, b2);
                            }
                        } else {                                // two 64-column MN chunks per 128-wide tile
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                tma_load_3d(sa + j * 8192, &map_a_hi, bar, a0 + 64 * j, a1, a2);
                                tma_load_3d(sb + j * 8192, &map_b_hi, bar, b0 + 64 * j, b1, b2);
                                if (p.passes >= 2) tma_load_3d(sa + TILE_BYTES + j * 8192, &map_a_lo, bar, a0 + 64 * j, a1, a2);
                                if (p.passes >= 3) tma_load_3d(sb + TILE_BYTES + j * 8192, &map_b_lo, bar, b0 + 64 * j, b1, b2);
                            }
                        }
                        if (p.trace && blockIdx.x == 0 && gk < 256) p.trace[0 * 256 + gk] = clock64();
                        ++gk;
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // whole warp in the loop, one elected lane issues (see MODE_HALO above)
        constexpr uint32_t idesc = umma_idesc(PREC == PREC_TF32 ? 2 : 1, TM, TN, MODE == MODE_WGRAD ? 1 : 0);
        // descriptor step per K=16 (bf16) / K=8 (tf32) instruction, in 16-byte units:
        //   K-major: 32 bytes along the 128-byte row;  MN-major: 16 rows of 128 bytes
        constexpr uint32_t kstep = MODE == MODE_WGRAD ? (16 * 128) >> 4 : 2;
        int stage = 0, phase = 0, it = 0, gk = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const TileCoord t = tile_coord<MODE>(p, tile);
            if (tile_dead<MODE>(p, t)) continue;             // (the accumulator sequence `it` only counts live tiles)
            const int acc = it & 1, acc_phase = (it >> 1) & 1;
            ++it;
            mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);     // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * TN;
            uint32_t started = 0;
            for (int k = t.k_begin; k < t.k_end; ++k) {
                if (k_dead<MODE>(p, k)) continue;
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + PLANES * TILE_BYTES;
                const uint64_t a_hi = MODE == MODE_WGRAD ? umma_desc_mn_sw128(sa) : umma_desc_k_sw128(sa);
                const uint64_t b_hi = MODE == MODE_WGRAD ? umma_desc_mn_sw128(sb) : umma_desc_k_sw128(sb);
                if (elect_one()) {
                    if (p.trace && blockIdx.x == 0 && gk < 256) p.trace[1 * 256 + gk] = clock64();
#pragma unroll
                    for (int kk = 0; kk < ((p.debug & 1) ? 0 : 4); ++kk) {
                        umma<PREC>(tmem_d, a_hi + kstep * kk, b_hi + kstep * kk, idesc, started | (uint32_t)kk);
                        if constexpr (PLANES == 2) {
                            const uint64_t a_lo = a_hi + (TILE_BYTES >> 4), b_lo = b_hi + (TILE_BYTES >> 4);
                            if (p.passes >= 2) umma<PREC>(tmem_d, a_lo + kstep * kk, b_hi + kstep * kk, idesc, 1u);
                            if (p.passes >= 3) umma<PREC>(tmem_d, a_hi + kstep * kk, b_lo + kstep * kk, idesc, 1u);
                        }
                    }
                    if (p.debug & 1) mbar_arrive(empty0 + 8 * stage);
                    else umma_commit(empty0 + 8 * stage);     // frees the smem stage when these MMAs retire
                }
                __syncwarp();
                ++gk;
                started = 1;
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) {
                if (p.debug & 1) mbar_arrive(tfull0 + 8 * acc);
                else umma_commit(tfull0 + 8 * acc);           // accumulator complete -> epilogue
            }
            __syncwarp();
        }
    } else {
        // epilogue: NUM_EPI_WARPS = 16 warps, FOUR per SM sub-partition.  Measured (round 2, tools/tc_trace.py + DX_TC_DEBUG masks): the
        // epilogue of one 128 x 128 tile took 3-5 k cycles with two warps per scheduler — dependent-issue latency of ~1 k instructions
        // per warp, not TMEM loads, bias loads or the TMA stores — and bounded every GEMM of the path (1 tensor-core pass instead of 3
        // bought 12 %).  Warp w may only touch TMEM lanes [32*(w%4), +32) (thread <-> output row); the four warps of a lane quarter
        // take one 32-column chunk each.  Staging: 2 KB per warp, used in two rounds (hi plane then lo plane; or columns 0-15 then
        // 16-31 of an fp32 chunk), each a 32-row x 64-byte SWIZZLE_64B box stored by TMA.
        const int quad = warp & 3;
        const int part = (warp - 2) >> 2;                       // 32-column chunk of the tile
        const int row = quad * 32 + lane;
        const uint32_t wbuf = sOut + (uint32_t)(warp - 2) * 2048u;
        const uint32_t rbuf = wbuf + lane * 64;
        // one round: 32 rows x 64 bytes (16 bytes per unit u) into the SWIZZLE_64B staging tile, then one TMA store
        auto stage64 = [&](const uint32_t* w16) {   // w16: 16 x 32-bit words of this lane's row
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t off = (uint32_t)((c ^ ((lane >> 1) & 3)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbuf + off), "r"(w16[4 * c]), "r"(w16[4 * c + 1]), "r"(w16[4 * c + 2]),
                             "r"(w16[4 * c + 3]) : "memory");
            }
            if (!(p.debug & 64)) fence_async_smem();
            __syncwarp();
        };
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const TileCoord t = tile_coord<MODE>(p, tile);
            const bool dead = tile_dead<MODE>(p, t);             // dead tiles are stored as zeros, no accumulator involved
            const int acc = it & 1, acc_phase = (it >> 1) & 1;
            uint32_t v[32];
            if (!dead) {
                mbar_wait(tfull0 + 8 * acc, acc_phase);
                if (p.trace && blockIdx.x == 0 && threadIdx.x == 64 && it < 256) p.trace[2 * 256 + it] = clock64();
                tc_fence_after();
            }
            if (dead || (p.debug & 4)) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0u;
            } else {
                tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TN + part * 32, v);
                tmem_ld_wait(v);
            }
            if (!dead) {   // this warp's TMEM reads are done: hand the accumulator back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            }
            const int s = t.o1 + row;
            const bool row_ok = MODE != MODE_WGRAD && s < p.S;
            const size_t grow = (size_t)t.o2 * p.S + s;
            const int nb = t.o0 + part * 32;
            const int row0 = t.o1 + quad * 32;
            if constexpr (EPI == EPI_LN) {
                // ---- LayerNorm epilogue: this thread owns columns [part * 32, +32) of output row `row` (Cout == 128) --------------
                const bool valid = row_ok && (p.lens == nullptr || s < (int)p.lens[t.o2]);
                float mean = 0.f, rstd = 0.f;
                if (!dead) {
                    const unsigned long long seed_in = dyn_seed(p.ln_seed_in, p.dyn);
                    const float inv_keep = p.ln_p_in > 0.f ? 1.f / (1.f - p.ln_p_in) : 1.f;
                    float psum = 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + c4);
                        float a[4] = {__uint_as_float(v[4 * c4]) * p.alpha + bv.x, __uint_as_float(v[4 * c4 + 1]) * p.alpha + bv.y,
                                      __uint_as_float(v[4 * c4 + 2]) * p.alpha + bv.z, __uint_as_float(v[4 * c4 + 3]) * p.alpha + bv.w};
                        if (p.ln_p_in > 0.f) {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                a[e] *= dropout_scale(seed_in, (unsigned long long)grow * 128ull + (unsigned)(nb + 4 * c4 + e), p.ln_p_in, inv_keep);
                        }
                        if (p.ln_res && row_ok) {
                            const float4 r = __ldg(reinterpret_cast<const float4*>(p.ln_res + grow * 128 + nb) + c4);
                            a[0] += r.x; a[1] += r.y; a[2] += r.z; a[3] += r.w;
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) { v[4 * c4 + e] = __float_as_uint(a[e]); psum += a[e]; }
                    }
                    // exchange the partial row sums among the four warps that own the four column chunks of the same rows
                    const uint32_t xrow = xch + (uint32_t)row * 16u;          // [slot][row][part] fp32
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(xrow + 4u * part), "f"(psum) : "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");
                    float4 q4;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q4.x), "=f"(q4.y), "=f"(q4.z), "=f"(q4.w) : "r"(xrow) : "memory");
                    mean = ((q4.x + q4.y) + (q4.z + q4.w)) * (1.f / 128.f);
                    float psq = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const float d = __uint_as_float(v[i]) - mean; psq += d * d; }
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(xrow + 2048u + 4u * part), "f"(psq) : "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory");
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q4.x), "=f"(q4.y), "=f"(q4.z), "=f"(q4.w) : "r"(xrow + 2048u) : "memory");
                    rstd = rsqrtf(((q4.x + q4.y) + (q4.z + q4.w)) * (1.f / 128.f) + 1e-5f);
                }
                if (part == 0 && row_ok) p.ln_rstd[grow] = (valid && !dead) ? rstd : 0.f;
                const float keep = (valid && !dead) ? 1.f : 0.f;
                float o[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = (__uint_as_float(v[i]) - mean) * rstd * keep;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {   // xhat: columns [nb + 16 hh, +16) -> map_y2
                    stage64(reinterpret_cast<const uint32_t*>(o) + 16 * hh);
                    if (lane == 0) { tma_store_3d(&map_y2, wbuf, nb + 16 * hh, row0, t.o2); tma_store_commit(); }
                }
                // y = mask(film_g * (xhat * w + b) + film_b)
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.ln_w + nb) + c4);
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ln_b + nb) + c4);
                    o[4 * c4] = o[4 * c4] * w4.x + b4.x; o[4 * c4 + 1] = o[4 * c4 + 1] * w4.y + b4.y;
                    o[4 * c4 + 2] = o[4 * c4 + 2] * w4.z + b4.z; o[4 * c4 + 3] = o[4 * c4 + 3] * w4.w + b4.w;
                }
                if (p.film) {
                    const float* fg = p.film + (size_t)t.o2 * p.film_stride + nb;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 g4 = __ldg(reinterpret_cast<const float4*>(fg) + c4);
                        const float4 e4 = __ldg(reinterpret_cast<const float4*>(fg + 128) + c4);
                        o[4 * c4] = g4.x * o[4 * c4] + e4.x; o[4 * c4 + 1] = g4.y * o[4 * c4 + 1] + e4.y;
                        o[4 * c4 + 2] = g4.z * o[4 * c4 + 2] + e4.z; o[4 * c4 + 3] = g4.w * o[4 * c4 + 3] + e4.w;
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] *= keep;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {   // y -> map_y
                    stage64(reinterpret_cast<const uint32_t*>(o) + 16 * hh);
                    if (lane == 0) { tma_store_3d(&map_y, wbuf, nb + 16 * hh, row0, t.o2); tma_store_commit(); }
                }
                if (p.ln_planes) {
                    uint32_t hl[32];                 // 16 hi words | 16 lo words
#pragma unroll
                    for (int i = 0; i < 16; ++i) split_pair_u32(o[2 * i], o[2 * i + 1], hl[i], hl[16 + i]);
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl) {
                        stage64(hl + 16 * pl);
                        if (lane == 0) { tma_store_3d(&map_y3, wbuf, nb, row0, pl * p.B + t.o2); tma_store_commit(); }
                    }
                }
                if (!dead) ++it;
                continue;
            }
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(v[i]);
            if (MODE != MODE_WGRAD && !dead) {
                // every condition below is warp-uniform except row_ok; columns >= Cout are clipped by the TMA store
                const bool full = nb + 32 <= p.Cout;
                if (p.alpha != 1.f) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] *= p.alpha;
                }
                if (p.bias && !(p.debug & 256)) {
                    if (full) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + c);
                            o[4 * c] += bv.x; o[4 * c + 1] += bv.y; o[4 * c + 2] += bv.z; o[4 * c + 3] += bv.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (nb + i < p.Cout) o[i] += __ldg(p.bias + nb + i);
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = fmaxf(o[i], 0.f);
                }
                if (p.relu_src && row_ok) {
                    const float* src = p.relu_src + grow * p.Cout + nb;
                    if (full) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 m = __ldg(reinterpret_cast<const float4*>(src) + c);
                            o[4 * c] = m.x > 0.f ? o[4 * c] : 0.f; o[4 * c + 1] = m.y > 0.f ? o[4 * c + 1] : 0.f;
                            o[4 * c + 2] = m.z > 0.f ? o[4 * c + 2] : 0.f; o[4 * c + 3] = m.w > 0.f ? o[4 * c + 3] : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (nb + i < p.Cout) o[i] = __ldg(src + i) > 0.f ? o[i] : 0.f;
                    }
                }
                if (p.relu_src_hi && row_ok) {   // mask from the bf16 hi plane of the forward activation: hi > 0 <=> fp32 value > 0
                    const uint4* src = reinterpret_cast<const uint4*>(p.relu_src_hi + grow * p.Cout + nb);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 m = __ldg(src + c);
                        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            o[8 * c + 2 * k] = (short)(mw[k] & 0xffffu) > 0 ? o[8 * c + 2 * k] : 0.f;
                            o[8 * c + 2 * k + 1] = ((int)mw[k] >> 16) > 0 ? o[8 * c + 2 * k + 1] : 0.f;
                        }
                    }
                }
                if (p.add_src && row_ok) {
                    const float* src = p.add_src + grow * p.ldy + nb;
                    if (full) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 a4 = __ldg(reinterpret_cast<const float4*>(src) + c);
                            o[4 * c] += a4.x; o[4 * c + 1] += a4.y; o[4 * c + 2] += a4.z; o[4 * c + 3] += a4.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (nb + i < p.Cout) o[i] += __ldg(src + i);
                    }
                }
                if (p.round_tf32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = round_tf32(o[i]);
                }
                if (p.hp_dot_src && row_ok) {
                    const float4* src = reinterpret_cast<const float4*>(p.hp_dot_src + grow * p.Cout + nb);
                    float d0 = 0.f, d1 = 0.f;                         // columns [0, 16) and [16, 32) of the chunk
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 a4 = __ldg(src + c);
                        const float t = o[4 * c] * a4.x + o[4 * c + 1] * a4.y + o[4 * c + 2] * a4.z + o[4 * c + 3] * a4.w;
                        if (c < 4) d0 += t; else d1 += t;
                    }
                    if (p.hp_dh == 16) {                              // two heads per chunk: each entry has exactly one writer
                        float* dst = p.hp_dot_out + ((size_t)t.o2 * p.hp_NH + (nb >> 4)) * p.S + s;
                        dst[0] = d0;
                        dst[p.S] = d1;
                    } else {                                          // dh / 32 chunks per head: order-independent sum of <= 2 addends
                        atomicAdd(p.hp_dot_out + ((size_t)t.o2 * p.hp_NH + nb / p.hp_dh) * p.S + s, d0 + d1);
                    }
                }
                if (p.hp_dh) {
                    if (!row_ok) {                                   // rows in [S, Sp): +0 (like the conversion pass wrote)
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = 0.f;
                    } else if (nb < p.hp_scale_cols) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] *= p.hp_scale;
                    }
                }
            }
            if (MODE != MODE_WGRAD && p.colsum && !dead) {
                // column sums of this warp's 32 rows x 32 columns by recursive halving across the lanes (31 shuffles):
                // afterwards lane l holds the sum of column l
                const float rk = row_ok ? 1.f : 0.f;
                float r16[16], r8[8], r4[4], r2[2];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const bool up = lane & 16;
                    const float keep = (up ? o[i + 16] : o[i]) * rk, send = (up ? o[i] : o[i + 16]) * rk;
                    r16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool up = lane & 8;
                    r8[i] = (up ? r16[i + 8] : r16[i]) + __shfl_xor_sync(0xffffffffu, up ? r16[i] : r16[i + 8], 8);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool up = lane & 4;
                    r4[i] = (up ? r8[i + 4] : r8[i]) + __shfl_xor_sync(0xffffffffu, up ? r8[i] : r8[i + 4], 4);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const bool up = lane & 2;
                    r2[i] = (up ? r4[i + 2] : r4[i]) + __shfl_xor_sync(0xffffffffu, up ? r4[i] : r4[i + 2], 2);
                }
                const bool up1 = lane & 1;
                const float tot = (up1 ? r2[1] : r2[0]) + __shfl_xor_sync(0xffffffffu, up1 ? r2[0] : r2[1], 1);
                if (nb + lane < p.Cout) atomicAdd(p.colsum + nb + lane, tot);
            }
            if (MODE != MODE_WGRAD && p.y_planes) {
                // the consumer GEMMs read bf16 hi|lo operand planes: emit them here instead of a later split pass over an
                // fp32 copy (32 columns = 64 bytes per row and plane)
                uint32_t hl[32];                     // 16 hi words | 16 lo words
#pragma unroll
                for (int i = 0; i < 16; ++i) split_pair_u32(o[2 * i], o[2 * i + 1], hl[i], hl[16 + i]);
                if (p.skip_y && p.hp_dh == 16) {
                    // head planes, head_dim 16: the chunk covers two heads; per plane two dense [32 rows][32 B] tiles (the map over R
                    // has a 16-column box and no swizzle)
                    const int hh = nb >> 4, pls = p.B * p.hp_NH;
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl) {
                        if (lane == 0) tma_store_wait_read<0>();
                        __syncwarp();
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
                            const uint32_t rb = wbuf + sub * 1024 + lane * 32;
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rb + 16 * c), "r"(hl[16 * pl + 8 * sub + 4 * c]),
                                             "r"(hl[16 * pl + 8 * sub + 4 * c + 1]), "r"(hl[16 * pl + 8 * sub + 4 * c + 2]),
                                             "r"(hl[16 * pl + 8 * sub + 4 * c + 3]) : "memory");
                        }
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_3d(&map_y, wbuf, 0, row0, pl * pls + t.o2 * p.hp_NH + hh);
                            tma_store_3d(&map_y, wbuf + 1024, 0, row0, pl * pls + t.o2 * p.hp_NH + hh + 1);
                            tma_store_commit();
                        }
                    }
                    if (!dead) ++it;
                    continue;
                }
                if (p.skip_y) {
                    // planes only: hi tile then lo tile (2 KB each) through this warp's buffer (map_y is the bf16 planes map; rows >= S
                    // are clipped by the store)
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl) {
                        stage64(hl + 16 * pl);
                        if (lane == 0 && !(p.debug & 8)) {
                            if (p.hp_dh) {   // head planes, head_dim 32 / 64: one head per chunk (column d0 inside the head)
                                const int hh = nb / p.hp_dh, d0 = nb - hh * p.hp_dh;
                                tma_store_3d(&map_y, wbuf, d0, row0, (pl * p.B + t.o2) * p.hp_NH + hh);
                            } else {
                                tma_store_3d(&map_y, wbuf, nb, row0, pl * p.B + t.o2);
                            }
                            tma_store_commit();
                        }
                    }
                    if (!dead) ++it;
                    continue;
                }
                if (row_ok) {   // fp32 output AND planes requested: direct (slower) global stores for the planes
                    uint4* dh = reinterpret_cast<uint4*>(p.y_planes + grow * p.Cout + nb);
                    uint4* dl = reinterpret_cast<uint4*>(p.y_planes + p.y_plane_elems + grow * p.Cout + nb);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        dh[c] = make_uint4(hl[4 * c], hl[4 * c + 1], hl[4 * c + 2], hl[4 * c + 3]);
                        dl[c] = make_uint4(hl[16 + 4 * c], hl[16 + 4 * c + 1], hl[16 + 4 * c + 2], hl[16 + 4 * c + 3]);
                    }
                }
            }
            // fp32 output: two rounds of 32 rows x 16 columns (64-byte rows) through this warp's 2 KB buffer
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                stage64(reinterpret_cast<const uint32_t*>(o) + 16 * hh);
                if (lane == 0 && !(p.debug & 8)) {
                    tma_store_3d(&map_y, wbuf, nb + 16 * hh, row0, t.o2);
                    tma_store_commit();
                }
            }
            if (p.trace && blockIdx.x == 0 && threadIdx.x == 64 && it < 256) p.trace[3 * 256 + it] = clock64();
            if (!dead) ++it;
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1 && !(p.debug & 16)) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- weight gradient, KW = 3, ONE bf16 pass: all three taps per CTA from one x tile fetched WITH ITS HALO -----------------
// dW[co][ci][tap] = sum_s dy[s][co] * x[s + tap - 1][ci].  The generic MODE_WGRAD kernel makes one 128 x 128 output tile per
// (tap, K split) and streams a dy tile and an x tile for each: per 64-row chunk of the reduction 6 tile loads serve 3 taps, and at one
// tensor-core pass the 128 x 128 x 64 MMAs (~260 cycles) finish long before L2 has delivered the next 32 KB (measured: the K = 32 k
// row weight gradients run at 0.46 - 0.65 PFLOP/s whatever the pass count buys).  Here a CTA owns (Cout tile, Cin tile, K split) for
// ALL taps: per chunk it loads the dy tile once and the x tile once with one halo row on each side (66 rows; rows outside the
// utterance arrive as zeros), and the three taps read it through MN-major UMMA descriptors whose start address is advanced by
// `tap` rows (the 128-byte swizzle is a function of the absolute shared-memory address, see umma_desc_k_sw128_rows), into three
// TMEM accumulators.  2 tile loads per chunk instead of 6, 34 KB stages (five of them), 12 MMAs per stage.
constexpr int WG3_STAGES = 5;
constexpr int WG3_A_BYTES = 2 * 8192;                            // dy: two 64-channel MN chunks of [64 rows][128 B]
constexpr int WG3_B_CHUNK = 9 * 1024;                            // x: 66 rows x 128 B, rounded up to the 1024-byte swizzle repeat
constexpr int WG3_STAGE_BYTES = WG3_A_BYTES + 2 * WG3_B_CHUNK;   // 34816
constexpr int WG3_SMEM = WG3_STAGES * WG3_STAGE_BYTES + 2 * OUT_BYTES + 256 + 1024;
// The same kernel serves k = 1 (KW = 1: a 64-row x box, one accumulator): five 32 KB stages in flight and K splits over live chunks
// instead of the generic kernel's three 64 KB stages, half of each unused at one pass.
constexpr int WG3_TMEM_COLS = 512;                               // three 128-column accumulators

struct Wg3Params {
    const long long* lens;
    int B, S, Cin, Cout, halo, tiles_m, tiles_n, nsplit, k_chunks, num_tiles;
};
// The K splits divide the LIVE 64-row chunks evenly (rows at or beyond len + halo carry dy == 0 exactly, see live_rows above, and are
// never loaded): with the bench batch (lengths 513..1000 of 1000) equal shares of the (utterance, chunk) grid differ by up to 2x in work.
struct Wg3Tile { int co0, ci0, split, n_begin, n_end; };        // [n_begin, n_end) in live-chunk order

__device__ __forceinline__ int wg3_live_chunks(const Wg3Params& p, int b) {
    if (p.lens == nullptr) return p.k_chunks;
    return min(p.k_chunks, (min((int)p.lens[b], p.S) + p.halo + 63) >> 6);
}
__device__ __forceinline__ Wg3Tile wg3_tile(const Wg3Params& p, int tile, int total_live) {
    Wg3Tile t;
    const int tci = tile % p.tiles_n;
    const int r = tile / p.tiles_n;
    t.ci0 = tci * TN; t.co0 = (r % p.tiles_m) * TM; t.split = r / p.tiles_m;
    t.n_begin = (int)(((long long)t.split * total_live) / p.nsplit);
    t.n_end = (int)(((long long)(t.split + 1) * total_live) / p.nsplit);
    return t;
}
// walks the live chunks of a tile: (b, kc) of live chunk n_begin, then next()
struct Wg3Walk {
    int b, kc, lc;
    __device__ __forceinline__ void start(const Wg3Params& p, int n) {
        b = 0; lc = wg3_live_chunks(p, 0);
        while (n >= lc) { n -= lc; ++b; lc = wg3_live_chunks(p, b); }
        kc = n;
    }
    __device__ __forceinline__ void next(const Wg3Params& p) {
        if (++kc == lc) { kc = 0; ++b; if (b < p.B) lc = wg3_live_chunks(p, b); }
    }
};
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_lbo(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int KW>
__global__ void __launch_bounds__(NTHREADS, 1) wgrad_halo3_kernel(const __grid_constant__ CUtensorMap map_dy,     // (co, s, b), box 64 x 64
                                                                  const __grid_constant__ CUtensorMap map_x,      // (ci, s, b), box 64 x 66
                                                                  const __grid_constant__ CUtensorMap map_part,   // (ci, co, split * 3 + tap) fp32
                                                                  Wg3Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sOut = base + WG3_STAGES * WG3_STAGE_BYTES;
    const uint32_t bars = sOut + 2 * OUT_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * WG3_STAGES, tfull = bars + 16 * WG3_STAGES, tempty = tfull + 8, tmem_slot = tempty + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    volatile int* total_slot = reinterpret_cast<volatile int*>(smem_raw + (tmem_slot + 8 - smem_u32(smem_raw)));

    if (threadIdx.x == 0) {
        for (int i = 0; i < WG3_STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        mbar_init(tfull, 1);
        mbar_init(tempty, NUM_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int tot = 0;
        for (int b = 0; b < p.B; ++b) tot += wg3_live_chunks(p, b);
        *total_slot = tot;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(WG3_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int total_live = *total_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const Wg3Tile t = wg3_tile(p, tile, total_live);
                if (t.n_begin == t.n_end) continue;
                Wg3Walk w;
                w.start(p, t.n_begin);
                for (int n = t.n_begin; n < t.n_end; ++n, w.next(p)) {
                    const int b = w.b, s0 = w.kc * 64;
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t bar = full0 + 8 * stage;
                    const uint32_t sa = base + stage * WG3_STAGE_BYTES, sb = sa + WG3_A_BYTES;
                    mbar_expect_tx(bar, WG3_A_BYTES + 2 * (64 + KW - 1) * 128);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        tma_load_3d(sa + j * 8192, &map_dy, bar, t.co0 + 64 * j, s0, b);
                        tma_load_3d(sb + j * WG3_B_CHUNK, &map_x, bar, t.ci0 + 64 * j, s0 - (KW - 1) / 2, b);
                    }
                    if (++stage == WG3_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc(1, TM, TN, 1);
        constexpr uint32_t kstep = (16 * 128) >> 4;              // 16 reduction rows of 128 bytes per K = 16 instruction
        int stage = 0, phase = 0, it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const Wg3Tile t = wg3_tile(p, tile, total_live);
            if (t.n_begin == t.n_end) continue;
            mbar_wait(tempty, (it & 1) ^ 1);                      // the epilogue has read the three accumulators of the previous live tile
            ++it;
            tc_fence_after();
            uint32_t started = 0;
            for (int n = t.n_begin; n < t.n_end; ++n) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const uint32_t sa = base + stage * WG3_STAGE_BYTES, sb = sa + WG3_A_BYTES;
                const uint64_t a = umma_desc_mn_sw128(sa);
                if (elect_one()) {
#pragma unroll
                    for (uint32_t tap = 0; tap < (uint32_t)KW; ++tap) {
                        const uint64_t bd = umma_desc_mn_sw128_lbo(sb + tap * 128u, WG3_B_CHUNK);
#pragma unroll
                        for (uint32_t kk = 0; kk < 4; ++kk)
                            umma<PREC_BF16X3>(tmem_base + tap * TN, a + kstep * kk, bd + kstep * kk, idesc, started | kk);
                    }
                    umma_commit(empty0 + 8 * stage);
                }
                __syncwarp();
                started = 1;
                if (++stage == WG3_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(tfull);
            __syncwarp();
        }
    } else {
        const int quad = warp & 3, part = (warp - 2) >> 2;
        const uint32_t wbuf = sOut + (uint32_t)(warp - 2) * 2048u;
        const uint32_t rbuf = wbuf + lane * 64;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const Wg3Tile t = wg3_tile(p, tile, total_live);
            const bool dead = t.n_begin == t.n_end;              // a split without live chunks stores zeros
            if (!dead) {
                mbar_wait(tfull, it & 1);
                tc_fence_after();
            }
#pragma unroll 1
            for (int tap = 0; tap < KW; ++tap) {
                uint32_t v[32];
                if (dead) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                } else {
                    tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + tap * TN + part * 32, v);
                    tmem_ld_wait(v);
                    if (tap == KW - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty);
                    }
                }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {                  // 32 rows x 16 fp32 columns per round through the SWIZZLE_64B staging tile
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t off = (uint32_t)((c ^ ((lane >> 1) & 3)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbuf + off), "r"(v[16 * hh + 4 * c]), "r"(v[16 * hh + 4 * c + 1]),
                                     "r"(v[16 * hh + 4 * c + 2]), "r"(v[16 * hh + 4 * c + 3]) : "memory");
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&map_part, wbuf, t.ci0 + part * 32 + 16 * hh, t.co0 + quad * 32, t.split * KW + tap);
                        tma_store_commit();
                    }
                }
            }
            if (!dead) ++it;
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(WG3_TMEM_COLS) : "memory");
    }
}

// ---- operand planes: fp32 [B, S, ld] -> bf16 hi/lo, row-major [B*S][C] and/or transposed [C][B][Sp] -------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// grid (ceil(S/64), ceil(C/64), B), 256 threads; one 64(s) x 64(c) tile per block
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, int ld, int S, int C, int Sp,
                                                           __nv_bfloat16* __restrict__ row_hi, __nv_bfloat16* __restrict__ row_lo,
                                                           __nv_bfloat16* __restrict__ t_hi, __nv_bfloat16* __restrict__ t_lo, int B,
                                                           float* __restrict__ colsum = nullptr) {
    __shared__ __nv_bfloat16 th[64][72], tl[64][72];   // [c][s], padded rows (144 B: 16-byte aligned chunks)
    __shared__ float csum[16][68];
    const int b = blockIdx.z, s0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int t = threadIdx.x;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = t + 256 * i, r = idx >> 4, cq = (idx & 15) * 4;
        const int s = s0 + r, c = c0 + cq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < S && c < C) v = *reinterpret_cast<const float4*>(src + ((size_t)b * S + s) * ld + c);   // C % 4 == 0
        cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
        __nv_bfloat16 h[4], l[4];
        split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
        if (row_hi && s < S && c < C) {
            const size_t o = ((size_t)b * S + s) * C + c;
            *reinterpret_cast<uint2*>(row_hi + o) = *reinterpret_cast<uint2*>(h);
            *reinterpret_cast<uint2*>(row_lo + o) = *reinterpret_cast<uint2*>(l);
        }
        if (t_hi) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { th[cq + e][r] = h[e]; tl[cq + e][r] = l[e]; }
        }
    }
    if (colsum) {   // thread (t >> 4) covers rows (t >> 4) + 16 i of column quad (t & 15): reduce the 16 row groups in smem
        *reinterpret_cast<float4*>(&csum[t >> 4][(t & 15) * 4]) = cs;
        __syncthreads();
        if (t < 64 && c0 + t < C) {
            float tot = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) tot += csum[k][t];
            atomicAdd(colsum + c0 + t, tot);
        }
    }
    if (!t_hi) return;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = t + 256 * i, c = idx >> 3, sq = (idx & 7) * 8;   // 64 c-rows x 8 chunks of 8 bf16 (16 bytes)
        if (c0 + c < C && s0 + sq < Sp) {
            const size_t o = ((size_t)(c0 + c) * B + b) * Sp + s0 + sq;
            *reinterpret_cast<uint4*>(t_hi + o) = *reinterpret_cast<uint4*>(&th[c][sq]);
            *reinterpret_cast<uint4*>(t_lo + o) = *reinterpret_cast<uint4*>(&tl[c][sq]);
        }
    }
}

// weight planes: fp32 packed [KW][N][K] -> bf16 hi/lo (same layout)
__global__ void split_flat_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                  size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        __nv_bfloat16 h, l;
        split_bf16(src[i], h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

// ---- host side: tensor maps -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 3-D tensor (d0 contiguous), box (b0, b1, b2), SWIZZLE_128B (or 64B), zero OOB fill.  esz = 4 (fp32) or 2 (bf16).
int make_map_3d(CUtensorMap* map, const void* ptr, int esz, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, bool swizzle64 = false, bool swizzle32 = false,
                bool swizzle_none = false) {
    EncodeTiledFn enc = get_encode();
    DX_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                     const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_none ? CU_TENSOR_MAP_SWIZZLE_NONE
                                  : (swizzle32 ? CU_TENSOR_MAP_SWIZZLE_32B : (swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B)),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) esz=%d dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u,%u)",
               (int)r, esz, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
               (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, b0, b1, b2);
    return DX_OK;
}

int g_num_sms = 0;
int g_prec = PREC_BF16X3;
unsigned long long g_tc_launches = 0;
int g_passes_conv = 3, g_passes_wgrad = 0;   // BF16X3 passes issued by forward / dgrad GEMMs and by weight-gradient GEMMs (wgrad 0 = by reduction length)
// wgrad_passes = 0: a weight gradient sums B*S products per element in fp32; with this many rows or more the rounding of the operands
// to bf16 (unbiased, independent from row to row) is taken as is: ONE pass on the hi planes, the lo planes are not loaded
constexpr long long kWgradSinglePassRows = 4096;

int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int PREC, int MODE, int EPI = EPI_STD>
int launch(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, const CUtensorMap& my,
           const TcParams& p, cudaStream_t st, const CUtensorMap* my2 = nullptr, const CUtensorMap* my3 = nullptr) {
    static bool configured = false;
    if (!configured) {
        DX_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<PREC, MODE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     smem_bytes<PREC, MODE, EPI>()));
        configured = true;
    }
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    gemm_tc_kernel<PREC, MODE, EPI><<<grid, NTHREADS, smem_bytes<PREC, MODE, EPI>(), st>>>(ah, al, bh, bl, my, my2 ? *my2 : my,
                                                                                           my3 ? *my3 : my, p);
    ++g_tc_launches;
    return check_launch("gemm_tc");
}

long long* g_trace = nullptr;

int tc_debug_mask() {
    static int mask = -1;
    if (mask < 0) {
        const char* e = getenv("DX_TC_DEBUG");
        mask = e ? atoi(e) : 0;
    }
    return mask;
}

bool halo_enabled() {   // DX_TC_HALO=0 falls back to per-tap activation loads (bring-up / A-B timing)
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("DX_TC_HALO");
        on = (e && atoi(e) == 0) ? 0 : 1;
    }
    return on == 1;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline int round8(int x) { return (x + 7) & ~7; }

}  // namespace

bool tma_available() { return get_encode() != nullptr; }
int make_tma_map_3d(void* map, const void* ptr, int esz, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                    unsigned long long stride1_bytes, unsigned long long stride2_bytes, unsigned b0, unsigned b1, unsigned b2,
                    int swizzle_bytes) {
    return make_map_3d((CUtensorMap*)map, ptr, esz, d0, d1, d2, stride1_bytes, stride2_bytes, b0, b1, b2, swizzle_bytes == 64,
                       swizzle_bytes == 32);
}

void set_tc_precision(int tf32) { g_prec = tf32 ? PREC_TF32 : PREC_BF16X3; }
void set_tc_trace(long long* buf) { g_trace = buf; }
void set_tc_passes(int conv, int wgrad) { g_passes_conv = conv; g_passes_wgrad = wgrad; }
void get_tc_passes(int* conv, int* wgrad) { *conv = g_passes_conv; *wgrad = g_passes_wgrad; }
unsigned long long tc_gemm_launches() { return g_tc_launches; }
long long* tc_trace_buffer() { return g_trace; }

int split_activation_planes(const float* x, int ld, void* planes, float* colsum_out, int rows, int C, cudaStream_t st) {
    DX_REQUIRE(C % 8 == 0 && ld % 4 == 0, "split_activation_planes: C=%d ld=%d", C, ld);
    __nv_bfloat16* hi = (__nv_bfloat16*)planes;
    if (colsum_out) DX_CUDA(cudaMemsetAsync(colsum_out, 0, (size_t)C * sizeof(float), st));
    dim3 grid(ceil_div(rows, 64), ceil_div(C, 64), 1);
    split_planes_kernel<<<grid, 256, 0, st>>>(x, ld, rows, C, 0, hi, hi + (size_t)rows * C, nullptr, nullptr, 1, colsum_out);
    return check_launch("split_planes");
}

int split_weight_planes(const float* w, void* planes, size_t n, cudaStream_t st) {
    __nv_bfloat16* hi = (__nv_bfloat16*)planes;
    split_flat_kernel<<<grid_1d(n), 256, 0, st>>>(w, hi, hi + n, n);
    return check_launch("split_flat");
}

bool conv_gemm_tc_supported(const ConvGemmArgs& a) {
    if (a.Cin % 8 != 0 || a.Cout % 4 != 0 || a.ldx % 4 != 0 || a.ldy % 4 != 0) return false;   // 16-byte global strides
    if (((uintptr_t)a.x | (uintptr_t)a.w | (uintptr_t)a.y) & 15) return false;
    if (a.Cin < 16 || a.Cout < 16) return false;
    return get_encode() != nullptr;
}

// workspace: bf16 hi/lo planes of x ([B*S][Cin]) and of the packed weight ([KW][Cout][Cin]) when running BF16X3
size_t conv_gemm_tc_workspace(const ConvGemmArgs& a) {
    if (g_prec == PREC_TF32) return 0;
    const size_t xe = (size_t)a.B * a.S * a.Cin, we = (size_t)a.KW * a.Cout * a.Cin;
    return (a.x_planes ? 0 : 2 * align256(xe * 2)) + (a.w_planes ? 0 : 2 * align256(we * 2)) + 256;
}

int conv_gemm_tc(const ConvGemmArgs& a, cudaStream_t st) {
    TcParams p;
    p.bias = a.bias; p.relu_src = a.relu_src; p.add_src = a.add_src;
    p.relu_src_hi = (const __nv_bfloat16*)a.relu_src_hi;
    p.y_planes = (__nv_bfloat16*)a.y_planes;
    p.colsum = a.y_colsum;
    if (a.y_colsum) DX_CUDA(cudaMemsetAsync(a.y_colsum, 0, (size_t)a.Cout * sizeof(float), st));
    p.y_plane_elems = (long long)a.B * a.S * a.Cout;
    p.skip_y = a.y == nullptr;
    p.hp_dh = 0; p.hp_NH = 0; p.hp_scale_cols = 0; p.hp_scale = 1.f; p.hp_dot_src = nullptr; p.hp_dot_out = nullptr;
    p.ln_res = nullptr; p.ln_w = nullptr; p.ln_b = nullptr; p.film = nullptr; p.film_stride = 0; p.ln_rstd = nullptr;
    p.ln_p_in = 0.f; p.ln_seed_in = 0; p.dyn = nullptr; p.ln_planes = 0;
    p.passes = g_passes_conv;
    if (a.ln) {
        const LnEpilogueArgs& l = *a.ln;
        DX_REQUIRE(g_prec == PREC_BF16X3 && a.Cout == 128 && a.y && l.xhat && l.rstd && l.ln_w && l.ln_b && a.bias && !a.relu && !a.relu_src &&
                       !a.relu_src_hi && !a.add_src && !a.y_colsum && !a.head_planes && a.ldy == 128,
                   "conv_gemm_tc: the LayerNorm epilogue needs the bf16x3 backend, Cout == 128, bias, y / xhat / rstd outputs (Cout = %d)", a.Cout);
        DX_REQUIRE((((uintptr_t)l.res | (uintptr_t)l.ln_w | (uintptr_t)l.ln_b | (uintptr_t)l.film | (uintptr_t)a.bias | (uintptr_t)l.xhat) & 15) == 0 &&
                       (l.film == nullptr || l.film_stride % 4 == 0),
                   "conv_gemm_tc: LayerNorm epilogue operands must be 16-byte aligned");
        p.ln_res = l.res; p.ln_w = l.ln_w; p.ln_b = l.ln_b; p.film = l.film; p.film_stride = l.film_stride; p.ln_rstd = l.rstd;
        p.ln_p_in = l.p_in; p.ln_seed_in = l.seed_in; p.dyn = l.dyn; p.ln_planes = a.y_planes != nullptr;
        p.skip_y = 0;
    }
    if (a.head_planes) {
        DX_REQUIRE(!a.y && !a.y_planes && !a.y_colsum && g_prec == PREC_BF16X3, "conv_gemm_tc: head_planes is a planes-only bf16x3 output");
        DX_REQUIRE((a.head_dim == 16 || a.head_dim == 32 || a.head_dim == 64) && a.Cout % a.head_dim == 0 && a.Cout % 32 == 0 &&
                       a.head_Sp >= a.S && a.head_Sp % 8 == 0 && ((uintptr_t)a.head_planes & 15) == 0,
                   "conv_gemm_tc: head_planes needs head_dim in {16, 32, 64}, Cout %% 32 == 0, Sp >= S (dh=%d Cout=%d Sp=%d S=%d)", a.head_dim,
                   a.Cout, a.head_Sp, a.S);
        p.y_planes = (__nv_bfloat16*)a.head_planes;
        p.skip_y = 1;
        p.hp_dh = a.head_dim; p.hp_NH = a.Cout / a.head_dim; p.hp_scale_cols = a.head_scale_cols; p.hp_scale = a.head_scale;
        if (a.head_dot_src) {
            DX_REQUIRE(a.head_dot_out && (((uintptr_t)a.head_dot_src | (uintptr_t)a.head_dot_out) & 15) == 0, "conv_gemm_tc: head_dot_src / head_dot_out");
            p.hp_dot_src = a.head_dot_src; p.hp_dot_out = a.head_dot_out;
            DX_CUDA(cudaMemsetAsync(a.head_dot_out, 0, (size_t)a.B * p.hp_NH * a.S * sizeof(float), st));
        }
    }
    DX_REQUIRE(a.y || a.y_planes || a.head_planes, "conv_gemm_tc: no output (y and y_planes are both NULL)");
    DX_REQUIRE(!(a.y_planes || a.relu_src_hi) || (a.Cout % 32 == 0 && (((uintptr_t)a.y_planes | (uintptr_t)a.relu_src_hi) & 15) == 0),
               "conv_gemm_tc: y_planes / relu_src_hi need Cout %% 32 == 0 and 16-byte alignment (Cout = %d)", a.Cout);
    p.B = a.B; p.S = a.S; p.Cin = a.Cin; p.Cout = a.Cout; p.KW = a.KW; p.ldy = a.ldy;
    p.tiles_m_per_b = ceil_div(a.S, TM);
    p.tiles_n = ceil_div(a.Cout, TN);
    p.num_tiles = p.tiles_m_per_b * a.B * p.tiles_n;
    p.nsplit = 1;
    p.lens = a.lens; p.halo = a.halo;
    p.alpha = a.alpha; p.relu = a.relu; p.round_tf32 = a.round_tf32;
    p.debug = tc_debug_mask();
    p.trace = g_trace;
    CUtensorMap mxh, mxl, mwh, mwl, my;
    int rc;
    if (a.head_planes) {
        // (dh, Sp, 2 * B * NH) over R[2][B][NH][Sp][dh]; 32-row boxes of min(32, dh) columns (64-byte rows: SWIZZLE_64B like the
        // plain planes map; 32-byte rows for head_dim 16: dense)
        const uint32_t bw = a.head_dim < 32 ? a.head_dim : 32;
        if ((rc = make_map_3d(&my, a.head_planes, 2, a.head_dim, a.head_Sp, 2ull * a.B * p.hp_NH, (uint64_t)a.head_dim * 2,
                              (uint64_t)a.head_Sp * a.head_dim * 2, bw, 32, 1, bw == 32, false, bw != 32)))
            return rc;
    } else if (a.y) {
        // fp32 output: 32-row x 16-column boxes (64-byte rows, SWIZZLE_64B), two per epilogue warp and tile
        if ((rc = make_map_3d(&my, a.y, 4, a.Cout, a.S, a.B, (uint64_t)a.ldy * 4, (uint64_t)a.S * a.ldy * 4, 16, 32, 1, true))) return rc;
    } else {
        // planes-only output: ONE bf16 map over [2][B*S][Cout] seen as (Cout, S, 2B): the lo plane is utterance index B + b;
        // 32 x 32 boxes of 64-byte rows, SWIZZLE_64B (the epilogue writes the staging tile with the matching xor)
        if ((rc = make_map_3d(&my, a.y_planes, 2, a.Cout, a.S, 2 * (uint64_t)a.B, (uint64_t)a.Cout * 2, (uint64_t)a.S * a.Cout * 2, 32, 32, 1, true)))
            return rc;
    }
    if (g_prec == PREC_TF32) {
        p.k_chunks = ceil_div(a.Cin, Cfg<PREC_TF32>::TKB);
        if ((rc = make_map_3d(&mxh, a.x, 4, a.Cin, a.S, a.B, (uint64_t)a.ldx * 4, (uint64_t)a.S * a.ldx * 4, 32, TM, 1))) return rc;
        if ((rc = make_map_3d(&mwh, a.w, 4, a.Cin, a.Cout, a.KW, (uint64_t)a.Cin * 4, (uint64_t)a.Cout * a.Cin * 4, 32, TN, 1))) return rc;
        return launch<PREC_TF32, MODE_CONV>(mxh, mxh, mwh, mwh, my, p, st);
    }
    const size_t need = conv_gemm_tc_workspace(a);
    DX_REQUIRE(need <= 256 || (a.workspace && a.workspace_bytes >= need), "conv_gemm_tc: workspace %zu < %zu bytes", a.workspace_bytes, need);
    const size_t xe = (size_t)a.B * a.S * a.Cin, we = (size_t)a.KW * a.Cout * a.Cin;
    uint8_t* ws = (uint8_t*)(((uintptr_t)a.workspace + 255) & ~(uintptr_t)255);
    const size_t xbytes = a.x_planes ? 0 : 2 * align256(xe * 2);
    const __nv_bfloat16* xh = (const __nv_bfloat16*)ws;
    const __nv_bfloat16* xl = (const __nv_bfloat16*)(ws + align256(xe * 2));
    const __nv_bfloat16* wh = (const __nv_bfloat16*)(ws + xbytes);
    const __nv_bfloat16* wl = (const __nv_bfloat16*)(ws + xbytes + align256(we * 2));
    {
        if (a.x_planes) {
            xh = (const __nv_bfloat16*)a.x_planes;
            xl = xh + xe;
        } else {
            dim3 grid(ceil_div(a.S, 64), ceil_div(a.Cin, 64), a.B);
            split_planes_kernel<<<grid, 256, 0, st>>>(a.x, a.ldx, a.S, a.Cin, 0, (__nv_bfloat16*)xh, (__nv_bfloat16*)xl, nullptr, nullptr, a.B);
            if ((rc = check_launch("split_planes"))) return rc;
        }
        if (a.w_planes) {   // cached by the caller: hi plane followed by lo plane
            wh = (const __nv_bfloat16*)a.w_planes;
            wl = wh + we;
        } else {
            split_flat_kernel<<<grid_1d(we), 256, 0, st>>>(a.w, (__nv_bfloat16*)wh, (__nv_bfloat16*)wl, we);
            if ((rc = check_launch("split_flat"))) return rc;
        }
    }
    if (((uintptr_t)xh | (uintptr_t)xl | (uintptr_t)wh | (uintptr_t)wl) & 15) {
        set_last_error("conv_gemm_tc: operand planes must be 16-byte aligned");
        return DX_ERR_ARG;
    }
    p.k_chunks = ceil_div(a.Cin, Cfg<PREC_BF16X3>::TKB);
    const uint64_t xs1 = (uint64_t)a.Cin * 2, xs2 = (uint64_t)a.S * a.Cin * 2;
    const uint64_t ws1 = (uint64_t)a.Cin * 2, ws2 = (uint64_t)a.Cout * a.Cin * 2;
    const bool halo_mode = a.KW == 3 && halo_enabled();
    const uint32_t box_rows = halo_mode ? HALO_ROWS : TM;
    if ((rc = make_map_3d(&mxh, xh, 2, a.Cin, a.S, a.B, xs1, xs2, 64, box_rows, 1))) return rc;
    if ((rc = make_map_3d(&mxl, xl, 2, a.Cin, a.S, a.B, xs1, xs2, 64, box_rows, 1))) return rc;
    if ((rc = make_map_3d(&mwh, wh, 2, a.Cin, a.Cout, a.KW, ws1, ws2, 64, TN, 1))) return rc;
    if ((rc = make_map_3d(&mwl, wl, 2, a.Cin, a.Cout, a.KW, ws1, ws2, 64, TN, 1))) return rc;
    if (a.ln) {
        CUtensorMap mxhat, mplanes;
        if ((rc = make_map_3d(&mxhat, a.ln->xhat, 4, a.Cout, a.S, a.B, (uint64_t)a.Cout * 4, (uint64_t)a.S * a.Cout * 4, 16, 32, 1, true))) return rc;
        if (a.y_planes) {
            if ((rc = make_map_3d(&mplanes, a.y_planes, 2, a.Cout, a.S, 2 * (uint64_t)a.B, (uint64_t)a.Cout * 2, (uint64_t)a.S * a.Cout * 2, 32, 32,
                                  1, true)))
                return rc;
        } else {
            mplanes = mxhat;
        }
        DX_REQUIRE(a.KW == 1 || halo_mode, "conv_gemm_tc: LayerNorm epilogue with KW = %d needs the halo kernel", a.KW);
        if (halo_mode) return launch<PREC_BF16X3, MODE_HALO, EPI_LN>(mxh, mxl, mwh, mwl, my, p, st, &mxhat, &mplanes);
        return launch<PREC_BF16X3, MODE_CONV, EPI_LN>(mxh, mxl, mwh, mwl, my, p, st, &mxhat, &mplanes);
    }
    if (halo_mode) return launch<PREC_BF16X3, MODE_HALO>(mxh, mxl, mwh, mwl, my, p, st);
    return launch<PREC_BF16X3, MODE_CONV>(mxh, mxl, mwh, mwl, my, p, st);
}

// ---- wgrad -----------------------------------------------------------------------------------------------------------
bool conv_wgrad_tc_supported(const ConvWgradArgs& a) {
    if (a.Cin % 8 != 0 || a.Cout % 8 != 0 || a.ldx % 4 != 0) return false;   // 16-byte pitches of the bf16 row planes
    if (a.Cin < 16 || a.Cout < 16) return false;
    if (((uintptr_t)a.x | (uintptr_t)a.dy) & 15) return false;
    return get_encode() != nullptr;
}

static int wgrad_nsplit(const ConvWgradArgs& a) {
    const int tiles = ceil_div(a.Cout, TM) * ceil_div(a.Cin, TN) * a.KW;
    const int sms = 148;
    int best = 1;
    double best_eff = 0.0;
    const int total_k = a.B * ceil_div(a.S, 64);
    for (int ns = 1; ns <= total_k / 4 && ns <= 64; ++ns) {
        const int ctas = tiles * ns, waves = ceil_div(ctas, sms);
        const double eff = (double)ctas / ((double)waves * sms);
        if (eff > best_eff + 0.02) { best_eff = eff; best = ns; }
        if (eff >= 0.9 && ctas >= sms) { best = ns; break; }
    }
    return best;
}

// the three-tap kernel (wgrad_halo3_kernel) has KW times fewer tiles per split: it splits the reduction finer to fill the SMs
static int wgrad3_nsplit(const ConvWgradArgs& a) {
    const int tiles = ceil_div(a.Cout, TM) * ceil_div(a.Cin, TN);
    const int sms = 148;
    int best = 1;
    double best_eff = 0.0;
    const int total_k = a.B * ceil_div(a.S, 64);
    for (int ns = 1; ns <= total_k / 4 && ns <= 64; ++ns) {
        const int ctas = tiles * ns, waves = ceil_div(ctas, sms);
        const double eff = (double)ctas / ((double)waves * sms);
        if (eff > best_eff + 0.02) { best_eff = eff; best = ns; }
        if (eff >= 0.9 && ctas >= sms) { best = ns; break; }
    }
    return best;
}

bool wgrad_halo3_enabled() {   // DX_WGRAD_HALO3=0: one tile per tap through the generic kernel (A/B timing)
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("DX_WGRAD_HALO3");
        on = (e && atoi(e) == 0) ? 0 : 1;
    }
    return on != 0;
}

size_t conv_wgrad_tc_workspace(const ConvWgradArgs& a) {
    const size_t dye = (size_t)a.B * a.S * a.Cout, xe = (size_t)a.B * a.S * a.Cin;
    const int ns = wgrad_nsplit(a) > wgrad3_nsplit(a) ? wgrad_nsplit(a) : wgrad3_nsplit(a);   // either kernel
    const size_t part = (size_t)ns * a.KW * a.Cout * a.Cin * 4;
    return (a.dy_planes ? 0 : 2 * align256(dye * 2)) + (a.x_planes ? 0 : 2 * align256(xe * 2)) + align256(part) + 256;
}

__global__ void wgrad_reduce_tc_kernel(const float* __restrict__ part, float* __restrict__ dw, int nsplit, int KW, int Cout,
                                       int Cin, float alpha) {
    // part[(split*KW + tap)][co][ci]  ->  dw[co][ci][tap]
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

// ---- deferred, batched split-K reduction ------------------------------------------------------------------------------
// A backward pass issues ~58 weight-gradient GEMMs, each followed by a tiny (latency-bound, ~9 us) reduction of its split-K
// partials into the parameter layout.  With deferral on, the reductions are only RECORDED (descriptor = kernel-parameter data,
// so a CUDA-graph capture needs no side memory) and one launch performs all of them before the gradients are consumed.
struct ReduceDesc {
    const float* part;
    float* dw;
    int nsplit, KW, Cout, Cin;
    float alpha;
    int block0;
};
constexpr int kMaxReduce = 72;          // 72 * 40 B < the 4 KB kernel-parameter space
constexpr int kReduceChunk = 1024;      // elements per block
struct ReduceBatch {
    ReduceDesc d[kMaxReduce];
    int n;
};
bool g_defer_reduce = false;
ReduceBatch g_reduce = {};
int g_reduce_blocks = 0;

// One block = kReduceChunk consecutive (co, ci) positions of one weight gradient, ALL taps.  The partials are read once from DRAM
// (~1.5 GB per backward pass at the bench shapes, the whole cost of this kernel): 16-byte streaming loads, eight split slices in
// flight per thread; k = 3 gradients leave as twelve contiguous floats of dw[co][ci..ci+3][0..2].
__device__ __forceinline__ float4 reduce_slices4(const float4* src, size_t stride4, int nsplit) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int k = 0; k < nsplit; ++k) {
        const float4 a = __ldcs(src + (size_t)k * stride4);
        s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
    return s;
}

__global__ void __launch_bounds__(256) wgrad_reduce_batched_kernel(const __grid_constant__ ReduceBatch b) {
    int lo = 0, hi = b.n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (b.d[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const ReduceDesc& d = b.d[lo];
    const size_t cc = (size_t)d.Cout * d.Cin, per = (size_t)d.KW * cc;
    const size_t i0 = (size_t)((int)blockIdx.x - d.block0) * kReduceChunk;
    const float al = d.alpha;
    if ((d.Cin & 3) == 0 && (reinterpret_cast<uintptr_t>(d.dw) & 15) == 0 && (d.KW == 1 || d.KW == 3)) {
        const size_t i = i0 + 4 * threadIdx.x;                    // kReduceChunk = 4 * 256
        if (i >= cc) return;
        const float4* src = reinterpret_cast<const float4*>(d.part + i);
        if (d.KW == 1) {
            const float4 s = reduce_slices4(src, per / 4, d.nsplit);
            *reinterpret_cast<float4*>(d.dw + i) = make_float4(al * s.x, al * s.y, al * s.z, al * s.w);
        } else {
            const float4 s0 = reduce_slices4(src, per / 4, d.nsplit), s1 = reduce_slices4(src + cc / 4, per / 4, d.nsplit),
                         s2 = reduce_slices4(src + cc / 2, per / 4, d.nsplit);
            float4* dst = reinterpret_cast<float4*>(d.dw + 3 * i);
            dst[0] = make_float4(al * s0.x, al * s1.x, al * s2.x, al * s0.y);
            dst[1] = make_float4(al * s1.y, al * s2.y, al * s0.z, al * s1.z);
            dst[2] = make_float4(al * s2.z, al * s0.w, al * s1.w, al * s2.w);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < kReduceChunk / 256; ++j) {
        const size_t i = i0 + threadIdx.x + 256 * j;              // (co, ci) position
        if (i >= cc) break;
        for (int tap = 0; tap < d.KW; ++tap) {
            float s = 0.f;
            for (int k = 0; k < d.nsplit; ++k) s += d.part[(size_t)k * per + (size_t)tap * cc + i];
            d.dw[i * d.KW + tap] = al * s;
        }
    }
}

int wgrad_reduce_flush(cudaStream_t st) {
    if (g_reduce.n == 0) return DX_OK;
    wgrad_reduce_batched_kernel<<<g_reduce_blocks, 256, 0, st>>>(g_reduce);
    g_reduce.n = 0;
    g_reduce_blocks = 0;
    return check_launch("wgrad_reduce_batched");
}

void wgrad_reduce_defer(bool on) {
    g_defer_reduce = on;
    g_reduce.n = 0;   // anything still pending belongs to an aborted pass
    g_reduce_blocks = 0;
}

int conv_wgrad_tc(const ConvWgradArgs& a, cudaStream_t st) {
    const size_t need = conv_wgrad_tc_workspace(a);
    DX_REQUIRE(a.workspace && a.workspace_bytes >= need, "conv_wgrad_tc: workspace %zu < %zu bytes", a.workspace_bytes, need);
    const int passes = g_passes_wgrad ? g_passes_wgrad : ((long long)a.B * a.S >= kWgradSinglePassRows ? 1 : 3);
    const bool three_tap = passes == 1 && (a.KW == 3 || a.KW == 1) && wgrad_halo3_enabled();   // all taps per CTA (wgrad_halo3_kernel<KW>)
    const int nsplit = three_tap ? wgrad3_nsplit(a) : wgrad_nsplit(a);
    const size_t dye = (size_t)a.B * a.S * a.Cout, xe = (size_t)a.B * a.S * a.Cin;
    uint8_t* ws = (uint8_t*)(((uintptr_t)a.workspace + 255) & ~(uintptr_t)255);
    const size_t dyb = a.dy_planes ? 0 : 2 * align256(dye * 2), xb = a.x_planes ? 0 : 2 * align256(xe * 2);
    const __nv_bfloat16* dh = (const __nv_bfloat16*)ws;
    const __nv_bfloat16* dl = (const __nv_bfloat16*)(ws + align256(dye * 2));
    const __nv_bfloat16* xh = (const __nv_bfloat16*)(ws + dyb);
    const __nv_bfloat16* xl = (const __nv_bfloat16*)(ws + dyb + align256(xe * 2));
    float* part = (float*)(ws + dyb + xb);
    int rc;
    if (a.dy_planes) {
        dh = (const __nv_bfloat16*)a.dy_planes;
        dl = dh + dye;
    } else {
        dim3 g1(ceil_div(a.S, 64), ceil_div(a.Cout, 64), a.B);
        split_planes_kernel<<<g1, 256, 0, st>>>(a.dy, a.Cout, a.S, a.Cout, 0, (__nv_bfloat16*)dh, (__nv_bfloat16*)dl, nullptr, nullptr, a.B);
        if ((rc = check_launch("split_planes(dy)"))) return rc;
    }
    if (a.x_planes) {
        xh = (const __nv_bfloat16*)a.x_planes;
        xl = xh + xe;
    } else {
        dim3 g2(ceil_div(a.S, 64), ceil_div(a.Cin, 64), a.B);
        split_planes_kernel<<<g2, 256, 0, st>>>(a.x, a.ldx, a.S, a.Cin, 0, (__nv_bfloat16*)xh, (__nv_bfloat16*)xl, nullptr, nullptr, a.B);
        if ((rc = check_launch("split_planes(x)"))) return rc;
    }
    if (((uintptr_t)dh | (uintptr_t)dl | (uintptr_t)xh | (uintptr_t)xl) & 15) {
        set_last_error("conv_wgrad_tc: operand planes must be 16-byte aligned");
        return DX_ERR_ARG;
    }
    TcParams p;
    p.bias = nullptr; p.relu_src = nullptr; p.add_src = nullptr;
    p.relu_src_hi = nullptr; p.y_planes = nullptr; p.colsum = nullptr; p.y_plane_elems = 0; p.skip_y = 0;
    p.hp_dh = 0; p.hp_NH = 0; p.hp_scale_cols = 0; p.hp_scale = 1.f; p.hp_dot_src = nullptr; p.hp_dot_out = nullptr;
    p.ln_res = nullptr; p.ln_w = nullptr; p.ln_b = nullptr; p.film = nullptr; p.film_stride = 0; p.ln_rstd = nullptr;
    p.ln_p_in = 0.f; p.ln_seed_in = 0; p.dyn = nullptr; p.ln_planes = 0;
    p.passes = passes;
    p.B = a.B; p.S = a.S; p.Cin = a.Cin; p.Cout = a.Cout; p.KW = a.KW; p.ldy = a.Cin;
    p.tiles_m_per_b = ceil_div(a.Cout, TM);
    p.tiles_n = ceil_div(a.Cin, TN);
    p.nsplit = nsplit;
    p.lens = a.lens; p.halo = a.halo;
    p.num_tiles = p.tiles_m_per_b * p.tiles_n * a.KW * nsplit;
    p.k_chunks = ceil_div(a.S, Cfg<PREC_BF16X3>::TKB);
    p.alpha = 1.f; p.relu = 0; p.round_tf32 = 0;
    p.debug = tc_debug_mask();
    p.trace = nullptr;
    CUtensorMap mah, mal, mbh, mbl, my;
    // (c, s, b) maps over the row planes; box = 64 channels (128 B) x 64 rows: one MN chunk of an MN-major operand tile
    const uint64_t d1 = (uint64_t)a.Cout * 2, d2 = (uint64_t)a.S * a.Cout * 2, x1 = (uint64_t)a.Cin * 2, x2 = (uint64_t)a.S * a.Cin * 2;
    if ((rc = make_map_3d(&mah, dh, 2, a.Cout, a.S, a.B, d1, d2, 64, 64, 1))) return rc;
    if ((rc = make_map_3d(&mal, dl, 2, a.Cout, a.S, a.B, d1, d2, 64, 64, 1))) return rc;
    if ((rc = make_map_3d(&mbh, xh, 2, a.Cin, a.S, a.B, x1, x2, 64, 64, 1))) return rc;
    if ((rc = make_map_3d(&mbl, xl, 2, a.Cin, a.S, a.B, x1, x2, 64, 64, 1))) return rc;
    if ((rc = make_map_3d(&my, part, 4, a.Cin, a.Cout, (uint64_t)nsplit * a.KW, (uint64_t)a.Cin * 4, (uint64_t)a.Cout * a.Cin * 4, 16, 32, 1, true)))
        return rc;
    if (three_tap) {
        CUtensorMap mx66;   // x with one halo row on each side of the 64-row chunk (k = 3)
        if ((rc = make_map_3d(&mx66, xh, 2, a.Cin, a.S, a.B, x1, x2, 64, 64 + a.KW - 1, 1))) return rc;
        Wg3Params q;
        q.lens = a.lens; q.B = a.B; q.S = a.S; q.Cin = a.Cin; q.Cout = a.Cout; q.halo = a.halo;
        q.tiles_m = p.tiles_m_per_b; q.tiles_n = p.tiles_n; q.nsplit = nsplit; q.k_chunks = p.k_chunks;
        q.num_tiles = q.tiles_m * q.tiles_n * nsplit;
        static bool configured = false;
        if (!configured) {
            DX_CUDA(cudaFuncSetAttribute(wgrad_halo3_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG3_SMEM));
            DX_CUDA(cudaFuncSetAttribute(wgrad_halo3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG3_SMEM));
            configured = true;
        }
        const int grid = q.num_tiles < num_sms() ? q.num_tiles : num_sms();
        if (a.KW == 3) wgrad_halo3_kernel<3><<<grid, NTHREADS, WG3_SMEM, st>>>(mah, mx66, my, q);
        else wgrad_halo3_kernel<1><<<grid, NTHREADS, WG3_SMEM, st>>>(mah, mx66, my, q);
        ++g_tc_launches;
        if ((rc = check_launch("wgrad_halo3"))) return rc;
    } else if ((rc = launch<PREC_BF16X3, MODE_WGRAD>(mah, mal, mbh, mbl, my, p, st))) {
        return rc;
    }
    const size_t per = (size_t)a.KW * a.Cout * a.Cin;
    if (g_defer_reduce) {   // the caller keeps the workspace alive until dx_wgrad_flush
        if (g_reduce.n == kMaxReduce && (rc = wgrad_reduce_flush(st))) return rc;
        ReduceDesc& d = g_reduce.d[g_reduce.n++];
        d.part = part; d.dw = a.dw; d.nsplit = nsplit; d.KW = a.KW; d.Cout = a.Cout; d.Cin = a.Cin; d.alpha = a.alpha;
        d.block0 = g_reduce_blocks;
        g_reduce_blocks += (int)(((size_t)a.Cout * a.Cin + kReduceChunk - 1) / kReduceChunk);   // a block = 1024 (co, ci) positions, all taps
    } else {
        wgrad_reduce_tc_kernel<<<grid_1d(per, 256, 148 * 8), 256, 0, st>>>(part, a.dw, nsplit, a.KW, a.Cout, a.Cin, a.alpha);
        if ((rc = check_launch("wgrad_reduce_tc"))) return rc;
    }
    if (a.dbias) return colsum(a.dy, a.dbias, a.B * a.S, a.Cout, a.alpha, st);
    return DX_OK;
}

}  // namespace dx
