// tcgen05 / TMEM / TMA flash-attention FORWARD for sm_100a (head_dim 64 and 16), fp32-grade accuracy (bf16x3 operand split).
//
// Replaces the scaled-dot-product core of nn.MultiheadAttention as called at model.py:182-186 (key padding mask from the
// lengths, dropout on the attention weights); S x S scores never leave the SM.
//
// CTA = (128-query tile, head, utterance), 576 threads, one CTA per SM:
//   warp 0      TMA producer: the Q tile once, a K ring and a V ring of 128-key tiles, all from the per-head bf16 hi|lo operand
//               planes R[plane][b][3H][Sp][dh] written by attn_prep_kernel (q pre-scaled by 1/sqrt(dh)); 128-byte swizzle
//               for dh = 64, 32-byte swizzle for dh = 16; rows beyond the tensor arrive as zeros
//   warp 1      MMA issuer (one thread): S = Q K^T as three `tcgen05.mma kind::f16` per K-step (hi*hi + lo*hi + hi*lo, fp32
//               accumulation in TMEM) into one of TWO score buffers, O += P V with P read straight from TMEM (A operand in
//               tensor memory) and V consumed MN-major from its row-major tile (no transposed copy)
//   warps 2-17  softmax: thread <-> (query row = TMEM lane, 32-column quarter of the key tile); tcgen05.ld of the scores,
//               exp2, row sums, dropout, bf16 hi|lo split, tcgen05.st of P IN PLACE over the scores it came from (hi pairs
//               in the first 16 columns of the quarter, lo pairs in the last 16)
// The score buffers ping-pong: the MMAs of key tile t+1 run while the softmax warps work on tile t.
//
// Softmax, default (ONLINE): ONE pass over the key tiles with a LAZY running maximum.  Row r keeps a reference maximum M_r; a tile
// whose row maximum exceeds M_r by more than 8 / log2(e) raises it and rescales the row sum and the O row in TMEM, otherwise the tile
// is exponentiated against the old M_r (weights <= 2^8, cancelled by the normalisation).  After the first tile hardly any tile moves
// M_r, so the O correction (TMEM load / scale / store behind the P V of the previous tile) runs a few times per CTA.
// DX_ATTN_FWD_ONLINE=0 selects the exact two-pass form of round 1 (pass A forms S and keeps the row maxima, pass B forms S again and
// exponentiates against the final maximum: 1.5x the Q K^T flops and a second TMEM read).  Measured at the bench shapes: head_dim 64
// 105.5 -> 96.4 us per layer call, head_dim 16 199.6 -> 195.4 us; the kernel is bound by the ~13 ALU instructions per score of the
// exp / row sum / dropout / hi|lo split (issue slots 65 % busy), so dropping the max pass buys less than its share of the tiles.
// TMEM columns: [0,128) and [128,256) the two S/P buffers, [256, 256+dh) O.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace dx {

namespace {

using namespace tcptx;

constexpr int TQ = 128, TKEY = 128;
// Two shapes of the same kernel:
//   NBUF = 2 score buffers, NSW = 16 softmax warps (thread = row x 32-key quarter), all 512 TMEM columns, one CTA per SM  (dh = 64)
//   NBUF = 1 score buffer,  NSW = 8 softmax warps (thread = row x 64-key half), 256 TMEM columns, TWO CTAs per SM       (dh = 16):
//            the small K/V tiles leave room for a second CTA, whose softmax arithmetic fills this one's max pass, prologue,
//            epilogue and the gaps between consecutive CTAs (the per-score ALU work is the bound for dh = 16)
template <int NSW> constexpr int atc_threads() { return 64 + 32 * NSW; }   // + TMA warp, MMA warp
template <int NBUF> constexpr int atc_tmem_cols() { return NBUF == 2 ? 512 : 256; }
constexpr float kLog2e = 1.4426950408889634f;

template <int DH, int NBUF = 2>
struct AtcCfg {
    static_assert(DH == 64 || DH == 16, "tcgen05 attention: head_dim 64 (128-byte rows) or 16 (32-byte rows)");
    static constexpr int ROW_BYTES = DH * 2;
    static constexpr int TILE_BYTES = 128 * ROW_BYTES;              // one plane of a 128-row operand tile
    static constexpr uint64_t LAYOUT = DH == 64 ? 2ull : 6ull;       // SWIZZLE_128B / SWIZZLE_32B
    static constexpr uint32_t SBO = 8 * ROW_BYTES;                    // 8-row groups
    static constexpr int KS = DH / 16;                                // K-steps of Q K^T
    static constexpr uint32_t V_KSTEP = (16 * ROW_BYTES) >> 4;        // 16 key rows per K-step of P V (MN-major B)
    // head_dim 64 with one score buffer (two CTAs per SM): one K and one V stage are enough — S(t+1) cannot be issued before
    // P V(t) anyway, K(t+1) lands during the softmax of tile t and V(t+1) during S(t+1) and its softmax
    static constexpr int K_STAGES = DH == 64 ? (NBUF == 2 ? 3 : 1) : 6, V_STAGES = DH == 64 ? (NBUF == 2 ? 2 : 1) : 4;
    static constexpr int SMEM_TILES = 2 + 2 * K_STAGES + 2 * V_STAGES;   // Q, K ring, V ring; hi|lo planes each
    static constexpr int SMEM_BYTES = SMEM_TILES * TILE_BYTES + 256 + 2 * 4 * TQ * 4 + 64 + 1024;   // + barriers + max/sum exchange (x2)
};
// K-major tile (rows x dh): 8-row groups SBO apart
template <int DH>
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(AtcCfg<DH>::SBO >> 4) << 32) | (1ull << 46) |
           (AtcCfg<DH>::LAYOUT << 61);
}
// MN-major tile (K = rows, N = dh contiguous, one swizzle-wide chunk): 8-row K groups SBO apart; LBO (between MN chunks) unused
template <int DH>
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(AtcCfg<DH>::SBO >> 4) << 32) | (1ull << 46) |
           (AtcCfg<DH>::LAYOUT << 61);
}
// fp32 accumulate, bf16 operands, A K-major (or TMEM), B K-major (b_mn = 0) or MN-major (b_mn = 1)
__host__ __device__ constexpr uint32_t atc_idesc(int M, int N, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct AtcParams {
    float* ctx;              // [B, S, D]
    float* lse;              // [B, H, S]
    __nv_bfloat16* ctx_planes;   // nullable [2][B*S][D]
    const long long* lens;
    int B, S, H, Sp;
    float dropout_p;
    unsigned long long seed;
    const StepState* dyn;
    int cta_trace;
    long long* trace;        // optional [4][256] clock64 trace of block (0,0,0): softmax warp 2 lane 0 (rows 0-2), MMA thread (row 3)
};

template <int DH, int NBUF, int NSW, bool ONLINE>
__global__ void __launch_bounds__(atc_threads<NSW>(), NBUF == 1 ? 2 : 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_r, AtcParams p) {
    using C = AtcCfg<DH, NBUF>;
    constexpr int NCG = NSW / 4;             // column groups of a 128-key tile
    constexpr int CPT = 4 / NCG;             // 32-column chunks per thread
    constexpr int OC = DH == 64 ? DH / NCG : 16;   // O columns per thread in the correction step and the epilogue (head_dim 16: column group 0 takes all)
    constexpr int O_COL0 = NBUF * TKEY;      // O accumulator behind the score buffer(s)
    constexpr int ATC_TMEM_COLS = atc_tmem_cols<NBUF>();
    constexpr int ATC_SOFTMAX_WARPS = NSW;
    constexpr int TB = C::TILE_BYTES, KST = C::K_STAGES, VST = C::V_STAGES;
    extern __shared__ uint8_t atc_smem_raw[];
    const uint32_t base = (smem_u32(atc_smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = base;                       // [plane]
    const uint32_t sK = sQ + 2 * TB;                // [stage][plane]
    const uint32_t sV = sK + KST * 2 * TB;          // [stage][plane]
    const uint32_t bars = sV + VST * 2 * TB;
    // barriers (8 bytes each): q_full | k_full[8] k_empty[8] | v_full[4] v_empty[4] | s_full[2] s_free[2] p_full[2] | o_full
    const uint32_t q_full = bars, k_full0 = bars + 8, k_empty0 = bars + 72, v_full0 = bars + 136, v_empty0 = bars + 168;
    const uint32_t s_full0 = bars + 200, s_free0 = bars + 216, p_full0 = bars + 232, o_full = bars + 248;
    const uint32_t red0 = bars + 256;               // float [2][4 column quarters][128 rows]
    const uint32_t tmem_slot = red0 + 2 * 4 * TQ * 4;
    const uint32_t o_ready = s_free0;               // ONLINE: P V of tile t retired (count 1); the max pass and its barrier do not exist
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(atc_smem_raw + (tmem_slot - smem_u32(atc_smem_raw)));
    float* xch = reinterpret_cast<float*>(atc_smem_raw + (red0 - smem_u32(atc_smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
    const int D = p.H * DH, NH = 3 * p.H;
    const int len = min((int)p.lens[b], p.S);
    float* ctx = p.ctx + (size_t)b * p.S * D + h * DH;
    float* lse = p.lse + ((size_t)b * p.H + h) * p.S;
    __nv_bfloat16* cph = p.ctx_planes ? p.ctx_planes + (size_t)b * p.S * D + h * DH : nullptr;
    const size_t cplane = (size_t)p.B * p.S * D;

    const bool tr0 = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
    if (tr0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.trace[3 * 256 + 200] = clock64(); p.trace[3 * 256 + 204] = (long long)gt;
    }
    // DX_ATTN_CTA_TRACE (debug): per-CTA {start ns, end ns, SM id} records behind the [4][256] trace (buffer must hold 1024 + 3 * #CTAs)
    const long long cta_lin = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (p.trace && p.cta_trace && threadIdx.x == 0) {
        unsigned long long gt; uint32_t smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[1024 + 3 * cta_lin] = (long long)gt; p.trace[1024 + 3 * cta_lin + 2] = smid;
    }
    if (q0 >= len) {   // whole tile is padding: zeros (their only consumer masks them, model.py:259)
        if (threadIdx.x < TQ) {
            const int r = q0 + threadIdx.x;
            if (r < p.S) {
                for (int c = 0; c < DH; c += 4) {
                    *reinterpret_cast<float4*>(ctx + (size_t)r * D + c) = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (cph) {
                        *reinterpret_cast<uint2*>(cph + (size_t)r * D + c) = make_uint2(0u, 0u);
                        *reinterpret_cast<uint2*>(cph + cplane + (size_t)r * D + c) = make_uint2(0u, 0u);
                    }
                }
                lse[r] = 0.f;
            }
        }
        return;
    }

    const int n_tiles = (len + TKEY - 1) / TKEY;
    // step t: key tile t % n_tiles, score buffer t % NBUF.  Two-pass: t < n_tiles = pass A (maxima), else pass B (P V).  ONLINE: one pass
    const int n_steps = ONLINE ? n_tiles : 2 * n_tiles;
    const int t_b0 = ONLINE ? 0 : n_tiles;   // first P V step
    // plane-head slices of R: (plane, b, which, h) -> index ((plane * B + b) * 3H + which * H + h)
    const int sl_q = b * NH + h, sl_k = b * NH + p.H + h, sl_v = b * NH + 2 * p.H + h, sl_lo = p.B * NH;
    // TMA producer state (thread 0 only) and one producer step: K tile of step t (+ its V tile in a P V step)
    int pks = 0, pkph = 0, pvs = 0, pvph = 0;
    auto produce = [&](int t) {
        const int j = t < n_tiles ? t : t - n_tiles;
        mbar_wait(k_empty0 + 8 * pks, pkph ^ 1);
        mbar_expect_tx(k_full0 + 8 * pks, 2 * TB);
        tma_load_3d(sK + (pks * 2 + 0) * TB, &map_r, k_full0 + 8 * pks, 0, j * TKEY, sl_k);
        tma_load_3d(sK + (pks * 2 + 1) * TB, &map_r, k_full0 + 8 * pks, 0, j * TKEY, sl_lo + sl_k);
        if (++pks == KST) { pks = 0; pkph ^= 1; }
        if (t >= t_b0) {
            mbar_wait(v_empty0 + 8 * pvs, pvph ^ 1);
            mbar_expect_tx(v_full0 + 8 * pvs, 2 * TB);
            tma_load_3d(sV + (pvs * 2 + 0) * TB, &map_r, v_full0 + 8 * pvs, 0, j * TKEY, sl_v);
            tma_load_3d(sV + (pvs * 2 + 1) * TB, &map_r, v_full0 + 8 * pvs, 0, j * TKEY, sl_lo + sl_v);
            if (++pvs == VST) { pvs = 0; pvph ^= 1; }
        }
    };
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_r)) : "memory");
        mbar_init(q_full, 1);
        for (int i = 0; i < KST; ++i) { mbar_init(k_full0 + 8 * i, 1); mbar_init(k_empty0 + 8 * i, 1); }
        for (int i = 0; i < VST; ++i) { mbar_init(v_full0 + 8 * i, 1); mbar_init(v_empty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_full0 + 8 * i, 1); mbar_init(s_free0 + 8 * i, ONLINE ? 1 : ATC_SOFTMAX_WARPS); mbar_init(p_full0 + 8 * i, ATC_SOFTMAX_WARPS);
        }
        mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the first loads go out BEFORE the TMEM allocation and the CTA-wide barrier (this thread initialised the mbarriers they
        // signal): ~0.9 k cycles of the ~4 k-cycle entry -> first MMA latency (3 CTAs per SM back to back at the bench shapes)
        mbar_expect_tx(q_full, 2 * TB);
        tma_load_3d(sQ, &map_r, q_full, 0, q0, sl_q);
        tma_load_3d(sQ + TB, &map_r, q_full, 0, q0, sl_lo + sl_q);
        produce(0);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(ATC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (tr0) p.trace[3 * 256 + 201] = clock64();
    if (warp == 0) {
        if (lane == 0) {
            for (int t = 1; t < n_steps; ++t) produce(t);
        }
    } else if (warp == 1) {
        // The WHOLE warp runs the issue loop (warp-uniform control flow lets the compiler keep descriptors and addresses in
        // uniform registers: a divergent `if (lane == 0)` costs ~12 instructions + a broadcast loop per MMA, which starves
        // behind the softmax warps sharing this scheduler); one elected lane issues the MMAs and commits.
        constexpr uint32_t idesc_s = atc_idesc(TQ, TKEY, 0), idesc_o = atc_idesc(TQ, DH, 1);
        int ks = 0, kph = 0, vs = 0, vph = 0;
        const uint64_t qh = desc_k<DH>(sQ), ql = desc_k<DH>(sQ + TB);
        const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
        // S(t) = Q K^T (3 passes per K-step) into score buffer t & 1; frees the K stage when the MMAs retire
        auto issue_s = [&](int t) {
            mbar_wait(k_full0 + 8 * ks, kph);
            tc_fence_after();
            const uint32_t tmem_s = tmem_base + (t % NBUF) * TKEY;
            const uint64_t kh = desc_k<DH>(sK + (ks * 2 + 0) * TB), kl = desc_k<DH>(sK + (ks * 2 + 1) * TB);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < C::KS; ++kk) {
                    umma_ss(tmem_s, qh + 2 * kk, kh + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);
                    umma_ss(tmem_s, ql + 2 * kk, kh + 2 * kk, idesc_s, 1u);
                    umma_ss(tmem_s, qh + 2 * kk, kl + 2 * kk, idesc_s, 1u);
                }
                umma_commit(s_full0 + 8 * (t % NBUF));
                umma_commit(k_empty0 + 8 * ks);
                if (tr && t < 256) p.trace[3 * 256 + t] = clock64();
            }
            __syncwarp();
            if (++ks == KST) { ks = 0; kph ^= 1; }
        };
        mbar_wait(q_full, 0);
        issue_s(0);
        if (NBUF > 1 && n_steps > 1) issue_s(1);
        uint32_t sfree_cnt0 = 0, sfree_cnt1 = 0, pfull_cnt0 = 0, pfull_cnt1 = 0;
        for (int t = 0; t < n_steps; ++t) {
            const int buf = t % NBUF;
            if (t < t_b0) {      // pass A: the softmax warps have read S(t): its buffer may take S(t + 2)
                const uint32_t par = (buf ? sfree_cnt1 : sfree_cnt0) & 1;
                if (buf) ++sfree_cnt1; else ++sfree_cnt0;
                mbar_wait(s_free0 + 8 * buf, par);
                tc_fence_after();
            } else {             // pass B: P(t) is in TMEM (in place of S(t)): O += P V, then the buffer may take S(t + 2)
                const uint32_t par = (buf ? pfull_cnt1 : pfull_cnt0) & 1;
                if (buf) ++pfull_cnt1; else ++pfull_cnt0;
                mbar_wait(p_full0 + 8 * buf, par);
                mbar_wait(v_full0 + 8 * vs, vph);
                tc_fence_after();
                const uint64_t vh = desc_mn<DH>(sV + (vs * 2 + 0) * TB), vl = desc_mn<DH>(sV + (vs * 2 + 1) * TB);
                const uint32_t tmem_o = tmem_base + O_COL0, tmem_p = tmem_base + buf * TKEY;
                const uint32_t acc0 = t > t_b0 ? 1u : 0u;
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < TKEY / 16; ++kk) {
                        const uint32_t ph = tmem_p + 32 * (kk >> 1) + 8 * (kk & 1), pl = ph + 16;
                        umma_ts(tmem_o, ph, vh + C::V_KSTEP * kk, idesc_o, kk > 0 ? 1u : acc0);
                        umma_ts(tmem_o, pl, vh + C::V_KSTEP * kk, idesc_o, 1u);
                        umma_ts(tmem_o, ph, vl + C::V_KSTEP * kk, idesc_o, 1u);
                    }
                    umma_commit(v_empty0 + 8 * vs);
                    if (ONLINE) umma_commit(o_ready);   // phase t: O holds the sum over tiles 0..t
                }
                __syncwarp();
                if (++vs == VST) { vs = 0; vph ^= 1; }
            }
            if (t + NBUF < n_steps) issue_s(t + NBUF);   // (the tensor pipe is in order: it runs after the P V that read this buffer)
        }
        if (elect_one()) umma_commit(o_full);
        __syncwarp();
    } else {
        // 16 softmax warps: TMEM lane quarter = warp % 4 (hardware rule), column quarter cq = keys [32 cq, 32 cq + 32) of every tile;
        // the four column quarters of a row keep separate running maxima / sums and combine them once per pass through smem
        const int quad = warp & 3, cq = (warp - 2) >> 2;
        const int rl = quad * 32 + lane;                         // row within the tile = TMEM lane
        const int r = q0 + rl;                                   // this thread's query row
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
        const unsigned long long seed = dyn_seed(p.seed, p.dyn);
        const bool drop = p.dropout_p > 0.f;
        const uint32_t thresh = drop_threshold(p.dropout_p);
        const uint32_t rk = hash_u32(seed, ((unsigned long long)b * p.H + h) * (unsigned long long)p.S + r);   // same keys as the backward
        float m, l;
        if constexpr (!ONLINE) {
        // ---- pass A: row maximum over all valid keys ----
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int t = 0; t < n_tiles; ++t) {
            const int buf = t % NBUF;
            mbar_wait(s_full0 + 8 * buf, (t / NBUF) & 1);
            if (tr) p.trace[t] = clock64();
            tc_fence_after();
            uint32_t v[32 * CPT];
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) tmem_ld32(t_lane + buf * TKEY + 32 * (cq * CPT + cc), v + 32 * cc);
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) tmem_ld_wait32(v + 32 * cc);
            if (tr) p.trace[256 + t] = clock64();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free0 + 8 * buf);
            if ((t + 1) * TKEY > len) {
                const int key0 = t * TKEY + 32 * cq * CPT;
#pragma unroll
                for (int i = 0; i < 32 * CPT; ++i) m4[i & 3] = fmaxf(m4[i & 3], key0 + i < len ? __uint_as_float(v[i]) : -INFINITY);
            } else {
#pragma unroll
                for (int i = 0; i < 32 * CPT; i += 2) m4[(i >> 1) & 3] = fmaxf(m4[(i >> 1) & 3], fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
            }
            if (tr) p.trace[512 + t] = clock64();
        }
        m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        xch[cq * TQ + rl] = m;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NSW) : "memory");
#pragma unroll
        for (int c = 0; c < NCG; ++c) m = fmaxf(m, xch[c * TQ + rl]);   // finite: key 0 is always valid
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NSW) : "memory");
        // ---- pass B: P = exp(S - m) in place, row sums ----
        const float m2 = m * kLog2e;
        float l0 = 0.f, l1 = 0.f;
        for (int t = n_tiles; t < n_steps; ++t) {
            const int buf = t % NBUF, j = t - n_tiles;
            mbar_wait(s_full0 + 8 * buf, (t / NBUF) & 1);
            if (tr) p.trace[t] = clock64();
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) {
            uint32_t v[32], o[32];
            const uint32_t taddr = t_lane + buf * TKEY + 32 * (cq * CPT + cc);
            tmem_ld32(taddr, v);
            tmem_ld_wait32(v);
            if (tr) p.trace[256 + t] = clock64();
            const int key0 = j * TKEY + 32 * (cq * CPT + cc);
            if ((j + 1) * TKEY > len) {   // last tile: keys >= len score -inf (a separate, rarely taken pre-pass keeps the hot loop unpredicated)
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (key0 + i >= len) v[i] = 0xff800000u;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float p0 = ex2(fmaf(__uint_as_float(v[2 * i]), kLog2e, -m2));
                float p1 = ex2(fmaf(__uint_as_float(v[2 * i + 1]), kLog2e, -m2));
                l0 += p0; l1 += p1;
                if (drop) {   // dropout on the attention weights: applied to what multiplies V, not to the row sum
                    p0 = drop_keep(rk, drop_col_term((uint32_t)(key0 + 2 * i)), thresh) ? p0 : 0.f;
                    p1 = drop_keep(rk, drop_col_term((uint32_t)(key0 + 2 * i + 1)), thresh) ? p1 : 0.f;
                }
                split_pair(p0, p1, o[i], o[16 + i]);
            }
            tmem_st32(taddr, o);
            }   // cc
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full0 + 8 * buf);
            if (tr) p.trace[512 + t] = clock64();
        }
        xch[cq * TQ + rl] = l0 + l1;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NSW) : "memory");
        l = 0.f;
#pragma unroll
        for (int c = 0; c < NCG; ++c) l += xch[c * TQ + rl];
        } else {
        // ---- ONE pass, online softmax with a LAZY reference maximum: row r keeps M_r; a key tile whose row maximum exceeds
        // M_r by more than kTau raises M_r (and rescales the row sum and the O row in TMEM by exp(M_old - M_new)); otherwise the
        // tile is exponentiated against the old M_r (P <= 2^8: harmless in fp32 / bf16x3, and the normalisation cancels it).
        // After the first tile almost no tile moves M_r, so the O correction (TMEM load, scale, store, ordered behind the P V
        // of the previous tile) runs for a few tiles per CTA only.  The threads sharing a row exchange their tile maxima
        // through shared memory with one named barrier per TMEM lane quarter.
        constexpr float kTau = 8.f / kLog2e;
        m = -INFINITY;
        float l0 = 0.f, l1 = 0.f;
        for (int t = 0; t < n_tiles; ++t) {
            const int buf = t % NBUF;
            mbar_wait(s_full0 + 8 * buf, (t / NBUF) & 1);
            if (tr) p.trace[t] = clock64();
            tc_fence_after();
            // thread's scores: kept in registers between the max and the exp for one 32-column chunk per thread; with two
            // chunks per thread (two CTAs per SM, 96 registers) the chunks are read from TMEM again instead of spilling
            const int key0 = t * TKEY + 32 * cq * CPT;
            const bool last = (t + 1) * TKEY > len;
            auto load_chunk = [&](int cc, uint32_t* d) {
                tmem_ld32(t_lane + buf * TKEY + 32 * (cq * CPT + cc), d);
                tmem_ld_wait32(d);
                if (last) {   // keys >= len score -inf
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (key0 + 32 * cc + i >= len) d[i] = 0xff800000u;
                }
            };
            uint32_t v[32];
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) {
                load_chunk(cc, v);
#pragma unroll
                for (int i = 0; i < 32; i += 2) m4[(i >> 1) & 3] = fmaxf(m4[(i >> 1) & 3], fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
            }
            if (tr) p.trace[256 + t] = clock64();
            float* x = xch + (t & 1) * 4 * TQ;
            x[cq * TQ + rl] = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * NCG) : "memory");
            float tm = x[rl];
#pragma unroll
            for (int c = 1; c < NCG; ++c) tm = fmaxf(tm, x[c * TQ + rl]);
            const bool need = tm > m + kTau;                       // first tile: m = -inf, tm finite or -inf (row of a dead tile part)
            const float mn = need ? tm : m;
            const float alpha = need ? ex2((m - mn) * kLog2e) : 1.f;   // first tile: 0
            l0 *= alpha; l1 *= alpha;
            if (t > 0 && __any_sync(0xffffffffu, need)) {         // warp-uniform: the O rows of this warp take their factors
                if (DH == 64 || cq == 0) {
                    mbar_wait(o_ready, (t - 1) & 1);               // P V of tile t-1 has retired (at most one phase behind: P V(t) waits for us)
                    tc_fence_after();
#pragma unroll
                    for (int c0 = cq * OC; c0 < cq * OC + OC; c0 += 16) {
                        uint32_t ov[16];
                        tmem_ld16(t_lane + O_COL0 + c0, ov);
                        tmem_ld_wait16(ov);
#pragma unroll
                        for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                        tmem_st16(t_lane + O_COL0 + c0, ov);
                    }
                }
            }
            m = mn;
            const float m2 = m * kLog2e;
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) {
                uint32_t o[32];
                if (CPT > 1) load_chunk(cc, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float p0 = ex2(fmaf(__uint_as_float(v[2 * i]), kLog2e, -m2));
                    float p1 = ex2(fmaf(__uint_as_float(v[2 * i + 1]), kLog2e, -m2));
                    l0 += p0; l1 += p1;
                    if (drop) {
                        p0 = drop_keep(rk, drop_col_term((uint32_t)(key0 + 32 * cc + 2 * i)), thresh) ? p0 : 0.f;
                        p1 = drop_keep(rk, drop_col_term((uint32_t)(key0 + 32 * cc + 2 * i + 1)), thresh) ? p1 : 0.f;
                    }
                    split_pair(p0, p1, o[i], o[16 + i]);
                }
                tmem_st32(t_lane + buf * TKEY + 32 * (cq * CPT + cc), o);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full0 + 8 * buf);
            if (tr) p.trace[512 + t] = clock64();
        }
        float* x = xch + (n_tiles & 1) * 4 * TQ;
        x[cq * TQ + rl] = l0 + l1;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * NCG) : "memory");
        l = 0.f;
#pragma unroll
        for (int c = 0; c < NCG; ++c) l += x[c * TQ + rl];
        }
        // ---- epilogue: O / l -> ctx (+ operand planes), lse; the column quarters split the head dimension ----
        mbar_wait(o_full, 0);
        tc_fence_after();
        const bool valid = r < len;
        const float inv_keep = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
        const float sc = valid ? inv_keep / l : 0.f;
        if (DH == 64 || cq == 0)
#pragma unroll
        for (int c0 = cq * OC; c0 < cq * OC + OC; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_lane + O_COL0 + c0, v);
            tmem_ld_wait16(v);
            if (r < p.S) {
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * sc;
                float* dst = ctx + (size_t)r * D + c0;
#pragma unroll
                for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                if (cph) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) split_pair(f[2 * i], f[2 * i + 1], hi[i], lo[i]);
                    uint4* dh_ = reinterpret_cast<uint4*>(cph + (size_t)r * D + c0);
                    uint4* dl_ = reinterpret_cast<uint4*>(cph + cplane + (size_t)r * D + c0);
                    dh_[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); dh_[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    dl_[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); dl_[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                }
            }
        }
        if (cq == 0 && r < p.S) lse[r] = valid ? m + logf(l) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (tr0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.trace[3 * 256 + 202] = clock64(); p.trace[3 * 256 + 205] = (long long)gt;
    }
    if (p.trace && p.cta_trace && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.trace[1024 + 3 * cta_lin + 1] = (long long)gt;
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ATC_TMEM_COLS) : "memory");
    }
}

template <int DH, int NBUF, int NSW, bool ONLINE>
int launch_fwd_tc(const AttnArgs& a, cudaStream_t st) {
    using C = AtcCfg<DH, NBUF>;
    // NBUF = 2: one CTA per SM (the kernel owns all 512 TMEM columns): request more than half of the shared memory.
    // NBUF = 1: two CTAs per SM (256 TMEM columns each): request what is needed (< half).
    const int smem = NBUF == 2 ? (C::SMEM_BYTES > 120 * 1024 ? C::SMEM_BYTES : 120 * 1024) : C::SMEM_BYTES;
    static_assert(NBUF == 2 || C::SMEM_BYTES <= 110 * 1024, "two CTAs per SM need at most half of the shared memory each");
    static bool configured = false;
    if (!configured) {
        DX_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<DH, NBUF, NSW, ONLINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    CUtensorMap map;
    // R as (dh, Sp, 2 * B * 3H): box = dh x 128 rows of one (plane, utterance, q|k|v, head) slice
    const int rc = make_tma_map_3d(&map, a.R, 2, (unsigned long long)DH, (unsigned long long)a.Sp, (unsigned long long)2 * a.B * 3 * a.H,
                                   (unsigned long long)DH * 2, (unsigned long long)a.Sp * DH * 2, DH, 128, 1, DH == 64 ? 128 : 32);
    if (rc) return rc;
    AtcParams p;
    p.ctx = a.ctx; p.lse = a.lse; p.ctx_planes = (__nv_bfloat16*)a.ctx_planes; p.lens = a.lens;
    p.B = a.B; p.S = a.S; p.H = a.H; p.Sp = a.Sp; p.dropout_p = a.dropout_p; p.seed = a.seed; p.dyn = a.dyn;
    p.trace = tc_trace_buffer();
    p.cta_trace = getenv("DX_ATTN_CTA_TRACE") != nullptr;
    dim3 grid(ceil_div(a.S, TQ), a.H, a.B);
    attn_fwd_tc_kernel<DH, NBUF, NSW, ONLINE><<<grid, atc_threads<NSW>(), smem, st>>>(map, p);
    return check_launch("attn_fwd_tc");
}

}  // namespace

bool attention_fwd_tc_supported(const AttnArgs& a) {
    return (a.dh == 64 || a.dh == 16) && a.lens != nullptr && a.R != nullptr && tma_available();
}

// a.R / a.Sp must be bound (attention_mma.cu: bind_planes + prep)
int attention_fwd_tc(const AttnArgs& a, cudaStream_t st) {
    static int online = -1;   // DX_ATTN_FWD_ONLINE=0: the exact two-pass softmax (max pass + P V pass) instead of the one-pass lazy-maximum form
    if (online < 0) { const char* e = getenv("DX_ATTN_FWD_ONLINE"); online = (e && atoi(e) == 0) ? 0 : 1; }
    if (a.dh == 64) {
        static int two64 = -1;   // DX_ATTN_FWD64_2CTA=1: head_dim 64 as two single-buffer CTAs per SM (A/B timing)
        if (two64 < 0) { const char* e = getenv("DX_ATTN_FWD64_2CTA"); two64 = (e && atoi(e) != 0) ? 1 : 0; }
        if (two64) return online ? launch_fwd_tc<64, 1, 8, true>(a, st) : launch_fwd_tc<64, 1, 8, false>(a, st);
        return online ? launch_fwd_tc<64, 2, 16, true>(a, st) : launch_fwd_tc<64, 2, 16, false>(a, st);
    }
    if (a.dh == 16) {
        static int two = -1;   // DX_ATTN_FWD_2CTA=0: the one-CTA-per-SM shape for head_dim 16 as well (A/B timing)
        if (two < 0) { const char* e = getenv("DX_ATTN_FWD_2CTA"); two = (e && atoi(e) == 0) ? 0 : 1; }
        if (online) return two ? launch_fwd_tc<16, 1, 8, true>(a, st) : launch_fwd_tc<16, 2, 16, true>(a, st);
        return two ? launch_fwd_tc<16, 1, 8, false>(a, st) : launch_fwd_tc<16, 2, 16, false>(a, st);
    }
    set_last_error("attention_fwd_tc: unsupported head_dim %d", a.dh);
    return DX_ERR_UNSUPPORTED;
}

}  // namespace dx
