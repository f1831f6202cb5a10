// tcgen05 / TMEM / TMA flash-attention BACKWARD for sm_100a (head_dim 64 and 16), fp32-grade accuracy (bf16x3 operand split).
//
// Gradient of the scaled-dot-product core of nn.MultiheadAttention as called at model.py:182-186 (key padding mask, dropout on
// the attention weights) with respect to q, k, v; scores are recomputed, S x S never leaves the SM.
//
// CTA = (128-key tile, head, utterance), 576 threads, one CTA per SM; it walks the valid 128-query tiles:
//   warp 0      TMA producer: K, V tiles once; Q and dO tiles per query tile (per-head bf16 hi|lo planes R / GR)
//   warp 1      MMA issuer (warp-uniform loop, one elected lane), five products per (key tile, query tile), each as three
//               `tcgen05.mma kind::f16` per K-step (hi*hi + lo*hi + hi*lo), fp32 accumulators in TMEM:
//                 S^T  = K Q^T        [keys x queries]   (A, B K-major from smem)
//                 dP^T = V dO^T       [keys x queries]
//                 dV  += P^T dO       A = P^T (dropped) read from TMEM in place of S^T, B = dO tile MN-major
//                 dK  += dS^T Q       A = dS^T read from TMEM in place of dP^T,       B = Q tile MN-major (q carries 1/sqrt(dh))
//                 dQ   = dS K         A = dS^T staged in shared memory [keys][queries], consumed MN-major; B = K tile MN-major
//   warps 2-17  thread <-> (key row = TMEM lane, 32-query quarter): tcgen05.ld S^T and dP^T, P^T = exp(S^T - lse),
//               dS^T = P^T (dP^T * drop - delta), bf16 hi|lo split, tcgen05.st in place + swizzled st.shared of dS^T; then
//               drain dQ (TMEM lane = query row) with fp32 atomics
// TMEM columns: [0,128) S^T / P^T, [128,256) dP^T / dS^T, [256,256+dh) dK, [320,320+dh) dV, [384,384+dh) dQ.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace dx {

namespace {

using namespace tcptx;

constexpr int BT = 128;                                   // keys per CTA, queries per tile
constexpr int ABT_SOFTMAX_WARPS = 16;
constexpr int ABT_THREADS = 64 + 32 * ABT_SOFTMAX_WARPS;
constexpr int ABT_TMEM_COLS = 512;
constexpr float kLog2eB = 1.4426950408889634f;

template <int DH>
struct AbtCfg {
    static_assert(DH == 64 || DH == 16, "tcgen05 attention backward: head_dim 64 or 16");
    static constexpr int ROW_BYTES = DH * 2;
    static constexpr int TB = 128 * ROW_BYTES;                        // one plane of a 128-row operand tile
    static constexpr uint64_t LAYOUT = DH == 64 ? 2ull : 6ull;         // SWIZZLE_128B / SWIZZLE_32B
    static constexpr uint32_t SBO = 8 * ROW_BYTES;
    static constexpr int KS = DH / 16;                                  // K-steps over the head dimension
    static constexpr uint32_t ROW_KSTEP = (16 * ROW_BYTES) >> 4;        // 16 rows per K-step of an MN-major (dh-contiguous) tile
    static constexpr int QSTAGES = DH == 64 ? 1 : 2;                    // Q / dO ring
    static constexpr int DS_BYTES = 2 * 2 * 128 * 128;                  // dS^T staging: [plane][query chunk of 64][128 keys][128 B]
    static constexpr int DS_BUFS = DH == 64 ? 1 : 2;                    // the pipelined dh = 16 kernel double-buffers it (no wait on dQ before refilling)
    static constexpr int AUX_BYTES = 2 * 3 * BT * 4;                    // lse*log2e, delta, dropout row keys; double buffered
    static constexpr int SMEM_BYTES = (4 + 4 * QSTAGES) * TB + DS_BUFS * DS_BYTES + 256 + AUX_BYTES + 64 + 1024;
    static constexpr int COL_DK = 256, COL_DV = 320, COL_DQ = 384;
};
template <int DH>
__device__ __forceinline__ uint64_t bdesc_k(uint32_t saddr) {   // K-major tile (rows x dh)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(AbtCfg<DH>::SBO >> 4) << 32) | (1ull << 46) | (AbtCfg<DH>::LAYOUT << 61);
}
template <int DH>
__device__ __forceinline__ uint64_t bdesc_mn(uint32_t saddr) {  // the same tile consumed MN-major (K = rows, N = dh contiguous)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(AbtCfg<DH>::SBO >> 4) << 32) | (1ull << 46) |
           (AbtCfg<DH>::LAYOUT << 61);
}
// dS^T staging as the A operand of dQ = dS K: MN-major, M = queries in two 64-element (128-byte) chunks 16 KB apart (LBO), K = key
// rows in 8-row groups 1024 B apart (SBO), SWIZZLE_128B
__device__ __forceinline__ uint64_t bdesc_ds(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(16384 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct AbtParams {
    float* dqkv;             // [B, S, 3D], zero-initialised by the caller (dq is accumulated with atomics)
    const float* lse;        // [B, H, S]
    const float* delta;      // [B, H, S]
    const long long* lens;
    int B, S, H, Sp;
    float dropout_p;
    unsigned long long seed;
    const StepState* dyn;
    long long* trace;        // optional [4][256] clock64 trace of CTA (0,0,0), softmax warp 2 lane 0: s_full seen / P^T,dS^T published / dq_full seen / dQ drained
};

template <int DH>
__global__ void __launch_bounds__(ABT_THREADS, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_r,
                                                                     const __grid_constant__ CUtensorMap map_g, AbtParams p) {
    using C = AbtCfg<DH>;
    constexpr int TB = C::TB, QST = C::QSTAGES;
    extern __shared__ uint8_t abt_smem_raw[];
    const uint32_t base = (smem_u32(abt_smem_raw) + 1023u) & ~1023u;
    const uint32_t sK = base, sV = base + 2 * TB;                 // [plane]
    const uint32_t sQG = base + 4 * TB;                           // [stage][Q hi, Q lo, dO hi, dO lo]
    const uint32_t sDS = sQG + QST * 4 * TB;                      // [plane][chunk][128 keys][128 B]
    const uint32_t bars = sDS + C::DS_BYTES;
    const uint32_t kv_full = bars, qg_full0 = bars + 8, qg_empty0 = bars + 24, s_full = bars + 40, pds_full = bars + 48;
    const uint32_t dq_full = bars + 56, dq_free = bars + 64, dkv_full = bars + 72, k_ready = bars + 80;
    const uint32_t aux0 = bars + 256;                             // float/u32 [2 buffers][3][128]
    const uint32_t tmem_slot = aux0 + C::AUX_BYTES;
    uint8_t* smem_gen = abt_smem_raw + (base - smem_u32(abt_smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(abt_smem_raw + (tmem_slot - smem_u32(abt_smem_raw)));
    float* aux = reinterpret_cast<float*>(abt_smem_raw + (aux0 - smem_u32(abt_smem_raw)));
    uint8_t* ds_gen = smem_gen + (sDS - base);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * BT, h = blockIdx.y, b = blockIdx.z;
    const int D = p.H * DH, ld = 3 * D, NH = 3 * p.H;
    const int len = min((int)p.lens[b], p.S);
    if (k0 >= len) return;   // dK = dV = 0 for padded keys (dqkv is zero-initialised)

    if (threadIdx.x == 0) {
        mbar_init(kv_full, 1);
        for (int i = 0; i < QST; ++i) { mbar_init(qg_full0 + 8 * i, 1); mbar_init(qg_empty0 + 8 * i, 1); }
        mbar_init(s_full, 1); mbar_init(pds_full, ABT_SOFTMAX_WARPS);
        mbar_init(dq_full, 1); mbar_init(dq_free, ABT_SOFTMAX_WARPS); mbar_init(dkv_full, 1); mbar_init(k_ready, ABT_SOFTMAX_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(ABT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int n_q = (len + BT - 1) / BT;
    const int sl_q = b * NH + h, sl_k = b * NH + p.H + h, sl_v = b * NH + 2 * p.H + h, sl_lo = p.B * NH;
    const int sl_g = b * p.H + h, sl_glo = p.B * p.H;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(kv_full, 4 * TB);
            tma_load_3d(sK, &map_r, kv_full, 0, k0, sl_k);
            tma_load_3d(sK + TB, &map_r, kv_full, 0, k0, sl_lo + sl_k);
            tma_load_3d(sV, &map_r, kv_full, 0, k0, sl_v);
            tma_load_3d(sV + TB, &map_r, kv_full, 0, k0, sl_lo + sl_v);
            int st = 0, ph = 0;
            for (int i = 0; i < n_q; ++i) {
                mbar_wait(qg_empty0 + 8 * st, ph ^ 1);
                const uint32_t s0 = sQG + st * 4 * TB, bar = qg_full0 + 8 * st;
                mbar_expect_tx(bar, 4 * TB);
                tma_load_3d(s0, &map_r, bar, 0, i * BT, sl_q);
                tma_load_3d(s0 + TB, &map_r, bar, 0, i * BT, sl_lo + sl_q);
                tma_load_3d(s0 + 2 * TB, &map_g, bar, 0, i * BT, sl_g);
                tma_load_3d(s0 + 3 * TB, &map_g, bar, 0, i * BT, sl_glo + sl_g);
                if (++st == QST) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // whole warp in the loop, one elected lane issues (descriptors stay in uniform registers)
        constexpr uint32_t idesc_s = idesc_bf16(BT, BT, 0, 0);     // S^T, dP^T: A, B K-major, N = 128 queries
        constexpr uint32_t idesc_g = idesc_bf16(BT, DH, 0, 1);     // dV, dK: A in TMEM, B MN-major, N = dh
        constexpr uint32_t idesc_q = idesc_bf16(BT, DH, 1, 1);     // dQ: A (dS^T staging) MN-major, B MN-major, N = dh
        const uint64_t kh = bdesc_k<DH>(sK), kl = bdesc_k<DH>(sK + TB), vh = bdesc_k<DH>(sV), vl = bdesc_k<DH>(sV + TB);
        const uint64_t kmh = bdesc_mn<DH>(sK), kml = bdesc_mn<DH>(sK + TB);
        const uint64_t dsh = bdesc_ds(sDS), dsl = bdesc_ds(sDS + 2 * 128 * 128);
        const uint32_t t_st = tmem_base, t_dpt = tmem_base + BT, t_dk = tmem_base + C::COL_DK, t_dv = tmem_base + C::COL_DV,
                       t_dq = tmem_base + C::COL_DQ;
        mbar_wait(k_ready, 0);   // K, V landed and the K rows of padded keys are zeroed
        int st = 0, ph = 0;
        for (int i = 0; i < n_q; ++i) {
            mbar_wait(qg_full0 + 8 * st, ph);
            tc_fence_after();
            const uint32_t s0 = sQG + st * 4 * TB;
            const uint64_t qh = bdesc_k<DH>(s0), ql = bdesc_k<DH>(s0 + TB), gh = bdesc_k<DH>(s0 + 2 * TB), gl = bdesc_k<DH>(s0 + 3 * TB);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < C::KS; ++kk) {   // S^T = K Q^T
                    umma_ss(t_st, kh + 2 * kk, qh + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);
                    umma_ss(t_st, kl + 2 * kk, qh + 2 * kk, idesc_s, 1u);
                    umma_ss(t_st, kh + 2 * kk, ql + 2 * kk, idesc_s, 1u);
                }
#pragma unroll
                for (int kk = 0; kk < C::KS; ++kk) {   // dP^T = V dO^T
                    umma_ss(t_dpt, vh + 2 * kk, gh + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);
                    umma_ss(t_dpt, vl + 2 * kk, gh + 2 * kk, idesc_s, 1u);
                    umma_ss(t_dpt, vh + 2 * kk, gl + 2 * kk, idesc_s, 1u);
                }
                umma_commit(s_full);
            }
            __syncwarp();
            mbar_wait(pds_full, i & 1);                         // P^T, dS^T in TMEM, dS^T staged in shared memory
            if (i > 0) mbar_wait(dq_free, (i - 1) & 1);         // the previous tile's dQ has been drained
            tc_fence_after();
            const uint64_t qmh = bdesc_mn<DH>(s0), qml = bdesc_mn<DH>(s0 + TB), gmh = bdesc_mn<DH>(s0 + 2 * TB), gml = bdesc_mn<DH>(s0 + 3 * TB);
            const uint32_t acc0 = i > 0 ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < BT / 16; ++kk) {   // K-steps over the 128 queries (dV, dK) / 128 keys (dQ)
                    const uint32_t ah = 32 * (kk >> 1) + 8 * (kk & 1), al = ah + 16;   // in-place hi / lo pair columns
                    umma_ts(t_dv, t_st + ah, gmh + C::ROW_KSTEP * kk, idesc_g, kk > 0 ? 1u : acc0);
                    umma_ts(t_dv, t_st + al, gmh + C::ROW_KSTEP * kk, idesc_g, 1u);
                    umma_ts(t_dv, t_st + ah, gml + C::ROW_KSTEP * kk, idesc_g, 1u);
                }
#pragma unroll
                for (int kk = 0; kk < BT / 16; ++kk) {
                    const uint32_t ah = 32 * (kk >> 1) + 8 * (kk & 1), al = ah + 16;
                    umma_ts(t_dk, t_dpt + ah, qmh + C::ROW_KSTEP * kk, idesc_g, kk > 0 ? 1u : acc0);
                    umma_ts(t_dk, t_dpt + al, qmh + C::ROW_KSTEP * kk, idesc_g, 1u);
                    umma_ts(t_dk, t_dpt + ah, qml + C::ROW_KSTEP * kk, idesc_g, 1u);
                }
                umma_commit(qg_empty0 + 8 * st);   // Q / dO are free once dV and dK retire: the next tile's loads overlap dQ and its drain
#pragma unroll
                for (int kk = 0; kk < BT / 16; ++kk) {   // dQ = dS K: A K-step = 16 key rows of 128 B in the staging tile
                    umma_ss(t_dq, dsh + 128 * kk, kmh + C::ROW_KSTEP * kk, idesc_q, kk > 0 ? 1u : 0u);
                    umma_ss(t_dq, dsl + 128 * kk, kmh + C::ROW_KSTEP * kk, idesc_q, 1u);
                    umma_ss(t_dq, dsh + 128 * kk, kml + C::ROW_KSTEP * kk, idesc_q, 1u);
                }
                umma_commit(dq_full);
                if (i == n_q - 1) umma_commit(dkv_full);
            }
            __syncwarp();
            if (++st == QST) { st = 0; ph ^= 1; }
        }
    } else {
        const int quad = warp & 3, cq = (warp - 2) >> 2;
        const int rl = quad * 32 + lane;                         // key row within the tile (S^T / dP^T lanes), query row for dQ
        const int key = k0 + rl;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int st_tid = threadIdx.x - 64;                      // 0..511
        const float* lse = p.lse + ((size_t)b * p.H + h) * p.S;
        const float* delta = p.delta + ((size_t)b * p.H + h) * p.S;
        float* dbase = p.dqkv + (size_t)b * p.S * ld + h * DH;
        const unsigned long long seed = dyn_seed(p.seed, p.dyn);
        const bool drop = p.dropout_p > 0.f;
        const float inv_keep = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
        const uint32_t thresh = drop_threshold(p.dropout_p);
        const unsigned long long bh = ((unsigned long long)b * p.H + h) * (unsigned long long)p.S;
        const uint32_t kc = drop_col_term((uint32_t)key);
        const float scale = rsqrtf((float)DH);
        const bool key_ok = key < len;
        const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;

        // padded keys (rows >= len of this utterance, only in its last key tile): zero their K rows (both planes) once, so that
        // whatever their dS^T rows hold adds nothing to dQ = dS K; generic-proxy writes -> proxy fence -> the MMA warp's k_ready
        mbar_wait(kv_full, 0);
        if (!key_ok && cq == 0) {
#pragma unroll
            for (int c = 0; c < C::ROW_BYTES; c += 16) {
                *reinterpret_cast<uint4*>(smem_gen + (sK - base) + (size_t)rl * C::ROW_BYTES + c) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(smem_gen + (sK - base) + C::TB + (size_t)rl * C::ROW_BYTES + c) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(k_ready);

        const float keep_prob = drop ? 1.f - p.dropout_p : 1.f;
        // per-query vectors of a tile (lse * log2e, delta * keep_prob, dropout row key): written one tile ahead into the other buffer
        // the global loads are issued one tile ahead into registers (fetch) and only written to shared memory after this tile's
        // arithmetic (store): their latency never sits in front of the score loads
        float n_lse = 0.f, n_delta = 0.f;
        auto fetch_aux = [&](int i) {
            if (st_tid < BT) {
                const int q = i * BT + st_tid;
                n_lse = q < len ? lse[q] : 0.f;
                n_delta = q < len ? delta[q] : 0.f;
            }
        };
        auto store_aux = [&](int i) {
            if (st_tid < BT) {
                float* ax = aux + (i & 1) * 3 * BT;
                // padded queries: lse = +inf makes P = exp2(S - inf) = 0 (and with it dS) without any per-score select
                ax[st_tid] = i * BT + st_tid < len ? n_lse * kLog2eB : INFINITY;
                ax[BT + st_tid] = n_delta * keep_prob;
                reinterpret_cast<uint32_t*>(ax)[2 * BT + st_tid] = drop ? hash_u32(seed, bh + i * BT + st_tid) : 0u;
            }
        };
        fetch_aux(0);
        store_aux(0);
        for (int i = 0; i < n_q; ++i) {
            const int q0 = i * BT;
            const float* ax = aux + (i & 1) * 3 * BT;             // [lse*log2e | delta*keep | row key]
            asm volatile("bar.sync 1, 512;" ::: "memory");        // this tile's vectors are visible; the other buffer is free
            if (i + 1 < n_q) fetch_aux(i + 1);
            mbar_wait(s_full, i & 1);
            if (tr && i < 256) p.trace[i] = clock64();
            tc_fence_after();
            uint32_t s[32], g[32];
            tmem_ld32(t_lane + 32 * cq, s);
            tmem_ld32(t_lane + BT + 32 * cq, g);
            tmem_ld_wait32(s);
            tmem_ld_wait32(g);
            // in place: s <- P^T pairs (hi | lo), g <- dS^T pairs (hi | lo); both WITHOUT the 1/keep factor (dK, dV, dQ are scaled once)
            // no per-score masking: padded queries carry lse = +inf (P = 0), padded keys have their K rows zeroed in shared memory
            // (their dS^T rows then add nothing to dQ, and their dK / dV rows are never stored)
            uint32_t pt[32], dst[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int c = 32 * cq + 2 * j;                    // query column within the tile
                const float2 ls = *reinterpret_cast<const float2*>(ax + c);
                const float2 dl = *reinterpret_cast<const float2*>(ax + BT + c);
                const float p0 = ex2(fmaf(__uint_as_float(s[2 * j]), kLog2eB, -ls.x));
                const float p1 = ex2(fmaf(__uint_as_float(s[2 * j + 1]), kLog2eB, -ls.y));
                float d0 = __uint_as_float(g[2 * j]), d1 = __uint_as_float(g[2 * j + 1]);
                float pd0 = p0, pd1 = p1;
                if (drop) {
                    const uint2 rk = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(ax) + 2 * BT + c);
                    const bool k0_ = drop_keep(rk.x, kc, thresh), k1_ = drop_keep(rk.y, kc, thresh);
                    pd0 = k0_ ? p0 : 0.f; pd1 = k1_ ? p1 : 0.f;
                    d0 = k0_ ? d0 : 0.f; d1 = k1_ ? d1 : 0.f;
                }
                split_pair(pd0, pd1, pt[j], pt[16 + j]);                                  // P^T (dropped)
                split_pair(p0 * (d0 - dl.x), p1 * (d1 - dl.y), dst[j], dst[16 + j]);      // dS^T * keep_prob
            }
            tmem_st32(t_lane + 32 * cq, pt);
            tmem_st32(t_lane + BT + 32 * cq, dst);
            // dS^T staging for dQ: row = key, 32 queries = 4 x 16 B per plane at units (cq & 1) * 4 + u of query chunk cq >> 1
            {
                uint8_t* rowp = ds_gen + (size_t)(cq >> 1) * 16384 + (size_t)rl * 128;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int unit = ((cq & 1) * 4 + u) ^ (rl & 7);
                    *reinterpret_cast<uint4*>(rowp + unit * 16) = make_uint4(dst[4 * u], dst[4 * u + 1], dst[4 * u + 2], dst[4 * u + 3]);
                    *reinterpret_cast<uint4*>(rowp + 32768 + unit * 16) = make_uint4(dst[16 + 4 * u], dst[16 + 4 * u + 1], dst[16 + 4 * u + 2], dst[16 + 4 * u + 3]);
                }
            }
            tmem_st_wait();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pds_full);
            if (i + 1 < n_q) store_aux(i + 1);
            if (tr && i < 256) p.trace[256 + i] = clock64();
            // ---- drain dQ of this query tile: TMEM lane = query row, vector fp32 reductions (q carries 1/sqrt(dh)) ----
            mbar_wait(dq_full, i & 1);
            if (tr && i < 256) p.trace[512 + i] = clock64();
            tc_fence_after();
            if (DH == 64 || cq == 0) {
                const int c0 = DH == 64 ? cq * 16 : 0;
                uint32_t v[16];
                tmem_ld16(t_lane + C::COL_DQ + c0, v);
                tmem_ld_wait16(v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dq_free);              // the accumulator is back with the MMA warp before the reductions start
                const int q = q0 + rl;
                if (q < len) {
                    float* dq = dbase + (size_t)q * ld + c0;
                    const float sc = scale * inv_keep;
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dq + e), "f"(sc * __uint_as_float(v[e])),
                                     "f"(sc * __uint_as_float(v[e + 1])), "f"(sc * __uint_as_float(v[e + 2])), "f"(sc * __uint_as_float(v[e + 3]))
                                     : "memory");
                }
            } else {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dq_free);
            }
            if (tr && i < 256) p.trace[768 + i] = clock64();
        }
        // ---- dK, dV of this key tile ----
        mbar_wait(dkv_full, 0);
        tc_fence_after();
        if (DH == 64 || cq == 0) {
            const int c0 = DH == 64 ? cq * 16 : 0;
            uint32_t vk[16], vv[16];
            tmem_ld16(t_lane + C::COL_DK + c0, vk);
            tmem_ld16(t_lane + C::COL_DV + c0, vv);
            tmem_ld_wait16(vk);
            tmem_ld_wait16(vv);
            if (key_ok) {
                float* dk = dbase + (size_t)key * ld + D + c0;
                float* dv = dbase + (size_t)key * ld + 2 * D + c0;
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    *reinterpret_cast<float4*>(dk + e) = make_float4(__uint_as_float(vk[e]) * inv_keep, __uint_as_float(vk[e + 1]) * inv_keep,
                                                                    __uint_as_float(vk[e + 2]) * inv_keep, __uint_as_float(vk[e + 3]) * inv_keep);
                    *reinterpret_cast<float4*>(dv + e) = make_float4(__uint_as_float(vv[e]) * inv_keep, __uint_as_float(vv[e + 1]) * inv_keep,
                                                                    __uint_as_float(vv[e + 2]) * inv_keep, __uint_as_float(vv[e + 3]) * inv_keep);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ABT_TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// Pipelined variant (head_dim 16): the query tile is processed as two 64-query halves through TWO S^T / dP^T buffer pairs in
// TMEM, so the five products of half g and the loads/handshakes around them run while the 16 element-wise warps work on half
// g + 1 (the serial kernel above pays ~3 barrier round trips per tile, which for dh = 16 is as long as the arithmetic).
// dQ is still one M = 128 product per query tile (its A operand = both staged halves), drained one half-step later.
// TMEM columns: buffer u < 2: S^T / P^T [128u, 128u + 64), dP^T / dS^T [128u + 64, 128u + 128); dK 256.., dV 320.., dQ 384...
// =====================================================================================================================
template <int DH>
__global__ void __launch_bounds__(ABT_THREADS, 1) attn_bwd_tc_pipe_kernel(const __grid_constant__ CUtensorMap map_r,
                                                                          const __grid_constant__ CUtensorMap map_g, AbtParams p) {
    using C = AbtCfg<DH>;
    static_assert(DH == 16, "pipelined backward: shared-memory budget (two Q / dO stages + dS^T staging) is laid out for head_dim 16");
    // (a third buffer pair at TMEM columns [256, 384) with the accumulators moved to 384 / 416 / 448 gave wrong products on B200 and
    // only 6 % more speed: two buffers already cover the handshake latency)
    constexpr int TB = C::TB, QST = 2, HQ = 64, NBUF = 2;
    constexpr int COL_DK = C::COL_DK, COL_DV = C::COL_DV, COL_DQ = C::COL_DQ;
    constexpr uint32_t HALF_BYTES = HQ * C::ROW_BYTES;           // 64 rows of a Q / dO tile
    extern __shared__ uint8_t abt_smem_raw[];
    const uint32_t base = (smem_u32(abt_smem_raw) + 1023u) & ~1023u;
    const uint32_t sK = base, sV = base + 2 * TB;
    const uint32_t sQG = base + 4 * TB;                           // [stage][Q hi, Q lo, dO hi, dO lo]
    const uint32_t sDS = sQG + QST * 4 * TB;                      // [plane][64-query chunk][128 keys][128 B]
    const uint32_t bars = sDS + C::DS_BUFS * C::DS_BYTES;     // dS^T staging is double buffered by query tile parity
    const uint32_t kv_full = bars, qg_full0 = bars + 8, qg_empty0 = bars + 32, s_full0 = bars + 56, pds_full0 = bars + 80;
    const uint32_t dq_full = bars + 104, dq_free = bars + 112, dkv_full = bars + 120, k_ready = bars + 128;
    const uint32_t aux0 = bars + 256;
    const uint32_t tmem_slot = aux0 + C::AUX_BYTES;
    uint8_t* smem_gen = abt_smem_raw + (base - smem_u32(abt_smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(abt_smem_raw + (tmem_slot - smem_u32(abt_smem_raw)));
    float* aux = reinterpret_cast<float*>(abt_smem_raw + (aux0 - smem_u32(abt_smem_raw)));
    uint8_t* ds_gen = smem_gen + (sDS - base);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * BT, h = blockIdx.y, b = blockIdx.z;
    const int D = p.H * DH, ld = 3 * D, NH = 3 * p.H;
    const int len = min((int)p.lens[b], p.S);
    if (k0 >= len) return;

    if (threadIdx.x == 0) {
        mbar_init(kv_full, 1);
        for (int i = 0; i < QST; ++i) { mbar_init(qg_full0 + 8 * i, 1); mbar_init(qg_empty0 + 8 * i, 1); }
        for (int i = 0; i < NBUF; ++i) { mbar_init(s_full0 + 8 * i, 1); mbar_init(pds_full0 + 8 * i, ABT_SOFTMAX_WARPS); }
        mbar_init(dq_full, 1); mbar_init(dq_free, ABT_SOFTMAX_WARPS); mbar_init(dkv_full, 1); mbar_init(k_ready, ABT_SOFTMAX_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(ABT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int n_q = (len + BT - 1) / BT, n_g = 2 * n_q;
    const int sl_q = b * NH + h, sl_k = b * NH + p.H + h, sl_v = b * NH + 2 * p.H + h, sl_lo = p.B * NH;
    const int sl_g = b * p.H + h, sl_glo = p.B * p.H;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(kv_full, 4 * TB);
            tma_load_3d(sK, &map_r, kv_full, 0, k0, sl_k);
            tma_load_3d(sK + TB, &map_r, kv_full, 0, k0, sl_lo + sl_k);
            tma_load_3d(sV, &map_r, kv_full, 0, k0, sl_v);
            tma_load_3d(sV + TB, &map_r, kv_full, 0, k0, sl_lo + sl_v);
            int st = 0, ph = 0;
            for (int i = 0; i < n_q; ++i) {
                mbar_wait(qg_empty0 + 8 * st, ph ^ 1);
                const uint32_t s0 = sQG + st * 4 * TB, bar = qg_full0 + 8 * st;
                mbar_expect_tx(bar, 4 * TB);
                tma_load_3d(s0, &map_r, bar, 0, i * BT, sl_q);
                tma_load_3d(s0 + TB, &map_r, bar, 0, i * BT, sl_lo + sl_q);
                tma_load_3d(s0 + 2 * TB, &map_g, bar, 0, i * BT, sl_g);
                tma_load_3d(s0 + 3 * TB, &map_g, bar, 0, i * BT, sl_glo + sl_g);
                if (++st == QST) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_s = idesc_bf16(BT, HQ, 0, 0);     // S^T, dP^T halves: N = 64 queries
        constexpr uint32_t idesc_g = idesc_bf16(BT, DH, 0, 1);     // dV, dK: A in TMEM, B MN-major
        constexpr uint32_t idesc_q = idesc_bf16(BT, DH, 1, 1);     // dQ: A (staging) MN-major, B MN-major
        const uint64_t kh = bdesc_k<DH>(sK), kl = bdesc_k<DH>(sK + TB), vh = bdesc_k<DH>(sV), vl = bdesc_k<DH>(sV + TB);
        const uint64_t kmh = bdesc_mn<DH>(sK), kml = bdesc_mn<DH>(sK + TB);
        const uint32_t t_dk = tmem_base + COL_DK, t_dv = tmem_base + COL_DV, t_dq = tmem_base + COL_DQ;
        mbar_wait(k_ready, 0);   // K, V landed and the K rows of padded keys are zeroed
        // S^T_g = K Q_h^T, dP^T_g = V dO_h^T into buffer g & 1 (rows [64 hh, 64 hh + 64) of query tile g >> 1)
        auto issue_sdp = [&](int g) {
            const int i = g >> 1, hh = g & 1, stg = i % QST, u = g % NBUF;
            if (hh == 0) mbar_wait(qg_full0 + 8 * stg, (i / QST) & 1);
            tc_fence_after();
            const uint32_t s0 = sQG + stg * 4 * TB + hh * HALF_BYTES;
            const uint64_t qh = bdesc_k<DH>(s0), ql = bdesc_k<DH>(s0 + TB), gh = bdesc_k<DH>(s0 + 2 * TB), gl = bdesc_k<DH>(s0 + 3 * TB);
            const uint32_t t_st = tmem_base + u * 128, t_dpt = t_st + 64;
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < C::KS; ++kk) {
                    umma_ss(t_st, kh + 2 * kk, qh + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);
                    umma_ss(t_st, kl + 2 * kk, qh + 2 * kk, idesc_s, 1u);
                    umma_ss(t_st, kh + 2 * kk, ql + 2 * kk, idesc_s, 1u);
                }
#pragma unroll
                for (int kk = 0; kk < C::KS; ++kk) {
                    umma_ss(t_dpt, vh + 2 * kk, gh + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);
                    umma_ss(t_dpt, vl + 2 * kk, gh + 2 * kk, idesc_s, 1u);
                    umma_ss(t_dpt, vh + 2 * kk, gl + 2 * kk, idesc_s, 1u);
                }
                umma_commit(s_full0 + 8 * u);
            }
            __syncwarp();
        };
        issue_sdp(0);
        issue_sdp(1);
        if (NBUF > 2 && n_g > 2) issue_sdp(2);
        for (int g = 0; g < n_g; ++g) {
            const int i = g >> 1, hh = g & 1, stg = i % QST, u = g % NBUF;
            mbar_wait(pds_full0 + 8 * u, (g / NBUF) & 1);           // P^T_g, dS^T_g in TMEM (+ dS^T_g staged)
            if (hh == 1 && i > 0) mbar_wait(dq_free, (i - 1) & 1);  // the previous tile's dQ has been drained
            tc_fence_after();
            const uint32_t s0 = sQG + stg * 4 * TB + hh * HALF_BYTES;
            const uint64_t qmh = bdesc_mn<DH>(s0), qml = bdesc_mn<DH>(s0 + TB), gmh = bdesc_mn<DH>(s0 + 2 * TB), gml = bdesc_mn<DH>(s0 + 3 * TB);
            const uint32_t t_st = tmem_base + u * 128, t_dpt = t_st + 64;
            const uint32_t acc0 = g > 0 ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < HQ / 16; ++kk) {   // K-steps over the 64 queries of this half: pairs (hi | lo) of column group kk
                    umma_ts(t_dv, t_st + 16 * kk, gmh + C::ROW_KSTEP * kk, idesc_g, kk > 0 ? 1u : acc0);
                    umma_ts(t_dv, t_st + 16 * kk + 8, gmh + C::ROW_KSTEP * kk, idesc_g, 1u);
                    umma_ts(t_dv, t_st + 16 * kk, gml + C::ROW_KSTEP * kk, idesc_g, 1u);
                }
#pragma unroll
                for (int kk = 0; kk < HQ / 16; ++kk) {
                    umma_ts(t_dk, t_dpt + 16 * kk, qmh + C::ROW_KSTEP * kk, idesc_g, kk > 0 ? 1u : acc0);
                    umma_ts(t_dk, t_dpt + 16 * kk + 8, qmh + C::ROW_KSTEP * kk, idesc_g, 1u);
                    umma_ts(t_dk, t_dpt + 16 * kk, qml + C::ROW_KSTEP * kk, idesc_g, 1u);
                }
                if (hh == 1) {
                    umma_commit(qg_empty0 + 8 * stg);   // Q / dO of this tile are free once dV and dK retire
                    const uint64_t dsh = bdesc_ds(sDS + (i & 1) * C::DS_BYTES), dsl = bdesc_ds(sDS + (i & 1) * C::DS_BYTES + 2 * 128 * 128);
#pragma unroll
                    for (int kk = 0; kk < BT / 16; ++kk) {   // dQ = dS K over the 128 keys, both staged halves as M = 128
                        umma_ss(t_dq, dsh + 128 * kk, kmh + C::ROW_KSTEP * kk, idesc_q, kk > 0 ? 1u : 0u);
                        umma_ss(t_dq, dsl + 128 * kk, kmh + C::ROW_KSTEP * kk, idesc_q, 1u);
                        umma_ss(t_dq, dsh + 128 * kk, kml + C::ROW_KSTEP * kk, idesc_q, 1u);
                    }
                    umma_commit(dq_full);
                    if (g == n_g - 1) umma_commit(dkv_full);
                }
            }
            __syncwarp();
            if (g + NBUF < n_g) issue_sdp(g + NBUF);   // into the buffer this half just released (the tensor pipe is in order)
        }
    } else {
        const int quad = warp & 3, cq = (warp - 2) >> 2;
        const int rl = quad * 32 + lane;
        const int key = k0 + rl;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int st_tid = threadIdx.x - 64;
        const float* lse = p.lse + ((size_t)b * p.H + h) * p.S;
        const float* delta = p.delta + ((size_t)b * p.H + h) * p.S;
        float* dbase = p.dqkv + (size_t)b * p.S * ld + h * DH;
        const unsigned long long seed = dyn_seed(p.seed, p.dyn);
        const bool drop = p.dropout_p > 0.f;
        const float inv_keep = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
        const float keep_prob = drop ? 1.f - p.dropout_p : 1.f;
        const uint32_t thresh = drop_threshold(p.dropout_p);
        const unsigned long long bh = ((unsigned long long)b * p.H + h) * (unsigned long long)p.S;
        const uint32_t kc = drop_col_term((uint32_t)key);
        const float scale = rsqrtf((float)DH);
        const bool key_ok = key < len;

        // padded keys (rows >= len of this utterance, only in its last key tile): zero their K rows (both planes) once, so that
        // whatever their dS^T rows hold adds nothing to dQ = dS K; generic-proxy writes -> proxy fence -> the MMA warp's k_ready
        mbar_wait(kv_full, 0);
        if (!key_ok && cq == 0) {
#pragma unroll
            for (int c = 0; c < C::ROW_BYTES; c += 16) {
                *reinterpret_cast<uint4*>(smem_gen + (sK - base) + (size_t)rl * C::ROW_BYTES + c) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(smem_gen + (sK - base) + C::TB + (size_t)rl * C::ROW_BYTES + c) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(k_ready);
        float n_lse = 0.f, n_delta = 0.f;   // next tile's per-query values: fetched at half 0, stored to smem after half 1
        auto fetch_aux = [&](int i) {
            if (st_tid < BT) {
                const int q = i * BT + st_tid;
                n_lse = q < len ? lse[q] : 0.f;
                n_delta = q < len ? delta[q] : 0.f;
            }
        };
        auto store_aux = [&](int i) {
            if (st_tid < BT) {
                float* ax = aux + (i & 1) * 3 * BT;
                // padded queries: lse = +inf makes P = exp2(S - inf) = 0 (and with it dS) without any per-score select
                ax[st_tid] = i * BT + st_tid < len ? n_lse * kLog2eB : INFINITY;
                ax[BT + st_tid] = n_delta * keep_prob;
                reinterpret_cast<uint32_t*>(ax)[2 * BT + st_tid] = drop ? hash_u32(seed, bh + i * BT + st_tid) : 0u;
            }
        };
        // dQ of query tile i: TMEM lane = query row; vector fp32 reductions (q carries 1/sqrt(dh)); hands the accumulator back first
        auto drain_dq = [&](int i, bool release) {
            mbar_wait(dq_full, i & 1);   // (already observed by the staging guard except for the last tile)
            tc_fence_after();
            if (DH == 64 || cq == 0) {
                const int c0 = DH == 64 ? cq * 16 : 0;
                uint32_t v[16];
                tmem_ld16(t_lane + COL_DQ + c0, v);
                tmem_ld_wait16(v);
                tc_fence_before();
                __syncwarp();
                if (release && lane == 0) mbar_arrive(dq_free);
                const int q = i * BT + rl;
                if (q < len) {
                    float* dq = dbase + (size_t)q * ld + c0;
                    const float sc = scale * inv_keep;
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dq + e), "f"(sc * __uint_as_float(v[e])),
                                     "f"(sc * __uint_as_float(v[e + 1])), "f"(sc * __uint_as_float(v[e + 2])), "f"(sc * __uint_as_float(v[e + 3]))
                                     : "memory");
                }
            } else {
                tc_fence_before();
                __syncwarp();
                if (release && lane == 0) mbar_arrive(dq_free);
            }
        };
        const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
        fetch_aux(0);
        store_aux(0);
        for (int g = 0; g < n_g; ++g) {
            const int i = g >> 1, hh = g & 1, q0 = i * BT;
            if (tr && g < 256) p.trace[g] = clock64();                       // half-step begins
            const float* ax = aux + (i & 1) * 3 * BT;
            if (hh == 0) {
                asm volatile("bar.sync 1, 512;" ::: "memory");
                if (i + 1 < n_q) fetch_aux(i + 1);
            }
            const int u = g % NBUF;
            mbar_wait(s_full0 + 8 * u, (g / NBUF) & 1);
            if (tr && g < 256) p.trace[256 + g] = clock64();                 // S^T / dP^T of this half available
            tc_fence_after();
            const uint32_t t_st = t_lane + u * 128 + 16 * cq, t_dpt = t_st + 64;
            uint32_t s[16], gq[16];
            tmem_ld16(t_st, s);
            tmem_ld16(t_dpt, gq);
            tmem_ld_wait16(s);
            tmem_ld_wait16(gq);
            const int cbase = 64 * hh + 16 * cq;                   // first query column of this thread within the tile
            uint32_t pt[16], dst[16];                              // 8 hi pairs | 8 lo pairs
            // no per-score masking: padded queries carry lse = +inf (P = 0), padded keys have their K rows zeroed in shared memory
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = cbase + 2 * j;
                const float2 ls = *reinterpret_cast<const float2*>(ax + c);
                const float2 dl = *reinterpret_cast<const float2*>(ax + BT + c);
                const float p0 = ex2(fmaf(__uint_as_float(s[2 * j]), kLog2eB, -ls.x));
                const float p1 = ex2(fmaf(__uint_as_float(s[2 * j + 1]), kLog2eB, -ls.y));
                float d0 = __uint_as_float(gq[2 * j]), d1 = __uint_as_float(gq[2 * j + 1]);
                float pd0 = p0, pd1 = p1;
                if (drop) {
                    const uint2 rk = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(ax) + 2 * BT + c);
                    const bool k0_ = drop_keep(rk.x, kc, thresh), k1_ = drop_keep(rk.y, kc, thresh);
                    pd0 = k0_ ? p0 : 0.f; pd1 = k1_ ? p1 : 0.f;
                    d0 = k0_ ? d0 : 0.f; d1 = k1_ ? d1 : 0.f;
                }
                split_pair(pd0, pd1, pt[j], pt[8 + j]);
                split_pair(p0 * (d0 - dl.x), p1 * (d1 - dl.y), dst[j], dst[8 + j]);
            }
            tmem_st16(t_st, pt);
            tmem_st16(t_dpt, dst);
            if (tr && g < 256) p.trace[512 + g] = clock64();                 // arithmetic done
            {   // dS^T staging (buffer i & 1: its previous reader dQ(i-2) retired before dQ(i-1), whose completion this thread saw
                // when it drained it): row = key, chunk hh (64 queries = 128 B), this thread's 16 queries = units 2 cq, 2 cq + 1
                uint8_t* rowp = ds_gen + (size_t)(i & 1) * C::DS_BYTES + (size_t)hh * 16384 + (size_t)rl * 128;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int unit = (2 * cq + u) ^ (rl & 7);
                    *reinterpret_cast<uint4*>(rowp + unit * 16) = make_uint4(dst[4 * u], dst[4 * u + 1], dst[4 * u + 2], dst[4 * u + 3]);
                    *reinterpret_cast<uint4*>(rowp + 32768 + unit * 16) = make_uint4(dst[8 + 4 * u], dst[8 + 4 * u + 1], dst[8 + 4 * u + 2], dst[8 + 4 * u + 3]);
                }
            }
            tmem_st_wait();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pds_full0 + 8 * u);
            if (tr && g < 256) p.trace[768 + g] = clock64();                 // published
            if (hh == 1 && i + 1 < n_q) store_aux(i + 1);
            if (hh == 1 && i > 0) drain_dq(i - 1, true);           // a whole tile late: dQ(i-1) retired long ago, nothing to wait for
        }
        drain_dq(n_q - 1, false);
        // ---- dK, dV of this key tile ----
        mbar_wait(dkv_full, 0);
        tc_fence_after();
        if (DH == 64 || cq == 0) {
            const int c0 = DH == 64 ? cq * 16 : 0;
            uint32_t vk[16], vv[16];
            tmem_ld16(t_lane + COL_DK + c0, vk);
            tmem_ld16(t_lane + COL_DV + c0, vv);
            tmem_ld_wait16(vk);
            tmem_ld_wait16(vv);
            if (key_ok) {
                float* dk = dbase + (size_t)key * ld + D + c0;
                float* dv = dbase + (size_t)key * ld + 2 * D + c0;
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    *reinterpret_cast<float4*>(dk + e) = make_float4(__uint_as_float(vk[e]) * inv_keep, __uint_as_float(vk[e + 1]) * inv_keep,
                                                                    __uint_as_float(vk[e + 2]) * inv_keep, __uint_as_float(vk[e + 3]) * inv_keep);
                    *reinterpret_cast<float4*>(dv + e) = make_float4(__uint_as_float(vv[e]) * inv_keep, __uint_as_float(vv[e + 1]) * inv_keep,
                                                                    __uint_as_float(vv[e + 2]) * inv_keep, __uint_as_float(vv[e + 3]) * inv_keep);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ABT_TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// Pipelined variant for head_dim 64 (round 2): the same half-tile pipeline as the head_dim 16 kernel above, re-laid-out to fit
// 227 KB of shared memory next to the resident K / V tiles (64 KB):
//   * Q / dO arrive as 64-query HALF tiles (32 KB each: Q hi, Q lo, dO hi, dO lo) through a ring of THREE half-stages (96 KB): the
//     S^T / dP^T products of half g + 2 are issued while half g's dV / dK products still read their half-stage;
//   * ONE dS^T staging buffer (64 KB): before half 0 of tile i + 1 overwrites it the element-wise warps wait for dQ(i) to retire
//     (it was issued a half-step earlier, so the wait is normally over when the arithmetic of that half is done);
//   * the per-query vectors (lse, delta, dropout row keys) are kept per HALF (2 x 64 entries, 1.5 KB) with one 512-thread barrier
//     per half-step.
// The serial kernel (one 128-query tile at a time: ~3.4 k cycles of MMA waits + 2.6 k of dQ drain per 4.4 k of arithmetic) stays as
// DX_ATTN_BWD_PIPE64=0.
// TMEM columns: buffer u < 2: S^T / P^T [128u, 128u + 64), dP^T / dS^T [128u + 64, 128u + 128); dK 256.., dV 320.., dQ 384...
// =====================================================================================================================
constexpr int ABT64_HB = 64 * 128;                       // one plane of a 64-query half tile (head_dim 64: 128-byte rows)
constexpr int ABT64_HSTAGES = 3;
constexpr int ABT64_DS_BYTES = 2 * 2 * 128 * 128;
constexpr int ABT64_AUX_BYTES = 2 * 3 * 64 * 4;
constexpr int ABT64_SMEM_BYTES = 4 * 128 * 128 + ABT64_HSTAGES * 4 * ABT64_HB + ABT64_DS_BYTES + 256 + ABT64_AUX_BYTES + 64 + 1024;
static_assert(ABT64_SMEM_BYTES <= 232448, "pipelined head_dim 64 backward: shared-memory budget");

__global__ void __launch_bounds__(ABT_THREADS, 1) attn_bwd_tc_pipe64_kernel(const __grid_constant__ CUtensorMap map_r,    // K / V: 128-row boxes
                                                                            const __grid_constant__ CUtensorMap map_rq,   // Q: 64-row boxes
                                                                            const __grid_constant__ CUtensorMap map_g,    // dO: 64-row boxes
                                                                            AbtParams p) {
    constexpr int DH = 64;
    using C = AbtCfg<DH>;
    constexpr int TB = C::TB, HB = ABT64_HB, QST = ABT64_HSTAGES, HQ = 64, NBUF = 2;
    constexpr int COL_DK = C::COL_DK, COL_DV = C::COL_DV, COL_DQ = C::COL_DQ;
    extern __shared__ uint8_t abt_smem_raw[];
    const uint32_t base = (smem_u32(abt_smem_raw) + 1023u) & ~1023u;
    const uint32_t sK = base, sV = base + 2 * TB;
    const uint32_t sQG = base + 4 * TB;                           // [half-stage][Q hi, Q lo, dO hi, dO lo] of 64 rows
    const uint32_t sDS = sQG + QST * 4 * HB;                      // [plane][64-query chunk][128 keys][128 B], ONE buffer
    const uint32_t bars = sDS + ABT64_DS_BYTES;
    const uint32_t kv_full = bars, qg_full0 = bars + 8, qg_empty0 = bars + 32, s_full0 = bars + 56, pds_full0 = bars + 80;
    const uint32_t dq_full = bars + 104, dq_free = bars + 112, dkv_full = bars + 120, k_ready = bars + 128;
    const uint32_t aux0 = bars + 256;                             // float/u32 [2 half buffers][3][64]
    const uint32_t tmem_slot = aux0 + ABT64_AUX_BYTES;
    uint8_t* smem_gen = abt_smem_raw + (base - smem_u32(abt_smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(abt_smem_raw + (tmem_slot - smem_u32(abt_smem_raw)));
    float* aux = reinterpret_cast<float*>(abt_smem_raw + (aux0 - smem_u32(abt_smem_raw)));
    uint8_t* ds_gen = smem_gen + (sDS - base);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * BT, h = blockIdx.y, b = blockIdx.z;
    const int D = p.H * DH, ld = 3 * D, NH = 3 * p.H;
    const int len = min((int)p.lens[b], p.S);
    if (k0 >= len) return;

    if (threadIdx.x == 0) {
        mbar_init(kv_full, 1);
        for (int i = 0; i < QST; ++i) { mbar_init(qg_full0 + 8 * i, 1); mbar_init(qg_empty0 + 8 * i, 1); }
        for (int i = 0; i < NBUF; ++i) { mbar_init(s_full0 + 8 * i, 1); mbar_init(pds_full0 + 8 * i, ABT_SOFTMAX_WARPS); }
        mbar_init(dq_full, 1); mbar_init(dq_free, ABT_SOFTMAX_WARPS); mbar_init(dkv_full, 1); mbar_init(k_ready, ABT_SOFTMAX_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(ABT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int n_q = (len + BT - 1) / BT, n_g = 2 * n_q;
    const int sl_q = b * NH + h, sl_k = b * NH + p.H + h, sl_v = b * NH + 2 * p.H + h, sl_lo = p.B * NH;
    const int sl_g = b * p.H + h, sl_glo = p.B * p.H;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(kv_full, 4 * TB);
            tma_load_3d(sK, &map_r, kv_full, 0, k0, sl_k);
            tma_load_3d(sK + TB, &map_r, kv_full, 0, k0, sl_lo + sl_k);
            tma_load_3d(sV, &map_r, kv_full, 0, k0, sl_v);
            tma_load_3d(sV + TB, &map_r, kv_full, 0, k0, sl_lo + sl_v);
            int st = 0, ph = 0;
            for (int g = 0; g < n_g; ++g) {   // one 64-query half tile per ring slot
                mbar_wait(qg_empty0 + 8 * st, ph ^ 1);
                const uint32_t s0 = sQG + st * 4 * HB, bar = qg_full0 + 8 * st;
                mbar_expect_tx(bar, 4 * HB);
                tma_load_3d(s0, &map_rq, bar, 0, g * HQ, sl_q);
                tma_load_3d(s0 + HB, &map_rq, bar, 0, g * HQ, sl_lo + sl_q);
                tma_load_3d(s0 + 2 * HB, &map_g, bar, 0, g * HQ, sl_g);
                tma_load_3d(s0 + 3 * HB, &map_g, bar, 0, g * HQ, sl_glo + sl_g);
                if (++st == QST) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_s = idesc_bf16(BT, HQ, 0, 0);     // S^T, dP^T halves: N = 64 queries
        constexpr uint32_t idesc_g = idesc_bf16(BT, DH, 0, 1);     // dV, dK: A in TMEM, B MN-major
        constexpr uint32_t idesc_q = idesc_bf16(BT, DH, 1, 1);     // dQ: A (staging) MN-major, B MN-major
        const uint64_t kh = bdesc_k<DH>(sK), kl = bdesc_k<DH>(sK + TB), vh = bdesc_k<DH>(sV), vl = bdesc_k<DH>(sV + TB);
        const uint64_t kmh = bdesc_mn<DH>(sK), kml = bdesc_mn<DH>(sK + TB);
        const uint32_t t_dk = tmem_base + COL_DK, t_dv = tmem_base + COL_DV, t_dq = tmem_base + COL_DQ;
        mbar_wait(k_ready, 0);   // K, V landed and the K rows of padded keys are zeroed
        // S^T_g = K Q_h^T, dP^T_g = V dO_h^T into buffer g & 1 (rows [64 hh, 64 hh + 64) of query tile g >> 1)
        auto issue_sdp = [&](int g) {
            const int stg = g % QST, u = g % NBUF;
            mbar_wait(qg_full0 + 8 * stg, (g / QST) & 1);
            tc_fence_after();
            const uint32_t s0 = sQG + stg * 4 * HB;
            const uint64_t qh = bdesc_k<DH>(s0), ql = bdesc_k<DH>(s0 + HB), gh = bdesc_k<DH>(s0 + 2 * HB), gl = bdesc_k<DH>(s0 + 3 * HB);
            const uint32_t t_st = tmem_base + u * 128, t_dpt = t_st + 64;
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < C::KS; ++kk) {
                    umma_ss(t_st, kh + 2 * kk, qh + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);
                    umma_ss(t_st, kl + 2 * kk, qh + 2 * kk, idesc_s, 1u);
                    umma_ss(t_st, kh + 2 * kk, ql + 2 * kk, idesc_s, 1u);
                }
#pragma unroll
                for (int kk = 0; kk < C::KS; ++kk) {
                    umma_ss(t_dpt, vh + 2 * kk, gh + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);
                    umma_ss(t_dpt, vl + 2 * kk, gh + 2 * kk, idesc_s, 1u);
                    umma_ss(t_dpt, vh + 2 * kk, gl + 2 * kk, idesc_s, 1u);
                }
                umma_commit(s_full0 + 8 * u);
            }
            __syncwarp();
        };
        issue_sdp(0);
        issue_sdp(1);
        for (int g = 0; g < n_g; ++g) {
            const int i = g >> 1, hh = g & 1, stg = g % QST, u = g % NBUF;
            mbar_wait(pds_full0 + 8 * u, (g / NBUF) & 1);           // P^T_g, dS^T_g in TMEM (+ dS^T_g staged)
            if (hh == 1 && i > 0) mbar_wait(dq_free, (i - 1) & 1);  // the previous tile's dQ has been drained
            tc_fence_after();
            const uint32_t s0 = sQG + stg * 4 * HB;
            const uint64_t qmh = bdesc_mn<DH>(s0), qml = bdesc_mn<DH>(s0 + HB), gmh = bdesc_mn<DH>(s0 + 2 * HB), gml = bdesc_mn<DH>(s0 + 3 * HB);
            const uint32_t t_st = tmem_base + u * 128, t_dpt = t_st + 64;
            const uint32_t acc0 = g > 0 ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < HQ / 16; ++kk) {   // K-steps over the 64 queries of this half: pairs (hi | lo) of column group kk
                    umma_ts(t_dv, t_st + 16 * kk, gmh + C::ROW_KSTEP * kk, idesc_g, kk > 0 ? 1u : acc0);
                    umma_ts(t_dv, t_st + 16 * kk + 8, gmh + C::ROW_KSTEP * kk, idesc_g, 1u);
                    umma_ts(t_dv, t_st + 16 * kk, gml + C::ROW_KSTEP * kk, idesc_g, 1u);
                }
#pragma unroll
                for (int kk = 0; kk < HQ / 16; ++kk) {
                    umma_ts(t_dk, t_dpt + 16 * kk, qmh + C::ROW_KSTEP * kk, idesc_g, kk > 0 ? 1u : acc0);
                    umma_ts(t_dk, t_dpt + 16 * kk + 8, qmh + C::ROW_KSTEP * kk, idesc_g, 1u);
                    umma_ts(t_dk, t_dpt + 16 * kk, qml + C::ROW_KSTEP * kk, idesc_g, 1u);
                }
                umma_commit(qg_empty0 + 8 * stg);       // this half's Q / dO are free once dV and dK retire
                if (hh == 1) {
                    const uint64_t dsh = bdesc_ds(sDS), dsl = bdesc_ds(sDS + 2 * 128 * 128);
#pragma unroll
                    for (int kk = 0; kk < BT / 16; ++kk) {   // dQ = dS K over the 128 keys, both staged halves as M = 128
                        umma_ss(t_dq, dsh + 128 * kk, kmh + C::ROW_KSTEP * kk, idesc_q, kk > 0 ? 1u : 0u);
                        umma_ss(t_dq, dsl + 128 * kk, kmh + C::ROW_KSTEP * kk, idesc_q, 1u);
                        umma_ss(t_dq, dsh + 128 * kk, kml + C::ROW_KSTEP * kk, idesc_q, 1u);
                    }
                    umma_commit(dq_full);
                    if (g == n_g - 1) umma_commit(dkv_full);
                }
            }
            __syncwarp();
            if (g + NBUF < n_g) issue_sdp(g + NBUF);   // into the buffer this half just released (the tensor pipe is in order)
        }
    } else {
        const int quad = warp & 3, cq = (warp - 2) >> 2;
        const int rl = quad * 32 + lane;
        const int key = k0 + rl;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int st_tid = threadIdx.x - 64;
        const float* lse = p.lse + ((size_t)b * p.H + h) * p.S;
        const float* delta = p.delta + ((size_t)b * p.H + h) * p.S;
        float* dbase = p.dqkv + (size_t)b * p.S * ld + h * DH;
        const unsigned long long seed = dyn_seed(p.seed, p.dyn);
        const bool drop = p.dropout_p > 0.f;
        const float inv_keep = drop ? 1.f / (1.f - p.dropout_p) : 1.f;
        const float keep_prob = drop ? 1.f - p.dropout_p : 1.f;
        const uint32_t thresh = drop_threshold(p.dropout_p);
        const unsigned long long bh = ((unsigned long long)b * p.H + h) * (unsigned long long)p.S;
        const uint32_t kc = drop_col_term((uint32_t)key);
        const float scale = rsqrtf((float)DH);
        const bool key_ok = key < len;

        // padded keys (rows >= len of this utterance, only in its last key tile): zero their K rows (both planes) once, so that
        // whatever their dS^T rows hold adds nothing to dQ = dS K; generic-proxy writes -> proxy fence -> the MMA warp's k_ready
        mbar_wait(kv_full, 0);
        if (!key_ok && cq == 0) {
#pragma unroll
            for (int c = 0; c < C::ROW_BYTES; c += 16) {
                *reinterpret_cast<uint4*>(smem_gen + (sK - base) + (size_t)rl * C::ROW_BYTES + c) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(smem_gen + (sK - base) + C::TB + (size_t)rl * C::ROW_BYTES + c) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(k_ready);
        float n_lse = 0.f, n_delta = 0.f;   // next HALF's per-query values: fetched at the start of a half-step, stored at its end
        auto fetch_aux = [&](int g) {
            if (st_tid < HQ) {
                const int q = g * HQ + st_tid;
                n_lse = q < len ? lse[q] : 0.f;
                n_delta = q < len ? delta[q] : 0.f;
            }
        };
        auto store_aux = [&](int g) {
            if (st_tid < HQ) {
                float* ax = aux + (g & 1) * 3 * HQ;
                // padded queries: lse = +inf makes P = exp2(S - inf) = 0 (and with it dS) without any per-score select
                ax[st_tid] = g * HQ + st_tid < len ? n_lse * kLog2eB : INFINITY;
                ax[HQ + st_tid] = n_delta * keep_prob;
                reinterpret_cast<uint32_t*>(ax)[2 * HQ + st_tid] = drop ? hash_u32(seed, bh + g * HQ + st_tid) : 0u;
            }
        };
        // dQ of query tile i: TMEM lane = query row; vector fp32 reductions (q carries 1/sqrt(dh)); hands the accumulator back first
        auto drain_dq = [&](int i, bool release) {
            mbar_wait(dq_full, i & 1);   // (already observed by the staging guard except for the last tile)
            tc_fence_after();
            if (DH == 64 || cq == 0) {
                const int c0 = DH == 64 ? cq * 16 : 0;
                uint32_t v[16];
                tmem_ld16(t_lane + COL_DQ + c0, v);
                tmem_ld_wait16(v);
                tc_fence_before();
                __syncwarp();
                if (release && lane == 0) mbar_arrive(dq_free);
                const int q = i * BT + rl;
                if (q < len) {
                    float* dq = dbase + (size_t)q * ld + c0;
                    const float sc = scale * inv_keep;
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dq + e), "f"(sc * __uint_as_float(v[e])),
                                     "f"(sc * __uint_as_float(v[e + 1])), "f"(sc * __uint_as_float(v[e + 2])), "f"(sc * __uint_as_float(v[e + 3]))
                                     : "memory");
                }
            } else {
                tc_fence_before();
                __syncwarp();
                if (release && lane == 0) mbar_arrive(dq_free);
            }
        };
        const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
        fetch_aux(0);
        store_aux(0);
        for (int g = 0; g < n_g; ++g) {
            const int i = g >> 1, hh = g & 1, q0 = i * BT;
            if (tr && g < 256) p.trace[g] = clock64();                       // half-step begins
            const float* ax = aux + (g & 1) * 3 * HQ;           // [lse*log2e | delta*keep | row key] of this half
            asm volatile("bar.sync 1, 512;" ::: "memory");      // this half's vectors are visible; the other half buffer is free
            if (g + 1 < n_g) fetch_aux(g + 1);
            const int u = g % NBUF;
            mbar_wait(s_full0 + 8 * u, (g / NBUF) & 1);
            if (tr && g < 256) p.trace[256 + g] = clock64();                 // S^T / dP^T of this half available
            tc_fence_after();
            const uint32_t t_st = t_lane + u * 128 + 16 * cq, t_dpt = t_st + 64;
            uint32_t s[16], gq[16];
            tmem_ld16(t_st, s);
            tmem_ld16(t_dpt, gq);
            tmem_ld_wait16(s);
            tmem_ld_wait16(gq);
            const int cbase = 16 * cq;                             // first query column of this thread within the half
            uint32_t pt[16], dst[16];                              // 8 hi pairs | 8 lo pairs
            // no per-score masking: padded queries carry lse = +inf (P = 0), padded keys have their K rows zeroed in shared memory
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = cbase + 2 * j;
                const float2 ls = *reinterpret_cast<const float2*>(ax + c);
                const float2 dl = *reinterpret_cast<const float2*>(ax + HQ + c);
                const float p0 = ex2(fmaf(__uint_as_float(s[2 * j]), kLog2eB, -ls.x));
                const float p1 = ex2(fmaf(__uint_as_float(s[2 * j + 1]), kLog2eB, -ls.y));
                float d0 = __uint_as_float(gq[2 * j]), d1 = __uint_as_float(gq[2 * j + 1]);
                float pd0 = p0, pd1 = p1;
                if (drop) {
                    const uint2 rk = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(ax) + 2 * HQ + c);
                    const bool k0_ = drop_keep(rk.x, kc, thresh), k1_ = drop_keep(rk.y, kc, thresh);
                    pd0 = k0_ ? p0 : 0.f; pd1 = k1_ ? p1 : 0.f;
                    d0 = k0_ ? d0 : 0.f; d1 = k1_ ? d1 : 0.f;
                }
                split_pair(pd0, pd1, pt[j], pt[8 + j]);
                split_pair(p0 * (d0 - dl.x), p1 * (d1 - dl.y), dst[j], dst[8 + j]);
            }
            tmem_st16(t_st, pt);
            tmem_st16(t_dpt, dst);
            if (tr && g < 256) p.trace[512 + g] = clock64();                 // arithmetic done
            if (hh == 0 && i > 0) mbar_wait(dq_full, (i - 1) & 1);           // the ONE staging buffer: dQ(i-1) has finished reading it
            {   // dS^T staging: row = key, chunk hh (64 queries = 128 B), this thread's 16 queries = units 2 cq, 2 cq + 1
                uint8_t* rowp = ds_gen + (size_t)hh * 16384 + (size_t)rl * 128;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int unit = (2 * cq + u) ^ (rl & 7);
                    *reinterpret_cast<uint4*>(rowp + unit * 16) = make_uint4(dst[4 * u], dst[4 * u + 1], dst[4 * u + 2], dst[4 * u + 3]);
                    *reinterpret_cast<uint4*>(rowp + 32768 + unit * 16) = make_uint4(dst[8 + 4 * u], dst[8 + 4 * u + 1], dst[8 + 4 * u + 2], dst[8 + 4 * u + 3]);
                }
            }
            tmem_st_wait();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pds_full0 + 8 * u);
            if (tr && g < 256) p.trace[768 + g] = clock64();                 // published
            if (g + 1 < n_g) store_aux(g + 1);                     // into the half buffer every thread left at this half-step's barrier
            if (hh == 1 && i > 0) drain_dq(i - 1, true);           // a whole tile late: dQ(i-1) retired long ago, nothing to wait for
        }
        drain_dq(n_q - 1, false);
        // ---- dK, dV of this key tile ----
        mbar_wait(dkv_full, 0);
        tc_fence_after();
        if (DH == 64 || cq == 0) {
            const int c0 = DH == 64 ? cq * 16 : 0;
            uint32_t vk[16], vv[16];
            tmem_ld16(t_lane + COL_DK + c0, vk);
            tmem_ld16(t_lane + COL_DV + c0, vv);
            tmem_ld_wait16(vk);
            tmem_ld_wait16(vv);
            if (key_ok) {
                float* dk = dbase + (size_t)key * ld + D + c0;
                float* dv = dbase + (size_t)key * ld + 2 * D + c0;
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    *reinterpret_cast<float4*>(dk + e) = make_float4(__uint_as_float(vk[e]) * inv_keep, __uint_as_float(vk[e + 1]) * inv_keep,
                                                                    __uint_as_float(vk[e + 2]) * inv_keep, __uint_as_float(vk[e + 3]) * inv_keep);
                    *reinterpret_cast<float4*>(dv + e) = make_float4(__uint_as_float(vv[e]) * inv_keep, __uint_as_float(vv[e + 1]) * inv_keep,
                                                                    __uint_as_float(vv[e + 2]) * inv_keep, __uint_as_float(vv[e + 3]) * inv_keep);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ABT_TMEM_COLS) : "memory");
    }
}

template <int DH>
int launch_bwd_tc(const AttnArgs& a, cudaStream_t st) {
    using C = AbtCfg<DH>;
    const int smem = C::SMEM_BYTES > 120 * 1024 ? C::SMEM_BYTES : 120 * 1024;   // one CTA per SM (all TMEM columns)
    static bool configured = false;
    if (!configured) {
        DX_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    CUtensorMap map_r, map_g;
    const int sw = DH == 64 ? 128 : 32;
    int rc = make_tma_map_3d(&map_r, a.R, 2, (unsigned long long)DH, (unsigned long long)a.Sp, (unsigned long long)2 * a.B * 3 * a.H,
                             (unsigned long long)DH * 2, (unsigned long long)a.Sp * DH * 2, DH, 128, 1, sw);
    if (rc) return rc;
    rc = make_tma_map_3d(&map_g, a.GR, 2, (unsigned long long)DH, (unsigned long long)a.Sp, (unsigned long long)2 * a.B * a.H,
                         (unsigned long long)DH * 2, (unsigned long long)a.Sp * DH * 2, DH, 128, 1, sw);
    if (rc) return rc;
    AbtParams p;
    p.dqkv = a.dqkv; p.lse = a.lse; p.delta = a.delta; p.lens = a.lens;
    p.B = a.B; p.S = a.S; p.H = a.H; p.Sp = a.Sp; p.dropout_p = a.dropout_p; p.seed = a.seed; p.dyn = a.dyn;
    p.trace = tc_trace_buffer();
    dim3 grid(ceil_div(a.S, BT), a.H, a.B);
    if constexpr (DH == 16) {
        static bool configured_pipe = false;
        if (!configured_pipe) {
            DX_CUDA(cudaFuncSetAttribute(attn_bwd_tc_pipe_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured_pipe = true;
        }
        attn_bwd_tc_pipe_kernel<DH><<<grid, ABT_THREADS, smem, st>>>(map_r, map_g, p);
        return check_launch("attn_bwd_tc_pipe");
    }
    if constexpr (DH == 64) {
        static int pipe64 = -1;
        if (pipe64 < 0) {
            const char* e = getenv("DX_ATTN_BWD_PIPE64");
            pipe64 = (e && atoi(e) == 0) ? 0 : 1;
        }
        if (pipe64) {
            static bool configured64 = false;
            if (!configured64) {
                DX_CUDA(cudaFuncSetAttribute(attn_bwd_tc_pipe64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ABT64_SMEM_BYTES));
                configured64 = true;
            }
            CUtensorMap map_rq, map_gq;   // 64-row boxes: Q and dO arrive as half tiles
            rc = make_tma_map_3d(&map_rq, a.R, 2, (unsigned long long)DH, (unsigned long long)a.Sp, (unsigned long long)2 * a.B * 3 * a.H,
                                 (unsigned long long)DH * 2, (unsigned long long)a.Sp * DH * 2, DH, 64, 1, sw);
            if (rc) return rc;
            rc = make_tma_map_3d(&map_gq, a.GR, 2, (unsigned long long)DH, (unsigned long long)a.Sp, (unsigned long long)2 * a.B * a.H,
                                 (unsigned long long)DH * 2, (unsigned long long)a.Sp * DH * 2, DH, 64, 1, sw);
            if (rc) return rc;
            attn_bwd_tc_pipe64_kernel<<<grid, ABT_THREADS, ABT64_SMEM_BYTES, st>>>(map_r, map_rq, map_gq, p);
            return check_launch("attn_bwd_tc_pipe64");
        }
    }
    attn_bwd_tc_kernel<DH><<<grid, ABT_THREADS, smem, st>>>(map_r, map_g, p);
    return check_launch("attn_bwd_tc");
}

}  // namespace

bool attention_bwd_tc_supported(const AttnArgs& a) {
    return (a.dh == 64 || a.dh == 16) && a.lens != nullptr && a.R != nullptr && a.GR != nullptr && tma_available();
}

// a.R / a.GR / a.Sp / a.delta bound and filled, dqkv zeroed (attention_mma.cu: attention_bwd_mma)
int attention_bwd_tc(const AttnArgs& a, cudaStream_t st) {
    if (a.dh == 64) return launch_bwd_tc<64>(a, st);
    if (a.dh == 16) return launch_bwd_tc<16>(a, st);
    set_last_error("attention_bwd_tc: unsupported head_dim %d", a.dh);
    return DX_ERR_UNSUPPORTED;
}

}  // namespace dx
