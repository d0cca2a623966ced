// K4, pass 1 of the tensor-core IPA operator for L <= 96: the warp-specialised "ping-pong" edition.
//
// What is different from the one-item-per-CTA edition in ipa_tc.cu (structure_module.py:168-186):
//
//  * The point term  sum_p |q_p(i) - k_p(j)|  needs 4 square roots per (i, j, head) -- they stay on the XU (MUFU) pipe -- but
//    the SQUARED distances are contractions:  d^2 = |q|^2 + |k|^2 - 2 q.k.  Each of the 4 point channels is one K = 16
//    tcgen05.mma (kind::f16, fp16 operands, fp32 accumulator in TMEM) on hi + lo split coordinates:
//        a(i) = [-2qh(3) | -2qh(3) | -2ql(3) | -2ql(3) | nq_h nq_l | 256 256]
//        b(j) = [  kh(3) |   kl(3) |   kh(3) |   kl(3) | 256 256   | nk_h nk_l]      a.b = |q|^2 + |k|^2 - 2 q.k
//    with x = xh + xl (fp16 pair: 22 significant bits, coordinates re-centred on the sample's first residue) and
//    n = |x|^2 / 256 split the same way (products with the exact constant 256 restore the scale; |x - x_0| < 4000 nm).
//    The logit pass then costs 4 MUFU.SQRT + ~9 other instructions per key instead of ~29 (no subtractions, no squares, no
//    key points in shared memory).
//  * One persistent CTA (512 threads, 128 registers each) per SM runs TWO warpgroups on different items.  In a warpgroup the six
//    warps that own TMEM lanes 0..95 are the consumers: two threads per query row -- "groups" 0 and 1 -- taking the even / odd
//    16-key chunks.  The warp of lane quadrant 3 (rows 96..127 do not exist for L <= 96) is the ISSUER: one lane issues
//    every TMA load and every tcgen05.mma of the warpgroup's items and never touches data.  The distance accumulators
//    are produced chunk by chunk into one 64-column TMEM stage per group: a chunk is pulled into registers (80 of them),
//    the stage is handed back to the issuer, and the group's next chunk is contracted under the square roots of this one.
//    The warpgroups can hand the XU-bound phase (logits + exponentials) to each other through a pair of named barriers,
//    so that one warpgroup's staging / frame transforms / P.V product / epilogue run under the other's square roots.
//
// Per item (sample b, head h) and warpgroup:
//   issuer   TMA: bf16 q | k | v records as UMMA K-chunks, the head's point records, the sample's frames (prefetched one item
//            ahead), the head's pair-bias slab; S = Q.K^T as soon as q / k have landed
//   group 0  4 key points -> global frame, re-centred, hi/lo fp16 rows of the B operands     } then the issuer contracts
//   group 1  4 query points -> rows of the A operands                                        } D_p chunks 0 and 1
//   both     value points 4g .. 4g+3 -> hi/lo bf16 channels of the MN-major value operand (off the critical path)
//   XU phase
//     pass A per 16-key chunk: tcgen05.ld S and the four D_p, release the stage, logits =
//            S + hw * sum_p sqrt|D_p| + pair_bias + key_bias (log2 domain) parked back into the S columns, row maximum
//     pass B: P = exp2(l - max) -> bf16 -> shared memory (A operand of P.V; overlays the dead point operands) and the row
//            of the probability workspace for pass 2
//   issuer   O = P.[v | v_pts hi | lo | 1] into the stage columns -> commit
//   epilogue group 0: scalars + points 0..3, group 1: points 4..7 (normalise, undo the re-centring, inverse frame, norms)
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

#include "ipa_tc_shared.cuh"

using namespace se3;
using namespace se3::ipa_tc;

namespace {

constexpr int kThreadsPP = 512;          // 2 warpgroups x 8 warps: 6 consumers, 1 issuer, 1 helper
constexpr int kConsumers = 192;          // consumer threads per warpgroup
constexpr int kMaxLenPP = 96;
constexpr int kTurn0 = 1, kTurn1 = 2;    // named barriers: permission to enter the XU phase
constexpr int kWgBar0 = 3;               // + wg: barrier over the consumers of one warpgroup
constexpr float kNormScale = 1.0f / 256.0f;

struct PpPlan {
    uint32_t raw, q, k, vs, vs_step, frm, frm_step, vp, ops, bias, kb, hmax, total;   // vs / frm: two buffers `step` bytes apart
    int Lq, Lpi, nck;                                                                 // keys padded to 16, bias row pitch, 16-key chunks
};
// Shared memory of ONE warpgroup.
__host__ __device__ inline PpPlan pp_plan(int L, bool pts_bf16) {
    const uint32_t Lq = (uint32_t)(L + 15) & ~15u, Lpi = (uint32_t)ipa_bias_pitch(L);
    auto up = [](uint32_t x) { return (x + 127u) & ~127u; };
    PpPlan p;
    uint32_t o = 0;
    p.raw = o; o += up((uint32_t)L * (pts_bf16 ? 96u : 192u));       // [L][qp 12 | kp 12 | vp 24] local point records
    p.q = o; o += 4096;                                              // [2][128][16 B]
    p.k = o; o += Lq * 32;                                           // [2][Lq][16 B]
    p.vs = o; p.vs_step = Lq * 32; o += 2 * p.vs_step;               // [2][Lq][16 B], double buffered (P.V of item n runs under the loads of n + 1)
    p.frm = o; p.frm_step = up((uint32_t)L * 48); o += 2 * p.frm_step;   // rotations [L][9] then translations [L][3], double buffered (epilogue)
    p.vp = o; o += Lq * (NVP * 2);                                   // [Lq/8][NVP/8][8][8] bf16
    p.ops = o;                                                       // A operands 4 x [2][128][16 B] fp16, B operands 4 x [2][Lq][16 B]; later P [Lq/8][128][16 B]
    { const uint32_t a = 16384u + Lq * 128u, b = Lq * 256u; o += a > b ? a : b; }
    p.bias = o; o += up((uint32_t)L * Lpi * 2);                      // bf16 [L queries][Lpi keys] (common.cuh: ipa_bias_pitch)
    p.kb = o; o += up(Lq * 4);
    p.hmax = o; o += 1024;
    p.total = o;
    p.Lq = (int)Lq; p.Lpi = (int)Lpi; p.nck = (int)(Lq >> 4);
    return p;
}

struct PpBars {   // per warpgroup
    uint64_t in_full, bias_full, ops_ready, raw_free, st_full[2], st_free[2], p_ready, o_full;
};

__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
// (a, b) -> fp16 pair of the leading halves and fp16 pair of the remainders (F2FP packs; the scalar cvt runs on the XU pipe)
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    hi = h2_bits(h);
    lo = h2_bits(__floats2half2_rn(a - f.x, b - f.y));
}
// bf16 at a shared-window address -> float (32-bit address arithmetic; a generic pointer costs 64-bit adds and LD instead of LDS)
__device__ __forceinline__ float lds_bf16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return __uint_as_float((uint32_t)v << 16);
}
__device__ __forceinline__ float sqrt_abs(float x) { return fast_sqrt(fabsf(x)); }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---- issuer-side helpers (one thread) ---------------------------------------------------------------------------------
// every input of item (b, h) except the pair-bias slab: point records, q / k / v K-chunks, frames -> `in_full`
template <bool kPtsBf16>
__device__ __forceinline__ void pp_issue_loads(int item, int buf, int L, int Lq, int H, uint8_t* base, const PpPlan& pl, const CUtensorMap* map_q,
                                               const CUtensorMap* map_kv, const CUtensorMap* map_pts, const float* rot, const float* trans,
                                               bool bulk_frames, uint64_t* in_full) {
    constexpr int kRawRow = kPtsBf16 ? 96 : 192;
    const int b_ = item / H, h_ = item - b_ * H, row0 = b_ * L;
    tc::mbar_expect_tx(in_full, (uint32_t)(L * kRawRow + 4096 + Lq * 64 + (bulk_frames ? L * 48 : 0)));
    tc::tma_tile_2d_g2s(base + pl.raw, map_pts, h_ * 48, row0, in_full);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        tc::tma_tile_2d_g2s(base + pl.k + (size_t)half * Lq * 16, map_kv, h_ * 48 + 16 + half * 8, row0, in_full);
        tc::tma_tile_2d_g2s(base + pl.q + (size_t)half * 2048, map_q, h_ * 48 + half * 8, row0, in_full);
        tc::tma_tile_2d_g2s(base + pl.vs + buf * pl.vs_step + (size_t)half * Lq * 16, map_kv, h_ * 48 + 32 + half * 8, row0, in_full);
    }
    if (bulk_frames) {
        tc::tma_bulk_g2s(base + pl.frm + buf * pl.frm_step, rot + (int64_t)row0 * 9, (uint32_t)(L * 36), in_full);
        tc::tma_bulk_g2s(base + pl.frm + buf * pl.frm_step + (size_t)L * 36, trans + (int64_t)row0 * 3, (uint32_t)(L * 12), in_full);
    }
}
__device__ __forceinline__ void pp_issue_bias(int item, int L, int Lpi, int H, uint8_t* dst, const __nv_bfloat16* pair_bias_t, uint64_t* bias_full) {
    const int h_ = item % H;
    tc::mbar_expect_tx(bias_full, (uint32_t)(L * Lpi * 2));
    tc::tma_bulk_g2s(dst, pair_bias_t + (int64_t)h_ * L * Lpi, (uint32_t)(L * Lpi * 2), bias_full);
}
// squared distances of 16-key chunk c, four point channels -> the stage of group c & 1
__device__ __forceinline__ void pp_issue_dist(int c, uint32_t tmem_st, uint32_t a_ops, uint32_t b_ops, int Lq, uint32_t idesc_d, uint64_t* st_full) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
        tc::mma_bf16(tmem_st + (uint32_t)((c & 1) * 64 + p * 16), tc::make_desc(a_ops + (uint32_t)p * 4096u, 128),
                     tc::make_desc_raw(b_ops + (uint32_t)p * (uint32_t)Lq * 32u + (uint32_t)c * 256u, (uint32_t)Lq * 16u, 128u), idesc_d, false);
    tc::mma_commit(st_full);
}

template <typename OutT, bool kPtsBf16>
__global__ void __launch_bounds__(kThreadsPP, 1)
k_ipa_tc_pass1_pp(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_pts,
                  const float* __restrict__ rot, const float* __restrict__ trans, const __nv_bfloat16* __restrict__ pair_bias_t,
                  const float* __restrict__ key_bias, const float* __restrict__ head_weight, OutT* __restrict__ out,
                  __nv_bfloat16* __restrict__ pbuf, float* __restrict__ inv_sum, const __grid_constant__ se3_ipa_shape sh, const __grid_constant__ PpPlan pl,
                  int Bpad, int use_turns, long long* __restrict__ dbg) {
    constexpr int kRawRow = kPtsBf16 ? 96 : 192;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ PpBars bars[2];
    __shared__ uint32_t tmem_slot;
    const int L = sh.len, H = sh.heads;
    const int Lq = pl.Lq, Lpi = pl.Lpi, LpT = pl.Lq;
    const int nck = pl.nck;                                  // 16-key chunks; chunk c belongs to group c & 1
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wg = warp >> 3, wi = warp & 7;                 // warpgroup, warp within it
    const int quad = wi & 3;                                 // TMEM lane quadrant this warp may touch (= warp % 4)
    uint8_t* base = smem_raw + (size_t)wg * pl.total;
    PpBars& bar = bars[wg];
    const bool bulk_frames = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(rot) | reinterpret_cast<uintptr_t>(trans)) & 15) == 0;

    // work items (sample b, head h): item = b * H + h; worker = 2 * CTA + warpgroup takes every (2 * gridDim)-th item, so that
    // the workers running at any moment read neighbouring (sample, head) records (shared DRAM pages)
    const int n_items = H * sh.batch;
    const int stride = 2 * (int)gridDim.x;
    const int first = 2 * (int)blockIdx.x + wg;
    const int n_mine = first < n_items ? (n_items - first + stride - 1) / stride : 0;
    const int n_other = (first ^ 1) < n_items ? (n_items - (first ^ 1) + stride - 1) / stride : 0;   // items of the other warpgroup of this CTA

    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            tc::mbar_init(&bars[w].in_full, 1);
            tc::mbar_init(&bars[w].bias_full, 1);
            tc::mbar_init(&bars[w].ops_ready, 6);
            tc::mbar_init(&bars[w].raw_free, 6);
            tc::mbar_init(&bars[w].st_full[0], 1);
            tc::mbar_init(&bars[w].st_full[1], 1);
            tc::mbar_init(&bars[w].st_free[0], 3);
            tc::mbar_init(&bars[w].st_free[1], 3);
            tc::mbar_init(&bars[w].p_ready, 6);
            tc::mbar_init(&bars[w].o_full, 1);
        }
        tc::mbar_fence_init();
    }
    if (wi == 7) {
        // helper warp, once per kernel: the value operand's padding keys must be exactly zero (0 * garbage could be NaN) and keep
        // logits of -inf; nothing else ever writes those rows.  (Padding rows of the distance operands only feed accumulator
        // entries that are never read.)
        float* s_kb = reinterpret_cast<float*>(base + pl.kb);
        for (int j = lane; j < Lq; j += 32) s_kb[j] = j < L ? 0.f : -CUDART_INF_F;
        for (int j = L + lane; j < Lq; j += 32) {
            uint8_t* vcol = base + pl.vp + (size_t)(j >> 3) * (NVP * 16) + (size_t)(j & 7) * 16;
#pragma unroll
            for (int c = 0; c < NVP / 8; ++c) *reinterpret_cast<uint4*>(vcol + c * 128) = make_uint4(0, 0, 0, 0);
        }
        tc::fence_async_smem();
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_s = tmem_slot + (uint32_t)wg * 256u;   // S / logits: Lq columns
    const uint32_t tmem_st = tmem_s + 128u;                    // two stages of 4 x 16 distance columns (one per group), later the O accumulator (80)

    if (wi == 3) {
        // ================================ issuer warp: TMA + tcgen05.mma of warpgroup `wg` ================================
        if (lane == 0 && n_mine > 0) {
            pp_issue_loads<kPtsBf16>(first, 0, L, Lq, H, base, pl, &map_q, &map_kv, &map_pts, rot, trans, bulk_frames, &bar.in_full);
            pp_issue_bias(first, L, Lpi, H, base + pl.bias, pair_bias_t, &bar.bias_full);
            const uint32_t a_ops = tc::smem_u32(base + pl.ops), b_ops = a_ops + 16384u;
            const uint32_t idesc_s = tc::make_idesc_bf16(128, Lq), idesc_d = tc::make_idesc_f16(128, 16);
            const uint32_t idesc_vs = tc::make_idesc_bf16(128, DK, /*b_mn_major=*/true), idesc_vp = tc::make_idesc_bf16(128, NVP, /*b_mn_major=*/true);
            uint32_t n_full0 = 0, n_free0 = 0, n_free1 = 0;   // completed phases of st_full[0] / phases of st_free[g] already waited for
            for (int it = 0; it < n_mine; ++it) {
                const int item = first + it * stride, buf = it & 1;
                tc::mbar_wait(&bar.in_full, (uint32_t)(it & 1));
                tc::fence_after();                             // (the previous item's logits left the S columns before its p_ready)
                tc::mma_bf16(tmem_s, tc::make_desc(tc::smem_u32(base + pl.q), 128), tc::make_desc(tc::smem_u32(base + pl.k), (uint32_t)Lq), idesc_s, false);
                tc::mbar_wait(&bar.ops_ready, (uint32_t)(it & 1));
                tc::fence_after();
                if (dbg) dbg[(int64_t)item * 32 + 24] = clock64();
                pp_issue_dist(0, tmem_st, a_ops, b_ops, Lq, idesc_d, &bar.st_full[0]);
                if (nck > 1) pp_issue_dist(1, tmem_st, a_ops, b_ops, Lq, idesc_d, &bar.st_full[1]);
                if (it + 1 < n_mine) {         // q / k are dead once S exists, the point records once the value points are transformed
                    tc::mbar_wait(&bar.st_full[0], n_full0 & 1u);
                    tc::mbar_wait(&bar.raw_free, (uint32_t)(it & 1));
                    pp_issue_loads<kPtsBf16>(item + stride, buf ^ 1, L, Lq, H, base, pl, &map_q, &map_kv, &map_pts, rot, trans, bulk_frames, &bar.in_full);
                }
                n_full0 += (uint32_t)((nck + 1) >> 1);
                for (int c = 2; c < nck; ++c) {
                    if (c & 1) tc::mbar_wait(&bar.st_free[1], n_free1++ & 1u);
                    else tc::mbar_wait(&bar.st_free[0], n_free0++ & 1u);
                    tc::fence_after();
                    if (dbg) dbg[(int64_t)item * 32 + 16 + c] = clock64();
                    pp_issue_dist(c, tmem_st, a_ops, b_ops, Lq, idesc_d, &bar.st_full[c & 1]);
                    if (dbg && c == 2) dbg[(int64_t)item * 32 + 30] = clock64();
                }
                tc::mbar_wait(&bar.p_ready, (uint32_t)(it & 1));
                tc::fence_after();
                if (dbg) dbg[(int64_t)item * 32 + 25] = clock64();
                const uint32_t p_addr = a_ops, vs_addr = tc::smem_u32(base + pl.vs + buf * pl.vs_step), vp_addr = tc::smem_u32(base + pl.vp);
                for (int ks = 0; ks < Lq / 16; ++ks) {
                    const uint64_t a_desc = tc::make_desc_kstep(p_addr, 128, ks);
                    tc::mma_bf16(tmem_st, a_desc, tc::make_desc_raw(vs_addr + (uint32_t)ks * 256u, 128u, (uint32_t)Lq * 16u), idesc_vs, ks > 0);
                    tc::mma_bf16(tmem_st + DK, a_desc, tc::make_desc_raw(vp_addr + (uint32_t)ks * 2u * NVP * 16u, NVP * 16u, 128u), idesc_vp, ks > 0);
                }
                tc::mma_commit(&bar.o_full);
                if (dbg) dbg[(int64_t)item * 32 + 26] = clock64();
                if (it + 1 < n_mine) pp_issue_bias(item + stride, L, Lpi, H, base + pl.bias, pair_bias_t, &bar.bias_full);   // the logit pass of this item is over: its slab may be replaced
            }
        }
        __syncwarp();
    } else if (wi != 7) {
        // ================================ consumer warps (TMEM lane quadrants 0..2) =========================================
        const int g = wi >> 2;                               // group: even / odd 16-key chunks; which side of the transform
        const int row = quad * 32 + lane;                    // query row = TMEM lane (also: residue of the transform), 0..95
        const int t = g * 96 + row;                          // consumer index within the warpgroup
        const uint32_t lane_base = (uint32_t)(quad * 32);
        const bool row_ok = row < L;
        const bool warp_ok = quad * 32 < L;
        const int wgbar = kWgBar0 + wg;
        const int my_turn = wg ? kTurn1 : kTurn0, other_turn = wg ? kTurn0 : kTurn1;
        float* s_kb = reinterpret_cast<float*>(base + pl.kb);
        float* s_hmax = reinterpret_cast<float*>(base + pl.hmax);
        const __nv_bfloat16* s_bias = reinterpret_cast<const __nv_bfloat16*>(base + pl.bias);
        uint8_t* s_ops = base + pl.ops;
        uint32_t n_full = 0;
#define SE3_STAMP(k) do { if (dbg && t == 0) dbg[((int64_t)(first + it * stride)) * 32 + (k)] = clock64(); } while (0)
        if (use_turns && wg == 1 && n_other > 0) tc::bar_arrive(kTurn0, 2 * kConsumers);   // warpgroup 0 goes first

        for (int it = 0; it < n_mine; ++it) {
            const int item = first + it * stride, buf = it & 1;
            const int b = item / H, h = item - b * H;
            SE3_STAMP(0);
            tc::mbar_wait(&bar.in_full, (uint32_t)(it & 1));
            SE3_STAMP(1);
            const float* s_rot = reinterpret_cast<const float*>(base + pl.frm + buf * pl.frm_step);
            const float* f_rot = bulk_frames ? s_rot : rot + (int64_t)b * L * 9;        // unaligned sample blocks: frames straight from global memory
            const float* f_trn = bulk_frames ? s_rot + L * 9 : trans + (int64_t)b * L * 3;
            const float cx = f_trn[0], cy = f_trn[1], cz = f_trn[2];   // re-centring: the sample's first residue

            // ---- frame transform, one thread per (residue, group) ------------------------------------------------------------
            float R[9], T[3];
            const uint8_t* rawrow = base + pl.raw + (size_t)(row_ok ? row : 0) * kRawRow;
            {
                const int r = row_ok ? row : 0;
#pragma unroll
                for (int c = 0; c < 9; ++c) R[c] = f_rot[r * 9 + c];
                T[0] = f_trn[r * 3] - cx; T[1] = f_trn[r * 3 + 1] - cy; T[2] = f_trn[r * 3 + 2] - cz;
            }
            if (row_ok) {
                {
                    // group 0: the key points = rows of the B operands; group 1: the query points = rows of the A operands.
                    // chunk 0 at +0, chunk 1 at +k1 (16 bytes per row), point p at +p * pstep
                    const uint32_t pstep = g ? 4096u : (uint32_t)Lq * 32u, k1 = g ? 2048u : (uint32_t)Lq * 16u;
                    uint8_t* drow = s_ops + (g ? 0 : 16384) + (size_t)row * 16;
                    float l[12];
                    load_coords<kPtsBf16, 12>(rawrow, g ? 0 : 12, l);
                    const float sc = g ? -2.f : 1.f;           // the query side carries the factor of -2 q.k (exact, commutes with the split)
                    constexpr uint32_t c256 = 0x5C005C00u;     // fp16 (256, 256)
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        float gx, gy, gz;
                        to_global(R, T, l[3 * p], l[3 * p + 1], l[3 * p + 2], gx, gy, gz);
                        uint32_t hxy, lxy, hzn, lzn;           // (x, y) and (z, |.|^2 / 256): leading halves, remainders
                        split_h2(sc * gx, sc * gy, hxy, lxy);
                        split_h2(sc * gz, (gx * gx + gy * gy + gz * gz) * kNormScale, hzn, lzn);
                        const uint32_t nn = __byte_perm(hzn, lzn, 0x7632);    // (n_h, n_l)
                        const uint32_t lyz = __byte_perm(lxy, lzn, 0x5432);   // (l_y, l_z)
                        uint8_t* dst = drow + (size_t)p * pstep;
                        if (g) {   // a = [hx hy hz hx hy hz lx ly | lz lx ly lz n_h n_l 256 256]
                            *reinterpret_cast<uint4*>(dst) = make_uint4(hxy, __byte_perm(hzn, hxy, 0x5410), __byte_perm(hxy, hzn, 0x5432), lxy);
                            *reinterpret_cast<uint4*>(dst + k1) = make_uint4(__byte_perm(lzn, lxy, 0x5410), lyz, nn, c256);
                        } else {   // b = [hx hy hz lx ly lz hx hy | hz lx ly lz 256 256 n_h n_l]
                            const uint32_t hzlx = __byte_perm(hzn, lxy, 0x5410);
                            *reinterpret_cast<uint4*>(dst) = make_uint4(hxy, hzlx, lyz, hxy);
                            *reinterpret_cast<uint4*>(dst + k1) = make_uint4(hzlx, lyz, c256, nn);
                        }
                    }
                }
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar.ops_ready);
            if (row_ok) {
                {
                    // value points 4g .. 4g+3 = channels 12g .. 12g+11 of the 24 hi (and of the 24 lo) channels:
                    // group 0: chunk 0 and the first half of chunk 1; group 1: the second half of chunk 1 and chunk 2
                    uint8_t* vcol = base + pl.vp + (size_t)(row >> 3) * (NVP * 16) + (size_t)(row & 7) * 16;   // 16-byte channel chunk c of this key at + c * 128
                    float l[12], gv[12];
                    load_coords<kPtsBf16, 12>(rawrow, 24 + 12 * g, l);
#pragma unroll
                    for (int p = 0; p < 4; ++p) to_global(R, T, l[3 * p], l[3 * p + 1], l[3 * p + 2], gv[3 * p], gv[3 * p + 1], gv[3 * p + 2]);
                    uint32_t hi[6], lo[6];
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        const __nv_bfloat162 hh = __floats2bfloat162_rn(gv[2 * c], gv[2 * c + 1]);
                        hi[c] = *reinterpret_cast<const uint32_t*>(&hh);
                        lo[c] = tc::pack_bf16(gv[2 * c] - __bfloat162float(hh.x), gv[2 * c + 1] - __bfloat162float(hh.y));
                    }
                    if (g == 0) {
                        *reinterpret_cast<uint4*>(vcol) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint2*>(vcol + 128) = make_uint2(hi[4], hi[5]);
                        *reinterpret_cast<uint4*>(vcol + 3 * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        *reinterpret_cast<uint2*>(vcol + 4 * 128) = make_uint2(lo[4], lo[5]);
                        *reinterpret_cast<uint4*>(vcol + 6 * 128) = make_uint4(0x00003F80u, 0, 0, 0);     // bf16 1.0 in channel 48: row sum of the rounded probabilities
                        *reinterpret_cast<uint4*>(vcol + 7 * 128) = make_uint4(0, 0, 0, 0);
                        if (key_bias) s_kb[row] = key_bias[(int64_t)b * L + row] * kLog2e;
                    } else {
                        *reinterpret_cast<uint2*>(vcol + 128 + 8) = make_uint2(hi[0], hi[1]);
                        *reinterpret_cast<uint4*>(vcol + 2 * 128) = make_uint4(hi[2], hi[3], hi[4], hi[5]);
                        *reinterpret_cast<uint2*>(vcol + 4 * 128 + 8) = make_uint2(lo[0], lo[1]);
                        *reinterpret_cast<uint4*>(vcol + 5 * 128) = make_uint4(lo[2], lo[3], lo[4], lo[5]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar.raw_free);        // the point records are consumed
            SE3_STAMP(2);

            // ---- XU phase ------------------------------------------------------------------------------------------------
            if (use_turns) tc::bar_sync(my_turn, 2 * kConsumers);
            tc::mbar_wait(&bar.bias_full, (uint32_t)(it & 1));
            SE3_STAMP(3);
            float m = -CUDART_INF_F;
            const float hw = head_weight[h] * kLog2e;
            const uint32_t bias_col = tc::smem_u32(s_bias + min(row, L - 1) * Lpi);   // [query][j]: this row's biases, key by key
            for (int c = g; c < nck; c += 2) {
                const int j0 = c * 16;                         // this group's chunk: 16 keys, S and four D_p = 80 registers
                tc::mbar_wait(&bar.st_full[g], n_full & 1u);
                ++n_full;
                tc::fence_after();
                if (dbg && (t == 0 || t == 96)) dbg[((int64_t)(first + it * stride)) * 32 + 8 + c] = clock64();
                const bool active = warp_ok && j0 < L;         // warp-uniform
                uint32_t d[4][16];
                if (active) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) tc::tmem_ld16(tc::tmem_addr(tmem_st, lane_base, (uint32_t)(g * 64 + p * 16)), d[p]);
                    tc::tmem_wait_ld();
                }
                if (dbg && t == 0 && c == 0) dbg[((int64_t)(first + it * stride)) * 32 + 27] = clock64();
                if (c + 2 < nck) {                             // the stage may be overwritten by this group's next chunk
                    tc::fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bar.st_free[g]);
                }
                if (active) {
                    // sum of the four distances per key; the S chunk is fetched under the second half of the square roots
                    // (64 + 16 live registers would not fit beside the loop's own state)
                    const int nk = min(16, L - j0);            // real keys in this chunk (warp-uniform)
                    float ds[16];
                    uint32_t s[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        if (u == 8) tc::tmem_ld16(tc::tmem_addr(tmem_s, lane_base, (uint32_t)j0), s);
                        ds[u] = (u < nk) ? (sqrt_abs(__uint_as_float(d[0][u])) + sqrt_abs(__uint_as_float(d[1][u]))) +
                                               (sqrt_abs(__uint_as_float(d[2][u])) + sqrt_abs(__uint_as_float(d[3][u])))
                                         : 0.f;
                    }
                    tc::tmem_wait_ld();
                    uint32_t bj = bias_col + (uint32_t)(j0 * 2);   // walked key by key
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const float4 kb4 = *reinterpret_cast<const float4*>(s_kb + j0 + 4 * q4);   // -inf on padding keys
                        const float kbv[4] = {kb4.x, kb4.y, kb4.z, kb4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int u = 4 * q4 + e;
                            float lg = fmaf(hw, ds[u], __uint_as_float(s[u]) + kbv[e]);
                            if (u < nk) lg = fmaf(lds_bf16(bj), kLog2e, lg);
                            asm volatile("" : "+r"(bj));       // keep the address a loop-carried register
                            bj += 2u;
                            m = fmaxf(m, lg);
                            s[u] = __float_as_uint(lg);
                        }
                    }
                    tc::tmem_st16(tc::tmem_addr(tmem_s, lane_base, (uint32_t)j0), s);
                    if (dbg && t == 0 && c == 0) dbg[((int64_t)(first + it * stride)) * 32 + 28] = clock64();
                } else if (warp_ok) {                          // a whole chunk of padding keys
                    uint32_t s[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) s[u] = __float_as_uint(-CUDART_INF_F);
                    tc::tmem_st16(tc::tmem_addr(tmem_s, lane_base, (uint32_t)j0), s);
                }
            }
            tc::tmem_wait_st();
            s_hmax[t] = m;
            tc::bar_sync(wgbar, kConsumers);                   // both key halves of every row have their maximum; every D chunk is consumed
            m = fmaxf(s_hmax[row], s_hmax[96 + row]);
            if (m == -CUDART_INF_F) m = 0.f;
            SE3_STAMP(4);
            if (warp_ok) {
                // pass B: P = exp2(l - m) -> bf16 -> A operand of P.V (overlays the point operands) + this row of the workspace
                uint8_t* prow = reinterpret_cast<uint8_t*>(pbuf) + ((((int64_t)h * L + (row_ok ? row : 0)) * Bpad + b) * LpT) * 2;
                for (int c = g; c < nck; c += 2) {
                    const int j0 = c * 16;
                    uint32_t r[16], pk[8];
                    tc::tmem_ld16(tc::tmem_addr(tmem_s, lane_base, (uint32_t)j0), r);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int u = 0; u < 8; ++u) pk[u] = tc::pack_bf16(fast_ex2(__uint_as_float(r[2 * u]) - m), fast_ex2(__uint_as_float(r[2 * u + 1]) - m));
                    *reinterpret_cast<uint4*>(s_ops + ((size_t)(j0 >> 3) * 128 + row) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(s_ops + ((size_t)((j0 >> 3) + 1) * 128 + row) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    if (row_ok)
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                                     :: "l"(prow + (size_t)j0 * 2), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                }
            }
            tc::fence_async_smem();
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar.p_ready);
            SE3_STAMP(5);
            if (use_turns && it < n_other - (wg ? 1 : 0)) tc::bar_arrive(other_turn, 2 * kConsumers);   // the other warpgroup still has an item waiting for the XU phase

            // ---- epilogue: group 0 scalars + points 0..3, group 1 points 4..7 ---------------------------------------------------
            tc::mbar_wait(&bar.o_full, (uint32_t)(it & 1));
            tc::fence_after();
            SE3_STAMP(6);
            if (warp_ok) {
                // O columns: v 0..15 | point hi 16..39 | point lo 40..63 | row sum 64
                uint32_t rs[4], rh[16], rl[16], rv[16];
                tc::tmem_ld4(tc::tmem_addr(tmem_st, lane_base, 64), rs);
                tc::tmem_ld16(tc::tmem_addr(tmem_st, lane_base, (uint32_t)(g ? 28 : 16)), rh);   // [0, 12): hi halves of this group's four points
                tc::tmem_ld16(tc::tmem_addr(tmem_st, lane_base, (uint32_t)(g ? 52 : 40)), rl);   // [0, 12): lo halves
                if (g == 0) tc::tmem_ld16(tc::tmem_addr(tmem_st, lane_base, 0), rv);
                tc::tmem_wait_ld();
                if (row_ok) {
                    float Ri[9], Ti[3];                        // the row's own frame (the frames are double buffered for this)
#pragma unroll
                    for (int c = 0; c < 9; ++c) Ri[c] = f_rot[row * 9 + c];
                    Ti[0] = cx - f_trn[row * 3]; Ti[1] = cy - f_trn[row * 3 + 1]; Ti[2] = cz - f_trn[row * 3 + 2];
                    const float inv = fast_rcp(__uint_as_float(rs[0]));
                    const int HD = H * DK;
                    OutT* orow = out + ((int64_t)b * L + row) * (int64_t)(2 * HD + 4 * H * PV);
                    float pl4[12], nr[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float gx = (__uint_as_float(rh[p * 3]) + __uint_as_float(rl[p * 3])) * inv + Ti[0];
                        const float gy = (__uint_as_float(rh[p * 3 + 1]) + __uint_as_float(rl[p * 3 + 1])) * inv + Ti[1];
                        const float gz = (__uint_as_float(rh[p * 3 + 2]) + __uint_as_float(rl[p * 3 + 2])) * inv + Ti[2];
                        pl4[p * 3] = Ri[0] * gx + Ri[3] * gy + Ri[6] * gz;
                        pl4[p * 3 + 1] = Ri[1] * gx + Ri[4] * gy + Ri[7] * gz;
                        pl4[p * 3 + 2] = Ri[2] * gx + Ri[5] * gy + Ri[8] * gz;
                        nr[p] = fast_sqrt(pl4[p * 3] * pl4[p * 3] + pl4[p * 3 + 1] * pl4[p * 3 + 1] + pl4[p * 3 + 2] * pl4[p * 3 + 2]);
                    }
                    OutT* pdst = orow + HD + h * PV * 3 + g * 12;
                    OutT* ndst = orow + 2 * HD + 3 * H * PV + h * PV + g * 4;
                    if constexpr (std::is_same<OutT, float>::value) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) reinterpret_cast<float4*>(pdst)[c] = make_float4(pl4[4 * c], pl4[4 * c + 1], pl4[4 * c + 2], pl4[4 * c + 3]);
                        *reinterpret_cast<float4*>(ndst) = make_float4(nr[0], nr[1], nr[2], nr[3]);
                    } else {
                        *reinterpret_cast<uint2*>(pdst) = make_uint2(tc::pack_bf16(pl4[0], pl4[1]), tc::pack_bf16(pl4[2], pl4[3]));
                        *reinterpret_cast<uint2*>(pdst + 4) = make_uint2(tc::pack_bf16(pl4[4], pl4[5]), tc::pack_bf16(pl4[6], pl4[7]));
                        *reinterpret_cast<uint2*>(pdst + 8) = make_uint2(tc::pack_bf16(pl4[8], pl4[9]), tc::pack_bf16(pl4[10], pl4[11]));
                        *reinterpret_cast<uint2*>(ndst) = make_uint2(tc::pack_bf16(nr[0], nr[1]), tc::pack_bf16(nr[2], nr[3]));
                    }
                    if (g == 0) {
                        inv_sum[((int64_t)h * L + row) * Bpad + b] = inv;
                        float sc[DK];
#pragma unroll
                        for (int c = 0; c < DK; ++c) sc[c] = __uint_as_float(rv[c]) * inv;
                        store_vec<DK>(orow + h * DK, sc);
                    }
                }
            }
            tc::fence_before();                                // the accumulator reads are ordered before the next item's ops_ready arrive
            SE3_STAMP(7);
        }
#undef SE3_STAMP
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_slot, 512);
}

long long* g_pp_dbg = nullptr;
int pp_use_turns() {
    const char* v = getenv("SE3DIFF_B200_IPA_PP_TURNS");
    return v ? (v[0] != '0') : 1;
}

template <typename OutT, bool kPtsBf16>
int launch_pp(const Pass1Args& a) {
    const int L = a.sh.len, Lq = (L + 15) & ~15;
    const PpPlan pl = pp_plan(L, kPtsBf16);
    size_t smem = (size_t)pl.total * 2;
    if (smem + 1024 > 227 * 1024) return SE3_EUNSUPPORTED;
    if (smem < 120 * 1024) smem = 120 * 1024;               // one CTA per SM (each allocates all 512 TMEM columns)
    CUtensorMap map_q, map_kv, map_pts;
    const uint64_t rows = (uint64_t)a.sh.batch * L, width = (uint64_t)a.sh.heads * 48;
    if (int rc = make_map_2d(&map_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.scal, width, rows, (uint64_t)a.scal_stride, 8, 128, "q tiles")) return rc;
    if (int rc = make_map_2d(&map_kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.scal, width, rows, (uint64_t)a.scal_stride, 8, (uint32_t)Lq, "k / v tiles")) return rc;
    if (int rc = make_map_2d(&map_pts, kPtsBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, kPtsBf16 ? 2 : 4, a.pts, width, rows,
                             (uint64_t)a.pts_stride, 48, (uint32_t)L, "point records")) return rc;
    auto k1 = k_ipa_tc_pass1_pp<OutT, kPtsBf16>;
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("ipa_tc pass1 (ping-pong) smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return SE3_ECUDA; }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    const int n_items = a.sh.heads * a.sh.batch;
    const int grid = (n_items + 1) / 2 < sms ? (n_items + 1) / 2 : sms;
    k1<<<grid, kThreadsPP, smem, a.stream>>>(map_q, map_kv, map_pts, a.rot, a.trans, a.pair_bias, a.key_bias, a.head_weight, static_cast<OutT*>(a.out), a.pbuf,
                                              a.inv_sum, a.sh, pl, a.Bpad, pp_use_turns(), g_pp_dbg);
    count_launch();
    return check_launch("se3_ipa_attention_tc_fwd(pass 1, ping-pong)");
}

}  // namespace

namespace se3 {
namespace ipa_tc {

int launch_pass1_pingpong(const Pass1Args& a) {
    if (a.sh.len > kMaxLenPP) return SE3_EUNSUPPORTED;       // rows 96..127 belong to the issuer / helper warps
    if (a.out_bf16) return a.pts_bf16 ? launch_pp<__nv_bfloat16, true>(a) : launch_pp<__nv_bfloat16, false>(a);
    return a.pts_bf16 ? launch_pp<float, true>(a) : launch_pp<float, false>(a);
}

}  // namespace ipa_tc
}  // namespace se3

extern "C" {
/* developer hook (not in the public header): 32 clock64 slots per item of the ping-pong pass 1 (scripts/ipa_pp_phase_times.py) */
void se3_debug_set_pp_phase_buffer(long long* buf) { g_pp_dbg = buf; }
}
