// K4 (tensor-core edition) -- DiG invariant point attention on tcgen05 / TMEM, two passes.
//
// Pass 1, one work item per (sample b, head h, 128-query-row tile).  L <= 128: four persistent 128-thread CTAs per SM draw items
// from an atomic work queue (in the caller's workspace) and fetch the next item's operands under the current item's epilogue;
// longer sequences: one persistent 256-thread CTA (or 2-CTA cluster) per SM.  Per item:
//   stage  Q (x scalar_weight x log2e), K as bf16 UMMA operands; key points to the global frame (fp32, smem);
//          V^T = [v (16) | v_pt - c hi (24) | v_pt - c lo (24)] as the bf16 B operand of the second MMA
//          (global point coordinates re-centred on the sample's first residue and split hi + lo so the
//          fp32 aggregation demanded by structure_module.py:193-196 keeps ~16 mantissa bits)
//   MMA 1  S[128 x Lp] = Q.K^T                     tcgen05.mma kind::f16, fp32 accumulator in TMEM
//   SIMT   one thread per TMEM lane (= query row): S += head_w * sum_p |Qp_i - Kp_j| + pair_bias + key_bias
//          (the un-squared norm of structure_module.py:170 is not a contraction: 128 sqrt per (i,j), MUFU),
//          logits parked back in TMEM, row max, P = exp2(l - max) -> bf16 -> shared memory (A operand of
//          MMA 2) and -> global P[h][i][b][Lp] for pass 2
//   MMA 2  O[128 x 80] = P.[V | 1]                 accumulator reuses the TMEM columns of S; the ones column gives the row
//          sum of the rounded probabilities -> 1/rowsum -> inv[h][i][b]
//   SIMT   normalise, undo the re-centring, inverse frame R_i^T(. - T_i), norms; writes the scalar | point |
//          norm columns of the concat layout (structure_module.py:216)
// Sequences of 257..512 residues ("split" edition of pass 1): the keys of one (sample, head, query tile) are divided between the
// two CTAs of a thread-block cluster, each running the schedule above on its half.  They exchange the row maxima through
// distributed shared memory before the probabilities are formed (so both halves of P carry the same scale and pass 2 is
// unchanged), and rank 1 hands its partial O = P.[V | 1] to rank 0 the same way for the common epilogue.
// Pass 2, one CTA per (128-sample tile, query i, head h):
//   out_pair[b, i, h, :] = sum_j P[h,i,b,j] * pair_value[i,j,h,:]   (structure_module.py:209-213)
//   as a tensor-core GEMM with the SAMPLE index as M.  Pass 1 writes P row-major, bf16 [h][i][round_up(B,128)][Lp]: per (head,
//   query) a [samples][keys] matrix whose rows are written as whole 32-byte sectors.  Pass 2 fetches 128-sample x 64-key boxes of
//   it through a tensor map with CU_TENSOR_MAP_SWIZZLE_128B (UTMALDG) and consumes them through a SWIZZLE_128B K-major UMMA
//   descriptor; the pre-packed pair_value tile (16 x Lp bf16, se3_ipa_tc_pack_pair) is one bulk copy (UBLKCP); D[128 x 16]
//   accumulates in TMEM.  pair_value is therefore read once per 128 samples instead of once per sample (3.7 GB -> 30 MB per
//   layer at B=256, L=84).
// L <= 96 can also run pass 1 as the warp-specialised "ping-pong" edition of ipa_tc_pp.cu (SE3DIFF_B200_IPA_PP=1, experimental).
#include <cuda.h>
#include <math_constants.h>
#include <stdlib.h>

#include <type_traits>

#include "ipa_tc_shared.cuh"

using namespace se3;
using namespace se3::ipa_tc;

namespace {

// thread-block cluster primitives (split edition)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_peer_f32(float* own_smem_ptr, uint32_t peer_rank, float v) {   // same variable in the peer CTA
    uint32_t a = tc::smem_u32(own_smem_ptr), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(peer_rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(ra), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_peer_f32(const float* own_smem_ptr, uint32_t peer_rank) {   // same variable in the peer CTA
    uint32_t a = tc::smem_u32(own_smem_ptr), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(peer_rank));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

// Work queue of the persistent 128-thread edition: {items handed out beyond the first gridDim.x, CTAs that have left} in the
// 64 bytes that follow the row sums in the caller's workspace (se3_ipa_tc_workspace_bytes) -- per workspace, so concurrent
// calls on other streams (which need their own workspaces anyway) and graph replays cannot meet in it.  The bytes are zero
// before the first call (the caller's one-time memset, see the header) and the last CTA to leave zeroes them again (a memset
// node ahead of every launch instead was measured at +4 us per call inside a CUDA graph).
constexpr int64_t kQueueBytes = 64;

struct Pass1Smem {
    uint8_t *q, *k, *vs, *vp, *p;
    float *kp, *kb, *frm, *raw, *xo;
};
// Shared-memory plan of pass 1.  The P operand (written in the second half of the kernel) overlays everything that
// is dead by then -- the raw point records / bias slab, the key points, the key bias and the K operand -- when those
// fit into its 256*Lp bytes (`alias`).
// L = sequence length (query side), LK = key rows staged by the CTA (= L, or the cluster rank's share in the split edition),
// Lp = LK rounded up to 16
__host__ __device__ inline uint32_t bias_slab_bytes(int L, int LK) {
    const int lpi = (L + 7) & ~7;
    return (uint32_t)((LK * (lpi < 128 ? lpi : 128) * 2 + 15) & ~15);
}
__host__ __device__ inline uint32_t pass1_front_bytes(int L, int LK) {   // bias slab and (earlier) the raw local points share it
    const uint32_t raw = (uint32_t)LK * 48 * 4;
    const uint32_t m = bias_slab_bytes(L, LK) > raw ? bias_slab_bytes(L, LK) : raw;
    return (m + 127u) & ~127u;
}
__host__ __device__ inline bool pass1_can_alias(int L, int LK, int Lp) { return pass1_front_bytes(L, LK) + (uint32_t)Lp * (48 + 4 + 32) <= (uint32_t)Lp * 256; }
// The 256-thread editions (one CTA per SM) walk several items per CTA and prefetch the next item's inputs while the current
// item's probability tile is still draining from the P region: the raw point records get their own buffer instead of
// borrowing the front of the P region, and the split edition adds the [65][128] fp32 exchange buffer.
constexpr size_t kExchangeBytes = 65 * 128 * 4;
__host__ __device__ inline size_t pass1_smem_bytes(int L, int LK, int Lp, bool wide = false, bool split = false) {
    const size_t base = (size_t)Lp * (32 + NVP * 2 + 256) + 4096 + (pass1_can_alias(L, LK, Lp) ? 0 : (size_t)Lp * (48 + 4 + 32)) + (((size_t)LK * 48 + 127) & ~(size_t)127);
    return base + (wide ? (size_t)LK * 192 + 128 : 0) + (split ? kExchangeBytes : 0);
}
__device__ __forceinline__ Pass1Smem carve1(uint8_t* base, int L, int LK, int Lp, bool wide = false) {
    Pass1Smem s;
    s.vp = base;                                 // [Lp/8][NVP/8][8][8] bf16, MN-major point-value operand
    s.vs = s.vp + (size_t)Lp * (NVP * 2);        // [2][Lp][16 B] scalar values as they arrive (MN-major through the descriptor strides)
    s.q = s.vs + (size_t)Lp * 32;                // [2][128][16 B]
    s.p = s.q + 4096;                            // [Lp/8][128][16 B]; raw points, then the bias slab, live at its start until pass B
    uint8_t* rest = pass1_can_alias(L, LK, Lp) ? s.p + pass1_front_bytes(L, LK) : s.p + (size_t)Lp * 256;
    s.k = rest;                                  // [2][Lp][16 B]
    s.kp = reinterpret_cast<float*>(s.k + (size_t)Lp * 32);   // [Lp/2][12][2] fp32: negated global key points, interleaved by key pair
    s.kb = s.kp + Lp * 12;                       // [Lp]
    uint8_t* tail = pass1_can_alias(L, LK, Lp) ? s.p + (size_t)Lp * 256 : reinterpret_cast<uint8_t*>(s.kb + Lp);
    s.frm = reinterpret_cast<float*>(tail);      // rotations [LK][9] then translations [LK][3], as they lie in global memory
    // (TMA tile destinations need 128-byte alignment; `tail` is only 64-byte aligned when Lp is an odd multiple of 16)
    uint8_t* after = base + (((size_t)(tail - base) + (((size_t)LK * 48 + 127) & ~(size_t)127) + 127) & ~(size_t)127);
    s.raw = wide ? reinterpret_cast<float*>(after) : reinterpret_cast<float*>(s.p);
    s.xo = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s.raw) + (size_t)LK * 192);   // split edition only
    return s;
}

// kSplit = false: one CTA per (sample, head, query tile), all keys (L <= 256).   LpB = LpT = L rounded up to 16.
// kSplit = true : a cluster of two CTAs per (sample, head, query tile); rank 0 takes keys [0, LpB), rank 1 keys [LpB, LpT)
//                 (LpB = half of LpT rounded up to 16 = rows of every staging buffer; blockIdx.x = 2 * tile + rank).
// kWide (always with kSplit, and alone for 129..256 key rows, where shared memory admits one CTA per SM anyway): 256 threads;
//                 warps w and w + 4 share a TMEM lane quadrant (= 32 query rows) and take one half of the key columns each.
// kPtsBf16: the point records are bf16 (one projection GEMM writes scalar and point records side by side) instead of fp32.
// kPersist (always with kWide; alone = the 128-thread edition for L <= 128 with four persistent CTAs per SM): the CTA walks
//                 every gridDim.x-th item and issues the next item's copies as soon as the second product has consumed the
//                 operands.  Without it: one item per CTA, taken from the grid coordinates.
// kKeyBias = false: the caller passed no key bias (no padded / unknown residues: every BASELINE configuration) -- the logit pass then
//                 neither loads nor adds the staged zeros (2 LDS.128 + 4 FADD2 of the ~225 instructions of an 8-column step).
template <typename OutT, bool kSplit, bool kWide, bool kPtsBf16, bool kPersist = kWide, bool kKeyBias = true>
__global__ void __launch_bounds__(kWide ? 256 : 128, kWide ? 1 : 4)
k_ipa_tc_pass1(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_pts,
               const __grid_constant__ CUtensorMap map_bias, const float* __restrict__ rot, const float* __restrict__ trans, const __nv_bfloat16* __restrict__ pair_bias_t,
               const float* __restrict__ key_bias, const float* __restrict__ head_weight, OutT* __restrict__ out,
               __nv_bfloat16* __restrict__ pbuf, float* __restrict__ inv_sum, const se3_ipa_shape sh, int LpB, int LpT, int Bpad, int tmem_cols,
               const void* __restrict__ pts, int pts_stride, long long* __restrict__ dbg, int head0, unsigned int* __restrict__ sched) {
    static_assert(kPersist || !kWide, "the 256-thread editions are persistent");
    constexpr int kRawRow = kPtsBf16 ? 96 : 192;            // bytes of one staged point record [qp 12 | kp 12 | vp 24]
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t bar, bar_bias, bar_in;
    __shared__ float s_xmax[kSplit ? 128 : 1];             // row maxima offered to the peer CTA (split edition)
    __shared__ float s_hmax[kWide ? 256 : 1];              // row maxima of the two key-column halves (wide edition)
    static_assert(kWide || !kSplit, "the split edition runs 256 threads");
    constexpr int kThreads = kWide ? 256 : 128;
    __shared__ uint32_t tmem_slot;
    __shared__ int s_next;                                 // next item of this CTA (dynamic schedule)
    const int L = sh.len, H = sh.heads;
    const uint32_t rank = kSplit ? cluster_ctarank() : 0u;
    const int k0 = kSplit ? (int)rank * LpB : 0;                      // first key of this CTA
    const int Lp = kSplit ? (rank ? LpT - LpB : LpB) : LpB;           // its keys, padded to 16 ...
    const int LK = kSplit ? min(L - k0, Lp) : L;                      // ... of which real
    const int LKbox = kSplit ? LpB : L;                               // rows of the staged point records
    const Pass1Smem s = carve1(smem_raw, L, LKbox, LpB, kWide);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int qrow = kWide ? (tid & 127) : tid;            // query row of the tile = TMEM lane
    const int khalf = kWide ? (tid >> 7) : 0;              // which half of the key columns this thread walks in passes A and B
    const int ntile = (L + 127) >> 7;
    // Work items (sample b, head h, query tile).  Narrow edition: one per CTA, from the grid coordinates.  256-thread editions:
    // item = (b * H + h) * ntile + tile, this CTA (cluster) takes every `item_step`-th one starting at its own index, so that
    // the CTAs running at any moment work on neighbouring (sample, head) pairs whose record slices share DRAM pages.
    const int item_step = kPersist ? (int)(kSplit ? gridDim.x >> 1 : gridDim.x) : 1;
    const int item_first = kPersist ? (int)(kSplit ? blockIdx.x >> 1 : blockIdx.x) : 0;
    const int n_items = kPersist ? ntile * H * sh.batch : 1;
    int item = item_first;
    // optional phase timestamps: 16 clock64 slots per item, written by thread 0 (scripts/ipa_phase_times.py)
#define SE3_STAMP(k) do { if (dbg && threadIdx.x == 0) dbg[(kPersist ? (int64_t)item * (kSplit ? 2 : 1) + (int64_t)rank : ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (k)] = clock64(); } while (0)

    // pair-bias tile of this (head, query tile): bf16 [L keys][ncol queries], fetched by TMA into the region that
    // later holds P (P is only written after every warp has finished the logit pass)
    const int Lpi = ipa_bias_pitch(L);                    // row pitch of the packed bias matrix (common.cuh)
    // L <= 128: the whole QUERY-major [L][Lpi keys] matrix of the head, one bulk copy (the thread of a query row reads the eight
    // biases of a logit step as one 16-byte LDS).  Longer sequences: a [keys][128 queries] box of the 2-D tensor map over the
    // key-major matrix (columns past the matrix edge are zero-filled), slab pitch 128
    const int ncol = kWide ? 128 : Lpi;                   // multiple of 8 -> 16-byte rows
    const __nv_bfloat16* s_bias = reinterpret_cast<const __nv_bfloat16*>(s.p);
    // ---- staging: six TMA tile copies + two bulk copies per item, issued by one thread --------------------------------
    // scalars: bf16 head-major records [q 16 | k 16 | v 16] (q already carries scalar_weight * log2 e).  A 16-byte wide,
    //          R-row box of the 2-D tensor map lands as [R][16 B]: exactly one K-chunk of a UMMA operand, so q, k and v
    //          go from global memory into their operand tiles with no thread touching them.  Rows past the end of the
    //          matrix are zero-filled by the TMA unit; rows past this sample's L hold the next sample's (finite)
    //          values, which only ever meet logits forced to -inf / probabilities that are exactly zero.
    // points : fp32 records [qp 12 | kp 12 | vp 24] of this head, one [L][192 B] box, parked raw in the (still unused) P region
    //          (256-thread editions: in their own buffer)
    // frames : the [LK][9] and [LK][3] blocks of this CTA's key residues, two 1-D bulk copies
    const uint8_t* s_raw = reinterpret_cast<const uint8_t*>(s.raw);   // [LKbox][48] raw local points
    float* s_rot = s.frm;
    float* s_trn = s.frm + LK * 9;
    const bool bulk_frames = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(rot) | reinterpret_cast<uintptr_t>(trans)) & 15) == 0;
    auto decode = [&](const int it, int& b_, int& h_, int& q0_) {
        if constexpr (kPersist) {
            const int bh = it / ntile;
            b_ = bh / H; h_ = bh - b_ * H; q0_ = (it - bh * ntile) * 128;
        } else {
            b_ = blockIdx.z; h_ = blockIdx.y + head0; q0_ = blockIdx.x * 128;      // head0: first head of this launch's head group
        }
    };
    auto issue_loads = [&](const int it) {                 // thread 0 only
        int b_, h_, q0_;
        decode(it, b_, h_, q0_);
        const int row0 = b_ * L;
        tc::mbar_expect_tx(&bar_in, (uint32_t)(4096 + LpB * 64 + LKbox * kRawRow + (bulk_frames ? LK * 48 : 0)));
        tc::tma_tile_2d_g2s(s.raw, &map_pts, h_ * 48, row0 + k0, &bar_in);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            tc::tma_tile_2d_g2s(s.k + (size_t)half * LpB * 16, &map_kv, h_ * 48 + 16 + half * 8, row0 + k0, &bar_in);
            tc::tma_tile_2d_g2s(s.q + (size_t)half * 2048, &map_q, h_ * 48 + half * 8, row0 + q0_, &bar_in);
            tc::tma_tile_2d_g2s(s.vs + (size_t)half * LpB * 16, &map_kv, h_ * 48 + 32 + half * 8, row0 + k0, &bar_in);
        }
        if (bulk_frames) {
            tc::tma_bulk_g2s(s_rot, rot + ((int64_t)b_ * L + k0) * 9, (uint32_t)(LK * 36), &bar_in);
            tc::tma_bulk_g2s(s_trn, trans + ((int64_t)b_ * L + k0) * 3, (uint32_t)(LK * 12), &bar_in);
        }
    };
    if (warp == 0) tc::tmem_alloc(&tmem_slot, (uint32_t)tmem_cols);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_init(&bar_bias, 1);
        tc::mbar_init(&bar_in, 1);
        tc::mbar_fence_init();
        if (item < n_items) issue_loads(item);
    }
    // 128-thread persistent edition: items beyond the first gridDim.x are handed out by an atomic counter, one item ahead of
    // their use (the SMs do not run at the same pace: with a fixed round-robin schedule the slowest CTA took 13 % longer than
    // the average one and set the kernel's duration).  Consecutive tickets go to whichever CTAs ask next, so the CTAs running at
    // any moment still work on neighbouring (sample, head) records.
    constexpr bool kDynamic = kPersist && !kWide;
    int ticket = 0;                                        // thread 0: drawn early, published (s_next) at the end of the item
    auto fetch_next = [&]() {                              // thread 0 only
        if constexpr (kDynamic) ticket = (int)gridDim.x + (int)atomicAdd(&sched[0], 1u);
    };
    if (kDynamic && tid == 0) { fetch_next(); s_next = ticket; }
    // who issues the next item's copies and draws the next ticket: L <= 96 leaves the warp of lane quadrant 3 without query rows,
    // so its first thread does it while warps 0..2 are in the epilogue
    const int prefetcher = (kDynamic && L <= 96) ? 96 : 0;
    tc::fence_before();
    __syncthreads();   // the barriers are initialised, the TMEM base address is published
    tc::fence_after();
    const uint32_t tmem = tmem_slot;
    const int nchunk = Lp / 16;
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32;

  for (uint32_t par = 0; item < n_items; par ^= 1u) {
    SE3_STAMP(0);
    int item_next = item + item_step;
    if constexpr (kDynamic) item_next = s_next;            // written by thread 0 before the barrier that ended the previous item
    int b, h, q0;
    decode(item, b, h, q0);
    const int i = q0 + qrow;
    const bool row_ok = i < L;
    const bool warp_ok = q0 + (warp & 3) * 32 < L;
    if (!bulk_frames) {                                    // unaligned sample block: 4-byte asynchronous copies
        const float* rsrc = rot + ((int64_t)b * L + k0) * 9;
        const float* tsrc = trans + ((int64_t)b * L + k0) * 3;
        for (int idx = tid; idx < LK * 9; idx += kThreads) tc::cp_async4(s_rot + idx, rsrc + idx);
        for (int idx = tid; idx < LK * 3; idx += kThreads) tc::cp_async4(s_trn + idx, tsrc + idx);
        tc::cp_async_commit();
    }
    for (int j = tid; j < Lp; j += kThreads) s.kb[j] = (j < LK) ? (key_bias ? key_bias[(int64_t)b * L + k0 + j] * kLog2e : 0.f) : -CUDART_INF_F;
    SE3_STAMP(9);
    if (!bulk_frames) {
        tc::cp_async_wait<0>();
        __syncthreads();   // the fallback frame copies of every thread are done
    }
    SE3_STAMP(10);
    tc::mbar_wait(&bar_in, par);   // frames, raw points and the scalar operands are in shared memory
    SE3_STAMP(11);

    // ---- local -> global frame, one thread per residue -------------------------------------------------------------
    // key points  : negated, interleaved by key pair ([pair][component][2]) so that pass A forms q + (-k) for two keys
    //               with one packed add; lanes 2m / 2m+1 swap halves by shuffle and each writes 12 contiguous floats
    // value points: re-centred on the sample's first residue, split hi + lo bf16 (the fp32 aggregation demanded by
    //               structure_module.py:193-196 keeps ~16 mantissa bits), written as whole 16-byte operand chunks
    // ones column : channel 64 of V is 1 for real keys, so MMA 2 also returns the row sum of the ROUNDED probabilities
    const float* t0 = kSplit ? trans + (int64_t)b * L * 3 : s_trn;     // the sample's first residue (rank 1 does not stage it)
    const float cx = t0[0], cy = t0[1], cz = t0[2];
    for (int base = 0; base < Lp; base += kThreads) {
        if (base + warp * 32 >= Lp) break;                 // warp-uniform: the shuffles below need whole warps
        const int row = base + tid;
        float nk[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) nk[c] = 0.f;
        if (row < LK) {
            float R[9], T[3];
#pragma unroll
            for (int c = 0; c < 9; ++c) R[c] = s_rot[row * 9 + c];
#pragma unroll
            for (int c = 0; c < 3; ++c) T[c] = s_trn[row * 3 + c];
            const uint8_t* rawrow = s_raw + (size_t)row * kRawRow;
            {
                float l[12];
                load_coords<kPtsBf16, 12>(rawrow, 12, l);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    float gx, gy, gz;
                    to_global(R, T, l[3 * p], l[3 * p + 1], l[3 * p + 2], gx, gy, gz);
                    nk[3 * p] = -gx; nk[3 * p + 1] = -gy; nk[3 * p + 2] = -gz;
                }
            }
            T[0] -= cx; T[1] -= cy; T[2] -= cz;
            uint32_t hi[12], lo[12];
            float gv[24];
            {
                float l[24];
                load_coords<kPtsBf16, 24>(rawrow, 24, l);
#pragma unroll
                for (int p = 0; p < 8; ++p) to_global(R, T, l[3 * p], l[3 * p + 1], l[3 * p + 2], gv[3 * p], gv[3 * p + 1], gv[3 * p + 2]);
            }
#pragma unroll
            for (int c = 0; c < 12; ++c) {
                const __nv_bfloat162 hh = __floats2bfloat162_rn(gv[2 * c], gv[2 * c + 1]);
                hi[c] = *reinterpret_cast<const uint32_t*>(&hh);
                lo[c] = tc::pack_bf16(gv[2 * c] - __bfloat162float(hh.x), gv[2 * c + 1] - __bfloat162float(hh.y));
            }
            uint8_t* col = s.vp + (size_t)(row >> 3) * (NVP * 16) + (size_t)(row & 7) * 16;   // chunk g of this row at + g*128
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                *reinterpret_cast<uint4*>(col + g * 128) = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
                *reinterpret_cast<uint4*>(col + (3 + g) * 128) = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
            }
            *reinterpret_cast<uint4*>(col + 6 * 128) = make_uint4(0x00003F80u, 0, 0, 0);     // bf16 1.0 in channel 48 of the point operand
            *reinterpret_cast<uint4*>(col + 7 * 128) = make_uint4(0, 0, 0, 0);
        } else if (row < Lp) {                             // padding keys: the point operand must be exactly zero
            uint8_t* col = s.vp + (size_t)(row >> 3) * (NVP * 16) + (size_t)(row & 7) * 16;
#pragma unroll
            for (int g = 0; g < NVP / 8; ++g) *reinterpret_cast<uint4*>(col + g * 128) = make_uint4(0, 0, 0, 0);
        }
        {
            const bool odd = tid & 1;
            float mine[6], recv[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                mine[c] = odd ? nk[c + 6] : nk[c];
                recv[c] = __shfl_xor_sync(0xffffffffu, odd ? nk[c] : nk[c + 6], 1);
            }
            if (row < Lp) {
                float4* dst = reinterpret_cast<float4*>(s.kp + (size_t)(row >> 1) * 24 + (odd ? 12 : 0));
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    dst[c] = odd ? make_float4(recv[2 * c], mine[2 * c], recv[2 * c + 1], mine[2 * c + 1])
                                 : make_float4(mine[2 * c], recv[2 * c], mine[2 * c + 1], recv[2 * c + 1]);
            }
        }
    }
    float qp[12], Ri[9], Ti[3];
    {
        // the thread's own query residue: from the staged blocks, or (split edition: the CTA stages key residues only) from global
        const int qi = row_ok ? i : 0;
        const float* qrot = kSplit ? rot + ((int64_t)b * L + qi) * 9 : s_rot + qi * 9;
        const float* qtrn = kSplit ? trans + ((int64_t)b * L + qi) * 3 : s_trn + qi * 3;
#pragma unroll
        for (int c = 0; c < 9; ++c) Ri[c] = qrot[c];
#pragma unroll
        for (int c = 0; c < 3; ++c) Ti[c] = qtrn[c];
        const void* qrow = kSplit ? static_cast<const void*>(reinterpret_cast<const uint8_t*>(pts) + (((int64_t)b * L + qi) * pts_stride + h * 48) * (kPtsBf16 ? 2 : 4))
                                  : static_cast<const void*>(s_raw + (size_t)qi * kRawRow);
        float l[12];
        load_coords<kPtsBf16, 12>(qrow, 0, l);
#pragma unroll
        for (int p = 0; p < 4; ++p) to_global(Ri, Ti, l[3 * p], l[3 * p + 1], l[3 * p + 2], qp[3 * p], qp[3 * p + 1], qp[3 * p + 2]);
    }
    SE3_STAMP(12);
    SE3_STAMP(1);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();

    // ---- MMA 1: S = Q.K^T -----------------------------------------------------------------------------------
    if (tid == 0) {
        // the raw-point area is dead: fetch the pair-bias slab into it (TMA), it lands while the MMA runs
        const __nv_bfloat16* src = pair_bias_t + ((int64_t)h * L + k0) * Lpi + q0;
        if (ncol == Lpi) {
            tc::mbar_expect_tx(&bar_bias, (uint32_t)(LK * ncol * 2));
            tc::tma_bulk_g2s(s.p, src, (uint32_t)(LK * ncol * 2), &bar_bias);
        } else {   // one tile copy (a loop of per-key 256-byte bulk copies issued by this thread was measured at ~25k cycles)
            tc::mbar_expect_tx(&bar_bias, (uint32_t)(LKbox * 128 * 2));
            tc::tma_tile_2d_g2s(s.p, &map_bias, q0, h * L + k0, &bar_bias);
        }
        tc::mma_bf16(tmem, tc::make_desc(tc::smem_u32(s.q), 128), tc::make_desc(tc::smem_u32(s.k), (uint32_t)LpB),
                     tc::make_idesc_bf16(128, Lp), false);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after();
    SE3_STAMP(2);

    float m = -CUDART_INF_F;
    // 16-key column chunks of this thread: all of them, or one half each for the two threads of a row (wide edition)
    const int c_begin = khalf ? (nchunk + 1) / 2 : 0, c_end = (kWide && !khalf) ? (nchunk + 1) / 2 : nchunk;
    tc::mbar_wait(&bar_bias, par);
    if (warp_ok) {
        // ---- pass A: logits (log2 domain) -> TMEM, row max -------------------------------------------------
        const float hw = head_weight[h] * kLog2e;
        // 256-thread editions: [j][query] slab, conflict-free 2-byte LDS per key.  128-thread edition: this row's [keys] line
        const __nv_bfloat16* bias_col = kWide ? s_bias + min(qrow, ncol - 1) : s_bias + min(qrow, L - 1) * Lpi;
        // Packed fp32x2 arithmetic (FADD2/FMUL2/FFMA2, sm_100): two keys per instruction.  The key points are staged
        // NEGATED and interleaved by key pair ([pair][component][2]) so that q + (-k) is a single packed add.
        float2 q2[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) q2[k] = make_float2(qp[k], qp[k]);
        const float2 hw2 = make_float2(hw, hw), l2e2 = make_float2(kLog2e, kLog2e);
        // A step is 8 key columns (4 key pairs), ONE body walked by a rolled loop.  The kernel's loop is ~2400 instructions
        // (38 KB) for a 32 KB instruction cache per SM, the four resident CTAs sit in different phases, and ncu shows 7 % of the
        // instruction-line requests missing there and the GPC-level instruction cache behind it at 75 % of its request rate
        // (profiles/r3a_ipa_tc_ncu_raw.csv: sm__icc_request_hit_rate, gcc__cache_requests_type_instruction): every instruction
        // of this loop that exists twice costs.  Unrolling 16 columns bought no better schedule (two key pairs are in flight
        // either way, the register budget bounds that); the step that straddles the last key computes its padding columns like
        // real ones (their key points are zero) and overwrites them afterwards; steps past the last key are not computed.
        auto chunk = [&](const int col0) {
            uint32_t r[8];
            tc::tmem_ld8(tc::tmem_addr(tmem, lane_base, col0), r);
            uint4 pbq = make_uint4(0u, 0u, 0u, 0u);          // 128-thread edition: the step's eight pair biases, one LDS.128
            if constexpr (!kWide) pbq = *reinterpret_cast<const uint4*>(bias_col + col0);
            tc::tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = col0 + 2 * u;
                const float4* kp4 = reinterpret_cast<const float4*>(s.kp + j * 12);   // pair block: 24 floats
                float2 ds;                                                            // sum over the four points, started by the first
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    // components 3p, 3p+1, 3p+2 of the pair: float4 #(3p/2) ... laid out {c_j, c_j1, c'_j, c'_j1}
                    float2 nk[3];
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int comp = 3 * p + t;
                        const float4 v = kp4[comp >> 1];
                        nk[t] = (comp & 1) ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
                    }
                    const float2 dx = __fadd2_rn(q2[3 * p], nk[0]), dy = __fadd2_rn(q2[3 * p + 1], nk[1]), dz = __fadd2_rn(q2[3 * p + 2], nk[2]);
                    float2 d2 = __fmul2_rn(dx, dx);
                    d2 = __ffma2_rn(dy, dy, d2);
                    d2 = __ffma2_rn(dz, dz, d2);
                    const float2 dn = make_float2(fast_sqrt(d2.x), fast_sqrt(d2.y));
                    ds = p == 0 ? dn : __fadd2_rn(ds, dn);
                }
                float pb0, pb1;
                if constexpr (kWide) {
                    pb0 = __bfloat162float(bias_col[j * ncol]);
                    pb1 = __bfloat162float(bias_col[(j + 1) * ncol]);              // (past the last key: stale shared memory, overwritten below)
                } else {
                    const uint32_t w = u == 0 ? pbq.x : u == 1 ? pbq.y : u == 2 ? pbq.z : pbq.w;   // keys j | j + 1 (zero past the last key)
                    pb0 = __uint_as_float(w << 16);
                    pb1 = __uint_as_float(w & 0xffff0000u);
                }
                float2 l2 = __ffma2_rn(hw2, ds, make_float2(__uint_as_float(r[2 * u]), __uint_as_float(r[2 * u + 1])));
                l2 = __ffma2_rn(make_float2(pb0, pb1), l2e2, l2);
                if constexpr (kKeyBias) l2 = __fadd2_rn(l2, *reinterpret_cast<const float2*>(s.kb + j));
                r[2 * u] = __float_as_uint(l2.x);
                r[2 * u + 1] = __float_as_uint(l2.y);
            }
            if (col0 + 8 > LK) {                           // warp-uniform: padding keys of the straddling step
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (col0 + u >= LK) r[u] = __float_as_uint(-CUDART_INF_F);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) m = fmaxf(m, fmaxf(__uint_as_float(r[2 * u]), __uint_as_float(r[2 * u + 1])));
            tc::tmem_st8(tc::tmem_addr(tmem, lane_base, col0), r);
        };
#pragma unroll 1
        for (int col0 = c_begin * 16; col0 < c_end * 16; col0 += 8) {
            if (col0 < LK) {
                chunk(col0);
            } else {                                       // only padding keys
                uint32_t r[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) r[u] = __float_as_uint(-CUDART_INF_F);
                tc::tmem_st8(tc::tmem_addr(tmem, lane_base, col0), r);
            }
        }
        tc::tmem_wait_st();
    }
    if constexpr (kWide) {
        // the two column halves of a row were walked by different threads
        s_hmax[tid] = m;
        __syncthreads();
        m = fmaxf(s_hmax[qrow], s_hmax[128 + qrow]);
    }
    if constexpr (kSplit) {
        // the two key halves must form their probabilities against the same row maximum: swap maxima through DSMEM
        if (khalf == 0) s_xmax[qrow] = m;
        cluster_sync();
        m = fmaxf(m, ld_peer_f32(&s_xmax[qrow], rank ^ 1u));
    }
    if (m == -CUDART_INF_F) m = 0.f;
    __syncthreads();  // every warp is done with the bias tile: its shared memory becomes the P operand
    SE3_STAMP(3);
    if (warp_ok) {
        // ---- pass B: P = exp2(l - m) -> bf16 -> smem (A operand) + global (pass 2) ------------------------------
        // Probability workspace: row-major [h][i][b][LpT] bf16 -- per (head, query) a [samples][keys] matrix that pass 2 fetches
        // as 128-sample x 64-key swizzled TMA boxes.  This thread owns row (h, i, b): every 16-key chunk is ONE aligned 32-byte
        // store, i.e. whole L2 sectors (the earlier [j/8][b%128][8] operand layout scattered 16-byte half-sectors 2 KB apart:
        // 2x DRAM write amplification and a 9k-cycle drain per 128 x 256 tile, by per-thread stores and by a TMA tensor store alike).
        uint8_t* prow = reinterpret_cast<uint8_t*>(pbuf) + ((((int64_t)h * L + (row_ok ? i : 0)) * Bpad + b) * LpT + k0) * 2;
        for (int c = c_begin; c < c_end; ++c) {
            uint32_t r[16], pk[8];
            tc::tmem_ld16(tc::tmem_addr(tmem, lane_base, c * 16), r);
            tc::tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float2 d = __fadd2_rn(make_float2(__uint_as_float(r[2 * u]), __uint_as_float(r[2 * u + 1])), make_float2(-m, -m));   // one packed subtract per key pair
                pk[u] = tc::pack_bf16(fast_ex2(d.x), fast_ex2(d.y));
            }
            const uint4 lo = make_uint4(pk[0], pk[1], pk[2], pk[3]), hi = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            *reinterpret_cast<uint4*>(s.p + ((size_t)(2 * c) * 128 + qrow) * 16) = lo;
            *reinterpret_cast<uint4*>(s.p + ((size_t)(2 * c + 1) * 128 + qrow) * 16) = hi;
            if (row_ok)
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                             :: "l"(prow + (size_t)c * 32), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
        }
    }
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();

    SE3_STAMP(4);
    // ---- MMA 2: O = P.V (accumulator overwrites the consumed S columns) -------------------------------------------
    if (tid == 0) {
        // two MN-major B operands: the scalar values as TMA delivered them ([channel group][key][16 B]: 8 keys = 128 B apart,
        // channel groups Lp*16 B apart) -> columns 0..15; the point operand ([key group][channel group][8][8]) -> columns 16..79
        const uint32_t idesc_s = tc::make_idesc_bf16(128, DK, /*b_mn_major=*/true), idesc_p = tc::make_idesc_bf16(128, NVP, /*b_mn_major=*/true);
        for (int ks = 0; ks < nchunk; ++ks) {
            const uint64_t a_desc = tc::make_desc_kstep(tc::smem_u32(s.p), 128, ks);
            tc::mma_bf16(tmem, a_desc, tc::make_desc_raw(tc::smem_u32(s.vs) + (uint32_t)ks * 256u, /*K-group*/ 128u, /*MN-group*/ (uint32_t)LpB * 16u), idesc_s, ks > 0);
            tc::mma_bf16(tmem + DK, a_desc, tc::make_desc_raw(tc::smem_u32(s.vp) + (uint32_t)ks * 2u * NVP * 16u, /*K-group*/ NVP * 16u, /*MN-group*/ 128u), idesc_p, ks > 0);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 1);
    tc::fence_after();
    SE3_STAMP(5);
    if constexpr (kPersist) {
        // both products have consumed their shared-memory operands (the P operand too, whose front the 128-thread edition reuses
        // for the raw point records): the next item's copies fly under this item's epilogue
        if (tid == prefetcher && item_next < n_items) {
            issue_loads(item_next);
            fetch_next();                                  // the ticket after that one: its latency hides under the next item
        }
    }

    auto load_acc = [&](float (&o)[NV]) {
#pragma unroll
        for (int c = 0; c < NV / 16; ++c) {
            uint32_t r[16];
            tc::tmem_ld16(tc::tmem_addr(tmem, lane_base, c * 16), r);
            tc::tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 16; ++u) o[c * 16 + u] = __uint_as_float(r[u]);
        }
    };
    auto finish = [&](const float (&o)[NV]) {
        if (row_ok) {
            // column 64 = sum_j P_ij * 1 over the ROUNDED probabilities: the weights the tensor core applied sum to one
            // exactly, which the translation-covariant point aggregate needs
            const float inv = 1.0f / o[64];
            inv_sum[((int64_t)h * L + i) * Bpad + b] = inv;
            const int HD = H * DK;
            OutT* orow = out + ((int64_t)b * L + i) * (int64_t)(2 * HD + 4 * H * PV);
            float sc[DK], pl[3 * PV], nr[PV];
#pragma unroll
            for (int c = 0; c < DK; ++c) sc[c] = o[c] * inv;
#pragma unroll
            for (int p = 0; p < PV; ++p) {
                const float gx = (o[DK + p * 3] + o[DK + 3 * PV + p * 3]) * inv + (cx - Ti[0]);
                const float gy = (o[DK + p * 3 + 1] + o[DK + 3 * PV + p * 3 + 1]) * inv + (cy - Ti[1]);
                const float gz = (o[DK + p * 3 + 2] + o[DK + 3 * PV + p * 3 + 2]) * inv + (cz - Ti[2]);
                pl[p * 3] = Ri[0] * gx + Ri[3] * gy + Ri[6] * gz;
                pl[p * 3 + 1] = Ri[1] * gx + Ri[4] * gy + Ri[7] * gz;
                pl[p * 3 + 2] = Ri[2] * gx + Ri[5] * gy + Ri[8] * gz;
                nr[p] = sqrtf(pl[p * 3] * pl[p * 3] + pl[p * 3 + 1] * pl[p * 3 + 1] + pl[p * 3 + 2] * pl[p * 3 + 2]);
            }
            store_vec<DK>(orow + h * DK, sc);
            store_vec<3 * PV>(orow + HD + h * PV * 3, pl);
            store_vec<PV>(orow + 2 * HD + 3 * H * PV + h * PV, nr);
        }
    };
    if constexpr (kSplit) {
        // rank 1 pushes its partial accumulator (65 live columns: v 16 | points hi 24 | lo 24 | row sum) into rank 0's dead P
        // operand through DSMEM, [column][row] (conflict-free on both sides); rank 0 adds and finishes the rows
        constexpr int kLive = 65;
        float* s_xo = s.xo;
        cluster_sync();   // rank 0 has consumed the exchange buffer of the previous item
        if (rank == 1 && warp_ok && khalf == 0) {
            float o[NV];
            load_acc(o);
#pragma unroll
            for (int c = 0; c < kLive; ++c) st_peer_f32(&s_xo[c * 128 + qrow], 0u, o[c]);
        }
        cluster_sync();   // the pushed values are visible to rank 0
        if (rank == 0 && warp_ok && khalf == 0) {
            float o[NV];
            load_acc(o);
#pragma unroll
            for (int c = 0; c < kLive; ++c) o[c] += s_xo[c * 128 + qrow];
            finish(o);
        }
    } else {
        if (warp_ok && khalf == 0) {
            float o[NV];
            load_acc(o);
            finish(o);
        }
    }
    if (kDynamic && tid == prefetcher) s_next = ticket;
    tc::fence_before();
    __syncthreads();   // every warp has read its accumulator rows: TMEM and the operand buffers belong to the next item
    tc::fence_after();
    SE3_STAMP(6);
    SE3_STAMP(7);
    if (dbg && threadIdx.x == 0) {   // wall-clock time of the same instant: clock64 / globaltimer differences give the SM clock
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        dbg[(kPersist ? (int64_t)item * (kSplit ? 2 : 1) + (int64_t)rank : ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + 14] = gt;
    }
    if constexpr (!kPersist) break;
    item = item_next;
  }
    if constexpr (kDynamic) {
        if (tid == 0 && atomicAdd(&sched[1], 1u) == gridDim.x - 1) {   // last CTA out: the queue is zero again for the next call
            sched[0] = 0;
            sched[1] = 0;
        }
    }
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)tmem_cols);
#undef SE3_STAMP
}

// ---------------------------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(128)
k_ipa_tc_pass2(const __grid_constant__ CUtensorMap map_p, const float* __restrict__ inv_sum, const __nv_bfloat16* __restrict__ pvc,
               OutT* __restrict__ out, const se3_ipa_shape sh, int Lp, int Bpad, int head0) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int L = sh.len, H = sh.heads, B = sh.batch;
    const int bt = blockIdx.x, i = blockIdx.y, h = blockIdx.z + head0;
    const int tid = threadIdx.x, warp = tid >> 5;
    // A = probabilities of (h, i): rows = 128 samples, K = keys, fetched from the row-major workspace as 64-key boxes in the
    // 128-byte-swizzle operand layout (16 KB each, keys past LpT zero-filled by the TMA unit); B = pre-packed pair values
    const int nblk = (Lp + 63) >> 6;
    const uint32_t a_bytes = (uint32_t)nblk * 16384u, b_bytes = (uint32_t)Lp * 32u;
    uint8_t* sA = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // the swizzle pattern is tied to 1024-byte alignment
    uint8_t* sB = sA + a_bytes;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 32);
    if (tid == 0) {
        tc::mbar_init(&bar_tma, 1);
        tc::mbar_init(&bar_mma, 1);
        tc::mbar_fence_init();
        tc::mbar_expect_tx(&bar_tma, a_bytes + b_bytes);
        const int row0 = (h * L + i) * Bpad + bt * 128;
        for (int kb = 0; kb < nblk; ++kb) tc::tma_tile_2d_g2s(sA + (size_t)kb * 16384, &map_p, kb * 64, row0, &bar_tma);
        tc::tma_bulk_g2s(sB, reinterpret_cast<const uint8_t*>(pvc) + ((int64_t)i * H + h) * (int64_t)b_bytes, b_bytes, &bar_tma);
    }
    const int b = bt * 128 + tid;
    const float inv = (b < B) ? inv_sum[((int64_t)h * L + i) * Bpad + b] : 0.f;
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = tmem_slot;
    tc::mbar_wait(&bar_tma, 0);
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(128, DK);
        const uint32_t a_addr = tc::smem_u32(sA), b_addr = tc::smem_u32(sB);
        for (int ks = 0; ks < Lp / 16; ++ks)
            tc::mma_bf16(tmem, tc::make_desc_sw128(a_addr + (uint32_t)(ks >> 2) * 16384u, ks & 3), tc::make_desc_kstep(b_addr, DK, ks), idesc, ks > 0);
        tc::mma_commit(&bar_mma);
    }
    tc::mbar_wait(&bar_mma, 0);
    tc::fence_after();
    uint32_t r[16];
    tc::tmem_ld16(tc::tmem_addr(tmem, (uint32_t)warp * 32, 0), r);
    tc::tmem_wait_ld();
    if (b < B) {
        const int HD = H * DK;
        float v[DK];
#pragma unroll
        for (int c = 0; c < DK; ++c) v[c] = __uint_as_float(r[c]) * inv;
        store_vec<DK>(out + ((int64_t)b * L + i) * (int64_t)(2 * HD + 4 * H * PV) + HD + 3 * H * PV + h * DK, v);
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

long long* g_phase_dbg = nullptr;  // set by se3_debug_set_phase_buffer

bool pingpong_disabled() {   // work in progress: opt-in with SE3DIFF_B200_IPA_PP=1 until it beats the one-item-per-CTA edition
    const char* v = getenv("SE3DIFF_B200_IPA_PP");
    return !(v && v[0] == '1');
}

// SE3DIFF_B200_IPA_PERSIST=0: L <= 128 runs the one-item-per-CTA edition instead of the persistent one (measurement switch)
bool narrow_persistent() {
    static const bool v = [] { const char* e = getenv("SE3DIFF_B200_IPA_PERSIST"); return !(e && e[0] == '0'); }();
    return v;
}

template <typename OutT, bool kPtsBf16>
int launch_tc(const __nv_bfloat16* scal, int scal_stride, const void* pts, int pts_stride, const float* rot, const float* trans,
              const __nv_bfloat16* pair_bias, const __nv_bfloat16* pvc, const float* key_bias, const float* head_weight, OutT* out,
              __nv_bfloat16* pbuf, float* inv_sum, const se3_ipa_shape& sh, int Lp, int Bpad, cudaStream_t st) {
    const int L = sh.len;
    const bool split = Lp > 256;                           // keys divided between the two CTAs of a cluster
    const int LpB = split ? ((Lp / 2 + 15) & ~15) : Lp;    // key rows per CTA (rank 1 of a split gets Lp - LpB)
    const int LKbox = split ? LpB : L;
    int cols = 128;                                        // S needs LpB columns, the second accumulator NV = 80
    while (cols < LpB) cols *= 2;
    const bool wide = split || LpB > 128;                  // one CTA per SM by shared memory anyway: eight warps, several items per CTA
    size_t smem1 = pass1_smem_bytes(L, LKbox, LpB, wide, split);
    if (!split) {   // tensor memory (512 columns per SM) allows 512/cols resident CTAs; a CTA that is resident but blocked in
                    // tcgen05.alloc only steals issue slots, so shared memory is padded to admit exactly that many
        const size_t per_cta = (size_t)(227 * 1024) / (size_t)(512 / cols) - 1024;
        const size_t floor_bytes = (size_t)(227 * 1024) / (size_t)(512 / cols + 1) + 1;
        if (smem1 < floor_bytes && floor_bytes <= per_cta) smem1 = floor_bytes;
    }
    // tensor maps: 16-byte wide boxes of the scalar records (one UMMA K-chunk each), 192-byte wide boxes of the point records
    CUtensorMap map_q, map_kv, map_pts, map_bias, map_p;
    const uint64_t rows = (uint64_t)sh.batch * L, width = (uint64_t)sh.heads * 48;
    if (int rc = make_map_2d(&map_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, scal, width, rows, (uint64_t)scal_stride, 8, 128, "q tiles")) return rc;
    if (int rc = make_map_2d(&map_kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, scal, width, rows, (uint64_t)scal_stride, 8, (uint32_t)LpB, "k / v tiles")) return rc;
    if (int rc = make_map_2d(&map_pts, kPtsBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, kPtsBf16 ? 2 : 4, pts, width, rows,
                             (uint64_t)pts_stride, 48, (uint32_t)LKbox, "point records")) return rc;
    // probability workspace: row-major [h][i][b][Lp] = a [heads * L * Bpad][Lp] matrix; pass 2 takes 64-key x 128-sample boxes
    if (int rc = make_map_2d(&map_p, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pbuf, (uint64_t)Lp, (uint64_t)sh.heads * L * Bpad, (uint64_t)Lp, 64, 128,
                             "probability workspace", CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    const int ntile = (L + 127) / 128;
    const int Lpi = (L + 7) & ~7;
    if (Lpi > 128) {
        if (int rc = make_map_2d(&map_bias, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pair_bias, (uint64_t)Lpi, (uint64_t)sh.heads * L, (uint64_t)Lpi, 128,
                                 (uint32_t)LKbox, "pair-bias slabs")) return rc;
    } else {
        map_bias = map_q;                                  // unused by the kernel for L <= 128
    }
    cudaError_t e;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    const int n_items = ntile * sh.heads * sh.batch;
    // L <= 128: the warp-specialised ping-pong edition (ipa_tc_pp.cu) when its shared-memory plan fits; SE3DIFF_B200_IPA_PP=0 keeps
    // the one-item-per-CTA edition below (A/B timing)
    bool pass1_done = false;
    if (L <= 128 && !pingpong_disabled()) {
        Pass1Args a;
        a.scal = scal; a.scal_stride = scal_stride; a.pts = pts; a.pts_stride = pts_stride; a.pts_bf16 = kPtsBf16;
        a.rot = rot; a.trans = trans; a.pair_bias = pair_bias; a.key_bias = key_bias; a.head_weight = head_weight;
        a.out = out; a.out_bf16 = std::is_same<OutT, __nv_bfloat16>::value; a.pbuf = pbuf; a.inv_sum = inv_sum;
        a.sh = sh; a.Lp = Lp; a.Bpad = Bpad; a.stream = st;
        const int rc = launch_pass1_pingpong(a);
        if (rc == SE3_OK) pass1_done = true;
        else if (rc != SE3_EUNSUPPORTED) return rc;
    }
    const size_t smem2 = (size_t)((Lp + 63) / 64) * 16384 + (size_t)Lp * 32 + 1024;   // + slack for the 1024-byte alignment of the swizzled tiles
    auto k2 = k_ipa_tc_pass2<OutT>;
    e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) { set_error("ipa_tc pass2 smem attribute (%zu B): %s", smem2, cudaGetErrorString(e)); return SE3_ECUDA; }
    auto pass2 = [&](int head0, int nheads) {
        dim3 g2(Bpad / 128, L, nheads);
        k2<<<g2, 128, smem2, st>>>(map_p, inv_sum, pvc, out, sh, Lp, Bpad, head0);
        count_launch();
        return check_launch("se3_ipa_attention_tc_fwd(pass 2)");
    };
    if (pass1_done) {
        return pass2(0, sh.heads);
    } else if (!split) {
        auto k1 = wide ? (key_bias ? k_ipa_tc_pass1<OutT, false, true, kPtsBf16, true, true> : k_ipa_tc_pass1<OutT, false, true, kPtsBf16, true, false>)
                       : k_ipa_tc_pass1<OutT, false, false, kPtsBf16>;
        e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e != cudaSuccess) { set_error("ipa_tc pass1 smem attribute (%zu B): %s", smem1, cudaGetErrorString(e)); return SE3_ECUDA; }
        // One-item-per-CTA edition: optionally run pass 1 / pass 2 per GROUP of heads, so that pass 2 finds the group's probability
        // tiles (33 MB for 8 heads at L = 84, B = 256) still in the 126 MB L2 instead of reading them back from DRAM.
        // SE3DIFF_B200_IPA_HEAD_GROUP = heads per group (0 or unset: one group = all heads).
        int group = sh.heads;
        if (!wide) {
            const char* v = getenv("SE3DIFF_B200_IPA_HEAD_GROUP");
            const int gsz = v ? atoi(v) : 0;
            if (gsz > 0 && gsz < sh.heads) group = gsz;
        }
        if (!wide && group == sh.heads && narrow_persistent()) {     // four persistent CTAs per SM
            auto kp = key_bias ? k_ipa_tc_pass1<OutT, false, false, kPtsBf16, true, true> : k_ipa_tc_pass1<OutT, false, false, kPtsBf16, true, false>;
            e = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
            if (e != cudaSuccess) { set_error("ipa_tc pass1 smem attribute (%zu B): %s", smem1, cudaGetErrorString(e)); return SE3_ECUDA; }
            const int resident = (512 / cols) * sms;
            unsigned int* sched = reinterpret_cast<unsigned int*>(inv_sum + (int64_t)sh.heads * L * Bpad);
            kp<<<dim3((unsigned)(n_items < resident ? n_items : resident), 1, 1), 128, smem1, st>>>(map_q, map_kv, map_pts, map_bias, rot, trans, pair_bias, key_bias,
                                                                                                     head_weight, out, pbuf, inv_sum, sh, LpB, Lp, Bpad, cols, pts,
                                                                                                     pts_stride, g_phase_dbg, 0, sched);
            count_launch();
            if (int rc = check_launch("se3_ipa_attention_tc_fwd(pass 1)")) return rc;
            return pass2(0, sh.heads);
        }
        for (int h0 = 0; h0 < sh.heads; h0 += group) {
            const int nh = sh.heads - h0 < group ? sh.heads - h0 : group;
            const dim3 g1 = wide ? dim3((unsigned)(n_items < sms ? n_items : sms), 1, 1) : dim3(ntile, nh, sh.batch);
            k1<<<g1, wide ? 256 : 128, smem1, st>>>(map_q, map_kv, map_pts, map_bias, rot, trans, pair_bias, key_bias, head_weight, out, pbuf, inv_sum, sh, LpB, Lp, Bpad,
                                                    cols, pts, pts_stride, g_phase_dbg, h0, nullptr);
            count_launch();
            if (int rc = check_launch("se3_ipa_attention_tc_fwd(pass 1)")) return rc;
            if (int rc = pass2(h0, nh)) return rc;
        }
        return SE3_OK;
    } else {
        auto k1 = key_bias ? k_ipa_tc_pass1<OutT, true, true, kPtsBf16, true, true> : k_ipa_tc_pass1<OutT, true, true, kPtsBf16, true, false>;
        e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e != cudaSuccess) { set_error("ipa_tc pass1 (split) smem attribute (%zu B): %s", smem1, cudaGetErrorString(e)); return SE3_ECUDA; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (sms / 2), 1, 1);
        cfg.blockDim = dim3(256, 1, 1);
        cfg.dynamicSmemBytes = smem1;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        // persistent clusters: as many as can be co-resident (both CTAs of a cluster need SMs of one GPC)
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, k1, &cfg) != cudaSuccess || ncl < 1) { (void)cudaGetLastError(); ncl = sms / 2; }
        if (ncl > n_items) ncl = n_items;
        cfg.gridDim = dim3(2 * ncl, 1, 1);
        e = cudaLaunchKernelEx(&cfg, k1, map_q, map_kv, map_pts, map_bias, rot, trans, pair_bias, key_bias, head_weight, out, pbuf, inv_sum, sh, LpB, Lp, Bpad, cols,
                               pts, pts_stride, g_phase_dbg, 0, (unsigned int*)nullptr);
        if (e != cudaSuccess) { set_error("ipa_tc pass1 (split) launch: %s", cudaGetErrorString(e)); return SE3_ECUDA; }
        count_launch();
        if (int rc = check_launch("se3_ipa_attention_tc_fwd(pass 1)")) return rc;
        return pass2(0, sh.heads);
    }
}

}  // namespace

extern "C" {

/* developer hook (not in the public header): 8 clock64 stamps per pass-1 CTA are written to `buf` when non-null */
void se3_debug_set_phase_buffer(long long* buf) { g_phase_dbg = buf; }

int64_t se3_ipa_tc_workspace_bytes(const se3_ipa_shape* h_shape, int64_t* p_bytes, int64_t* inv_bytes) {
    if (!h_shape) return SE3_EINVAL;
    const int64_t Lp = (h_shape->len + 15) / 16 * 16, Bpad = (h_shape->batch + 127) / 128 * 128;
    const int64_t pb = (int64_t)h_shape->heads * h_shape->len * Bpad * Lp * 2, ib = (int64_t)h_shape->heads * h_shape->len * Bpad * 4 + kQueueBytes;
    if (p_bytes) *p_bytes = pb;
    if (inv_bytes) *inv_bytes = ib;
    return pb + ib;
}

int se3_ipa_attention_tc_fwd(const void* scalars_bf16, int64_t scalar_stride, const void* points, int points_are_bf16, int64_t point_stride,
                             const float* rot, const float* trans, const void* pair_bias_packed, const void* pair_value_packed,
                             const float* key_bias, const float* head_weight, void* out, int out_is_bf16, void* p_workspace,
                             float* inv_workspace, const se3_ipa_shape* h_shape, se3_stream_t stream) {
    SE3_REQUIRE(h_shape, "null shape");
    const se3_ipa_shape& sh = *h_shape;
    if (sh.batch == 0 || sh.len == 0) return SE3_OK;
    SE3_REQUIRE(scalars_bf16 && points && rot && trans && pair_bias_packed && pair_value_packed && head_weight && out && p_workspace && inv_workspace,
                "null pointer");
    if (sh.dk != DK || sh.pq != PQ || sh.pv != PV || sh.pair_batch != 1 || sh.len > 512 || sh.heads > 65535 || sh.batch > 65535) {
        set_error("se3_ipa_attention_tc_fwd: needs dk=16, 4/8 points, shared pair tensors, L <= 512 (got dk=%d L=%d H=%d pair_batch=%d); "
                  "use se3_ipa_attention_fwd", sh.dk, sh.len, sh.heads, sh.pair_batch);
        return SE3_EUNSUPPORTED;
    }
    SE3_REQUIRE(scalar_stride >= (int64_t)sh.heads * 48 && scalar_stride % 8 == 0 && scalar_stride < (1ll << 28) &&
                (reinterpret_cast<uintptr_t>(scalars_bf16) & 15) == 0, "scalar records: bf16 [rows][>= H*48], stride a multiple of 8, 16-byte aligned");
    SE3_REQUIRE(point_stride >= (int64_t)sh.heads * 48 && point_stride % (points_are_bf16 ? 8 : 4) == 0 && point_stride < (1ll << 28) &&
                (reinterpret_cast<uintptr_t>(points) & 15) == 0, "point records: [rows][>= H*48] fp32 (stride % 4) or bf16 (stride % 8), 16-byte aligned");
    const int Lp = (sh.len + 15) / 16 * 16, Bpad = (sh.batch + 127) / 128 * 128;
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16* sc = (const __nv_bfloat16*)scalars_bf16;
    const __nv_bfloat16 *pb = (const __nv_bfloat16*)pair_bias_packed, *pv = (const __nv_bfloat16*)pair_value_packed;
    __nv_bfloat16* pw = (__nv_bfloat16*)p_workspace;
    const int ss = (int)scalar_stride, ps = (int)point_stride;
    if (out_is_bf16) {
        __nv_bfloat16* o = (__nv_bfloat16*)out;
        return points_are_bf16 ? launch_tc<__nv_bfloat16, true>(sc, ss, points, ps, rot, trans, pb, pv, key_bias, head_weight, o, pw, inv_workspace, sh, Lp, Bpad, st)
                               : launch_tc<__nv_bfloat16, false>(sc, ss, points, ps, rot, trans, pb, pv, key_bias, head_weight, o, pw, inv_workspace, sh, Lp, Bpad, st);
    }
    float* o = (float*)out;
    return points_are_bf16 ? launch_tc<float, true>(sc, ss, points, ps, rot, trans, pb, pv, key_bias, head_weight, o, pw, inv_workspace, sh, Lp, Bpad, st)
                           : launch_tc<float, false>(sc, ss, points, ps, rot, trans, pb, pv, key_bias, head_weight, o, pw, inv_workspace, sh, Lp, Bpad, st);
}

}  // extern "C"
