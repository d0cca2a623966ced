// K4 (tensor-core edition) -- DiG invariant point attention on tcgen05 / TMEM, two passes.
//
// Pass 1, one CTA per (sample b, head h, 128-query-row tile):
//   stage  Q (x scalar_weight x log2e), K as bf16 UMMA operands; key points to the global frame (fp32, smem);
//          V^T = [v (16) | v_pt - c hi (24) | v_pt - c lo (24)] as the bf16 B operand of the second MMA
//          (global point coordinates re-centred on the sample's first residue and split hi + lo so the
//          fp32 aggregation demanded by structure_module.py:193-196 keeps ~16 mantissa bits)
//   MMA 1  S[128 x Lp] = Q.K^T                     tcgen05.mma kind::f16, fp32 accumulator in TMEM
//   SIMT   one thread per TMEM lane (= query row): S += head_w * sum_p |Qp_i - Kp_j| + pair_bias + key_bias
//          (the un-squared norm of structure_module.py:170 is not a contraction: 128 sqrt per (i,j), MUFU),
//          logits parked back in TMEM, row max, P = exp2(l - max) -> bf16 -> shared memory (A operand of
//          MMA 2) and -> global P[h][i][b][Lp] for pass 2, 1/rowsum -> inv[h][i][b]
//   MMA 2  O[128 x 64] = P.V                       accumulator reuses the TMEM columns of S
//   SIMT   normalise, undo the re-centring, inverse frame R_i^T(. - T_i), norms; writes the scalar | point |
//          norm columns of the concat layout (structure_module.py:216)
// Pass 2, one CTA per (128-sample tile, query i, 8-head group):
//   out_pair[b, i, h, :] = sum_j P[h,i,b,j] * pair_value[i,j,h,:]   (structure_module.py:209-213)
//   as a tensor-core GEMM with the SAMPLE index as M: A = P[h][i][b0:b0+128][Lp] (cp.async into the UMMA
//   layout, double buffered), B = pair_value pre-packed per (i,h) in UMMA layout, D[128 x 16] per head in TMEM.
//   pair_value is therefore read once per 128 samples instead of once per sample (3.7 GB -> 30 MB per layer
//   at B=256, L=84).
#include <math_constants.h>

#include "common.cuh"
#include "tc_common.cuh"

using namespace se3;

namespace {

constexpr int DK = 16, PQ = 4, PV = 8;
constexpr int NV = DK + 2 * 3 * PV;  // 64 columns of the value operand
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <typename T> __device__ __forceinline__ T to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

struct Pass1Smem {
    uint8_t *q, *k, *vt, *p;
    float *kp, *kb;
};
__device__ __forceinline__ Pass1Smem carve1(uint8_t* base, int Lp) {
    Pass1Smem s;
    s.q = base;                                  // [2][128][16 B]
    s.k = s.q + 2 * 128 * 16;                    // [2][Lp][16 B]
    s.vt = s.k + 2 * Lp * 16;                    // [Lp/8][64][16 B]
    s.p = s.vt + (Lp / 8) * NV * 16;             // [Lp/8][128][16 B]
    s.kp = reinterpret_cast<float*>(s.p + (Lp / 8) * 128 * 16);  // [Lp][12]
    s.kb = s.kp + Lp * 12;                       // [Lp]
    return s;
}
inline size_t pass1_smem_bytes(int Lp) { return 2 * 128 * 16 + (size_t)Lp * (32 + 128 + 256 + 48 + 4); }

template <typename OutT>
__global__ void __launch_bounds__(128)
k_ipa_tc_pass1(const float* __restrict__ proj, const float* __restrict__ rot, const float* __restrict__ trans,
               const float* __restrict__ pair_bias, const float* __restrict__ key_bias, const float* __restrict__ head_weight,
               float scalar_weight, OutT* __restrict__ out, __nv_bfloat16* __restrict__ pbuf, float* __restrict__ inv_sum,
               const se3_ipa_shape sh, int Lp, int Bpad, int tmem_cols) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const Pass1Smem s = carve1(smem_raw, Lp);
    const int L = sh.len, H = sh.heads;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int i = q0 + tid;
    const bool row_ok = i < L;
    const bool warp_ok = q0 + warp * 32 < L;

    if (warp == 0) tc::tmem_alloc(&tmem_slot, (uint32_t)tmem_cols);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }

    const float cx = trans[(int64_t)b * L * 3], cy = trans[(int64_t)b * L * 3 + 1], cz = trans[(int64_t)b * L * 3 + 2];

    // ---- stage keys / values of (b, h) --------------------------------------------------------------
    for (int j = tid; j < Lp; j += 128) {
        uint4 k0 = make_uint4(0, 0, 0, 0), k1 = k0;
        float kpg[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) kpg[c] = 0.f;
        __nv_bfloat16* vt_col = reinterpret_cast<__nv_bfloat16*>(s.vt + (size_t)(j >> 3) * NV * 16) + (j & 7);  // + row*8
        if (j < L) {
            const int64_t rj = (int64_t)b * L + j;
            const float* pr = proj + rj * sh.proj_stride;
            const float4* kq = reinterpret_cast<const float4*>(pr + sh.off_k + h * DK);
            const float4 a0 = __ldg(kq), a1 = __ldg(kq + 1), a2 = __ldg(kq + 2), a3 = __ldg(kq + 3);
            k0 = make_uint4(tc::pack_bf16(a0.x, a0.y), tc::pack_bf16(a0.z, a0.w), tc::pack_bf16(a1.x, a1.y), tc::pack_bf16(a1.z, a1.w));
            k1 = make_uint4(tc::pack_bf16(a2.x, a2.y), tc::pack_bf16(a2.z, a2.w), tc::pack_bf16(a3.x, a3.y), tc::pack_bf16(a3.z, a3.w));
            const float4* vq = reinterpret_cast<const float4*>(pr + sh.off_v + h * DK);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 v = __ldg(vq + c4);
                vt_col[(c4 * 4 + 0) * 8] = __float2bfloat16_rn(v.x);
                vt_col[(c4 * 4 + 1) * 8] = __float2bfloat16_rn(v.y);
                vt_col[(c4 * 4 + 2) * 8] = __float2bfloat16_rn(v.z);
                vt_col[(c4 * 4 + 3) * 8] = __float2bfloat16_rn(v.w);
            }
            float R[9], T[3];
#pragma unroll
            for (int c = 0; c < 9; ++c) R[c] = __ldg(rot + rj * 9 + c);
#pragma unroll
            for (int c = 0; c < 3; ++c) T[c] = __ldg(trans + rj * 3 + c);
            const float* kpl = pr + sh.off_kp + h * PQ * 3;
#pragma unroll
            for (int p = 0; p < PQ; ++p) {
                const float x = __ldg(kpl + p * 3), y = __ldg(kpl + p * 3 + 1), z = __ldg(kpl + p * 3 + 2);
#pragma unroll
                for (int r = 0; r < 3; ++r) kpg[p * 3 + r] = R[r * 3] * x + R[r * 3 + 1] * y + R[r * 3 + 2] * z + T[r];
            }
            const float* vpl = pr + sh.off_vp + h * PV * 3;
            const float cc[3] = {cx, cy, cz};
#pragma unroll
            for (int p = 0; p < PV; ++p) {
                const float x = __ldg(vpl + p * 3), y = __ldg(vpl + p * 3 + 1), z = __ldg(vpl + p * 3 + 2);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float g = R[r * 3] * x + R[r * 3 + 1] * y + R[r * 3 + 2] * z + (T[r] - cc[r]);
                    const __nv_bfloat16 hi = __float2bfloat16_rn(g);
                    vt_col[(DK + p * 3 + r) * 8] = hi;
                    vt_col[(DK + 3 * PV + p * 3 + r) * 8] = __float2bfloat16_rn(g - __bfloat162float(hi));
                }
            }
        } else {
#pragma unroll 8
            for (int c = 0; c < NV; ++c) vt_col[c * 8] = __float2bfloat16_rn(0.f);
        }
        *reinterpret_cast<uint4*>(s.k + (size_t)j * 16) = k0;
        *reinterpret_cast<uint4*>(s.k + (size_t)(Lp + j) * 16) = k1;
#pragma unroll
        for (int c4 = 0; c4 < 3; ++c4) reinterpret_cast<float4*>(s.kp + j * 12)[c4] = make_float4(kpg[c4 * 4], kpg[c4 * 4 + 1], kpg[c4 * 4 + 2], kpg[c4 * 4 + 3]);
        s.kb[j] = (j < L) ? (key_bias ? key_bias[(int64_t)b * L + j] * kLog2e : 0.f) : -CUDART_INF_F;
    }
    // ---- this thread's query row -----------------------------------------------------------------------
    float qp[12], Ri[9], Ti[3];
    {
        uint4 q0v = make_uint4(0, 0, 0, 0), q1v = q0v;
        const int64_t ri = (int64_t)b * L + (row_ok ? i : 0);
        const float* pr = proj + ri * sh.proj_stride;
#pragma unroll
        for (int c = 0; c < 9; ++c) Ri[c] = __ldg(rot + ri * 9 + c);
#pragma unroll
        for (int c = 0; c < 3; ++c) Ti[c] = __ldg(trans + ri * 3 + c);
        const float* qpl = pr + sh.off_qp + h * PQ * 3;
#pragma unroll
        for (int p = 0; p < PQ; ++p) {
            const float x = __ldg(qpl + p * 3), y = __ldg(qpl + p * 3 + 1), z = __ldg(qpl + p * 3 + 2);
#pragma unroll
            for (int r = 0; r < 3; ++r) qp[p * 3 + r] = Ri[r * 3] * x + Ri[r * 3 + 1] * y + Ri[r * 3 + 2] * z + Ti[r];
        }
        if (row_ok) {
            const float sc = scalar_weight * kLog2e;
            const float4* qq = reinterpret_cast<const float4*>(pr + sh.off_q + h * DK);
            const float4 a0 = __ldg(qq), a1 = __ldg(qq + 1), a2 = __ldg(qq + 2), a3 = __ldg(qq + 3);
            q0v = make_uint4(tc::pack_bf16(a0.x * sc, a0.y * sc), tc::pack_bf16(a0.z * sc, a0.w * sc), tc::pack_bf16(a1.x * sc, a1.y * sc),
                             tc::pack_bf16(a1.z * sc, a1.w * sc));
            q1v = make_uint4(tc::pack_bf16(a2.x * sc, a2.y * sc), tc::pack_bf16(a2.z * sc, a2.w * sc), tc::pack_bf16(a3.x * sc, a3.y * sc),
                             tc::pack_bf16(a3.z * sc, a3.w * sc));
        }
        *reinterpret_cast<uint4*>(s.q + (size_t)tid * 16) = q0v;
        *reinterpret_cast<uint4*>(s.q + (size_t)(128 + tid) * 16) = q1v;
    }
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = tmem_slot;

    // ---- MMA 1: S = Q.K^T -----------------------------------------------------------------------------------
    if (tid == 0) {
        tc::mma_bf16(tmem, tc::make_desc(tc::smem_u32(s.q), 128), tc::make_desc(tc::smem_u32(s.k), (uint32_t)Lp),
                     tc::make_idesc_bf16(128, Lp), false);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after();

    const int nchunk = Lp / 16;
    const uint32_t lane_base = (uint32_t)warp * 32;
    float inv = 0.f;
    if (warp_ok) {
        // ---- pass A: logits (log2 domain) -> TMEM, row max -------------------------------------------------
        const float hw = head_weight[h] * kLog2e;
        const float* bias_row = pair_bias + ((int64_t)h * L + (row_ok ? i : 0)) * L;
        float m = -CUDART_INF_F;
        for (int c = 0; c < nchunk; ++c) {
            uint32_t r[16];
            tc::tmem_ld16(tc::tmem_addr(tmem, lane_base, c * 16), r);
            tc::tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int j = c * 16 + u;
                const float4* kp4 = reinterpret_cast<const float4*>(s.kp + j * 12);
                const float4 k0 = kp4[0], k1 = kp4[1], k2 = kp4[2];
                float dx, dy, dz, ds;
                dx = qp[0] - k0.x; dy = qp[1] - k0.y; dz = qp[2] - k0.z;
                ds = fast_sqrt(dx * dx + dy * dy + dz * dz);
                dx = qp[3] - k0.w; dy = qp[4] - k1.x; dz = qp[5] - k1.y;
                ds += fast_sqrt(dx * dx + dy * dy + dz * dz);
                dx = qp[6] - k1.z; dy = qp[7] - k1.w; dz = qp[8] - k2.x;
                ds += fast_sqrt(dx * dx + dy * dy + dz * dz);
                dx = qp[9] - k2.y; dy = qp[10] - k2.z; dz = qp[11] - k2.w;
                ds += fast_sqrt(dx * dx + dy * dy + dz * dz);
                const float pb = (j < L) ? __ldg(bias_row + j) : 0.f;
                const float l2 = fmaf(pb, kLog2e, fmaf(hw, ds, __uint_as_float(r[u]))) + s.kb[j];
                m = fmaxf(m, l2);
                r[u] = __float_as_uint(l2);
            }
            tc::tmem_st16(tc::tmem_addr(tmem, lane_base, c * 16), r);
        }
        tc::tmem_wait_st();
        if (m == -CUDART_INF_F) m = 0.f;
        // ---- pass B: P = exp2(l - m) -> bf16 -> smem (A operand) + global (pass 2) ------------------------------
        float sum = 0.f;
        __nv_bfloat16* prow = pbuf + (((int64_t)h * L + (row_ok ? i : 0)) * Bpad + b) * Lp;
        for (int c = 0; c < nchunk; ++c) {
            uint32_t r[16], pk[8];
            tc::tmem_ld16(tc::tmem_addr(tmem, lane_base, c * 16), r);
            tc::tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float p0 = fast_ex2(__uint_as_float(r[2 * u]) - m), p1 = fast_ex2(__uint_as_float(r[2 * u + 1]) - m);
                const __nv_bfloat162 v = __floats2bfloat162_rn(p0, p1);
                sum += __bfloat162float(v.x) + __bfloat162float(v.y);
                pk[u] = *reinterpret_cast<const uint32_t*>(&v);
            }
            const uint4 lo = make_uint4(pk[0], pk[1], pk[2], pk[3]), hi = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            *reinterpret_cast<uint4*>(s.p + ((size_t)(2 * c) * 128 + tid) * 16) = lo;
            *reinterpret_cast<uint4*>(s.p + ((size_t)(2 * c + 1) * 128 + tid) * 16) = hi;
            if (row_ok) {
                reinterpret_cast<uint4*>(prow + c * 16)[0] = lo;
                reinterpret_cast<uint4*>(prow + c * 16)[1] = hi;
            }
        }
        inv = 1.0f / sum;
        if (row_ok) inv_sum[((int64_t)h * L + i) * Bpad + b] = inv;
    }
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();

    // ---- MMA 2: O = P.V (accumulator overwrites the consumed S columns) -------------------------------------------
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(128, NV);
        for (int ks = 0; ks < nchunk; ++ks)
            tc::mma_bf16(tmem, tc::make_desc_kstep(tc::smem_u32(s.p), 128, ks), tc::make_desc_kstep(tc::smem_u32(s.vt), NV, ks), idesc, ks > 0);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 1);
    tc::fence_after();

    if (warp_ok) {
        float o[NV];
#pragma unroll
        for (int c = 0; c < NV / 16; ++c) {
            uint32_t r[16];
            tc::tmem_ld16(tc::tmem_addr(tmem, lane_base, c * 16), r);
            tc::tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < 16; ++u) o[c * 16 + u] = __uint_as_float(r[u]);
        }
        if (row_ok) {
            const int HD = H * DK;
            OutT* orow = out + ((int64_t)b * L + i) * (int64_t)(2 * HD + 4 * H * PV);
#pragma unroll
            for (int c = 0; c < DK; ++c) orow[h * DK + c] = to_out<OutT>(o[c] * inv);
#pragma unroll
            for (int p = 0; p < PV; ++p) {
                const float gx = (o[DK + p * 3] + o[DK + 3 * PV + p * 3]) * inv + (cx - Ti[0]);
                const float gy = (o[DK + p * 3 + 1] + o[DK + 3 * PV + p * 3 + 1]) * inv + (cy - Ti[1]);
                const float gz = (o[DK + p * 3 + 2] + o[DK + 3 * PV + p * 3 + 2]) * inv + (cz - Ti[2]);
                const float lx = Ri[0] * gx + Ri[3] * gy + Ri[6] * gz;
                const float ly = Ri[1] * gx + Ri[4] * gy + Ri[7] * gz;
                const float lz = Ri[2] * gx + Ri[5] * gy + Ri[8] * gz;
                orow[HD + (h * PV + p) * 3] = to_out<OutT>(lx);
                orow[HD + (h * PV + p) * 3 + 1] = to_out<OutT>(ly);
                orow[HD + (h * PV + p) * 3 + 2] = to_out<OutT>(lz);
                orow[2 * HD + 3 * H * PV + h * PV + p] = to_out<OutT>(sqrtf(lx * lx + ly * ly + lz * lz));
            }
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int HG = 8;  // heads per pass-2 CTA -> 128 TMEM columns

template <typename OutT>
__global__ void __launch_bounds__(128)
k_ipa_tc_pass2(const __nv_bfloat16* __restrict__ pbuf, const float* __restrict__ inv_sum, const __nv_bfloat16* __restrict__ pvc,
               OutT* __restrict__ out, const se3_ipa_shape sh, int Lp, int Bpad) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t bar[3];
    __shared__ uint32_t tmem_slot;
    const int L = sh.len, H = sh.heads, B = sh.batch;
    const int b0 = blockIdx.x * 128, i = blockIdx.y, h0 = blockIdx.z * HG;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t a_bytes = (size_t)Lp * 256, b_bytes = (size_t)Lp * 32, stage_bytes = a_bytes + b_bytes;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 128);
    if (tid == 0) { tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1); tc::mbar_init(&bar[2], 1); tc::mbar_fence_init(); }

    auto load_head = [&](int hl, int stage) {
        uint8_t* sA = smem_raw + stage * stage_bytes;
        uint8_t* sB = sA + a_bytes;
        const int h = h0 + hl;
        const __nv_bfloat16* src = pbuf + (((int64_t)h * L + i) * Bpad + b0) * Lp;
        // a warp instruction covers 16 rows x 2 K-chunks: 32-byte global sectors, <= 2-way smem conflicts
        const int r16 = lane & 15, kcs = lane >> 4;
        for (int kcp = 0; kcp < Lp / 16; ++kcp)
#pragma unroll
            for (int rb = 0; rb < 2; ++rb) {
                const int row = (rb * 4 + warp) * 16 + r16, kc = kcp * 2 + kcs;
                tc::cp_async16(sA + ((size_t)kc * 128 + row) * 16, src + (size_t)row * Lp + kc * 8);
            }
        const uint8_t* bsrc = reinterpret_cast<const uint8_t*>(pvc + ((int64_t)i * H + h) * Lp * 16);
        for (int p = tid; p < (int)(b_bytes / 16); p += 128) tc::cp_async16(sB + (size_t)p * 16, bsrc + (size_t)p * 16);
        tc::cp_async_commit();
    };

    load_head(0, 0);
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = tc::make_idesc_bf16(128, DK);
    for (int hl = 0; hl < HG; ++hl) {
        const int stage = hl & 1;
        if (hl + 1 < HG) {
            if (hl >= 1) tc::mbar_wait(&bar[stage ^ 1], ((hl - 1) >> 1) & 1);  // MMAs of head hl-1 have drained that stage
            load_head(hl + 1, stage ^ 1);
            tc::cp_async_wait<1>();
        } else {
            tc::cp_async_wait<0>();
        }
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after();
            const uint32_t a_addr = tc::smem_u32(smem_raw + stage * stage_bytes), b_addr = a_addr + (uint32_t)a_bytes;
            for (int ks = 0; ks < Lp / 16; ++ks)
                tc::mma_bf16(tmem + hl * DK, tc::make_desc_kstep(a_addr, 128, ks), tc::make_desc_kstep(b_addr, DK, ks), idesc, ks > 0);
            tc::mma_commit(&bar[stage]);
            if (hl == HG - 1) tc::mma_commit(&bar[2]);
        }
    }
    tc::mbar_wait(&bar[2], 0);
    tc::fence_after();
    const int b = b0 + tid;
    const int HD = H * DK;
#pragma unroll 1
    for (int hl = 0; hl < HG; ++hl) {
        uint32_t r[16];
        tc::tmem_ld16(tc::tmem_addr(tmem, (uint32_t)warp * 32, hl * DK), r);
        tc::tmem_wait_ld();
        if (b < B) {
            const int h = h0 + hl;
            const float inv = inv_sum[((int64_t)h * L + i) * Bpad + b];
            OutT* o = out + ((int64_t)b * L + i) * (int64_t)(2 * HD + 4 * H * PV) + HD + 3 * H * PV + h * DK;
#pragma unroll
            for (int c = 0; c < DK; ++c) o[c] = to_out<OutT>(__uint_as_float(r[c]) * inv);
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

template <typename OutT>
int launch_tc(const float* proj, const float* rot, const float* trans, const float* pair_bias, const __nv_bfloat16* pvc,
              const float* key_bias, const float* head_weight, float scalar_weight, OutT* out, __nv_bfloat16* pbuf, float* inv_sum,
              const se3_ipa_shape& sh, int Lp, int Bpad, cudaStream_t st) {
    const int L = sh.len;
    int cols = 64;
    while (cols < Lp) cols *= 2;
    const size_t smem1 = pass1_smem_bytes(Lp);
    auto k1 = k_ipa_tc_pass1<OutT>;
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e != cudaSuccess) { set_error("ipa_tc pass1 smem attribute (%zu B): %s", smem1, cudaGetErrorString(e)); return SE3_ECUDA; }
    dim3 g1((L + 127) / 128, sh.heads, sh.batch);
    k1<<<g1, 128, smem1, st>>>(proj, rot, trans, pair_bias, key_bias, head_weight, scalar_weight, out, pbuf, inv_sum, sh, Lp, Bpad, cols);
    count_launch();
    int rc = check_launch("se3_ipa_attention_tc_fwd(pass 1)");
    if (rc) return rc;
    const size_t smem2 = 2 * ((size_t)Lp * 256 + (size_t)Lp * 32);
    auto k2 = k_ipa_tc_pass2<OutT>;
    e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) { set_error("ipa_tc pass2 smem attribute (%zu B): %s", smem2, cudaGetErrorString(e)); return SE3_ECUDA; }
    dim3 g2(Bpad / 128, L, sh.heads / HG);
    k2<<<g2, 128, smem2, st>>>(pbuf, inv_sum, pvc, out, sh, Lp, Bpad);
    count_launch();
    return check_launch("se3_ipa_attention_tc_fwd(pass 2)");
}

}  // namespace

extern "C" {

int64_t se3_ipa_tc_workspace_bytes(const se3_ipa_shape* h_shape, int64_t* p_bytes, int64_t* inv_bytes) {
    if (!h_shape) return SE3_EINVAL;
    const int64_t Lp = (h_shape->len + 15) / 16 * 16, Bpad = (h_shape->batch + 127) / 128 * 128;
    const int64_t pb = (int64_t)h_shape->heads * h_shape->len * Bpad * Lp * 2, ib = (int64_t)h_shape->heads * h_shape->len * Bpad * 4;
    if (p_bytes) *p_bytes = pb;
    if (inv_bytes) *inv_bytes = ib;
    return pb + ib;
}

int se3_ipa_attention_tc_fwd(const float* proj, const float* rot, const float* trans, const float* pair_bias,
                             const void* pair_value_packed, const float* key_bias, const float* head_weight, float scalar_weight,
                             void* out, int out_is_bf16, void* p_workspace, float* inv_workspace, const se3_ipa_shape* h_shape,
                             se3_stream_t stream) {
    SE3_REQUIRE(h_shape, "null shape");
    const se3_ipa_shape& sh = *h_shape;
    if (sh.batch == 0 || sh.len == 0) return SE3_OK;
    SE3_REQUIRE(proj && rot && trans && pair_bias && pair_value_packed && head_weight && out && p_workspace && inv_workspace, "null pointer");
    if (sh.dk != DK || sh.pq != PQ || sh.pv != PV || sh.pair_batch != 1 || sh.len > 256 || sh.heads % HG != 0 || sh.batch > 65535) {
        set_error("se3_ipa_attention_tc_fwd: needs dk=16, 4/8 points, shared pair tensors, L <= 256, heads %% 8 == 0 "
                  "(got dk=%d L=%d H=%d pair_batch=%d); use se3_ipa_attention_fwd", sh.dk, sh.len, sh.heads, sh.pair_batch);
        return SE3_EUNSUPPORTED;
    }
    SE3_REQUIRE(sh.proj_stride % 4 == 0 && sh.off_q % 4 == 0 && sh.off_k % 4 == 0 && sh.off_v % 4 == 0 &&
                (reinterpret_cast<uintptr_t>(proj) & 15) == 0, "projection matrix must be 16-byte aligned with offsets % 4 == 0");
    const int Lp = (sh.len + 15) / 16 * 16, Bpad = (sh.batch + 127) / 128 * 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_is_bf16)
        return launch_tc<__nv_bfloat16>(proj, rot, trans, pair_bias, (const __nv_bfloat16*)pair_value_packed, key_bias, head_weight,
                                        scalar_weight, (__nv_bfloat16*)out, (__nv_bfloat16*)p_workspace, inv_workspace, sh, Lp, Bpad, st);
    return launch_tc<float>(proj, rot, trans, pair_bias, (const __nv_bfloat16*)pair_value_packed, key_bias, head_weight, scalar_weight,
                            (float*)out, (__nv_bfloat16*)p_workspace, inv_workspace, sh, Lp, Bpad, st);
}

}  // extern "C"
