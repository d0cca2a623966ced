// K4 backward (fp32 SIMT) -- gradient of se3_ipa_attention_fwd (ipa_simt.cu), i.e. of SAAttention.forward between the
// input projections and fc_out (structure_module.py:131-216), for the fine-tune loss side (finetune.py:338-393: the
// control model is re-evaluated on stored rollout states with autograd; frames carry no gradient there).
//
//   s_ij   = (sw q_i).k_j + hw * sum_p |Qp_i - Kp_j| + pair_bias_hij + key_bias_j          P = softmax_j(s)
//   o_s    = sum_j P_ij v_j         o_pg = sum_j P_ij Vp_j      o_pl = R_i^T (o_pg - T_i)     o_n = |o_pl|
//   o_pair = sum_j P_ij z_ij
// With g_pl = dO_pl + dO_n o_pl/|o_pl|, g_pg = R_i g_pl:
//   dP_ij - D_i = dO_s.(v_j - o_s) + g_pg.(Vp_j - o_pg) + dO_pair.(z_ij - o_pair)          dS_ij = P_ij (dP_ij - D_i)
//   dq_i = sw sum_j dS_ij k_j        dk_j = sum_i dS_ij (sw q_i)      dv_j = sum_i P_ij dO_s,i
//   dQp_i = hw sum_j dS_ij sum_p (Qp_i - Kp_j)/|.|   (dKp_j: minus the same, summed over i)   dVp_j = sum_i P_ij g_pg,i
//   d hw  = sum_ij dS_ij sum_p |Qp_i - Kp_j|        point gradients return to the local frame through R^T.
// One CTA per (sample, head), all L <= 128 keys resident.  Phase 1: TWO lanes per QUERY row (even / odd keys; row maximum,
// row sum and the partial gradients joined by one shuffle) recompute the logits and the softmax (rows of P and dS stay in
// shared memory) and accumulate the query-side gradients; phase 2: two lanes per KEY column (even / odd rows) walk the
// same two matrices down the rows for the key-side gradients -- no atomics, deterministic.  Value points are staged
// relative to the sample's first residue, so that dO.(Vp_j - o_pg) is formed from nm-sized numbers.  P and
// dS also go to global memory (coalesced): the caller reduces them over the samples into the gradients of the shared
// pair tensors (d pair_bias = sum_b dS; d pair_value[i,j,h,:] = sum_b P_hij dO_pair_i -- a GEMM with K = samples).
//
// Longer sequences (L > 128, or a record size whose L x L matrices do not fit) run the same two phases as TWO kernels that meet in
// the P / dS workspaces the caller wants anyway: k_ipa_bwd_rows -- one CTA per (sample, head, 64 query rows), keys staged in
// chunks of 32 or 64, the rows' logits parked in shared memory between the three sweeps (maximum, sum, gradients), P / dS tiles
// written coalesced; k_ipa_bwd_cols -- one CTA per (sample, head, 64 key columns) walks P / dS down the rows (coalesced global
// reads) with the query records staged in chunks of 64 rows.
#include <math_constants.h>

#include "common.cuh"

using namespace se3;

namespace {

constexpr int PQ = 4, PV = 8;

__device__ __forceinline__ float pair_sum(float v) { return v + __shfl_xor_sync(0xffffffffu, v, 1); }

// Operands of one query row (sample b, head h, residue row_i = b * L + i): sw q, global query points, the incoming gradients
// dO_s / dO_pair, g_pg = R_i (dO_pl + dO_n o_pl / |o_pl|), the rotation, and D_i = dO_s.o_s + dO_pair.o_pair + g_pg.(o_pg - c0)
template <int DK>
struct QueryRow {
    float q[DK], qp[3 * PQ], gs[DK], gp[3 * PV], gzp[DK], Ri[9], Ds;
};
template <int DK>
__device__ __forceinline__ void load_query_row(QueryRow<DK>& r, const float* __restrict__ proj, const float* __restrict__ rot,
                                               const float* __restrict__ trans, const float* __restrict__ out, const float* __restrict__ d_out,
                                               float scalar_weight, const se3_ipa_shape& sh, int b, int h, int64_t row_i) {
    const int L = sh.len, H = sh.heads, HD = H * DK, W = 2 * HD + 4 * H * PV;
    const int C_S = h * DK, C_P = HD + h * PV * 3, C_Z = HD + 3 * H * PV + h * DK, C_N = 2 * HD + 3 * H * PV + h * PV;
    const float* pr = proj + row_i * sh.proj_stride;
    const float* o = out + row_i * (int64_t)W;
    const float* go = d_out + row_i * (int64_t)W;
    float Ti[3];
    r.Ds = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) r.Ri[k] = rot[row_i * 9 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) Ti[k] = trans[row_i * 3 + k];
#pragma unroll
    for (int c = 0; c < DK; ++c) {
        r.q[c] = pr[sh.off_q + h * sh.hs_scalar + c] * scalar_weight;
        r.gs[c] = go[C_S + c];
        r.gzp[c] = go[C_Z + c];
        r.Ds += r.gs[c] * o[C_S + c] + r.gzp[c] * o[C_Z + c];
    }
#pragma unroll
    for (int p = 0; p < PQ; ++p) {
        const float x = pr[sh.off_qp + h * sh.hs_point + p * 3], y = pr[sh.off_qp + h * sh.hs_point + p * 3 + 1], z = pr[sh.off_qp + h * sh.hs_point + p * 3 + 2];
#pragma unroll
        for (int c = 0; c < 3; ++c) r.qp[p * 3 + c] = ((r.Ri[c * 3] * x + r.Ri[c * 3 + 1] * y) + r.Ri[c * 3 + 2] * z) + Ti[c];
    }
#pragma unroll
    for (int p = 0; p < PV; ++p) {
        const float lx = o[C_P + p * 3], ly = o[C_P + p * 3 + 1], lz = o[C_P + p * 3 + 2];
        float gx = go[C_P + p * 3], gy = go[C_P + p * 3 + 1], gz = go[C_P + p * 3 + 2];
        const float nrm = sqrtf(lx * lx + ly * ly + lz * lz);
        if (nrm > 0.f) {                                           // torch.norm's backward is 0 at the origin
            const float sc = go[C_N + p] / nrm;
            gx += sc * lx; gy += sc * ly; gz += sc * lz;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            r.gp[p * 3 + c] = (r.Ri[c * 3] * gx + r.Ri[c * 3 + 1] * gy) + r.Ri[c * 3 + 2] * gz;
            const float opg = ((r.Ri[c * 3] * lx + r.Ri[c * 3 + 1] * ly) + r.Ri[c * 3 + 2] * lz) + (Ti[c] - trans[(int64_t)b * L * 3 + c]);
            r.Ds += r.gp[p * 3 + c] * opg;
        }
    }
}

// Keys [j0, j0 + n) of (sample b, head h) into `keys` ([n][KW]: k | Kp | v | Vp, global frame; value points relative to the
// sample's first residue), cooperatively by the CTA
template <int DK>
__device__ __forceinline__ void stage_keys(float* keys, const float* __restrict__ proj, const float* __restrict__ rot, const float* __restrict__ trans,
                                           const se3_ipa_shape& sh, int b, int h, int j0, int n) {
    constexpr int KW = 2 * DK + 3 * PQ + 3 * PV, O_KS = 0, O_KP = DK, O_VS = DK + 3 * PQ, O_VP = 2 * DK + 3 * PQ;
    const int L = sh.len;
#pragma unroll 4
    for (int idx = threadIdx.x; idx < n * DK; idx += blockDim.x) {
        const int j = idx / DK, c = idx - j * DK;
        const float* pr = proj + ((int64_t)b * L + j0 + j) * sh.proj_stride;
        keys[j * KW + O_KS + c] = pr[sh.off_k + h * sh.hs_scalar + c];
        keys[j * KW + O_VS + c] = pr[sh.off_v + h * sh.hs_scalar + c];
    }
#pragma unroll 4
    for (int idx = threadIdx.x; idx < n * (PQ + PV); idx += blockDim.x) {
        const int j = idx / (PQ + PV), p = idx - j * (PQ + PV);
        const int64_t rj = (int64_t)b * L + j0 + j;
        const float* pr = proj + rj * sh.proj_stride + (p < PQ ? sh.off_kp + h * sh.hs_point + p * 3 : sh.off_vp + h * sh.hs_vpoint + (p - PQ) * 3);
        const float x = pr[0], y = pr[1], z = pr[2];
        const float* R = rot + rj * 9;
        const float* T = trans + rj * 3;
        const float* C0 = trans + (int64_t)b * L * 3;                 // centre of the value points: the sample's first residue
        float* dst = keys + j * KW + (p < PQ ? O_KP + p * 3 : O_VP + (p - PQ) * 3);
#pragma unroll
        for (int r = 0; r < 3; ++r)
            dst[r] = ((R[r * 3] * x + R[r * 3 + 1] * y) + R[r * 3 + 2] * z) + (p < PQ ? T[r] : T[r] - C0[r]);
    }
}

template <int DK, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 192 ? 2 : 1)
k_ipa_bwd(const float* __restrict__ proj, const float* __restrict__ rot, const float* __restrict__ trans,
          const float* __restrict__ pair_bias, const float* __restrict__ pair_value, const float* __restrict__ key_bias,
          const float* __restrict__ head_weight, float scalar_weight, const float* __restrict__ out,
          const float* __restrict__ d_out, float* __restrict__ d_proj, float* __restrict__ p_ws, float* __restrict__ ds_ws,
          float* __restrict__ d_hw_rows, const se3_ipa_shape sh) {
    constexpr int KW = 2 * DK + 3 * PQ + 3 * PV;                                  // floats per staged record
    constexpr int O_KS = 0, O_KP = DK, O_VS = DK + 3 * PQ, O_VP = 2 * DK + 3 * PQ;  // key record: k | Kp | v | Vp (global frame)
    constexpr int O_Q = 0, O_QP = DK, O_GS = DK + 3 * PQ, O_GP = 2 * DK + 3 * PQ;   // query record: sw q | Qp | dO_s | g_pg
    extern __shared__ __align__(16) float smem[];
    const int L = sh.len, H = sh.heads, LS = L | 1;                               // odd row pitch: conflict-free both ways
    float* keys = smem;               // [L][KW]
    float* qrec = keys + L * KW;      // [L][KW]
    float* Pm = qrec + L * KW;        // [L][LS]
    float* Sm = Pm + L * LS;          // [L][LS]
    float* kbias = Sm + L * LS;       // [L]

    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int lane2 = tid & 1;                                                    // which half of the keys (phase 1) / rows (phase 2)
    const bool active = (tid >> 1) < L;
    const int i = active ? (tid >> 1) : 0;
    const int n_act = active ? L : 0;                                             // idle lanes run empty loops, join the shuffles
    const int pb = sh.pair_batch == 1 ? 0 : b;
    const int64_t row_i = (int64_t)b * L + i;
    const float hw = head_weight[h];
    const int HD = H * DK;
    const int W = 2 * HD + 4 * H * PV;
    const int C_S = h * DK, C_P = HD + h * PV * 3, C_Z = HD + 3 * H * PV + h * DK, C_N = 2 * HD + 3 * H * PV + h * PV;

    // ---- phase 0: stage the keys of this (sample, head) in the global frame --------------------------------
#pragma unroll 4
    for (int idx = tid; idx < L * DK; idx += blockDim.x) {
        const int j = idx / DK, c = idx - j * DK;
        const float* pr = proj + ((int64_t)b * L + j) * sh.proj_stride;
        keys[j * KW + O_KS + c] = pr[sh.off_k + h * sh.hs_scalar + c];
        keys[j * KW + O_VS + c] = pr[sh.off_v + h * sh.hs_scalar + c];
    }
#pragma unroll 4
    for (int idx = tid; idx < L * (PQ + PV); idx += blockDim.x) {
        const int j = idx / (PQ + PV), p = idx - j * (PQ + PV);
        const int64_t rj = (int64_t)b * L + j;
        const float* pr = proj + rj * sh.proj_stride + (p < PQ ? sh.off_kp + h * sh.hs_point + p * 3 : sh.off_vp + h * sh.hs_vpoint + (p - PQ) * 3);
        const float x = pr[0], y = pr[1], z = pr[2];
        const float* R = rot + rj * 9;
        const float* T = trans + rj * 3;
        const float* C0 = trans + (int64_t)b * L * 3;                 // centre of the value points: the sample's first residue
        float* dst = keys + j * KW + (p < PQ ? O_KP + p * 3 : O_VP + (p - PQ) * 3);
#pragma unroll
        for (int r = 0; r < 3; ++r)
            dst[r] = ((R[r * 3] * x + R[r * 3 + 1] * y) + R[r * 3 + 2] * z) + (p < PQ ? T[r] : T[r] - C0[r]);
    }
    for (int idx = tid; idx < L; idx += blockDim.x) kbias[idx] = key_bias ? key_bias[(int64_t)b * L + idx] : 0.f;

    // ---- the thread's query row: operands, incoming gradients, D_i ---------------------------------------------
    float q[DK], qp[3 * PQ], gs[DK], gp[3 * PV], gzp[DK], Ri[9];
    float Ds = 0.f;                                                    // D_i = dO_s.o_s + dO_pair.o_pair + g_pg.(o_pg - c0)
    if (active) {
        const float* pr = proj + row_i * sh.proj_stride;
        const float* o = out + row_i * (int64_t)W;
        const float* go = d_out + row_i * (int64_t)W;
        float Ti[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) Ri[k] = rot[row_i * 9 + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) Ti[k] = trans[row_i * 3 + k];
#pragma unroll
        for (int c = 0; c < DK; ++c) {
            q[c] = pr[sh.off_q + h * sh.hs_scalar + c] * scalar_weight;
            gs[c] = go[C_S + c];
            gzp[c] = go[C_Z + c];
            Ds += gs[c] * o[C_S + c] + gzp[c] * o[C_Z + c];
        }
#pragma unroll
        for (int p = 0; p < PQ; ++p) {
            const float x = pr[sh.off_qp + h * sh.hs_point + p * 3], y = pr[sh.off_qp + h * sh.hs_point + p * 3 + 1], z = pr[sh.off_qp + h * sh.hs_point + p * 3 + 2];
#pragma unroll
            for (int r = 0; r < 3; ++r) qp[p * 3 + r] = ((Ri[r * 3] * x + Ri[r * 3 + 1] * y) + Ri[r * 3 + 2] * z) + Ti[r];
        }
#pragma unroll
        for (int p = 0; p < PV; ++p) {
            const float lx = o[C_P + p * 3], ly = o[C_P + p * 3 + 1], lz = o[C_P + p * 3 + 2];
            float gx = go[C_P + p * 3], gy = go[C_P + p * 3 + 1], gz = go[C_P + p * 3 + 2];
            const float nrm = sqrtf(lx * lx + ly * ly + lz * lz);
            if (nrm > 0.f) {                                           // torch.norm's backward is 0 at the origin
                const float sc = go[C_N + p] / nrm;
                gx += sc * lx; gy += sc * ly; gz += sc * lz;
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                gp[p * 3 + r] = (Ri[r * 3] * gx + Ri[r * 3 + 1] * gy) + Ri[r * 3 + 2] * gz;
                const float opg = ((Ri[r * 3] * lx + Ri[r * 3 + 1] * ly) + Ri[r * 3 + 2] * lz) + (Ti[r] - trans[(int64_t)b * L * 3 + r]);
                Ds += gp[p * 3 + r] * opg;
            }
        }
        if (lane2 == 0) {
            float* qr = qrec + i * KW;
#pragma unroll
            for (int c = 0; c < DK; ++c) { qr[O_Q + c] = q[c]; qr[O_GS + c] = gs[c]; }
#pragma unroll
            for (int c = 0; c < 3 * PQ; ++c) qr[O_QP + c] = qp[c];
#pragma unroll
            for (int c = 0; c < 3 * PV; ++c) qr[O_GP + c] = gp[c];
        }
    }
    __syncthreads();

    // ---- phase 1: query rows ----------------------------------------------------------------------------------------
    {
        const float* bias_row = pair_bias + (((int64_t)pb * H + h) * L + i) * L;
        const float* pv_row = pair_value + (((int64_t)pb * L + i) * L) * (int64_t)HD + h * DK;
        float* Prow = Pm + i * LS;
        float* Srow = Sm + i * LS;
        float m = -CUDART_INF_F;
        for (int j = lane2; j < n_act; j += 2) {                      // logits (ipa_simt.cu evaluates the same expression)
            const float4* kr = reinterpret_cast<const float4*>(keys + j * KW);
            float dot = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 kv = kr[c4];
                dot += q[c4 * 4] * kv.x; dot += q[c4 * 4 + 1] * kv.y; dot += q[c4 * 4 + 2] * kv.z; dot += q[c4 * 4 + 3] * kv.w;
            }
            float kp[3 * PQ];
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) {
                const float4 kv = kr[DK / 4 + c4];
                kp[c4 * 4] = kv.x; kp[c4 * 4 + 1] = kv.y; kp[c4 * 4 + 2] = kv.z; kp[c4 * 4 + 3] = kv.w;
            }
            float dsum = 0.f;
#pragma unroll
            for (int p = 0; p < PQ; ++p) {
                const float dx = qp[p * 3] - kp[p * 3], dy = qp[p * 3 + 1] - kp[p * 3 + 1], dz = qp[p * 3 + 2] - kp[p * 3 + 2];
                dsum += sqrtf(dx * dx + dy * dy + dz * dz);
            }
            const float s = ((dot + hw * dsum) + __ldg(bias_row + j)) + kbias[j];
            Srow[j] = s;
            m = fmaxf(m, s);
        }
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        float l = 0.f;
        for (int j = lane2; j < n_act; j += 2) {
            const float e = m == -CUDART_INF_F ? 0.f : expf(Srow[j] - m);
            Prow[j] = e;
            l += e;
        }
        l = pair_sum(l);
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        float dq[DK], dQp[3 * PQ], dhw = 0.f;
#pragma unroll
        for (int c = 0; c < DK; ++c) dq[c] = 0.f;
#pragma unroll
        for (int c = 0; c < 3 * PQ; ++c) dQp[c] = 0.f;
        for (int j = lane2; j < n_act; j += 2) {
            const float p = Prow[j] * inv;
            const float4* kr = reinterpret_cast<const float4*>(keys + j * KW);
            float dP = -Ds;
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 v = kr[O_VS / 4 + c4];
                dP += gs[c4 * 4] * v.x; dP += gs[c4 * 4 + 1] * v.y; dP += gs[c4 * 4 + 2] * v.z; dP += gs[c4 * 4 + 3] * v.w;
            }
#pragma unroll
            for (int c4 = 0; c4 < 3 * PV / 4; ++c4) {
                const float4 v = kr[O_VP / 4 + c4];
                dP += gp[c4 * 4] * v.x; dP += gp[c4 * 4 + 1] * v.y; dP += gp[c4 * 4 + 2] * v.z; dP += gp[c4 * 4 + 3] * v.w;
            }
            const float4* zr = reinterpret_cast<const float4*>(pv_row + (int64_t)j * HD);
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 v = __ldg(zr + c4);
                dP += gzp[c4 * 4] * v.x; dP += gzp[c4 * 4 + 1] * v.y; dP += gzp[c4 * 4 + 2] * v.z; dP += gzp[c4 * 4 + 3] * v.w;
            }
            const float ds = p * dP;
            Prow[j] = p;
            Srow[j] = ds;
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 kv = kr[O_KS / 4 + c4];
                dq[c4 * 4] += ds * kv.x; dq[c4 * 4 + 1] += ds * kv.y; dq[c4 * 4 + 2] += ds * kv.z; dq[c4 * 4 + 3] += ds * kv.w;
            }
            float kp[3 * PQ];
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) {
                const float4 kv = kr[O_KP / 4 + c4];
                kp[c4 * 4] = kv.x; kp[c4 * 4 + 1] = kv.y; kp[c4 * 4 + 2] = kv.z; kp[c4 * 4 + 3] = kv.w;
            }
            float dsum = 0.f;
            const float hds = hw * ds;
#pragma unroll
            for (int pt = 0; pt < PQ; ++pt) {
                const float dx = qp[pt * 3] - kp[pt * 3], dy = qp[pt * 3 + 1] - kp[pt * 3 + 1], dz = qp[pt * 3 + 2] - kp[pt * 3 + 2];
                const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                dsum += d;
                const float coef = d > 0.f ? hds / d : 0.f;            // torch.norm's backward is 0 at distance 0
                dQp[pt * 3] += coef * dx; dQp[pt * 3 + 1] += coef * dy; dQp[pt * 3 + 2] += coef * dz;
            }
            dhw += ds * dsum;
        }
#pragma unroll
        for (int c = 0; c < DK; ++c) dq[c] = pair_sum(dq[c]);
#pragma unroll
        for (int c = 0; c < 3 * PQ; ++c) dQp[c] = pair_sum(dQp[c]);
        dhw = pair_sum(dhw);
        float* gr = d_proj + row_i * sh.proj_stride;
        if (active && lane2 == 0) {
#pragma unroll
            for (int c = 0; c < DK; ++c) gr[sh.off_q + h * sh.hs_scalar + c] = dq[c] * scalar_weight;
            d_hw_rows[row_i * H + h] = dhw;
        }
        if (active && lane2 == 1) {
#pragma unroll
            for (int pt = 0; pt < PQ; ++pt)
#pragma unroll
                for (int c = 0; c < 3; ++c)                             // local = R^T global
                    gr[sh.off_qp + h * sh.hs_point + pt * 3 + c] = (Ri[c] * dQp[pt * 3] + Ri[3 + c] * dQp[pt * 3 + 1]) + Ri[6 + c] * dQp[pt * 3 + 2];
        }
    }
    __syncthreads();

    // ---- P and dS of this (sample, head) to global memory, coalesced ---------------------------------------------
    {
        float* pg = p_ws + ((int64_t)b * H + h) * L * L;
        float* sg = ds_ws + ((int64_t)b * H + h) * L * L;
#pragma unroll 4
        for (int idx = tid; idx < L * L; idx += blockDim.x) {
            const int r = idx / L, c = idx - r * L;
            pg[idx] = Pm[r * LS + c];
            sg[idx] = Sm[r * LS + c];
        }
    }

    // ---- phase 2: key columns -----------------------------------------------------------------------------------------
    {
        const int j = i;
        float kp[3 * PQ], dk[DK], dv[DK], dVp[3 * PV], dKp[3 * PQ];
#pragma unroll
        for (int c = 0; c < 3 * PQ; ++c) { kp[c] = keys[j * KW + O_KP + c]; dKp[c] = 0.f; }
#pragma unroll
        for (int c = 0; c < DK; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
#pragma unroll
        for (int c = 0; c < 3 * PV; ++c) dVp[c] = 0.f;
        for (int r = lane2; r < n_act; r += 2) {
            const float p = Pm[r * LS + j], ds = Sm[r * LS + j];
            const float4* qr = reinterpret_cast<const float4*>(qrec + r * KW);
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 v = qr[O_Q / 4 + c4];
                dk[c4 * 4] += ds * v.x; dk[c4 * 4 + 1] += ds * v.y; dk[c4 * 4 + 2] += ds * v.z; dk[c4 * 4 + 3] += ds * v.w;
            }
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 v = qr[O_GS / 4 + c4];
                dv[c4 * 4] += p * v.x; dv[c4 * 4 + 1] += p * v.y; dv[c4 * 4 + 2] += p * v.z; dv[c4 * 4 + 3] += p * v.w;
            }
#pragma unroll
            for (int c4 = 0; c4 < 3 * PV / 4; ++c4) {
                const float4 v = qr[O_GP / 4 + c4];
                dVp[c4 * 4] += p * v.x; dVp[c4 * 4 + 1] += p * v.y; dVp[c4 * 4 + 2] += p * v.z; dVp[c4 * 4 + 3] += p * v.w;
            }
            float qv[3 * PQ];
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) {
                const float4 v = qr[O_QP / 4 + c4];
                qv[c4 * 4] = v.x; qv[c4 * 4 + 1] = v.y; qv[c4 * 4 + 2] = v.z; qv[c4 * 4 + 3] = v.w;
            }
            const float hds = hw * ds;
#pragma unroll
            for (int pt = 0; pt < PQ; ++pt) {
                const float dx = qv[pt * 3] - kp[pt * 3], dy = qv[pt * 3 + 1] - kp[pt * 3 + 1], dz = qv[pt * 3 + 2] - kp[pt * 3 + 2];
                const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                const float coef = d > 0.f ? hds / d : 0.f;
                dKp[pt * 3] -= coef * dx; dKp[pt * 3 + 1] -= coef * dy; dKp[pt * 3 + 2] -= coef * dz;
            }
        }
#pragma unroll
        for (int c = 0; c < DK; ++c) { dk[c] = pair_sum(dk[c]); dv[c] = pair_sum(dv[c]); }
#pragma unroll
        for (int c = 0; c < 3 * PQ; ++c) dKp[c] = pair_sum(dKp[c]);
#pragma unroll
        for (int c = 0; c < 3 * PV; ++c) dVp[c] = pair_sum(dVp[c]);
        // Ri is this thread's own residue (i == j): gradients of the local points are R^T (global gradient)
        float* gr = d_proj + row_i * sh.proj_stride;
        if (active && lane2 == 0) {
#pragma unroll
            for (int c = 0; c < DK; ++c) {
                gr[sh.off_k + h * sh.hs_scalar + c] = dk[c];
                gr[sh.off_v + h * sh.hs_scalar + c] = dv[c];
            }
#pragma unroll
            for (int pt = 0; pt < PQ; ++pt)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    gr[sh.off_kp + h * sh.hs_point + pt * 3 + c] = (Ri[c] * dKp[pt * 3] + Ri[3 + c] * dKp[pt * 3 + 1]) + Ri[6 + c] * dKp[pt * 3 + 2];
        }
        if (active && lane2 == 1) {
#pragma unroll
            for (int pt = 0; pt < PV; ++pt)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    gr[sh.off_vp + h * sh.hs_vpoint + pt * 3 + c] = (Ri[c] * dVp[pt * 3] + Ri[3 + c] * dVp[pt * 3 + 1]) + Ri[6 + c] * dVp[pt * 3 + 2];
        }
    }
}


// ---- tiled edition (L > 128): phase 1 ---------------------------------------------------------------------------------
// (key chunks of 32 where that lets two CTAs of the rows kernel share an SM -- with 64 it needed 116 KB at L = 256: one CTA of four
// warps per SM, issue-active 17 %; chunks of 64 where only one fits anyway; column tiles of 64 with three CTAs per SM instead of
// one 256-thread CTA at 215 registers)
constexpr int kRowTile = 64, kColTile = 64, kRowChunk = 64;

template <int DK, int kKeyChunk>
__global__ void __launch_bounds__(2 * kRowTile, kKeyChunk == 32 ? 2 : 1)
k_ipa_bwd_rows(const float* __restrict__ proj, const float* __restrict__ rot, const float* __restrict__ trans,
               const float* __restrict__ pair_bias, const float* __restrict__ pair_value, const float* __restrict__ key_bias,
               const float* __restrict__ head_weight, float scalar_weight, const float* __restrict__ out,
               const float* __restrict__ d_out, float* __restrict__ d_proj, float* __restrict__ p_ws, float* __restrict__ ds_ws,
               float* __restrict__ d_hw_rows, const se3_ipa_shape sh) {
    constexpr int KW = 2 * DK + 3 * PQ + 3 * PV, O_KS = 0, O_KP = DK, O_VS = DK + 3 * PQ, O_VP = 2 * DK + 3 * PQ;
    constexpr int RT = kRowTile, KC = kKeyChunk, TP = KC + 1;
    extern __shared__ __align__(16) float smem[];
    const int L = sh.len, H = sh.heads, LS = L | 1;
    float* keys = smem;                  // [KC][KW]
    float* Srow = keys + KC * KW;        // [RT][LS]: logits, then exp(logit - max), of this CTA's rows
    float* Pt = Srow + RT * LS;          // [RT][KC + 1] tiles of P / dS on their way to global memory
    float* St = Pt + RT * TP;
    float* kbias = St + RT * TP;         // [L]
    const int h = blockIdx.x, b = blockIdx.y, i0 = blockIdx.z * RT, tid = threadIdx.x;
    const int lane2 = tid & 1, rl = tid >> 1;
    const bool active = i0 + rl < L;
    const int i = active ? i0 + rl : 0;
    const int pb = sh.pair_batch == 1 ? 0 : b;
    const int64_t row_i = (int64_t)b * L + i;
    const float hw = head_weight[h];
    const int HD = H * DK;
    QueryRow<DK> qr;
    if (active) load_query_row<DK>(qr, proj, rot, trans, out, d_out, scalar_weight, sh, b, h, row_i);
    for (int idx = tid; idx < L; idx += blockDim.x) kbias[idx] = key_bias ? key_bias[(int64_t)b * L + idx] : 0.f;
    const float* bias_row = pair_bias + (((int64_t)pb * H + h) * L + i) * L;
    const float* pv_row = pair_value + (((int64_t)pb * L + i) * L) * (int64_t)HD + h * DK;
    float* Sr = Srow + rl * LS;

    // sweep 1: logits (the expression of k_ipa_bwd / ipa_simt.cu), row maximum
    float m = -CUDART_INF_F;
    for (int c0 = 0; c0 < L; c0 += KC) {
        const int n = min(KC, L - c0);
        __syncthreads();
        stage_keys<DK>(keys, proj, rot, trans, sh, b, h, c0, n);
        __syncthreads();
        for (int jj = lane2; jj < (active ? n : 0); jj += 2) {
            const int j = c0 + jj;
            const float4* kr = reinterpret_cast<const float4*>(keys + jj * KW);
            float dot = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 kv = kr[c4];
                dot += qr.q[c4 * 4] * kv.x; dot += qr.q[c4 * 4 + 1] * kv.y; dot += qr.q[c4 * 4 + 2] * kv.z; dot += qr.q[c4 * 4 + 3] * kv.w;
            }
            float kp[3 * PQ];
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) {
                const float4 kv = kr[DK / 4 + c4];
                kp[c4 * 4] = kv.x; kp[c4 * 4 + 1] = kv.y; kp[c4 * 4 + 2] = kv.z; kp[c4 * 4 + 3] = kv.w;
            }
            float dsum = 0.f;
#pragma unroll
            for (int p = 0; p < PQ; ++p) {
                const float dx = qr.qp[p * 3] - kp[p * 3], dy = qr.qp[p * 3 + 1] - kp[p * 3 + 1], dz = qr.qp[p * 3 + 2] - kp[p * 3 + 2];
                dsum += sqrtf(dx * dx + dy * dy + dz * dz);
            }
            const float sv = ((dot + hw * dsum) + __ldg(bias_row + j)) + kbias[j];
            Sr[j] = sv;
            m = fmaxf(m, sv);
        }
    }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    // sweep 2: exponentials and row sum (each lane revisits exactly the entries it wrote)
    float l = 0.f;
    for (int c0 = 0; c0 < L; c0 += KC) {
        const int n = min(KC, L - c0);
        for (int jj = lane2; jj < (active ? n : 0); jj += 2) {
            const float e = m == -CUDART_INF_F ? 0.f : expf(Sr[c0 + jj] - m);
            Sr[c0 + jj] = e;
            l += e;
        }
    }
    l = pair_sum(l);
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    // sweep 3: P, dS, query-side gradients
    float dq[DK], dQp[3 * PQ], dhw = 0.f;
#pragma unroll
    for (int c = 0; c < DK; ++c) dq[c] = 0.f;
#pragma unroll
    for (int c = 0; c < 3 * PQ; ++c) dQp[c] = 0.f;
    float* pg = p_ws + ((int64_t)b * H + h) * L * L;
    float* sg = ds_ws + ((int64_t)b * H + h) * L * L;
    for (int c0 = 0; c0 < L; c0 += KC) {
        const int n = min(KC, L - c0);
        __syncthreads();   // the previous chunk's tiles have left, its keys are no longer read
        stage_keys<DK>(keys, proj, rot, trans, sh, b, h, c0, n);
        __syncthreads();
        for (int jj = lane2; jj < (active ? n : 0); jj += 2) {
            const int j = c0 + jj;
            const float p = Sr[j] * inv;
            const float4* kr = reinterpret_cast<const float4*>(keys + jj * KW);
            float dP = -qr.Ds;
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 v = kr[O_VS / 4 + c4];
                dP += qr.gs[c4 * 4] * v.x; dP += qr.gs[c4 * 4 + 1] * v.y; dP += qr.gs[c4 * 4 + 2] * v.z; dP += qr.gs[c4 * 4 + 3] * v.w;
            }
#pragma unroll
            for (int c4 = 0; c4 < 3 * PV / 4; ++c4) {
                const float4 v = kr[O_VP / 4 + c4];
                dP += qr.gp[c4 * 4] * v.x; dP += qr.gp[c4 * 4 + 1] * v.y; dP += qr.gp[c4 * 4 + 2] * v.z; dP += qr.gp[c4 * 4 + 3] * v.w;
            }
            const float4* zr = reinterpret_cast<const float4*>(pv_row + (int64_t)j * HD);
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 v = __ldg(zr + c4);
                dP += qr.gzp[c4 * 4] * v.x; dP += qr.gzp[c4 * 4 + 1] * v.y; dP += qr.gzp[c4 * 4 + 2] * v.z; dP += qr.gzp[c4 * 4 + 3] * v.w;
            }
            const float ds = p * dP;
            Pt[rl * TP + jj] = p;
            St[rl * TP + jj] = ds;
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 kv = kr[O_KS / 4 + c4];
                dq[c4 * 4] += ds * kv.x; dq[c4 * 4 + 1] += ds * kv.y; dq[c4 * 4 + 2] += ds * kv.z; dq[c4 * 4 + 3] += ds * kv.w;
            }
            float kp[3 * PQ];
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) {
                const float4 kv = kr[O_KP / 4 + c4];
                kp[c4 * 4] = kv.x; kp[c4 * 4 + 1] = kv.y; kp[c4 * 4 + 2] = kv.z; kp[c4 * 4 + 3] = kv.w;
            }
            float dsum = 0.f;
            const float hds = hw * ds;
#pragma unroll
            for (int pt = 0; pt < PQ; ++pt) {
                const float dx = qr.qp[pt * 3] - kp[pt * 3], dy = qr.qp[pt * 3 + 1] - kp[pt * 3 + 1], dz = qr.qp[pt * 3 + 2] - kp[pt * 3 + 2];
                const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                dsum += d;
                const float coef = d > 0.f ? hds / d : 0.f;            // torch.norm's backward is 0 at distance 0
                dQp[pt * 3] += coef * dx; dQp[pt * 3 + 1] += coef * dy; dQp[pt * 3 + 2] += coef * dz;
            }
            dhw += ds * dsum;
        }
        __syncthreads();
        const int nrow = min(RT, L - i0);
        for (int idx = tid; idx < nrow * n; idx += blockDim.x) {      // tile -> global, rows of n contiguous floats
            const int r = idx / n, c = idx - r * n;
            pg[(int64_t)(i0 + r) * L + c0 + c] = Pt[r * TP + c];
            sg[(int64_t)(i0 + r) * L + c0 + c] = St[r * TP + c];
        }
    }
#pragma unroll
    for (int c = 0; c < DK; ++c) dq[c] = pair_sum(dq[c]);
#pragma unroll
    for (int c = 0; c < 3 * PQ; ++c) dQp[c] = pair_sum(dQp[c]);
    dhw = pair_sum(dhw);
    float* gr = d_proj + row_i * sh.proj_stride;
    if (active && lane2 == 0) {
#pragma unroll
        for (int c = 0; c < DK; ++c) gr[sh.off_q + h * sh.hs_scalar + c] = dq[c] * scalar_weight;
        d_hw_rows[row_i * H + h] = dhw;
    }
    if (active && lane2 == 1) {
#pragma unroll
        for (int pt = 0; pt < PQ; ++pt)
#pragma unroll
            for (int c = 0; c < 3; ++c)                                 // local = R^T global
                gr[sh.off_qp + h * sh.hs_point + pt * 3 + c] = (qr.Ri[c] * dQp[pt * 3] + qr.Ri[3 + c] * dQp[pt * 3 + 1]) + qr.Ri[6 + c] * dQp[pt * 3 + 2];
    }
}

// ---- tiled edition: phase 2 (runs after k_ipa_bwd_rows on the same stream) -----------------------------------------------
template <int DK>
__global__ void __launch_bounds__(2 * kColTile, 3)
k_ipa_bwd_cols(const float* __restrict__ proj, const float* __restrict__ rot, const float* __restrict__ trans,
               const float* __restrict__ head_weight, float scalar_weight, const float* __restrict__ out,
               const float* __restrict__ d_out, float* __restrict__ d_proj, const float* __restrict__ p_ws,
               const float* __restrict__ ds_ws, const se3_ipa_shape sh) {
    constexpr int KW = 2 * DK + 3 * PQ + 3 * PV, O_Q = 0, O_QP = DK, O_GS = DK + 3 * PQ, O_GP = 2 * DK + 3 * PQ;
    constexpr int RC = kRowChunk;
    extern __shared__ __align__(16) float smem[];
    float* qrec = smem;                  // [RC][KW]: sw q | Qp | dO_s | g_pg of a chunk of query rows
    const int L = sh.len, H = sh.heads;
    const int h = blockIdx.x, b = blockIdx.y, j0 = blockIdx.z * kColTile, tid = threadIdx.x;
    const int lane2 = tid & 1;
    const bool active = j0 + (tid >> 1) < L;
    const int j = active ? j0 + (tid >> 1) : 0;
    const int64_t row_j = (int64_t)b * L + j;
    const float hw = head_weight[h];
    float kp[3 * PQ], Rj[9], dk[DK], dv[DK], dVp[3 * PV], dKp[3 * PQ];
    {
        const float* pr = proj + row_j * sh.proj_stride;
#pragma unroll
        for (int k = 0; k < 9; ++k) Rj[k] = rot[row_j * 9 + k];
#pragma unroll
        for (int p = 0; p < PQ; ++p) {
            const float x = pr[sh.off_kp + h * sh.hs_point + p * 3], y = pr[sh.off_kp + h * sh.hs_point + p * 3 + 1], z = pr[sh.off_kp + h * sh.hs_point + p * 3 + 2];
#pragma unroll
            for (int c = 0; c < 3; ++c) kp[p * 3 + c] = ((Rj[c * 3] * x + Rj[c * 3 + 1] * y) + Rj[c * 3 + 2] * z) + trans[row_j * 3 + c];
        }
    }
#pragma unroll
    for (int c = 0; c < 3 * PQ; ++c) dKp[c] = 0.f;
#pragma unroll
    for (int c = 0; c < DK; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
#pragma unroll
    for (int c = 0; c < 3 * PV; ++c) dVp[c] = 0.f;
    const float* pg = p_ws + ((int64_t)b * H + h) * L * L;
    const float* sg = ds_ws + ((int64_t)b * H + h) * L * L;
    for (int r0 = 0; r0 < L; r0 += RC) {
        const int n = min(RC, L - r0);
        __syncthreads();
        if (tid < n) {
            QueryRow<DK> qr;
            load_query_row<DK>(qr, proj, rot, trans, out, d_out, scalar_weight, sh, b, h, (int64_t)b * L + r0 + tid);
            float* q = qrec + tid * KW;
#pragma unroll
            for (int c = 0; c < DK; ++c) { q[O_Q + c] = qr.q[c]; q[O_GS + c] = qr.gs[c]; }
#pragma unroll
            for (int c = 0; c < 3 * PQ; ++c) q[O_QP + c] = qr.qp[c];
#pragma unroll
            for (int c = 0; c < 3 * PV; ++c) q[O_GP + c] = qr.gp[c];
        }
        __syncthreads();
        for (int rr = lane2; rr < (active ? n : 0); rr += 2) {
            const int r = r0 + rr;
            const float p = pg[(int64_t)r * L + j], ds = sg[(int64_t)r * L + j];
            const float4* q4 = reinterpret_cast<const float4*>(qrec + rr * KW);
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 v = q4[O_Q / 4 + c4];
                dk[c4 * 4] += ds * v.x; dk[c4 * 4 + 1] += ds * v.y; dk[c4 * 4 + 2] += ds * v.z; dk[c4 * 4 + 3] += ds * v.w;
            }
#pragma unroll
            for (int c4 = 0; c4 < DK / 4; ++c4) {
                const float4 v = q4[O_GS / 4 + c4];
                dv[c4 * 4] += p * v.x; dv[c4 * 4 + 1] += p * v.y; dv[c4 * 4 + 2] += p * v.z; dv[c4 * 4 + 3] += p * v.w;
            }
#pragma unroll
            for (int c4 = 0; c4 < 3 * PV / 4; ++c4) {
                const float4 v = q4[O_GP / 4 + c4];
                dVp[c4 * 4] += p * v.x; dVp[c4 * 4 + 1] += p * v.y; dVp[c4 * 4 + 2] += p * v.z; dVp[c4 * 4 + 3] += p * v.w;
            }
            float qv[3 * PQ];
#pragma unroll
            for (int c4 = 0; c4 < 3; ++c4) {
                const float4 v = q4[O_QP / 4 + c4];
                qv[c4 * 4] = v.x; qv[c4 * 4 + 1] = v.y; qv[c4 * 4 + 2] = v.z; qv[c4 * 4 + 3] = v.w;
            }
            const float hds = hw * ds;
#pragma unroll
            for (int pt = 0; pt < PQ; ++pt) {
                const float dx = qv[pt * 3] - kp[pt * 3], dy = qv[pt * 3 + 1] - kp[pt * 3 + 1], dz = qv[pt * 3 + 2] - kp[pt * 3 + 2];
                const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                const float coef = d > 0.f ? hds / d : 0.f;
                dKp[pt * 3] -= coef * dx; dKp[pt * 3 + 1] -= coef * dy; dKp[pt * 3 + 2] -= coef * dz;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < DK; ++c) { dk[c] = pair_sum(dk[c]); dv[c] = pair_sum(dv[c]); }
#pragma unroll
    for (int c = 0; c < 3 * PQ; ++c) dKp[c] = pair_sum(dKp[c]);
#pragma unroll
    for (int c = 0; c < 3 * PV; ++c) dVp[c] = pair_sum(dVp[c]);
    float* gr = d_proj + row_j * sh.proj_stride;
    if (active && lane2 == 0) {
#pragma unroll
        for (int c = 0; c < DK; ++c) {
            gr[sh.off_k + h * sh.hs_scalar + c] = dk[c];
            gr[sh.off_v + h * sh.hs_scalar + c] = dv[c];
        }
#pragma unroll
        for (int pt = 0; pt < PQ; ++pt)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                gr[sh.off_kp + h * sh.hs_point + pt * 3 + c] = (Rj[c] * dKp[pt * 3] + Rj[3 + c] * dKp[pt * 3 + 1]) + Rj[6 + c] * dKp[pt * 3 + 2];
    }
    if (active && lane2 == 1) {
#pragma unroll
        for (int pt = 0; pt < PV; ++pt)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                gr[sh.off_vp + h * sh.hs_vpoint + pt * 3 + c] = (Rj[c] * dVp[pt * 3] + Rj[3 + c] * dVp[pt * 3 + 1]) + Rj[6 + c] * dVp[pt * 3 + 2];
    }
}

size_t resident_smem_bytes(int L, int KW) { return ((size_t)2 * L * KW + (size_t)2 * L * (L | 1) + L) * sizeof(float); }
size_t rows_smem_bytes(int L, int KW, int kc) { return ((size_t)kc * KW + (size_t)kRowTile * (L | 1) + (size_t)2 * kRowTile * (kc + 1) + L) * sizeof(float); }

template <int DK>
int launch(const float* proj, const float* rot, const float* trans, const float* pair_bias, const float* pair_value,
           const float* key_bias, const float* head_weight, float scalar_weight, const float* out, const float* d_out,
           float* d_proj, float* p_ws, float* ds_ws, float* d_hw_rows, const se3_ipa_shape& sh, cudaStream_t st) {
    constexpr int KW = 2 * DK + 3 * PQ + 3 * PV;
    const int L = sh.len;
    if (L > 128 || resident_smem_bytes(L, KW) > (size_t)227 * 1024) {   // tiled edition: two kernels meeting in the P / dS workspaces
        const bool two_fit = 2 * (rows_smem_bytes(L, KW, 32) + 1024) <= (size_t)227 * 1024;
        const size_t smem_r = rows_smem_bytes(L, KW, two_fit ? 32 : 64), smem_c = (size_t)kRowChunk * KW * sizeof(float);
        if (smem_r > (size_t)227 * 1024) { set_error("se3_ipa_attention_bwd: len=%d needs %zu bytes of shared memory per CTA", L, smem_r); return SE3_EUNSUPPORTED; }
        auto kr = two_fit ? k_ipa_bwd_rows<DK, 32> : k_ipa_bwd_rows<DK, 64>;
        auto kc = k_ipa_bwd_cols<DK>;
        cudaError_t e = cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
        if (e != cudaSuccess) { set_error("ipa bwd (rows) smem attribute (%zu bytes): %s", smem_r, cudaGetErrorString(e)); return SE3_ECUDA; }
        kr<<<dim3(sh.heads, sh.batch, (L + kRowTile - 1) / kRowTile), 2 * kRowTile, smem_r, st>>>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight,
                                                                                                   scalar_weight, out, d_out, d_proj, p_ws, ds_ws, d_hw_rows, sh);
        count_launch();
        if (int rc = check_launch("se3_ipa_attention_bwd(rows)")) return rc;
        kc<<<dim3(sh.heads, sh.batch, (L + kColTile - 1) / kColTile), 2 * kColTile, smem_c, st>>>(proj, rot, trans, head_weight, scalar_weight, out, d_out, d_proj,
                                                                                                   p_ws, ds_ws, sh);
        count_launch();
        return check_launch("se3_ipa_attention_bwd(columns)");
    }
    const int threads = ((2 * L + 31) / 32) * 32;                     // two lanes per row / column
    const size_t smem = resident_smem_bytes(L, KW);
    auto kern = threads <= 192 ? k_ipa_bwd<DK, 192> : k_ipa_bwd<DK, 256>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("ipa bwd smem attribute (%zu bytes): %s", smem, cudaGetErrorString(e)); return SE3_ECUDA; }
    }
    dim3 grid(sh.heads, sh.batch);
    kern<<<grid, threads, smem, st>>>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, d_out,
                                      d_proj, p_ws, ds_ws, d_hw_rows, sh);
    count_launch();
    return check_launch("se3_ipa_attention_bwd");
}

}  // namespace

extern "C" int se3_ipa_attention_bwd(const float* proj, const float* rot, const float* trans, const float* pair_bias,
                                     const float* pair_value, const float* key_bias, const float* head_weight,
                                     float scalar_weight, const float* out, const float* d_out, float* d_proj, float* p_ws,
                                     float* ds_ws, float* d_hw_rows, const se3_ipa_shape* h_shape, se3_stream_t stream) {
    SE3_REQUIRE(h_shape, "null shape");
    const se3_ipa_shape& sh = *h_shape;
    SE3_REQUIRE(sh.batch >= 0 && sh.len >= 0 && sh.heads > 0, "bad shape");
    if (sh.batch == 0 || sh.len == 0) return SE3_OK;
    SE3_REQUIRE(proj && rot && trans && pair_bias && pair_value && head_weight && out && d_out && d_proj && p_ws && ds_ws && d_hw_rows,
                "null pointer");
    SE3_REQUIRE(sh.pq == PQ && sh.pv == PV, "only 4 query/key points and 8 value points (structure_module.py:85-93)");
    SE3_REQUIRE(sh.pair_batch == 1 || sh.pair_batch == sh.batch, "pair_batch must be 1 or batch");
    SE3_REQUIRE(sh.len <= SE3_IPA_BWD_MAX_LEN, "se3_ipa_attention_bwd parks the logits of 64 query rows in shared memory: len <= 512");
    SE3_REQUIRE(sh.batch <= 65535, "grid limit");
    SE3_REQUIRE((reinterpret_cast<uintptr_t>(pair_value) & 15) == 0, "pair_value must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    switch (sh.dk) {
        case 4: return launch<4>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, d_out, d_proj, p_ws, ds_ws, d_hw_rows, sh, st);
        case 8: return launch<8>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, d_out, d_proj, p_ws, ds_ws, d_hw_rows, sh, st);
        case 16: return launch<16>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, d_out, d_proj, p_ws, ds_ws, d_hw_rows, sh, st);
        case 32: return launch<32>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, d_out, d_proj, p_ws, ds_ws, d_hw_rows, sh, st);
        default: set_error("se3_ipa_attention_bwd: unsupported dk=%d (4, 8, 16, 32)", sh.dk); return SE3_EUNSUPPORTED;
    }
}
