// K4 (fp32 SIMT edition) -- DiG invariant point attention, SAAttention.forward between the input
// projections and fc_out (structure_module.py:131-216), fused into one kernel:
//   logits = (q*sw).k + head_w * sum_p |Qp_i - Kp_j| + pair_bias + key_bias        :137-182
//   softmax over keys, streamed with a running max (flash-style, keys tiled through smem)   :186
//   out = [ P.v | R_i^T (P.Vp_global - T_i) | sum_j P_ij pair_value[i,j] | |.| ]            :189-216
// Mapping: one thread owns one (sample, head, query row); keys/values of the (sample, head) are
// staged per tile in shared memory in the GLOBAL frame and read as 128-bit broadcasts; all
// accumulation is fp32 (the reference forces fp32 for the point term, :193-196).
// This is the all-fp32 parity path; the tcgen05 edition (ipa_tc.cu) is the throughput path.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"

using namespace se3;

namespace {

template <bool FAST> __device__ __forceinline__ float f_sqrt(float x) {
    if (FAST) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    return sqrtf(x);
}
template <bool FAST> __device__ __forceinline__ float f_exp(float x) { return FAST ? __expf(x) : expf(x); }

constexpr int PQ = 4, PV = 8;

template <int DK, bool FAST>
__global__ void __launch_bounds__(128)
k_ipa_rows(const float* __restrict__ proj, const float* __restrict__ rot, const float* __restrict__ trans,
           const float* __restrict__ pair_bias, const float* __restrict__ pair_value, const float* __restrict__ key_bias,
           const float* __restrict__ head_weight, float scalar_weight, float* __restrict__ out, const se3_ipa_shape sh,
           int tile_keys) {
    constexpr int KW = 2 * DK + 3 * PQ + 3 * PV;  // floats per staged key: k_s | k_pt | v_s | v_pt
    constexpr int O_KS = 0, O_KP = DK, O_VS = DK + 3 * PQ, O_VP = 2 * DK + 3 * PQ;
    extern __shared__ __align__(16) float smem[];
    constexpr int ZS = 8 * DK + 4;            // floats per thread: two chunks of 4 keys x DK pair values, +4: conflict-free LDS.128
    float* keys = smem;                       // [tile_keys][KW]
    float* kbias = smem + tile_keys * KW;     // [tile_keys]
    float* zslot = kbias + tile_keys + threadIdx.x * ZS;   // this thread's pair-value ring (cp.async, one chunk ahead)

    const int L = sh.len, H = sh.heads;
    const int b = blockIdx.z, h = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < L;
    const int pb = sh.pair_batch == 1 ? 0 : b;
    const int64_t row_i = (int64_t)b * L + (active ? i : 0);
    const float hw = head_weight[h];

    // query: scaled scalar part and global-frame points
    float q[DK], qp[3 * PQ], Ri[9], Ti[3];
    {
        const float* pr = proj + row_i * sh.proj_stride;
#pragma unroll
        for (int c = 0; c < DK; ++c) q[c] = pr[sh.off_q + h * sh.hs_scalar + c] * scalar_weight;
#pragma unroll
        for (int k = 0; k < 9; ++k) Ri[k] = rot[row_i * 9 + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) Ti[k] = trans[row_i * 3 + k];
#pragma unroll
        for (int p = 0; p < PQ; ++p) {
            const float x = pr[sh.off_qp + h * sh.hs_point + p * 3], y = pr[sh.off_qp + h * sh.hs_point + p * 3 + 1], z = pr[sh.off_qp + h * sh.hs_point + p * 3 + 2];
#pragma unroll
            for (int r = 0; r < 3; ++r) qp[p * 3 + r] = ((Ri[r * 3] * x + Ri[r * 3 + 1] * y) + Ri[r * 3 + 2] * z) + Ti[r];
        }
    }

    float m = -CUDART_INF_F, l = 0.f;
    float acc_s[DK], acc_p[3 * PV], acc_z[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) { acc_s[c] = 0.f; acc_z[c] = 0.f; }
#pragma unroll
    for (int c = 0; c < 3 * PV; ++c) acc_p[c] = 0.f;

    const float* bias_row = pair_bias + (((int64_t)pb * H + h) * L + (active ? i : 0)) * L;
    const float* pv_row = pair_value + (((int64_t)pb * L + (active ? i : 0)) * L) * ((int64_t)H * DK) + h * DK;
    // pair_value[i, j, h, :] is this thread's own 4*DK-byte slice per key, 4*H*DK bytes apart: fetched with ordinary loads
    // inside the accumulation it exposed one L2 round trip per chunk (ncu at B = 64: long-scoreboard 7.2 of 10.6 stall
    // cycles per issue).  The slices of the NEXT 4-key chunk are copied into a thread-private shared-memory ring by cp.async
    // while the current chunk is computed; the arithmetic is unchanged.
    auto prefetch = [&](int jg) {
        float* dst = zslot + ((jg >> 2) & 1) * 4 * DK;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (jg + u < L) {
                const float* src = pv_row + (int64_t)(jg + u) * H * DK;
#pragma unroll
                for (int c4 = 0; c4 < DK / 4; ++c4)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst + u * DK + c4 * 4)), "l"(src + c4 * 4) : "memory");
            }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (active) prefetch(0);

    for (int j0 = 0; j0 < L; j0 += tile_keys) {
        const int nk = min(tile_keys, L - j0);
        __syncthreads();
        // ---- stage this key tile (all threads) -------------------------------------------------
#pragma unroll 4
        for (int idx = threadIdx.x; idx < nk * DK; idx += blockDim.x) {
            const int j = idx / DK, c = idx - j * DK;
            const float* pr = proj + ((int64_t)b * L + j0 + j) * sh.proj_stride;
            keys[j * KW + O_KS + c] = pr[sh.off_k + h * sh.hs_scalar + c];
            keys[j * KW + O_VS + c] = pr[sh.off_v + h * sh.hs_scalar + c];
        }
#pragma unroll 4
        for (int idx = threadIdx.x; idx < nk * (PQ + PV); idx += blockDim.x) {
            const int j = idx / (PQ + PV), p = idx - j * (PQ + PV);
            const int64_t rj = (int64_t)b * L + j0 + j;
            const float* pr = proj + rj * sh.proj_stride + (p < PQ ? sh.off_kp + h * sh.hs_point + p * 3 : sh.off_vp + h * sh.hs_vpoint + (p - PQ) * 3);
            const float x = pr[0], y = pr[1], z = pr[2];
            const float* R = rot + rj * 9;
            const float* T = trans + rj * 3;
            float* dst = keys + j * KW + (p < PQ ? O_KP + p * 3 : O_VP + (p - PQ) * 3);
#pragma unroll
            for (int r = 0; r < 3; ++r) dst[r] = ((R[r * 3] * x + R[r * 3 + 1] * y) + R[r * 3 + 2] * z) + T[r];
        }
        for (int idx = threadIdx.x; idx < nk; idx += blockDim.x) kbias[idx] = key_bias ? key_bias[(int64_t)b * L + j0 + idx] : 0.f;
        __syncthreads();
        if (!active) continue;

        // ---- stream the tile in chunks of 4 keys -----------------------------------------------
        for (int jj = 0; jj < nk; jj += 4) {
            prefetch(j0 + jj + 4);                                     // (an empty group past the last key)
            asm volatile("cp.async.wait_group 1;" ::: "memory");      // this chunk's slices have landed
            const float* zc = zslot + (((j0 + jj) >> 2) & 1) * 4 * DK;
            float s[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = jj + u;
                if (j < nk) {
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
                        dsum += f_sqrt<FAST>(dx * dx + dy * dy + dz * dz);
                    }
                    s[u] = ((dot + hw * dsum) + __ldg(bias_row + j0 + j)) + kbias[j];
                } else {
                    s[u] = -CUDART_INF_F;
                }
            }
            const float cm = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
            const float m_new = fmaxf(m, cm);
            if (m_new == -CUDART_INF_F) continue;  // everything masked so far
            const float scale = f_exp<FAST>(m - m_new);
            m = m_new;
            l *= scale;
#pragma unroll
            for (int c = 0; c < DK; ++c) { acc_s[c] *= scale; acc_z[c] *= scale; }
#pragma unroll
            for (int c = 0; c < 3 * PV; ++c) acc_p[c] *= scale;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = jj + u;
                if (j < nk) {
                    const float p = f_exp<FAST>(s[u] - m);
                    l += p;
                    const float4* vr = reinterpret_cast<const float4*>(keys + j * KW + O_VS);
#pragma unroll
                    for (int c4 = 0; c4 < DK / 4; ++c4) {
                        const float4 v = vr[c4];
                        acc_s[c4 * 4] += p * v.x; acc_s[c4 * 4 + 1] += p * v.y; acc_s[c4 * 4 + 2] += p * v.z; acc_s[c4 * 4 + 3] += p * v.w;
                    }
#pragma unroll
                    for (int c4 = 0; c4 < 3 * PV / 4; ++c4) {
                        const float4 v = vr[DK / 4 + c4];
                        acc_p[c4 * 4] += p * v.x; acc_p[c4 * 4 + 1] += p * v.y; acc_p[c4 * 4 + 2] += p * v.z; acc_p[c4 * 4 + 3] += p * v.w;
                    }
                    const float4* zr = reinterpret_cast<const float4*>(zc + u * DK);
#pragma unroll
                    for (int c4 = 0; c4 < DK / 4; ++c4) {
                        const float4 v = zr[c4];
                        acc_z[c4 * 4] += p * v.x; acc_z[c4 * 4 + 1] += p * v.y; acc_z[c4 * 4 + 2] += p * v.z; acc_z[c4 * 4 + 3] += p * v.w;
                    }
                }
            }
        }
    }
    if (!active) return;

    // ---- epilogue: normalise, inverse frame on the points, norms, concat layout (:198-216) --------
    const float inv = 1.0f / l;
    const int HD = H * DK;
    float* o = out + row_i * (int64_t)(2 * HD + 4 * H * PV);
#pragma unroll
    for (int c = 0; c < DK; ++c) o[h * DK + c] = acc_s[c] * inv;
#pragma unroll
    for (int p = 0; p < PV; ++p) {
        const float dx = acc_p[p * 3] * inv - Ti[0], dy = acc_p[p * 3 + 1] * inv - Ti[1], dz = acc_p[p * 3 + 2] * inv - Ti[2];
        float loc[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) loc[r] = (Ri[r] * dx + Ri[3 + r] * dy) + Ri[6 + r] * dz;
        o[HD + (h * PV + p) * 3] = loc[0];
        o[HD + (h * PV + p) * 3 + 1] = loc[1];
        o[HD + (h * PV + p) * 3 + 2] = loc[2];
        o[2 * HD + 3 * H * PV + h * PV + p] = sqrtf(loc[0] * loc[0] + loc[1] * loc[1] + loc[2] * loc[2]);
    }
#pragma unroll
    for (int c = 0; c < DK; ++c) o[HD + 3 * H * PV + h * DK + c] = acc_z[c] * inv;
}


// Small grids (the 0.19 M-parameter control model of the fine-tune rollout: B = 64, H = 4 -> 256 (sample, head) pairs): one thread
// per row leaves five warps per SM, each walking all L keys serially (125 us per call at L = 84, latency-bound).  This edition
// gives every query row KS lanes of one warp; lane s takes the 4-key chunks s, s + KS, ... with its own running maximum, sum and
// accumulators, and the lanes' flash-attention states are merged by a shuffle butterfly at the end.  128 threads = 128 / KS rows
// per CTA, so a (sample, head) pair spreads over several CTAs; the pair values are read with ordinary loads (the per-thread
// prefetch ring of k_ipa_rows would cost 67 KB of shared memory per CTA here, and there are four times the warps to cover the
// loads).  Same expressions per key as k_ipa_rows; the order of the sums over keys differs (fp32 either way).
template <int DK, bool FAST, int KS>
__global__ void __launch_bounds__(128)
k_ipa_rows_split(const float* __restrict__ proj, const float* __restrict__ rot, const float* __restrict__ trans,
                 const float* __restrict__ pair_bias, const float* __restrict__ pair_value, const float* __restrict__ key_bias,
                 const float* __restrict__ head_weight, float scalar_weight, float* __restrict__ out, const se3_ipa_shape sh,
                 int tile_keys) {
    constexpr int KW = 2 * DK + 3 * PQ + 3 * PV;
    constexpr int O_KS = 0, O_KP = DK, O_VS = DK + 3 * PQ, O_VP = 2 * DK + 3 * PQ;
    constexpr int kRows = 128 / KS;
    extern __shared__ __align__(16) float smem[];
    float* keys = smem;                       // [tile_keys][KW]
    float* kbias = smem + tile_keys * KW;     // [tile_keys]
    const int L = sh.len, H = sh.heads;
    const int b = blockIdx.z, h = blockIdx.y;
    const int part = threadIdx.x % KS;
    const int i = blockIdx.x * kRows + threadIdx.x / KS;
    const bool active = i < L;
    const int pb = sh.pair_batch == 1 ? 0 : b;
    const int64_t row_i = (int64_t)b * L + (active ? i : 0);
    const float hw = head_weight[h];
    float q[DK], qp[3 * PQ], Ri[9], Ti[3];
    {
        const float* pr = proj + row_i * sh.proj_stride;
#pragma unroll
        for (int c = 0; c < DK; ++c) q[c] = pr[sh.off_q + h * sh.hs_scalar + c] * scalar_weight;
#pragma unroll
        for (int k = 0; k < 9; ++k) Ri[k] = rot[row_i * 9 + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) Ti[k] = trans[row_i * 3 + k];
#pragma unroll
        for (int p = 0; p < PQ; ++p) {
            const float x = pr[sh.off_qp + h * sh.hs_point + p * 3], y = pr[sh.off_qp + h * sh.hs_point + p * 3 + 1], z = pr[sh.off_qp + h * sh.hs_point + p * 3 + 2];
#pragma unroll
            for (int r = 0; r < 3; ++r) qp[p * 3 + r] = ((Ri[r * 3] * x + Ri[r * 3 + 1] * y) + Ri[r * 3 + 2] * z) + Ti[r];
        }
    }
    float m = -CUDART_INF_F, l = 0.f;
    float acc_s[DK], acc_p[3 * PV], acc_z[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) { acc_s[c] = 0.f; acc_z[c] = 0.f; }
#pragma unroll
    for (int c = 0; c < 3 * PV; ++c) acc_p[c] = 0.f;
    const float* bias_row = pair_bias + (((int64_t)pb * H + h) * L + (active ? i : 0)) * L;
    const float* pv_row = pair_value + (((int64_t)pb * L + (active ? i : 0)) * L) * ((int64_t)H * DK) + h * DK;

    for (int j0 = 0; j0 < L; j0 += tile_keys) {
        const int nk = min(tile_keys, L - j0);
        __syncthreads();
#pragma unroll 4
        for (int idx = threadIdx.x; idx < nk * DK; idx += blockDim.x) {
            const int j = idx / DK, c = idx - j * DK;
            const float* pr = proj + ((int64_t)b * L + j0 + j) * sh.proj_stride;
            keys[j * KW + O_KS + c] = pr[sh.off_k + h * sh.hs_scalar + c];
            keys[j * KW + O_VS + c] = pr[sh.off_v + h * sh.hs_scalar + c];
        }
#pragma unroll 4
        for (int idx = threadIdx.x; idx < nk * (PQ + PV); idx += blockDim.x) {
            const int j = idx / (PQ + PV), p = idx - j * (PQ + PV);
            const int64_t rj = (int64_t)b * L + j0 + j;
            const float* pr = proj + rj * sh.proj_stride + (p < PQ ? sh.off_kp + h * sh.hs_point + p * 3 : sh.off_vp + h * sh.hs_vpoint + (p - PQ) * 3);
            const float x = pr[0], y = pr[1], z = pr[2];
            const float* R = rot + rj * 9;
            const float* T = trans + rj * 3;
            float* dst = keys + j * KW + (p < PQ ? O_KP + p * 3 : O_VP + (p - PQ) * 3);
#pragma unroll
            for (int r = 0; r < 3; ++r) dst[r] = ((R[r * 3] * x + R[r * 3 + 1] * y) + R[r * 3 + 2] * z) + T[r];
        }
        for (int idx = threadIdx.x; idx < nk; idx += blockDim.x) kbias[idx] = key_bias ? key_bias[(int64_t)b * L + j0 + idx] : 0.f;
        __syncthreads();
        if (!active) continue;
        for (int jj = 4 * part; jj < nk; jj += 4 * KS) {              // this lane's 4-key chunks of the tile
            float s[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = jj + u;
                if (j < nk) {
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
                        dsum += f_sqrt<FAST>(dx * dx + dy * dy + dz * dz);
                    }
                    s[u] = ((dot + hw * dsum) + __ldg(bias_row + j0 + j)) + kbias[j];
                } else {
                    s[u] = -CUDART_INF_F;
                }
            }
            const float cm = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
            const float m_new = fmaxf(m, cm);
            if (m_new == -CUDART_INF_F) continue;  // everything masked so far
            const float scale = f_exp<FAST>(m - m_new);
            m = m_new;
            l *= scale;
#pragma unroll
            for (int c = 0; c < DK; ++c) { acc_s[c] *= scale; acc_z[c] *= scale; }
#pragma unroll
            for (int c = 0; c < 3 * PV; ++c) acc_p[c] *= scale;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = jj + u;
                if (j < nk) {
                    const float p = f_exp<FAST>(s[u] - m);
                    l += p;
                    const float4* vr = reinterpret_cast<const float4*>(keys + j * KW + O_VS);
#pragma unroll
                    for (int c4 = 0; c4 < DK / 4; ++c4) {
                        const float4 v = vr[c4];
                        acc_s[c4 * 4] += p * v.x; acc_s[c4 * 4 + 1] += p * v.y; acc_s[c4 * 4 + 2] += p * v.z; acc_s[c4 * 4 + 3] += p * v.w;
                    }
#pragma unroll
                    for (int c4 = 0; c4 < 3 * PV / 4; ++c4) {
                        const float4 v = vr[DK / 4 + c4];
                        acc_p[c4 * 4] += p * v.x; acc_p[c4 * 4 + 1] += p * v.y; acc_p[c4 * 4 + 2] += p * v.z; acc_p[c4 * 4 + 3] += p * v.w;
                    }
                    const float4* zr = reinterpret_cast<const float4*>(pv_row + (int64_t)(j0 + j) * H * DK);
#pragma unroll
                    for (int c4 = 0; c4 < DK / 4; ++c4) {
                        const float4 v = __ldg(zr + c4);
                        acc_z[c4 * 4] += p * v.x; acc_z[c4 * 4 + 1] += p * v.y; acc_z[c4 * 4 + 2] += p * v.z; acc_z[c4 * 4 + 3] += p * v.w;
                    }
                }
            }
        }
    }
    // ---- merge the KS partial states of a row (all lanes of the warp take part; rows past L carry empty states) ----------------
#pragma unroll
    for (int d = 1; d < KS; d <<= 1) {
        const float m_o = __shfl_xor_sync(0xffffffffu, m, d), l_o = __shfl_xor_sync(0xffffffffu, l, d);
        const float m_n = fmaxf(m, m_o);
        const float a = m == -CUDART_INF_F ? 0.f : f_exp<FAST>(m - m_n), bq = m_o == -CUDART_INF_F ? 0.f : f_exp<FAST>(m_o - m_n);
        l = l * a + l_o * bq;
#pragma unroll
        for (int c = 0; c < DK; ++c) {
            acc_s[c] = acc_s[c] * a + __shfl_xor_sync(0xffffffffu, acc_s[c], d) * bq;
            acc_z[c] = acc_z[c] * a + __shfl_xor_sync(0xffffffffu, acc_z[c], d) * bq;
        }
#pragma unroll
        for (int c = 0; c < 3 * PV; ++c) acc_p[c] = acc_p[c] * a + __shfl_xor_sync(0xffffffffu, acc_p[c], d) * bq;
        m = m_n;
    }
    if (!active || part != 0) return;
    const float inv = 1.0f / l;
    const int HD = H * DK;
    float* o = out + row_i * (int64_t)(2 * HD + 4 * H * PV);
#pragma unroll
    for (int c = 0; c < DK; ++c) o[h * DK + c] = acc_s[c] * inv;
#pragma unroll
    for (int p = 0; p < PV; ++p) {
        const float dx = acc_p[p * 3] * inv - Ti[0], dy = acc_p[p * 3 + 1] * inv - Ti[1], dz = acc_p[p * 3 + 2] * inv - Ti[2];
        float loc[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) loc[r] = (Ri[r] * dx + Ri[3 + r] * dy) + Ri[6 + r] * dz;
        o[HD + (h * PV + p) * 3] = loc[0];
        o[HD + (h * PV + p) * 3 + 1] = loc[1];
        o[HD + (h * PV + p) * 3 + 2] = loc[2];
        o[2 * HD + 3 * H * PV + h * PV + p] = sqrtf(loc[0] * loc[0] + loc[1] * loc[1] + loc[2] * loc[2]);
    }
#pragma unroll
    for (int c = 0; c < DK; ++c) o[HD + 3 * H * PV + h * DK + c] = acc_z[c] * inv;
}

template <int DK>
int launch(const float* proj, const float* rot, const float* trans, const float* pair_bias, const float* pair_value,
           const float* key_bias, const float* head_weight, float scalar_weight, float* out, const se3_ipa_shape& sh,
           bool fast, cudaStream_t st) {
    constexpr int KW = 2 * DK + 3 * PQ + 3 * PV;
    const int L = sh.len;
    const int threads = L >= 128 ? 128 : ((L + 31) / 32) * 32;
    const int tile_keys = L < 128 ? ((L + 3) / 4) * 4 : 128;
    {
        // small grids: four lanes per query row (k_ipa_rows_split).  SE3DIFF_B200_IPA_SPLIT=0 keeps one thread per row.
        static const bool split_on = [] { const char* v = getenv("SE3DIFF_B200_IPA_SPLIT"); return !(v && v[0] == '0'); }();
        const int64_t ctas = (int64_t)((L + threads - 1) / threads) * sh.heads * sh.batch;
        if (split_on && ctas < 4 * 148 && L >= 16) {
            constexpr int KS = 4;
            const size_t smem_s = (size_t)tile_keys * (KW + 1) * sizeof(float);
            dim3 grid_s((L + 128 / KS - 1) / (128 / KS), sh.heads, sh.batch);
            auto ks = fast ? k_ipa_rows_split<DK, true, KS> : k_ipa_rows_split<DK, false, KS>;
            if (smem_s > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
                if (e != cudaSuccess) { set_error("ipa smem attribute: %s", cudaGetErrorString(e)); return SE3_ECUDA; }
            }
            ks<<<grid_s, 128, smem_s, st>>>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, sh, tile_keys);
            count_launch();
            return check_launch("se3_ipa_attention_fwd");
        }
    }
    const size_t smem = ((size_t)tile_keys * (KW + 1) + (size_t)threads * (8 * DK + 4)) * sizeof(float);
    dim3 grid((L + threads - 1) / threads, sh.heads, sh.batch);
    auto kern = fast ? k_ipa_rows<DK, true> : k_ipa_rows<DK, false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("ipa smem attribute: %s", cudaGetErrorString(e)); return SE3_ECUDA; }
    }
    kern<<<grid, threads, smem, st>>>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, sh, tile_keys);
    count_launch();
    return check_launch("se3_ipa_attention_fwd");
}

}  // namespace

namespace se3 {
int ipa_simt_fwd(const float* proj, const float* rot, const float* trans, const float* pair_bias, const float* pair_value,
                 const float* key_bias, const float* head_weight, float scalar_weight, float* out, const se3_ipa_shape* h,
                 bool fast, se3_stream_t stream) {
    SE3_REQUIRE(h, "null shape");
    const se3_ipa_shape& sh = *h;
    SE3_REQUIRE(sh.batch >= 0 && sh.len >= 0 && sh.heads > 0, "bad shape");
    if (sh.batch == 0 || sh.len == 0) return SE3_OK;
    SE3_REQUIRE(proj && rot && trans && pair_bias && pair_value && head_weight && out, "null pointer");
    SE3_REQUIRE(sh.pq == PQ && sh.pv == PV, "only 4 query/key points and 8 value points (structure_module.py:85-93)");
    SE3_REQUIRE(sh.pair_batch == 1 || sh.pair_batch == sh.batch, "pair_batch must be 1 or batch");
    SE3_REQUIRE(sh.batch <= 65535 && sh.heads <= 65535, "grid limit");
    SE3_REQUIRE((reinterpret_cast<uintptr_t>(pair_value) & 15) == 0, "pair_value must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    switch (sh.dk) {
        case 4: return launch<4>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, sh, fast, st);
        case 8: return launch<8>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, sh, fast, st);
        case 16: return launch<16>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, sh, fast, st);
        case 32: return launch<32>(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, sh, fast, st);
        default: set_error("se3_ipa_attention_fwd: unsupported dk=%d (4, 8, 16, 32)", sh.dk); return SE3_EUNSUPPORTED;
    }
}
}  // namespace se3

extern "C" int se3_ipa_attention_fwd(const float* proj, const float* rot, const float* trans, const float* pair_bias,
                                     const float* pair_value, const float* key_bias, const float* head_weight,
                                     float scalar_weight, float* out, const se3_ipa_shape* h_shape, int flags,
                                     se3_stream_t stream) {
    if (flags != SE3_IPA_EXACT && flags != SE3_IPA_FAST_MATH) {
        se3::set_error("se3_ipa_attention_fwd: unknown flags %d", flags);
        return SE3_EINVAL;
    }
    return se3::ipa_simt_fwd(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, out, h_shape,
                             flags == SE3_IPA_FAST_MATH, stream);
}
