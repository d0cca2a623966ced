// K3 + frame update -- fused per-step SDE algebra of the three samplers (denoiser.py:30-166,
// 245-262, 401-459, 668-762) including the score conversion of _get_score (denoiser.py:169-203).
//
// One thread per residue, one CTA per 256 residues, one WARP per 32: every operand array ([n,9] rotations, [n,3]
// vectors) is moved with 128-bit coalesced accesses through shared memory by the warp that owns the residues
// (common.cuh: warp_tile_load / warp_tile_store), so only __syncwarp() separates load, compute and store.
// HBM-bound: 120-168 algorithmic bytes per residue against ~150 flop + 2 sincos.
// Compiled with -fmad=false: each R3 expression is evaluated in the reference's order, so the R3
// half is bit-identical to the fp32 torch path.  The SO(3) half (which could never be bit-identical: other libm) uses
// explicit FMAs and a short sincos (common.cuh: so3_apply_rotvec_fused) and agrees to ~1e-6.
#include "common.cuh"

using namespace se3;

namespace {

inline dim3 grid_for(int64_t n) { return dim3((unsigned)((n + kTile - 1) / kTile)); }

struct Vec3 { float x, y, z; };

__device__ __forceinline__ Vec3 ld3(const float* s, int t) { return {s[t * 3], s[t * 3 + 1], s[t * 3 + 2]}; }
__device__ __forceinline__ void st3(float* s, int t, Vec3 v) { s[t * 3] = v.x; s[t * 3 + 1] = v.y; s[t * 3 + 2] = v.z; }
__device__ __forceinline__ void ld9(const float* s, int t, float r[9]) {
#pragma unroll
    for (int k = 0; k < 9; ++k) r[k] = s[t * 9 + k];
}
__device__ __forceinline__ void st9(float* s, int t, const float r[9]) {
#pragma unroll
    for (int k = 0; k < 9; ++k) s[t * 9 + k] = r[k];
}

// R . Exp(v)   (apply_rotvec_to_rotmat, so3_sde.py:782-802), fused edition (common.cuh)
__device__ __forceinline__ void apply_rotvec(const float r[9], Vec3 v, float tol, float out[9]) {
    so3_apply_rotvec_fused(r, v.x, v.y, v.z, tol, out);
}

// SO(3) reverse drift (denoiser.py:64-68 with so3_sde.py:173-194): 0 - g^2*score*w [+ g*u*w]
__device__ __forceinline__ float rot_drift(float g, float score, float w, bool has_u, float u) {
    float d = 0.0f - ((g * g) * score) * w;
    if (has_u) d = d + (g * u) * w;
    return d;
}
// R3 reverse drift (denoiser.py:64-68 with sde_lib.py:140-150): -0.5*beta*x - sqrt(beta)^2*score*w [+ sqrt(beta)*u*w]
__device__ __forceinline__ float pos_drift(float beta, float sqb, float x, float score, float w, bool has_u, float u) {
    float d = ((-0.5f * beta) * x) - ((sqb * sqb) * score) * w;
    if (has_u) d = d + (sqb * u) * w;
    return d;
}

// ---------------------------------------------------------------------------------------------
// Euler-Maruyama step (denoiser.py:54-116)
// ---------------------------------------------------------------------------------------------
// One residue of the Euler-Maruyama step, in place in shared memory: s_rot [.][9], sv[0..4(6)] = pos, m_rot, m_pos, z_rot, z_pos
// [, u_rot, u_pos] as [.][3]; results overwrite rot / pos (and z_rot / z_pos with the Brownian increments when OUT_DW)
template <bool HAS_U, bool OUT_DW>
__device__ __forceinline__ void em_element(float* s_rot, float* const* s_v, int t, const se3_em_scalars& c) {
    float r[9], mean[9], out[9];
    ld9(s_rot, t, r);
    const Vec3 x = ld3(s_v[0], t), mr = ld3(s_v[1], t), mp = ld3(s_v[2], t), zr = ld3(s_v[3], t), zp = ld3(s_v[4], t);
    Vec3 ur = {0, 0, 0}, up = {0, 0, 0};
    if (HAS_U) { ur = ld3(s_v[5], t); up = ld3(s_v[6], t); }
    const float w = c.score_weight, g = c.rot_g, nsd = c.noise_weight * c.sqrt_abs_dt;
    // rotations: mean = R.Exp(drift*dt); sample = mean.Exp(g*dW)
    const Vec3 sr = {mr.x * c.rot_scale, mr.y * c.rot_scale, mr.z * c.rot_scale};
    const Vec3 dr = {rot_drift(g, sr.x, w, HAS_U, ur.x), rot_drift(g, sr.y, w, HAS_U, ur.y), rot_drift(g, sr.z, w, HAS_U, ur.z)};
    const Vec3 dwr = {nsd * zr.x, nsd * zr.y, nsd * zr.z};
    apply_rotvec(r, {dr.x * c.dt, dr.y * c.dt, dr.z * c.dt}, c.tol, mean);
    apply_rotvec(mean, {g * dwr.x, g * dwr.y, g * dwr.z}, c.tol, out);
    st9(s_rot, t, out);
    // positions: mean = x + drift*dt; sample = mean + sqrt(beta)*dW
    const Vec3 sp = {mp.x / c.pos_std, mp.y / c.pos_std, mp.z / c.pos_std};
    const float b = c.pos_beta, q = c.pos_sqrt_beta;
    const Vec3 dp = {pos_drift(b, q, x.x, sp.x, w, HAS_U, up.x), pos_drift(b, q, x.y, sp.y, w, HAS_U, up.y),
                     pos_drift(b, q, x.z, sp.z, w, HAS_U, up.z)};
    const Vec3 dwp = {nsd * zp.x, nsd * zp.y, nsd * zp.z};
    st3(s_v[0], t, {(x.x + dp.x * c.dt) + q * dwp.x, (x.y + dp.y * c.dt) + q * dwp.y, (x.z + dp.z * c.dt) + q * dwp.z});
    if (OUT_DW) { st3(s_v[3], t, dwr); st3(s_v[4], t, dwp); }
}

template <bool HAS_U, bool OUT_DW>
__global__ void __launch_bounds__(kTile)
k_em(const float* __restrict__ rot, const float* __restrict__ pos, const float* __restrict__ m_rot,
     const float* __restrict__ m_pos, const float* __restrict__ u_rot, const float* __restrict__ u_pos,
     const float* __restrict__ z_rot, const float* __restrict__ z_pos, float* __restrict__ rot_out,
     float* __restrict__ pos_out, float* __restrict__ dw_rot, float* __restrict__ dw_pos, int64_t n,
     const se3_em_scalars c) {
    __shared__ __align__(16) float s_rot[kTile * 9];
    __shared__ __align__(16) float s_v[HAS_U ? 7 : 5][kTile * 3];  // pos, m_rot, m_pos, z_rot, z_pos, [u_rot, u_pos]
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    warp_tile_load<9>(rot, s_rot, first, count);
    warp_tile_load<3>(pos, s_v[0], first, count);
    warp_tile_load<3>(m_rot, s_v[1], first, count);
    warp_tile_load<3>(m_pos, s_v[2], first, count);
    warp_tile_load<3>(z_rot, s_v[3], first, count);
    warp_tile_load<3>(z_pos, s_v[4], first, count);
    if (HAS_U) {
        warp_tile_load<3>(u_rot, s_v[5], first, count);
        warp_tile_load<3>(u_pos, s_v[6], first, count);
    }
    __syncwarp();
    const int t = threadIdx.x;
    if (t < count) {
        float* sv[7] = {s_v[0], s_v[1], s_v[2], s_v[3], s_v[4], HAS_U ? s_v[5] : nullptr, HAS_U ? s_v[6] : nullptr};
        em_element<HAS_U, OUT_DW>(s_rot, sv, t, c);
    }
    __syncwarp();
    warp_tile_store<9>(rot_out, s_rot, first, count);
    warp_tile_store<3>(pos_out, s_v[0], first, count);
    if (OUT_DW) {
        if (dw_rot) warp_tile_store<3>(dw_rot, s_v[3], first, count);
        if (dw_pos) warp_tile_store<3>(dw_pos, s_v[4], first, count);
    }
}


// Pipelined edition of k_em for large n: every WARP walks 32-residue groups on its own, with the NEXT group's operand tiles in
// flight (cp.async into the other half of a two-stage shared-memory ring) while it computes and stores the current one.
// k_em alone keeps ~half of its resident warps in the compute phase, so the bytes in flight cover ~0.6 of the HBM roof
// (profiles/r1l_elementwise_per_warp_ncu_summary.txt); here a warp always has a group's worth of loads outstanding.
template <bool HAS_U, bool OUT_DW>
__global__ void __launch_bounds__(kTile, 4)
k_em_pipe(const float* __restrict__ rot, const float* __restrict__ pos, const float* __restrict__ m_rot,
          const float* __restrict__ m_pos, const float* __restrict__ u_rot, const float* __restrict__ u_pos,
          const float* __restrict__ z_rot, const float* __restrict__ z_pos, float* __restrict__ rot_out,
          float* __restrict__ pos_out, float* __restrict__ dw_rot, float* __restrict__ dw_pos, int64_t n,
          const se3_em_scalars c) {
    constexpr int NV = HAS_U ? 7 : 5;
    constexpr int kStage = 32 * 9 + NV * 32 * 3;             // floats per stage: one group of every operand
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* wbase = smem + (size_t)warp * 2 * kStage;
    const float* srcs[7] = {pos, m_rot, m_pos, z_rot, z_pos, u_rot, u_pos};
    const int64_t ngroups = (n + 31) >> 5, gstride = (int64_t)gridDim.x * (kTile / 32);
    auto issue = [&](int64_t grp, float* st) {               // the whole warp; full groups only (the launcher sends ragged tails to k_em)
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(st);
        const float4* r4 = reinterpret_cast<const float4*>(rot + grp * 32 * 9);
#pragma unroll
        for (int i = lane; i < 72; i += 32) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + (uint32_t)i * 16u), "l"(r4 + i) : "memory");
#pragma unroll
        for (int a = 0; a < NV; ++a) {
            const float4* v4 = reinterpret_cast<const float4*>(srcs[a] + grp * 32 * 3);
            if (lane < 24) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + (uint32_t)(288 + a * 96) * 4u + (uint32_t)lane * 16u), "l"(v4 + lane) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int64_t g = (int64_t)blockIdx.x * (kTile / 32) + warp;
    if (g < ngroups) issue(g, wbase);
    for (int k = 0; g < ngroups; g += gstride, k ^= 1) {
        float* st = wbase + k * kStage;
        const int64_t next = g + gstride;
        if (next < ngroups) {
            issue(next, wbase + (k ^ 1) * kStage);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        float* sv[7];
#pragma unroll
        for (int a = 0; a < 7; ++a) sv[a] = a < NV ? st + 288 + a * 96 : nullptr;
        em_element<HAS_U, OUT_DW>(st, sv, lane, c);
        __syncwarp();
        float4* ro = reinterpret_cast<float4*>(rot_out + g * 32 * 9);
#pragma unroll
        for (int i = lane; i < 72; i += 32) ro[i] = reinterpret_cast<const float4*>(st)[i];
        if (lane < 24) {
            reinterpret_cast<float4*>(pos_out + g * 32 * 3)[lane] = reinterpret_cast<const float4*>(sv[0])[lane];
            if (OUT_DW) {
                if (dw_rot) reinterpret_cast<float4*>(dw_rot + g * 32 * 3)[lane] = reinterpret_cast<const float4*>(sv[3])[lane];
                if (dw_pos) reinterpret_cast<float4*>(dw_pos + g * 32 * 3)[lane] = reinterpret_cast<const float4*>(sv[4])[lane];
            }
        }
        __syncwarp();                                          // the stage is refilled by the copies issued at the top of the next round
    }
}

// SO(3)-only Euler-Maruyama step: the rotation half of k_em, for samplers whose state is a bare rotation
// (se3diff/train.py:54-70, se3diff/finetune.py:33-56 call EulerMaruyamaPredictor.update_given_score on [B,3,3]).
template <bool HAS_U, bool OUT_DW>
__global__ void __launch_bounds__(kTile)
k_em_so3(const float* __restrict__ rot, const float* __restrict__ m_rot, const float* __restrict__ u_rot, const float* __restrict__ z_rot,
         float* __restrict__ rot_out, float* __restrict__ dw_rot, int64_t n, const se3_em_scalars c) {
    __shared__ __align__(16) float s_rot[kTile * 9];
    __shared__ __align__(16) float s_v[HAS_U ? 3 : 2][kTile * 3];  // m_rot, z_rot, [u_rot]
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    warp_tile_load<9>(rot, s_rot, first, count);
    warp_tile_load<3>(m_rot, s_v[0], first, count);
    warp_tile_load<3>(z_rot, s_v[1], first, count);
    if (HAS_U) warp_tile_load<3>(u_rot, s_v[2], first, count);
    __syncwarp();
    const int t = threadIdx.x;
    if (t < count) {
        float r[9], mean[9], out[9];
        ld9(s_rot, t, r);
        const Vec3 mr = ld3(s_v[0], t), zr = ld3(s_v[1], t);
        Vec3 ur = {0, 0, 0};
        if (HAS_U) ur = ld3(s_v[2], t);
        const float w = c.score_weight, g = c.rot_g, nsd = c.noise_weight * c.sqrt_abs_dt;
        const Vec3 sr = {mr.x * c.rot_scale, mr.y * c.rot_scale, mr.z * c.rot_scale};
        const Vec3 dr = {rot_drift(g, sr.x, w, HAS_U, ur.x), rot_drift(g, sr.y, w, HAS_U, ur.y), rot_drift(g, sr.z, w, HAS_U, ur.z)};
        const Vec3 dwr = {nsd * zr.x, nsd * zr.y, nsd * zr.z};
        apply_rotvec(r, {dr.x * c.dt, dr.y * c.dt, dr.z * c.dt}, c.tol, mean);
        apply_rotvec(mean, {g * dwr.x, g * dwr.y, g * dwr.z}, c.tol, out);
        st9(s_rot, t, out);
        if (OUT_DW) st3(s_v[1], t, dwr);
    }
    __syncwarp();
    warp_tile_store<9>(rot_out, s_rot, first, count);
    if (OUT_DW) warp_tile_store<3>(dw_rot, s_v[1], first, count);
}

// ---------------------------------------------------------------------------------------------
// DPM-Solver-2 (denoiser.py:676-762)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile)
k_dpm_mid(const float* __restrict__ rot, const float* __restrict__ pos, const float* __restrict__ m_rot,
          const float* __restrict__ m_pos, float* __restrict__ rot_u, float* __restrict__ pos_u, int64_t n,
          const se3_dpm_scalars c) {
    __shared__ __align__(16) float s_rot[kTile * 9];
    __shared__ __align__(16) float s_v[3][kTile * 3];
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    warp_tile_load<9>(rot, s_rot, first, count);
    warp_tile_load<3>(pos, s_v[0], first, count);
    warp_tile_load<3>(m_rot, s_v[1], first, count);
    warp_tile_load<3>(m_pos, s_v[2], first, count);
    __syncwarp();
    const int t = threadIdx.x;
    if (t < count) {
        float r[9], out[9];
        ld9(s_rot, t, r);
        const Vec3 x = ld3(s_v[0], t), mr = ld3(s_v[1], t), mp = ld3(s_v[2], t);
        const float g = c.rot_g_t;
        const Vec3 d = {rot_drift(g, mr.x * c.rot_scale_t, 0.5f, false, 0.f), rot_drift(g, mr.y * c.rot_scale_t, 0.5f, false, 0.f),
                        rot_drift(g, mr.z * c.rot_scale_t, 0.5f, false, 0.f)};
        apply_rotvec(r, {d.x * c.dt_mid, d.y * c.dt_mid, d.z * c.dt_mid}, c.tol, out);
        st9(s_rot, t, out);
        st3(s_v[0], t, {c.pos_c_x_mid * x.x + c.pos_c_s_mid * (mp.x / c.pos_std_t),
                        c.pos_c_x_mid * x.y + c.pos_c_s_mid * (mp.y / c.pos_std_t),
                        c.pos_c_x_mid * x.z + c.pos_c_s_mid * (mp.z / c.pos_std_t)});
    }
    __syncwarp();
    warp_tile_store<9>(rot_u, s_rot, first, count);
    warp_tile_store<3>(pos_u, s_v[0], first, count);
}

__global__ void __launch_bounds__(kTile)
k_dpm_final(const float* __restrict__ rot, const float* __restrict__ pos, const float* __restrict__ m_rot_t,
            const float* __restrict__ m_rot_l, const float* __restrict__ m_pos_l, float* __restrict__ rot_out,
            float* __restrict__ pos_out, int64_t n, const se3_dpm_scalars c) {
    __shared__ __align__(16) float s_rot[kTile * 9];
    __shared__ __align__(16) float s_v[4][kTile * 3];
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    warp_tile_load<9>(rot, s_rot, first, count);
    warp_tile_load<3>(pos, s_v[0], first, count);
    warp_tile_load<3>(m_rot_t, s_v[1], first, count);
    warp_tile_load<3>(m_rot_l, s_v[2], first, count);
    warp_tile_load<3>(m_pos_l, s_v[3], first, count);
    __syncwarp();
    const int t = threadIdx.x;
    if (t < count) {
        float r[9], out[9];
        ld9(s_rot, t, r);
        const Vec3 x = ld3(s_v[0], t), m1 = ld3(s_v[1], t), m2 = ld3(s_v[2], t), mp = ld3(s_v[3], t);
        const float g = c.rot_g_lam;
        float dv[3];
        const float a1[3] = {m1.x, m1.y, m1.z}, a2[3] = {m2.x, m2.y, m2.z};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float s1 = a1[k] * c.rot_scale_t, s2 = a2[k] * c.rot_scale_lam;
            // node_score = s2 + 0.5*(s2 - s1)/(t_lambda - t)*dt            denoiser.py:741-747
            const float ns = s2 + ((0.5f * (s2 - s1)) / c.dt_mid) * c.dt;
            dv[k] = rot_drift(g, ns, 0.5f, false, 0.f) * c.dt;
        }
        apply_rotvec(r, {dv[0], dv[1], dv[2]}, c.tol, out);
        st9(s_rot, t, out);
        st3(s_v[0], t, {c.pos_c_x_fin * x.x + c.pos_c_s_fin * (mp.x / c.pos_std_lam),
                        c.pos_c_x_fin * x.y + c.pos_c_s_fin * (mp.y / c.pos_std_lam),
                        c.pos_c_x_fin * x.z + c.pos_c_s_fin * (mp.z / c.pos_std_lam)});
    }
    __syncwarp();
    warp_tile_store<9>(rot_out, s_rot, first, count);
    warp_tile_store<3>(pos_out, s_v[0], first, count);
}

// ---------------------------------------------------------------------------------------------
// Heun (denoiser.py:401-459)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile)
k_heun_churn(const float* __restrict__ rot, const float* __restrict__ pos, const float* __restrict__ z_rot,
             const float* __restrict__ z_pos, float* __restrict__ rot_hat, float* __restrict__ pos_hat, int64_t n,
             const se3_heun_scalars c) {
    __shared__ __align__(16) float s_rot[kTile * 9];
    __shared__ __align__(16) float s_v[3][kTile * 3];
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    warp_tile_load<9>(rot, s_rot, first, count);
    warp_tile_load<3>(pos, s_v[0], first, count);
    warp_tile_load<3>(z_rot, s_v[1], first, count);
    warp_tile_load<3>(z_pos, s_v[2], first, count);
    __syncwarp();
    const int t = threadIdx.x;
    if (t < count) {
        float r[9], mean[9], out[9];
        ld9(s_rot, t, r);
        const Vec3 x = ld3(s_v[0], t), zr = ld3(s_v[1], t), zp = ld3(s_v[2], t);
        const float nsd = 1.0f * c.churn_sqrt_abs_dt, g = c.churn_rot_g, q = c.churn_pos_sqrt_beta;
        // forward SDE: SO(3) drift is the zero vector, R3 drift is -0.5*beta*x
        apply_rotvec(r, {0.0f * c.churn_dt, 0.0f * c.churn_dt, 0.0f * c.churn_dt}, c.tol, mean);
        apply_rotvec(mean, {g * (nsd * zr.x), g * (nsd * zr.y), g * (nsd * zr.z)}, c.tol, out);
        st9(s_rot, t, out);
        const float hb = -0.5f * c.churn_pos_beta;
        st3(s_v[0], t, {(x.x + (hb * x.x) * c.churn_dt) + q * (nsd * zp.x), (x.y + (hb * x.y) * c.churn_dt) + q * (nsd * zp.y),
                        (x.z + (hb * x.z) * c.churn_dt) + q * (nsd * zp.z)});
    }
    __syncwarp();
    warp_tile_store<9>(rot_hat, s_rot, first, count);
    warp_tile_store<3>(pos_hat, s_v[0], first, count);
}

// CORRECT=false: first-order step from (rot_hat,pos_hat) with the drift at t_hat.
// CORRECT=true : same start point, drift = (drift(t_next; pos_pred, m_next) + drift(t_hat))/2.
template <bool CORRECT>
__global__ void __launch_bounds__(kTile)
k_heun_step(const float* __restrict__ rot_hat, const float* __restrict__ pos_hat, const float* __restrict__ m_rot_hat,
            const float* __restrict__ m_pos_hat, const float* __restrict__ pos_pred, const float* __restrict__ m_rot_next,
            const float* __restrict__ m_pos_next, float* __restrict__ rot_out, float* __restrict__ pos_out, int64_t n,
            const se3_heun_scalars c) {
    __shared__ __align__(16) float s_rot[kTile * 9];
    __shared__ __align__(16) float s_v[CORRECT ? 6 : 3][kTile * 3];
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    warp_tile_load<9>(rot_hat, s_rot, first, count);
    warp_tile_load<3>(pos_hat, s_v[0], first, count);
    warp_tile_load<3>(m_rot_hat, s_v[1], first, count);
    warp_tile_load<3>(m_pos_hat, s_v[2], first, count);
    if (CORRECT) {
        warp_tile_load<3>(pos_pred, s_v[3], first, count);
        warp_tile_load<3>(m_rot_next, s_v[4], first, count);
        warp_tile_load<3>(m_pos_next, s_v[5], first, count);
    }
    __syncwarp();
    const int t = threadIdx.x;
    if (t < count) {
        float r[9], out[9];
        ld9(s_rot, t, r);
        const Vec3 xv = ld3(s_v[0], t), mrv = ld3(s_v[1], t), mpv = ld3(s_v[2], t);
        const float x[3] = {xv.x, xv.y, xv.z}, mr[3] = {mrv.x, mrv.y, mrv.z}, mp[3] = {mpv.x, mpv.y, mpv.z};
        float x1[3] = {0, 0, 0}, mr1[3] = {0, 0, 0}, mp1[3] = {0, 0, 0};
        if (CORRECT) {
            const Vec3 a = ld3(s_v[3], t), b = ld3(s_v[4], t), d = ld3(s_v[5], t);
            x1[0] = a.x; x1[1] = a.y; x1[2] = a.z; mr1[0] = b.x; mr1[1] = b.y; mr1[2] = b.z; mp1[0] = d.x; mp1[1] = d.y; mp1[2] = d.z;
        }
        float dr[3], xo[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float d_r = rot_drift(c.hat_rot_g, mr[k] * c.hat_rot_scale, 0.5f, false, 0.f);
            float d_p = pos_drift(c.hat_pos_beta, c.hat_pos_sqrt_beta, x[k], mp[k] / c.hat_pos_std, 0.5f, false, 0.f);
            if (CORRECT) {
                const float n_r = rot_drift(c.next_rot_g, mr1[k] * c.next_rot_scale, 0.5f, false, 0.f);
                const float n_p = pos_drift(c.next_pos_beta, c.next_pos_sqrt_beta, x1[k], mp1[k] / c.next_pos_std, 0.5f, false, 0.f);
                d_r = (n_r + d_r) / 2.0f;  // denoiser.py:452
                d_p = (n_p + d_p) / 2.0f;
            }
            dr[k] = d_r * c.step_dt;
            xo[k] = x[k] + d_p * c.step_dt;
        }
        apply_rotvec(r, {dr[0], dr[1], dr[2]}, c.tol, out);
        st9(s_rot, t, out);
        st3(s_v[0], t, {xo[0], xo[1], xo[2]});
    }
    __syncwarp();
    warp_tile_store<9>(rot_out, s_rot, first, count);
    warp_tile_store<3>(pos_out, s_v[0], first, count);
}

// traceback_brownian_motion (denoiser.py:133-166)
template <bool HAS_U>
__global__ void __launch_bounds__(kTile)
k_traceback(const float* __restrict__ rot, const float* __restrict__ pos, const float* __restrict__ rot_next,
            const float* __restrict__ pos_next, const float* __restrict__ m_rot, const float* __restrict__ m_pos,
            const float* __restrict__ u_rot, const float* __restrict__ u_pos, float* __restrict__ dw_rot,
            float* __restrict__ dw_pos, int64_t n, const se3_em_scalars c) {
    __shared__ __align__(16) float s_rot[2][kTile * 9];
    __shared__ __align__(16) float s_v[HAS_U ? 6 : 4][kTile * 3];
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    warp_tile_load<9>(rot, s_rot[0], first, count);
    warp_tile_load<9>(rot_next, s_rot[1], first, count);
    warp_tile_load<3>(pos, s_v[0], first, count);
    warp_tile_load<3>(pos_next, s_v[1], first, count);
    warp_tile_load<3>(m_rot, s_v[2], first, count);
    warp_tile_load<3>(m_pos, s_v[3], first, count);
    if (HAS_U) { warp_tile_load<3>(u_rot, s_v[4], first, count); warp_tile_load<3>(u_pos, s_v[5], first, count); }
    __syncwarp();
    const int t = threadIdx.x;
    if (t < count) {
        float r[9], rn[9], mean[9], rel[9], lg[3];
        ld9(s_rot[0], t, r);
        ld9(s_rot[1], t, rn);
        const Vec3 x = ld3(s_v[0], t), xn = ld3(s_v[1], t), mr = ld3(s_v[2], t), mp = ld3(s_v[3], t);
        Vec3 ur = {0, 0, 0}, up = {0, 0, 0};
        if (HAS_U) { ur = ld3(s_v[4], t); up = ld3(s_v[5], t); }
        const float w = c.score_weight, g = c.rot_g, b = c.pos_beta, q = c.pos_sqrt_beta;
        const Vec3 dr = {rot_drift(g, mr.x * c.rot_scale, w, HAS_U, ur.x), rot_drift(g, mr.y * c.rot_scale, w, HAS_U, ur.y),
                         rot_drift(g, mr.z * c.rot_scale, w, HAS_U, ur.z)};
        apply_rotvec(r, {dr.x * c.dt, dr.y * c.dt, dr.z * c.dt}, c.tol, mean);
        so3_mul<float, true>(mean, rn, rel);
        so3_log(rel, lg);
        st3(s_v[2], t, {lg[0] / g, lg[1] / g, lg[2] / g});
        const Vec3 dp = {pos_drift(b, q, x.x, mp.x / c.pos_std, w, HAS_U, up.x), pos_drift(b, q, x.y, mp.y / c.pos_std, w, HAS_U, up.y),
                         pos_drift(b, q, x.z, mp.z / c.pos_std, w, HAS_U, up.z)};
        st3(s_v[3], t, {(xn.x - (x.x + dp.x * c.dt)) / q, (xn.y - (x.y + dp.y * c.dt)) / q, (xn.z - (x.z + dp.z * c.dt)) / q});
    }
    __syncwarp();
    warp_tile_store<3>(dw_rot, s_v[2], first, count);
    warp_tile_store<3>(dw_pos, s_v[3], first, count);
}

}  // namespace

#define SE3_LAUNCH_CHECK(name) \
    count_launch();            \
    return check_launch(name)

extern "C" {

int se3_frame_update_em(const float* rot, const float* pos, const float* m_rot, const float* m_pos,
                        const float* u_rot, const float* u_pos, const float* z_rot, const float* z_pos,
                        float* rot_out, float* pos_out, float* dw_rot, float* dw_pos, int64_t n,
                        const se3_em_scalars* h, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(rot && pos && m_rot && m_pos && z_rot && z_pos && rot_out && pos_out, "null pointer");
    SE3_REQUIRE((u_rot == nullptr) == (u_pos == nullptr), "u_rot and u_pos must be given together");
    cudaStream_t st = (cudaStream_t)stream;
    const bool has_u = u_rot != nullptr, out_dw = dw_rot != nullptr || dw_pos != nullptr;
    // Large inputs: the pipelined edition over the whole groups of 32 (every pointer 16-byte aligned), k_em for the rest.
    int64_t n_pipe = 0;
    {
        uintptr_t bits = 0;
        for (const void* p : {(const void*)rot, (const void*)pos, (const void*)m_rot, (const void*)m_pos, (const void*)u_rot, (const void*)u_pos, (const void*)z_rot,
                              (const void*)z_pos, (const void*)rot_out, (const void*)pos_out, (const void*)dw_rot, (const void*)dw_pos})
            bits |= reinterpret_cast<uintptr_t>(p);
        if ((bits & 15) == 0 && n >= 148 * 4 * (int64_t)kTile * 4) n_pipe = n & ~(int64_t)31;
    }
    if (n_pipe) {
        const size_t smem = (size_t)(kTile / 32) * 2 * (32 * 9 + (has_u ? 7 : 5) * 32 * 3) * sizeof(float);
        const int64_t want = (n_pipe / 32 + kTile / 32 - 1) / (kTile / 32);
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
#define EM_PIPE(U, D)                                                                                                            \
        do {                                                                                                                    \
            int occ = 0;                                                                                                        \
            cudaFuncSetAttribute(k_em_pipe<U, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_em_pipe<U, D>, kTile, smem) != cudaSuccess || occ < 1) occ = 1; \
            const unsigned grid = (unsigned)(want < (int64_t)sms * occ ? want : (int64_t)sms * occ);   /* one resident wave of persistent CTAs */ \
            k_em_pipe<U, D><<<grid, kTile, smem, st>>>(rot, pos, m_rot, m_pos, u_rot, u_pos, z_rot, z_pos, rot_out, pos_out, dw_rot, dw_pos, n_pipe, *h); \
        } while (0)
        if (has_u && out_dw) EM_PIPE(true, true);
        else if (has_u) EM_PIPE(true, false);
        else if (out_dw) EM_PIPE(false, true);
        else EM_PIPE(false, false);
#undef EM_PIPE
        count_launch();
        if (int rc = check_launch("se3_frame_update_em(pipelined)")) return rc;
        if (n_pipe == n) return SE3_OK;
    }
    const int64_t o = n_pipe, m = n - n_pipe;                     // the remainder (everything for small or unaligned inputs)
#define EM_ARGS rot + o * 9, pos + o * 3, m_rot + o * 3, m_pos + o * 3, u_rot ? u_rot + o * 3 : nullptr, u_pos ? u_pos + o * 3 : nullptr, z_rot + o * 3, \
                z_pos + o * 3, rot_out + o * 9, pos_out + o * 3, dw_rot ? dw_rot + o * 3 : nullptr, dw_pos ? dw_pos + o * 3 : nullptr, m, *h
    if (has_u && out_dw) k_em<true, true><<<grid_for(m), kTile, 0, st>>>(EM_ARGS);
    else if (has_u) k_em<true, false><<<grid_for(m), kTile, 0, st>>>(EM_ARGS);
    else if (out_dw) k_em<false, true><<<grid_for(m), kTile, 0, st>>>(EM_ARGS);
    else k_em<false, false><<<grid_for(m), kTile, 0, st>>>(EM_ARGS);
#undef EM_ARGS
    SE3_LAUNCH_CHECK("se3_frame_update_em");
}

int se3_so3_update_em(const float* rot, const float* m_rot, const float* u_rot, const float* z_rot, float* rot_out, float* dw_rot,
                      int64_t n, const se3_em_scalars* h, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(rot && m_rot && z_rot && rot_out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const bool has_u = u_rot != nullptr, out_dw = dw_rot != nullptr;
#define EM_ARGS rot, m_rot, u_rot, z_rot, rot_out, dw_rot, n, *h
    if (has_u && out_dw) k_em_so3<true, true><<<grid_for(n), kTile, 0, st>>>(EM_ARGS);
    else if (has_u) k_em_so3<true, false><<<grid_for(n), kTile, 0, st>>>(EM_ARGS);
    else if (out_dw) k_em_so3<false, true><<<grid_for(n), kTile, 0, st>>>(EM_ARGS);
    else k_em_so3<false, false><<<grid_for(n), kTile, 0, st>>>(EM_ARGS);
#undef EM_ARGS
    SE3_LAUNCH_CHECK("se3_so3_update_em");
}

int se3_frame_update_dpm_mid(const float* rot, const float* pos, const float* m_rot, const float* m_pos,
                             float* rot_u, float* pos_u, int64_t n, const se3_dpm_scalars* h, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(rot && pos && m_rot && m_pos && rot_u && pos_u, "null pointer");
    k_dpm_mid<<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rot, pos, m_rot, m_pos, rot_u, pos_u, n, *h);
    SE3_LAUNCH_CHECK("se3_frame_update_dpm_mid");
}

int se3_frame_update_dpm_final(const float* rot, const float* pos, const float* m_rot_t, const float* m_rot_lam,
                               const float* m_pos_lam, float* rot_out, float* pos_out, int64_t n,
                               const se3_dpm_scalars* h, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(rot && pos && m_rot_t && m_rot_lam && m_pos_lam && rot_out && pos_out, "null pointer");
    k_dpm_final<<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rot, pos, m_rot_t, m_rot_lam, m_pos_lam, rot_out, pos_out, n, *h);
    SE3_LAUNCH_CHECK("se3_frame_update_dpm_final");
}

int se3_frame_heun_churn(const float* rot, const float* pos, const float* z_rot, const float* z_pos, float* rot_hat,
                         float* pos_hat, int64_t n, const se3_heun_scalars* h, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(rot && pos && z_rot && z_pos && rot_hat && pos_hat, "null pointer");
    k_heun_churn<<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rot, pos, z_rot, z_pos, rot_hat, pos_hat, n, *h);
    SE3_LAUNCH_CHECK("se3_frame_heun_churn");
}

int se3_frame_heun_predict(const float* rot_hat, const float* pos_hat, const float* m_rot_hat, const float* m_pos_hat,
                           float* rot_out, float* pos_out, int64_t n, const se3_heun_scalars* h, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(rot_hat && pos_hat && m_rot_hat && m_pos_hat && rot_out && pos_out, "null pointer");
    k_heun_step<false><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rot_hat, pos_hat, m_rot_hat, m_pos_hat, nullptr, nullptr,
                                                                        nullptr, rot_out, pos_out, n, *h);
    SE3_LAUNCH_CHECK("se3_frame_heun_predict");
}

int se3_frame_heun_correct(const float* rot_hat, const float* pos_hat, const float* m_rot_hat, const float* m_pos_hat,
                           const float* pos_pred, const float* m_rot_next, const float* m_pos_next, float* rot_out,
                           float* pos_out, int64_t n, const se3_heun_scalars* h, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(rot_hat && pos_hat && m_rot_hat && m_pos_hat && pos_pred && m_rot_next && m_pos_next && rot_out && pos_out,
                "null pointer");
    k_heun_step<true><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rot_hat, pos_hat, m_rot_hat, m_pos_hat, pos_pred,
                                                                       m_rot_next, m_pos_next, rot_out, pos_out, n, *h);
    SE3_LAUNCH_CHECK("se3_frame_heun_correct");
}

int se3_frame_traceback(const float* rot, const float* pos, const float* rot_next, const float* pos_next,
                        const float* m_rot, const float* m_pos, const float* u_rot, const float* u_pos, float* dw_rot,
                        float* dw_pos, int64_t n, const se3_em_scalars* h, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(rot && pos && rot_next && pos_next && m_rot && m_pos && dw_rot && dw_pos, "null pointer");
    SE3_REQUIRE((u_rot == nullptr) == (u_pos == nullptr), "u_rot and u_pos must be given together");
    if (u_rot) k_traceback<true><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rot, pos, rot_next, pos_next, m_rot, m_pos, u_rot, u_pos, dw_rot, dw_pos, n, *h);
    else k_traceback<false><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rot, pos, rot_next, pos_next, m_rot, m_pos, u_rot, u_pos, dw_rot, dw_pos, n, *h);
    SE3_LAUNCH_CHECK("se3_frame_traceback");
}

}  // extern "C"
