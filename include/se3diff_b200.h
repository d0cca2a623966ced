/*
 * se3diff_b200 -- C ABI of the B200 (sm_100a) kernels behind the SE(3) reverse-diffusion
 * sampling path of ddrichman/SE3Diff (vendored BioEmu 0.1.12).
 *
 * Boundary rules:
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless the
 *     parameter is named `h_*`;
 *   - the caller owns and allocates every buffer; kernels never allocate and never synchronise
 *     the host (CUDA-graph capturable); work is enqueued on `stream` (a cudaStream_t cast to void*);
 *   - return 0 on success, a negative SE3_E* code otherwise, text via se3_last_error();
 *   - no global mutable state other than the thread-local last-error string.
 *
 * Each entry point cites the reference interface it replaces as file:line relative to
 * /root/reference/bioemu/src/bioemu/ .  Arrays use the reference's own memory layout: rotations
 * are row-major [n,3,3] fp32 (ChemGraph.node_orientations), vectors [n,3] fp32 (ChemGraph.pos and
 * the axis-angle scores).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 */
#ifndef SE3DIFF_B200_H
#define SE3DIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SE3_OK 0
#define SE3_EINVAL (-1)   /* bad argument (null pointer, negative size, unsupported shape) */
#define SE3_ECUDA (-2)    /* CUDA runtime error; see se3_last_error() */
#define SE3_EUNSUPPORTED (-3)

typedef void* se3_stream_t; /* cudaStream_t */

const char* se3_last_error(void);
/* bumped whenever a signature or a by-value struct of this header changes; the loader compares it with the header it was written against */
#define SE3_ABI_VERSION 5
int se3_abi_version(void);
/* number of kernels launched by this library in the calling thread since the last reset
 * (bench.py's `gpu_launches`). */
int64_t se3_launch_count(void);
void se3_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------
 * K2 -- SO(3) exp / log / composition                         so3_sde.py:478-554, 557-676, 782-911
 * ------------------------------------------------------------------------------------------ */
/* rotvec_to_rotmat (so3_sde.py:533-554): Rodrigues, Taylor coefficients when |theta| < tol. */
int se3_so3_exp(const float* rotvec, float* rotmat, int64_t n, float tol, se3_stream_t stream);
int se3_so3_exp_f64(const double* rotvec, double* rotmat, int64_t n, double tol, se3_stream_t stream);
/* rotmat_to_rotvec (so3_sde.py:557-648): atan2 angle, three regimes with the reference's isclose
 * thresholds (|theta| <= 1e-8 ; |theta-pi| <= 1e-2 + 1e-5*pi ; else). */
int se3_so3_log(const float* rotmat, float* rotvec, int64_t n, se3_stream_t stream);
int se3_so3_log_f64(const double* rotmat, double* rotvec, int64_t n, se3_stream_t stream);
/* angle_from_rotmat (so3_sde.py:651-676): any of angle/sin/cos may be NULL. */
int se3_so3_angle(const float* rotmat, float* angle, float* sin_out, float* cos_out, int64_t n,
                  se3_stream_t stream);
/* apply_rotvec_to_rotmat (so3_sde.py:782-802): out = R . Exp(v).  `out` may alias `rotmat`. */
int se3_so3_compose_rotvec(const float* rotmat, const float* rotvec, float* out, int64_t n, float tol,
                           se3_stream_t stream);
/* rot_mult with optional transpose of the left factor (so3_sde.py:870-877): out = op(A) . B */
int se3_so3_matmul(const float* a, const float* b, float* out, int64_t n, int transpose_a,
                   se3_stream_t stream);
/* rot_vf (so3_sde.py:880-891): out = Log(base^T . target) */
int se3_so3_rel_log(const float* base, const float* target, float* rotvec, int64_t n, se3_stream_t stream);
/* scale_rotmat / geodesic_t (so3_sde.py:406-425, 894-911): out = base . Exp(t * Log(base^T . target)) */
int se3_so3_geodesic(const float* base, const float* target, float t, float* out, int64_t n, float tol,
                     se3_stream_t stream);
/* rotquat_to_rotvec / rotquat_to_rotmat (so3_sde.py:725-779); quaternion [r,i,j,k]; either output
 * may be NULL. */
int se3_so3_from_quat(const float* quat, float* rotvec, float* rotmat, int64_t n, float tol,
                      se3_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3 + frame update -- the per-step SDE algebra fused with the score conversion of
 * _get_score (denoiser.py:169-203).  `m_rot`/`m_pos` are the RAW score-model outputs
 * (out["node_orientations"], out["pos"]); the conversion rot_score = m_rot*rot_scale and
 * pos_score = m_pos/pos_std happens in-kernel.  All per-graph schedule quantities are identical
 * for every graph of a batch on this path (one t per step), so they travel as by-value scalars
 * computed on the host in fp32 (0 bytes of HBM traffic, graph-capturable).
 * ------------------------------------------------------------------------------------------ */
typedef struct se3_em_scalars {
    float dt;            /* dts[i]                                   denoiser.py:232,258       */
    float sqrt_abs_dt;   /* sqrt(|dt|)                               denoiser.py:81            */
    float noise_weight;  /* EulerMaruyamaPredictor.noise_weight      denoiser.py:51            */
    float score_weight;  /* 0.5*mcf*(1+nw^2)                         denoiser.py:64            */
    float rot_g;         /* DiGSO3SDE.beta(t) = g(t)                 so3_sde.py:346-363        */
    float rot_scale;     /* get_score_scaling(t)                     so3_sde.py:142-161        */
    float pos_beta;      /* CosineVPSDE.beta(t)                      sde_lib.py:160-162        */
    float pos_sqrt_beta; /* sqrt(beta(t))  (the R3 diffusion)        sde_lib.py:148            */
    float pos_std;       /* sqrt(1-alpha(t)^2)                       sde_lib.py:128            */
    float tol;           /* SO3SDE.tol                               so3_sde.py:78             */
} se3_em_scalars;

/* One Euler-Maruyama reverse step on both fields (EulerMaruyamaPredictor.update_given_score,
 * denoiser.py:54-116; loop body denoiser.py:254-262 and 324-338).
 *   u_rot/u_pos   : optional control from the fine-tune model (finetune_score), NULL for plain EM
 *   z_rot/z_pos   : standard normals [n,3] (the randn_like of denoiser.py:80), drawn by the caller
 *   dw_rot/dw_pos : optional outputs dW = nw*sqrt|dt|*z (denoiser.py:81,335), NULL to skip
 * rot_out/pos_out may alias rot/pos. */
int se3_frame_update_em(const float* rot, const float* pos, const float* m_rot, const float* m_pos,
                        const float* u_rot, const float* u_pos, const float* z_rot, const float* z_pos,
                        float* rot_out, float* pos_out, float* dw_rot, float* dw_pos, int64_t n,
                        const se3_em_scalars* h_scalars, se3_stream_t stream);

/* The rotation half alone, for samplers whose state is a bare [n,3,3] rotation (the se3diff toy: se3diff/train.py:54-70,
 * se3diff/finetune.py:33-56).  The pos_* members of the scalars are ignored.  u_rot, dw_rot optional. */
int se3_so3_update_em(const float* rot, const float* m_rot, const float* u_rot, const float* z_rot, float* rot_out,
                      float* dw_rot, int64_t n, const se3_em_scalars* h_scalars, se3_stream_t stream);

typedef struct se3_dpm_scalars {
    float pos_std_t;      /* sigma_t = sqrt(1-alpha_t^2)              denoiser.py:677          */
    float pos_c_x_mid;    /* alpha_lambda/alpha_t                     denoiser.py:700          */
    float pos_c_s_mid;    /* sigma_lambda*sigma_t*(exp(h/2)-1)        denoiser.py:701          */
    float pos_std_lam;    /* sigma_lambda                             denoiser.py:693          */
    float pos_c_x_fin;    /* alpha_next/alpha_t                       denoiser.py:734          */
    float pos_c_s_fin;    /* sigma_next*sigma_lambda*(exp(h)-1)       denoiser.py:735          */
    float rot_scale_t;    /* get_score_scaling(t)                                              */
    float rot_scale_lam;  /* get_score_scaling(t_lambda)                                       */
    float rot_g_t;        /* g(t)                                                              */
    float rot_g_lam;      /* g(t_lambda)                                                       */
    float dt_mid;         /* (t_lambda - t)[0]                        denoiser.py:721,745      */
    float dt;             /* dts[i]                                   denoiser.py:746,756      */
    float tol;
} se3_dpm_scalars;

/* DPM-Solver-2 first half (denoiser.py:699-727): pos_u = c_x*pos + c_s*(m_pos/std_t),
 * rot_u = rot . Exp(-0.5*g_t^2*(m_rot*scale_t) * dt_mid). */
int se3_frame_update_dpm_mid(const float* rot, const float* pos, const float* m_rot, const float* m_pos,
                             float* rot_u, float* pos_u, int64_t n, const se3_dpm_scalars* h_scalars,
                             se3_stream_t stream);
/* DPM-Solver-2 second half (denoiser.py:733-762): needs the first call's raw rotation output
 * (m_rot_t) and the second call's outputs at (pos_u, rot_u, t_lambda); updates from the ORIGINAL
 * rot/pos.  rot_out/pos_out may alias rot/pos. */
int se3_frame_update_dpm_final(const float* rot, const float* pos, const float* m_rot_t,
                               const float* m_rot_lam, const float* m_pos_lam, float* rot_out,
                               float* pos_out, int64_t n, const se3_dpm_scalars* h_scalars,
                               se3_stream_t stream);

typedef struct se3_heun_scalars {
    /* churn t -> t_hat (forward SDE step, denoiser.py:413-418, 118-131) */
    float churn_dt;         /* (t_hat - t)[0] */
    float churn_sqrt_abs_dt;
    float churn_rot_g;      /* g(t) */
    float churn_pos_beta;   /* beta(t) */
    float churn_pos_sqrt_beta;
    /* deterministic Euler step t_hat -> t_next (denoiser.py:423-437), noise_weight 0 => w = 0.5 */
    float step_dt;          /* (t_next - t_hat)[0] */
    float hat_rot_g, hat_rot_scale, hat_pos_beta, hat_pos_sqrt_beta, hat_pos_std;
    /* second-order correction at t_next (denoiser.py:440-459) */
    float next_rot_g, next_rot_scale, next_pos_beta, next_pos_sqrt_beta, next_pos_std;
    float tol;
} se3_heun_scalars;

int se3_frame_heun_churn(const float* rot, const float* pos, const float* z_rot, const float* z_pos,
                         float* rot_hat, float* pos_hat, int64_t n, const se3_heun_scalars* h_scalars,
                         se3_stream_t stream);
int se3_frame_heun_predict(const float* rot_hat, const float* pos_hat, const float* m_rot_hat,
                           const float* m_pos_hat, float* rot_out, float* pos_out, int64_t n,
                           const se3_heun_scalars* h_scalars, se3_stream_t stream);
/* pos_pred is the predictor output (needed by the R3 drift at t_next); m_*_next the raw model output
 * evaluated at (rot_pred, pos_pred, t_next). */
int se3_frame_heun_correct(const float* rot_hat, const float* pos_hat, const float* m_rot_hat,
                           const float* m_pos_hat, const float* pos_pred, const float* m_rot_next,
                           const float* m_pos_next, float* rot_out, float* pos_out, int64_t n,
                           const se3_heun_scalars* h_scalars, se3_stream_t stream);

/* traceback_brownian_motion (denoiser.py:133-166): dW_rot = Log(mean^T x_next)/g with
 * mean = rot . Exp(drift*dt); dW_pos = (x_next - mean)/sqrt(beta).  u_* optional (NULL). */
int se3_frame_traceback(const float* rot, const float* pos, const float* rot_next, const float* pos_next,
                        const float* m_rot, const float* m_pos, const float* u_rot, const float* u_pos,
                        float* dw_rot, float* dw_pos, int64_t n, const se3_em_scalars* h_scalars,
                        se3_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Translation (R^3) updates alone -- the position half of the steps above on bare [n, 3] positions, for callers that step
 * the two fields separately (EulerMaruyamaPredictor with a CosineVPSDE corruption, denoiser.py:72-97; the position lines
 * of dpm_solver :699-701, 733-735 and of heun_denoiser :413-459).  Same expressions, same rounding: results are
 * bit-identical to the pos outputs of the fused kernels.  pos_out may alias pos.  Scalar structs as above (the rotation
 * members are ignored).
 * ------------------------------------------------------------------------------------------ */
/* pos_out = (pos + drift*dt) + sqrt(beta)*dW, drift = -0.5*beta*pos - beta*(m_pos/std)*w [+ sqrt(beta)*u_pos*w],
 * dW = noise_weight*sqrt|dt|*z_pos; u_pos, dw_pos optional (NULL). */
int se3_r3_update_em(const float* pos, const float* m_pos, const float* u_pos, const float* z_pos, float* pos_out,
                     float* dw_pos, int64_t n, const se3_em_scalars* h_scalars, se3_stream_t stream);
/* DPM-Solver-2 position update: final_half = 0: c_x_mid*pos + c_s_mid*(m_pos/std_t) (denoiser.py:699-701);
 * final_half = 1: c_x_fin*pos + c_s_fin*(m_pos/std_lambda) with m_pos the model output at t_lambda (:733-735). */
int se3_r3_update_dpm(const float* pos, const float* m_pos, float* pos_out, int64_t n, const se3_dpm_scalars* h_scalars,
                      int final_half, se3_stream_t stream);
/* Heun: churn t -> t_hat (forward SDE step, noise weight 1, denoiser.py:413-418); first-order step from pos_hat
 * (pos_pred = m_pos_next = NULL, :423-437) or the corrected step averaging the drifts at t_hat and at
 * (t_next, pos_pred) (:440-459). */
int se3_r3_heun_churn(const float* pos, const float* z_pos, float* pos_hat, int64_t n, const se3_heun_scalars* h_scalars,
                      se3_stream_t stream);
int se3_r3_heun_step(const float* pos_hat, const float* m_pos_hat, const float* pos_pred, const float* m_pos_next,
                     float* pos_out, int64_t n, const se3_heun_scalars* h_scalars, se3_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * K1 -- IGSO(3): truncated series, lookup tables, inverse-CDF sampling       so3_sde.py:993-2042
 * ------------------------------------------------------------------------------------------ */
/* igso3_expansion / digso3_expansion / dlog_igso3_expansion (so3_sde.py:1731-1940) for n
 * (omega, sigma) pairs, l = 0..l_max inclusive.  One warp per element, lanes stride l, terms whose
 * exponential underflows to exactly 0 are skipped (bit-identical sum), warp-shuffle reduction.
 * Any of f/df/dlog may be NULL. */
int se3_igso3_series_f32(const float* omega, const float* sigma, float* f, float* df, float* dlog,
                         int64_t n, int l_max, float tol, se3_stream_t stream);
int se3_igso3_series_f64(const double* omega, const double* sigma, double* f, double* df, double* dlog,
                         int64_t n, int l_max, double tol, se3_stream_t stream);
/* ScoreSO3.forward (so3_sde.py:1698-1715): score = q/(|q|+tol) * dlog f(|q|, sigma). */
int se3_igso3_score(const float* rotvec, const float* sigma, float* score, int64_t n, int l_max, float tol,
                    se3_stream_t stream);
/* igso3_marginal_pdf (so3_sde.py:1795-1854), l = 0..l_count-1 (se3diff/train.py:90 passes
 * arange(l_max), i.e. l_count = l_max). */
int se3_igso3_marginal_pdf(const float* omega, const float* omega0, const float* sigma, float* pdf, int64_t n,
                           int l_count, float tol, se3_stream_t stream);
/* BaseSampleSO3._generate_lookup (so3_sde.py:1131-1187): fp64 series + cumulative trapezoid,
 * normalised.  sigma_grid [num_sigma] fp32; omega_pts [n_pts] fp64 is the full angle grid
 * pi*linspace(0,1,num_omega+1)^k, built by the caller exactly as so3_sde.py:1165-1170 does (the
 * fp32 linspace is part of the reference's numerics); cdf [num_sigma, n_pts-1] fp32 out.
 * uniform != 0 builds the one-row USO3 table (so3_sde.py:1455-1472; cdf is [1, n_pts-1]). */
int se3_igso3_build_cdf(const float* sigma_grid, int num_sigma, const double* omega_pts, int n_pts, int l_max,
                        double tol, int uniform, float* cdf, se3_stream_t stream);
/* ScoreSO3._compute_score_scaling (so3_sde.py:1637-1696) on the caller's grid
 * pi*linspace(0,1,num_omega)^k; score_scaling [num_sigma] fp32 out. */
int se3_igso3_build_score_scaling(const float* sigma_grid, int num_sigma, const double* omega_pts, int n_pts,
                                  int l_max, double tol, float* score_scaling, se3_stream_t stream);
/* BaseSampleSO3.sample for one sample per element (so3_sde.py:1189-1286, 1374-1391) fused with the
 * optional left-multiplication of sample_marginal (so3_sde.py:283): out = x . Exp(axis*omega).
 *   sigma   : [n] per-element std dev; NULL => uniform SO(3) (row 0, no small-sigma zeroing)
 *   normals : [n,3] axis normals, u : [n] uniforms in [0,1) -- drawn by the caller (parity mode);
 *             both NULL => in-kernel Philox4x32-7 keyed by (seed, element index)
 *   x       : optional [n,3,3]; NULL => out = Exp(axis*omega)
 *   angle_out optional [n]
 *   cdf_index: optional guide records built by se3_igso3_build_cdf_index (same row order as cdf); NULL => binary
 *             search.  Either way the angle index equals the reference's `sum(cdf < u)` exactly. */
int se3_igso3_sample(const float* sigma, const float* sigma_grid, int num_sigma, const float* cdf,
                     const float* omega_grid, int num_omega, const float* normals, const float* u,
                     uint64_t seed, const float* x, float* out, float* angle_out, int64_t n, float tol,
                     const float* cdf_index, se3_stream_t stream);
/* Guide records over the CDF rows: [0,1) is cut into G = 2^k >= num_omega/2 bins and each (row, bin) owns one aligned
 * 32-byte record {lo | hi << 16, cdf[lo-1 .. lo+5]} with lo/hi = lower_bound(row, bin edges): a lookup reads ONE L2
 * sector instead of 11 scattered probes.  Bins that hold more than five grid points (flat stretches of the CDF) carry the
 * lower bounds of their eight sub-bin edges instead of CDF values (ABI version 4: records built by an older library must be
 * rebuilt; they are derived data and are never stored in the npz caches).
 * index: se3_igso3_cdf_index_floats(num_rows, num_omega) floats, 32-byte aligned; num_omega <= 65535. */
int64_t se3_igso3_cdf_index_floats(int num_rows, int num_omega);
int se3_igso3_build_cdf_index(const float* cdf, int num_rows, int num_omega, float* index, se3_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4 -- structure-module (invariant point) attention         structure_module.py:109-220
 * ------------------------------------------------------------------------------------------ */
typedef struct se3_ipa_shape {
    int32_t batch;       /* B samples */
    int32_t len;         /* L residues (dense, padded) */
    int32_t heads;       /* H */
    int32_t dk;          /* scalar channels per head (16) */
    int32_t pq;          /* query/key points per head (4) */
    int32_t pv;          /* value points per head (8) */
    int32_t proj_stride; /* row stride (elements) of the fused projection matrix */
    int32_t off_q, off_k, off_v, off_qp, off_kp, off_vp; /* column offsets of head 0's blocks inside a projection row */
    int32_t hs_scalar, hs_point, hs_vpoint; /* column stride between consecutive heads of the q/k/v blocks, of the
                            q/k point blocks and of the v point block.  Block-major rows (reference weight order):
                            dk, 3*pq, 3*pv.  Head-major rows [h][q|k|v|qp|kp|vp] (what the tensor-core kernel wants,
                            one contiguous 4*(3dk+6pq+3pv)-byte record per head): 3*dk+6*pq+3*pv for all three. */
    int32_t pair_batch;  /* 1: pair tensors shared by all samples (B copies of one sequence,
                            sample.py:223); B: one per sample */
} se3_ipa_shape;

/* SAAttention.forward between the input projections and fc_out (structure_module.py:131-216).
 *   proj      [B*L, proj_stride] fp32, LOCAL frame (straight out of the fused Linear): element c of head h's q block at
 *             off_q + h*hs_scalar + c (k, v alike); point p of head h at off_qp + h*hs_point + 3p + xyz (kp alike),
 *             value points at off_vp + h*hs_vpoint + 3p + xyz
 *   rot [B*L,9], trans [B*L,3]: frames (rotation, not inverse)           structure_module.py:125-166
 *   pair_bias [pair_batch,H,L,L] fp32 = pair_weight * Linear(x2d)        structure_module.py:179
 *   pair_value[pair_batch,L,L,H*dk] fp32 = pair_value(x2d)               structure_module.py:209
 *   key_bias  [B,L] fp32 additive mask (0 / -inf) or NULL                models.py:285-293
 *   head_weight [H] fp32 = -0.5*point_weight*softplus(gamma_h)           structure_module.py:171-176
 *   scalar_weight = 1/sqrt(3*dk)
 *   out [B*L, H*(2*dk + 4*pv)] fp32 = [scalar | point_local | pair | point_norm] (structure_module.py:216)
 *   flags : SE3_IPA_EXACT (fp32 SIMT, IEEE sqrt/exp -- the parity path) or SE3_IPA_FAST_MATH
 *           (same kernel with sqrt.approx / ex2.approx).
 */
#define SE3_IPA_EXACT 0
#define SE3_IPA_FAST_MATH 1
int se3_ipa_attention_fwd(const float* proj, const float* rot, const float* trans, const float* pair_bias,
                          const float* pair_value, const float* key_bias, const float* head_weight,
                          float scalar_weight, float* out, const se3_ipa_shape* h_shape, int flags,
                          se3_stream_t stream);

/* Gradient of se3_ipa_attention_fwd (SE3_IPA_EXACT) for the fine-tune loss side, where the control model is re-evaluated
 * with autograd on stored rollout states (finetune.py:338-393 `_chunk_update`; torch autograd of structure_module.py:131-216
 * in the reference).  Frames (rot, trans) are constants there and receive no gradient.
 *   out, d_out : the forward result and its incoming gradient, [B*L, H*(2*dk + 4*pv)] fp32
 *   d_proj     : [B*L, proj_stride] fp32, every column of the q | k | v | q_pt | k_pt | v_pt blocks is written
 *   p_ws, ds_ws: [B, H, L, L] fp32 each -- the attention probabilities and the logit gradients dS.  They ARE outputs:
 *                d pair_bias = ds_ws (summed over B when pair_batch = 1), d pair_value[b, i, j, h*dk + c] =
 *                p_ws[b, h, i, j] * d_out[b*L + i, pair block h, c] (summed over B when pair_batch = 1: a GEMM over the
 *                samples, left to the caller's BLAS)
 *   d_hw_rows  : [B*L, H] fp32, per-row partials of d head_weight (sum over rows = the gradient; deterministic)
 * len <= 128 (and (2*len*(2*dk+36) + 2*len*(len|1) + len) * 4 bytes <= 227 KB): one CTA per (sample, head) keeps the L keys,
 * the L query records and both L x L matrices in shared memory.  Longer sequences up to SE3_IPA_BWD_MAX_LEN run as two
 * kernels (query rows in tiles of 64 with the keys staged in chunks; then key columns in tiles of 128) that meet in
 * p_ws / ds_ws. */
#define SE3_IPA_BWD_MAX_LEN 512
int se3_ipa_attention_bwd(const float* proj, const float* rot, const float* trans, const float* pair_bias,
                          const float* pair_value, const float* key_bias, const float* head_weight,
                          float scalar_weight, const float* out, const float* d_out, float* d_proj, float* p_ws,
                          float* ds_ws, float* d_hw_rows, const se3_ipa_shape* h_shape, se3_stream_t stream);

/* Tensor-core edition of the same operator (tcgen05.mma + TMEM, bf16 operands, fp32 accumulation): two passes, see
 * se3diff_b200/csrc/ipa_tc.cu.  Requirements: dk = 16, 4 / 8 points, shared pair tensors (pair_batch = 1), L <= 512 (257..512: the
 * keys of a query tile are split over a 2-CTA cluster).
 * The projection (structure_module.py:131-135) is delivered as two head-major record sets -- column slices of ONE bf16 GEMM
 * over row-permuted slices of the fused projection weight (se3_ipa_split_perm), or two separate matrices:
 *   scalars_bf16 : bf16 [B*L][scalar_stride], head h owns elements [h*48, h*48+48) = q 16 | k 16 | v 16; the q block
 *                  must already carry the factor scalar_weight * log2(e) (folded into the weight rows by the caller).
 *                  Each 16-byte piece is one chunk of a UMMA operand and is copied verbatim (cp.async), no conversion.
 *   points       : fp32 or bf16 (points_are_bf16) [B*L][point_stride], head h owns [h*48, h*48+48) = q_pts 12 | k_pts 12 |
 *                  v_pts 24 (local frame).  bf16 records let ONE projection GEMM write scalar and point records side by side
 *                  (row pitch = scalar_stride = point_stride); their 2^-9 rounding is of the size of the bf16 GEMM-operand
 *                  rounding the local points already carry
 * Other differences from se3_ipa_attention_fwd:
 *   pair_bias_packed  : bf16 pair_weight*pair_bias(x2d), zero padded, as se3_ipa_tc_pack_pair / se3_pair_project write it
 *                       (se3_ipa_tc_packed_pair_bytes gives the size): L <= 128: query-major [H][L (query i)][pitch (key j)], pitch =
 *                       a whole number of 8-key chunks, odd where 8 * chunks <= 128 allows (one conflict-free 16-byte shared-memory
 *                       read per query row and 8-key logit step); L > 128: TRANSPOSED [H][L (key j)][round_up(L,8) (query i)].
 *                       The (head, query-tile) slab is fetched by TMA into shared memory
 *   pair_value_packed : bf16 [L][H][Lp/8][16][8] with Lp = round_up(L,16): pair_value[i, j, h*16+c] stored at
 *                       [i][h][j/8][c][j%8], zero for j >= L (the UMMA K-major operand layout, built once per
 *                       sequence by the caller)
 *   out               : fp32 or bf16 (out_is_bf16) concat layout
 *   p_workspace / inv_workspace : scratch of the sizes reported by se3_ipa_tc_workspace_bytes (un-normalised probabilities,
 *                       bf16 row-major [H][L][round_up(B,128)][Lp], and 1/rowsum fp32 [H][L][round_up(B,128)] followed by
 *                       64 bytes of work-queue state for the persistent pass-1 kernel: ABI version 3).  Those 64 bytes must be
 *                       ZERO before the first call with a workspace (one cudaMemset after allocating it); every call leaves
 *                       them zero.  Calls that may run concurrently need their own workspaces.
 * Of h_shape only batch, len, heads, dk, pq, pv, pair_batch are read. */
int64_t se3_ipa_tc_workspace_bytes(const se3_ipa_shape* h_shape, int64_t* p_bytes, int64_t* inv_bytes);
int se3_ipa_attention_tc_fwd(const void* scalars_bf16, int64_t scalar_stride, const void* points, int points_are_bf16,
                             int64_t point_stride, const float* rot, const float* trans, const void* pair_bias_packed,
                             const void* pair_value_packed, const float* key_bias, const float* head_weight, void* out,
                             int out_is_bf16, void* p_workspace, float* inv_workspace, const se3_ipa_shape* h_shape,
                             se3_stream_t stream);

/* The two operand packs above, from the tensors the per-sequence pair precompute produces (models.py:243-293 feeding
 * structure_module.py:179, 209; SURVEY.md 8b `pair_precompute`), so that a caller holding only device pointers can feed
 * se3_ipa_attention_tc_fwd:
 *   pair_bias  [len(i)][len(j)][heads] fp32 = pair_weight * Linear(x2d)   -> bias_packed  (se3_ipa_tc_packed_pair_bytes: bias_bytes)
 *   pair_value [len(i)][len(j)][heads*16] fp32 = Linear(x2d)              -> value_packed (value_bytes), 16-byte aligned
 * Either pair (input, output) may be NULL to run only the other pack.  Pure byte movement plus one round-to-nearest-even. */
int64_t se3_ipa_tc_packed_pair_bytes(int len, int heads, int64_t* bias_bytes, int64_t* value_bytes);
int se3_ipa_tc_pack_pair(const float* pair_bias, const float* pair_value, void* bias_packed, void* value_packed, int len,
                         int heads, se3_stream_t stream);
/* The per-sequence pair precompute itself (SURVEY.md 8b `pair_precompute`), fp32, once per sequence:
 *   se3_pair_embed:   x2d[r, :] = LayerNorm(pair_embeds[r, :]; ln_gamma, ln_beta, ln_eps) . w_x2d^T + relpos_table[bucket[r % len^2], :]
 *                     (models.py:243-293: `x2d_proj` on the dense pair embedding [pair_batch*len*len, dim_embed] plus the T5-style
 *                     relative-position bias; `bucket` [len*len] int32 is the integer-exact bucket table the host evaluates,
 *                     relpos_table [num_buckets, dim_pair], w_x2d [dim_pair, dim_embed]); x2d [pair_batch*len*len, dim_pair];
 *                     stats_workspace: pair_batch*len*len*2 floats
 *   se3_pair_project: one layer's pair tensors from x2d: bias = pair_weight * x2d . W_bias^T (structure_module.py:179) and value =
 *                     x2d . W_value^T (:209); w_bias_value [heads + heads*dk, dim_pair] = the two weights stacked (bias rows first).
 *                     packed = 0: fp32 bias_out [pair_batch, heads, len, len], value_out [pair_batch, len, len, heads*dk]  (se3_ipa_attention_fwd)
 *                     packed = 1: the bf16 operands of se3_ipa_attention_tc_fwd (sizes: se3_ipa_tc_packed_pair_bytes; pair_batch = 1, dk = 16)
 * Plain fp32 tile GEMMs (k ascending): results agree with a BLAS evaluation to fp32 summation-order differences. */
int se3_pair_embed(const float* pair_embeds, const float* ln_gamma, const float* ln_beta, float ln_eps, const float* w_x2d,
                   const float* relpos_table, const int32_t* bucket, float* x2d, float* stats_workspace, int64_t pair_batch,
                   int len, int dim_embed, int dim_pair, se3_stream_t stream);
int se3_pair_project(const float* x2d, const float* w_bias_value, float pair_weight, void* bias_out, void* value_out, int packed,
                     int64_t pair_batch, int len, int heads, int dk, int dim_pair, se3_stream_t stream);
/* HOST function (no device work): row index sets of the reference's fused projection weight
 * [scalar_query | scalar_key | scalar_value | point_query | point_key | point_value] (structure_module.py:56-107, rows in the
 * reference's parameter order) for the head-major records of se3_ipa_attention_tc_fwd:
 *   h_scalar_rows [heads*3*dk] : head h -> q dk | k dk | v dk;   h_point_rows [heads*48] : head h -> q_pts 12 | k_pts 12 | v_pts 24;
 *   h_q_positions [heads*dk] (optional): positions of the q rows inside h_scalar_rows -- the rows that carry scalar_weight * log2(e). */
int se3_ipa_split_perm(int heads, int dk, int32_t* h_scalar_rows, int32_t* h_point_rows, int32_t* h_q_positions);

/* Fused residual update + pre-LayerNorm of the next block (bf16 throughput mode of the score network):
 *   x[rows,dim] += y[rows,dim] + bias[dim]      (y fp32 or bf16, bias optional: NULL skips the update; structure_module.py:247-248)
 *   out[rows,dim] = LayerNorm(x; gamma, beta, eps)   written as fp32 or bf16
 * dim in {128, 256, 512, 1024}; all pointers 16-byte aligned. */
int se3_residual_layernorm(float* x, const void* y, int y_is_bf16, const float* bias, const float* gamma, const float* beta, float eps,
                           void* out, int out_is_bf16, int64_t rows, int dim, se3_stream_t stream);
/* Tail of a diffusion head, structure_module.py:12-22 (`... Linear(D, D) -> ReLU -> Linear(D, 3)`):
 * out[r, k] = sum_c relu(y[r, c] + b1[c]) * w3[k, c] + b3[k], k < 3; y [rows, dim] fp32 = the first Linear without its bias,
 * w3 [3, dim] row-major; dim in {128, 256, 512, 1024}; out [rows, 3] fp32.  rot: optional [rows, 3, 3] frames -- the result is
 * rotated into the global frame, out[r] = R_r . out[r] (models.py:305: the translation score); NULL for none. */
int se3_bias_relu_project3(const float* y, const float* b1, const float* w3, const float* b3, const float* rot, float* out,
                           int64_t rows, int dim, se3_stream_t stream);
/* erf GELU of FeedForward (structure_module.py:25-40) on bf16 data, fp32 math, erfc by Abramowitz & Stegun 7.1.26
 * (absolute error 1.5e-7, i.e. below the bf16 output rounding); n % 8 == 0; in == out allowed. */
int se3_gelu_bf16(const void* in, void* out, int64_t n, se3_stream_t stream);

/* Folded-state indicator of the fine-tune objective (observables/folding_stability.py:52-81, called at finetune.py:452 on
 * the last batch of the rollout): dRMSD of every sample's C-alpha distance matrix to the reference's,
 * p = clamp(sigmoid(k (dRMSD - d_0)), tol, 1 - tol).  coords [batch, len, 3], ref_coords [len, 3] (nm); drmsd optional. */
int se3_folded_proportion(const float* coords, const float* ref_coords, float* p_folded, float* drmsd, int64_t batch, int len,
                          float k, float d_0, float tol, se3_stream_t stream);

/* Backbone atoms from frames (get_atom37_from_frames / compute_backbone / _adjust_oxygen_pos, convert_chemgraph.py:139-293):
 * pos [batch, len, 3] in Angstrom, rot [batch, len, 3, 3], aatype [len] (openfold order ARNDCQEGHILKMFPSTWYV), pos_is_known
 * [len] bytes or NULL -> atoms [batch, len, 5, 3] = N, CA, C, CB, O (the first five slots of the atom37 layout; glycine CB
 * is exactly zero, i.e. masked out by the reference's `any(pos != 0)` rule). */
int se3_backbone_atoms(const float* pos, const float* rot, const int32_t* aatype, const uint8_t* pos_is_known, float* atoms,
                       int64_t batch, int len, se3_stream_t stream);
/* The three statistics of the physicality filter (_filter_unphysical_traj_masks, convert_chemgraph.py:296-345) per sample:
 * out [batch, 3] = max sequential CA-CA distance, max sequential C-N distance, min heavy-atom distance between residues at
 * least three apart, in the unit of `atoms` ([batch, len, 5, 3] as written by se3_backbone_atoms). */
int se3_physicality(const float* atoms, const int32_t* aatype, float* out, int64_t batch, int len, se3_stream_t stream);

/* tcgen05 self-test: d[128,n] fp32 = a[128,k] . b[n,k]^T with bf16 operands, through the operand staging,
 * UMMA descriptors, TMEM allocation and mbarrier completion the attention kernels use. */
int se3_debug_umma_gemm(const void* a_bf16, const void* b_bf16, float* d, int n, int k, se3_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SE3DIFF_B200_H */
