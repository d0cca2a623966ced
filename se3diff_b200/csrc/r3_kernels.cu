// Translation (R^3) update kernels -- the position half of every sampler step on bare [n, 3] positions, for callers
// that step the two fields separately (EulerMaruyamaPredictor on a CosineVPSDE corruption, denoiser.py:72-97; the
// position lines of dpm_solver :699-701, 733-735 and of heun_denoiser :413-459).  The fused frame kernels
// (frame_kernels.cu) evaluate exactly these expressions next to the rotation update; here they stand alone.
//
// Purely componentwise: the [n, 3] arrays are walked as flat float streams with 128-bit loads and stores (every thread
// owns 4 consecutive floats, several vectors in flight per thread), scalar tail.  HBM-bound: EM 36 B in (+12 control)
// and 12 B out (+12 dW) per residue, DPM 24 in + 12 out, Heun 24..48 in + 12 out.  Compiled with -fmad=false: each
// expression rounds like the reference's torch expression (and like frame_kernels.cu: results are bit-identical).
#include "common.cuh"

using namespace se3;

namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;   // float4 vectors in flight per thread

// R3 reverse drift (denoiser.py:64-68 with sde_lib.py:140-150): -0.5*beta*x - sqrt(beta)^2*score*w [+ sqrt(beta)*u*w]
__device__ __forceinline__ float pos_drift(float beta, float sqb, float x, float score, float w, bool has_u, float u) {
    float d = ((-0.5f * beta) * x) - ((sqb * sqb) * score) * w;
    if (has_u) d = d + (sqb * u) * w;
    return d;
}

struct EmOp {           // update_given_score, CosineVPSDE branch (denoiser.py:54-97)
    se3_em_scalars c;
    bool has_u;
    __device__ __forceinline__ void operator()(float x, float m, float u, float z, float& out, float& dw) const {
        const float nsd = c.noise_weight * c.sqrt_abs_dt;
        const float d = pos_drift(c.pos_beta, c.pos_sqrt_beta, x, m / c.pos_std, c.score_weight, has_u, u);
        dw = nsd * z;
        out = (x + d * c.dt) + c.pos_sqrt_beta * dw;
    }
};

// out[e] = f(a[e], b[e], c[e], d[e]) (unused inputs null), optional second output
template <int NIN, bool OUT2, class F>
__global__ void __launch_bounds__(kThreads)
k_stream(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, const float* __restrict__ d,
         float* __restrict__ out, float* __restrict__ out2, int64_t e, bool vec_ok, const F f) {
    const int64_t nvec = vec_ok ? e / 4 : 0;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t v0 = (int64_t)blockIdx.x * kThreads + threadIdx.x; v0 < nvec; v0 += stride * kVecPerThread) {
        float4 va[kVecPerThread], vb[kVecPerThread], vc[kVecPerThread], vd[kVecPerThread];
#pragma unroll
        for (int k = 0; k < kVecPerThread; ++k) {
            const int64_t v = v0 + k * stride;
            if (v < nvec) {
                va[k] = __ldg(reinterpret_cast<const float4*>(a) + v);
                if (NIN > 1) vb[k] = __ldg(reinterpret_cast<const float4*>(b) + v);
                if (NIN > 2) vc[k] = __ldg(reinterpret_cast<const float4*>(c) + v);
                if (NIN > 3) vd[k] = __ldg(reinterpret_cast<const float4*>(d) + v);
            }
        }
#pragma unroll
        for (int k = 0; k < kVecPerThread; ++k) {
            const int64_t v = v0 + k * stride;
            if (v < nvec) {
                float4 o, o2;
                const float4 zb = NIN > 1 ? vb[k] : make_float4(0, 0, 0, 0), zc = NIN > 2 ? vc[k] : make_float4(0, 0, 0, 0),
                             zd = NIN > 3 ? vd[k] : make_float4(0, 0, 0, 0);
                f(va[k].x, zb.x, zc.x, zd.x, o.x, o2.x);
                f(va[k].y, zb.y, zc.y, zd.y, o.y, o2.y);
                f(va[k].z, zb.z, zc.z, zd.z, o.z, o2.z);
                f(va[k].w, zb.w, zc.w, zd.w, o.w, o2.w);
                reinterpret_cast<float4*>(out)[v] = o;
                if (OUT2) reinterpret_cast<float4*>(out2)[v] = o2;
            }
        }
    }
    // scalar tail (or everything, for unaligned pointers)
    for (int64_t i = nvec * 4 + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < e; i += stride) {
        float o, o2;
        f(a[i], NIN > 1 ? b[i] : 0.f, NIN > 2 ? c[i] : 0.f, NIN > 3 ? d[i] : 0.f, o, o2);
        out[i] = o;
        if (OUT2) out2[i] = o2;
    }
}

struct EmArgs4 {        // (x, m, z, u) order for k_stream
    EmOp op;
    __device__ __forceinline__ void operator()(float x, float m, float z, float u, float& out, float& dw) const { op(x, m, u, z, out, dw); }
};

struct DpmOp {          // denoiser.py:699-701 / :733-735: c_x * x + c_s * (m / std)
    float c_x, c_s, std;
    __device__ __forceinline__ void operator()(float x, float m, float, float, float& out, float&) const { out = c_x * x + c_s * (m / std); }
};

struct HeunChurnOp {    // forward SDE step t -> t_hat (denoiser.py:413-418, 118-131): drift -0.5*beta*x, noise weight 1
    float dt, sq_dt, beta, sqb;
    __device__ __forceinline__ void operator()(float x, float z, float, float, float& out, float&) const {
        const float nsd = 1.0f * sq_dt, hb = -0.5f * beta;
        out = (x + (hb * x) * dt) + sqb * (nsd * z);
    }
};

struct HeunStepOp {     // denoiser.py:423-459: first-order step, or the corrected one with the drift at (t_next, pos_pred)
    se3_heun_scalars c;
    bool correct;
    __device__ __forceinline__ void operator()(float x, float m, float x1, float m1, float& out, float&) const {
        float d = pos_drift(c.hat_pos_beta, c.hat_pos_sqrt_beta, x, m / c.hat_pos_std, 0.5f, false, 0.f);
        if (correct) {
            const float n = pos_drift(c.next_pos_beta, c.next_pos_sqrt_beta, x1, m1 / c.next_pos_std, 0.5f, false, 0.f);
            d = (n + d) / 2.0f;
        }
        out = x + d * c.step_dt;
    }
};

inline bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int NIN, bool OUT2, class F>
int run(const char* what, const float* a, const float* b, const float* c, const float* d, float* out, float* out2, int64_t n, const F& f,
        cudaStream_t st) {
    const int64_t e = 3 * n;
    const bool vec_ok = aligned16(a) && aligned16(b) && aligned16(c) && aligned16(d) && aligned16(out) && aligned16(out2);
    const int64_t work = vec_ok ? (e / 4 + kVecPerThread - 1) / kVecPerThread + 4 : e;
    int64_t blocks = (work + kThreads - 1) / kThreads;
    if (blocks > 148 * 32) blocks = 148 * 32;          // grid-stride beyond 32 CTAs per SM's worth
    if (blocks < 1) blocks = 1;
    k_stream<NIN, OUT2, F><<<(unsigned)blocks, kThreads, 0, st>>>(a, b, c, d, out, out2, e, vec_ok, f);
    count_launch();
    return check_launch(what);
}

}  // namespace

extern "C" int se3_r3_update_em(const float* pos, const float* m_pos, const float* u_pos, const float* z_pos, float* pos_out,
                                float* dw_pos, int64_t n, const se3_em_scalars* h, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(pos && m_pos && z_pos && pos_out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const EmArgs4 f{EmOp{*h, u_pos != nullptr}};
    if (u_pos && dw_pos) return run<4, true>("se3_r3_update_em", pos, m_pos, z_pos, u_pos, pos_out, dw_pos, n, f, st);
    if (u_pos) return run<4, false>("se3_r3_update_em", pos, m_pos, z_pos, u_pos, pos_out, nullptr, n, f, st);
    if (dw_pos) return run<3, true>("se3_r3_update_em", pos, m_pos, z_pos, nullptr, pos_out, dw_pos, n, f, st);
    return run<3, false>("se3_r3_update_em", pos, m_pos, z_pos, nullptr, pos_out, nullptr, n, f, st);
}

extern "C" int se3_r3_update_dpm(const float* pos, const float* m_pos, float* pos_out, int64_t n, const se3_dpm_scalars* h,
                                 int final_half, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(pos && m_pos && pos_out, "null pointer");
    const DpmOp f = final_half ? DpmOp{h->pos_c_x_fin, h->pos_c_s_fin, h->pos_std_lam} : DpmOp{h->pos_c_x_mid, h->pos_c_s_mid, h->pos_std_t};
    return run<2, false>("se3_r3_update_dpm", pos, m_pos, nullptr, nullptr, pos_out, nullptr, n, f, (cudaStream_t)stream);
}

extern "C" int se3_r3_heun_churn(const float* pos, const float* z_pos, float* pos_hat, int64_t n, const se3_heun_scalars* h,
                                 se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(pos && z_pos && pos_hat, "null pointer");
    const HeunChurnOp f{h->churn_dt, h->churn_sqrt_abs_dt, h->churn_pos_beta, h->churn_pos_sqrt_beta};
    return run<2, false>("se3_r3_heun_churn", pos, z_pos, nullptr, nullptr, pos_hat, nullptr, n, f, (cudaStream_t)stream);
}

extern "C" int se3_r3_heun_step(const float* pos_hat, const float* m_pos_hat, const float* pos_pred, const float* m_pos_next,
                                float* pos_out, int64_t n, const se3_heun_scalars* h, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && h, "negative n or null scalars");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(pos_hat && m_pos_hat && pos_out, "null pointer");
    SE3_REQUIRE((pos_pred == nullptr) == (m_pos_next == nullptr), "pos_pred and m_pos_next come together (the corrected step) or not at all");
    const bool correct = pos_pred != nullptr;
    const HeunStepOp f{*h, correct};
    if (correct) return run<4, false>("se3_r3_heun_step", pos_hat, m_pos_hat, pos_pred, m_pos_next, pos_out, nullptr, n, f, (cudaStream_t)stream);
    return run<2, false>("se3_r3_heun_step", pos_hat, m_pos_hat, nullptr, nullptr, pos_out, nullptr, n, f, (cudaStream_t)stream);
}
