// Shared device helpers for the se3diff_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/se3diff_b200.h"

namespace se3 {

void set_error(const char* fmt, ...);
int check_launch(const char* what);
void count_launch(int n = 1);

// Packed pair bias of the tensor-core IPA operator (bf16, one slab per head; written by se3_ipa_tc_pack_pair / se3_pair_project,
// read by pass 1 of se3_ipa_attention_tc_fwd).
//   L <= 128 (one query tile): QUERY-major [H][L(i)][pitch(j)], keys >= L zero.  The thread of query row i reads the eight biases of
//            an 8-column logit step as ONE 16-byte LDS; the pitch is a whole number of 16-byte chunks, odd where the slab allows it
//            (eight consecutive rows then start in eight different bank groups: the quarter-warp phases of LDS.128 are conflict-free).
//   L  > 128: KEY-major [H][L(j)][round_up(L, 8)(i)], queries >= L zero: a (keys x 128 queries) box of a 2-D tensor map per query tile.
__host__ __device__ constexpr bool ipa_bias_query_major(int L) { return L <= 128; }
__host__ __device__ constexpr int ipa_bias_pitch(int L) {
    const int chunks = (L + 7) / 8;
    if (!ipa_bias_query_major(L)) return chunks * 8;
    return (chunks | 1) * 8 <= 128 ? (chunks | 1) * 8 : chunks * 8;   // the slab [L][pitch] shares the 256 * round_up(L, 16) bytes of the P operand
}

#define SE3_REQUIRE(cond, msg)                       \
    do {                                             \
        if (!(cond)) {                               \
            se3::set_error("%s: %s", __func__, msg); \
            return SE3_EINVAL;                       \
        }                                            \
    } while (0)

constexpr int kTile = 256;  // elements (rotations / residues) per CTA == threads per CTA

// ---------------------------------------------------------------------------------------------
// Coalesced tile movement.  A CTA owns `kTile` consecutive elements of W scalars each (W = 3, 9,
// 4...).  Global<->shared traffic is issued as 128-bit accesses over the contiguous byte range of
// the tile (every lane of a warp touches consecutive 16 B => full 128 B sectors), then each thread
// reads its own element from shared memory at stride W (W odd => bank-conflict free).
// Falls back to scalar accesses on the ragged last tile or for unaligned base pointers.
// ---------------------------------------------------------------------------------------------
template <int W, typename T>
__device__ __forceinline__ void tile_load(const T* __restrict__ g, T* __restrict__ s, int64_t first, int count) {
    const T* src = g + first * W;
    constexpr int kPerVec = 16 / sizeof(T);
    if (count == kTile && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (kTile * W) % kPerVec == 0) {
        const float4* v = reinterpret_cast<const float4*>(src);
        float4* d = reinterpret_cast<float4*>(s);
        constexpr int nvec = kTile * W / kPerVec;
#pragma unroll
        for (int i = threadIdx.x; i < nvec; i += kTile) d[i] = __ldg(v + i);
    } else {
        for (int i = threadIdx.x; i < count * W; i += kTile) s[i] = src[i];
    }
}

// Asynchronous edition (cp.async, no register staging): the copy is in flight while the thread does independent work;
// tile_load_wait() before the __syncthreads() that publishes the tile.
template <int W, typename T>
__device__ __forceinline__ void tile_load_async(const T* __restrict__ g, T* __restrict__ s, int64_t first, int count) {
    const T* src = g + first * W;
    constexpr int kPerVec = 16 / sizeof(T);
    if (count == kTile && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (kTile * W) % kPerVec == 0) {
        constexpr int nvec = kTile * W / kPerVec;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s);
#pragma unroll
        for (int i = threadIdx.x; i < nvec; i += kTile)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + (uint32_t)i * 16u), "l"(reinterpret_cast<const float4*>(src) + i) : "memory");
    } else {
        for (int i = threadIdx.x; i < count * W; i += kTile) s[i] = src[i];
    }
}
__device__ __forceinline__ void tile_load_wait() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int W, typename T>
__device__ __forceinline__ void tile_store(T* __restrict__ g, const T* __restrict__ s, int64_t first, int count) {
    T* dst = g + first * W;
    constexpr int kPerVec = 16 / sizeof(T);
    if (count == kTile && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (kTile * W) % kPerVec == 0) {
        float4* v = reinterpret_cast<float4*>(dst);
        const float4* d = reinterpret_cast<const float4*>(s);
        constexpr int nvec = kTile * W / kPerVec;
#pragma unroll
        for (int i = threadIdx.x; i < nvec; i += kTile) v[i] = d[i];
    } else {
        for (int i = threadIdx.x; i < count * W; i += kTile) dst[i] = s[i];
    }
}

// Per-warp edition of the same movement: warp w of the CTA moves its own 32 elements [first + 32w, first + 32w + 32)
// to / from the SAME shared-memory layout (offset 32*w*W), so a kernel whose threads only ever touch their own element of
// arrays with one fixed W needs __syncwarp() instead of __syncthreads(): the warps of a CTA run their load / compute /
// store phases independently instead of meeting at two CTA barriers per tile (ncu on k_dpm_mid: 5.4 warps stalled at
// the barrier per issued instruction, DRAM at 53 %).
// SLOT: elements of shared memory each thread owns (>= W): a kernel that reads W_in and writes W_out values per element
// through one buffer gives every warp a private 32*SLOT slice, SLOT = max(W_in, W_out).
template <int W, int SLOT = W, typename T>
__device__ __forceinline__ void warp_tile_load(const T* __restrict__ g, T* __restrict__ s, int64_t first, int count) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wcount = min(32, count - 32 * w);
    if (wcount <= 0) return;
    const T* src = g + (first + 32 * w) * W;
    T* dst = s + 32 * w * SLOT;
    constexpr int kPerVec = 16 / sizeof(T);
    if (wcount == 32 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (32 * W) % kPerVec == 0) {
        constexpr int nvec = 32 * W / kPerVec;
#pragma unroll
        for (int i = lane; i < nvec; i += 32) reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    } else {
        for (int i = lane; i < wcount * W; i += 32) dst[i] = src[i];
    }
}

// cp.async edition of warp_tile_load (tile_load_wait(), then __syncwarp(), publishes the warp's slice)
template <int W, int SLOT = W, typename T>
__device__ __forceinline__ void warp_tile_load_async(const T* __restrict__ g, T* __restrict__ s, int64_t first, int count) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wcount = min(32, count - 32 * w);
    if (wcount <= 0) return;
    const T* src = g + (first + 32 * w) * W;
    T* dst = s + 32 * w * SLOT;
    constexpr int kPerVec = 16 / sizeof(T);
    if (wcount == 32 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (32 * W) % kPerVec == 0) {
        constexpr int nvec = 32 * W / kPerVec;
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
#pragma unroll
        for (int i = lane; i < nvec; i += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + (uint32_t)i * 16u), "l"(reinterpret_cast<const float4*>(src) + i) : "memory");
    } else {
        for (int i = lane; i < wcount * W; i += 32) dst[i] = src[i];
    }
}

template <int W, int SLOT = W, typename T>
__device__ __forceinline__ void warp_tile_store(T* __restrict__ g, const T* __restrict__ s, int64_t first, int count) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wcount = min(32, count - 32 * w);
    if (wcount <= 0) return;
    T* dst = g + (first + 32 * w) * W;
    const T* src = s + 32 * w * SLOT;
    constexpr int kPerVec = 16 / sizeof(T);
    if (wcount == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (32 * W) % kPerVec == 0) {
        constexpr int nvec = 32 * W / kPerVec;
#pragma unroll
        for (int i = lane; i < nvec; i += 32) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
    } else {
        for (int i = lane; i < wcount * W; i += 32) dst[i] = src[i];
    }
}

// ---------------------------------------------------------------------------------------------
// SO(3) maps on register-resident 3-vectors / row-major 3x3 matrices.
// Operation order follows the reference expression by expression (so3_sde.py) so that, with FMA
// contraction disabled for these translation units, the arithmetic differs from the CPU reference
// only in the libm calls (sin, cos, atan2).
// ---------------------------------------------------------------------------------------------
template <typename T> struct Math;
template <> struct Math<float> {
    static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
    static __device__ __forceinline__ void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float atan2(float y, float x) { return atan2f(y, x); }
    static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
    static __device__ __forceinline__ float pi() { return 3.14159274101257324f; }
    // torch.isclose(theta, pi, atol=1e-2): |theta - pi| <= atol + rtol*|pi| evaluated in fp32
    static __device__ __forceinline__ float pi_band() { return 1e-2f + 1e-5f * 3.14159274101257324f; }
};
template <> struct Math<double> {
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
    static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
    static __device__ __forceinline__ double abs(double x) { return fabs(x); }
    static __device__ __forceinline__ double pi() { return 3.141592653589793; }
    static __device__ __forceinline__ double pi_band() { return 1e-2 + 1e-5 * 3.141592653589793; }
};

// rotvec_to_rotmat / skew_matrix_exponential_map (so3_sde.py:478-554)
template <typename T>
__device__ __forceinline__ void so3_exp(const T v[3], T tol, T r[9]) {
    const T x = v[0], y = v[1], z = v[2];
    const T th = Math<T>::sqrt(x * x + y * y + z * z);
    const T th2 = th * th;
    T a, b;
    if (Math<T>::abs(th) < tol) {
        a = T(1) - th2 / T(6);
        b = T(0.5) - th2 / T(24);
    } else {
        T s, c;
        Math<T>::sincos(th, &s, &c);
        a = s / th;
        b = (T(1) - c) / th2;
    }
    // K = [[0,-z,y],[z,0,-x],[-y,x,0]];  K2 = K.K summed in k order as the einsum does
    const T k2_00 = (-z) * z + y * (-y), k2_01 = y * x, k2_02 = (-z) * (-x);
    const T k2_10 = (-x) * (-y), k2_11 = z * (-z) + (-x) * x, k2_12 = z * y;
    const T k2_20 = x * z, k2_21 = (-y) * (-z), k2_22 = (-y) * y + x * (-x);
    r[0] = (T(1) + a * T(0)) + b * k2_00;
    r[1] = (T(0) + a * (-z)) + b * k2_01;
    r[2] = (T(0) + a * y) + b * k2_02;
    r[3] = (T(0) + a * z) + b * k2_10;
    r[4] = (T(1) + a * T(0)) + b * k2_11;
    r[5] = (T(0) + a * (-x)) + b * k2_12;
    r[6] = (T(0) + a * (-y)) + b * k2_20;
    r[7] = (T(0) + a * x) + b * k2_21;
    r[8] = (T(1) + a * T(0)) + b * k2_22;
}

// rot_mult (so3_sde.py:875-877): c = a.b (TA: a^T.b)
template <typename T, bool TA = false>
__device__ __forceinline__ void so3_mul(const T a[9], const T b[9], T c[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const T a0 = TA ? a[0 * 3 + i] : a[i * 3 + 0];
            const T a1 = TA ? a[1 * 3 + i] : a[i * 3 + 1];
            const T a2 = TA ? a[2 * 3 + i] : a[i * 3 + 2];
            c[i * 3 + j] = (a0 * b[0 * 3 + j] + a1 * b[1 * 3 + j]) + a2 * b[2 * 3 + j];
        }
}

// ---- fused editions for the frame-update kernels ---------------------------------------------------------------------
// The rotation half of a frame update is never bit-identical to the CPU reference (its sin / cos come from another libm), so
// nothing is lost by spelling it with explicit fused multiply-adds (these translation units are compiled with -fmad=false,
// which only stops CONTRACTION of a*b+c; fmaf() is always fused) and a short sincos: two-constant Cody-Waite reduction by
// pi/2 and the degree-7 / degree-8 minimax polynomials of Cephes' sinf / cosf on [-pi/4, pi/4] (approximation error 3e-9 /
// 1e-10, i.e. 1 ulp-class results for the |theta| << 1e4 a step's rotation vector has).  ~3x fewer instructions than the
// reference-order so3_exp + so3_mul: ncu showed k_em issue-bound at 0.6 of the HBM roof, not latency-bound.
__device__ __forceinline__ void sincos_fused(float x, float* s, float* c) {
    const float k = rintf(x * 0.636619772367581343f);
    float r = fmaf(-k, 1.5707963705062866f, x);
    r = fmaf(-k, -4.371138828673793e-08f, r);
    const float r2 = r * r;
    const float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
    const float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2 * r2, fmaf(-0.5f, r2, 1.0f));
    const int q = (int)k;
    const float ss = (q & 1) ? cp : sp, cc = (q & 1) ? sp : cp;
    *s = (q & 2) ? -ss : ss;
    *c = ((q + 1) & 2) ? -cc : cc;
}
// out = R . Exp(v)  (apply_rotvec_to_rotmat, so3_sde.py:782-802; Rodrigues with the Taylor branch below `tol`, :533-554)
__device__ __forceinline__ void so3_apply_rotvec_fused(const float r[9], float x, float y, float z, float tol, float out[9]) {
    const float th2 = fmaf(x, x, fmaf(y, y, z * z));
    float th;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(th) : "f"(th2));      // 1 ulp-class; the correctly rounded form is a 7-instruction sequence
    float a, b;
    if (th < tol) {
        a = fmaf(th2, -1.0f / 6.0f, 1.0f);
        b = fmaf(th2, -1.0f / 24.0f, 0.5f);
    } else {
        float s, c;
        sincos_fused(th, &s, &c);
        a = __fdividef(s, th);
        b = __fdividef(1.0f - c, th2);
    }
    // E = I + a K + b K^2,  K = [[0,-z,y],[z,0,-x],[-y,x,0]],  K^2 = v v^T - |v|^2 I
    const float bx = b * x, by = b * y, bz = b * z;
    const float e00 = fmaf(-b, fmaf(y, y, z * z), 1.0f), e11 = fmaf(-b, fmaf(x, x, z * z), 1.0f), e22 = fmaf(-b, fmaf(x, x, y * y), 1.0f);
    const float e01 = fmaf(bx, y, -a * z), e10 = fmaf(bx, y, a * z);
    const float e02 = fmaf(bx, z, a * y), e20 = fmaf(bx, z, -a * y);
    const float e12 = fmaf(by, z, -a * x), e21 = fmaf(by, z, a * x);
    (void)bz;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float r0 = r[i * 3], r1 = r[i * 3 + 1], r2 = r[i * 3 + 2];
        out[i * 3] = fmaf(r2, e20, fmaf(r1, e10, r0 * e00));
        out[i * 3 + 1] = fmaf(r2, e21, fmaf(r1, e11, r0 * e01));
        out[i * 3 + 2] = fmaf(r2, e22, fmaf(r1, e12, r0 * e02));
    }
}

// angle_from_rotmat (so3_sde.py:651-676)
template <typename T>
__device__ __forceinline__ T so3_angle(const T r[9], T w[3], T* sin_out, T* cos_out) {
    w[0] = r[7] - r[5];
    w[1] = r[2] - r[6];
    w[2] = r[3] - r[1];
    const T s = Math<T>::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]) / T(2);
    const T c = (((r[0] + r[4]) + r[8]) - T(1)) / T(2);
    *sin_out = s;
    *cos_out = c;
    return Math<T>::atan2(s, c);
}

// rotmat_to_rotvec (so3_sde.py:557-648).  The reference blends the three regimes with 0/1 masks;
// here they are branches with the same thresholds (isclose(theta,0): |theta| <= 1e-8;
// isclose(theta,pi,atol=1e-2): |theta-pi| <= 1e-2+1e-5*pi), which is equivalent whenever the
// disabled branches are finite and avoids NaN poisoning when they are not.
template <typename T>
__device__ __forceinline__ void so3_log(const T r[9], T v[3]) {
    T w[3], s, c;
    const T th = so3_angle(r, w, &s, &c);
    const bool is_zero = Math<T>::abs(th) <= T(1e-8);
    const bool is_pi = Math<T>::abs(th - Math<T>::pi()) <= Math<T>::pi_band();
    if (is_pi) {
        // outer = (I + R)/2 with the diagonal clamped at 0; axis = sqrt(max(diag, 1e-8));
        // signs from the row of largest norm (first maximum wins, as torch.argmax)
        T o[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = ((i % 4 == 0 ? T(1) : T(0)) + r[i]) / T(2);
        o[0] = o[0] > T(0) ? o[0] : T(0);
        o[4] = o[4] > T(0) ? o[4] : T(0);
        o[8] = o[8] > T(0) ? o[8] : T(0);
        T n0 = Math<T>::sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
        T n1 = Math<T>::sqrt(o[3] * o[3] + o[4] * o[4] + o[5] * o[5]);
        T n2 = Math<T>::sqrt(o[6] * o[6] + o[7] * o[7] + o[8] * o[8]);
        int row = 0;
        T best = n0;
        if (n1 > best) { best = n1; row = 1; }
        if (n2 > best) { row = 2; }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const T d = o[k * 4];
            const T ax = Math<T>::sqrt(d > T(1e-8) ? d : T(1e-8));
            const T l = o[row * 3 + k];
            const T sg = l > T(0) ? T(1) : (l < T(0) ? T(-1) : T(0));
            // reference: vector*prefactor (prefactor = 0/1 = 0 in the pi regime) + vector_pi
            v[k] = w[k] * T(0) + (ax * th) * sg;
        }
    } else if (is_zero) {
        const T pre = T(0.5) / (T(1) - th * th / T(6));
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = w[k] * pre;
    } else {
        const T pre = th / (T(2) * s);
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = w[k] * pre;
    }
}

}  // namespace se3
