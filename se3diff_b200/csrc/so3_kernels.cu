// K2 -- SO(3) exp / log / composition kernels (so3_sde.py:406-911), HBM-bound elementwise work.
// One thread per rotation, one CTA per 256 rotations; all global traffic goes through shared
// memory as 128-bit coalesced accesses (common.cuh: tile_load / tile_store).
// Compiled with -fmad=false so the arithmetic follows the reference expression order.
#include "common.cuh"

using namespace se3;

namespace {

inline dim3 grid_for(int64_t n) { return dim3((unsigned)((n + kTile - 1) / kTile)); }

// k_exp / k_log / k_compose: every warp owns a private 32 x 9 slice of the buffer (common.cuh: warp_tile_load / _store with
// SLOT = 9), so only __syncwarp() separates its load, compute and store phases.
template <typename T>
__global__ void __launch_bounds__(kTile) k_exp(const T* __restrict__ v, T* __restrict__ out, int64_t n, T tol) {
    __shared__ __align__(16) T s[kTile * 9];
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    const int t = threadIdx.x, lane = t & 31;
    T* sw = s + (t >> 5) * 32 * 9;
    warp_tile_load<3, 9>(v, s, first, count);
    __syncwarp();
    T a[3], r[9];
    if (t < count) {
        a[0] = sw[lane * 3]; a[1] = sw[lane * 3 + 1]; a[2] = sw[lane * 3 + 2];
        so3_exp(a, tol, r);
    }
    __syncwarp();
    if (t < count) {
#pragma unroll
        for (int k = 0; k < 9; ++k) sw[lane * 9 + k] = r[k];
    }
    __syncwarp();
    warp_tile_store<9, 9>(out, s, first, count);
}

template <typename T>
__global__ void __launch_bounds__(kTile) k_log(const T* __restrict__ rm, T* __restrict__ out, int64_t n) {
    __shared__ __align__(16) T s[kTile * 9];
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    const int t = threadIdx.x, lane = t & 31;
    T* sw = s + (t >> 5) * 32 * 9;
    warp_tile_load<9, 9>(rm, s, first, count);
    __syncwarp();
    T r[9], v[3];
    if (t < count) {
#pragma unroll
        for (int k = 0; k < 9; ++k) r[k] = sw[lane * 9 + k];
        so3_log(r, v);
    }
    __syncwarp();
    if (t < count) { sw[lane * 3] = v[0]; sw[lane * 3 + 1] = v[1]; sw[lane * 3 + 2] = v[2]; }
    __syncwarp();
    warp_tile_store<3, 9>(out, s, first, count);
}

__global__ void __launch_bounds__(kTile) k_angle(const float* __restrict__ rm, float* __restrict__ ang,
                                                 float* __restrict__ sn, float* __restrict__ cs, int64_t n) {
    __shared__ __align__(16) float s[kTile * 9];
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    tile_load<9>(rm, s, first, count);
    __syncthreads();
    const int t = threadIdx.x;
    if (t < count) {
        float r[9], w[3], si, co;
#pragma unroll
        for (int k = 0; k < 9; ++k) r[k] = s[t * 9 + k];
        const float th = so3_angle(r, w, &si, &co);
        if (ang) ang[first + t] = th;
        if (sn) sn[first + t] = si;
        if (cs) cs[first + t] = co;
    }
}

// MODE 0: out = R.Exp(v)   MODE 1: out = op(A).B   MODE 2: out(vec) = Log(A^T.B)
// MODE 3: out = A.Exp(t*Log(A^T.B))
template <int MODE>
__global__ void __launch_bounds__(kTile) k_compose(const float* __restrict__ a, const float* __restrict__ b,
                                                   float* __restrict__ out, int64_t n, float tol, float tpar,
                                                   int transpose_a) {
    __shared__ __align__(16) float sa[kTile * 9];
    __shared__ __align__(16) float sb[kTile * 9];
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    const int t = threadIdx.x, lane = t & 31;
    float* swa = sa + (t >> 5) * 32 * 9;
    float* swb = sb + (t >> 5) * 32 * 9;
    warp_tile_load<9, 9>(a, sa, first, count);
    if (MODE == 0) warp_tile_load<3, 9>(b, sb, first, count); else warp_tile_load<9, 9>(b, sb, first, count);
    __syncwarp();
    float ra[9], rb[9], rc[9], v[3];
    if (t < count) {
#pragma unroll
        for (int k = 0; k < 9; ++k) ra[k] = swa[lane * 9 + k];
        if (MODE == 0) {
            v[0] = swb[lane * 3]; v[1] = swb[lane * 3 + 1]; v[2] = swb[lane * 3 + 2];
            so3_exp(v, tol, rb);
            so3_mul<float, false>(ra, rb, rc);
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) rb[k] = swb[lane * 9 + k];
            if (MODE == 1) {
                if (transpose_a) so3_mul<float, true>(ra, rb, rc); else so3_mul<float, false>(ra, rb, rc);
            } else {
                so3_mul<float, true>(ra, rb, rc);
                so3_log(rc, v);
                if (MODE == 3) {
                    v[0] = tpar * v[0]; v[1] = tpar * v[1]; v[2] = tpar * v[2];
                    so3_exp(v, tol, rb);
                    so3_mul<float, false>(ra, rb, rc);
                }
            }
        }
    }
    __syncwarp();
    if (MODE == 2) {
        if (t < count) { swa[lane * 3] = v[0]; swa[lane * 3 + 1] = v[1]; swa[lane * 3 + 2] = v[2]; }
        __syncwarp();
        warp_tile_store<3, 9>(out, sa, first, count);
    } else {
        if (t < count) {
#pragma unroll
            for (int k = 0; k < 9; ++k) swa[lane * 9 + k] = rc[k];
        }
        __syncwarp();
        warp_tile_store<9, 9>(out, sa, first, count);
    }
}

// rotquat_to_rotvec / rotquat_to_rotmat (so3_sde.py:725-779)
__global__ void __launch_bounds__(kTile) k_quat(const float* __restrict__ q, float* __restrict__ rv,
                                                float* __restrict__ rm, int64_t n, float tol) {
    __shared__ __align__(16) float s[kTile * 9];
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    tile_load<4>(q, s, first, count);
    __syncthreads();
    const int t = threadIdx.x;
    float v[3], r[9];
    if (t < count) {
        const float qr = s[t * 4], qi = s[t * 4 + 1], qj = s[t * 4 + 2], qk = s[t * 4 + 3];
        const float nrm = sqrtf(qi * qi + qj * qj + qk * qk);
        const float ang = 2.0f * atan2f(nrm, qr);
        const float d = nrm + tol;
        v[0] = (qi / d) * ang; v[1] = (qj / d) * ang; v[2] = (qk / d) * ang;
        if (rm) {
            // the reference evaluates Rodrigues with `ang` as the angle and hat(axis*ang) as K
            const float th2 = ang * ang;
            float a, b;
            if (fabsf(ang) < 1e-7f) { a = 1.0f - th2 / 6.0f; b = 0.5f - th2 / 24.0f; }
            else { float sn, cs; sincosf(ang, &sn, &cs); a = sn / ang; b = (1.0f - cs) / th2; }
            const float x = v[0], y = v[1], z = v[2];
            r[0] = 1.0f + b * ((-z) * z + y * (-y)); r[1] = a * (-z) + b * (y * x); r[2] = a * y + b * ((-z) * (-x));
            r[3] = a * z + b * ((-x) * (-y)); r[4] = 1.0f + b * (z * (-z) + (-x) * x); r[5] = a * (-x) + b * (z * y);
            r[6] = a * (-y) + b * (x * z); r[7] = a * x + b * ((-y) * (-z)); r[8] = 1.0f + b * ((-y) * y + x * (-x));
        }
    }
    __syncthreads();
    if (rv) {
        if (t < count) { s[t * 3] = v[0]; s[t * 3 + 1] = v[1]; s[t * 3 + 2] = v[2]; }
        __syncthreads();
        tile_store<3>(rv, s, first, count);
        __syncthreads();
    }
    if (rm) {
        if (t < count) {
#pragma unroll
            for (int k = 0; k < 9; ++k) s[t * 9 + k] = r[k];
        }
        __syncthreads();
        tile_store<9>(rm, s, first, count);
    }
}

}  // namespace

#define SE3_LAUNCH_CHECK(name) \
    count_launch();            \
    return check_launch(name)

extern "C" {

int se3_so3_exp(const float* rotvec, float* rotmat, int64_t n, float tol, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && (n == 0 || (rotvec && rotmat)), "null pointer or negative n");
    if (n == 0) return SE3_OK;
    k_exp<float><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rotvec, rotmat, n, tol);
    SE3_LAUNCH_CHECK("se3_so3_exp");
}

int se3_so3_exp_f64(const double* rotvec, double* rotmat, int64_t n, double tol, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && (n == 0 || (rotvec && rotmat)), "null pointer or negative n");
    if (n == 0) return SE3_OK;
    k_exp<double><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rotvec, rotmat, n, tol);
    SE3_LAUNCH_CHECK("se3_so3_exp_f64");
}

int se3_so3_log(const float* rotmat, float* rotvec, int64_t n, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && (n == 0 || (rotvec && rotmat)), "null pointer or negative n");
    if (n == 0) return SE3_OK;
    k_log<float><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rotmat, rotvec, n);
    SE3_LAUNCH_CHECK("se3_so3_log");
}

int se3_so3_log_f64(const double* rotmat, double* rotvec, int64_t n, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && (n == 0 || (rotvec && rotmat)), "null pointer or negative n");
    if (n == 0) return SE3_OK;
    k_log<double><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rotmat, rotvec, n);
    SE3_LAUNCH_CHECK("se3_so3_log_f64");
}

int se3_so3_angle(const float* rotmat, float* angle, float* sin_out, float* cos_out, int64_t n,
                  se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && (n == 0 || rotmat), "null pointer or negative n");
    if (n == 0) return SE3_OK;
    k_angle<<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rotmat, angle, sin_out, cos_out, n);
    SE3_LAUNCH_CHECK("se3_so3_angle");
}

int se3_so3_compose_rotvec(const float* rotmat, const float* rotvec, float* out, int64_t n, float tol,
                           se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && (n == 0 || (rotvec && rotmat && out)), "null pointer or negative n");
    if (n == 0) return SE3_OK;
    k_compose<0><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(rotmat, rotvec, out, n, tol, 0.f, 0);
    SE3_LAUNCH_CHECK("se3_so3_compose_rotvec");
}

int se3_so3_matmul(const float* a, const float* b, float* out, int64_t n, int transpose_a, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && (n == 0 || (a && b && out)), "null pointer or negative n");
    if (n == 0) return SE3_OK;
    k_compose<1><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(a, b, out, n, 0.f, 0.f, transpose_a);
    SE3_LAUNCH_CHECK("se3_so3_matmul");
}

int se3_so3_rel_log(const float* base, const float* target, float* rotvec, int64_t n, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && (n == 0 || (base && target && rotvec)), "null pointer or negative n");
    if (n == 0) return SE3_OK;
    k_compose<2><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(base, target, rotvec, n, 0.f, 0.f, 1);
    SE3_LAUNCH_CHECK("se3_so3_rel_log");
}

int se3_so3_geodesic(const float* base, const float* target, float t, float* out, int64_t n, float tol,
                     se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && (n == 0 || (base && target && out)), "null pointer or negative n");
    if (n == 0) return SE3_OK;
    k_compose<3><<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(base, target, out, n, tol, t, 1);
    SE3_LAUNCH_CHECK("se3_so3_geodesic");
}

int se3_so3_from_quat(const float* quat, float* rotvec, float* rotmat, int64_t n, float tol, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && (n == 0 || (quat && (rotvec || rotmat))), "null pointer or negative n");
    if (n == 0) return SE3_OK;
    k_quat<<<grid_for(n), kTile, 0, (cudaStream_t)stream>>>(quat, rotvec, rotmat, n, tol);
    SE3_LAUNCH_CHECK("se3_so3_from_quat");
}

}  // extern "C"
