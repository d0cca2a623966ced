// K1 -- IGSO(3) on SO(3): truncated angle series (score), lookup-table construction (fp64) and
// inverse-CDF sampling (so3_sde.py:993-2042).
//
//  * series / score  : one warp per (omega, sigma) element, lanes stride the expansion order l, the
//                      partial sums are combined with warp shuffles.  Terms whose exponential factor
//                      exp(-l(l+1)sigma^2/2) is exactly 0 in the working precision are not evaluated
//                      (the sum is unchanged bit for bit); compute-bound (sin/exp), not HBM-bound.
//  * table build     : one CTA per sigma row, fp64 like the reference (so3_sde.py:1156,1659); the
//                      exponential factors of the row are staged once in shared memory.
//  * sampling        : one thread per rotation: bucketize(sigma) + lower_bound on the CDF row
//                      (== the reference's `sum(cdf < u)`, so3_sde.py:1265) + lerp + Rodrigues
//                      [+ left-multiplication by x for sample_marginal]; output written through the
//                      128-bit tile path.  HBM-bound: 72-88 B/rotation.
#include "common.cuh"

using namespace se3;

namespace {

template <typename T> struct Lim;
template <> struct Lim<float> { static constexpr float cut = 104.0f; };     // expf(-104) == 0
template <> struct Lim<double> { static constexpr double cut = 745.2; };    // exp(-745.2) == 0

template <typename T> __device__ __forceinline__ T t_sin(T x);
template <> __device__ __forceinline__ float t_sin<float>(float x) { return sinf(x); }
template <> __device__ __forceinline__ double t_sin<double>(double x) { return sin(x); }
template <typename T> __device__ __forceinline__ T t_cos(T x);
template <> __device__ __forceinline__ float t_cos<float>(float x) { return cosf(x); }
template <> __device__ __forceinline__ double t_cos<double>(double x) { return cos(x); }
template <typename T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ bool t_bad(T x) { return isnan(x) || isinf(x); }

// number of leading terms that can be non-zero: smallest l with l(l+1)*s2/2 > cut, capped at l_max
template <typename T>
__device__ __forceinline__ int term_count(T s2, int l_max) {
    if (!(s2 > T(0))) return l_max + 1;
    const double lc = ::sqrt(2.0 * (double)Lim<T>::cut / (double)s2 + 0.25) + 1.5;
    return lc >= (double)(l_max + 1) ? l_max + 1 : (int)lc;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-term factors in the reference's rounding order (so3_sde.py:1760-1767, 1884-1895):
// (2l+1) and -l(l+1) are formed in fp32 (torch promotes the int64 l_grid with a python float to the
// default dtype) and only then meet omega/sigma in the working precision T.
template <typename T>
__device__ __forceinline__ T exp_factor(int l, T s2) {
    const float lf = (float)l;
    const float f1 = 2.0f * lf + 1.0f;
    const float f2 = -lf * (lf + 1.0f);
    return (T)f1 * t_exp<T>(((T)f2 * s2) / T(2));
}

// finalisers (so3_sde.py:1775-1792, 1899-1913, 1937-1940)
template <typename T>
__device__ __forceinline__ T finish_f(T fsum, T flim, T om, T tol) {
    T f = fsum / (t_sin<T>(T(0.5) * om) + tol);
    if (om <= tol) f = flim;
    if (t_bad(f)) f = T(0);
    return f > T(0) ? f : T(0);
}
template <typename T>
__device__ __forceinline__ T finish_df(T dsum, T om, T tol) {
    T d = dsum / ((T(1) - t_cos<T>(om)) + tol);
    if (om <= tol) d = T(0);
    if (t_bad(d)) d = T(0);
    return d;
}

// MODE 0: series (f, df, dlog);  MODE 1: score q/(|q|+tol)*dlog with omega=|q| (ScoreSO3.forward)
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_series(const T* __restrict__ omega, const T* __restrict__ sigma, T* __restrict__ f_out, T* __restrict__ df_out,
         T* __restrict__ dlog_out, int64_t n, int l_max, T tol) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp0; e < n; e += nwarps) {
        T om, q[3] = {T(0), T(0), T(0)};
        if (MODE == 1) {
            q[0] = omega[e * 3]; q[1] = omega[e * 3 + 1]; q[2] = omega[e * 3 + 2];
            om = Math<T>::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
        } else {
            om = omega[e];
        }
        const T sg = sigma[e];
        const T s2 = sg * sg;
        const int nt = term_count<T>(s2, l_max);
        const bool need_df = MODE == 1 || df_out || dlog_out;
        T fs = T(0), fl = T(0), ds = T(0);
        for (int l = lane; l < nt; l += 32) {
            const float lf = (float)l;
            const T ef = exp_factor<T>(l, s2);
            fs += ef * t_sin<T>((T)(lf + 0.5f) * om);
            fl += ef * (T)(2.0f * lf + 1.0f);
            if (need_df) {
                const float l1 = lf + 1.0f;
                ds += ef * ((T)lf * t_sin<T>((T)l1 * om) - (T)l1 * t_sin<T>((T)lf * om));
            }
        }
        fs = warp_sum(fs); fl = warp_sum(fl); ds = warp_sum(ds);
        const T f = finish_f(fs, fl, om, tol);
        const T d = finish_df(ds, om, tol);
        const T dl = d / (f + tol);
        if (MODE == 1) {
            if (lane < 3) f_out[e * 3 + lane] = q[lane] / (om + tol) * dl;
        } else if (lane == 0) {
            if (f_out) f_out[e] = f;
            if (df_out) df_out[e] = d;
            if (dlog_out) dlog_out[e] = dl;
        }
    }
}

// igso3_marginal_pdf (so3_sde.py:1795-1854)
__global__ void __launch_bounds__(256)
k_marginal(const float* __restrict__ omega, const float* __restrict__ omega0, const float* __restrict__ sigma,
           float* __restrict__ out, int64_t n, int l_count, float tol) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp0; e < n; e += nwarps) {
        const float om = omega[e], o0 = omega0[e], sg = sigma[e], s2 = sg * sg;
        const int nt = min(term_count<float>(s2, l_count - 1), l_count);
        float a = 0.f, b = 0.f;
        for (int l = lane; l < nt; l += 32) {
            const float lf = (float)l;
            const float ex = expf(((-lf * (lf + 1.0f)) * s2) / 2.0f);
            const float sn = sinf((lf + 0.5f) * om);
            a += (ex * sn) * sinf((lf + 0.5f) * o0);
            b += (ex * (2.0f * lf + 1.0f)) * sn;
        }
        a = warp_sum(a); b = warp_sum(b);
        if (lane == 0) {
            const float d = sinf(0.5f * om), d0 = sinf(0.5f * o0);
            float f = a * d / (d0 + tol);
            if (o0 <= tol) f = b * d;
            if (t_bad(f)) f = 0.f;
            f = f * 2.0f / 3.14159274101257324f;
            out[e] = f > 0.f ? f : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// lookup tables, fp64
// ---------------------------------------------------------------------------------------------
constexpr int kRowThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kRowThreads / 32; ++w) t += red[w];
    return t;
}

// One CTA per sigma row.  dyn smem: ef[l_max+1] | vals[n_pts]
// MODE 0: CDF row (so3_sde.py:1172-1187)   MODE 1: score scaling (so3_sde.py:1673-1692)
template <int MODE>
__global__ void __launch_bounds__(kRowThreads)
k_table_row(const float* __restrict__ sigma_grid, const double* __restrict__ omega, int n_pts, int l_max, double tol,
            int uniform, float* __restrict__ out) {
    extern __shared__ __align__(16) double sm[];
    double* ef = sm;
    double* vals = sm + (l_max + 1);
    __shared__ double red[kRowThreads / 32];
    const int row = blockIdx.x;
    const double sg = (double)sigma_grid[row];
    const double s2 = sg * sg;
    const int nt = uniform ? 0 : term_count<double>(s2, l_max);
    for (int l = threadIdx.x; l < nt; l += kRowThreads) ef[l] = exp_factor<double>(l, s2);
    __syncthreads();
    double acc_num = 0.0, acc_den = 0.0;
    for (int k = threadIdx.x; k < n_pts; k += kRowThreads) {
        const double om = omega[k];
        double f;
        double dsum = 0.0;
        if (uniform) {
            f = 1.0;
        } else {
            double fs = 0.0, fl = 0.0;
            for (int l = 0; l < nt; ++l) {
                const float lf = (float)l;
                const double e = ef[l];
                fs += e * sin((double)(lf + 0.5f) * om);
                fl += e * (double)(2.0f * lf + 1.0f);
                if (MODE == 1) {
                    const float l1 = lf + 1.0f;
                    dsum += e * ((double)lf * sin((double)l1 * om) - (double)l1 * sin((double)lf * om));
                }
            }
            f = finish_f<double>(fs, fl, om, tol);
        }
        // so3_sde.py:1176 evaluates (f*(1-cos))/pi, so3_sde.py:1678-1679 evaluates f*((1-cos)/pi)
        const double pdf = MODE == 0 ? f * (1.0 - cos(om)) / 3.141592653589793 : f * ((1.0 - cos(om)) / 3.141592653589793);
        if (MODE == 0) {
            vals[k] = pdf;
        } else {
            const double dl = finish_df<double>(dsum, om, tol) / (f + tol);
            const double ap = fabs(pdf);
            acc_num += dl * dl * ap;
            acc_den += ap;
        }
    }
    if (MODE == 0) {
        __syncthreads();
        if (threadIdx.x == 0) {  // sequential cumulative trapezoid, the order torch.cumsum uses
            double run = 0.0;
            for (int k = 0; k + 1 < n_pts; ++k) {
                run += ((vals[k] + vals[k + 1]) * (omega[k + 1] - omega[k])) / 2.0;
                vals[k] = run;
            }
        }
        __syncthreads();
        const double last = vals[n_pts - 2];
        for (int k = threadIdx.x; k + 1 < n_pts; k += kRowThreads) out[(int64_t)row * (n_pts - 1) + k] = (float)(vals[k] / last);
    } else {
        const double num = block_sum(acc_num, red);
        const double den = block_sum(acc_den, red);
        if (threadIdx.x == 0) out[row] = (float)sqrt(num / (3.0 * den + tol));
    }
}

// ---------------------------------------------------------------------------------------------
// inverse-CDF sampling
// ---------------------------------------------------------------------------------------------
// first index with a[idx] >= v  == count of a[k] < v for a sorted a (torch.bucketize / `sum(cdf < u)`).
// Every probe is a 32-byte L2 sector of its own: the sampling kernel is bound by the number of probes (L2 sector
// bandwidth), so a plain binary search (fewest probes) beats wider k-ary searches here.
__device__ __forceinline__ int lower_bound(const float* __restrict__ a, int n, float v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// Same result for a (nearly) geometric grid such as sigma(t) = s_min (s_max/s_min)^t on linspace t: a log-linear
// guess from the end points followed by an exact local fix-up costs 2-3 probes instead of log2(n) = 10.
__device__ __forceinline__ int lower_bound_geometric(const float* __restrict__ a, int n, float v) {
    const float a0 = __ldg(a), a1 = __ldg(a + n - 1);
    if (!(v > a0)) return 0;
    if (v > a1) return n;
    int k = (int)ceilf(__logf(v / a0) / __logf(a1 / a0) * (float)(n - 1));
    k = min(max(k, 0), n - 1);
    while (k > 0 && !(__ldg(a + k - 1) < v)) --k;      // a[k-1] >= v: answer is further left
    while (k < n && __ldg(a + k) < v) ++k;             // a[k] < v: answer is further right
    return k;
}

// Blocked search index over a CDF row (fan-out 8, 4 levels, rows of up to 2048 entries): every level is ONE aligned
// 32-byte sector holding 8 separators, so a lookup costs 4 sector reads instead of the 11 scattered probes of a
// binary search -- the sampling kernel is bound by L2 sector traffic, not by arithmetic.
//   level 0: cdf[256(m+1)-1]                     m = 0..7      [8]
//   level 1: cdf[256 c0 + 32(m+1)-1]                           [8][8]
//   level 2: cdf[256 c0 + 32 c1 + 4(m+1)-1]                    [64][8]
//   level 3: cdf[256 c0 + 32 c1 + 4 c2 + 0..3]  (the row itself)
// Entries past the row end are +inf.  The result is exactly lower_bound(row, u).
constexpr int kIndexPitch = 8 + 64 + 512;   // floats per row

__global__ void k_build_cdf_index(const float* __restrict__ cdf, int rows, int n, float* __restrict__ index) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)rows * kIndexPitch) return;
    const int row = (int)(t / kIndexPitch), e = (int)(t - (int64_t)row * kIndexPitch);
    int src;
    if (e < 8) src = 256 * (e + 1) - 1;
    else if (e < 72) { const int c0 = (e - 8) >> 3, m = (e - 8) & 7; src = 256 * c0 + 32 * (m + 1) - 1; }
    else { const int blk = (e - 72) >> 3, m = (e - 72) & 7; src = 32 * blk + 4 * (m + 1) - 1; }
    index[t] = src < n ? cdf[(int64_t)row * n + src] : __int_as_float(0x7f800000);
}

__device__ __forceinline__ int count_less8(const float* __restrict__ node, float v) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(node)), b = __ldg(reinterpret_cast<const float4*>(node) + 1);
    return (a.x < v) + (a.y < v) + (a.z < v) + (a.w < v) + (b.x < v) + (b.y < v) + (b.z < v) + (b.w < v);
}
__device__ __forceinline__ int lower_bound_indexed(const float* __restrict__ row, const float* __restrict__ idx, int n, float v) {
    const int c0 = min(count_less8(idx, v), 7);
    const int c1 = min(count_less8(idx + 8 + c0 * 8, v), 7);
    const int c2 = min(count_less8(idx + 72 + (c0 * 8 + c1) * 8, v), 7);
    const int base = 256 * c0 + 32 * c1 + 4 * c2;
    int c3 = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) c3 += (base + m < n && __ldg(row + base + m) < v) ? 1 : 0;
    return min(base + c3, n);
}

__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t* hi) {
    const uint64_t p = (uint64_t)a * b;
    *hi = (uint32_t)(p >> 32);
    return (uint32_t)p;
}
// Philox4x32-10, counter = (idx_lo, idx_hi, stream, 0), key = seed
__device__ __forceinline__ void philox(uint64_t seed, uint64_t idx, uint32_t stream, uint32_t r[4]) {
    uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = stream, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t h0, h1;
        const uint32_t l0 = mulhilo(0xD2511F53u, c0, &h0), l1 = mulhilo(0xCD9E8D57u, c2, &h1);
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }  // [0,1)

__global__ void __launch_bounds__(kTile)
k_sample(const float* __restrict__ sigma, const float* __restrict__ sigma_grid, int num_sigma, const float* __restrict__ cdf,
         const float* __restrict__ omega_grid, int num_omega, const float* __restrict__ normals, const float* __restrict__ u,
         uint64_t seed, const float* __restrict__ x, float* __restrict__ out, float* __restrict__ angle_out, int64_t n,
         float tol, const float* __restrict__ cdf_index) {
    __shared__ __align__(16) float s_rot[kTile * 9];
    __shared__ __align__(16) float s_nrm[kTile * 3];
    const int64_t first = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, n - first);
    if (x) tile_load<9>(x, s_rot, first, count);
    if (normals) tile_load<3>(normals, s_nrm, first, count);
    __syncthreads();
    const int t = threadIdx.x;
    if (t < count) {
        const int64_t e = first + t;
        float nx, ny, nz, uu;
        if (normals) {
            nx = s_nrm[t * 3]; ny = s_nrm[t * 3 + 1]; nz = s_nrm[t * 3 + 2];
            uu = u[e];
        } else {
            uint32_t r[4], r2[4];
            philox(seed, (uint64_t)e, 0u, r);
            philox(seed, (uint64_t)e, 1u, r2);
            const float a0 = sqrtf(-2.0f * logf(1.0f - u01(r[0]))), a1 = sqrtf(-2.0f * logf(1.0f - u01(r[2])));
            float s0, c0, s1, c1;
            sincospif(2.0f * u01(r[1]), &s0, &c0);
            sincospif(2.0f * u01(r[3]), &s1, &c1);
            nx = a0 * c0; ny = a0 * s0; nz = a1 * c1;
            uu = u01(r2[0]);
            (void)s1;
        }
        int row = 0;
        float sg = 0.f;
        if (sigma) {
            sg = sigma[e];
            row = lower_bound_geometric(sigma_grid, num_sigma, sg);  // torch.bucketize(sigma, sigma_grid)
            row = row < num_sigma ? row : num_sigma - 1;   // the reference would raise (so3_sde.py:1633)
        }
        const float* c = cdf + (int64_t)row * num_omega;
        int stop = cdf_index ? lower_bound_indexed(c, cdf_index + (int64_t)row * kIndexPitch, num_omega, uu) : lower_bound(c, num_omega, uu);
        stop = stop < num_omega ? stop : num_omega - 1;
        const int start = stop > 0 ? stop - 1 : 0;
        const float c0 = __ldg(c + start), c1 = __ldg(c + stop);
        const float delta = fmaxf(c1 - c0, tol);
        float w = (uu - c0) / delta;
        w = fminf(fmaxf(w, 0.0f), 1.0f);
        const float o0 = __ldg(omega_grid + start), o1 = __ldg(omega_grid + stop);
        // torch.lerp (CPU/CUDA kernels): w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
        const float diff = o1 - o0;
        float ang = w < 0.5f ? o0 + w * diff : o1 - diff * (1.0f - w);
        if (sigma && sg < tol) ang = 0.0f;  // SampleIGSO3._process_angles
        const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
        float v[3] = {(nx / nn) * ang, (ny / nn) * ang, (nz / nn) * ang};
        float r[9];
        so3_exp(v, tol, r);
        if (x) {
            float xr[9], o[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) xr[k] = s_rot[t * 9 + k];
            so3_mul<float, false>(xr, r, o);
#pragma unroll
            for (int k = 0; k < 9; ++k) s_rot[t * 9 + k] = o[k];
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) s_rot[t * 9 + k] = r[k];
        }
        if (angle_out) angle_out[e] = ang;
    }
    __syncthreads();
    tile_store<9>(out, s_rot, first, count);
}

inline int series_grid(int64_t n) {
    const int64_t blocks = (n + 7) / 8;  // 8 warps per CTA
    return (int)(blocks < 148 * 16 ? (blocks > 0 ? blocks : 1) : 148 * 16);
}

}  // namespace

#define SE3_LAUNCH_CHECK(name) \
    count_launch();            \
    return check_launch(name)

extern "C" {

int se3_igso3_series_f32(const float* omega, const float* sigma, float* f, float* df, float* dlog, int64_t n, int l_max,
                         float tol, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && l_max >= 0, "negative size");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(omega && sigma && (f || df || dlog), "null pointer");
    k_series<float, 0><<<series_grid(n), 256, 0, (cudaStream_t)stream>>>(omega, sigma, f, df, dlog, n, l_max, tol);
    SE3_LAUNCH_CHECK("se3_igso3_series_f32");
}

int se3_igso3_series_f64(const double* omega, const double* sigma, double* f, double* df, double* dlog, int64_t n,
                         int l_max, double tol, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && l_max >= 0, "negative size");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(omega && sigma && (f || df || dlog), "null pointer");
    k_series<double, 0><<<series_grid(n), 256, 0, (cudaStream_t)stream>>>(omega, sigma, f, df, dlog, n, l_max, tol);
    SE3_LAUNCH_CHECK("se3_igso3_series_f64");
}

int se3_igso3_score(const float* rotvec, const float* sigma, float* score, int64_t n, int l_max, float tol,
                    se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && l_max >= 0, "negative size");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(rotvec && sigma && score, "null pointer");
    k_series<float, 1><<<series_grid(n), 256, 0, (cudaStream_t)stream>>>(rotvec, sigma, score, nullptr, nullptr, n, l_max, tol);
    SE3_LAUNCH_CHECK("se3_igso3_score");
}

int se3_igso3_marginal_pdf(const float* omega, const float* omega0, const float* sigma, float* pdf, int64_t n, int l_count,
                           float tol, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && l_count >= 1, "bad size");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(omega && omega0 && sigma && pdf, "null pointer");
    k_marginal<<<series_grid(n), 256, 0, (cudaStream_t)stream>>>(omega, omega0, sigma, pdf, n, l_count, tol);
    SE3_LAUNCH_CHECK("se3_igso3_marginal_pdf");
}

static int table_launch(int mode, const float* sigma_grid, int num_sigma, const double* omega, int n_pts, int l_max,
                        double tol, int uniform, float* out, cudaStream_t st) {
    const size_t smem = (size_t)(l_max + 1 + n_pts) * sizeof(double);
    auto kern = mode == 0 ? k_table_row<0> : k_table_row<1>;
    if (smem > 48 * 1024) {
        if (smem > 227 * 1024) { set_error("igso3 table: l_max + n_omega too large for shared memory"); return SE3_EUNSUPPORTED; }
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("igso3 table smem attribute: %s", cudaGetErrorString(e)); return SE3_ECUDA; }
    }
    kern<<<uniform ? 1 : num_sigma, kRowThreads, smem, st>>>(sigma_grid, omega, n_pts, l_max, tol, uniform, out);
    count_launch();
    return check_launch("se3_igso3_build_table");
}

int se3_igso3_build_cdf(const float* sigma_grid, int num_sigma, const double* omega_pts, int n_pts, int l_max, double tol,
                        int uniform, float* cdf, se3_stream_t stream) {
    SE3_REQUIRE(num_sigma >= 1 && n_pts >= 2 && l_max >= 0, "bad size");
    SE3_REQUIRE(sigma_grid && omega_pts && cdf, "null pointer");
    return table_launch(0, sigma_grid, num_sigma, omega_pts, n_pts, l_max, tol, uniform, cdf, (cudaStream_t)stream);
}

int se3_igso3_build_score_scaling(const float* sigma_grid, int num_sigma, const double* omega_pts, int n_pts, int l_max,
                                  double tol, float* score_scaling, se3_stream_t stream) {
    SE3_REQUIRE(num_sigma >= 1 && n_pts >= 1 && l_max >= 0, "bad size");
    SE3_REQUIRE(sigma_grid && omega_pts && score_scaling, "null pointer");
    return table_launch(1, sigma_grid, num_sigma, omega_pts, n_pts, l_max, tol, 0, score_scaling, (cudaStream_t)stream);
}

int se3_igso3_build_cdf_index(const float* cdf, int num_rows, int num_omega, float* index, se3_stream_t stream) {
    SE3_REQUIRE(cdf && index && num_rows >= 1 && num_omega >= 1, "bad argument");
    SE3_REQUIRE(num_omega <= 2048, "the blocked index covers rows of up to 2048 entries");
    const int64_t total = (int64_t)num_rows * kIndexPitch;
    k_build_cdf_index<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cdf, num_rows, num_omega, index);
    SE3_LAUNCH_CHECK("se3_igso3_build_cdf_index");
}

int se3_igso3_sample(const float* sigma, const float* sigma_grid, int num_sigma, const float* cdf, const float* omega_grid,
                     int num_omega, const float* normals, const float* u, uint64_t seed, const float* x, float* out,
                     float* angle_out, int64_t n, float tol, const float* cdf_index, se3_stream_t stream) {
    SE3_REQUIRE(n >= 0 && num_omega >= 1, "bad size");
    if (n == 0) return SE3_OK;
    SE3_REQUIRE(cdf && omega_grid && out, "null pointer");
    SE3_REQUIRE(!sigma || (sigma_grid && num_sigma >= 1), "sigma given without sigma_grid");
    SE3_REQUIRE((normals == nullptr) == (u == nullptr), "normals and u must be given together");
    k_sample<<<(unsigned)((n + kTile - 1) / kTile), kTile, 0, (cudaStream_t)stream>>>(sigma, sigma_grid, num_sigma, cdf, omega_grid,
                                                                                    num_omega, normals, u, seed, x, out, angle_out, n, tol, num_omega <= 2048 ? cdf_index : nullptr);
    SE3_LAUNCH_CHECK("se3_igso3_sample");
}

}  // extern "C"
