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
#include <stdlib.h>

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
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// how a small grid is read (the read-only path of global memory)
struct GlobalGrid { const float* p; __device__ __forceinline__ float operator()(int i) const { return __ldg(p + i); } };
template <typename Grid>
__device__ __forceinline__ int lower_bound_geometric(const Grid a, int n, float v) {
    const float a0 = a(0), a1 = a(n - 1);
    if (!(v > a0)) return 0;
    if (v > a1) return n;
    // a[0] < v <= a[n-1]: the answer lies in [1, n-1].  The guess (flush-to-zero MUFU forms: the plain __log2f / __fdividef carry
    // denormal guards, 16 more instructions per rotation) may be off by a slot or two; both neighbours of the guess are fetched
    // in ONE round trip and settle the usual case, the two loops below make every other case exact
    const float l0 = lg2_ftz(a0);
    int k = (int)ceilf((lg2_ftz(v) - l0) * rcp_ftz(lg2_ftz(a1) - l0) * (float)(n - 1));
    k = min(max(k, 1), n - 1);
    const float below = a(k - 1), at = a(k);
    if (below < v && !(at < v)) return k;
    while (k > 0 && !(a(k - 1) < v)) --k;              // a[k-1] >= v: answer is further left
    while (k < n && a(k) < v) ++k;                     // a[k] < v: answer is further right
    return k;
}

// Guide records over a CDF row: the unit interval is cut into G = 2^k bins (G ~ n/2) and bin g of a row owns ONE aligned
// 32-byte record
//     word 0    : lo | hi << 16,  lo = lower_bound(row, g/G), hi = lower_bound(row, (g+1)/G)
//     words 1..7: hi - lo <= 5 (the usual case: a bin holds ~2 grid points): cdf[lo-1 .. lo+5]  (+inf past the row end; word 1
//                 is unused when lo == 0) -- the record also contains both CDF values the interpolation needs, so the lookup
//                 costs ONE L2 sector instead of the 11 scattered probes of a binary search (the sampling kernel is bound by
//                 L2 sector traffic and by the latency of dependent probes, not by arithmetic);
//                 hi - lo > 5 (flat stretches of the CDF -- its x^9 start on the cubic omega grid and its Gaussian tail; ~3 % of
//                 the draws, i.e. about every second warp has one): lower_bound(row, (8g + k)/(8G)), k = 1..7 -- the bin's
//                 eight sub-bins, which narrow [lo, hi] eightfold; the seven entries around the narrowed range are then fetched
//                 in ONE round trip and a bounded binary search only remains for the few sub-bins that still hold more than five
//                 grid points (before: 4.4 dependent L2 probes per such warp).
// u * G, u * 8G and the bin edges are exact in fp32 (power-of-two scaling), so the result is exactly lower_bound(row, u) =
// sum(cdf < u).
__host__ __device__ inline int guide_bins(int n) {   // smallest power of two >= n/2, within [8, 4096]: n = 2000 -> 1024
    int g = 8;
    while (g < (n + 1) / 2 && g < 4096) g *= 2;
    return g;
}

__global__ void k_build_cdf_index(const float* __restrict__ cdf, int rows, int n, int G, float* __restrict__ index) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)rows * G) return;
    const int row = (int)(t / G), g = (int)(t - (int64_t)row * G);
    const float* c = cdf + (int64_t)row * n;
    const int lo = lower_bound(c, n, (float)g / (float)G);
    const int hi = g + 1 < G ? lower_bound(c, n, (float)(g + 1) / (float)G) : n;
    float w[8];
    const float inf = __int_as_float(0x7f800000);
    w[0] = __uint_as_float((uint32_t)lo | ((uint32_t)hi << 16));
    if (hi - lo <= 5) {
        for (int k = 0; k < 7; ++k) { const int i = lo - 1 + k; w[1 + k] = (i >= 0 && i < n) ? c[i] : inf; }
    } else {
        for (int k = 1; k < 8; ++k) w[k] = __uint_as_float((uint32_t)lower_bound(c, n, (float)(8 * g + k) / (float)(8 * G)));
    }
    float4* dst = reinterpret_cast<float4*>(index + t * 8);
    dst[0] = make_float4(w[0], w[1], w[2], w[3]);
    dst[1] = make_float4(w[4], w[5], w[6], w[7]);
}

// v[k] = cdf[lo - 1 + k] for a range [lo, hi] with hi - lo <= 5 that holds lower_bound(row, u): the index, and the two values the
// reference interpolates between (cdf[max(stop-1,0)], cdf[stop]; so3_sde.py:1268-1281).  False when u lies above the whole row.
__device__ __forceinline__ bool settle_short_range(const float (&v)[7], int lo, int hi, int n, float u, int& stop, float& c0, float& c1) {
    // p[k] = "cdf[lo+k] < u" is monotone in k (sorted row), cnt = number of true ones; the two values the interpolation needs
    // are v[cnt] (= cdf[stop-1]) and v[cnt+1] (= cdf[stop]): nested selects on the predicates
    bool p[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) p[k] = (k < hi - lo) && (v[1 + k] < u);
    const int cnt = (int)p[0] + (int)p[1] + (int)p[2] + (int)p[3] + (int)p[4];
    stop = lo + cnt;
    if (stop >= n) return false;                 // (stop == n: u above the whole row -> generic path)
    float lo_v = v[0], hi_v = v[1];
#pragma unroll
    for (int k = 0; k < 5; ++k) { lo_v = p[k] ? v[k + 1] : lo_v; hi_v = p[k] ? v[k + 2] : hi_v; }
    c1 = hi_v;
    c0 = stop > 0 ? lo_v : hi_v;                 // start = max(stop - 1, 0)
    return true;
}

// lower_bound(row, u) through the guide records, also returning cdf[max(stop-1,0)] and cdf[min(stop,n-1)] with `stop` already
// clipped to n-1
__device__ __forceinline__ int lookup_guided(const float* __restrict__ row, const float* __restrict__ rec_row, int n, int G, float u,
                                             float& c0, float& c1) {
    int lo = 0, hi = n;
    if (u >= 0.0f && u < 1.0f) {
        const int g = min((int)(u * (float)G), G - 1);
        const float* rec = rec_row + g * 8;
        const float4 a = __ldg(reinterpret_cast<const float4*>(rec)), b = __ldg(reinterpret_cast<const float4*>(rec) + 1);
        const uint32_t w = __float_as_uint(a.x);
        lo = (int)(w & 0xffffu); hi = (int)(w >> 16);
        int stop;
        if (hi - lo <= 5) {
            const float v[7] = {a.y, a.z, a.w, b.x, b.y, b.z, b.w};   // cdf[lo-1+k]
            if (settle_short_range(v, lo, hi, n, u, stop, c0, c1)) return stop;
        } else {
            // long bin: its sub-bin [ (8g+s)/(8G), (8g+s+1)/(8G) ) narrows the range (the record was just fetched: L1 hits)
            const int s = min(max((int)(u * (float)(8 * G)) - 8 * g, 0), 7);
            const int slo = s > 0 ? (int)__float_as_uint(__ldg(rec + s)) : lo;
            const int shi = s < 7 ? (int)__float_as_uint(__ldg(rec + s + 1)) : hi;
            lo = slo; hi = shi;
            if (hi - lo <= 5) {
                float v[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) { const int i = lo - 1 + k; v[k] = (i >= 0 && i < n) ? __ldg(row + i) : __int_as_float(0x7f800000); }
                if (settle_short_range(v, lo, hi, n, u, stop, c0, c1)) return stop;
            }
        }
    }
    while (lo < hi) {                                    // long run (or u outside [0,1)): bounded binary search
        const int mid = (lo + hi) >> 1;
        if (__ldg(row + mid) < u) lo = mid + 1; else hi = mid;
    }
    int stop = lo < n ? lo : n - 1;
    const int start = stop > 0 ? stop - 1 : 0;
    c0 = __ldg(row + start); c1 = __ldg(row + stop);
    return stop;
}

__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t* hi) {
    const uint64_t p = (uint64_t)a * b;
    *hi = (uint32_t)(p >> 32);
    return (uint32_t)p;
}
// Philox4x32-7, counter = (idx_lo, idx_hi, stream, 0), key = seed
__device__ __forceinline__ void philox(uint64_t seed, uint64_t idx, uint32_t stream, uint32_t r[4]) {
    uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = stream, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 7; ++i) {      // Philox4x32-7: the shortest variant that passes BigCrush (Salmon et al. 2011); 10 rounds cost 26 more instructions per rotation
        uint32_t h0, h1;
        const uint32_t l0 = mulhilo(0xD2511F53u, c0, &h0), l1 = mulhilo(0xCD9E8D57u, c2, &h1);
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }  // [0,1)

__device__ __forceinline__ float sin_mufu(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cos_mufu(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- per-rotation pieces of the sampler ------------------------------------------------------------------------------------------
struct SampleDraw { int stop; float c0, c1, sg, uu, nx, ny, nz; };

// everything that does not need the operand tiles: the uniform (passed in or Philox), the sigma row, the CDF lookup
template <typename Grid>
__device__ __forceinline__ SampleDraw sample_lookup(int64_t e, bool have_normals, float u_in, float sg_in, bool have_sigma, const Grid sigma_grid,
                                                    int num_sigma, const float* __restrict__ cdf, int num_omega, uint64_t seed,
                                                    const float* __restrict__ cdf_index, int guide_bins_n) {
    SampleDraw d = {0, 0.f, 0.f, sg_in, u_in, 0.f, 0.f, 0.f};
    if (!have_normals) {
        // One Philox4x32-7 block per rotation: a direction uniform on the sphere from two uniforms (z = 2a - 1, phi = 2 pi b) --
        // the same law as the reference's normalised Gaussian triple (so3_sde.py:1229-1242), which would take four uniforms,
        // two logarithms and a second block -- and the CDF uniform from the third.  (Bit parity with the reference's torch
        // generator is the business of the noise-passed-in mode; this mode only has to draw from the same distribution.)
        // The direction is a UNIT vector by construction (sample_compose does not normalise it); its trigonometry is the short
        // fused-FMA sincos of the frame kernels (libdevice sincospif is twice as long without FMA contraction in this unit).
        uint32_t r[4];
        philox(seed, (uint64_t)e, 0u, r);
        const float z = fmaf(2.0f, u01(r[0]), -1.0f);
        float rho;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rho) : "f"(fmaxf(fmaf(-z, z, 1.0f), 0.0f)));
        // azimuth in [-pi, pi): the MUFU sine / cosine (abs. error ~5e-7 on that range) -- a random direction needs no more
        const float phi = fmaf(6.283185307179586f, u01(r[1]), -3.141592653589793f);
        const float sp = sin_mufu(phi), cp = cos_mufu(phi);
        d.nx = rho * cp; d.ny = rho * sp; d.nz = z;
        d.uu = u01(r[2]);
    }
    int row = 0;
    if (have_sigma) {
        row = lower_bound_geometric(sigma_grid, num_sigma, d.sg);  // torch.bucketize(sigma, sigma_grid)
        row = row < num_sigma ? row : num_sigma - 1;   // the reference would raise (so3_sde.py:1633)
    }
    // (32-bit offsets: se3_igso3_sample checks num_sigma * num_omega and the record count against 2^31)
    const float* c = cdf + (uint32_t)row * (uint32_t)num_omega;
    if (cdf_index) {
        d.stop = lookup_guided(c, cdf_index + (uint32_t)row * (uint32_t)(guide_bins_n * 8), num_omega, guide_bins_n, d.uu, d.c0, d.c1);
    } else {
        int stop = lower_bound(c, num_omega, d.uu);
        stop = stop < num_omega ? stop : num_omega - 1;
        d.c0 = __ldg(c + (stop > 0 ? stop - 1 : 0)); d.c1 = __ldg(c + stop);
        d.stop = stop;
    }
    return d;
}

// angle by interpolation (bit-exact against torch.lerp), axis-angle -> rotation, [x .] r written over the rotation slot `s_rot9`
// unit_axis: (nx, ny, nz) already has norm one (Philox mode)
template <typename Grid>
__device__ __forceinline__ float sample_compose(const SampleDraw& d, bool have_sigma, bool have_x, bool unit_axis, const Grid omega_grid, float tol,
                                                float* s_rot9) {
    const int start = d.stop > 0 ? d.stop - 1 : 0;
    const float delta = fmaxf(d.c1 - d.c0, tol);
    float w = (d.uu - d.c0) / delta;
    w = fminf(fmaxf(w, 0.0f), 1.0f);
    const float o0 = omega_grid(start), o1 = omega_grid(d.stop);
    // torch.lerp (CPU/CUDA kernels): w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
    const float diff = o1 - o0;
    // ATen contracts both branches into one FMA (CPU: vec::fmadd in lerp_vec / -mfma scalar code; CUDA: nvcc -fmad), and this
    // translation unit is built with -fmad=false, so the FMAs are spelled out
    float ang = w < 0.5f ? fmaf(w, diff, o0) : fmaf(-diff, 1.0f - w, o1);
    if (have_sigma && d.sg < tol) ang = 0.0f;  // SampleIGSO3._process_angles
    // Axis-angle -> rotation straight from the UNIT axis and the angle (Rodrigues: E = I + sin(ang) K + (1 - cos(ang)) K^2 with
    // K^2 = n n^T - I): no rotation vector is formed, so neither its norm nor the two divisions by it (nor the Taylor branch
    // that guards them) exist.  The angle above is the bit-exact part of this kernel; the matrix entries depend on sin / cos and
    // agree with the reference to ~1e-6 either way (same tolerance as before).
    float ax = d.nx, ay = d.ny, az = d.nz;
    if (!unit_axis) {
        float inv;                                           // 1 / |n|: one MUFU.RSQ
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaf(d.nx, d.nx, fmaf(d.ny, d.ny, d.nz * d.nz))));
        ax *= inv; ay *= inv; az *= inv;
    }
    // the angle lies in [0, pi]: MUFU sine / cosine (abs. error ~5e-7 there, i.e. the same ~1e-6 class as every other rotation
    // entry of this library against the CPU reference's libm), two instructions each instead of the ~25 of a reduced polynomial
    const float sn = sin_mufu(ang), cs = cos_mufu(ang);
    const float b = 1.0f - cs;
    const float bx = b * ax, by = b * ay;
    const float sx = sn * ax, sy = sn * ay, sz = sn * az;
    const float e00 = fmaf(-b, fmaf(ay, ay, az * az), 1.0f), e11 = fmaf(-b, fmaf(ax, ax, az * az), 1.0f), e22 = fmaf(-b, fmaf(ax, ax, ay * ay), 1.0f);
    const float e01 = fmaf(bx, ay, -sz), e10 = fmaf(bx, ay, sz);
    const float e02 = fmaf(bx, az, sy), e20 = fmaf(bx, az, -sy);
    const float e12 = fmaf(by, az, -sx), e21 = fmaf(by, az, sx);
    if (have_x) {
        float xr[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) xr[k] = s_rot9[k];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float r0 = xr[i * 3], r1 = xr[i * 3 + 1], r2 = xr[i * 3 + 2];
            s_rot9[i * 3] = fmaf(r2, e20, fmaf(r1, e10, r0 * e00));
            s_rot9[i * 3 + 1] = fmaf(r2, e21, fmaf(r1, e11, r0 * e01));
            s_rot9[i * 3 + 2] = fmaf(r2, e22, fmaf(r1, e12, r0 * e02));
        }
    } else {
        s_rot9[0] = e00; s_rot9[1] = e01; s_rot9[2] = e02;
        s_rot9[3] = e10; s_rot9[4] = e11; s_rot9[5] = e12;
        s_rot9[6] = e20; s_rot9[7] = e21; s_rot9[8] = e22;
    }
    return ang;
}

struct SampleArgs {
    const float *sigma, *sigma_grid, *cdf, *omega_grid, *normals, *u, *x, *cdf_index;
    float *out, *angle_out;
    uint64_t seed;
    int64_t n;
    int num_sigma, num_omega, guide_bins_n;
    float tol;
};

// ragged last tile / unaligned operand arrays: every choice is made at run time (one tile per call takes this path, or all of them)
__device__ __noinline__ void sample_tile_generic(const SampleArgs& a, float* s_rot, float* s_nrm, int64_t first, int count) {
    if (a.x) warp_tile_load_async<9>(a.x, s_rot, first, count);
    if (a.normals) warp_tile_load_async<3>(a.normals, s_nrm, first, count);
    const int t = threadIdx.x;
    const int64_t e = first + t;
    SampleDraw d = {};
    if (t < count)
        d = sample_lookup(e, a.normals != nullptr, a.normals ? a.u[e] : 0.f, a.sigma ? a.sigma[e] : 0.f, a.sigma != nullptr, GlobalGrid{a.sigma_grid}, a.num_sigma, a.cdf,
                          a.num_omega, a.seed, a.cdf_index, a.guide_bins_n);
    tile_load_wait();
    __syncwarp();
    if (t < count) {
        if (a.normals) { d.nx = s_nrm[t * 3]; d.ny = s_nrm[t * 3 + 1]; d.nz = s_nrm[t * 3 + 2]; }
        const float ang = sample_compose(d, a.sigma != nullptr, a.x != nullptr, a.normals == nullptr, GlobalGrid{a.omega_grid}, a.tol, s_rot + t * 9);
        if (a.angle_out) a.angle_out[e] = ang;
    }
    __syncwarp();
    warp_tile_store<9>(a.out, s_rot, first, count);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
// One rotation per thread, kSampleTile per CTA; every warp moves, computes and stores its own 32 rotations (only __syncwarp() between the
// phases).  Three independent fetches per rotation -- the operand tiles, (sigma, u), and the guide record that depends on
// (sigma, u) -- are in flight together: the tiles by cp.async, so the table lookup proceeds under them.
// Which operands exist is a template parameter and full, 16-byte-aligned tiles take a straight-line path: the run-time edition of
// the same body spent ~150 of its ~510 instructions per rotation on null-pointer / raggedness / alignment tests, 64-bit address
// arithmetic and constant-bank reloads (ncu source counters of the r2 capture), and the kernel is bound by its instruction
// stream and the latency of its dependent lookups at 64 resident warps per SM: 32 registers are part of the design.
// Measured and dropped (r4): persistent CTAs that stage the sigma and omega grids (4 + 8 KB) in shared memory once, to take the
// scattered grid reads off the L1 data pipe (the busiest unit, 69 % in the r2 capture) -- 40 registers (48 warps per SM):
// 0.55 / 0.50 of the HBM roof against 0.61 / 0.55; capped at 32 registers: 0.37.  A single 256-bit load of the guide record
// (ld.global.v8.f32) crashes ptxas 12.9 in this kernel.
constexpr int kSampleTile = 128;   // rotations (= threads) per tile (256: -1 %, 64: the same; the warps of a CTA are independent)
// Tiles per CTA.  With two, the second tile's (sigma, u) and operand tiles are in flight while the first is computed: 40 registers
// instead of 32 (48 resident warps per SM instead of 64).  Measured at n = 1e7 (fraction of the HBM roof, random sigma | one sigma):
// Philox mode 0.556 | 0.767 with one tile, 0.591 | 0.745 with two; noise passed in 0.618 | 0.911 and 0.626 | 0.862 -- so the Philox
// mode takes two and the parity mode one.  (Two tiles capped at 32 registers: 0.35 with the noise passed in.  Philox mode with three /
// four tiles: 0.50 / 0.32 | 0.70 / 0.55 -- the registers of the extra chains cost more resident warps than the chains hide.)
__host__ __device__ constexpr int sample_tiles(bool noise_passed_in) { return noise_passed_in ? 1 : 2; }
template <bool kX, bool kNormals, bool kSigma>
__global__ void __launch_bounds__(kSampleTile)
k_sample(const __grid_constant__ SampleArgs a, const int fast) {
    constexpr int kSampleTiles = sample_tiles(kNormals);
    __shared__ __align__(16) float s_rot[kSampleTiles][kSampleTile * 9];
    __shared__ __align__(16) float s_nrm[kSampleTiles][kNormals ? kSampleTile * 3 : 4];
    const int64_t first0 = (int64_t)blockIdx.x * (kSampleTile * kSampleTiles);
    const int t = threadIdx.x;
    const int lane = t & 31;
    const int wbase = t & ~31;                              // first rotation of this warp inside a tile
    float sg[kSampleTiles], uin[kSampleTiles];
    bool full[kSampleTiles];
    // the lookup chains start from (sigma, u): issued before anything else, for every tile of the CTA
#pragma unroll
    for (int k = 0; k < kSampleTiles; ++k) {
        const int64_t first = first0 + k * kSampleTile;
        full[k] = fast && a.n - first >= kSampleTile;
        sg[k] = (kSigma && full[k]) ? __ldg(a.sigma + first + t) : 0.f;
        uin[k] = (kNormals && full[k]) ? __ldg(a.u + first + t) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kSampleTiles; ++k) {
        if (full[k]) {
            const int64_t wfirst = first0 + k * kSampleTile + wbase;
            if (kX) {
                const float4* src = reinterpret_cast<const float4*>(a.x + wfirst * 9) + lane;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_rot[k] + wbase * 9) + (uint32_t)lane * 16u;   // the warp's 32 x 9 floats = 72 16-byte pieces
                cp_async16(dst, src);
                cp_async16(dst + 512u, src + 32);
                if (lane < 8) cp_async16(dst + 1024u, src + 64);
            }
            if (kNormals) {
                if (lane < 24) cp_async16((uint32_t)__cvta_generic_to_shared(s_nrm[k] + wbase * 3) + (uint32_t)lane * 16u, reinterpret_cast<const float4*>(a.normals + wfirst * 3) + lane);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
#pragma unroll
    for (int k = 0; k < kSampleTiles; ++k) {
        const int64_t first = first0 + k * kSampleTile;
        if (first >= a.n) break;
        if (!full[k]) {                                     // ragged last tile / unaligned arrays: the run-time edition
            sample_tile_generic(a, s_rot[k], s_nrm[k], first, (int)min((int64_t)kSampleTile, a.n - first));
            continue;
        }
        const int64_t e = first + t;
        SampleDraw d = sample_lookup(e, kNormals, uin[k], sg[k], kSigma, GlobalGrid{a.sigma_grid}, a.num_sigma, a.cdf, a.num_omega, a.seed, a.cdf_index, a.guide_bins_n);
        if (k + 1 < kSampleTiles) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (kNormals) { d.nx = s_nrm[k][t * 3]; d.ny = s_nrm[k][t * 3 + 1]; d.nz = s_nrm[k][t * 3 + 2]; }
        const float ang = sample_compose(d, kSigma, kX, !kNormals, GlobalGrid{a.omega_grid}, a.tol, s_rot[k] + t * 9);
        if (a.angle_out) a.angle_out[e] = ang;
        __syncwarp();
        float4* dst = reinterpret_cast<float4*>(a.out + (first + wbase) * 9) + lane;
        const float4* s4 = reinterpret_cast<const float4*>(s_rot[k] + wbase * 9) + lane;
        dst[0] = s4[0];
        dst[32] = s4[32];
        if (lane < 8) dst[64] = s4[64];
    }
}

// Persisting-L2 access window over a lookup table for the launches enqueued on `st` until clear_l2_window (SE3DIFF_B200_L2_WINDOW=0
// disables).  The device's persisting carve-out is raised once to what the table needs (at most what the device allows).
inline bool l2_window_enabled() {
    static const bool on = [] { const char* v = getenv("SE3DIFF_B200_L2_WINDOW"); return !(v && v[0] == '0'); }();
    return on;
}
inline bool set_l2_window(cudaStream_t st, const void* base, size_t bytes) {
    int dev = 0, max_persist = 0, max_window = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev) != cudaSuccess || max_persist <= 0 || max_window <= 0) {
        (void)cudaGetLastError();
        return false;
    }
    size_t want = bytes < (size_t)max_persist ? bytes : (size_t)max_persist, have = 0;
    if (cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize) != cudaSuccess || (have < want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess)) {
        (void)cudaGetLastError();
        return false;
    }
    cudaStreamAttrValue a = {};
    a.accessPolicyWindow.base_ptr = const_cast<void*>(base);
    a.accessPolicyWindow.num_bytes = bytes < (size_t)max_window ? bytes : (size_t)max_window;
    a.accessPolicyWindow.hitRatio = want >= bytes ? 1.0f : (float)want / (float)bytes;
    a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return true;
}
inline void clear_l2_window(cudaStream_t st) {
    cudaStreamAttrValue a = {};
    a.accessPolicyWindow.num_bytes = 0;
    (void)cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a);
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

int64_t se3_igso3_cdf_index_floats(int num_rows, int num_omega) {
    if (num_rows < 1 || num_omega < 1 || num_omega > 65535) return SE3_EINVAL;
    return (int64_t)num_rows * guide_bins(num_omega) * 8;
}

int se3_igso3_build_cdf_index(const float* cdf, int num_rows, int num_omega, float* index, se3_stream_t stream) {
    SE3_REQUIRE(cdf && index && num_rows >= 1 && num_omega >= 1, "bad argument");
    SE3_REQUIRE(num_omega <= 65535, "the guide records hold 16-bit grid indices");
    SE3_REQUIRE((reinterpret_cast<uintptr_t>(index) & 31) == 0, "index must be 32-byte aligned");
    const int G = guide_bins(num_omega);
    const int64_t total = (int64_t)num_rows * G;
    k_build_cdf_index<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cdf, num_rows, num_omega, G, index);
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
    SE3_REQUIRE((int64_t)(sigma ? num_sigma : 1) * num_omega < (int64_t)1 << 31, "table too large (32-bit offsets inside the kernel)");
    SE3_REQUIRE((int64_t)(sigma ? num_sigma : 1) * guide_bins(num_omega) * 8 < (int64_t)1 << 31, "guide records too large (32-bit offsets inside the kernel)");
    // (A pipelined per-warp edition like k_em_pipe was measured and dropped: 0.529 -> 0.536 of the HBM roof with the noise passed
    // in, 0.478 -> 0.399 in Philox mode -- this kernel is bound by its ~250 instructions per rotation at 64 resident warps, and
    // the ring costs occupancy.)
    cudaStream_t st = (cudaStream_t)stream;
    const float* index = num_omega <= 65535 ? cdf_index : nullptr;
    // Large draws with per-rotation sigma touch the guide records (32 MB for the shipped 1000 x 2000 table) at random while
    // ~1 GB of operands streams through L2 and evicts them: the records are given a persisting-L2 access window for this launch.
    bool window = false;
    if (index && sigma && n >= (1 << 20) && l2_window_enabled()) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone)
            window = set_l2_window(st, index, (size_t)num_sigma * guide_bins(num_omega) * 32);
    }
    const SampleArgs a = {sigma, sigma_grid, cdf, omega_grid, normals, u, x, index, out, angle_out, seed, n, num_sigma, num_omega, guide_bins(num_omega), tol};
    // straight-line path: 16-byte-aligned operand arrays
    const int fast = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(normals) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int per_cta = kSampleTile * sample_tiles(normals != nullptr);
    const unsigned grid = (unsigned)((n + per_cta - 1) / per_cta);
    const int which = (x ? 4 : 0) | (normals ? 2 : 0) | (sigma ? 1 : 0);
    switch (which) {
#define SE3_SAMPLE_CASE(w, X, N, S) case w: k_sample<X, N, S><<<grid, kSampleTile, 0, st>>>(a, fast); break;
        SE3_SAMPLE_CASE(0, false, false, false) SE3_SAMPLE_CASE(1, false, false, true) SE3_SAMPLE_CASE(2, false, true, false) SE3_SAMPLE_CASE(3, false, true, true)
        SE3_SAMPLE_CASE(4, true, false, false) SE3_SAMPLE_CASE(5, true, false, true) SE3_SAMPLE_CASE(6, true, true, false) SE3_SAMPLE_CASE(7, true, true, true)
#undef SE3_SAMPLE_CASE
    }
    if (window) clear_l2_window(st);
    SE3_LAUNCH_CHECK("se3_igso3_sample");
}

}  // extern "C"
