// TMEM read throughput, small-N tcgen05.mma cadence and sqrt-pipe mix probes (developer diagnostics).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../se3diff_b200/csrc tmem.cu -o tmem
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

// MODE 0: x16 loads, wait after each;  MODE 1: x32 loads;  MODE 2: 4 x16 loads then one wait
template <int MODE>
__global__ void k_ld(float* out, long long* cyc, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tc::tmem_alloc(&slot, 256);
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = slot;
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[16];
                tc::tmem_ld16(tc::tmem_addr(tmem, lane_base, ((it & 3) * 4 + c) * 16), r);
                tc::tmem_wait_ld();
#pragma unroll
                for (int u = 0; u < 16; ++u) acc ^= r[u];
            }
        } else if (MODE == 1) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld32(tc::tmem_addr(tmem, lane_base, ((it & 3) * 2 + c) * 32), r);
                tc::tmem_wait_ld();
#pragma unroll
                for (int u = 0; u < 32; ++u) acc ^= r[u];
            }
        } else {
            uint32_t r[4][16];
#pragma unroll
            for (int c = 0; c < 4; ++c) tc::tmem_ld16(tc::tmem_addr(tmem, lane_base, ((it & 3) * 4 + c) * 16), r[c]);
            tc::tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int u = 0; u < 16; ++u) acc ^= r[c][u];
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// loads of 64 columns + the sqrt work the IPA logit pass would do on them (4 points x 16 keys) -- do the two overlap?
template <bool WITH_LD, bool WITH_SQRT>
__global__ void k_mix(float* out, long long* cyc, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tc::tmem_alloc(&slot, 256);
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = slot;
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32;
    float accf = 0.f;
    uint32_t r[4][16];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int u = 0; u < 16; ++u) r[c][u] = __float_as_uint(1.0f + threadIdx.x + c + u);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (WITH_LD) {
#pragma unroll
            for (int c = 0; c < 4; ++c) tc::tmem_ld16(tc::tmem_addr(tmem, lane_base, ((it & 3) * 4 + c) * 16), r[c]);
            tc::tmem_wait_ld();
        }
        if (WITH_SQRT) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v;
                    asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(fabsf(__uint_as_float(r[c][u]) + accf)));
                    s += v;
                }
                accf = fmaf(s, 1e-9f, accf);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int u = 0; u < 16; ++u) accf += __uint_as_float(r[c][u]);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = accf;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// n back-to-back MMAs (M = 128, N, K = 16, bf16) issued by one thread, then a commit: cycles per MMA
__global__ void k_mma(long long* cyc, int n_mma, int N) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (warp == 0) tc::tmem_alloc(&slot, 256);
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(128, N);
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i)
            tc::mma_bf16(tmem + (uint32_t)((i & 3) * 64), tc::make_desc(tc::smem_u32(sm), 128), tc::make_desc(tc::smem_u32(sm) + 8192, (uint32_t)N), idesc, false);
        tc::mma_commit(&bar);
        const long long t1 = clock64();
        tc::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        cyc[0] = t1 - t0;
        cyc[1] = t2 - t0;
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// sqrt through the FMA pipe: bit-trick seed + Newton steps on rsqrt, d = x * y.  OP 0: MUFU.SQRT, 1: one Newton step, 2: two steps, 3: half MUFU half FMA (2 steps)
__device__ __forceinline__ float sqrt_fma(float x, int steps) {
    float y = __uint_as_float(0x5f375a86u - (__float_as_uint(x) >> 1));
    const float hx = 0.5f * x;
    y = y * fmaf(-hx * y, y, 1.5f);
    if (steps > 1) y = y * fmaf(-hx * y, y, 1.5f);
    return x * y;
}
template <int OP>
__global__ void k_sqrt(float* out, long long* cyc, int iters) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = 1.5f + threadIdx.x * 0.001f + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float r;
            if (OP == 0 || (OP == 3 && (i & 1))) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v[i]));
            else r = sqrt_fma(v[i], OP == 1 ? 1 : 2);
            v[i] = r + 1.25f;
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 2 * 1024 * 4); cudaMalloc(&cyc, 64);
    long long c[2];
    const int iters = 4000;
    for (int threads : {128, 256, 512}) {
        for (int ctas : {1, 2}) {
            if (threads * ctas > 1024) continue;
            auto rep = [&](const char* nm, double bytes_per_iter_per_warp) {
                cudaDeviceSynchronize();
                cudaMemcpy(c, cyc, 8, cudaMemcpyDeviceToHost);
                const double warps = threads / 32.0 * ctas;
                printf("%-22s threads=%3d ctas/SM=%d: %8lld cyc, %.1f B/cyc/SM, %.1f cyc per warp-load\n", nm, threads, ctas, c[0],
                       bytes_per_iter_per_warp * iters * warps / (double)c[0], (double)c[0] / iters / 4.0);
                cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
            };
            for (int w = 0; w < 2; ++w) k_ld<0><<<148 * ctas, threads>>>(out, cyc, iters);
            rep("ld x16 + wait each", 4 * 2048.0);
            for (int w = 0; w < 2; ++w) k_ld<1><<<148 * ctas, threads>>>(out, cyc, iters);
            rep("ld x32 + wait each", 4 * 2048.0);
            for (int w = 0; w < 2; ++w) k_ld<2><<<148 * ctas, threads>>>(out, cyc, iters);
            rep("4 x ld x16, one wait", 4 * 2048.0);
            for (int w = 0; w < 2; ++w) k_mix<true, false><<<148 * ctas, threads>>>(out, cyc, iters);
            rep("mix: ld + adds", 4 * 2048.0);
            for (int w = 0; w < 2; ++w) k_mix<false, true><<<148 * ctas, threads>>>(out, cyc, iters);
            rep("mix: sqrt only", 4 * 2048.0);
            for (int w = 0; w < 2; ++w) k_mix<true, true><<<148 * ctas, threads>>>(out, cyc, iters);
            rep("mix: ld + sqrt", 4 * 2048.0);
        }
    }
    cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int N : {16, 32, 64, 96, 128}) {
        for (int n : {1, 8, 64}) {
            for (int w = 0; w < 2; ++w) k_mma<<<148, 128, 65536>>>(cyc, n, N);
            cudaDeviceSynchronize();
            cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
            printf("mma M=128 N=%3d K=16 x %2d: issue %lld cyc, to completion %lld cyc (%.1f per MMA)\n", N, n, c[0], c[1], (double)c[1] / n);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
        }
    }
    for (int threads : {128, 512, 1024}) {
        auto rep = [&](const char* nm) {
            cudaDeviceSynchronize();
            cudaMemcpy(c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-26s threads=%4d: %.2f cycles per warp-sqrt per SMSP\n", nm, threads, (double)c[0] / (2000.0 * 8 * (threads / 128.0)));
        };
        for (int w = 0; w < 2; ++w) k_sqrt<0><<<148, threads>>>(out, cyc, 2000); rep("sqrt MUFU");
        for (int w = 0; w < 2; ++w) k_sqrt<1><<<148, threads>>>(out, cyc, 2000); rep("sqrt FMA 1 Newton");
        for (int w = 0; w < 2; ++w) k_sqrt<2><<<148, threads>>>(out, cyc, 2000); rep("sqrt FMA 2 Newton");
        for (int w = 0; w < 2; ++w) k_sqrt<3><<<148, threads>>>(out, cyc, 2000); rep("half MUFU half FMA(2)");
    }
    return 0;
}
