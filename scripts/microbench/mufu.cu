// MUFU / FFMA2 throughput and latency probe (developer diagnostics; build: nvcc -arch=sm_100a -O3 mufu.cu -o mufu)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float fsqrt(float x) { float r; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float frsq(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <int OP, int ILP>
__global__ void k(float* out, long long* cyc, int iters) {
    float v[ILP];
    for (int i = 0; i < ILP; ++i) v[i] = 1.5f + threadIdx.x * 0.001f + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) v[i] = fsqrt(v[i]);
            else if (OP == 1) v[i] = frsq(v[i]);
            else if (OP == 2) v[i] = fex2(v[i]);
            else if (OP == 3) v[i] = fmaf(v[i], 1.0001f, 0.5f);

        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP, int ILP>
__global__ void k2(float* out, long long* cyc, int iters) {
    float2 v[ILP];
    for (int i = 0; i < ILP; ++i) v[i] = make_float2(1.5f + threadIdx.x * 0.001f + i, 0.5f + i);
    const float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(0.5f, 0.25f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) v[i] = __ffma2_rn(v[i], a, b);
            else if (OP == 1) v[i] = __fadd2_rn(v[i], b);
            else if (OP == 2) v[i] = __fmul2_rn(v[i], a);
            else if (OP == 3) { v[i].x = fmaf(v[i].x, a.x, b.x); v[i].y = fmaf(v[i].y, a.y, b.y); }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; ++i) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP, int ILP> void run2(const char* name, int threads) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k2<OP, ILP><<<148, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    k2<OP, ILP><<<148, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_per_smsp = threads / 32 / 4.0;
    printf("%-10s ILP=%d threads=%4d: %.2f cycles per packed warp-instruction per SMSP; dependent-chain step %.2f cycles\n", name, ILP, threads,
           (double)c / ((double)iters * ILP * warps_per_smsp), (double)c / ((double)iters));
    cudaFree(out); cudaFree(cyc);
}
template <int OP, int ILP> void run(const char* name, int threads) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<OP, ILP><<<148, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    k<OP, ILP><<<148, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_per_smsp = threads / 32 / 4.0;
    printf("%-10s ILP=%d threads=%4d: %.2f cycles per warp-instruction per SMSP (%.2f cycles/iter/thread-op)\n", name, ILP, threads,
           (double)c / ((double)iters * ILP * warps_per_smsp), (double)c / ((double)iters * ILP));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0, 1>("sqrt", 128); run<0, 8>("sqrt", 128); run<0, 8>("sqrt", 512); run<0, 8>("sqrt", 1024);
    run<1, 1>("rsqrt", 128); run<1, 8>("rsqrt", 128); run<1, 8>("rsqrt", 1024);
    run<2, 1>("ex2", 128); run<2, 8>("ex2", 128); run<2, 8>("ex2", 1024);
    run<3, 1>("ffma", 128); run<3, 8>("ffma", 128); run<3, 8>("ffma", 1024);
    run2<0, 1>("ffma2", 128); run2<0, 2>("ffma2", 128); run2<0, 4>("ffma2", 128); run2<0, 8>("ffma2", 128); run2<0, 8>("ffma2", 1024);
    run2<1, 1>("fadd2", 128); run2<1, 8>("fadd2", 128); run2<1, 8>("fadd2", 1024);
    run2<2, 1>("fmul2", 128); run2<2, 8>("fmul2", 1024);
    run2<3, 1>("2x ffma", 128); run2<3, 8>("2x ffma", 128); run2<3, 8>("2x ffma", 1024);
    return 0;
}
