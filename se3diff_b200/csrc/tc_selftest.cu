// tcgen05 self-test: D[128 x N] = A[128 x K] . B[N x K]^T (bf16 in, fp32 out) through the same
// operand staging, descriptors, TMEM and mbarrier helpers the attention kernels use.
#include "common.cuh"
#include "tc_common.cuh"

using namespace se3;

namespace {

__global__ void __launch_bounds__(128) k_umma_selftest(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                       float* __restrict__ d, int N, int K) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    uint8_t* sA = smem;                       // [K/8][128][16 B]
    uint8_t* sB = smem + (size_t)K * 128 * 2;  // [K/8][N][16 B]
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base_slot, 256);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    // stage operands: 16-byte pieces (8 consecutive k) of row-major global rows
    for (int idx = tid; idx < 128 * (K / 8); idx += 128) {
        const int r = idx % 128, kc = idx / 128;
        *reinterpret_cast<uint4*>(sA + ((size_t)kc * 128 + r) * 16) = *reinterpret_cast<const uint4*>(a + (size_t)r * K + kc * 8);
    }
    for (int idx = tid; idx < N * (K / 8); idx += 128) {
        const int r = idx % N, kc = idx / N;
        *reinterpret_cast<uint4*>(sB + ((size_t)kc * N + r) * 16) = *reinterpret_cast<const uint4*>(b + (size_t)r * K + kc * 8);
    }
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = tmem_base_slot;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(128, N);
        for (int ks = 0; ks < K / 16; ++ks)
            tc::mma_bf16(tmem, tc::make_desc_kstep(tc::smem_u32(sA), 128, ks), tc::make_desc_kstep(tc::smem_u32(sB), N, ks), idesc, ks > 0);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tc::tmem_ld16(tc::tmem_addr(tmem, warp * 32, c0), r);
        tc::tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; ++c) d[(size_t)tid * N + c0 + c] = __uint_as_float(r[c]);
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace

extern "C" int se3_debug_umma_gemm(const void* a_bf16, const void* b_bf16, float* d, int n, int k, se3_stream_t stream) {
    SE3_REQUIRE(a_bf16 && b_bf16 && d, "null pointer");
    SE3_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && k >= 16 && k % 16 == 0, "need 16 <= N <= 256, N % 16 == 0, K % 16 == 0");
    const size_t smem = (size_t)k * (128 + n) * 2;
    SE3_REQUIRE(smem <= 200 * 1024, "tile too large for shared memory");
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("selftest smem attribute: %s", cudaGetErrorString(e)); return SE3_ECUDA; }
    }
    k_umma_selftest<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)b_bf16, d, n, k);
    count_launch();
    return check_launch("se3_debug_umma_gemm");
}
