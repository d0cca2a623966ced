// Observables evaluated on the sampled ensemble right after the sampling path (fine-tune objective).
//
// se3_folded_proportion: the f_dRMSD folded-state indicator of observables/folding_stability.py:52-81 --
//   dRMSD_b = sqrt( mean_{i,j} ( |x_bi - x_bj| - |r_i - r_j| )^2 ),   p_b = clamp( sigmoid( k (dRMSD_b - d_0) ), tol, 1 - tol )
// One CTA per sample: the sample's C-alpha coordinates and the reference coordinates are staged in shared memory once
// (24 L bytes), the L^2 pair terms are strided over the threads, block reduction in fp32 with a Kahan-free pairwise tree
// (L^2 <= 10^6 terms of similar size).  Distances are evaluated directly (not through the |a|^2 + |b|^2 - 2ab expansion
// torch.cdist switches to above 25 points), which is the more accurate of the two.
#include "common.cuh"

using namespace se3;

namespace {

__global__ void __launch_bounds__(256)
k_folded_proportion(const float* __restrict__ coords, const float* __restrict__ ref, float* __restrict__ p_out,
                    float* __restrict__ drmsd_out, int L, float k, float d0, float tol) {
    extern __shared__ float sm[];
    float* sx = sm;              // [L][3] sample
    float* sr = sm + 3 * L;      // [L][3] reference
    __shared__ float red[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < 3 * L; i += 256) {
        sx[i] = coords[(int64_t)b * 3 * L + i];
        sr[i] = ref[i];
    }
    __syncthreads();
    float acc = 0.f;
    const int64_t pairs = (int64_t)L * L;
    for (int64_t idx = tid; idx < pairs; idx += 256) {
        const int i = (int)(idx / L), j = (int)(idx - (int64_t)i * L);
        const float ax = sx[3 * i] - sx[3 * j], ay = sx[3 * i + 1] - sx[3 * j + 1], az = sx[3 * i + 2] - sx[3 * j + 2];
        const float bx = sr[3 * i] - sr[3 * j], by = sr[3 * i + 1] - sr[3 * j + 1], bz = sr[3 * i + 2] - sr[3 * j + 2];
        const float d = sqrtf(ax * ax + ay * ay + az * az) - sqrtf(bx * bx + by * by + bz * bz);
        acc += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w];
        const float drmsd = sqrtf(s / (float)pairs);
        float p = 1.0f / (1.0f + expf(-(k * (drmsd - d0))));
        p = fminf(fmaxf(p, tol), 1.0f - tol);
        p_out[b] = p;
        if (drmsd_out) drmsd_out[b] = drmsd;
    }
}

}  // namespace

extern "C" int se3_folded_proportion(const float* coords, const float* ref_coords, float* p_folded, float* drmsd, int64_t batch, int len,
                                     float k, float d_0, float tol, se3_stream_t stream) {
    SE3_REQUIRE(batch >= 0 && len >= 0, "negative size");
    if (batch == 0) return SE3_OK;
    SE3_REQUIRE(coords && ref_coords && p_folded && len > 0, "null pointer or empty sequence");
    const size_t smem = (size_t)len * 6 * sizeof(float);
    SE3_REQUIRE(smem <= 200 * 1024, "sequence too long for the shared-memory staging of this kernel");
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_folded_proportion, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("se3_folded_proportion smem attribute: %s", cudaGetErrorString(e)); return SE3_ECUDA; }
    }
    k_folded_proportion<<<(unsigned)batch, 256, smem, (cudaStream_t)stream>>>(coords, ref_coords, p_folded, drmsd, len, k, d_0, tol);
    count_launch();
    return check_launch("se3_folded_proportion");
}
