// Row-wise fused glue of the score network (bf16 throughput mode):
//   x += y + bias                      (residual update, structure_module.py:247-248: x1d = x1d + attn(...)/ffn(...))
//   out = LayerNorm(x) * gamma + beta  (the pre-norm of the NEXT block: norm1 / norm2 / diff-head norms)
// in one pass: one warp per row, the row lives in registers (D <= 1024), two-pass mean/variance, 128-bit
// accesses.  Replaces 4-5 ATen launches (broadcast bias add, residual add, layer_norm, dtype cast) and
// their ~7 passes over the [N, D] activations by one kernel that reads x, y once and writes x, out once.
#include "common.cuh"
#include "tc_common.cuh"

using namespace se3;

namespace {

template <int VPL, typename OutT>  // VPL float4 per lane: D = 128 * VPL
__global__ void __launch_bounds__(256)
k_residual_layernorm(float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ bias,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps, OutT* __restrict__ out,
                     int64_t rows) {
    constexpr int D = 128 * VPL;
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    float4 v[VPL];
    float4* xr = reinterpret_cast<float4*>(x + row * D);
#pragma unroll
    for (int k = 0; k < VPL; ++k) v[k] = xr[k * 32 + lane];
    if (y != nullptr) {
        const float4* yr = reinterpret_cast<const float4*>(y + row * D);
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const float4 a = __ldg(yr + k * 32 + lane);
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(bias) + k * 32 + lane);
            v[k].x += a.x + b4.x; v[k].y += a.y + b4.y; v[k].z += a.z + b4.z; v[k].w += a.w + b4.w;
            xr[k * 32 + lane] = v[k];
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / D) + eps);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + k * 32 + lane);
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + k * 32 + lane);
        const float o0 = (v[k].x - mean) * rstd * g.x + bt.x, o1 = (v[k].y - mean) * rstd * g.y + bt.y;
        const float o2 = (v[k].z - mean) * rstd * g.z + bt.z, o3 = (v[k].w - mean) * rstd * g.w + bt.w;
        if constexpr (sizeof(OutT) == 2) {
            reinterpret_cast<uint2*>(out + row * D)[k * 32 + lane] = make_uint2(tc::pack_bf16(o0, o1), tc::pack_bf16(o2, o3));
        } else {
            reinterpret_cast<float4*>(out + row * D)[k * 32 + lane] = make_float4(o0, o1, o2, o3);
        }
    }
}

template <typename OutT>
int launch(float* x, const float* y, const float* bias, const float* gamma, const float* beta, float eps, OutT* out, int64_t rows,
           int dim, cudaStream_t st) {
    const unsigned grid = (unsigned)((rows * 32 + 255) / 256);
    switch (dim / 128) {
        case 1: k_residual_layernorm<1, OutT><<<grid, 256, 0, st>>>(x, y, bias, gamma, beta, eps, out, rows); break;
        case 2: k_residual_layernorm<2, OutT><<<grid, 256, 0, st>>>(x, y, bias, gamma, beta, eps, out, rows); break;
        case 4: k_residual_layernorm<4, OutT><<<grid, 256, 0, st>>>(x, y, bias, gamma, beta, eps, out, rows); break;
        case 8: k_residual_layernorm<8, OutT><<<grid, 256, 0, st>>>(x, y, bias, gamma, beta, eps, out, rows); break;
        default: set_error("se3_residual_layernorm: dim must be 128, 256, 512 or 1024 (got %d)", dim); return SE3_EUNSUPPORTED;
    }
    count_launch();
    return check_launch("se3_residual_layernorm");
}

}  // namespace

extern "C" int se3_residual_layernorm(float* x, const float* y, const float* bias, const float* gamma, const float* beta, float eps,
                                      void* out, int out_is_bf16, int64_t rows, int dim, se3_stream_t stream) {
    SE3_REQUIRE(rows >= 0, "negative rows");
    if (rows == 0) return SE3_OK;
    SE3_REQUIRE(x && gamma && beta && out, "null pointer");
    SE3_REQUIRE(dim % 128 == 0, "dim must be a multiple of 128");
    SE3_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias) |
                  reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "pointers must be 16-byte aligned");
    if (out_is_bf16) return launch<__nv_bfloat16>(x, y, bias, gamma, beta, eps, (__nv_bfloat16*)out, rows, dim, (cudaStream_t)stream);
    return launch<float>(x, y, bias, gamma, beta, eps, (float*)out, rows, dim, (cudaStream_t)stream);
}
