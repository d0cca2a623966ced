// Helpers shared by the editions of the tensor-core IPA operator (ipa_tc.cu: one item per CTA / persistent wide / cluster-split;
// ipa_tc_pp.cu: warp-specialised ping-pong edition for L <= 128).
#pragma once
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace se3 {
namespace ipa_tc {

constexpr int DK = 16, PQ = 4, PV = 8;
constexpr int NV = 80;   // accumulator columns of the second product: v 16 | v_pt hi 24 | v_pt lo 24 | ones 1 | 15 x zero
constexpr int NVP = 64;  // ... of which the point operand (hi | lo | ones | zero) is a separate N = 64 MMA
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// N contiguous outputs (N % 8 == 0, destination 16-byte aligned) as 128-bit stores
template <int N> __device__ __forceinline__ void store_vec(float* dst, const float (&v)[N]) {
#pragma unroll
    for (int c = 0; c < N / 4; ++c) reinterpret_cast<float4*>(dst)[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
template <int N> __device__ __forceinline__ void store_vec(__nv_bfloat16* dst, const float (&v)[N]) {
#pragma unroll
    for (int c = 0; c < N / 8; ++c)
        reinterpret_cast<uint4*>(dst)[c] = make_uint4(tc::pack_bf16(v[8 * c], v[8 * c + 1]), tc::pack_bf16(v[8 * c + 2], v[8 * c + 3]),
                                                      tc::pack_bf16(v[8 * c + 4], v[8 * c + 5]), tc::pack_bf16(v[8 * c + 6], v[8 * c + 7]));
}

// N consecutive coordinates (N % 4 == 0) of a point record starting at element `first` (a multiple of 4): fp32 records are read
// as float4, bf16 records (4 values per 8-byte load) are widened
template <bool kBf16, int N>
__device__ __forceinline__ void load_coords(const void* row, int first, float (&l)[N]) {
    if constexpr (kBf16) {
        const uint2* p = reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(row) + first * 2);
#pragma unroll
        for (int c = 0; c < N / 4; ++c) {
            const uint2 v = p[c];
            l[4 * c] = __uint_as_float(v.x << 16); l[4 * c + 1] = __uint_as_float(v.x & 0xffff0000u);
            l[4 * c + 2] = __uint_as_float(v.y << 16); l[4 * c + 3] = __uint_as_float(v.y & 0xffff0000u);
        }
    } else {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + first);
#pragma unroll
        for (int c = 0; c < N / 4; ++c) {
            const float4 v = p[c];
            l[4 * c] = v.x; l[4 * c + 1] = v.y; l[4 * c + 2] = v.z; l[4 * c + 3] = v.w;
        }
    }
}

// global = R.local + T for one point
__device__ __forceinline__ void to_global(const float (&R)[9], const float (&T)[3], float x, float y, float z, float& gx, float& gy, float& gz) {
    // three FMAs per component, the translation inside the chain (one instruction fewer than (R.x) + T; the extra roundings at the
    // translation's magnitude are 2^-24 relative, far below the 2^-9 of the bf16 point records)
    gx = fmaf(R[0], x, fmaf(R[1], y, fmaf(R[2], z, T[0])));
    gy = fmaf(R[3], x, fmaf(R[4], y, fmaf(R[5], z, T[1])));
    gz = fmaf(R[6], x, fmaf(R[7], y, fmaf(R[8], z, T[2])));
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// row-major [rows][cols] matrix with a row pitch; box = box_cols x box_rows elements, dense in shared memory
inline int make_map_2d(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_elems,
                uint32_t box_cols, uint32_t box_rows, const char* what, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_NONE) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return SE3_ECUDA; }
    const cuuint64_t dims[2] = {cols, rows}, strides[1] = {pitch_elems * (uint64_t)elem_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
    const CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tensor map for %s: cuTensorMapEncodeTiled failed with %d", what, (int)r); return SE3_ECUDA; }
    return SE3_OK;
}


// Arguments of one pass-1 launch, edition independent (ipa_tc.cu fills them in)
struct Pass1Args {
    const __nv_bfloat16* scal; int scal_stride;            // bf16 scalar records [rows][>= H*48]: q16 | k16 | v16 per head
    const void* pts; int pts_stride; bool pts_bf16;        // point records [rows][>= H*48]: qp12 | kp12 | vp24 per head
    const float *rot, *trans;                              // frames [B*L][9], [B*L][3]
    const __nv_bfloat16* pair_bias;                        // bf16 [H][L(j)][round_up(L,8)(i)]
    const float *key_bias, *head_weight;                   // [B][L] or null; [H]
    void* out; bool out_bf16;                              // concat layout [B*L][H*(2*dk + 4*pv)]
    __nv_bfloat16* pbuf; float* inv_sum;                   // probability workspace [H][L][Bpad][Lp] and 1/rowsum [H][L][Bpad]
    se3_ipa_shape sh; int Lp, Bpad;                        // Lp = L rounded up to 16
    cudaStream_t stream;
};
// ping-pong edition (ipa_tc_pp.cu).  Returns SE3_EUNSUPPORTED without launching when the shape does not fit it.
int launch_pass1_pingpong(const Pass1Args& a);

}  // namespace ipa_tc
}  // namespace se3
