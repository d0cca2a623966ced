// Operand packing of the tensor-core IPA operator's SHARED pair tensors (once per sequence and layer), and the row index sets that
// turn the reference's fused projection weight into the operator's head-major records.
//   pair bias  : pair_weight * Linear(x2d)  [L(i)][L(j)][H] fp32   (structure_module.py:179)
//             -> bf16, zero padded, the (head, query tile) slab pass 1 fetches by TMA: L <= 128 query-major [H][L(i)][pitch(j)],
//                longer chains key-major [H][L(j)][round_up(L, 8)(i)]  (common.cuh: ipa_bias_pitch)
//   pair value : Linear(x2d)                [L(i)][L(j)][H*16] fp32 (structure_module.py:209)
//             -> bf16 [L(i)][H][Lp/8][16][8] with Lp = round_up(L, 16): element (i, j, h*16+c) at [i][h][j/8][c][j%8], zero for
//                j >= L -- the K-major UMMA operand of pass 2, one contiguous block per (query, head)
// Byte movement + one rounding: the results are bit-identical to torch's permute / pad / .to(bfloat16).
#include <cuda_bf16.h>

#include "common.cuh"

using namespace se3;

namespace {

__global__ void __launch_bounds__(256) k_pack_pair_bias(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int L, int H, int Lpi) {
    // one thread per output element, the padded index fastest (coalesced 2-byte stores; the reads of one warp touch 32 rows of H floats)
    const int64_t n = (int64_t)H * L * Lpi;
    const bool qmajor = ipa_bias_query_major(L);
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int fast = (int)(o % Lpi);
        const int64_t r = o / Lpi;
        const int slow = (int)(r % L), h = (int)(r / L);
        const int i = qmajor ? slow : fast, j = qmajor ? fast : slow;
        out[o] = __float2bfloat16_rn(fast < L ? in[((int64_t)i * L + j) * H + h] : 0.f);
    }
}

__global__ void __launch_bounds__(256) k_pack_pair_value(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int L, int H, int Lp) {
    // one thread per 16-byte output chunk = 8 consecutive keys of one channel: [i][h][j/8][c][0..8)
    const int64_t n = (int64_t)L * H * (Lp / 8) * 16;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(o & 15);
        int64_t r = o >> 4;
        const int jc = (int)(r % (Lp / 8));
        r /= (Lp / 8);
        const int h = (int)(r % H), i = (int)(r / H);
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j0 = jc * 8 + 2 * u;
            const float a = j0 < L ? in[((int64_t)i * L + j0) * (H * 16) + h * 16 + c] : 0.f;
            const float b = j0 + 1 < L ? in[((int64_t)i * L + j0 + 1) * (H * 16) + h * 16 + c] : 0.f;
            const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
            w[u] = *reinterpret_cast<const uint32_t*>(&v);
        }
        reinterpret_cast<uint4*>(out)[o] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace

extern "C" {

int64_t se3_ipa_tc_packed_pair_bytes(int len, int heads, int64_t* bias_bytes, int64_t* value_bytes) {
    if (len <= 0 || heads <= 0) return SE3_EINVAL;
    const int64_t lpi = ipa_bias_pitch(len), lp = (len + 15) / 16 * 16;
    const int64_t bb = (int64_t)heads * len * lpi * 2, vb = (int64_t)len * heads * lp * 16 * 2;
    if (bias_bytes) *bias_bytes = bb;
    if (value_bytes) *value_bytes = vb;
    return bb + vb;
}

int se3_ipa_tc_pack_pair(const float* pair_bias, const float* pair_value, void* bias_packed, void* value_packed, int len, int heads,
                         se3_stream_t stream) {
    SE3_REQUIRE(len > 0 && len <= 512 && heads > 0, "1 <= len <= 512, heads >= 1");
    SE3_REQUIRE((pair_bias == nullptr) == (bias_packed == nullptr) && (pair_value == nullptr) == (value_packed == nullptr),
                "each input needs its output (pass both NULL to skip one of the two packs)");
    SE3_REQUIRE(value_packed == nullptr || (reinterpret_cast<uintptr_t>(value_packed) & 15) == 0, "value_packed must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int lpi = ipa_bias_pitch(len), lp = (len + 15) / 16 * 16;
    if (pair_bias) {
        const int64_t n = (int64_t)heads * len * lpi;
        k_pack_pair_bias<<<(unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16), 256, 0, st>>>(pair_bias, (__nv_bfloat16*)bias_packed, len, heads, lpi);
        count_launch();
        if (int rc = check_launch("se3_ipa_tc_pack_pair(bias)")) return rc;
    }
    if (pair_value) {
        const int64_t n = (int64_t)len * heads * (lp / 8) * 16;
        k_pack_pair_value<<<(unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16), 256, 0, st>>>(pair_value, (__nv_bfloat16*)value_packed, len, heads, lp);
        count_launch();
        if (int rc = check_launch("se3_ipa_tc_pack_pair(value)")) return rc;
    }
    return SE3_OK;
}

int se3_ipa_split_perm(int heads, int dk, int32_t* h_scalar_rows, int32_t* h_point_rows, int32_t* h_q_positions) {
    SE3_REQUIRE(heads > 0 && dk > 0 && h_scalar_rows && h_point_rows, "heads, dk >= 1 and both row arrays");
    const int hd = heads * dk;
    for (int h = 0; h < heads; ++h) {
        for (int c = 0; c < dk; ++c) {
            h_scalar_rows[h * 3 * dk + c] = h * dk + c;                     // q
            h_scalar_rows[h * 3 * dk + dk + c] = hd + h * dk + c;           // k
            h_scalar_rows[h * 3 * dk + 2 * dk + c] = 2 * hd + h * dk + c;   // v
            if (h_q_positions) h_q_positions[h * dk + c] = h * 3 * dk + c;
        }
        for (int c = 0; c < 12; ++c) {
            h_point_rows[h * 48 + c] = 3 * hd + h * 12 + c;                 // query points
            h_point_rows[h * 48 + 12 + c] = 3 * hd + 12 * heads + h * 12 + c;   // key points
        }
        for (int c = 0; c < 24; ++c) h_point_rows[h * 48 + 24 + c] = 3 * hd + 24 * heads + h * 24 + c;   // value points
    }
    return SE3_OK;
}

}  // extern "C"
