// Per-sequence pair precompute (SURVEY.md 8b `pair_precompute`; models.py:243-293 feeding structure_module.py:179, 209), fp32:
//   se3_pair_embed   : x2d[r, :] = LayerNorm(pair_embeds[r, :]) . W_x2d^T + relpos_table[bucket[r % L^2], :]        rows r of [Bp*L*L]
//   se3_pair_project : per layer  bias = pair_weight * x2d . W_bias^T  and  value = x2d . W_value^T, written straight in the layout the
//                      attention kernel of choice reads (fp32 [Bp,H,L,L] + [Bp,L,L,H*dk] for se3_ipa_attention_fwd, or the bf16 TMA slab /
//                      UMMA operand packs of se3_ipa_attention_tc_fwd -- the same bytes se3_ipa_tc_pack_pair produces)
// Once per sequence (the reference recomputes all of it B times per layer per score evaluation), so these are plain fp32 SIMT tile
// GEMMs (64 x 64 tiles, 4 x 4 outputs per thread, k ascending): 16 GFLOP per sequence at L = 84, a few milliseconds.
#include <cuda_bf16.h>

#include "common.cuh"

using namespace se3;

namespace {

constexpr int TM = 64, TN = 64, TK = 32;

// C[m, n] = sum_k A[m, k] * B[n, k]: one 64 x 64 tile per CTA (256 threads, 4 x 4 per thread).  `a_row(m, k)` / `emit(m, n, acc)` are
// supplied by the caller; K % TK == 0.
template <typename LoadA, typename Emit>
__device__ __forceinline__ void tile_gemm(int M, int N, int K, const float* __restrict__ Bw, LoadA a_at, Emit emit) {
    __shared__ float sA[TK][TM + 1], sB[TK][TN + 1];
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += TK) {
        for (int e = threadIdx.x; e < TM * TK; e += 256) {
            const int r = e / TK, k = e % TK;
            sA[k][r] = (m0 + r < M) ? a_at(m0 + r, k0 + k) : 0.f;
            sB[k][r] = (n0 + r < N) ? Bw[(int64_t)(n0 + r) * K + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < TK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; b[i] = sB[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) emit(m, n, acc[i][j]);
        }
}

// LayerNorm statistics per row (mean, 1/sqrt(var + eps)), one warp per row
__global__ void __launch_bounds__(256) k_row_stats(const float* __restrict__ x, float2* __restrict__ stats, int64_t rows, int dim, float eps) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    float s = 0.f;
    for (int k = lane; k < dim; k += 32) s += x[r * dim + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)dim;
    float v = 0.f;
    for (int k = lane; k < dim; k += 32) { const float d = x[r * dim + k] - mean; v = fmaf(d, d, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) stats[r] = make_float2(mean, rsqrtf(v / (float)dim + eps));
}

__global__ void __launch_bounds__(256)
k_pair_embed(const float* __restrict__ pair, const float2* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
             const float* __restrict__ w, const float* __restrict__ relpos, const int32_t* __restrict__ bucket, float* __restrict__ x2d, int64_t rows,
             int ll, int de, int dp) {
    tile_gemm((int)rows, dp, de, w,
              [&](int m, int k) { const float2 st = stats[m]; return fmaf((pair[(int64_t)m * de + k] - st.x) * st.y, gamma[k], beta[k]); },
              [&](int m, int n, float acc) { x2d[(int64_t)m * dp + n] = acc + relpos[(int64_t)bucket[m % ll] * dp + n]; });
}

// n < H: bias channel (scaled by pair_weight); n >= H: value channel n - H = h * dk + c
template <bool kPacked>
__global__ void __launch_bounds__(256)
k_pair_project(const float* __restrict__ x2d, const float* __restrict__ w_cat, float pair_weight, void* __restrict__ bias_out, void* __restrict__ value_out,
               int64_t rows, int L, int H, int dk, int dp) {
    const int lpi = ipa_bias_pitch(L), lp = (L + 15) & ~15;
    const bool qmajor = ipa_bias_query_major(L);
    tile_gemm((int)rows, H + H * dk, dp, w_cat, [&](int m, int k) { return x2d[(int64_t)m * dp + k]; },
              [&](int m, int n, float acc) {
                  const int b = m / (L * L), ij = m - b * L * L, i = ij / L, j = ij - i * L;
                  if (n < H) {
                      const float v = pair_weight * acc;
                      if (kPacked) reinterpret_cast<__nv_bfloat16*>(bias_out)[((int64_t)n * L + (qmajor ? i : j)) * lpi + (qmajor ? j : i)] = __float2bfloat16_rn(v);   // [H][i][j] (L <= 128) or [H][j][i]
                      else reinterpret_cast<float*>(bias_out)[(((int64_t)b * H + n) * L + i) * L + j] = v;                                         // [Bp][H][i][j]
                  } else {
                      const int hc = n - H, h = hc / dk, c = hc - h * dk;
                      if (kPacked)                                                                                                                 // [i][h][j/8][c][j%8]
                          reinterpret_cast<__nv_bfloat16*>(value_out)[((((int64_t)i * H + h) * (lp / 8) + (j >> 3)) * 16 + c) * 8 + (j & 7)] = __float2bfloat16_rn(acc);
                      else reinterpret_cast<float*>(value_out)[(int64_t)m * (H * dk) + hc] = acc;                                                 // [Bp][i][j][H*dk]
                  }
              });
}

}  // namespace

extern "C" {

int se3_pair_embed(const float* pair_embeds, const float* ln_gamma, const float* ln_beta, float ln_eps, const float* w_x2d, const float* relpos_table,
                   const int32_t* bucket, float* x2d, float* stats_workspace, int64_t pair_batch, int len, int dim_embed, int dim_pair, se3_stream_t stream) {
    SE3_REQUIRE(pair_embeds && ln_gamma && ln_beta && w_x2d && relpos_table && bucket && x2d && stats_workspace, "null pointer");
    SE3_REQUIRE(pair_batch >= 1 && len >= 1 && dim_embed >= 1 && dim_pair >= 1 && dim_embed % TK == 0, "bad shape (dim_embed must be a multiple of 32)");
    const int64_t rows = pair_batch * len * len;
    SE3_REQUIRE(rows < (1ll << 31), "pair_batch * len^2 must fit 31 bits");
    cudaStream_t st = (cudaStream_t)stream;
    k_row_stats<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(pair_embeds, reinterpret_cast<float2*>(stats_workspace), rows, dim_embed, ln_eps);
    count_launch();
    if (int rc = check_launch("se3_pair_embed(stats)")) return rc;
    k_pair_embed<<<dim3((unsigned)((rows + TM - 1) / TM), (unsigned)((dim_pair + TN - 1) / TN)), 256, 0, st>>>(
        pair_embeds, reinterpret_cast<const float2*>(stats_workspace), ln_gamma, ln_beta, w_x2d, relpos_table, bucket, x2d, rows, len * len, dim_embed, dim_pair);
    count_launch();
    return check_launch("se3_pair_embed");
}

int se3_pair_project(const float* x2d, const float* w_bias_value, float pair_weight, void* bias_out, void* value_out, int packed, int64_t pair_batch, int len,
                     int heads, int dk, int dim_pair, se3_stream_t stream) {
    SE3_REQUIRE(x2d && w_bias_value && bias_out && value_out, "null pointer");
    SE3_REQUIRE(pair_batch >= 1 && len >= 1 && heads >= 1 && dk >= 1 && dim_pair % TK == 0, "bad shape (dim_pair must be a multiple of 32)");
    SE3_REQUIRE(!packed || (pair_batch == 1 && dk == 16 && len <= 512), "the packed layouts are those of se3_ipa_attention_tc_fwd: shared pair tensors, dk = 16, L <= 512");
    const int64_t rows = pair_batch * len * len;
    SE3_REQUIRE(rows < (1ll << 31), "pair_batch * len^2 must fit 31 bits");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((unsigned)((rows + TM - 1) / TM), (unsigned)((heads + heads * dk + TN - 1) / TN));
    if (packed) {
        // padding (queries >= L of the bias slabs, keys >= L of the value operand) must read as zero
        const int64_t lpi = ipa_bias_pitch(len), lp = (len + 15) / 16 * 16;
        if (lpi != len && cudaMemsetAsync(bias_out, 0, (size_t)heads * len * lpi * 2, st) != cudaSuccess) { set_error("se3_pair_project: memset"); return SE3_ECUDA; }
        if (lp != len && cudaMemsetAsync(value_out, 0, (size_t)len * heads * lp * 16 * 2, st) != cudaSuccess) { set_error("se3_pair_project: memset"); return SE3_ECUDA; }
        k_pair_project<true><<<grid, 256, 0, st>>>(x2d, w_bias_value, pair_weight, bias_out, value_out, rows, len, heads, dk, dim_pair);
    } else {
        k_pair_project<false><<<grid, 256, 0, st>>>(x2d, w_bias_value, pair_weight, bias_out, value_out, rows, len, heads, dk, dim_pair);
    }
    count_launch();
    return check_launch("se3_pair_project");
}

}  // extern "C"
