// Backbone reconstruction and the physicality filter that follow the sampler on the wall-clock of sample.sh
// (convert_chemgraph.py:139-293 and :296-395).
//
// se3_backbone_atoms: N, CA, C, CB, O of every residue from its frame.  The reference goes through openfold's
//   torsion-angle machinery with zero torsions (compute_backbone) -- for the five backbone atoms that reduces to
//   x = R . ideal_local(aatype) + T for N / CA / C / CB (rigid group 0; glycine has no CB: exact zero, masked out) --
//   and then imputes O from the neighbouring frame (_adjust_oxygen_pos): in the CA-C-N(next) plane, 1.23 A from C, pointing
//   away from the triangle; at the C-terminus (or next residue unknown) from CA->C and CA->N of the same residue.
//   One thread per (sample, residue); frames are read once, the neighbour's N is recomputed (9 flops) instead of exchanged.
// se3_physicality: per sample max sequential CA-CA distance, max sequential C-N distance, min heavy-atom distance between
//   residues at least 3 apart (mdtraj.compute_contacts default: contacts='all', scheme='closest-heavy').  One CTA per sample,
//   atoms staged in shared memory, pair terms strided over threads, block max / min reductions.
#include <math_constants.h>

#include "common.cuh"

using namespace se3;

namespace {

// ideal backbone geometry per residue type, openfold order ARNDCQEGHILKMFPSTWYV: local N, CA, C, CB (Angstrom)
// (openfold/np/residue_constants.py: rigid_group_atom_positions, rigid group 0 entries)
__constant__ float kLocal[20][4][3] = {
    {{-0.525f, 1.363f, 0.f}, {0.f, 0.f, 0.f}, {1.526f, 0.f, 0.f}, {-0.529f, -0.774f, -1.205f}},
    {{-0.524f, 1.362f, 0.f}, {0.f, 0.f, 0.f}, {1.525f, 0.f, 0.f}, {-0.524f, -0.778f, -1.209f}},
    {{-0.536f, 1.357f, 0.f}, {0.f, 0.f, 0.f}, {1.526f, 0.f, 0.f}, {-0.531f, -0.787f, -1.200f}},
    {{-0.525f, 1.362f, 0.f}, {0.f, 0.f, 0.f}, {1.527f, 0.f, 0.f}, {-0.526f, -0.778f, -1.208f}},
    {{-0.522f, 1.362f, 0.f}, {0.f, 0.f, 0.f}, {1.524f, 0.f, 0.f}, {-0.519f, -0.773f, -1.212f}},
    {{-0.526f, 1.361f, 0.f}, {0.f, 0.f, 0.f}, {1.526f, 0.f, 0.f}, {-0.525f, -0.779f, -1.207f}},
    {{-0.528f, 1.361f, 0.f}, {0.f, 0.f, 0.f}, {1.526f, 0.f, 0.f}, {-0.526f, -0.781f, -1.207f}},
    {{-0.572f, 1.337f, 0.f}, {0.f, 0.f, 0.f}, {1.517f, 0.f, 0.f}, {0.f, 0.f, 0.f}},
    {{-0.527f, 1.360f, 0.f}, {0.f, 0.f, 0.f}, {1.525f, 0.f, 0.f}, {-0.525f, -0.778f, -1.208f}},
    {{-0.493f, 1.373f, 0.f}, {0.f, 0.f, 0.f}, {1.527f, 0.f, 0.f}, {-0.536f, -0.793f, -1.213f}},
    {{-0.520f, 1.363f, 0.f}, {0.f, 0.f, 0.f}, {1.525f, 0.f, 0.f}, {-0.522f, -0.773f, -1.214f}},
    {{-0.526f, 1.362f, 0.f}, {0.f, 0.f, 0.f}, {1.526f, 0.f, 0.f}, {-0.524f, -0.778f, -1.208f}},
    {{-0.521f, 1.364f, 0.f}, {0.f, 0.f, 0.f}, {1.525f, 0.f, 0.f}, {-0.523f, -0.776f, -1.210f}},
    {{-0.518f, 1.363f, 0.f}, {0.f, 0.f, 0.f}, {1.524f, 0.f, 0.f}, {-0.525f, -0.776f, -1.212f}},
    {{-0.566f, 1.351f, 0.f}, {0.f, 0.f, 0.f}, {1.527f, 0.f, 0.f}, {-0.546f, -0.611f, -1.293f}},
    {{-0.529f, 1.360f, 0.f}, {0.f, 0.f, 0.f}, {1.525f, 0.f, 0.f}, {-0.518f, -0.777f, -1.211f}},
    {{-0.517f, 1.364f, 0.f}, {0.f, 0.f, 0.f}, {1.526f, 0.f, 0.f}, {-0.516f, -0.793f, -1.215f}},
    {{-0.521f, 1.363f, 0.f}, {0.f, 0.f, 0.f}, {1.525f, 0.f, 0.f}, {-0.523f, -0.776f, -1.212f}},
    {{-0.522f, 1.362f, 0.f}, {0.f, 0.f, 0.f}, {1.524f, 0.f, 0.f}, {-0.522f, -0.776f, -1.213f}},
    {{-0.494f, 1.373f, 0.f}, {0.f, 0.f, 0.f}, {1.527f, 0.f, 0.f}, {-0.533f, -0.795f, -1.213f}},
};
constexpr int kGly = 7;
constexpr float kCOBond = 1.23f;  // convert_chemgraph.py:16

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 unit_eps(V3 a) {   // a / (|a| + 1e-7), convert_chemgraph.py:213-230
    const float n = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z) + 1e-7f;
    return {a.x / n, a.y / n, a.z / n};
}
__device__ __forceinline__ V3 place(const float* R, const float* T, const float* l) {
    return {R[0] * l[0] + R[1] * l[1] + R[2] * l[2] + T[0], R[3] * l[0] + R[4] * l[1] + R[5] * l[2] + T[1],
            R[6] * l[0] + R[7] * l[1] + R[8] * l[2] + T[2]};
}

__global__ void __launch_bounds__(256)
k_backbone(const float* __restrict__ pos, const float* __restrict__ rot, const int* __restrict__ aatype, const uint8_t* __restrict__ known,
           float* __restrict__ out, int64_t n_total, int L) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n_total) return;
    const int i = (int)(idx % L);
    const int aa = aatype[i];
    float R[9], T[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = rot[idx * 9 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) T[k] = pos[idx * 3 + k];
    const V3 n = place(R, T, kLocal[aa][0]), ca = place(R, T, kLocal[aa][1]), c = place(R, T, kLocal[aa][2]);
    V3 cb = place(R, T, kLocal[aa][3]);
    if (aa == kGly) cb = {0.f, 0.f, 0.f};
    const bool next_gone = (i == L - 1) || (known && !known[i + 1]);
    V3 dir;
    if (!next_gone) {
        float Rn[9], Tn[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) Rn[k] = rot[(idx + 1) * 9 + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) Tn[k] = pos[(idx + 1) * 3 + k];
        const V3 n_next = place(Rn, Tn, kLocal[aatype[i + 1]][0]);
        dir = unit_eps(add(unit_eps(sub(c, ca)), unit_eps(sub(c, n_next))));
    } else {
        dir = unit_eps(add(unit_eps(sub(c, ca)), unit_eps(sub(n, ca))));
    }
    const V3 o = {c.x + dir.x * kCOBond, c.y + dir.y * kCOBond, c.z + dir.z * kCOBond};
    float* dst = out + idx * 15;
    const V3 atoms[5] = {n, ca, c, cb, o};
#pragma unroll
    for (int a = 0; a < 5; ++a) { dst[a * 3] = atoms[a].x; dst[a * 3 + 1] = atoms[a].y; dst[a * 3 + 2] = atoms[a].z; }
}

__global__ void __launch_bounds__(256)
k_physicality(const float* __restrict__ atoms, const int* __restrict__ aatype, float* __restrict__ out, int L) {
    extern __shared__ float sa[];   // [L][5][3]
    __shared__ float red[3][8];
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < L * 15; i += 256) sa[i] = atoms[(int64_t)b * L * 15 + i];
    __syncthreads();
    float ca_max = 0.f, cn_max = 0.f, heavy_min = CUDART_INF_F;
    for (int i = tid; i < L - 1; i += 256) {
        const float* a = sa + i * 15;
        const float* n = sa + (i + 1) * 15;
        const float dx = a[3] - n[3], dy = a[4] - n[4], dz = a[5] - n[5];           // CA_i - CA_{i+1}
        ca_max = fmaxf(ca_max, sqrtf(dx * dx + dy * dy + dz * dz));
        const float ex = a[6] - n[0], ey = a[7] - n[1], ez = a[8] - n[2];           // C_i - N_{i+1}
        cn_max = fmaxf(cn_max, sqrtf(ex * ex + ey * ey + ez * ez));
    }
    const int64_t pairs = (int64_t)L * L;
    for (int64_t idx = tid; idx < pairs; idx += 256) {
        const int i = (int)(idx / L), j = (int)(idx - (int64_t)i * L);
        if (j - i < 3) continue;                                                     // residues at least 3 apart, each pair once
        const bool cb_i = aatype[i] != kGly, cb_j = aatype[j] != kGly;
        float best = CUDART_INF_F;
#pragma unroll
        for (int p = 0; p < 5; ++p) {
            if (p == 3 && !cb_i) continue;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                if (q == 3 && !cb_j) continue;
                const float dx = sa[i * 15 + p * 3] - sa[j * 15 + q * 3], dy = sa[i * 15 + p * 3 + 1] - sa[j * 15 + q * 3 + 1],
                            dz = sa[i * 15 + p * 3 + 2] - sa[j * 15 + q * 3 + 2];
                best = fminf(best, dx * dx + dy * dy + dz * dz);
            }
        }
        heavy_min = fminf(heavy_min, best);
    }
    heavy_min = sqrtf(heavy_min);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ca_max = fmaxf(ca_max, __shfl_xor_sync(0xffffffffu, ca_max, o));
        cn_max = fmaxf(cn_max, __shfl_xor_sync(0xffffffffu, cn_max, o));
        heavy_min = fminf(heavy_min, __shfl_xor_sync(0xffffffffu, heavy_min, o));
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = ca_max; red[1][tid >> 5] = cn_max; red[2][tid >> 5] = heavy_min; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, c = 0.f, h = CUDART_INF_F;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a = fmaxf(a, red[0][w]); c = fmaxf(c, red[1][w]); h = fminf(h, red[2][w]); }
        out[b * 3] = a; out[b * 3 + 1] = c; out[b * 3 + 2] = h;
    }
}

}  // namespace

extern "C" {

int se3_backbone_atoms(const float* pos, const float* rot, const int32_t* aatype, const uint8_t* pos_is_known, float* atoms, int64_t batch,
                       int len, se3_stream_t stream) {
    SE3_REQUIRE(batch >= 0 && len >= 0, "negative size");
    if (batch == 0 || len == 0) return SE3_OK;
    SE3_REQUIRE(pos && rot && aatype && atoms, "null pointer");
    const int64_t n = batch * len;
    k_backbone<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pos, rot, aatype, pos_is_known, atoms, n, len);
    count_launch();
    return check_launch("se3_backbone_atoms");
}

int se3_physicality(const float* atoms, const int32_t* aatype, float* out, int64_t batch, int len, se3_stream_t stream) {
    SE3_REQUIRE(batch >= 0 && len >= 0, "negative size");
    if (batch == 0) return SE3_OK;
    SE3_REQUIRE(atoms && aatype && out && len > 0, "null pointer or empty sequence");
    const size_t smem = (size_t)len * 15 * sizeof(float);
    SE3_REQUIRE(smem <= 200 * 1024, "sequence too long for the shared-memory staging of this kernel");
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_physicality, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("se3_physicality smem attribute: %s", cudaGetErrorString(e)); return SE3_ECUDA; }
    }
    k_physicality<<<(unsigned)batch, 256, smem, (cudaStream_t)stream>>>(atoms, aatype, out, len);
    count_launch();
    return check_launch("se3_physicality");
}

}  // extern "C"
