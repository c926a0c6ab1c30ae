// FSQ bottleneck (bandwidth-bound, eight lanes per token, 128-bit coalesced HBM access).
//   quantize    VQEmbed.forward      l3ac/vq/__init__.py:25-30 ; SuperFSQ.forward l3ac/vq/fsq.py:30-68
//   dequantize  VQEmbed.to_features  l3ac/vq/__init__.py:20-23 ; l3ac/vq/fsq.py:70-81
// The codebook (prod(levels) = 117,649 / 250,047 entries) is implicit: index = sum_d level_d * basis_d,
// basis = cumprod([1, L0, ..., L_{D-2}]) (dimension 0 least significant, l3ac/vq/fsq.py:15).
// Arithmetic mirrors the reference op by op in fp32 (tanh, +1, /2, *(L-1), round-half-even, /(L-1), *2-1);
// every step except tanhf is correctly rounded on both sides, so indices are bit-exact given the same
// latents unless a value sits within an ulp of a rounding tie.
#include "common.cuh"

namespace l3ac {

constexpr int kFsqMaxD = 8;

struct FsqLevels {
    int levels[kFsqMaxD];
    int basis[kFsqMaxD];
    int D;
};

__device__ __forceinline__ void fsq_round(float z, int L, float& q_z, float& level) {
    const float lm1 = (float)(L - 1);
    const float act = __fdiv_rn(__fadd_rn(tanhf(z), 1.0f), 2.0f);   // (tanh z + 1) / 2   fsq_act.py:39
    level = rintf(__fmul_rn(act, lm1));                              // round half to even  fsq.py:59
    const float q_act = __fdiv_rn(level, lm1);                       // fsq.py:60
    q_z = __fsub_rn(__fmul_rn(q_act, 2.0f), 1.0f);                   // fsq.py:21
}

// F == 128.  EIGHT LANES PER TOKEN, four tokens per warp and iteration: lane j of a group loads the float4 pieces j, j + 8,
// j + 16, j + 24 of its token's row (every load instruction covers 128 contiguous bytes per token), accumulates its 16 features
// into the D partial dot products with the projection weights read from shared memory (all four groups read the same
// addresses: broadcast, conflict-free), and a reduce-scatter over the group (4 + 2 + 1 exchange shuffles) leaves lane j with
// the complete latent of dimension j -- tanh / rounding run once per dimension and token.  The D codes are broadcast back,
// every lane produces its 16 output features, and the stores mirror the loads.  Two iterations (eight tokens, 4 KB) are
// requested per warp before the first is consumed.  ~80 instructions per token instead of ~150 for the warp-per-token kernel.
constexpr int kFsqIters = 2;

// (volatile: ptxas would otherwise hoist all 48 loop-invariant weight vectors into registers -- 228 of them, one block per SM)
__device__ __forceinline__ float4 lds_f4(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

__global__ void __launch_bounds__(256, 2) fsq_quantize_kernel(const float* __restrict__ x, long long M,
                                                           const float* __restrict__ w_in,
                                                           const float* __restrict__ b_in,
                                                           const float* __restrict__ w_out,
                                                           const float* __restrict__ b_out, FsqLevels lv,
                                                           float* __restrict__ q_feature, int32_t* __restrict__ indices,
                                                           float* __restrict__ level_indices, float* __restrict__ z_out) {
    constexpr int F = 128;
    __shared__ __align__(16) float s_win[kFsqMaxD * F];       // [d][f]
    __shared__ __align__(16) float s_wout[kFsqMaxD * F];      // [d][f] (transposed from the (F, D) weight)
    const int lane = threadIdx.x & 31, j = lane & 7, grp = lane >> 3;
    const int warps_per_block = blockDim.x >> 5;
    for (int i = threadIdx.x; i < kFsqMaxD * F; i += blockDim.x) {
        const int d = i / F, f = i - d * F;
        s_win[i] = d < lv.D ? __ldg(w_in + i) : 0.f;
        s_wout[i] = d < lv.D ? __ldg(w_out + f * lv.D + d) : 0.f;
    }
    __syncthreads();
    float4 bo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) bo[i] = __ldg(reinterpret_cast<const float4*>(b_out) + j + 8 * i);
    const bool has_dim = j < lv.D;
    const float my_bin = has_dim ? __ldg(b_in + j) : 0.f;
    const int my_L = has_dim ? lv.levels[j] : 2, my_basis = has_dim ? lv.basis[j] : 0;
    const unsigned full = 0xffffffffu;
    const bool hA = j & 4, hB = j & 2, hC = j & 1;

    const long long tokens_per_iter = 4;
    const long long stride = (long long)gridDim.x * warps_per_block * tokens_per_iter;
    for (long long base = ((long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5)) * tokens_per_iter; base < M; base += kFsqIters * stride) {
        float4 xv[kFsqIters][4];
        long long rows[kFsqIters];
#pragma unroll
        for (int u = 0; u < kFsqIters; ++u) {
            rows[u] = base + u * stride + grp;
            const bool ok = rows[u] < M;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                xv[u][i] = ok ? __ldcs(reinterpret_cast<const float4*>(x + rows[u] * F) + j + 8 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kFsqIters; ++u) {
            if (base + u * stride >= M) break;                 // warp-uniform: no token of this iteration exists
            const long long row = rows[u];
            const bool ok = row < M;
            float p[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                p[d] = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 w = lds_f4(s_win + d * F + 4 * (j + 8 * i));
                    p[d] = fmaf(xv[u][i].x, w.x, p[d]);
                    p[d] = fmaf(xv[u][i].y, w.y, p[d]);
                    p[d] = fmaf(xv[u][i].z, w.z, p[d]);
                    p[d] = fmaf(xv[u][i].w, w.w, p[d]);
                }
            }
            // reduce-scatter over the 8 lanes of the group: lane j ends up with dimension j
            float a[4], b2[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = (hA ? p[i + 4] : p[i]) + __shfl_xor_sync(full, hA ? p[i] : p[i + 4], 4);
#pragma unroll
            for (int i = 0; i < 2; ++i) b2[i] = (hB ? a[i + 2] : a[i]) + __shfl_xor_sync(full, hB ? a[i] : a[i + 2], 2);
            const float z = (hC ? b2[1] : b2[0]) + __shfl_xor_sync(full, hC ? b2[0] : b2[1], 1) + my_bin;
            float my_qz, level;
            fsq_round(z, my_L, my_qz, level);
            int index = has_dim ? (int)level * my_basis : 0;
            index += __shfl_xor_sync(full, index, 4);
            index += __shfl_xor_sync(full, index, 2);
            index += __shfl_xor_sync(full, index, 1);
            if (ok && has_dim) {
                if (z_out) z_out[row * lv.D + j] = z;
                if (level_indices) level_indices[row * lv.D + j] = level;
            }
            if (ok && j == 0) indices[row] = index;
            float4 o[4] = {bo[0], bo[1], bo[2], bo[3]};
#pragma unroll
            for (int d = 0; d < kFsqMaxD; ++d) {
                if (d < lv.D) {
                    const float qz = __shfl_sync(full, my_qz, (lane & ~7) + d);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 w = lds_f4(s_wout + d * F + 4 * (j + 8 * i));
                        o[i].x = fmaf(w.x, qz, o[i].x);
                        o[i].y = fmaf(w.y, qz, o[i].y);
                        o[i].z = fmaf(w.z, qz, o[i].z);
                        o[i].w = fmaf(w.w, qz, o[i].w);
                    }
                }
            }
            if (ok) {
#pragma unroll
                for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(q_feature + row * F)[j + 8 * i] = o[i];   // read next by the decoder: keep in L2
            }
        }
    }
}

__global__ void __launch_bounds__(256) fsq_latents_kernel(const float* __restrict__ z, long long M, FsqLevels lv,
                                                          float* __restrict__ q_z, int32_t* __restrict__ indices,
                                                          float* __restrict__ level_indices) {
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < M;
         row += (long long)gridDim.x * blockDim.x) {
        int index = 0;
        for (int d = 0; d < lv.D; ++d) {
            float q, level;
            fsq_round(z[row * lv.D + d], lv.levels[d], q, level);
            index += (int)level * lv.basis[d];
            if (q_z) q_z[row * lv.D + d] = q;
            if (level_indices) level_indices[row * lv.D + d] = level;
        }
        indices[row] = index;
    }
}

// Same lane mapping as the quantizer: eight lanes per token, lane j < D decodes digit j of the token's index, the D codes are
// broadcast within the group and every lane produces 16 output features (four 128-byte-per-token store instructions).
template <typename IdxT>
__global__ void __launch_bounds__(256, 2) fsq_dequantize_kernel(const IdxT* __restrict__ indices, long long M,
                                                                const float* __restrict__ w_out,
                                                                const float* __restrict__ b_out, FsqLevels lv,
                                                                float* __restrict__ q_feature) {
    constexpr int F = 128;
    constexpr int kIters = 4;
    __shared__ __align__(16) float s_wout[kFsqMaxD * F];      // [d][f]
    const int lane = threadIdx.x & 31, j = lane & 7, grp = lane >> 3;
    const int warps_per_block = blockDim.x >> 5;
    for (int i = threadIdx.x; i < kFsqMaxD * F; i += blockDim.x) {
        const int d = i / F, f = i - d * F;
        s_wout[i] = d < lv.D ? __ldg(w_out + f * lv.D + d) : 0.f;
    }
    __syncthreads();
    float4 bo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) bo[i] = __ldg(reinterpret_cast<const float4*>(b_out) + j + 8 * i);
    // the divisions are by per-lane constants hoisted out of the loop
    const bool has_dim = j < lv.D;
    const long long my_basis = has_dim ? lv.basis[j] : 1;
    const int my_L = has_dim ? lv.levels[j] : 2;
    const float my_lm1 = (float)(my_L - 1);
    const long long stride = (long long)gridDim.x * warps_per_block * 4;
    for (long long base = ((long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5)) * 4; base < M; base += kIters * stride) {
        long long idx[kIters];
#pragma unroll
        for (int u = 0; u < kIters; ++u) idx[u] = base + u * stride + grp < M ? (long long)__ldg(indices + base + u * stride + grp) : 0;
#pragma unroll
        for (int u = 0; u < kIters; ++u) {
            if (base + u * stride >= M) break;                 // warp-uniform
            const long long row = base + u * stride + grp;
            // (idx // basis) % L with floor semantics (l3ac/vq/fsq.py:70-71); indices are non-negative
            const int level = sizeof(IdxT) == 4 ? (int)(((unsigned)idx[u] / (unsigned)my_basis) % (unsigned)my_L)
                                                : (int)((idx[u] / my_basis) % my_L);
            const float my_qz = __fsub_rn(__fmul_rn(__fdiv_rn((float)level, my_lm1), 2.0f), 1.0f);
            float4 o[4] = {bo[0], bo[1], bo[2], bo[3]};
#pragma unroll
            for (int d = 0; d < kFsqMaxD; ++d) {
                if (d < lv.D) {
                    const float qz = __shfl_sync(0xffffffffu, my_qz, (lane & ~7) + d);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 w = lds_f4(s_wout + d * F + 4 * (j + 8 * i));
                        o[i].x = fmaf(w.x, qz, o[i].x);
                        o[i].y = fmaf(w.y, qz, o[i].y);
                        o[i].z = fmaf(w.z, qz, o[i].z);
                        o[i].w = fmaf(w.w, qz, o[i].w);
                    }
                }
            }
            if (row < M) {
#pragma unroll
                for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(q_feature + row * F)[j + 8 * i] = o[i];   // read next by the decoder: keep in L2
            }
        }
    }
}

static int make_levels(const int* levels, int D, FsqLevels* out) {
    if (!levels || D < 1 || D > kFsqMaxD) return L3AC_EINVAL;
    long long basis = 1;
    for (int d = 0; d < kFsqMaxD; ++d) {
        out->levels[d] = d < D ? levels[d] : 1;
        out->basis[d] = (int)basis;
        if (d < D) {
            if (levels[d] < 2) return L3AC_EINVAL;
            basis *= levels[d];
            if (basis >= (1LL << 24)) return L3AC_EINVAL;   // keeps the reference's fp32 index sum exact
        }
    }
    out->D = D;
    return L3AC_OK;
}

static int fsq_grid(long long M, int rows_per_block) {
    long long g = (M + rows_per_block - 1) / rows_per_block;
    const long long cap = (long long)l3ac_sm_count() * 8;           // 8 resident 256-thread blocks per SM, each looping
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace l3ac

using namespace l3ac;

extern "C" int l3ac_fsq_quantize(const float* x, long long M, int F, const float* w_in, const float* b_in,
                                 const float* w_out, const float* b_out, const int* levels, int D, float* q_feature,
                                 int32_t* indices, float* level_indices, float* z, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(x && w_in && b_in && w_out && b_out && q_feature && indices && M > 0);
    if (F != 128) return L3AC_EUNSUPPORTED;
    FsqLevels lv;
    const int rc = make_levels(levels, D, &lv);
    if (rc != L3AC_OK) return rc;
    fsq_quantize_kernel<<<fsq_grid(M, 32), 256, 0, (cudaStream_t)stream>>>(x, M, w_in, b_in, w_out, b_out, lv, q_feature,
                                                                          indices, level_indices, z);
    return l3ac_launch_status();
}

extern "C" int l3ac_fsq_quantize_latents(const float* z, long long M, const int* levels, int D, float* q_z,
                                         int32_t* indices, float* level_indices, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(z && indices && M > 0);
    FsqLevels lv;
    const int rc = make_levels(levels, D, &lv);
    if (rc != L3AC_OK) return rc;
    fsq_latents_kernel<<<fsq_grid(M, 256), 256, 0, (cudaStream_t)stream>>>(z, M, lv, q_z, indices, level_indices);
    return l3ac_launch_status();
}

extern "C" int l3ac_fsq_dequantize(const void* indices, int indices_are_i64, long long M, int F, const float* w_out,
                                   const float* b_out, const int* levels, int D, float* q_feature,
                                   l3ac_stream_t stream) {
    L3AC_CHECK_ARG(indices && w_out && b_out && q_feature && M > 0);
    if (F != 128) return L3AC_EUNSUPPORTED;
    FsqLevels lv;
    const int rc = make_levels(levels, D, &lv);
    if (rc != L3AC_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (indices_are_i64)
        fsq_dequantize_kernel<long long><<<fsq_grid(M, 32), 256, 0, st>>>((const long long*)indices, M, w_out, b_out, lv,
                                                                         q_feature);
    else
        fsq_dequantize_kernel<int32_t><<<fsq_grid(M, 32), 256, 0, st>>>((const int32_t*)indices, M, w_out, b_out, lv,
                                                                       q_feature);
    return l3ac_launch_status();
}
