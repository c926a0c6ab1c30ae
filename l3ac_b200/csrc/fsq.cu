// FSQ bottleneck (bandwidth-bound, warp-per-token, 128-bit coalesced HBM access).
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

// F == 128: lane owns features 4*lane .. 4*lane+3.
__global__ void __launch_bounds__(256) fsq_quantize_kernel(const float* __restrict__ x, long long M,
                                                           const float* __restrict__ w_in,
                                                           const float* __restrict__ b_in,
                                                           const float* __restrict__ w_out,
                                                           const float* __restrict__ b_out, FsqLevels lv,
                                                           float* __restrict__ q_feature, int32_t* __restrict__ indices,
                                                           float* __restrict__ level_indices, float* __restrict__ z_out) {
    constexpr int F = 128;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    // per-lane weight slices live in registers for the whole grid-stride loop
    float4 win[kFsqMaxD];
    float wout[4][kFsqMaxD];
    float bin[kFsqMaxD];
#pragma unroll
    for (int d = 0; d < kFsqMaxD; ++d) {
        if (d < lv.D) {
            win[d] = __ldg(reinterpret_cast<const float4*>(w_in + d * F) + lane);
            bin[d] = __ldg(b_in + d);
#pragma unroll
            for (int i = 0; i < 4; ++i) wout[i][d] = __ldg(w_out + (4 * lane + i) * lv.D + d);
        }
    }
    const float4 bo = __ldg(reinterpret_cast<const float4*>(b_out) + lane);

    for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M;
         row += (long long)gridDim.x * warps_per_block) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + row * F) + lane);
        float qz[kFsqMaxD];
        int index = 0;
#pragma unroll
        for (int d = 0; d < kFsqMaxD; ++d) {
            if (d < lv.D) {
                float p = xv.x * win[d].x;
                p = fmaf(xv.y, win[d].y, p);
                p = fmaf(xv.z, win[d].z, p);
                p = fmaf(xv.w, win[d].w, p);
                const float z = warp_sum(p) + bin[d];
                float level;
                fsq_round(z, lv.levels[d], qz[d], level);
                index += (int)level * lv.basis[d];
                if (lane == 0) {
                    if (z_out) z_out[row * lv.D + d] = z;
                    if (level_indices) level_indices[row * lv.D + d] = level;
                }
            }
        }
        if (lane == 0) indices[row] = index;
        float o[4] = {bo.x, bo.y, bo.z, bo.w};
#pragma unroll
        for (int d = 0; d < kFsqMaxD; ++d) {
            if (d < lv.D) {
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = fmaf(wout[i][d], qz[d], o[i]);
            }
        }
        reinterpret_cast<float4*>(q_feature + row * F)[lane] = make_float4(o[0], o[1], o[2], o[3]);
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

template <typename IdxT>
__global__ void __launch_bounds__(256) fsq_dequantize_kernel(const IdxT* __restrict__ indices, long long M,
                                                             const float* __restrict__ w_out,
                                                             const float* __restrict__ b_out, FsqLevels lv,
                                                             float* __restrict__ q_feature) {
    constexpr int F = 128;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    float wout[4][kFsqMaxD];
#pragma unroll
    for (int d = 0; d < kFsqMaxD; ++d)
        if (d < lv.D) {
#pragma unroll
            for (int i = 0; i < 4; ++i) wout[i][d] = __ldg(w_out + (4 * lane + i) * lv.D + d);
        }
    const float4 bo = __ldg(reinterpret_cast<const float4*>(b_out) + lane);
    for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M;
         row += (long long)gridDim.x * warps_per_block) {
        const long long idx = (long long)__ldg(indices + row);
        float o[4] = {bo.x, bo.y, bo.z, bo.w};
#pragma unroll
        for (int d = 0; d < kFsqMaxD; ++d) {
            if (d < lv.D) {
                // (idx // basis) % L with floor semantics (l3ac/vq/fsq.py:70-71); indices are non-negative
                const int level = (int)((idx / lv.basis[d]) % lv.levels[d]);
                const float q_act = __fdiv_rn((float)level, (float)(lv.levels[d] - 1));
                const float qz = __fsub_rn(__fmul_rn(q_act, 2.0f), 1.0f);
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = fmaf(wout[i][d], qz, o[i]);
            }
        }
        reinterpret_cast<float4*>(q_feature + row * F)[lane] = make_float4(o[0], o[1], o[2], o[3]);
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
    if (g > 148LL * 16) g = 148LL * 16;
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
    fsq_quantize_kernel<<<fsq_grid(M, 8), 256, 0, (cudaStream_t)stream>>>(x, M, w_in, b_in, w_out, b_out, lv, q_feature,
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
        fsq_dequantize_kernel<long long><<<fsq_grid(M, 8), 256, 0, st>>>((const long long*)indices, M, w_out, b_out, lv,
                                                                         q_feature);
    else
        fsq_dequantize_kernel<int32_t><<<fsq_grid(M, 8), 256, 0, st>>>((const int32_t*)indices, M, w_out, b_out, lv,
                                                                       q_feature);
    return l3ac_launch_status();
}
