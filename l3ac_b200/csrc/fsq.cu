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

// Lane group of dimension d: the four lanes whose bits (4, 3, 2) spell d.  The eight per-dimension partial sums of a token
// are reduced reduce-scatter style (4 + 2 + 1 exchange shuffles leave every lane with ONE dimension summed over 8 lanes, two
// butterfly steps finish it): 9 shuffles instead of 8 x 5, and tanh / rounding run once per dimension instead of 32 times.
__device__ __forceinline__ int fsq_lane_dim(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }
__device__ __forceinline__ int fsq_dim_lane(int d) { return ((d >> 2) & 1) * 16 + ((d >> 1) & 1) * 8 + (d & 1) * 4; }

__device__ __forceinline__ float fsq_reduce8(const float (&p)[8], int lane) {
    const unsigned full = 0xffffffffu;
    const bool hA = lane & 16, hB = lane & 8, hC = lane & 4;
    float a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = (hA ? p[i + 4] : p[i]) + __shfl_xor_sync(full, hA ? p[i] : p[i + 4], 16);
#pragma unroll
    for (int i = 0; i < 2; ++i) b[i] = (hB ? a[i + 2] : a[i]) + __shfl_xor_sync(full, hB ? a[i] : a[i + 2], 8);
    float c = (hC ? b[1] : b[0]) + __shfl_xor_sync(full, hC ? b[0] : b[1], 4);
    c += __shfl_xor_sync(full, c, 2);
    c += __shfl_xor_sync(full, c, 1);
    return c;
}

// F == 128: lane owns features 4*lane .. 4*lane+3.  One warp per token, kInFlight tokens (512 B each) requested per warp before
// the first is consumed: with ~16 resident warps per SM (the per-lane weight slices cost ~64 registers) that is 32 KB in
// flight per SM -- what 6.4 TB/s x ~800 ns of HBM latency needs; two in flight ran at 0.28 of the HBM peak.
constexpr int kFsqInFlight = 4;
__global__ void __launch_bounds__(256) fsq_quantize_kernel(const float* __restrict__ x, long long M,
                                                           const float* __restrict__ w_in,
                                                           const float* __restrict__ b_in,
                                                           const float* __restrict__ w_out,
                                                           const float* __restrict__ b_out, FsqLevels lv,
                                                           float* __restrict__ q_feature, int32_t* __restrict__ indices,
                                                           float* __restrict__ level_indices, float* __restrict__ z_out) {
    constexpr int F = 128;
    __shared__ __align__(16) float s_win[kFsqMaxD * F];
    __shared__ float s_wout[F * kFsqMaxD];
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int i = threadIdx.x; i < lv.D * F; i += blockDim.x) {          // coalesced staging; lanes then take their slices
        s_win[i] = __ldg(w_in + i);
        s_wout[i] = __ldg(w_out + i);
    }
    __syncthreads();
    float4 win[kFsqMaxD];
    float wout[4][kFsqMaxD];
#pragma unroll
    for (int d = 0; d < kFsqMaxD; ++d) {
        win[d] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d < lv.D) {
            win[d] = reinterpret_cast<const float4*>(s_win + d * F)[lane];
#pragma unroll
            for (int i = 0; i < 4; ++i) wout[i][d] = s_wout[(4 * lane + i) * lv.D + d];
        }
    }
    const float4 bo = __ldg(reinterpret_cast<const float4*>(b_out) + lane);
    const int my_d = fsq_lane_dim(lane);
    const bool has_dim = my_d < lv.D;
    const float my_bin = has_dim ? __ldg(b_in + my_d) : 0.f;
    const int my_L = has_dim ? lv.levels[my_d] : 2, my_basis = has_dim ? lv.basis[my_d] : 0;
    const bool writer = has_dim && (lane & 3) == 0;

    const long long stride = (long long)gridDim.x * warps_per_block;
    for (long long row0 = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row0 < M; row0 += kFsqInFlight * stride) {
        long long rows[kFsqInFlight];
        float4 xv[kFsqInFlight];
#pragma unroll
        for (int u = 0; u < kFsqInFlight; ++u) {
            rows[u] = row0 + u * stride;
            xv[u] = rows[u] < M ? __ldcs(reinterpret_cast<const float4*>(x + rows[u] * F) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kFsqInFlight; ++u) {
            if (rows[u] >= M) break;                           // warp-uniform
            const long long row = rows[u];
            float p[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                p[d] = xv[u].x * win[d].x;
                p[d] = fmaf(xv[u].y, win[d].y, p[d]);
                p[d] = fmaf(xv[u].z, win[d].z, p[d]);
                p[d] = fmaf(xv[u].w, win[d].w, p[d]);
            }
            const float z = fsq_reduce8(p, lane) + my_bin;
            float my_qz, level;
            fsq_round(z, my_L, my_qz, level);
            int index = has_dim ? (int)level * my_basis : 0;
            index += __shfl_xor_sync(0xffffffffu, index, 4);
            index += __shfl_xor_sync(0xffffffffu, index, 8);
            index += __shfl_xor_sync(0xffffffffu, index, 16);
            if (writer) {
                if (z_out) z_out[row * lv.D + my_d] = z;
                if (level_indices) level_indices[row * lv.D + my_d] = level;
            }
            if (lane == 0) indices[row] = index;
            float o[4] = {bo.x, bo.y, bo.z, bo.w};
#pragma unroll
            for (int d = 0; d < kFsqMaxD; ++d) {
                if (d < lv.D) {
                    const float qz = __shfl_sync(0xffffffffu, my_qz, fsq_dim_lane(d));
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[i] = fmaf(wout[i][d], qz, o[i]);
                }
            }
            reinterpret_cast<float4*>(q_feature + row * F)[lane] = make_float4(o[0], o[1], o[2], o[3]);   // read next by the decoder: keep in L2
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

template <typename IdxT>
__global__ void __launch_bounds__(256) fsq_dequantize_kernel(const IdxT* __restrict__ indices, long long M,
                                                             const float* __restrict__ w_out,
                                                             const float* __restrict__ b_out, FsqLevels lv,
                                                             float* __restrict__ q_feature) {
    constexpr int F = 128;
    __shared__ float s_wout[F * kFsqMaxD];
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int i = threadIdx.x; i < lv.D * F; i += blockDim.x) s_wout[i] = __ldg(w_out + i);
    __syncthreads();
    float wout[4][kFsqMaxD];
#pragma unroll
    for (int d = 0; d < kFsqMaxD; ++d)
        if (d < lv.D) {
#pragma unroll
            for (int i = 0; i < 4; ++i) wout[i][d] = s_wout[(4 * lane + i) * lv.D + d];
        }
    const float4 bo = __ldg(reinterpret_cast<const float4*>(b_out) + lane);
    // lane d (< D) decodes digit d of the token's index; the D codes are then broadcast.  The divisions are by per-lane
    // constants hoisted out of the loop.
    const bool has_dim = lane < lv.D;
    const long long my_basis = has_dim ? lv.basis[lane] : 1;
    const int my_L = has_dim ? lv.levels[lane] : 2;
    const float my_lm1 = (float)(my_L - 1);
    const long long stride = (long long)gridDim.x * warps_per_block;
    for (long long row0 = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row0 < M; row0 += 4 * stride) {
        long long idx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) idx[u] = row0 + u * stride < M ? (long long)__ldg(indices + row0 + u * stride) : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long row = row0 + u * stride;
            if (row >= M) break;                               // warp-uniform
            // (idx // basis) % L with floor semantics (l3ac/vq/fsq.py:70-71); indices are non-negative
            const int level = sizeof(IdxT) == 4 ? (int)(((unsigned)idx[u] / (unsigned)my_basis) % (unsigned)my_L)
                                                : (int)((idx[u] / my_basis) % my_L);
            const float my_qz = __fsub_rn(__fmul_rn(__fdiv_rn((float)level, my_lm1), 2.0f), 1.0f);
            float o[4] = {bo.x, bo.y, bo.z, bo.w};
#pragma unroll
            for (int d = 0; d < kFsqMaxD; ++d) {
                if (d < lv.D) {
                    const float qz = __shfl_sync(0xffffffffu, my_qz, d);
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[i] = fmaf(wout[i][d], qz, o[i]);
                }
            }
            reinterpret_cast<float4*>(q_feature + row * F)[lane] = make_float4(o[0], o[1], o[2], o[3]);   // read next by the decoder: keep in L2
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
