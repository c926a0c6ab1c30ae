// Fused thin-channel Residual(ConvUnit) in fp32 (l3ac/modules.py:10-44 for C = 24, the full-rate encoder stage):
//   out = x + pw_conv2( GRN( snake( pw_conv1( LayerNorm( dwconv7(x) ) ) ) ) )
// With 24 channels the two point-wise GEMMs are 2 x 2304 MAC per time step: as tensor-core launches they are bound by
// the TMA row rate (48-byte rows) and by a 4C-wide hidden tensor (1.8 GB per 24 clips as a split-bf16 pair) that has to
// round-trip through HBM.  Here one thread owns one time step: depthwise conv and LayerNorm are in-thread (no
// shuffles) for two time steps, the hidden activation lives one value at a time in a register, the weights are
// warp-broadcast float4 reads from shared memory (each feeding both rows), and HBM sees only x in and x out (192 B per time step).  Exact fp32 arithmetic, so the
// encode side needs no operand splitting here.
#include "common.cuh"

namespace l3ac {
namespace thin {

constexpr int kC = 24;
constexpr int kH = 96;
constexpr int kThreads = 128;
constexpr int kSPT = 2;                 // time steps per thread: every weight LDS feeds two rows (the kernel is LDS-bound at one)
constexpr int kTile = kThreads * kSPT;  // time steps per CTA
constexpr int kPitch = 28;              // floats per staged x row: float4 reads by consecutive threads are conflict-free

struct Smem {
    float xs[(kTile + 6) * kPitch];
    float w1[kH * kC];       // [u][c]
    float w2t[kH * kC];      // [u][c] = w2[c][u]
    float dw[7 * kC];
    float par[kH * 8];       // per hidden unit: b1, alpha, 1/(alpha+eps), scale, shift, 3 pad -> two 16-byte broadcast loads
    float c[4 * kC];         // dw_b, ln_w, ln_b, b2
};

// SPLIT: write the result as the split-bf16 pair (hi = bf16(v), lo = bf16(v - hi)) the next tcgen05 GEMM consumes, instead
// of fp32 -- same bytes, and the separate fp32 -> split conversion pass (a full read + write of the tensor) disappears.
template <bool SPLIT>
__global__ void __launch_bounds__(kThreads, 4) convunit_thin_kernel(const float* __restrict__ x, int B, int T,
                                                                 const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                                 const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                                 float eps, const float* __restrict__ w1,
                                                                 const float* __restrict__ b1, const float* __restrict__ alpha,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 const float* __restrict__ w2, const float* __restrict__ b2,
                                                                 void* __restrict__ out_v, void* __restrict__ out_lo_v) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int b = blockIdx.y, t0 = blockIdx.x * kTile, tid = threadIdx.x;
    const float* xb = x + (long long)b * T * kC;
    for (int i = tid; i < (kTile + 6) * (kC / 4); i += kThreads) {
        const int r = i / (kC / 4), c4 = i - r * (kC / 4);
        const int t = t0 + r - 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < T) v = __ldg(reinterpret_cast<const float4*>(xb + (long long)t * kC) + c4);
        *reinterpret_cast<float4*>(sm.xs + r * kPitch + 4 * c4) = v;
    }
    for (int i = tid; i < kH * kC; i += kThreads) {
        sm.w1[i] = __ldg(w1 + i);
        const int u = i / kC, c = i - u * kC;
        sm.w2t[i] = __ldg(w2 + c * kH + u);
    }
    for (int i = tid; i < 7 * kC; i += kThreads) sm.dw[i] = __ldg(dw_w + i);
    for (int i = tid; i < kH; i += kThreads) {
        const float a = __ldg(alpha + i);
        sm.par[8 * i] = __ldg(b1 + i);
        sm.par[8 * i + 1] = a;
        sm.par[8 * i + 2] = 1.0f / (a + kEps);
        sm.par[8 * i + 3] = __ldg(scale + i);
        sm.par[8 * i + 4] = __ldg(shift + i);
    }
    if (tid < kC) {
        sm.c[tid] = __ldg(dw_b + tid);
        sm.c[kC + tid] = __ldg(ln_w + tid);
        sm.c[2 * kC + tid] = __ldg(ln_b + tid);
        sm.c[3 * kC + tid] = __ldg(b2 + tid);
    }
    __syncthreads();

    // Thread `tid` owns time steps t0 + tid + s * kThreads.  Depthwise conv k7 (zero padded) + LayerNorm over the 24
    // channels, all in registers.  Channel pairs are kept as float2 so that every multiply-add below is one packed FFMA2.
    float2 a2[kSPT][kC / 2];
#pragma unroll
    for (int s = 0; s < kSPT; ++s) {
#pragma unroll
        for (int c = 0; c < kC / 2; ++c) a2[s][c] = make_float2(sm.c[2 * c], sm.c[2 * c + 1]);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const float* xr = sm.xs + (tid + s * kThreads + j) * kPitch;
#pragma unroll
            for (int c4 = 0; c4 < kC / 4; ++c4) {
                const float4 xv = *reinterpret_cast<const float4*>(xr + 4 * c4);
                const float4 wv = *reinterpret_cast<const float4*>(sm.dw + j * kC + 4 * c4);
                a2[s][2 * c4] = ffma2(make_float2(wv.x, wv.y), make_float2(xv.x, xv.y), a2[s][2 * c4]);
                a2[s][2 * c4 + 1] = ffma2(make_float2(wv.z, wv.w), make_float2(xv.z, xv.w), a2[s][2 * c4 + 1]);
            }
        }
        float mean = 0.f;
#pragma unroll
        for (int c = 0; c < kC / 2; ++c) mean += a2[s][c].x + a2[s][c].y;
        mean *= (1.0f / kC);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < kC / 2; ++c) {
            a2[s][c].x -= mean;
            a2[s][c].y -= mean;
            var = fmaf(a2[s][c].x, a2[s][c].x, var);
            var = fmaf(a2[s][c].y, a2[s][c].y, var);
        }
        const float rstd = rsqrt_nr(var * (1.0f / kC) + eps);
#pragma unroll
        for (int c = 0; c < kC / 2; ++c) {
            a2[s][c].x = fmaf(a2[s][c].x * rstd, sm.c[kC + 2 * c], sm.c[2 * kC + 2 * c]);
            a2[s][c].y = fmaf(a2[s][c].y * rstd, sm.c[kC + 2 * c + 1], sm.c[2 * kC + 2 * c + 1]);
        }
    }

    // MLP: for every hidden unit  h = affine(snake(w1[u] . a + b1[u]))  and  acc += w2[:, u] * h
    float2 acc2[kSPT][kC / 2];
#pragma unroll
    for (int s = 0; s < kSPT; ++s) {
        const float* xr = sm.xs + (tid + s * kThreads + 3) * kPitch;
#pragma unroll
        for (int c = 0; c < kC / 2; ++c)
            acc2[s][c] = make_float2(xr[2 * c] + sm.c[3 * kC + 2 * c], xr[2 * c + 1] + sm.c[3 * kC + 2 * c + 1]);      // residual + b2
    }
#pragma unroll 2
    for (int u = 0; u < kH; ++u) {
        const float4 p0 = *reinterpret_cast<const float4*>(sm.par + 8 * u);       // b1, alpha, 1/(alpha+eps), scale
        const float shift_u = sm.par[8 * u + 4];
        float2 hp[kSPT];                                       // two partial dot products (even / odd channel pairs) per row
#pragma unroll
        for (int s = 0; s < kSPT; ++s) hp[s] = make_float2(p0.x, 0.f);
        const float4* wr = reinterpret_cast<const float4*>(sm.w1 + u * kC);
#pragma unroll
        for (int c4 = 0; c4 < kC / 4; ++c4) {
            const float4 wv = wr[c4];
#pragma unroll
            for (int s = 0; s < kSPT; ++s) {
                hp[s] = ffma2(make_float2(wv.x, wv.y), a2[s][2 * c4], hp[s]);
                hp[s] = ffma2(make_float2(wv.z, wv.w), a2[s][2 * c4 + 1], hp[s]);
            }
        }
        float2 hh[kSPT];
#pragma unroll
        for (int s = 0; s < kSPT; ++s) {
            float h = hp[s].x + hp[s].y;
            const float sn = sinf(p0.y * h);            // parity mode: libm-accurate sine (not the MUFU approximation)
            h = fmaf(p0.z, sn * sn, h);
            h = fmaf(h, p0.w, shift_u);
            hh[s] = make_float2(h, h);
        }
        const float4* w2r = reinterpret_cast<const float4*>(sm.w2t + u * kC);
#pragma unroll
        for (int c4 = 0; c4 < kC / 4; ++c4) {
            const float4 wv = w2r[c4];
#pragma unroll
            for (int s = 0; s < kSPT; ++s) {
                acc2[s][2 * c4] = ffma2(make_float2(wv.x, wv.y), hh[s], acc2[s][2 * c4]);
                acc2[s][2 * c4 + 1] = ffma2(make_float2(wv.z, wv.w), hh[s], acc2[s][2 * c4 + 1]);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < kSPT; ++s) {
        const int t = t0 + tid + s * kThreads;
        if (t >= T) continue;
        const long long base = ((long long)b * T + t) * kC;
        if (!SPLIT) {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_v) + base);
#pragma unroll
            for (int c4 = 0; c4 < kC / 4; ++c4)
                o[c4] = make_float4(acc2[s][2 * c4].x, acc2[s][2 * c4].y, acc2[s][2 * c4 + 1].x, acc2[s][2 * c4 + 1].y);
        } else {
            uint4* oh = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out_v) + base);
            uint4* ol = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out_lo_v) + base);
#pragma unroll
            for (int c8 = 0; c8 < kC / 8; ++c8) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 v = acc2[s][4 * c8 + e];
                    const __nv_bfloat162 hi = __floats2bfloat162_rn(v.x, v.y);
                    const float2 hf = __bfloat1622float2(hi);
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x - hf.x, v.y - hf.y);
                    h[e] = *reinterpret_cast<const uint32_t*>(&hi);
                    l[e] = *reinterpret_cast<const uint32_t*>(&lo);
                }
                oh[c8] = make_uint4(h[0], h[1], h[2], h[3]);
                ol[c8] = make_uint4(l[0], l[1], l[2], l[3]);
            }
        }
    }
}

}  // namespace thin
}  // namespace l3ac

extern "C" int l3ac_convunit_thin_f32(const float* x, int B, int T, int C, const float* dw_w, const float* dw_b,
                                      const float* ln_w, const float* ln_b, float eps, const float* w1, const float* b1,
                                      const float* alpha, const float* scale, const float* shift, const float* w2,
                                      const float* b2, void* out, void* out_lo, int out_dtype, l3ac_stream_t stream) {
    using namespace l3ac::thin;
    L3AC_CHECK_ARG(x && dw_w && dw_b && ln_w && ln_b && w1 && b1 && alpha && scale && shift && w2 && b2 && out);
    L3AC_CHECK_ARG(B > 0 && B <= 65535 && T > 0);
    L3AC_CHECK_ARG(out_dtype == L3AC_F32 || (out_dtype == L3AC_BF16X2 && out_lo));
    if (C != kC) return L3AC_EUNSUPPORTED;
    L3AC_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out_lo)) & 15) == 0);
    dim3 grid(l3ac_cdiv(T, kTile), B);
    auto launch = [&](auto kernel) -> int {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) return (int)e;
        kernel<<<grid, kThreads, sizeof(Smem), (cudaStream_t)stream>>>(x, B, T, dw_w, dw_b, ln_w, ln_b, eps, w1, b1, alpha, scale, shift,
                                                                       w2, b2, out, out_lo);
        return l3ac_launch_status();
    };
    return out_dtype == L3AC_F32 ? launch(convunit_thin_kernel<false>) : launch(convunit_thin_kernel<true>);
}
