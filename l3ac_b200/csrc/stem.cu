// Encoder stem = V3FirstBlock (l3ac/tconv/__init__.py:8-27) fused into one kernel:
//   5 x [TrendPool(k) -> Conv1d(1->4,k7,pad 3)]  (k = 1,5,11,21,45; l3ac/tconv/base.py:8-45)
//   -> Conv1d 1x1 20->80 -> exact GELU -> cat raw x -> Conv1d 1x1 81->C
// audio (B,T) -> out (B,T,C) channels-last.  One block = 256 consecutive samples of one clip (2 per thread),
// halo 47 = 44 (max+avg pool of 45) + 3 (conv k7).
#include "common.cuh"

namespace l3ac {

constexpr int kStemThreads = 128;
constexpr int kStemSPT = 1;                       // samples per thread (2 halves the weight LDS traffic but 201 registers cut occupancy: measured 3.65 vs 3.29 ms)
constexpr int kStemTile = kStemThreads * kStemSPT;
constexpr int kStemReach = 47;
constexpr int kStemW = kStemTile + 2 * kStemReach;
constexpr int kStemH = 80;

template <int CO>
__global__ void __launch_bounds__(kStemThreads) stem_kernel(const float* __restrict__ audio, int B, int T,
                                                         const float* __restrict__ branch_w,
                                                         const float* __restrict__ branch_b,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         float* __restrict__ out) {
    __shared__ float xs[kStemW], ms[kStemW], ps[kStemW];
    __shared__ __align__(16) float s_w1[kStemH * 20];
    __shared__ __align__(16) float s_w2t[(kStemH + 1) * CO];   // [u][c], u = 80 is the raw-x column
    __shared__ float s_b1[kStemH], s_b2[CO], s_bw[5 * 4 * 7], s_bb[20];

    const int b = blockIdx.y, t0 = blockIdx.x * kStemTile;
    const float* xb = audio + (long long)b * T;
    for (int i = threadIdx.x; i < kStemW; i += blockDim.x) {
        const int t = t0 - kStemReach + i;
        xs[i] = (t >= 0 && t < T) ? __ldg(xb + t) : 0.f;
    }
    for (int i = threadIdx.x; i < kStemH * 20; i += blockDim.x) s_w1[i] = w1[i];
    for (int i = threadIdx.x; i < (kStemH + 1) * CO; i += blockDim.x) {
        const int u = i / CO, c = i - u * CO;
        s_w2t[i] = w2[c * (kStemH + 1) + u];
    }
    for (int i = threadIdx.x; i < kStemH; i += blockDim.x) s_b1[i] = b1[i];
    for (int i = threadIdx.x; i < CO; i += blockDim.x) s_b2[i] = b2[i];
    for (int i = threadIdx.x; i < 140; i += blockDim.x) s_bw[i] = branch_w[i];
    for (int i = threadIdx.x; i < 20; i += blockDim.x) s_bb[i] = branch_b[i];
    __syncthreads();

    float h[kStemSPT][20];
    const int pool_k[5] = {1, 5, 11, 21, 45};
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int k = pool_k[j], half = k >> 1;
        const float* src = xs;
        if (k > 1) {
            for (int i = threadIdx.x; i < kStemW; i += blockDim.x) {
                const int t = t0 - kStemReach + i;
                float m = 0.f;
                if (t >= 0 && t < T && i >= half && i < kStemW - half) {
                    for (int e = -half; e <= half; ++e) m = fmaxf(m, fabsf(xs[i + e]));
                }
                ms[i] = m;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < kStemW; i += blockDim.x) {
                const int t = t0 - kStemReach + i;
                float a = 0.f;
                if (t >= 0 && t < T && i >= 2 * half && i < kStemW - 2 * half) {
                    for (int e = -half; e <= half; ++e) a += ms[i + e];
                    a = a / (float)k;
                }
                ps[i] = a;
            }
            __syncthreads();
            src = ps;
        }
#pragma unroll
        for (int sp = 0; sp < kStemSPT; ++sp) {
            const int li = threadIdx.x + sp * kStemThreads + kStemReach;   // this thread's sample inside xs/ps
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float acc = s_bb[j * 4 + c];
#pragma unroll
                for (int q = 0; q < 7; ++q) acc = fmaf(s_bw[(j * 4 + c) * 7 + q], src[li + q - 3], acc);
                h[sp][j * 4 + c] = acc;
            }
        }
        __syncthreads();   // ms/ps are rewritten by the next branch
    }

    // The two 1x1 convs (20 -> 80 -> GELU -> 24) as packed FFMA2: inputs and accumulators are float2 pairs.
    float2 acc2[kStemSPT][CO / 2];
#pragma unroll
    for (int sp = 0; sp < kStemSPT; ++sp) {
        const float xv = xs[threadIdx.x + sp * kStemThreads + kStemReach];
#pragma unroll
        for (int c = 0; c < CO / 2; ++c)
            acc2[sp][c] = make_float2(fmaf(s_w2t[kStemH * CO + 2 * c], xv, s_b2[2 * c]),
                                      fmaf(s_w2t[kStemH * CO + 2 * c + 1], xv, s_b2[2 * c + 1]));
    }
    for (int u = 0; u < kStemH; ++u) {
        float2 ap[kStemSPT];
#pragma unroll
        for (int sp = 0; sp < kStemSPT; ++sp) ap[sp] = make_float2(s_b1[u], 0.f);
        const float4* wr = reinterpret_cast<const float4*>(s_w1 + u * 20);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float4 w4 = wr[i];
#pragma unroll
            for (int sp = 0; sp < kStemSPT; ++sp) {
                ap[sp] = ffma2(make_float2(w4.x, w4.y), make_float2(h[sp][4 * i], h[sp][4 * i + 1]), ap[sp]);
                ap[sp] = ffma2(make_float2(w4.z, w4.w), make_float2(h[sp][4 * i + 2], h[sp][4 * i + 3]), ap[sp]);
            }
        }
        float2 g[kStemSPT];
#pragma unroll
        for (int sp = 0; sp < kStemSPT; ++sp) {
            const float gv = gelu_erf(ap[sp].x + ap[sp].y);          // parity mode: erff (not the A&S / MUFU approximation)
            g[sp] = make_float2(gv, gv);
        }
        const float4* w2r = reinterpret_cast<const float4*>(s_w2t + u * CO);
#pragma unroll
        for (int i = 0; i < CO / 4; ++i) {
            const float4 w4 = w2r[i];
#pragma unroll
            for (int sp = 0; sp < kStemSPT; ++sp) {
                acc2[sp][2 * i] = ffma2(make_float2(w4.x, w4.y), g[sp], acc2[sp][2 * i]);
                acc2[sp][2 * i + 1] = ffma2(make_float2(w4.z, w4.w), g[sp], acc2[sp][2 * i + 1]);
            }
        }
    }
    float acc[kStemSPT][CO];
#pragma unroll
    for (int sp = 0; sp < kStemSPT; ++sp)
#pragma unroll
        for (int c = 0; c < CO / 2; ++c) {
            acc[sp][2 * c] = acc2[sp][c].x;
            acc[sp][2 * c + 1] = acc2[sp][c].y;
        }
#pragma unroll
    for (int sp = 0; sp < kStemSPT; ++sp) {
        const int t = t0 + threadIdx.x + sp * kStemThreads;
        if (t >= T) continue;
        float4* o = reinterpret_cast<float4*>(out + ((long long)b * T + t) * CO);
#pragma unroll
        for (int i = 0; i < CO / 4; ++i)
            o[i] = make_float4(acc[sp][4 * i], acc[sp][4 * i + 1], acc[sp][4 * i + 2], acc[sp][4 * i + 3]);
    }
}

}  // namespace l3ac

extern "C" int l3ac_stem(const float* audio, int B, int T, const float* branch_w, const float* branch_b,
                         const float* w1, const float* b1, const float* w2, const float* b2, int C, float* out,
                         l3ac_stream_t stream) {
    L3AC_CHECK_ARG(audio && branch_w && branch_b && w1 && b1 && w2 && b2 && out);
    L3AC_CHECK_ARG(B > 0 && B <= 65535 && T > 0);
    if (C != 24) return L3AC_EUNSUPPORTED;
    dim3 grid(l3ac_cdiv(T, l3ac::kStemTile), B);
    l3ac::stem_kernel<24><<<grid, l3ac::kStemThreads, 0, (cudaStream_t)stream>>>(audio, B, T, branch_w, branch_b, w1, b1,
                                                                              w2, b2, out);
    return l3ac_launch_status();
}
