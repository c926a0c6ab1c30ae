// Block-local causal attention with a Toeplitz (relative-distance) bias, fp32 SIMT flash-style.
// Semantics: LocalAttention(window w, causal, look_backward=1, exact_windowsize=False, autopad) of
// local-attention==1.11.2 as configured at l3ac/local_trans.py:34-38, with DynamicPositionBias
// gathered as bias[h][q_pos - k_pos] (l3ac/local_trans.py:43):
//   query p sees keys j with  max(0, (p/w - 1) * w) <= j <= p.
// Right zero-padding of the sequence (autopad) never reaches a real query (causal), so it is skipped.
// Block = 128 threads, one (b, h, 64-query tile); key tiles of 64; online softmax.
#include "common.cuh"

namespace l3ac {

constexpr int kAttD = 32;
constexpr int kAttBQ = 64;
constexpr int kAttBK = 64;

__global__ void __launch_bounds__(128) local_attention_kernel(const float* __restrict__ qkv,
                                                              const float* __restrict__ bias_table, int B, int T,
                                                              int H, int window, float* __restrict__ out) {
    __shared__ float Qs[kAttBQ][kAttD + 1];
    __shared__ float Ks[kAttBK][kAttD + 1];
    __shared__ __align__(8) float Vs[kAttBK][kAttD];
    __shared__ float Ps[kAttBQ][kAttBK + 1];

    const int q0 = blockIdx.x * kAttBQ, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, tk = tid & 15, tq = tid >> 4;   // tq in 0..7 -> rows 8 tq + i
    const int ld = 3 * H * kAttD;
    const float* base = qkv + (long long)b * T * ld;
    const float* bt = bias_table + (long long)h * 2 * window;
    const float qscale = 0.17677669529663687f;   // 32 ** -0.5, the python double the reference multiplies q by

    for (int e = tid; e < kAttBQ * kAttD; e += blockDim.x) {
        const int r = e / kAttD, d = e - r * kAttD;
        const int p = q0 + r;
        Qs[r][d] = (p < T) ? base[(long long)p * ld + h * kAttD + d] * qscale : 0.f;
    }

    float m_run[8], l_run[8], o[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        m_run[i] = -INFINITY;
        l_run[i] = 0.f;
        o[i][0] = o[i][1] = 0.f;
    }

    const int q_last = min(q0 + kAttBQ, T) - 1;
    int k_begin = (q0 / window - 1) * window;
    if (k_begin < 0) k_begin = 0;

    for (int k0 = k_begin; k0 <= q_last; k0 += kAttBK) {
        __syncthreads();   // previous tile fully consumed (also orders the Qs fill on the first pass)
        for (int e = tid; e < kAttBK * kAttD; e += blockDim.x) {
            const int r = e / kAttD, d = e - r * kAttD;
            const int j = k0 + r;
            float kv = 0.f, vv = 0.f;
            if (j < T) {
                const float* row = base + (long long)j * ld + h * kAttD + d;
                kv = row[H * kAttD];
                vv = row[2 * H * kAttD];
            }
            Ks[r][d] = kv;
            Vs[r][d] = vv;
        }
        __syncthreads();

        // S = Q K^T for rows 8 tq + i, keys tk + 16 j
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
        for (int d = 0; d < kAttD; ++d) {
            float kk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) kk[j] = Ks[tk + 16 * j][d];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float qv = Qs[8 * tq + i][d];
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv, kk[j], s[i][j]);
            }
        }
        // bias + mask + online softmax (row statistics are replicated over the 16 lanes of a row group)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = q0 + 8 * tq + i;
            int lo = (p / window - 1) * window;
            if (lo < 0) lo = 0;
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kpos = k0 + tk + 16 * j;
                const bool ok = (p < T) && (kpos <= p) && (kpos >= lo);
                s[i][j] = ok ? s[i][j] + __ldg(bt + (p - kpos)) : -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m_run[i], mx);
            float corr = 1.f, rs = 0.f;
            if (m_new == -INFINITY) {
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
            } else {
                corr = expf(m_run[i] - m_new);   // m_run = -inf -> 0
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[i][j] = expf(s[i][j] - m_new);   // masked (-inf) -> 0
                    rs += s[i][j];
                }
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
            l_run[i] = l_run[i] * corr + rs;
            m_run[i] = m_new;
            o[i][0] *= corr;
            o[i][1] *= corr;
#pragma unroll
            for (int j = 0; j < 4; ++j) Ps[8 * tq + i][tk + 16 * j] = s[i][j];
        }
        __syncthreads();
        // O += P V for rows 8 tq + i, columns 2 tk, 2 tk + 1
#pragma unroll 4
        for (int j = 0; j < kAttBK; ++j) {
            const float2 v2 = *reinterpret_cast<const float2*>(&Vs[j][2 * tk]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float pv = Ps[8 * tq + i][j];
                o[i][0] = fmaf(pv, v2.x, o[i][0]);
                o[i][1] = fmaf(pv, v2.y, o[i][1]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int p = q0 + 8 * tq + i;
        if (p < T) {
            const float inv = 1.0f / l_run[i];
            float2* dst = reinterpret_cast<float2*>(out + ((long long)b * T + p) * (H * kAttD) + h * kAttD + 2 * tk);
            *dst = make_float2(o[i][0] * inv, o[i][1] * inv);
        }
    }
}

}  // namespace l3ac

extern "C" int l3ac_local_attention_f32(const float* qkv, const float* bias_table, int B, int T, int H, int D,
                                        int window, float* out, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(qkv && bias_table && out && B > 0 && B <= 65535 && T > 0 && H > 0 && H <= 65535 && window > 0);
    if (D != l3ac::kAttD) return L3AC_EUNSUPPORTED;
    dim3 grid(l3ac_cdiv(T, l3ac::kAttBQ), H, B);
    l3ac::local_attention_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(qkv, bias_table, B, T, H, window, out);
    return l3ac_launch_status();
}
