// fp32 SIMT GEMM / conv-as-GEMM with fused epilogue (reference-precision path; see l3ac_b200.h).
// out[m,n] = epi(bias[n] + sum_s sum_k A[b, t + shift_s, k] * W[n, s*K + k]),  m = b*T + t.
// Tile 128 x BN x 16, 256 threads; thread (ty, tx) owns rows ty + 16 i and column pairs
// (2 tx + 32 j, +1) so that (value, gate) pairs of the GEGLU epilogue stay inside one thread.
#include "common.cuh"

namespace l3ac {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 16;

struct EpiParams {
    const float* bias;
    const float* alpha;
    const float* scale;
    const float* shift;
    const float* residual;
    long long ldr, ldo;
    int act;
};

__device__ __forceinline__ float epi_scalar(float v, int n, const EpiParams& e) {
    if (e.bias) v += __ldg(e.bias + n);
    if (e.act == L3AC_ACT_SNAKE) {
        const float a = __ldg(e.alpha + n);
        v = snake_f(v, a, 1.0f / (a + kEps));
        if (e.scale) v = fmaf(v, __ldg(e.scale + n), __ldg(e.shift + n));
    } else if (e.act == L3AC_ACT_GELU) {
        v = gelu_erf(v);
    } else if (e.act == L3AC_ACT_TANH) {
        v = tanhf(v);
    }
    return v;
}

template <int BN, typename OutT>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                       long long lda, int B, int T, int K, int N, int taps,
                                                       int tap_shift0, int tap_step, EpiParams e,
                                                       OutT* __restrict__ out) {
    constexpr int BM = kGemmBM, BK = kGemmBK, TN = BN / 16, TM = BM / 16;
    __shared__ float As[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][BN + 4];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long M = (long long)B * T;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int ldw = taps * K;

    // Two-level accumulation: 64 products into `acc`, blocks into `tot`.  A single running fp32 sum over K*taps (up to 2048)
    // terms carries ~sqrt(K) ulps of rounding noise, several times what the reference's blocked / vectorised CPU GEMM has;
    // with 64-term blocks the parity mode sits at the reference's own distance from exact arithmetic (tests/golden/fp32_floor.json).
    float acc[TM][TN], tot[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = tot[i][j] = 0.f;
    int blk = 0;

    // loader mapping: k = tid % 16, rows tid/16 + 16 r
    const int lk = tid & 15, lr = tid >> 4;

    for (int s = 0; s < taps; ++s) {
        const int shift = tap_shift0 + s * tap_step;
        for (int k0 = 0; k0 < K; k0 += BK) {
            const int k = k0 + lk;
#pragma unroll
            for (int r = 0; r < BM / 16; ++r) {
                const int row = lr + 16 * r;
                const long long m = m0 + row;
                float v = 0.f;
                if (m < M && k < K) {
                    const int t = (int)(m % T) + shift;
                    if (t >= 0 && t < T) v = __ldg(A + (m + shift) * lda + k);
                }
                As[lk][row] = v;
            }
#pragma unroll
            for (int r = 0; r < BN / 16; ++r) {
                const int col = lr + 16 * r;
                const int n = n0 + col;
                Ws[lk][col] = (n < N && k < K) ? __ldg(W + (long long)n * ldw + s * K + k) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float a[TM], w[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
                for (int j = 0; j < TN / 2; ++j) {
                    const float2 w2 = *reinterpret_cast<const float2*>(&Ws[kk][2 * tx + 32 * j]);
                    w[2 * j] = w2.x;
                    w[2 * j + 1] = w2.y;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
            __syncthreads();
            if ((++blk & 3) == 0) {
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        tot[i][j] += acc[i][j];
                        acc[i][j] = 0.f;
                    }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] += tot[i][j];

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long m = m0 + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) {
            const int n = n0 + 2 * tx + 32 * j;
            if (n >= N) continue;
            if (e.act == L3AC_ACT_GEGLU) {
                float v = acc[i][2 * j], g = acc[i][2 * j + 1];
                if (e.bias) {
                    v += __ldg(e.bias + n);
                    g += __ldg(e.bias + n + 1);
                }
                float r = v * gelu_erf(g);
                const int no = n >> 1;
                if (e.residual) r += e.residual[m * e.ldr + no];
                out[m * e.ldo + no] = cvt_out<OutT>(r);
            } else {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (n + q >= N) continue;
                    float v = epi_scalar(acc[i][2 * j + q], n + q, e);
                    if (e.residual) v += e.residual[m * e.ldr + n + q];
                    out[m * e.ldo + n + q] = cvt_out<OutT>(v);
                }
            }
        }
    }
}

}  // namespace l3ac

using namespace l3ac;

template <int BN>
static int launch_gemm_f32(const l3ac_gemm_desc* d, cudaStream_t st) {
    const long long M = (long long)d->B * d->T;
    dim3 grid((unsigned)((M + kGemmBM - 1) / kGemmBM), (unsigned)((d->N + BN - 1) / BN));
    EpiParams e{d->bias, d->alpha, d->scale, d->shift, d->residual, d->ldr, d->ldo, d->act};
    if (d->out_dtype == L3AC_F32)
        gemm_f32_kernel<BN, float><<<grid, 256, 0, st>>>((const float*)d->A, (const float*)d->W, d->lda, d->B, d->T,
                                                         d->K, d->N, d->taps, d->tap_shift0, d->tap_step, e,
                                                         (float*)d->out);
    else
        gemm_f32_kernel<BN, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)d->A, (const float*)d->W, d->lda, d->B,
                                                                 d->T, d->K, d->N, d->taps, d->tap_shift0,
                                                                 d->tap_step, e, (__nv_bfloat16*)d->out);
    return l3ac_launch_status();
}

int l3ac_validate_gemm_desc(const l3ac_gemm_desc* d) {
    L3AC_CHECK_ARG(d && d->A && d->W && d->out);
    L3AC_CHECK_ARG(d->B > 0 && d->T > 0 && d->K > 0 && d->N > 0 && d->taps >= 1);
    L3AC_CHECK_ARG(d->lda >= d->K);
    L3AC_CHECK_ARG(d->out_dtype == L3AC_F32 || d->out_dtype == L3AC_BF16 || d->out_dtype == L3AC_BF16X2);
    L3AC_CHECK_ARG(d->act >= L3AC_ACT_NONE && d->act <= L3AC_ACT_TANH);
    if (d->act == L3AC_ACT_SNAKE) {
        L3AC_CHECK_ARG(d->alpha != nullptr);
        L3AC_CHECK_ARG((d->scale == nullptr) == (d->shift == nullptr));
    }
    if (d->act == L3AC_ACT_GEGLU) L3AC_CHECK_ARG(d->N % 2 == 0);
    const int n_out = d->act == L3AC_ACT_GEGLU ? d->N / 2 : d->N;
    L3AC_CHECK_ARG(d->ldo >= n_out);
    if (d->residual) L3AC_CHECK_ARG(d->ldr >= n_out);
    return L3AC_OK;
}

extern "C" int l3ac_gemm_f32(const l3ac_gemm_desc* d, l3ac_stream_t stream) {
    const int rc = l3ac_validate_gemm_desc(d);
    if (rc != L3AC_OK) return rc;
    if (d->A_lo || d->W_lo || d->out_dtype == L3AC_BF16X2) return L3AC_EUNSUPPORTED;   // split pairs are a tcgen05-path feature
    const long long M = (long long)d->B * d->T;
    L3AC_CHECK_ARG((M + kGemmBM - 1) / kGemmBM <= 2147483647LL);
    cudaStream_t st = (cudaStream_t)stream;
    if (d->N <= 32) return launch_gemm_f32<32>(d, st);
    if (d->N <= 64 || d->N == 192) return launch_gemm_f32<64>(d, st);
    return launch_gemm_f32<128>(d, st);
}
