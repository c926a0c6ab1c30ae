// Encoder stem = V3FirstBlock (l3ac/tconv/__init__.py:8-27) on the tensor cores at fp32-class precision:
//   5 x [TrendPool(k) -> Conv1d(1->4,k7,pad 3)]  (k = 1,5,11,21,45; l3ac/tconv/base.py:8-45)
//   -> Conv1d 1x1 20->80 -> exact GELU -> cat raw x -> Conv1d 1x1 81->24
// audio (B,T) -> out (B,T,24) channels-last.  The fp32 SIMT stem (stem.cu) spends 3.5 kMAC of FFMA per sample on the two
// 1x1 convs; here they are 3-term split-bf16 MMAs (hi*Whi + lo*Whi + hi*Wlo, fp32 accumulate -- the scheme of the
// split tcgen05 GEMMs and of convunit_tc_split.cu), with everything between them in registers:
//   * the pooled signals of a 256-sample tile are built in shared memory in four passes (max over 4 neighbours, the
//     window maxima from those, sums over 4 neighbours, the window sums from those): ~70 reads per sample instead of
//     the 330 of the direct double loop,
//   * every warp owns 2 x 16 samples; the 20 branch-conv outputs are computed straight into the mma A-fragment layout
//     (a thread needs 3 channel pairs of 2 rows), split, and multiplied by W1 (K = 20 as a k16 + a k8 step),
//   * the 80 hidden columns are produced 16 at a time: GELU on the accumulators, split, and the accumulator registers are
//     the A fragments of the second conv, accumulated over the 5 chunks; the raw-x column (k = 80) is one FMA per output.
// Persistent CTAs: the split weight fragments are staged once, the next tile's samples are prefetched into registers.
#include "common.cuh"

namespace l3ac {
namespace stemtc {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kMT = 2;                          // 16-sample m-tiles per warp
constexpr int kTile = kWarps * kMT * 16;        // 256 samples per CTA tile
constexpr int kReach = 47;                      // 44 (max + avg pool of 45) + 3 (conv k7)
constexpr int kW = kTile + 2 * kReach;          // staged samples
constexpr int kWP = 352;                        // array pitch (>= kW)
constexpr int kH = 80, kCin = 20, kCo = 24;
constexpr int kW1Vec = 2 * (kH / 8) * 32;       // uint4 {hi.b0, hi.b1, lo.b0, lo.b1} per (k-step, n-tile, lane)
constexpr int kW2Vec = (kH / 16) * (kCo / 8) * 32;

struct Smem {
    float sig[5][kWP];       // 0: raw x; 1..4: TrendPool(k) of it, k = 5, 11, 21, 45
    float ax4[kWP];          // max(|x[i..i+3]|)
    float mx[4][kWP];        // window maxima
    float s4[4][kWP];        // mx[i] + ... + mx[i+3]
    uint4 w1f[kW1Vec];
    uint4 w2f[kW2Vec];
    float2 bw2[7][12];       // branch conv weights of channel pair c/2 at tap q (pairs 10, 11 are zero padding)
    float2 bb2[12];
    float b1[kH];
    float2 b2[kCo / 2], wx[kCo / 2];     // second conv: bias and the raw-x column w2[:, 80]
};

__device__ __forceinline__ void mma_k16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(b0));
}

__device__ __forceinline__ void split2(float2 v, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v.x - hf.x, v.y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// split B fragment of W[n][k] (row stride ld, k < kmax real): lane l of (k-step ks, n-tile nt)
__device__ __forceinline__ uint4 split_b_frag(const float* __restrict__ w, int ld, int kmax, int ks, int nt, int l) {
    const int n = nt * 8 + (l >> 2), k0 = ks * 16 + (l & 3) * 2;
    const int ko[4] = {k0, k0 + 1, k0 + 8, k0 + 9};
    float hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float v = ko[j] < kmax ? __ldg(w + n * ld + ko[j]) : 0.f;
        hi[j] = __bfloat162float(__float2bfloat16_rn(v));
        lo[j] = v - hi[j];
    }
    uint4 r;
    uint32_t d;
    split2(make_float2(hi[0], hi[1]), r.x, d);
    split2(make_float2(hi[2], hi[3]), r.y, d);
    split2(make_float2(lo[0], lo[1]), r.z, d);
    split2(make_float2(lo[2], lo[3]), r.w, d);
    return r;
}

// exact-erf GELU on a pair (erf from Abramowitz & Stegun 7.1.26 on MUFU.RCP / MUFU.EX2, |error| <= 1.5e-7: the formula
// of gelu_erf_fast in common.cuh with the polynomial in packed fp32)
__device__ __forceinline__ float2 gelu2(float2 x) {
    const float2 z = fmul2(x, make_float2(0.70710678118654752440f, 0.70710678118654752440f));
    const float2 a = make_float2(fabsf(z.x), fabsf(z.y));
    const float2 d = ffma2(a, make_float2(0.3275911f, 0.3275911f), make_float2(1.0f, 1.0f));
    float2 t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
    float2 p = ffma2(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
    p = ffma2(p, t, make_float2(1.421413741f, 1.421413741f));
    p = ffma2(p, t, make_float2(-0.284496736f, -0.284496736f));
    p = ffma2(p, t, make_float2(0.254829592f, 0.254829592f));
    const float2 q = fmul2(fmul2(a, a), make_float2(-1.4426950408889634f, -1.4426950408889634f));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
    const float2 m = fmul2(fmul2(p, t), e);                                        // 1 - erf(|z|)
    const float2 erf_abs = make_float2(1.0f - m.x, 1.0f - m.y);
    const float2 s = make_float2(1.0f + copysignf(erf_abs.x, z.x), 1.0f + copysignf(erf_abs.y, z.y));
    return fmul2(fmul2(x, make_float2(0.5f, 0.5f)), s);
}

__global__ void __launch_bounds__(kThreads, 2) stem_tc_kernel(const float* __restrict__ audio, int B, int T,
                                                              const float* __restrict__ branch_w, const float* __restrict__ branch_b,
                                                              const float* __restrict__ w1, const float* __restrict__ b1,
                                                              const float* __restrict__ w2, const float* __restrict__ b2,
                                                              float* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int tiles_per_clip = (T + kTile - 1) / kTile;
    const int n_tiles = tiles_per_clip * B;

    // ---- once per CTA: split weight fragments and parameters
    for (int i = tid; i < kW1Vec; i += kThreads)            // conv 20 -> 80: B[k = channel][n = hidden] = w1[n][k]
        sm.w1f[i] = split_b_frag(w1, kCin, kCin, (i >> 5) / (kH / 8), (i >> 5) % (kH / 8), i & 31);
    for (int i = tid; i < kW2Vec; i += kThreads)            // conv 81 -> 24 without its raw-x column: B[k = hidden][n = channel] = w2[n][k]
        sm.w2f[i] = split_b_frag(w2, kH + 1, kH, (i >> 5) / (kCo / 8), (i >> 5) % (kCo / 8), i & 31);
    for (int i = tid; i < 7 * 12; i += kThreads) {
        const int q = i / 12, cp = i - q * 12;
        sm.bw2[q][cp] = cp < 10 ? make_float2(__ldg(branch_w + (2 * cp) * 7 + q), __ldg(branch_w + (2 * cp + 1) * 7 + q)) : make_float2(0.f, 0.f);
    }
    if (tid < 12) sm.bb2[tid] = tid < 10 ? make_float2(__ldg(branch_b + 2 * tid), __ldg(branch_b + 2 * tid + 1)) : make_float2(0.f, 0.f);
    if (tid < kH) sm.b1[tid] = __ldg(b1 + tid);
    if (tid < kCo / 2) {
        sm.b2[tid] = make_float2(__ldg(b2 + 2 * tid), __ldg(b2 + 2 * tid + 1));
        sm.wx[tid] = make_float2(__ldg(w2 + (2 * tid) * (kH + 1) + kH), __ldg(w2 + (2 * tid + 1) * (kH + 1) + kH));
    }

    // this thread's staged samples i = tid and tid + kThreads of a tile, prefetched one tile ahead
    auto fetch = [&](int tile, float (&v)[2]) {
        const int clip = tile / tiles_per_clip;
        const int t0 = (tile - clip * tiles_per_clip) * kTile;
        const float* xb = audio + (long long)clip * T;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int i = tid + s * kThreads;
            const int t = t0 - kReach + i;
            v[s] = (i < kW && t >= 0 && t < T) ? __ldg(xb + t) : 0.f;
        }
    };
    float nxt[2];
    int tile = blockIdx.x;
    if (tile < n_tiles) fetch(tile, nxt);

    // channel pairs of this thread in the A-fragment layout and the pooled signal each pair convolves
    const int cpair[3] = {t4, 4 + t4, 8 + t4};                       // channels 2 t4, 8 + 2 t4, 16 + 2 t4 (+1)
    const int branch[3] = {t4 >> 1, 2 + (t4 >> 1), 4};

    for (; tile < n_tiles; tile += gridDim.x) {
        const int clip = tile / tiles_per_clip;
        const int t0 = (tile - clip * tiles_per_clip) * kTile;
        __syncthreads();                                   // the previous tile's signals are no longer read (first pass: set-up visible)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int i = tid + s * kThreads;
            if (i < kW) sm.sig[0][i] = nxt[s];
        }
        __syncthreads();
        if (tile + (int)gridDim.x < n_tiles) fetch(tile + gridDim.x, nxt);
        // ---- TrendPool(k) = avg_pool1d(max_pool1d(|x|, k, 1, k/2), k, 1, k/2): max pads -inf (equivalent to 0 on |x|), avg pads 0
        // and divides by k; positions outside the clip hold 0 (l3ac/tconv/base.py:8-14).
        for (int i = tid; i < kW; i += kThreads) {
            float m = 0.f;
            if (i + 3 < kW) m = fmaxf(fmaxf(fabsf(sm.sig[0][i]), fabsf(sm.sig[0][i + 1])), fmaxf(fabsf(sm.sig[0][i + 2]), fabsf(sm.sig[0][i + 3])));
            sm.ax4[i] = m;
        }
        __syncthreads();
        for (int i = tid; i < kW; i += kThreads) {
            const int t = t0 - kReach + i;
            const bool in = t >= 0 && t < T;
            float m5 = 0.f, m11 = 0.f, m21 = 0.f, m45 = 0.f;
            if (in && i >= 2 && i < kW - 2) m5 = fmaxf(sm.ax4[i - 2], sm.ax4[i - 1]);
            if (in && i >= 5 && i < kW - 5) m11 = fmaxf(fmaxf(sm.ax4[i - 5], sm.ax4[i - 1]), sm.ax4[i + 2]);
            if (in && i >= 10 && i < kW - 10) {
                const float* a = sm.ax4 + i - 10;
                m21 = fmaxf(fmaxf(fmaxf(a[0], a[4]), fmaxf(a[8], a[12])), fmaxf(a[16], a[17]));
            }
            if (in && i >= 22 && i < kW - 22) {
                const float* a = sm.ax4 + i - 22;
                float m = a[41];
#pragma unroll
                for (int j = 0; j < 11; ++j) m = fmaxf(m, a[4 * j]);
                m45 = m;
            }
            sm.mx[0][i] = m5;
            sm.mx[1][i] = m11;
            sm.mx[2][i] = m21;
            sm.mx[3][i] = m45;
        }
        __syncthreads();
        for (int i = tid; i < kW; i += kThreads) {
            const bool ok = i + 3 < kW;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float* m = sm.mx[k] + i;
                sm.s4[k][i] = ok ? (m[0] + m[1]) + (m[2] + m[3]) : 0.f;
            }
        }
        __syncthreads();
        for (int i = tid; i < kW; i += kThreads) {
            const int t = t0 - kReach + i;
            float p5 = 0.f, p11 = 0.f, p21 = 0.f, p45 = 0.f;
            if (t < 0 || t >= T) {                         // the branch convs zero-pad the pooled signals outside the clip
            } else {
            if (i >= 4 && i < kW - 4) p5 = (sm.s4[0][i - 2] + sm.mx[0][i + 2]) / 5.0f;
            if (i >= 10 && i < kW - 10) {
                const float* m = sm.mx[1] + i - 5;
                p11 = ((sm.s4[1][i - 5] + sm.s4[1][i - 1]) + ((m[8] + m[9]) + m[10])) / 11.0f;
            }
            if (i >= 20 && i < kW - 20) {
                const float* s = sm.s4[2] + i - 10;
                p21 = (((s[0] + s[4]) + (s[8] + s[12])) + (s[16] + sm.mx[2][i + 10])) / 21.0f;
            }
            if (i >= 44 && i < kW - 44) {
                const float* s = sm.s4[3] + i - 22;
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int j = 0; j < 10; j += 2) {
                    a += s[4 * j];
                    b += s[4 * j + 4];
                }
                p45 = ((a + b) + (s[40] + sm.mx[3][i + 22])) / 45.0f;
            }
            }
            sm.sig[1][i] = p5;
            sm.sig[2][i] = p11;
            sm.sig[3][i] = p21;
            sm.sig[4][i] = p45;
        }
        __syncthreads();

        // ---- branch convs (1 -> 4, k7) into the A-fragment layout, split
        const int li = kReach + warp * (kMT * 16) + g;         // + 16 i + 8 h: this thread's samples inside the staged arrays
        uint32_t ahi[kMT][3][2], alo[kMT][3][2];
        {
            float2 hv[kMT][2][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float2 bb = sm.bb2[cpair[c]];
#pragma unroll
                for (int i = 0; i < kMT; ++i) hv[i][0][c] = hv[i][1][c] = bb;
            }
#pragma unroll
            for (int q = 0; q < 7; ++q)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float2 w = sm.bw2[q][cpair[c]];
                    const float* src = sm.sig[branch[c]] + li + q - 3;
#pragma unroll
                    for (int i = 0; i < kMT; ++i)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float v = src[i * 16 + h * 8];
                            hv[i][h][c] = ffma2(w, make_float2(v, v), hv[i][h][c]);
                        }
                }
#pragma unroll
            for (int i = 0; i < kMT; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int c = 0; c < 3; ++c) split2(hv[i][h][c], ahi[i][c][h], alo[i][c][h]);
        }

        // ---- the two 1x1 convs, 16 hidden columns at a time
        float acc[kMT][3][4];
#pragma unroll
        for (int i = 0; i < kMT; ++i)
#pragma unroll
            for (int n = 0; n < 3; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[i][n][e] = 0.f;
#pragma unroll 1
        for (int hc = 0; hc < kH / 16; ++hc) {
            float hacc[kMT][2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const float2 bv = *reinterpret_cast<const float2*>(sm.b1 + hc * 16 + nt * 8 + t4 * 2);
#pragma unroll
                for (int i = 0; i < kMT; ++i) {
                    hacc[i][nt][0] = hacc[i][nt][2] = bv.x;
                    hacc[i][nt][1] = hacc[i][nt][3] = bv.y;
                }
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const uint4 wa = sm.w1f[(hc * 2 + nt) * 32 + lane];                       // k-step 0: channels 0..15
                const uint4 wb = sm.w1f[((kH / 8) + hc * 2 + nt) * 32 + lane];            // k-step 1: channels 16..19 (k8)
#pragma unroll
                for (int i = 0; i < kMT; ++i) {
                    mma_k16(hacc[i][nt], ahi[i][0][0], ahi[i][0][1], ahi[i][1][0], ahi[i][1][1], wa.x, wa.y);
                    mma_k16(hacc[i][nt], alo[i][0][0], alo[i][0][1], alo[i][1][0], alo[i][1][1], wa.x, wa.y);
                    mma_k16(hacc[i][nt], ahi[i][0][0], ahi[i][0][1], ahi[i][1][0], ahi[i][1][1], wa.z, wa.w);
                    mma_k8(hacc[i][nt], ahi[i][2][0], ahi[i][2][1], wb.x);
                    mma_k8(hacc[i][nt], alo[i][2][0], alo[i][2][1], wb.x);
                    mma_k8(hacc[i][nt], ahi[i][2][0], ahi[i][2][1], wb.z);
                }
            }
            uint32_t ghi[kMT][2][2], glo[kMT][2][2];
#pragma unroll
            for (int i = 0; i < kMT; ++i)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        split2(gelu2(make_float2(hacc[i][nt][2 * h], hacc[i][nt][2 * h + 1])), ghi[i][nt][h], glo[i][nt][h]);
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                const uint4 w = sm.w2f[(hc * 3 + n) * 32 + lane];
#pragma unroll
                for (int i = 0; i < kMT; ++i) {
                    mma_k16(acc[i][n], ghi[i][0][0], ghi[i][0][1], ghi[i][1][0], ghi[i][1][1], w.x, w.y);
                    mma_k16(acc[i][n], glo[i][0][0], glo[i][0][1], glo[i][1][0], glo[i][1][1], w.x, w.y);
                    mma_k16(acc[i][n], ghi[i][0][0], ghi[i][0][1], ghi[i][1][0], ghi[i][1][1], w.z, w.w);
                }
            }
        }

        // ---- + bias + raw-x column, store (B, T, 24) fp32
#pragma unroll
        for (int i = 0; i < kMT; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = warp * (kMT * 16) + i * 16 + h * 8 + g;
                const int t = t0 + r;
                if (t >= T) continue;
                const float xv = sm.sig[0][kReach + r];
                float* o = out + ((long long)clip * T + t) * kCo + t4 * 2;
#pragma unroll
                for (int n = 0; n < 3; ++n) {
                    const float2 v = ffma2(sm.wx[n * 4 + t4], make_float2(xv, xv),
                                           fadd2(make_float2(acc[i][n][2 * h], acc[i][n][2 * h + 1]), sm.b2[n * 4 + t4]));
                    *reinterpret_cast<float2*>(o + n * 8) = v;
                }
            }
    }
}

}  // namespace stemtc
}  // namespace l3ac

extern "C" int l3ac_stem_tc(const float* audio, int B, int T, const float* branch_w, const float* branch_b,
                            const float* w1, const float* b1, const float* w2, const float* b2, int C, float* out,
                            l3ac_stream_t stream) {
    using namespace l3ac::stemtc;
    L3AC_CHECK_ARG(audio && branch_w && branch_b && w1 && b1 && w2 && b2 && out);
    L3AC_CHECK_ARG(B > 0 && T > 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0);
    if (C != kCo) return L3AC_EUNSUPPORTED;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return L3AC_EDRIVER;
    cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return (int)e;
    const long long n_tiles = (long long)l3ac_cdiv(T, kTile) * B;
    const long long ctas = 2ll * sms;
    stem_tc_kernel<<<(int)(n_tiles < ctas ? n_tiles : ctas), kThreads, sizeof(Smem), (cudaStream_t)stream>>>(
        audio, B, T, branch_w, branch_b, w1, b1, w2, b2, out);
    return l3ac_launch_status();
}
