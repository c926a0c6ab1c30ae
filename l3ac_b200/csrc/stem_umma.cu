// Encoder stem = V3FirstBlock (l3ac/tconv/__init__.py:8-27) on tcgen05 / TMEM at fp32-class precision:
//   5 x [TrendPool(k) -> Conv1d(1->4,k7,pad 3)]  (k = 1,5,11,21,45; l3ac/tconv/base.py:8-45)
//   -> Conv1d 1x1 20->80 -> exact GELU -> cat raw x -> Conv1d 1x1 81->24
// audio (B,T) -> out (B,T,24) channels-last.  Same arithmetic as stem_tc.cu (two-level pooling in shared memory, 3-term
// split-bf16 products hi*Whi + lo*Whi + hi*Wlo with fp32 accumulation, A&S erf on MUFU), but the two 1x1 convs are
// tcgen05.mma with TMEM accumulators instead of mma.sync, whose HMMAs occupied ~55 % of the issue slots of that kernel:
//   * one CTA owns 256 consecutive samples (+ 47 of context per side) as two 128-row blocks (two CTAs per SM); ONE THREAD owns one sample
//     = one TMEM lane.  All per-channel parameters (branch conv taps, biases) are kernel-parameter constants: with a row
//     per thread they are warp-uniform operands of the FFMA / FADD instructions.
//   * S1: the thread convolves the five pooled signals into its 20 branch outputs, splits them (hi, lo) and writes them as
//     operand planes [channel / 8][row][8] (umma.cuh); conv 20->80 = 3 terms x 2 K-steps of tcgen05.mma (N = 80) per block.
//   * S2: the 80 hidden columns come back 16 at a time (tcgen05.ld), + bias, GELU, split, into a two-deep ring of K = 16
//     operand chunks; conv 81->24 accumulates one K-step (3 terms, N = 32) per chunk while the thread works on the next
//     one.  The raw-x column (k = 80) is a sixth chunk.
//   * S3: + bias, 96 contiguous bytes per sample to HBM.
// Warps 0-7: row owners (warp >> 2 = block, warp & 3 = TMEM lane quadrant); warps 8-9 issue the MMAs of one block each
// (a hand-over costs the issuing warp ~300 cycles of barrier polling and commit: with one issuer for all four blocks the
// 28 hand-overs of a tile, not the MMAs, set the pace -- 19 k cycles per tile measured, tools/stem_probe.py).
// The weights are converted and uploaded once into a plan (l3ac_stem_plan).
#include "common.cuh"
#include "umma.cuh"

#include <cstring>
#include <new>
#include <vector>

namespace l3ac {
namespace stemu {

using namespace l3ac::umma;

constexpr int kBlocks = 2;
constexpr int kRows = kBlocks * 128;            // 256 samples per CTA tile; two CTAs per SM cover each other's pooling / hand-over phases
constexpr int kReach = 47;                      // 44 (max + avg pool of 45) + 3 (conv k7)
constexpr int kW = kRows + 2 * kReach;          // staged samples (350)
constexpr int kWP = 352;                        // array pitch
constexpr int kH = 80, kCin = 20, kCo = 24;
constexpr int kChunks = 6;                      // K = 16 chunks of the second conv: 5 x 16 hidden columns + the raw-x column
constexpr int kPlane = 128 * 16;                // one 8-channel plane of a 128-row block
constexpr int kRowWarps = 4 * kBlocks;
constexpr int kThreads = 32 * (kRowWarps + kBlocks);      // + one MMA issuer warp per block
// shared memory carve-up
constexpr int kPoolFloats = 14 * kWP;                               // sig[5], ax4, mx[4], s4[4]
constexpr int kOffA1 = kPoolFloats * 4;                             // [block][part][3 planes]
constexpr int kOffZero = kOffA1 + kBlocks * 2 * 3 * kPlane;         // one shared zero plane (K padding 24 -> 32)
constexpr int kOffA2 = kOffZero + kPlane;                           // [block][buf 2][part 2][2 planes]
constexpr int kOffW1 = kOffA2 + kBlocks * 2 * 2 * 2 * kPlane;       // [part][kstep 2][half 2][80][8] bf16
constexpr int kW1Bytes = 2 * 2 * 2 * kH * 16;
constexpr int kOffW2 = kOffW1 + kW1Bytes;                           // [part][chunk 6][half 2][32][8] bf16
constexpr int kW2Bytes = 2 * kChunks * 2 * 32 * 16;
constexpr int kOffBars = kOffW2 + kW2Bytes;
constexpr int kNumBars = kBlocks * (1 + 1 + 2 + 2 + 1);
constexpr int kSmemBytes = kOffBars + 8 * kNumBars + 16;
static_assert(2 * (kSmemBytes + 1024) <= 227 * 1024, "two CTAs per SM");
constexpr int kTmemCols = 128 * kBlocks;
constexpr int kD2Col = 96;                                          // TMEM columns of a block: D1 [0, 80), D2 [96, 128)

struct Params {
    const float* audio;
    float* out;
    const uint8_t* wblob;        // device: W1 (hi, lo) | W2 (hi, lo) in operand order
    int B, T;
    float bw[kCin][7], bb[kCin], b1[kH], b2[kCo];
};

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    lo = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

// exact-erf GELU on a pair, erf from Abramowitz & Stegun 7.1.26 on MUFU.RCP / MUFU.EX2 (|error| <= 1.5e-7) as in stem_tc.cu,
// rearranged to 7 FMA-pipe + 2 MUFU instructions per element: with m = 1 - erf(|z|) = poly(t) t exp(-z^2), t = 1 / (1 + p |z|),
//     gelu(x) = x/2 (1 + erf(x / sqrt 2)) = relu(x) - |x| m / 2          (both signs of x)
// and everything is expressed in z' = z sqrt(log2 e) so that exp(-z^2) = 2^(-z'^2) needs no further scaling; the 1/2 sits in
// the polynomial coefficients.  |.| and the negations are operand modifiers.
__device__ __forceinline__ float2 gelu2(float2 x) {
    constexpr float kC = 0.84932180028801904272f;                  // sqrt(log2 e) / sqrt 2
    constexpr float kP = 0.3275911f / 1.20112240878645f;           // p / sqrt(log2 e)
    const float2 zp = fmul2(x, make_float2(kC, kC));
    float2 t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(fmaf(fabsf(zp.x), kP, 1.0f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(fmaf(fabsf(zp.y), kP, 1.0f)));
    float2 p = ffma2(make_float2(0.5f * 1.061405429f, 0.5f * 1.061405429f), t, make_float2(0.5f * -1.453152027f, 0.5f * -1.453152027f));
    p = ffma2(p, t, make_float2(0.5f * 1.421413741f, 0.5f * 1.421413741f));
    p = ffma2(p, t, make_float2(0.5f * -0.284496736f, 0.5f * -0.284496736f));
    p = ffma2(p, t, make_float2(0.5f * 0.254829592f, 0.5f * 0.254829592f));
    const float2 w = fmul2(zp, zp);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(-w.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(-w.y));
    const float2 m = fmul2(fmul2(p, t), e);                                        // (1 - erf(|z|)) / 2
    return make_float2(fmaf(-fabsf(x.x), m.x, fmaxf(x.x, 0.f)), fmaf(-fabsf(x.y), m.y, fmaxf(x.y, 0.f)));
}

template <int C>
struct Chunk {
    // hidden columns 16 C .. 16 C + 15 of this thread's row: + bias, GELU, split -> two planes hi, two planes lo
    static __device__ __forceinline__ void gelu_chunk(const Params& p, uint32_t taddr, uint32_t dst) {
        uint32_t v[16];
        tmem_ld16(taddr + 16 * C, v);
        tmem_ld_wait();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 g = gelu2(make_float2(__uint_as_float(v[2 * i]) + p.b1[16 * C + 2 * i], __uint_as_float(v[2 * i + 1]) + p.b1[16 * C + 2 * i + 1]));
            split2(g.x, g.y, hi[i], lo[i]);
        }
        st_shared_v4(dst, hi[0], hi[1], hi[2], hi[3]);
        st_shared_v4(dst + kPlane, hi[4], hi[5], hi[6], hi[7]);
        st_shared_v4(dst + 2 * kPlane, lo[0], lo[1], lo[2], lo[3]);
        st_shared_v4(dst + 3 * kPlane, lo[4], lo[5], lo[6], lo[7]);
    }
};

__global__ void __launch_bounds__(kThreads, 2) stem_umma_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    float* pool = reinterpret_cast<float*>(smem);
    float* sig = pool;                       // [5][kWP]: 0 raw x; 1..4 TrendPool(k), k = 5, 11, 21, 45
    float* ax4 = pool + 5 * kWP;
    float* mx = pool + 6 * kWP;              // [4][kWP]
    float* s4 = pool + 10 * kWP;             // [4][kWP]
    const uint32_t a1_s = sbase + kOffA1, zero_s = sbase + kOffZero, a2_s = sbase + kOffA2, w1_s = sbase + kOffW1, w2_s = sbase + kOffW2;
    const uint32_t bars = sbase + kOffBars;
    const uint32_t a1_ready = bars, d1_ready = a1_ready + 8 * kBlocks, a2_full = d1_ready + 8 * kBlocks,
                   a2_empty = a2_full + 16 * kBlocks, d2_ready = a2_empty + 16 * kBlocks, tmem_slot = d2_ready + 8 * kBlocks;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - sbase));

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int tiles_per_clip = (p.T + kRows - 1) / kRows;
    const int n_tiles = tiles_per_clip * p.B;

    // ---- once per CTA: weights, the zero plane, barriers, TMEM
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.wblob);
        uint4* dst = reinterpret_cast<uint4*>(smem + kOffW1);
        for (int i = tid; i < (kW1Bytes + kW2Bytes) / 16; i += kThreads) dst[i] = __ldg(src + i);
        uint4* z = reinterpret_cast<uint4*>(smem + kOffZero);
        for (int i = tid; i < kPlane / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int b = 0; b < kBlocks; ++b) {
            mbar_init(a1_ready + 8 * b, 4);
            mbar_init(d1_ready + 8 * b, 1);
            for (int u = 0; u < 2; ++u) {
                mbar_init(a2_full + 16 * b + 8 * u, 4);
                mbar_init(a2_empty + 16 * b + 8 * u, 1);
            }
            mbar_init(d2_ready + 8 * b, 1);
        }
        fence_mbar_init();
    }
    if (warp == kRowWarps) tmem_alloc(tmem_slot, kTmemCols);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    L3AC_PDL_SYNC();      // weight staging, barriers and TMEM above may overlap the previous kernel's tail

    if (warp >= kRowWarps) {
        // =============================================================== MMA issuers: warp kRowWarps + b owns block b
        const int b = warp - kRowWarps;
        const bool leader = elect_one();
        constexpr uint64_t kDescHi = (uint64_t)(((128u >> 4) & 0x3FFF) | (1u << 14)) << 32;     // SBO = 128 B, sm_100 descriptor version
        const uint32_t idesc80 = make_idesc_bf16(kH), idesc32 = make_idesc_bf16(32);
        const uint32_t lbo_plane = (uint32_t)(kPlane >> 4) << 16, lbo_w1 = ((uint32_t)(kH * 16) >> 4) << 16, lbo_w2 = (512u >> 4) << 16;
        const uint32_t d1 = tmem_base + 128 * b, d2 = d1 + kD2Col;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            // conv 20 -> 80: D1 = A1 . W1^T, terms (hi, Whi), (lo, Whi), (hi, Wlo); K-step 1 = plane 2 + the zero plane
            mbar_wait(a1_ready + 8 * b, it & 1);
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const uint32_t a = a1_s + (b * 2 + (term == 1 ? 1 : 0)) * 3 * kPlane;
                    const uint32_t w = w1_s + (term == 2 ? kW1Bytes / 2 : 0);
                    tc_mma_bf16(d1, kDescHi | ((a >> 4) | lbo_plane), kDescHi | ((w >> 4) | lbo_w1), idesc80, term ? 1u : 0u);
                    const uint32_t a2p = a + 2 * kPlane;
                    tc_mma_bf16(d1, kDescHi | ((a2p >> 4) | (((zero_s - a2p) >> 4) << 16)), kDescHi | (((w + 2 * kH * 16) >> 4) | lbo_w1), idesc80, 1u);
                }
                tc_commit(d1_ready + 8 * b);
            }
            __syncwarp();
            // conv 81 -> 24: one K = 16 chunk at a time
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c) {
                const int u = c & 1;
                mbar_wait(a2_full + 16 * b + 8 * u, (it + (c >> 1)) & 1);
                tc_fence_after();
                if (leader) {
                    const uint32_t a = a2_s + ((b * 2 + u) * 4) * kPlane;              // hi planes 0, 1; lo planes 2, 3
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t at = a + (term == 1 ? 2 * kPlane : 0);
                        const uint32_t w = w2_s + (term == 2 ? kW2Bytes / 2 : 0) + c * 1024;
                        tc_mma_bf16(d2, kDescHi | ((at >> 4) | lbo_plane), kDescHi | ((w >> 4) | lbo_w2), idesc32, (c | term) ? 1u : 0u);
                    }
                    tc_commit(a2_empty + 16 * b + 8 * u);
                    if (c == kChunks - 1) tc_commit(d2_ready + 8 * b);
                }
                __syncwarp();
            }
        }
    } else {
        // =============================================================== row owners
        const int blk = warp >> 2, quad = warp & 3;
        const int r = blk * 128 + quad * 32 + lane;                // this thread's sample within the tile
        const uint32_t tl = tmem_base + ((uint32_t)(quad * 32) << 16) + 128 * blk;
        const uint32_t row16 = (uint32_t)((quad * 32 + lane) * 16);
        constexpr int kRowThreads = kRowWarps * 32;

        auto fetch = [&](int tile, float (&v)[2]) {                // this thread's staged samples of a tile, prefetched one tile ahead
            const int clip = tile / tiles_per_clip;
            const int t0 = (tile - clip * tiles_per_clip) * kRows;
            const float* xb = p.audio + (long long)clip * p.T;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int i = tid + s * kRowThreads;
                const int t = t0 - kReach + i;
                v[s] = (i < kW && t >= 0 && t < p.T) ? __ldg(xb + t) : 0.f;
            }
        };
        auto row_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(kRowThreads) : "memory"); };

        float nxt[2];
        int tile = blockIdx.x;
        if (tile < n_tiles) fetch(tile, nxt);
        int it = 0;
        for (; tile < n_tiles; tile += gridDim.x, ++it) {
            const int clip = tile / tiles_per_clip;
            const int t0 = (tile - clip * tiles_per_clip) * kRows;
            row_sync();                                            // the previous tile's signals are no longer read
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int i = tid + s * kRowThreads;
                if (i < kW) sig[i] = nxt[s];
            }
            row_sync();
            if (tile + (int)gridDim.x < n_tiles) fetch(tile + gridDim.x, nxt);
            // ---- TrendPool(k) = avg_pool1d(max_pool1d(|x|, k, 1, k/2), k, 1, k/2) in two levels (stem_tc.cu): max pads -inf
            // (equivalent to 0 on |x|), avg pads 0 and divides by k; positions outside the clip hold 0 (l3ac/tconv/base.py:8-14)
            for (int i = tid; i < kW; i += kRowThreads) {
                float m = 0.f;
                if (i + 3 < kW) m = fmaxf(fmaxf(fabsf(sig[i]), fabsf(sig[i + 1])), fmaxf(fabsf(sig[i + 2]), fabsf(sig[i + 3])));
                ax4[i] = m;
            }
            row_sync();
            for (int i = tid; i < kW; i += kRowThreads) {
                const int t = t0 - kReach + i;
                const bool in = t >= 0 && t < p.T;
                float m5 = 0.f, m11 = 0.f, m21 = 0.f, m45 = 0.f;
                if (in && i >= 2 && i < kW - 2) m5 = fmaxf(ax4[i - 2], ax4[i - 1]);
                if (in && i >= 5 && i < kW - 5) m11 = fmaxf(fmaxf(ax4[i - 5], ax4[i - 1]), ax4[i + 2]);
                if (in && i >= 10 && i < kW - 10) {
                    const float* a = ax4 + i - 10;
                    m21 = fmaxf(fmaxf(fmaxf(a[0], a[4]), fmaxf(a[8], a[12])), fmaxf(a[16], a[17]));
                }
                if (in && i >= 22 && i < kW - 22) {
                    const float* a = ax4 + i - 22;
                    float m = a[41];
#pragma unroll
                    for (int j = 0; j < 11; ++j) m = fmaxf(m, a[4 * j]);
                    m45 = m;
                }
                mx[i] = m5;
                mx[kWP + i] = m11;
                mx[2 * kWP + i] = m21;
                mx[3 * kWP + i] = m45;
            }
            row_sync();
            for (int i = tid; i < kW; i += kRowThreads) {
                const bool ok = i + 3 < kW;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float* m = mx + k * kWP + i;
                    s4[k * kWP + i] = ok ? (m[0] + m[1]) + (m[2] + m[3]) : 0.f;
                }
            }
            row_sync();
            for (int i = tid; i < kW; i += kRowThreads) {
                const int t = t0 - kReach + i;
                float p5 = 0.f, p11 = 0.f, p21 = 0.f, p45 = 0.f;
                if (t >= 0 && t < p.T) {                           // the branch convs zero-pad the pooled signals outside the clip
                    if (i >= 4 && i < kW - 4) p5 = (s4[i - 2] + mx[i + 2]) / 5.0f;
                    if (i >= 10 && i < kW - 10) {
                        const float* m = mx + kWP + i - 5;
                        p11 = ((s4[kWP + i - 5] + s4[kWP + i - 1]) + ((m[8] + m[9]) + m[10])) / 11.0f;
                    }
                    if (i >= 20 && i < kW - 20) {
                        const float* s = s4 + 2 * kWP + i - 10;
                        p21 = (((s[0] + s[4]) + (s[8] + s[12])) + (s[16] + mx[2 * kWP + i + 10])) / 21.0f;
                    }
                    if (i >= 44 && i < kW - 44) {
                        const float* s = s4 + 3 * kWP + i - 22;
                        float a = 0.f, b = 0.f;
#pragma unroll
                        for (int j = 0; j < 10; j += 2) {
                            a += s[4 * j];
                            b += s[4 * j + 4];
                        }
                        p45 = ((a + b) + (s[40] + mx[3 * kWP + i + 22])) / 45.0f;
                    }
                }
                sig[kWP + i] = p5;
                sig[2 * kWP + i] = p11;
                sig[3 * kWP + i] = p21;
                sig[4 * kWP + i] = p45;
            }
            row_sync();

            // ---- S1: branch convs (1 -> 4, k7) of this sample, split, operand planes of conv 20 -> 80
            const float xv = sig[kReach + r];
            {
                float hv[kCin];
#pragma unroll
                for (int c = 0; c < kCin; ++c) hv[c] = p.bb[c];
#pragma unroll
                for (int br = 0; br < 5; ++br)
#pragma unroll
                    for (int q = 0; q < 7; ++q) {
                        const float v = sig[br * kWP + kReach + r + q - 3];
#pragma unroll
                        for (int j = 0; j < 4; ++j) hv[4 * br + j] = fmaf(p.bw[4 * br + j][q], v, hv[4 * br + j]);
                    }
                uint32_t hi[12], lo[12];
#pragma unroll
                for (int i = 0; i < 10; ++i) split2(hv[2 * i], hv[2 * i + 1], hi[i], lo[i]);
                hi[10] = hi[11] = lo[10] = lo[11] = 0u;
                const uint32_t dst = a1_s + (blk * 2) * 3 * kPlane + row16;
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    st_shared_v4(dst + pl * kPlane, hi[4 * pl], hi[4 * pl + 1], hi[4 * pl + 2], hi[4 * pl + 3]);
                    st_shared_v4(dst + (3 + pl) * kPlane, lo[4 * pl], lo[4 * pl + 1], lo[4 * pl + 2], lo[4 * pl + 3]);
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a1_ready + 8 * blk);

            // ---- S2: hidden columns 16 at a time through the two-deep chunk ring
            mbar_wait(d1_ready + 8 * blk, it & 1);
            tc_fence_after();
#define L3AC_STEM_CHUNK(C)                                                                                    \
            {                                                                                                 \
                constexpr int u = (C) & 1;                                                                    \
                if ((C) >= 2) mbar_wait(a2_empty + 16 * blk + 8 * u, (it + (((C) - 2) >> 1)) & 1);            \
                Chunk<(C)>::gelu_chunk(p, tl, a2_s + ((blk * 2 + u) * 4) * kPlane + row16);                   \
                fence_async_smem();                                                                           \
                tc_fence_before();                                                                            \
                __syncwarp();                                                                                 \
                if (lane == 0) mbar_arrive(a2_full + 16 * blk + 8 * u);                                       \
            }
            L3AC_STEM_CHUNK(0)
            L3AC_STEM_CHUNK(1)
            L3AC_STEM_CHUNK(2)
            L3AC_STEM_CHUNK(3)
            L3AC_STEM_CHUNK(4)
#undef L3AC_STEM_CHUNK
            {   // chunk 5: the raw-x column (k = 80) and 15 zero columns
                constexpr int u = 1;
                mbar_wait(a2_empty + 16 * blk + 8 * u, (it + 1) & 1);
                uint32_t hi, lo;
                split2(xv, 0.f, hi, lo);
                const uint32_t dst = a2_s + ((blk * 2 + u) * 4) * kPlane + row16;
                st_shared_v4(dst, hi, 0u, 0u, 0u);
                st_shared_v4(dst + kPlane, 0u, 0u, 0u, 0u);
                st_shared_v4(dst + 2 * kPlane, lo, 0u, 0u, 0u);
                st_shared_v4(dst + 3 * kPlane, 0u, 0u, 0u, 0u);
                fence_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a2_full + 16 * blk + 8 * u);
            }

            // ---- S3: + bias, store (B, T, 24) fp32
            mbar_wait(d2_ready + 8 * blk, it & 1);
            tc_fence_after();
            {
                uint32_t v[16], w[8];
                tmem_ld16(tl + kD2Col, v);
                tmem_ld8(tl + kD2Col + 16, w);
                tmem_ld_wait();
                tc_fence_before();
                const int t = t0 + r;
                if (t < p.T) {
                    float4* o = reinterpret_cast<float4*>(p.out + ((long long)clip * p.T + t) * kCo);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        o[i] = make_float4(__uint_as_float(v[4 * i]) + p.b2[4 * i], __uint_as_float(v[4 * i + 1]) + p.b2[4 * i + 1],
                                           __uint_as_float(v[4 * i + 2]) + p.b2[4 * i + 2], __uint_as_float(v[4 * i + 3]) + p.b2[4 * i + 3]);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        o[4 + i] = make_float4(__uint_as_float(w[4 * i]) + p.b2[16 + 4 * i], __uint_as_float(w[4 * i + 1]) + p.b2[16 + 4 * i + 1],
                                               __uint_as_float(w[4 * i + 2]) + p.b2[16 + 4 * i + 2], __uint_as_float(w[4 * i + 3]) + p.b2[16 + 4 * i + 3]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kRowWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace stemu
}  // namespace l3ac

struct l3ac_stem_plan {
    l3ac::stemu::Params params;
    void* dev_blob;
    int device;
};

extern "C" int l3ac_stem_plan_create(const float* branch_w, const float* branch_b, const float* w1, const float* b1,
                                     const float* w2, const float* b2, int C, l3ac_stem_plan** plan_out) {
    using namespace l3ac::stemu;
    L3AC_CHECK_ARG(branch_w && branch_b && w1 && b1 && w2 && b2 && plan_out);
    if (C != kCo) return L3AC_EUNSUPPORTED;
    auto bits = [](float v) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        uint16_t b;
        memcpy(&b, &h, 2);
        return b;
    };
    auto rnd = [](float v) { return __bfloat162float(__float2bfloat16_rn(v)); };
    std::vector<uint16_t> blob((kW1Bytes + kW2Bytes) / 2, 0);
    // W1: [part][kstep][half][n 80][k 8]; channel = 16 kstep + 8 half + k (< 20); w1[n][channel]
    for (int part = 0; part < 2; ++part)
        for (int ks = 0; ks < 2; ++ks)
            for (int h = 0; h < 2; ++h)
                for (int n = 0; n < kH; ++n)
                    for (int k = 0; k < 8; ++k) {
                        const int ch = 16 * ks + 8 * h + k;
                        if (ch >= kCin) continue;
                        const float w = w1[n * kCin + ch], hi = rnd(w);
                        blob[(size_t)part * (kW1Bytes / 4) + ((ks * 2 + h) * kH + n) * 8 + k] = bits(part == 0 ? hi : w - hi);
                    }
    // W2: [part][chunk][half][n 32][k 8]; hidden = 16 chunk + 8 half + k (chunk < 5), chunk 5: k = 0 of half 0 is the raw-x column
    const size_t w2o = kW1Bytes / 2;
    for (int part = 0; part < 2; ++part)
        for (int c = 0; c < kChunks; ++c)
            for (int h = 0; h < 2; ++h)
                for (int n = 0; n < kCo; ++n)
                    for (int k = 0; k < 8; ++k) {
                        int col;
                        if (c < 5) col = 16 * c + 8 * h + k;
                        else if (h == 0 && k == 0) col = kH;
                        else continue;
                        const float w = w2[n * (kH + 1) + col], hi = rnd(w);
                        blob[w2o + (size_t)part * (kW2Bytes / 4) + ((c * 2 + h) * 32 + n) * 8 + k] = bits(part == 0 ? hi : w - hi);
                    }
    l3ac_stem_plan* plan = new (std::nothrow) l3ac_stem_plan();
    if (!plan) return L3AC_EINVAL;
    if (cudaGetDevice(&plan->device) != cudaSuccess) { delete plan; return L3AC_EDRIVER; }
    cudaError_t e = cudaMalloc(&plan->dev_blob, kW1Bytes + kW2Bytes);
    if (e != cudaSuccess) { delete plan; return (int)e; }
    e = cudaMemcpy(plan->dev_blob, blob.data(), kW1Bytes + kW2Bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(plan->dev_blob); delete plan; return (int)e; }
    Params& p = plan->params;
    p = Params{};
    p.wblob = static_cast<const uint8_t*>(plan->dev_blob);
    for (int c = 0; c < kCin; ++c) {
        p.bb[c] = branch_b[c];
        for (int q = 0; q < 7; ++q) p.bw[c][q] = branch_w[c * 7 + q];
    }
    for (int i = 0; i < kH; ++i) p.b1[i] = b1[i];
    for (int i = 0; i < kCo; ++i) p.b2[i] = b2[i];
    *plan_out = plan;
    return L3AC_OK;
}

extern "C" int l3ac_stem_plan_destroy(l3ac_stem_plan* plan) {
    if (!plan) return L3AC_OK;
    cudaFree(plan->dev_blob);
    delete plan;
    return L3AC_OK;
}

extern "C" int l3ac_stem_umma(const l3ac_stem_plan* plan, const float* audio, int B, int T, float* out, l3ac_stream_t stream) {
    using namespace l3ac::stemu;
    L3AC_CHECK_ARG(plan && audio && out && B > 0 && T > 0);
    L3AC_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return L3AC_EDRIVER;
    L3AC_CHECK_ARG(dev == plan->device);
    const long long n_tiles = (long long)l3ac_cdiv(T, kRows) * B;
    L3AC_CHECK_ARG(n_tiles < (1LL << 30));
    Params p = plan->params;
    p.audio = audio;
    p.out = out;
    p.B = B;
    p.T = T;
    cudaError_t e = cudaFuncSetAttribute(stem_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    const int sms = l3ac_sm_count();
    const long long ctas = 2LL * sms;
    l3ac_launch(stem_umma_kernel, dim3((int)(n_tiles < ctas ? n_tiles : ctas)), dim3(kThreads), kSmemBytes, (cudaStream_t)stream, p);
    return l3ac_launch_status();
}
