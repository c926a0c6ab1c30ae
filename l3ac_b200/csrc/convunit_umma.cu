// Whole Residual(ConvUnit) (l3ac/modules.py:10-44) of the thin encode-side stages (C = 24 at the full sample rate, C = 48)
// on tcgen05 / TMEM at fp32-class precision:
//   x + pw_conv2( GRN( Snake( pw_conv1( LayerNorm( dwconv7(x) ) ) ) ) )
// Same arithmetic as convunit_tc_split.cu (3-term split-bf16 products hi*Whi + lo*Whi + hi*Wlo, fp32 accumulation, GRN folded
// to a per-channel affine, MUFU sine), with the two point-wise convs as tcgen05.mma instead of mma.sync; structure of
// stem_umma.cu:
//   * a CTA owns 256 consecutive time steps as two 128-row blocks; the fp32 x tile (+ 3 rows of context per side, zero outside
//     the clip) is staged in shared memory with cp.async; ONE THREAD owns one time step = one TMEM lane, so the LayerNorm
//     statistics are thread-local and every per-channel parameter is a warp-uniform kernel-parameter constant;
//   * S1: dwconv7 + LayerNorm of the row, split (hi, lo), operand planes [channel / 8][row][8] (umma.cuh);
//     pw_conv1 = 3 terms x ceil(C / 16) K-steps of tcgen05.mma (N = 4C) into TMEM;
//   * S2: the 4C hidden columns come back 16 at a time: + bias, snake, GRN affine, split, into a two-deep ring of K = 16
//     operand chunks; pw_conv2 accumulates one K-step (3 terms) per chunk while the thread works on the next one;
//   * S3: + bias + residual (from the staged tile) -> fp32 rows or the split pair the strided down-conv GEMM consumes.
// Warps 0-7: row owners; warps 8-9: one MMA issuer per block.  C = 24: 256 TMEM columns and ~107 KB of shared memory, two
// CTAs per SM; C = 48: 512 columns, one CTA per SM.  Weights are converted and uploaded once into a plan.
#include "common.cuh"
#include "umma.cuh"

#include <cstring>
#include <new>
#include <vector>

namespace l3ac {
namespace cuu {

using namespace l3ac::umma;

constexpr int kBlocks = 2;
constexpr int kRows = kBlocks * 128;
constexpr int kPlane = 128 * 16;
constexpr int kRowWarps = 4 * kBlocks;
constexpr int kRowThreads = 32 * kRowWarps;
constexpr int kThreads = 32 * (kRowWarps + kBlocks);

template <int C>
struct Cfg {
    static_assert(C == 24 || C == 48, "thin stages only");
    static constexpr int kH = 4 * C;
    static constexpr int kPlanes1 = C / 8;                       // operand planes of pw_conv1's A
    static constexpr int kSteps1 = (C + 15) / 16;                // K = 16 steps (C = 24: the second one is plane 2 + the zero plane)
    static constexpr int kChunks = kH / 16;
    static constexpr int kN2 = C == 24 ? 32 : 48;                // UMMA N of pw_conv2 (a multiple of 16)
    static constexpr int kBlkCols = C == 24 ? 128 : 256;         // TMEM columns per block: D1 [0, 4C), D2 [kD2, kD2 + kN2)
    static constexpr int kD2 = C == 24 ? 96 : 192;
    static constexpr int kTileRows = kRows + 6;
    static constexpr int kPitch = C + 4;                         // floats per staged row: 16-byte row reads of 8 consecutive rows hit distinct banks
    static constexpr int kOffA1 = kTileRows * kPitch * 4;                              // after the x tile
    static constexpr int kOffZero = kOffA1 + kBlocks * 2 * kPlanes1 * kPlane;
    static constexpr int kOffA2 = kOffZero + kPlane;
    static constexpr int kOffW1 = kOffA2 + kBlocks * 2 * 4 * kPlane;
    static constexpr int kW1Bytes = 2 * kSteps1 * 2 * kH * 16;                         // [part][kstep][half][4C][8]
    static constexpr int kOffW2 = kOffW1 + kW1Bytes;
    static constexpr int kW2Bytes = 2 * kChunks * 2 * kN2 * 16;                        // [part][chunk][half][kN2][8]
    static constexpr int kOffBars = kOffW2 + kW2Bytes;
    static constexpr int kSmemBytes = kOffBars + 8 * kBlocks * 7 + 16;
    static constexpr int kCtasPerSm = C == 24 ? 2 : 1;
    static_assert(kCtasPerSm * (kSmemBytes + 1024) <= 227 * 1024, "shared memory budget");
    static_assert(kOffA1 % 16 == 0, "operand planes are 16-byte aligned (no-swizzle descriptors address 16-byte units)");
};

template <int C>
struct Params {
    const float* x;
    void* out;
    void* out_lo;
    const uint8_t* wblob;
    int B, T, out_split;
    float eps;
    float dw_w[7][C], dw_b[C], ln_w[C], ln_b[C], b1[4 * C], alpha[4 * C], ialpha[4 * C], scale[4 * C], shift[4 * C], b2[C];
};

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    lo = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

template <int C, int CH>
struct Chunk {
    // hidden columns 16 CH .. + 15 of this thread's row: + bias, snake, GRN affine, split -> two planes hi, two planes lo
    static __device__ __forceinline__ void run(const Params<C>& p, uint32_t taddr, uint32_t dst) {
        uint32_t v[16];
        tmem_ld16(taddr + 16 * CH, v);
        tmem_ld_wait();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float r[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = 16 * CH + 2 * i + j;
                const float u = __uint_as_float(v[2 * i + j]) + p.b1[n];
                const float s = __sinf(p.alpha[n] * u);
                r[j] = fmaf(fmaf(p.ialpha[n], s * s, u), p.scale[n], p.shift[n]);
            }
            split2(r[0], r[1], hi[i], lo[i]);
        }
        st_shared_v4(dst, hi[0], hi[1], hi[2], hi[3]);
        st_shared_v4(dst + kPlane, hi[4], hi[5], hi[6], hi[7]);
        st_shared_v4(dst + 2 * kPlane, lo[0], lo[1], lo[2], lo[3]);
        st_shared_v4(dst + 3 * kPlane, lo[4], lo[5], lo[6], lo[7]);
    }
};

template <int C, int CH>
struct ChunkLoop {
    static __device__ __forceinline__ void run(const Params<C>& p, uint32_t tl, uint32_t a2_blk, uint32_t a2_full, uint32_t a2_empty, int it,
                                               int lane) {
        using cfg = Cfg<C>;
        constexpr int u = CH & 1;
        // chunk CH reuses the ring slot of chunk CH - 2: wait for that chunk's MMAs (kChunks / 2 phases per tile and slot)
        if (CH >= 2) mbar_wait(a2_empty + 8 * u, (it * (cfg::kChunks / 2) + ((CH - 2) >> 1)) & 1);
        Chunk<C, CH>::run(p, tl, a2_blk + u * 4 * kPlane);
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_full + 8 * u);
        if constexpr (CH + 1 < cfg::kChunks) ChunkLoop<C, CH + 1>::run(p, tl, a2_blk, a2_full, a2_empty, it, lane);
    }
};

template <int C>
__global__ void __launch_bounds__(kThreads, Cfg<C>::kCtasPerSm) convunit_umma_kernel(const __grid_constant__ Params<C> p) {
    using cfg = Cfg<C>;
    constexpr int kH = cfg::kH;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const float* xs = reinterpret_cast<const float*>(smem);                 // [kTileRows][kPitch]: rows t0 - 3 .. t0 + 258
    const uint32_t a1_s = sbase + cfg::kOffA1, zero_s = sbase + cfg::kOffZero, a2_s = sbase + cfg::kOffA2;
    const uint32_t w1_s = sbase + cfg::kOffW1, w2_s = sbase + cfg::kOffW2, bars = sbase + cfg::kOffBars;
    const uint32_t a1_ready = bars, d1_ready = a1_ready + 8 * kBlocks, a2_full = d1_ready + 8 * kBlocks,
                   a2_empty = a2_full + 16 * kBlocks, d2_ready = a2_empty + 16 * kBlocks, tmem_slot = d2_ready + 8 * kBlocks;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - sbase));

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int tiles_per_clip = (p.T + kRows - 1) / kRows;
    const int n_tiles = tiles_per_clip * p.B;

    {   // once per CTA: weights, the zero plane, barriers, TMEM
        const uint4* src = reinterpret_cast<const uint4*>(p.wblob);
        uint4* dst = reinterpret_cast<uint4*>(smem + cfg::kOffW1);
        for (int i = tid; i < (cfg::kW1Bytes + cfg::kW2Bytes) / 16; i += kThreads) dst[i] = __ldg(src + i);
        uint4* z = reinterpret_cast<uint4*>(smem + cfg::kOffZero);
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
    if (warp == kRowWarps) tmem_alloc(tmem_slot, kBlocks * cfg::kBlkCols);
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
        const uint32_t idesc1 = make_idesc_bf16(kH), idesc2 = make_idesc_bf16(cfg::kN2);
        const uint32_t lbo_plane = (uint32_t)(kPlane >> 4) << 16, lbo_w1 = ((uint32_t)(kH * 16) >> 4) << 16,
                       lbo_w2 = ((uint32_t)(cfg::kN2 * 16) >> 4) << 16;
        const uint32_t d1 = tmem_base + cfg::kBlkCols * b, d2 = d1 + cfg::kD2;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            mbar_wait(a1_ready + 8 * b, it & 1);
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int term = 0; term < 3; ++term) {          // (hi, Whi), (lo, Whi), (hi, Wlo)
                    const uint32_t a = a1_s + (b * 2 + (term == 1 ? 1 : 0)) * cfg::kPlanes1 * kPlane;
                    const uint32_t w = w1_s + (term == 2 ? cfg::kW1Bytes / 2 : 0);
#pragma unroll
                    for (int ks = 0; ks < cfg::kSteps1; ++ks) {
                        const uint32_t ak = a + 2 * ks * kPlane;
                        const bool padded = 2 * ks + 1 >= cfg::kPlanes1;        // odd plane count: the second K half is the zero plane
                        const uint32_t lbo = padded ? (((zero_s - ak) >> 4) << 16) : lbo_plane;
                        tc_mma_bf16(d1, kDescHi | ((ak >> 4) | lbo), kDescHi | (((w + ks * 2 * kH * 16) >> 4) | lbo_w1), idesc1, (term | ks) ? 1u : 0u);
                    }
                }
                tc_commit(d1_ready + 8 * b);
            }
            __syncwarp();
#pragma unroll 1
            for (int c = 0; c < cfg::kChunks; ++c) {
                const int u = c & 1;
                mbar_wait(a2_full + 16 * b + 8 * u, (it * (cfg::kChunks / 2) + (c >> 1)) & 1);
                tc_fence_after();
                if (leader) {
                    const uint32_t a = a2_s + ((b * 2 + u) * 4) * kPlane;              // hi planes 0, 1; lo planes 2, 3
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t at = a + (term == 1 ? 2 * kPlane : 0);
                        const uint32_t w = w2_s + (term == 2 ? cfg::kW2Bytes / 2 : 0) + c * (2 * cfg::kN2 * 16);
                        tc_mma_bf16(d2, kDescHi | ((at >> 4) | lbo_plane), kDescHi | ((w >> 4) | lbo_w2), idesc2, (c | term) ? 1u : 0u);
                    }
                    tc_commit(a2_empty + 16 * b + 8 * u);
                    if (c == cfg::kChunks - 1) tc_commit(d2_ready + 8 * b);
                }
                __syncwarp();
            }
        }
    } else {
        // =============================================================== row owners
        const int blk = warp >> 2, quad = warp & 3;
        const int r = blk * 128 + quad * 32 + lane;                // this thread's time step within the tile
        const uint32_t tl = tmem_base + ((uint32_t)(quad * 32) << 16) + cfg::kBlkCols * blk;
        const uint32_t row16 = (uint32_t)((quad * 32 + lane) * 16);
        auto row_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(kRowThreads) : "memory"); };
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int clip = tile / tiles_per_clip;
            const int t0 = (tile - clip * tiles_per_clip) * kRows;
            const float* xb = p.x + (long long)clip * p.T * C;
            row_sync();                                            // the previous tile's rows are no longer read
            for (int i = tid; i < cfg::kTileRows * (C / 4); i += kRowThreads) {      // 16-byte pieces; rows outside the clip: zero fill
                const int row = i / (C / 4), piece = i - row * (C / 4);
                const int t = t0 - 3 + row;
                const bool ok = t >= 0 && t < p.T;
                const float* src = xb + (long long)(ok ? t : 0) * C + piece * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + (uint32_t)(row * cfg::kPitch + piece * 4) * 4), "l"(src),
                             "r"(ok ? 16 : 0)
                             : "memory");
            }
            cp_async_wait_all();
            row_sync();

            // ---- S1: dwconv7 + LayerNorm of this row, split, operand planes of pw_conv1
            {
                float y[C];
#pragma unroll
                for (int c = 0; c < C; ++c) y[c] = p.dw_b[c];
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    const float4* src = reinterpret_cast<const float4*>(xs + (r + q) * cfg::kPitch);
#pragma unroll
                    for (int g = 0; g < C / 4; ++g) {
                        const float4 v = src[g];
                        y[4 * g] = fmaf(p.dw_w[q][4 * g], v.x, y[4 * g]);
                        y[4 * g + 1] = fmaf(p.dw_w[q][4 * g + 1], v.y, y[4 * g + 1]);
                        y[4 * g + 2] = fmaf(p.dw_w[q][4 * g + 2], v.z, y[4 * g + 2]);
                        y[4 * g + 3] = fmaf(p.dw_w[q][4 * g + 3], v.w, y[4 * g + 3]);
                    }
                }
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) sum += y[c];
                const float mean = sum * (1.0f / C);
                float sq = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    y[c] -= mean;
                    sq = fmaf(y[c], y[c], sq);
                }
                const float rstd = rsqrt_nr(sq * (1.0f / C) + p.eps);
                uint32_t hi[C / 2], lo[C / 2];
#pragma unroll
                for (int i = 0; i < C / 2; ++i)
                    split2(fmaf(y[2 * i] * rstd, p.ln_w[2 * i], p.ln_b[2 * i]), fmaf(y[2 * i + 1] * rstd, p.ln_w[2 * i + 1], p.ln_b[2 * i + 1]),
                           hi[i], lo[i]);
                const uint32_t dst = a1_s + (blk * 2) * cfg::kPlanes1 * kPlane + row16;
#pragma unroll
                for (int pl = 0; pl < cfg::kPlanes1; ++pl) {
                    st_shared_v4(dst + pl * kPlane, hi[4 * pl], hi[4 * pl + 1], hi[4 * pl + 2], hi[4 * pl + 3]);
                    st_shared_v4(dst + (cfg::kPlanes1 + pl) * kPlane, lo[4 * pl], lo[4 * pl + 1], lo[4 * pl + 2], lo[4 * pl + 3]);
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a1_ready + 8 * blk);

            // ---- S2: hidden columns 16 at a time through the two-deep chunk ring
            mbar_wait(d1_ready + 8 * blk, it & 1);
            tc_fence_after();
            ChunkLoop<C, 0>::run(p, tl, a2_s + (blk * 2 * 4) * kPlane + row16, a2_full + 16 * blk, a2_empty + 16 * blk, it, lane);

            // ---- S3: + bias + residual, store fp32 rows or the split pair
            mbar_wait(d2_ready + 8 * blk, it & 1);
            tc_fence_after();
            {
                const int t = t0 + r;
                const float4* res = reinterpret_cast<const float4*>(xs + (r + 3) * cfg::kPitch);
#pragma unroll
                for (int g8 = 0; g8 < C / 8; ++g8) {
                    uint32_t v[8];
                    tmem_ld8(tl + cfg::kD2 + 8 * g8, v);
                    tmem_ld_wait();
                    const float4 r0 = res[2 * g8], r1 = res[2 * g8 + 1];
                    float o[8];
                    o[0] = __uint_as_float(v[0]) + p.b2[8 * g8] + r0.x;
                    o[1] = __uint_as_float(v[1]) + p.b2[8 * g8 + 1] + r0.y;
                    o[2] = __uint_as_float(v[2]) + p.b2[8 * g8 + 2] + r0.z;
                    o[3] = __uint_as_float(v[3]) + p.b2[8 * g8 + 3] + r0.w;
                    o[4] = __uint_as_float(v[4]) + p.b2[8 * g8 + 4] + r1.x;
                    o[5] = __uint_as_float(v[5]) + p.b2[8 * g8 + 5] + r1.y;
                    o[6] = __uint_as_float(v[6]) + p.b2[8 * g8 + 6] + r1.z;
                    o[7] = __uint_as_float(v[7]) + p.b2[8 * g8 + 7] + r1.w;
                    if (t < p.T) {
                        const long long off = ((long long)clip * p.T + t) * C + 8 * g8;
                        if (!p.out_split) {
                            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
                            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                        } else {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) split2(o[2 * i], o[2 * i + 1], hi[i], lo[i]);
                            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out_lo) + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                }
                tc_fence_before();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kRowWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kBlocks * cfg::kBlkCols);
    }
}

}  // namespace cuu
}  // namespace l3ac

struct l3ac_convunit_plan {
    int C;
    void* params;            // Params<24> or Params<48> (host)
    void* dev_blob;
    int device;
};

namespace {

template <int C>
int make_plan(const float* dw_w, const float* dw_b, const float* ln_w, const float* ln_b, float eps, const float* w1, const float* b1,
              const float* alpha, const float* scale, const float* shift, const float* w2, const float* b2, l3ac_convunit_plan* plan) {
    using namespace l3ac::cuu;
    using cfg = Cfg<C>;
    constexpr int kH = cfg::kH;
    auto bits = [](float v) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        uint16_t b;
        memcpy(&b, &h, 2);
        return b;
    };
    auto rnd = [](float v) { return __bfloat162float(__float2bfloat16_rn(v)); };
    std::vector<uint16_t> blob((cfg::kW1Bytes + cfg::kW2Bytes) / 2, 0);
    // W1: [part][kstep][half][n 4C][k 8]; channel = 16 kstep + 8 half + k (< C); w1[n][channel]
    for (int part = 0; part < 2; ++part)
        for (int ks = 0; ks < cfg::kSteps1; ++ks)
            for (int h = 0; h < 2; ++h)
                for (int n = 0; n < kH; ++n)
                    for (int k = 0; k < 8; ++k) {
                        const int ch = 16 * ks + 8 * h + k;
                        if (ch >= C) continue;
                        const float w = w1[n * C + ch], hi = rnd(w);
                        blob[(size_t)part * (cfg::kW1Bytes / 4) + ((ks * 2 + h) * kH + n) * 8 + k] = bits(part == 0 ? hi : w - hi);
                    }
    // W2: [part][chunk][half][n kN2][k 8]; hidden = 16 chunk + 8 half + k; w2[n][hidden], n < C
    const size_t w2o = cfg::kW1Bytes / 2;
    for (int part = 0; part < 2; ++part)
        for (int c = 0; c < cfg::kChunks; ++c)
            for (int h = 0; h < 2; ++h)
                for (int n = 0; n < C; ++n)
                    for (int k = 0; k < 8; ++k) {
                        const float w = w2[n * kH + 16 * c + 8 * h + k], hi = rnd(w);
                        blob[w2o + (size_t)part * (cfg::kW2Bytes / 4) + ((c * 2 + h) * cfg::kN2 + n) * 8 + k] = bits(part == 0 ? hi : w - hi);
                    }
    cudaError_t e = cudaMalloc(&plan->dev_blob, cfg::kW1Bytes + cfg::kW2Bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(plan->dev_blob, blob.data(), cfg::kW1Bytes + cfg::kW2Bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(plan->dev_blob); return (int)e; }
    Params<C>* p = new (std::nothrow) Params<C>();
    if (!p) { cudaFree(plan->dev_blob); return L3AC_EINVAL; }
    p->wblob = static_cast<const uint8_t*>(plan->dev_blob);
    p->eps = eps;
    for (int q = 0; q < 7; ++q)
        for (int c = 0; c < C; ++c) p->dw_w[q][c] = dw_w[q * C + c];
    for (int c = 0; c < C; ++c) {
        p->dw_b[c] = dw_b[c];
        p->ln_w[c] = ln_w[c];
        p->ln_b[c] = ln_b[c];
        p->b2[c] = b2[c];
    }
    for (int n = 0; n < kH; ++n) {
        p->b1[n] = b1[n];
        p->alpha[n] = alpha[n];
        p->ialpha[n] = 1.0f / (alpha[n] + l3ac::kEps);
        p->scale[n] = scale[n];
        p->shift[n] = shift[n];
    }
    plan->params = p;
    return L3AC_OK;
}

template <int C>
int run_plan(const l3ac_convunit_plan* plan, const float* x, int B, int T, void* out, void* out_lo, int out_dtype, cudaStream_t stream) {
    using namespace l3ac::cuu;
    using cfg = Cfg<C>;
    Params<C> p = *static_cast<const Params<C>*>(plan->params);
    p.x = x;
    p.out = out;
    p.out_lo = out_lo;
    p.B = B;
    p.T = T;
    p.out_split = out_dtype == L3AC_BF16X2 ? 1 : 0;
    cudaError_t e = cudaFuncSetAttribute(convunit_umma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    const long long n_tiles = (long long)l3ac_cdiv(T, kRows) * B;
    if (n_tiles >= (1LL << 30)) return L3AC_EINVAL;
    const long long ctas = (long long)cfg::kCtasPerSm * l3ac_sm_count();
    l3ac_launch(convunit_umma_kernel<C>, dim3((int)(n_tiles < ctas ? n_tiles : ctas)), dim3(kThreads), cfg::kSmemBytes, stream, p);
    return l3ac_launch_status();
}

}  // namespace

extern "C" int l3ac_convunit_plan_create(int C, const float* dw_w, const float* dw_b, const float* ln_w, const float* ln_b, float eps,
                                         const float* w1, const float* b1, const float* alpha, const float* scale, const float* shift,
                                         const float* w2, const float* b2, l3ac_convunit_plan** plan_out) {
    L3AC_CHECK_ARG(dw_w && dw_b && ln_w && ln_b && w1 && b1 && alpha && scale && shift && w2 && b2 && plan_out);
    if (C != 24 && C != 48) return L3AC_EUNSUPPORTED;
    l3ac_convunit_plan* plan = new (std::nothrow) l3ac_convunit_plan();
    if (!plan) return L3AC_EINVAL;
    plan->C = C;
    if (cudaGetDevice(&plan->device) != cudaSuccess) { delete plan; return L3AC_EDRIVER; }
    const int rc = C == 24 ? make_plan<24>(dw_w, dw_b, ln_w, ln_b, eps, w1, b1, alpha, scale, shift, w2, b2, plan)
                           : make_plan<48>(dw_w, dw_b, ln_w, ln_b, eps, w1, b1, alpha, scale, shift, w2, b2, plan);
    if (rc != L3AC_OK) { delete plan; return rc; }
    *plan_out = plan;
    return L3AC_OK;
}

extern "C" int l3ac_convunit_plan_destroy(l3ac_convunit_plan* plan) {
    if (!plan) return L3AC_OK;
    cudaFree(plan->dev_blob);
    if (plan->C == 24) delete static_cast<l3ac::cuu::Params<24>*>(plan->params);
    else delete static_cast<l3ac::cuu::Params<48>*>(plan->params);
    delete plan;
    return L3AC_OK;
}

extern "C" int l3ac_convunit_umma(const l3ac_convunit_plan* plan, const float* x, int B, int T, void* out, void* out_lo, int out_dtype,
                                  l3ac_stream_t stream) {
    L3AC_CHECK_ARG(plan && x && out && B > 0 && T > 0);
    L3AC_CHECK_ARG(out_dtype == L3AC_F32 || out_dtype == L3AC_BF16X2);
    L3AC_CHECK_ARG((out_dtype == L3AC_BF16X2) == (out_lo != nullptr));
    L3AC_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out_lo)) & 15) == 0);
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return L3AC_EDRIVER;
    L3AC_CHECK_ARG(dev == plan->device);
    return plan->C == 24 ? run_plan<24>(plan, x, B, T, out, out_lo, out_dtype, (cudaStream_t)stream)
                         : run_plan<48>(plan, x, B, T, out, out_lo, out_dtype, (cudaStream_t)stream);
}
