// Fused thin-channel Residual(ConvUnit) on the tensor cores at fp32-class precision (l3ac/modules.py:10-44 for the
// encode-side stages with C = 24 (full sample rate) and C = 48):
//   out = x + pw_conv2( GRN( snake( pw_conv1( LayerNorm( dwconv7(x) ) ) ) ) )
// As tcgen05 launches these layers are bound by the TMA row rate (48/96-byte rows) and by a 4C-wide hidden tensor that has
// to round-trip through HBM as a split-bf16 pair; as an fp32 SIMT kernel (convunit_thin.cu) the two point-wise GEMMs are
// 4.6 kMAC of FFMA per time step and shared-memory-broadcast bound.  Here every warp owns 16-row tiles and keeps the
// whole unit in REGISTERS in the mma.m16n8 accumulator-fragment layout:
//   * dwconv7 + LayerNorm are computed straight into that layout from an fp32 x tile in shared memory (the LayerNorm
//     reduction over channels is two quad shuffles),
//   * the result is split into a bf16 (hi, lo) pair = the A fragments of pw_conv1, which runs as 3-term split MMAs
//     (hi*Whi + lo*Whi + hi*Wlo, fp32 accumulate: the same fp32-class scheme as the tcgen05 split GEMMs),
//   * the hidden activation is produced 32 columns at a time; bias + snake + folded GRN run on the accumulators, which are
//     then split again and are -- without any data movement -- the A fragments of pw_conv2, accumulated over the chunks,
//   * bias + residual are added from the same x tile and the result is stored as fp32 or directly as the split pair the
//     next tcgen05 GEMM consumes.
// HBM sees x in and the result out (8 C bytes per time step).  Weights are split and laid out in B-fragment order in
// shared memory once per (persistent) CTA; every warp double-buffers its own x rows with cp.async and, after the
// one-time set-up, synchronises with nobody but itself.
#include "common.cuh"

namespace l3ac {
namespace thintc {

// Launch shapes: (C, m-tiles per warp, warps per CTA, CTAs per SM).  Measured for C = 24 on 24 clips: two m-tiles per warp with
// 1 x 16 warps per SM (weight fragments loaded once per two tiles) 479 us; one m-tile per warp at <= 85 registers with 3 x 8 or
// 2 x 12 warps per SM 545 / 541 us -- the extra resident warps do not pay for the doubled fragment traffic.  C = 48 needs 74 KB
// of weight fragments and runs one m-tile per warp.
template <int C, int kMT, int kWarps, int kCtas>
struct Geo {
    static constexpr int kThreads = kWarps * 32;
    static constexpr int H = 4 * C;
    static constexpr int kNT = C / 8;                    // 8-channel n-tiles of a C-wide row
    static constexpr int kK16 = C / 16;                  // full k16 steps over C
    static constexpr bool kK8 = (C % 16) != 0;           // trailing k8 step (C = 24)
    static constexpr int kKS1 = kK16 + (kK8 ? 1 : 0);
    static constexpr int kChunks = H / 32;               // hidden columns are produced 32 at a time
    static constexpr int kPitch = C;                     // floats per staged row (C = 24: the fragment-pattern float2 reads of 4 rows x
                                                         // 4 lanes hit 32 distinct banks; C = 48: 2-way conflicts on the 84 x-tile reads
                                                         // of a 432-MMA tile, accepted to keep two buffers per warp in shared memory)
    static constexpr int kRows = kMT * 16;               // rows per warp tile
    static constexpr int kXsFloats = (kRows + 6) * kPitch;   // one staging buffer of one warp
    static constexpr int kW1Vec = kKS1 * (H / 8) * 32;   // uint4 {hi.b0, hi.b1, lo.b0, lo.b1} per (k-step, n-tile, lane)
    static constexpr int kW2Vec = (H / 16) * kNT * 32;
    static constexpr int kParFloats = (H / 2) * 12;      // per hidden column pair: b1, alpha, 1/alpha, scale, shift (x2), 2 pad
    static constexpr int kCParFloats = 7 * C + 4 * C;    // dw [7][C], dw_b, ln_w, ln_b, b2
    static constexpr size_t kSmemBytes =
        2 * (size_t)kWarps * kXsFloats * 4 + ((size_t)kW1Vec + kW2Vec) * 16 + (kParFloats + kCParFloats) * 4;
    static_assert(C % 8 == 0 && (kPitch * 4) % 16 == 0 && (kXsFloats * 4) % 16 == 0, "rows must be 16-byte chunks");
};

struct Params {
    const float* x;
    const float *dw_w, *dw_b, *ln_w, *ln_b, *w1, *b1, *alpha, *scale, *shift, *w2, *b2;
    float eps;
    int B, T;
    void* out;
    void* out_lo;
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

// v -> (bf16x2(v), bf16x2(v - hi)): the split pair of a channel pair
__device__ __forceinline__ void split2(float2 v, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v.x - hf.x, v.y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ uint32_t pack2(float2 v) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;                       // src-size 0: the 16 bytes are zero-filled (conv zero padding)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n));
}

// SPLIT: write the result as the split-bf16 pair the next tcgen05 GEMM consumes instead of fp32.
// kSplitOps: 3-term split-bf16 products (fp32-class, encode side); false = plain bf16 operands with fp32 accumulation (the
// decode side's arithmetic: A = bf16(LayerNorm(..)), hidden = bf16(snake(..)), bf16 weights), one MMA per product.
template <int C, int kMT, int kWarps, int kCtas, bool SPLIT, bool kSplitOps>
__global__ void __launch_bounds__(kWarps * 32, kCtas) convunit_tc_split_kernel(const Params p) {
    using G = Geo<C, kMT, kWarps, kCtas>;
    constexpr int kThreads = G::kThreads;
    constexpr int H = G::H, kNT = G::kNT, kK16 = G::kK16, kKS1 = G::kKS1, kPitch = G::kPitch, kRows = G::kRows;
    extern __shared__ __align__(16) uint8_t smem[];
    float* xs = reinterpret_cast<float*>(smem);                               // [kWarps][2][kRows + 6][kPitch]
    uint4* w1f = reinterpret_cast<uint4*>(xs + 2 * kWarps * G::kXsFloats);    // [kKS1][H/8][32]
    uint4* w2f = w1f + G::kW1Vec;                                             // [H/16][kNT][32]
    float* par = reinterpret_cast<float*>(w2f + G::kW2Vec);                   // [H/2][12]
    float* cpar = par + G::kParFloats;                                        // dw [7][C], dw_b, ln_w, ln_b, b2

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int tiles_per_clip = (p.T + kRows - 1) / kRows;
    const long long n_tiles = (long long)tiles_per_clip * p.B;
    const long long tile_step = (long long)gridDim.x * kWarps;
    float* xw = xs + warp * 2 * G::kXsFloats;                                 // this warp's two staging buffers
    const uint32_t xw_addr = (uint32_t)__cvta_generic_to_shared(xw);

    // Every warp stages its own rows t0 - 3 ... t0 + kRows + 2 (zero outside the clip = the conv's zero padding) and only
    // synchronises with itself: after the one-time weight set-up there is no CTA-wide barrier, so the warps of an SM drift
    // apart and the tensor-core phases of some overlap the SIMT phases (LayerNorm, snake, splits) of others.
    auto stage = [&](long long tile, int buf) {
        const int clip = (int)(tile / tiles_per_clip);
        const int t0 = (int)(tile - (long long)clip * tiles_per_clip) * kRows;
        const float* xb = p.x + (long long)clip * p.T * C;
        constexpr int kVecRow = C / 4;
        for (int i = lane; i < (kRows + 6) * kVecRow; i += 32) {
            const int r = i / kVecRow, c4 = i - r * kVecRow;
            const int t = t0 + r - 3;
            const bool ok = t >= 0 && t < p.T;
            cp_async16(xw_addr + (uint32_t)((buf * G::kXsFloats + r * kPitch + 4 * c4) * 4), xb + (long long)(ok ? t : 0) * C + 4 * c4, ok);
        }
        asm volatile("cp.async.commit_group;");
    };

    long long tile = (long long)blockIdx.x * kWarps + warp;
    if (tile < n_tiles) stage(tile, 0);

    // ---- once per CTA: weights as split B fragments, parameters
    for (int i = tid; i < G::kW1Vec; i += kThreads) {          // pw_conv1: B[k = channel][n = hidden column] = w1[n][k]
        const int l = i & 31, nt = (i >> 5) % (H / 8), ks = (i >> 5) / (H / 8);
        const int n = nt * 8 + (l >> 2), k0 = ks * 16 + (l & 3) * 2;
        const int ko[4] = {k0, k0 + 1, k0 + 8, k0 + 9};
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float w = ko[j] < C ? __ldg(p.w1 + n * C + ko[j]) : 0.f;
            hi[j] = __bfloat162float(__float2bfloat16_rn(w));
            lo[j] = w - hi[j];
        }
        uint4 v;
        uint32_t d;
        split2(make_float2(hi[0], hi[1]), v.x, d);
        split2(make_float2(hi[2], hi[3]), v.y, d);
        split2(make_float2(lo[0], lo[1]), v.z, d);
        split2(make_float2(lo[2], lo[3]), v.w, d);
        w1f[i] = v;
    }
    for (int i = tid; i < G::kW2Vec; i += kThreads) {          // pw_conv2: B[k = hidden column][n = channel] = w2[n][k]
        const int l = i & 31, n2 = (i >> 5) % kNT, ks = (i >> 5) / kNT;
        const int n = n2 * 8 + (l >> 2), k0 = ks * 16 + (l & 3) * 2;
        const int ko[4] = {k0, k0 + 1, k0 + 8, k0 + 9};
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float w = __ldg(p.w2 + n * H + ko[j]);
            hi[j] = __bfloat162float(__float2bfloat16_rn(w));
            lo[j] = w - hi[j];
        }
        uint4 v;
        uint32_t d;
        split2(make_float2(hi[0], hi[1]), v.x, d);
        split2(make_float2(hi[2], hi[3]), v.y, d);
        split2(make_float2(lo[0], lo[1]), v.z, d);
        split2(make_float2(lo[2], lo[3]), v.w, d);
        w2f[i] = v;
    }
    for (int i = tid; i < H; i += kThreads) {
        float* q = par + (i >> 1) * 12 + (i & 1);
        const float a = __ldg(p.alpha + i);
        q[0] = __ldg(p.b1 + i);
        q[2] = a;
        q[4] = 1.0f / (a + kEps);
        q[6] = __ldg(p.scale + i);
        q[8] = __ldg(p.shift + i);
    }
    for (int i = tid; i < 7 * C; i += kThreads) cpar[i] = __ldg(p.dw_w + i);
    for (int i = tid; i < C; i += kThreads) {
        cpar[7 * C + i] = __ldg(p.dw_b + i);
        cpar[8 * C + i] = __ldg(p.ln_w + i);
        cpar[9 * C + i] = __ldg(p.ln_b + i);
        cpar[10 * C + i] = __ldg(p.b2 + i);
    }

    __syncthreads();                           // weights and parameters staged (the only CTA-wide barrier)

    for (int it = 0; tile < n_tiles; ++it, tile += tile_step) {
        asm volatile("cp.async.wait_all;");
        __syncwarp();                          // this tile's rows are visible to the whole warp; the other buffer is free
        if (tile + tile_step < n_tiles) stage(tile + tile_step, (it + 1) & 1);
        const int clip = (int)(tile / tiles_per_clip);
        const int t0 = (int)(tile - (long long)clip * tiles_per_clip) * kRows;
        const float* xt = xw + (it & 1) * G::kXsFloats + g * kPitch + t4 * 2;   // + (16 i + 8 h + tap) rows, + 8 n

        // ---- dwconv7 + bias and LayerNorm over the C channels, in the fragment layout; split -> A fragments of pw_conv1
        uint32_t ahi[kMT][kNT][2], alo[kMT][kNT][2];
        {
            float2 y[kMT][2][kNT];
#pragma unroll
            for (int n = 0; n < kNT; ++n) {
                const float2 db = *reinterpret_cast<const float2*>(cpar + 7 * C + n * 8 + t4 * 2);
#pragma unroll
                for (int i = 0; i < kMT; ++i) y[i][0][n] = y[i][1][n] = db;
            }
#pragma unroll
            for (int j = 0; j < 7; ++j)
#pragma unroll
                for (int n = 0; n < kNT; ++n) {
                    const float2 w = *reinterpret_cast<const float2*>(cpar + j * C + n * 8 + t4 * 2);
#pragma unroll
                    for (int i = 0; i < kMT; ++i)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            y[i][h][n] = ffma2(w, *reinterpret_cast<const float2*>(xt + (i * 16 + h * 8 + j) * kPitch + n * 8), y[i][h][n]);
                }
#pragma unroll
            for (int i = 0; i < kMT; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float sum = 0.f;
#pragma unroll
                    for (int n = 0; n < kNT; ++n) sum += y[i][h][n].x + y[i][h][n].y;
                    const float mean = quad_sum(sum) * (1.0f / C);
                    float v = 0.f;
#pragma unroll
                    for (int n = 0; n < kNT; ++n) {
                        y[i][h][n].x -= mean;
                        y[i][h][n].y -= mean;
                        v = fmaf(y[i][h][n].x, y[i][h][n].x, v);
                        v = fmaf(y[i][h][n].y, y[i][h][n].y, v);
                    }
                    const float rstd = rsqrt_nr(quad_sum(v) * (1.0f / C) + p.eps);
#pragma unroll
                    for (int n = 0; n < kNT; ++n) {
                        const float2 lw = *reinterpret_cast<const float2*>(cpar + 8 * C + n * 8 + t4 * 2);
                        const float2 lb = *reinterpret_cast<const float2*>(cpar + 9 * C + n * 8 + t4 * 2);
                        const float2 a = ffma2(fmul2(y[i][h][n], make_float2(rstd, rstd)), lw, lb);
                        if (kSplitOps) {
                            split2(a, ahi[i][n][h], alo[i][n][h]);
                        } else {
                            ahi[i][n][h] = pack2(a);
                            alo[i][n][h] = 0;
                        }
                    }
                }
        }

        // ---- the MLP, 32 hidden columns at a time
        float acc[kMT][kNT][4];
#pragma unroll
        for (int i = 0; i < kMT; ++i)
#pragma unroll
            for (int n = 0; n < kNT; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[i][n][e] = 0.f;

#pragma unroll 1
        for (int hc = 0; hc < G::kChunks; ++hc) {
            float hacc[kMT][4][4];
            const float* pq = par + (hc * 16 + t4) * 12;        // column pair (32 hc + 8 nt + 2 t4) / 2 = 16 hc + 4 nt + t4
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float2 b1 = *reinterpret_cast<const float2*>(pq + nt * 48);
#pragma unroll
                for (int i = 0; i < kMT; ++i) {
                    hacc[i][nt][0] = hacc[i][nt][2] = b1.x;
                    hacc[i][nt][1] = hacc[i][nt][3] = b1.y;
                }
            }
            // pw_conv1: K = C
#pragma unroll
            for (int ks = 0; ks < kKS1; ++ks)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const uint4 w = w1f[(ks * (H / 8) + hc * 4 + nt) * 32 + lane];
#pragma unroll
                    for (int i = 0; i < kMT; ++i) {
                        if (ks < kK16) {
                            mma_k16(hacc[i][nt], ahi[i][2 * ks][0], ahi[i][2 * ks][1], ahi[i][2 * ks + 1][0], ahi[i][2 * ks + 1][1], w.x, w.y);
                            if (kSplitOps) {
                                mma_k16(hacc[i][nt], alo[i][2 * ks][0], alo[i][2 * ks][1], alo[i][2 * ks + 1][0], alo[i][2 * ks + 1][1], w.x, w.y);
                                mma_k16(hacc[i][nt], ahi[i][2 * ks][0], ahi[i][2 * ks][1], ahi[i][2 * ks + 1][0], ahi[i][2 * ks + 1][1], w.z, w.w);
                            }
                        } else {                         // trailing 8 channels
                            mma_k8(hacc[i][nt], ahi[i][kNT - 1][0], ahi[i][kNT - 1][1], w.x);
                            if (kSplitOps) {
                                mma_k8(hacc[i][nt], alo[i][kNT - 1][0], alo[i][kNT - 1][1], w.x);
                                mma_k8(hacc[i][nt], ahi[i][kNT - 1][0], ahi[i][kNT - 1][1], w.z);
                            }
                        }
                    }
                }
            // snake + folded GRN on the accumulators, split -> A fragments of pw_conv2 (k-step s = hidden n-tiles 2 s, 2 s + 1)
            uint32_t hhi[kMT][4][2], hlo[kMT][4][2];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float4 q0 = *reinterpret_cast<const float4*>(pq + nt * 48);          // b1.x b1.y al.x al.y
                const float4 q1 = *reinterpret_cast<const float4*>(pq + nt * 48 + 4);      // ia.x ia.y sc.x sc.y
                const float2 sh = *reinterpret_cast<const float2*>(pq + nt * 48 + 8);
#pragma unroll
                for (int i = 0; i < kMT; ++i)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float2 v = snake_affine2(make_float2(hacc[i][nt][2 * h], hacc[i][nt][2 * h + 1]), make_float2(q0.z, q0.w),
                                                       make_float2(q1.x, q1.y), make_float2(q1.z, q1.w), sh);
                        if (kSplitOps) {
                            split2(v, hhi[i][nt][h], hlo[i][nt][h]);
                        } else {
                            hhi[i][nt][h] = pack2(v);
                            hlo[i][nt][h] = 0;
                        }
                    }
            }
            // pw_conv2: this chunk's 32 hidden columns are two k16 steps
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int n = 0; n < kNT; ++n) {
                    const uint4 w = w2f[((hc * 2 + ks) * kNT + n) * 32 + lane];
#pragma unroll
                    for (int i = 0; i < kMT; ++i) {
                        mma_k16(acc[i][n], hhi[i][2 * ks][0], hhi[i][2 * ks][1], hhi[i][2 * ks + 1][0], hhi[i][2 * ks + 1][1], w.x, w.y);
                        if (kSplitOps) {
                            mma_k16(acc[i][n], hlo[i][2 * ks][0], hlo[i][2 * ks][1], hlo[i][2 * ks + 1][0], hlo[i][2 * ks + 1][1], w.x, w.y);
                            mma_k16(acc[i][n], hhi[i][2 * ks][0], hhi[i][2 * ks][1], hhi[i][2 * ks + 1][0], hhi[i][2 * ks + 1][1], w.z, w.w);
                        }
                    }
                }
        }

        // ---- + b2 + x, store
#pragma unroll
        for (int i = 0; i < kMT; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int t = t0 + i * 16 + h * 8 + g;
                if (t >= p.T) continue;
                const float* xr = xt + (i * 16 + h * 8 + 3) * kPitch;
                const long long base = ((long long)clip * p.T + t) * C + t4 * 2;
#pragma unroll
                for (int n = 0; n < kNT; ++n) {
                    const float2 b2 = *reinterpret_cast<const float2*>(cpar + 10 * C + n * 8 + t4 * 2);
                    const float2 xv = *reinterpret_cast<const float2*>(xr + n * 8);
                    const float2 r = fadd2(fadd2(make_float2(acc[i][n][2 * h], acc[i][n][2 * h + 1]), b2), xv);
                    if (!SPLIT) {
                        *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + base + n * 8) = r;
                    } else {
                        uint32_t hi, lo;
                        split2(r, hi, lo);
                        *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.out) + base + n * 8) = hi;
                        *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.out_lo) + base + n * 8) = lo;
                    }
                }
            }
    }
    asm volatile("cp.async.wait_all;");
}

template <int C, int kMT, int kWarps, int kCtas, bool kSplitOps>
static int launch(const Params& p, bool split, cudaStream_t stream) {
    using G = Geo<C, kMT, kWarps, kCtas>;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return L3AC_EDRIVER;
    const long long n_tiles = (long long)l3ac_cdiv(p.T, G::kRows) * p.B;           // warp tiles
    const long long ctas = (n_tiles + kWarps - 1) / kWarps;
    const int grid = (int)(ctas < (long long)sms * kCtas ? ctas : (long long)sms * kCtas);
    auto go = [&](auto kernel) -> int {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        kernel<<<grid, G::kThreads, G::kSmemBytes, stream>>>(p);
        return l3ac_launch_status();
    };
    return split ? go(convunit_tc_split_kernel<C, kMT, kWarps, kCtas, true, kSplitOps>)
                 : go(convunit_tc_split_kernel<C, kMT, kWarps, kCtas, false, kSplitOps>);
}

}  // namespace thintc
}  // namespace l3ac

extern "C" int l3ac_convunit_thin_tc(const float* x, int B, int T, int C, const float* dw_w, const float* dw_b,
                                     const float* ln_w, const float* ln_b, float eps, const float* w1, const float* b1,
                                     const float* alpha, const float* scale, const float* shift, const float* w2,
                                     const float* b2, void* out, void* out_lo, int out_dtype, int operand_dtype,
                                     l3ac_stream_t stream) {
    using namespace l3ac::thintc;
    L3AC_CHECK_ARG(x && dw_w && dw_b && ln_w && ln_b && w1 && b1 && alpha && scale && shift && w2 && b2 && out);
    L3AC_CHECK_ARG(B > 0 && T > 0);
    L3AC_CHECK_ARG(out_dtype == L3AC_F32 || (out_dtype == L3AC_BF16X2 && out_lo));
    L3AC_CHECK_ARG(operand_dtype == L3AC_BF16X2 || operand_dtype == L3AC_BF16);
    L3AC_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out_lo)) & 15) == 0);
    Params p{x, dw_w, dw_b, ln_w, ln_b, w1, b1, alpha, scale, shift, w2, b2, eps, B, T, out, out_lo};
    const bool split = out_dtype == L3AC_BF16X2;
    if (operand_dtype == L3AC_BF16X2) {
        if (C == 24) return launch<24, 2, 16, 1, true>(p, split, (cudaStream_t)stream);
        if (C == 48) return launch<48, 1, 16, 1, true>(p, split, (cudaStream_t)stream);
    } else {
        if (C == 24) return launch<24, 2, 16, 1, false>(p, split, (cudaStream_t)stream);
        if (C == 48) return launch<48, 1, 16, 1, false>(p, split, (cudaStream_t)stream);
    }
    return L3AC_EUNSUPPORTED;
}
