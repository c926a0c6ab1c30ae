// Fused full-rate decoder tail (l3ac/modules.py:174-194): three Residual(LegacyUnit) blocks with dilations d0,d1,d2
//   x += Conv1x1( snake( Conv_k7,dil d( snake(x, a0) ) + b, a1 ) ) + b'          (l3ac/modules.py:47-64)
// followed by Snake -> Conv1d(C -> 1, k7, pad 3) -> tanh, in ONE kernel: the (B, T, 24) fp32 stream is read once from
// HBM and only the (B, T) waveform is written.  As separate launches these nine full-rate tensors are pure HBM
// traffic ("thin full-rate layers are HBM-bound").
//
// One CTA owns kRows consecutive samples of one clip (kRows - 84 outputs + a 42-sample halo on each side = the
// receptive field 3*(1+3+9)+3); each warp owns kMT tiles of 16 consecutive samples.  The fp32 residual stream never
// leaves REGISTERS: it is held in the mma.m16n8 accumulator-fragment layout (row = lane/4 (+8), channel pair =
// 2*(lane%4) of each 8-channel n-tile), which is at once
//   * the layout the k7 conv's accumulators come out in, so bias + snake run on registers,
//   * the A-fragment layout of the following 1x1 conv (accumulator n-tiles 0,1 = k-step 0, n-tile 2 = a k8 step), so
//     the hidden activation goes from one MMA to the next without touching shared memory,
//   * the layout of the 1x1 conv's result, so the residual add is register-to-register.
// Only bf16(snake(x)) -- the operand every OTHER warp needs through the conv's time taps -- goes through shared memory,
// double-buffered so that one __syncthreads per LegacyUnit suffices.  Rows are stored unpadded (24 bf16 = 48 B): a
// 12-word row stride puts 8 consecutive rows on 8 distinct 4-bank groups, so ldmatrix and the quad-wise 4-byte stores
// are conflict-free without a swizzle.  The conv's K dimension is tap-major (k = tap*24 + channel, 168 -> 176 = 11
// k-steps instead of 7 x 2 channel-padded ones): the two 8-channel halves of an ldmatrix k-step may come from
// different taps because ldmatrix takes one row address per lane.  27 guard rows on either side of the buffers replace
// per-tap row clamping.  The final Conv1d(24 -> 1, k7) runs on the tensor cores as well: its 7 taps sit in the N
// dimension (P[t'][j] = s[t'] . w[j], A = snake(x) straight from registers as a 3-term split-bf16 pair, fp32-class) and
// a diagonal sum y[t] = sum_j P[t+j-3][j] through shared memory finishes it.
// The kernel is persistent (weights staged once per CTA, next tile's rows prefetched into the dead residual registers).
// (24 channels cannot feed a 128xN tcgen05 tile and the kernel is bound by its SIMT snake work and latency, so it uses
// the register-level MMA path; the fat layers use tcgen05 in gemm_tc.cu / mlp_fused.cu.)
#include "common.cuh"

namespace l3ac {
namespace tail {

constexpr int kC = 24;
constexpr int kHalo = 42;
constexpr int kGuard = 27;                           // largest conv reach (3 taps x dilation 9)
constexpr int kRowBytes = kC * 2;                    // 48
constexpr int kKSteps = 11;                          // ceil(7 * 24 / 16)
constexpr int kConvFragWords = kKSteps * 3 * 32 * 2; // uint32 per unit: [kstep][ntile][lane][2]
constexpr int kPwFragWords = 2 * 3 * 32 * 2;         // [kstep][ntile][lane][2] (K padded 24 -> 32)
constexpr int kParFloats = 3 * 6 * kC + 2 * kC;      // per unit: conv_bias, pw_bias, alpha0, 1/alpha0, alpha1, 1/alpha1; final alpha, 1/alpha

struct Params {
    const float* x;
    const uint32_t* conv_frags;   // [3][kConvFragWords]
    const uint32_t* pw_frags;     // [3][kPwFragWords]
    const float* conv_bias;       // [3][24]
    const float* pw_bias;         // [3][24]
    const float* alpha0;          // [3][24]
    const float* alpha1;          // [3][24]
    const float* alpha_f;         // [24]
    const float* w_f;             // [7][24]
    float bias_f;
    int dil[3];
    int B, T;
    float* out;
};

template <int kWarps, int kMT>
struct Cfg {
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kRows = kWarps * kMT * 16;
    static constexpr int kOut = kRows - 2 * kHalo;
    static constexpr int kBufBytes = (kRows + 2 * kGuard) * kRowBytes;
    static constexpr int kCtasPerSm = (kWarps * kMT >= 32) ? 1 : 2;
    static constexpr int kPtStride = (kRows + 15) / 16 * 16 + 4;      // P^T row stride in floats, = 4 mod 16: conflict-free stores
    static constexpr size_t kSmemBytes =
        2 * (size_t)kBufBytes + 3 * (kConvFragWords + kPwFragWords) * 4 + 2 * 32 * 16 + kParFloats * 4;
    static_assert(kBufBytes % 16 == 0, "buffers must keep 16-byte alignment");
    static_assert(8 * kPtStride * 4 <= kBufBytes, "P^T must fit an operand buffer");
};

template <int kOff>
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4 + %5];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr), "n"(kOff));
}

__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_bf16_k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(b0));
}

__device__ __forceinline__ uint32_t pack_bf16(float2 v) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// x + sin^2(alpha x) / (alpha + eps) on a channel pair (MUFU sine, packed fp32 ops)
__device__ __forceinline__ float2 snake2(float2 v, float2 alpha, float2 inv_alpha) {
    const float2 t = fmul2(alpha, v);
    const float2 s = make_float2(__sinf(t.x), __sinf(t.y));
    return ffma2(inv_alpha, fmul2(s, s), v);
}

// The tile's rows of the fp32 stream, straight from global memory into the accumulator-fragment layout: thread (g, t4)
// holds rows row_w + 16 i + g + 8 h, channels 8 n + 2 t4 + {0, 1}.  Rows outside the clip are zero (every conv on the
// path zero-pads its input); bit 2 i + h of the returned mask says the row lies inside the clip.
template <int kMT>
__device__ __forceinline__ uint32_t load_rows(float (&xr)[kMT][3][4], const float* __restrict__ xb, int t_row0, int T, int t4) {
    uint32_t valid = 0;
#pragma unroll
    for (int i = 0; i < kMT; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = t_row0 + i * 16 + h * 8;
            const bool ok = t >= 0 && t < T;
            valid |= ok ? (1u << (2 * i + h)) : 0u;
            const float* src = xb + (long long)(ok ? t : 0) * kC + t4 * 2;
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                const float2 v = ok ? __ldg(reinterpret_cast<const float2*>(src + n * 8)) : make_float2(0.f, 0.f);
                xr[i][n][2 * h] = v.x;
                xr[i][n][2 * h + 1] = v.y;
            }
        }
    return valid;
}

// Persistent: CTA c processes tiles c, c + gridDim.x, ... of the (clip, time-tile) grid; the weights of all three units
// are staged in shared memory once per CTA and the next tile's rows are fetched from global memory while the final
// conv of the current tile runs (the residual registers are dead by then).
template <int kWarps, int kMT>
__global__ void __launch_bounds__(kWarps * 32, Cfg<kWarps, kMT>::kCtasPerSm) decoder_tail_kernel(const Params p) {
    using cfg = Cfg<kWarps, kMT>;
    constexpr int kThreads = cfg::kThreads, kRows = cfg::kRows, kOut = cfg::kOut, kBufBytes = cfg::kBufBytes;
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* buf0 = smem;                                                     // 2 x [kGuard + kRows + kGuard][24] bf16
    uint32_t* w_conv = reinterpret_cast<uint32_t*>(smem + 2 * kBufBytes);     // [3][kConvFragWords]
    uint32_t* w_pw = w_conv + 3 * kConvFragWords;                             // [3][kPwFragWords]
    uint4* w_fin = reinterpret_cast<uint4*>(w_pw + 3 * kPwFragWords);         // [2][32] {hi.b0, hi.b1, lo.b0, lo.b1}
    float* s_par = reinterpret_cast<float*>(w_fin + 2 * 32);            // [3][6][24], then final alpha / 1/alpha

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int row_w = warp * kMT * 16;                   // first tile row of this warp
    const int tiles_per_clip = (p.T + kOut - 1) / kOut;
    const int n_tiles = tiles_per_clip * p.B;

    int tile = blockIdx.x;
    int clip = tile / tiles_per_clip;
    int t_first = (tile - clip * tiles_per_clip) * kOut - kHalo;       // global sample of tile row 0
    float xr[kMT][3][4];                                 // the residual stream of this thread
    uint32_t valid = 0;
    if (tile < n_tiles) valid = load_rows<kMT>(xr, p.x + (long long)clip * p.T * kC, t_first + row_w + g, p.T, t4);

    // ---- once per CTA: weights of all three units, parameters, final-conv fragments, guard rows
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.conv_frags);
        uint4* dst = reinterpret_cast<uint4*>(w_conv);
        for (int i = tid; i < 3 * kConvFragWords / 4; i += kThreads) dst[i] = __ldg(src + i);
        src = reinterpret_cast<const uint4*>(p.pw_frags);
        dst = reinterpret_cast<uint4*>(w_pw);
        for (int i = tid; i < 3 * kPwFragWords / 4; i += kThreads) dst[i] = __ldg(src + i);
    }
    for (int i = tid; i < 3 * kC; i += kThreads) {
        const int u = i / kC, c = i - u * kC;
        float* sp = s_par + u * 6 * kC;
        sp[c] = __ldg(p.conv_bias + i);
        sp[kC + c] = __ldg(p.pw_bias + i);
        const float a0 = __ldg(p.alpha0 + i), a1 = __ldg(p.alpha1 + i);
        sp[2 * kC + c] = a0;
        sp[3 * kC + c] = 1.0f / (a0 + kEps);
        sp[4 * kC + c] = a1;
        sp[5 * kC + c] = 1.0f / (a1 + kEps);
    }
    if (tid < kC) {
        const float a = __ldg(p.alpha_f + tid);
        s_par[18 * kC + tid] = a;
        s_par[19 * kC + tid] = 1.0f / (a + kEps);
    }
    for (int i = tid; i < 2 * 32; i += kThreads) {   // B fragments of the final conv: k = channel (24 -> 32), n = tap (7 -> 8)
        const int ks = i >> 5, l = i & 31;
        const int n = l >> 2, k0 = ks * 16 + (l & 3) * 2;
        float hi[4], lo[4];
        const int ko[4] = {k0, k0 + 1, k0 + 8, k0 + 9};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float w = (n < 7 && ko[j] < kC) ? __ldg(p.w_f + n * kC + ko[j]) : 0.f;
            hi[j] = __bfloat162float(__float2bfloat16_rn(w));
            lo[j] = w - hi[j];
        }
        w_fin[i] = make_uint4(pack_bf16(make_float2(hi[0], hi[1])), pack_bf16(make_float2(hi[2], hi[3])),
                              pack_bf16(make_float2(lo[0], lo[1])), pack_bf16(make_float2(lo[2], lo[3])));
    }
    {
        constexpr int kGuardVec = kGuard * kRowBytes / 16;      // 81 uint4 per guard region, never written again
        for (int i = tid; i < 4 * kGuardVec; i += kThreads) {
            const int r = i / kGuardVec, o = i - r * kGuardVec;
            uint8_t* base = buf0 + (r >> 1) * kBufBytes + ((r & 1) ? (kGuard + kRows) * kRowBytes : 0);
            reinterpret_cast<uint4*>(base)[o] = make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();

    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;      // ldmatrix: lanes 0-15 address rows 0-15 of the low k half,
    const int lhalf = lane >> 4;                               //           lanes 16-31 rows 0-15 of the high k half
    const uint32_t buf_addr = (uint32_t)__cvta_generic_to_shared(buf0);
    const uint32_t lane_off = (uint32_t)((kGuard + row_w + lrow) * kRowBytes);
    const uint32_t st_off = (uint32_t)((kGuard + row_w + g) * kRowBytes + t4 * 4);

    // Shared-memory operand buffers and the four barriers of a tile: unit 0 reads buffer 0, unit 1 buffer 1, unit 2
    // buffer 0, the final diagonal sum reads P^T in buffer 1.  Every write to a buffer is separated from the previous
    // reads of the same buffer by at least one of the barriers below.
    while (tile < n_tiles) {
        const uint32_t cur_valid = valid;
        const int cur_t_first = t_first, cur_clip = clip;
#pragma unroll 1
        for (int u = 0; u < 3; ++u) {
            const float* sp = s_par + u * 6 * kC;
            const uint32_t abuf = (u & 1) ? kBufBytes : 0;
            // ---- a = bf16(snake(x, alpha0)) -> shared (the operand the time taps of every warp read)
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                const float2 al = *reinterpret_cast<const float2*>(sp + 2 * kC + n * 8 + t4 * 2);
                const float2 ia = *reinterpret_cast<const float2*>(sp + 3 * kC + n * 8 + t4 * 2);
#pragma unroll
                for (int i = 0; i < kMT; ++i)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float2 r = snake2(make_float2(xr[i][n][2 * h], xr[i][n][2 * h + 1]), al, ia);
                        *reinterpret_cast<uint32_t*>(buf0 + abuf + st_off + (i * 16 + h * 8) * kRowBytes + n * 16) = pack_bf16(r);
                    }
            }
            __syncthreads();
            // ---- conv_k7 (dilation d) + bias: M = time, N = 24 (3 n-tiles), K = 7 taps x 24 channels in 11 k-steps
            float acc[kMT][3][4];
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                const float2 cb = *reinterpret_cast<const float2*>(sp + n * 8 + t4 * 2);
#pragma unroll
                for (int i = 0; i < kMT; ++i) {
                    acc[i][n][0] = acc[i][n][2] = cb.x;
                    acc[i][n][1] = acc[i][n][3] = cb.y;
                }
            }
            {
                const int d_bytes = (u == 0 ? p.dil[0] : (u == 1 ? p.dil[1] : p.dil[2])) * kRowBytes;   // no dynamic index into the parameter struct
                const uint32_t a_lane = buf_addr + abuf + lane_off;
                const uint32_t* wc = w_conv + u * kConvFragWords + lane * 2;
#pragma unroll
                for (int ks = 0; ks < kKSteps; ++ks) {
                    // the two 8-channel chunks of this k-step: q = 2 ks (+1), tap q / 3, channels 8 (q % 3) ...
                    constexpr int kLast = 7 * 3 - 1;
                    const int qa = 2 * ks, qb = (2 * ks + 1 > kLast) ? kLast : 2 * ks + 1;    // chunk 21 is K padding (zero weights)
                    const int off_a = (qa / 3 - 3) * d_bytes + (qa % 3) * 16;
                    const int off_b = (qb / 3 - 3) * d_bytes + (qb % 3) * 16;
                    const uint32_t addr = a_lane + (uint32_t)(lhalf ? off_b : off_a);
                    uint2 bw[3];
#pragma unroll
                    for (int n = 0; n < 3; ++n) bw[n] = *reinterpret_cast<const uint2*>(wc + (ks * 3 + n) * 64);
                    uint32_t a0, a1, a2, a3;
                    ldmatrix_x4<0>(addr, a0, a1, a2, a3);
#pragma unroll
                    for (int n = 0; n < 3; ++n) mma_bf16(acc[0][n], a0, a1, a2, a3, bw[n].x, bw[n].y);
                    if constexpr (kMT > 1) {
                        ldmatrix_x4<16 * kRowBytes>(addr, a0, a1, a2, a3);
#pragma unroll
                        for (int n = 0; n < 3; ++n) mma_bf16(acc[1][n], a0, a1, a2, a3, bw[n].x, bw[n].y);
                    }
                    if constexpr (kMT > 2) {
                        ldmatrix_x4<32 * kRowBytes>(addr, a0, a1, a2, a3);
#pragma unroll
                        for (int n = 0; n < 3; ++n) mma_bf16(acc[2][n], a0, a1, a2, a3, bw[n].x, bw[n].y);
                    }
                    static_assert(kMT <= 3, "add an ldmatrix step");
                }
            }
            // ---- h = bf16(snake(acc, alpha1)) in registers = A fragments of the 1x1 conv; x += conv1x1(h) + bias
            {
                const uint32_t* wp = w_pw + u * kPwFragWords + lane * 2;
                uint2 pw0[3];
                uint32_t pw1[3];
#pragma unroll
                for (int n = 0; n < 3; ++n) {
                    pw0[n] = *reinterpret_cast<const uint2*>(wp + n * 64);
                    pw1[n] = wp[(3 + n) * 64];
                }
#pragma unroll
                for (int i = 0; i < kMT; ++i) {
                    uint32_t ha[3][2];
#pragma unroll
                    for (int n = 0; n < 3; ++n) {
                        const float2 al = *reinterpret_cast<const float2*>(sp + 4 * kC + n * 8 + t4 * 2);
                        const float2 ia = *reinterpret_cast<const float2*>(sp + 5 * kC + n * 8 + t4 * 2);
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            ha[n][h] = pack_bf16(snake2(make_float2(acc[i][n][2 * h], acc[i][n][2 * h + 1]), al, ia));
                    }
#pragma unroll
                    for (int n = 0; n < 3; ++n) {
                        const float2 pb = *reinterpret_cast<const float2*>(sp + kC + n * 8 + t4 * 2);
                        float o[4] = {pb.x, pb.y, pb.x, pb.y};
                        mma_bf16(o, ha[0][0], ha[0][1], ha[1][0], ha[1][1], pw0[n].x, pw0[n].y);
                        mma_bf16_k8(o, ha[2][0], ha[2][1], pw1[n]);
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            if (cur_valid & (1u << (2 * i + h))) {                  // rows outside the clip stay zero
                                const float2 r = fadd2(make_float2(xr[i][n][2 * h], xr[i][n][2 * h + 1]), make_float2(o[2 * h], o[2 * h + 1]));
                                xr[i][n][2 * h] = r.x;
                                xr[i][n][2 * h + 1] = r.y;
                            }
                    }
                }
            }
        }

        // ---- final: s = snake(x, alpha_f); Conv1d(24 -> 1, k7, pad 3) + tanh.  The conv is factored as
        //   P[t'][j] = sum_c s[t'][c] w[j][c]   (an MMA with the 7 taps in the N dimension, A = s straight from registers
        //                                        as a split-bf16 pair, 3 terms hi*Whi + lo*Whi + hi*Wlo: fp32-class)
        //   y[t] = sum_j P[t + j - 3][j]        (a diagonal sum through shared memory, P stored transposed)
        // (bf16 decode path: the MUFU sine / tanh.approx errors, ~5e-7 absolute / 2^-11 relative, are far below the bf16
        // operand rounding upstream.)
        float* pt = reinterpret_cast<float*>(buf0 + kBufBytes);     // buffer 1 is free: its last reader was unit 1's conv
        {
            const float* sp = s_par + 18 * kC;
            const uint4 wf0 = w_fin[lane], wf1 = w_fin[32 + lane];
#pragma unroll
            for (int i = 0; i < kMT; ++i) {
                uint32_t hi[3][2], lo[3][2];
#pragma unroll
                for (int n = 0; n < 3; ++n) {
                    const float2 al = *reinterpret_cast<const float2*>(sp + n * 8 + t4 * 2);
                    const float2 ia = *reinterpret_cast<const float2*>(sp + kC + n * 8 + t4 * 2);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float2 sv = snake2(make_float2(xr[i][n][2 * h], xr[i][n][2 * h + 1]), al, ia);
                        const __nv_bfloat162 hb = __floats2bfloat162_rn(sv.x, sv.y);
                        const float2 hf = __bfloat1622float2(hb);
                        hi[n][h] = *reinterpret_cast<const uint32_t*>(&hb);
                        lo[n][h] = pack_bf16(make_float2(sv.x - hf.x, sv.y - hf.y));
                    }
                }
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                mma_bf16(o, hi[0][0], hi[0][1], hi[1][0], hi[1][1], wf0.x, wf0.y);
                mma_bf16_k8(o, hi[2][0], hi[2][1], wf1.x);
                mma_bf16(o, lo[0][0], lo[0][1], lo[1][0], lo[1][1], wf0.x, wf0.y);
                mma_bf16_k8(o, lo[2][0], lo[2][1], wf1.x);
                mma_bf16(o, hi[0][0], hi[0][1], hi[1][0], hi[1][1], wf0.z, wf0.w);
                mma_bf16_k8(o, hi[2][0], hi[2][1], wf1.z);
                float* dst = pt + (2 * t4) * cfg::kPtStride + row_w + i * 16 + g;       // P^T[tap][row]
                dst[0] = o[0];
                dst[cfg::kPtStride] = o[1];
                dst[8] = o[2];
                dst[cfg::kPtStride + 8] = o[3];
            }
        }
        // the residual registers are dead: fetch the next tile's rows while the diagonal sums run
        tile += gridDim.x;
        if (tile < n_tiles) {
            clip = tile / tiles_per_clip;
            t_first = (tile - clip * tiles_per_clip) * kOut - kHalo;
            valid = load_rows<kMT>(xr, p.x + (long long)clip * p.T * kC, t_first + row_w + g, p.T, t4);
        }
        __syncthreads();
        for (int r = tid; r < kOut; r += kThreads) {
            const int t = cur_t_first + kHalo + r;
            if (t >= p.T) break;
            const float* src = pt + kHalo + r - 3;
            float acc = p.bias_f;
#pragma unroll
            for (int j = 0; j < 7; ++j) acc += src[j * cfg::kPtStride + j];
            float y;
            asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(acc));
            p.out[(long long)cur_clip * p.T + t] = y;
        }
    }
}

template <int kWarps, int kMT>
static int launch(const Params& p, int sms, cudaStream_t stream) {
    using cfg = Cfg<kWarps, kMT>;
    cudaError_t e = cudaFuncSetAttribute(decoder_tail_kernel<kWarps, kMT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)cfg::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    const long long n_tiles = (long long)l3ac_cdiv(p.T, cfg::kOut) * p.B;
    const long long ctas = (long long)sms * cfg::kCtasPerSm;
    decoder_tail_kernel<kWarps, kMT><<<(int)(n_tiles < ctas ? n_tiles : ctas), cfg::kThreads, cfg::kSmemBytes, stream>>>(p);
    return l3ac_launch_status();
}

}  // namespace tail
}  // namespace l3ac

extern "C" int l3ac_decoder_tail(const float* x, int B, int T, int C, const void* conv_frags, const float* conv_bias,
                                 const void* pw_frags, const float* pw_bias, const float* alpha0, const float* alpha1,
                                 const int* dilations, const float* alpha_f, const float* w_f, float bias_f, float* out,
                                 l3ac_stream_t stream) {
    using namespace l3ac::tail;
    L3AC_CHECK_ARG(x && conv_frags && conv_bias && pw_frags && pw_bias && alpha0 && alpha1 && dilations && alpha_f && w_f && out);
    L3AC_CHECK_ARG(B > 0 && B <= 65535 && T > 0);
    if (C != kC) return L3AC_EUNSUPPORTED;
    L3AC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    L3AC_CHECK_ARG((reinterpret_cast<uintptr_t>(conv_frags) & 15) == 0 && (reinterpret_cast<uintptr_t>(pw_frags) & 15) == 0);
    Params p{};
    p.x = x;
    p.conv_frags = (const uint32_t*)conv_frags;
    p.pw_frags = (const uint32_t*)pw_frags;
    p.conv_bias = conv_bias;
    p.pw_bias = pw_bias;
    p.alpha0 = alpha0;
    p.alpha1 = alpha1;
    p.alpha_f = alpha_f;
    p.w_f = w_f;
    p.bias_f = bias_f;
    int reach = 3;
    for (int i = 0; i < 3; ++i) {
        p.dil[i] = dilations[i];
        L3AC_CHECK_ARG(dilations[i] >= 1);
        if (3 * dilations[i] > kGuard) return L3AC_EUNSUPPORTED;   // a conv's reach must fit the guard rows
        reach += 3 * dilations[i];
    }
    if (reach > kHalo) return L3AC_EUNSUPPORTED;      // receptive field must fit the 42-sample halo (dilations 1,3,9)
    p.B = B;
    p.T = T;
    p.out = out;
    // 16 warps x 3 m-tiles (768-row tiles, 12 % halo recompute, one CTA per SM) once there is a tile for every SM;
    // 8 warps x 3 (384-row tiles, two CTAs per SM) for small launches.  Measured on 24 x 10 s clips: 538 vs 569 us.
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return L3AC_EDRIVER;
    if ((long long)l3ac_cdiv(T, Cfg<16, 3>::kOut) * B >= sms) return launch<16, 3>(p, sms, (cudaStream_t)stream);
    return launch<8, 3>(p, sms, (cudaStream_t)stream);
}
