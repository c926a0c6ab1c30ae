// Fused full-rate decoder tail on tcgen05 / TMEM (l3ac/modules.py:47-64,174-179,192-194): three Residual(LegacyUnit)
// blocks (dilations d0,d1,d2)
//   x += Conv1x1( snake( Conv_k7,dil d( snake(x, a0) ) + b, a1 ) ) + b'
// then Snake -> Conv1d(24 -> 1, k7, pad 3) -> tanh, in ONE persistent kernel: (B, T, 24) fp32 is read once from HBM and
// only the (B, T) waveform is written.
//
// The contractions run on the 5th-generation tensor cores with the A operand written by threads (umma.cuh):
//   * one CTA owns 896 consecutive samples of one clip (812 outputs + a 42-sample halo per side = the receptive field
//     3*(1+3+9)+3) as seven 128-row blocks; one THREAD owns one row, i.e. exactly the TMEM lane its tcgen05.ld reads;
//   * the fp32 RESIDUAL STREAM LIVES IN TMEM (24 columns of the row's lane): the 1x1 conv of every LegacyUnit is issued with
//     accumulate = 1 INTO it, so `x += conv1x1(h)` costs no instruction at all, nothing is carried in registers between
//     stages (64 registers per thread, 31 warps per SM), and a stage that needs x reads the row back and adds the biases
//     accumulated so far (kernel-parameter constants).  With the residual in registers (two rows per thread, 16 row-owner
//     warps) the same kernel needed 1843 instead of ~1200 warp-instructions per row and ran at 494 us per 24 clips; this
//     version takes 347 us (mma.sync kernel: 537 us);
//   * bf16(snake(x)) is stored as 8-channel planes [3 (+1 zero)][27 guard + 896 + 27][8]; the k7 conv of a block is 11
//     tcgen05.mma (M=128, N=32, K=16) whose A descriptors are the SAME tile shifted by (tap-3)*dilation rows -- nine of
//     them pair the taps (2p, 2p+1) of one 8-channel plane through LBO = dilation * 16 B, two cover tap 6 -- so the
//     conv's 168-long K axis costs 11 K-steps instead of 7 x 2 channel-padded ones and no im2col is ever materialised;
//   * the accumulator row comes back with tcgen05.ld, bias + snake run on registers, bf16(h) goes to the second plane
//     buffer (h = 0 on rows outside the clip, whose residual row must stay zero) and two more MMAs (K = 24 -> 32) do the 1x1 conv;
//   * the final Conv1d(24 -> 1, k7) has its 7 taps in the N dimension (P[t'][j] = s[t'] . w[j], A = snake(x) as a
//     split-bf16 pair, 3 terms hi*Whi + lo*Whi + hi*Wlo: fp32-class) + a diagonal sum y[t] = sum_j P[t+j-3][j] through
//     shared memory.
// Warps 0-27: row owners (warp & 3 = TMEM lane quadrant, warp >> 2 = block).  Warps 28-30 issue the MMAs through one elected
// lane each (28 / 29: the k7 convs and the final conv of the even / odd blocks, 30: the 1x1 convs, so that a block's 1x1
// conv never queues behind the other blocks' k7 convs and one issuer's barrier polls overlap the other's MMAs); all
// hand-overs are mbarriers per 128-row block, so the tensor pipe, the SFU (168 sines per sample) and the FMA pipe overlap
// across blocks without any CTA-wide barrier in the steady state.  Per-channel parameters live in the kernel-parameter
// constant bank: with one row per thread they are warp-uniform operands of the FMUL / FFMA instructions.
#include "common.cuh"
#include "umma.cuh"

#include <cstring>
#include <new>
#include <vector>

namespace l3ac {
namespace tailtc {

using namespace l3ac::umma;

constexpr int kC = 24;
constexpr int kHalo = 42;
constexpr int kGuard = 27;                       // largest conv reach (3 taps x dilation 9)
constexpr int kConvMmas = 11;
constexpr int kWConvBytes = kConvMmas * 1024;    // per unit and part: 11 x [2 halves][32 n][8 k] bf16
constexpr int kWPwBytes = 2 * 1024;              // per unit and part: 2 x [2][32][8]
constexpr int kWFinBytes = 2 * 2 * 512;          // {hi, lo} x 2 x [2][16][8]

// kSplit = false: bf16 operands (the decode side's default arithmetic), seven 128-row blocks per CTA.
// kSplit = true:  3-term split-bf16 operands (a = hi + lo, W = Whi + Wlo; hi*Whi + lo*Whi + hi*Wlo, fp32 accumulate):
//                 fp32-class results for precision="split".  Twice the operand planes and weights in shared memory, so four
//                 blocks per CTA; three times the MMAs, which then set the pace (the row owners' work grows by ~30 %).
template <bool kSplit>
struct Cfg {
    static constexpr int kBlocks = kSplit ? 4 : 7;
    static constexpr int kRows = kBlocks * 128;             // 896 / 512
    static constexpr int kOut = kRows - 2 * kHalo;          // 812 / 428
    static constexpr int kRowsTot = kRows + 2 * kGuard;
    static constexpr int kPlaneBytes = kRowsTot * 16;       // one 8-channel plane
    static constexpr int kBufBytes = 3 * kPlaneBytes;       // 24 channels; the K-padding plane (zeros) is shared by all buffers
    static constexpr int kParts = kSplit ? 2 : 1;           // operand parts: hi (, lo)
    static constexpr int kTerms = kSplit ? 3 : 1;
    static constexpr int kNumBufs = 2 * kParts;             // A hi, H hi (, A lo, H lo)
    static constexpr int kWBytes = kParts * 3 * (kWConvBytes + kWPwBytes) + kWFinBytes;
    static constexpr int kPtStride = kRows + 8;             // P^T[tap][row + 4]
    static constexpr int kPtBytes = 7 * kPtStride * 4;
    static constexpr int kNumBars = 5 * kBlocks;
    static constexpr int kWorkerWarps = 4 * kBlocks;        // one row per thread
    static constexpr int kConvWarp = kWorkerWarps, kPwWarp = kWorkerWarps + 2;     // MMA issuers: two for the k7 convs (even / odd blocks), one for the 1x1 convs
    static constexpr int kThreads = 32 * (kWorkerWarps + 3);
    static constexpr int kSmemBytes = kNumBufs * kBufBytes + kPlaneBytes + kWBytes + kPtBytes + 8 * kNumBars + 16;
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
    static_assert(kBufBytes % 16 == 0 && kWBytes % 16 == 0 && kPtBytes % 16 == 0, "16-byte carve-up");
    static_assert(kBlocks * 64 <= 512, "TMEM columns");
};

struct Params {
    const float* x;
    float* out;
    const uint8_t* wblob;      // device: conv[3] | pw[3] | fin, already in operand order
    int B, T;
    int dil[3];
    float bias_f;
    float conv_b[3][kC], cum_b[3][kC], a0[3][kC], ia0[3][kC], a1[3][kC], ia1[3][kC], af[kC], iaf[kC];      // cum_b[u] = pw_b[0] + .. + pw_b[u]
};

#ifdef L3AC_TAIL_TRACE
// Debug build only (tools/tail_trace.py): CTA 0 stamps clock64() at the hand-overs of its second tile.
__device__ unsigned long long g_tail_trace[6 * 256];
#define TAIL_TRACE(role, ev, unit, blk)                                                                              \
    do {                                                                                                             \
        if (blockIdx.x == 0 && it == 1 && trace_n < 255) {                                                           \
            g_tail_trace[(role) * 256 + 1 + trace_n++] = ((unsigned long long)clock64() << 16) | ((ev) << 8) | ((unit) << 4) | (blk); \
            g_tail_trace[(role) * 256] = trace_n;                                                                    \
        }                                                                                                            \
    } while (0)
#define TAIL_TRACE_DECL unsigned int trace_n = 0;
#else
#define TAIL_TRACE(role, ev, unit, blk) do {} while (0)
#define TAIL_TRACE_DECL
#endif

__device__ __forceinline__ float snake1(float v, float a, float ia) {
    const float s = __sinf(a * v);
    return fmaf(ia, s * s, v);
}

// this thread's row of the fp32 stream (zero outside the clip: every conv on the path zero-pads its input)
__device__ __forceinline__ bool load_row(float (&xr)[kC], const float* __restrict__ xb, int t, int T) {
    const bool ok = t >= 0 && t < T;
    const float4* src = reinterpret_cast<const float4*>(xb + (long long)(ok ? t : 0) * kC);
#pragma unroll
    for (int i = 0; i < kC / 4; ++i) {
        const float4 v = ok ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        xr[4 * i] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w;
    }
    return ok;
}

// 24 accumulator columns of this thread's row
__device__ __forceinline__ void ld_row24(uint32_t taddr, float (&v)[kC]) {
    uint32_t lo[16], hi[8];
    tmem_ld16(taddr, lo);
    tmem_ld8(taddr + 16, hi);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(lo[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[16 + i] = __uint_as_float(hi[i]);
}

// bf16 pair (hi) and, for the split operand, the bf16 pair of the remainders (lo)
template <bool kSplit>
__device__ __forceinline__ void pack_parts(float s0, float s1, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(s0, s1);
    if (kSplit) lo = pack_bf16x2(s0 - __uint_as_float(hi << 16), s1 - __uint_as_float(hi & 0xffff0000u));
}

template <int U, bool kSplit>
struct UnitPhase {
    using K = Cfg<kSplit>;
    // a = snake(x, alpha0) as bf16 (or the split pair) -> plane buffer A (dst_lo: the lo-part buffer)
    static __device__ __forceinline__ void snake_in(const Params& p, const float (&xr)[kC], uint32_t dst, uint32_t dst_lo) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t pk[4], pl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = 8 * c + 2 * i;
                pack_parts<kSplit>(snake1(xr[e], p.a0[U][e], p.ia0[U][e]), snake1(xr[e + 1], p.a0[U][e + 1], p.ia0[U][e + 1]), pk[i], pl[i]);
            }
            st_shared_v4(dst + c * K::kPlaneBytes, pk[0], pk[1], pk[2], pk[3]);
            if (kSplit) st_shared_v4(dst_lo + c * K::kPlaneBytes, pl[0], pl[1], pl[2], pl[3]);
        }
    }
    // h = snake(conv + bias, alpha1) -> plane buffer H   (8 accumulator columns at a time; rows outside the clip: h = 0,
    // so that the 1x1 conv adds nothing to their -- zero -- residual row)
    static __device__ __forceinline__ void snake_mid(const Params& p, uint32_t taddr, uint32_t dst, uint32_t dst_lo, bool valid) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t v[8];
            tmem_ld8(taddr + 8 * c, v);
            tmem_ld_wait();
            uint32_t pk[4], pl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = 8 * c + 2 * i;
                pack_parts<kSplit>(snake1(__uint_as_float(v[2 * i]) + p.conv_b[U][e], p.a1[U][e], p.ia1[U][e]),
                                   snake1(__uint_as_float(v[2 * i + 1]) + p.conv_b[U][e + 1], p.a1[U][e + 1], p.ia1[U][e + 1]), pk[i], pl[i]);
                pk[i] = valid ? pk[i] : 0u;
                if (kSplit) pl[i] = valid ? pl[i] : 0u;
            }
            st_shared_v4(dst + c * K::kPlaneBytes, pk[0], pk[1], pk[2], pk[3]);
            if (kSplit) st_shared_v4(dst_lo + c * K::kPlaneBytes, pl[0], pl[1], pl[2], pl[3]);
        }
    }
    // the residual stream after unit U: the TMEM row the 1x1 convs accumulated into + their biases (zero outside the clip)
    static __device__ __forceinline__ void load_x(const Params& p, uint32_t taddr, float (&xr)[kC], bool valid) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t v[8];
            tmem_ld8(taddr + 8 * c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) xr[8 * c + i] = valid ? __uint_as_float(v[i]) + p.cum_b[U][8 * c + i] : 0.f;
        }
    }
};

template <bool kSplit>
__global__ void __launch_bounds__(Cfg<kSplit>::kThreads, 1) decoder_tail_tc_kernel(const __grid_constant__ Params p) {
    using K = Cfg<kSplit>;
    constexpr int kBlocks = K::kBlocks, kPlaneBytes = K::kPlaneBytes, kBufBytes = K::kBufBytes, kRows = K::kRows, kOut = K::kOut;
    constexpr int kPtStride = K::kPtStride, kThreads = K::kThreads, kWorkerWarps = K::kWorkerWarps, kConvWarp = K::kConvWarp, kPwWarp = K::kPwWarp;
    constexpr int kTerms = K::kTerms;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    // plane buffers: A hi, H hi (, A lo, H lo), then the shared all-zero K-padding plane
    const uint32_t buf_a = sbase, buf_h = buf_a + kBufBytes;
    const uint32_t buf_a_lo = kSplit ? buf_h + kBufBytes : buf_a, buf_h_lo = kSplit ? buf_a_lo + kBufBytes : buf_h;
    const uint32_t zero_plane = sbase + K::kNumBufs * kBufBytes;
    const uint32_t fin_lo = kSplit ? buf_a_lo : buf_h;        // lo part of the final conv's operand (bf16 mode borrows buffer H)
    // weights: conv [part][unit] | pw [part][unit] | fin
    const uint32_t w_conv = zero_plane + kPlaneBytes, w_pw = w_conv + K::kParts * 3 * kWConvBytes, w_fin = w_pw + K::kParts * 3 * kWPwBytes;
    const uint32_t pt_addr = w_fin + kWFinBytes;
    float* pt = reinterpret_cast<float*>(smem + (pt_addr - sbase));
    const uint32_t bars = pt_addr + K::kPtBytes;
    const uint32_t a_ready = bars, d_ready = a_ready + 8 * kBlocks, h_ready = d_ready + 8 * kBlocks,
                   o_ready = h_ready + 8 * kBlocks, pt_ready = o_ready + 8 * kBlocks;
    const uint32_t tmem_slot = pt_ready + 8 * kBlocks;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - sbase));

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    // ---- once per CTA: zero the plane buffers (guard rows and the K-padding plane stay zero for good), stage the
    // weights, barriers, TMEM
    {
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < (K::kNumBufs * kBufBytes + kPlaneBytes) / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
        const uint4* src = reinterpret_cast<const uint4*>(p.wblob);
        uint4* dst = reinterpret_cast<uint4*>(smem + (w_conv - sbase));
        for (int i = tid; i < K::kWBytes / 16; i += kThreads) dst[i] = __ldg(src + i);
        for (int i = tid; i < K::kPtBytes / 4; i += kThreads) pt[i] = 0.f;
    }
    if (tid == 0) {
        for (int b = 0; b < kBlocks; ++b) {
            mbar_init(a_ready + 8 * b, 4);       // the four warps that own the block's rows
            mbar_init(d_ready + 8 * b, 1);       // tcgen05.commit
            mbar_init(h_ready + 8 * b, 4);
            mbar_init(o_ready + 8 * b, 1);
            mbar_init(pt_ready + 8 * b, 4);
        }
        fence_mbar_init();
    }
    if (warp == kConvWarp) tmem_alloc(tmem_slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    L3AC_PDL_SYNC();      // weight staging, zeroed planes, barriers and TMEM above may overlap the previous kernel's tail

    const int tiles_per_clip = (p.T + kOut - 1) / kOut;
    const int n_tiles = tiles_per_clip * p.B;

    constexpr uint64_t kDescHi = (uint64_t)(((128u >> 4) & 0x3FFF) | (1u << 14)) << 32;     // SBO = 128 B, sm_100 descriptor version
    constexpr uint32_t kBlockStep = (128 * 16) >> 4;                                        // 128 rows further down the planes
    const uint32_t lbo_plane = (uint32_t)(kPlaneBytes >> 4) << 16, lbo_w32 = (512u >> 4) << 16, lbo_w16 = (256u >> 4) << 16;
    // K halves (plane 2, zero plane): LBO = distance from the buffer's third plane to the shared zero plane
    auto lbo_zero = [&](uint32_t buf) { return ((zero_plane - (buf + 2 * kPlaneBytes)) >> 4) << 16; };

    if (warp == kConvWarp || warp == kConvWarp + 1) {
        // =============================================================== MMA issuers 1a / 1b: the k7 convs and the final conv of the
        // even / odd blocks.  One issuer needs ~770 cycles per block (barrier polls ~250, eleven MMAs through the uniform
        // datapath ~300, commit) for 440 cycles of tensor-pipe work (tools/tail_trace.py); two of them keep the pipe fed.
        const int b_first = warp - kConvWarp;
        // Everything this warp computes runs on the uniform datapath, whose dependent-instruction latency is long: the low
        // descriptor words (address and LBO fields) of a unit's eleven conv MMAs are computed once per unit and a block
        // only adds its row offset -- with the descriptors rebuilt per MMA the issue loop cost ~85 cycles per MMA
        // against the 40 the tensor core needs (tools/tail_trace.py, tools/umma_rate.cu).
        const bool leader = elect_one();
        const uint32_t idesc32 = make_idesc_bf16(32), idesc16 = make_idesc_bf16(16);
        TAIL_TRACE_DECL
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
#pragma unroll 1
            for (int u = 0; u < 3; ++u) {
                const int d = u == 0 ? p.dil[0] : (u == 1 ? p.dil[1] : p.dil[2]);
                uint32_t a_lo[K::kParts][kConvMmas];
#pragma unroll
                for (int part = 0; part < K::kParts; ++part) {
                    const uint32_t buf = part == 0 ? buf_a : buf_a_lo;
                    const uint32_t rows = (buf + kGuard * 16) >> 4;
#pragma unroll
                    for (int j = 0; j < kConvMmas; ++j) {
                        // j < 9: taps (2 (j / 3), + 1) of plane j % 3; j = 9: tap 6 of planes 0, 1; j = 10: tap 6 of plane 2 + the zero plane
                        const int plane = j < 9 ? j % 3 : (j == 9 ? 0 : 2);
                        const int tap = j < 9 ? 2 * (j / 3) : 6;
                        a_lo[part][j] = (rows + (uint32_t)(plane * (kPlaneBytes >> 4)) + (uint32_t)((tap - 3) * d)) |
                                        (j < 9 ? (uint32_t)d << 16 : (j == 9 ? lbo_plane : lbo_zero(buf)));
                    }
                }
                const uint32_t wc_lo = ((w_conv + u * kWConvBytes) >> 4) | lbo_w32;               // part 0 (hi); part 1 follows the three hi units
#pragma unroll 1
                for (int b = b_first; b < kBlocks; b += 2) {
                    if (b > 0) mbar_wait(a_ready + 8 * (b - 1), u & 1);          // the taps reach into both neighbour blocks
                    mbar_wait(a_ready + 8 * b, u & 1);
                    if (b + 1 < kBlocks) mbar_wait(a_ready + 8 * (b + 1), u & 1);
                    tc_fence_after();
                    if (leader && b_first == 0) TAIL_TRACE(4, 10, u, b);
                    if (leader) {
                        const uint32_t dcol = tmem_base + 64 * b + 32, boff = kBlockStep * b;
#pragma unroll
                        for (int term = 0; term < kTerms; ++term) {                              // hi*Whi (, lo*Whi, hi*Wlo)
                            const uint32_t wsel = wc_lo + (term == 2 ? (3 * kWConvBytes) >> 4 : 0);
#pragma unroll
                            for (int j = 0; j < kConvMmas; ++j)
                                tc_mma_bf16(dcol, kDescHi | (a_lo[term == 1 ? K::kParts - 1 : 0][j] + boff), kDescHi | (wsel + j * (1024 >> 4)), idesc32,
                                            (term | j) ? 1u : 0u);
                        }
                        TAIL_TRACE(4, 11, u, b);
                        tc_commit(d_ready + 8 * b);
                        TAIL_TRACE(4, 13, u, b);
                    }
                    __syncwarp();
                }
            }
            // final conv: P = s . Wf^T with s = hi (buffer A) + lo (fin_lo), Wf = hi + lo
            const uint32_t wf_lo = (w_fin >> 4) | lbo_w16;
#pragma unroll 1
            for (int b = b_first; b < kBlocks; b += 2) {
                // (only block b's rows are read, but waiting for the neighbours as well keeps every d_ready(b) phase behind the
                // a_ready phases of blocks b-1 .. b+1: the row owners' parity waits on a neighbour's d_ready rely on that)
                if (b > 0) mbar_wait(a_ready + 8 * (b - 1), 1);
                mbar_wait(a_ready + 8 * b, 1);
                if (b + 1 < kBlocks) mbar_wait(a_ready + 8 * (b + 1), 1);
                tc_fence_after();
                if (leader && b_first == 0) TAIL_TRACE(4, 10, 3, b);
                if (leader) {
                    const uint32_t ra = ((buf_a + kGuard * 16) >> 4) + kBlockStep * b, rl = ((fin_lo + kGuard * 16) >> 4) + kBlockStep * b;
                    const uint32_t dcol = tmem_base + 64 * b + 32;
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t arow = term == 1 ? rl : ra;
                        const uint32_t abuf = term == 1 ? fin_lo : buf_a;
                        const uint32_t wsel = wf_lo + (term == 2 ? (1024 >> 4) : 0);
#pragma unroll
                        for (int m = 0; m < 2; ++m)
                            tc_mma_bf16(dcol, kDescHi | ((arow + m * (2 * kPlaneBytes >> 4)) | (m == 0 ? lbo_plane : lbo_zero(abuf))),
                                        kDescHi | (wsel + m * (512 >> 4)), idesc16, (term | m) ? 1u : 0u);
                    }
                    tc_commit(d_ready + 8 * b);
                }
                __syncwarp();
            }
        }
    } else if (warp == kPwWarp) {
        // =============================================================== MMA issuer 2: the 1x1 convs, as soon as a block's h is written
        // (its own warp: behind the conv issuer's program order the 1x1 convs of a unit would wait for all eight k7 convs)
        const bool leader = elect_one();
        const uint32_t idesc32 = make_idesc_bf16(32);
        TAIL_TRACE_DECL
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
#pragma unroll 1
            for (int u = 0; u < 3; ++u) {
                const uint32_t wp_lo = ((w_pw + u * kWPwBytes) >> 4) | lbo_w32;                   // part 0 (hi); part 1 follows the three hi units
#pragma unroll 1
                for (int b = 0; b < kBlocks; ++b) {
                    mbar_wait(h_ready + 8 * b, (it + u) & 1);
                    tc_fence_after();
                    if (leader) TAIL_TRACE(5, 12, u, b);
                    if (leader) {
#pragma unroll
                        for (int term = 0; term < kTerms; ++term) {
                            const uint32_t hbuf = term == 1 ? buf_h_lo : buf_h;
                            const uint32_t a0 = ((hbuf + kGuard * 16) >> 4) + kBlockStep * b;
                            const uint32_t wsel = wp_lo + (term == 2 ? (3 * kWPwBytes) >> 4 : 0);
#pragma unroll
                            for (int m = 0; m < 2; ++m)
                                tc_mma_bf16(tmem_base + 64 * b, kDescHi | ((a0 + m * (2 * kPlaneBytes >> 4)) | (m == 0 ? lbo_plane : lbo_zero(hbuf))),
                                            kDescHi | (wsel + m * (1024 >> 4)), idesc32, 1u);     // accumulates INTO the residual row: x += conv1x1(h)
                        }
                        tc_commit(o_ready + 8 * b);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < kWorkerWarps) {
        // =============================================================== row owners: one row of one block per thread.  The fp32
        // residual row lives in TMEM columns [64 b, 64 b + 24) of the thread's lane: the 1x1 convs accumulate into it, a stage
        // that needs it reads it back (+ the accumulated biases), nothing is carried in registers between stages.
        const int blk = warp >> 2, quad = warp & 3;
        const int r = 128 * blk + 32 * quad + lane;
        const uint32_t tx = tmem_base + ((uint32_t)(quad * 32) << 16) + 64 * blk, td = tx + 32;
        const uint32_t row_off = (uint32_t)(kGuard + r) * 16;
        float xr[kC];
        bool valid = false;
        TAIL_TRACE_DECL
        const bool tracer = quad == 0 && lane == 0 && (blk & 1) == 0;
        [[maybe_unused]] const int wg = blk >> 1;                     // (trace role)
        int tile = blockIdx.x;
        int clip = 0, t_first = 0;
        if (tile < n_tiles) {
            clip = tile / tiles_per_clip;
            t_first = (tile - clip * tiles_per_clip) * kOut - kHalo;
            valid = load_row(xr, p.x + (long long)clip * p.T * kC, t_first + r, p.T);
        }
        int it = 0;
        for (; tile < n_tiles; ++it) {
            const int cur_clip = clip, cur_t_first = t_first;
            const bool cur_valid = valid;
            // ---- the tile's rows -> TMEM; unit 0's operand straight from the registers
            if (tracer) TAIL_TRACE(wg, 1, 0, blk);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                uint32_t v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(xr[8 * c + i]);
                tmem_st8(tx + 8 * c, v);
            }
            UnitPhase<0, kSplit>::snake_in(p, xr, buf_a + row_off, buf_a_lo + row_off);
            tmem_st_wait();
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready + 8 * blk);
            if (tracer) TAIL_TRACE(wg, 2, 0, blk);
#define L3AC_TAIL_UNIT(U)                                                                                     \
            {                                                                                                 \
                mbar_wait_tag(d_ready + 8 * blk, U & 1, 10);                                                  \
                tc_fence_after();                                                                             \
                if (tracer) TAIL_TRACE(wg, 3, U, blk);                                                        \
                UnitPhase<U, kSplit>::snake_mid(p, td, buf_h + row_off, buf_h_lo + row_off, cur_valid);       \
                fence_async_smem();                                                                           \
                tc_fence_before();                                                                            \
                __syncwarp();                                                                                 \
                if (lane == 0) mbar_arrive(h_ready + 8 * blk);                                                \
                if (tracer) TAIL_TRACE(wg, 4, U, blk);                                                        \
                mbar_wait_tag(o_ready + 8 * blk, (it + U) & 1, 11);                                           \
                /* This block's rows of buffer A are also read by the k7 convs of both neighbour blocks (other issuers than   */ \
                /* the one whose commit was seen above): they must be through before the rows are overwritten.  Neither      */ \
                /* barrier can be a phase ahead: every conv of a neighbour block waits for THIS block's next a_ready first.  */ \
                if (blk + 1 < kBlocks) mbar_wait_tag(d_ready + 8 * (blk + 1), U & 1, 12);                     \
                if (blk > 0) mbar_wait_tag(d_ready + 8 * (blk - 1), U & 1, 13);                               \
                tc_fence_after();                                                                             \
                if (tracer) TAIL_TRACE(wg, 5, U, blk);                                                        \
                UnitPhase<U, kSplit>::load_x(p, tx, xr, cur_valid);                                           \
                if (tracer) TAIL_TRACE(wg, 1, U + 1, blk);                                                    \
                if (U < 2) {                                                                                  \
                    UnitPhase<(U < 2 ? U + 1 : 0), kSplit>::snake_in(p, xr, buf_a + row_off, buf_a_lo + row_off); \
                } else {    /* final conv operand: s = snake(x, alpha_f) as a split pair, hi -> buffer A, lo -> fin_lo */ \
                    _Pragma("unroll") for (int c = 0; c < 3; ++c) {                                           \
                        uint32_t ph[4], pl[4];                                                                \
                        _Pragma("unroll") for (int i = 0; i < 4; ++i) {                                       \
                            const int e = 8 * c + 2 * i;                                                      \
                            pack_parts<true>(snake1(xr[e], p.af[e], p.iaf[e]), snake1(xr[e + 1], p.af[e + 1], p.iaf[e + 1]), ph[i], pl[i]); \
                        }                                                                                     \
                        st_shared_v4(buf_a + row_off + c * kPlaneBytes, ph[0], ph[1], ph[2], ph[3]);          \
                        st_shared_v4(fin_lo + row_off + c * kPlaneBytes, pl[0], pl[1], pl[2], pl[3]);         \
                    }                                                                                         \
                }                                                                                             \
                fence_async_smem();                                                                           \
                tc_fence_before();                                                                            \
                __syncwarp();                                                                                 \
                if (lane == 0) mbar_arrive(a_ready + 8 * blk);                                                \
                if (tracer) TAIL_TRACE(wg, 2, U + 1, blk);                                                    \
            }
            L3AC_TAIL_UNIT(0)
            L3AC_TAIL_UNIT(1)
            L3AC_TAIL_UNIT(2)
#undef L3AC_TAIL_UNIT
            // the row registers are dead: fetch the next tile's row while the final conv runs
            tile += gridDim.x;
            if (tile < n_tiles) {
                clip = tile / tiles_per_clip;
                t_first = (tile - clip * tiles_per_clip) * kOut - kHalo;
                valid = load_row(xr, p.x + (long long)clip * p.T * kC, t_first + r, p.T);
            }
            // P row -> P^T in shared memory
            {
                mbar_wait(d_ready + 8 * blk, 1);
                tc_fence_after();
                if (tracer) TAIL_TRACE(wg, 7, 3, blk);
                uint32_t v[8];
                tmem_ld8(td, v);
                tmem_ld_wait();
                float* dst = pt + 4 + r;
#pragma unroll
                for (int j = 0; j < 7; ++j) dst[j * kPtStride] = __uint_as_float(v[j]);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(pt_ready + 8 * blk);
            }
            // y[t] = tanh(bias + sum_j P[t + j - 3][j])
            {
                if (blk > 0) mbar_wait(pt_ready + 8 * (blk - 1), it & 1);
                mbar_wait(pt_ready + 8 * blk, it & 1);
                if (blk + 1 < kBlocks) mbar_wait(pt_ready + 8 * (blk + 1), it & 1);
                const int t = cur_t_first + r;
                if (r >= kHalo && r < kRows - kHalo && t < p.T) {
                    const float* src = pt + 4 + r - 3;
                    float acc = p.bias_f;
#pragma unroll
                    for (int j = 0; j < 7; ++j) acc += src[j * kPtStride + j];
                    float y;
                    if (kSplit) y = tanhf(acc);                                  // fp32-class mode: libm-accurate tanh
                    else asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(acc));
                    p.out[(long long)cur_clip * p.T + t] = y;
                }
                if (tracer) TAIL_TRACE(wg, 8, 3, blk);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kConvWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tailtc
}  // namespace l3ac

#ifdef L3AC_TAIL_TRACE
extern "C" int l3ac_debug_tail_trace(unsigned long long* host_buf) {      // host_buf: 6 * 256 entries
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_buf, l3ac::tailtc::g_tail_trace, 6 * 256 * sizeof(unsigned long long));
    return 0;
}
#endif

struct l3ac_tail_plan {
    l3ac::tailtc::Params params;
    void* dev_blob;          // bf16 operand weights
    void* dev_blob_split;    // (hi, lo) operand weights of the 3-term split variant
    int device;
};

static inline uint16_t bf16_bits(float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    uint16_t b;
    memcpy(&b, &h, 2);
    return b;
}
static inline float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
// part 0: bf16(w); part 1: bf16(w - bf16(w))
static inline uint16_t bf16_part(float w, int part) { return bf16_bits(part == 0 ? w : w - bf16_round(w)); }

// Weights in operand order: conv [part][unit][mma j][half h][n 32][k 8] | pw [part][unit][m][h][n 32][k 8] | fin {hi, lo}[m][h][n 16][k 8]
static std::vector<uint16_t> tail_weight_blob(int parts, const float* conv_w, const float* pw_w, const float* w_f) {
    using namespace l3ac::tailtc;
    const size_t total = ((size_t)parts * 3 * (kWConvBytes + kWPwBytes) + kWFinBytes) / 2;
    std::vector<uint16_t> blob(total, 0);
    // W[u][n][ch][tap] at conv_w[((u * 24 + n) * 24 + ch) * 7 + tap]
    for (int part = 0; part < parts; ++part)
        for (int u = 0; u < 3; ++u)
            for (int j = 0; j < kConvMmas; ++j)
                for (int h = 0; h < 2; ++h) {
                    int tap, plane;
                    if (j < 9) { tap = 2 * (j / 3) + h; plane = j % 3; }
                    else if (j == 9) { tap = 6; plane = h; }
                    else { tap = 6; plane = h == 0 ? 2 : -1; }
                    for (int n = 0; n < kC && plane >= 0; ++n)
                        for (int k = 0; k < 8; ++k)
                            blob[(size_t)(part * 3 + u) * kWConvBytes / 2 + j * 512 + h * 256 + n * 8 + k] =
                                bf16_part(conv_w[(((size_t)u * kC + n) * kC + 8 * plane + k) * 7 + tap], part);
                }
    const size_t pw0 = (size_t)parts * 3 * kWConvBytes / 2;
    for (int part = 0; part < parts; ++part)
        for (int u = 0; u < 3; ++u)
            for (int m = 0; m < 2; ++m)
                for (int h = 0; h < 2; ++h)
                    for (int n = 0; n < kC; ++n)
                        for (int k = 0; k < 8; ++k) {
                            const int ch = 16 * m + 8 * h + k;
                            if (ch < kC)
                                blob[pw0 + (size_t)(part * 3 + u) * kWPwBytes / 2 + m * 512 + h * 256 + n * 8 + k] =
                                    bf16_part(pw_w[((size_t)u * kC + n) * kC + ch], part);
                        }
    const size_t fin0 = pw0 + (size_t)parts * 3 * kWPwBytes / 2;
    for (int part = 0; part < 2; ++part)                    // hi, lo
        for (int m = 0; m < 2; ++m)
            for (int h = 0; h < 2; ++h)
                for (int n = 0; n < 7; ++n)
                    for (int k = 0; k < 8; ++k) {
                        const int ch = 16 * m + 8 * h + k;
                        if (ch < kC) blob[fin0 + part * 512 + m * 256 + h * 128 + n * 8 + k] = bf16_part(w_f[n * kC + ch], part);
                    }
    return blob;
}

extern "C" int l3ac_tail_plan_create(const float* conv_w, const float* conv_b, const float* pw_w, const float* pw_b,
                                     const float* alpha0, const float* alpha1, const int* dilations, const float* alpha_f,
                                     const float* w_f, float bias_f, int C, l3ac_tail_plan** plan_out) {
    using namespace l3ac::tailtc;
    L3AC_CHECK_ARG(conv_w && conv_b && pw_w && pw_b && alpha0 && alpha1 && dilations && alpha_f && w_f && plan_out);
    if (C != kC) return L3AC_EUNSUPPORTED;
    int reach = 3;
    for (int i = 0; i < 3; ++i) {
        L3AC_CHECK_ARG(dilations[i] >= 1);
        if (3 * dilations[i] > kGuard) return L3AC_EUNSUPPORTED;      // a conv's reach must fit the guard rows
        reach += 3 * dilations[i];
    }
    if (reach > kHalo) return L3AC_EUNSUPPORTED;                      // receptive field must fit the 42-sample halo
    const std::vector<uint16_t> blob = tail_weight_blob(1, conv_w, pw_w, w_f), blob2 = tail_weight_blob(2, conv_w, pw_w, w_f);
    if (blob.size() * 2 != (size_t)Cfg<false>::kWBytes || blob2.size() * 2 != (size_t)Cfg<true>::kWBytes) return L3AC_EINVAL;
    l3ac_tail_plan* plan = new (std::nothrow) l3ac_tail_plan();
    if (!plan) return L3AC_EINVAL;
    plan->dev_blob = plan->dev_blob_split = nullptr;
    if (cudaGetDevice(&plan->device) != cudaSuccess) { delete plan; return L3AC_EDRIVER; }
    cudaError_t e = cudaMalloc(&plan->dev_blob, blob.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&plan->dev_blob_split, blob2.size() * 2);
    if (e == cudaSuccess) e = cudaMemcpy(plan->dev_blob, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(plan->dev_blob_split, blob2.data(), blob2.size() * 2, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(plan->dev_blob);
        cudaFree(plan->dev_blob_split);
        delete plan;
        return (int)e;
    }
    Params& p = plan->params;
    p = Params{};
    p.wblob = static_cast<const uint8_t*>(plan->dev_blob);
    p.bias_f = bias_f;
    for (int u = 0; u < 3; ++u) {
        p.dil[u] = dilations[u];
        for (int c = 0; c < kC; ++c) {
            p.conv_b[u][c] = conv_b[u * kC + c];
            p.cum_b[u][c] = pw_b[u * kC + c] + (u > 0 ? p.cum_b[u - 1][c] : 0.f);
            p.a0[u][c] = alpha0[u * kC + c];
            p.ia0[u][c] = 1.0f / (alpha0[u * kC + c] + l3ac::kEps);
            p.a1[u][c] = alpha1[u * kC + c];
            p.ia1[u][c] = 1.0f / (alpha1[u * kC + c] + l3ac::kEps);
        }
    }
    for (int c = 0; c < kC; ++c) {
        p.af[c] = alpha_f[c];
        p.iaf[c] = 1.0f / (alpha_f[c] + l3ac::kEps);
    }
    *plan_out = plan;
    return L3AC_OK;
}

extern "C" int l3ac_tail_plan_destroy(l3ac_tail_plan* plan) {
    if (!plan) return L3AC_OK;
    cudaFree(plan->dev_blob);
    cudaFree(plan->dev_blob_split);
    delete plan;
    return L3AC_OK;
}

template <bool kSplit>
static int launch_tail(const l3ac_tail_plan* plan, const float* x, int B, int T, float* out, l3ac_stream_t stream) {
    using namespace l3ac::tailtc;
    using K = Cfg<kSplit>;
    L3AC_CHECK_ARG(plan && x && out && B > 0 && T > 0);
    L3AC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return L3AC_EDRIVER;
    L3AC_CHECK_ARG(dev == plan->device);
    const long long n_tiles = (long long)l3ac_cdiv(T, K::kOut) * B;
    L3AC_CHECK_ARG(n_tiles < (1LL << 30));
    Params p = plan->params;
    p.wblob = static_cast<const uint8_t*>(kSplit ? plan->dev_blob_split : plan->dev_blob);
    p.x = x;
    p.out = out;
    p.B = B;
    p.T = T;
    cudaError_t e = cudaFuncSetAttribute(decoder_tail_tc_kernel<kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    const int sms = l3ac_sm_count();
    l3ac_launch(decoder_tail_tc_kernel<kSplit>, dim3((int)(n_tiles < sms ? n_tiles : sms)), dim3(K::kThreads), K::kSmemBytes, (cudaStream_t)stream, p);
    return l3ac_launch_status();
}

extern "C" int l3ac_decoder_tail_tc(const l3ac_tail_plan* plan, const float* x, int B, int T, float* out, l3ac_stream_t stream) {
    return launch_tail<false>(plan, x, B, T, out, stream);
}

extern "C" int l3ac_decoder_tail_tc_split(const l3ac_tail_plan* plan, const float* x, int B, int T, float* out, l3ac_stream_t stream) {
    return launch_tail<true>(plan, x, B, T, out, stream);
}
