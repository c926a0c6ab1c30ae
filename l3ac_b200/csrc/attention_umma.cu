// Block-local causal attention on tcgen05 / TMEM (flash-style, online softmax).  Same semantics and interface as
// attention_tc.cu (LocalAttention of local-attention==1.11.2 as configured at l3ac/local_trans.py:34-38 + the
// DynamicPositionBias Toeplitz table, l3ac/local_trans.py:43):
//   query p attends keys j with max(0, (p/w - 1) w) <= j <= p;  logit = q.k / sqrt(32) + bias[h][p - j].
//
// One CTA = one (batch, head, 128-query tile); key/value tiles of 128 keys stream through a two-stage TMA ring.
//   S = Q K^T      two tcgen05.mma (M = 128, N = 128, K = 16) into 128 TMEM columns; Q and K are 8-channel planes
//                  [d / 8][row][8] (umma.cuh): a TMA box of [128 rows x 8 channels] of the (B T, 3 H D) tensor IS one plane;
//   softmax        ONE THREAD PER QUERY ROW (= its TMEM lane): the row maximum and sum need no shuffles.  Pass 1 reads the
//                  row in 32-column chunks, applies scale + Toeplitz bias (+ causal / window mask on the few partial tiles)
//                  and writes the result back to TMEM; pass 2 re-reads it, exponentiates in the base-2 domain (one FADD +
//                  MUFU.EX2 per probability) and stores bf16 P as K-major planes [key / 8][row][8];
//   O_j = P V      eight tcgen05.mma (N = 48) with V in its NATURAL layout as an MN-major B operand (planes
//                  [d / 8][key][8], tools/umma_probe_mn.cu); a constant fifth plane holds a ones column, so column 32 of
//                  the result is the row sum of the (bf16-rounded) probabilities -- the normaliser comes out of the
//                  tensor core instead of 128 FADDs per row and tile;
//   the per-tile result is folded into fp32 register accumulators (o = o * 2^(m_old - m_new) + O_j) while the next tile's
//   logits are already in TMEM.
// Warps 0-7: softmax, TWO threads per query row (64 keys of the tile each, one row-maximum exchange through shared memory per
// tile): four softmax warps per scheduler with two CTAs per SM hide the TMEM / MUFU / LDS latencies that one warp per
// scheduler could not (4100 -> cycles per tile measured with one thread per row).  Warp 8: MMA issuer (one elected lane; it
// issues S(j+1) before P(j) V(j)), warp 9: TMA loader.  Two CTAs per SM (256 TMEM columns, ~90 KB shared memory each).
// SPLIT (encode side): q/k/v arrive as (hi, lo) bf16 pairs and both products are issued as hi*hi + lo*hi + hi*lo, P is
// split the same way -- fp32-class logits and outputs (token indices are sensitive to bf16 rounding).
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace l3ac {
namespace attu {

using namespace l3ac::umma;

constexpr int kD = 32;
constexpr int kBQ = 128;
constexpr int kBK = 128;
constexpr int kPlane = 128 * 16;              // one 8-element group x 128 rows
constexpr int kQBytes = 4 * kPlane;
constexpr int kKBytes = 4 * kPlane;
constexpr int kVBytes = 6 * kPlane;           // 4 V planes + ones-column plane + zero plane (N = 48)
constexpr int kPBytes = 16 * kPlane;
constexpr int kPadLo = 256, kPadHi = 128;     // slack around the bias table: index q - k of masked elements stays in bounds
constexpr int kSoftmaxWarps = 8, kMmaWarp = 8, kLoadWarp = 9;      // two softmax threads per query row (64 keys of a tile each)
constexpr int kThreads = 32 * 10;
// TMEM: S buffers of 128 columns, then O (48).  Plain: one S buffer, 256 columns, two CTAs per SM.  SPLIT: shared memory
// allows one CTA per SM only, so S is double-buffered (512 columns): S(j+1) is computed while the softmax warps work on S(j).

template <bool SPLIT>
struct Smem {
    static constexpr int kParts = SPLIT ? 2 : 1;
    // K/V ring depth.  SPLIT (one CTA per SM, 3x the MMA work per tile): three stages, the load of tile j+2 must not wait for
    // P V of tile j.  Plain: two stages keep two CTAs per SM resident, and the other CTA covers the gap.
    static constexpr int kStages = SPLIT ? 3 : 2;
    static constexpr int q = 0;
    static constexpr int k = q + kParts * kQBytes;                              // [stage][part]
    static constexpr int v = k + kStages * kParts * kKBytes;                    // [stage]: hi (6 planes), lo (4 planes)
    static constexpr int kVStage = kVBytes + (SPLIT ? kKBytes : 0);
    static constexpr int p = v + kStages * kVStage;                             // [part]
    static constexpr int table = p + kParts * kPBytes;
    // bias table with slack, 64-wide window maxima of the table, barriers, row-maximum exchange
    static size_t bytes(int window) { return (size_t)table + (size_t)(2 * window + kPadLo + kPadHi) * 4 + (size_t)2 * window * 4 + 8 * 16 + 16 + 2 * 2 * 128 * 4; }
};

#ifdef L3AC_ATTU_TRACE
// Debug build only (tools/attn_trace.py): one CTA stamps clock64() at its pipeline hand-overs.
__device__ unsigned long long g_attu_trace[4 * 512];
#define ATTU_TRACE(role, ev, j)                                                                                          \
    do {                                                                                                                 \
        if (blockIdx.x == 0 && blockIdx.y == 2 && trace_n < 511) {                                  \
            g_attu_trace[(role) * 512 + 1 + trace_n++] = ((unsigned long long)clock64() << 16) | ((ev) << 8) | (j);     \
            g_attu_trace[(role) * 512] = trace_n;                                                                       \
        }                                                                                                                \
    } while (0)
#define ATTU_TRACE_DECL unsigned int trace_n = 0;
#else
#define ATTU_TRACE(role, ev, j) do {} while (0)
#define ATTU_TRACE_DECL
#endif

__device__ __forceinline__ float ex2f(float x) {       // 2^x on MUFU.EX2 (ex2(-inf) = +0)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool SPLIT, int OUT>
__global__ void __launch_bounds__(kThreads, SPLIT ? 1 : 2)
local_attention_umma_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                            const __nv_bfloat16* __restrict__ qkv_hi, const __nv_bfloat16* __restrict__ qkv_lo,
                            const float* __restrict__ bias_table, int B, int T, int H, int window, void* __restrict__ out,
                            void* __restrict__ out_lo) {
    using L = Smem<SPLIT>;
    constexpr int kParts = L::kParts;
    constexpr int kStages = L::kStages;
    constexpr int kSBufs = SPLIT ? 2 : 1;
    constexpr int kTmemCols = SPLIT ? 512 : 256;
    constexpr int kOCol = 128 * kSBufs;
    auto sbuf = [](int t) { return kSBufs == 2 ? (t & 1) : 0; };                 // S buffer of tile t and the phase parity of its
    auto sphase = [](int t) { return kSBufs == 2 ? ((t >> 1) & 1) : (t & 1); };  // s_full / s_free barrier for that tile
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t q_s = sbase + L::q, k_s = sbase + L::k, v_s = sbase + L::v, p_s = sbase + L::p;
    float* s_table = reinterpret_cast<float*>(smem + L::table) + kPadLo;
    float* s_wmax = s_table + 2 * window + kPadHi;      // s_wmax[i] = max(table[i .. i+63]), 0 <= i <= 2w - 64
    const uint32_t bars = (sbase + L::table + (uint32_t)(2 * window + kPadLo + kPadHi) * 4 + (uint32_t)2 * window * 4 + 7u) & ~7u;
    const uint32_t kv_full = bars, kv_empty = bars + 32, s_full = bars + 64, s_free = bars + 80, p_full = bars + 96,
                   pv_done = bars + 104, tmem_slot = bars + 112;
    float* s_mx = reinterpret_cast<float*>(smem + (bars + 128 - sbase));          // [tile parity][column half][row]: row-maximum exchange
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - sbase));

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    // grid: x = (batch, head) fastest, y = query tile from the LAST one down -- late tiles see the most keys, so the heavy CTAs
    // are scheduled first and the light ones fill the tail of the launch
    const int q0 = ((int)gridDim.y - 1 - (int)blockIdx.y) * kBQ, h = blockIdx.x % H, b = blockIdx.x / H;
    const long long ld = 3LL * H * kD;
    const __nv_bfloat16* base_hi = qkv_hi + (long long)b * T * ld + h * kD;
    const __nv_bfloat16* base_lo = SPLIT ? qkv_lo + (long long)b * T * ld + h * kD : nullptr;
    const int koff = H * kD, voff = 2 * H * kD;
    const int q_last = min(q0 + kBQ, T) - 1;
    int k_begin = (q0 / window - 1) * window;
    if (k_begin < 0) k_begin = 0;
    const int n_tiles = (q_last - k_begin) / kBK + 1;
    ATTU_TRACE_DECL
    if (tid == 0) ATTU_TRACE(3, 1, 0);

    // ---- set-up: bias table (base-2 domain, zero slack on both sides), the constant V planes, Q, barriers, TMEM
    for (int i = tid; i < 2 * window + kPadLo + kPadHi; i += kThreads) {
        const int idx = i - kPadLo;
        s_table[idx] = (idx >= 0 && idx < 2 * window) ? __ldg(bias_table + (long long)h * 2 * window + idx) * 1.4426950408889634f : -INFINITY;
    }
    for (int i = tid; i < kStages * 2 * 128; i += kThreads) {
        const int st = i / 256, r = i & 255;                    // rows 0..127: ones-column plane, 128..255: zero plane
        uint4* dst = reinterpret_cast<uint4*>(smem + L::v + st * L::kVStage + 4 * kPlane) + r;
        *dst = make_uint4(r < 128 ? 0x00003F80u : 0u, 0u, 0u, 0u);      // bf16 {1, 0, 0, 0, 0, 0, 0, 0}
    }
    L3AC_PDL_SYNC();      // the bias table and the constant planes above are weights; q / k / v below come from the previous kernel
    for (int i = tid; i < 128 * 4; i += kThreads) {
        const int r = i >> 2, g = i & 3;
        const int t = min(q0 + r, T - 1);
        cp_async16(q_s + g * kPlane + r * 16, base_hi + (long long)t * ld + g * 8);
        if (SPLIT) cp_async16(q_s + kQBytes + g * kPlane + r * 16, base_lo + (long long)t * ld + g * 8);
    }
    cp_async_wait_all();
    __syncthreads();
    {   // 64-wide sliding maxima of the bias table in two levels (8 + 8 reads per entry); the 8-wide level lives in the P buffer
        float* m8 = reinterpret_cast<float*>(smem + L::p);
        for (int i = tid; i <= 2 * window - 8; i += kThreads) {
            float m = s_table[i];
#pragma unroll
            for (int k = 1; k < 8; ++k) m = fmaxf(m, s_table[i + k]);
            m8[i] = m;
        }
        __syncthreads();
        for (int i = tid; i <= 2 * window - 64; i += kThreads) {
            float m = m8[i];
#pragma unroll
            for (int k = 1; k < 8; ++k) m = fmaxf(m, m8[i + 8 * k]);
            s_wmax[i] = m;
        }
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(kv_full + 8 * s, 1);        // the loader's arrive.expect_tx (+ the TMA transaction bytes)
            mbar_init(kv_empty + 8 * s, 1);       // tcgen05.commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(s_full + 8 * s, 1);
            mbar_init(s_free + 8 * s, kSoftmaxWarps);
        }
        mbar_init(p_full, kSoftmaxWarps);
        mbar_init(pv_done, 1);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    if (tid == 0) ATTU_TRACE(3, 2, 0);

    if (warp == kLoadWarp) {
        // =============================================================== TMA loader: a [128 keys x 8 channels] box of the (B T, 3 H D)
        // tensor lands as one operand plane; rows past the clip belong to the next clip (or are zero-filled past the tensor):
        // finite values under probabilities that the causal mask has set to zero
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi) : "memory");
            if (SPLIT) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo) : "memory");
            const int c_k = koff + h * kD, c_v = voff + h * kD;
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % kStages;
                if (j >= kStages) mbar_wait(kv_empty + 8 * st, ((j / kStages) - 1) & 1);      // P V of tile j - kStages has read the stage
                ATTU_TRACE(2, 31, j);
                const int r0 = b * T + k_begin + j * kBK;
                const uint32_t kd = k_s + st * kParts * kKBytes, vd = v_s + st * L::kVStage, bar = kv_full + 8 * st;
                mbar_arrive_expect_tx(bar, kParts * 8 * kPlane);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    tma_load_2d(kd + g * kPlane, &tm_hi, c_k + 8 * g, r0, bar);
                    tma_load_2d(vd + g * kPlane, &tm_hi, c_v + 8 * g, r0, bar);
                    if (SPLIT) {
                        tma_load_2d(kd + kKBytes + g * kPlane, &tm_lo, c_k + 8 * g, r0, bar);
                        tma_load_2d(vd + kVBytes + g * kPlane, &tm_lo, c_v + 8 * g, r0, bar);
                    }
                }
                ATTU_TRACE(2, 30, j);
            }
        }
    } else if (warp == kMmaWarp) {
        // =============================================================== MMA issuer
        const bool leader = elect_one();
        const uint32_t idesc_s = make_idesc_bf16(128), idesc_o = make_idesc_bf16_bmn(48), idesc_o32 = make_idesc_bf16_bmn(32);
        const uint32_t o_tmem = tmem_base + kOCol;
        // Issue order: S(0); then per tile j:  S(j+1) as soon as the softmax warps have read S(j) out of TMEM, THEN P(j) V(j) --
        // the next tile's logits are computed while the probabilities of this one are still being written.
        auto issue_qk = [&](int j) {
            const int st = j % kStages;
            mbar_wait(kv_full + 8 * st, (j / kStages) & 1);
            if (j >= kSBufs) mbar_wait(s_free + 8 * sbuf(j), sphase(j - kSBufs));   // the softmax warps have read the tile that used this S buffer
            const uint32_t s_tmem = tmem_base + 128 * sbuf(j);
            tc_fence_after();
            const uint32_t kd = k_s + st * kParts * kKBytes;
            if (leader) {
                // S = Q K^T  (SPLIT: Qhi Khi + Qlo Khi + Qhi Klo)
#pragma unroll
                for (int term = 0; term < (SPLIT ? 3 : 1); ++term) {
                    const uint32_t qa = q_s + (term == 1 ? kQBytes : 0), kb = kd + (term == 2 ? kKBytes : 0);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        tc_mma_bf16(s_tmem, make_desc(qa + 2 * ks * kPlane, kPlane, 128), make_desc(kb + 2 * ks * kPlane, kPlane, 128),
                                    idesc_s, (term | ks) ? 1u : 0u);
                }
                tc_commit(s_full + 8 * sbuf(j));
                ATTU_TRACE(1, 20, j);
            }
            __syncwarp();
        };
        issue_qk(0);
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % kStages;
            if (j + 1 < n_tiles) issue_qk(j + 1);
            const uint32_t vd = v_s + st * L::kVStage;
            mbar_wait(p_full, j & 1);
            tc_fence_after();
            if (leader) {
                // O_j = P V'  (V' = [V | 1 | 0]: column 32 = row sums; SPLIT: Phi V'hi + Plo V'hi + Phi Vlo)
#pragma unroll
                for (int term = 0; term < (SPLIT ? 3 : 1); ++term) {
                    const uint32_t pa = p_s + (term == 1 ? kPBytes : 0), vb = vd + (term == 2 ? kVBytes : 0);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        tc_mma_bf16(o_tmem, make_desc(pa + 2 * kk * kPlane, kPlane, 128), make_desc(vb + kk * 16 * 16, 128, kPlane),
                                    term == 2 ? idesc_o32 : idesc_o, (term | kk) ? 1u : 0u);
                }
                tc_commit(pv_done);
                tc_commit(kv_empty + 8 * st);
                ATTU_TRACE(1, 21, j);
            }
            __syncwarp();
        }
    } else {
        // =============================================================== softmax: two threads per query row (warps 0-3: keys
        // 0..63 of every tile, warps 4-7: keys 64..127; both on the row's TMEM lane)
        const int quad = warp & 3, half = warp >> 2;
        const int row = quad * 32 + lane;
        const int qpos = q0 + row;
        int lo = (qpos / window - 1) * window;
        lo = lo < 0 ? 0 : lo;
        const int qmin_w = q0 + quad * 32, qmax_w = qmin_w + 31;
        int lo_max_w = (qmax_w / window - 1) * window;              // first visible key of the warp's last row (monotone in the row)
        lo_max_w = lo_max_w < 0 ? 0 : lo_max_w;
        const uint32_t tl = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int col0 = 64 * half;                                  // this thread's columns of S / keys of the tile
        const float scale = 0.17677669529663687f * 1.4426950408889634f;        // 32 ** -0.5, base-2 domain
        float l_run = 0.f;
        float o_acc[kD / 2];                                         // output columns 16 * half ..
#pragma unroll
        for (int i = 0; i < kD / 2; ++i) o_acc[i] = 0.f;

        // o = o * corr + O_j, l = l * corr + rowsum_j  (O_j of the tile whose P V has completed)
        auto fold = [&](float corr) {
            uint32_t ov[16], sv[8];
            tmem_ld16(tl + kOCol + 16 * half, ov);
            tmem_ld8(tl + kOCol + 32, sv);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < kD / 2; ++i) o_acc[i] = fmaf(o_acc[i], corr, __uint_as_float(ov[i]));
            l_run = fmaf(l_run, corr, __uint_as_float(sv[0]));
        };

        // Per tile: (1) an upper bound R_half of this thread's 64 biased logits (see below), exchanged with the partner thread of
        // the row (same scheduler: a 64-thread named barrier per lane quadrant); (2) the reference R = max(R_old, both halves):
        // softmax is shift-invariant, so any common reference >= the logits is exact and keeps p <= 1 -- no write-back of biased
        // logits to TMEM and no second bias lookup; (3) the main pass: s' = s * scale + bias (+ mask), p = 2^(s' - R) -> bf16.
        float r_acc = -INFINITY;                    // reference of the accumulators (o_acc, l_run)
        const uint32_t span = (uint32_t)(qpos - lo);                   // visible keys: lo <= key <= qpos
        for (int j = 0; j < n_tiles; ++j) {
            const int k0 = k_begin + j * kBK + col0;                 // first key of this thread's half tile
            const int rel = qpos - k0;
            if (tid == 0) ATTU_TRACE(0, 9, j);
            mbar_wait(s_full + 8 * sbuf(j), sphase(j));
            tc_fence_after();
            const uint32_t ts = tl + 128 * sbuf(j) + col0;           // this thread's 64 columns of S(j)
            if (tid == 0) ATTU_TRACE(0, 10, j);
            const bool full = (qmax_w < T) && (k0 + 63 <= qmin_w) && (k0 >= lo_max_w);      // warp-uniform
            // ---- (1) an upper bound of this half row's biased logits.  Full tile (all 64 keys visible): max(raw) * scale + the
            // maximum of the 64 bias values in play -- it exceeds the true maximum by at most the spread of the raw logits,
            // which is checked (> 60 in the base-2 domain: take the exact path).  Partial tile: the exact masked maximum.
            float mx = -INFINITY;
            bool exact = !full;
            if (full) {
                float mn = INFINITY;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld32(ts + 32 * c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        mx = fmaxf(mx, __uint_as_float(v[i]));
                        mn = fminf(mn, __uint_as_float(v[i]));
                    }
                }
                exact = __any_sync(0xffffffffu, !((mx - mn) * scale <= 60.f));        // (also catches NaN / inf logits)
                mx = fmaf(mx, scale, s_wmax[rel - 63]);
            }
            if (exact) {
                mx = -INFINITY;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld32(ts + 32 * c, v);
                    tmem_ld_wait();
                    const float* tp = s_table + (rel - 32 * c);
                    const int first = lo - k0 - 32 * c;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const bool ok = (uint32_t)(i - first) <= span;
                        mx = fmaxf(mx, ok ? fmaf(__uint_as_float(v[i]), scale, tp[-i]) : -INFINITY);
                    }
                }
            }
            float* xch = s_mx + (j & 1) * 256;
            xch[half * 128 + row] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
            mx = fmaxf(mx, xch[(half ^ 1) * 128 + row]);
            if (tid == 0) ATTU_TRACE(0, 11, j);
            // ---- (2) reference and the factor that moves the accumulators to it
            const float r_new = fmaxf(r_acc, mx);                    // (-inf while no key of the row has been visible yet)
            const float corr = (r_new == -INFINITY) ? 1.f : ex2f(r_acc - r_new);      // first visible tile: 2^(-inf) = 0 (accumulators are zero)
            // ---- the previous tile's P V has finished: add its result (same reference as the accumulators), free the P buffer
            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
                fold(1.f);
            }
            if (__any_sync(0xffffffffu, corr != 1.f)) {
#pragma unroll
                for (int i = 0; i < kD / 2; ++i) o_acc[i] *= corr;
                l_run *= corr;
            }
            r_acc = r_new;
            if (tid == 0) ATTU_TRACE(0, 13, j);
            // ---- (3) main pass: p = 2^(s * scale + bias - R) -> bf16 planes [key / 8][row][8]
            const float r_sub = (r_new == -INFINITY) ? 0.f : r_new;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld32(ts + 32 * c, v);
                tmem_ld_wait();
                if (c == 1) {                          // last read of S(j): a later tile's Q K^T may overwrite the buffer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free + 8 * sbuf(j));
                }
                const float* tp = s_table + (rel - 32 * c);
                const int first = lo - k0 - 32 * c;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float pf[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int e = 8 * g + i;
                        float sv = fmaf(__uint_as_float(v[e]), scale, tp[-e] - r_sub);
                        if (!full) sv = ((uint32_t)(e - first) <= span) ? sv : -INFINITY;
                        pf[i] = ex2f(sv);                                      // masked (-inf) -> 0
                    }
                    uint32_t ph[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) ph[i] = pack_bf16x2(pf[2 * i], pf[2 * i + 1]);
                    const uint32_t dst = p_s + (8 * half + 4 * c + g) * kPlane + row * 16;
                    st_shared_v4(dst, ph[0], ph[1], ph[2], ph[3]);
                    if (SPLIT) {
                        uint32_t pl[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            pl[i] = pack_bf16x2(pf[2 * i] - __uint_as_float(ph[i] << 16), pf[2 * i + 1] - __uint_as_float(ph[i] & 0xffff0000u));
                        st_shared_v4(dst + kPBytes, pl[0], pl[1], pl[2], pl[3]);
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            if (tid == 0) ATTU_TRACE(0, 14, j);
        }
        mbar_wait(pv_done, (n_tiles - 1) & 1);
        tc_fence_after();
        fold(1.f);

        if (qpos < T) {
            const float inv = 1.0f / l_run;
            const long long off = ((long long)b * T + qpos) * (H * kD) + h * kD + 16 * half;
            if (OUT == L3AC_F32) {
                float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + off);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    dst[i] = make_float4(o_acc[4 * i] * inv, o_acc[4 * i + 1] * inv, o_acc[4 * i + 2] * inv, o_acc[4 * i + 3] * inv);
            } else {
                uint32_t hi[8], lo16[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float v0 = o_acc[2 * i] * inv, v1 = o_acc[2 * i + 1] * inv;
                    hi[i] = pack_bf16x2(v0, v1);
                    lo16[i] = pack_bf16x2(v0 - __uint_as_float(hi[i] << 16), v1 - __uint_as_float(hi[i] & 0xffff0000u));
                }
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + off);
#pragma unroll
                for (int i = 0; i < 2; ++i) dst[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                if (OUT == L3AC_BF16X2) {
                    uint4* dl = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out_lo) + off);
#pragma unroll
                    for (int i = 0; i < 2; ++i) dl[i] = make_uint4(lo16[4 * i], lo16[4 * i + 1], lo16[4 * i + 2], lo16[4 * i + 3]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <bool SPLIT>
static int launch(const void* hi, const void* lo, const float* table, int B, int T, int H, int window, void* out,
                  void* out_lo, int out_dtype, cudaStream_t st) {
    const size_t smem = Smem<SPLIT>::bytes(window);
    if (smem > 227 * 1024) return L3AC_EUNSUPPORTED;
    dim3 grid((unsigned)B * (unsigned)H, l3ac_cdiv(T, kBQ));
    CUtensorMap tm_hi, tm_lo;
    const long long ld = 3LL * H * kD;
    if (!l3ac::tma::encode_planes_2d(&tm_hi, hi, ld, (long long)B * T, ld, kBK)) return L3AC_EDRIVER;
    if (!l3ac::tma::encode_planes_2d(&tm_lo, SPLIT ? lo : hi, ld, (long long)B * T, ld, kBK)) return L3AC_EDRIVER;
#define L3AC_ATTU_LAUNCH(OUTV)                                                                                          \
    do {                                                                                                                \
        cudaError_t e = cudaFuncSetAttribute(local_attention_umma_kernel<SPLIT, OUTV>,                                   \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                    \
        if (e != cudaSuccess) return (int)e;                                                                            \
        l3ac_launch(local_attention_umma_kernel<SPLIT, OUTV>, grid, dim3(kThreads), smem, st, tm_hi, tm_lo, (const __nv_bfloat16*)hi,   \
                    (const __nv_bfloat16*)lo, table, B, T, H, window, out, out_lo);                                      \
    } while (0)
    if (out_dtype == L3AC_F32) L3AC_ATTU_LAUNCH(L3AC_F32);
    else if (out_dtype == L3AC_BF16) L3AC_ATTU_LAUNCH(L3AC_BF16);
    else L3AC_ATTU_LAUNCH(L3AC_BF16X2);
#undef L3AC_ATTU_LAUNCH
    return l3ac_launch_status();
}

}  // namespace attu
}  // namespace l3ac

#ifdef L3AC_ATTU_TRACE
extern "C" int l3ac_debug_attu_trace(unsigned long long* host_buf) {      // host_buf: 4 * 512 entries
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_buf, l3ac::attu::g_attu_trace, 4 * 512 * sizeof(unsigned long long));
    return 0;
}
#endif

extern "C" int l3ac_local_attention_umma(const void* qkv_hi, const void* qkv_lo, const float* bias_table, int B, int T, int H,
                                         int D, int window, void* out, void* out_lo, int out_dtype, l3ac_stream_t stream) {
    using namespace l3ac::attu;
    L3AC_CHECK_ARG(qkv_hi && bias_table && out && B > 0 && B <= 65535 && T > 0 && H > 0 && H <= 65535 && window > 0);
    L3AC_CHECK_ARG(out_dtype == L3AC_F32 || out_dtype == L3AC_BF16 || out_dtype == L3AC_BF16X2);
    L3AC_CHECK_ARG((out_dtype == L3AC_BF16X2) == (out_lo != nullptr));
    L3AC_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(qkv_lo) & 15) == 0);
    L3AC_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_lo) & 15) == 0);
    if (D != kD || window > 4096) return L3AC_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    return qkv_lo ? launch<true>(qkv_hi, qkv_lo, bias_table, B, T, H, window, out, out_lo, out_dtype, st)
                  : launch<false>(qkv_hi, nullptr, bias_table, B, T, H, window, out, out_lo, out_dtype, st);
}
