// Fused ConvUnit MLP on tcgen05/TMEM (the point-wise half of l3ac/modules.py:36-44):
//
//   out[m, :] = x[m, :] + b2 + W2 . f( W1 . a[m, :] + b1 ),   f(v) = (1+gamma) * snake(v; alpha) + beta
//
// a = LayerNorm(dwconv7(x)) in bf16 (l3ac_dwconv7_ln), W1 (4C x C) and W2 (C x 4C) bf16, x / out fp32.  The 4C-wide
// hidden activation -- the largest tensor of the whole path -- never exists in HBM: per 128-row tile it is produced in
// 64-column chunks into TMEM, passed through the snake epilogue in registers, written as bf16 straight into the
// SWIZZLE_128B K-major shared-memory layout the tensor core reads, and consumed by the second GEMM whose C-wide
// accumulator stays in TMEM for the whole tile.
//
// One persistent CTA per SM, 20 warps.  The two GEMMs are issued by different warps from different weight rings, so the
// only coupling between them is the data flow D1 -> snake -> A2:
//   warp 0      TMA producer 1: the 128 x C activation tile and the W1 ring (one [128 hidden x 64 k] box per 16 KB slot)
//   warp 1      GEMM1 issuer: D1[b] = A . W1[super-chunk]^T, N = 128 (two 64-column chunks per MMA: every tcgen05
//               instruction costs the issuing thread ~100 cycles here, so fewer and wider is better), into one of TWO
//               128-column TMEM buffers, as soon as the buffer has been drained and the slots have landed
//   warp 2      TMA producer 2: the W2 ring (one [<=128 out x 64 k] box per 16 KB slot)
//   warp 3      GEMM2 issuer: D2 += A2[b] . W2[:, chunk]^T as soon as a hidden chunk has been written
//   warps 4-19  epilogue, four groups of four warps (one per TMEM lane quadrant).  Groups 2b and 2b+1 share D1[b] (one
//               64-column half each); group g owns A2[g].  All sixteen warps also drain D2 (+ b2 + residual -> fp32)
//               through 16-column staging slabs.
// C <= 128 ("pipelined"): D2 is double-buffered and the staging slabs have their own shared memory, so the epilogue warps
// run the snake chunks of tile t+1 BEFORE the output of tile t -- the residual loads and the stores of a tile are off the
// critical path and the tensor core never waits for them.  C > 128: one D2, the slabs alias the group's A2 buffer and the
// output of tile t comes before the chunks of tile t+1.
//
// All issuing warps run warp-uniform loops and issue through one elected lane: under a divergent `if (lane == 0)` the
// compiler cannot prove the shared-memory descriptors uniform and wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST
// waterfall (~200 cycles per MMA, measured with tools/mlp_trace.py).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace l3ac {
namespace mlp {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kTileBytes = kBM * kBK * 2;      // 16 KB: one [128 x 64] bf16 SW128 tile (A k-block, A2 buffer, ring slot)
constexpr int kNB = 4;                         // D1 / A2 buffers = epilogue groups
constexpr int kMaxRing = 8;                    // per ring
constexpr int kEpiWarp0 = 4;                   // first epilogue warp (a multiple of 4: warp & 3 is the TMEM lane quadrant)
constexpr int kEpiWarps = 4 * kNB;
constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);
constexpr int kSlabPitch = 20;                 // floats per staged row: 16 columns + 4 pad (conflict-free float4 access)
constexpr int kSlabBytes = 32 * kSlabPitch * 4;   // 2560 B per warp
constexpr int kD1Stride = 64;                  // TMEM columns per D1 buffer
constexpr int kD2Col = kNB * kD1Stride;        // D2 starts after the D1 buffers: 256 + C (or 2 x 128) <= 512
constexpr int kNumBars = 4 * kMaxRing + 4 + 4 * kNB + 4;
constexpr int kSmemLimit = 227 * 1024;

struct Params {
    const float* b1;
    const float* alpha;
    const float* ialpha;
    const float* scale;
    const float* shift;
    const float* b2;
    const float* residual;
    float* out;
    float* ch0;               // optional: channel 0 of the output as a compact (M) plane (what EnhanceBlock's statistics pass reads)
    long long M;
    int C, H4, HN, NC;        // channels, hidden = 4C, hidden chunk width (64), number of chunks
    int a_kb;                 // k-blocks of the activation tile = ceil(C / 64)
    int NS;                   // super-chunks (128 hidden columns = two chunks) per tile = ceil(NC / 2)
    int n_halves;             // GEMM2 N splits of <= 128 output columns (one W2 ring slot each)
    int ring1, ring2;         // ring depths
    int slot2_bytes;          // W2 ring slot: one [w2_rows x 64] box
    int w2_rows;              // rows of a W2 box = min(C, 128) rounded up to 16
    int a_bufs;               // activation tile buffers (2: the next tile's load overlaps this tile's GEMM1s)
    int resident;             // each ring holds a whole tile's slots: weights are loaded once per CTA and stay
    int pipelined;            // C <= 128: two D2 buffers, dedicated staging, output(t) after chunks(t+1)
    int num_m_tiles;
};

#ifdef L3AC_MLP_TRACE
// Debug build only: CTA 0 stamps clock64() at pipeline hand-overs of two steady-state tiles (tools/mlp_trace.py).  Every
// tracing thread owns a 512-entry region and a private index, so a stamp is one fire-and-forget store.
__device__ unsigned long long g_trace_buf[8 * 512];
#define MLP_TRACE_DECL(role) unsigned int trace_n = 0; const unsigned int trace_role = (role);
#define MLP_TRACE(ev, j)                                                                          \
    do {                                                                                          \
        if (blockIdx.x == 0 && it >= 2 && it <= 3 && trace_n < 511) {                             \
            g_trace_buf[trace_role * 512 + 1 + trace_n++] = ((unsigned long long)clock64() << 20) | ((unsigned long long)(it) << 16) | ((unsigned long long)(ev) << 8) | (unsigned long long)(j); \
            g_trace_buf[trace_role * 512] = trace_n;                                              \
        }                                                                                         \
    } while (0)
#else
#define MLP_TRACE_DECL(role)
#define MLP_TRACE(ev, j) do {} while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B smem matrix descriptor (see gemm_tc.cu)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1)
convunit_mlp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                    const __grid_constant__ CUtensorMap tmW2, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    // layout (1024-aligned tiles first): A [a_bufs][a_kb], A2 [kNB], W1 ring, W2 ring, staging (pipelined), parameters, barriers
    const uint32_t a_base = smem_base;
    const uint32_t a2_base = a_base + p.a_bufs * p.a_kb * kTileBytes;
    const uint32_t ring1_base = a2_base + kNB * kTileBytes;
    const uint32_t ring2_base = ring1_base + p.ring1 * kTileBytes;
    const uint32_t stage_base = ring2_base + p.ring2 * p.slot2_bytes;
    // per-column epilogue parameters [5][H4] (b1, alpha, 1/(alpha+eps), scale, shift) and b2 [C]: staged once -- with ~220 KB of
    // shared memory carved out there is next to no L1 left, and a global load per use is an exposed L2 round trip
    const uint32_t par_base = stage_base + (p.pipelined ? kEpiWarps * kSlabBytes : 0);
    float* s_par = reinterpret_cast<float*>(smem_gen + (par_base - smem_base));
    const uint32_t bar_base = par_base + (5 * p.H4 + p.C) * 4;
    const uint32_t r1_full = bar_base;                         // [kMaxRing]
    const uint32_t r1_empty = r1_full + 8 * kMaxRing;          // [kMaxRing]
    const uint32_t r2_full = r1_empty + 8 * kMaxRing;          // [kMaxRing]
    const uint32_t r2_empty = r2_full + 8 * kMaxRing;          // [kMaxRing]
    const uint32_t a_full = r2_empty + 8 * kMaxRing;           // [2]
    const uint32_t a_empty = a_full + 16;                      // [2]
    const uint32_t d1_full = a_empty + 16;                     // [kNB]
    const uint32_t d1_empty = d1_full + 8 * kNB;               // [kNB]
    const uint32_t a2_full = d1_empty + 8 * kNB;               // [kNB]
    const uint32_t a2_empty = a2_full + 8 * kNB;               // [kNB]
    const uint32_t d2_full = a2_empty + 8 * kNB;               // [2]
    const uint32_t d2_empty = d2_full + 16;                    // [2]
    const uint32_t tmem_slot = d2_empty + 16;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    // broadcast from lane 0: lets the compiler treat the role branches below as warp-uniform
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
        for (int s = 0; s < kMaxRing; ++s) {
            mbar_init(r1_full + 8 * s, 1);
            mbar_init(r1_empty + 8 * s, 1);
            mbar_init(r2_full + 8 * s, 1);
            mbar_init(r2_empty + 8 * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full + 8 * i, 1);
            mbar_init(a_empty + 8 * i, 1);
        }
        for (int i = 0; i < kNB; ++i) {
            mbar_init(d1_full + 8 * i, 1);        // (only [0..1] are used: two 128-column D1 buffers)
            mbar_init(d1_empty + 8 * i, 8);       // a D1 buffer is shared by a pair of groups (eight warps)
            mbar_init(a2_full + 8 * i, 4);        // an A2 buffer belongs to one group of four warps
            mbar_init(a2_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(d2_full + 8 * i, 1);
            mbar_init(d2_empty + 8 * i, kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.H4; i += kThreads) {
        s_par[i] = p.b1[i];
        s_par[p.H4 + i] = p.alpha[i];
        s_par[2 * p.H4 + i] = p.ialpha[i];
        s_par[3 * p.H4 + i] = p.scale[i];
        s_par[4 * p.H4 + i] = p.shift[i];
        if (i < p.C) s_par[5 * p.H4 + i] = p.b2[i];
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    L3AC_PDL_SYNC();      // the prologue above (barriers, parameters, TMEM) may overlap the previous kernel's tail

    const int n_my_tiles = (p.num_m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer 1: activation tile + W1 ring
        MLP_TRACE_DECL(0)
        const bool leader = elect_one();
        int rs = 0;
        uint32_t rphase = 0;
        for (int it = 0; it < n_my_tiles; ++it) {
            const int m_tile = blockIdx.x + it * gridDim.x;
            const int ab = it % p.a_bufs;
            mbar_wait(a_empty + 8 * ab, ((it / p.a_bufs) & 1) ^ 1);      // the GEMM1s that read this buffer have finished
            if (leader) MLP_TRACE(1, 0);
            if (leader) {
                mbar_arrive_expect_tx(a_full + 8 * ab, p.a_kb * kTileBytes);
                for (int kb = 0; kb < p.a_kb; ++kb)
                    tma_load_2d(a_base + (ab * p.a_kb + kb) * kTileBytes, &tmA, kb * kBK, m_tile * kBM, a_full + 8 * ab);
            }
            if (p.resident && it > 0) continue;               // resident weights are loaded once per CTA
            for (int sc = 0; sc < p.NS; ++sc)
                for (int kb = 0; kb < p.a_kb; ++kb) {         // one [128 hidden x 64 k] box per slot (rows past 4C are zero fill)
                    mbar_wait(r1_empty + 8 * rs, rphase ^ 1);
                    if (leader) {
                        mbar_arrive_expect_tx(r1_full + 8 * rs, kTileBytes);
                        tma_load_2d(ring1_base + rs * kTileBytes, &tmW1, kb * kBK, sc * 128, r1_full + 8 * rs);
                    }
                    if (++rs == p.ring1) {
                        rs = 0;
                        rphase ^= 1;
                    }
                }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ TMA producer 2: W2 ring
        const bool leader = elect_one();
        int rs = 0;
        uint32_t rphase = 0;
        const uint32_t w2_box = p.w2_rows * kBK * 2;
        for (int it = 0; it < (p.resident ? 1 : n_my_tiles); ++it)
            for (int j = 0; j < p.NC; ++j)
                for (int h = 0; h < p.n_halves; ++h) {
                    mbar_wait(r2_empty + 8 * rs, rphase ^ 1);
                    if (leader) {
                        mbar_arrive_expect_tx(r2_full + 8 * rs, w2_box);
                        tma_load_2d(ring2_base + rs * p.slot2_bytes, &tmW2, j * p.HN, h * 128, r2_full + 8 * rs);
                    }
                    if (++rs == p.ring2) {
                        rs = 0;
                        rphase ^= 1;
                    }
                }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ GEMM1 issuer: D1[sseq % 2] = A . W1[super-chunk]^T
        MLP_TRACE_DECL(1)
        const bool leader = elect_one();
        int rs = 0;
        uint32_t rphase = 0;
        uint32_t sseq = 0;                         // super-chunk sequence number of this CTA, across tiles
        const uint32_t idesc1 = make_idesc(128);
        // descriptors differ only in the 14-bit (address >> 4) field: keep the bases and add offsets in the issue loop
        const uint64_t a_desc0 = make_sw128_desc(a_base), w_desc0 = make_sw128_desc(ring1_base);
        const int k16_last = (p.C - (p.a_kb - 1) * kBK + 15) / 16;      // K16 steps of the last k-block
        for (int it = 0; it < n_my_tiles; ++it) {
            const bool ring_sync = !(p.resident && it > 0);    // resident mode: only the first tile waits for weight slots
            const int ab = it % p.a_bufs;
            mbar_wait(a_full + 8 * ab, (it / p.a_bufs) & 1);
            tc_fence_after();
            if (leader) MLP_TRACE(10, 0);
            const uint64_t a_tile = a_desc0 + (uint64_t)((uint32_t)(ab * p.a_kb) * (kTileBytes >> 4));
            for (int sc = 0; sc < p.NS; ++sc, ++sseq) {
                const int sb = sseq & 1;
                mbar_wait(d1_empty + 8 * sb, ((sseq >> 1) & 1) ^ 1);        // both groups of the pair have drained this D1 buffer
                tc_fence_after();
                if (leader) MLP_TRACE(15, sc);
                const uint32_t d1 = tmem_base + sb * 128;
                for (int kb = 0; kb < p.a_kb; ++kb) {
                    if (ring_sync) {
                        mbar_wait(r1_full + 8 * rs, rphase);
                        tc_fence_after();
                    }
                    if (leader) {
                        const uint64_t a_desc = a_tile + (uint64_t)((uint32_t)kb * (kTileBytes >> 4));
                        const uint64_t b_desc = w_desc0 + (uint64_t)((uint32_t)rs * (kTileBytes >> 4));
                        const int k16 = kb == p.a_kb - 1 ? k16_last : kBK / 16;
                        for (int k = 0; k < k16; ++k) tc_mma_f16(d1, a_desc + 2 * k, b_desc + 2 * k, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
                        if (!p.resident) tc_commit(r1_empty + 8 * rs);
                    }
                    if (++rs == p.ring1) {
                        rs = 0;
                        rphase ^= 1;
                    }
                }
                if (leader) {
                    tc_commit(d1_full + 8 * sb);
                    if (sc == p.NS - 1) tc_commit(a_empty + 8 * ab);     // last GEMM1 on this A buffer: free once it completes
                }
                if (leader) MLP_TRACE(16, sc);
            }
            __syncwarp();
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------ GEMM2 issuer: D2 += A2[group] . W2[:, chunk]^T
        MLP_TRACE_DECL(6)
        const bool leader = elect_one();
        int rs = 0;
        uint32_t rphase = 0;
        uint32_t sseq = 0, a2_par = 0;             // super-chunk sequence number; phase parity bit per A2 buffer
        const uint64_t a2_desc0 = make_sw128_desc(a2_base), w_desc0 = make_sw128_desc(ring2_base);
        const uint32_t slot2_q = (uint32_t)p.slot2_bytes >> 4;
        const uint32_t idesc_full = make_idesc(p.w2_rows);                                  // every half but possibly the last
        const uint32_t idesc_last = make_idesc((p.C - (p.n_halves - 1) * 128 + 15) & ~15);  // UMMA N is a multiple of 16; extra W2 rows are TMA zero fill
        for (int it = 0; it < n_my_tiles; ++it) {
            const bool ring_sync = !(p.resident && it > 0);
            const int d2b = p.pipelined ? (it & 1) : 0;
            const uint32_t d2_tmem = tmem_base + kD2Col + d2b * 128;
            for (int j = 0; j < p.NC; ++j) {
                // chunk j is half (j & 1) of super-chunk sseq: it was written by group 2 * (sseq & 1) + (j & 1) into its A2 buffer
                const int buf = 2 * (int)(sseq & 1) + (j & 1);
                mbar_wait(a2_full + 8 * buf, (a2_par >> buf) & 1);           // the group wrote the bf16 hidden chunk
                a2_par ^= 1u << buf;
                if (j == 0)                                                  // the output epilogue has drained this D2 buffer
                    mbar_wait(d2_empty + 8 * d2b, ((p.pipelined ? (it >> 1) : it) & 1) ^ 1);
                tc_fence_after();
                if (leader) MLP_TRACE(12, j);
                const uint64_t a_desc = a2_desc0 + (uint64_t)((uint32_t)buf * (kTileBytes >> 4));
                for (int h = 0; h < p.n_halves; ++h) {
                    if (ring_sync) {
                        mbar_wait(r2_full + 8 * rs, rphase);
                        tc_fence_after();
                    }
                    if (leader) {
                        const uint64_t b_desc = w_desc0 + (uint64_t)((uint32_t)rs * slot2_q);
                        const uint32_t idesc2 = h == p.n_halves - 1 ? idesc_last : idesc_full;
#pragma unroll
                        for (int k = 0; k < 4; ++k)      // HN = 64: four K16 steps
                            tc_mma_f16(d2_tmem + h * 128, a_desc + 2 * k, b_desc + 2 * k, idesc2, (j > 0 || k > 0) ? 1u : 0u);
                        if (!p.resident) tc_commit(r2_empty + 8 * rs);
                    }
                    if (++rs == p.ring2) {
                        rs = 0;
                        rphase ^= 1;
                    }
                }
                if (leader) tc_commit(a2_empty + 8 * buf);
                if (leader) MLP_TRACE(13, j);
                if (j & 1) ++sseq;
            }
            if (p.NC & 1) ++sseq;                 // a tile with an odd chunk count ends on a half-filled super-chunk
            if (leader) tc_commit(d2_full + 8 * d2b);
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------------------- epilogue warps
        const int quad = warp & 3;
        const int grp = (warp - kEpiWarp0) >> 2;                  // group index = D1 / A2 buffer index
        MLP_TRACE_DECL(2 + grp)
        const int row = quad * 32 + lane;                         // accumulator row of this thread
        float* stg = p.pipelined ? reinterpret_cast<float*>(smem_gen + (stage_base - smem_base) + (grp * 4 + quad) * kSlabBytes)
                                 : reinterpret_cast<float*>(smem_gen + (a2_base - smem_base) + grp * kTileBytes + quad * kSlabBytes);
        const int lane_r = lane >> 2, ci = lane & 3;              // coalesced output phase: 8 rows x 4 float4 (16 columns) per pass
        const float* stg_rd = stg + lane_r * kSlabPitch + 4 * ci;
        float* stg_wr = stg + lane * kSlabPitch;
        uint32_t my_use = 0, sb_use = 0;                          // chunks this group has processed; super-chunks seen on its D1 buffer
        const int n_passes = p.HN / 32;                           // 32-column passes per chunk
        const uint32_t a2_row = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);   // SW128 K-major tile, row = accumulator row
        const int n_out_tasks = p.C / 16;                         // 16-column output slices per tile (C % 16 == 0)

        // D2 of tile `ot` (+ b2 + residual) -> fp32.  32-column TMEM chunks, staged 16 columns at a time.
        auto output_tile = [&](int ot) {
            const int it = ot;                                    // (trace macro)
            const int d2b = p.pipelined ? (ot & 1) : 0;
            if (quad == 0 && lane == 0) MLP_TRACE(50 + grp, 0);
            mbar_wait(d2_full + 8 * d2b, (p.pipelined ? (ot >> 1) : ot) & 1);
            tc_fence_after();
            if (quad == 0 && lane == 0) MLP_TRACE(60 + grp, 0);
            const uint32_t d2_tmem = tmem_base + kD2Col + d2b * 128;
            const long long row_base = (long long)(blockIdx.x + ot * gridDim.x) * kBM;
            const int rows_valid = (int)((p.M - row_base) < kBM ? (p.M - row_base) : kBM);
            const int slab_rows = rows_valid - quad * 32;
            const long long row_lane = row_base + quad * 32 + lane_r;
            // task = one 16-column slice of the tile; the slices rotate over the four groups from tile to tile so that
            // narrow layers (C = 48: three slices) do not pin the whole output phase on the same groups
            for (int t = (grp - ot % kNB + kNB) % kNB; t < n_out_tasks; t += kNB) {
                const int nb = 16 * t;
                float4 res[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    res[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lane_r + 8 * q < slab_rows) res[q] = __ldg(reinterpret_cast<const float4*>(p.residual + (row_lane + 8 * q) * p.C + nb + 4 * ci));
                }
                uint32_t v[16];
                tmem_ld16(d2_tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)nb, v);
                __syncwarp();                          // the previous slice's staged rows have been read
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b4 = reinterpret_cast<const float4*>(s_par + 5 * p.H4 + nb)[i];
                    *reinterpret_cast<float4*>(stg_wr + 4 * i) =
                        make_float4(__uint_as_float(v[4 * i]) + b4.x, __uint_as_float(v[4 * i + 1]) + b4.y,
                                    __uint_as_float(v[4 * i + 2]) + b4.z, __uint_as_float(v[4 * i + 3]) + b4.w);
                }
                __syncwarp();
                const int col = nb + 4 * ci;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (!(lane_r + 8 * q < slab_rows)) continue;
                    float4 val = *reinterpret_cast<const float4*>(stg_rd + 8 * q * kSlabPitch);
                    val.x += res[q].x; val.y += res[q].y; val.z += res[q].z; val.w += res[q].w;
                    *reinterpret_cast<float4*>(p.out + (row_lane + 8 * q) * p.C + col) = val;
                    if (col == 0 && p.ch0) p.ch0[row_lane + 8 * q] = val.x;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_empty + 8 * d2b);
            if (quad == 0 && lane == 0) MLP_TRACE(70 + grp, 0);
        };

        for (int it = 0; it < n_my_tiles; ++it) {
            {   // pull this warp's share of the residual tile towards L2 now; the output epilogue reads it microseconds later
                const long long rb = (long long)(blockIdx.x + it * gridDim.x) * kBM + quad * 32 + lane_r;
                for (int t = (grp - it % kNB + kNB) % kNB; t < n_out_tasks; t += kNB)
                    if (ci == 0)
                        for (int q = 0; q < 4; ++q)
                            if (rb + 8 * q < p.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + (rb + 8 * q) * p.C + 16 * t));
            }
            // super-chunk sc of tile `it` is number it * NS + sc of this CTA's sequence and lives in D1[(it * NS + sc) & 1];
            // this group reads half (grp & 1) of the buffers with index grp >> 1
            for (int sc = ((grp >> 1) + it * p.NS) & 1; sc < p.NS; sc += 2) {
                const int j = 2 * sc + (grp & 1);
                if (quad == 0 && lane == 0) MLP_TRACE(20 + grp, j);
                mbar_wait(d1_full + 8 * (grp >> 1), sb_use & 1);
                ++sb_use;
                tc_fence_after();
                if (j >= p.NC) {                                         // half-filled super-chunk: nothing to read, hand the buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(d1_empty + 8 * (grp >> 1));
                    continue;
                }
                if (quad == 0 && lane == 0) MLP_TRACE(30 + grp, j);
                mbar_wait(a2_empty + 8 * grp, (my_use & 1) ^ 1);         // GEMM2 of this group's previous chunk has finished reading the A2 buffer
                ++my_use;
                for (int cc = 0; cc < n_passes; ++cc) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * kD1Stride + cc * 32), v);   // D1[grp >> 1], half grp & 1
                    if (quad == 0 && lane == 0) MLP_TRACE(80 + grp, cc);
                    if (cc + 1 == n_passes) {                            // last read of D1[grp]: the GEMM1 kNB chunks later may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(d1_empty + 8 * (grp >> 1));
                    }
                    const int n0 = j * p.HN + cc * 32;                   // first hidden column of this pass
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b4 = reinterpret_cast<const float4*>(s_par + n0)[i];
                        const float4 a4 = reinterpret_cast<const float4*>(s_par + p.H4 + n0)[i];
                        const float4 i4 = reinterpret_cast<const float4*>(s_par + 2 * p.H4 + n0)[i];
                        const float4 c4 = reinterpret_cast<const float4*>(s_par + 3 * p.H4 + n0)[i];
                        const float4 h4 = reinterpret_cast<const float4*>(s_par + 4 * p.H4 + n0)[i];
                        // packed fp32 (FADD2 / FMUL2 / FFMA2): the epilogue warps are issue-bound, and the packed ops round
                        // exactly like the scalar ones
                        const float2 x01 = fadd2(make_float2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), make_float2(b4.x, b4.y));
                        const float2 x23 = fadd2(make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), make_float2(b4.z, b4.w));
                        const float2 r01 = snake_affine2(x01, make_float2(a4.x, a4.y), make_float2(i4.x, i4.y), make_float2(c4.x, c4.y), make_float2(h4.x, h4.y));
                        const float2 r23 = snake_affine2(x23, make_float2(a4.z, a4.w), make_float2(i4.z, i4.w), make_float2(c4.z, c4.w), make_float2(h4.z, h4.w));
                        const float r[4] = {r01.x, r01.y, r23.x, r23.y};
                        const __nv_bfloat162 h01 = __floats2bfloat162_rn(r[0], r[1]), h23 = __floats2bfloat162_rn(r[2], r[3]);
                        pk[2 * i] = *reinterpret_cast<const uint32_t*>(&h01);
                        pk[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                    }
                    const uint32_t dst = a2_base + grp * kTileBytes + a2_row;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t chunk = (uint32_t)((4 * cc + q) ^ (row & 7));
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + chunk * 16), "r"(pk[4 * q]),
                                     "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                                     : "memory");
                    }
                }
                if (quad == 0 && lane == 0) MLP_TRACE(90 + grp, 0);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(a2_full + 8 * grp);
                if (quad == 0 && lane == 0) MLP_TRACE(40 + grp, j);
            }
            if (p.pipelined) {
                if (it > 0) output_tile(it - 1);      // tile it-1 is certainly through GEMM2 by now: no wait, off the critical path
            } else {
                output_tile(it);
                // The staging slabs alias this group's A2 buffer: no warp of the group may start writing the next tile's
                // hidden chunk into it before every warp of the group has finished reading its slab.
                asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
            }
        }
        if (p.pipelined && n_my_tiles > 0) output_tile(n_my_tiles - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

static bool encode_2d(EncodeTiledFn enc, CUtensorMap* tm, const void* ptr, long long inner, long long rows, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)inner * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


}  // namespace mlp
}  // namespace l3ac

#ifdef L3AC_MLP_TRACE
extern "C" int l3ac_debug_mlp_trace(unsigned long long* host_buf) {      // host_buf: 8 * 512 entries
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_buf, l3ac::mlp::g_trace_buf, 8 * 512 * sizeof(unsigned long long));
    static unsigned long long zeros[8 * 512];
    cudaMemcpyToSymbol(l3ac::mlp::g_trace_buf, zeros, sizeof(zeros));
    return 0;
}
#endif

extern "C" int l3ac_convunit_mlp_tc(const void* a, const void* w1, const float* b1, const float* alpha, const float* ialpha,
                                    const float* scale, const float* shift, const void* w2, const float* b2,
                                    const float* residual, float* out, long long M, int C, l3ac_stream_t stream) {
    return l3ac_convunit_mlp_tc_ch0(a, w1, b1, alpha, ialpha, scale, shift, w2, b2, residual, out, nullptr, M, C, stream);
}

extern "C" int l3ac_convunit_mlp_tc_ch0(const void* a, const void* w1, const float* b1, const float* alpha, const float* ialpha,
                                        const float* scale, const float* shift, const void* w2, const float* b2,
                                        const float* residual, float* out, float* ch0_out, long long M, int C, l3ac_stream_t stream) {
    using namespace l3ac::mlp;
    L3AC_CHECK_ARG(a && w1 && b1 && alpha && ialpha && scale && shift && w2 && b2 && residual && out && M > 0);
    // C = 512 does not fit shared memory / TMEM (two-GEMM path); the 16-column output staging wants whole 16-column groups
    if (C < 16 || C > 256 || C % 16 != 0) return L3AC_EUNSUPPORTED;
    const int H4 = 4 * C;
    const int HN = (H4 % 64 == 0) ? 64 : 32;
    if (H4 % HN != 0) return L3AC_EUNSUPPORTED;
    L3AC_CHECK_ARG(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(w2) |
                     reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(b1) |
                     reinterpret_cast<uintptr_t>(alpha) | reinterpret_cast<uintptr_t>(ialpha) | reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift) |
                     reinterpret_cast<uintptr_t>(b2)) & 15) == 0);
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return L3AC_EDRIVER;
    static int a_bufs_override = -1;
    if (a_bufs_override < 0) {
        const char* e = getenv("L3AC_MLP_A_BUFS");       // tuning knob: force 1 or 2 activation buffers
        a_bufs_override = e ? atoi(e) : 0;
    }
    Params p{};
    p.b1 = b1; p.alpha = alpha; p.ialpha = ialpha; p.scale = scale; p.shift = shift; p.b2 = b2; p.residual = residual; p.out = out;
    p.ch0 = ch0_out;
    p.M = M; p.C = C; p.H4 = H4; p.HN = HN; p.NC = H4 / HN;
    p.a_kb = (C + kBK - 1) / kBK;
    p.n_halves = (C + 127) / 128;
    p.NS = (p.NC + 1) / 2;
    if (HN != 64) return L3AC_EUNSUPPORTED;
    p.pipelined = C <= 128 ? 1 : 0;
    p.w2_rows = ((C < 128 ? C : 128) + 15) & ~15;
    p.slot2_bytes = p.w2_rows * kBK * 2;
    const int need1 = p.NS * p.a_kb, need2 = p.NC * p.n_halves;
    // Try two activation buffers first (the next tile's load then overlaps this tile's GEMM1s), fall back to one.
    bool placed = false;
    size_t smem_bytes = 0;
    for (p.a_bufs = (a_bufs_override > 0 ? a_bufs_override : 2); p.a_bufs >= 1 && !placed; --p.a_bufs) {
        const int fixed = 1024 + p.a_bufs * p.a_kb * kTileBytes + kNB * kTileBytes + (p.pipelined ? kEpiWarps * kSlabBytes : 0) +
                          (5 * H4 + C) * 4 + 8 * kNumBars + 64;
        const int left = kSmemLimit - fixed;
        if (need1 <= kMaxRing && need2 <= kMaxRing && need1 * kTileBytes + need2 * p.slot2_bytes <= left) {
            p.resident = 1;
            p.ring1 = need1;
            p.ring2 = need2;
        } else {
            // Split what is left between the rings in proportion to the bytes each streams per pair of chunks (W1: a_kb
            // slots of 16 KB, W2: 2 * n_halves slots).  Two A buffers are only worth it if both rings stay >= 3 deep.
            p.resident = 0;
            const double w1 = (double)p.a_kb * kTileBytes, w2 = 2.0 * p.n_halves * p.slot2_bytes;
            p.ring1 = (int)(left * (w1 / (w1 + w2))) / kTileBytes;
            if (p.ring1 < 2) p.ring1 = 2;
            if (p.ring1 > kMaxRing) p.ring1 = kMaxRing;
            p.ring2 = (left - p.ring1 * kTileBytes) / p.slot2_bytes;
            if (p.ring2 > kMaxRing) p.ring2 = kMaxRing;
            const int min_depth = p.a_bufs == 2 ? 3 : 2;
            if (left < 0 || p.ring1 < min_depth || p.ring2 < min_depth) continue;
            while (p.ring1 < kMaxRing && fixed + (p.ring1 + 1) * kTileBytes + p.ring2 * p.slot2_bytes <= kSmemLimit) ++p.ring1;
        }
        smem_bytes = (size_t)fixed + (size_t)p.ring1 * kTileBytes + (size_t)p.ring2 * p.slot2_bytes;
        placed = true;
        break;
    }
    if (!placed) return L3AC_EUNSUPPORTED;
    const long long mt = (M + kBM - 1) / kBM;
    L3AC_CHECK_ARG(mt < (1LL << 30));
    p.num_m_tiles = (int)mt;
    CUtensorMap tmA, tmW1, tmW2;
    if (!encode_2d(enc, &tmA, a, C, M, kBM) || !encode_2d(enc, &tmW1, w1, C, H4, 128) || !encode_2d(enc, &tmW2, w2, H4, C, p.w2_rows))
        return L3AC_EINVAL;
    cudaError_t e = cudaFuncSetAttribute(convunit_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return (int)e;
    const int sms = l3ac_sm_count();
    const int grid = (int)(mt < sms ? mt : sms);
    l3ac_launch(convunit_mlp_kernel, dim3(grid), dim3(kThreads), smem_bytes, (cudaStream_t)stream, tmA, tmW1, tmW2, p);
    return l3ac_launch_status();
}
