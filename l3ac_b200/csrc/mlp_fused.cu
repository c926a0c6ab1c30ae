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
// One persistent CTA per SM, 10 warps:
//   warp 0     TMA producer: the 128 x C activation tile (resident for the tile) and a ring of 16 KB weight boxes
//              (W1: 64 hidden rows x 64 k, W2: <=128 output rows x 64 k) in exactly the order the MMA warp consumes them
//   warp 1     MMA issuer (one thread): GEMM1(j+1) is issued before GEMM2(j), so the snake epilogue of chunk j overlaps
//              tensor-core work; D1 is double-buffered (2 x 64 TMEM columns), D2 owns C columns
//   warps 2-9  epilogue: tcgen05.ld D1 -> bias/snake/affine -> bf16 -> swizzled smem A2 (fence.proxy.async) ; at the end
//              of the tile D2 -> + b2 + residual -> coalesced fp32 stores
#include <cuda.h>

#include "common.cuh"

namespace l3ac {
namespace mlp {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kTileBytes = kBM * kBK * 2;      // 16 KB: one [128 x 64] bf16 SW128 tile (A k-block, A2 buffer, ring stage)
constexpr int kMaxAKb = 4;                     // C <= 256
constexpr int kMaxRing = 12;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kStagePitch = 36;
constexpr int kStageBytes = kEpiWarps * 32 * kStagePitch * 4;   // output staging, aliased onto the A2 buffers
constexpr int kD1Cols = 256;                   // two D1 buffers of up to 128 columns
constexpr int kNumBars = 2 * kMaxRing + 2 + 4 + 4 + 2;
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
    long long M;
    int C, H4, HN, NC;        // channels, hidden = 4C, hidden chunk width, number of chunks
    int a_kb;                 // k-blocks of the activation tile = ceil(C / 64)
    int n_halves;             // GEMM2 N splits of <= 128 output columns
    int hn_kb;                // k-blocks of one hidden chunk = ceil(HN / 64)
    int ring;                 // weight ring depth (16 KB slots)
    int resident;             // ring == boxes per tile: every weight box is loaded once per CTA and stays in shared memory
    int num_m_tiles;
};

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

__host__ __device__ inline int a2_stride_bytes(int hn_kb) {
    const int need = hn_kb * kTileBytes, slabs = 20 * 1024;      // 4 slabs of 32 x 36 floats = 18 432 B, rounded to 1 KB
    return need > slabs ? need : slabs;
}

__global__ void __launch_bounds__(kThreads, 1)
convunit_mlp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                    const __grid_constant__ CUtensorMap tmW2, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    // layout (all 1024-aligned): A tile [a_kb], A2 [2][hn_kb] (the output staging slabs alias it), ring [ring], barriers
    const uint32_t a_base = smem_base;
    const uint32_t a2_base = a_base + p.a_kb * kTileBytes;
    // each epilogue group owns one A2 buffer; its four output-staging slabs (4 x 4608 B) alias that same buffer, so a
    // group never touches shared memory the other group may be writing for the next tile
    const uint32_t a2_buf_bytes = (uint32_t)a2_stride_bytes(p.hn_kb);
    const uint32_t ring_base = a2_base + 2 * a2_buf_bytes;
    float* s_stage = reinterpret_cast<float*>(smem_gen + (a2_base - smem_base));
    const uint32_t bar_base = ring_base + p.ring * kTileBytes;
    const uint32_t ring_full = bar_base;                       // [kMaxRing]
    const uint32_t ring_empty = ring_full + 8 * kMaxRing;      // [kMaxRing]
    const uint32_t a_full = ring_empty + 8 * kMaxRing;         // [1]
    const uint32_t a_empty = a_full + 8;                       // [1]
    const uint32_t d1_full = a_empty + 8;                      // [2]
    const uint32_t d1_empty = d1_full + 16;                    // [2]
    const uint32_t a2_full = d1_empty + 16;                    // [2]
    const uint32_t a2_empty = a2_full + 16;                    // [2]
    const uint32_t d2_full = a2_empty + 16;                    // [1]
    const uint32_t d2_empty = d2_full + 8;                     // [1]
    const uint32_t tmem_slot = d2_empty + 8;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
        for (int s = 0; s < p.ring; ++s) {
            mbar_init(ring_full + 8 * s, 1);
            mbar_init(ring_empty + 8 * s, 1);
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(d1_full + 8 * i, 1);
            mbar_init(d1_empty + 8 * i, kEpiWarps / 2);      // each D1 / A2 buffer belongs to one group of four warps
            mbar_init(a2_full + 8 * i, kEpiWarps / 2);
            mbar_init(a2_empty + 8 * i, 1);
        }
        mbar_init(d2_full, 1);
        mbar_init(d2_empty, kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    const uint32_t d2_tmem = tmem_base + kD1Cols;

    const int n_my_tiles = (p.num_m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer
            int rs = 0;
            uint32_t rphase = 0;
            bool first_tile = true;
            auto ring_load = [&](const CUtensorMap* map, int c0, int c1, uint32_t bytes) {
                if (p.resident && !first_tile) return;            // weights already resident in their slots
                mbar_wait(ring_empty + 8 * rs, rphase ^ 1);
                mbar_arrive_expect_tx(ring_full + 8 * rs, bytes);
                tma_load_2d(ring_base + rs * kTileBytes, map, c0, c1, ring_full + 8 * rs);
                if (++rs == p.ring) {
                    rs = 0;
                    rphase ^= 1;
                }
            };
            const uint32_t w1_bytes = p.HN * kBK * 2, w2_bytes = kBM * kBK * 2;     // HN <= 128 rows per W1 box
            auto gemm1_boxes = [&](int j) {
                for (int kb = 0; kb < p.a_kb; ++kb) ring_load(&tmW1, kb * kBK, j * p.HN, w1_bytes);
            };
            for (int it = 0; it < n_my_tiles; ++it) {
                const int m_tile = blockIdx.x + it * gridDim.x;
                mbar_wait(a_empty, (it & 1) ^ 1);                // previous tile's GEMM1s have finished reading A
                mbar_arrive_expect_tx(a_full, p.a_kb * kTileBytes);
                for (int kb = 0; kb < p.a_kb; ++kb) tma_load_2d(a_base + kb * kTileBytes, &tmA, kb * kBK, m_tile * kBM, a_full);
                gemm1_boxes(0);
                for (int j = 0; j < p.NC; ++j) {
                    if (j + 1 < p.NC) gemm1_boxes(j + 1);
                    for (int kb = 0; kb < p.hn_kb; ++kb)
                        for (int h = 0; h < p.n_halves; ++h) ring_load(&tmW2, j * p.HN + kb * kBK, h * 128, w2_bytes);
                }
                first_tile = false;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------------ MMA issuer
            int rs = 0;
            uint32_t rphase = 0;
            uint32_t d1_use[2] = {0, 0}, a2_use[2] = {0, 0};      // how many times each buffer has been handed over
            bool ring_sync = true;                                // resident mode: only the first tile waits for the weight boxes
            const uint32_t idesc1 = make_idesc(p.HN);
            auto gemm1 = [&](int j) {
                const int buf = j & 1;
                mbar_wait(d1_empty + 8 * buf, (d1_use[buf] & 1) ^ 1);     // epilogue has drained this D1 buffer
                tc_fence_after();
                const uint32_t d1 = tmem_base + buf * 128;
                for (int kb = 0; kb < p.a_kb; ++kb) {
                    if (ring_sync) {
                        mbar_wait(ring_full + 8 * rs, rphase);
                        tc_fence_after();
                    }
                    const uint64_t a_desc = make_sw128_desc(a_base + kb * kTileBytes);
                    const uint64_t b_desc = make_sw128_desc(ring_base + rs * kTileBytes);
                    const int k_left = p.C - kb * kBK;
                    const int k16 = k_left >= kBK ? kBK / 16 : (k_left + 15) / 16;
                    for (int k = 0; k < k16; ++k) tc_mma_f16(d1, a_desc + 2 * k, b_desc + 2 * k, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
                    if (!p.resident) tc_commit(ring_empty + 8 * rs);
                    if (++rs == p.ring) {
                        rs = 0;
                        rphase ^= 1;
                    }
                }
                tc_commit(d1_full + 8 * buf);
                ++d1_use[buf];
            };
            for (int it = 0; it < n_my_tiles; ++it) {
                mbar_wait(a_full, it & 1);
                tc_fence_after();
                gemm1(0);
                for (int j = 0; j < p.NC; ++j) {
                    if (j + 1 < p.NC) gemm1(j + 1);
                    if (j + 1 == p.NC) tc_commit(a_empty);        // all GEMM1s of this tile issued: A is free once they complete
                    const int buf = j & 1;
                    mbar_wait(a2_full + 8 * buf, a2_use[buf] & 1);       // epilogue wrote the bf16 hidden chunk
                    tc_fence_after();
                    if (j == 0) {
                        mbar_wait(d2_empty, (it & 1) ^ 1);               // previous tile's output epilogue has drained D2
                        tc_fence_after();
                    }
                    for (int kb = 0; kb < p.hn_kb; ++kb) {
                        const uint64_t a_desc = make_sw128_desc(a2_base + buf * a2_buf_bytes + kb * kTileBytes);
                        const int k_left = p.HN - kb * kBK;
                        const int k16 = k_left >= kBK ? kBK / 16 : k_left / 16;
                        for (int h = 0; h < p.n_halves; ++h) {
                            if (ring_sync) {
                                mbar_wait(ring_full + 8 * rs, rphase);
                                tc_fence_after();
                            }
                            const uint64_t b_desc = make_sw128_desc(ring_base + rs * kTileBytes);
                            const int n = min(128, (p.C - h * 128 + 15) & ~15);    // UMMA N is a multiple of 16; extra W2 rows are TMA zero fill
                            const uint32_t idesc2 = make_idesc(n);
                            for (int k = 0; k < k16; ++k)
                                tc_mma_f16(d2_tmem + h * 128, a_desc + 2 * k, b_desc + 2 * k, idesc2, (j > 0 || kb > 0 || k > 0) ? 1u : 0u);
                            if (!p.resident) tc_commit(ring_empty + 8 * rs);
                            if (++rs == p.ring) {
                                rs = 0;
                                rphase ^= 1;
                            }
                        }
                    }
                    tc_commit(a2_empty + 8 * buf);
                    ++a2_use[buf];
                }
                tc_commit(d2_full);
                if (p.resident) ring_sync = false;
            }
        }
    } else {
        // ---------------------------------------------------------------------- epilogue warps
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;                         // accumulator row of this thread
        float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_stage) + half * a2_buf_bytes) + quad * 32 * kStagePitch;
        const int lane_r = lane >> 3, ci = lane & 7;              // coalesced output phase: 4 rows x 8 float4 per pass
        const float* stg_rd = stg + lane_r * kStagePitch + 4 * ci;
        float* stg_wr = stg + lane * kStagePitch;
        uint32_t my_use = 0;                                      // chunks this group has processed (its buffer index == half)
        const int n_passes = p.HN / 32;                           // 32-column passes per chunk
        // A2 write address: SW128 K-major tiles of 64 columns; row = accumulator row
        const uint32_t a2_row = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        for (int it = 0; it < n_my_tiles; ++it) {
            const int m_tile = blockIdx.x + it * gridDim.x;
            for (int j = half; j < p.NC; j += 2) {                // group `half` owns D1[half] / A2[half] = chunks j == half (mod 2)
                const int buf = half;
                mbar_wait(d1_full + 8 * buf, my_use & 1);
                tc_fence_after();
                mbar_wait(a2_empty + 8 * buf, (my_use & 1) ^ 1);         // GEMM2(j-2) has finished reading this A2 buffer
                ++my_use;
                for (int cc = 0; cc < n_passes; ++cc) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 128 + cc * 32), v);
                    if (cc + 1 == n_passes) {                            // last read of D1[buf]: GEMM1(j+2) may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(d1_empty + 8 * buf);
                    }
                    const int n0 = j * p.HN + cc * 32;                   // first hidden column of this pass
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b1 + n0) + i);
                        const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.alpha + n0) + i);
                        const float4 i4 = __ldg(reinterpret_cast<const float4*>(p.ialpha + n0) + i);
                        const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.scale + n0) + i);
                        const float4 h4 = __ldg(reinterpret_cast<const float4*>(p.shift + n0) + i);
                        float r[4];
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, aa[4] = {a4.x, a4.y, a4.z, a4.w};
                        const float ii[4] = {i4.x, i4.y, i4.z, i4.w};
                        const float cc4[4] = {c4.x, c4.y, c4.z, c4.w}, hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float x = __uint_as_float(v[4 * i + e]) + bb[e];
                            const float sn = __sinf(aa[e] * x);
                            x = fmaf(ii[e], sn * sn, x);
                            r[e] = fmaf(x, cc4[e], hh[e]);
                        }
                        const __nv_bfloat162 h01 = __floats2bfloat162_rn(r[0], r[1]), h23 = __floats2bfloat162_rn(r[2], r[3]);
                        pk[2 * i] = *reinterpret_cast<const uint32_t*>(&h01);
                        pk[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                    }
                    const uint32_t dst = a2_base + buf * a2_buf_bytes + (cc >> 1) * kTileBytes + a2_row;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t chunk = (uint32_t)((4 * (cc & 1) + q) ^ (row & 7));
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + chunk * 16), "r"(pk[4 * q]),
                                     "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                                     : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(a2_full + 8 * buf);
            }
            // ---- output epilogue: D2 (+ b2 + residual) -> fp32, 32-column chunks, coalesced through the staging slab
            mbar_wait(d2_full, it & 1);
            tc_fence_after();
            const long long row_base = (long long)m_tile * kBM;
            const int rows_valid = (int)((p.M - row_base) < kBM ? (p.M - row_base) : kBM);
            const int slab_rows = rows_valid - quad * 32;
            const long long row_lane = row_base + quad * 32 + lane_r;
            const int n_chunks = (p.C + 31) / 32;
            for (int c = half; c < n_chunks; c += 2) {
                uint32_t v[32];
                tmem_ld32(d2_tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), v);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int n = c * 32 + 4 * i;
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (n < p.C) b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + n));
                    *reinterpret_cast<float4*>(stg_wr + 4 * i) =
                        make_float4(__uint_as_float(v[4 * i]) + b4.x, __uint_as_float(v[4 * i + 1]) + b4.y,
                                    __uint_as_float(v[4 * i + 2]) + b4.z, __uint_as_float(v[4 * i + 3]) + b4.w);
                }
                __syncwarp();
                const int col = c * 32 + 4 * ci;
                const bool col_ok = col < p.C;
                float4 res[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    res[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (col_ok && lane_r + 4 * q < slab_rows)
                        res[q] = __ldg(reinterpret_cast<const float4*>(p.residual + (row_lane + 4 * q) * p.C + col));
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (!(col_ok && lane_r + 4 * q < slab_rows)) continue;
                    float4 val = *reinterpret_cast<const float4*>(stg_rd + 4 * q * kStagePitch);
                    val.x += res[q].x; val.y += res[q].y; val.z += res[q].z; val.w += res[q].w;
                    *reinterpret_cast<float4*>(p.out + (row_lane + 4 * q) * p.C + col) = val;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_empty);
            // The staging slabs alias this group's A2 buffer: no warp of the group may start writing the next tile's
            // hidden chunk into it before every warp of the group has finished reading its slab.
            asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
        }
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

extern "C" int l3ac_convunit_mlp_tc(const void* a, const void* w1, const float* b1, const float* alpha, const float* ialpha,
                                    const float* scale, const float* shift, const void* w2, const float* b2,
                                    const float* residual, float* out, long long M, int C, l3ac_stream_t stream) {
    using namespace l3ac::mlp;
    L3AC_CHECK_ARG(a && w1 && b1 && alpha && ialpha && scale && shift && w2 && b2 && residual && out && M > 0);
    if (C < 16 || C > 256 || C % 8 != 0) return L3AC_EUNSUPPORTED;       // C = 512 does not fit shared memory: use the two-GEMM path
    const int H4 = 4 * C;
    const int HN = (H4 % 128 == 0) ? 128 : (H4 % 64 == 0) ? 64 : 32;
    if (H4 % HN != 0) return L3AC_EUNSUPPORTED;
    L3AC_CHECK_ARG(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(w2) |
                     reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(b1) |
                     reinterpret_cast<uintptr_t>(alpha) | reinterpret_cast<uintptr_t>(ialpha) | reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift) |
                     reinterpret_cast<uintptr_t>(b2)) & 15) == 0);
    L3AC_CHECK_ARG(C % 4 == 0);
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return L3AC_EDRIVER;
    Params p{};
    p.b1 = b1; p.alpha = alpha; p.ialpha = ialpha; p.scale = scale; p.shift = shift; p.b2 = b2; p.residual = residual; p.out = out;
    p.M = M; p.C = C; p.H4 = H4; p.HN = HN; p.NC = H4 / HN;
    p.a_kb = (C + kBK - 1) / kBK;
    p.n_halves = (C + 127) / 128;
    p.hn_kb = (HN + kBK - 1) / kBK;
    const int a2_region = 2 * a2_stride_bytes(p.hn_kb);
    const int fixed = 1024 + p.a_kb * kTileBytes + a2_region + 8 * kNumBars + 64;
    p.ring = (kSmemLimit - fixed) / kTileBytes;
    if (p.ring > kMaxRing) p.ring = kMaxRing;
    if (p.ring < 2) return L3AC_EUNSUPPORTED;
    const int boxes_per_tile = p.NC * (p.a_kb + p.hn_kb * p.n_halves);
    p.resident = boxes_per_tile <= p.ring ? 1 : 0;
    if (p.resident) p.ring = boxes_per_tile;
    const size_t smem_bytes = (size_t)fixed + (size_t)p.ring * kTileBytes;
    const long long mt = (M + kBM - 1) / kBM;
    L3AC_CHECK_ARG(mt < (1LL << 30));
    p.num_m_tiles = (int)mt;
    CUtensorMap tmA, tmW1, tmW2;
    if (!encode_2d(enc, &tmA, a, C, M, kBM) || !encode_2d(enc, &tmW1, w1, C, H4, HN) || !encode_2d(enc, &tmW2, w2, H4, C, 128))
        return L3AC_EINVAL;
    cudaError_t e = cudaFuncSetAttribute(convunit_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return (int)e;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = (int)(mt < sms ? mt : sms);
    convunit_mlp_kernel<<<grid, kThreads, smem_bytes, (cudaStream_t)stream>>>(tmA, tmW1, tmW2, p);
    return l3ac_launch_status();
}
