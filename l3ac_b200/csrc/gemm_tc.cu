// bf16 x bf16 -> fp32 GEMM / conv-as-GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
//   out[m, n] = epi( bias[n] + sum_{s<taps} sum_{k<K} A[b, t + shift_s, k] * W[n, s*K + k] ),  m = b*T + t
//
// Persistent, warp-specialised, one CTA per SM (10 warps):
//   warp 0      TMA producer   : A tile (128 rows x 64 k) via a 3-D map (k, t, b) -- out-of-range rows/columns are
//                                zero-filled by the TMA unit, which implements the conv zero padding and all K/M/N
//                                tails -- and W tile (BN rows x 64 k) via a 2-D map; SWIZZLE_128B, mbarrier tx-count.
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//                                accumulators in TMEM, double-buffered (2 x BN columns); tcgen05.commit frees smem
//                                stages and publishes finished accumulators.
//   warps 2..9  epilogue       : tcgen05.ld (32 lanes x 32 columns per instruction) -> bias / snake+affine / GEGLU /
//                                residual -> global stores; overlaps the next tile's main loop.
// Everything a conv layer needs beyond a GEMM (k-tap shifts with dilation, per-sample zero padding, strided
// "patchify" convs as a reshape) is expressed through the TMA coordinates, never by reshaping data in HBM.
#include <cuda.h>

#include "common.cuh"

namespace l3ac {
namespace tc {

constexpr int kBM = 128;
constexpr int kBK = 64;                       // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kATileBytes = kBM * kBK * 2;    // 16 KB
constexpr int kMaxBN = 256;
constexpr int kMaxStages = 8;
constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kSmemBudget = 200 * 1024;

struct Params {
    const float* bias;
    const float* alpha;
    const float* scale;
    const float* shift;
    const float* residual;
    void* out;
    long long ldr, ldo;
    int B, T, K, N, taps, tap_shift0, tap_step;
    int BN, stages, flat;
    int tiles_per_b, num_m_tiles, num_n_tiles, k_blocks;
    int act, out_dtype;
};

// ------------------------------------------------------------------------------------------ PTX wrappers
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
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

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in
// bits [0,14), SBO (8 rows x 128 B = 1024 B) >> 4 in bits [32,46), version 1 in bits [46,48), layout type
// SWIZZLE_128B (= 2) in bits [61,64).  LBO is unused for swizzled K-major operands.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major.
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

__device__ __forceinline__ void tile_coords(const Params& p, int tile, int& m_tile, int& n_tile) {
    m_tile = tile / p.num_n_tiles;
    n_tile = tile - m_tile * p.num_n_tiles;
}

__device__ __forceinline__ float epi_fast(float v, int act, float bias, float alpha, float inv_alpha, float scale,
                                          float shift) {
    v += bias;
    if (act == L3AC_ACT_SNAKE) {
        const float s = __sinf(alpha * v);
        v = fmaf(inv_alpha, s * s, v);
        v = fmaf(v, scale, shift);
    } else if (act == L3AC_ACT_GELU) {
        v = gelu_erf(v);
    } else if (act == L3AC_ACT_TANH) {
        v = tanhf(v);
    }
    return v;
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int w_tile_bytes = p.BN * kBK * 2;
    const uint32_t a_base = smem_base;
    const uint32_t w_base = a_base + p.stages * kATileBytes;
    const uint32_t tail = w_base + p.stages * w_tile_bytes;     // 1024-aligned (both tile sizes are multiples of 1 KB)
    // tail layout: params [5][256] floats, then barriers
    float* s_par = reinterpret_cast<float*>(smem_gen + (tail - smem_base));
    const uint32_t bar_base = tail + 5 * kMaxBN * 4;
    const uint32_t full_bar = bar_base;                          // [kMaxStages]
    const uint32_t empty_bar = bar_base + 8 * kMaxStages;        // [kMaxStages]
    const uint32_t tfull_bar = bar_base + 16 * kMaxStages;       // [2]
    const uint32_t tempty_bar = tfull_bar + 16;                  // [2]
    const uint32_t tmem_slot = tempty_bar + 16;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tmem_cols = (2 * p.BN <= 32) ? 32 : (2 * p.BN <= 64) ? 64 : (2 * p.BN <= 128) ? 128 : (2 * p.BN <= 256) ? 256 : 512;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar + 8 * a, 1);
            mbar_init(tempty_bar + 8 * a, kEpiThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int total_kb = p.taps * p.k_blocks;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = kATileBytes + w_tile_bytes;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_tile, n_tile;
                tile_coords(p, tile, m_tile, n_tile);
                int b = 0, row0 = m_tile * kBM;
                if (!p.flat) {
                    b = m_tile / p.tiles_per_b;
                    row0 = (m_tile - b * p.tiles_per_b) * kBM;
                }
                const int n0 = n_tile * p.BN;
                for (int s = 0; s < p.taps; ++s) {
                    const int shift = p.tap_shift0 + s * p.tap_step;
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                        mbar_arrive_expect_tx(full_bar + 8 * stage, tx_bytes);
                        tma_load_3d(a_base + stage * kATileBytes, &tmA, kb * kBK, row0 + shift, b, full_bar + 8 * stage);
                        tma_load_2d(w_base + stage * w_tile_bytes, &tmW, s * p.K + kb * kBK, n0, full_bar + 8 * stage);
                        if (++stage == p.stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------------ MMA issuer
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t idesc = make_idesc(p.BN);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.BN;
                int it = 0;
                for (int s = 0; s < p.taps; ++s) {
                    for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
                        mbar_wait(full_bar + 8 * stage, phase);
                        tc_fence_after();
                        const uint64_t a_desc = make_sw128_desc(a_base + stage * kATileBytes);
                        const uint64_t b_desc = make_sw128_desc(w_base + stage * w_tile_bytes);
                        const int k_left = p.K - kb * kBK;
                        const int k16 = k_left >= kBK ? kBK / 16 : (k_left + 15) / 16;
                        for (int k = 0; k < k16; ++k) {
                            // advance 16 bf16 = 32 B inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                            tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
                        }
                        tc_commit(empty_bar + 8 * stage);
                        if (++stage == p.stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
                (void)total_kb;
                tc_commit(tfull_bar + 8 * acc);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ---------------------------------------------------------------------- epilogue (8 warps)
        const int ep_tid = threadIdx.x - 64;
        const int quad = warp & 3;            // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;     // which half of the column chunks
        const int n_chunks = p.BN / 32;
        float* s_bias = s_par;
        float* s_alpha = s_par + kMaxBN;
        float* s_ialpha = s_par + 2 * kMaxBN;
        float* s_scale = s_par + 3 * kMaxBN;
        float* s_shift = s_par + 4 * kMaxBN;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int m_tile, n_tile;
            tile_coords(p, tile, m_tile, n_tile);
            const int n0 = n_tile * p.BN;
            asm volatile("bar.sync 1, 256;" ::: "memory");     // previous tile's parameter reads are done
            for (int i = ep_tid; i < p.BN; i += kEpiThreads) {
                const int n = n0 + i;
                const bool ok = n < p.N;
                s_bias[i] = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
                const float a = (ok && p.alpha) ? __ldg(p.alpha + n) : 1.f;
                s_alpha[i] = a;
                s_ialpha[i] = 1.0f / (a + kEps);
                s_scale[i] = (ok && p.scale) ? __ldg(p.scale + n) : 1.f;
                s_shift[i] = (ok && p.shift) ? __ldg(p.shift + n) : 0.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");

            const int row = quad * 32 + lane;
            long long m;
            bool row_ok;
            if (p.flat) {
                m = (long long)m_tile * kBM + row;
                row_ok = m < (long long)p.B * p.T;
            } else {
                const int b = m_tile / p.tiles_per_b;
                const int t = (m_tile - b * p.tiles_per_b) * kBM + row;
                row_ok = t < p.T;
                m = (long long)b * p.T + t;
            }

            mbar_wait(tfull_bar + 8 * acc, acc_phase);
            tc_fence_after();
            for (int c = half; c < n_chunks; c += 2) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.BN + c * 32), v);
                if (!row_ok) continue;
                const int cb = c * 32;              // column offset inside the tile
                const int nb = n0 + cb;             // global column
                if (nb >= p.N) continue;
                if (p.act == L3AC_ACT_GEGLU) {
                    float r[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float val = __uint_as_float(v[2 * i]) + s_bias[cb + 2 * i];
                        const float gate = __uint_as_float(v[2 * i + 1]) + s_bias[cb + 2 * i + 1];
                        r[i] = val * gelu_erf(gate);
                    }
                    const int no = nb >> 1;
                    const int n_out = p.N >> 1;
                    if (p.residual) {
                        const float* rr = p.residual + m * p.ldr + no;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (no + i < n_out) r[i] += rr[i];
                    }
                    if (p.out_dtype == L3AC_F32) {
                        float* o = reinterpret_cast<float*>(p.out) + m * p.ldo + no;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (no + i < n_out) o[i] = r[i];
                    } else {
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + no;
                        if (no + 16 <= n_out && ((p.ldo | no) & 7) == 0) {
                            uint4 pk[2];
                            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
                            for (int i = 0; i < 8; ++i) h2[i] = __floats2bfloat162_rn(r[2 * i], r[2 * i + 1]);
                            reinterpret_cast<uint4*>(o)[0] = pk[0];
                            reinterpret_cast<uint4*>(o)[1] = pk[1];
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (no + i < n_out) o[i] = __float2bfloat16_rn(r[i]);
                        }
                    }
                    continue;
                }
                float r[32];
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    r[i] = epi_fast(__uint_as_float(v[i]), p.act, s_bias[cb + i], s_alpha[cb + i], s_ialpha[cb + i],
                                    s_scale[cb + i], s_shift[cb + i]);
                const bool full = nb + 32 <= p.N;
                if (p.residual) {
                    const float* rr = p.residual + m * p.ldr + nb;
                    if (full && ((p.ldr | nb) & 3) == 0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 q4 = reinterpret_cast<const float4*>(rr)[i];
                            r[4 * i] += q4.x;
                            r[4 * i + 1] += q4.y;
                            r[4 * i + 2] += q4.z;
                            r[4 * i + 3] += q4.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (nb + i < p.N) r[i] += rr[i];
                    }
                }
                if (p.out_dtype == L3AC_F32) {
                    float* o = reinterpret_cast<float*>(p.out) + m * p.ldo + nb;
                    if (full && ((p.ldo | nb) & 3) == 0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            reinterpret_cast<float4*>(o)[i] = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (nb + i < p.N) o[i] = r[i];
                    }
                } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + nb;
                    if (full && ((p.ldo | nb) & 7) == 0) {
                        uint4 pk[4];
                        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
                        for (int i = 0; i < 16; ++i) h2[i] = __floats2bfloat162_rn(r[2 * i], r[2 * i + 1]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(o)[i] = pk[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (nb + i < p.N) o[i] = __float2bfloat16_rn(r[i]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
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

static int pick_bn(int N) {
    if (N <= kMaxBN) return ((N + 31) / 32) * 32;
    int best = 256;
    long long best_cost = -1;
    for (int bn = 256; bn >= 128; bn -= 32) {
        const long long cost = (long long)((N + bn - 1) / bn) * bn;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

}  // namespace tc
}  // namespace l3ac

using namespace l3ac::tc;

extern "C" int l3ac_gemm_bf16_tc(const l3ac_gemm_desc* d, l3ac_stream_t stream) {
    const int rc = l3ac_validate_gemm_desc(d);
    if (rc != L3AC_OK) return rc;
    const long long ktot = (long long)d->taps * d->K;
    L3AC_CHECK_ARG(d->lda % 8 == 0 && ktot % 8 == 0);
    L3AC_CHECK_ARG((reinterpret_cast<uintptr_t>(d->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->W) & 15) == 0);
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return L3AC_EDRIVER;

    Params p{};
    p.bias = d->bias; p.alpha = d->alpha; p.scale = d->scale; p.shift = d->shift; p.residual = d->residual;
    p.out = d->out; p.ldr = d->ldr; p.ldo = d->ldo;
    p.B = d->B; p.T = d->T; p.K = d->K; p.N = d->N;
    p.taps = d->taps; p.tap_shift0 = d->tap_shift0; p.tap_step = d->tap_step;
    p.act = d->act; p.out_dtype = d->out_dtype;
    p.BN = pick_bn(d->N);
    const int stage_bytes = kATileBytes + p.BN * kBK * 2;
    p.stages = kSmemBudget / stage_bytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    p.flat = (d->taps == 1 && d->tap_shift0 == 0) ? 1 : 0;
    p.k_blocks = (d->K + kBK - 1) / kBK;
    const long long M = (long long)d->B * d->T;
    if (p.flat) {
        p.tiles_per_b = 0;
        const long long mt = (M + kBM - 1) / kBM;
        L3AC_CHECK_ARG(mt < (1LL << 30));
        p.num_m_tiles = (int)mt;
    } else {
        p.tiles_per_b = (d->T + kBM - 1) / kBM;
        const long long mt = (long long)p.tiles_per_b * d->B;
        L3AC_CHECK_ARG(mt < (1LL << 30));
        p.num_m_tiles = (int)mt;
    }
    p.num_n_tiles = (d->N + p.BN - 1) / p.BN;
    L3AC_CHECK_ARG((long long)p.num_m_tiles * p.num_n_tiles < (1LL << 31));

    // A: (k, t, b) bf16, row pitch lda.  Flat GEMMs view all B*T rows as one sample so tiles never straddle padding.
    CUtensorMap tmA, tmW;
    {
        const cuuint64_t rows = p.flat ? (cuuint64_t)M : (cuuint64_t)d->T;
        const cuuint64_t batches = p.flat ? 1 : (cuuint64_t)d->B;
        cuuint64_t dims[3] = {(cuuint64_t)d->K, rows, batches};
        cuuint64_t strides[2] = {(cuuint64_t)d->lda * 2, (cuuint64_t)d->lda * 2 * rows};
        cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)kBM, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->A), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return L3AC_EINVAL;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)d->N};
        cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)p.BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->W), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return L3AC_EINVAL;
    }

    const size_t smem = 1024 /* alignment slack */ + (size_t)p.stages * stage_bytes + 5 * kMaxBN * 4 + 16 * kMaxStages + 64;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (long long)p.num_m_tiles * p.num_n_tiles;
    const int grid = (int)(tiles < sms ? tiles : sms);
    gemm_tc_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmA, tmW, p);
    return l3ac_launch_status();
}
