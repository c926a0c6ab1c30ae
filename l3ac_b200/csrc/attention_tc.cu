// Tensor-core block-local causal attention (flash-style, online softmax) over bf16 q/k/v planes.
// Semantics are those of attention.cu (LocalAttention of local-attention==1.11.2 as configured at
// l3ac/local_trans.py:34-38 + the DynamicPositionBias Toeplitz table, l3ac/local_trans.py:43):
//   query p attends keys j with max(0, (p/w - 1) w) <= j <= p;  logit = q.k / sqrt(32) + bias[h][p - j].
//
// One CTA = one (batch, head, 64-query tile), 4 warps x 16 queries.  K/V tiles of 64 keys are streamed with cp.async
// (double-buffered) into XOR-swizzled shared memory; S = Q K^T and O += P V run as warp-level mma.m16n8k16
// (bf16 in, fp32 accumulate); the softmax works directly on the accumulator fragments and P is re-used as the A
// operand of the second MMA without leaving registers.
// SPLIT: q/k/v arrive as (hi, lo) bf16 pairs and both products are issued as hi*hi + lo*hi + hi*lo, which gives the
// encode side fp32-class logits (token indices are sensitive to bf16 rounding).
#include "common.cuh"

namespace l3ac {
namespace att {

constexpr int kD = 32;
constexpr int kBQ = 64;
constexpr int kBK = 64;
constexpr int kTile = kBK * 64;   // bytes of one [64][32] bf16 tile

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {       // 2^x on MUFU.EX2 (2 ulp; ex2(-inf) = +0)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t v) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}

// Loads rows [row0, row0+64) (clamped to T-1) of one 32-wide head slice into a swizzled [64][32] bf16 tile.
__device__ __forceinline__ void load_tile_async(uint32_t dst, const __nv_bfloat16* base, long long ld, int row0, int T) {
    for (int i = threadIdx.x; i < 64 * 4; i += 128) {
        const int r = i >> 2, ch = i & 3;
        int t = row0 + r;
        t = t < T ? t : T - 1;
        cp_async16(dst + swz(r, ch), base + (long long)t * ld + ch * 8);
    }
}

template <bool SPLIT, int OUT>
__global__ void __launch_bounds__(128) local_attention_tc_kernel(const __nv_bfloat16* __restrict__ qkv_hi,
                                                                 const __nv_bfloat16* __restrict__ qkv_lo,
                                                                 const float* __restrict__ bias_table, int B, int T,
                                                                 int H, int window, void* __restrict__ out,
                                                                 void* __restrict__ out_lo) {
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: [stage 2][K hi, V hi, (K lo, V lo)] tiles, then Q hi (, Q lo), then the bias table (2w floats)
    constexpr int kPlanes = SPLIT ? 4 : 2;
    uint8_t* q_smem = smem + 2 * kPlanes * kTile;
    float* s_table = reinterpret_cast<float*>(q_smem + (SPLIT ? 2 : 1) * kTile);
    const uint32_t kv_addr = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t q_addr = (uint32_t)__cvta_generic_to_shared(q_smem);

    const int q0 = blockIdx.x * kBQ, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ld = 3LL * H * kD;
    const __nv_bfloat16* base_hi = qkv_hi + (long long)b * T * ld + h * kD;
    const __nv_bfloat16* base_lo = SPLIT ? qkv_lo + (long long)b * T * ld + h * kD : nullptr;
    const int koff = H * kD, voff = 2 * H * kD;

    for (int i = threadIdx.x; i < 2 * window; i += 128)       // base-2 domain: bias * log2 e
        s_table[i] = __ldg(bias_table + (long long)h * 2 * window + i) * 1.4426950408889634f;

    const int q_last = min(q0 + kBQ, T) - 1;
    int k_begin = (q0 / window - 1) * window;
    if (k_begin < 0) k_begin = 0;
    const int n_tiles = (q_last - k_begin) / kBK + 1;

    // prologue: Q tile + first K/V tile in flight
    load_tile_async(q_addr, base_hi, ld, q0, T);
    if (SPLIT) load_tile_async(q_addr + kTile, base_lo, ld, q0, T);
    load_tile_async(kv_addr, base_hi + koff, ld, k_begin, T);
    load_tile_async(kv_addr + kTile, base_hi + voff, ld, k_begin, T);
    if (SPLIT) {
        load_tile_async(kv_addr + 2 * kTile, base_lo + koff, ld, k_begin, T);
        load_tile_async(kv_addr + 3 * kTile, base_lo + voff, ld, k_begin, T);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // Q fragments (A operand): rows warp*16 .., two k-steps of 16
    uint32_t qa[2][4], ql[2][4];
    {
        const int lrow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int lchunk = lane >> 4;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            ldsm_x4(q_addr + swz(lrow, 2 * ks + lchunk), qa[ks]);
            if (SPLIT) ldsm_x4(q_addr + kTile + swz(lrow, 2 * ks + lchunk), ql[ks]);
        }
    }

    const int r_lo = lane >> 2;                      // accumulator rows r_lo and r_lo + 8 of the warp's 16 queries
    const int qpos0 = q0 + warp * 16 + r_lo, qpos1 = qpos0 + 8;
    int lo0 = (qpos0 / window - 1) * window, lo1 = (qpos1 / window - 1) * window;
    lo0 = lo0 < 0 ? 0 : lo0;
    lo1 = lo1 < 0 ? 0 : lo1;
    constexpr float kLog2e = 1.4426950408889634f;
    const float scale = 0.17677669529663687f * kLog2e;        // 32 ** -0.5, base-2 domain
    const int qmin_w = q0 + warp * 16, qmax_w = qmin_w + 15;    // this warp's query rows
    int lo_max_w = (qmax_w / window - 1) * window;              // first visible key of the last row (monotone in the row)
    lo_max_w = lo_max_w < 0 ? 0 : lo_max_w;

    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[n][e] = 0.f;

    for (int it = 0; it < n_tiles; ++it) {
        const int k0 = k_begin + it * kBK;
        const uint32_t st = kv_addr + (it & 1) * kPlanes * kTile;
        if (it + 1 < n_tiles) {       // prefetch the next K/V tile into the other stage
            const uint32_t nx = kv_addr + ((it + 1) & 1) * kPlanes * kTile;
            load_tile_async(nx, base_hi + koff, ld, k0 + kBK, T);
            load_tile_async(nx + kTile, base_hi + voff, ld, k0 + kBK, T);
            if (SPLIT) {
                load_tile_async(nx + 2 * kTile, base_lo + koff, ld, k0 + kBK, T);
                load_tile_async(nx + 3 * kTile, base_lo + voff, ld, k0 + kBK, T);
            }
        }
        cp_async_commit();

        // ---- S = Q K^T (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
            uint32_t kb[4], kl[4];
            const int krow = nt * 8 + (lane & 7), kch = lane >> 3;
            ldsm_x4(st + swz(krow, kch), kb);
            mma_bf16(s[nt], qa[0], kb[0], kb[1]);
            mma_bf16(s[nt], qa[1], kb[2], kb[3]);
            if (SPLIT) {
                ldsm_x4(st + 2 * kTile + swz(krow, kch), kl);
                mma_bf16(s[nt], ql[0], kb[0], kb[1]);
                mma_bf16(s[nt], ql[1], kb[2], kb[3]);
                mma_bf16(s[nt], qa[0], kl[0], kl[1]);
                mma_bf16(s[nt], qa[1], kl[2], kl[3]);
            }
        }
        // ---- scale + bias + mask, online softmax on the fragments.  Everything is kept in the base-2 domain (the scale and
        // the bias table carry a factor log2 e), so that a probability is one FADD + MUFU.EX2 instead of the ~10
        // instructions of expf.  Tiles that lie entirely inside every row's visible range (all but the diagonal and the
        // window-boundary tiles) skip the per-element position / mask logic; the choice is warp-uniform.
        float mx[2] = {-INFINITY, -INFINITY};
        const bool full = (qmax_w < T) && (k0 + kBK - 1 <= qmin_w) && (k0 >= lo_max_w);
        if (full) {
            const float* t0p = s_table + (qpos0 - k0 - (lane & 3) * 2);      // index qpos0 - kpos for e = 0; e = 1: one less
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float* tp = t0p - nt * 8;
                s[nt][0] = fmaf(s[nt][0], scale, tp[0]);
                s[nt][1] = fmaf(s[nt][1], scale, tp[-1]);
                s[nt][2] = fmaf(s[nt][2], scale, tp[8]);
                s[nt][3] = fmaf(s[nt][3], scale, tp[7]);
                mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
            }
        } else {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int kpos = k0 + nt * 8 + (lane & 3) * 2 + (e & 1);
                    const int qpos = (e >> 1) ? qpos1 : qpos0;
                    const int lo = (e >> 1) ? lo1 : lo0;
                    const bool ok = (qpos < T) && (kpos <= qpos) && (kpos >= lo);
                    const float v = ok ? fmaf(s[nt][e], scale, s_table[qpos - kpos]) : -INFINITY;
                    s[nt][e] = v;
                    mx[e >> 1] = fmaxf(mx[e >> 1], v);
                }
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            corr[r] = (m_new == -INFINITY) ? 1.f : ex2f(m_run[r] - m_new);
            m_run[r] = m_new;
            l_run[r] *= corr[r];
        }
        const float mref[2] = {m_run[0] == -INFINITY ? 0.f : m_run[0], m_run[1] == -INFINITY ? 0.f : m_run[1]};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float pv = ex2f(s[nt][e] - mref[e >> 1]);      // masked (-inf) -> 0
                s[nt][e] = pv;
                l_run[e >> 1] += pv;
            }
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            o[n][0] *= corr[0];
            o[n][1] *= corr[0];
            o[n][2] *= corr[1];
            o[n][3] *= corr[1];
        }
        // ---- O += P V : P fragments come straight from the S accumulators
#pragma unroll
        for (int j = 0; j < 4; ++j) {           // 16 keys per k-step
            uint32_t pa[4], pl[4];
            pa[0] = pack_bf16(s[2 * j][0], s[2 * j][1]);
            pa[1] = pack_bf16(s[2 * j][2], s[2 * j][3]);
            pa[2] = pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]);
            pa[3] = pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3]);
            if (SPLIT) {
                const float2 f0 = unpack_bf16(pa[0]), f1 = unpack_bf16(pa[1]), f2 = unpack_bf16(pa[2]), f3 = unpack_bf16(pa[3]);
                pl[0] = pack_bf16(s[2 * j][0] - f0.x, s[2 * j][1] - f0.y);
                pl[1] = pack_bf16(s[2 * j][2] - f1.x, s[2 * j][3] - f1.y);
                pl[2] = pack_bf16(s[2 * j + 1][0] - f2.x, s[2 * j + 1][1] - f2.y);
                pl[3] = pack_bf16(s[2 * j + 1][2] - f3.x, s[2 * j + 1][3] - f3.y);
            }
            const int vrow = 16 * j + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
            for (int nd = 0; nd < 4; nd += 2) {
                uint32_t vb[4], vl[4];
                ldsm_x4_trans(st + kTile + swz(vrow, nd + (lane >> 4)), vb);
                mma_bf16(o[nd], pa, vb[0], vb[1]);
                mma_bf16(o[nd + 1], pa, vb[2], vb[3]);
                if (SPLIT) {
                    ldsm_x4_trans(st + 3 * kTile + swz(vrow, nd + (lane >> 4)), vl);
                    mma_bf16(o[nd], pl, vb[0], vb[1]);
                    mma_bf16(o[nd + 1], pl, vb[2], vb[3]);
                    mma_bf16(o[nd], pa, vl[0], vl[1]);
                    mma_bf16(o[nd + 1], pa, vl[2], vl[3]);
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();      // next stage landed; everyone is done reading this one
    }

    // ---- finalise: divide by the row sums (partial per lane -> quad reduce) and write
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const int hd = H * kD;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qpos = r ? qpos1 : qpos0;
        if (qpos >= T) continue;
        const float inv = 1.0f / l_run[r];
        const long long off = ((long long)b * T + qpos) * hd + h * kD + (lane & 3) * 2;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const float v0 = o[n][2 * r] * inv, v1 = o[n][2 * r + 1] * inv;
            if (OUT == L3AC_F32) {
                *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + off + n * 8) = make_float2(v0, v1);
            } else {
                const uint32_t hi = pack_bf16(v0, v1);
                *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + off + n * 8) = hi;
                if (OUT == L3AC_BF16X2) {
                    const float2 f = unpack_bf16(hi);
                    *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out_lo) + off + n * 8) = pack_bf16(v0 - f.x, v1 - f.y);
                }
            }
        }
    }
}

template <bool SPLIT>
static int launch(const void* hi, const void* lo, const float* table, int B, int T, int H, int window, void* out,
                  void* out_lo, int out_dtype, cudaStream_t st) {
    const size_t smem = (size_t)(2 * (SPLIT ? 4 : 2) + (SPLIT ? 2 : 1)) * kTile + (size_t)2 * window * sizeof(float);
    dim3 grid(l3ac_cdiv(T, kBQ), H, B);
#define L3AC_ATT_LAUNCH(OUTV)                                                                                          \
    do {                                                                                                               \
        cudaError_t e = cudaFuncSetAttribute(local_attention_tc_kernel<SPLIT, OUTV>,                                    \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                   \
        if (e != cudaSuccess) return (int)e;                                                                           \
        local_attention_tc_kernel<SPLIT, OUTV><<<grid, 128, smem, st>>>((const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo, \
                                                                        table, B, T, H, window, out, out_lo);           \
    } while (0)
    if (out_dtype == L3AC_F32) L3AC_ATT_LAUNCH(L3AC_F32);
    else if (out_dtype == L3AC_BF16) L3AC_ATT_LAUNCH(L3AC_BF16);
    else L3AC_ATT_LAUNCH(L3AC_BF16X2);
#undef L3AC_ATT_LAUNCH
    return l3ac_launch_status();
}

}  // namespace att
}  // namespace l3ac

extern "C" int l3ac_local_attention_tc(const void* qkv_hi, const void* qkv_lo, const float* bias_table, int B, int T, int H,
                                       int D, int window, void* out, void* out_lo, int out_dtype, l3ac_stream_t stream) {
    using namespace l3ac::att;
    L3AC_CHECK_ARG(qkv_hi && bias_table && out && B > 0 && B <= 65535 && T > 0 && H > 0 && H <= 65535 && window > 0);
    L3AC_CHECK_ARG(out_dtype == L3AC_F32 || out_dtype == L3AC_BF16 || out_dtype == L3AC_BF16X2);
    L3AC_CHECK_ARG((out_dtype == L3AC_BF16X2) == (out_lo != nullptr));
    L3AC_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(qkv_lo) & 15) == 0);
    if (D != kD || window > 4096) return L3AC_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    return qkv_lo ? launch<true>(qkv_hi, qkv_lo, bias_table, B, T, H, window, out, out_lo, out_dtype, st)
                  : launch<false>(qkv_hi, nullptr, bias_table, B, T, H, window, out, out_lo, out_dtype, st);
}
