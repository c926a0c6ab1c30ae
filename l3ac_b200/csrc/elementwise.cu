// HBM-bound stencil / normalisation kernels of the L3AC conv stack (channels-last activations).
//   dwconv7 + LayerNorm      ConvUnit prologue          l3ac/modules.py:33-35
//   LayerNorm                ChannelNorm / nn.LayerNorm l3ac/layers.py:50-56
//   snake                    Snake1d                    l3ac/layers.py:29-47
//   linear upsample (+CN)    nn.Upsample + ChannelNorm  l3ac/modules.py:160-164
//   EnhanceBlock             stats + apply              l3ac/tconv/__init__.py:30-44
//   tail                     Snake -> Conv(C->1,k7) -> tanh   l3ac/modules.py:192-194
#include "common.cuh"

#include <new>

namespace l3ac {

// Writes v as OutT; for the split pair (OutT = bf16 and lo != nullptr) also the low-order plane bf16(v - hi).
template <typename OutT>
__device__ __forceinline__ void store_act(OutT* hi, OutT* lo, long long i, float v) {
    hi[i] = cvt_out<OutT>(v);
}
template <>
__device__ __forceinline__ void store_act<__nv_bfloat16>(__nv_bfloat16* hi, __nv_bfloat16* lo, long long i, float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// Eight simultaneous sums over a group of GS lanes (GS = 8, 16 or 32 consecutive lanes of a warp).  Three halving steps
// (a lane keeps half of its values and trades the other half with its partner) leave every lane with ONE row's partial
// sum, log2(GS) - 3 butterfly steps finish it: 7 + (log2(GS) - 3) shuffles instead of 8 log2(GS).  After the call lane l
// of a group holds the group total of row (l * 8 / GS) & 7 ... see group_row8; group_bcast8 hands every total to every lane.
template <int GS>
__device__ __forceinline__ float group_reduce8(const float (&v)[8]) {
    const int gl = (threadIdx.x & 31) & (GS - 1);
    const bool b2 = gl & (GS / 2), b1 = gl & (GS / 4), b0 = gl & (GS / 8);
    float w[4], u[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = (b2 ? v[k + 4] : v[k]) + __shfl_xor_sync(0xffffffffu, b2 ? v[k] : v[k + 4], GS / 2);
#pragma unroll
    for (int k = 0; k < 2; ++k) u[k] = (b1 ? w[k + 2] : w[k]) + __shfl_xor_sync(0xffffffffu, b1 ? w[k] : w[k + 2], GS / 4);
    float t = (b0 ? u[1] : u[0]) + __shfl_xor_sync(0xffffffffu, b0 ? u[0] : u[1], GS / 8);
#pragma unroll
    for (int o = GS / 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;           // total of row 4 b2 + 2 b1 + b0
}
// all eight totals in every lane of the group: row r sits in the group's lane r * (GS / 8)
template <int GS>
__device__ __forceinline__ void group_bcast8(float t, float (&v)[8]) {
    const int base = (threadIdx.x & 31) & ~(GS - 1);
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = __shfl_sync(0xffffffffu, t, base + r * (GS / 8));
}

template <int GS>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = GS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Four consecutive values (16-byte aligned destination for fp32, 8-byte for bf16).
template <typename OutT>
__device__ __forceinline__ void store_act4(OutT* hi, OutT* lo, long long i, float4 v) {
    *reinterpret_cast<float4*>(hi + i) = v;
}
template <>
__device__ __forceinline__ void store_act4<__nv_bfloat16>(__nv_bfloat16* hi, __nv_bfloat16* lo, long long i, float4 v) {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&h01);
    pk.y = *reinterpret_cast<const uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(hi + i) = pk;
    if (lo) {
        const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
        pk.x = *reinterpret_cast<const uint32_t*>(&l01);
        pk.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(lo + i) = pk;
    }
}

// ------------------------------------------------------------------------------------------
// depthwise conv k7 (pad 3) + LayerNorm over C.  One warp per (b, t) row; lane owns channels
// lane, lane+32, ...  Weights are staged in shared memory ([7][C] taps, bias, ln w/b).
// ------------------------------------------------------------------------------------------
template <int CPL, typename OutT>
__global__ void __launch_bounds__(256) dwconv7_ln_kernel(const float* __restrict__ x, int B, int T, int C,
                                                         const float* __restrict__ dw_w,
                                                         const float* __restrict__ dw_b,
                                                         const float* __restrict__ ln_w,
                                                         const float* __restrict__ ln_b, float eps,
                                                         OutT* __restrict__ out, OutT* __restrict__ out_lo) {
    extern __shared__ float sm[];
    float* s_w = sm;             // [7][C]
    float* s_b = sm + 7 * C;     // [C]
    float* s_lw = s_b + C;       // [C]
    float* s_lb = s_lw + C;      // [C]
    for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) s_w[i] = dw_w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        s_b[i] = dw_b[i];
        s_lw[i] = ln_w[i];
        s_lb[i] = ln_b[i];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const long long rows = (long long)B * T;
    for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows;
         row += (long long)gridDim.x * warps_per_block) {
        const int t = (int)(row % T);
        const float* xr = x + row * C;
        float acc[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
            acc[i] = (c < C) ? s_b[c] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int tt = t + j - 3;
            if (tt < 0 || tt >= T) continue;   // warp-uniform
            const float* xs = xr + (long long)(j - 3) * C;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = lane + 32 * i;
                if (c < C) acc[i] = fmaf(s_w[j * C + c], __ldg(xs + c), acc[i]);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) s += (lane + 32 * i < C) ? acc[i] : 0.f;
        const float mean = warp_sum(s) / (float)C;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const float dlt = acc[i] - mean;
            v += (lane + 32 * i < C) ? dlt * dlt : 0.f;
        }
        const float rstd = rsqrt_nr(warp_sum(v) / (float)C + eps);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) store_act<OutT>(out, out_lo, row * C + c, (acc[i] - mean) * rstd * s_lw[c] + s_lb[c]);
        }
    }
}

// Thin-channel variant (C <= 96): one warp handles R consecutive time steps of one sample per iteration and issues
// all (R + 6) * CPL input loads up front.  With 96..384 B rows the one-row-per-warp kernel above is latency-bound
// (one DRAM round trip per row and warp); this keeps ~8x more bytes in flight and reads every input row once per warp
// instead of seven times.
template <int CPL, int R, typename OutT>
__global__ void __launch_bounds__(256) dwconv7_ln_rows_kernel(const float* __restrict__ x, int B, int T, int C,
                                                              const float* __restrict__ dw_w,
                                                              const float* __restrict__ dw_b,
                                                              const float* __restrict__ ln_w,
                                                              const float* __restrict__ ln_b, float eps,
                                                              OutT* __restrict__ out, OutT* __restrict__ out_lo) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    float w[7][CPL], bias[CPL], lw[CPL], lb[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const int c = lane + 32 * i;
        const bool ok = c < C;
#pragma unroll
        for (int j = 0; j < 7; ++j) w[j][i] = ok ? __ldg(dw_w + j * C + c) : 0.f;
        bias[i] = ok ? __ldg(dw_b + c) : 0.f;
        lw[i] = ok ? __ldg(ln_w + c) : 0.f;
        lb[i] = ok ? __ldg(ln_b + c) : 0.f;
    }
    const int runs_per_sample = (T + R - 1) / R;
    const long long total_runs = (long long)B * runs_per_sample;
    const float inv_c = 1.0f / (float)C;
    for (long long run = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); run < total_runs;
         run += (long long)gridDim.x * warps_per_block) {
        const int b = (int)(run / runs_per_sample);
        const int t0 = (int)(run - (long long)b * runs_per_sample) * R;
        const float* xb = x + (long long)b * T * C;
        float xr[R + 6][CPL];
#pragma unroll
        for (int r = 0; r < R + 6; ++r) {
            const int t = t0 + r - 3;
            const bool row_ok = t >= 0 && t < T;      // warp-uniform
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = lane + 32 * i;
                xr[r][i] = (row_ok && c < C) ? __ldg(xb + (long long)t * C + c) : 0.f;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (t0 + r >= T) break;                   // warp-uniform
            float acc[CPL];
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                float a = bias[i];
#pragma unroll
                for (int j = 0; j < 7; ++j) a = fmaf(w[j][i], xr[r + j][i], a);
                acc[i] = a;
                s += (lane + 32 * i < C) ? a : 0.f;
            }
            const float mean = warp_sum(s) * inv_c;
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const float dlt = acc[i] - mean;
                v += (lane + 32 * i < C) ? dlt * dlt : 0.f;
            }
            const float rstd = rsqrt_nr(warp_sum(v) * inv_c + eps);
            const long long row = (long long)b * T + t0 + r;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = lane + 32 * i;
                if (c < C) store_act<OutT>(out, out_lo, row * C + c, (acc[i] - mean) * rstd * lw[i] + lb[i]);
            }
        }
    }
}

// Thin-channel vector variant (C % 4 == 0, C <= 128): a row is owned by a group of GS lanes (one float4 of channels
// per lane), a warp handles 32/GS independent runs of R consecutive time steps, every lane issues its R+6 128-bit
// loads up front, and the LayerNorm statistics are log2(GS)-step shuffle reductions.  ~14 instructions per output
// element instead of ~42 for the scalar one-channel-per-lane kernel (which is instruction-bound at C = 24).
template <int GS, int R, typename OutT>
__global__ void __launch_bounds__(256) dwconv7_ln_vec_kernel(const float* __restrict__ x, int B, int T, int C,
                                                             const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                             const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                             float eps, OutT* __restrict__ out, OutT* __restrict__ out_lo) {
    constexpr int RW = 32 / GS;
    const int lane = threadIdx.x & 31, g = lane % GS, sub = lane / GS;
    const int C4 = C >> 2;
    const bool act = g < C4;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 w[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) w[j] = act ? __ldg(reinterpret_cast<const float4*>(dw_w + j * C) + g) : z4;
    const float4 bias = act ? __ldg(reinterpret_cast<const float4*>(dw_b) + g) : z4;
    const float4 lw = act ? __ldg(reinterpret_cast<const float4*>(ln_w) + g) : z4;
    const float4 lb = act ? __ldg(reinterpret_cast<const float4*>(ln_b) + g) : z4;
    const int runs_per_sample = (T + R - 1) / R;
    const long long total_runs = (long long)B * runs_per_sample;
    const float inv_c = 1.0f / (float)C;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long base = warp_global * RW; base < total_runs; base += nwarps * RW) {
        const long long run = base + sub;
        const bool run_ok = run < total_runs;
        const long long rr = run_ok ? run : total_runs - 1;
        int b;
        if (total_runs <= 0x7fffffffLL) b = (int)((unsigned)rr / (unsigned)runs_per_sample);     // the 64-bit division costs ~60 instructions
        else b = (int)(rr / runs_per_sample);
        const int t0 = (int)(rr - (long long)b * runs_per_sample) * R;
        const float4* xb = reinterpret_cast<const float4*>(x + (long long)b * T * C) + g;
        float4 xr[R + 6];
        // Interior runs (no clip edge inside the R + 6 rows, true for all but two runs per clip) take a warp-uniform fast
        // path: one base pointer, 32-bit row offsets, no per-row range predicates or zero-fill selects.
        const bool interior = __all_sync(0xffffffffu, run_ok && t0 >= 3 && t0 + R + 3 <= T);
        if (interior) {
            const float4* p0 = xb + (long long)(t0 - 3) * C4;
#pragma unroll
            for (int r = 0; r < R + 6; ++r) xr[r] = act ? __ldg(p0 + r * C4) : z4;
        } else {
#pragma unroll
            for (int r = 0; r < R + 6; ++r) {
                const int t = t0 + r - 3;
                xr[r] = (act && t >= 0 && t < T) ? __ldg(xb + (long long)t * C4) : z4;
            }
        }
        const long long out_row0 = ((long long)b * T + t0) * C + 4 * g;
        static_assert(R == 8, "the statistics use the 8-row group reduction");
        float4 a[R];
        float st[8];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float4 acc = bias;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const float2 lo = ffma2(make_float2(w[j].x, w[j].y), make_float2(xr[r + j].x, xr[r + j].y), make_float2(acc.x, acc.y));
                const float2 hi = ffma2(make_float2(w[j].z, w[j].w), make_float2(xr[r + j].z, xr[r + j].w), make_float2(acc.z, acc.w));
                acc = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
            a[r] = acc;
            st[r] = (acc.x + acc.y) + (acc.z + acc.w);                     // inactive lanes hold zeros
        }
        group_bcast8<GS>(group_reduce8<GS>(st), st);                       // row sums: 8 rows in one reduction
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float mean = st[r] * inv_c;
            a[r].x -= mean; a[r].y -= mean; a[r].z -= mean; a[r].w -= mean;
            st[r] = act ? (a[r].x * a[r].x + a[r].y * a[r].y) + (a[r].z * a[r].z + a[r].w * a[r].w) : 0.f;
        }
        group_bcast8<GS>(group_reduce8<GS>(st), st);                       // centred sums of squares
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float rstd = rsqrt_nr(st[r] * inv_c + eps);
            if (act && (interior || (run_ok && t0 + r < T))) {
                const long long i = out_row0 + r * C;
                store_act4<OutT>(out, out_lo, i,
                                 make_float4(a[r].x * rstd * lw.x + lb.x, a[r].y * rstd * lw.y + lb.y, a[r].z * rstd * lw.z + lb.z,
                                             a[r].w * rstd * lw.w + lb.w));
            }
        }
    }
}

// Thread-per-row variant for the decode side's thin stages (C = 48 / 96, bf16 out; 2.6 M / 0.85 M rows per 32 clips).  The
// lane-group kernel above spends 36 issue slots per element there (a quarter of its lanes idle: 12 / 24 float4 columns do not
// divide a warp; 124 registers, 25 % occupancy, ncu: profiles/r02_row_kernels_ncu.txt).  Here a block stages its 128 rows
// (+ 3 halo rows per side, zero outside the clip) in shared memory with cp.async at a (C + 4)-float pitch -- a thread's
// 16-byte reads of consecutive rows are then bank-conflict-free -- and ONE THREAD owns one row: the seven taps are FFMAs
// against kernel-parameter constants (no weight registers, no weight loads), the LayerNorm statistics are thread-local (no
// shuffles), ~15 issue slots per element.
template <int C>
struct DwRowParams {
    const float* x;
    __nv_bfloat16* out;
    int B, T;
    float eps;
    float w[7][C];
    float b[C], lw[C], lb[C];
};

template <int C>
__global__ void __launch_bounds__(128) dwconv7_ln_thread_kernel(const __grid_constant__ DwRowParams<C> p) {
    constexpr int RB = 128, PITCH = C + 4, CH = C / 4;
    extern __shared__ __align__(16) float dw_tile[];                 // (RB + 6) x PITCH
    const int b = blockIdx.y, t0 = blockIdx.x * RB;
    const float* xb = p.x + (long long)b * p.T * C;
    for (int i = threadIdx.x; i < (RB + 6) * CH; i += 128) {
        const int r = i / CH, c = i - r * CH;
        const int t = t0 - 3 + r;
        float* dst = dw_tile + r * PITCH + 4 * c;
        if (t >= 0 && t < p.T) {
            const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dst);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32), "l"(xb + (long long)t * C + 4 * c) : "memory");
        } else {
            *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const int r = threadIdx.x, t = t0 + r;
    if (t >= p.T) return;
    float acc[C];
    const float* row = dw_tile + r * PITCH;
#pragma unroll
    for (int g = 0; g < CH; ++g) {
        float4 a = make_float4(p.b[4 * g], p.b[4 * g + 1], p.b[4 * g + 2], p.b[4 * g + 3]);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const float4 xv = *reinterpret_cast<const float4*>(row + j * PITCH + 4 * g);
            a.x = fmaf(p.w[j][4 * g], xv.x, a.x);
            a.y = fmaf(p.w[j][4 * g + 1], xv.y, a.y);
            a.z = fmaf(p.w[j][4 * g + 2], xv.z, a.z);
            a.w = fmaf(p.w[j][4 * g + 3], xv.w, a.w);
        }
        acc[4 * g] = a.x; acc[4 * g + 1] = a.y; acc[4 * g + 2] = a.z; acc[4 * g + 3] = a.w;
    }
    // LayerNorm over the row: four interleaved partial sums (the lane-group kernel sums four channels per lane first, too)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int g = 0; g < CH; ++g) { s0 += acc[4 * g]; s1 += acc[4 * g + 1]; s2 += acc[4 * g + 2]; s3 += acc[4 * g + 3]; }
    const float inv_c = 1.0f / (float)C;
    const float mean = ((s0 + s1) + (s2 + s3)) * inv_c;
    s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
    for (int g = 0; g < CH; ++g) {
        acc[4 * g] -= mean; acc[4 * g + 1] -= mean; acc[4 * g + 2] -= mean; acc[4 * g + 3] -= mean;
        s0 = fmaf(acc[4 * g], acc[4 * g], s0); s1 = fmaf(acc[4 * g + 1], acc[4 * g + 1], s1);
        s2 = fmaf(acc[4 * g + 2], acc[4 * g + 2], s2); s3 = fmaf(acc[4 * g + 3], acc[4 * g + 3], s3);
    }
    const float rstd = rsqrt_nr(((s0 + s1) + (s2 + s3)) * inv_c + p.eps);
    uint4* orow = reinterpret_cast<uint4*>(p.out + ((long long)b * p.T + t) * C);
#pragma unroll
    for (int k = 0; k < C / 8; ++k) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = 8 * k + 2 * i;
            const __nv_bfloat162 h = __floats2bfloat162_rn(acc[e] * rstd * p.lw[e] + p.lb[e], acc[e + 1] * rstd * p.lw[e + 1] + p.lb[e + 1]);
            pk[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        orow[k] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// Up-layer tail fused with the next unit's prologue (decode side, bf16 operands, C = 48 / 96):
//   y (B, T, C) fp32 -- the 1x1 up conv's output -- -> x_up = ChannelNorm(Upsample_linear(y, S))  (B, T*S, C) fp32   (l3ac/modules.py:162-163)
//                                                   -> a = LayerNorm(dwconv7(x_up))              (B, T*S, C) bf16   (l3ac/modules.py:33-35)
// x_up is the residual stream of the next stage, so it has to be written, but it no longer has to be READ back for the
// depthwise conv: a block builds its 128 (+ 2 x 3 halo) rows of x_up in shared memory from the ~47 rows of y they interpolate,
// one thread per row (lerp weights like ATen's upsample_linear1d, ChannelNorm statistics thread-local), copies the core rows
// out coalesced and runs the thread-per-row dwconv7 + LayerNorm of dwconv7_ln_thread_kernel on the tile.
template <int C>
struct UpDwParams {
    const float* y;
    float* xup;
    __nv_bfloat16* a;
    int B, T, S;
    float cn_eps, ln_eps;
    float cw[C], cb[C];
    float w[7][C];
    float b[C], lw[C], lb[C];
};

constexpr int kUpDwThreads = 160;           // 134 tile rows in one pass (five warps), 128 dwconv rows

template <int C, int S>
__global__ void __launch_bounds__(kUpDwThreads, C >= 96 ? 3 : 5) upsample_cn_dwconv7_ln_kernel(const __grid_constant__ UpDwParams<C> p) {
    constexpr int RB = 128, XR = RB + 6, PITCH = C + 4, CH = C / 4;
    constexpr int YR = (XR + S - 1) / S + 2;                       // rows of y a tile interpolates between
    extern __shared__ __align__(16) float updw_smem[];
    float* ytile = updw_smem;                                        // YR x PITCH
    float* xtile = updw_smem + YR * PITCH;                           // XR x PITCH
    const int b = blockIdx.y, t0 = blockIdx.x * RB;
    const int T = p.T, To = T * S;
    const float* yb = p.y + (long long)b * T * C;
    // source row of output j (exact integer form of floor((j + 0.5) / S - 0.5)); before row 0 the index clamps to 0
    auto src_row = [](int j) { const int q = 2 * j + 1 - S; return q >= 0 ? q / (2 * S) : 0; };
    const int j_first = t0 - 3 > 0 ? t0 - 3 : 0;
    const int iy0 = src_row(j_first);
    for (int i = threadIdx.x; i < YR * CH; i += kUpDwThreads) {
        const int r = i / CH, c = i - r * CH;
        int t = iy0 + r;
        t = t > T - 1 ? T - 1 : t;
        const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(ytile + r * PITCH + 4 * c);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32), "l"(yb + (long long)t * C + 4 * c) : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // ---- x_up rows t0 - 3 .. t0 + 130 -> xtile (zeros outside the clip: the conv's padding)
    if (threadIdx.x < XR) {
        const int r = threadIdx.x, j = t0 - 3 + r;
        float* xrow = xtile + r * PITCH;
        if (j < 0 || j >= To) {
#pragma unroll
            for (int g = 0; g < CH; ++g) *reinterpret_cast<float4*>(xrow + 4 * g) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const int q = 2 * j + 1 - S;
            int i0 = q >= 0 ? q / (2 * S) : 0;
            const float rscale = (float)(1.0 / (double)S);
            const float src = fmaf(rscale, (float)j + 0.5f, -0.5f);            // ATen contracts this to one fma
            float l1 = src - (float)i0;
            if (q < 0 || src < 0.f) l1 = 0.f;                                   // clamped source index: weight 0 on row 0
            int i1 = i0 + 1;
            i0 = i0 > T - 1 ? T - 1 : i0;
            i1 = i1 > T - 1 ? T - 1 : i1;
            if (q < 0) i1 = 0;
            const float w1 = l1, w0 = 1.0f - l1;
            const float* r0 = ytile + (i0 - iy0) * PITCH;
            const float* r1 = ytile + (i1 - iy0) * PITCH;
            float v[C];
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int g = 0; g < CH; ++g) {
                const float4 a0 = *reinterpret_cast<const float4*>(r0 + 4 * g), a1 = *reinterpret_cast<const float4*>(r1 + 4 * g);
                v[4 * g] = __fmaf_rn(w1, a1.x, __fmul_rn(w0, a0.x));
                v[4 * g + 1] = __fmaf_rn(w1, a1.y, __fmul_rn(w0, a0.y));
                v[4 * g + 2] = __fmaf_rn(w1, a1.z, __fmul_rn(w0, a0.z));
                v[4 * g + 3] = __fmaf_rn(w1, a1.w, __fmul_rn(w0, a0.w));
                s0 += v[4 * g]; s1 += v[4 * g + 1]; s2 += v[4 * g + 2]; s3 += v[4 * g + 3];
            }
            const float inv_c = 1.0f / (float)C;
            const float mean = ((s0 + s1) + (s2 + s3)) * inv_c;
            s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
            for (int g = 0; g < CH; ++g) {
                v[4 * g] -= mean; v[4 * g + 1] -= mean; v[4 * g + 2] -= mean; v[4 * g + 3] -= mean;
                s0 = fmaf(v[4 * g], v[4 * g], s0); s1 = fmaf(v[4 * g + 1], v[4 * g + 1], s1);
                s2 = fmaf(v[4 * g + 2], v[4 * g + 2], s2); s3 = fmaf(v[4 * g + 3], v[4 * g + 3], s3);
            }
            const float rstd = rsqrt_nr(((s0 + s1) + (s2 + s3)) * inv_c + p.cn_eps);
#pragma unroll
            for (int g = 0; g < CH; ++g)
                *reinterpret_cast<float4*>(xrow + 4 * g) =
                    make_float4(v[4 * g] * rstd * p.cw[4 * g] + p.cb[4 * g], v[4 * g + 1] * rstd * p.cw[4 * g + 1] + p.cb[4 * g + 1],
                                v[4 * g + 2] * rstd * p.cw[4 * g + 2] + p.cb[4 * g + 2], v[4 * g + 3] * rstd * p.cw[4 * g + 3] + p.cb[4 * g + 3]);
        }
    }
    __syncthreads();
    // ---- the tile's own rows of x_up -> HBM, coalesced (they are one contiguous block of the (B, To, C) tensor)
    const int n_rows = To - t0 < RB ? To - t0 : RB;
    {
        float4* dst = reinterpret_cast<float4*>(p.xup + ((long long)b * To + t0) * C);
        for (int i = threadIdx.x; i < n_rows * CH; i += kUpDwThreads) {
            const int r = i / CH, c = i - r * CH;
            dst[i] = *reinterpret_cast<const float4*>(xtile + (r + 3) * PITCH + 4 * c);
        }
    }
    // ---- dwconv7 + LayerNorm, one thread per row (dwconv7_ln_thread_kernel)
    const int r = threadIdx.x;
    if (r >= n_rows) return;
    float acc[C];
    const float* row = xtile + r * PITCH;
#pragma unroll
    for (int g = 0; g < CH; ++g) {
        float4 a = make_float4(p.b[4 * g], p.b[4 * g + 1], p.b[4 * g + 2], p.b[4 * g + 3]);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const float4 xv = *reinterpret_cast<const float4*>(row + j * PITCH + 4 * g);
            a.x = fmaf(p.w[j][4 * g], xv.x, a.x);
            a.y = fmaf(p.w[j][4 * g + 1], xv.y, a.y);
            a.z = fmaf(p.w[j][4 * g + 2], xv.z, a.z);
            a.w = fmaf(p.w[j][4 * g + 3], xv.w, a.w);
        }
        acc[4 * g] = a.x; acc[4 * g + 1] = a.y; acc[4 * g + 2] = a.z; acc[4 * g + 3] = a.w;
    }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int g = 0; g < CH; ++g) { s0 += acc[4 * g]; s1 += acc[4 * g + 1]; s2 += acc[4 * g + 2]; s3 += acc[4 * g + 3]; }
    const float inv_c = 1.0f / (float)C;
    const float mean = ((s0 + s1) + (s2 + s3)) * inv_c;
    s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
    for (int g = 0; g < CH; ++g) {
        acc[4 * g] -= mean; acc[4 * g + 1] -= mean; acc[4 * g + 2] -= mean; acc[4 * g + 3] -= mean;
        s0 = fmaf(acc[4 * g], acc[4 * g], s0); s1 = fmaf(acc[4 * g + 1], acc[4 * g + 1], s1);
        s2 = fmaf(acc[4 * g + 2], acc[4 * g + 2], s2); s3 = fmaf(acc[4 * g + 3], acc[4 * g + 3], s3);
    }
    const float rstd = rsqrt_nr(((s0 + s1) + (s2 + s3)) * inv_c + p.ln_eps);
    uint4* orow = reinterpret_cast<uint4*>(p.a + ((long long)b * To + t0 + r) * C);
#pragma unroll
    for (int k = 0; k < C / 8; ++k) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = 8 * k + 2 * i;
            const __nv_bfloat162 h = __floats2bfloat162_rn(acc[e] * rstd * p.lw[e] + p.lb[e], acc[e + 1] * rstd * p.lw[e + 1] + p.lb[e + 1]);
            pk[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        orow[k] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// Wide-channel vector variant (C % 128 == 0, C <= 512): one CTA of C/4 threads (one float4 of channels per thread, one
// warp per 128 channels) handles R consecutive time steps.  The per-time-step statistics are a warp shuffle reduction
// followed by one shared-memory exchange between the CTA's warps for all R steps at once (two exchanges: mean, then
// centred sum of squares).  ~16 instructions per output element instead of ~34 for the one-channel-per-thread tile.
template <int R, int NW, typename OutT>
__global__ void __launch_bounds__(32 * NW) dwconv7_ln_wide_kernel(const float* __restrict__ x, int B, int T, int C,
                                                                  const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                                  const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                                  float eps, OutT* __restrict__ out, OutT* __restrict__ out_lo) {
    __shared__ __align__(16) float s_part[2][R][4];      // per-warp partial sums, read back as one float4 per row
    const int g = threadIdx.x, lane = g & 31, warp = g >> 5;
    const int C4 = C >> 2;
    const int b = blockIdx.y, t0 = blockIdx.x * R;
    float4 w[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) w[j] = __ldg(reinterpret_cast<const float4*>(dw_w + j * C) + g);
    const float4 bias = __ldg(reinterpret_cast<const float4*>(dw_b) + g);
    const float4* xb = reinterpret_cast<const float4*>(x + (long long)b * T * C) + g;
    float4 xr[R + 6];
    const bool interior = t0 >= 3 && t0 + R + 3 <= T;      // block-uniform: no clip edge inside the R + 6 rows
    if (interior) {
        const float4* p0 = xb + (long long)(t0 - 3) * C4;
#pragma unroll
        for (int r = 0; r < R + 6; ++r) xr[r] = __ldg(p0 + r * C4);
    } else {
#pragma unroll
        for (int r = 0; r < R + 6; ++r) {
            const int t = t0 + r - 3;
            xr[r] = (t >= 0 && t < T) ? __ldg(xb + (long long)t * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float4 y[R];
    float st[8];
    static_assert(R == 8, "the statistics use the 8-row group reduction");
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float4 a = bias;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const float2 lo = ffma2(make_float2(w[j].x, w[j].y), make_float2(xr[r + j].x, xr[r + j].y), make_float2(a.x, a.y));
            const float2 hi = ffma2(make_float2(w[j].z, w[j].w), make_float2(xr[r + j].z, xr[r + j].w), make_float2(a.z, a.w));
            a = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        y[r] = a;
        st[r] = (a.x + a.y) + (a.z + a.w);
    }
    {   // eight row sums in one reduction: lane 4 r ends up with row r's warp total
        const float t = group_reduce8<32>(st);
        if ((lane & 3) == 0) s_part[0][lane >> 2][warp] = t;
    }
    __syncthreads();
    const float inv_c = 1.0f / (float)C;
    auto total = [](const float4 p) {      // sum of the NW per-warp partials, in warp order
        float m = p.x;
        if (NW > 1) m += p.y;
        if (NW > 2) m += p.z;
        if (NW > 3) m += p.w;
        return m;
    };
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float m = total(*reinterpret_cast<const float4*>(s_part[0][r])) * inv_c;
        y[r].x -= m; y[r].y -= m; y[r].z -= m; y[r].w -= m;
        st[r] = (y[r].x * y[r].x + y[r].y * y[r].y) + (y[r].z * y[r].z + y[r].w * y[r].w);
    }
    {
        const float t = group_reduce8<32>(st);
        if ((lane & 3) == 0) s_part[1][lane >> 2][warp] = t;
    }
    __syncthreads();
    const float4 lw = __ldg(reinterpret_cast<const float4*>(ln_w) + g), lb = __ldg(reinterpret_cast<const float4*>(ln_b) + g);
    const long long out_row0 = ((long long)b * T + t0) * C + 4 * g;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (!interior && t0 + r >= T) break;
        const float rstd = rsqrt_nr(total(*reinterpret_cast<const float4*>(s_part[1][r])) * inv_c + eps);
        store_act4<OutT>(out, out_lo, out_row0 + (long long)r * C,
                         make_float4(y[r].x * rstd * lw.x + lb.x, y[r].y * rstd * lw.y + lb.y, y[r].z * rstd * lw.z + lb.z,
                                     y[r].w * rstd * lw.w + lb.w));
    }
}

// Wide-channel variant (C a multiple of 32, 128 <= C <= 512): one CTA = TT consecutive time steps x all channels,
// one thread per channel.  Each input element is loaded once (38 coalesced loads in flight per thread), the conv
// outputs stay in registers, and the per-time-step LayerNorm statistics are reduced warp -> shared memory -> CTA
// (two passes: mean, then centred sum of squares, like F.layer_norm).
template <int TT, typename OutT>
__global__ void __launch_bounds__(512) dwconv7_ln_tile_kernel(const float* __restrict__ x, int B, int T, int C,
                                                              const float* __restrict__ dw_w,
                                                              const float* __restrict__ dw_b,
                                                              const float* __restrict__ ln_w,
                                                              const float* __restrict__ ln_b, float eps,
                                                              OutT* __restrict__ out, OutT* __restrict__ out_lo) {
    __shared__ float s_part[TT][16];     // [time][warp]
    __shared__ float s_stat[TT];
    const int c = threadIdx.x, lane = c & 31, warp = c >> 5, nwarps = blockDim.x >> 5;
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    const float* xb = x + (long long)b * T * C + c;
    float w[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) w[j] = __ldg(dw_w + j * C + c);
    const float bias = __ldg(dw_b + c), lw = __ldg(ln_w + c), lb = __ldg(ln_b + c);
    float xr[TT + 6];
#pragma unroll
    for (int r = 0; r < TT + 6; ++r) {
        const int t = t0 + r - 3;
        xr[r] = (t >= 0 && t < T) ? __ldg(xb + (long long)t * C) : 0.f;
    }
    float y[TT];
#pragma unroll
    for (int i = 0; i < TT; ++i) {
        float a = bias;
#pragma unroll
        for (int j = 0; j < 7; ++j) a = fmaf(w[j], xr[i + j], a);
        y[i] = a;
    }
    const float inv_c = 1.0f / (float)C;
    // pass 1: mean over channels for every time step
#pragma unroll
    for (int i = 0; i < TT; ++i) {
        const float s = warp_sum(y[i]);
        if (lane == 0) s_part[i][warp] = s;
    }
    __syncthreads();
    if (c < TT) {
        float s = 0.f;
        for (int q = 0; q < nwarps; ++q) s += s_part[c][q];
        s_stat[c] = s * inv_c;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TT; ++i) y[i] -= s_stat[i];
    __syncthreads();
    // pass 2: biased variance
#pragma unroll
    for (int i = 0; i < TT; ++i) {
        const float s = warp_sum(y[i] * y[i]);
        if (lane == 0) s_part[i][warp] = s;
    }
    __syncthreads();
    if (c < TT) {
        float s = 0.f;
        for (int q = 0; q < nwarps; ++q) s += s_part[c][q];
        s_stat[c] = rsqrt_nr(s * inv_c + eps);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TT; ++i) {
        const int t = t0 + i;
        if (t < T) store_act<OutT>(out, out_lo, ((long long)b * T + t) * C + c, y[i] * s_stat[i] * lw + lb);
    }
}

template <int CPL, typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long M, int C,
                                                        const float* __restrict__ w,
                                                        const float* __restrict__ b, float eps,
                                                        OutT* __restrict__ out, OutT* __restrict__ out_lo) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M;
         row += (long long)gridDim.x * warps_per_block) {
        const float* xr = x + row * C;
        float v[CPL];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
            v[i] = (c < C) ? xr[c] : 0.f;
            s += v[i];
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const float dlt = v[i] - mean;
            q += (lane + 32 * i < C) ? dlt * dlt : 0.f;
        }
        const float rstd = rsqrt_nr(warp_sum(q) / (float)C + eps);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) store_act<OutT>(out, out_lo, row * C + c, (v[i] - mean) * rstd * __ldg(w + c) + __ldg(b + c));
        }
    }
}

// fp32 -> (hi, lo) bf16 planes, 4 elements per thread
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, long long n4,
                                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
        const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
        const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
        uint2 ph, pl;
        ph.x = *reinterpret_cast<const uint32_t*>(&h01);
        ph.y = *reinterpret_cast<const uint32_t*>(&h23);
        pl.x = *reinterpret_cast<const uint32_t*>(&l01);
        pl.y = *reinterpret_cast<const uint32_t*>(&l23);
        reinterpret_cast<uint2*>(hi)[i] = ph;
        reinterpret_cast<uint2*>(lo)[i] = pl;
    }
}

template <typename OutT>
__global__ void __launch_bounds__(256) snake_kernel(const float* __restrict__ x, long long n, int C,
                                                    const float* __restrict__ alpha, OutT* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float a = __ldg(alpha + c);
        out[i] = cvt_out<OutT>(snake_f(x[i], a, 1.0f / (a + kEps)));
    }
}

// ------------------------------------------------------------------------------------------
// linear upsample x`scale` along time (+ optional LayerNorm over C).  One warp per output row.
// Index math follows ATen's area_pixel_compute_source_index (align_corners=False):
//   src = max(0, (1/scale) * (j + 0.5) - 0.5), i0 = floor(src), i1 = min(i0 + 1, T - 1), lam = src - i0.
// ------------------------------------------------------------------------------------------
template <int CPL>
__global__ void __launch_bounds__(256) upsample_cn_kernel(const float* __restrict__ x, int B, int T, int C,
                                                          int scale, const float* __restrict__ cn_w,
                                                          const float* __restrict__ cn_b, float eps,
                                                          float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int To = T * scale;
    const long long rows = (long long)B * To;    // scalar fallback (C % 4 != 0 or unaligned pointers): 1-D grid over all clips
    const float rscale = (float)(1.0 / (double)scale);
    for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows;
         row += (long long)gridDim.x * warps_per_block) {
        const int b = (int)(row / To);
        const int j = (int)(row % To);
        float src = fmaf(rscale, (float)j + 0.5f, -0.5f);   // ATen contracts this to one fma
        src = src < 0.f ? 0.f : src;
        int i0 = (int)src;
        if (i0 > T - 1) i0 = T - 1;
        const int i1 = i0 + (i0 < T - 1 ? 1 : 0);
        const float l1 = src - (float)i0;
        const float l0 = 1.0f - l1;
        const float* x0 = x + ((long long)b * T + i0) * C;
        const float* x1 = x + ((long long)b * T + i1) * C;
        float v[CPL];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
            v[i] = (c < C) ? __fmaf_rn(l1, __ldg(x1 + c), __fmul_rn(l0, __ldg(x0 + c))) : 0.f;
            s += v[i];
        }
        float* orow = out + row * C;
        if (cn_w == nullptr) {
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = lane + 32 * i;
                if (c < C) orow[c] = v[i];
            }
            continue;
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const float dlt = v[i] - mean;
            q += (lane + 32 * i < C) ? dlt * dlt : 0.f;
        }
        const float rstd = rsqrt_nr(warp_sum(q) / (float)C + eps);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) orow[c] = (v[i] - mean) * rstd * __ldg(cn_w + c) + __ldg(cn_b + c);
        }
    }
}

// ------------------------------------------------------------------------------------------
// EnhanceBlock.  Branch j in {0..3}: pool kernel k_j in {1,3,5,9}, conv dilation d_j in {1,2,3,5}.
//   p_j = trend_pool(x[:, 0], k_j)  (k=1: identity without abs; else avg_k(max_k(|x|)), max ignores
//   out-of-range samples, avg zero-pads and always divides by k), zero outside [0,T) for the conv.
//   y_j[t] = cb_j + sum_i cw_j[i] * p_j[t + (i-3) d_j]
// kReach = 8 (pool) + 15 (conv) = 23 samples on each side of a tile.
// ------------------------------------------------------------------------------------------
constexpr int kEnhReach = 23;
constexpr int kEnhChunk = 1024;   // time steps per stats block (partials granularity)
constexpr int kEnhTile = 512;     // time steps per apply block (128 made ~30k tiny CTAs per full-rate launch: 1.4 TB/s)

// Computes y_j[t0 + i] for i in [0, CH) into ys[j][i] (shared, [4][CH]).  xs/ms/ps: [CH + 2*kEnhReach].
template <int CH>
__device__ __forceinline__ void enhance_branches(const float* __restrict__ xb /* x + b*T*C */, int T, int C, int t0,
                                                 const float* __restrict__ conv_w,
                                                 const float* __restrict__ conv_b, float* xs, float* ms, float* ps,
                                                 float* ys) {
    constexpr int R = kEnhReach;
    constexpr int W = CH + 2 * R;
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
        const int t = t0 - R + i;
        xs[i] = (t >= 0 && t < T) ? __ldg(xb + (long long)t * C) : 0.f;
    }
    __syncthreads();
    const int pool_k[4] = {1, 3, 5, 9};
    const int dil[4] = {1, 2, 3, 5};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = pool_k[j], half = k >> 1, d = dil[j];
        const float* src = xs;
        if (k > 1) {
            for (int i = threadIdx.x; i < W; i += blockDim.x) {
                const int t = t0 - R + i;
                float m = 0.f;
                if (t >= 0 && t < T && i >= half && i < W - half) {
                    for (int e = -half; e <= half; ++e) m = fmaxf(m, fabsf(xs[i + e]));   // OOB xs are 0 <= |x|
                }
                ms[i] = m;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < W; i += blockDim.x) {
                const int t = t0 - R + i;
                float a = 0.f;
                if (t >= 0 && t < T && i >= 2 * half && i < W - 2 * half) {
                    for (int e = -half; e <= half; ++e) a += ms[i + e];
                    a = a / (float)k;
                }
                ps[i] = a;
            }
            __syncthreads();
            src = ps;
        }
        float w[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) w[i] = __ldg(conv_w + j * 7 + i);
        const float cb = __ldg(conv_b + j);
        for (int i = threadIdx.x; i < CH; i += blockDim.x) {
            float acc = cb;
#pragma unroll
            for (int q = 0; q < 7; ++q) acc = fmaf(w[q], src[i + R + (q - 3) * d], acc);
            ys[j * CH + i] = acc;
        }
        __syncthreads();
    }
}

// partials layout: [B][nchunk][8] = (sum_j, sumsq_j) for j = 0..3
__global__ void __launch_bounds__(256) enhance_stats_kernel(const float* __restrict__ x, int B, int T, int C,
                                                            const float* __restrict__ conv_w,
                                                            const float* __restrict__ conv_b,
                                                            float* __restrict__ partials, float4* __restrict__ branches) {
    constexpr int CH = kEnhChunk;
    __shared__ float xs[CH + 2 * kEnhReach], ms[CH + 2 * kEnhReach], ps[CH + 2 * kEnhReach];
    __shared__ float ys[4 * CH];
    __shared__ float red[8][8];
    const int b = blockIdx.y, chunk = blockIdx.x, t0 = chunk * CH;
    enhance_branches<CH>(x + (long long)b * T * C, T, C, t0, conv_w, conv_b, xs, ms, ps, ys);
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    for (int i = threadIdx.x; i < CH; i += blockDim.x) {
        if (t0 + i < T) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float y = ys[j * CH + i];
                s[j] += y;
                q[j] = fmaf(y, y, q[j]);
            }
            // the four (un-normalised) branch signals of this sample, so that the apply pass is a pure streaming kernel
            if (branches) branches[(long long)b * T + t0 + i] = make_float4(ys[i], ys[CH + i], ys[2 * CH + i], ys[3 * CH + i]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s[j] = warp_sum(s[j]);
        q[j] = warp_sum(q[j]);
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[warp][2 * j] = s[j];
            red[warp][2 * j + 1] = q[j];
        }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float a = 0.f;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) a += red[wv][threadIdx.x];
        partials[((long long)b * gridDim.x + chunk) * 8 + threadIdx.x] = a;
    }
}

template <typename OutT>
__global__ void __launch_bounds__(256) enhance_apply_kernel(const float* __restrict__ x, int B, int T, int C,
                                                            const float* __restrict__ conv_w,
                                                            const float* __restrict__ conv_b,
                                                            const float* __restrict__ in_w,
                                                            const float* __restrict__ in_b,
                                                            const float* __restrict__ merge_w,
                                                            const float* __restrict__ merge_b,
                                                            const float* __restrict__ partials, int nchunk,
                                                            OutT* __restrict__ out) {
    constexpr int CH = kEnhTile;
    __shared__ float xs[CH + 2 * kEnhReach], ms[CH + 2 * kEnhReach], ps[CH + 2 * kEnhReach];
    __shared__ float ys[4 * CH];
    __shared__ float s_scale[4], s_shift[4];
    const int b = blockIdx.y, t0 = blockIdx.x * CH;
    if (threadIdx.x < 4) {
        const int j = threadIdx.x;
        double s = 0.0, q = 0.0;
        for (int c = 0; c < nchunk; ++c) {
            s += (double)partials[((long long)b * nchunk + c) * 8 + 2 * j];
            q += (double)partials[((long long)b * nchunk + c) * 8 + 2 * j + 1];
        }
        const double mean = s / (double)T;
        double var = q / (double)T - mean * mean;   // biased variance (InstanceNorm1d)
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + 1e-5));
        const float g = in_w[j] * rstd;
        s_scale[j] = g;
        s_shift[j] = in_b[j] - (float)mean * g;
    }
    enhance_branches<CH>(x + (long long)b * T * C, T, C, t0, conv_w, conv_b, xs, ms, ps, ys);
    // (enhance_branches ends with __syncthreads, so s_scale/s_shift are visible)
    for (int i = threadIdx.x; i < 4 * CH; i += blockDim.x) {
        const int j = i / CH;
        ys[i] = fmaf(ys[i], s_scale[j], s_shift[j]);
    }
    __syncthreads();
    const int nt = min(CH, T - t0);
    const long long base = ((long long)b * T + t0) * C;
    for (int e = threadIdx.x; e < nt * C; e += blockDim.x) {
        const int i = e / C, c = e - i * C;
        const float4 mw = __ldg(reinterpret_cast<const float4*>(merge_w) + c);
        float y = __ldg(merge_b + c);
        y = fmaf(mw.x, ys[i], y);
        y = fmaf(mw.y, ys[CH + i], y);
        y = fmaf(mw.z, ys[2 * CH + i], y);
        y = fmaf(mw.w, ys[3 * CH + i], y);
        const float xv = x[base + e];
        out[base + e] = cvt_out<OutT>(fmaf(y, xv, xv));
    }
}

// ------------------------------------------------------------------------------------------
// Vectorised row kernels.  A row of C channels is handled by a group of GS lanes (GS = 8, 16 or 32), each lane
// owning VPL float4 (chunk index g + GS*v); a warp therefore processes 32/GS rows at once and U such row sets per
// iteration, with every load issued before the first use.  Row statistics are xor-shuffle reductions inside the group.
// ------------------------------------------------------------------------------------------

template <int GS, int VPL, int U>
__global__ void __launch_bounds__(256) upsample_cn_vec_kernel(const float* __restrict__ x, int B, int T, int C, int scale,
                                                              const float* __restrict__ cn_w, const float* __restrict__ cn_b,
                                                              float eps, float* __restrict__ out) {
    constexpr int RW = 32 / GS;                 // rows per warp
    const int lane = threadIdx.x & 31, g = lane % GS, sub = lane / GS;
    const int C4 = C >> 2;
    const int To = T * scale;
    const int bclip = blockIdx.y;                // one clip per grid row: no per-row division by To (~20 instructions per row and lane)
    const long long rows = To;
    const float rscale = (float)(1.0 / (double)scale);
    const float inv_c = 1.0f / (float)C;
    float4 w4[VPL], b4[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c4 = g + GS * v;
        w4[v] = (cn_w && c4 < C4) ? __ldg(reinterpret_cast<const float4*>(cn_w) + c4) : make_float4(1.f, 1.f, 1.f, 1.f);
        b4[v] = (cn_b && c4 < C4) ? __ldg(reinterpret_cast<const float4*>(cn_b) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long base = warp_global * (RW * U); base < rows; base += nwarps * (RW * U)) {
        float4 a0[U][VPL], a1[U][VPL];
        float l1[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = base + u * RW + sub;
            ok[u] = row < rows;
            const int j = (int)(ok[u] ? row : rows - 1);
            const int b = bclip;
            float src = fmaf(rscale, (float)j + 0.5f, -0.5f);   // ATen contracts this to one fma
            src = src < 0.f ? 0.f : src;
            int i0 = (int)src;
            if (i0 > T - 1) i0 = T - 1;
            const int i1 = i0 + (i0 < T - 1 ? 1 : 0);
            l1[u] = src - (float)i0;
            const float4* p0 = reinterpret_cast<const float4*>(x + ((long long)b * T + i0) * C);
            const float4* p1 = reinterpret_cast<const float4*>(x + ((long long)b * T + i1) * C);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c4 = g + GS * v;
                a0[u][v] = (c4 < C4) ? __ldg(p0 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                a1[u][v] = (c4 < C4) ? __ldg(p1 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float w1 = l1[u], w0 = 1.0f - l1[u];
            float4 val[VPL];
            float s = 0.f;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                val[v].x = __fmaf_rn(w1, a1[u][v].x, __fmul_rn(w0, a0[u][v].x));
                val[v].y = __fmaf_rn(w1, a1[u][v].y, __fmul_rn(w0, a0[u][v].y));
                val[v].z = __fmaf_rn(w1, a1[u][v].z, __fmul_rn(w0, a0[u][v].z));
                val[v].w = __fmaf_rn(w1, a1[u][v].w, __fmul_rn(w0, a0[u][v].w));
                s += (val[v].x + val[v].y) + (val[v].z + val[v].w);      // lanes beyond C4 hold zeros
            }
            float mean = 0.f, rstd = 1.f;
            if (cn_w) {
                mean = group_sum<GS>(s) * inv_c;
                float q = 0.f;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    if (g + GS * v < C4) {
                        const float dx = val[v].x - mean, dy = val[v].y - mean, dz = val[v].z - mean, dw = val[v].w - mean;
                        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                    }
                }
                rstd = rsqrt_nr(group_sum<GS>(q) * inv_c + eps);
            }
            if (!ok[u]) continue;
            float4* orow = reinterpret_cast<float4*>(out + ((long long)bclip * To + base + u * RW + sub) * C);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c4 = g + GS * v;
                if (c4 < C4) {
                    float4 o;
                    o.x = (val[v].x - mean) * rstd * w4[v].x + b4[v].x;
                    o.y = (val[v].y - mean) * rstd * w4[v].y + b4[v].y;
                    o.z = (val[v].z - mean) * rstd * w4[v].z + b4[v].z;
                    o.w = (val[v].w - mean) * rstd * w4[v].w + b4[v].w;
                    orow[c4] = o;
                }
            }
        }
    }
}

template <int GS, int VPL, int U, typename OutT>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const float* __restrict__ x, long long M, int C,
                                                            const float* __restrict__ w, const float* __restrict__ b,
                                                            float eps, OutT* __restrict__ out, OutT* __restrict__ out_lo) {
    constexpr int RW = 32 / GS;
    const int lane = threadIdx.x & 31, g = lane % GS, sub = lane / GS;
    const int C4 = C >> 2;
    const float inv_c = 1.0f / (float)C;
    float4 w4[VPL], b4[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        const int c4 = g + GS * v;
        w4[v] = (c4 < C4) ? __ldg(reinterpret_cast<const float4*>(w) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        b4[v] = (c4 < C4) ? __ldg(reinterpret_cast<const float4*>(b) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long base = warp_global * (RW * U); base < M; base += nwarps * (RW * U)) {
        float4 val[U][VPL];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = base + u * RW + sub;
            const float4* p = reinterpret_cast<const float4*>(x + (row < M ? row : M - 1) * C);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c4 = g + GS * v;
                val[u][v] = (c4 < C4) ? __ldg(p + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float s = 0.f;
#pragma unroll
            for (int v = 0; v < VPL; ++v) s += (val[u][v].x + val[u][v].y) + (val[u][v].z + val[u][v].w);
            const float mean = group_sum<GS>(s) * inv_c;
            float q = 0.f;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if (g + GS * v < C4) {
                    const float dx = val[u][v].x - mean, dy = val[u][v].y - mean, dz = val[u][v].z - mean, dw = val[u][v].w - mean;
                    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                }
            }
            const float rstd = rsqrt_nr(group_sum<GS>(q) * inv_c + eps);
            const long long row = base + u * RW + sub;
            if (row >= M) continue;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c4 = g + GS * v;
                if (c4 < C4) {
                    const long long i = row * C + 4 * c4;
                    store_act4<OutT>(out, out_lo, i,
                                     make_float4((val[u][v].x - mean) * rstd * w4[v].x + b4[v].x,
                                                 (val[u][v].y - mean) * rstd * w4[v].y + b4[v].y,
                                                 (val[u][v].z - mean) * rstd * w4[v].z + b4[v].z,
                                                 (val[u][v].w - mean) * rstd * w4[v].w + b4[v].w));
                }
            }
        }
    }
}

// Linear upsample (+ ChannelNorm) by an integer factor S in {2,3,4,5}, run variant (C % 4 == 0, C <= 128): a group of GS lanes
// (one float4 of channels per lane) produces a run of R consecutive output rows of one clip, R a multiple of S, so that the
// input rows every output interpolates between are compile-time offsets into the NR = R/S + 2 rows the group loads ONCE
// (0.6 - 0.75 loads per output row instead of 2), the per-row index arithmetic shrinks to one FMA for the weight, and the
// statistics of the run share one 8-row group reduction each.  The interpolation weight is computed per row exactly like
// ATen's upsample_linear1d (src = scale^-1 (j + 0.5) - 0.5 in fp32, clamped at 0); rows clamped at the clip edges make the
// edge cases fall out of the same formula.
__host__ __device__ constexpr int ups_off(int r, int S) {      // floor((r + 0.5) / S - 0.5) + 1 = i0 - ibase for output j0 + r, j0 % S == 0
    return (2 * r + 1 - S >= 0) ? (2 * r + 1 - S) / (2 * S) + 1 : 0;
}

template <int GS, int S>
__global__ void __launch_bounds__(256) upsample_cn_run_kernel(const float* __restrict__ x, int B, int T, int C,
                                                              const float* __restrict__ cn_w, const float* __restrict__ cn_b,
                                                              float eps, float* __restrict__ out) {
    constexpr int R = (S == 3) ? 6 : (S == 5) ? 5 : 8;
    constexpr int NR = ups_off(R - 1, S) + 2;
    constexpr int RW = 32 / GS;
    static_assert(R % S == 0 && R <= 8, "a run is a whole number of input intervals and fits the 8-row reduction");
    const int lane = threadIdx.x & 31, g = lane % GS, sub = lane / GS;
    const int C4 = C >> 2;
    const bool act = g < C4;
    const int b = blockIdx.y;
    const int To = T * S;
    const int runs = (To + R - 1) / R;
    const float rscale = (float)(1.0 / (double)S);
    const float inv_c = 1.0f / (float)C;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 w4 = (cn_w && act) ? __ldg(reinterpret_cast<const float4*>(cn_w) + g) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 b4 = (cn_b && act) ? __ldg(reinterpret_cast<const float4*>(cn_b) + g) : z4;
    const float4* xb = reinterpret_cast<const float4*>(x + (long long)b * T * C) + g;
    float* ob = out + (long long)b * To * C + 4 * g;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int base = warp_global * RW; base < runs; base += nwarps * RW) {
        const int run = base + sub;
        const bool run_ok = run < runs;
        const int rr = run_ok ? run : runs - 1;
        const int j0 = rr * R, ibase = rr * (R / S) - 1;
        float4 xr[NR];
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            int t = ibase + k;
            t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
            xr[k] = act ? __ldg(xb + (long long)t * C4) : z4;
        }
        float4 val[R];
        float st[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) st[r] = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int o = ups_off(r, S);
            float src = fmaf(rscale, (float)(j0 + r) + 0.5f, -0.5f);      // ATen contracts this to one fma
            float l1 = src - (float)(ibase + o);
            if (src < 0.f) l1 = 0.f;                                       // clamped source index: weight 0 on row 0
            const float w1 = l1, w0 = 1.0f - l1;
            const float4 a0 = (src < 0.f) ? xr[o + 1] : xr[o], a1 = xr[o + 1];
            float4 v;
            v.x = __fmaf_rn(w1, a1.x, __fmul_rn(w0, a0.x));
            v.y = __fmaf_rn(w1, a1.y, __fmul_rn(w0, a0.y));
            v.z = __fmaf_rn(w1, a1.z, __fmul_rn(w0, a0.z));
            v.w = __fmaf_rn(w1, a1.w, __fmul_rn(w0, a0.w));
            val[r] = v;
            st[r] = (v.x + v.y) + (v.z + v.w);                              // inactive lanes hold zeros
        }
        if (cn_w) {
            group_bcast8<GS>(group_reduce8<GS>(st), st);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float mean = st[r] * inv_c;
                val[r].x -= mean; val[r].y -= mean; val[r].z -= mean; val[r].w -= mean;
                st[r] = act ? (val[r].x * val[r].x + val[r].y * val[r].y) + (val[r].z * val[r].z + val[r].w * val[r].w) : 0.f;
            }
            group_bcast8<GS>(group_reduce8<GS>(st), st);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!act || !run_ok || j0 + r >= To) continue;
            const float rstd = cn_w ? rsqrt_nr(st[r] * inv_c + eps) : 1.0f;
            float4 o4;
            o4.x = val[r].x * rstd * w4.x + b4.x;
            o4.y = val[r].y * rstd * w4.y + b4.y;
            o4.z = val[r].z * rstd * w4.z + b4.z;
            o4.w = val[r].w * rstd * w4.w + b4.w;
            *reinterpret_cast<float4*>(ob + (long long)(j0 + r) * C) = o4;
        }
    }
}

// EnhanceBlock gating, vectorised: a thread owns one float4 of channels (its merge weights stay in registers) and walks
// the rows of the tile with a stride, four rows in flight.
template <typename OutT, int CH>
__global__ void __launch_bounds__(256) enhance_apply_vec_kernel(const float* __restrict__ x, int B, int T, int C,
                                                                const float* __restrict__ conv_w,
                                                                const float* __restrict__ conv_b,
                                                                const float* __restrict__ in_w, const float* __restrict__ in_b,
                                                                const float* __restrict__ merge_w,
                                                                const float* __restrict__ merge_b,
                                                                const float* __restrict__ partials, int nchunk,
                                                                OutT* __restrict__ out) {
    __shared__ float xs[CH + 2 * kEnhReach], ms[CH + 2 * kEnhReach], ps[CH + 2 * kEnhReach];
    __shared__ float ys[4 * CH];
    __shared__ float s_scale[4], s_shift[4];
    const int b = blockIdx.y, t0 = blockIdx.x * CH;
    {   // The branch signals below need only channel 0; start pulling the whole (tile x C) block towards L2 now so that the
        // streaming phase at the end does not begin with a cold HBM round trip.
        const char* tile = reinterpret_cast<const char*>(x + ((long long)b * T + t0) * C);
        const long long tile_bytes = (long long)min(CH, T - t0) * C * 4;
        for (long long off = (long long)threadIdx.x * 128; off < tile_bytes; off += (long long)blockDim.x * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(tile + off));
    }
    if (threadIdx.x < 4) {
        const int j = threadIdx.x;
        double s = 0.0, q = 0.0;
        for (int c = 0; c < nchunk; ++c) {
            s += (double)partials[((long long)b * nchunk + c) * 8 + 2 * j];
            q += (double)partials[((long long)b * nchunk + c) * 8 + 2 * j + 1];
        }
        const double mean = s / (double)T;
        double var = q / (double)T - mean * mean;   // biased variance (InstanceNorm1d)
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + 1e-5));
        const float gg = in_w[j] * rstd;
        s_scale[j] = gg;
        s_shift[j] = in_b[j] - (float)mean * gg;
    }
    enhance_branches<CH>(x + (long long)b * T * C, T, C, t0, conv_w, conv_b, xs, ms, ps, ys);
    for (int i = threadIdx.x; i < 4 * CH; i += blockDim.x) {
        const int j = i / CH;
        ys[i] = fmaf(ys[i], s_scale[j], s_shift[j]);
    }
    __syncthreads();
    const int nt = min(CH, T - t0);
    const int C4 = C >> 2;
    const int rpp = blockDim.x / C4;                  // rows per pass
    const int c4 = threadIdx.x % C4, r0 = threadIdx.x / C4;
    if (r0 >= rpp) return;
    float4 mw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) mw[q] = __ldg(reinterpret_cast<const float4*>(merge_w) + 4 * c4 + q);   // merge_w[c][0..3]
    const float4 mb = __ldg(reinterpret_cast<const float4*>(merge_b) + c4);
    const float4* xrow = reinterpret_cast<const float4*>(x + ((long long)b * T + t0) * C) + c4;
    OutT* obase = out + ((long long)b * T + t0) * C + 4 * c4;
    for (int i0 = r0; i0 < nt; i0 += 4 * rpp) {
        float4 xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * rpp;
            xv[u] = (i < nt) ? __ldg(xrow + (long long)i * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * rpp;
            if (i >= nt) break;
            const float y0 = ys[i], y1 = ys[CH + i], y2 = ys[2 * CH + i], y3 = ys[3 * CH + i];
            float4 r;
            r.x = fmaf(fmaf(mw[0].w, y3, fmaf(mw[0].z, y2, fmaf(mw[0].y, y1, fmaf(mw[0].x, y0, mb.x)))), xv[u].x, xv[u].x);
            r.y = fmaf(fmaf(mw[1].w, y3, fmaf(mw[1].z, y2, fmaf(mw[1].y, y1, fmaf(mw[1].x, y0, mb.y)))), xv[u].y, xv[u].y);
            r.z = fmaf(fmaf(mw[2].w, y3, fmaf(mw[2].z, y2, fmaf(mw[2].y, y1, fmaf(mw[2].x, y0, mb.z)))), xv[u].z, xv[u].z);
            r.w = fmaf(fmaf(mw[3].w, y3, fmaf(mw[3].z, y2, fmaf(mw[3].y, y1, fmaf(mw[3].x, y0, mb.w)))), xv[u].w, xv[u].w);
            OutT* o = obase + (long long)i * C;
            if (sizeof(OutT) == 4) {
                *reinterpret_cast<float4*>(o) = r;
            } else {
                const __nv_bfloat162 h01 = __floats2bfloat162_rn(r.x, r.y), h23 = __floats2bfloat162_rn(r.z, r.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&h01);
                pk.y = *reinterpret_cast<const uint32_t*>(&h23);
                *reinterpret_cast<uint2*>(o) = pk;
            }
        }
    }
}

// EnhanceBlock gating, streaming variant: the branch signals come from the compact (B, T, 4) array the stats pass wrote, so
// there is no per-tile pooling / convolution phase (strided channel-0 reads, five barriers) in front of the stream.  Warp 0
// reduces the clip's partial sums (fp64), then a thread owns one float4 of channels and walks the tile's rows, four in flight.
constexpr int kEnhStreamTile = 1024;

template <typename OutT>
__global__ void __launch_bounds__(256) enhance_apply_stream_kernel(const float* __restrict__ x, int B, int T, int C,
                                                                   const float* __restrict__ in_w, const float* __restrict__ in_b,
                                                                   const float* __restrict__ merge_w,
                                                                   const float* __restrict__ merge_b,
                                                                   const float* __restrict__ partials, int nchunk,
                                                                   const float4* __restrict__ branches, int tile,
                                                                   OutT* __restrict__ out) {
    __shared__ float s_scale[4], s_shift[4];
    const int b = blockIdx.y, t0 = blockIdx.x * tile;
    if (threadIdx.x < 32) {
        double acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0;
        for (int c = threadIdx.x; c < nchunk; c += 32) {
            const float4* pp = reinterpret_cast<const float4*>(partials + ((long long)b * nchunk + c) * 8);
            const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
            acc[0] += p0.x; acc[1] += p0.y; acc[2] += p0.z; acc[3] += p0.w;
            acc[4] += p1.x; acc[5] += p1.y; acc[6] += p1.z; acc[7] += p1.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (threadIdx.x < 4) {
            const int j = threadIdx.x;
            const double sum = j == 0 ? acc[0] : j == 1 ? acc[2] : j == 2 ? acc[4] : acc[6];
            const double sq = j == 0 ? acc[1] : j == 1 ? acc[3] : j == 2 ? acc[5] : acc[7];
            const double mean = sum / (double)T;
            double var = sq / (double)T - mean * mean;   // biased variance (InstanceNorm1d)
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + 1e-5));
            const float gg = in_w[j] * rstd;
            s_scale[j] = gg;
            s_shift[j] = in_b[j] - (float)mean * gg;
        }
    }
    __syncthreads();
    const int nt = min(tile, T - t0);
    const int C4 = C >> 2;
    const int rpp = blockDim.x / C4;                  // rows per pass
    const int c4 = threadIdx.x % C4, r0 = threadIdx.x / C4;
    if (r0 >= rpp) return;
    const float4 sc = make_float4(s_scale[0], s_scale[1], s_scale[2], s_scale[3]);
    const float4 sh = make_float4(s_shift[0], s_shift[1], s_shift[2], s_shift[3]);
    float4 mw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) mw[q] = __ldg(reinterpret_cast<const float4*>(merge_w) + 4 * c4 + q);   // merge_w[c][0..3]
    const float4 mb = __ldg(reinterpret_cast<const float4*>(merge_b) + c4);
    const float4* xrow = reinterpret_cast<const float4*>(x + ((long long)b * T + t0) * C) + c4;
    const float4* yrow = branches + (long long)b * T + t0;
    OutT* obase = out + ((long long)b * T + t0) * C + 4 * c4;
    for (int i0 = r0; i0 < nt; i0 += 4 * rpp) {
        float4 xv[4], yv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * rpp;
            const bool ok = i < nt;
            xv[u] = ok ? __ldg(xrow + (long long)i * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
            yv[u] = ok ? __ldg(yrow + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * rpp;
            if (i >= nt) break;
            const float y0 = fmaf(yv[u].x, sc.x, sh.x), y1 = fmaf(yv[u].y, sc.y, sh.y), y2 = fmaf(yv[u].z, sc.z, sh.z),
                        y3 = fmaf(yv[u].w, sc.w, sh.w);
            float4 r;
            r.x = fmaf(fmaf(mw[0].w, y3, fmaf(mw[0].z, y2, fmaf(mw[0].y, y1, fmaf(mw[0].x, y0, mb.x)))), xv[u].x, xv[u].x);
            r.y = fmaf(fmaf(mw[1].w, y3, fmaf(mw[1].z, y2, fmaf(mw[1].y, y1, fmaf(mw[1].x, y0, mb.y)))), xv[u].y, xv[u].y);
            r.z = fmaf(fmaf(mw[2].w, y3, fmaf(mw[2].z, y2, fmaf(mw[2].y, y1, fmaf(mw[2].x, y0, mb.z)))), xv[u].z, xv[u].z);
            r.w = fmaf(fmaf(mw[3].w, y3, fmaf(mw[3].z, y2, fmaf(mw[3].y, y1, fmaf(mw[3].x, y0, mb.w)))), xv[u].w, xv[u].w);
            OutT* o = obase + (long long)i * C;
            if (sizeof(OutT) == 4) {
                *reinterpret_cast<float4*>(o) = r;
            } else {
                const __nv_bfloat162 h01 = __floats2bfloat162_rn(r.x, r.y), h23 = __floats2bfloat162_rn(r.z, r.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&h01);
                pk.y = *reinterpret_cast<const uint32_t*>(&h23);
                *reinterpret_cast<uint2*>(o) = pk;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// decoder tail: snake -> conv(C->1, k7, pad 3) -> tanh.  Block = 256 output samples.
// ------------------------------------------------------------------------------------------
constexpr int kTailTile = 256;
constexpr int kTailMaxC = 32;

__global__ void __launch_bounds__(kTailTile) tail_kernel(const float* __restrict__ x, int B, int T, int C,
                                                         const float* __restrict__ alpha,
                                                         const float* __restrict__ w, float bias,
                                                         float* __restrict__ out) {
    __shared__ float s[(kTailTile + 6) * (kTailMaxC + 1)];
    __shared__ float s_w[7 * kTailMaxC];
    const int b = blockIdx.y, t0 = blockIdx.x * kTailTile;
    const int ld = C + 1;
    for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) s_w[i] = w[i];
    const float* xb = x + (long long)b * T * C;
    for (int e = threadIdx.x; e < (kTailTile + 6) * C; e += blockDim.x) {
        const int i = e / C, c = e - i * C;
        const int t = t0 - 3 + i;
        float v = 0.f;
        if (t >= 0 && t < T) {
            const float a = __ldg(alpha + c);
            v = snake_f(__ldg(xb + (long long)t * C + c), a, 1.0f / (a + kEps));
        }
        s[i * ld + c] = v;
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t < T) {
        float acc = bias;
        for (int j = 0; j < 7; ++j) {
            const float* sr = s + (threadIdx.x + j) * ld;
            for (int c = 0; c < C; ++c) acc = fmaf(s_w[j * C + c], sr[c], acc);
        }
        out[(long long)b * T + t] = tanhf(acc);
    }
}

static inline int grid_for_rows(long long rows, int rows_per_block) {
    long long g = (rows + rows_per_block - 1) / rows_per_block;
    const long long cap = 148LL * 32;   // persistent-style cap: 32 blocks per SM worth of waves
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// grid.x of a (row blocks, clip) grid: the same ~32 blocks per SM in total as grid_for_rows (each block loops)
static inline int ups_grid_x(long long rows_per_clip, int rows_per_block, int B) {
    long long g = (rows_per_clip + rows_per_block - 1) / rows_per_block;
    const long long cap = (148LL * 32 + B - 1) / B;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace l3ac

using namespace l3ac;

#define DISPATCH_CPL(C, ...)                                        \
    do {                                                            \
        const int cpl_ = ((C) + 31) / 32;                           \
        if (cpl_ <= 1) { constexpr int CPL = 1; __VA_ARGS__; }      \
        else if (cpl_ <= 2) { constexpr int CPL = 2; __VA_ARGS__; } \
        else if (cpl_ <= 3) { constexpr int CPL = 3; __VA_ARGS__; } \
        else if (cpl_ <= 4) { constexpr int CPL = 4; __VA_ARGS__; } \
        else if (cpl_ <= 6) { constexpr int CPL = 6; __VA_ARGS__; } \
        else if (cpl_ <= 8) { constexpr int CPL = 8; __VA_ARGS__; } \
        else { constexpr int CPL = 16; __VA_ARGS__; }               \
    } while (0)

extern "C" int l3ac_dwconv7_ln(const float* x, int B, int T, int C, const float* dw_w, const float* dw_b,
                               const float* ln_w, const float* ln_b, float eps, void* out, void* out_lo,
                               int out_dtype, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(x && dw_w && dw_b && ln_w && ln_b && out);
    L3AC_CHECK_ARG(B > 0 && T > 0 && C > 0 && C <= 512);
    L3AC_CHECK_ARG(out_dtype == L3AC_F32 || out_dtype == L3AC_BF16 || out_dtype == L3AC_BF16X2);
    L3AC_CHECK_ARG((out_dtype == L3AC_BF16X2) == (out_lo != nullptr));
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)B * T;
    if (C % 4 == 0 && C <= 128 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
        ((reinterpret_cast<uintptr_t>(dw_w) | reinterpret_cast<uintptr_t>(dw_b) | reinterpret_cast<uintptr_t>(ln_w) |
          reinterpret_cast<uintptr_t>(ln_b) | reinterpret_cast<uintptr_t>(out_lo)) & 15) == 0) {
        constexpr int R = 8;
        const long long runs = (long long)B * ((T + R - 1) / R);
#define L3AC_DWV(GS)                                                                                                   \
    do {                                                                                                               \
        const int g_ = grid_for_rows(runs, 8 * (32 / GS));                                                             \
        if (out_dtype == L3AC_F32)                                                                                     \
            dwconv7_ln_vec_kernel<GS, R, float><<<g_, 256, 0, st>>>(x, B, T, C, dw_w, dw_b, ln_w, ln_b, eps, (float*)out, \
                                                                    nullptr);                                          \
        else                                                                                                           \
            dwconv7_ln_vec_kernel<GS, R, __nv_bfloat16><<<g_, 256, 0, st>>>(x, B, T, C, dw_w, dw_b, ln_w, ln_b, eps,     \
                                                                            (__nv_bfloat16*)out, (__nv_bfloat16*)out_lo); \
    } while (0)
        if (C <= 32) L3AC_DWV(8);
        else if (C <= 64) L3AC_DWV(16);
        else L3AC_DWV(32);
#undef L3AC_DWV
        return l3ac_launch_status();
    }
    if (C <= 96) {
        constexpr int R = 8;
        const int grid = grid_for_rows((rows + R - 1) / R + B, 8);
#define L3AC_ROWS_LAUNCH(CPLV)                                                                                          \
    do {                                                                                                               \
        if (out_dtype == L3AC_F32)                                                                                     \
            dwconv7_ln_rows_kernel<CPLV, R, float><<<grid, 256, 0, st>>>(x, B, T, C, dw_w, dw_b, ln_w, ln_b, eps,       \
                                                                         (float*)out, nullptr);                        \
        else                                                                                                           \
            dwconv7_ln_rows_kernel<CPLV, R, __nv_bfloat16><<<grid, 256, 0, st>>>(                                      \
                x, B, T, C, dw_w, dw_b, ln_w, ln_b, eps, (__nv_bfloat16*)out, (__nv_bfloat16*)out_lo);                 \
    } while (0)
        if (C <= 32) L3AC_ROWS_LAUNCH(1);
        else if (C <= 64) L3AC_ROWS_LAUNCH(2);
        else L3AC_ROWS_LAUNCH(3);
#undef L3AC_ROWS_LAUNCH
        return l3ac_launch_status();
    }
    if (C % 128 == 0 && C <= 512 && B <= 65535 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
        ((reinterpret_cast<uintptr_t>(dw_w) | reinterpret_cast<uintptr_t>(dw_b) | reinterpret_cast<uintptr_t>(ln_w) |
          reinterpret_cast<uintptr_t>(ln_b) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out_lo)) & 15) == 0) {
        constexpr int R = 8;
        dim3 wgrid(l3ac_cdiv(T, R), B);
#define L3AC_DWLN_WIDE(NW)                                                                                              \
    do {                                                                                                                \
        if (out_dtype == L3AC_F32)                                                                                      \
            dwconv7_ln_wide_kernel<R, NW, float><<<wgrid, 32 * NW, 0, st>>>(x, B, T, C, dw_w, dw_b, ln_w, ln_b, eps, (float*)out, \
                                                                          nullptr);                                     \
        else                                                                                                            \
            dwconv7_ln_wide_kernel<R, NW, __nv_bfloat16><<<wgrid, 32 * NW, 0, st>>>(                                    \
                x, B, T, C, dw_w, dw_b, ln_w, ln_b, eps, (__nv_bfloat16*)out, (__nv_bfloat16*)out_lo);                  \
    } while (0)
        switch (C / 128) {
            case 1: L3AC_DWLN_WIDE(1); break;
            case 2: L3AC_DWLN_WIDE(2); break;
            case 3: L3AC_DWLN_WIDE(3); break;
            default: L3AC_DWLN_WIDE(4); break;
        }
#undef L3AC_DWLN_WIDE
        return l3ac_launch_status();
    }
    if (C % 32 == 0 && C >= 128 && C <= 512 && B <= 65535) {
        constexpr int TT = 32;
        dim3 tgrid(l3ac_cdiv(T, TT), B);
        if (out_dtype == L3AC_F32)
            dwconv7_ln_tile_kernel<TT, float><<<tgrid, C, 0, st>>>(x, B, T, C, dw_w, dw_b, ln_w, ln_b, eps, (float*)out, nullptr);
        else
            dwconv7_ln_tile_kernel<TT, __nv_bfloat16><<<tgrid, C, 0, st>>>(x, B, T, C, dw_w, dw_b, ln_w, ln_b, eps,
                                                                           (__nv_bfloat16*)out, (__nv_bfloat16*)out_lo);
        return l3ac_launch_status();
    }
    const int grid = grid_for_rows(rows, 8);
    const size_t smem = (size_t)10 * C * sizeof(float);
    DISPATCH_CPL(C, {
        if (out_dtype == L3AC_F32)
            dwconv7_ln_kernel<CPL, float><<<grid, 256, smem, st>>>(x, B, T, C, dw_w, dw_b, ln_w, ln_b, eps, (float*)out,
                                                                   nullptr);
        else
            dwconv7_ln_kernel<CPL, __nv_bfloat16><<<grid, 256, smem, st>>>(x, B, T, C, dw_w, dw_b, ln_w, ln_b, eps,
                                                                           (__nv_bfloat16*)out, (__nv_bfloat16*)out_lo);
    });
    return l3ac_launch_status();
}

extern "C" int l3ac_layernorm(const float* x, long long M, int C, const float* w, const float* b, float eps,
                              void* out, void* out_lo, int out_dtype, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(x && w && b && out);
    L3AC_CHECK_ARG(M > 0 && C > 0 && C <= 512);
    L3AC_CHECK_ARG(out_dtype == L3AC_F32 || out_dtype == L3AC_BF16 || out_dtype == L3AC_BF16X2);
    L3AC_CHECK_ARG((out_dtype == L3AC_BF16X2) == (out_lo != nullptr));
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
        ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) {
        const int C4 = C / 4;
#define L3AC_LN(GS, VPL, U)                                                                                        \
    do {                                                                                                           \
        const int g_ = grid_for_rows(M, 8 * (32 / GS) * U);                                                        \
        if (out_dtype == L3AC_F32)                                                                                 \
            layernorm_vec_kernel<GS, VPL, U, float><<<g_, 256, 0, st>>>(x, M, C, w, b, eps, (float*)out, nullptr);   \
        else                                                                                                       \
            layernorm_vec_kernel<GS, VPL, U, __nv_bfloat16><<<g_, 256, 0, st>>>(x, M, C, w, b, eps, (__nv_bfloat16*)out, \
                                                                                (__nv_bfloat16*)out_lo);             \
    } while (0)
        if (C4 <= 8) L3AC_LN(8, 1, 4);
        else if (C4 <= 16) L3AC_LN(16, 1, 4);
        else if (C4 <= 32) L3AC_LN(32, 1, 4);
        else if (C4 <= 64) L3AC_LN(32, 2, 2);
        else L3AC_LN(32, 4, 1);
#undef L3AC_LN
        return l3ac_launch_status();
    }
    const int grid = grid_for_rows(M, 8);
    DISPATCH_CPL(C, {
        if (out_dtype == L3AC_F32)
            layernorm_kernel<CPL, float><<<grid, 256, 0, st>>>(x, M, C, w, b, eps, (float*)out, nullptr);
        else
            layernorm_kernel<CPL, __nv_bfloat16><<<grid, 256, 0, st>>>(x, M, C, w, b, eps, (__nv_bfloat16*)out,
                                                                       (__nv_bfloat16*)out_lo);
    });
    return l3ac_launch_status();
}

extern "C" int l3ac_split_bf16(const float* x, long long n, void* hi, void* lo, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(x && hi && lo && n > 0 && n % 4 == 0);
    L3AC_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(lo) & 7) == 0);
    split_bf16_kernel<<<grid_for_rows(n / 4, 256 * 2), 256, 0, (cudaStream_t)stream>>>(x, n / 4, (__nv_bfloat16*)hi,
                                                                                      (__nv_bfloat16*)lo);
    return l3ac_launch_status();
}

extern "C" int l3ac_snake(const float* x, long long M, int C, const float* alpha, void* out, int out_dtype,
                          l3ac_stream_t stream) {
    L3AC_CHECK_ARG(x && alpha && out && M > 0 && C > 0);
    L3AC_CHECK_ARG(out_dtype == L3AC_F32 || out_dtype == L3AC_BF16);
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = M * C;
    const int grid = grid_for_rows(n, 256 * 4);
    if (out_dtype == L3AC_F32)
        snake_kernel<float><<<grid, 256, 0, st>>>(x, n, C, alpha, (float*)out);
    else
        snake_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, n, C, alpha, (__nv_bfloat16*)out);
    return l3ac_launch_status();
}

extern "C" int l3ac_upsample_linear_cn(const float* x, int B, int T, int C, int scale, const float* cn_w,
                                       const float* cn_b, float eps, float* out, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(x && out && B > 0 && B <= 65535 && T > 0 && C > 0 && C <= 512 && scale >= 1 && (long long)T * scale <= 0x7fffffffLL);
    L3AC_CHECK_ARG((cn_w == nullptr) == (cn_b == nullptr));
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)B * T * scale;
    if (C % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
        (!cn_w || ((reinterpret_cast<uintptr_t>(cn_w) | reinterpret_cast<uintptr_t>(cn_b)) & 15) == 0)) {
        const int C4 = C / 4;
        if (C4 <= 32 && scale >= 2 && scale <= 5) {        // run variant: input rows loaded once per run of outputs
#define L3AC_UPR(GS, SV)                                                                                                 \
    upsample_cn_run_kernel<GS, SV><<<dim3(ups_grid_x(((long long)T * SV + 4) / 5, 8 * (32 / GS), B), B), 256, 0, st>>>(   \
        x, B, T, C, cn_w, cn_b, eps, out)
#define L3AC_UPR_S(GS)                        \
    do {                                      \
        if (scale == 2) L3AC_UPR(GS, 2);      \
        else if (scale == 3) L3AC_UPR(GS, 3); \
        else if (scale == 4) L3AC_UPR(GS, 4); \
        else L3AC_UPR(GS, 5);                 \
    } while (0)
            if (C4 <= 8) L3AC_UPR_S(8);
            else if (C4 <= 16) L3AC_UPR_S(16);
            else L3AC_UPR_S(32);
#undef L3AC_UPR_S
#undef L3AC_UPR
            return l3ac_launch_status();
        }
#define L3AC_UPS(GS, VPL, U)                                                                                      \
    upsample_cn_vec_kernel<GS, VPL, U><<<dim3(ups_grid_x((long long)T * scale, 8 * (32 / GS) * U, B), B), 256, 0, st>>>( \
        x, B, T, C, scale, cn_w, cn_b, eps, out)
        if (C4 <= 8) L3AC_UPS(8, 1, 4);
        else if (C4 <= 16) L3AC_UPS(16, 1, 4);
        else if (C4 <= 32) L3AC_UPS(32, 1, 4);
        else if (C4 <= 64) L3AC_UPS(32, 2, 2);
        else L3AC_UPS(32, 4, 1);
#undef L3AC_UPS
        return l3ac_launch_status();
    }
    const int grid = grid_for_rows(rows, 8);
    DISPATCH_CPL(C, { upsample_cn_kernel<CPL><<<grid, 256, 0, st>>>(x, B, T, C, scale, cn_w, cn_b, eps, out); });
    return l3ac_launch_status();
}

extern "C" long long l3ac_enhance_partials_floats(int B, int T) {
    return (long long)B * l3ac_cdiv(T, kEnhChunk) * 8;
}

extern "C" int l3ac_enhance_stats(const float* x, int B, int T, int C, const float* conv_w, const float* conv_b,
                                  float* partials, float* branches, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(x && conv_w && conv_b && partials && B > 0 && B <= 65535 && T > 0 && C > 0);
    L3AC_CHECK_ARG((reinterpret_cast<uintptr_t>(branches) & 15) == 0);
    dim3 grid(l3ac_cdiv(T, kEnhChunk), B);
    enhance_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, B, T, C, conv_w, conv_b, partials,
                                                                 reinterpret_cast<float4*>(branches));
    return l3ac_launch_status();
}

extern "C" int l3ac_enhance_apply(const float* x, int B, int T, int C, const float* conv_w, const float* conv_b,
                                  const float* in_w, const float* in_b, const float* merge_w,
                                  const float* merge_b, const float* partials, const float* branches, void* out,
                                  int out_dtype, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(x && conv_w && conv_b && in_w && in_b && merge_w && merge_b && partials && out);
    L3AC_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && C > 0);
    L3AC_CHECK_ARG(out_dtype == L3AC_F32 || out_dtype == L3AC_BF16);
    L3AC_CHECK_ARG((reinterpret_cast<uintptr_t>(merge_w) & 15) == 0);
    dim3 grid(l3ac_cdiv(T, kEnhTile), B);
    const int nchunk = l3ac_cdiv(T, kEnhChunk);
    cudaStream_t st = (cudaStream_t)stream;
    if (branches && C % 4 == 0 && C / 4 <= 256 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                                                    reinterpret_cast<uintptr_t>(merge_b) | reinterpret_cast<uintptr_t>(branches) |
                                                    reinterpret_cast<uintptr_t>(partials)) & 15) == 0) {
        // streaming variant; short clips take 128-row tiles so that every SM gets a CTA
        const int tile = ((long long)l3ac_cdiv(T, kEnhStreamTile) * B >= 2 * 148) ? kEnhStreamTile : 128;
        dim3 g1(l3ac_cdiv(T, tile), B);
        if (out_dtype == L3AC_F32)
            enhance_apply_stream_kernel<float><<<g1, 256, 0, st>>>(x, B, T, C, in_w, in_b, merge_w, merge_b, partials, nchunk,
                                                                   reinterpret_cast<const float4*>(branches), tile, (float*)out);
        else
            enhance_apply_stream_kernel<__nv_bfloat16><<<g1, 256, 0, st>>>(x, B, T, C, in_w, in_b, merge_w, merge_b, partials, nchunk,
                                                                           reinterpret_cast<const float4*>(branches), tile,
                                                                           (__nv_bfloat16*)out);
        return l3ac_launch_status();
    }
    if (C % 4 == 0 && C / 4 <= 256 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(merge_b) & 15) == 0) {
        // 512-step tiles amortise the branch signals; short clips at the coarse decoder stages (T = 1779: 4 tiles per clip)
        // would leave most SMs without a CTA, so they take 128-step tiles.
        const bool small = (long long)grid.x * B < 2 * 148;
        dim3 grid_s(l3ac_cdiv(T, 128), B);
#define L3AC_ENH(OUTT, CHV, GRID)                                                                                       \
    enhance_apply_vec_kernel<OUTT, CHV><<<GRID, 256, 0, st>>>(x, B, T, C, conv_w, conv_b, in_w, in_b, merge_w, merge_b, \
                                                              partials, nchunk, (OUTT*)out)
        if (out_dtype == L3AC_F32) {
            if (small) L3AC_ENH(float, 128, grid_s);
            else L3AC_ENH(float, kEnhTile, grid);
        } else {
            if (small) L3AC_ENH(__nv_bfloat16, 128, grid_s);
            else L3AC_ENH(__nv_bfloat16, kEnhTile, grid);
        }
#undef L3AC_ENH
        return l3ac_launch_status();
    }
    if (out_dtype == L3AC_F32)
        enhance_apply_kernel<float><<<grid, 256, 0, st>>>(x, B, T, C, conv_w, conv_b, in_w, in_b, merge_w, merge_b,
                                                          partials, nchunk, (float*)out);
    else
        enhance_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, B, T, C, conv_w, conv_b, in_w, in_b, merge_w,
                                                                  merge_b, partials, nchunk, (__nv_bfloat16*)out);
    return l3ac_launch_status();
}

extern "C" int l3ac_tail_conv_tanh(const float* x, int B, int T, int C, const float* alpha, const float* w,
                                   float bias, float* out, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(x && alpha && w && out && B > 0 && B <= 65535 && T > 0 && C > 0 && C <= kTailMaxC);
    dim3 grid(l3ac_cdiv(T, kTailTile), B);
    tail_kernel<<<grid, kTailTile, 0, (cudaStream_t)stream>>>(x, B, T, C, alpha, w, bias, out);
    return l3ac_launch_status();
}

// ------------------------------------------------------------------------------------------
// dwconv7 + LayerNorm through a plan (C = 48 / 96, bf16 out): the per-channel parameters are kept on the HOST in the plan and
// travel as kernel parameters, so the thread-per-row kernel reads them from the constant bank.
// ------------------------------------------------------------------------------------------
struct l3ac_dwconv_plan {
    int C;
    l3ac::DwRowParams<48> p48;
    l3ac::DwRowParams<96> p96;
};

template <int C>
static void fill_dw_params(l3ac::DwRowParams<C>& p, const float* dw_w, const float* dw_b, const float* ln_w, const float* ln_b, float eps) {
    p.eps = eps;
    for (int j = 0; j < 7; ++j)
        for (int c = 0; c < C; ++c) p.w[j][c] = dw_w[j * C + c];
    for (int c = 0; c < C; ++c) {
        p.b[c] = dw_b[c];
        p.lw[c] = ln_w[c];
        p.lb[c] = ln_b[c];
    }
}

extern "C" int l3ac_dwconv_plan_create(int C, const float* dw_w, const float* dw_b, const float* ln_w, const float* ln_b, float eps,
                                       l3ac_dwconv_plan** plan_out) {
    L3AC_CHECK_ARG(dw_w && dw_b && ln_w && ln_b && plan_out);
    if (C != 48 && C != 96) return L3AC_EUNSUPPORTED;
    l3ac_dwconv_plan* plan = new (std::nothrow) l3ac_dwconv_plan();
    if (!plan) return L3AC_EINVAL;
    plan->C = C;
    if (C == 48) fill_dw_params(plan->p48, dw_w, dw_b, ln_w, ln_b, eps);
    else fill_dw_params(plan->p96, dw_w, dw_b, ln_w, ln_b, eps);
    *plan_out = plan;
    return L3AC_OK;
}

extern "C" int l3ac_dwconv_plan_destroy(l3ac_dwconv_plan* plan) {
    delete plan;
    return L3AC_OK;
}

template <int C>
static int launch_dw_thread(l3ac::DwRowParams<C> p, const float* x, int B, int T, void* out, cudaStream_t st) {
    using namespace l3ac;
    p.x = x;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.B = B;
    p.T = T;
    constexpr int smem = (128 + 6) * (C + 4) * 4;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(dwconv7_ln_thread_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    dwconv7_ln_thread_kernel<C><<<dim3(l3ac_cdiv(T, 128), B), 128, smem, st>>>(p);
    return l3ac_launch_status();
}

extern "C" int l3ac_dwconv7_ln_plan(const l3ac_dwconv_plan* plan, const float* x, int B, int T, void* out, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(plan && x && out && B > 0 && B <= 65535 && T > 0);
    L3AC_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    if (plan->C == 48) return launch_dw_thread<48>(plan->p48, x, B, T, out, (cudaStream_t)stream);
    return launch_dw_thread<96>(plan->p96, x, B, T, out, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// Upsample + ChannelNorm + (next unit's) dwconv7 + LayerNorm through a plan (C = 48 / 96, scale 2 / 3, bf16 operand out).
// ------------------------------------------------------------------------------------------
struct l3ac_updw_plan {
    int C, S;
    l3ac::UpDwParams<48> p48;
    l3ac::UpDwParams<96> p96;
};

template <int C>
static void fill_updw_params(l3ac::UpDwParams<C>& p, int S, const float* cn_w, const float* cn_b, float cn_eps, const float* dw_w,
                             const float* dw_b, const float* ln_w, const float* ln_b, float ln_eps) {
    p.S = S;
    p.cn_eps = cn_eps;
    p.ln_eps = ln_eps;
    for (int j = 0; j < 7; ++j)
        for (int c = 0; c < C; ++c) p.w[j][c] = dw_w[j * C + c];
    for (int c = 0; c < C; ++c) {
        p.cw[c] = cn_w[c];
        p.cb[c] = cn_b[c];
        p.b[c] = dw_b[c];
        p.lw[c] = ln_w[c];
        p.lb[c] = ln_b[c];
    }
}

extern "C" int l3ac_updw_plan_create(int C, int scale, const float* cn_w, const float* cn_b, float cn_eps, const float* dw_w,
                                     const float* dw_b, const float* ln_w, const float* ln_b, float ln_eps, l3ac_updw_plan** plan_out) {
    L3AC_CHECK_ARG(cn_w && cn_b && dw_w && dw_b && ln_w && ln_b && plan_out);
    if ((C != 48 && C != 96) || (scale != 2 && scale != 3)) return L3AC_EUNSUPPORTED;
    l3ac_updw_plan* plan = new (std::nothrow) l3ac_updw_plan();
    if (!plan) return L3AC_EINVAL;
    plan->C = C;
    plan->S = scale;
    if (C == 48) fill_updw_params(plan->p48, scale, cn_w, cn_b, cn_eps, dw_w, dw_b, ln_w, ln_b, ln_eps);
    else fill_updw_params(plan->p96, scale, cn_w, cn_b, cn_eps, dw_w, dw_b, ln_w, ln_b, ln_eps);
    *plan_out = plan;
    return L3AC_OK;
}

extern "C" int l3ac_updw_plan_destroy(l3ac_updw_plan* plan) {
    delete plan;
    return L3AC_OK;
}

template <int C, int S>
static int launch_updw(l3ac::UpDwParams<C> p, const float* y, int B, int T, float* xup, void* a, cudaStream_t st) {
    using namespace l3ac;
    p.y = y;
    p.xup = xup;
    p.a = static_cast<__nv_bfloat16*>(a);
    p.B = B;
    p.T = T;
    constexpr int XR = 128 + 6, YR = (XR + S - 1) / S + 2;
    constexpr int smem = (XR + YR) * (C + 4) * 4;
    cudaError_t e = cudaFuncSetAttribute(upsample_cn_dwconv7_ln_kernel<C, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    upsample_cn_dwconv7_ln_kernel<C, S><<<dim3(l3ac_cdiv((long long)T * S, 128), B), kUpDwThreads, smem, st>>>(p);
    return l3ac_launch_status();
}

extern "C" int l3ac_upsample_cn_dwconv7_ln(const l3ac_updw_plan* plan, const float* y, int B, int T, float* x_up, void* a_out,
                                           l3ac_stream_t stream) {
    L3AC_CHECK_ARG(plan && y && x_up && a_out && B > 0 && B <= 65535 && T > 0 && (long long)T * plan->S < (1LL << 30));
    L3AC_CHECK_ARG(((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x_up) | reinterpret_cast<uintptr_t>(a_out)) & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->C == 48) return plan->S == 2 ? launch_updw<48, 2>(plan->p48, y, B, T, x_up, a_out, st) : launch_updw<48, 3>(plan->p48, y, B, T, x_up, a_out, st);
    return plan->S == 2 ? launch_updw<96, 2>(plan->p96, y, B, T, x_up, a_out, st) : launch_updw<96, 3>(plan->p96, y, B, T, x_up, a_out, st);
}

// ------------------------------------------------------------------------------------------
// EnhanceBlock gate + the up layer's 1x1 conv in one kernel (decode side, bf16 operands; (C_in, C_out) = (48, 24) / (96, 48)):
//   a = bf16(x + y(t, c) * x),  y = merge(InstanceNorm(branches))      (l3ac/tconv/__init__.py:40-44)
//   out = a . W^T + b                                                   (l3ac/modules.py:161)
// These two stages are HBM-bound row kernels around a tiny contraction (K = 48 / 96, N = 24 / 48: 0.15 % of the model's FLOPs):
// the gated activation goes straight into mma.sync A fragments in registers -- every lane loads exactly the (row, k) pairs of
// its fragment, 8 bytes at a time, a quad covering one 32-byte sector per row -- so the bf16 activation tensor is never
// written or read; W sits in shared memory at a (K + 8)-element pitch (conflict-free fragment reads).
// ------------------------------------------------------------------------------------------
namespace l3ac {

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kEnhUpTile = 2048;       // rows of one clip per block

template <int CI, int CO>
struct EnhUpBlob {                     // device blob, copied to shared memory by every block
    __nv_bfloat16 w[CO][CI + 8];       // up conv weight, bf16
    float4 mw[CI];                     // merge_layer.1 weight [c][0..3]
    float mb[CI];                      // merge_layer.1 bias
    float bias[CO];                    // up conv bias
    float in_w[4], in_b[4];            // InstanceNorm affine
};

template <int CI, int CO>
__global__ void __launch_bounds__(256) enhance_up_kernel(const float* __restrict__ x, int B, int T, const float* __restrict__ partials,
                                                         int nchunk, const float4* __restrict__ branches,
                                                         const EnhUpBlob<CI, CO>* __restrict__ blob, float* __restrict__ out) {
    __shared__ __align__(16) EnhUpBlob<CI, CO> sb;
    __shared__ float s_scale[4], s_shift[4];
    const int b = blockIdx.y, t0 = blockIdx.x * kEnhUpTile;
    {
        const uint4* src = reinterpret_cast<const uint4*>(blob);
        uint4* dst = reinterpret_cast<uint4*>(&sb);
        for (int i = threadIdx.x; i < (int)(sizeof(EnhUpBlob<CI, CO>) / 16); i += 256) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if (threadIdx.x < 32) {            // InstanceNorm statistics of the clip from the stats pass' partial sums (fp64)
        double acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0;
        for (int c = threadIdx.x; c < nchunk; c += 32) {
            const float4* pp = reinterpret_cast<const float4*>(partials + ((long long)b * nchunk + c) * 8);
            const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
            acc[0] += p0.x; acc[1] += p0.y; acc[2] += p0.z; acc[3] += p0.w;
            acc[4] += p1.x; acc[5] += p1.y; acc[6] += p1.z; acc[7] += p1.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (threadIdx.x < 4) {
            const int j = threadIdx.x;
            const double sum = j == 0 ? acc[0] : j == 1 ? acc[2] : j == 2 ? acc[4] : acc[6];
            const double sq = j == 0 ? acc[1] : j == 1 ? acc[3] : j == 2 ? acc[5] : acc[7];
            const double mean = sum / (double)T;
            double var = sq / (double)T - mean * mean;   // biased variance (InstanceNorm1d)
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + 1e-5));
            const float gg = sb.in_w[j] * rstd;
            s_scale[j] = gg;
            s_shift[j] = sb.in_b[j] - (float)mean * gg;
        }
    }
    __syncthreads();
    const int nt = min(kEnhUpTile, T - t0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    const float4 sc = make_float4(s_scale[0], s_scale[1], s_scale[2], s_scale[3]);
    const float4 sh = make_float4(s_shift[0], s_shift[1], s_shift[2], s_shift[3]);
    const float* xb = x + ((long long)b * T + t0) * CI;
    const float4* yb = branches + (long long)b * T + t0;
    float* ob = out + ((long long)b * T + t0) * CO;
    constexpr int KK = CI / 16, NT = CO / 8;
    for (int m0 = warp * 16; m0 < nt; m0 += 8 * 16) {
        const int r_lo = m0 + g, r_hi = r_lo + 8;
        const int l_lo = r_lo < nt ? r_lo : nt - 1, l_hi = r_hi < nt ? r_hi : nt - 1;       // (clamped loads, masked stores)
        float2 xv[KK][4];
#pragma unroll
        for (int kk = 0; kk < KK; ++kk)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = 16 * kk + 8 * h + 2 * tig;
                xv[kk][2 * h] = __ldg(reinterpret_cast<const float2*>(xb + (long long)l_lo * CI + c));
                xv[kk][2 * h + 1] = __ldg(reinterpret_cast<const float2*>(xb + (long long)l_hi * CI + c));
            }
        const float4 b_lo = __ldg(yb + l_lo), b_hi = __ldg(yb + l_hi);
        const float yl0 = fmaf(b_lo.x, sc.x, sh.x), yl1 = fmaf(b_lo.y, sc.y, sh.y), yl2 = fmaf(b_lo.z, sc.z, sh.z), yl3 = fmaf(b_lo.w, sc.w, sh.w);
        const float yh0 = fmaf(b_hi.x, sc.x, sh.x), yh1 = fmaf(b_hi.y, sc.y, sh.y), yh2 = fmaf(b_hi.z, sc.z, sh.z), yh3 = fmaf(b_hi.w, sc.w, sh.w);
        uint32_t a[KK][4];
#pragma unroll
        for (int kk = 0; kk < KK; ++kk)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = 16 * kk + 8 * h + 2 * tig;
                const float4 m0w = sb.mw[c], m1w = sb.mw[c + 1];
                const float mb0 = sb.mb[c], mb1 = sb.mb[c + 1];
                // gate(t, c) = mb[c] + sum_k mw[c][k] y_k(t), nested like enhance_apply_stream_kernel; r = x + gate * x
                const float gl0 = fmaf(m0w.w, yl3, fmaf(m0w.z, yl2, fmaf(m0w.y, yl1, fmaf(m0w.x, yl0, mb0))));
                const float gl1 = fmaf(m1w.w, yl3, fmaf(m1w.z, yl2, fmaf(m1w.y, yl1, fmaf(m1w.x, yl0, mb1))));
                const float gh0 = fmaf(m0w.w, yh3, fmaf(m0w.z, yh2, fmaf(m0w.y, yh1, fmaf(m0w.x, yh0, mb0))));
                const float gh1 = fmaf(m1w.w, yh3, fmaf(m1w.z, yh2, fmaf(m1w.y, yh1, fmaf(m1w.x, yh0, mb1))));
                const float2 xl = xv[kk][2 * h], xh = xv[kk][2 * h + 1];
                const __nv_bfloat162 pl = __floats2bfloat162_rn(fmaf(gl0, xl.x, xl.x), fmaf(gl1, xl.y, xl.y));
                const __nv_bfloat162 ph = __floats2bfloat162_rn(fmaf(gh0, xh.x, xh.x), fmaf(gh1, xh.y, xh.y));
                a[kk][2 * h] = *reinterpret_cast<const uint32_t*>(&pl);          // a0 / a2: row g
                a[kk][2 * h + 1] = *reinterpret_cast<const uint32_t*>(&ph);      // a1 / a3: row g + 8
            }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sb.w[8 * n + g][16 * kk + 2 * tig]);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sb.w[8 * n + g][16 * kk + 8 + 2 * tig]);
                mma_bf16_16816(d, a[kk], b0, b1);
            }
            const int col = 8 * n + 2 * tig;
            const float bias0 = sb.bias[col], bias1 = sb.bias[col + 1];
            if (r_lo < nt) *reinterpret_cast<float2*>(ob + (long long)r_lo * CO + col) = make_float2(d[0] + bias0, d[1] + bias1);
            if (r_hi < nt) *reinterpret_cast<float2*>(ob + (long long)r_hi * CO + col) = make_float2(d[2] + bias0, d[3] + bias1);
        }
    }
}

}  // namespace l3ac

struct l3ac_enhup_plan {
    int C_in, C_out, device;
    void* dev_blob;
};

template <int CI, int CO>
static int make_enhup_blob(l3ac_enhup_plan* plan, const float* in_w, const float* in_b, const float* merge_w, const float* merge_b,
                           const float* up_w, const float* up_b) {
    using Blob = l3ac::EnhUpBlob<CI, CO>;
    Blob* h = new (std::nothrow) Blob();
    if (!h) return L3AC_EINVAL;
    memset(h, 0, sizeof(Blob));
    for (int n = 0; n < CO; ++n)
        for (int k = 0; k < CI; ++k) h->w[n][k] = __float2bfloat16_rn(up_w[(size_t)n * CI + k]);
    for (int c = 0; c < CI; ++c) {
        h->mw[c] = make_float4(merge_w[4 * c], merge_w[4 * c + 1], merge_w[4 * c + 2], merge_w[4 * c + 3]);
        h->mb[c] = merge_b[c];
    }
    for (int n = 0; n < CO; ++n) h->bias[n] = up_b[n];
    for (int j = 0; j < 4; ++j) {
        h->in_w[j] = in_w[j];
        h->in_b[j] = in_b[j];
    }
    cudaError_t e = cudaMalloc(&plan->dev_blob, sizeof(Blob));
    if (e == cudaSuccess) e = cudaMemcpy(plan->dev_blob, h, sizeof(Blob), cudaMemcpyHostToDevice);
    delete h;
    return e == cudaSuccess ? L3AC_OK : (int)e;
}

extern "C" int l3ac_enhup_plan_create(int C_in, int C_out, const float* in_w, const float* in_b, const float* merge_w, const float* merge_b,
                                      const float* up_w, const float* up_b, l3ac_enhup_plan** plan_out) {
    L3AC_CHECK_ARG(in_w && in_b && merge_w && merge_b && up_w && up_b && plan_out);
    if (!((C_in == 48 && C_out == 24) || (C_in == 96 && C_out == 48))) return L3AC_EUNSUPPORTED;
    l3ac_enhup_plan* plan = new (std::nothrow) l3ac_enhup_plan();
    if (!plan) return L3AC_EINVAL;
    plan->C_in = C_in;
    plan->C_out = C_out;
    plan->dev_blob = nullptr;
    if (cudaGetDevice(&plan->device) != cudaSuccess) { delete plan; return L3AC_EDRIVER; }
    const int rc = C_in == 48 ? make_enhup_blob<48, 24>(plan, in_w, in_b, merge_w, merge_b, up_w, up_b)
                              : make_enhup_blob<96, 48>(plan, in_w, in_b, merge_w, merge_b, up_w, up_b);
    if (rc != L3AC_OK) {
        cudaFree(plan->dev_blob);
        delete plan;
        return rc;
    }
    *plan_out = plan;
    return L3AC_OK;
}

extern "C" int l3ac_enhup_plan_destroy(l3ac_enhup_plan* plan) {
    if (!plan) return L3AC_OK;
    cudaFree(plan->dev_blob);
    delete plan;
    return L3AC_OK;
}

extern "C" int l3ac_enhance_up(const l3ac_enhup_plan* plan, const float* x, int B, int T, const float* partials, const float* branches,
                               float* out, l3ac_stream_t stream) {
    using namespace l3ac;
    L3AC_CHECK_ARG(plan && x && partials && branches && out && B > 0 && B <= 65535 && T > 0);
    L3AC_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(partials) | reinterpret_cast<uintptr_t>(branches) |
                     reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return L3AC_EDRIVER;
    L3AC_CHECK_ARG(dev == plan->device);
    const int nchunk = l3ac_cdiv(T, kEnhChunk);
    dim3 grid(l3ac_cdiv(T, kEnhUpTile), B);
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->C_in == 48)
        enhance_up_kernel<48, 24><<<grid, 256, 0, st>>>(x, B, T, partials, nchunk, reinterpret_cast<const float4*>(branches),
                                                         static_cast<const EnhUpBlob<48, 24>*>(plan->dev_blob), out);
    else
        enhance_up_kernel<96, 48><<<grid, 256, 0, st>>>(x, B, T, partials, nchunk, reinterpret_cast<const float4*>(branches),
                                                         static_cast<const EnhUpBlob<96, 48>*>(plan->dev_blob), out);
    return l3ac_launch_status();
}
