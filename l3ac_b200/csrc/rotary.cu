// Rotary-position path of LocalMHA (en_coder_dynamic_pos = false: l3ac/local_trans.py:29,36; the ModelConfig default,
// l3ac/en_codec.py:12).  local-attention rotates keys by their position 0..2w-1 inside the bucket [previous window ; own
// window] and queries by w..2w-1 (apply_rotary_pos_emb after look_around), so a key carries TWO rotations: one as a member of
// its own window, one as the look-back of the next.  Instead of teaching the attention kernels two key versions, the rotated
// q/k/v are laid out as one independent "segment" per window and the unchanged block-local attention kernels run on those:
//
//   segment n >= 1 (2w rows):  rows [0, w)  = window n-1: k rotated by r, v          (q rows are zero: their outputs are unused)
//                              rows [w, 2w) = window n  : q and k rotated by r, v
//   segment 0:                 rows [0, w)  = window 0  : q and k rotated by w + r, v;  rows [w, 2w) zero (unused)
//
// With T' = 2w and window' = w the kernels' mask "keys max(0, (r/w - 1) w) .. r" gives every real query exactly its
// reference key set, and a zero bias table stands in for the absent DynamicPositionBias.  l3ac_rotary_unpack gathers the
// useful output rows back to (B, T, H*D).  Rows past T are zero, as autopad's zero rows are after rotation.
// Arithmetic per element follows apply_rotary_pos_emb: fl(fl(x cos) + fl(rotate_half(x) sin)), no FMA contraction; the
// cos/sin tables are computed on the host exactly as SinusoidalEmbeddings does (fp32 angle = t * inv_freq).
#include "common.cuh"

namespace l3ac {

template <int KIND>   // 0 fp32, 1 bf16, 2 bf16 (hi, lo) pair
__device__ __forceinline__ void rot_store(void* hi, void* lo, long long i, float v) {
    if (KIND == 0) {
        reinterpret_cast<float*>(hi)[i] = v;
    } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        reinterpret_cast<__nv_bfloat16*>(hi)[i] = h;
        if (KIND == 2) reinterpret_cast<__nv_bfloat16*>(lo)[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) rotary_pack_kernel(const float* __restrict__ qkv, int B, int T, int H, int window,
                                                          int nw, const float* __restrict__ cos_t,
                                                          const float* __restrict__ sin_t, void* __restrict__ out_hi,
                                                          void* __restrict__ out_lo) {
    constexpr int D = 32, HD = 16;
    const long long total = (long long)B * nw * 2 * window * H * HD;
    const int ld = 3 * H * D;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(e % HD);
        long long t = e / HD;
        const int h = (int)(t % H);
        t /= H;
        const int r = (int)(t % (2 * window));
        t /= 2 * window;
        const int n = (int)(t % nw), b = (int)(t / nw);
        int p, ang;
        bool want_q;
        if (n == 0) {
            p = r < window ? r : -1;
            ang = window + r;
            want_q = true;
        } else if (r < window) {
            p = (n - 1) * window + r;
            ang = r;
            want_q = false;
        } else {
            p = n * window + (r - window);
            ang = r;
            want_q = true;
        }
        const long long orow = (((long long)b * nw + n) * 2 * window + r) * ld + h * D + d;
        float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
        if (p >= 0 && p < T) {
            const float* src = qkv + ((long long)b * T + p) * ld + h * D + d;
            const float c = __ldg(cos_t + ang * D + d), s = __ldg(sin_t + ang * D + d);     // table columns d and d + 16 are equal
            const float ka = __ldg(src + H * D), kb = __ldg(src + H * D + HD);
            k0 = __fadd_rn(__fmul_rn(ka, c), __fmul_rn(-kb, s));          // rotate_half: (-x2, x1)
            k1 = __fadd_rn(__fmul_rn(kb, c), __fmul_rn(ka, s));
            if (want_q) {
                const float qa = __ldg(src), qb = __ldg(src + HD);
                q0 = __fadd_rn(__fmul_rn(qa, c), __fmul_rn(-qb, s));
                q1 = __fadd_rn(__fmul_rn(qb, c), __fmul_rn(qa, s));
            }
            v0 = __ldg(src + 2 * H * D);
            v1 = __ldg(src + 2 * H * D + HD);
        }
        rot_store<KIND>(out_hi, out_lo, orow, q0);
        rot_store<KIND>(out_hi, out_lo, orow + HD, q1);
        rot_store<KIND>(out_hi, out_lo, orow + H * D, k0);
        rot_store<KIND>(out_hi, out_lo, orow + H * D + HD, k1);
        rot_store<KIND>(out_hi, out_lo, orow + 2 * H * D, v0);
        rot_store<KIND>(out_hi, out_lo, orow + 2 * H * D + HD, v1);
    }
}

// out row p of clip b  <-  segment n = p / w, row (n == 0 ? p : w + p mod w); rows are `row_vec` 16-byte vectors long.
__global__ void __launch_bounds__(256) rotary_unpack_kernel(const uint4* __restrict__ seg, uint4* __restrict__ out, int B, int T,
                                                            int window, int nw, int row_vec) {
    const long long total = (long long)B * T * row_vec;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % row_vec);
        const long long row = e / row_vec;
        const int p = (int)(row % T), b = (int)(row / T);
        const int n = p / window, i = p - n * window;
        const long long srow = ((long long)b * nw + n) * 2 * window + (n == 0 ? i : window + i);
        out[e] = __ldg(seg + srow * row_vec + c);
    }
}

}  // namespace l3ac

extern "C" int l3ac_rotary_pack(const float* qkv, int B, int T, int H, int D, int window, const float* cos_table,
                                const float* sin_table, void* out, void* out_lo, int out_dtype, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(qkv && cos_table && sin_table && out && B > 0 && T > 0 && H > 0 && window > 0);
    L3AC_CHECK_ARG(out_dtype == L3AC_F32 || out_dtype == L3AC_BF16 || (out_dtype == L3AC_BF16X2 && out_lo));
    if (D != 32) return L3AC_EUNSUPPORTED;
    const int nw = l3ac_cdiv(T, window);
    const long long total = (long long)B * nw * 2 * window * H * 16;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)l3ac_sm_count() * 16;
    const int grid = (int)(blocks < cap ? blocks : cap);
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == L3AC_F32)
        l3ac::rotary_pack_kernel<0><<<grid, 256, 0, st>>>(qkv, B, T, H, window, nw, cos_table, sin_table, out, out_lo);
    else if (out_dtype == L3AC_BF16)
        l3ac::rotary_pack_kernel<1><<<grid, 256, 0, st>>>(qkv, B, T, H, window, nw, cos_table, sin_table, out, out_lo);
    else
        l3ac::rotary_pack_kernel<2><<<grid, 256, 0, st>>>(qkv, B, T, H, window, nw, cos_table, sin_table, out, out_lo);
    return l3ac_launch_status();
}

extern "C" int l3ac_rotary_unpack(const void* seg, void* out, int B, int T, int window, int row_bytes, l3ac_stream_t stream) {
    L3AC_CHECK_ARG(seg && out && B > 0 && T > 0 && window > 0 && row_bytes > 0 && row_bytes % 16 == 0);
    L3AC_CHECK_ARG(((reinterpret_cast<uintptr_t>(seg) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    const int nw = l3ac_cdiv(T, window);
    const long long total = (long long)B * T * (row_bytes / 16);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)l3ac_sm_count() * 16;
    const int grid = (int)(blocks < cap ? blocks : cap);
    l3ac::rotary_unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(seg), reinterpret_cast<uint4*>(out),
                                                                         B, T, window, nw, row_bytes / 16);
    return l3ac_launch_status();
}
