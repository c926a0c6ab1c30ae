// Shared device/host helpers for the l3ac_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>

#include "../../include/l3ac_b200.h"

#define L3AC_CHECK_ARG(cond)            \
    do {                                \
        if (!(cond)) return L3AC_EINVAL; \
    } while (0)

// Kernel launches are asynchronous; report launch-configuration errors to the caller.
static inline int l3ac_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? L3AC_OK : (int)e;
}

// SM count of the CURRENT device (cached per device id: one process may drive several GPUs).
static inline int l3ac_sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }
    if (cache[dev] == 0) cudaDeviceGetAttribute(&cache[dev], cudaDevAttrMultiProcessorCount, dev);
    return cache[dev];
}

// Programmatic dependent launch for the kernels that support it (L3AC_PDL=0 disables): the prologue of kernel N+1 overlaps
// the tail of kernel N; see the griddepcontrol instructions in the kernels.
static inline bool l3ac_pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("L3AC_PDL");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on == 1;
}

// Launch with (or, when disabled, without) the programmatic-stream-serialization attribute.
template <typename... KArgs, typename... Args>
static inline void l3ac_launch(void (*fn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = l3ac_pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, fn, KArgs(args)...);       // errors surface through l3ac_launch_status()
}

// Device side: call once per thread after the part of the prologue that touches only kernel parameters, weights and shared
// memory.  No-ops when the launch does not carry the attribute.
#define L3AC_PDL_SYNC()                                                      \
    do {                                                                     \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      \
        asm volatile("griddepcontrol.wait;" ::: "memory");                   \
    } while (0)

static inline int l3ac_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// shared argument validation of the two GEMM entry points (defined in gemm_f32.cu)
int l3ac_validate_gemm_desc(const l3ac_gemm_desc* d);

namespace l3ac {

constexpr float kEps = 1e-8f;   // l3ac/xtract/nn/utils.py:33 (float32(1e-8))

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// x + sin^2(alpha x) / (alpha + 1e-8)      (l3ac/layers.py:29-33)
__device__ __forceinline__ float snake_f(float x, float alpha, float inv_alpha) {
    float s = sinf(alpha * x);
    return x + inv_alpha * (s * s);
}

// exact-erf GELU (nn.GELU() default / F.gelu default)
__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// GELU with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7 absolute, i.e. ~1 ulp of the O(1) values it produces)
// on MUFU.RCP / MUFU.EX2: 13 instructions against ~28 for erff's two-branch polynomial.  Used where erf dominates the
// instruction count (the 80-wide hidden layer of the encoder stem).
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float z = x * 0.70710678118654752440f, a = fabsf(z);
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, a, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * a * -1.4426950408889634f));
    const float erf_abs = fmaf(-p * t, e, 1.0f);
    return 0.5f * x * (1.0f + copysignf(erf_abs, z));
}

// 1/sqrt(v) for v >= 1e-8 (variance + eps): MUFU.RSQ and one Newton-Raphson step (<= 1 ulp, like sqrt followed by a divide)
// in 5 instructions instead of the ~20 of the IEEE `1.0f / sqrtf(v)` sequence with its slow-path branches.
__device__ __forceinline__ float rsqrt_nr(float v) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
    const float h = 0.5f * v * y;
    return fmaf(y, fmaf(-h, y, 0.5f), y);        // y * (1.5 - 0.5 v y^2)
}

// Packed fp32 FMA (sm_100 `fma.rn.f32x2`, SASS FFMA2): two independent IEEE fp32 FMAs per issued instruction.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
// Two lanes of  f(v) = (v + sin^2(alpha v) / (alpha + eps)) * scale + shift  (snake + folded GRN affine) with the MUFU
// sine; the packed ops round exactly like their scalar counterparts, so this is bit-identical to the scalar formula.
__device__ __forceinline__ float2 snake_affine2(float2 v, float2 alpha, float2 inv_alpha, float2 scale, float2 shift) {
    const float2 t = fmul2(alpha, v);
    const float2 s = make_float2(__sinf(t.x), __sinf(t.y));
    v = ffma2(inv_alpha, fmul2(s, s), v);
    return ffma2(v, scale, shift);
}

// Output-type adapters used by kernels that can emit fp32 or bf16 activations.
template <typename T> __device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace l3ac
