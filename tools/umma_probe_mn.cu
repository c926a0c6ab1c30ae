// Probe: tcgen05.mma with an MN-MAJOR, no-swizzle B operand -- the layout of V in P.V of the attention kernel: V is stored
// as planes [d / 8][key][8 d-values] (a key's 8 consecutive head-dim values are 16 contiguous bytes, keys 16 B apart), so
// the contraction dimension K = keys is the STRIDED one.  Tests both readings of (LBO, SBO) and a key-row offset
// (K-step k starts at key 16 k), with A (= P, K-major planes [key / 8][row][8]) as in umma_probe.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/build/umma_probe_mn tools/umma_probe_mn.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

constexpr int kKeys = 128;     // rows per V plane
constexpr int kN = 48;         // head dim 32 + 16 (ones column + padding)

// A: planes [kKeys / 8][128 rows][8 keys] (K-major).  V: planes [kN / 8][kKeys][8] (MN-major B).  D[128][kN] over `ksteps` K=16 steps.
__global__ void probe(const __nv_bfloat16* a, const __nv_bfloat16* v, float* d, int ksteps, int swap, int a_f16) {
    __shared__ __align__(128) __nv_bfloat16 sa[(kKeys / 8) * 128 * 8];
    __shared__ __align__(128) __nv_bfloat16 sv[(kN / 8) * kKeys * 8];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (kKeys / 8) * 128 * 8; i += blockDim.x) sa[i] = a[i];
    for (int i = threadIdx.x; i < (kN / 8) * kKeys * 8; i += blockDim.x) sv[i] = v[i];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        // idesc: D fp32, A/B bf16, A K-major, B MN-major (bit 16), N = 48, M = 128
        const uint32_t idesc = (1u << 4) | (a_f16 ? 0u : (1u << 7)) | (1u << 10) | (1u << 16) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int k = 0; k < ksteps; ++k) {
            const uint64_t ad = make_desc(smem_u32(sa) + 2 * k * (128 * 16), 128 * 16, 128);
            const uint32_t vaddr = smem_u32(sv) + 16 * k * 16;        // keys 16 k ..
            const uint64_t bd = swap ? make_desc(vaddr, kKeys * 16, 128) : make_desc(vaddr, 128, kKeys * 16);
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                "l"(ad), "l"(bd), "r"(idesc), "r"(k > 0 ? 1u : 0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE;\nbra WAIT;\nDONE:\n}\n" ::"r"(
            smem_u32(&bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < kN; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) d[threadIdx.x * kN + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main() {
    const int na = (kKeys / 8) * 128 * 8, nv = (kN / 8) * kKeys * 8;
    __nv_bfloat16 *ha = new __nv_bfloat16[na], *hv = new __nv_bfloat16[nv];
    float *fa = new float[na], *fv = new float[nv];
    srand(11);
    for (int i = 0; i < na; ++i) { ha[i] = __float2bfloat16((rand() % 17 - 8) / 8.0f); fa[i] = __bfloat162float(ha[i]); }
    for (int i = 0; i < nv; ++i) { hv[i] = __float2bfloat16((rand() % 13 - 6) / 4.0f); fv[i] = __bfloat162float(hv[i]); }
    __nv_bfloat16 *da, *dv;
    float* dd;
    cudaMalloc(&da, na * 2); cudaMalloc(&dv, nv * 2); cudaMalloc(&dd, 128 * kN * 4);
    cudaMemcpy(da, ha, na * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dv, hv, nv * 2, cudaMemcpyHostToDevice);
    float* hd = new float[128 * kN];
    __half* hah = new __half[na];
    for (int i = 0; i < na; ++i) hah[i] = __float2half(fa[i]);       // same values, fp16 encoding
    __nv_bfloat16* dah;
    cudaMalloc(&dah, na * 2);
    cudaMemcpy(dah, hah, na * 2, cudaMemcpyHostToDevice);
    for (int a_f16 = 0; a_f16 < 2; ++a_f16)
    for (int swap = 0; swap < 2 - a_f16; ++swap)
        for (int ksteps : {1, 8}) {
            cudaMemset(dd, 0, 128 * kN * 4);
            probe<<<1, 128>>>(a_f16 ? dah : da, dv, dd, ksteps, swap, a_f16);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("swap=%d ksteps=%d: CUDA error %s\n", swap, ksteps, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hd, dd, 128 * kN * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < kN; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 16 * ksteps; ++k)      // A[m][k] = plane k/8, row m, elem k%8;  V[k][n] = plane n/8, row k, elem n%8
                        ref += (double)fa[((k / 8) * 128 + m) * 8 + (k & 7)] * fv[((n / 8) * kKeys + k) * 8 + (n & 7)];
                    maxerr = fmax(maxerr, fabs(ref - hd[m * kN + n]));
                }
            printf("A %s, B bf16 MN-major: %s  ksteps=%d  max_err %.3g\n", a_f16 ? "fp16" : "bf16", swap ? "LBO = plane stride, SBO = 128" : "LBO = 128 (next 8 keys), SBO = plane stride", ksteps, maxerr);
        }
    return 0;
}
