// Probe: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) as a function of N and of the shared-memory
// layout of the operands, measured in isolation (one CTA, one issuing thread, `reps` MMAs back to back, one commit).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/build/umma_rate tools/umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t swz) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)swz << 61;
    return d;
}

// mode 0: no-swizzle planes, A rows at 16 B stride (SBO 128, LBO = plane); mode 1: SWIZZLE_128B K-major tiles (SBO 1024)
// distinct != 0: every MMA reads a different A tile (row shift), else the same one
__device__ volatile int g_stop;
__global__ void rate(int n, int mode, int reps, int distinct, int ndst, int busy, int mma_warp, long long* out, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(smem + 190 * 1024) = 0;
    if (threadIdx.x < 32) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(smem) + 191 * 1024 + 8 * threadIdx.x));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (warp != mma_warp) {
        if (busy == 0) goto done;
        float acc[8];
        for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 0.001f + i;
        volatile int* flag = reinterpret_cast<volatile int*>(smem + 190 * 1024);
        const uint32_t my = smem_u32(smem) + 100 * 1024 + threadIdx.x * 16;
        for (int it = 0; it < 100000 && *flag == 0; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i] = __sinf(acc[i] * 1.0001f) * acc[i] + 0.5f;
                acc[i] = fmaf(acc[i], 0.999f, 0.001f);
                acc[i] = fmaf(acc[i], 1.001f, -0.001f);
            }
            if (busy >= 2) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my), "r"(__float_as_uint(acc[0])), "r"(1u), "r"(2u), "r"(3u) : "memory");
            if (busy == 4 || busy == 6) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (busy == 5 || busy == 6) { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncwarp(); if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(smem) + 191 * 1024 + 8 * warp) : "memory"); }
            if (busy == 3 && warp < 4) {
                uint32_t v[8];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(tmem + ((uint32_t)(warp * 32) << 16) + 256));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc[1] += __uint_as_float(v[0]) * 1e-30f;
            }
        }
        float t = 0;
        for (int i = 0; i < 8; ++i) t += acc[i];
        sink[threadIdx.x] = t;
    } else {
        uint32_t leader;
        asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(leader));
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 96 * 1024;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t ad0 = mode == 0 ? make_desc(a0, 20000, 128, 0) : make_desc(a0, 16, 1024, 2);
        const uint64_t bd0 = mode == 0 ? make_desc(b0, n * 16, 128, 0) : make_desc(b0, 16, 1024, 2);
        const uint32_t astep = distinct ? (mode == 0 ? 5u : 1024u) : 0u;     // in 16-byte units
        const uint32_t dstep = ndst > 1 ? (uint32_t)n : 0u;
        long long t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
            if (leader) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint64_t ad = ad0 + (uint64_t)(astep * (q & 3));
                    const uint32_t d = tmem + dstep * (q & 1);
                    asm volatile(
                        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                        "l"(ad), "l"(bd0), "r"(idesc), "r"(1u)
                        : "memory");
                }
            }
            __syncwarp();
        }
        long long t1 = clock64();
        if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile(
            "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE;\nbra WAIT;\nDONE:\n}\n" ::"r"(
                smem_u32(&bar))
            : "memory");
        long long t2 = clock64();
        if (leader) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
        *reinterpret_cast<volatile int*>(smem + 190 * 1024) = 1;
    }
done:
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* d;
    float* sink;
    cudaMalloc(&d, 16);
    cudaMalloc(&sink, 4096 * 4);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 2048;
    const int ns[] = {16, 32, 64, 128, 256};
    for (int busy = 0; busy < 7; ++busy)
        for (int cfg = 0; cfg < 3; ++cfg) {
            const int threads = busy == 0 ? 128 : 544, mma_warp = cfg == 0 ? 0 : (busy == 0 ? 3 : 16);
            const int mode = cfg == 2 ? 1 : 0;
            for (int n : ns) {
                long long h[2];
                for (int w = 0; w < 2; ++w) {
                    rate<<<1, threads, 200 * 1024>>>(n, mode, reps, 1, 1, busy, mma_warp, d, sink);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("busy=%d (%s) %s mma_warp=%2d N=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma (floor %.0f)\n", busy,
                       busy == 0 ? "alone" : busy == 1 ? "16 warps sin+fma" : busy == 2 ? "+st.shared" : busy == 3 ? "+tcgen05.ld" : busy == 4 ? "+st.shared+fence.proxy.async" : busy == 5 ? "+st.shared+mbar arrive" : "+st+fence+arrive",
                       mode ? "SW128    " : "noswizzle", mma_warp, n, (double)h[0] / reps, (double)h[1] / reps, 128.0 * n / 256);
            }
        }
    return 0;
}
