// Probe: tcgen05.mma with a NO-SWIZZLE K-major shared-memory layout whose operand tile is written by threads, with
// (1) arbitrary row shifts of the A tile expressed as start-address offsets, (2) the two 8-element K halves of one
// K=16 step taken from different planes / taps through the LBO field.  Prints the max error against the host result
// for both readings of the (LBO, SBO) descriptor fields.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/build/umma_probe tools/umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;      // swizzle mode 0
}

constexpr int kRowsTot = 256;     // rows per plane
constexpr int kPlanes = 4;

// A: [kPlanes][kRowsTot][8] bf16, B: [2][32][8] bf16.  D[128][32] = sum_k A'[m][k] B[n][k] with
// A'[m][0:8] = A[plane0][row0 + m][:], A'[m][8:16] = A[plane0][row0 + m][:] + lbo bytes.
__global__ void probe(const __nv_bfloat16* a, const __nv_bfloat16* b, float* d, int row0, int plane0, int lbo_bytes, int swap) {
    __shared__ __align__(128) __nv_bfloat16 sa[kPlanes * kRowsTot * 8];
    __shared__ __align__(128) __nv_bfloat16 sb[2 * 32 * 8];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < kPlanes * kRowsTot * 8; i += blockDim.x) sa[i] = a[i];
    for (int i = threadIdx.x; i < 2 * 32 * 8; i += blockDim.x) sb[i] = b[i];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a_addr = smem_u32(sa) + (plane0 * kRowsTot + row0) * 16;
        const uint32_t b_addr = smem_u32(sb);
        uint64_t ad, bd;
        if (!swap) {
            ad = make_desc(a_addr, lbo_bytes, 128);
            bd = make_desc(b_addr, 32 * 16, 128);
        } else {
            ad = make_desc(a_addr, 128, lbo_bytes);
            bd = make_desc(b_addr, 128, 32 * 16);
        }
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
            "l"(ad), "l"(bd), "r"(idesc), "r"(0u)
            : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE;\nbra WAIT;\nDONE:\n}\n" ::"r"(
            smem_u32(&bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
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
    for (int i = 0; i < 32; ++i) d[threadIdx.x * 32 + i] = __uint_as_float(v[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}

int main() {
    const int na = kPlanes * kRowsTot * 8, nb = 2 * 32 * 8;
    __nv_bfloat16 *ha = new __nv_bfloat16[na], *hb = new __nv_bfloat16[nb];
    float *fa = new float[na], *fb = new float[nb];
    srand(7);
    for (int i = 0; i < na; ++i) { ha[i] = __float2bfloat16((rand() % 17 - 8) / 8.0f); fa[i] = __bfloat162float(ha[i]); }
    for (int i = 0; i < nb; ++i) { hb[i] = __float2bfloat16((rand() % 13 - 6) / 4.0f); fb[i] = __bfloat162float(hb[i]); }
    __nv_bfloat16 *da, *db;
    float* dd;
    cudaMalloc(&da, na * 2); cudaMalloc(&db, nb * 2); cudaMalloc(&dd, 128 * 32 * 4);
    cudaMemcpy(da, ha, na * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb, nb * 2, cudaMemcpyHostToDevice);
    float* hd = new float[128 * 32];
    struct Case { int row0, plane0, lbo; const char* name; };
    const Case cases[] = {{0, 0, kRowsTot * 16, "aligned rows, next plane"},
                          {5, 0, kRowsTot * 16, "row shift 5, next plane"},
                          {27, 1, 9 * 16, "row shift 27, second half 9 rows later (tap pair)"},
                          {3, 2, 3 * 16, "row shift 3, second half 3 rows later"},
                          {100, 0, 2 * kRowsTot * 16, "row shift 100, plane + 2"}};
    for (int swap = 0; swap < 2; ++swap)
        for (const Case& c : cases) {
            cudaMemset(dd, 0, 128 * 32 * 4);
            probe<<<1, 128>>>(da, db, dd, c.row0, c.plane0, c.lbo, swap);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("swap=%d %s: CUDA error %s\n", swap, c.name, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hd, dd, 128 * 32 * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 32; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 16; ++k) {
                        const int base = (c.plane0 * kRowsTot + c.row0 + m) * 8 + (k >= 8 ? c.lbo / 2 : 0) + (k & 7);
                        ref += (double)fa[base] * fb[((k >> 3) * 32 + n) * 8 + (k & 7)];
                    }
                    maxerr = fmax(maxerr, fabs(ref - hd[m * 32 + n]));
                }
            printf("swap=%d  %-55s max_err %.3g\n", swap, c.name, maxerr);
        }
    return 0;
}
