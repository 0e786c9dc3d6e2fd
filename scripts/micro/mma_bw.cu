// Microbenchmark: issue rate of tcgen05.mma kind::tf32 (SS operands, K-major no-swizzle layout)
// for a few N, M=128, K=8 per instruction.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MODE 0: tf32 no-swizzle; MODE 1: bf16 no-swizzle (K=16 per instruction, same 32 B of K per row)
template <int MODE>
__global__ void mma_rate(int n, int ksteps, int jobs, int nslots, unsigned long long *cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 + 256) * 32 * 10 / 4; i += blockDim.x) ((float *)smem)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = slot;
    // K-major no swizzle: [kchunk 16B][row][16B]; LBO = rows*16, SBO = 128
    const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 128 * 32 * 10;
    const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)n * 16;
    const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    const uint32_t a_lo0 = ((a_addr >> 4) & 0x3fff) | (((a_lbo >> 4) & 0x3fff) << 16);
    const uint32_t b_lo0 = ((b_addr >> 4) & 0x3fff) | (((b_lbo >> 4) & 0x3fff) << 16);
    uint32_t idesc;
    if (MODE == 0) idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    else idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);  // bf16 x bf16 -> f32
    unsigned long long t0 = 0, t1 = 0;
    if (warp == 0) {
        uint32_t elected;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(elected));
        t0 = clock64();
        if (elected) {
            for (int j = 0; j < jobs; ++j) {
                const uint32_t d = tbase + (uint32_t)((j % nslots) * n);
                uint32_t a_lo = a_lo0, b_lo = b_lo0;
                for (int ks = 0; ks < ksteps; ++ks) {
                    if (MODE == 0) mma_tf32(d, hi | a_lo, hi | b_lo, idesc, ks > 0);
                    else mma_bf16(d, hi | a_lo, hi | b_lo, idesc, ks > 0);
                    a_lo += (2 * a_lbo) >> 4;
                    b_lo += (2 * b_lbo) >> 4;
                }
            }
            commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        t1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}

template <int MODE>
void run(const char *name, int n, int ksteps, int nslots) {
    const int jobs = 4000, grid = 148;
    unsigned long long *cyc; cudaMalloc(&cyc, grid * 8);
    const size_t smem = (128 + 256) * 32 * 10 + 1024;
    cudaError_t ea = cudaFuncSetAttribute(mma_rate<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (ea != cudaSuccess) printf("attr: %s\n", cudaGetErrorString(ea));
    mma_rate<MODE><<<grid, 128, smem>>>(n, ksteps, 10, nslots, cyc);
    mma_rate<MODE><<<grid, 128, smem>>>(n, ksteps, jobs, nslots, cyc);
    cudaError_t e = cudaGetLastError(); cudaError_t e2 = cudaDeviceSynchronize(); if (e == cudaSuccess) e = e2;
    unsigned long long h[148]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < grid; ++i) c += (double)h[i]; c /= grid;
    const double per = c / ((double)jobs * ksteps);
    const double kper = MODE == 0 ? 8 : 16;
    printf("%-10s N=%3d ksteps=%2d slots=%d: %7.1f cycles/MMA  -> %7.1f TFLOP/s at 1.9 GHz x148  (%s)\n", name, n, ksteps, nslots,
           per, 2.0 * 128 * n * kper / per * 1.9e9 * 148 / 1e12, cudaGetErrorString(e));
    cudaFree(cyc);
}

int main() {
    for (int n : {64, 128, 256}) run<0>("tf32 SS", n, 5, 512 / n >= 4 ? 4 : 512 / n);
    run<0>("tf32 SS", 128, 5, 1);
    run<0>("tf32 SS", 128, 9, 4);
    run<0>("tf32 SS", 128, 1, 4);
    for (int n : {64, 128, 256}) run<1>("bf16 SS", n, 5, 512 / n >= 4 ? 4 : 512 / n);
    return 0;
}
