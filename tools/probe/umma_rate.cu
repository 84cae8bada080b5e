// umma_rate.cu -- how many cycles does one tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) take when
// the operands come from shared memory, as a function of N and of how the A/B tiles are walked?
// One CTA per SM (grid = 148), one thread issues REPS x 4 MMAs over a ring of A tiles, commit, wait.
// Garbage operands (values irrelevant).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// mode 0: A walks a ring of `atiles` 16 KB tiles, B walks 8 chunks (the GRU pattern); mode 1: same A tile and same B chunk every time
__global__ void __launch_bounds__(128) rate(int N, int reps, int atiles, int mode, long long *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem;                    // atiles * 16 KB
    unsigned char *sB = smem + atiles * 16384;   // 8 chunks of N rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (atiles * 16384 + 8 * N * 128) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        long long t0 = clock64();
        int at = 0, bc = 0;
        uint32_t sink = 0;
        for (int r = 0; r < reps; ++r) {
            const uint32_t a = a0 + (mode ? 0 : at * 16384), b = b0 + (mode ? 0 : bc * N * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma(tmem_base, make_desc(a + k * 32), make_desc(b + k * 32), idesc, 1u);
            if (++at == atiles) at = 0;
            if (++bc == 8) bc = 0;
            if (mode >= 2) {
                // a dependent long-latency instruction between tiles (the GRU issue thread's barrier probe):
                // does it overlap with the queued MMAs or add to them?
                for (int q = 0; q < mode - 1; ++q) {
                    uint32_t ok;
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(1u) : "memory");
                    sink += ok;
                }
            }
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0 + (sink == 0xffffffffu); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}
int main()
{
    long long *d, h[2];
    CK(cudaMalloc(&d, 16));
    const int reps = 2000;
    const int Ns[] = {64, 128};
    for (int mode = 0; mode < 5; ++mode)
        for (int N : Ns) {
            const int atiles = 5;
            const size_t smem = atiles * 16384 + 8 * N * 128 + 1024;
            CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            for (int it = 0; it < 2; ++it) {
                rate<<<148, 128, smem>>>(N, reps, atiles, mode, d);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
            printf("mode %d  M=128 N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d), smem read %d B/MMA -> %.0f B/clk\n", mode, N,
                   h[0] / (4.0 * reps), h[1] / (4.0 * reps), 128 * N / 256, 4096 + N * 32, (4096 + N * 32) / (h[1] / (4.0 * reps)));
        }
    return 0;
}
