// umma_shift.cu -- does a row-shifted A-operand start address (the convNet kernel's implicit-GEMM trick: tap t reads
// the same activation buffer (t-3)*dilation rows further on) change the cost of an M=128 N=64 K=16 tcgen05.mma?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_shift umma_shift.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
template <int N>
__global__ void __launch_bounds__(128) issue(int reps, int shift_rows, int two_tiles, long long *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (65536 + N * 128) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t a = make_desc(smem_u32(smem) + shift_rows * 128), b = make_desc(smem_u32(smem + 65536));
        const uint64_t a2 = a + (two_tiles ? (128 * 128 >> 4) : 0);  // second 128-row tile 16 KB further on
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int k = 0; k < 16; ++k) umma(tmem_base + ((k & 4) ? 128 : 0), ((k & 4) ? a2 : a) + (k & 3) * 2, b + (k & 3) * 2, idesc);
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}
template <int N> void run(long long *d, int shift, int two)
{
    long long h[2];
    const int reps = 500;
    const size_t smem = 65536 + N * 128 + 1024;
    CK(cudaFuncSetAttribute(issue<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int it = 0; it < 2; ++it) { issue<N><<<148, 128, smem>>>(reps, shift, two, d); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("N=%3d A start shifted by %2d rows%s: %.1f cyc/MMA\n", N, shift, two ? " (two row tiles alternating)" : "", h[1] / (16.0 * reps));
}
int main()
{
    long long *d;
    CK(cudaMalloc(&d, 16));
    for (int sh : {0, 1, 3, 4, 6, 8, 12}) run<64>(d, sh, 0);
    run<64>(d, 3, 1);
    for (int sh : {0, 3, 12}) run<128>(d, sh, 0);
    return 0;
}
