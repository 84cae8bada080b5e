// umma2_probe.cu -- bring-up of tcgen05.mma.cta_group::2 (CTA pair): D[256 x 128] = A[256 x K] * B[128 x K]^T with
// A rows 128r..128r+127 in CTA r's shared memory, B rows 64r..64r+63 in CTA r's shared memory (K-major SWIZZLE_128B,
// generic stores), fp16 operands.  Checks where D lands (each CTA dumps its 128 TMEM lanes x 128 columns) and measures
// the issue/execute interval of back-to-back M=256 N=128 K=16 MMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma2_probe umma2_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr int KCH = 2;  // K = 128

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe(const __half *__restrict__ A, const __half *__restrict__ B,
                                                                       float *__restrict__ D, int reps, long long *cyc)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem;                   // KCH x [128 rows x 64 k]
    unsigned char *sB = smem + KCH * 16384;     // KCH x [64 rows x 64 k]
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int K = KCH * 64;
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K, kc = k >> 6, kk = k & 63;
        *reinterpret_cast<__half *>(sA + kc * 16384 + r * 128 + (((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2) = A[(size_t)(rank * 128 + r) * K + k];
    }
    for (int i = tid; i < 64 * K; i += 128) {
        const int r = i / K, k = i % K, kc = k >> 6, kk = k & 63;
        *reinterpret_cast<__half *>(sB + kc * 8192 + r * 128 + (((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2) = B[(size_t)(rank * 64 + r) * K + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;

    if (rank == 0 && tid == 0) {
        // idesc: c=F32 (1<<4), a=b=F16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int c = 0; c < KCH; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t da = make_desc(smem_u32(sA + c * 16384) + k * 32);
                    const uint64_t db = make_desc(smem_u32(sB + c * 8192) + k * 32);
                    const uint32_t acc = (r | c | k) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar_done)), "h"((uint16_t)3) : "memory");
        mbar_wait(smem_u32(&bar_done), 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar_done), 0);   // both CTAs: the multicast commit arrives on each CTA's barrier
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) D[((size_t)(blockIdx.x >> 1) * 256 + rank * 128 + warp * 32 + lane) * 128 + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

int main()
{
    const int K = KCH * 64, M = 256, N = 128;
    std::vector<float> A(M * K), B(N * K);
    srand(1);
    for (auto &v : A) v = (rand() % 2001 - 1000) / 1000.0f;
    for (auto &v : B) v = (rand() % 2001 - 1000) / 1000.0f;
    std::vector<__half> Ah(M * K), Bh(N * K);
    for (int i = 0; i < M * K; ++i) { Ah[i] = __float2half(A[i]); A[i] = __half2float(Ah[i]); }
    for (int i = 0; i < N * K; ++i) { Bh[i] = __float2half(B[i]); B[i] = __half2float(Bh[i]); }
    __half *dA, *dB; float *dD; long long *dC;
    const int pairs = 74;
    CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dD, (size_t)pairs * M * N * 4)); CK(cudaMalloc(&dC, 16));
    CK(cudaMemcpy(dA, Ah.data(), M * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bh.data(), N * K * 2, cudaMemcpyHostToDevice));
    const int smem = KCH * (16384 + 8192) + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // correctness: one repetition
    CK(cudaMemset(dD, 0, (size_t)pairs * M * N * 4));
    probe<<<2, 128, smem>>>(dA, dB, dD, 1, dC);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(M * N);
    CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
            maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
        }
    printf("cta_group::2 M=256 N=128 K=%d: max |D - ref| = %.3e (D row m = unit m of CTA m/128, column n = B row n)\n", K, maxerr);
    // rate: all 74 pairs busy
    long long h[2];
    for (int it = 0; it < 2; ++it) { probe<<<2 * pairs, 128, smem>>>(dA, dB, dD, 1000, dC); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, dC, 16, cudaMemcpyDeviceToHost));
    printf("back-to-back: issue %.1f cyc/MMA, complete %.1f cyc/MMA (M=256 N=128 K=16 per MMA = 2 x (128x128x16) per SM pair)\n",
           h[0] / (8.0 * 1000), h[1] / (8.0 * 1000));
    // latency: 8 MMAs (two tiles) issued once, then commit (multicast) -> wait
    for (int it = 0; it < 2; ++it) { probe<<<2 * pairs, 128, smem>>>(dA, dB, dD, 1, dC); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, dC, 16, cudaMemcpyDeviceToHost));
    printf("8 MMAs: issued after %lld cycles, commit observed after %lld cycles (8 x 64 = 512 of them are execution)\n", h[0], h[1]);
    return 0;
}
