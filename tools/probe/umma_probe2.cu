// umma_probe2.cu (row-shifted B operand start, base_offset variants) -- derived from umma_probe.cu -- bring-up check for the hand-built tcgen05 path used by the GRU kernel:
// K-major SWIZZLE_128B operand tiles written by generic stores (B) and by cp.async.bulk (A),
// hand-packed smem/instruction descriptors, tcgen05.mma cta_group::1 kind::f16 (bf16 -> fp32),
// tcgen05.commit -> mbarrier, tcgen05.ld 32x32b.  D[128 x 64] = A[128 x K] * B[64 x K]^T.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // K-major, SWIZZLE_128B: start>>4 | LBO 0 | SBO (1024 B >> 4) << 32 | version 1 << 46 | layout 2 << 61
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int KCH>
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16 *__restrict__ Apacked,
                                             const __nv_bfloat16 *__restrict__ Brow,     // [96][K] row-major
                                             float *__restrict__ D, int shift, int use_base_offset)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem;                       // KCH * 16 KB
    unsigned char *sB = smem + KCH * 16384;         // KCH * 12 KB (96 rows)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + KCH * 12288);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = KCH * 64;

    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);  // A landed
        mbar_init(smem_u32(&bars[1]), 1);  // MMA done
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // B: generic stores into the swizzled K-major layout (what the GRU epilogue does for h)
    for (int i = tid; i < 96 * K; i += 128) {
        const int row = i / K, k = i % K;
        const int kc = k >> 6, kk = k & 63;
        const uint32_t off = kc * 12288 + row * 128 + ((((kk >> 3) ^ (row & 7)) << 4)) + (kk & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(sB + off) = Brow[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        mbar_expect_tx(smem_u32(&bars[0]), KCH * 16384);
        for (int c = 0; c < KCH; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sA + c * 16384)), "l"(Apacked + (size_t)c * 8192), "r"(16384u), "r"(smem_u32(&bars[0])) : "memory");
        mbar_wait(smem_u32(&bars[0]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // idesc: c=F32 (1<<4), a=BF16 (1<<7), b=BF16 (1<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        for (int c = 0; c < KCH; ++c) {
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = make_desc(smem_u32(sA + c * 16384) + k * 32);
                const uint32_t baddr = smem_u32(sB + c * 12288) + shift * 128 + k * 32;
                uint64_t db = make_desc(baddr);
                if (use_base_offset) db |= (uint64_t)((baddr >> 7) & 7) << 49;
                const uint32_t acc = (c | k) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
    }
    __syncwarp();
    mbar_wait(smem_u32(&bars[1]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // each warp reads its 32 lanes (rows), 64 columns in 4 x16 chunks
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) D[(warp * 32 + lane) * 64 + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u));
}

int main()
{
    const int KCH = 8, K = KCH * 64, M = 128, N = 64;
    std::vector<float> A(M * K), B(96 * K);
    srand(1);
    for (auto &v : A) v = (rand() % 2001 - 1000) / 1000.0f;
    for (auto &v : B) v = (rand() % 2001 - 1000) / 1000.0f;
    std::vector<__nv_bfloat16> Ap(M * K), Bb(96 * K);
    std::vector<float> Ar(M * K), Br(96 * K);
    for (int r = 0; r < M; ++r)
        for (int k = 0; k < K; ++k) {
            __nv_bfloat16 b = __float2bfloat16(A[r * K + k]);
            Ar[r * K + k] = __bfloat162float(b);
            const int kc = k >> 6, kk = k & 63;
            const size_t off = (size_t)kc * 16384 + r * 128 + (((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2;
            Ap[off / 2] = b;
        }
    for (int i = 0; i < 96 * K; ++i) { Bb[i] = __float2bfloat16(B[i]); Br[i] = __bfloat162float(Bb[i]); }
    __nv_bfloat16 *dA, *dB; float *dD;
    CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, 96 * K * 2)); CK(cudaMalloc(&dD, M * N * 4));
    CK(cudaMemcpy(dA, Ap.data(), M * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bb.data(), 96 * K * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, M * N * 4));
    const int smem = KCH * (16384 + 12288) + 256;
    CK(cudaFuncSetAttribute(probe<KCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int bad = 0;
    for (int ubo = 0; ubo < 2; ++ubo)
        for (int shift : {0, 8, 3, 12, 13, 29}) {
            CK(cudaMemset(dD, 0, M * N * 4));
            probe<KCH><<<1, 128, smem>>>(dA, dB, dD, shift, ubo);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            std::vector<float> D(M * N);
            CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0, maxref = 0;
            for (int r = 0; r < M; ++r)
                for (int c = 0; c < N; ++c) {
                    double s = 0;
                    for (int k = 0; k < K; ++k) s += (double)Ar[r * K + k] * Br[(c + shift) * K + k];
                    maxerr = fmax(maxerr, fabs(s - D[r * N + c]));
                    maxref = fmax(maxref, fabs(s));
                }
            printf("shift %2d base_offset_field %d: max |err| = %.3e (ref %.2f) -> %s\n", shift, ubo, maxerr, maxref, maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
            if (ubo == 0 && maxerr >= 1e-3 * maxref) bad = 1;
        }
    return bad;
}
