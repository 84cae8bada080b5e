// tc_common.cuh -- thin PTX wrappers for the sm_100a tensor-core path (tcgen05 / TMEM / mbarrier /
// bulk async copies) shared by the neural decoders (conv_net.cu; gru_decode.cu carries its own copy).
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// linear (non-tensor) bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 x fp16 -> fp32
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row groups 1024 B apart.  The
// swizzle is a function of the absolute shared-memory address, so the start may be any 128 B row of a
// 1024 B-aligned buffer (row-shifted views; verified on B200 by tools/probe/umma_probe2.cu).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// the same descriptor from its low word ((address & 0x3FFFF) >> 4): consecutive K steps / row shifts are plain
// adds on that word, which keeps the single issuing thread ahead of the tensor pipe
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return (saddr & 0x3FFFF) >> 4; }
__device__ __forceinline__ uint64_t umma_desc_from_lo(uint32_t lo)
{
    return (uint64_t)lo | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: fp16 A/B (format 0), fp32 accumulate, both operands K-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem_addr, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem_addr), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// GELU.  The reference's nn.GELU() is the exact erf form (models.py:703).  The usual tanh form
// 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) is off by up to 4.7e-4, which on the reference-TRAINED convNet
// accumulates through the ten conv layers to 3e-3 on the logits -- by itself the whole tolerance (round 1 measured 1e-5 on
// random-init weights).  Here: x * sigmoid(2 x (c1 + c3 x^2 + c5 x^4)) with the odd quintic fitted to the erf form
// (minimax over [-8, 8]: |gelu_fit - gelu_erf| <= 2.6e-5), the sigmoid through ex2.approx + rcp.approx (relative
// error ~2^-22): 2 MUFU + 7 FP32 operations.  The polynomial argument is clamped to [-8, 8] (its quintic term turns it over
// beyond |x| ~ 10); outside, sigmoid is 0 or 1 to fp32 precision and the result is 0 or x like the erf form.
__device__ __forceinline__ float gelu_f(float x)
{
    const float c1 = -2.0f * 1.4426950409f * 7.97507884e-01f;   // -2 log2(e) c_k
    const float c3 = -2.0f * 1.4426950409f * 3.70056460e-02f;
    const float c5 = -2.0f * 1.4426950409f * -3.51516783e-04f;
    const float xc = fminf(fmaxf(x, -8.0f), 8.0f);
    const float x2 = xc * xc;
    const float w = xc * fmaf(x2, fmaf(x2, c5, c3), c1);
    return x * rcp_approx(1.0f + ex2_approx(w));
}
// (round 1, kept for reference measurements) tanh-form GELU on two fp16 lanes: 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) with one MUFU.TANH for the pair
// (6 half2 instructions per two elements).  Inputs are the fp32 accumulator + bias rounded to fp16; the result
// is the fp16 operand of the next layer, so every intermediate carries the precision of its consumer.
__device__ __forceinline__ uint32_t gelu_h2(uint32_t xb)
{
    const __half2 x = *reinterpret_cast<const __half2 *>(&xb);
    const __half2 k0 = __floats2half2_rn(0.7978845608f, 0.7978845608f);
    const __half2 k1 = __floats2half2_rn(0.0356774081f, 0.0356774081f);
    const __half2 hf = __floats2half2_rn(0.5f, 0.5f);
    const __half2 u = __hmul2(__hfma2(__hmul2(x, x), k1, k0), x);
    uint32_t tb;
    const uint32_t ub = *reinterpret_cast<const uint32_t *>(&u);
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(tb) : "r"(ub));
    const __half2 t = *reinterpret_cast<const __half2 *>(&tb);
    const __half2 h = __hmul2(x, hf);
    const __half2 g = __hfma2(h, t, h);
    return *reinterpret_cast<const uint32_t *>(&g);
}
__device__ __forceinline__ uint32_t hadd2_u32(uint32_t a, uint32_t b)
{
    const __half2 r = __hadd2(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b));
    return *reinterpret_cast<const uint32_t *>(&r);
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi)
{
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t v)
{
    return __half22float2(*reinterpret_cast<const __half2 *>(&v));
}

// host: byte offset of element (row r, k) inside a K-major SWIZZLE_128B tile of 128-byte rows (k < 64 fp16)
__host__ __device__ inline size_t swz_off(int r, int k)
{
    return (size_t)r * 128 + (size_t)((((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2);
}

}  // namespace tc
