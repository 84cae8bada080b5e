// npd_common.cuh -- shared declarations for libnpd.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "npd.h"

#define NPD_API extern "C" __attribute__((visibility("default")))

// ---- error plumbing ---------------------------------------------------------------------------
void npd_set_error(const char *fmt, ...);

#define NPD_CHECK_CUDA(expr)                                                                  \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            npd_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return NPD_ECUDA;                                                                 \
        }                                                                                     \
    } while (0)

#define NPD_REQUIRE(cond, ...)                                                                \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            npd_set_error(__VA_ARGS__);                                                       \
            return NPD_EINVAL;                                                                \
        }                                                                                     \
    } while (0)

// ---- experiment switches ---------------------------------------------------------------------------
// Kernel-selection / trace / "skip part of the work" switches read from the environment exist only in builds made
// with -DNPD_DEBUG_KNOBS (`make DEBUG_KNOBS=1`, used by tools/exp_*.sh and the trace tools).  The shipped libnpd.so
// never reads the environment: what it launches and what it returns depend on the call's arguments only.
#ifdef NPD_DEBUG_KNOBS
#include <stdlib.h>
static inline const char *npd_knob(const char *name) { return getenv(name); }
#else
static inline const char *npd_knob(const char *) { return nullptr; }
#endif

// ---- code object ------------------------------------------------------------------------------
struct npd_code {
    int n, N, K;
    float infty;
    uint32_t pac_g;     // 0 = plain polar
    int pac_M;          // number of taps incl. the leading one (pac_code.py:101)
    uint32_t pac_taps;  // bit (j-1) set <=> g[j] == -1, j = 1..M-1 (state[j-1] multiplies u)
    int device;
    int sm_count;
    int32_t *d_info;          // [K] sorted info positions
    uint32_t *d_frozen_words; // [max(N/32,1)] bit i = position i frozen
    uint32_t *d_info_words;   // complement of frozen within N
    int32_t *h_info;          // host copy
};

struct DeviceProps {
    int device;
    int sm_count;
    int smem_optin;
    int cc_major, cc_minor;
};
int npd_get_device_props(DeviceProps *p);

// library-owned stream-ordered memory pool that keeps its memory across synchronisations (sc_decode.cu)
int npd_scratch_pool(cudaMemPool_t *out);

// fused-sweep internals (encode_channel.cu / sc_decode.cu, used by count_sweep.cu)
int npd_gen_encode_awgn_bits(const npd_code *code, uint32_t *ubits, float *y, int64_t B, float sigma, uint64_t seed,
                             uint32_t point, uint64_t cw_offset, cudaStream_t st);
int npd_sc_decode_count(const npd_code *code, const float *y, float llr_scale, const uint32_t *ubits, float *decoded_scratch,
                        unsigned char *flags, int64_t B, uint64_t *counts, cudaStream_t st);

// code length and channel count of forward()'s `in4` output of a convNet handle (conv_net.cu)
void npd_conv_dims(const npd_conv *cv, int *N, int *in4_channels);

// ---- device helpers ---------------------------------------------------------------------------
#define NPD_FULL 0xffffffffu

// utils.py:272-275: min(|a|,|b|) * sign(a) * sign(b).  Multiplying by +-1 is exact and a zero operand
// makes the min zero, so the product equals min(|a|,|b|) carrying sign(a) xor sign(b) (a signed zero
// where the reference yields one) -- bit-identical for all finite inputs.
__device__ __forceinline__ float npd_f_minsum(float a, float b)
{
    float m = fminf(fabsf(a), fabsf(b));
    uint32_t sg = (__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u;
    return __uint_as_float(__float_as_uint(m) | sg);
}

// polar.py:414/445: u*a + b with u in {-1,0,+1} carried as (s = "u is -1", z = "u is 0").
// u*a is exact, so (s ? -a : a) + b rounds identically; u = 0 gives (+-0) + b == b.
__device__ __forceinline__ float npd_g_trit(uint32_t s, uint32_t z, float a, float b)
{
    float t = __uint_as_float(__float_as_uint(a) ^ (s << 31));
    float r = t + b;
    return z ? b : r;
}

// ---- Philox4x32-10 (counter-based RNG; restated in oracle/npd_oracle.c) ------------------------
#define NPD_STREAM_MSG 0u
#define NPD_STREAM_NOISE 1u  // + SNR point index

__device__ __forceinline__ uint4 npd_philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        uint4 nxt;
        nxt.x = hi1 ^ c.y ^ k.x;
        nxt.y = lo1;
        nxt.z = hi0 ^ c.w ^ k.y;
        nxt.w = lo0;
        c = nxt;
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// Box-Muller on two 32-bit words -> two N(0,1) samples.  u = (float(r) + 0.5) * 2^-32 in (0,1].
__device__ __forceinline__ float2 npd_box_muller(uint32_t r0, uint32_t r1)
{
    // u = (r + 1/2) 2^-32 as one fused multiply-add (this file is compiled without FMA contraction)
    float u0 = fmaf((float)r0, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    float u1 = fmaf((float)r1, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    // u0 lies in [2^-33, 1], never denormal: lg2.approx.ftz without __logf's denormal pre-scaling; sqrt.approx (one
    // MUFU, max error 1 ulp) instead of the IEEE square root's Newton sequence.  -2 ln u = (-2 ln 2) lg2 u.
    float l2, rad;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u0));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-1.3862943611198906f * l2));
    float sn, cs;
    __sincosf(6.283185307179586f * u1, &sn, &cs);
    return make_float2(rad * cs, rad * sn);
}
