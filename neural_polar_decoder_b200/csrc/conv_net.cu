// conv_net.cu -- convNet one-shot decoder (reference models.py:691-772) as two tcgen05 kernels.
//
// convNet.forward: y[B,64] -> ten dilated k=7 Conv1d + GELU (channels 1->64->...->64->128->128, three
// residual adds) -> flatten [B,8192] -> Linear 8192->256, GELU, 256->64, GELU, 64->64 -> LayerNorm(64).
//
// Kernel 1, conv_stack_kernel (implicit GEMM, no im2col): a persistent CTA decodes 6 codewords at a time as
// two independent groups of 3.  A group's activations live in shared memory as fp16 rows of 64 channels
// (128 B, the K-major SWIZZLE_128B operand row) -- one row per position, codewords 76 rows apart so that the
// 12 zero rows between them are the padding of every dilation.  Conv tap t of a layer with dilation d is
// then ONE accumulating MMA per 128-row tile: D[row, c_out] += A[row + (t-3)d, c_in] * W_t[c_out, c_in]^T with
// A = the same buffer addressed (t-3)d rows further on (the swizzle depends only on the absolute address, so
// a row-shifted descriptor start is legal) and W_t = a pre-swizzled fp16 tile streamed from L2 by
// cp.async.bulk through an mbarrier ring.  Accumulators (2 tiles x <=128 channels per group) sit in TMEM.
// While the 16 epilogue warps turn group A's accumulators into the next layer's operand rows (bias, GELU,
// residual, fp16, swizzled store), the MMA warp runs the same layer for group B, reusing the weight slots
// still resident in the ring (64-channel layers are loaded once per pair of groups).  The last layer's
// epilogue writes the 128-channel rows straight into the A-operand tiles of kernel 2.
//
// Kernel 2, conv_fc_kernel: a persistent GEMM CTA per 128 codewords: [128 x 8192] x W1^T (N = 256,
// 3-stage bulk-copy pipeline), then GELU -> fp16 operand in shared memory -> W2 (N = 64) -> GELU -> W3
// (N = 64) as two more MMA passes, and LayerNorm(64, eps 1e-6) in registers (one thread = one codeword).
//
// Precision: fp16 operands (weights, activations), fp32 accumulation, bias, GELU and LayerNorm.  GELU is an
// erf-accurate fitted form evaluated in fp32 (tc_common.cuh: the tanh form alone costs 3e-3 on the logits of the
// reference-trained checkpoint); parity target is the reference's logits within 1e-2 relative + 2e-3 absolute.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <type_traits>

#include "npd_common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int CN = 64;                     // positions per codeword (max_len = N)
constexpr int CC = 64;                     // embed_dim / 2
constexpr int CV_G = 3;                    // codewords per group
constexpr int CV_HALO = 12;                // 3 * max dilation
constexpr int CV_PITCH = CN + CV_HALO;     // rows from one codeword to the next
constexpr int CV_TILES = 2;                // 128-row MMA tiles per group
constexpr int CV_ROWS = CV_HALO + 128 * CV_TILES + CV_HALO;
constexpr int CV_BUF = CV_ROWS * 128;      // one 64-channel activation buffer
constexpr int CV_SLOT = 16384;             // weight ring slot: 128 rows x 64 k fp16
constexpr int CV_STAGES = 5;
constexpr int CV_CW = 2 * CV_G;            // codewords per CTA pass
constexpr int CV_EPI_WARPS = 16;
constexpr int CV_THREADS = (CV_EPI_WARPS + 2) * 32;
constexpr int CV_LAYERS = 10;
static_assert(CV_G * CV_PITCH <= 128 * CV_TILES, "group does not fit its tiles");
static_assert(CV_BUF % 1024 == 0, "activation buffers must keep the 1024 B swizzle phase");

constexpr int CV_OFF_RING = 0;
constexpr int CV_OFF_BUFS = CV_OFF_RING + CV_STAGES * CV_SLOT;
constexpr int CV_OFF_BIAS = CV_OFF_BUFS + 4 * CV_BUF;
constexpr int CV_OFF_BARS = CV_OFF_BIAS + CV_LAYERS * 128 * 4;
constexpr int CV_SMEM = CV_OFF_BARS + 256;

constexpr int FC_KC = 2 * CN;              // K chunks of 64 over k' = position * 128 + channel
constexpr int FC_F1 = 4 * CN;              // 256
constexpr int FC_STAGES = 3;
constexpr int FC_A_BYTES = 128 * 128;      // 128 codewords x 64 k
constexpr int FC_B_BYTES = FC_F1 * 128;    // 256 outputs x 64 k
constexpr int FC_STAGE = FC_A_BYTES + FC_B_BYTES;
constexpr int FC_OFF_W2 = FC_STAGES * FC_STAGE;
constexpr int FC_W2_BYTES = (FC_F1 / 64) * CN * 128;   // 4 chunks of [64 x 64]
constexpr int FC_OFF_W3 = FC_OFF_W2 + FC_W2_BYTES;
constexpr int FC_W3_BYTES = CN * 128;
constexpr int FC_OFF_CONST = FC_OFF_W3 + FC_W3_BYTES;  // b1[256] b2[64] b3[64] lnw[64] lnb[64]
constexpr int FC_NCONST = FC_F1 + 4 * CN;
constexpr int FC_OFF_BARS = FC_OFF_CONST + FC_NCONST * 4;
constexpr int FC_SMEM = FC_OFF_BARS + 256;
constexpr int FC_OFF_A2 = 0;                            // reuses the pipeline stages after the main loop
constexpr int FC_OFF_A3 = FC_OFF_A2 + (FC_F1 / 64) * FC_A_BYTES;
constexpr int FC_THREADS = 192;
static_assert(FC_OFF_A3 + FC_A_BYTES <= FC_OFF_W2, "A2/A3 must fit in the stage area");

struct LayerDesc {
    int dil, chunks, cout, ksteps, nslots, slot0;
};

struct ConvParams {
    const unsigned char *wpack;  // weight slots, layer after layer
    const float *bias;           // [10][128]
    const float *y;              // [B,64]
    unsigned char *act;          // kernel-2 A operand tiles: [ceil(B/128)][FC_KC][128 rows][128 B]
    float *in4;                  // optional [B,64,64]
    int64_t B, n_pass;
    int dbg;  // bench-only experiments (NPD_CONV_DBG): 1 = no MMAs, 2 = no epilogue math, 4 = no weight copies
    long long *trace;  // bench-only (NPD_CONV_TRACE): clock64 stamps of CTA 0's third pass, [layer][group][4], or null
    LayerDesc layers[CV_LAYERS];
};

// slot k of (layer L, group g): 0 = MMA warp: activations ready, 1 = MMA warp: all slots issued + committed,
// 2 = epilogue warp 0: accumulators seen, 3 = epilogue warp 0: next layer's operand rows written
__device__ __forceinline__ void cv_trace(const ConvParams &p, int64_t ps, int L, int g, int k)
{
    if (p.trace && blockIdx.x == 0 && ps == 2 * (int64_t)gridDim.x && (threadIdx.x & 31) == 0)
        p.trace[(L * 2 + g) * 4 + k] = clock64();
}

// One epilogue block: 32 accumulator columns [c, c+32) of the thread's TMEM lane -> + bias (fp32) -> GELU on fp16
// pairs (-> + residual row, in place) -> four 16-byte chunks of the K-major SWIZZLE_128B operand row `orow`
// (swizzle phase `sw`).  IN4 also stores the fp32 copy that forward() returns as input4.
template <bool RES, bool IN4, bool GLOBAL_OUT>
__device__ __forceinline__ void epi_block(uint32_t taddr, int c, const float *bias, unsigned char *orow, uint32_t sw,
                                          bool store, float *in4)
{
    float v[32];
    tmem_ld32(taddr + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int cc = c + 8 * j;
        const float4 b0 = *reinterpret_cast<const float4 *>(bias + cc);
        const float4 b1 = *reinterpret_cast<const float4 *>(bias + cc + 4);
        uint4 pk;
#ifdef NPD_CONV_GELU_TANH_H2  // round 1: tanh form on fp16 pairs (3e-3 logit error on the trained checkpoint)
        pk.x = gelu_h2(pack_half2(v[8 * j + 0] + b0.x, v[8 * j + 1] + b0.y));
        pk.y = gelu_h2(pack_half2(v[8 * j + 2] + b0.z, v[8 * j + 3] + b0.w));
        pk.z = gelu_h2(pack_half2(v[8 * j + 4] + b1.x, v[8 * j + 5] + b1.y));
        pk.w = gelu_h2(pack_half2(v[8 * j + 6] + b1.z, v[8 * j + 7] + b1.w));
#else  // erf-accurate GELU in fp32 (tc_common.cuh), rounded once to the next layer's fp16 operand
        pk.x = pack_half2(gelu_f(v[8 * j + 0] + b0.x), gelu_f(v[8 * j + 1] + b0.y));
        pk.y = pack_half2(gelu_f(v[8 * j + 2] + b0.z), gelu_f(v[8 * j + 3] + b0.w));
        pk.z = pack_half2(gelu_f(v[8 * j + 4] + b1.x), gelu_f(v[8 * j + 5] + b1.y));
        pk.w = pack_half2(gelu_f(v[8 * j + 6] + b1.z), gelu_f(v[8 * j + 7] + b1.w));
#endif
        uint4 *dst = reinterpret_cast<uint4 *>(orow + ((((cc & 63) >> 3) ^ sw) << 4));
        if (RES) {  // input_{k+1} = layers_k(input_k) + input_k (models.py:748-755)
            const uint4 xr = *dst;
            pk.x = hadd2_u32(pk.x, xr.x); pk.y = hadd2_u32(pk.y, xr.y);
            pk.z = hadd2_u32(pk.z, xr.z); pk.w = hadd2_u32(pk.w, xr.w);
        }
        if (IN4) {
            if (in4 != nullptr) {
                const float2 f0 = unpack_half2(pk.x), f1 = unpack_half2(pk.y), f2 = unpack_half2(pk.z), f3 = unpack_half2(pk.w);
                float *d4 = in4 + (size_t)(8 * j) * CN;
                d4[0] = f0.x; d4[CN] = f0.y; d4[2 * CN] = f1.x; d4[3 * CN] = f1.y;
                d4[4 * CN] = f2.x; d4[5 * CN] = f2.y; d4[6 * CN] = f3.x; d4[7 * CN] = f3.y;
            }
        }
        if (store) *dst = pk;
    }
}

// All MMAs of one weight slot of a layer whose shape is known at compile time (the clock-stamp trace showed the MMA warp
// issuing for 92 % of a pass at 75 cycles per N = 64 MMA against the 48-cycle operand-read floor: runtime layer
// descriptors made every MMA pay descriptor arithmetic and branches).  Here tap, chunk, K step and row tile are
// compile-time, so every descriptor is one add of an immediate to a_in0 / a_in1 / b_slot.
//   COUT 64: two taps per slot (rows 0-63 / 64-127 of the slot); COUT 128: one (tap, 64-channel input chunk) per slot
template <int COUT, int KSTEPS, int CHUNKS, int DIL, int J>
__device__ __forceinline__ void conv_issue_slot(uint32_t d0, uint32_t a_in0, uint32_t a_in1, uint32_t b_slot, uint32_t idesc)
{
    constexpr int NENT = COUT == 64 ? 2 : 1;
#pragma unroll
    for (int e = 0; e < NENT; ++e) {
        constexpr int dummy = 0;
        (void)dummy;
        const int tap = COUT == 64 ? 2 * J + e : (CHUNKS == 1 ? J : J >> 1);
        const int chunk = (COUT == 64 || CHUNKS == 1) ? 0 : (J & 1);
        if (tap >= 7) break;
        const uint32_t a_lo = (chunk ? a_in1 : a_in0) + (uint32_t)(((CV_HALO + (tap - 3) * DIL) * 128) >> 4);
        const uint32_t b_lo = b_slot + (uint32_t)((e * 8192) >> 4);
#pragma unroll
        for (int m = 0; m < CV_TILES; ++m)
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
                umma_f16(d0 + m * 128, umma_desc_from_lo(a_lo + m * 1024 + k * 2), umma_desc_from_lo(b_lo + k * 2), idesc,
                         (J | e | k) ? 1u : 0u);
    }
}

struct ConvIssue {
    uint32_t bar_full, bar_empty, ring_lo;
    uint32_t st, ph;
    bool no_mma;
};

// every slot of one (layer, group): wait for the slot's weights, issue its MMAs, release it unless the other group re-uses it
template <int COUT, int KSTEPS, int CHUNKS, int DIL, int NSLOTS>
__device__ __forceinline__ void conv_issue_group(ConvIssue &c, uint32_t d0, uint32_t a_in0, uint32_t a_in1, bool hold)
{
    constexpr uint32_t idesc = umma_idesc_f16(128, COUT);
    auto slot = [&](auto jc) {
        constexpr int J = decltype(jc)::value;
        mbar_wait(c.bar_full + 8 * c.st, c.ph);
        tc_fence_after();
        if (elect_one()) {
            if (!c.no_mma) conv_issue_slot<COUT, KSTEPS, CHUNKS, DIL, J>(d0, a_in0, a_in1, c.ring_lo + c.st * (CV_SLOT >> 4), idesc);
            if (!hold) umma_commit(c.bar_empty + 8 * c.st);
        }
        __syncwarp();
        if (++c.st == CV_STAGES) { c.st = 0; c.ph ^= 1; }
    };
    // NSLOTS is 4, 7 or 14
    slot(std::integral_constant<int, 0>{}); slot(std::integral_constant<int, 1>{}); slot(std::integral_constant<int, 2>{});
    slot(std::integral_constant<int, 3>{});
    if constexpr (NSLOTS > 4) {
        slot(std::integral_constant<int, 4>{}); slot(std::integral_constant<int, 5>{}); slot(std::integral_constant<int, 6>{});
    }
    if constexpr (NSLOTS > 7) {
        slot(std::integral_constant<int, 7>{}); slot(std::integral_constant<int, 8>{}); slot(std::integral_constant<int, 9>{});
        slot(std::integral_constant<int, 10>{}); slot(std::integral_constant<int, 11>{}); slot(std::integral_constant<int, 12>{});
        slot(std::integral_constant<int, 13>{});
    }
}

__global__ void __launch_bounds__(CV_THREADS, 1) conv_stack_kernel(const ConvParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char *s_bufs = smem + CV_OFF_BUFS;
    float *s_bias = reinterpret_cast<float *>(smem + CV_OFF_BIAS);
    uint64_t *s_bars = reinterpret_cast<uint64_t *>(smem + CV_OFF_BARS);
    const uint32_t bar_full = smem_u32(s_bars), bar_empty = bar_full + 8 * CV_STAGES,
                   bar_acc = bar_empty + 8 * CV_STAGES, bar_act = bar_acc + 16;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bars + 2 * CV_STAGES + 4);

    if (tid == 0) {
        for (int i = 0; i < CV_STAGES; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(bar_acc + 8 * g, 1);
            mbar_init(bar_act + 8 * g, CV_EPI_WARPS * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == CV_EPI_WARPS + 1) tmem_alloc(smem_u32(s_tmem), 512);
    for (int i = tid; i < 4 * CV_BUF / 16; i += CV_THREADS) reinterpret_cast<uint4 *>(s_bufs)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < CV_LAYERS * 128; i += CV_THREADS) s_bias[i] = p.bias[i];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == CV_EPI_WARPS) {
        // ================= producer: weight slots in consumption order =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t ps = blockIdx.x; ps < p.n_pass; ps += gridDim.x)
                for (int L = 0; L < CV_LAYERS; ++L) {
                    const LayerDesc ld = p.layers[L];
                    // 64-channel layers fit the ring: both groups use the same residency
                    const int loads = ld.cout == 64 ? 1 : 2;
                    for (int rep = 0; rep < loads; ++rep) {
                        const unsigned char *src = p.wpack + (size_t)ld.slot0 * CV_SLOT;
                        for (int j = 0; j < ld.nslots; ++j, src += CV_SLOT) {
                            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                            if (p.dbg & 4) {
                                mbar_arrive(bar_full + 8 * stage);
                            } else {
                                mbar_expect_tx(bar_full + 8 * stage, CV_SLOT);
                                bulk_g2s(smem_u32(smem + CV_OFF_RING + stage * CV_SLOT), src, CV_SLOT, bar_full + 8 * stage);
                            }
                            if (++stage == CV_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
        }
    } else if (warp == CV_EPI_WARPS + 1) {
        // ================= MMA issuer (warp-uniform schedule, one elected lane issues) =================
        // layer shapes are compile-time (models.py:701-730: (C_out, C_in, dilation) = kConvShape); the ring stage, the
        // activation buffers of the group and the TMEM base are the only runtime terms of a descriptor
        const uint32_t bufs0 = umma_desc_lo(smem_u32(s_bufs));
        constexpr uint32_t BUF_LO = CV_BUF >> 4;
        ConvIssue c;
        c.bar_full = bar_full; c.bar_empty = bar_empty; c.ring_lo = umma_desc_lo(smem_u32(smem + CV_OFF_RING));
        c.st = 0; c.ph = 0; c.no_mma = (p.dbg & 1) != 0;
        uint32_t nact0 = 0, nact1 = 0;
        for (int64_t ps = blockIdx.x; ps < p.n_pass; ps += gridDim.x) {
#pragma unroll
            for (int L = 0; L < CV_LAYERS; ++L) {
                uint32_t st0 = c.st, ph0 = c.ph;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const bool hold = (L < 8) && g == 0;  // 64-channel layers: group 1 re-uses and then releases the slots
                    uint32_t &nact = g ? nact1 : nact0;
                    mbar_wait(bar_act + 8 * g, nact & 1);
                    ++nact;
                    tc_fence_after();
                    cv_trace(p, ps, L, g, 0);
                    const uint32_t bx = bufs0 + (g * 2) * BUF_LO, bt = bx + BUF_LO;
                    const uint32_t in0 = (L & 1) ? bt : bx, in1 = bx;
                    const uint32_t d0 = tmem_base + (uint32_t)(g * 2) * 128;
                    if (L < 8 && g == 1) { c.st = st0; c.ph = ph0; }  // the slots are still resident from group 0
                    if (L == 0) conv_issue_group<64, 1, 1, 1, 4>(c, d0, in0, in1, hold);
                    else if (L == 1 || L == 4 || L == 7) conv_issue_group<64, 4, 1, 2, 4>(c, d0, in0, in1, hold);
                    else if (L == 2 || L == 5) conv_issue_group<64, 4, 1, 4, 4>(c, d0, in0, in1, hold);
                    else if (L == 3 || L == 6) conv_issue_group<64, 4, 1, 1, 4>(c, d0, in0, in1, hold);
                    else if (L == 8) conv_issue_group<128, 4, 1, 4, 7>(c, d0, in0, in1, hold);
                    else conv_issue_group<128, 4, 2, 1, 14>(c, d0, in0, in1, hold);
                    if (elect_one()) umma_commit(bar_acc + 8 * g);
                    __syncwarp();
                    cv_trace(p, ps, L, g, 1);
                }
            }
        }
    } else {
        // ================= epilogue warps: one thread = one row (position) of one tile, half its channels ===========
        const int q = warp & 3, m = (warp >> 2) & 1, hc = warp >> 3;
        const int o = m * 128 + q * 32 + lane;
        const int gi = o / CV_PITCH, l = o - gi * CV_PITCH;
        const bool in_cw = gi < CV_G && l < CN;
        const int row = CV_HALO + o;
        const uint32_t sw = row & 7;
        uint32_t nacc0 = 0, nacc1 = 0;
        for (int64_t ps = blockIdx.x; ps < p.n_pass; ps += gridDim.x) {
            // ---- layer-1 input: y as channel 0 of the row (k-step 0 = chunks 0 and 1) ----
            for (int g = 0; g < 2; ++g) {
                const int64_t cw = ps * CV_CW + g * CV_G + gi;
                if (in_cw && hc == 0) {
                    const float yv = cw < p.B ? p.y[cw * CN + l] : 0.0f;
                    unsigned char *xr = s_bufs + (g * 2) * CV_BUF + row * 128;
                    *reinterpret_cast<uint4 *>(xr + ((0 ^ sw) << 4)) = make_uint4(pack_half2(yv, 0.0f), 0, 0, 0);
                    *reinterpret_cast<uint4 *>(xr + ((1 ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
                }
                fence_async_smem();
                tc_fence_before();
                mbar_arrive(bar_act + 8 * g);
            }
            for (int L = 0; L < CV_LAYERS; ++L) {
                const float *bias = s_bias + L * 128;
                for (int g = 0; g < 2; ++g) {
                    uint32_t &nacc = g ? nacc1 : nacc0;
                    mbar_wait(bar_acc + 8 * g, nacc & 1);
                    ++nacc;
                    tc_fence_after();
                    if (warp == 0) cv_trace(p, ps, L, g, 2);
                    const int64_t cw = ps * CV_CW + g * CV_G + gi;
                    const bool valid = in_cw && cw < p.B;
                    unsigned char *bx = s_bufs + (g * 2) * CV_BUF, *bt = bx + CV_BUF;
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 2 + m) * 128;
                    if (p.dbg & 2) {
                    } else if (L < 8) {
                        // 64 channels -> the other buffer (odd layers) or, with the residual, in place (even layers)
                        unsigned char *orow = ((L & 1) ? bx : bt) + row * 128;
                        const int c = hc * 32;
                        if (L == 3 || L == 7 || (L == 5 && p.in4 == nullptr))
                            epi_block<true, false, false>(taddr, c, bias, orow, sw, in_cw, nullptr);
                        else if (L == 5)
                            epi_block<true, true, false>(taddr, c, bias, orow, sw, in_cw, valid ? p.in4 + (cw * CC + c) * CN + l : nullptr);
                        else
                            epi_block<false, false, false>(taddr, c, bias, orow, sw, in_cw, nullptr);
                    } else if (L == 8) {
                        // 128 channels: 0-63 -> T, 64-127 -> X (free once this layer's MMAs have read it)
                        unsigned char *orow = (hc ? bx : bt) + row * 128;
                        epi_block<false, false, false>(taddr, hc * 64, bias, orow, sw, in_cw, nullptr);
                        epi_block<false, false, false>(taddr, hc * 64 + 32, bias, orow, sw, in_cw, nullptr);
                    } else {
                        // flatten + Linear operand: tile (cw / 128, k chunk 2 l + c / 64), row cw % 128
                        unsigned char *orow = p.act + ((size_t)(cw >> 7) * FC_KC + (size_t)(2 * l + hc)) * FC_A_BYTES + (size_t)(cw & 127) * 128;
                        epi_block<false, false, true>(taddr, hc * 64, bias, orow, (uint32_t)(cw & 7), valid, nullptr);
                        epi_block<false, false, true>(taddr, hc * 64 + 32, bias, orow, (uint32_t)(cw & 7), valid, nullptr);
                    }
                    tc_fence_before();
                    if (L < CV_LAYERS - 1) {
                        fence_async_smem();
                        mbar_arrive(bar_act + 8 * g);
                    }
                    if (warp == 0) cv_trace(p, ps, L, g, 3);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == CV_EPI_WARPS + 1) tmem_dealloc(tmem_base, 512);
}

struct FcParams {
    const unsigned char *act, *w1, *w23;
    const float *consts;  // b1[256] b2[64] b3[64] lnw[64] lnb[64]
    float *logits;        // [B,64]
    int sign_out;         // store sign(logit) instead of the logit (convNet.decode, models.py:769-772)
    int64_t B, n_tiles;
};

__global__ void __launch_bounds__(FC_THREADS, 1) conv_fc_kernel(const FcParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *s_const = reinterpret_cast<float *>(smem + FC_OFF_CONST);
    uint64_t *s_bars = reinterpret_cast<uint64_t *>(smem + FC_OFF_BARS);
    const uint32_t bar_full = smem_u32(s_bars), bar_empty = bar_full + 8 * FC_STAGES, bar_w = bar_empty + 8 * FC_STAGES,
                   bar_d1 = bar_w + 8, bar_a2 = bar_d1 + 8, bar_d2 = bar_a2 + 8, bar_a3 = bar_d2 + 8, bar_d3 = bar_a3 + 8;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bars + 2 * FC_STAGES + 6);

    if (tid == 0) {
        for (int i = 0; i < FC_STAGES; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        mbar_init(bar_w, 1);
        mbar_init(bar_d1, 1);
        mbar_init(bar_a2, 128);
        mbar_init(bar_d2, 1);
        mbar_init(bar_a3, 128);
        mbar_init(bar_d3, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) tmem_alloc(smem_u32(s_tmem), 512);
    for (int i = tid; i < FC_NCONST; i += FC_THREADS) s_const[i] = p.consts[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t d1 = tmem_base, d2 = tmem_base + 256, d3 = tmem_base + 320;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(bar_w, FC_W2_BYTES + FC_W3_BYTES);
            bulk_g2s(smem_u32(smem + FC_OFF_W2), p.w23, FC_W2_BYTES + FC_W3_BYTES, bar_w);
            uint32_t stage = 0, phase = 0, it = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                // the stage area doubles as the A2/A3 operands of the previous tile until its last MMA is done
                if (it > 0) mbar_wait(bar_d3, (it - 1) & 1);
                const unsigned char *a = p.act + (size_t)tile * FC_KC * FC_A_BYTES;
                const unsigned char *b = p.w1;
                for (int kc = 0; kc < FC_KC; ++kc, a += FC_A_BYTES, b += FC_B_BYTES) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    mbar_expect_tx(bar_full + 8 * stage, FC_STAGE);
                    const uint32_t dst = smem_u32(smem + stage * FC_STAGE);
                    bulk_g2s(dst, a, FC_A_BYTES, bar_full + 8 * stage);
                    bulk_g2s(dst + FC_A_BYTES, b, FC_B_BYTES, bar_full + 8 * stage);
                    if (++stage == FC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        const uint32_t s0 = smem_u32(smem);
        const uint32_t idesc1 = umma_idesc_f16(128, FC_F1), idesc2 = umma_idesc_f16(128, CN);
        uint32_t stage = 0, phase = 0, it = 0;
        mbar_wait(bar_w, 0);
        for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            for (int kc = 0; kc < FC_KC; ++kc) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a = s0 + stage * FC_STAGE, b = a + FC_A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(d1, umma_desc(a + k * 32), umma_desc(b + k * 32), idesc1, (kc | k) ? 1u : 0u);
                    umma_commit(bar_empty + 8 * stage);
                }
                __syncwarp();
                if (++stage == FC_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(bar_d1);
            __syncwarp();
            // ---- Linear(256 -> 64) on the GELU'd fp16 rows the epilogue warps just wrote ----
            mbar_wait(bar_a2, it & 1);
            tc_fence_after();
            if (elect_one()) {
                for (int ch = 0; ch < FC_F1 / 64; ++ch)
                    for (int k = 0; k < 4; ++k)
                        umma_f16(d2, umma_desc(s0 + FC_OFF_A2 + ch * FC_A_BYTES + k * 32),
                                 umma_desc(s0 + FC_OFF_W2 + ch * (CN * 128) + k * 32), idesc2, (ch | k) ? 1u : 0u);
                umma_commit(bar_d2);
            }
            __syncwarp();
            // ---- Linear(64 -> 64) ----
            mbar_wait(bar_a3, it & 1);
            tc_fence_after();
            if (elect_one()) {
                for (int k = 0; k < 4; ++k)
                    umma_f16(d3, umma_desc(s0 + FC_OFF_A3 + k * 32), umma_desc(s0 + FC_OFF_W3 + k * 32), idesc2, k ? 1u : 0u);
                umma_commit(bar_d3);
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue warps: one thread = one codeword =================
        const int row = warp * 32 + lane;
        const uint32_t sw = row & 7;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        const float *b1 = s_const, *b2 = s_const + FC_F1, *b3 = b2 + CN, *lnw = b3 + CN, *lnb = lnw + CN;
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const int64_t cw = tile * 128 + row;
            mbar_wait(bar_d1, it & 1);
            tc_fence_after();
            for (int c0 = 0; c0 < FC_F1; c0 += 32) {
                float v[32];
                tmem_ld32(d1 + lane_addr + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = c0 + 8 * j;
                    float r[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) r[i] = gelu_f(v[8 * j + i] + b1[c + i]);
                    *reinterpret_cast<uint4 *>(smem + FC_OFF_A2 + (c >> 6) * FC_A_BYTES + row * 128 + ((((c & 63) >> 3) ^ sw) << 4)) =
                        make_uint4(pack_half2(r[0], r[1]), pack_half2(r[2], r[3]), pack_half2(r[4], r[5]), pack_half2(r[6], r[7]));
                }
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(bar_a2);

            mbar_wait(bar_d2, it & 1);
            tc_fence_after();
            for (int c0 = 0; c0 < CN; c0 += 32) {
                float v[32];
                tmem_ld32(d2 + lane_addr + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = c0 + 8 * j;
                    float r[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) r[i] = gelu_f(v[8 * j + i] + b2[c + i]);
                    *reinterpret_cast<uint4 *>(smem + FC_OFF_A3 + row * 128 + (((c >> 3) ^ sw) << 4)) =
                        make_uint4(pack_half2(r[0], r[1]), pack_half2(r[2], r[3]), pack_half2(r[4], r[5]), pack_half2(r[6], r[7]));
                }
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(bar_a3);

            // ---- LayerNorm(64, eps = 1e-6) over the thread's own row (models.py:739, 760) ----
            mbar_wait(bar_d3, it & 1);
            tc_fence_after();
            float x0[32], x1[32];
            tmem_ld32(d3 + lane_addr, x0);
            tmem_ld32(d3 + lane_addr + 32, x1);
            tmem_ld_wait();
            tc_fence_before();
            float mean = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                x0[i] += b3[i];
                x1[i] += b3[32 + i];
                mean += x0[i] + x1[i];
            }
            mean *= (1.0f / CN);
            float var = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                x0[i] -= mean;
                x1[i] -= mean;
                var += x0[i] * x0[i] + x1[i] * x1[i];
            }
            const float rstd = rsqrtf(var * (1.0f / CN) + 1e-6f);
            if (cw < p.B) {
                float4 *dst = reinterpret_cast<float4 *>(p.logits + cw * CN);
                const bool sg = p.sign_out != 0;
                auto out = [&](float v) { return sg ? (float)((v > 0.0f) - (v < 0.0f)) : v; };  // torch.sign
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    dst[i / 4] = make_float4(out(x0[i] * rstd * lnw[i] + lnb[i]), out(x0[i + 1] * rstd * lnw[i + 1] + lnb[i + 1]),
                                             out(x0[i + 2] * rstd * lnw[i + 2] + lnb[i + 2]), out(x0[i + 3] * rstd * lnw[i + 3] + lnb[i + 3]));
                    dst[8 + i / 4] = make_float4(out(x1[i] * rstd * lnw[32 + i] + lnb[32 + i]), out(x1[i + 1] * rstd * lnw[33 + i] + lnb[33 + i]),
                                                 out(x1[i + 2] * rstd * lnw[34 + i] + lnb[34 + i]), out(x1[i + 3] * rstd * lnw[35 + i] + lnb[35 + i]));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// ---- host side ----------------------------------------------------------------------------------
struct npd_conv {
    int N, embed_dim, sm_count;
    unsigned char *d_wpack, *d_w1, *d_w23;
    float *d_bias, *d_consts;
    LayerDesc layers[CV_LAYERS];
};

namespace {

unsigned short h16(float f)
{
    __half h = __float2half_rn(f);
    return *reinterpret_cast<unsigned short *>(&h);
}

// reference models.py:701-730: (C_out, C_in, dilation) of the ten Conv1d layers
const int kConvShape[CV_LAYERS][3] = {{64, 1, 1},  {64, 64, 2}, {64, 64, 4}, {64, 64, 1},  {64, 64, 2},
                                      {64, 64, 4}, {64, 64, 1}, {64, 64, 2}, {128, 64, 4}, {128, 128, 1}};

}  // namespace

NPD_API int npd_conv_create(int N, int embed_dim, const float *h_params, size_t n_params, npd_conv_t **out)
{
    NPD_REQUIRE(out, "npd_conv_create: null out");
    *out = nullptr;
    NPD_REQUIRE(h_params, "npd_conv_create: null parameter blob");
    if (N != CN || embed_dim != 2 * CC) {
        npd_set_error("npd_conv_create: supported envelope is N = 64, embed_dim = 128 (got N=%d embed_dim=%d)", N, embed_dim);
        return NPD_EUNSUPPORTED;
    }
    size_t need = 0;
    for (int L = 0; L < CV_LAYERS; ++L) need += (size_t)kConvShape[L][0] * kConvShape[L][1] * 7 + kConvShape[L][0];
    need += (size_t)FC_F1 * (2 * CC * CN) + FC_F1 + (size_t)CN * FC_F1 + CN + (size_t)CN * CN + CN + 2 * CN;
    NPD_REQUIRE(n_params == need, "npd_conv_create: parameter blob has %zu floats, expected %zu", n_params, need);
    for (size_t i = 0; i < n_params; ++i)
        if (!(fabsf(h_params[i]) <= 65504.0f)) {
            npd_set_error("npd_conv_create: parameter %zu = %g is outside the fp16 range", i, (double)h_params[i]);
            return NPD_EUNSUPPORTED;
        }
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    if ((size_t)dp.smem_optin < (size_t)CV_SMEM || (size_t)dp.smem_optin < (size_t)FC_SMEM) {
        npd_set_error("npd_conv_create: needs %d B of shared memory per block (limit %d)", CV_SMEM > FC_SMEM ? CV_SMEM : FC_SMEM, dp.smem_optin);
        return NPD_EUNSUPPORTED;
    }

    npd_conv *cv = (npd_conv *)calloc(1, sizeof(npd_conv));
    if (!cv) return NPD_ENOMEM;
    cv->N = N;
    cv->embed_dim = embed_dim;
    cv->sm_count = dp.sm_count;

    // ---- conv weights -> ring slots of 128 rows x 64 k (fp16, K-major SWIZZLE_128B) ----
    std::vector<unsigned short> wpack;
    std::vector<float> bias((size_t)CV_LAYERS * 128, 0.0f);
    const float *src = h_params;
    int slot = 0;
    for (int L = 0; L < CV_LAYERS; ++L) {
        const int co = kConvShape[L][0], ci = kConvShape[L][1], dil = kConvShape[L][2];
        const float *W = src;            // [co][ci][7]
        const float *b = src + (size_t)co * ci * 7;
        src = b + co;
        for (int c = 0; c < co; ++c) bias[(size_t)L * 128 + c] = b[c];
        LayerDesc &ld = cv->layers[L];
        ld.dil = dil;
        ld.cout = co;
        ld.chunks = ci > 64 ? 2 : 1;
        ld.ksteps = ci >= 64 ? 4 : 1;
        ld.slot0 = slot;
        ld.nslots = co == 64 ? 4 : 7 * ld.chunks;
        wpack.resize((size_t)(slot + ld.nslots) * (CV_SLOT / 2), 0);
        for (int j = 0; j < ld.nslots; ++j) {
            unsigned short *dst = wpack.data() + (size_t)(slot + j) * (CV_SLOT / 2);
            for (int r = 0; r < 128; ++r) {
                int tap, c, chunk = 0;
                if (co == 64) { tap = 2 * j + (r >> 6); c = r & 63; }        // two taps per slot
                else if (ld.chunks == 1) { tap = j; c = r; }
                else { tap = j >> 1; chunk = j & 1; c = r; }
                if (tap >= 7) continue;
                for (int k = 0; k < 64; ++k) {
                    const int cin = chunk * 64 + k;
                    if (cin >= ci) break;
                    dst[swz_off(r, k) / 2] = h16(W[((size_t)c * ci + cin) * 7 + tap]);
                }
            }
        }
        slot += ld.nslots;
    }
    // ---- Linear layers ----
    const float *W1 = src, *B1 = W1 + (size_t)FC_F1 * (2 * CC * CN);
    const float *W2 = B1 + FC_F1, *B2 = W2 + (size_t)CN * FC_F1;
    const float *W3 = B2 + CN, *B3 = W3 + (size_t)CN * CN;
    const float *LNW = B3 + CN, *LNB = LNW + CN;
    std::vector<unsigned short> w1((size_t)FC_KC * (FC_B_BYTES / 2));
    for (int kc = 0; kc < FC_KC; ++kc)
        for (int j = 0; j < FC_F1; ++j)
            for (int k = 0; k < 64; ++k) {
                // kernel K order k' = position * 128 + channel; reference flatten order = channel * N + position
                const int l = kc >> 1, c = (kc & 1) * 64 + k;
                w1[(size_t)kc * (FC_B_BYTES / 2) + swz_off(j, k) / 2] = h16(W1[(size_t)j * (2 * CC * CN) + (size_t)c * CN + l]);
            }
    std::vector<unsigned short> w23((FC_W2_BYTES + FC_W3_BYTES) / 2);
    for (int ch = 0; ch < FC_F1 / 64; ++ch)
        for (int j = 0; j < CN; ++j)
            for (int k = 0; k < 64; ++k) w23[(size_t)ch * (CN * 64) + swz_off(j, k) / 2] = h16(W2[(size_t)j * FC_F1 + ch * 64 + k]);
    for (int j = 0; j < CN; ++j)
        for (int k = 0; k < 64; ++k) w23[FC_W2_BYTES / 2 + swz_off(j, k) / 2] = h16(W3[(size_t)j * CN + k]);
    std::vector<float> consts(FC_NCONST);
    memcpy(consts.data(), B1, FC_F1 * 4);
    memcpy(consts.data() + FC_F1, B2, CN * 4);
    memcpy(consts.data() + FC_F1 + CN, B3, CN * 4);
    memcpy(consts.data() + FC_F1 + 2 * CN, LNW, CN * 4);
    memcpy(consts.data() + FC_F1 + 3 * CN, LNB, CN * 4);

    cudaError_t e = cudaMalloc(&cv->d_wpack, wpack.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&cv->d_w1, w1.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&cv->d_w23, w23.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&cv->d_bias, bias.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&cv->d_consts, consts.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(cv->d_wpack, wpack.data(), wpack.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(cv->d_w1, w1.data(), w1.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(cv->d_w23, w23.data(), w23.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(cv->d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(cv->d_consts, consts.data(), consts.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_fc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM);
    if (e != cudaSuccess) {
        npd_set_error("npd_conv_create: %s", cudaGetErrorString(e));
        npd_conv_destroy(cv);
        return NPD_ECUDA;
    }
    *out = cv;
    return NPD_OK;
}

void npd_conv_dims(const npd_conv *cv, int *N, int *in4_channels)
{
    *N = cv->N;
    *in4_channels = CC;
}

NPD_API int npd_conv_destroy(npd_conv_t *cv)
{
    if (!cv) return NPD_OK;
    cudaFree(cv->d_wpack);
    cudaFree(cv->d_w1);
    cudaFree(cv->d_w23);
    cudaFree(cv->d_bias);
    cudaFree(cv->d_consts);
    free(cv);
    return NPD_OK;
}

// fp16 rows of the flattened last conv activation: 128 codewords x 8192 x 2 B per tile, at most 2^16 codewords
// (1 GiB) in flight; larger batches are processed in chunks of that size
NPD_API size_t npd_conv_workspace_bytes(const npd_conv_t *cv, int64_t B)
{
    if (!cv || B <= 0) return 0;
    int64_t tiles = (B + 127) / 128;
    if (tiles > 512) tiles = 512;
    return (size_t)tiles * FC_KC * FC_A_BYTES;
}

static int conv_forward_impl(const npd_conv_t *cv, const float *y, float *logits, float *in4, int64_t B, void *workspace,
                             size_t workspace_bytes, void *stream, int sign_out);

NPD_API int npd_conv_forward(const npd_conv_t *cv, const float *y, float *logits, float *in4, int64_t B,
                             void *workspace, size_t workspace_bytes, void *stream)
{
    return conv_forward_impl(cv, y, logits, in4, B, workspace, workspace_bytes, stream, 0);
}

NPD_API int npd_conv_decode(const npd_conv_t *cv, const float *y, float *bits, int64_t B, void *workspace,
                            size_t workspace_bytes, void *stream)
{
    return conv_forward_impl(cv, y, bits, nullptr, B, workspace, workspace_bytes, stream, 1);
}

static int conv_forward_impl(const npd_conv_t *cv, const float *y, float *logits, float *in4, int64_t B, void *workspace,
                             size_t workspace_bytes, void *stream, int sign_out)
{
    NPD_REQUIRE(cv && y && logits, "npd_conv_forward: null argument");
    NPD_REQUIRE(B >= 0, "npd_conv_forward: negative batch");
    if (B == 0) return NPD_OK;
    const size_t tile_bytes = (size_t)FC_KC * FC_A_BYTES;
    NPD_REQUIRE(workspace && workspace_bytes >= tile_bytes,
                "npd_conv_forward: workspace of %zu B is smaller than one 128-codeword tile (%zu B); see npd_conv_workspace_bytes",
                workspace_bytes, tile_bytes);
    const int64_t chunk = (int64_t)(workspace_bytes / tile_bytes) * 128;
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t nb = B - b0 < chunk ? B - b0 : chunk;
        ConvParams cp{};
        cp.wpack = cv->d_wpack; cp.bias = cv->d_bias; cp.y = y + b0 * CN; cp.act = (unsigned char *)workspace;
        cp.in4 = in4 ? in4 + b0 * CC * CN : nullptr;
        cp.B = nb; cp.n_pass = (nb + CV_CW - 1) / CV_CW;
        memcpy(cp.layers, cv->layers, sizeof(cp.layers));
        { const char *d = npd_knob("NPD_CONV_DBG"); cp.dbg = d ? atoi(d) : 0; }
        const char *trace_path = (b0 == 0) ? npd_knob("NPD_CONV_TRACE") : nullptr;  // bench-only (synchronises!)
        if (trace_path) {
            NPD_CHECK_CUDA(cudaMalloc(&cp.trace, sizeof(long long) * CV_LAYERS * 8));
            NPD_CHECK_CUDA(cudaMemsetAsync(cp.trace, 0, sizeof(long long) * CV_LAYERS * 8, st));
        }
        const unsigned g1 = (unsigned)(cp.n_pass < cv->sm_count ? cp.n_pass : cv->sm_count);
        conv_stack_kernel<<<g1, CV_THREADS, CV_SMEM, st>>>(cp);
        NPD_CHECK_CUDA(cudaGetLastError());
        if (trace_path) {
            long long h[CV_LAYERS * 8];
            NPD_CHECK_CUDA(cudaStreamSynchronize(st));
            NPD_CHECK_CUDA(cudaMemcpy(h, cp.trace, sizeof(h), cudaMemcpyDeviceToHost));
            cudaFree(cp.trace);
            if (FILE *f = fopen(trace_path, "w")) {
                for (int i = 0; i < CV_LAYERS * 2; ++i) fprintf(f, "%lld %lld %lld %lld\n", h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
                fclose(f);
            }
        }
        FcParams fp{};
        fp.act = (const unsigned char *)workspace; fp.w1 = cv->d_w1; fp.w23 = cv->d_w23; fp.consts = cv->d_consts;
        fp.logits = logits + b0 * CN; fp.sign_out = sign_out; fp.B = nb; fp.n_tiles = (nb + 127) / 128;
        const unsigned g2 = (unsigned)(fp.n_tiles < cv->sm_count ? fp.n_tiles : cv->sm_count);
        conv_fc_kernel<<<g2, FC_THREADS, FC_SMEM, st>>>(fp);
        NPD_CHECK_CUDA(cudaGetLastError());
    }
    return NPD_OK;
}
