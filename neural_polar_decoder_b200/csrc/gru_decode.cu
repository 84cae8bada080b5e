// gru_decode.cu -- CRISP GRU sequential decoder as one persistent tcgen05 kernel per codeword tile.
//
// Replaces RNN_decoder.decode(net, False, y) (reference rnn_all.py:514-521, 532-547; decoding_type
// 'y_input', onehot) with RNN_Model.forward (387-398) = 2-layer nn.GRU + Linear(H,1) head, for all N
// autoregressive steps in ONE launch.  PyTorch GRU cell, gate order r,z,n:
//     r = s(W_ir x + b_ir + W_hr h + b_hr)       z = s(W_iz x + b_iz + W_hz h + b_hz)
//     n = tanh(W_in x + b_in + r * (W_hn h + b_hn))       h' = (1 - z) * n + z * h
// Layer-0 input x = [y (N) | onehot(prev decision) (2)]: the y part is an MMA over K = N (padded to
// 64), the one-hot part selects one of two weight columns and is added in the epilogue.
//
// Mapping (DESIGN.md "GRU decoder"): a CTA owns TILE_B = 64 codewords for the whole decode.  GEMMs
// are computed transposed, D[gate unit, codeword] = W[gate unit, k] * h[codeword, k]^T, so that
//   * the weights are the A operand: 128-row x 64-k fp16 tiles (16 KB, pre-swizzled on the host into
//     the canonical K-major SWIZZLE_128B layout) streamed from L2 by cp.async.bulk through an mbarrier
//     ring, in exactly the order the MMA warp consumes them (a 4.8 MB "program" per step),
//   * the hidden states are the B operand: h0, h1 and y live in shared memory as fp16
//     [64 codewords x K] K-major swizzled tiles for the whole kernel and are rewritten in place by the
//     epilogue warps (generic stores + fence.proxy.async),
//   * accumulators live in TMEM: one "job" = 128 hidden units x 4 accumulators (r, z, n_i, n_h) x 64
//     codewords = 256 columns; two jobs are in flight (512 columns) so the MMAs of job k+1 overlap the
//     gate math of job k,
//   * every epilogue thread owns one hidden unit (TMEM lane) and 32 codewords (columns): biases and
//     the one-hot columns are per-thread constants, the GRU cell is thread-local, the new state of a
//     layer is staged in registers until the layer's last MMA has read the old state, and the head
//     dot product is a butterfly reduction across lanes.
// Warp roles: warps 0-15 epilogue (TMEM lane quarter = warp % 4, 16-codeword column quarter = warp / 4), warp 16 =
// bulk copy producer, warp 17 = MMA issuer + TMEM allocator.
#include <cuda.h>  // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint; libcuda is not linked)
#include <cuda_fp16.h>

#include <vector>

#include "npd_common.cuh"

namespace {

constexpr int TILE_B = 64;         // codewords per CTA
constexpr int JOB_UNITS = 128;     // hidden units per job (UMMA M)
constexpr int A_TILE_BYTES = 128 * 128;  // 128 rows x 64 fp16
constexpr int B_CHUNK_BYTES = TILE_B * 128;  // 64 rows x 64 fp16
constexpr int NUM_STAGES = 6;  // 96 KB weight ring: the static MMA schedule needs 3*KH tiles to be whole ring passes
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int NUM_PRODUCERS = 1;  // bulk-copy producer warps (a second one changes nothing: the copies are not the limiter)
constexpr int MMA_WARP = EPI_WARPS + NUM_PRODUCERS;
constexpr int NUM_THREADS = (MMA_WARP + 1) * 32;
constexpr int CW_PER_THREAD = TILE_B / (EPI_WARPS / 4);  // 16 codewords (accumulator columns) per epilogue thread

struct GruParams {
    const unsigned char *wpack;  // tiles_per_step * 16 KB, in consumption order
    const float *w_iyT;          // [N][3H] fp32: the y columns of weight_ih_l0, transposed (hoisted input projection)
    const unsigned char *wpack2; // CTA-pair kernel: [2 ranks][tiles_per_step2] 16 KB half-tiles, or null
    int tiles_per_step2;
    int use_tmap;                // pair kernel: weight halves arrive by 2-SM tensor copies that signal the leader directly
    int quad;                    // pair kernel: clusters of 4 = two pairs that share every half-tile by multicast
    const float *consts0;        // [H][12]: b_r b_z b_in b_hn cr0 cr1 cz0 cz1 cn0 cn1 - -
    const float *consts1;        // [H][4] : b_r b_z b_in b_hn
    const float *w_out;          // [H]
    float b_out;
    const float *y;              // [B,N]
    const float *forced;         // [B,N] or null
    const float *genie;          // [B,N] or null: decoded starts as this tensor (rnn_all.py:521-522)
    const float *h0;             // [2][B][H] or null: initial hidden state (decoding_type 'y_h0', rnn_all.py:523-524)
    // MLP head (out_linear_depth > 1, rnn_all.py:335-343): Linear(H,Yh) SELU [Linear(Yh,Yh) SELU]* Linear(Yh,1).
    // Single-CTA kernel only; head_depth <= 1 means the plain Linear(H,1) head (w_out / b_out).
    int head_depth, head_yh;
    const __half *head_w1;       // [Yh][H]
    const __half *head_wh;       // [head_depth - 2][Yh][Yh]
    const float *head_b;         // [head_depth - 1][Yh]
    const float *head_wl;        // [Yh]; its bias is b_out
    __half *head_act;            // workspace: [CTA][2][64][Yh] activations (L2-resident ping-pong)
    signed char *h_lo;           // pair kernel: rounding residual of the fp16 state in units of 2^-19, [pair][2 layers][H/32][16 chunks][32 units][8]
    const uint32_t *info_words;  // bit i = position i is an info (loss) position
    float *logits;               // [B,N] or null
    float *decoded;              // [B,N]
    int64_t B;
    int N, H, tiles_per_step;
    int dbg;  // bench-only experiments (results are garbage): 1 = no bulk copies, 2 = no MMAs, 4 = no per-tile work
    long long *trace;  // bench-only (NPD_GRU_TRACE): [N][TRACE_SLOTS] clock64 stamps of CTA 0, or null
};

constexpr int TRACE_SLOTS = 360;  // per step: MMA warp 0..15 (job begin / issued), 16-17 (h_ready waits); epilogue 20..35, 36 step end
__device__ __forceinline__ long long gtime_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// bench-only: wall-clock (ns) stamp of cluster 0's two CTAs, for the ring round-trip breakdown of the pair kernel
__device__ __forceinline__ void trace_ns(const GruParams &p, int step, int slot)
{
    if (p.trace && (blockIdx.x >> 1) == 0 && slot < TRACE_SLOTS) p.trace[step * TRACE_SLOTS + slot] = gtime_ns();
}
__device__ __forceinline__ void trace_ev(const GruParams &p, int step, int slot)
{
    if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) p.trace[step * TRACE_SLOTS + slot] = clock64();
}

// ---- PTX wrappers -------------------------------------------------------------------------------
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
// one non-blocking probe of a phase (the blocking wait costs ~90 cycles even when the phase is complete)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 2-CTA cluster plumbing: the two CTAs of a pair decode different codewords with the SAME weight stream, so every
// weight tile is fetched from L2 once and multicast into both CTAs' rings (halves the L2 -> SM traffic, which is
// what bounds the kernel once the MMA issue is lean: 16 KB per ~200 cycles per SM)
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint16_t mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_fp16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row groups 1024 B apart, version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// descriptor from its low word ((address & 0x3FFFF) >> 4): K steps and ring stages become plain adds on that word
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return (saddr & 0x3FFFF) >> 4; }
__device__ __forceinline__ uint64_t umma_desc_from_lo(uint32_t lo)
{
    return (uint64_t)lo | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

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
// Precision of the recurrent path (measured on the reference-TRAINED Polar(64,22) checkpoint, where |logit| ~ 1 and the
// update gate z sits near 1 for long-memory units).  h' = (1 - z) n + z h accumulates over N steps whatever
// perturbs z, n or the stored h by ~2^-12: with tanh.approx (max relative error 2^-11) for the gates, r / z parked as fp16
// between their accumulators and the n accumulator, and the state kept only as the fp16 MMA operand, the forced-feedback
// logit error reached 7e-3 (1.5-2.4 x the north-star tolerance 1e-2 |ref| + 2e-3).  Therefore:
//   NPD_GRU_ZC    the update gate is computed as zc = 1 - z = sigmoid(-x) through ex2.approx + rcp.approx (two MUFU, relative
//                 error ~1e-7) and parked as fp16: its RELATIVE rounding keeps long-memory units (z -> 1) exact where a parked
//                 fp16 z (or tanh.approx's 2^-11) perturbs h by 2.4e-4 |h - n| every step; h' = h - zc (h - n)
//   NPD_GRU_LO    the rounding residual of the fp16 state (hnew - fp16(hnew); |h| < 1, so |residual| <= 2^-12: one signed byte in
//                 units of 2^-19) lives in an L2-resident global buffer owned by the thread that wrote it; the update reads
//                 hi + lo, the tensor cores read hi
//   NPD_GRU_ACT   bits 1 / 2: r / n through ex2 + rcp as well (measured: no gain in logit error, +3 % time; off)
#ifndef NPD_GRU_ACT
#define NPD_GRU_ACT 0
#endif
#ifndef NPD_GRU_ZC
#define NPD_GRU_ZC 1
#endif
#ifndef NPD_GRU_LO
#define NPD_GRU_LO 1
#endif
__device__ __forceinline__ float tanh_mufu(float x)
{
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// tanh(x) = 1 - 2 / (1 + e^(2x)); e = inf gives 1, e = 0 gives -1; absolute error ~1e-7
__device__ __forceinline__ float tanh_exact(float x) { return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(x * 2.885390082f)), 1.0f); }
__device__ __forceinline__ float tanh_f(float x) { return (NPD_GRU_ACT & 4) ? tanh_exact(x) : tanh_mufu(x); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b)
{
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t w)
{
    return __half22float2(*reinterpret_cast<const __half2 *>(&w));
}
// sigmoid(x) = 0.5 + 0.5 tanh(x / 2) takes the pre-halved argument; EXACT: 1 / (1 + e^(-x)) = rcp(1 + ex2(-2 half_x log2 e))
template <bool EXACT>
__device__ __forceinline__ float sigmoid_half_arg_t(float half_x)
{
    return EXACT ? rcp_approx(1.0f + ex2_approx(half_x * -2.885390082f)) : fmaf(tanh_mufu(half_x), 0.5f, 0.5f);
}
__device__ __forceinline__ float sigmoid_half_arg(float half_x) { return sigmoid_half_arg_t<false>(half_x); }
// 1 - sigmoid(x) = 1 / (1 + e^x) from the pre-halved argument, relative error ~1e-7
__device__ __forceinline__ float sigmoid_compl_half_arg(float half_x) { return rcp_approx(1.0f + ex2_approx(half_x * 2.885390082f)); }

// byte offset of element (row c, k) inside a K-major SWIZZLE_128B operand buffer of 64-row chunks
__device__ __forceinline__ uint32_t b_off(int c, int k)
{
    return (uint32_t)((k >> 6) * B_CHUNK_BYTES + c * 128 + ((((k & 63) >> 3) ^ (c & 7)) << 4) + (k & 7) * 2);
}

// ---- MLP head on the legacy warp-level tensor-core path: 0.13 MFLOP per codeword and step (2.8 % of the step's GRU
// work), so plain mma.sync from the epilogue warps is enough and leaves the tcgen05 schedule untouched ----
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float selu_f(float x)  // torch.nn.SELU
{
    return 1.0507009873554805f * (fmaxf(x, 0.0f) + fminf(1.6732632423543772f * expm1f(x), 0.0f));
}
// out[c][j] = selu(sum_k W[j][k] in[c][k] + bias[j]) for the CTA's 64 codewords c, j < M; in(c, k) returns the fp16 pair
// (k, k+1) of codeword c; all 16 epilogue warps call it, a warp item = 16 units x 32 codewords
template <class LoadPair>
__device__ __forceinline__ void head_layer(int warp, int lane, const __half *W, const float *bias, int M, int K, LoadPair in,
                                           __half *out)
{
    const int g = lane >> 2, t = lane & 3;
    for (int item = warp; item < (M / 16) * 2; item += EPI_WARPS) {
        const int j0 = (item >> 1) * 16, c0 = (item & 1) * 32;
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = 0.0f;
        const __half *w_lo = W + (size_t)(j0 + g) * K + 2 * t, *w_hi = w_lo + (size_t)8 * K;
        for (int k0 = 0; k0 < K; k0 += 16) {
            uint32_t a[4];
            a[0] = __ldg(reinterpret_cast<const uint32_t *>(w_lo + k0));
            a[1] = __ldg(reinterpret_cast<const uint32_t *>(w_hi + k0));
            a[2] = __ldg(reinterpret_cast<const uint32_t *>(w_lo + k0 + 8));
            a[3] = __ldg(reinterpret_cast<const uint32_t *>(w_hi + k0 + 8));
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int c = c0 + nt * 8 + g;
                mma16816(acc[nt], a, in(c, k0 + 2 * t), in(c, k0 + 2 * t + 8));
            }
        }
        const float b_lo = __ldg(bias + j0 + g), b_hi = __ldg(bias + j0 + g + 8);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int c = c0 + nt * 8 + 2 * t;
            __half *o = out + (size_t)c * M + j0 + g;
            o[0] = __float2half_rn(selu_f(acc[nt][0] + b_lo));
            o[M] = __float2half_rn(selu_f(acc[nt][1] + b_lo));
            o[8] = __float2half_rn(selu_f(acc[nt][2] + b_hi));
            o[M + 8] = __float2half_rn(selu_f(acc[nt][3] + b_hi));
        }
    }
}

struct Smem {
    // dynamic shared memory carve-up (all operand regions 1024-byte aligned)
    static __host__ __device__ size_t ring() { return 0; }
    static __host__ __device__ size_t h0(int) { return (size_t)NUM_STAGES * A_TILE_BYTES; }
    static __host__ __device__ size_t h1(int H) { return h0(H) + (size_t)(H / 64) * B_CHUNK_BYTES; }
    static __host__ __device__ size_t red(int H) { return h1(H) + (size_t)(H / 64) * B_CHUNK_BYTES; }   // [4][64] floats
    static __host__ __device__ size_t bars(int H) { return red(H) + 4 * TILE_B * 4; }
    static __host__ __device__ size_t total(int H) { return bars(H) + 256; }
};

// ---- MMA issue: one weight tile (4 K-steps) with the probe of the NEXT ring stage overlapped ----------------
// The issuing thread is the throughput limit of this kernel (a tile is only 4 x 48 tensor-pipe cycles, measured by
// tools/probe/umma_issue.cu), so its per-tile instruction stream is kept minimal: ring stage, barrier parity and all
// descriptor offsets are compile-time constants (static schedule, see mma_block), and the ~150-cycle latency of the
// barrier probe hides behind the four MMAs because its predicate is only consumed after they are issued (one asm
// block keeps that order).  Returns 1 when the next stage's weights have already landed.
__device__ __forceinline__ uint32_t tile_issue(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc0,
                                               uint32_t bar_empty_cur, uint32_t bar_full_next, uint32_t parity_next,
                                               bool no_mma)
{
    uint32_t ok;
    const uint64_t a0 = umma_desc_from_lo(a_lo), b0 = umma_desc_from_lo(b_lo);
    asm volatile(
        "{\n\t"
        ".reg .pred q, pacc, pone, pm;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q, [%7], %8;\n\t"
        "setp.ne.b32 pacc, %4, 0;\n\t"
        "setp.eq.b32 pone, 0, 0;\n\t"
        "setp.eq.b32 pm, %9, 0;\n\t"
        "@pm tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %5, pacc;\n\t"
        "@pm tcgen05.mma.cta_group::1.kind::f16 [%1], %10, %11, %5, pone;\n\t"
        "@pm tcgen05.mma.cta_group::1.kind::f16 [%1], %12, %13, %5, pone;\n\t"
        "@pm tcgen05.mma.cta_group::1.kind::f16 [%1], %14, %15, %5, pone;\n\t"
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%6], %16;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t"
        "}"
        : "=r"(ok)
        : "r"(d_tmem), "l"(a0), "l"(b0), "r"(acc0), "r"(idesc), "r"(bar_empty_cur), "r"(bar_full_next),
          "r"(parity_next), "r"((uint32_t)no_mma), "l"(a0 + 2), "l"(b0 + 2), "l"(a0 + 4), "l"(b0 + 4), "l"(a0 + 6), "l"(b0 + 6),
          "h"((uint16_t)3)  // the ring slot is released in BOTH CTAs of the pair (their empty barriers count 2 commits)
        : "memory");
    return ok;
}

// Three runs of KH tiles (= 3*KH tiles, a multiple of the 6 ring stages): every block starts at ring stage 0, so the
// stage of tile t is t % 6 and its barrier parity ((t / 6) & 1) ^ pb with pb the block's starting parity.
// Run r accumulates into TMEM columns d[r] from B operand chunks b[r] .. b[r] + KH - 1; first[r] = overwrite.
struct MmaCtx {
    uint32_t bar_full, bar_empty;  // shared addresses of full[0] / empty[0]
    uint32_t ring_lo;              // descriptor low word of ring stage 0
    uint32_t idesc;
    uint32_t ok;                   // the current tile's weights are known to have landed
    uint32_t pb;                   // barrier parity of the next block's first pass over the ring
    bool no_mma;
    int trace_tile, trace_step;    // bench-only: per-tile clock stamps (slots 40..) of one job per step
    int tile_in_step;
    uint16_t empty_mask;           // pair kernel: CTAs whose ring slot a tile's completion releases
    const GruParams *prm;
};

template <int KH>
__device__ __forceinline__ void mma_block(MmaCtx &c, uint32_t d0, uint32_t b0, bool f0, uint32_t d1, uint32_t b1, bool f1,
                                          uint32_t d2, uint32_t b2, bool f2)
{
    constexpr uint32_t A_TILE_LO = A_TILE_BYTES >> 4, B_CHUNK_LO = B_CHUNK_BYTES >> 4;
    constexpr int T = 3 * KH;
    static_assert(T % NUM_STAGES == 0, "a block must be a whole number of ring passes");
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int r = t / KH, kc = t % KH;
        const int stage = t % NUM_STAGES, pass = t / NUM_STAGES;
        const int nstage = (t + 1) % NUM_STAGES, npass = (t + 1) / NUM_STAGES;  // t + 1 == T: stage 0 of the next block
        const uint32_t d = r == 0 ? d0 : r == 1 ? d1 : d2;
        const uint32_t b = (r == 0 ? b0 : r == 1 ? b1 : b2) + kc * B_CHUNK_LO;
        const bool first = (r == 0 ? f0 : r == 1 ? f1 : f2) && kc == 0;
        if (!c.ok) mbar_wait(c.bar_full + 8 * stage, (pass & 1) ^ c.pb);
        c.ok = tile_issue(d, c.ring_lo + stage * A_TILE_LO, b, c.idesc, first ? 0u : 1u, c.bar_empty + 8 * stage,
                          c.bar_full + 8 * nstage, (npass & 1) ^ c.pb, c.no_mma);
    }
    if ((T / NUM_STAGES) & 1) c.pb ^= 1;
}

template <int KH>
__global__ void __launch_bounds__(NUM_THREADS, 1) gru_decode_kernel(const GruParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int H = KH * 64, JOBS = H / JOB_UNITS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N;
    const int64_t cw0 = (int64_t)blockIdx.x * TILE_B;
    if (warp == 0) trace_ev(p, 0, 37);

    unsigned char *s_ring = smem + Smem::ring();
    unsigned char *s_h0 = smem + Smem::h0(H);
    unsigned char *s_h1 = smem + Smem::h1(H);
    float *s_red = reinterpret_cast<float *>(smem + Smem::red(H));
    uint64_t *s_bars = reinterpret_cast<uint64_t *>(smem + Smem::bars(H));
    // barriers: full[S], empty[S], tmem_full[2], tmem_empty[2], h_ready[2]
    const uint32_t bar_full = smem_u32(s_bars), bar_empty = bar_full + 8 * NUM_STAGES,
                   bar_tfull = bar_empty + 8 * NUM_STAGES, bar_tempty = bar_tfull + 16, bar_hready = bar_tempty + 16;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bars + 2 * NUM_STAGES + 6);
    uint32_t *s_bits = s_tmem + 1;  // [2]: feedback bits of the 64 codewords (1 = previous decision was +1)

    // ---- one-time setup ----
    if (tid == 0) {
        for (int i = 0; i < NUM_STAGES; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 2);  // this CTA's and the peer CTA's MMAs on the slot have retired
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, EPI_THREADS);
            mbar_init(bar_hready + 8 * i, EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_bits[0] = 0xffffffffu;  // step 0 feeds back +1 (rnn_all.py:542-543)
        s_bits[1] = 0xffffffffu;
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // h0 = h1 = 0 (rnn_all.py:538) or the caller's initial state (net.get_h0(y), 523-524); the y tile goes, as fp32
    // [k][codeword], into the (still idle) ring memory for the input-projection prologue below
    if (p.h0 == nullptr) {
        for (int i = tid; i < (2 * KH * B_CHUNK_BYTES) / 16; i += NUM_THREADS)
            reinterpret_cast<uint4 *>(s_h0)[i] = make_uint4(0, 0, 0, 0);
    } else {
        for (int i = tid; i < 2 * TILE_B * H; i += NUM_THREADS) {
            const int layer = i / (TILE_B * H), c = (i / H) % TILE_B, u = i % H;
            const float v = (cw0 + c < p.B) ? p.h0[((size_t)layer * p.B + cw0 + c) * H + u] : 0.0f;
            *reinterpret_cast<__half *>((layer ? s_h1 : s_h0) + b_off(c, u)) = __float2half_rn(v);
        }
    }
    float *s_yT = reinterpret_cast<float *>(s_ring);
    for (int i = tid; i < TILE_B * N; i += NUM_THREADS) {
        const int c = i / N, k = i % N;
        s_yT[k * TILE_B + c] = (cw0 + c < p.B) ? p.y[(cw0 + c) * N + k] : 0.0f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t rank = cluster_ctarank();

    // ---- input projection, hoisted out of the step loop: Gy = W_iy y is the same in all N steps.
    // Every epilogue thread computes, in fp32, exactly the (unit, 16 codewords) entries it will add to its layer-0
    // accumulators later and keeps them, rounded to fp16, in thread-local memory (3 gates x JOBS x 16 halves = 384 B
    // per thread, 29 MB for a full wave: L2-resident; as fp32 the wave's 58 MB fell out of L2 -- ncu showed 7 GB of
    // DRAM traffic per launch).
    uint4 gy[JOBS * 3 * (CW_PER_THREAD / 8)];
    if (warp < EPI_WARPS) {
        const int q = warp & 3, col0 = (warp >> 2) * CW_PER_THREAD;
        for (int jg = 0; jg < JOBS * 3; ++jg) {
            const int j = jg / 3, g = jg % 3;
            const int row = g * H + j * JOB_UNITS + q * 32 + lane;
            float acc[CW_PER_THREAD];
#pragma unroll
            for (int i = 0; i < CW_PER_THREAD; ++i) acc[i] = 0.0f;
#pragma unroll 8
            for (int k = 0; k < N; ++k) {
                const float w = __ldg(p.w_iyT + (size_t)k * (3 * H) + row);
#pragma unroll
                for (int i4 = 0; i4 < CW_PER_THREAD / 4; ++i4) {
                    const float4 yv = *reinterpret_cast<const float4 *>(s_yT + k * TILE_B + col0 + 4 * i4);
                    acc[4 * i4 + 0] = fmaf(w, yv.x, acc[4 * i4 + 0]);
                    acc[4 * i4 + 1] = fmaf(w, yv.y, acc[4 * i4 + 1]);
                    acc[4 * i4 + 2] = fmaf(w, yv.z, acc[4 * i4 + 2]);
                    acc[4 * i4 + 3] = fmaf(w, yv.w, acc[4 * i4 + 3]);
                }
            }
#pragma unroll
            for (int i8 = 0; i8 < CW_PER_THREAD / 8; ++i8)
                gy[jg * (CW_PER_THREAD / 8) + i8] =
                    make_uint4(pack_h2(acc[8 * i8 + 0], acc[8 * i8 + 1]), pack_h2(acc[8 * i8 + 2], acc[8 * i8 + 3]),
                               pack_h2(acc[8 * i8 + 4], acc[8 * i8 + 5]), pack_h2(acc[8 * i8 + 6], acc[8 * i8 + 7]));
        }
    }
    // the ring memory (y tile) is free for the weight stream from here on, in both CTAs of the pair (the peer
    // multicasts into this CTA's ring and arrives on its barriers)
    cluster_sync_all();
    if (warp == 0) trace_ev(p, 0, 38);

    if (warp >= EPI_WARPS && warp < MMA_WARP) {
        // ================= producer: stream the weight program, once per step =================
        if (lane == 0) {
            const uint32_t T = (uint32_t)p.tiles_per_step;
            const uint32_t total = (uint32_t)N * T;
            uint32_t t = 0, stage = 0, phase = 0;
            for (uint32_t g = 0; g < total; ++g) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);  // both CTAs are done with the slot
                if (p.dbg & 1) {
                    mbar_arrive(bar_full + 8 * stage);
                } else {
                    mbar_expect_tx(bar_full + 8 * stage, A_TILE_BYTES);
                    if ((g & 1u) == rank)  // the CTAs take turns fetching a tile for both
                        bulk_g2s_multicast(smem_u32(s_ring + stage * A_TILE_BYTES), p.wpack + (size_t)t * A_TILE_BYTES,
                                           A_TILE_BYTES, bar_full + 8 * stage, (uint16_t)3);
                }
                if (++t == T) t = 0;
                if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer: one thread runs the static schedule =================
        // M = 128, N = 64, fp16 x fp16 -> fp32 (a_format = b_format = 0), both operands K-major
        if (elect_one()) {
            MmaCtx c;
            c.bar_full = bar_full; c.bar_empty = bar_empty; c.ring_lo = umma_desc_lo(smem_u32(s_ring));
            c.idesc = (1u << 4) | ((uint32_t)(TILE_B >> 3) << 17) | ((128u >> 4) << 24);
            c.ok = 0; c.pb = 0; c.no_mma = (p.dbg & 2) != 0; c.trace_tile = -1; c.trace_step = 0; c.prm = &p; c.tile_in_step = 1 << 30; c.empty_mask = 3;
            const uint32_t b_h0 = umma_desc_lo(smem_u32(s_h0)), b_h1 = umma_desc_lo(smem_u32(s_h1));
            uint32_t job = 0;  // global job counter -> TMEM slot job & 1
            uint32_t hphase0 = 0, hphase1 = 0;
            for (int step = 0; step < N; ++step) {
                // ---- layer 0: R, Z, NH from h0 (the y part of R, Z and all of NI are the hoisted Gy) ----
                for (int j = 0; j < JOBS; ++j, ++job) {
                    const uint32_t slot = job & 1, d0 = tmem_base + slot * 256;
                    mbar_wait(bar_tempty + 8 * slot, ((job >> 1) & 1) ^ 1);
                    tc_fence_after();
                    trace_ev(p, step, 2 * j);
                    mma_block<KH>(c, d0 + 0 * TILE_B, b_h0, true, d0 + 1 * TILE_B, b_h0, true, d0 + 3 * TILE_B, b_h0, true);
                    umma_commit(bar_tfull + 8 * slot);
                    trace_ev(p, step, 2 * j + 1);
                }
                // ---- layer 1, two jobs at a time: both hidden-state streams first (they need only the previous
                // step's h1 and cover the wait for layer 0's state update), then both input streams (this step's h0)
                for (int jp = 0; jp < JOBS; jp += 2) {
                    const int nj = (JOBS - jp) < 2 ? (JOBS - jp) : 2;
                    for (int i = 0; i < nj; ++i) {
                        const uint32_t jb = job + i, slot = jb & 1, d0 = tmem_base + slot * 256;
                        mbar_wait(bar_tempty + 8 * slot, ((jb >> 1) & 1) ^ 1);
                        tc_fence_after();
                        trace_ev(p, step, 8 + 2 * (jp + i));
                        if (jp == 0 && i == 0 && step > 0) {
                            mbar_wait(bar_hready + 8, hphase1);
                            hphase1 ^= 1;
                            tc_fence_after();
                            trace_ev(p, step, 17);
                        }
                        mma_block<KH>(c, d0 + 3 * TILE_B, b_h1, true, d0 + 0 * TILE_B, b_h1, true, d0 + 1 * TILE_B, b_h1, true);
                    }
                    if (jp == 0) {
                        trace_ev(p, step, 18);
                        mbar_wait(bar_hready + 0, hphase0);
                        hphase0 ^= 1;
                        tc_fence_after();
                        trace_ev(p, step, 16);
                    }
                    for (int i = 0; i < nj; ++i) {
                        const uint32_t jb = job + i, slot = jb & 1, d0 = tmem_base + slot * 256;
                        mma_block<KH>(c, d0 + 0 * TILE_B, b_h0, false, d0 + 1 * TILE_B, b_h0, false, d0 + 2 * TILE_B, b_h0, true);
                        umma_commit(bar_tfull + 8 * slot);
                        trace_ev(p, step, 8 + 2 * (jp + i) + 1);
                    }
                    job += nj;
                }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue warps: gate math, state update, head, feedback =================
        constexpr int CW = CW_PER_THREAD;                 // 16 codewords per thread
        const int q = warp & 3, cq = warp >> 2;
        const int col0 = cq * CW;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint32_t job = 0;
        uint32_t staged[3 * (CW / 2)];  // all jobs but the last of a layer x 16 codewords, fp16 pairs (codeword 2i, 2i+1)
        float head[CW];
        const uint32_t info0 = p.info_words[0], info1 = N > 32 ? p.info_words[1] : 0u,
                       info2 = N > 64 ? p.info_words[2] : 0u, info3 = N > 96 ? p.info_words[3] : 0u;

        for (int step = 0; step < N; ++step) {
            const uint32_t bits = s_bits[cq >> 1] >> ((cq & 1) * CW);
            for (int layer = 0; layer < 2; ++layer) {
                unsigned char *s_h = layer ? s_h1 : s_h0;
                if (layer == 1) {
#pragma unroll
                    for (int i = 0; i < CW; ++i) head[i] = 0.0f;
                }
#pragma unroll
                for (int j = 0; j < JOBS; ++j) {
                    const int u = j * JOB_UNITS + q * 32 + lane;  // hidden unit of this thread
                    // per-unit constants; sigmoid arguments are pre-halved (sigmoid(x) = 0.5 + 0.5 tanh(x / 2)) and the
                    // one-hot input column of the previous decision is folded into the layer-0 biases
                    float hr0, hr1, hz0, hz1, bn0, bn1, b_hn, wo = 0.f;
                    if (layer == 0) {
                        const float4 c0 = __ldg(reinterpret_cast<const float4 *>(p.consts0 + (size_t)u * 12));
                        const float4 c1 = __ldg(reinterpret_cast<const float4 *>(p.consts0 + (size_t)u * 12 + 4));
                        const float2 c2 = __ldg(reinterpret_cast<const float2 *>(p.consts0 + (size_t)u * 12 + 8));
                        hr0 = 0.5f * (c0.x + c1.x); hr1 = 0.5f * (c0.x + c1.y);
                        hz0 = 0.5f * (c0.y + c1.z); hz1 = 0.5f * (c0.y + c1.w);
                        bn0 = c0.z + c2.x; bn1 = c0.z + c2.y;
                        b_hn = c0.w;
                    } else {
                        const float4 c0 = __ldg(reinterpret_cast<const float4 *>(p.consts1 + (size_t)u * 4));
                        hr0 = hr1 = 0.5f * c0.x; hz0 = hz1 = 0.5f * c0.y; bn0 = bn1 = c0.z; b_hn = c0.w;
                        wo = __ldg(p.w_out + u);
                    }
                    const uint32_t slot = job & 1;
                    mbar_wait(bar_tfull + 8 * slot, (job >> 1) & 1);
                    tc_fence_after();
                    if (warp == 0) trace_ev(p, step, 20 + 2 * (layer * 4 + j));
                    const uint32_t t0 = tmem_base + lane_addr + slot * 256 + col0;
#pragma unroll
                    for (int cc = 0; cc < CW; cc += 8) {
                        float aR[8], aZ[8], aNI[8], aNH[8];
                        tmem_ld8(t0 + 0 * TILE_B + cc, aR);
                        tmem_ld8(t0 + 1 * TILE_B + cc, aZ);
                        if (layer == 1) tmem_ld8(t0 + 2 * TILE_B + cc, aNI);
                        tmem_ld8(t0 + 3 * TILE_B + cc, aNH);
                        tmem_ld_wait();
                        if (cc + 8 == CW) {
                            // all accumulators of this job are in registers: the MMA warp may refill the slot
                            tc_fence_before();
                            mbar_arrive(bar_tempty + 8 * slot);
                        }
                        if (layer == 0) {
                            // add the hoisted input projection of this (unit, codewords)
                            const uint4 gr = gy[(j * 3 + 0) * (CW / 8) + cc / 8];
                            const uint4 gz = gy[(j * 3 + 1) * (CW / 8) + cc / 8];
                            const uint4 gn = gy[(j * 3 + 2) * (CW / 8) + cc / 8];
                            const uint32_t wr[4] = {gr.x, gr.y, gr.z, gr.w}, wz[4] = {gz.x, gz.y, gz.z, gz.w},
                                           wn[4] = {gn.x, gn.y, gn.z, gn.w};
#pragma unroll
                            for (int h2 = 0; h2 < 4; ++h2) {
                                const float2 fr = unpack_h2(wr[h2]), fz = unpack_h2(wz[h2]), fn = unpack_h2(wn[h2]);
                                aR[2 * h2] += fr.x; aR[2 * h2 + 1] += fr.y;
                                aZ[2 * h2] += fz.x; aZ[2 * h2 + 1] += fz.y;
                                aNI[2 * h2] = fn.x; aNI[2 * h2 + 1] = fn.y;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const bool plus = (layer == 0) && ((bits >> (cc + i)) & 1u);
                            const float r = sigmoid_half_arg(fmaf(aR[i], 0.5f, plus ? hr1 : hr0));
                            const float nn = tanh_f(fmaf(r, aNH[i] + b_hn, aNI[i] + (plus ? bn1 : bn0)));
                            const float hold = __half2float(*reinterpret_cast<const __half *>(s_h + b_off(col0 + cc + i, u)));
#if NPD_GRU_ZC
                            // update gate as zc = 1 - z through ex2 + rcp (see NPD_GRU_ZC above): h' = h - zc (h - n)
                            const float zc = sigmoid_compl_half_arg(fmaf(aZ[i], 0.5f, plus ? hz1 : hz0));
                            const float hnew = fmaf(-zc, hold - nn, hold);
#else
                            const float z = sigmoid_half_arg(fmaf(aZ[i], 0.5f, plus ? hz1 : hz0));
                            const float hnew = fmaf(z, hold - nn, nn);  // (1 - z) n + z h
#endif
                            if (layer == 1) head[cc + i] = fmaf(wo, hnew, head[cc + i]);
                            const unsigned short hb = __half_as_ushort(__float2half_rn(hnew));
                            if (j == JOBS - 1) {
                                // tmem_full of the layer's last job: every MMA reading the old state has retired
                                *reinterpret_cast<unsigned short *>(s_h + b_off(col0 + cc + i, u)) = hb;
                            } else {
                                const int si = j * (CW / 2) + ((cc + i) >> 1);
                                if ((i & 1) == 0) staged[si] = hb; else staged[si] |= (uint32_t)hb << 16;
                            }
                        }
                    }
                    if (warp == 0) trace_ev(p, step, 21 + 2 * (layer * 4 + j));
                    ++job;
                }
                // every MMA that reads the old state of this layer has retired (tmem_full of the last
                // job): write the new state in place
#pragma unroll
                for (int j = 0; j < JOBS - 1; ++j) {
                    const int u = j * JOB_UNITS + q * 32 + lane;
#pragma unroll
                    for (int i = 0; i < CW / 2; ++i) {
                        const uint32_t pr = staged[j * (CW / 2) + i];
                        *reinterpret_cast<unsigned short *>(s_h + b_off(col0 + 2 * i, u)) = (unsigned short)(pr & 0xffffu);
                        *reinterpret_cast<unsigned short *>(s_h + b_off(col0 + 2 * i + 1, u)) = (unsigned short)(pr >> 16);
                    }
                }
                fence_async_smem();
                mbar_arrive(bar_hready + 8 * layer);
            }

            // ---- head: logit[c] = w_out . h1[c] + b_out ; transpose-reduce over the 32 lanes (units).  A position
            // outside the loss set decides +1 (or its genie value) whatever the logit: skipped unless logits are wanted.
            const uint32_t iw = step < 32 ? info0 : step < 64 ? info1 : step < 96 ? info2 : info3;
            const bool is_info = (iw >> (step & 31)) & 1u;
            const bool need_head = is_info || p.logits != nullptr;
            if (need_head && p.head_depth > 1) {
                // MLP head: h1 of this step is complete in shared memory once every epilogue thread has passed here
                epi_bar_sync();
                const int Yh = p.head_yh;
                __half *act0 = p.head_act + (size_t)blockIdx.x * 2 * TILE_B * Yh, *act1 = act0 + (size_t)TILE_B * Yh;
                head_layer(warp, lane, p.head_w1, p.head_b, Yh, H,
                           [&](int c, int k) { return *reinterpret_cast<const uint32_t *>(s_h1 + b_off(c, k)); }, act0);
                epi_bar_sync();
                for (int l = 0; l < p.head_depth - 2; ++l) {
                    const __half *src = act0;
                    head_layer(warp, lane, p.head_wh + (size_t)l * Yh * Yh, p.head_b + (size_t)(l + 1) * Yh, Yh, Yh,
                               [&](int c, int k) { return __ldcg(reinterpret_cast<const uint32_t *>(src + (size_t)c * Yh + k)); },
                               act1);
                    epi_bar_sync();
                    __half *tmp = act0; act0 = act1; act1 = tmp;
                }
#pragma unroll
                for (int i = 0; i < TILE_B / EPI_WARPS; ++i) {
                    const int c = warp * (TILE_B / EPI_WARPS) + i;
                    float sacc = 0.0f;
                    for (int j = lane; j < Yh; j += 32)
                        sacc = fmaf(__ldg(p.head_wl + j), __half2float(__ldcg(act0 + (size_t)c * Yh + j)), sacc);
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) sacc += __shfl_xor_sync(NPD_FULL, sacc, o);
                    if (lane < 4) s_red[lane * TILE_B + c] = lane == 0 ? sacc : 0.0f;
                }
                epi_bar_sync();
            } else if (need_head) {
#pragma unroll
                for (int s = CW / 2; s >= 1; s >>= 1) {
#pragma unroll
                    for (int i = 0; i < s; ++i) {
                        const bool up = (lane & s) != 0;
                        const float send = up ? head[i] : head[i + s];
                        const float keep = up ? head[i + s] : head[i];
                        head[i] = keep + __shfl_xor_sync(NPD_FULL, send, s);
                    }
                }
                head[0] += __shfl_xor_sync(NPD_FULL, head[0], CW);  // lanes l and l ^ 16 hold the two halves of the units
                if (lane < CW) s_red[q * TILE_B + col0 + lane] = head[0];  // lane l holds column col0 + l of this lane quarter
                epi_bar_sync();
            }
            if (warp < 2) {
                const int c = warp * 32 + lane;
                const float logit = need_head ? ((s_red[c] + s_red[TILE_B + c]) + (s_red[2 * TILE_B + c] + s_red[3 * TILE_B + c])) + p.b_out : 0.0f;
                const bool valid = cw0 + c < p.B;
                // decoded = ones, or gt.clone() in genie mode; only loss positions are overwritten
                // (rnn_all.py:519-522, 546-547)
                float dec = (p.genie && valid) ? p.genie[(cw0 + c) * N + step] : 1.0f;
                if (is_info) dec = (logit > 0.0f) ? 1.0f : ((logit < 0.0f) ? -1.0f : 0.0f);
                if (valid) {
                    if (p.logits) p.logits[(cw0 + c) * N + step] = logit;
                    p.decoded[(cw0 + c) * N + step] = dec;
                }
                // next step feeds back sign(decoded[:, step]) (a genie value need not be +-1) ...
                float prev = (dec > 0.0f) ? 1.0f : ((dec < 0.0f) ? -1.0f : 0.0f);
                if (p.forced && valid) prev = p.forced[(cw0 + c) * N + step];  // ... or the forced sequence
                // get_onehot (rnn_all.py:258-260): index = (0.5 + 0.5*prev).long() -> 1 only for prev = +1
                const uint32_t m = __ballot_sync(NPD_FULL, prev >= 1.0f);
                if (lane == 0) s_bits[warp] = m;
            }
            epi_bar_sync();
            if (warp == 0) trace_ev(p, step, 36);
        }
    }

    tc_fence_before();
    cluster_sync_all();  // the peer may still signal this CTA's barriers / write its ring until its own schedule ends
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// =====================================================================================================================
// CTA-pair kernel (cta_group::2): M = 256 hidden units x N = 128 codewords per tcgen05.mma
// =====================================================================================================================
// tools/probe/umma_issue.cu / umma2_probe.cu measured what bounds the single-CTA kernel above: an M=128 N=64 MMA costs
// 48 cycles (its 6 KB of operand reads at the 128 B/clk shared-memory port) against 32 at the tensor rate, while one
// M=256 N=128 cta_group::2 MMA takes 64 cycles for four times the work -- each SM reads its 128-unit half of the
// weight tile once for 128 codewords.  Here a cluster of two CTAs decodes 128 codewords:
//   * CTA r keeps the hidden states of its own 64 codewords (the B operand rows 64r .. 64r+63) and streams only the
//     unit rows 256J + 128r .. + 127 of every weight tile (its half of the A operand): half the L2 traffic per SM;
//   * the leader CTA's elected thread issues every MMA and commits (multicast) to both CTAs' barriers; the peer's MMA
//     warp relays "my half of the tile has landed" to the leader's full barriers;
//   * CTA r's TMEM receives D rows = its 128 units x ALL 128 codewords; a job's four accumulators (R, Z, NI, NH) x
//     128 columns fill TMEM, so the hand-off between MMA and epilogue is per accumulator (4 full / 4 empty barriers):
//     R and Z are drained as soon as they complete, while the MMAs of the later accumulators still run;
//   * an epilogue thread owns one unit and 32 codewords; half of the warps update codewords whose state lives in the
//     peer CTA: they read the old and write the new state through distributed shared memory, and every CTA sends its
//     partial head sums (over its 256 units) to the peer so that both compute identical decisions for all 128
//     codewords (the one-hot feedback is needed by both CTAs' layer-0 epilogues).
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// a pure signal (no data published through it): a release at cluster scope costs the arriving thread ~0.8 us
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP_C:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_C;\n\t"
        "bra WAIT_LOOP_C;\n\t"
        "DONE_C:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_cluster_u16(uint32_t addr, unsigned short v)
{
    asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ unsigned short ld_cluster_u16(uint32_t addr)
{
    unsigned short v;
    asm volatile("ld.shared::cluster.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v)
{
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void umma_commit2(uint32_t bar, uint16_t mask)  // arrives on the same barrier of every CTA in mask
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

// one weight tile (4 K-steps of M=256 N=128) + probe of the next ring stage; see tile_issue
__device__ __forceinline__ uint32_t tile_issue2(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc0,
                                                uint32_t bar_empty_cur, uint32_t bar_full_next, uint32_t parity_next,
                                                uint32_t skip_mma, uint16_t empty_mask)
{
    uint32_t ok;
    const uint64_t a0 = umma_desc_from_lo(a_lo), b0 = umma_desc_from_lo(b_lo);
    asm volatile(
        "{\n\t"
        ".reg .pred q, pacc, pone, pm;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q, [%7], %8;\n\t"
        "setp.ne.b32 pacc, %4, 0;\n\t"
        "setp.eq.b32 pone, 0, 0;\n\t"
        "setp.eq.b32 pm, %16, 0;\n\t"
        "@pm tcgen05.mma.cta_group::2.kind::f16 [%1], %2, %3, %5, pacc;\n\t"
        "@pm tcgen05.mma.cta_group::2.kind::f16 [%1], %9, %10, %5, pone;\n\t"
        "@pm tcgen05.mma.cta_group::2.kind::f16 [%1], %11, %12, %5, pone;\n\t"
        "@pm tcgen05.mma.cta_group::2.kind::f16 [%1], %13, %14, %5, pone;\n\t"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%6], %15;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t"
        "}"
        : "=r"(ok)
        : "r"(d_tmem), "l"(a0), "l"(b0), "r"(acc0), "r"(idesc), "r"(bar_empty_cur), "r"(bar_full_next), "r"(parity_next),
          "l"(a0 + 2), "l"(b0 + 2), "l"(a0 + 4), "l"(b0 + 4), "l"(a0 + 6), "l"(b0 + 6), "h"(empty_mask), "r"(skip_mma)
        : "memory");
    return ok;
}

// run RUN (0..2) of a 3-run block: KH tiles accumulating into TMEM columns d from B chunks b .. b + KH - 1
template <int KH, int RUN>
__device__ __forceinline__ void mma_run2(MmaCtx &c, uint32_t d, uint32_t b, bool first)
{
    constexpr uint32_t A_TILE_LO = A_TILE_BYTES >> 4, B_CHUNK_LO = B_CHUNK_BYTES >> 4;
#pragma unroll
    for (int kc = 0; kc < KH; ++kc) {
        const int t = RUN * KH + kc;
        const int stage = t % NUM_STAGES, pass = t / NUM_STAGES;
        const int nstage = (t + 1) % NUM_STAGES, npass = (t + 1) / NUM_STAGES;
        // (cluster-scope acquires cost ~1000 cycles each; the issuing thread reads none of the data itself -- each SM's
        // tensor core reads its own shared memory, which the bulk copy completed before signalling)
        if (!c.ok && !c.no_mma) mbar_wait(c.bar_full + 8 * stage, (pass & 1) ^ c.pb);
        if (c.tile_in_step < 48) trace_ns(*c.prm, c.trace_step, 100 + c.tile_in_step);        // leader: data of both halves seen
        c.ok = tile_issue2(d, c.ring_lo + stage * A_TILE_LO, b + kc * B_CHUNK_LO, c.idesc, (first && kc == 0) ? 0u : 1u,
                           c.bar_empty + 8 * stage, c.bar_full + 8 * nstage, (npass & 1) ^ c.pb, (uint32_t)(c.prm->dbg & 32),
                           c.empty_mask);
        if (c.trace_tile >= 0 && c.trace_tile < 48 && c.prm->trace) c.prm->trace[c.trace_step * TRACE_SLOTS + 40 + c.trace_tile++] = clock64();
        if (c.tile_in_step < 48) trace_ns(*c.prm, c.trace_step, 150 + c.tile_in_step);        // leader: MMAs + commit issued
        ++c.tile_in_step;
    }
    if (RUN == 2 && ((3 * KH / NUM_STAGES) & 1)) c.pb ^= 1;
}

constexpr int PAIR_CW = 2 * TILE_B;   // codewords per cluster
constexpr int CW3 = PAIR_CW / 4;      // 32 accumulator columns (codewords) per epilogue thread

struct Smem3 {
    static __host__ __device__ size_t ring() { return 0; }
    static __host__ __device__ size_t h0(int) { return (size_t)NUM_STAGES * A_TILE_BYTES; }
    static __host__ __device__ size_t h1(int H) { return h0(H) + (size_t)(H / 64) * B_CHUNK_BYTES; }
    static __host__ __device__ size_t headacc(int H) { return h1(H) + (size_t)(H / 64) * B_CHUNK_BYTES; }  // [4][128] floats
    static __host__ __device__ size_t part(int H) { return headacc(H) + 4 * PAIR_CW * 4; }                  // [128] floats: the peer's partial logits
    static __host__ __device__ size_t bars(int H) { return part(H) + PAIR_CW * 4; }
    static __host__ __device__ size_t total(int H) { return bars(H) + 256; }
};

// 2-SM tensor copy of one 16 KB half-tile ([128 rows][128 B] box of the pre-swizzled weight stream) into THIS CTA's
// ring; its transaction bytes complete on the LEADER CTA's barrier (peer bit of the shared::cluster address cleared)
__device__ __forceinline__ void tmap_g2s_pair(uint32_t dst, const CUtensorMap *tmap, int row, uint32_t bar_local)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_local & 0xFEFFFFFFu), "r"(0), "r"(row) : "memory");
}

template <int KH>
#ifndef NPD_GRU_K3_BOUNDS
#define NPD_GRU_K3_BOUNDS __launch_bounds__(NUM_THREADS, 1)
#endif
__global__ void NPD_GRU_K3_BOUNDS gru_decode_kernel3(const GruParams p, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int H = KH * 64, JOBS2 = H / 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N;
    const uint32_t crank = cluster_ctarank();  // rank in the cluster (2 CTAs = one pair, or 4 = two pairs sharing weights)
    const uint32_t rank = crank & 1u;          // rank in the MMA pair: 0 = leader
    const uint32_t lead = crank & ~1u, sibling = crank ^ 2u;
    const uint16_t pair_mask = (uint16_t)(3u << lead);
    const int64_t pair0 = (int64_t)(blockIdx.x >> 1) * PAIR_CW;  // first codeword of the pair
    const int64_t cw0 = pair0 + rank * TILE_B;                   // first codeword whose state lives in this CTA

    unsigned char *s_ring = smem + Smem3::ring();
    unsigned char *s_h0 = smem + Smem3::h0(H);
    unsigned char *s_h1 = smem + Smem3::h1(H);
    float *s_headacc = reinterpret_cast<float *>(smem + Smem3::headacc(H));
    float *s_part = reinterpret_cast<float *>(smem + Smem3::part(H));
    uint64_t *s_bars = reinterpret_cast<uint64_t *>(smem + Smem3::bars(H));
    // barriers: full[S], empty[S], acc_full[4], acc_empty[4] (leader's are used), h_ready[2] (leader's), head[1]
    const uint32_t bar_full = smem_u32(s_bars), bar_empty = bar_full + 8 * NUM_STAGES,
                   bar_afull = bar_empty + 8 * NUM_STAGES, bar_aempty = bar_afull + 32, bar_hready = bar_aempty + 32,
                   bar_head = bar_hready + 16;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bars + 2 * NUM_STAGES + 11);
    uint32_t *s_bits = s_tmem + 1;  // [4]: feedback bits of the pair's 128 codewords (1 = previous decision was +1)

    if (tid == 0) {
        for (int i = 0; i < NUM_STAGES; ++i) {
            // relay variant: leader = own copy + the peer's relay; tensor-copy variant: one expect_tx of both halves
            mbar_init(bar_full + 8 * i, (rank == 0 && !p.use_tmap) ? 2 : 1);
            mbar_init(bar_empty + 8 * i, p.quad ? 2 : 1);  // both pairs' MMAs on a shared (multicast) slot have retired
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(bar_afull + 8 * i, 1);
            mbar_init(bar_aempty + 8 * i, 2 * EPI_WARPS);  // one arrive per epilogue warp of both CTAs
        }
        for (int i = 0; i < 2; ++i) mbar_init(bar_hready + 8 * i, 2 * EPI_WARPS);
        mbar_init(bar_head, 4);  // the peer's four warps that send partial logits
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 4; ++i) s_bits[i] = 0xffffffffu;  // step 0 feeds back +1 (rnn_all.py:542-543)
    }
    // h0 = h1 = 0 (rnn_all.py:538) or the caller's initial state of this CTA's 64 codewords (net.get_h0(y), 523-524);
    // y of all 128 codewords of the pair as fp32 [k][codeword] in the idle ring memory
    if (p.h0 == nullptr) {
        for (int i = tid; i < (2 * KH * B_CHUNK_BYTES) / 16; i += NUM_THREADS)
            reinterpret_cast<uint4 *>(s_h0)[i] = make_uint4(0, 0, 0, 0);
    } else {
        for (int i = tid; i < 2 * TILE_B * H; i += NUM_THREADS) {
            const int layer = i / (TILE_B * H), c = (i / H) % TILE_B, u = i % H;
            const float v = (cw0 + c < p.B) ? p.h0[((size_t)layer * p.B + cw0 + c) * H + u] : 0.0f;
            *reinterpret_cast<__half *>((layer ? s_h1 : s_h0) + b_off(c, u)) = __float2half_rn(v);
        }
    }
    float *s_yT = reinterpret_cast<float *>(s_ring);
    for (int i = tid; i < PAIR_CW * N; i += NUM_THREADS) {
        const int c = i / N, k = i % N;
        s_yT[k * PAIR_CW + c] = (pair0 + c < p.B) ? p.y[(pair0 + c) * N + k] : 0.0f;
    }
    fence_async_smem();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers exist before any cross-CTA signal; TMEM allocation is pair-wide
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    // ---- hoisted input projection for (this thread's unit) x (its 32 codeword columns), fp16 in local memory ----
    uint4 gy[JOBS2 * 3 * (CW3 / 8)];
    if (warp < EPI_WARPS) {
        const int q = warp & 3, col0 = (warp >> 2) * CW3;
        for (int jg = 0; jg < JOBS2 * 3; ++jg) {
            const int j = jg / 3, g = jg % 3;
            const int row = g * H + j * 256 + (int)rank * 128 + q * 32 + lane;
            float acc[CW3];
#pragma unroll
            for (int i = 0; i < CW3; ++i) acc[i] = 0.0f;
#pragma unroll 4
            for (int k = 0; k < N; ++k) {
                const float w = __ldg(p.w_iyT + (size_t)k * (3 * H) + row);
#pragma unroll
                for (int i4 = 0; i4 < CW3 / 4; ++i4) {
                    const float4 yv = *reinterpret_cast<const float4 *>(s_yT + k * PAIR_CW + col0 + 4 * i4);
                    acc[4 * i4 + 0] = fmaf(w, yv.x, acc[4 * i4 + 0]);
                    acc[4 * i4 + 1] = fmaf(w, yv.y, acc[4 * i4 + 1]);
                    acc[4 * i4 + 2] = fmaf(w, yv.z, acc[4 * i4 + 2]);
                    acc[4 * i4 + 3] = fmaf(w, yv.w, acc[4 * i4 + 3]);
                }
            }
#pragma unroll
            for (int i8 = 0; i8 < CW3 / 8; ++i8)
                gy[jg * (CW3 / 8) + i8] =
                    make_uint4(pack_h2(acc[8 * i8 + 0], acc[8 * i8 + 1]), pack_h2(acc[8 * i8 + 2], acc[8 * i8 + 3]),
                               pack_h2(acc[8 * i8 + 4], acc[8 * i8 + 5]), pack_h2(acc[8 * i8 + 6], acc[8 * i8 + 7]));
        }
    }
    cluster_sync_all();  // the ring memory (y tile) of both CTAs is free for the weight stream from here on

    const uint32_t lead_aempty = mapa_u32(bar_aempty, lead), lead_hready = mapa_u32(bar_hready, lead);

    if (warp >= EPI_WARPS && warp < MMA_WARP) {
        // ================= producer: this CTA's half of every weight tile =================
        // PRODUCER_LANES threads take the ring stages in turn (stage = tile % 6, so each lane owns fixed stages): one
        // thread's wake-up + issue latency per tile is otherwise what the ring round trip waits for
        constexpr int PRODUCER_LANES = 2;
        static_assert(NUM_STAGES % PRODUCER_LANES == 0, "each producer lane owns fixed ring stages");
        if (lane < PRODUCER_LANES) {
            const uint32_t T = (uint32_t)p.tiles_per_step2;
            const uint32_t total = (p.dbg & 1) ? 0u : (uint32_t)N * T;
            const unsigned char *src = p.wpack2 + (size_t)rank * T * A_TILE_BYTES;
            uint32_t t = (uint32_t)lane, stage = (uint32_t)lane, phase = 0;
            for (uint32_t g = (uint32_t)lane; g < total; g += PRODUCER_LANES) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                if (t < 48) trace_ns(p, (int)(g / T), (rank ? 250 : 200) + (int)t);  // producer: slot free, copy goes out
                if (p.use_tmap) {
                    // both halves complete on the leader's barrier: no relay, the leader expects 2 x 16 KB
                    if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * A_TILE_BYTES);
                    tmap_g2s_pair(smem_u32(s_ring + stage * A_TILE_BYTES), &tmap, (int)((rank * T + t) * 128), bar_full + 8 * stage);
                } else {
                    // every CTA's copy signals its OWN barrier (a plain bulk copy whose mbarrier operand points into the
                    // other CTA never completes -- tried: the kernel hangs), hence the relay in the peer's MMA warp
                    mbar_expect_tx(bar_full + 8 * stage, A_TILE_BYTES);
                    if (!p.quad)
                        bulk_g2s(smem_u32(s_ring + stage * A_TILE_BYTES), src + (size_t)t * A_TILE_BYTES, A_TILE_BYTES, bar_full + 8 * stage);
                    else if (((g & 1u) << 1 | rank) == crank)  // the two CTAs with this pair rank take turns fetching for both
                        bulk_g2s_multicast(smem_u32(s_ring + stage * A_TILE_BYTES), src + (size_t)t * A_TILE_BYTES, A_TILE_BYTES,
                                           bar_full + 8 * stage, (uint16_t)((1u << crank) | (1u << sibling)));
                }
                t += PRODUCER_LANES;
                if (t >= T) t -= T;
                stage += PRODUCER_LANES;
                if (stage >= NUM_STAGES) { stage -= NUM_STAGES; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        if (elect_one()) {
            if (rank != 0) {
                // ================= peer: relay "my half has landed" to the leader's full barriers =================
                const uint32_t total = ((p.dbg & 1) || p.use_tmap) ? 0u : (uint32_t)N * (uint32_t)p.tiles_per_step2;
                const uint32_t lead_full = mapa_u32(bar_full, lead);
                uint32_t stage = 0, phase = 0;
                const uint32_t T = (uint32_t)p.tiles_per_step2;
                for (uint32_t g = 0; g < total; ++g) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    if (g % T < 48) trace_ns(p, (int)(g / T), 300 + (int)(g % T));  // relay: the peer's half has landed
                    mbar_arrive_cluster_relaxed(lead_full + 8 * stage);
                    if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
                }
            } else {
                // ================= leader: the static MMA schedule for both SMs =================
                MmaCtx c;
                c.bar_full = bar_full; c.bar_empty = bar_empty; c.ring_lo = umma_desc_lo(smem_u32(s_ring));
                c.idesc = (1u << 4) | ((uint32_t)(PAIR_CW >> 3) << 17) | ((256u >> 4) << 24);  // M = 256, N = 128
                c.ok = 0; c.pb = 0; c.no_mma = (p.dbg & 17) != 0;  /* bench-only: 1 = no weight stream at all, 16 = stream runs but is not waited for */ c.trace_tile = -1; c.trace_step = 0; c.prm = &p; c.tile_in_step = 0;
                c.empty_mask = p.quad ? (uint16_t)0xF : pair_mask;
                const uint32_t b_h0 = umma_desc_lo(smem_u32(s_h0)), b_h1 = umma_desc_lo(smem_u32(s_h1));
                const uint32_t dR = tmem_base, dZ = tmem_base + PAIR_CW, dNI = tmem_base + 2 * PAIR_CW, dNH = tmem_base + 3 * PAIR_CW;
                uint32_t use[4] = {0, 0, 0, 0};  // how often each accumulator region has been filled
                uint32_t hphase0 = 0, hphase1 = 0;
                auto acquire = [&](int a) {  // TMEM write-after-read only: no data flows through this wait
                    mbar_wait(bar_aempty + 8 * a, (use[a] & 1) ^ 1);
                    tc_fence_after();
                };
                auto publish = [&](int a) {
                    umma_commit2(bar_afull + 8 * a, pair_mask);
                    ++use[a];
                };
                for (int step = 0; step < N; ++step) {
                    c.tile_in_step = 0; c.trace_step = step;
                    for (int j = 0; j < JOBS2; ++j) {  // layer 0: R, Z, NH from h0
                        trace_ev(p, step, 2 * j);
                        acquire(0); mma_run2<KH, 0>(c, dR, b_h0, true); publish(0);
                        acquire(1); mma_run2<KH, 1>(c, dZ, b_h0, true); publish(1);
                        acquire(3); mma_run2<KH, 2>(c, dNH, b_h0, true); publish(3);
                        trace_ev(p, step, 2 * j + 1);
                    }
                    for (int j = 0; j < JOBS2; ++j) {  // layer 1: R, Z, NH from h1 (previous step), then R, Z, NI from h0
                        trace_ev(p, step, 8 + 2 * j);
                        c.trace_tile = (j == JOBS2 - 1) ? 0 : -1;
                        c.trace_step = step;
                        if (j == 0 && step > 0) {
                            mbar_wait_cluster(bar_hready + 8, hphase1);
                            hphase1 ^= 1;
                            tc_fence_after();
                        }
                        acquire(0); mma_run2<KH, 0>(c, dR, b_h1, true);
                        acquire(1); mma_run2<KH, 1>(c, dZ, b_h1, true);
                        acquire(3); mma_run2<KH, 2>(c, dNH, b_h1, true); publish(3);
                        if (j == 0) {
                            trace_ev(p, step, 18);
                            mbar_wait_cluster(bar_hready + 0, hphase0);
                            hphase0 ^= 1;
                            tc_fence_after();
                            trace_ev(p, step, 16);
                        }
                        mma_run2<KH, 0>(c, dR, b_h0, false); publish(0);
                        mma_run2<KH, 1>(c, dZ, b_h0, false); publish(1);
                        acquire(2); mma_run2<KH, 2>(c, dNI, b_h0, true); publish(2);
                        trace_ev(p, step, 8 + 2 * j + 1);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3, cq = warp >> 2;
        const int col0 = cq * CW3;                       // first of this thread's 32 codeword columns (pair index)
        const uint32_t owner = (uint32_t)(cq >> 1);      // CTA of the pair that holds these codewords' hidden states
        const int row0 = col0 & (TILE_B - 1);            // their rows in the owner's B-operand buffers
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        // bench-only experiment (results are garbage): dbg & 8 = every state access goes to this CTA's own memory
        const uint32_t h0_owner = mapa_u32(smem_u32(s_h0), lead + ((p.dbg & 8) ? rank : owner)),
                       h1_owner = mapa_u32(smem_u32(s_h1), lead + ((p.dbg & 8) ? rank : owner));
        const uint32_t peer = crank ^ 1u;
        const uint32_t part_peer = mapa_u32(smem_u32(s_part), peer), head_peer = mapa_u32(bar_head, peer);
        const uint32_t bits_peer = mapa_u32(smem_u32(s_bits), peer);
        uint32_t use[4] = {0, 0, 0, 0};
        uint32_t staged[(JOBS2 > 1 ? JOBS2 - 1 : 1) * (CW3 / 2)];
        const uint32_t info0 = p.info_words[0], info1 = N > 32 ? p.info_words[1] : 0u,
                       info2 = N > 64 ? p.info_words[2] : 0u, info3 = N > 96 ? p.info_words[3] : 0u;

        // drain accumulator a (32 columns) into v[], then hand the region back to the leader's MMA thread
        auto wait_acc = [&](int a) {
            mbar_wait(bar_afull + 8 * a, use[a] & 1);
            ++use[a];
            tc_fence_after();
        };
        auto release_acc = [&](int a) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(lead_aempty + 8 * a);
        };

        uint32_t head_phase = 0;  // phases of bar_head consumed so far (not every step exchanges partial logits)
        for (int step = 0; step < N; ++step) {
            const uint32_t bits = s_bits[cq];
            // a position outside the loss set decides +1 (or its genie value) whatever the logit: unless the caller
            // wants the logits, the head reduction and the cross-CTA exchange are skipped on those steps
            const uint32_t iw_step = step < 32 ? info0 : step < 64 ? info1 : step < 96 ? info2 : info3;
            const bool need_head = ((iw_step >> (step & 31)) & 1u) || p.logits != nullptr;
            for (int layer = 0; layer < 2; ++layer) {
                const uint32_t h_owner = layer ? h1_owner : h0_owner;
                if (layer == 1 && need_head && lane < 8) {
#pragma unroll
                    for (int c8 = 0; c8 < CW3 / 8; ++c8) s_headacc[q * PAIR_CW + col0 + c8 * 8 + lane] = 0.0f;
                }
#pragma unroll
                for (int j = 0; j < JOBS2; ++j) {
                    const int u = j * 256 + (int)rank * 128 + q * 32 + lane;  // hidden unit of this thread
                    float hr0, hr1, hz0, hz1, bn0, bn1, b_hn, wo = 0.f;
                    if (layer == 0) {
                        const float4 c0 = __ldg(reinterpret_cast<const float4 *>(p.consts0 + (size_t)u * 12));
                        const float4 c1 = __ldg(reinterpret_cast<const float4 *>(p.consts0 + (size_t)u * 12 + 4));
                        const float2 c2 = __ldg(reinterpret_cast<const float2 *>(p.consts0 + (size_t)u * 12 + 8));
                        hr0 = 0.5f * (c0.x + c1.x); hr1 = 0.5f * (c0.x + c1.y);
                        hz0 = 0.5f * (c0.y + c1.z); hz1 = 0.5f * (c0.y + c1.w);
                        bn0 = c0.z + c2.x; bn1 = c0.z + c2.y;
                        b_hn = c0.w;
                    } else {
                        const float4 c0 = __ldg(reinterpret_cast<const float4 *>(p.consts1 + (size_t)u * 4));
                        hr0 = hr1 = 0.5f * c0.x; hz0 = hz1 = 0.5f * c0.y; bn0 = bn1 = c0.z; b_hn = c0.w;
                        wo = __ldg(p.w_out + u);
                    }
                    const uint32_t t0 = tmem_base + lane_addr + col0;
                    uint32_t rp[CW3 / 2], zp[CW3 / 2];  // r and z (NPD_GRU_ZC: 1 - z) of the 32 columns as fp16 pairs
                    // ---- R, then Z: drained as soon as they complete ----
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        wait_acc(a);
#pragma unroll
                        for (int cc = 0; cc < CW3; cc += 8) {
                            float v[8];
                            tmem_ld8(t0 + a * PAIR_CW + cc, v);
                            tmem_ld_wait();
                            if (cc + 8 == CW3) release_acc(a);
                            if (layer == 0) {
                                const uint4 g4 = gy[(j * 3 + a) * (CW3 / 8) + cc / 8];
                                const uint32_t w4[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                                for (int h2 = 0; h2 < 4; ++h2) {
                                    const float2 f = unpack_h2(w4[h2]);
                                    v[2 * h2] += f.x; v[2 * h2 + 1] += f.y;
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 8; i += 2) {
                                const bool p0 = (layer == 0) && ((bits >> (cc + i)) & 1u), p1 = (layer == 0) && ((bits >> (cc + i + 1)) & 1u);
                                const float x0 = fmaf(v[i], 0.5f, a == 0 ? (p0 ? hr1 : hr0) : (p0 ? hz1 : hz0));
                                const float x1 = fmaf(v[i + 1], 0.5f, a == 0 ? (p1 ? hr1 : hr0) : (p1 ? hz1 : hz0));
                                float s0, s1;
                                if (a == 0) {
                                    s0 = sigmoid_half_arg_t<(NPD_GRU_ACT & 2) != 0>(x0);
                                    s1 = sigmoid_half_arg_t<(NPD_GRU_ACT & 2) != 0>(x1);
                                } else if (NPD_GRU_ZC) {
                                    s0 = sigmoid_compl_half_arg(x0);
                                    s1 = sigmoid_compl_half_arg(x1);
                                } else {
                                    s0 = sigmoid_half_arg(x0);
                                    s1 = sigmoid_half_arg(x1);
                                }
                                (a == 0 ? rp : zp)[(cc + i) >> 1] = pack_h2(s0, s1);
                            }
                        }
                    }
                    // ---- the job's last accumulators: NH (layer 0; NI is the hoisted projection), NH + NI (layer 1) ----
                    if (warp == 0) trace_ev(p, step, 19 + 0 * (layer * 4 + j));
#if NPD_GRU_LO
                    // residual of this thread's 32 state values, 8 bytes (8 codewords) per chunk, laid out
                    // [pair][layer][unit / 32][chunk of 8 codewords][unit % 32][8] so that a warp's access is 256 contiguous bytes
                    // (the kernel streams ~8 TB/s of weights from L2: half-used 32-byte sectors here cost 4.5 ms per launch)
                    uint2 *lo_base = reinterpret_cast<uint2 *>(p.h_lo) +
                                     ((((size_t)(blockIdx.x >> 1) * 2 + layer) * (H / 32) + (u >> 5)) * (PAIR_CW / 8) + (col0 >> 3)) * 32 + (u & 31);
#ifndef NPD_GRU_LO_AHEAD
#define NPD_GRU_LO_AHEAD 1  // chunks fetched ahead of their use (1..4): measured flat, the cost is L2 bandwidth, not latency
#endif
                    uint2 lo_q[NPD_GRU_LO_AHEAD];
#pragma unroll
                    for (int c = 0; c < NPD_GRU_LO_AHEAD; ++c) lo_q[c] = make_uint2(0u, 0u);
                    const bool use_lo = p.h_lo != nullptr;  // npd_gru_set_option(NPD_GRU_OPT_RESIDUAL_STATE): uniform
                    if (use_lo && step > 0) {
#pragma unroll
                        for (int c = 0; c < NPD_GRU_LO_AHEAD; ++c) lo_q[c] = __ldcg(lo_base + 32 * c);
                    }
#endif
                    wait_acc(3);
                    if (layer == 1) wait_acc(2);
                    if (warp == 0) trace_ev(p, step, 20 + 2 * (layer * 4 + j));
#pragma unroll
                    for (int cc = 0; cc < CW3; cc += 8) {
                        float aNH[8], aNI[8];
                        unsigned short hold[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) hold[i] = ld_cluster_u16(h_owner + b_off(row0 + cc + i, u));
#if NPD_GRU_LO
                        uint2 *lo_ptr = lo_base + 32 * (cc / 8);
                        const uint2 lo4 = lo_q[(cc / 8) % NPD_GRU_LO_AHEAD];  // fetched NPD_GRU_LO_AHEAD chunks ahead
                        if (use_lo && cc / 8 + NPD_GRU_LO_AHEAD < CW3 / 8 && step > 0) lo_q[(cc / 8) % NPD_GRU_LO_AHEAD] = __ldcg(lo_ptr + 32 * NPD_GRU_LO_AHEAD);
#endif
                        tmem_ld8(t0 + 3 * PAIR_CW + cc, aNH);
                        if (layer == 1) tmem_ld8(t0 + 2 * PAIR_CW + cc, aNI);
                        tmem_ld_wait();
                        if (cc + 8 == CW3) {
                            release_acc(3);
                            if (layer == 1) release_acc(2);
                        }
                        if (layer == 0) {
                            const uint4 g4 = gy[(j * 3 + 2) * (CW3 / 8) + cc / 8];
                            const uint32_t w4[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                            for (int h2 = 0; h2 < 4; ++h2) {
                                const float2 f = unpack_h2(w4[h2]);
                                aNI[2 * h2] = f.x; aNI[2 * h2 + 1] = f.y;
                            }
                        }
                        float hsum[8];
#if NPD_GRU_LO
                        const uint32_t lo_in[2] = {lo4.x, lo4.y};
                        uint32_t lo_out[2] = {0u, 0u};
#endif
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const bool plus = (layer == 0) && ((bits >> (cc + i)) & 1u);
                            const float2 r2 = unpack_h2(rp[(cc + i) >> 1]), z2 = unpack_h2(zp[(cc + i) >> 1]);
                            const float r = (i & 1) ? r2.y : r2.x, z = (i & 1) ? z2.y : z2.x;
                            const float nn = tanh_f(fmaf(r, aNH[i] + b_hn, aNI[i] + (plus ? bn1 : bn0)));
                            const float ho = __half2float(__ushort_as_half(hold[i]));
                            float hnew = NPD_GRU_ZC ? fmaf(-z, ho - nn, ho)   // h - (1 - z) (h - n), `z` holds 1 - z
                                                    : fmaf(z, ho - nn, nn);   // (1 - z) n + z h
#if NPD_GRU_LO
                            // the residual of the old state enters last (its load has the whole gate math to land): + z lo
                            {
                                int q8;  // signed byte i of the chunk, sign-extended
                                asm("bfe.s32 %0, %1, %2, 8;" : "=r"(q8) : "r"(lo_in[i >> 2]), "r"(8 * (i & 3)));
                                hnew = fmaf((float)q8 * 1.9073486328125e-06f, NPD_GRU_ZC ? 1.0f - z : z, hnew);  // 2^-19
                            }
#endif
                            hsum[i] = wo * hnew;
                            const unsigned short hb = __half_as_ushort(__float2half_rn(hnew));
#if NPD_GRU_LO
                            {
                                int q8;  // residual in units of 2^-19, saturated to a signed byte (|residual| <= 2^-12 while |h| < 1)
                                asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(q8) : "f"((hnew - __half2float(__ushort_as_half(hb))) * 524288.0f));
                                asm("bfi.b32 %0, %1, %0, %2, 8;" : "+r"(lo_out[i >> 2]) : "r"(q8), "r"(8 * (i & 3)));
                            }
#endif
                            if (j == JOBS2 - 1) {
                                // the last accumulator of the layer's last job is full: every MMA reading the old state retired
                                st_cluster_u16(h_owner + b_off(row0 + cc + i, u), hb);
                            } else {
                                const int si = j * (CW3 / 2) + ((cc + i) >> 1);
                                if ((i & 1) == 0) staged[si] = hb; else staged[si] |= (uint32_t)hb << 16;
                            }
                        }
#if NPD_GRU_LO
                        if (use_lo) __stcg(lo_ptr, make_uint2(lo_out[0], lo_out[1]));
#endif
                        if (layer == 1 && need_head) {
                            // head: reduce the 8 columns over the warp's 32 units; lane l < 8 ends with column cc + l
#pragma unroll
                            for (int s = 4; s >= 1; s >>= 1) {
#pragma unroll
                                for (int i = 0; i < s; ++i) {
                                    const bool up = (lane & s) != 0;
                                    const float send = up ? hsum[i] : hsum[i + s];
                                    const float keep = up ? hsum[i + s] : hsum[i];
                                    hsum[i] = keep + __shfl_xor_sync(NPD_FULL, send, s);
                                }
                            }
                            hsum[0] += __shfl_xor_sync(NPD_FULL, hsum[0], 8);
                            hsum[0] += __shfl_xor_sync(NPD_FULL, hsum[0], 16);
                            if (lane < 8) s_headacc[q * PAIR_CW + col0 + cc + lane] += hsum[0];
                        }
                    }
                    if (warp == 0) trace_ev(p, step, 21 + 2 * (layer * 4 + j));
                }
#pragma unroll
                for (int j = 0; j < JOBS2 - 1; ++j) {
                    const int u = j * 256 + (int)rank * 128 + q * 32 + lane;
#pragma unroll
                    for (int i = 0; i < CW3 / 2; ++i) {
                        const uint32_t pr = staged[j * (CW3 / 2) + i];
                        st_cluster_u16(h_owner + b_off(row0 + 2 * i, u), (unsigned short)(pr & 0xffffu));
                        st_cluster_u16(h_owner + b_off(row0 + 2 * i + 1, u), (unsigned short)(pr >> 16));
                    }
                }
                // the state rows (in either CTA) are read by the tensor cores of both SMs: make the stores visible to
                // the async proxy, then one release-arrive per warp on the leader's barrier
                asm volatile("fence.proxy.async;" ::: "memory");
                __syncwarp();
                // layer 1's state is first read by the NEXT step's layer-1 MMAs: its (slow) release is deferred until the
                // head exchange below is on its way
                if (layer == 0 && lane == 0) {
                    if (owner == rank) {
                        // the state rows this warp wrote live in this CTA: order them locally, then a plain signal
                        asm volatile("fence.acq_rel.cta;" ::: "memory");
                        mbar_arrive_cluster_relaxed(lead_hready + 8 * layer);
                    } else {
                        mbar_arrive_cluster(lead_hready + 8 * layer);  // remote rows: release at cluster scope (~0.8 us)
                    }
                }
            }

            // ---- head: this CTA's partial logits (over its 256 units) go to both CTAs ----
            if (need_head) epi_bar_sync();
            auto publish_h1 = [&]() {
                if (lane == 0) {
                    if (owner == rank) {
                        asm volatile("fence.acq_rel.cta;" ::: "memory");
                        mbar_arrive_cluster_relaxed(lead_hready + 8);
                    } else {
                        mbar_arrive_cluster(lead_hready + 8);
                    }
                }
            };
            if (warp >= 4) publish_h1();
            float part = 0.0f;
            if (warp < 4 && need_head) {
                const int c = warp * 32 + lane;  // codeword of the pair
                part = (s_headacc[c] + s_headacc[PAIR_CW + c]) + (s_headacc[2 * PAIR_CW + c] + s_headacc[3 * PAIR_CW + c]);
                st_cluster_f32(part_peer + c * 4, part);
            }
            if (warp < 4) {
                float logit = 0.0f;
                if (need_head) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(head_peer);  // release: the partial logits above are visible to the peer
                    mbar_wait_cluster(bar_head, head_phase & 1);
                    logit = (part + s_part[warp * 32 + lane]) + p.b_out;  // a + b == b + a: both CTAs get the same logit
                }
                const int c = warp * 32 + lane;
                const bool is_info = (iw_step >> (step & 31)) & 1u;
                const bool valid = pair0 + c < p.B;
                const bool mine = (uint32_t)(c >> 6) == rank;  // the owner CTA writes the outputs
                float dec = (p.genie && valid) ? p.genie[(pair0 + c) * N + step] : 1.0f;
                if (is_info) dec = (logit > 0.0f) ? 1.0f : ((logit < 0.0f) ? -1.0f : 0.0f);
                if (valid && mine) {
                    if (p.logits) p.logits[(pair0 + c) * N + step] = logit;
                    p.decoded[(pair0 + c) * N + step] = dec;
                }
                float prev = (dec > 0.0f) ? 1.0f : ((dec < 0.0f) ? -1.0f : 0.0f);
                if (p.forced && valid) prev = p.forced[(pair0 + c) * N + step];
                const uint32_t m = __ballot_sync(NPD_FULL, prev >= 1.0f);
                if (lane == 0) s_bits[warp] = m;  // both CTAs compute all 128 decisions identically
                publish_h1();
            }
            if (need_head) ++head_phase;
            epi_bar_sync();
            if (warp == 0) trace_ev(p, step, 36);
        }
        (void)bits_peer;
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

}  // namespace

// ---- host side ----------------------------------------------------------------------------------
static const void *gru_kernel3_for(int H)
{
    return H == 256 ? (const void *)gru_decode_kernel3<4> : (const void *)gru_decode_kernel3<8>;
}

static const void *gru_kernel_for(int H)
{
    switch (H) {
    case 128: return (const void *)gru_decode_kernel<2>;
    case 256: return (const void *)gru_decode_kernel<4>;
    case 384: return (const void *)gru_decode_kernel<6>;
    default: return (const void *)gru_decode_kernel<8>;
    }
}

struct npd_gru {
    int N, H, tiles_per_step;
    float b_out;
    unsigned char *d_wpack, *d_wpack2;
    CUtensorMap tmap2;   // [2 * tiles_per_step2 * 128 rows][128 B] view of d_wpack2 for the 2-SM tensor copies
    int have_tmap2;
    int tiles_per_step2;
    size_t smem_bytes3;
    float *d_w_iyT;
    float *d_consts0, *d_consts1, *d_w_out;
    size_t smem_bytes;
    int head_depth, head_yh;   // MLP head (npd_gru_set_head_mlp); depth 0 = the Linear(H,1) head of npd_gru_create
    int residual_state;        // npd_gru_set_option(NPD_GRU_OPT_RESIDUAL_STATE): 1 (default) = fp16 state + fp16 residual
    __half *d_head_w1, *d_head_wh;
    float *d_head_b, *d_head_wl;
};

namespace {

// write one 128 x 64 fp16 tile in K-major SWIZZLE_128B order; src(r, k) gives the fp32 weight
template <class F>
void pack_tile(std::vector<unsigned short> &out, F src)
{
    const size_t base = out.size();
    out.resize(base + 128 * 64);
    for (int r = 0; r < 128; ++r)
        for (int kk = 0; kk < 64; ++kk) {
            const size_t off = (size_t)r * 128 + ((((kk >> 3) ^ (r & 7)) << 4)) + (kk & 7) * 2;
            __half b = __float2half_rn(src(r, kk));
            out[base + off / 2] = *reinterpret_cast<unsigned short *>(&b);
        }
}

}  // namespace

NPD_API int npd_gru_create(int N, int H, const float *w_ih0, const float *w_hh0, const float *b_ih0,
                           const float *b_hh0, const float *w_ih1, const float *w_hh1, const float *b_ih1,
                           const float *b_hh1, const float *w_out, const float *b_out, npd_gru_t **out)
{
    NPD_REQUIRE(out, "npd_gru_create: null out");
    *out = nullptr;
    NPD_REQUIRE(w_ih0 && w_hh0 && b_ih0 && b_hh0 && w_ih1 && w_hh1 && b_ih1 && b_hh1 && w_out && b_out,
                "npd_gru_create: null weight pointer");
    if (N < 1 || N > 128 || H < 128 || H > 512 || (H % 128) != 0) {
        npd_set_error("npd_gru_create: supported envelope is 1 <= N <= 128, H in {128,256,384,512} (got N=%d H=%d)", N, H);
        return NPD_EUNSUPPORTED;
    }
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    {
        // operands are rounded to fp16 (fp32 accumulate): refuse weights outside its finite range
        const struct { const float *p; size_t n; } arrs[] = {
            {w_ih0, (size_t)3 * H * (N + 2)}, {w_hh0, (size_t)3 * H * H}, {w_ih1, (size_t)3 * H * H}, {w_hh1, (size_t)3 * H * H}};
        for (const auto &a : arrs)
            for (size_t i = 0; i < a.n; ++i)
                if (!(fabsf(a.p[i]) <= 65504.0f)) {
                    npd_set_error("npd_gru_create: weight magnitude %g is outside the fp16 range", (double)a.p[i]);
                    return NPD_EUNSUPPORTED;
                }
    }
    const int KH = H / 64, JOBS = H / 128, IN0 = N + 2;
    std::vector<unsigned short> pack;
    int n_tiles = 0;
    auto run = [&](const float *w, int gate, int j) {  // KH tiles: rows gate*H + j*128 .. +127 of a [3H, H] matrix
        for (int kc = 0; kc < KH; ++kc, ++n_tiles)
            pack_tile(pack, [&](int r, int kk) { return w[(size_t)(gate * H + j * 128 + r) * H + kc * 64 + kk]; });
    };
    // The order below is the MMA warp's static schedule (gru_decode_kernel): blocks of three runs.
    // layer 0, per job: R, Z, NH from h0 (the y part of R and Z and all of NI come from the hoisted projection)
    for (int j = 0; j < JOBS; ++j) {
        run(w_hh0, 0, j);
        run(w_hh0, 1, j);
        run(w_hh0, 2, j);
    }
    // layer 1, two jobs at a time: the hidden-state streams first (they need only the previous step's h1): NH, R_h,
    // Z_h of both jobs; then the input streams (they need this step's h0): R_x, Z_x, NI of both jobs
    for (int jp = 0; jp < JOBS; jp += 2) {
        const int nj = (JOBS - jp) < 2 ? (JOBS - jp) : 2;
        for (int i = 0; i < nj; ++i) {
            run(w_hh1, 2, jp + i);
            run(w_hh1, 0, jp + i);
            run(w_hh1, 1, jp + i);
        }
        for (int i = 0; i < nj; ++i) {
            run(w_ih1, 0, jp + i);
            run(w_ih1, 1, jp + i);
            run(w_ih1, 2, jp + i);
        }
    }
    // CTA-pair kernel (H a multiple of 256): rank r streams the unit rows 256 J + 128 r .. + 127 of every pair tile.
    // layer 0 per job: R, Z, NH (w_hh0); layer 1 per job: R, Z, NH (w_hh1, previous step's h1), then R, Z, NI (w_ih1)
    std::vector<unsigned short> pack2;
    int n_tiles2 = 0;
    if (H % 256 == 0) {
        for (int r = 0; r < 2; ++r) {
            int cnt = 0;
            auto run2 = [&](const float *w, int gate, int J) {
                for (int kc = 0; kc < KH; ++kc, ++cnt)
                    pack_tile(pack2, [&](int row, int kk) { return w[(size_t)(gate * H + J * 256 + r * 128 + row) * H + kc * 64 + kk]; });
            };
            for (int J = 0; J < H / 256; ++J) { run2(w_hh0, 0, J); run2(w_hh0, 1, J); run2(w_hh0, 2, J); }
            for (int J = 0; J < H / 256; ++J) {
                run2(w_hh1, 0, J); run2(w_hh1, 1, J); run2(w_hh1, 2, J);
                run2(w_ih1, 0, J); run2(w_ih1, 1, J); run2(w_ih1, 2, J);
            }
            n_tiles2 = cnt;
        }
    }
    std::vector<float> wiyT((size_t)N * 3 * H);
    for (int k = 0; k < N; ++k)
        for (int r = 0; r < 3 * H; ++r) wiyT[(size_t)k * 3 * H + r] = w_ih0[(size_t)r * IN0 + k];
    std::vector<float> c0((size_t)H * 12, 0.0f), c1((size_t)H * 4, 0.0f);
    for (int u = 0; u < H; ++u) {
        float *a = &c0[(size_t)u * 12];
        a[0] = b_ih0[u] + b_hh0[u];
        a[1] = b_ih0[H + u] + b_hh0[H + u];
        a[2] = b_ih0[2 * H + u];
        a[3] = b_hh0[2 * H + u];
        // onehot(-1) = [1,0] -> column N ; onehot(+1) = [0,1] -> column N+1 (rnn_all.py:258-260)
        a[4] = w_ih0[(size_t)u * IN0 + N];            a[5] = w_ih0[(size_t)u * IN0 + N + 1];
        a[6] = w_ih0[(size_t)(H + u) * IN0 + N];      a[7] = w_ih0[(size_t)(H + u) * IN0 + N + 1];
        a[8] = w_ih0[(size_t)(2 * H + u) * IN0 + N];  a[9] = w_ih0[(size_t)(2 * H + u) * IN0 + N + 1];
        float *b = &c1[(size_t)u * 4];
        b[0] = b_ih1[u] + b_hh1[u];
        b[1] = b_ih1[H + u] + b_hh1[H + u];
        b[2] = b_ih1[2 * H + u];
        b[3] = b_hh1[2 * H + u];
    }
    npd_gru *g = (npd_gru *)calloc(1, sizeof(npd_gru));
    if (!g) return NPD_ENOMEM;
    g->N = N; g->H = H; g->tiles_per_step = n_tiles; g->b_out = b_out[0];
    g->residual_state = 1;
    g->smem_bytes = Smem::total(H);
    g->tiles_per_step2 = n_tiles2;
    g->smem_bytes3 = Smem3::total(H);
    if (g->smem_bytes > (size_t)dp.smem_optin) {
        npd_set_error("npd_gru_create: needs %zu B of shared memory (limit %d)", g->smem_bytes, dp.smem_optin);
        free(g);
        return NPD_EUNSUPPORTED;
    }
    cudaError_t e = cudaMalloc(&g->d_wpack, pack.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_w_iyT, wiyT.size() * 4);
    if (e == cudaSuccess && n_tiles2) e = cudaMalloc(&g->d_wpack2, pack2.size() * 2);
    if (e == cudaSuccess && n_tiles2) e = cudaMemcpy(g->d_wpack2, pack2.data(), pack2.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_tiles2)
        e = cudaFuncSetAttribute(gru_kernel3_for(H), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_bytes3);
    if (e == cudaSuccess && n_tiles2) {
        // tensor map over the half-tile stream; without the driver entry point the pair kernel uses its relay variant
        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
            qres == cudaDriverEntryPointSuccess) {
            const cuuint64_t gdim[2] = {128, (cuuint64_t)2 * n_tiles2 * 128};
            const cuuint64_t gstride[1] = {128};
            const cuuint32_t box[2] = {128, 128}, estride[2] = {1, 1};
            const CUresult r = ((EncodeFn)fn)(&g->tmap2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, g->d_wpack2, gdim, gstride, box, estride,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            g->have_tmap2 = (r == CUDA_SUCCESS);
        } else {
            (void)cudaGetLastError();
        }
    }
    if (e == cudaSuccess) e = cudaMalloc(&g->d_consts0, c0.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_consts1, c1.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_w_out, (size_t)H * 4);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_wpack, pack.data(), pack.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_w_iyT, wiyT.data(), wiyT.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_consts0, c0.data(), c0.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_consts1, c1.data(), c1.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_w_out, w_out, (size_t)H * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const void *kern = gru_kernel_for(H);
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_bytes);
    }
    if (e != cudaSuccess) {
        npd_set_error("npd_gru_create: %s", cudaGetErrorString(e));
        npd_gru_destroy(g);
        return NPD_ECUDA;
    }
    *out = g;
    return NPD_OK;
}

NPD_API int npd_gru_destroy(npd_gru_t *g)
{
    if (!g) return NPD_OK;
    cudaFree(g->d_wpack);
    cudaFree(g->d_w_iyT);
    cudaFree(g->d_wpack2);
    cudaFree(g->d_consts0);
    cudaFree(g->d_consts1);
    cudaFree(g->d_w_out);
    cudaFree(g->d_head_w1); cudaFree(g->d_head_wh); cudaFree(g->d_head_b); cudaFree(g->d_head_wl);
    free(g);
    return NPD_OK;
}

NPD_API int npd_gru_set_head_mlp(npd_gru_t *g, int depth, int Yh, const float *h_params)
{
    NPD_REQUIRE(g && h_params, "npd_gru_set_head_mlp: null argument");
    if (depth < 2 || depth > 8 || Yh < 16 || Yh > 1024 || (Yh % 16) != 0) {
        npd_set_error("npd_gru_set_head_mlp: supported envelope is 2 <= depth <= 8, y_hidden_size a multiple of 16 in "
                      "[16, 1024] (got depth=%d, y_hidden_size=%d)", depth, Yh);
        return NPD_EUNSUPPORTED;
    }
    const int H = g->H;
    const size_t y = (size_t)Yh;
    std::vector<__half> w1(y * H), wh((size_t)(depth - 2) * y * y + 1);
    std::vector<float> b((size_t)(depth - 1) * y);
    const float *q = h_params;
    for (size_t i = 0; i < y * H; ++i) w1[i] = __float2half_rn(q[i]);
    q += y * H;
    for (size_t i = 0; i < y; ++i) b[i] = q[i];
    q += y;
    for (int l = 0; l < depth - 2; ++l) {
        for (size_t i = 0; i < y * y; ++i) wh[(size_t)l * y * y + i] = __float2half_rn(q[i]);
        q += y * y;
        for (size_t i = 0; i < y; ++i) b[(size_t)(l + 1) * y + i] = q[i];
        q += y;
    }
    cudaFree(g->d_head_w1); cudaFree(g->d_head_wh); cudaFree(g->d_head_b); cudaFree(g->d_head_wl);
    g->d_head_w1 = g->d_head_wh = nullptr; g->d_head_b = g->d_head_wl = nullptr;
    g->head_depth = 0;
    cudaError_t e = cudaMalloc(&g->d_head_w1, w1.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_head_wh, wh.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_head_b, b.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_head_wl, y * 4);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_head_w1, w1.data(), w1.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_head_wh, wh.data(), wh.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_head_b, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_head_wl, q, y * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        npd_set_error("npd_gru_set_head_mlp: %s", cudaGetErrorString(e));
        return NPD_ECUDA;
    }
    g->b_out = q[y];
    g->head_depth = depth; g->head_yh = Yh;
    return NPD_OK;
}

// bytes of the pair kernel's residual-state buffer: [pair][2 layers][H][128 codewords] signed bytes
static size_t gru_lo_bytes(const npd_gru *g, int64_t B)
{
    if (!NPD_GRU_LO || !g->residual_state || !g->d_wpack2 || g->head_depth > 1) return 0;
    return (size_t)((B + 2 * TILE_B - 1) / (2 * TILE_B) + 1) * 2 * g->H * (2 * TILE_B);
}

NPD_API int npd_gru_set_option(npd_gru_t *g, int option, int value)
{
    NPD_REQUIRE(g, "npd_gru_set_option: null handle");
    NPD_REQUIRE(option == NPD_GRU_OPT_RESIDUAL_STATE, "npd_gru_set_option: unknown option %d", option);
    g->residual_state = value != 0;
    return NPD_OK;
}

NPD_API size_t npd_gru_workspace_bytes(const npd_gru_t *g, int64_t B)
{
    if (!g || B <= 0) return 0;
    if (g->head_depth <= 1) return gru_lo_bytes(g, B);  // optional: without it the library's own pool is used
    return (size_t)((B + TILE_B - 1) / TILE_B + 1) * 2 * TILE_B * g->head_yh * sizeof(__half);  // +1: the padding CTA of a cluster
}

NPD_API int npd_gru_decode(const npd_gru_t *g, const npd_code_t *code, const float *y, const float *forced,
                           const float *genie, float *logits, float *decoded, int64_t B, void *ws, size_t ws_bytes,
                           void *stream)
{
    return npd_gru_decode_h0(g, code, y, nullptr, forced, genie, logits, decoded, B, ws, ws_bytes, stream);
}

NPD_API int npd_gru_decode_h0(const npd_gru_t *g, const npd_code_t *code, const float *y, const float *h0,
                              const float *forced, const float *genie, float *logits, float *decoded, int64_t B, void *ws,
                              size_t ws_bytes, void *stream)
{
    NPD_REQUIRE(g && code && y && decoded, "npd_gru_decode: null argument");
    NPD_REQUIRE(B >= 0, "npd_gru_decode: negative batch");
    NPD_REQUIRE(code->N == g->N, "npd_gru_decode: code length %d != decoder input length %d", code->N, g->N);
    if (B == 0) return NPD_OK;
    GruParams p{};
    p.wpack = g->d_wpack; p.w_iyT = g->d_w_iyT; p.consts0 = g->d_consts0; p.consts1 = g->d_consts1;
    p.w_out = g->d_w_out; p.b_out = g->b_out; p.y = y; p.h0 = h0; p.forced = forced; p.genie = genie; p.info_words = code->d_info_words;
    p.logits = logits; p.decoded = decoded; p.B = B; p.N = g->N; p.H = g->H;
    p.tiles_per_step = g->tiles_per_step;
    p.wpack2 = g->d_wpack2; p.tiles_per_step2 = g->tiles_per_step2;
    // CTA-pair kernel (cta_group::2) when the weights were packed for it; NPD_GRU_PAIR=0 selects the single-CTA kernel
    bool use_pair = g->d_wpack2 != nullptr;
    if (g->head_depth > 1) {
        const size_t need = npd_gru_workspace_bytes(g, B);
        NPD_REQUIRE(ws && ws_bytes >= need, "npd_gru_decode: a decoder with an MLP head needs %zu bytes of workspace (got %zu)",
                    need, ws ? ws_bytes : (size_t)0);
        use_pair = false;  // the MLP head lives in the single-CTA kernel
        p.head_depth = g->head_depth; p.head_yh = g->head_yh; p.head_w1 = g->d_head_w1; p.head_wh = g->d_head_wh;
        p.head_b = g->d_head_b; p.head_wl = g->d_head_wl; p.head_act = (__half *)ws;
    }
    { const char *d = npd_knob("NPD_GRU_PAIR"); if (d) use_pair = use_pair && atoi(d) != 0; }
    // residual of the fp16 recurrent state (pair kernel): the caller's workspace when it is large enough, else a
    // stream-ordered allocation from the library's pool (memory stays cached across calls)
    const size_t lo_need = use_pair ? gru_lo_bytes(g, B) : 0;
    void *lo_owned = nullptr;
    if (lo_need) {
        if (ws && ws_bytes >= lo_need) {
            p.h_lo = (signed char *)ws;
        } else {
            cudaMemPool_t pool;
            if (int rc = npd_scratch_pool(&pool)) return rc;
            NPD_CHECK_CUDA(cudaMallocFromPoolAsync(&lo_owned, lo_need, pool, (cudaStream_t)stream));
            p.h_lo = (signed char *)lo_owned;
        }
    }
    // the 2-SM tensor-copy variant (no relay) works but measures slower than relay + linear bulk copies (12.4 vs 11.4 ms
    // per 37888 codewords): opt-in with NPD_GRU_TMAP=1
    // NPD_GRU_QUAD=1: clusters of four (two MMA pairs that take turns fetching every half-tile and multicast it to the CTA
    // with the same pair rank in the other pair: half the L2 reads). Measured 12.7 ms against 10.95 ms for plain pairs at
    // 37888 codewords -- a ring slot then frees only when BOTH pairs have consumed it, which couples their schedules --
    // so plain pairs stay the default.
    p.quad = 0;
    { const char *d = npd_knob("NPD_GRU_QUAD"); if (d) p.quad = use_pair && atoi(d) != 0; }
    p.use_tmap = 0;
    { const char *d = npd_knob("NPD_GRU_TMAP"); if (d) p.use_tmap = g->have_tmap2 && atoi(d) != 0; }
    if (p.use_tmap) p.quad = 0;
    { const char *d = npd_knob("NPD_GRU_DBG"); p.dbg = d ? atoi(d) : 0; }
    int64_t grid = ((B + 2 * TILE_B - 1) / (2 * TILE_B)) * 2;  // CTA pairs (clusters of 2); an odd tile count pads with an idle-data CTA
    if (p.quad) grid = (grid + 3) / 4 * 4;
    const char *trace_path = npd_knob("NPD_GRU_TRACE");  // bench-only: dump CTA 0's event clocks (synchronises!)
    if (trace_path) NPD_CHECK_CUDA(cudaMalloc(&p.trace, sizeof(long long) * g->N * TRACE_SLOTS));
    if (trace_path) NPD_CHECK_CUDA(cudaMemsetAsync(p.trace, 0, sizeof(long long) * g->N * TRACE_SLOTS, (cudaStream_t)stream));
    CUtensorMap tm = g->tmap2;
    void *args[] = {&p, &tm};  // the single-CTA kernel takes only the first
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = use_pair ? g->smem_bytes3 : g->smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.quad ? 4 : 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NPD_CHECK_CUDA(cudaLaunchKernelExC(&cfg, use_pair ? gru_kernel3_for(g->H) : gru_kernel_for(g->H), args));
    NPD_CHECK_CUDA(cudaGetLastError());
    if (lo_owned) NPD_CHECK_CUDA(cudaFreeAsync(lo_owned, (cudaStream_t)stream));
    if (trace_path) {
        std::vector<long long> h((size_t)g->N * TRACE_SLOTS);
        NPD_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        NPD_CHECK_CUDA(cudaMemcpy(h.data(), p.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(p.trace);
        if (FILE *f = fopen(trace_path, "w")) {
            for (int s = 0; s < g->N; ++s) {
                for (int k = 0; k < TRACE_SLOTS; ++k) fprintf(f, "%lld ", h[(size_t)s * TRACE_SLOTS + k]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    return NPD_OK;
}
