// gru_train_tc.cuh -- one GRU layer-step of the training forward pass as ONE tcgen05 kernel (TF32 operands straight from
// the fp32 master weights and the saved fp32 state, fp32 accumulation in TMEM, gate math in the epilogue).
//
// Replaces, in GEMM mode 1 (TF32) and for H a multiple of 128, the pair
//     gh[B,3H] = h_prev[B,H] . W_hh[3H,H]^T   (library GEMM, 17.5 us at B = 4096, H = 512: 1.3 waves of 2-SM tiles)
//     cell_fwd_kernel                          (19-26 us: re-reads gh and writes the five saved quantities)
// of rnn_all.py:387-398's nn.GRU step (gate order r, z, n; h' = (1 - z) n + z h).  A CTA owns 128 batch rows x 128 hidden
// units: D[row, gate g of unit u] for the three gates = 384 TMEM columns, K = H in blocks of 32 floats.
//   warp 8   : producer -- per K block one 16 KB box of h_prev rows and three 16 KB boxes of W_hh rows (gate g, units
//              u0..u0+127) by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B: a row of 32 floats is one 128-byte swizzle
//              row, i.e. the canonical K-major operand layout) through a 3-stage mbarrier ring;
//   warp 9   : one elected thread issues tcgen05.mma.kind::tf32 (M = 128, N = 128 per gate, K = 8 per instruction: a K step
//              is 32 bytes inside the swizzle row, like K = 16 halves) and commits stages / the accumulator;
//   warps 0-7: epilogue -- thread = one batch row (TMEM lane) x 16 units of each 32-unit block: gates from the accumulators
//              + the input projection + biases, state update, the five saved quantities (r, z, n, W_hn h + b_hn, h').
//              All of its HBM traffic goes through the idle operand ring as TMA boxes (loads of the next block's inputs
//              and stores of the previous block's outputs run behind the gate math): with one 320-thread CTA per SM,
//              per-thread loads and stores cannot keep enough bytes in flight (first version: 61 us per launch, 60 % of
//              the stall samples on the epilogue's own loads).
// Step 0 (h_prev = 0) skips the GEMM.  The batch must be a multiple of 128 (a bulk store has no row mask).
#pragma once

#include <cuda.h>  // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint; libcuda is not linked)

#include "tc_common.cuh"

namespace gru_tc {

using namespace tc;

constexpr int TM = 128;                 // batch rows per CTA (TMEM lanes)
constexpr int TU = 128;                 // hidden units per CTA
constexpr int KB = 32;                  // floats per K block = one 128-byte swizzle row
constexpr int STAGES = 3;
constexpr int BOX_BYTES = TM * KB * 4;  // 16 KB: 128 rows x 32 floats, every TMA box of this kernel
constexpr int STAGE_BYTES = 4 * BOX_BYTES;  // h_prev rows + three gates' W_hh rows
// epilogue staging, on top of the (then idle) operand ring: inputs gi_r, gi_z, gi_n, h_prev of a block of 32 units, double
// buffered (boxes 0-7), outputs r, z, n, ghn, h' (boxes 8-12; box 12 lies behind the ring)
constexpr int N_BOXES = 13;
constexpr int OFF_BIAS = N_BOXES * BOX_BYTES;       // [6][TU] floats: b_ih (r, z, n), b_hh (r, z, n) of this CTA's units
constexpr int OFF_BARS = OFF_BIAS + 6 * TU * 4;
constexpr int SMEM_BYTES = OFF_BARS + 128 + 1024;   // + alignment slack
constexpr int THREADS = 320;
constexpr int EPI_WARPS = 8;

struct FwdParams {
    const float *b_ih, *b_hh;  // [3H]
    const float *wcol;     // layer 0: [2][3H] one-hot columns of W_ih0 (feedback -1 / 0, +1); else null
    const float *fb;       // layer 0: [B] feedback entering this step
    int64_t B;             // a multiple of TM
    int H;                 // a multiple of TU
    int gi_row0;           // first row of this step's input projection inside the gi tensor map
    int hprev_row0;        // first row of h_prev inside the state tensor map (unused when zero_h)
    int out_row0[5];       // first row of this step's r, z, n, ghn, h' inside the state tensor map
    int zero_h;            // step 0: h_prev = 0, no GEMM
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, int c0, int c1, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, tf32 x tf32 -> fp32 (operands are fp32 words; the tensor core reads 19 bits)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor: tf32 A/B (format 2), fp32 accumulate, both operands K-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// gate activations through ex2.approx + rcp.approx (relative error ~1e-7 on sigma, absolute ~1e-7 on tanh): TF32 inner
// products already carry 1e-3; the libm forms cost 7.7 k cycles per 32-unit block with 8 epilogue warps per SM
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950409f * x)); }
__device__ __forceinline__ float tanh_f(float x) { return 1.0f - 2.0f * rcp_approx(1.0f + ex2_approx(2.8853900818f * x)); }

// tm_s : the trainer's saved-state array as one [2 layers x 5 quantities x N steps x B, H] matrix -- operand A (h of the
//        previous step), the epilogue's h_prev and the five outputs are row ranges of it
// tm_w : W_hh of the layer, [3H, H]
// tm_gi: the layer's input projection, [B, 3H] (layer 0: the hoisted y part) or [N B, 3H] (layer 1: W_ih1 . h0 of all steps)
template <bool LAYER0>
__global__ void __launch_bounds__(THREADS, 1) gru_fwd_tc_kernel(const FwdParams p, const __grid_constant__ CUtensorMap tm_s,
                                                                const __grid_constant__ CUtensorMap tm_w,
                                                                const __grid_constant__ CUtensorMap tm_gi)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H;
    const int row0 = blockIdx.x * TM;
    const int u0 = blockIdx.y * TU;
    float *s_bias = reinterpret_cast<float *>(smem + OFF_BIAS);
    uint64_t *s_bars = reinterpret_cast<uint64_t *>(smem + OFF_BARS);
    const uint32_t bar_full = smem_u32(s_bars), bar_empty = bar_full + 8 * STAGES, bar_acc = bar_empty + 8 * STAGES,
                   bar_in_full = bar_acc + 8, bar_in_empty = bar_in_full + 16, bar_out_full = bar_in_empty + 16,
                   bar_out_empty = bar_out_full + 8;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bars + 2 * STAGES + 7);
    const int n_kb = p.zero_h ? 0 : H / KB;
    const uint32_t s0 = smem_u32(smem);
    constexpr int N_CB = TU / KB;  // column blocks of 32 units in the epilogue

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        mbar_init(bar_acc, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_in_full + 8 * i, 1);
            mbar_init(bar_in_empty + 8 * i, EPI_WARPS);
        }
        mbar_init(bar_out_full, EPI_WARPS);
        mbar_init(bar_out_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) tmem_alloc(smem_u32(s_tmem), 512);
    for (int i = tid; i < 6 * TU; i += THREADS) {
        const int q = i / TU, j = i - q * TU;  // q: 0-2 b_ih (r, z, n), 3-5 b_hh
        s_bias[i] = (q < 3 ? p.b_ih : p.b_hh)[(q % 3) * H + u0 + j];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 8) {
        if (lane == 0) {
            // ---- operand ring ----
            uint32_t stage = 0, phase = 0;
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                mbar_expect_tx(bar_full + 8 * stage, STAGE_BYTES);
                const uint32_t dst = s0 + stage * STAGE_BYTES;
                tma_load_2d(dst, &tm_s, kb * KB, p.hprev_row0 + row0, bar_full + 8 * stage);
#pragma unroll
                for (int g = 0; g < 3; ++g)
                    tma_load_2d(dst + (1 + g) * BOX_BYTES, &tm_w, kb * KB, g * H + u0, bar_full + 8 * stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            // ---- epilogue staging: the ring is idle once every MMA has completed ----
            if (n_kb > 0) mbar_wait(bar_acc, 0);
            auto load_inputs = [&](int cb) {
                const int buf = cb & 1;
                const uint32_t dst = s0 + buf * 4 * BOX_BYTES, bar = bar_in_full + 8 * buf;
                mbar_expect_tx(bar, (p.zero_h ? 3 : 4) * BOX_BYTES);
#pragma unroll
                for (int g = 0; g < 3; ++g) tma_load_2d(dst + g * BOX_BYTES, &tm_gi, g * H + u0 + cb * KB, p.gi_row0 + row0, bar);
                if (!p.zero_h) tma_load_2d(dst + 3 * BOX_BYTES, &tm_s, u0 + cb * KB, p.hprev_row0 + row0, bar);
            };
            load_inputs(0);
            load_inputs(1);
            for (int cb = 0; cb < N_CB; ++cb) {
                if (cb + 2 < N_CB) {  // refill the input buffer as soon as block cb's inputs sit in registers
                    mbar_wait(bar_in_empty + 8 * (cb & 1), (cb >> 1) & 1);
                    load_inputs(cb + 2);
                }
                mbar_wait(bar_out_full, cb & 1);
#pragma unroll
                for (int q = 0; q < 5; ++q) tma_store_2d(&tm_s, u0 + cb * KB, p.out_row0[q] + row0, s0 + (8 + q) * BOX_BYTES);
                bulk_commit();
                bulk_wait_read0();         // the stores have read the staging boxes
                mbar_arrive(bar_out_empty);
            }
            bulk_wait0();
        }
    } else if (warp == 9) {
        const uint32_t idesc = umma_idesc_tf32(TM, TU);
        uint32_t stage = 0, phase = 0;
        for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a = s0 + stage * STAGE_BYTES, b = a + BOX_BYTES;
#pragma unroll
                for (int k = 0; k < KB / 8; ++k)
#pragma unroll
                    for (int g = 0; g < 3; ++g)
                        umma_tf32(tmem_base + g * TU, umma_desc(a + k * 32), umma_desc(b + g * BOX_BYTES + k * 32), idesc,
                                  (kb | k) ? 1u : 0u);
                umma_commit(bar_empty + 8 * stage);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (n_kb > 0) {
            if (elect_one()) umma_commit(bar_acc);
            __syncwarp();
        }
    } else {
        // ================= epilogue: thread = batch row (TMEM lane) x 16 units of every 32-unit block =================
        const int quarter = warp & 3, half = warp >> 2;
        const int r_tile = quarter * 32 + lane;                    // row inside the tile = TMEM lane
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t row_off = (uint32_t)r_tile * 128;            // 128-byte rows; 16-byte chunk c sits at c ^ (row & 7)
        const uint32_t sw = (uint32_t)(r_tile & 7);
        const int64_t G = 3 * (int64_t)H;
        const float *wc = nullptr;
        if (LAYER0) wc = p.wcol + (p.fb[row0 + r_tile] > 0.0f ? G : 0);  // get_onehot (rnn_all.py:258-260)
        if (n_kb > 0) {
            mbar_wait(bar_acc, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int cb = 0; cb < N_CB; ++cb) {
            const int buf = cb & 1;
            const int c0 = cb * KB + half * 16;  // first unit (inside the tile) of this thread's 16
            float ar[16], az[16], an[16];
            if (n_kb > 0) {
                tmem_ld16(tmem_base + lane_addr + c0, ar);
                tmem_ld16(tmem_base + lane_addr + TU + c0, az);
                tmem_ld16(tmem_base + lane_addr + 2 * TU + c0, an);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) ar[i] = az[i] = an[i] = 0.0f;
            }
            mbar_wait(bar_in_full + 8 * buf, (cb >> 1) & 1);
            const unsigned char *in = smem + buf * 4 * BOX_BYTES + row_off;
            float4 gr[4], gz[4], gn[4], hp[4];
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
                const uint32_t ch = ((uint32_t)(half * 4 + v4) ^ sw) << 4;
                gr[v4] = *reinterpret_cast<const float4 *>(in + ch);
                gz[v4] = *reinterpret_cast<const float4 *>(in + BOX_BYTES + ch);
                gn[v4] = *reinterpret_cast<const float4 *>(in + 2 * BOX_BYTES + ch);
                hp[v4] = p.zero_h ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : *reinterpret_cast<const float4 *>(in + 3 * BOX_BYTES + ch);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_in_empty + 8 * buf);
            mbar_wait(bar_out_empty, (cb & 1) ^ 1);  // the previous block's stores have read the staging boxes
            unsigned char *out = smem + 8 * BOX_BYTES + row_off;
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
                float4 wr = make_float4(0.0f, 0.0f, 0.0f, 0.0f), wz = wr, wn = wr;
                if (LAYER0) {
                    const int u = u0 + c0 + 4 * v4;
                    wr = __ldg(reinterpret_cast<const float4 *>(wc + u));
                    wz = __ldg(reinterpret_cast<const float4 *>(wc + H + u));
                    wn = __ldg(reinterpret_cast<const float4 *>(wc + 2 * H + u));
                }
                const float *bi = s_bias + c0 + 4 * v4;  // [q][TU]
                const float gir[4] = {gr[v4].x, gr[v4].y, gr[v4].z, gr[v4].w}, giz[4] = {gz[v4].x, gz[v4].y, gz[v4].z, gz[v4].w},
                            gin[4] = {gn[v4].x, gn[v4].y, gn[v4].z, gn[v4].w}, hpv[4] = {hp[v4].x, hp[v4].y, hp[v4].z, hp[v4].w};
                const float wrv[4] = {wr.x, wr.y, wr.z, wr.w}, wzv[4] = {wz.x, wz.y, wz.z, wz.w}, wnv[4] = {wn.x, wn.y, wn.z, wn.w};
                float r4[4], z4[4], n4[4], g4[4], h4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = 4 * v4 + i;
                    float xr = gir[i] + bi[i], xz = giz[i] + bi[TU + i], xn = gin[i] + bi[2 * TU + i];
                    if (LAYER0) { xr += wrv[i]; xz += wzv[i]; xn += wnv[i]; }
                    const float ghr = ar[c] + bi[3 * TU + i], ghz = az[c] + bi[4 * TU + i], ghn = an[c] + bi[5 * TU + i];
                    const float r = sigmoid_f(xr + ghr);
                    const float z = sigmoid_f(xz + ghz);
                    const float n = tanh_f(xn + r * ghn);
                    r4[i] = r; z4[i] = z; n4[i] = n; g4[i] = ghn;
                    h4[i] = (1.0f - z) * n + z * hpv[i];
                }
                const uint32_t ch = ((uint32_t)(half * 4 + v4) ^ sw) << 4;
                *reinterpret_cast<float4 *>(out + ch) = make_float4(r4[0], r4[1], r4[2], r4[3]);
                *reinterpret_cast<float4 *>(out + BOX_BYTES + ch) = make_float4(z4[0], z4[1], z4[2], z4[3]);
                *reinterpret_cast<float4 *>(out + 2 * BOX_BYTES + ch) = make_float4(n4[0], n4[1], n4[2], n4[3]);
                *reinterpret_cast<float4 *>(out + 3 * BOX_BYTES + ch) = make_float4(g4[0], g4[1], g4[2], g4[3]);
                *reinterpret_cast<float4 *>(out + 4 * BOX_BYTES + ch) = make_float4(h4[0], h4[1], h4[2], h4[3]);
            }
            fence_async_smem();  // generic-proxy writes -> the bulk stores' reads
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_out_full);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// host: [rows, cols] row-major fp32 array as a 2-D tensor map with boxes of 32 floats x 128 rows, 128-byte swizzle
inline bool encode_map(CUtensorMap *tm, const float *base, uint64_t rows, uint64_t cols)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess || !f ||
            qres != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            return false;
        }
        fn = (EncodeFn)f;
    }
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {cols * 4};
    const cuuint32_t box[2] = {KB, 128}, estride[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estride,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gru_tc
