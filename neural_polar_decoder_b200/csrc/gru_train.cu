// gru_train.cu -- one training step of the CRISP GRU sequential decoder on the device.
//
// Replaces, for rnn_type GRU / decoding_type 'y_input' / onehot / 2 layers / Linear(H,1) head (the run_crisp.sh
// configuration), the body of the reference's training loop, rnn_all.py:1399-1437:
//     decoded = decoder.decode(net, True, y, gt, tfr)          teacher-forced (425-449) or student-forced (463-489)
//     loss    = MSELoss(decoded[:, info], msg_bits)             (1413)
//     loss.backward(); clip_grad_norm_(net.parameters(), clip)  (1431-1432)
//     optimizer.step()   [torch.optim.AdamW, default betas / eps / weight_decay]  (1346, 1435)
// All arithmetic is fp32 (the reference trains in fp32; parity = gradients and updated weights against the live
// reference at fp32 round-off in GEMM mode 0).  Structure (round 2, second version: layer-wise instead of step-wise
// wherever the data dependencies allow it):
//   forward   the y-part of W_ih0 hoisted out of the loop (one GEMM).  Teacher-forced: the feedback of
//             every step is known up front, so layer 0 runs its N steps (one recurrent GEMM [B,H] x [H,3H] + one fused
//             gate kernel per step), W_ih1 . h0 of ALL steps is ONE GEMM [N B,H] x [H,3H], layer 1 runs its N steps,
//             and the head (dot product, loss, d loss / d logit) is one launch over N B rows.  Student-forced: the
//             feedback of step t+1 is the sign of step t's logit, so the two layers and the head advance step by step.
//             The gate kernels SAVE r, z, n, (W_hn h + b_hn) and h, time-major per quantity ([layer][quantity][step][B H]),
//             so that "h of all steps" is one contiguous GEMM operand.
//   backward  layer 1 for t = N-1..0 (fused gate-gradient kernel -> dgi1[t], dgh1[t]; dh1 += dgh1[t] W_hh1), then for
//             ALL steps at once d x1 = dgi1 W_ih1, dW_ih1 = dgi1^T h0, dW_hh1 = dgh1[1:]^T h1[:-1]; layer 0 the same
//             way.  The detached feedback carries no gradient, so this order is valid for both forcing modes.  Weight
//             gradients are therefore 3 batched GEMMs (one slice per step, summed by a column-sum kernel) instead of 4 N
//             dependent accumulating GEMMs with K = B (55 % of the first version's time: 24 output tiles on 148 SMs).  Bias and one-hot-column gradients: every thread
//             of the gate-gradient kernel owns one unit of 8 consecutive rows and keeps their column sums in registers;
//             the [B/8, 7, H] partials are reduced once at the end (deterministic).
//   update    one fused kernel: global grad-norm clip coefficient + AdamW.
// The GEMMs are plain library GEMMs (cublasGemmEx on fp32 data; inner products in fp32, or on TF32 / bf16 / fp16 tensor
// cores with fp32 accumulation) -- except the forward pass's recurrent GEMMs in TF32 mode, which run together with their
// gate kernel as ONE hand-written tcgen05 kernel per layer-step (gru_train_tc.cuh: TMA-fed kind::tf32 MMAs into TMEM, gate
// math in the epilogue, TMA stores of the saved quantities) when H is a multiple of 128 and the batch of 128.
// Everything else is this file.
#include <cublas_v2.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "npd_common.cuh"
#include "gru_train_tc.cuh"

#define NPD_CHECK_CUBLAS(expr)                                                        \
    do {                                                                              \
        cublasStatus_t s__ = (expr);                                                  \
        if (s__ != CUBLAS_STATUS_SUCCESS) {                                           \
            npd_set_error("%s:%d: %s -> cuBLAS status %d", __FILE__, __LINE__, #expr, (int)s__); \
            return NPD_ECUDA;                                                         \
        }                                                                             \
    } while (0)

// fp32 in / fp32 out GEMM whose inner products run as: 0 = fp32 FMA (pedantic; the parity mode), 1 = TF32 tensor cores,
// 2 = bf16 tensor cores, 3 = fp16 tensor cores (cuBLAS down-converts the fp32 operands internally; accumulation stays fp32)
static cublasStatus_t gemm32(cublasHandle_t h, int mode, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                             const float *alpha, const float *A, int lda, const float *B, int ldb, const float *beta, float *C,
                             int ldc)
{
    const cublasComputeType_t ct = mode == 1 ? CUBLAS_COMPUTE_32F_FAST_TF32 : mode == 2 ? CUBLAS_COMPUTE_32F_FAST_16BF
                                 : mode == 3 ? CUBLAS_COMPUTE_32F_FAST_16F : CUBLAS_COMPUTE_32F_PEDANTIC;
    return cublasGemmEx(h, ta, tb, m, n, k, alpha, A, CUDA_R_32F, lda, B, CUDA_R_32F, ldb, beta, C, CUDA_R_32F, ldc, ct,
                        CUBLAS_GEMM_DEFAULT);
}

// the same GEMM for `batch` (A, B, C) triples at fixed strides
static cublasStatus_t gemm32_batched(cublasHandle_t h, int mode, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                                     const float *alpha, const float *A, int lda, long long sa, const float *B, int ldb,
                                     long long sb, const float *beta, float *C, int ldc, long long sc, int batch)
{
    const cublasComputeType_t ct = mode == 1 ? CUBLAS_COMPUTE_32F_FAST_TF32 : mode == 2 ? CUBLAS_COMPUTE_32F_FAST_16BF
                                 : mode == 3 ? CUBLAS_COMPUTE_32F_FAST_16F : CUBLAS_COMPUTE_32F_PEDANTIC;
    return cublasGemmStridedBatchedEx(h, ta, tb, m, n, k, alpha, A, CUDA_R_32F, lda, sa, B, CUDA_R_32F, ldb, sb, beta, C,
                                      CUDA_R_32F, ldc, sc, batch, ct, CUBLAS_GEMM_DEFAULT);
}

struct npd_gru_trainer {
    int N, H, I;        // code length (= steps), hidden size, layer-0 input width N + 2
    int64_t max_batch;
    size_t n_params;
    // parameter blob in state_dict order (rnn_all.py:307, 333-334): rnn.weight_ih_l0 [3H,I], weight_hh_l0 [3H,H],
    // bias_ih_l0, bias_hh_l0 [3H], weight_ih_l1 [3H,H], weight_hh_l1 [3H,H], bias_ih_l1, bias_hh_l1, linear.weight [H],
    // linear.bias [1]; grads / Adam moments use the same offsets
    float *p, *g, *m, *v;
    size_t o_wih0, o_whh0, o_bih0, o_bhh0, o_wih1, o_whh1, o_bih1, o_bhh1, o_wout, o_bout;
    int64_t step;       // optimizer step count (bias correction)
    float beta1, beta2, eps, weight_decay;
    cublasHandle_t blas;
    int gemm_mode;      // see gemm32
    // activations (sized for max_batch; addressed with the CALL's batch B, so every [step][B ..] block is contiguous)
    float *saved;       // [2 layers][5: r, z, n, ghn, h][N steps][B*H]
    float *gy;          // [B,3H]  y-part of the layer-0 input projection; after the backward pass: sum over steps of dgi0
    float *gh;          // [B,3H]  recurrent GEMM output of the current step
    float *dgi, *dgh;   // [N steps][B,3H]  gate gradients of the layer being processed; dgi doubles as W_ih1 . h0 of all
                        //                  steps in the forward pass
    float *dx1;         // [N steps][B,H]   gradient entering layer 0 from layer 1
    float *wpart;       // [N steps][3H,H]  per-step slices of a weight gradient (split-K by hand, see weight_grad)
    float *part;        // [2 layers][ceil(B/8)][7][H] column sums of the gate gradients, see cell_bwd_kernel
    float *dh, *zeros;  // [B,H]
    float *fb;          // [N][B] feedback entering step t (+-1)
    float *out, *dout, *lsq;  // [N][B] logits, d loss / d logit, squared error (0 off the loss set)
    float *scal;        // [4]: loss sum, grad norm^2, spare
    float *wcol;        // [2][3H] the one-hot columns of W_ih0, contiguous
    unsigned char *is_loss;    // [N] device copy of the loss set's indicator
    unsigned char *h_is_loss;  // [N] what is_loss holds (uploaded only when the loss set changes)
    int is_loss_valid;
    // fused TF32 layer-step kernel (gru_train_tc.cuh): GEMM mode 1, H a multiple of 128, tensor-map encoder available
    int tc_ok;
    int64_t tc_batch;            // batch the state tensor maps were built for
    CUtensorMap tm_s, tm_w[2], tm_gi[2];  // saved state [2 x 5 x N x B, H]; per layer W_hh [3H, H] and the input projection
};

namespace {

constexpr int kRowsPerThread = 8;  // cell_bwd_kernel: rows whose column sums one thread keeps in registers

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- forward gate kernel: one thread per (row b, unit j) -------------------------------------------------------
// LAYER0: gi = gy[b, g*H+j] + b_ih[g*H+j] + W_ih0[g*H+j, N + idx(fb[b])]   (one-hot feedback = column select; the two
//         columns are copied out contiguously once per iteration, wcol [2][3H] -- read in place they are 4 (N + 2) bytes
//         apart, one sector per lane)
// else  : gi = gi_buf[b, g*H+j] + b_ih[g*H+j]
// sv points at quantity 0 of this (layer, step); the five saved quantities are qs floats apart
template <bool LAYER0>
__global__ void __launch_bounds__(256) cell_fwd_kernel(const float *__restrict__ gi_src, const float *__restrict__ gh,
                                                       const float *__restrict__ b_ih, const float *__restrict__ b_hh,
                                                       const float *__restrict__ wcol,
                                                       const float *__restrict__ fb, const float *__restrict__ h_prev,
                                                       float *__restrict__ sv, size_t qs, int64_t B, int H)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int64_t b = idx / H;
    const int j = (int)(idx - b * H);
    float gir = gi_src[b * 3 * H + j] + b_ih[j];
    float giz = gi_src[b * 3 * H + H + j] + b_ih[H + j];
    float gin = gi_src[b * 3 * H + 2 * H + j] + b_ih[2 * H + j];
    if (LAYER0) {
        const float *wc = wcol + (fb[b] > 0.0f ? 3 * H : 0);  // get_onehot (rnn_all.py:258-260): +1 -> [0,1], -1 / 0 -> [1,0]
        gir += wc[j];
        giz += wc[H + j];
        gin += wc[2 * H + j];
    }
    const float ghr = gh[b * 3 * H + j] + b_hh[j];
    const float ghz = gh[b * 3 * H + H + j] + b_hh[H + j];
    const float ghn = gh[b * 3 * H + 2 * H + j] + b_hh[2 * H + j];
    const float r = sigmoidf_(gir + ghr);
    const float z = sigmoidf_(giz + ghz);
    const float n = tanhf(gin + r * ghn);
    const float hp = h_prev[idx];
    sv[idx] = r;
    sv[qs + idx] = z;
    sv[2 * qs + idx] = n;
    sv[3 * qs + idx] = ghn;
    sv[4 * qs + idx] = (1.0f - z) * n + z * hp;
}

// wcol[c][g] = W_ih0[g, N + c]
__global__ void __launch_bounds__(256) onehot_cols_kernel(const float *__restrict__ w_ih0, float *__restrict__ wcol, int I, int N, int G)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * G) wcol[i] = w_ih0[(size_t)(i % G) * I + N + i / G];
}

// feedback of all steps under teacher forcing: +1 into step 0 (rnn_all.py:444), gt[:, t-1] into step t (447)
__global__ void __launch_bounds__(256) teacher_fb_kernel(const float *__restrict__ gt, float *__restrict__ fb, int64_t B, int N)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * N) return;
    const int64_t s = idx / B, b = idx - s * B;
    fb[idx] = s == 0 ? 1.0f : gt[b * N + (s - 1)];
}

// ---- head: logit = h1 . w_out + b_out (one warp per row of [steps t0 .. t0 + nsteps) x B), loss terms, and under
// student forcing the next step's feedback ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_fwd_kernel(const float *__restrict__ h1, const float *__restrict__ w_out,
                                                       const float *__restrict__ b_out, const float *__restrict__ gt,
                                                       int N, int t0, int nsteps, const unsigned char *__restrict__ is_loss_t,
                                                       int write_fb, float inv_count, float *__restrict__ out,
                                                       float *__restrict__ dout, float *__restrict__ fb,
                                                       float *__restrict__ logits_out, float *__restrict__ lsq, int64_t B, int H)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= B * nsteps) return;
    float s = 0.0f;
    for (int j = lane; j < H; j += 32) s += h1[row * H + j] * w_out[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(NPD_FULL, s, o);
    if (lane == 0) {
        const int t = t0 + (int)(row / B);
        const int64_t b = row % B;
        const int is_loss = is_loss_t[t];
        const float o = s + b_out[0];
        out[(int64_t)t * B + b] = o;
        if (logits_out) logits_out[b * N + t] = o;
        float d = 0.0f, sq = 0.0f;
        if (is_loss) {
            const float e = o - gt[b * N + t];
            d = 2.0f * e * inv_count;           // d mean((o - target)^2) / d o
            sq = e * e;
        }
        dout[(int64_t)t * B + b] = d;
        lsq[(int64_t)t * B + b] = sq;           // summed once per iteration (90 k atomics on one word cost 0.1 ms here)
        if (write_fb && t + 1 < N)
            // student forcing feeds sign(decoded[:, t]) where decoded is the logit on loss (= info) positions and stays
            // +1 elsewhere (rnn_all.py:463-489); sign(0) = 0 one-hots like -1
            fb[(int64_t)(t + 1) * B + b] = is_loss ? (o > 0.0f ? 1.0f : (o < 0.0f ? -1.0f : 0.0f)) : 1.0f;
    }
}

// ---- backward gate kernel ------------------------------------------------------------------------------------------
// dh = dh_next[b,j] + extra, where extra = dout[b] * w_out[j] (layer 1) or dx1[b,j] (layer 0).
// Writes this step's dgi / dgh [B,3H] and overwrites dh_next with the direct path dh * z (the GEMM dgh W_hh is then
// accumulated on top with beta = 1).  One thread owns unit j of kRowsPerThread consecutive rows and adds their gate
// gradients into part[chunk][7][H] (stream-ordered launches, one owner per entry: no atomics): 0-2 = dgi (r, z, n) of rows
// whose feedback is +1 (all rows in layer 1), 3-5 = the same for feedback -1 / 0, 6 = dgh's n part.
template <bool LAYER0>
__global__ void __launch_bounds__(256) cell_bwd_kernel(const float *__restrict__ sv, size_t qs, const float *__restrict__ h_prev,
                                                       float *__restrict__ dh_next, const float *__restrict__ extra,
                                                       const float *__restrict__ dout, const float *__restrict__ w_out,
                                                       const float *__restrict__ fb, float *__restrict__ dgi,
                                                       float *__restrict__ dgh, float *__restrict__ part, int64_t B, int H)
{
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= H) return;
    const int64_t b0 = (int64_t)blockIdx.x * kRowsPerThread;
    float a[7] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    const float wo = LAYER0 ? 0.0f : w_out[j];
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
        const int64_t b = b0 + i;
        if (b >= B) break;
        const int64_t idx = b * H + j;
        const float r = sv[idx], z = sv[qs + idx], n = sv[2 * qs + idx], ghn = sv[3 * qs + idx];
        const float hp = h_prev[idx];
        float dh = dh_next[idx];
        dh += LAYER0 ? extra[idx] : dout[b] * wo;
        const float dn = dh * (1.0f - z);
        const float dz = dh * (hp - n);
        const float dn_pre = dn * (1.0f - n * n);
        const float dr_pre = dn_pre * ghn * r * (1.0f - r);
        const float dz_pre = dz * z * (1.0f - z);
        dh_next[idx] = dh * z;
        const int64_t o = b * 3 * H + j;
        dgi[o] = dr_pre;
        dgi[o + H] = dz_pre;
        dgi[o + 2 * H] = dn_pre;
        dgh[o] = dr_pre;
        dgh[o + H] = dz_pre;
        dgh[o + 2 * H] = dn_pre * r;
        if (LAYER0 && !(fb[b] > 0.0f)) { a[3] += dr_pre; a[4] += dz_pre; a[5] += dn_pre; }
        else { a[0] += dr_pre; a[1] += dz_pre; a[2] += dn_pre; }
        a[6] += dn_pre * r;
    }
    float *pp = part + (size_t)blockIdx.x * 7 * H + j;
#pragma unroll
    for (int k = 0; k < 7; ++k)
        if (LAYER0 || k < 3 || k == 6) pp[(size_t)k * H] += a[k];
}

// dst[c * dst_stride] = sum_b src[b * ld + c] (+ src2[b * ld + c]); one block per 32 columns, 8 warps striding over rows
__global__ void __launch_bounds__(256) colsum_kernel(const float *__restrict__ src, const float *__restrict__ src2, int64_t ld,
                                                     float *__restrict__ dst, int64_t dst_stride, int64_t B, int64_t C)
{
    __shared__ float part[8][33];
    const int64_t c = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
    const int w = threadIdx.x >> 5;
    float s = 0.0f;
    if (c < C)
        for (int64_t b = w; b < B; b += 8) {
            s += src[b * ld + c];
            if (src2) s += src2[b * ld + c];
        }
    part[w][threadIdx.x & 31] = s;
    __syncthreads();
    if (w == 0 && c < C) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x & 31];
        dst[c * dst_stride] = t;
    }
}

__global__ void __launch_bounds__(256) sum_kernel(const float *__restrict__ x, int64_t n, float *out, int square)
{
    float s = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        s += square ? x[i] * x[i] : x[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(NPD_FULL, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// clip_grad_norm_ (coef = clip / (norm + 1e-6), applied when < 1) + torch.optim.AdamW single-tensor update
__global__ void __launch_bounds__(256) adamw_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m,
                                                    float *__restrict__ v, int64_t n, const float *norm_sq, float clip,
                                                    float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float coef = 1.0f;
    if (clip > 0.0f) {
        const float c = clip / (sqrtf(*norm_sq) + 1e-6f);
        coef = c < 1.0f ? c : 1.0f;
    }
    const float gr = g[i] * coef;
    g[i] = gr;  // the clipped gradient stays readable (npd_gru_trainer_get)
    float w = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gr;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = w - (lr / bc1) * (mi / denom);
}

inline unsigned blocks_for(int64_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

}  // namespace

NPD_API size_t npd_gru_trainer_param_count(int N, int H)
{
    const size_t I = (size_t)N + 2, G = 3 * (size_t)H;
    return G * I + G * H + 2 * G + 2 * G * H + 2 * G + H + 1;
}

NPD_API int npd_gru_trainer_create(int N, int H, int64_t max_batch, const float *h_params, int tf32,
                                   npd_gru_trainer_t **out)
{
    NPD_REQUIRE(out && h_params, "npd_gru_trainer_create: null argument");
    NPD_REQUIRE(N >= 2 && N <= 4096 && H >= 8 && H <= 4096 && max_batch >= 1, "npd_gru_trainer_create: bad shape N=%d H=%d B=%lld",
                N, H, (long long)max_batch);
    // the all-steps GEMMs take N * B as a 32-bit dimension; the step-sum of dgi0 indexes B * 3H columns
    NPD_REQUIRE((int64_t)N * max_batch <= 0x7fffffffLL && max_batch * 3 * (int64_t)H <= 0x7fffffffLL,
                "npd_gru_trainer_create: N * max_batch = %lld (or max_batch * 3H) exceeds the GEMM index range",
                (long long)N * (long long)max_batch);
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    auto *t = new npd_gru_trainer();
    memset(t, 0, sizeof(*t));
    t->N = N; t->H = H; t->I = N + 2; t->max_batch = max_batch;
    const size_t I = t->I, G = 3 * (size_t)H, Hs = H;
    size_t o = 0;
    t->o_wih0 = o; o += G * I;
    t->o_whh0 = o; o += G * Hs;
    t->o_bih0 = o; o += G;
    t->o_bhh0 = o; o += G;
    t->o_wih1 = o; o += G * Hs;
    t->o_whh1 = o; o += G * Hs;
    t->o_bih1 = o; o += G;
    t->o_bhh1 = o; o += G;
    t->o_wout = o; o += Hs;
    t->o_bout = o; o += 1;
    t->n_params = o;
    t->beta1 = 0.9f; t->beta2 = 0.999f; t->eps = 1e-8f; t->weight_decay = 0.01f;  // torch.optim.AdamW defaults
    const size_t B = (size_t)max_batch, BH = B * Hs, BG = B * G, Ns = (size_t)N;
    const size_t chunks = (B + kRowsPerThread - 1) / kRowsPerThread;
    auto alloc = [&](float **p, size_t n) { return cudaMalloc((void **)p, n * sizeof(float)); };
#define TR_ALLOC(ptr, n) do { cudaError_t e__ = alloc(&(ptr), (n)); if (e__ != cudaSuccess) { \
        npd_set_error("npd_gru_trainer_create: cudaMalloc of %zu floats failed: %s", (size_t)(n), cudaGetErrorString(e__)); \
        npd_gru_trainer_destroy(t); return NPD_ENOMEM; } } while (0)
    TR_ALLOC(t->p, 4 * o);
    t->g = t->p + o; t->m = t->g + o; t->v = t->m + o;
    TR_ALLOC(t->saved, 2 * Ns * 5 * BH);
    TR_ALLOC(t->gy, BG); TR_ALLOC(t->gh, BG);
    TR_ALLOC(t->dgi, Ns * BG); TR_ALLOC(t->dgh, Ns * BG); TR_ALLOC(t->dx1, Ns * BH);
    TR_ALLOC(t->wpart, Ns * G * Hs);
    TR_ALLOC(t->part, 2 * chunks * 7 * Hs);
    TR_ALLOC(t->dh, BH); TR_ALLOC(t->zeros, BH);
    TR_ALLOC(t->fb, Ns * B); TR_ALLOC(t->out, Ns * B); TR_ALLOC(t->dout, Ns * B); TR_ALLOC(t->lsq, Ns * B);
    TR_ALLOC(t->scal, 4);
    TR_ALLOC(t->wcol, 2 * G);
#undef TR_ALLOC
    t->h_is_loss = (unsigned char *)malloc(Ns);
    if (!t->h_is_loss || cudaMalloc((void **)&t->is_loss, Ns) != cudaSuccess) {
        npd_set_error("npd_gru_trainer_create: cudaMalloc of %zu bytes failed", Ns);
        npd_gru_trainer_destroy(t);
        return NPD_ENOMEM;
    }
    NPD_CHECK_CUDA(cudaMemset(t->p, 0, 4 * o * sizeof(float)));
    NPD_CHECK_CUDA(cudaMemset(t->zeros, 0, BH * sizeof(float)));
    NPD_CHECK_CUDA(cudaMemcpy(t->p, h_params, o * sizeof(float), cudaMemcpyHostToDevice));
    NPD_CHECK_CUBLAS(cublasCreate(&t->blas));
    t->gemm_mode = tf32 < 0 ? 0 : tf32 > 3 ? 3 : tf32;
    NPD_CHECK_CUBLAS(cublasSetMathMode(t->blas, t->gemm_mode ? CUBLAS_DEFAULT_MATH : CUBLAS_PEDANTIC_MATH));
    NPD_CHECK_CUBLAS(cublasSetPointerMode(t->blas, CUBLAS_POINTER_MODE_HOST));
    if (t->gemm_mode == 1 && H % gru_tc::TU == 0) {
        t->tc_ok = gru_tc::encode_map(&t->tm_w[0], t->p + t->o_whh0, G, Hs) && gru_tc::encode_map(&t->tm_w[1], t->p + t->o_whh1, G, Hs);
        if (t->tc_ok &&
            (cudaFuncSetAttribute(gru_tc::gru_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gru_tc::SMEM_BYTES) != cudaSuccess ||
             cudaFuncSetAttribute(gru_tc::gru_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gru_tc::SMEM_BYTES) != cudaSuccess)) {
            (void)cudaGetLastError();
            t->tc_ok = 0;  // the library-GEMM path of the same mode
        }
    }
    *out = t;
    return NPD_OK;
}

NPD_API int npd_gru_trainer_destroy(npd_gru_trainer_t *t)
{
    if (!t) return NPD_OK;
    if (t->blas) cublasDestroy(t->blas);
    float *ptrs[] = {t->p, t->saved, t->gy, t->gh, t->dgi, t->dgh, t->dx1, t->wpart, t->part, t->dh, t->zeros, t->fb, t->out, t->dout, t->lsq,
                     t->scal, t->wcol};
    for (float *p : ptrs) if (p) cudaFree(p);
    if (t->is_loss) cudaFree(t->is_loss);
    free(t->h_is_loss);
    delete t;
    return NPD_OK;
}

// what: 0 = parameters, 1 = gradients of the last step (after clipping), 2 = exp_avg, 3 = exp_avg_sq
NPD_API int npd_gru_trainer_get(const npd_gru_trainer_t *t, int what, float *h_out)
{
    NPD_REQUIRE(t && h_out && what >= 0 && what <= 3, "npd_gru_trainer_get: bad argument");
    NPD_CHECK_CUDA(cudaDeviceSynchronize());
    NPD_CHECK_CUDA(cudaMemcpy(h_out, t->p + (size_t)what * t->n_params, t->n_params * sizeof(float), cudaMemcpyDeviceToHost));
    return NPD_OK;
}

NPD_API int npd_gru_trainer_set_params(npd_gru_trainer_t *t, const float *h_params, int reset_optimizer)
{
    NPD_REQUIRE(t && h_params, "npd_gru_trainer_set_params: null argument");
    NPD_CHECK_CUDA(cudaDeviceSynchronize());
    NPD_CHECK_CUDA(cudaMemcpy(t->p, h_params, t->n_params * sizeof(float), cudaMemcpyHostToDevice));
    if (reset_optimizer) {
        NPD_CHECK_CUDA(cudaMemset(t->m, 0, 2 * t->n_params * sizeof(float)));
        t->step = 0;
    }
    return NPD_OK;
}

NPD_API int npd_gru_train_step(npd_gru_trainer_t *t, const npd_code_t *loss_code, const float *y, const float *gt,
                               int teacher_forced, int64_t B, float lr, float clip, int apply_update,
                               float *loss_out, float *logits_out, void *stream)
{
    NPD_REQUIRE(t && loss_code && y && gt, "npd_gru_train_step: null argument");
    NPD_REQUIRE(loss_code->N == t->N, "npd_gru_train_step: code length %d != trainer N %d", loss_code->N, t->N);
    NPD_REQUIRE(B >= 1 && B <= t->max_batch, "npd_gru_train_step: batch %lld outside 1..%lld", (long long)B, (long long)t->max_batch);
    NPD_REQUIRE(loss_code->K >= 1, "npd_gru_train_step: empty loss set");
    cudaStream_t st = (cudaStream_t)stream;
    NPD_CHECK_CUBLAS(cublasSetStream(t->blas, st));
    const int N = t->N, H = t->H, I = t->I, G = 3 * H, mode = t->gemm_mode;
    const int64_t BH = B * H, BG = B * (int64_t)G;
    const int NB = (int)((int64_t)N * B);  // rows of the all-steps operands (checked at create)
    const float one = 1.0f, zero = 0.0f;
    std::vector<unsigned char> is_loss(N, 0);
    for (int k = 0; k < loss_code->K; ++k) is_loss[loss_code->h_info[k]] = 1;
    if (!t->is_loss_valid || memcmp(t->h_is_loss, is_loss.data(), (size_t)N) != 0) {  // once per curriculum stage
        memcpy(t->h_is_loss, is_loss.data(), (size_t)N);
        NPD_CHECK_CUDA(cudaMemcpyAsync(t->is_loss, t->h_is_loss, (size_t)N, cudaMemcpyHostToDevice, st));
        t->is_loss_valid = 1;
    }
    const float inv_count = 1.0f / ((float)B * (float)loss_code->K);  // nn.MSELoss(): mean over B x K
    float *P = t->p, *Gd = t->g;
    const size_t qs = (size_t)N * (size_t)BH;  // floats between two saved quantities of a layer
    auto sv = [&](int layer, int step) { return t->saved + (size_t)layer * 5 * qs + (size_t)step * (size_t)BH; };
    auto h_all = [&](int layer) { return t->saved + (size_t)layer * 5 * qs + 4 * qs; };  // [N B, H]
    auto h_of = [&](int layer, int step) -> const float * { return step < 0 ? t->zeros : h_all(layer) + (size_t)step * (size_t)BH; };
    const int64_t chunks = (B + kRowsPerThread - 1) / kRowsPerThread;
    float *part0 = t->part, *part1 = t->part + (size_t)chunks * 7 * H;

    NPD_CHECK_CUDA(cudaMemsetAsync(t->scal, 0, 4 * sizeof(float), st));
    NPD_CHECK_CUDA(cudaMemsetAsync(Gd, 0, t->n_params * sizeof(float), st));
    NPD_CHECK_CUDA(cudaMemsetAsync(t->part, 0, 2 * (size_t)chunks * 7 * H * sizeof(float), st));
    // ---- forward ----
    // row-major [rows, cols] arrays are cuBLAS column-major [cols, rows]: out[rows, G] = in[rows, K] . W[G, K]^T is
    // gemm(T, N, G, rows, K, W, in)
    if (t->tc_ok && t->tc_batch != B) {  // the saved state and the projections are laid out for the call's batch
        t->tc_ok = gru_tc::encode_map(&t->tm_s, t->saved, 10 * (uint64_t)N * (uint64_t)B, (uint64_t)H) &&
                   gru_tc::encode_map(&t->tm_gi[0], t->gy, (uint64_t)B, (uint64_t)G) &&
                   gru_tc::encode_map(&t->tm_gi[1], t->dgi, (uint64_t)N * (uint64_t)B, (uint64_t)G);
        t->tc_batch = B;
    }
    const bool tc = t->tc_ok != 0 && B % gru_tc::TM == 0 && 10 * (int64_t)N * B <= 0x7fffffffLL;
    const dim3 tc_grid((unsigned)(B / gru_tc::TM), (unsigned)(H / gru_tc::TU));
    auto tc_step = [&](int layer, int s) {
        gru_tc::FwdParams fp{};
        fp.b_ih = P + (layer == 0 ? t->o_bih0 : t->o_bih1);
        fp.b_hh = P + (layer == 0 ? t->o_bhh0 : t->o_bhh1);
        fp.wcol = layer == 0 ? t->wcol : nullptr;
        fp.fb = layer == 0 ? t->fb + (size_t)s * B : nullptr;
        fp.B = B; fp.H = H;
        fp.gi_row0 = layer == 0 ? 0 : (int)((int64_t)s * B);
        fp.hprev_row0 = s > 0 ? (int)((((int64_t)layer * 5 + 4) * N + (s - 1)) * B) : 0;
        for (int q = 0; q < 5; ++q) fp.out_row0[q] = (int)((((int64_t)layer * 5 + q) * N + s) * B);
        fp.zero_h = s == 0;
        if (layer == 0)
            gru_tc::gru_fwd_tc_kernel<true><<<tc_grid, gru_tc::THREADS, gru_tc::SMEM_BYTES, st>>>(fp, t->tm_s, t->tm_w[0], t->tm_gi[0]);
        else
            gru_tc::gru_fwd_tc_kernel<false><<<tc_grid, gru_tc::THREADS, gru_tc::SMEM_BYTES, st>>>(fp, t->tm_s, t->tm_w[1], t->tm_gi[1]);
    };
    auto layer0_step = [&](int s) -> cublasStatus_t {
        if (tc) { tc_step(0, s); return CUBLAS_STATUS_SUCCESS; }
        cublasStatus_t e = gemm32(t->blas, mode, CUBLAS_OP_T, CUBLAS_OP_N, G, (int)B, H, &one, P + t->o_whh0, H, h_of(0, s - 1), H, &zero, t->gh, G);
        cell_fwd_kernel<true><<<blocks_for(BH), 256, 0, st>>>(t->gy, t->gh, P + t->o_bih0, P + t->o_bhh0, t->wcol,
                                                              t->fb + (size_t)s * B, h_of(0, s - 1), sv(0, s), qs, B, H);
        return e;
    };
    auto layer1_step = [&](int s) -> cublasStatus_t {
        if (tc) { tc_step(1, s); return CUBLAS_STATUS_SUCCESS; }
        cublasStatus_t e = gemm32(t->blas, mode, CUBLAS_OP_T, CUBLAS_OP_N, G, (int)B, H, &one, P + t->o_whh1, H, h_of(1, s - 1), H, &zero, t->gh, G);
        cell_fwd_kernel<false><<<blocks_for(BH), 256, 0, st>>>(t->dgi + (size_t)s * BG, t->gh, P + t->o_bih1, P + t->o_bhh1, nullptr,
                                                               nullptr, h_of(1, s - 1), sv(1, s), qs, B, H);
        return e;
    };
    auto head = [&](int t0, int nsteps, int write_fb) {
        head_fwd_kernel<<<blocks_for((int64_t)nsteps * B * 32), 256, 0, st>>>(h_of(1, t0), P + t->o_wout, P + t->o_bout, gt, N, t0, nsteps,
                                                                             t->is_loss, write_fb, inv_count, t->out, t->dout, t->fb,
                                                                             logits_out, t->lsq, B, H);
    };
    // gy[B,3H] = y[B,N] . W_ih0[:, :N]^T
    NPD_CHECK_CUBLAS(gemm32(t->blas, mode, CUBLAS_OP_T, CUBLAS_OP_N, G, (int)B, N, &one, P + t->o_wih0, I, y, N, &zero, t->gy, G));
    onehot_cols_kernel<<<blocks_for(2 * G), 256, 0, st>>>(P + t->o_wih0, t->wcol, I, N, G);
    teacher_fb_kernel<<<blocks_for((int64_t)N * B), 256, 0, st>>>(gt, t->fb, B, N);  // student forcing overwrites steps >= 1
    if (teacher_forced) {
        for (int s = 0; s < N; ++s) NPD_CHECK_CUBLAS(layer0_step(s));
        // W_ih1 . h0 of all steps into the (still unused) dgi buffer
        NPD_CHECK_CUBLAS(gemm32(t->blas, mode, CUBLAS_OP_T, CUBLAS_OP_N, G, NB, H, &one, P + t->o_wih1, H, h_all(0), H, &zero, t->dgi, G));
        for (int s = 0; s < N; ++s) NPD_CHECK_CUBLAS(layer1_step(s));
        head(0, N, 0);
    } else {
        for (int s = 0; s < N; ++s) {
            NPD_CHECK_CUBLAS(layer0_step(s));
            NPD_CHECK_CUBLAS(gemm32(t->blas, mode, CUBLAS_OP_T, CUBLAS_OP_N, G, (int)B, H, &one, P + t->o_wih1, H, h_of(0, s), H, &zero,
                                    t->dgi + (size_t)s * BG, G));
            NPD_CHECK_CUBLAS(layer1_step(s));
            head(s, 1, 1);
        }
    }
    sum_kernel<<<64, 256, 0, st>>>(t->lsq, (int64_t)N * B, t->scal, 0);  // loss numerator
    NPD_CHECK_CUDA(cudaGetLastError());
    // ---- backward ----
    // in[rows, G] . W[G, K] -> out[rows, K] is gemm(N, N, K, rows, G, W, in); a[rows, K]^T-weighted sums
    // dW[G, K] = d[rows, G]^T a[rows, K] are gemm(N, T, K, G, rows, a, d)
    const dim3 bgrid((unsigned)chunks, (unsigned)((H + 255) / 256));
    // layer 1: dh = dh1 + dout_s (x) w_out
    NPD_CHECK_CUDA(cudaMemsetAsync(t->dh, 0, (size_t)BH * sizeof(float), st));
    for (int s = N - 1; s >= 0; --s) {
        cell_bwd_kernel<false><<<bgrid, 256, 0, st>>>(sv(1, s), qs, h_of(1, s - 1), t->dh, nullptr, t->dout + (size_t)s * B, P + t->o_wout,
                                                      nullptr, t->dgi + (size_t)s * BG, t->dgh + (size_t)s * BG, part1, B, H);
        if (s > 0)  // dh1 (for step s-1) = dh * z (already written) + dgh W_hh1
            NPD_CHECK_CUBLAS(gemm32(t->blas, mode, CUBLAS_OP_N, CUBLAS_OP_N, H, (int)B, G, &one, P + t->o_whh1, H, t->dgh + (size_t)s * BG, G,
                                    &one, t->dh, H));
    }
    // dW[G, H] = d[steps * B, G]^T a[steps * B, H].  As ONE GEMM the 3H x H output is 24 tiles for 148 SMs and the library does
    // not split K = steps * B on its own (2.9 ms each); one GEMM per step's rows into its own slice + a column sum over
    // the slices fills the machine (deterministic: fixed slice order).
    auto weight_grad = [&](const float *a, const float *d, int steps, float *dW) -> cublasStatus_t {
        cublasStatus_t e = gemm32_batched(t->blas, mode, CUBLAS_OP_N, CUBLAS_OP_T, H, G, (int)B, &one, a, H, (long long)BH, d, G,
                                          (long long)BG, &zero, t->wpart, H, (long long)G * H, steps);
        colsum_kernel<<<blocks_for((int64_t)G * H, 32), 256, 0, st>>>(t->wpart, nullptr, (int64_t)G * H, dW, 1, steps, (int64_t)G * H);
        return e;
    };
    // all steps at once: dx1 = dgi1 W_ih1 ; dW_ih1 = dgi1^T h0 ; dW_hh1 = dgh1[1:]^T h1[:-1] ; d w_out = h1^T dout
    NPD_CHECK_CUBLAS(gemm32(t->blas, mode, CUBLAS_OP_N, CUBLAS_OP_N, H, NB, G, &one, P + t->o_wih1, H, t->dgi, G, &zero, t->dx1, H));
    NPD_CHECK_CUBLAS(weight_grad(h_all(0), t->dgi, N, Gd + t->o_wih1));
    NPD_CHECK_CUBLAS(weight_grad(h_all(1), t->dgh + (size_t)BG, N - 1, Gd + t->o_whh1));
    NPD_CHECK_CUBLAS(cublasSgemv(t->blas, CUBLAS_OP_N, H, NB, &one, h_all(1), H, t->dout, 1, &zero, Gd + t->o_wout, 1));  // dout = 0 off the loss set
    // layer 0: dh = dh0 + dx1_s ; dgi / dgh are reused for this layer's gate gradients
    NPD_CHECK_CUDA(cudaMemsetAsync(t->dh, 0, (size_t)BH * sizeof(float), st));
    for (int s = N - 1; s >= 0; --s) {
        cell_bwd_kernel<true><<<bgrid, 256, 0, st>>>(sv(0, s), qs, h_of(0, s - 1), t->dh, t->dx1 + (size_t)s * BH, nullptr, nullptr,
                                                     t->fb + (size_t)s * B, t->dgi + (size_t)s * BG, t->dgh + (size_t)s * BG, part0, B, H);
        if (s > 0)
            NPD_CHECK_CUBLAS(gemm32(t->blas, mode, CUBLAS_OP_N, CUBLAS_OP_N, H, (int)B, G, &one, P + t->o_whh0, H, t->dgh + (size_t)s * BG, G,
                                    &one, t->dh, H));
    }
    NPD_CHECK_CUBLAS(weight_grad(h_all(0), t->dgh + (size_t)BG, N - 1, Gd + t->o_whh0));
    // d W_ih0[:, :N] = (sum_s dgi0_s)^T y  -- y is the same in every step; the sum over steps lands in gy
    colsum_kernel<<<blocks_for(BG, 32), 256, 0, st>>>(t->dgi, nullptr, BG, t->gy, 1, N, BG);
    NPD_CHECK_CUBLAS(gemm32(t->blas, mode, CUBLAS_OP_N, CUBLAS_OP_T, N, G, (int)B, &one, y, N, t->gy, G, &zero, Gd + t->o_wih0, I));
    // (rows of width N with leading dimension I: columns N, N+1 are set below)
    // bias gradients and the two one-hot columns of W_ih0 from the column-sum partials [chunks][7][H]
    const int64_t pld = 7 * (int64_t)H;
    const unsigned c3 = blocks_for(G, 32), c2 = blocks_for(2 * H, 32), c1 = blocks_for(H, 32);
    colsum_kernel<<<c3, 256, 0, st>>>(part0, nullptr, pld, Gd + t->o_wih0 + N + 1, I, chunks, G);           // feedback +1 -> column N+1
    colsum_kernel<<<c3, 256, 0, st>>>(part0 + 3 * H, nullptr, pld, Gd + t->o_wih0 + N, I, chunks, G);       // feedback -1 -> column N
    colsum_kernel<<<c3, 256, 0, st>>>(part0, part0 + 3 * H, pld, Gd + t->o_bih0, 1, chunks, G);
    colsum_kernel<<<c2, 256, 0, st>>>(part0, part0 + 3 * H, pld, Gd + t->o_bhh0, 1, chunks, 2 * H);        // r, z parts = dgi's
    colsum_kernel<<<c1, 256, 0, st>>>(part0 + 6 * H, nullptr, pld, Gd + t->o_bhh0 + 2 * H, 1, chunks, H);
    colsum_kernel<<<c3, 256, 0, st>>>(part1, nullptr, pld, Gd + t->o_bih1, 1, chunks, G);
    colsum_kernel<<<c2, 256, 0, st>>>(part1, nullptr, pld, Gd + t->o_bhh1, 1, chunks, 2 * H);
    colsum_kernel<<<c1, 256, 0, st>>>(part1 + 6 * H, nullptr, pld, Gd + t->o_bhh1 + 2 * H, 1, chunks, H);
    sum_kernel<<<64, 256, 0, st>>>(t->dout, (int64_t)N * B, Gd + t->o_bout, 0);
    NPD_CHECK_CUDA(cudaGetLastError());
    // ---- clip + AdamW ----
    sum_kernel<<<296, 256, 0, st>>>(Gd, (int64_t)t->n_params, t->scal + 1, 1);
    if (apply_update) {
        t->step += 1;
        const float bc1 = 1.0f - powf(t->beta1, (float)t->step);
        const float bc2s = sqrtf(1.0f - powf(t->beta2, (float)t->step));
        adamw_kernel<<<blocks_for((int64_t)t->n_params), 256, 0, st>>>(P, Gd, t->m, t->v, (int64_t)t->n_params, t->scal + 1, clip, lr,
                                                                       t->beta1, t->beta2, t->eps, t->weight_decay, bc1, bc2s);
    }
    NPD_CHECK_CUDA(cudaGetLastError());
    if (loss_out) {  // host float: mean squared error over the B x K loss entries (synchronises)
        float h[2];
        NPD_CHECK_CUDA(cudaMemcpyAsync(h, t->scal, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
        NPD_CHECK_CUDA(cudaStreamSynchronize(st));
        loss_out[0] = h[0] * inv_count;
        loss_out[1] = sqrtf(h[1]);  // total gradient norm before clipping
    }
    return NPD_OK;
}
