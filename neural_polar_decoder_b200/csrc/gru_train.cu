// gru_train.cu -- one training step of the CRISP GRU sequential decoder on the device.
//
// Replaces, for rnn_type GRU / decoding_type 'y_input' / onehot / 2 layers / Linear(H,1) head (the run_crisp.sh
// configuration), the body of the reference's training loop, rnn_all.py:1399-1437:
//     decoded = decoder.decode(net, True, y, gt, tfr)          teacher-forced (425-449) or student-forced (463-489)
//     loss    = MSELoss(decoded[:, info], msg_bits)             (1413)
//     loss.backward(); clip_grad_norm_(net.parameters(), clip)  (1431-1432)
//     optimizer.step()   [torch.optim.AdamW, default betas / eps / weight_decay]  (1346, 1435)
// All arithmetic is fp32 (the reference trains in fp32; parity = gradients and updated weights against the live
// reference at fp32 round-off).  Structure:
//   forward   N steps; per step 3 GEMMs [B,H] x [H,3H] (W_hh0, W_ih1, W_hh1; the y-part of W_ih0 is hoisted out of
//             the loop as in the decode kernel) + fused gate kernels that SAVE r, z, n, (W_hn h + b_hn)
//             and h for the backward pass + the head (dot product, loss, d loss / d logit, next feedback bit);
//   backward  N steps in reverse; per step and layer one fused gate-gradient kernel and two GEMMs
//             (d h_prev = dgh W_hh, d input = dgi W_ih), weight gradients accumulated as dg^T h GEMMs (beta = 1);
//             bias / one-hot-column gradients fall out of per-row accumulators reduced once at the end;
//   update    one fused kernel: global grad-norm clip coefficient + AdamW.
// The GEMMs are plain library GEMMs (cublasGemmEx on fp32 data; inner products in fp32, or on TF32 / bf16 / fp16 tensor
// cores with fp32 accumulation); everything else is this file.
#include <cublas_v2.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "npd_common.cuh"

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
    // activations (sized for max_batch)
    float *saved;       // [2 layers][N steps][5: r, z, n, ghn, h][B*H]
    float *gy;          // [B,3H]  y-part of the layer-0 input projection
    float *gi, *gh;     // [B,3H]  GEMM outputs of the current step
    float *dgi, *dgh;   // [B,3H]
    float *acc;         // [5][B,3H]: layer-0 dgi split by feedback sign (2), layer-0 dgh, layer-1 dgi, layer-1 dgh
    float *dh0, *dh1, *dx1, *zeros;  // [B,H]
    float *fb;          // [N][B] feedback entering step t (+-1)
    float *out, *dout;  // [N][B] logits, d loss / d logit
    float *scal;        // [4]: loss sum, grad norm^2, spare
};

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- forward gate kernel: one thread per (row b, unit j) -------------------------------------------------------
// LAYER0: gi = gy[b, g*H+j] + b_ih[g*H+j] + W_ih0[g*H+j, N + idx(fb[b])]   (one-hot feedback = column select)
// else  : gi = gi_buf[b, g*H+j] + b_ih[g*H+j]
template <bool LAYER0>
__global__ void __launch_bounds__(256) cell_fwd_kernel(const float *__restrict__ gi_src, const float *__restrict__ gh,
                                                       const float *__restrict__ b_ih, const float *__restrict__ b_hh,
                                                       const float *__restrict__ w_ih0, int I, int N,
                                                       const float *__restrict__ fb, const float *__restrict__ h_prev,
                                                       float *__restrict__ sv, int64_t B, int H)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int64_t b = idx / H;
    const int j = (int)(idx - b * H);
    const int64_t BH = B * H;
    float gir = gi_src[b * 3 * H + j] + b_ih[j];
    float giz = gi_src[b * 3 * H + H + j] + b_ih[H + j];
    float gin = gi_src[b * 3 * H + 2 * H + j] + b_ih[2 * H + j];
    if (LAYER0) {
        const int col = N + (fb[b] > 0.0f ? 1 : 0);  // get_onehot (rnn_all.py:258-260): +1 -> [0,1], -1 / 0 -> [1,0]
        gir += w_ih0[(size_t)j * I + col];
        giz += w_ih0[(size_t)(H + j) * I + col];
        gin += w_ih0[(size_t)(2 * H + j) * I + col];
    }
    const float ghr = gh[b * 3 * H + j] + b_hh[j];
    const float ghz = gh[b * 3 * H + H + j] + b_hh[H + j];
    const float ghn = gh[b * 3 * H + 2 * H + j] + b_hh[2 * H + j];
    const float r = sigmoidf_(gir + ghr);
    const float z = sigmoidf_(giz + ghz);
    const float n = tanhf(gin + r * ghn);
    const float hp = h_prev[idx];
    sv[idx] = r;
    sv[BH + idx] = z;
    sv[2 * BH + idx] = n;
    sv[3 * BH + idx] = ghn;
    sv[4 * BH + idx] = (1.0f - z) * n + z * hp;
}

// ---- head: logit = h1 . w_out + b_out (one warp per row), loss terms, next step's feedback ------------------------
__global__ void __launch_bounds__(256) head_fwd_kernel(const float *__restrict__ h1, const float *__restrict__ w_out,
                                                       const float *__restrict__ b_out, const float *__restrict__ gt,
                                                       int N, int t, int is_loss, int teacher, float inv_count,
                                                       float *__restrict__ out_t, float *__restrict__ dout_t,
                                                       float *__restrict__ fb_next, float *__restrict__ logits_out,
                                                       float *loss_sum, int64_t B, int H)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= B) return;
    float s = 0.0f;
    for (int j = lane; j < H; j += 32) s += h1[row * H + j] * w_out[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(NPD_FULL, s, o);
    if (lane == 0) {
        const float o = s + b_out[0];
        out_t[row] = o;
        if (logits_out) logits_out[row * N + t] = o;
        const float target = gt[row * N + t];
        float d = 0.0f;
        if (is_loss) {
            const float e = o - target;
            d = 2.0f * e * inv_count;           // d mean((o - target)^2) / d o
            atomicAdd(loss_sum, e * e);
        }
        dout_t[row] = d;
        if (fb_next) {
            // teacher forcing feeds gt[:, t] (rnn_all.py:447); student forcing feeds sign(decoded[:, t]) where decoded is
            // the logit on loss (= info) positions and stays +1 elsewhere (463-489); sign(0) = 0 one-hots like -1
            float f = teacher ? target : (is_loss ? (o > 0.0f ? 1.0f : (o < 0.0f ? -1.0f : 0.0f)) : 1.0f);
            fb_next[row] = f;
        }
    }
}

// ---- backward gate kernel ------------------------------------------------------------------------------------------
// dh = dh_next[b,j] + extra, where extra = dout[b] * w_out[j] (layer 1) or dx1[b,j] (layer 0).
// Writes dgi / dgh [B,3H], overwrites dh_next with the direct path dh * z (the GEMM dgh W_hh is then accumulated on top
// with beta = 1) and adds dgi / dgh into the per-row accumulators (bias and one-hot-column gradients).
template <bool LAYER0>
__global__ void __launch_bounds__(256) cell_bwd_kernel(const float *__restrict__ sv, const float *__restrict__ h_prev,
                                                       float *__restrict__ dh_next, const float *__restrict__ extra,
                                                       const float *__restrict__ dout, const float *__restrict__ w_out,
                                                       const float *__restrict__ fb, float *__restrict__ dgi,
                                                       float *__restrict__ dgh, float *__restrict__ acc_gi_pos,
                                                       float *__restrict__ acc_gi_neg, float *__restrict__ acc_gh,
                                                       int64_t B, int H)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int64_t b = idx / H;
    const int j = (int)(idx - b * H);
    const int64_t BH = B * H;
    const float r = sv[idx], z = sv[BH + idx], n = sv[2 * BH + idx], ghn = sv[3 * BH + idx];
    const float hp = h_prev[idx];
    float dh = dh_next[idx];
    dh += LAYER0 ? extra[idx] : dout[b] * w_out[j];
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
    float *ag = acc_gi_pos;
    if (LAYER0 && !(fb[b] > 0.0f)) ag = acc_gi_neg;
    ag[o] += dr_pre;
    ag[o + H] += dz_pre;
    ag[o + 2 * H] += dn_pre;
    acc_gh[o] += dr_pre;
    acc_gh[o + H] += dz_pre;
    acc_gh[o + 2 * H] += dn_pre * r;
}

// dst[c * dst_stride] (+)= sum_b src[b, c] (+ src2[b, c]); one block per 32 columns, 8 warps striding over rows
__global__ void __launch_bounds__(256) colsum_kernel(const float *__restrict__ src, const float *__restrict__ src2,
                                                     float *__restrict__ dst, int64_t dst_stride, int64_t B, int C, int accumulate)
{
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int w = threadIdx.x >> 5;
    float s = 0.0f;
    if (c < C)
        for (int64_t b = w; b < B; b += 8) {
            s += src[b * C + c];
            if (src2) s += src2[b * C + c];
        }
    part[w][threadIdx.x & 31] = s;
    __syncthreads();
    if (w == 0 && c < C) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x & 31];
        if (accumulate) dst[(int64_t)c * dst_stride] += t; else dst[(int64_t)c * dst_stride] = t;
    }
}

__global__ void __launch_bounds__(256) fill_kernel(float *__restrict__ a, float v, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

__global__ void __launch_bounds__(256) add_kernel(float *__restrict__ a, const float *__restrict__ b, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
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
    const size_t B = (size_t)max_batch, BH = B * Hs, BG = B * G;
    auto alloc = [&](float **p, size_t n) { return cudaMalloc((void **)p, n * sizeof(float)); };
#define TR_ALLOC(ptr, n) do { cudaError_t e__ = alloc(&(ptr), (n)); if (e__ != cudaSuccess) { \
        npd_set_error("npd_gru_trainer_create: cudaMalloc of %zu floats failed: %s", (size_t)(n), cudaGetErrorString(e__)); \
        npd_gru_trainer_destroy(t); return NPD_ENOMEM; } } while (0)
    TR_ALLOC(t->p, 4 * o);
    t->g = t->p + o; t->m = t->g + o; t->v = t->m + o;
    TR_ALLOC(t->saved, 2 * (size_t)N * 5 * BH);
    TR_ALLOC(t->gy, BG); TR_ALLOC(t->gi, BG); TR_ALLOC(t->gh, BG); TR_ALLOC(t->dgi, BG); TR_ALLOC(t->dgh, BG);
    TR_ALLOC(t->acc, 5 * BG);
    TR_ALLOC(t->dh0, BH); TR_ALLOC(t->dh1, BH); TR_ALLOC(t->dx1, BH); TR_ALLOC(t->zeros, BH);
    TR_ALLOC(t->fb, (size_t)N * B); TR_ALLOC(t->out, (size_t)N * B); TR_ALLOC(t->dout, (size_t)N * B);
    TR_ALLOC(t->scal, 4);
#undef TR_ALLOC
    NPD_CHECK_CUDA(cudaMemset(t->p, 0, 4 * o * sizeof(float)));
    NPD_CHECK_CUDA(cudaMemset(t->zeros, 0, BH * sizeof(float)));
    NPD_CHECK_CUDA(cudaMemcpy(t->p, h_params, o * sizeof(float), cudaMemcpyHostToDevice));
    NPD_CHECK_CUBLAS(cublasCreate(&t->blas));
    t->gemm_mode = tf32 < 0 ? 0 : tf32 > 3 ? 3 : tf32;
    NPD_CHECK_CUBLAS(cublasSetMathMode(t->blas, t->gemm_mode ? CUBLAS_DEFAULT_MATH : CUBLAS_PEDANTIC_MATH));
    NPD_CHECK_CUBLAS(cublasSetPointerMode(t->blas, CUBLAS_POINTER_MODE_HOST));
    *out = t;
    return NPD_OK;
}

NPD_API int npd_gru_trainer_destroy(npd_gru_trainer_t *t)
{
    if (!t) return NPD_OK;
    if (t->blas) cublasDestroy(t->blas);
    float *ptrs[] = {t->p, t->saved, t->gy, t->gi, t->gh, t->dgi, t->dgh, t->acc, t->dh0, t->dh1, t->dx1, t->zeros, t->fb,
                     t->out, t->dout, t->scal};
    for (float *p : ptrs) if (p) cudaFree(p);
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
    const int N = t->N, H = t->H, I = t->I, G = 3 * H;
    const int64_t BH = B * H, BG = B * (int64_t)G;
    const float one = 1.0f, zero = 0.0f;
    std::vector<char> is_loss(N, 0);
    for (int k = 0; k < loss_code->K; ++k) is_loss[loss_code->h_info[k]] = 1;
    const float inv_count = 1.0f / ((float)B * (float)loss_code->K);  // nn.MSELoss(): mean over B x K
    float *P = t->p, *Gd = t->g;
    auto sv = [&](int layer, int step) { return t->saved + ((size_t)layer * N + step) * 5 * (size_t)BH; };
    auto h_of = [&](int layer, int step) -> const float * { return step < 0 ? t->zeros : sv(layer, step) + 4 * (size_t)BH; };

    NPD_CHECK_CUDA(cudaMemsetAsync(t->scal, 0, 4 * sizeof(float), st));
    NPD_CHECK_CUDA(cudaMemsetAsync(Gd, 0, t->n_params * sizeof(float), st));
    NPD_CHECK_CUDA(cudaMemsetAsync(t->acc, 0, 5 * (size_t)BG * sizeof(float), st));
    // ---- forward ----
    // gy[B,3H] = y[B,N] . W_ih0[:, :N]^T   (row-major views as column-major: C^T = W . y^T)
    NPD_CHECK_CUBLAS(gemm32(t->blas, t->gemm_mode, CUBLAS_OP_T, CUBLAS_OP_N, G, (int)B, N, &one, P + t->o_wih0, I, y, N, &zero, t->gy, G));
    // feedback entering step 0 is +1 (rnn_all.py:444)
    fill_kernel<<<blocks_for(B), 256, 0, st>>>(t->fb, 1.0f, B);
    for (int s = 0; s < N; ++s) {
        // layer 0
        NPD_CHECK_CUBLAS(gemm32(t->blas, t->gemm_mode, CUBLAS_OP_T, CUBLAS_OP_N, G, (int)B, H, &one, P + t->o_whh0, H, h_of(0, s - 1), H,
                                     &zero, t->gh, G));
        cell_fwd_kernel<true><<<blocks_for(BH), 256, 0, st>>>(t->gy, t->gh, P + t->o_bih0, P + t->o_bhh0, P + t->o_wih0, I, N,
                                                              t->fb + (size_t)s * B, h_of(0, s - 1), sv(0, s), B, H);
        // layer 1
        NPD_CHECK_CUBLAS(gemm32(t->blas, t->gemm_mode, CUBLAS_OP_T, CUBLAS_OP_N, G, (int)B, H, &one, P + t->o_wih1, H, h_of(0, s), H, &zero,
                                     t->gi, G));
        NPD_CHECK_CUBLAS(gemm32(t->blas, t->gemm_mode, CUBLAS_OP_T, CUBLAS_OP_N, G, (int)B, H, &one, P + t->o_whh1, H, h_of(1, s - 1), H,
                                     &zero, t->gh, G));
        cell_fwd_kernel<false><<<blocks_for(BH), 256, 0, st>>>(t->gi, t->gh, P + t->o_bih1, P + t->o_bhh1, nullptr, I, N, nullptr,
                                                               h_of(1, s - 1), sv(1, s), B, H);
        head_fwd_kernel<<<blocks_for(B * 32), 256, 0, st>>>(h_of(1, s), P + t->o_wout, P + t->o_bout, gt, N, s, is_loss[s],
                                                            teacher_forced, inv_count, t->out + (size_t)s * B,
                                                            t->dout + (size_t)s * B, s + 1 < N ? t->fb + (size_t)(s + 1) * B : nullptr,
                                                            logits_out, t->scal, B, H);
    }
    NPD_CHECK_CUDA(cudaGetLastError());
    // ---- backward ----
    float *acc0p = t->acc, *acc0n = t->acc + BG, *acc0h = t->acc + 2 * BG, *acc1i = t->acc + 3 * BG, *acc1h = t->acc + 4 * BG;
    NPD_CHECK_CUDA(cudaMemsetAsync(t->dh0, 0, (size_t)BH * sizeof(float), st));
    NPD_CHECK_CUDA(cudaMemsetAsync(t->dh1, 0, (size_t)BH * sizeof(float), st));
    for (int s = N - 1; s >= 0; --s) {
        // layer 1: dh = dh1 + dout_s (x) w_out
        cell_bwd_kernel<false><<<blocks_for(BH), 256, 0, st>>>(sv(1, s), h_of(1, s - 1), t->dh1, nullptr, t->dout + (size_t)s * B,
                                                               P + t->o_wout, nullptr, t->dgi, t->dgh, acc1i, nullptr, acc1h, B, H);
        if (is_loss[s])  // d w_out += h1_s^T dout_s
            NPD_CHECK_CUBLAS(cublasSgemv(t->blas, CUBLAS_OP_N, H, (int)B, &one, h_of(1, s), H, t->dout + (size_t)s * B, 1, &one,
                                         Gd + t->o_wout, 1));
        // dW_hh1 += dgh^T h1_{s-1} ; dW_ih1 += dgi^T h0_s
        if (s > 0)
            NPD_CHECK_CUBLAS(gemm32(t->blas, t->gemm_mode, CUBLAS_OP_N, CUBLAS_OP_T, H, G, (int)B, &one, h_of(1, s - 1), H, t->dgh, G, &one,
                                         Gd + t->o_whh1, H));
        NPD_CHECK_CUBLAS(gemm32(t->blas, t->gemm_mode, CUBLAS_OP_N, CUBLAS_OP_T, H, G, (int)B, &one, h_of(0, s), H, t->dgi, G, &one,
                                     Gd + t->o_wih1, H));
        // dh1 (for step s-1) = dh * z (already written) + dgh W_hh1 ; dx1 = dgi W_ih1
        if (s > 0)
            NPD_CHECK_CUBLAS(gemm32(t->blas, t->gemm_mode, CUBLAS_OP_N, CUBLAS_OP_N, H, (int)B, G, &one, P + t->o_whh1, H, t->dgh, G, &one, t->dh1, H));
        NPD_CHECK_CUBLAS(gemm32(t->blas, t->gemm_mode, CUBLAS_OP_N, CUBLAS_OP_N, H, (int)B, G, &one, P + t->o_wih1, H, t->dgi, G, &zero, t->dx1, H));
        // layer 0: dh = dh0 + dx1
        cell_bwd_kernel<true><<<blocks_for(BH), 256, 0, st>>>(sv(0, s), h_of(0, s - 1), t->dh0, t->dx1, nullptr, nullptr,
                                                              t->fb + (size_t)s * B, t->dgi, t->dgh, acc0p, acc0n, acc0h, B, H);
        if (s > 0) {
            NPD_CHECK_CUBLAS(gemm32(t->blas, t->gemm_mode, CUBLAS_OP_N, CUBLAS_OP_T, H, G, (int)B, &one, h_of(0, s - 1), H, t->dgh, G, &one,
                                         Gd + t->o_whh0, H));
            NPD_CHECK_CUBLAS(gemm32(t->blas, t->gemm_mode, CUBLAS_OP_N, CUBLAS_OP_N, H, (int)B, G, &one, P + t->o_whh0, H, t->dgh, G, &one, t->dh0, H));
        }
    }
    // bias gradients and the two one-hot columns of W_ih0 from the per-row accumulators
    const unsigned cg = (unsigned)((G + 31) / 32);
    colsum_kernel<<<cg, 256, 0, st>>>(acc0p, nullptr, Gd + t->o_wih0 + N + 1, I, B, G, 0);  // feedback +1 -> column N+1
    colsum_kernel<<<cg, 256, 0, st>>>(acc0n, nullptr, Gd + t->o_wih0 + N, I, B, G, 0);      // feedback -1 -> column N
    colsum_kernel<<<cg, 256, 0, st>>>(acc0p, acc0n, Gd + t->o_bih0, 1, B, G, 0);
    colsum_kernel<<<cg, 256, 0, st>>>(acc0h, nullptr, Gd + t->o_bhh0, 1, B, G, 0);
    colsum_kernel<<<cg, 256, 0, st>>>(acc1i, nullptr, Gd + t->o_bih1, 1, B, G, 0);
    colsum_kernel<<<cg, 256, 0, st>>>(acc1h, nullptr, Gd + t->o_bhh1, 1, B, G, 0);
    // d W_ih0[:, :N] = (sum_s dgi0_s)^T y  -- y is the same in every step
    add_kernel<<<blocks_for(BG), 256, 0, st>>>(acc0p, acc0n, BG);
    NPD_CHECK_CUBLAS(gemm32(t->blas, t->gemm_mode, CUBLAS_OP_N, CUBLAS_OP_T, N, G, (int)B, &one, y, N, acc0p, G, &zero, Gd + t->o_wih0, I));
    // the GEMM above wrote rows of width N with leading dimension I: columns N, N+1 were untouched (set by the colsums)
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
