/*
 * npd.h -- C ABI of libnpd.so, the B200 (sm_100a) Monte-Carlo polar/PAC decode path.
 *
 * The reference (hebbarashwin/neural_polar_decoder) has no FFI layer: its boundary is the Python
 * method surface of PolarCode / PAC / RNN_decoder / convNet.  The Python drop-in in
 * neural_polar_decoder_b200/ keeps those signatures and forwards every hot-path call to the entry
 * points below through ctypes (INTEGRATION.md shows the binding).  Each entry point cites the
 * reference interface it replaces.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  Every tensor argument is a caller-owned DEVICE
 *    pointer to a contiguous row-major float32 array unless its name starts with `h_` (host).
 *  - BPSK convention of the reference: bit 0 <-> +1.0f, bit 1 <-> -1.0f (polar.py:130-132).
 *  - Kernels are enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *    on the CURRENT device; the call returns without synchronising.  Scratch is passed in by the
 *    caller where an entry point has a `workspace` argument; the one exception is npd_sc_decode at
 *    N >= 2048, whose level scratch is a stream-ordered allocation (cudaMallocFromPoolAsync /
 *    cudaFreeAsync on `stream`) from a library-owned pool that keeps its memory across calls.
 *    Handles own small device-side tables / repacked weights.  The library never reads the
 *    environment (experiment switches exist only in -DNPD_DEBUG_KNOBS builds).
 *  - Return value: 0 on success, a negative NPD_E* code otherwise; npd_last_error() gives a
 *    thread-local message.  There is no CPU fallback anywhere: without a CUDA device every compute
 *    entry point returns NPD_ECUDA.
 */
#ifndef NPD_H_
#define NPD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPD_VERSION 100 /* 0.1.0 */

#define NPD_OK 0
#define NPD_EINVAL (-1)       /* bad argument */
#define NPD_ECUDA (-2)        /* CUDA runtime error / no device */
#define NPD_EUNSUPPORTED (-3) /* valid request outside the implemented envelope */
#define NPD_ENOMEM (-4)

typedef struct npd_code npd_code_t; /* code object: (N, K, info set, frozen mask, PAC taps) */
typedef struct npd_gru npd_gru_t;   /* CRISP GRU decoder weights, repacked for the kernel */
typedef struct npd_conv npd_conv_t; /* convNet weights, repacked for the kernel */

int npd_version(void);
const char *npd_last_error(void);

/* Number of SMs / name of the current device (diagnostics; NPD_ECUDA if none). */
int npd_device_info(int *sm_count, int *cc_major, int *cc_minor, char *name, int name_len);

/* ---- code objects ------------------------------------------------------------------------------
 * Replaces PolarCode.__init__ (polar.py:66-117: info/frozen sets, infty) and PAC.__init__
 * (pac_code.py:97-119: g -> tap array, RM info set).  `h_info` = sorted info positions (host int32),
 * n = log2 N (1..12).  `pac_g` = 0 for a plain polar code, else the generator polynomial as the
 * reference passes it (e.g. 53 -> taps 1,1,0,1,0,1, MSB first).  `infty` = frozen prior
 * (polar.py:81, 471-472); ignored by the PAC decoder, which has no priors (pac_code.py:265-345). */
int npd_code_create(int n, int K, const int32_t *h_info, float infty, uint32_t pac_g,
                    npd_code_t **out);
int npd_code_destroy(npd_code_t *code);

/* ---- encoder + channel -------------------------------------------------------------------------
 * npd_polar_encode: PolarCode.encode_plotkin(message) (polar.py:128-148, scaling=None) and, for a
 * PAC code object, PAC.pac_encode(msg) (pac_code.py:220-224: rate profile -> rate-1 convolutional
 * pre-coder -> Plotkin transform).  msg[B,K] in {+1,-1} -> x[B,N] in {+1,-1}. */
int npd_polar_encode(const npd_code_t *code, const float *msg, float *x, int64_t B, void *stream);

/* npd_awgn: PolarCode.channel / PAC.channel (polar.py:201-207, pac_code.py:226-231):
 * y = x + sigma * z, z ~ N(0,1).  The reference draws z with torch.randn on the CPU generator; here
 * z is Philox4x32-10 + Box-Muller keyed by (seed, point) and counted by the GLOBAL codeword index
 * cw_offset + row, so a sweep gives identical noise for any GPU count / batch split. */
int npd_awgn(const float *x, float *y, int64_t B, int N, float sigma, uint64_t seed,
             uint32_t point, uint64_t cw_offset, void *stream);

/* npd_gen_encode_awgn: fused message generation (rnn_all.py:1771 / polar.py:1259: iid +-1) +
 * encode + channel for one SNR point.  Writes y[B,N]; msg[B,K] and x[B,N] are optional outputs
 * (NULL to skip).  Messages depend only on (seed, global codeword index), noise also on `point`. */
int npd_gen_encode_awgn(const npd_code_t *code, float *msg, float *x, float *y, int64_t B,
                        float sigma, uint64_t seed, uint32_t point, uint64_t cw_offset,
                        void *stream);

/* ---- successive-cancellation decoders ----------------------------------------------------------
 * npd_sc_decode: PolarCode.sc_decode_new(y, snr, use_gt) (polar.py:465-484 with 361-463 and
 * utils.py:272-275).  llr_scale = fp32(2/sigma^2) (torch rounds the Python scalar to fp32 before
 * the multiply).  Bit-exact fp32 min-sum SC: f = min(|a|,|b|)*sign(a)*sign(b), g = u*a + b with
 * u in {-1,0,+1}, leaf = L + infty*frozen, u = sign(leaf) (sign(0) = 0 and propagates).
 *   use_gt   [B,N] or NULL : genie decisions in {-1,0,+1} used for ALL positions (polar.py:480-481)
 *   leaf_llr [B,N] or NULL : llr_array[:,0,:] (includes the frozen prior)
 *   decoded  [B,K]         : u_hat[:, info_positions] in {-1,0,+1}
 * For a PAC code object use npd_pac_sc_decode. */
int npd_sc_decode(const npd_code_t *code, const float *y, float llr_scale, const float *use_gt,
                  float *leaf_llr, float *decoded, int64_t B, void *stream);

/* npd_sc_round_codewords: the decisions-only decoder for N >= 256 is a persistent kernel -- every resident warp decodes
 * groups of 8 codewords in rounds -- so a call (or a sweep chunk, npd_mc_sc_sweep) whose B is a multiple of this number
 * leaves no partly filled last round (B = 131072 at N = 1024 is 9.2 rounds and costs 10).  0 for codes decoded by the
 * dynamically scheduled kernels (N < 256, PAC).  The loops of polar.py:1258-1291 choose their own batch; sweep.py uses
 * this to size the chunks of the fused sweep. */
int64_t npd_sc_round_codewords(const npd_code_t *code);

/* npd_pac_sc_decode: PAC.pac_sc_decode(y, snr, use_gt_codeword) (pac_code.py:534-573): min-sum SC
 * without priors + convolutional-state tracking.  Outputs leaf_llr[B,N] (optional), v_hat[B,K]
 * (= v_hat[:, B-set], 0 where a tie left v undecided), u_hat[B,N] (optional). */
int npd_pac_sc_decode(const npd_code_t *code, const float *y, float llr_scale,
                      const float *use_gt_codeword, float *leaf_llr, float *v_hat, float *u_hat,
                      int64_t B, void *stream);

/* npd_scl_decode: PolarCode.scl_decode(y, snr, L, use_CRC=False) (polar.py:793-876 with pruneLists 777-791):
 * SC-list decoding with L paths -- min-sum LLR recursion WITHOUT the frozen prior, path metric += |L| for every
 * decision against sign(L), list pruned to the L smallest metrics (ascending list order), final pick = the
 * surviving path whose re-encoded codeword is closest to y.  1 <= list_size <= 32.
 *   leaf_llr [B,N] or NULL : llr_array[:,0,:] of the chosen path (frozen leaves include + infty)
 *   decoded  [B,K]         : its u_hat[:, info_positions]
 * The CRC-aided variant (use_CRC=True) needs the reference's per-codeword Python CRC and is not implemented. */
int npd_scl_decode(const npd_code_t *code, const float *y, float llr_scale, int list_size,
                   float *leaf_llr, float *decoded, int64_t B, void *stream);

/* ---- error counting ----------------------------------------------------------------------------
 * npd_count_errors: numerators of errors_ber (utils.py:17-25) and errors_bler (utils.py:37-51):
 * counts[0] += #{round(a) != round(b)}, counts[1] += #rows with any mismatch, over a[B,K], b[B,K].
 * counts is a device uint64[2] that the call ACCUMULATES into (zero it first). */
/* npd_count_errors_info: the same numerators for a decoder that returns full-length rows: msg[B,K] against
 * decoded_full[:, info_positions] with decoded_full [B,N] -- `errors_ber(msg_bits, decoded_bits[:, polar.info_positions]
 * .sign())` (rnn_all.py:875-879, run_models.py:338-346) without materialising the gathered tensor.  take_sign != 0
 * applies the `.sign()` to decoded_full first (logits in, e.g. convNet's), 0 compares the values as errors_ber does. */
int npd_count_errors_info(const npd_code_t *code, const float *msg, const float *decoded_full, int64_t B,
                          int take_sign, uint64_t *counts, void *stream);
int npd_count_errors(const float *a, const float *b, int64_t B, int K, uint64_t *counts,
                     void *stream);

/* ---- fused Monte-Carlo SC sweep ----------------------------------------------------------------
 * The inner loop of polar.py:1258-1291 / rnn_all.py:843-856 for one SNR point without any host
 * round trip: generate B messages (global indices cw_offset..cw_offset+B), encode, add noise,
 * SC-decode, count.  counts (device uint64[3]) accumulates bit errors, block errors, frames.
 * `workspace` is caller-owned device scratch of at least npd_mc_sc_workspace_bytes(code, chunk)
 * bytes (two chunks: one is generated on a library-owned second stream while the other is decoded and
 * counted on `stream`); the batch is processed in chunks of `chunk` codewords. */
size_t npd_mc_sc_workspace_bytes(const npd_code_t *code, int64_t chunk);
int npd_mc_sc_sweep(const npd_code_t *code, int64_t B, int64_t chunk, float sigma, float llr_scale,
                    uint64_t seed, uint32_t point, uint64_t cw_offset, void *workspace,
                    size_t workspace_bytes, uint64_t *counts, void *stream);

/* ---- CRISP GRU sequential decoder --------------------------------------------------------------
 * npd_gru_create: repack the parameters of RNN_Model('GRU', N+2, H, 1, L=2, ...) (rnn_all.py:294-343;
 * state_dict keys rnn.weight_ih_l{0,1}, rnn.weight_hh_l{0,1}, rnn.bias_*, linear.weight/bias) from
 * host fp32 into the kernel's fp16 tile streams (pre-swizzled 16 KB weight tiles in consumption order, one stream per
 * kernel variant).  h_* are host pointers, PyTorch layouts:
 *   w_ih0[3H, N+2], w_hh0[3H,H], b_ih0[3H], b_hh0[3H], w_ih1[3H,H], w_hh1[3H,H], b_ih1[3H],
 *   b_hh1[3H], w_out[H], b_out[1].  Gate order r,z,n. */
int npd_gru_create(int N, int H, const float *h_w_ih0, const float *h_w_hh0, const float *h_b_ih0,
                   const float *h_b_hh0, const float *h_w_ih1, const float *h_w_hh1,
                   const float *h_b_ih1, const float *h_b_hh1, const float *h_w_out,
                   const float *h_b_out, npd_gru_t **out);
int npd_gru_destroy(npd_gru_t *gru);

/* npd_gru_set_head_mlp: replace the Linear(H,1) head by the MLP head RNN_Model builds for out_linear_depth > 1
 * (rnn_all.py:335-343): Linear(H,Yh) SELU [Linear(Yh,Yh) SELU] x (depth-2) Linear(Yh,1), Yh = y_hidden_size.
 * h_params: host fp32 blob in state_dict order of `linear.*`: weight [Yh,H], bias [Yh], then (depth-2) x
 * (weight [Yh,Yh], bias [Yh]), then weight [1,Yh], bias [1].  Create the decoder with w_out = zeros first.
 * A decoder with an MLP head needs npd_gru_workspace_bytes(gru, B) > 0 bytes of workspace per decode call.
 * Envelope: 2 <= depth <= 8, Yh a multiple of 16 in [16, 1024]; NPD_EUNSUPPORTED otherwise. */
int npd_gru_set_head_mlp(npd_gru_t *gru, int depth, int y_hidden_size, const float *h_params);

/* Options of a GRU decoder handle.  NPD_GRU_OPT_RESIDUAL_STATE (default 1): the CTA-pair kernel (H = 256 / 512) keeps,
 * next to the fp16 recurrent state the tensor cores read, the rounding residual of that state (one signed byte) in an L2-resident
 * buffer (npd_gru_workspace_bytes, or the library's pool) and updates h' = h - (1 - z)(h - n) from hi + lo.  On the
 * reference-trained Polar(64,22) checkpoint the forced-feedback logit error drops from 9e-3 (2.6 x the tolerance
 * 1e-2 |ref| + 2e-3) to 3e-3 (at the fp16-operand floor) for +10 % decode time; 0 = the faster fp16-only state. */
#define NPD_GRU_OPT_RESIDUAL_STATE 1
int npd_gru_set_option(npd_gru_t *gru, int option, int value);

/* npd_gru_decode: RNN_decoder.decode(net, False, y) test branch, decoding_type 'y_input', onehot
 * (rnn_all.py:514-521, 532-547) with RNN_Model.forward (387-398): N autoregressive steps, hidden
 * state zero-initialised, input [y | onehot(prev decision)], decision = sign(logit) on info
 * positions, +1 elsewhere.
 *   y       [B,N]          raw channel output (not LLR; rnn_all.py:533-534)
 *   forced  [B,N] or NULL  when given, step i feeds back forced[:, i-1] instead of the decoder's own
 *                          decision: the teacher-forced pass of rnn_all.py:425-461 (logits = its output)
 *   genie   [B,N] or NULL  when given, `decoded` starts as this tensor instead of ones (gt.clone(),
 *                          rnn_all.py:519-522): positions outside the loss set keep and feed back their
 *                          genie value, loss positions the decoder's own decision
 *   logits  [B,N] or NULL  head output of every step
 *   decoded [B,N]          +-1/0 decisions on `loss positions` (the code's info set), +1 elsewhere
 * workspace: device scratch of npd_gru_workspace_bytes(gru, B) bytes. */
size_t npd_gru_workspace_bytes(const npd_gru_t *gru, int64_t B);
int npd_gru_decode(const npd_gru_t *gru, const npd_code_t *code, const float *y,
                   const float *forced, const float *genie, float *logits, float *decoded, int64_t B,
                   void *workspace, size_t workspace_bytes, void *stream);

/* npd_gru_decode_h0: the same loop started from a caller-supplied hidden state -- decoding_type 'y_h0'
 * (rnn_all.py:523-531): hidden = net.get_h0(y) (RNN_Model.get_h0, 362-375: the y-MLP reshaped to [L,B,H]) and the
 * step input is only the previous decision, i.e. a decoder created with all-zero y columns in w_ih0.
 *   h0      [2,B,H] or NULL  initial hidden state of both layers, fp32 (NULL = zeros = npd_gru_decode)
 * every other argument as npd_gru_decode. */
int npd_gru_decode_h0(const npd_gru_t *gru, const npd_code_t *code, const float *y, const float *h0,
                      const float *forced, const float *genie, float *logits, float *decoded, int64_t B,
                      void *workspace, size_t workspace_bytes, void *stream);

/* npd_mc_gru_sweep: the inner loop of polar_RNN_full_test (rnn_all.py:843-879) for one SNR point and the GRU decoder
 * without any host round trip: messages -> encode -> AWGN -> npd_gru_decode -> error counters, `chunk` codewords at a
 * time (a multiple of 148 x 64 fills whole waves of CTA pairs).  Philox streams as npd_gen_encode_awgn (seed, point,
 * global frame index cw_offset + row), so counts do not depend on the chunking or the GPU count.
 *   code      : the polar/PAC code object (encoder + info positions to score)
 *   loss_code : positions where the decoder decides (NULL = `code`: RNN_decoder.decode's default loss_inds)
 *   counts    : device uint64[3], ACCUMULATED: bit errors, block errors, frames
 *   workspace : >= npd_mc_gru_workspace_bytes(gru, code, chunk) bytes of device memory */
size_t npd_mc_gru_workspace_bytes(const npd_gru_t *gru, const npd_code_t *code, int64_t chunk);
int npd_mc_gru_sweep(const npd_gru_t *gru, const npd_code_t *code, const npd_code_t *loss_code, int64_t B,
                     int64_t chunk, float sigma, uint64_t seed, uint32_t point, uint64_t cw_offset,
                     void *workspace, size_t workspace_bytes, uint64_t *counts, void *stream);

/* ---- CRISP GRU training step --------------------------------------------------------
 * One iteration of the reference's training loop for rnn_type GRU / 'y_input' / onehot / 2 layers / Linear(H,1) head
 * (rnn_all.py:1399-1437): decoder.decode(net, True, y, gt, tfr) teacher-forced (425-449) or student-forced (463-489),
 * MSELoss on the loss positions (1413), backward, clip_grad_norm_(clip) (1432), torch.optim.AdamW step (1346, 1435;
 * betas 0.9 / 0.999, eps 1e-8, weight_decay 0.01).  fp32 data throughout.  The trainer owns parameters, gradients, Adam
 * moments and the saved activations for up to max_batch codewords.
 *   h_params : host fp32 blob in state_dict order: rnn.weight_ih_l0 [3H,N+2], rnn.weight_hh_l0 [3H,H], rnn.bias_ih_l0,
 *              rnn.bias_hh_l0 [3H], rnn.weight_ih_l1 [3H,H], rnn.weight_hh_l1 [3H,H], rnn.bias_ih_l1, rnn.bias_hh_l1,
 *              linear.weight [H], linear.bias [1]  (npd_gru_trainer_param_count(N, H) floats)
 *   tf32     : GEMM arithmetic on the fp32 data: 0 = fp32 (parity mode), 1 = TF32, 2 = bf16, 3 = fp16 tensor cores (fp32
 *              accumulation; master weights, saved activations, gate math and the optimizer stay fp32).  In mode 1, with
 *              H a multiple of 128 and a batch that is a multiple of 128, the forward pass's recurrent GEMM and gate math of
 *              a layer-step run as one tcgen05 kernel (csrc/gru_train_tc.cuh; sigma / tanh through ex2.approx + rcp.approx,
 *              relative error 1e-7); every other shape or mode uses library GEMMs + the exact gate kernels
 *   apply_update = 0 leaves the gradient UNCLIPPED (the clip coefficient is applied inside the update kernel) */
typedef struct npd_gru_trainer npd_gru_trainer_t;
size_t npd_gru_trainer_param_count(int N, int H);
int npd_gru_trainer_create(int N, int H, int64_t max_batch, const float *h_params, int tf32,
                           npd_gru_trainer_t **out);
int npd_gru_trainer_destroy(npd_gru_trainer_t *trainer);
/* what: 0 = parameters, 1 = gradients of the last step (after clipping), 2 = exp_avg, 3 = exp_avg_sq (synchronises) */
int npd_gru_trainer_get(const npd_gru_trainer_t *trainer, int what, float *h_out);
int npd_gru_trainer_set_params(npd_gru_trainer_t *trainer, const float *h_params, int reset_optimizer);
/* npd_gru_train_step:
 *   loss_code      : positions that enter the loss and, when student-forced, feed back sign(logit) (the info set)
 *   y, gt [B,N]    : channel output; genie tensor (+1 on frozen positions, the message on info positions, 1401-1402)
 *   teacher_forced : 1 = feedback gt[:, i-1] (tfr draw succeeded), 0 = the decoder's own detached decisions
 *   lr, clip       : this step's learning rate (the caller runs the scheduler) and clip_grad_norm_ max norm (<= 0: off)
 *   apply_update   : 0 = gradients only (readable through npd_gru_trainer_get(.., 1, ..))
 *   h_loss_and_norm: NULL or host float[2] <- MSE loss, total gradient norm before clipping (synchronises `stream`)
 *   logits [B,N]   : NULL or device output of every step's logit (decoded_vhat of the teacher-forced pass) */
int npd_gru_train_step(npd_gru_trainer_t *trainer, const npd_code_t *loss_code, const float *y, const float *gt,
                       int teacher_forced, int64_t B, float lr, float clip, int apply_update,
                       float *h_loss_and_norm, float *logits, void *stream);

/* ---- convNet one-shot decoder ------------------------------------------------------------------
 * npd_conv_create / npd_conv_forward: convNet.forward (models.py:742-767; layers 701-740) with
 * embed_dim = 2*C (C = 64), max_len = N: ten dilated k=7 Conv1d + GELU with three residual adds,
 * flatten, Linear(2C*N,4N) GELU Linear(4N,N) GELU Linear(N,N), LayerNorm(N, eps=1e-6).
 * h_params: host fp32 blob in state_dict order with every bias present (zeros under dont_use_bias):
 * layers1.0.weight, layers1.0.bias, layers1.2.*, layers2.0.*, ..., layers5.2.*, layersFin.0.*, layersFin.2.*,
 * layersFin.4.*, layer_norm.weight, layer_norm.bias (neural_polar_decoder_b200/models.py packs it).
 * logits[B,N] out (the LayerNorm output; bits = sign(logits) and probs = sigmoid(logits) are left to the
 * caller); in4[B,C,N] or NULL = forward()'s 5th return value (input4, models.py:752, 765).
 * workspace: device scratch of npd_conv_workspace_bytes(conv, B) bytes (fp16 rows of the flattened last
 * conv activation, at most 1 GiB; larger batches are processed in chunks). */
int npd_conv_create(int N, int embed_dim, const float *h_params, size_t n_params,
                    npd_conv_t **out);
int npd_conv_destroy(npd_conv_t *conv);
size_t npd_conv_workspace_bytes(const npd_conv_t *conv, int64_t B);
int npd_conv_forward(const npd_conv_t *conv, const float *y, float *logits, float *in4, int64_t B,
                     void *workspace, size_t workspace_bytes, void *stream);
/* npd_conv_decode: convNet.decode (models.py:769-772): bits[B,N] = sign(logits) (torch.sign: -1, 0, +1), same
 * kernels, the sign taken in the last kernel's epilogue. */
int npd_conv_decode(const npd_conv_t *conv, const float *y, float *bits, int64_t B, void *workspace,
                    size_t workspace_bytes, void *stream);

/* ---- host-buffer entry points -------------------------------------------------------------------
 * The same decoders for callers whose tensors live in HOST memory -- the reference's evaluation loops
 * build the test set and draw the noise on the CPU (rnn_all.py:1771-1773, polar.py:204) and read the
 * decisions back there (utils.py:41-45).  Every h_* argument is a HOST pointer with the shape of its
 * device counterpart above (pinned memory for full PCIe speed; pageable memory works).  The batch is
 * cut into chunks and chunk i+1's host->device copy, chunk i's kernels and chunk i-1's device->host
 * copy overlap on three private streams of the current device; staging memory belongs to the library
 * (grown on demand, reused across calls).  SYNCHRONOUS: outputs are complete on return.  Calls on the
 * same device are serialised. */
int npd_sc_decode_host(const npd_code_t *code, const float *h_y, float llr_scale,
                       const float *h_use_gt, float *h_leaf_llr, float *h_decoded, int64_t B);
/* Rows per pipeline chunk for every npd_*_host call of this process (0 = sized automatically from the row width;
 * rounded up to the kernel's tile).  A tuning / test hook: results never depend on it. */
int npd_host_set_chunk(int64_t rows);
int npd_scl_decode_host(const npd_code_t *code, const float *h_y, float llr_scale, int list_size,
                        float *h_leaf_llr, float *h_decoded, int64_t B);
int npd_pac_sc_decode_host(const npd_code_t *code, const float *h_y, float llr_scale,
                           const float *h_use_gt_codeword, float *h_leaf_llr, float *h_v_hat,
                           float *h_u_hat, int64_t B);
int npd_gru_decode_host(const npd_gru_t *gru, const npd_code_t *code, const float *h_y,
                        const float *h_forced, const float *h_genie, float *h_logits,
                        float *h_decoded, int64_t B);
int npd_conv_forward_host(const npd_conv_t *conv, const float *h_y, float *h_logits, float *h_in4,
                          int64_t B);
int npd_conv_decode_host(const npd_conv_t *conv, const float *h_y, float *h_bits, int64_t B);

#ifdef __cplusplus
}
#endif
#endif /* NPD_H_ */
