/*
 * npd_oracle.c -- CPU restatement of the CRISP reference's Monte-Carlo decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (neural_polar_decoder_b200/) may link,
 * load or call this file.  It is used by tests/, by __graft_entry__.smoke() and by bench.py's
 * cpu_baseline / --impl reference legs as the *checker* for the CUDA kernels.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against fixtures
 * minted from the live reference (oracle/gen_golden.py -> tests/golden/), and in the build container
 * tests/test_oracle_vs_reference.py re-runs the live reference side by side.
 *
 * All tensors are row-major float32, BPSK convention bit 0 <-> +1.0, bit 1 <-> -1.0
 * (reference polar.py:130-132).  Build: see oracle/Makefile (-O2 -ffp-contract=off: the reference's
 * torch eager ops round after every multiply and add, so no FMA contraction is allowed here).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NPDO_API __attribute__((visibility("default")))

/* torch.sign: -1, 0, +1 (0 for +-0).  reference utils.py:274, polar.py:479 */
static inline float sgnf(float v) { return (float)((v > 0.0f) - (v < 0.0f)); }

/* utils.py:272-275 min_sum_log_sum_exp: torch.min(|x|,|y|) * sign(x) * sign(y), left to right */
static inline float f_minsum(float a, float b)
{
    float m = fminf(fabsf(a), fabsf(b));
    float t = m * sgnf(a);
    return t * sgnf(b);
}

/* One Plotkin stage over a length-N row of +-1/0 floats: for every block of 2*h the left half
 * becomes left*right.  polar.py:140-144 (encoder) and 456-462 (partial sums), pac_code.py:213-217. */
static void plotkin_stage(float *u, int N, int h)
{
    for (int i = 0; i < N; i += 2 * h)
        for (int j = 0; j < h; ++j)
            u[i + j] = u[i + j] * u[i + h + j];
}

/* polar.py:128-148 encode_plotkin (scaling=None): u = ones; u[info] = msg; n Plotkin stages. */
NPDO_API void npdo_polar_encode(const float *msg, int64_t B, int n, int K, const int32_t *info,
                                float *x)
{
    const int N = 1 << n;
    for (int64_t b = 0; b < B; ++b) {
        float *u = x + b * N;
        for (int i = 0; i < N; ++i) u[i] = 1.0f;
        for (int k = 0; k < K; ++k) u[info[k]] = msg[b * K + k];
        for (int d = 0; d < n; ++d) plotkin_stage(u, N, 1 << d);
    }
}

/* Work arrays shared by the two SC decoders: llr[n+1][N], ps[n+1][N]
 * (polar.py:361-366 define_partial_arrays). */
typedef struct {
    int n, N;
    float *llr; /* (n+1) * N */
    float *ps;  /* (n+1) * N */
    float *tmp; /* N */
} sc_work_t;

static int work_alloc(sc_work_t *w, int n)
{
    memset(w, 0, sizeof(*w));
    w->n = n;
    w->N = 1 << n;
    w->llr = (float *)calloc((size_t)(n + 1) * w->N, sizeof(float));
    w->ps = (float *)calloc((size_t)(n + 1) * w->N, sizeof(float));
    w->tmp = (float *)calloc((size_t)w->N, sizeof(float));
    return (w->llr && w->ps && w->tmp) ? 0 : -1;
}

static void work_free(sc_work_t *w)
{
    free(w->llr);
    free(w->ps);
    free(w->tmp);
}

/* polar.py:369-449 updateLLR + partial_decode, flattened: every recursive call either computes the
 * left child with f (when the leaf lies in the left half) or the right child with g (û taken from the
 * partial-sum array one level down) and descends; depth 1 adds the prior to the stored leaf value.
 * `prior` may be NULL (pac_code.py:265-345 has no priors). */
static void update_llr(sc_work_t *w, int leaf, const float *prior)
{
    const int N = w->N;
    int bitpos = 0;
    for (int depth = w->n; depth >= 1; --depth) {
        const int h = 1 << (depth - 1);
        const int at_depth = leaf / h;
        const int left = 2 * bitpos;
        const float *par = w->llr + (size_t)depth * N;
        float *chl = w->llr + (size_t)(depth - 1) * N;
        const float *a = par + (size_t)left * h;
        const float *b = par + (size_t)(left + 1) * h;
        if (at_depth == left) {
            for (int j = 0; j < h; ++j) chl[(size_t)left * h + j] = f_minsum(a[j], b[j]);
            bitpos = left;
        } else {
            const float *uh = w->ps + (size_t)(depth - 1) * N + (size_t)left * h;
            for (int j = 0; j < h; ++j) {
                float t = uh[j] * a[j]; /* polar.py:414/445: u_hat * L_left + L_right, two roundings */
                chl[(size_t)(left + 1) * h + j] = t + b[j];
            }
            bitpos = left + 1;
        }
        if (depth == 1 && prior) {
            /* polar.py:399, 415: stored leaf = L + prior[pos] * ones */
            float p = prior[bitpos] * 1.0f;
            chl[bitpos] = chl[bitpos] + p;
        }
    }
}

/* polar.py:451-463 updatePartialSums: re-encode [u_0..u_leaf, 0, ..., 0] from scratch, keeping the
 * row after every stage. */
static void update_partial_sums(sc_work_t *w, int leaf, const float *u_hat)
{
    const int N = w->N;
    float *u = w->tmp;
    for (int i = 0; i < N; ++i) u[i] = (i <= leaf) ? u_hat[i] : 0.0f;
    for (int d = 0; d < w->n; ++d) {
        memcpy(w->ps + (size_t)d * N, u, sizeof(float) * N);
        plotkin_stage(u, N, 1 << d);
    }
    memcpy(w->ps + (size_t)w->n * N, u, sizeof(float) * N);
}

/* polar.py:465-484 sc_decode_new.
 *   y[B,N], scale = fp32(2/sigma^2) (the Python scalar is rounded to fp32 before the multiply),
 *   frozen[N] (1 = frozen), infty (polar.py:81), use_gt[B,N] or NULL,
 *   out: leaf_llr[B,N] (llr_array[:,0,:], includes the prior), u_hat[B,N] (all N decisions; callers
 *   index info positions), decoded[B,K] (u_hat[:, info]) -- any output may be NULL. */
NPDO_API int npdo_sc_decode(const float *y, int64_t B, int n, int K, const int32_t *info,
                            const uint8_t *frozen, float scale, float infty, const float *use_gt,
                            float *leaf_llr, float *u_hat_out, float *decoded)
{
    const int N = 1 << n;
    int fail = 0;
    /* codewords are independent: one work set per thread (bench.py's CPU baseline uses all cores) */
#pragma omp parallel
    {
        sc_work_t w;
        float *prior = (float *)calloc(N, sizeof(float));
        float *u_hat = (float *)calloc(N, sizeof(float));
        int bad = work_alloc(&w, n) || !prior || !u_hat;
        if (bad) {
#pragma omp atomic write
            fail = 1;
        } else {
            for (int i = 0; i < N; ++i) prior[i] = frozen[i] ? infty : 0.0f;
#pragma omp for schedule(dynamic, 1)
            for (int64_t b = 0; b < B; ++b) {
                memset(w.llr, 0, sizeof(float) * (size_t)(n + 1) * N);
                memset(w.ps, 0, sizeof(float) * (size_t)(n + 1) * N);
                memset(u_hat, 0, sizeof(float) * N);
                for (int i = 0; i < N; ++i) w.llr[(size_t)n * N + i] = scale * y[b * N + i];
                for (int ii = 0; ii < N; ++ii) {
                    update_llr(&w, ii, prior);
                    u_hat[ii] = use_gt ? use_gt[b * N + ii] : sgnf(w.llr[ii]);
                    update_partial_sums(&w, ii, u_hat);
                }
                if (leaf_llr) memcpy(leaf_llr + b * N, w.llr, sizeof(float) * N);
                if (u_hat_out) memcpy(u_hat_out + b * N, u_hat, sizeof(float) * N);
                if (decoded)
                    for (int k = 0; k < K; ++k) decoded[b * K + k] = u_hat[info[k]];
            }
        }
        free(prior);
        free(u_hat);
        if (!bad) work_free(&w);
    }
    return fail ? -1 : 0;
}

/* polar.py:793-876 scl_decode(y, snr, L, use_CRC=False) with pruneLists (777-791), literal structure: every path
 * owns its llr[n+1][N] / ps[n+1][N] arrays and its u_hat row; an information bit doubles the list (first copies take
 * sign(L), second copies -sign(L) and pay |L|), a list longer than L is pruned to the L smallest metrics kept in
 * ascending list-index order (torch.topk on -metric, then sort of the indices); a frozen bit pays |L| when
 * sign(L) != +1 and stores L + infty; the final pick re-encodes every surviving path and takes the codeword closest
 * to y in squared Euclidean distance (first minimum).  Priors are NOT part of the LLR recursion here (polar.py:801-803).
 *   out: leaf_llr[B,N] = llr_array of the chosen path (row 0), decoded[B,K] = its u_hat[info].
 * Ties: torch.topk's order among equal metrics and the float32 reduction order of the distance sum are
 * implementation details of torch; this restatement breaks metric ties towards the lower list index and sums the
 * distance in double precision.  Neither matters on real-valued noise (pinned on the fixtures). */
typedef struct {
    sc_work_t w;
    float *u_hat;
    float metric;
} scl_path_t;

static void scl_copy(scl_path_t *dst, const scl_path_t *src, int n)
{
    const int N = 1 << n;
    memcpy(dst->w.llr, src->w.llr, sizeof(float) * (size_t)(n + 1) * N);
    memcpy(dst->w.ps, src->w.ps, sizeof(float) * (size_t)(n + 1) * N);
    memcpy(dst->u_hat, src->u_hat, sizeof(float) * N);
    dst->metric = src->metric;
}

NPDO_API int npdo_scl_decode(const float *y, int64_t B, int n, int K, const int32_t *info,
                             const uint8_t *frozen, float scale, float infty, int L,
                             float *leaf_llr, float *decoded)
{
    const int N = 1 << n;
    if (L < 1 || L > 64) return -1;
    int fail = 0;
#pragma omp parallel
    {
        /* two generations of up to 2L paths (before pruning) */
        scl_path_t *cur = (scl_path_t *)calloc(2 * L, sizeof(scl_path_t));
        scl_path_t *nxt = (scl_path_t *)calloc(2 * L, sizeof(scl_path_t));
        float *x = (float *)calloc(N, sizeof(float));
        int *order = (int *)calloc(2 * L, sizeof(int));
        int bad = !cur || !nxt || !x || !order;
        for (int i = 0; !bad && i < 2 * L; ++i) {
            bad |= work_alloc(&cur[i].w, n) || work_alloc(&nxt[i].w, n);
            cur[i].u_hat = (float *)calloc(N, sizeof(float));
            nxt[i].u_hat = (float *)calloc(N, sizeof(float));
            bad |= !cur[i].u_hat || !nxt[i].u_hat;
        }
        if (bad) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int64_t b = 0; b < B; ++b) {
                int np = 1;
                memset(cur[0].w.llr, 0, sizeof(float) * (size_t)(n + 1) * N);
                memset(cur[0].w.ps, 0, sizeof(float) * (size_t)(n + 1) * N);
                memset(cur[0].u_hat, 0, sizeof(float) * N);
                cur[0].metric = 0.0f;
                for (int i = 0; i < N; ++i) cur[0].w.llr[(size_t)n * N + i] = scale * y[b * N + i];
                for (int ii = 0; ii < N; ++ii) {
                    for (int p = 0; p < np; ++p) update_llr(&cur[p].w, ii, NULL);
                    if (frozen[ii]) {
                        for (int p = 0; p < np; ++p) {
                            const float Lv = cur[p].w.llr[ii];
                            const float pen = fabsf(Lv) * ((sgnf(Lv) != 1.0f) ? 1.0f : 0.0f);
                            cur[p].w.llr[ii] = Lv + infty * 1.0f;
                            cur[p].u_hat[ii] = 1.0f;
                            update_partial_sums(&cur[p].w, ii, cur[p].u_hat);
                            cur[p].metric = cur[p].metric + pen;
                        }
                    } else {
                        /* duplicate: slot p keeps sign(L), slot np + p takes -sign(L) and pays |L| */
                        for (int p = 0; p < np; ++p) {
                            const float Lv = cur[p].w.llr[ii];
                            scl_copy(&cur[np + p], &cur[p], n);
                            cur[p].u_hat[ii] = sgnf(Lv);
                            cur[np + p].u_hat[ii] = -1.0f * sgnf(Lv);
                            cur[np + p].metric = cur[p].metric + fabsf(Lv);
                            update_partial_sums(&cur[p].w, ii, cur[p].u_hat);
                            update_partial_sums(&cur[np + p].w, ii, cur[np + p].u_hat);
                        }
                        np *= 2;
                        if (np > L) {
                            /* pruneLists: the L smallest metrics, kept in ascending index order */
                            int kept = 0;
                            for (int i = 0; i < np; ++i) {
                                int rank = 0;
                                for (int j = 0; j < np; ++j)
                                    rank += (cur[j].metric < cur[i].metric) || (cur[j].metric == cur[i].metric && j < i);
                                if (rank < L) order[kept++] = i;
                            }
                            for (int j = 0; j < L; ++j) scl_copy(&nxt[j], &cur[order[j]], n);
                            scl_path_t *t = cur; cur = nxt; nxt = t;
                            np = L;
                        }
                    }
                }
                /* ML pick among the list (polar.py:869-874) */
                int best = 0;
                double best_d = 0.0;
                for (int p = 0; p < np; ++p) {
                    for (int i = 0; i < N; ++i) x[i] = 1.0f;
                    for (int k = 0; k < K; ++k) x[info[k]] = cur[p].u_hat[info[k]];
                    for (int d = 0; d < n; ++d) plotkin_stage(x, N, 1 << d);
                    double dist = 0.0;
                    for (int i = 0; i < N; ++i) {
                        const float df = x[i] - y[b * N + i];
                        dist += (double)(df * df);
                    }
                    if (p == 0 || dist < best_d) { best = p; best_d = dist; }
                }
                if (leaf_llr) memcpy(leaf_llr + b * N, cur[best].w.llr, sizeof(float) * N);
                if (decoded)
                    for (int k = 0; k < K; ++k) decoded[b * K + k] = cur[best].u_hat[info[k]];
            }
        }
        for (int i = 0; cur && nxt && i < 2 * L; ++i) {
            work_free(&cur[i].w); work_free(&nxt[i].w);
            free(cur[i].u_hat); free(nxt[i].u_hat);
        }
        free(cur); free(nxt); free(x); free(order);
    }
    return fail ? -1 : 0;
}

/* pac_code.py:193-200 conv1bTrans_batch for one row.  g[M] holds +-1 (1 - 2*bit, MSB first,
 * pac_code.py:102-103); state[M-1] holds the previous inputs, newest first.
 * Returns u; writes next state into nxt (may alias neither). */
static float conv1b(float v, const float *state, const float *g, int M, float *nxt)
{
    float u = v * (0.5f * (1.0f - g[0]));
    for (int j = 1; j < M; ++j)
        if (g[j] == -1.0f) u = u * state[j - 1];
    if (nxt) {
        nxt[0] = v;
        for (int j = 1; j < M - 1; ++j) nxt[j] = state[j - 1];
    }
    return u;
}

/* pac_code.py:220-224 pac_encode = rate profile (u[B]=msg, rest +1; 171-172) ->
 * convolutional_encode (202-208) -> polar_encode (210-218). */
NPDO_API void npdo_pac_encode(const float *msg, int64_t B, int n, int K, const int32_t *info,
                              const float *g, int M, float *x)
{
    const int N = 1 << n;
    float *v = (float *)malloc(sizeof(float) * N);
    float st[64], nx[64];
    for (int64_t b = 0; b < B; ++b) {
        float *u = x + b * N;
        for (int i = 0; i < N; ++i) v[i] = 1.0f;
        for (int k = 0; k < K; ++k) v[info[k]] = msg[b * K + k];
        for (int j = 0; j < M - 1; ++j) st[j] = 1.0f;
        for (int i = 0; i < N; ++i) {
            u[i] = conv1b(v[i], st, g, M, nx);
            memcpy(st, nx, sizeof(float) * (M - 1));
        }
        for (int d = 0; d < n; ++d) plotkin_stage(u, N, 1 << d);
    }
    free(v);
}

/* pac_code.py:534-573 pac_sc_decode.
 *   out: leaf_llr[B,N] (no priors), v_hat[B,K] (= v_hat[:, B-set]), u_hat[B,N]. */
NPDO_API int npdo_pac_sc_decode(const float *y, int64_t B, int n, int K, const int32_t *info,
                                const uint8_t *frozen, const float *g, int M, float scale,
                                const float *use_gt_codeword, float *leaf_llr, float *v_hat_out,
                                float *u_hat_out)
{
    sc_work_t w;
    if (work_alloc(&w, n)) return -1;
    const int N = w.N;
    float *u_hat = (float *)calloc(N, sizeof(float));
    float *v_hat = (float *)calloc(N, sizeof(float));
    float st[64], s0[64], s1[64];
    for (int64_t b = 0; b < B; ++b) {
        memset(w.llr, 0, sizeof(float) * (size_t)(n + 1) * N);
        memset(w.ps, 0, sizeof(float) * (size_t)(n + 1) * N);
        memset(u_hat, 0, sizeof(float) * N);
        memset(v_hat, 0, sizeof(float) * N);
        for (int j = 0; j < M - 1; ++j) st[j] = 1.0f;
        for (int i = 0; i < N; ++i) w.llr[(size_t)n * N + i] = scale * y[b * N + i];
        for (int ii = 0; ii < N; ++ii) {
            update_llr(&w, ii, NULL);
            if (frozen[ii]) {
                v_hat[ii] = 1.0f;
                if (use_gt_codeword) {
                    u_hat[ii] = use_gt_codeword[b * N + ii]; /* state is NOT advanced (548-549) */
                } else {
                    u_hat[ii] = conv1b(1.0f, st, g, M, s0);
                    memcpy(st, s0, sizeof(float) * (M - 1));
                }
            } else {
                u_hat[ii] = use_gt_codeword ? use_gt_codeword[b * N + ii] : sgnf(w.llr[ii]);
                float u0 = conv1b(1.0f, st, g, M, s0);
                float u1 = conv1b(-1.0f, st, g, M, s1);
                int z = (u0 == u_hat[ii]);
                int o = (u1 == u_hat[ii]);
                /* 561-568: z branch applied first, then o branch (o wins if both match) */
                float cur[64];
                memcpy(cur, st, sizeof(float) * (M - 1));
                if (z) { v_hat[ii] = 1.0f; memcpy(cur, s0, sizeof(float) * (M - 1)); }
                if (o) { v_hat[ii] = -1.0f; memcpy(cur, s1, sizeof(float) * (M - 1)); }
                memcpy(st, cur, sizeof(float) * (M - 1));
            }
            update_partial_sums(&w, ii, u_hat);
        }
        if (leaf_llr) memcpy(leaf_llr + b * N, w.llr, sizeof(float) * N);
        if (u_hat_out) memcpy(u_hat_out + b * N, u_hat, sizeof(float) * N);
        if (v_hat_out)
            for (int k = 0; k < K; ++k) v_hat_out[b * K + k] = v_hat[info[k]];
    }
    free(u_hat);
    free(v_hat);
    work_free(&w);
    return 0;
}

/* utils.py:17-25 errors_ber numerator and utils.py:37-51 errors_bler numerator:
 * counts[0] = #{round(a) != round(b)}, counts[1] = #rows with any mismatch.
 * torch.round is round-half-to-even == rintf in the default rounding mode. */
NPDO_API void npdo_count_errors(const float *a, const float *b, int64_t B, int K, uint64_t *counts)
{
    uint64_t bit = 0, blk = 0;
    for (int64_t r = 0; r < B; ++r) {
        int any = 0;
        for (int k = 0; k < K; ++k) {
            int ne = rintf(a[r * K + k]) != rintf(b[r * K + k]);
            bit += ne;
            any |= ne;
        }
        blk += any;
    }
    counts[0] = bit;
    counts[1] = blk;
}

/* ---- counter-based RNG restatement (not reference behaviour: the reference draws torch.randn on
 * the CPU generator, polar.py:204; the product uses Philox4x32-10 so that results are independent
 * of the GPU count).  Integer part is bit-exact vs the kernel; the Box-Muller floats
 * are compared with a tolerance because device logf/sincosf are not bit-reproducible on the host. */
static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

NPDO_API void npdo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(uint32_t) * 4);
}

/* Message bits of codeword `cw`: stream = 0, one Philox block per 128 message bits.
 * counter = (cw_lo, cw_hi, block, NPD_STREAM_MSG), key = (seed_lo, seed_hi).
 * bit k of the message = bit (k % 32) of word (k / 32) % 4 of block k / 128; msg = 1 - 2*bit. */
NPDO_API void npdo_gen_msg(uint64_t seed, uint64_t cw0, int64_t B, int K, float *msg)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int64_t b = 0; b < B; ++b) {
        uint64_t cw = cw0 + (uint64_t)b;
        for (int blk = 0; blk * 128 < K; ++blk) {
            uint32_t ctr[4] = {(uint32_t)cw, (uint32_t)(cw >> 32), (uint32_t)blk, 0u};
            uint32_t r[4];
            npdo_philox4x32_10(ctr, key, r);
            for (int k = blk * 128; k < K && k < (blk + 1) * 128; ++k) {
                int bit = (r[(k >> 5) & 3] >> (k & 31)) & 1;
                msg[b * K + k] = bit ? -1.0f : 1.0f;
            }
        }
    }
}

/* Standard-normal noise for codeword `cw`, SNR-point index `pt`: one Philox block per 4 samples,
 * counter = (cw_lo, cw_hi, quad, NPD_STREAM_NOISE + pt), Box-Muller on (r0,r1) and (r2,r3):
 *   u = (float(r) + 0.5) * 2^-32  in (0,1]  (u0 = 1 gives radius 0, never a NaN);  z0 = sqrt(-2 ln u0) cos(2 pi u1), z1 = ... sin(2 pi u1). */
NPDO_API void npdo_gen_noise(uint64_t seed, uint64_t cw0, uint32_t pt, int64_t B, int N, float *z)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int64_t b = 0; b < B; ++b) {
        uint64_t cw = cw0 + (uint64_t)b;
        for (int q = 0; q * 4 < N; ++q) {
            uint32_t ctr[4] = {(uint32_t)cw, (uint32_t)(cw >> 32), (uint32_t)q, 1u + pt};
            uint32_t r[4];
            npdo_philox4x32_10(ctr, key, r);
            for (int h = 0; h < 2; ++h) {
                float u0 = ((float)r[2 * h] + 0.5f) * 2.3283064365386963e-10f;
                float u1 = ((float)r[2 * h + 1] + 0.5f) * 2.3283064365386963e-10f;
                float rad = sqrtf(-2.0f * logf(u0));
                float ang = 6.283185307179586f * u1;
                int e = q * 4 + 2 * h;
                if (e < N) z[b * N + e] = rad * cosf(ang);
                if (e + 1 < N) z[b * N + e + 1] = rad * sinf(ang);
            }
        }
    }
}

/* Number of threads the OpenMP loops above will use (1 without OpenMP). */
NPDO_API int npdo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
