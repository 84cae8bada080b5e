// host_pipe.cu -- host-buffer entry points (npd_*_host).
//
// The reference's evaluation loops hand the decoders HOST tensors whenever the model runs on the CPU
// (rnn_all.py:1771-1773 builds the test set on the CPU, polar.py:204 draws the noise there) and read
// the decisions back on the host (utils.py:41-45).  These entry points take those host buffers as
// they are and overlap the three legs of the trip: the batch is cut into chunks, and chunk i+1's
// host->device copy, chunk i's kernel and chunk i-1's device->host copy run on three private streams
// (PCIe is full duplex, so the copies in both directions overlap too).  Device staging memory is
// owned by a per-device pipe object and only ever grows; the calls are synchronous (the outputs are
// complete in host memory on return) and serialised per device by a mutex.
#include <atomic>
#include <stdlib.h>

#include <mutex>

#include "npd_common.cuh"

namespace {

constexpr int kSlots = 3;

struct Slot {
    cudaStream_t st = nullptr;
    char *buf = nullptr;
    size_t cap = 0;
};

struct HostPipe {
    std::mutex mu;
    bool ready = false;
    Slot slot[kSlots];
};

HostPipe g_pipe[64];

inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

int pipe_prepare(HostPipe &p, size_t bytes_per_slot)
{
    if (!p.ready) {
        for (int s = 0; s < kSlots; ++s)
            NPD_CHECK_CUDA(cudaStreamCreateWithFlags(&p.slot[s].st, cudaStreamNonBlocking));
        p.ready = true;
    }
    for (int s = 0; s < kSlots; ++s) {
        Slot &sl = p.slot[s];
        if (sl.cap >= bytes_per_slot) continue;
        NPD_CHECK_CUDA(cudaStreamSynchronize(sl.st));
        if (sl.buf) NPD_CHECK_CUDA(cudaFree(sl.buf));
        sl.buf = nullptr;
        sl.cap = 0;
        cudaError_t e = cudaMalloc(&sl.buf, bytes_per_slot);
        if (e != cudaSuccess) {
            npd_set_error("host pipe: cudaMalloc(%zu) -> %s", bytes_per_slot, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? NPD_ENOMEM : NPD_ECUDA;
        }
        sl.cap = bytes_per_slot;
    }
    return NPD_OK;
}

// Carves a slot's arena into 256-byte aligned regions.
struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(char *b) : base(b) {}
    float *take(size_t bytes, bool wanted = true)
    {
        if (!wanted) return nullptr;
        float *p = (float *)(base + off);
        off += al256(bytes);
        return p;
    }
};

std::atomic<int64_t> g_chunk_rows{0};  // npd_host_set_chunk: 0 = sized automatically

int64_t env_chunk() { return g_chunk_rows.load(std::memory_order_relaxed); }

// body(slot arena, lo, n, stream) enqueues H2D + kernels + D2H for rows [lo, lo+n) on `stream`.
template <class Body>
int run_pipe(int64_t B, int64_t chunk, size_t bytes_per_slot, Body body)
{
    int dev = 0;
    NPD_CHECK_CUDA(cudaGetDevice(&dev));
    NPD_REQUIRE(dev < 64, "host pipe: device index %d not supported", dev);
    HostPipe &p = g_pipe[dev];
    std::lock_guard<std::mutex> lk(p.mu);
    int rc = pipe_prepare(p, bytes_per_slot);
    if (rc) return rc;
    int s = 0;
    for (int64_t lo = 0; lo < B; lo += chunk, s = (s + 1) % kSlots) {
        const int64_t n = (B - lo < chunk) ? (B - lo) : chunk;
        // stream order protects the slot's arena: its previous chunk has fully drained before the
        // next host->device copy on the same stream starts
        rc = body(p.slot[s].buf, lo, n, p.slot[s].st);
        if (rc) break;
    }
    for (int i = 0; i < kSlots; ++i) {
        cudaError_t e = cudaStreamSynchronize(p.slot[i].st);
        if (e != cudaSuccess && rc == NPD_OK) {
            npd_set_error("host pipe: %s", cudaGetErrorString(e));
            rc = NPD_ECUDA;
        }
    }
    return rc;
}

#define H2D(dst, src, bytes, st) \
    NPD_CHECK_CUDA(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, (st)))
#define D2H(dst, src, bytes, st) \
    NPD_CHECK_CUDA(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, (st)))

int64_t pick_chunk(int64_t B, size_t row_bytes, int64_t granule, size_t target_bytes)
{
    int64_t c = env_chunk();
    if (c <= 0) c = (int64_t)(target_bytes / (row_bytes ? row_bytes : 1));
    if (c < granule) c = granule;
    c = (c + granule - 1) / granule * granule;
    // at least kSlots chunks when the batch allows, so that the three legs overlap
    if (c > granule && B / c < kSlots) {
        int64_t c2 = ((B + kSlots - 1) / kSlots + granule - 1) / granule * granule;
        if (c2 < c) c = c2 < granule ? granule : c2;
    }
    return c;
}

}  // namespace

NPD_API int npd_host_set_chunk(int64_t rows)
{
    NPD_REQUIRE(rows >= 0, "npd_host_set_chunk: rows < 0");
    g_chunk_rows.store(rows, std::memory_order_relaxed);
    return NPD_OK;
}

NPD_API int npd_sc_decode_host(const npd_code_t *code, const float *h_y, float llr_scale,
                               const float *h_use_gt, float *h_leaf_llr, float *h_decoded, int64_t B)
{
    NPD_REQUIRE(code && h_y && (h_decoded || code->K == 0), "npd_sc_decode_host: null argument");
    NPD_REQUIRE(B >= 0, "npd_sc_decode_host: B < 0");
    if (B == 0) return NPD_OK;
    const size_t N = code->N, K = code->K;
    const int64_t chunk = pick_chunk(B, N * 4, 256, (size_t)16 << 20);
    const size_t per = al256(chunk * N * 4) * (1 + (h_use_gt ? 1 : 0) + (h_leaf_llr ? 1 : 0)) +
                       al256(chunk * (K ? K : 1) * 4);
    return run_pipe(B, chunk, per, [&](char *arena, int64_t lo, int64_t n, cudaStream_t st) -> int {
        Carver cv(arena);
        float *d_y = cv.take(chunk * N * 4);
        float *d_gt = cv.take(chunk * N * 4, h_use_gt != nullptr);
        float *d_llr = cv.take(chunk * N * 4, h_leaf_llr != nullptr);
        float *d_dec = cv.take(chunk * (K ? K : 1) * 4);
        H2D(d_y, h_y + lo * N, n * N * 4, st);
        if (d_gt) H2D(d_gt, h_use_gt + lo * N, n * N * 4, st);
        int rc = npd_sc_decode(code, d_y, llr_scale, d_gt, d_llr, d_dec, n, st);
        if (rc) return rc;
        if (K) D2H(h_decoded + lo * K, d_dec, n * K * 4, st);
        if (d_llr) D2H(h_leaf_llr + lo * N, d_llr, n * N * 4, st);
        return NPD_OK;
    });
}

NPD_API int npd_scl_decode_host(const npd_code_t *code, const float *h_y, float llr_scale, int list_size,
                                float *h_leaf_llr, float *h_decoded, int64_t B)
{
    NPD_REQUIRE(code && h_y && (h_decoded || code->K == 0), "npd_scl_decode_host: null argument");
    NPD_REQUIRE(B >= 0, "npd_scl_decode_host: B < 0");
    if (B == 0) return NPD_OK;
    const size_t N = code->N, K = code->K;
    const int64_t chunk = pick_chunk(B, N * 4, 256, (size_t)4 << 20);
    const size_t per = al256(chunk * N * 4) * (1 + (h_leaf_llr ? 1 : 0)) + al256(chunk * (K ? K : 1) * 4);
    return run_pipe(B, chunk, per, [&](char *arena, int64_t lo, int64_t n, cudaStream_t st) -> int {
        Carver cv(arena);
        float *d_y = cv.take(chunk * N * 4);
        float *d_llr = cv.take(chunk * N * 4, h_leaf_llr != nullptr);
        float *d_dec = cv.take(chunk * (K ? K : 1) * 4);
        H2D(d_y, h_y + lo * N, n * N * 4, st);
        int rc = npd_scl_decode(code, d_y, llr_scale, list_size, d_llr, d_dec, n, st);
        if (rc) return rc;
        if (K) D2H(h_decoded + lo * K, d_dec, n * K * 4, st);
        if (d_llr) D2H(h_leaf_llr + lo * N, d_llr, n * N * 4, st);
        return NPD_OK;
    });
}

NPD_API int npd_pac_sc_decode_host(const npd_code_t *code, const float *h_y, float llr_scale,
                                   const float *h_use_gt_codeword, float *h_leaf_llr, float *h_v_hat,
                                   float *h_u_hat, int64_t B)
{
    NPD_REQUIRE(code && h_y && h_v_hat, "npd_pac_sc_decode_host: null argument");
    NPD_REQUIRE(B >= 0, "npd_pac_sc_decode_host: B < 0");
    if (B == 0) return NPD_OK;
    const size_t N = code->N, K = code->K;
    const int64_t chunk = pick_chunk(B, N * 4, 256, (size_t)16 << 20);
    const size_t per = al256(chunk * N * 4) * (1 + (h_use_gt_codeword ? 1 : 0) + (h_leaf_llr ? 1 : 0) +
                                              (h_u_hat ? 1 : 0)) +
                       al256(chunk * (K ? K : 1) * 4);
    return run_pipe(B, chunk, per, [&](char *arena, int64_t lo, int64_t n, cudaStream_t st) -> int {
        Carver cv(arena);
        float *d_y = cv.take(chunk * N * 4);
        float *d_gt = cv.take(chunk * N * 4, h_use_gt_codeword != nullptr);
        float *d_llr = cv.take(chunk * N * 4, h_leaf_llr != nullptr);
        float *d_u = cv.take(chunk * N * 4, h_u_hat != nullptr);
        float *d_v = cv.take(chunk * (K ? K : 1) * 4);
        H2D(d_y, h_y + lo * N, n * N * 4, st);
        if (d_gt) H2D(d_gt, h_use_gt_codeword + lo * N, n * N * 4, st);
        int rc = npd_pac_sc_decode(code, d_y, llr_scale, d_gt, d_llr, d_v, d_u, n, st);
        if (rc) return rc;
        if (K) D2H(h_v_hat + lo * K, d_v, n * K * 4, st);
        if (d_llr) D2H(h_leaf_llr + lo * N, d_llr, n * N * 4, st);
        if (d_u) D2H(h_u_hat + lo * N, d_u, n * N * 4, st);
        return NPD_OK;
    });
}

NPD_API int npd_gru_decode_host(const npd_gru_t *gru, const npd_code_t *code, const float *h_y,
                                const float *h_forced, const float *h_genie, float *h_logits,
                                float *h_decoded, int64_t B)
{
    NPD_REQUIRE(gru && code && h_y && h_decoded, "npd_gru_decode_host: null argument");
    NPD_REQUIRE(B >= 0, "npd_gru_decode_host: B < 0");
    if (B == 0) return NPD_OK;
    const size_t N = code->N;
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    // one chunk = one full wave of the persistent kernel (64 codewords per CTA, one CTA per SM)
    const int64_t wave = (int64_t)64 * dp.sm_count;
    const int64_t chunk = pick_chunk(B, N * 4, 64, (size_t)wave * N * 4);
    const size_t ws = al256(npd_gru_workspace_bytes(gru, chunk));
    const size_t per = al256(chunk * N * 4) * (2 + (h_forced ? 1 : 0) + (h_genie ? 1 : 0) + (h_logits ? 1 : 0)) + ws;
    return run_pipe(B, chunk, per, [&](char *arena, int64_t lo, int64_t n, cudaStream_t st) -> int {
        Carver cv(arena);
        float *d_y = cv.take(chunk * N * 4);
        float *d_dec = cv.take(chunk * N * 4);
        float *d_forced = cv.take(chunk * N * 4, h_forced != nullptr);
        float *d_genie = cv.take(chunk * N * 4, h_genie != nullptr);
        float *d_logits = cv.take(chunk * N * 4, h_logits != nullptr);
        float *d_ws = cv.take(ws, ws != 0);
        H2D(d_y, h_y + lo * N, n * N * 4, st);
        if (d_forced) H2D(d_forced, h_forced + lo * N, n * N * 4, st);
        if (d_genie) H2D(d_genie, h_genie + lo * N, n * N * 4, st);
        int rc = npd_gru_decode(gru, code, d_y, d_forced, d_genie, d_logits, d_dec, n, d_ws, ws, st);
        if (rc) return rc;
        D2H(h_decoded + lo * N, d_dec, n * N * 4, st);
        if (d_logits) D2H(h_logits + lo * N, d_logits, n * N * 4, st);
        return NPD_OK;
    });
}

static int conv_host_impl(const npd_conv_t *conv, const float *h_y, float *h_logits, float *h_in4, int64_t B, int sign_out);

NPD_API int npd_conv_forward_host(const npd_conv_t *conv, const float *h_y, float *h_logits,
                                  float *h_in4, int64_t B)
{
    return conv_host_impl(conv, h_y, h_logits, h_in4, B, 0);
}

NPD_API int npd_conv_decode_host(const npd_conv_t *conv, const float *h_y, float *h_bits, int64_t B)
{
    return conv_host_impl(conv, h_y, h_bits, nullptr, B, 1);
}

static int conv_host_impl(const npd_conv_t *conv, const float *h_y, float *h_logits, float *h_in4, int64_t B, int sign_out)
{
    NPD_REQUIRE(conv && h_y && h_logits, "npd_conv_forward_host: null argument");
    NPD_REQUIRE(B >= 0, "npd_conv_forward_host: B < 0");
    if (B == 0) return NPD_OK;
    int N = 0, in4_channels = 0;
    npd_conv_dims(conv, &N, &in4_channels);
    const size_t n_ = N, c4 = h_in4 ? (size_t)in4_channels : 0;
    const int64_t chunk = pick_chunk(B, n_ * 4, 128, (size_t)16384 * n_ * 4);
    const size_t ws = al256(npd_conv_workspace_bytes(conv, chunk));
    const size_t per = al256(chunk * n_ * 4) * 2 + al256(chunk * c4 * n_ * 4) + ws;
    return run_pipe(B, chunk, per, [&](char *arena, int64_t lo, int64_t n, cudaStream_t st) -> int {
        Carver cv(arena);
        float *d_y = cv.take(chunk * n_ * 4);
        float *d_logits = cv.take(chunk * n_ * 4);
        float *d_in4 = cv.take(chunk * c4 * n_ * 4, h_in4 != nullptr);
        float *d_ws = cv.take(ws, ws != 0);
        H2D(d_y, h_y + lo * n_, n * n_ * 4, st);
        int rc = sign_out ? npd_conv_decode(conv, d_y, d_logits, n, d_ws, ws, st)
                          : npd_conv_forward(conv, d_y, d_logits, d_in4, n, d_ws, ws, st);
        if (rc) return rc;
        D2H(h_logits + lo * n_, d_logits, n * n_ * 4, st);
        if (d_in4) D2H(h_in4 + lo * c4 * n_, d_in4, n * c4 * n_ * 4, st);
        return NPD_OK;
    });
}
