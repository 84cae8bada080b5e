// count_sweep.cu -- BER/BLER numerators and the fused Monte-Carlo SC sweep.
//
// Replaces errors_ber / errors_bler (reference utils.py:17-25, 37-51: round() both tensors, count
// mismatching elements and rows with any mismatch; the reference goes through CPU numpy for BLER) and
// the per-SNR inner loop of polar.py:1258-1291 / rnn_all.py:843-856.
#include <mutex>

#include "npd_common.cuh"

namespace {

// one warp per row chunk: lanes stride over K (coalesced), ballot for the row flag
__global__ void __launch_bounds__(256) count_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                                    int64_t B, int K, unsigned long long *counts,
                                                    unsigned long long add_frames)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long bits = 0, blocks = 0;
    const bool vec = (K & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    for (int64_t r = warp; r < B; r += nwarps) {
        uint32_t mism = 0;
        if (vec) {  // 16-byte loads: the kernel streams 8K bytes per row and is bandwidth-bound
            const float4 *a4 = reinterpret_cast<const float4 *>(a + r * K), *b4 = reinterpret_cast<const float4 *>(b + r * K);
            for (int q = lane; q < (K >> 2); q += 32) {
                const float4 x = __ldcs(a4 + q), y = __ldcs(b4 + q);
                mism += (rintf(x.x) != rintf(y.x)) + (rintf(x.y) != rintf(y.y)) + (rintf(x.z) != rintf(y.z)) +
                        (rintf(x.w) != rintf(y.w));
            }
        } else {
            for (int k = lane; k < K; k += 32)
                mism += rintf(a[r * K + k]) != rintf(b[r * K + k]);  // torch.round = half-to-even
        }
        const uint32_t any = __ballot_sync(NPD_FULL, mism != 0);
        bits += mism;
        if (lane == 0 && any) blocks += 1;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) bits += __shfl_xor_sync(NPD_FULL, bits, s);
    if (lane == 0) {
        if (bits) atomicAdd(counts + 0, bits);
        if (blocks) atomicAdd(counts + 1, blocks);
        if (warp == 0 && add_frames) atomicAdd(counts + 2, add_frames);
    }
}

// same counters for decisions that sit inside full-length rows: a[B,K] against b[r, cols[k]], b rows of length ldb
// (the GRU / convNet decoders return [B,N]; callers of the reference index [:, info_positions], rnn_all.py:875)
__global__ void __launch_bounds__(256) count_gather_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                                           const int32_t *__restrict__ cols, int64_t B, int K, int ldb,
                                                           int take_sign, unsigned long long *counts, unsigned long long add_frames)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long bits = 0, blocks = 0;
    for (int64_t r = warp; r < B; r += nwarps) {
        uint32_t mism = 0;
        for (int k = lane; k < K; k += 32) {
            float v = b[r * ldb + cols[k]];
            if (take_sign) v = v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f);  // `.sign()` of logits (run_models.py:338-339)
            mism += rintf(a[r * K + k]) != rintf(v);
        }
        const uint32_t any = __ballot_sync(NPD_FULL, mism != 0);
        bits += mism;
        if (lane == 0 && any) blocks += 1;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) bits += __shfl_xor_sync(NPD_FULL, bits, s);
    if (lane == 0) {
        if (bits) atomicAdd(counts + 0, bits);
        if (blocks) atomicAdd(counts + 1, blocks);
        if (warp == 0 && add_frames) atomicAdd(counts + 2, add_frames);
    }
}

int launch_count_gather(const float *a, const float *b, const int32_t *cols, int64_t B, int K, int ldb, int take_sign,
                        uint64_t *counts, uint64_t add_frames, cudaStream_t st)
{
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    int64_t grid = (B + 7) / 8;
    const int64_t cap = (int64_t)dp.sm_count * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    count_gather_kernel<<<(unsigned)grid, 256, 0, st>>>(a, b, cols, B, K, ldb, take_sign, (unsigned long long *)counts,
                                                        (unsigned long long)add_frames);
    NPD_CHECK_CUDA(cudaGetLastError());
    return NPD_OK;
}

int launch_count(const float *a, const float *b, int64_t B, int K, uint64_t *counts,
                 uint64_t add_frames, cudaStream_t st)
{
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    int64_t grid = (B + 7) / 8;
    const int64_t cap = (int64_t)dp.sm_count * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    count_kernel<<<(unsigned)grid, 256, 0, st>>>(a, b, B, K, (unsigned long long *)counts,
                                                 (unsigned long long)add_frames);
    NPD_CHECK_CUDA(cudaGetLastError());
    return NPD_OK;
}

}  // namespace

NPD_API int npd_count_errors(const float *a, const float *b, int64_t B, int K, uint64_t *counts,
                             void *stream)
{
    NPD_REQUIRE(B >= 0 && K >= 1, "npd_count_errors: bad shape");
    if (B == 0) return NPD_OK;  // an empty tensor has no storage: nothing to count, nothing to dereference
    NPD_REQUIRE(a && b && counts, "npd_count_errors: null argument");
    return launch_count(a, b, B, K, counts, 0, (cudaStream_t)stream);
}

NPD_API int npd_count_errors_info(const npd_code_t *code, const float *msg, const float *decoded_full, int64_t B,
                                  int take_sign, uint64_t *counts, void *stream)
{
    NPD_REQUIRE(code && B >= 0 && code->K >= 1, "npd_count_errors_info: bad shape");
    if (B == 0) return NPD_OK;
    NPD_REQUIRE(msg && decoded_full && counts, "npd_count_errors_info: null argument");
    return launch_count_gather(msg, decoded_full, code->d_info, B, code->K, code->N, take_sign, counts, 0, (cudaStream_t)stream);
}

namespace {
// one chunk's scratch: msg[chunk,K] + decoded[chunk,K] + y[chunk,N], each region 256-byte aligned
size_t mc_slot_bytes(const npd_code_t *code, int64_t chunk)
{
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    // plain layout: msg + decisions + y; fused-count layout (N >= 256): packed u words + flags + re-decode scratch + y
    const size_t plain = al((size_t)chunk * code->K * 4) * 2 + al((size_t)chunk * code->N * 4);
    const size_t fused = al((size_t)chunk * (code->N >> 5) * 4) + al((size_t)chunk) + al((size_t)chunk * code->K * 4) +
                         al((size_t)chunk * code->N * 4);
    return plain > fused ? plain : fused;
}

// the decoder counts its own errors (sc_quad_kernel's fused mode): no float messages / decisions, no count launch
bool mc_fused_count(const npd_code_t *code) { return code->pac_g == 0 && code->n >= 8 && code->K >= 1; }

// second stream + events of the sweep's two-deep pipeline, one set per device (created on first use)
struct SweepPipe {
    std::mutex mu;  // the enqueue of one sweep call is atomic with respect to other host threads on the device
    cudaStream_t gen = nullptr;
    cudaEvent_t generated[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr}, start = nullptr;
};
SweepPipe g_sweep[64];
}  // namespace

NPD_API size_t npd_mc_sc_workspace_bytes(const npd_code_t *code, int64_t chunk)
{
    if (!code || chunk <= 0) return 0;
    return 2 * mc_slot_bytes(code, chunk);  // two chunks in flight: one being generated, one being decoded
}

NPD_API int npd_mc_sc_sweep(const npd_code_t *code, int64_t B, int64_t chunk, float sigma,
                            float llr_scale, uint64_t seed, uint32_t point, uint64_t cw_offset,
                            void *workspace, size_t workspace_bytes, uint64_t *counts, void *stream)
{
    NPD_REQUIRE(code && workspace && counts, "npd_mc_sc_sweep: null argument");
    NPD_REQUIRE(B >= 0 && chunk > 0, "npd_mc_sc_sweep: bad sizes");
    NPD_REQUIRE(code->pac_g == 0, "npd_mc_sc_sweep: polar code objects only");
    NPD_REQUIRE(workspace_bytes >= npd_mc_sc_workspace_bytes(code, chunk),
                "npd_mc_sc_sweep: workspace too small (%zu < %zu)", workspace_bytes,
                npd_mc_sc_workspace_bytes(code, chunk));
    // Two chunks in flight: chunk i+1 is generated (message bits, encoder, channel) on a second stream while chunk i
    // is decoded and counted on the caller's stream -- the generator (Philox / Box-Muller, issue-bound, little shared
    // memory) and the SC kernel (shared-memory-bound occupancy, 12 warps per SM) fill different resources of an SM.
    int dev = 0;
    NPD_CHECK_CUDA(cudaGetDevice(&dev));
    NPD_REQUIRE(dev < 64, "npd_mc_sc_sweep: device index %d not supported", dev);
    SweepPipe &sp = g_sweep[dev];
    std::lock_guard<std::mutex> lk(sp.mu);
    if (!sp.gen) {
        NPD_CHECK_CUDA(cudaStreamCreateWithFlags(&sp.gen, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            NPD_CHECK_CUDA(cudaEventCreateWithFlags(&sp.generated[i], cudaEventDisableTiming));
            NPD_CHECK_CUDA(cudaEventCreateWithFlags(&sp.consumed[i], cudaEventDisableTiming));
        }
        NPD_CHECK_CUDA(cudaEventCreateWithFlags(&sp.start, cudaEventDisableTiming));
    }
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t slot_bytes = mc_slot_bytes(code, chunk);
    cudaStream_t main_st = (cudaStream_t)stream;
    // the generator stream starts after whatever the caller queued before this call (e.g. zeroing the counters)
    NPD_CHECK_CUDA(cudaEventRecord(sp.start, main_st));
    NPD_CHECK_CUDA(cudaStreamWaitEvent(sp.gen, sp.start, 0));
    int64_t n_chunks = 0;
    for (int64_t done = 0; done < B; done += chunk, ++n_chunks) {
        const int slot = (int)(n_chunks & 1);
        const int64_t b = (B - done < chunk) ? (B - done) : chunk;
        char *ws = (char *)workspace + slot * slot_bytes;
        if (n_chunks >= 2) NPD_CHECK_CUDA(cudaStreamWaitEvent(sp.gen, sp.consumed[slot], 0));  // slot's previous chunk counted
        int rc;
        if (mc_fused_count(code)) {
            uint32_t *ubits = (uint32_t *)ws;
            unsigned char *flags = (unsigned char *)(ws + al((size_t)chunk * (code->N >> 5) * 4));
            float *dec = (float *)((char *)flags + al((size_t)chunk));
            float *y = (float *)((char *)dec + al((size_t)chunk * code->K * 4));
            rc = npd_gen_encode_awgn_bits(code, ubits, y, b, sigma, seed, point, cw_offset + done, sp.gen);
            if (rc) return rc;
            NPD_CHECK_CUDA(cudaEventRecord(sp.generated[slot], sp.gen));
            NPD_CHECK_CUDA(cudaStreamWaitEvent(main_st, sp.generated[slot], 0));
            rc = npd_sc_decode_count(code, y, llr_scale, ubits, dec, flags, b, counts, main_st);
            if (rc) return rc;
        } else {
            float *msg = (float *)ws;
            float *dec = (float *)(ws + al((size_t)chunk * code->K * 4));
            float *y = (float *)(ws + 2 * al((size_t)chunk * code->K * 4));
            rc = npd_gen_encode_awgn(code, msg, nullptr, y, b, sigma, seed, point, cw_offset + done, sp.gen);
            if (rc) return rc;
            NPD_CHECK_CUDA(cudaEventRecord(sp.generated[slot], sp.gen));
            NPD_CHECK_CUDA(cudaStreamWaitEvent(main_st, sp.generated[slot], 0));
            rc = npd_sc_decode(code, y, llr_scale, nullptr, nullptr, dec, b, main_st);
            if (rc) return rc;
            rc = launch_count(msg, dec, b, code->K, counts, (uint64_t)b, main_st);
            if (rc) return rc;
        }
        NPD_CHECK_CUDA(cudaEventRecord(sp.consumed[slot], main_st));
    }
    return NPD_OK;
}

// ---- fused Monte-Carlo sweep for the CRISP GRU decoder ------------------------------------------------------------
// generate -> encode -> AWGN -> npd_gru_decode -> count, chunk by chunk on the caller's stream, no host round trip.
// The decode (0.3 us per codeword) is ~100x the generator and the counter, so the chunks simply run back to back.
NPD_API size_t npd_mc_gru_workspace_bytes(const npd_gru_t *gru, const npd_code_t *code, int64_t chunk)
{
    if (!gru || !code || chunk <= 0) return 0;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return al((size_t)chunk * code->K * 4) + 2 * al((size_t)chunk * code->N * 4) + al(npd_gru_workspace_bytes(gru, chunk));
}

NPD_API int npd_mc_gru_sweep(const npd_gru_t *gru, const npd_code_t *code, const npd_code_t *loss_code, int64_t B,
                             int64_t chunk, float sigma, uint64_t seed, uint32_t point, uint64_t cw_offset,
                             void *workspace, size_t workspace_bytes, uint64_t *counts, void *stream)
{
    NPD_REQUIRE(gru && code && workspace && counts, "npd_mc_gru_sweep: null argument");
    NPD_REQUIRE(B >= 0 && chunk > 0, "npd_mc_gru_sweep: bad sizes");
    NPD_REQUIRE(workspace_bytes >= npd_mc_gru_workspace_bytes(gru, code, chunk),
                "npd_mc_gru_sweep: workspace too small (%zu < %zu)", workspace_bytes,
                npd_mc_gru_workspace_bytes(gru, code, chunk));
    if (!loss_code) loss_code = code;  // decisions on the info positions (RNN_decoder.decode's default loss_inds)
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    char *ws = (char *)workspace;
    float *msg = (float *)ws;
    float *y = (float *)(ws + al((size_t)chunk * code->K * 4));
    float *dec = (float *)((char *)y + al((size_t)chunk * code->N * 4));
    void *gws = (char *)dec + al((size_t)chunk * code->N * 4);
    const size_t gws_bytes = npd_gru_workspace_bytes(gru, chunk);
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t done = 0; done < B; done += chunk) {
        const int64_t b = (B - done < chunk) ? (B - done) : chunk;
        int rc = npd_gen_encode_awgn(code, msg, nullptr, y, b, sigma, seed, point, cw_offset + done, stream);
        if (rc) return rc;
        rc = npd_gru_decode(gru, loss_code, y, nullptr, nullptr, nullptr, dec, b, gws_bytes ? gws : nullptr, gws_bytes, stream);
        if (rc) return rc;
        rc = launch_count_gather(msg, dec, code->d_info, b, code->K, code->N, 0, counts, (uint64_t)b, st);
        if (rc) return rc;
    }
    return NPD_OK;
}
