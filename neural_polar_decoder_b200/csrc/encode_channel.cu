// encode_channel.cu -- message generation, Plotkin-butterfly / PAC encoder, BPSK + AWGN channel.
//
// Replaces PolarCode.encode_plotkin (reference polar.py:128-148), PAC.pac_encode
// (pac_code.py:220-224 = rate_profiler 121-176 -> convolutional_encode 193-208 -> polar_encode 210-218)
// and PolarCode.channel / PAC.channel (polar.py:201-207, pac_code.py:226-231).
//
// The reference multiplies +-1 floats through N-1 torch.cat calls; here a codeword is N bits in
// shared-memory words (bit i of the row = position i): one Plotkin stage d is
//     x ^= (x >> 2^d) & M_d        (left half <- left xor right)
// inside a word for d < 5 and a word-wise xor for d >= 5.  One warp encodes one codeword; the BPSK map
// 1 - 2*bit, the Philox/Box-Muller noise and the add are fused into the 128-bit output stores.
#include "npd_common.cuh"

namespace {

struct EncParams {
    const float *msg_in;  // [B,K] or null (null => generate from Philox)
    float *msg_out;       // [B,K] or null
    float *x_out;         // [B,N] or null
    float *y_out;         // [B,N] or null
    uint32_t *ubits_out;  // [B, N/32] or null: the rate-profiled word V (message bits at the info positions), packed
    const int32_t *info;
    int64_t B;
    int n, K;
    uint32_t pac_taps;
    int pac_M;            // 0 = polar
    float sigma;
    uint64_t seed, cw_offset;
    uint32_t point;
};

__device__ __forceinline__ uint32_t word_stages(uint32_t x)
{
    x ^= (x >> 1) & 0x55555555u;
    x ^= (x >> 2) & 0x33333333u;
    x ^= (x >> 4) & 0x0F0F0F0Fu;
    x ^= (x >> 8) & 0x00FF00FFu;
    x ^= (x >> 16) & 0x0000FFFFu;
    return x;
}

// smem per warp: 2 * NW words
__global__ void __launch_bounds__(128) encode_kernel(const EncParams p)
{
    extern __shared__ uint32_t sm_words[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int n = p.n, N = 1 << n, NW = (N + 31) >> 5, K = p.K;
    uint32_t *W = sm_words + (size_t)warp * 2 * NW;
    uint32_t *V = W + NW;
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));

    for (int64_t r = (int64_t)blockIdx.x * wpb + warp; r < p.B; r += (int64_t)gridDim.x * wpb) {
        const uint64_t cw = p.cw_offset + (uint64_t)r;
        for (int i = lane; i < 2 * NW; i += 32) W[i] = 0u;
        __syncwarp();
        // rate profile: u[info] = msg, 0-bits (+1) elsewhere (polar.py:137-138, pac_code.py:171-172)
        uint4 rnd = make_uint4(0, 0, 0, 0);
        int rnd_block = -1;  // one Philox block serves 128 message bits: a lane's k = lane + 32 i stays in it for 4 turns
        if (!p.msg_in && (K & 3) == 0 && K <= 4096) {
            // generated messages, four bits per lane and turn (the per-bit loop below cost 2/3 of the kernel's
            // instructions at K = 512): bits 4q..4q+3 sit in one word of Philox block q >> 5, which is the same block
            // for the whole warp -- lane b draws block b once, every turn broadcasts its block -- and msg_out / info
            // move as 16-byte vectors
            const uint4 mine = npd_philox4x32_10(make_uint4((uint32_t)cw, (uint32_t)(cw >> 32), (uint32_t)lane, NPD_STREAM_MSG), key);
            for (int q0 = 0; q0 < (K >> 2); q0 += 32) {  // whole-warp turns: the shuffles need every lane
                const int q = q0 + lane, k0 = q << 2, src = q0 >> 5;
                rnd = make_uint4(__shfl_sync(NPD_FULL, mine.x, src), __shfl_sync(NPD_FULL, mine.y, src),
                                 __shfl_sync(NPD_FULL, mine.z, src), __shfl_sync(NPD_FULL, mine.w, src));
                if (q >= (K >> 2)) continue;
                const uint32_t wsel = (k0 >> 5) & 3;
                const uint32_t word = wsel == 0 ? rnd.x : wsel == 1 ? rnd.y : wsel == 2 ? rnd.z : rnd.w;
                const uint32_t nib = (word >> (k0 & 31)) & 0xFu;
                if (p.msg_out)
                    *reinterpret_cast<float4 *>(p.msg_out + r * K + k0) =
                        make_float4((nib & 1u) ? -1.0f : 1.0f, (nib & 2u) ? -1.0f : 1.0f, (nib & 4u) ? -1.0f : 1.0f,
                                    (nib & 8u) ? -1.0f : 1.0f);
                const int4 pos = __ldg(reinterpret_cast<const int4 *>(p.info + k0));
                if (nib & 1u) atomicOr(&V[pos.x >> 5], 1u << (pos.x & 31));
                if (nib & 2u) atomicOr(&V[pos.y >> 5], 1u << (pos.y & 31));
                if (nib & 4u) atomicOr(&V[pos.z >> 5], 1u << (pos.z & 31));
                if (nib & 8u) atomicOr(&V[pos.w >> 5], 1u << (pos.w & 31));
            }
        } else
        for (int k = lane; k < K; k += 32) {
            uint32_t bit;
            if (p.msg_in) {
                bit = p.msg_in[r * K + k] < 0.0f;
            } else {
                if ((k >> 7) != rnd_block) {
                    rnd_block = k >> 7;
                    rnd = npd_philox4x32_10(
                        make_uint4((uint32_t)cw, (uint32_t)(cw >> 32), (uint32_t)rnd_block, NPD_STREAM_MSG), key);
                }
                const uint32_t wsel = (k >> 5) & 3;
                const uint32_t word = wsel == 0 ? rnd.x : wsel == 1 ? rnd.y : wsel == 2 ? rnd.z : rnd.w;
                bit = (word >> (k & 31)) & 1u;
            }
            if (p.msg_out) p.msg_out[r * K + k] = bit ? -1.0f : 1.0f;
            const int pos = __ldg(p.info + k);
            if (bit) atomicOr(&V[pos >> 5], 1u << (pos & 31));
        }
        __syncwarp();
        if (p.ubits_out)
            for (int i = lane; i < NW; i += 32) p.ubits_out[r * NW + i] = V[i];
        if (p.pac_M) {
            // rate-1 convolutional pre-coder: u_i = v_i xor (xor over taps j of v_{i-j})
            for (int i0 = 0; i0 < N; i0 += 32) {
                const int e = i0 + lane;
                uint32_t bit = 0u;
                if (e < N) {
                    bit = (V[e >> 5] >> (e & 31)) & 1u;
                    for (int j = 1; j < p.pac_M; ++j)
                        if (((p.pac_taps >> (j - 1)) & 1u) && e - j >= 0)
                            bit ^= (V[(e - j) >> 5] >> ((e - j) & 31)) & 1u;
                }
                const uint32_t word = __ballot_sync(NPD_FULL, bit);
                if (lane == 0) W[i0 >> 5] = word;
            }
        } else {
            for (int i = lane; i < NW; i += 32) W[i] = V[i];
        }
        __syncwarp();
        for (int i = lane; i < NW; i += 32) W[i] = word_stages(W[i]);
        __syncwarp();
        for (int d = 5; d < n; ++d) {
            const int stride = 1 << (d - 5);
            for (int i = lane; i < NW; i += 32)
                if (!(i & stride)) W[i] ^= W[i + stride];
            __syncwarp();
        }
        // outputs
        if (N >= 4) {
            for (int q = lane; q < (N >> 2); q += 32) {
                const int e = q << 2;
                const uint32_t nib = (W[e >> 5] >> (e & 31)) & 0xFu;
                float4 xv;
                xv.x = (nib & 1u) ? -1.0f : 1.0f;
                xv.y = (nib & 2u) ? -1.0f : 1.0f;
                xv.z = (nib & 4u) ? -1.0f : 1.0f;
                xv.w = (nib & 8u) ? -1.0f : 1.0f;
                if (p.x_out) *reinterpret_cast<float4 *>(p.x_out + r * N + e) = xv;
                if (p.y_out) {
                    uint4 rnd = npd_philox4x32_10(
                        make_uint4((uint32_t)cw, (uint32_t)(cw >> 32), (uint32_t)q, NPD_STREAM_NOISE + p.point), key);
                    float2 z0 = npd_box_muller(rnd.x, rnd.y), z1 = npd_box_muller(rnd.z, rnd.w);
                    float4 yv;
                    yv.x = fmaf(p.sigma, z0.x, xv.x);
                    yv.y = fmaf(p.sigma, z0.y, xv.y);
                    yv.z = fmaf(p.sigma, z1.x, xv.z);
                    yv.w = fmaf(p.sigma, z1.y, xv.w);
                    *reinterpret_cast<float4 *>(p.y_out + r * N + e) = yv;
                }
            }
        } else if (lane == 0) {  // N == 2
            uint4 rnd = npd_philox4x32_10(
                make_uint4((uint32_t)cw, (uint32_t)(cw >> 32), 0u, NPD_STREAM_NOISE + p.point), key);
            float2 z0 = npd_box_muller(rnd.x, rnd.y);
            for (int e = 0; e < N; ++e) {
                float xv = ((W[0] >> e) & 1u) ? -1.0f : 1.0f;
                if (p.x_out) p.x_out[r * N + e] = xv;
                if (p.y_out) p.y_out[r * N + e] = xv + p.sigma * (e == 0 ? z0.x : z0.y);
            }
        }
        __syncwarp();
    }
}

// y = x + sigma * z over an existing x (the drop-in's channel()).  One thread per 4 samples.
__global__ void __launch_bounds__(256) awgn_kernel(const float *__restrict__ x, float *__restrict__ y,
                                                   int64_t B, int N, float sigma, uint64_t seed,
                                                   uint32_t point, uint64_t cw_offset)
{
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const int QN = (N + 3) >> 2;
    const int64_t total = B * QN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / QN;
        const int q = (int)(i - r * QN);
        const uint64_t cw = cw_offset + (uint64_t)r;
        uint4 rnd = npd_philox4x32_10(
            make_uint4((uint32_t)cw, (uint32_t)(cw >> 32), (uint32_t)q, NPD_STREAM_NOISE + point), key);
        float2 z0 = npd_box_muller(rnd.x, rnd.y), z1 = npd_box_muller(rnd.z, rnd.w);
        const int e = q << 2;
        if ((N & 3) == 0) {
            float4 xv = *reinterpret_cast<const float4 *>(x + r * N + e);
            float4 yv;
            yv.x = fmaf(sigma, z0.x, xv.x);
            yv.y = fmaf(sigma, z0.y, xv.y);
            yv.z = fmaf(sigma, z1.x, xv.z);
            yv.w = fmaf(sigma, z1.y, xv.w);
            *reinterpret_cast<float4 *>(y + r * N + e) = yv;
        } else {
            const float zz[4] = {z0.x, z0.y, z1.x, z1.y};
            for (int t = 0; t < 4 && e + t < N; ++t) y[r * N + e + t] = x[r * N + e + t] + sigma * zz[t];
        }
    }
}

// the kernels move rows as 16-byte vectors whenever the row length allows it: base pointers must then be 16-byte aligned
// (torch allocations are 256-byte aligned; a view with an odd storage offset is not)
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int launch_encode(const npd_code *code, const EncParams &p, cudaStream_t st)
{
    NPD_REQUIRE(((p.K & 3) != 0 || aligned16(p.msg_out)) && (code->N < 4 || (aligned16(p.x_out) && aligned16(p.y_out))),
                "npd encoder: output pointers must be 16-byte aligned (msg %p x %p y %p)", (void *)p.msg_out, (void *)p.x_out,
                (void *)p.y_out);
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    const int NW = (code->N + 31) / 32;
    const int wpb = 4;
    const size_t smem = (size_t)wpb * 2 * NW * 4;
    int64_t grid = (p.B + wpb - 1) / wpb;
    const int64_t cap = (int64_t)dp.sm_count * 16;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    encode_kernel<<<(unsigned)grid, 32 * wpb, smem, st>>>(p);
    NPD_CHECK_CUDA(cudaGetLastError());
    return NPD_OK;
}

}  // namespace

NPD_API int npd_polar_encode(const npd_code_t *code, const float *msg, float *x, int64_t B, void *stream)
{
    NPD_REQUIRE(code && msg && x, "npd_polar_encode: null argument");
    NPD_REQUIRE(B >= 0, "npd_polar_encode: negative batch");
    if (B == 0) return NPD_OK;
    EncParams p{};
    p.msg_in = msg; p.x_out = x; p.info = code->d_info; p.B = B; p.n = code->n; p.K = code->K;
    p.pac_taps = code->pac_taps; p.pac_M = code->pac_g ? code->pac_M : 0;
    return launch_encode(code, p, (cudaStream_t)stream);
}

NPD_API int npd_gen_encode_awgn(const npd_code_t *code, float *msg, float *x, float *y, int64_t B,
                                float sigma, uint64_t seed, uint32_t point, uint64_t cw_offset,
                                void *stream)
{
    NPD_REQUIRE(code && y, "npd_gen_encode_awgn: null argument");
    NPD_REQUIRE(B >= 0, "npd_gen_encode_awgn: negative batch");
    if (B == 0) return NPD_OK;
    EncParams p{};
    p.msg_out = msg; p.x_out = x; p.y_out = y; p.info = code->d_info; p.B = B; p.n = code->n;
    p.K = code->K; p.pac_taps = code->pac_taps; p.pac_M = code->pac_g ? code->pac_M : 0;
    p.sigma = sigma; p.seed = seed; p.point = point; p.cw_offset = cw_offset;
    return launch_encode(code, p, (cudaStream_t)stream);
}

// internal (count_sweep.cu): the fused sweep's generator -- y and the packed transmitted u words, no float messages
int npd_gen_encode_awgn_bits(const npd_code *code, uint32_t *ubits, float *y, int64_t B, float sigma, uint64_t seed,
                             uint32_t point, uint64_t cw_offset, cudaStream_t st)
{
    EncParams p{};
    p.ubits_out = ubits; p.y_out = y; p.info = code->d_info; p.B = B; p.n = code->n;
    p.K = code->K; p.pac_taps = code->pac_taps; p.pac_M = code->pac_g ? code->pac_M : 0;
    p.sigma = sigma; p.seed = seed; p.point = point; p.cw_offset = cw_offset;
    return launch_encode(code, p, st);
}

NPD_API int npd_awgn(const float *x, float *y, int64_t B, int N, float sigma, uint64_t seed,
                     uint32_t point, uint64_t cw_offset, void *stream)
{
    NPD_REQUIRE(x && y, "npd_awgn: null argument");
    NPD_REQUIRE(B >= 0 && N >= 1, "npd_awgn: bad shape");
    NPD_REQUIRE((N & 3) != 0 || (aligned16(x) && aligned16(y)), "npd_awgn: x and y must be 16-byte aligned when N %% 4 == 0");
    if (B == 0) return NPD_OK;
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    const int64_t total = B * ((N + 3) / 4);
    int64_t grid = (total + 255) / 256;
    const int64_t cap = (int64_t)dp.sm_count * 8;
    if (grid > cap) grid = cap;
    awgn_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, y, B, N, sigma, seed, point, cw_offset);
    NPD_CHECK_CUDA(cudaGetLastError());
    return NPD_OK;
}
