// sc_decode.cu -- bit-exact fp32 min-sum successive-cancellation decoders (polar + PAC).
//
// Replaces PolarCode.sc_decode_new (reference polar.py:465-484 with 361-463, utils.py:272-275) and
// PAC.pac_sc_decode (pac_code.py:534-573).  The reference re-encodes the partial sums from scratch
// after every bit (O(N^2) torch.cat calls); here they are kept as two bit planes (sign, is-zero) that
// are Plotkin-transformed in place, which is exact because products of {-1,0,+1} floats are exact.
//
// Mapping ("group" kernel): one warp decodes G codewords in lockstep (the SC schedule is data
// independent).  Every LLR array of the tree lives in shared memory interleaved as [element][G], so
// a min-sum update over a child array of h elements is G*h independent lane-items with conflict-free
// consecutive addresses -- lanes stay busy down to h = 32/G.  The root LLRs are never staged: the two
// top-level updates read y straight from global memory (coalesced for N >= 64).  G is chosen per N so
// that 4*G*(N-1) bytes of tree per warp leave >= 6 warps resident per SM ("SC decoder").
#include "npd_common.cuh"

namespace {

struct ScParams {
    const float *y;        // [B,N]
    const float *use_gt;   // [B,N] or null
    float *leaf_llr;       // [B,N] or null
    float *decoded;        // [B,K]   (polar: u_hat[:,info]; PAC: v_hat[:,info])
    float *u_hat;          // [B,N] or null (PAC only)
    const int32_t *info;   // [K]
    const uint32_t *frozen_words;
    int64_t B;
    int n, K;
    float scale, infty;
    uint32_t pac_taps, pac_state_mask;
};

template <int G>
struct Log2;
template <> struct Log2<1> { static constexpr int v = 0; };
template <> struct Log2<2> { static constexpr int v = 1; };
template <> struct Log2<4> { static constexpr int v = 2; };
template <> struct Log2<8> { static constexpr int v = 3; };
template <> struct Log2<16> { static constexpr int v = 4; };
template <> struct Log2<32> { static constexpr int v = 5; };

__host__ __device__ inline int plane_stride(int N) { return ((N + 31) >> 5) | 1; }  // odd: no bank clash

template <int G, bool PAC>
__host__ __device__ inline size_t sc_warp_smem_bytes(int N)
{
    const int planes = PAC ? 6 : 4;
    return (size_t)4 * G * (N - 1) + (size_t)4 * planes * G * plane_stride(N);
}

template <int G, bool PAC>
__global__ void __launch_bounds__(64) sc_group_kernel(const ScParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LG = Log2<G>::v;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int n = p.n;
    const int N = 1 << n;
    const int NW = (N + 31) >> 5;
    const int NWP = plane_stride(N);

    unsigned char *base = smem_raw + (size_t)warp * sc_warp_smem_bytes<G, PAC>(N);
    float *buf = reinterpret_cast<float *>(base);              // level lv at G*(2^lv - 1)
    uint32_t *PS = reinterpret_cast<uint32_t *>(buf + (size_t)G * (N - 1));  // transformed sums: sign
    uint32_t *PZ = PS + G * NWP;                                             // transformed sums: zero
    uint32_t *US = PZ + G * NWP;                                             // raw decisions: sign
    uint32_t *UZ = US + G * NWP;                                             // raw decisions: zero
    uint32_t *VS = UZ + G * NWP;                                             // PAC: v_hat sign
    uint32_t *VZ = VS + G * NWP;                                             // PAC: v_hat undecided

    const int64_t ngroups = (p.B + G - 1) / G;
    for (int64_t grp = (int64_t)blockIdx.x * wpb + warp; grp < ngroups;
         grp += (int64_t)gridDim.x * wpb) {
        const int64_t cw0 = grp * G;
        const int nvalid = (int)min((int64_t)G, p.B - cw0);

        for (int i = lane; i < (PAC ? 6 : 4) * G * NWP; i += 32) PS[i] = 0u;
        __syncwarp();

        uint32_t pac_state = 0u;  // lane c < G: shift register of codeword c, bit j = state[j] is -1
        uint32_t frozen_word = 0u;

        for (int o = 0; o < N; ++o) {
            const int top = (o == 0) ? n - 1 : (__ffs(o) - 1);
            for (int lv = top; lv >= 0; --lv) {
                const int h = 1 << lv;
                const bool is_g = (o != 0) && (lv == top);
                const int psbit0 = o - h;  // first partial-sum bit used by g
                float *ch = buf + (size_t)G * (h - 1);
                if (lv == n - 1) {
                    // parent = root = scale * y, read from global (polar.py:468-469)
                    if (h >= 32) {
                        for (int c = 0; c < G; ++c) {
                            const bool ok = c < nvalid;
                            const float *row = p.y + (cw0 + (ok ? c : 0)) * N;
                            for (int e = lane; e < h; e += 32) {
                                float a = ok ? p.scale * __ldg(row + e) : 0.f;
                                float b = ok ? p.scale * __ldg(row + e + h) : 0.f;
                                float r;
                                if (is_g) {
                                    const int bit = psbit0 + e;
                                    uint32_t s = (PS[c * NWP + (bit >> 5)] >> (bit & 31)) & 1u;
                                    uint32_t z = (PZ[c * NWP + (bit >> 5)] >> (bit & 31)) & 1u;
                                    r = npd_g_trit(s, z, a, b);
                                } else {
                                    r = npd_f_minsum(a, b);
                                }
                                ch[e * G + c] = r;
                            }
                        }
                    } else {
                        for (int idx = lane; idx < G * h; idx += 32) {
                            const int c = idx & (G - 1), e = idx >> LG;
                            const bool ok = c < nvalid;
                            const float *row = p.y + (cw0 + (ok ? c : 0)) * N;
                            float a = ok ? p.scale * __ldg(row + e) : 0.f;
                            float b = ok ? p.scale * __ldg(row + e + h) : 0.f;
                            float r;
                            if (is_g) {
                                const int bit = psbit0 + e;
                                uint32_t s = (PS[c * NWP + (bit >> 5)] >> (bit & 31)) & 1u;
                                uint32_t z = (PZ[c * NWP + (bit >> 5)] >> (bit & 31)) & 1u;
                                r = npd_g_trit(s, z, a, b);
                            } else {
                                r = npd_f_minsum(a, b);
                            }
                            ch[idx] = r;
                        }
                    }
                } else {
                    const float *par = buf + (size_t)G * (2 * h - 1);
                    if (is_g) {
                        for (int idx = lane; idx < G * h; idx += 32) {
                            const int c = idx & (G - 1), e = idx >> LG;
                            const int bit = psbit0 + e;
                            uint32_t s = (PS[c * NWP + (bit >> 5)] >> (bit & 31)) & 1u;
                            uint32_t z = (PZ[c * NWP + (bit >> 5)] >> (bit & 31)) & 1u;
                            ch[idx] = npd_g_trit(s, z, par[idx], par[idx + G * h]);
                        }
                    } else {
                        for (int idx = lane; idx < G * h; idx += 32)
                            ch[idx] = npd_f_minsum(par[idx], par[idx + G * h]);
                    }
                }
                __syncwarp();
            }

            // ---- leaf o: decision (polar.py:471-481 / pac_code.py:543-568) ----
            if ((o & 31) == 0) frozen_word = __ldg(p.frozen_words + (o >> 5));
            const bool frozen = (frozen_word >> (o & 31)) & 1u;
            const int w = o >> 5, bpos = o & 31;
            if (lane < G) {
                const int c = lane;
                const bool ok = c < nvalid;
                float L = buf[c];
                if (!PAC) L = L + (frozen ? p.infty : 0.0f);
                if (p.leaf_llr && ok) p.leaf_llr[(cw0 + c) * N + o] = L;
                uint32_t s = L < 0.0f, z = (L == 0.0f);
                bool have_gt = false;
                if (p.use_gt && ok) {
                    float t = p.use_gt[(cw0 + c) * N + o];
                    s = t < 0.0f;
                    z = (t == 0.0f);
                    have_gt = true;
                }
                if (PAC) {
                    const uint32_t par_bit = __popc(pac_state & p.pac_taps) & 1u;
                    uint32_t vs = 0u, vz = 0u;
                    if (frozen) {
                        if (!have_gt) {  // u = conv(+1, state), state <- shift in +1
                            s = par_bit;
                            z = 0u;
                            pac_state = (pac_state << 1) & p.pac_state_mask;
                        }
                    } else {
                        if (z) {
                            vz = 1u;  // neither branch matches a 0: v stays 0, state unchanged
                        } else {
                            vs = s ^ par_bit;
                            pac_state = ((pac_state << 1) | vs) & p.pac_state_mask;
                        }
                    }
                    VS[c * NWP + w] |= vs << bpos;
                    VZ[c * NWP + w] |= vz << bpos;
                }
                US[c * NWP + w] |= s << bpos;
                UZ[c * NWP + w] |= z << bpos;
                // in-word merges of the transformed partial sums (blocks < 32 bits)
                uint32_t ps = PS[c * NWP + w] | (s << bpos);
                uint32_t pz = PZ[c * NWP + w] | (z << bpos);
                const int m = __ffs(~o) - 1;  // trailing ones of o = number of completed merges
                const int mi = min(m, min(n, 5));
#pragma unroll 1
                for (int j = 0; j < mi; ++j) {
                    const int hb = 1 << j;
                    const uint32_t mask = ((1u << hb) - 1u) << ((o + 1 - 2 * hb) & 31);
                    ps ^= (ps >> hb) & mask;
                    pz |= (pz >> hb) & mask;
                }
                PS[c * NWP + w] = ps;
                PZ[c * NWP + w] = pz;
            }
            __syncwarp();
            // word-level merges (blocks >= 32 bits)
            {
                const int m = min(__ffs(~o) - 1, n);
                for (int j = 5; j < m; ++j) {
                    const int nw = 1 << (j - 5);
                    const int wl = (o + 1 - 2 * (1 << j)) >> 5;
                    for (int idx = lane; idx < G * nw; idx += 32) {
                        const int c = idx / nw, i = idx - c * nw;
                        PS[c * NWP + wl + i] ^= PS[c * NWP + wl + nw + i];
                        PZ[c * NWP + wl + i] |= PZ[c * NWP + wl + nw + i];
                    }
                    __syncwarp();
                }
            }
        }

        // ---- outputs ----
        const uint32_t *OS = PAC ? VS : US;
        const uint32_t *OZ = PAC ? VZ : UZ;
        for (int c = 0; c < nvalid; ++c) {
            float *dst = p.decoded + (cw0 + c) * p.K;
            for (int k = lane; k < p.K; k += 32) {
                const int pos = __ldg(p.info + k);
                uint32_t s = (OS[c * NWP + (pos >> 5)] >> (pos & 31)) & 1u;
                uint32_t z = (OZ[c * NWP + (pos >> 5)] >> (pos & 31)) & 1u;
                dst[k] = z ? 0.0f : (s ? -1.0f : 1.0f);
            }
            if (PAC && p.u_hat) {
                float *du = p.u_hat + (cw0 + c) * N;
                for (int e = lane; e < N; e += 32) {
                    uint32_t s = (US[c * NWP + (e >> 5)] >> (e & 31)) & 1u;
                    uint32_t z = (UZ[c * NWP + (e >> 5)] >> (e & 31)) & 1u;
                    du[e] = z ? 0.0f : (s ? -1.0f : 1.0f);
                }
            }
        }
        __syncwarp();
    }
    (void)NW;
}

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

template <int G, bool PAC>
int launch_group(const npd_code *code, const ScParams &p, cudaStream_t st)
{
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    const int N = code->N;
    const size_t per_warp = sc_warp_smem_bytes<G, PAC>(N);
    const size_t budget = (size_t)dp.smem_optin;
    if (per_warp + 1024 > budget) {
        npd_set_error("SC: N=%d needs %zu B of shared memory per warp at G=%d", N, per_warp, G);
        return NPD_EUNSUPPORTED;
    }
    // per-SM shared memory is 228 KB with 1 KB reserved per resident block
    int warps_per_sm = (int)((size_t)(228 * 1024) / (per_warp + 1024));
    if (warps_per_sm > 32) warps_per_sm = 32;
    if (warps_per_sm < 1) warps_per_sm = 1;
    int wpb = 1;
    if (warps_per_sm >= 16 && per_warp * 2 + 1024 <= budget) wpb = 2;
    const int forced_wpb = env_int("NPD_SC_WPB", 0);
    if (forced_wpb == 1 || forced_wpb == 2) wpb = forced_wpb;
    int blocks_per_sm = (int)((size_t)(228 * 1024) / (per_warp * wpb + 1024));
    if (blocks_per_sm > 32) blocks_per_sm = 32;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    const int64_t ngroups = (p.B + G - 1) / G;
    int64_t grid = (int64_t)dp.sm_count * blocks_per_sm;
    const int64_t need = (ngroups + wpb - 1) / wpb;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    auto kern = sc_group_kernel<G, PAC>;
    NPD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(per_warp * wpb)));
    kern<<<(unsigned)grid, 32 * wpb, per_warp * wpb, st>>>(p);
    NPD_CHECK_CUDA(cudaGetLastError());
    return NPD_OK;
}

int default_group(int n)
{
    // tree bytes per warp = 4*G*(N-1); keep it <= ~33 KB so >= 6 warps stay resident per SM
    if (n <= 8) return 32;
    if (n == 9) return 16;
    if (n == 10) return 8;
    if (n == 11) return 4;
    return 2;
}

template <bool PAC>
int dispatch(const npd_code *code, const ScParams &p, cudaStream_t st)
{
    int G = env_int("NPD_SC_G", 0);
    if (G != 1 && G != 2 && G != 4 && G != 8 && G != 16 && G != 32) G = default_group(code->n);
    switch (G) {
    case 1: return launch_group<1, PAC>(code, p, st);
    case 2: return launch_group<2, PAC>(code, p, st);
    case 4: return launch_group<4, PAC>(code, p, st);
    case 8: return launch_group<8, PAC>(code, p, st);
    case 16: return launch_group<16, PAC>(code, p, st);
    default: return launch_group<32, PAC>(code, p, st);
    }
}

}  // namespace

NPD_API int npd_sc_decode(const npd_code_t *code, const float *y, float llr_scale,
                          const float *use_gt, float *leaf_llr, float *decoded, int64_t B,
                          void *stream)
{
    NPD_REQUIRE(code && y && decoded, "npd_sc_decode: null argument");
    NPD_REQUIRE(B >= 0, "npd_sc_decode: negative batch");
    NPD_REQUIRE(code->pac_g == 0, "npd_sc_decode: PAC code object; use npd_pac_sc_decode");
    if (B == 0) return NPD_OK;
    ScParams p{};
    p.y = y; p.use_gt = use_gt; p.leaf_llr = leaf_llr; p.decoded = decoded; p.u_hat = nullptr;
    p.info = code->d_info; p.frozen_words = code->d_frozen_words;
    p.B = B; p.n = code->n; p.K = code->K; p.scale = llr_scale; p.infty = code->infty;
    return dispatch<false>(code, p, (cudaStream_t)stream);
}

NPD_API int npd_pac_sc_decode(const npd_code_t *code, const float *y, float llr_scale,
                              const float *use_gt_codeword, float *leaf_llr, float *v_hat,
                              float *u_hat, int64_t B, void *stream)
{
    NPD_REQUIRE(code && y && v_hat, "npd_pac_sc_decode: null argument");
    NPD_REQUIRE(B >= 0, "npd_pac_sc_decode: negative batch");
    NPD_REQUIRE(code->pac_g != 0, "npd_pac_sc_decode: not a PAC code object");
    if (B == 0) return NPD_OK;
    ScParams p{};
    p.y = y; p.use_gt = use_gt_codeword; p.leaf_llr = leaf_llr; p.decoded = v_hat; p.u_hat = u_hat;
    p.info = code->d_info; p.frozen_words = code->d_frozen_words;
    p.B = B; p.n = code->n; p.K = code->K; p.scale = llr_scale; p.infty = 0.0f;
    p.pac_taps = code->pac_taps;
    p.pac_state_mask = (code->pac_M - 1 >= 32) ? 0xffffffffu : ((1u << (code->pac_M - 1)) - 1u);
    return dispatch<true>(code, p, (cudaStream_t)stream);
}
