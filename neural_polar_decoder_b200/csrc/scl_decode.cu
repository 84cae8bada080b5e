// scl_decode.cu -- successive-cancellation LIST decoder, bit-exact fp32 min-sum.
//
// Replaces PolarCode.scl_decode(y, snr, L, use_CRC=False) (reference polar.py:793-876 with pruneLists 777-791, the
// LLR recursion of 369-449 and utils.py:272-275): the stronger classical baseline of the reference's sweeps
// (`--list_size` in rnn_all.py:866-872, hard-wired L = 4 in run_models.py:329).  Semantics kept exactly:
//   * the LLR recursion runs WITHOUT the frozen prior (polar.py:801-803); a frozen leaf pays |L| when
//     sign(L) != +1, stores L + infty and decides +1;
//   * an information leaf doubles the list: the first copies decide sign(L), the second copies -sign(L) and pay |L|
//     (float32 adds in the reference's order: metric + |L|);
//   * a list longer than L is pruned to the L smallest metrics, kept in ascending list-index order;
//   * the final pick re-encodes every surviving path and returns the one whose codeword is closest to y.
// Two torch implementation details are not part of the contract: the order torch.topk gives equal metrics (here: the
// lower list index wins) and the float32 reduction order of the distance sum (here: double precision).  Neither
// matters on real-valued noise; tests/golden/scl.npz pins decisions and leaf LLRs against the live reference.
//
// Mapping: one warp decodes one codeword; its paths are the G = 2^ceil(log2 L) "slots" of a shared-memory state laid
// out exactly like the exact SC group kernel's (sc_decode.cu: LLR tree [element][slot], partial sums as (sign, zero)
// bit planes that are Plotkin-transformed in place), so the f/g arithmetic and the tie handling are the same code
// shape.  Lanes run over (slot, element) pairs; cloning copies a slot in place while the list still fits, pruning
// gathers the survivors into the second state buffer.
#include "npd_common.cuh"

namespace {

struct SclParams {
    const float *y;        // [B,N]
    float *leaf_llr;       // [B,N] or null
    float *decoded;        // [B,K]
    const int32_t *info;   // [K]
    const uint32_t *frozen_words;
    int64_t B;
    int n, K, L;
    float scale, infty;
};

__host__ __device__ inline int scl_plane_stride(int N) { return ((N + 31) >> 5) | 1; }

// one state buffer: tree G*(N-1) floats | leaf history G*N floats | metric G floats | 4 bit planes G*NWP words each
template <int G>
__host__ __device__ inline size_t scl_state_words(int N)
{
    return (size_t)G * (N - 1) + (size_t)G * N + G + (size_t)4 * G * scl_plane_stride(N);
}
template <int G>
__host__ __device__ inline size_t scl_warp_smem_bytes(int N)
{
    return 4 * (2 * scl_state_words<G>(N) + 4 * G);  // two buffers + scratch (parent, type, saved L, pad)
}

template <int G>
struct SclState {
    float *tree, *leaf, *metric;
    uint32_t *PS, *PZ, *US, *UZ;
    __device__ SclState(float *base, int N, int NWP)
    {
        tree = base;
        leaf = tree + (size_t)G * (N - 1);
        metric = leaf + (size_t)G * N;
        PS = reinterpret_cast<uint32_t *>(metric + G);
        PZ = PS + G * NWP;
        US = PZ + G * NWP;
        UZ = US + G * NWP;
    }
};

template <int G>
struct LogG;
template <> struct LogG<1> { static constexpr int v = 0; };
template <> struct LogG<2> { static constexpr int v = 1; };
template <> struct LogG<4> { static constexpr int v = 2; };
template <> struct LogG<8> { static constexpr int v = 3; };
template <> struct LogG<16> { static constexpr int v = 4; };
template <> struct LogG<32> { static constexpr int v = 5; };

template <int G>
__device__ void scl_decode_one(const SclParams &p, float *base, const int64_t cw, const int lane)
{
    constexpr int LG = LogG<G>::v;
    const int n = p.n, N = 1 << n, NWP = scl_plane_stride(N), L = p.L;
    const size_t SW = scl_state_words<G>(N);
    int cur = 0;
    int *s_par = reinterpret_cast<int *>(base + 2 * SW);  // [G] parent slot of every survivor
    int *s_typ = s_par + G;                               // [G] 0 = keeps sign(L), 1 = takes -sign(L)
    float *s_L = reinterpret_cast<float *>(s_typ + G);    // [G] the parent's leaf LLR
    const float *yrow = p.y + cw * N;
    const bool want_leaf = p.leaf_llr != nullptr;

    {
        SclState<G> S(base, N, NWP);
        for (int i = lane; i < 4 * G * NWP; i += 32) S.PS[i] = 0u;
        if (lane < G) S.metric[lane] = 0.0f;
    }
    __syncwarp();
    int np = 1;  // live paths = slots 0 .. np-1
    uint32_t frozen_word = 0u;

    for (int o = 0; o < N; ++o) {
        SclState<G> S(base + cur * SW, N, NWP);
        // ---- LLR recursion down to leaf o for every live path (polar.py:369-449, no priors) ----
        const int top = (o == 0) ? n - 1 : (__ffs(o) - 1);
        for (int lv = top; lv >= 0; --lv) {
            const int h = 1 << lv;
            const bool is_g = (o != 0) && (lv == top);
            const int psbit0 = o - h;
            float *ch = S.tree + (size_t)G * (h - 1);
            const float *par = S.tree + (size_t)G * (2 * h - 1);
            for (int idx = lane; idx < G * h; idx += 32) {
                const int c = idx & (G - 1), e = idx >> LG;
                if (c >= np) continue;
                float a, b;
                if (lv == n - 1) {  // parent = root = scale * y, the same for every path (polar.py:797-798)
                    a = p.scale * __ldg(yrow + e);
                    b = p.scale * __ldg(yrow + e + h);
                } else {
                    a = par[idx];
                    b = par[idx + G * h];
                }
                float r;
                if (is_g) {
                    const int bit = psbit0 + e;
                    const uint32_t s = (S.PS[c * NWP + (bit >> 5)] >> (bit & 31)) & 1u;
                    const uint32_t z = (S.PZ[c * NWP + (bit >> 5)] >> (bit & 31)) & 1u;
                    r = npd_g_trit(s, z, a, b);
                } else {
                    r = npd_f_minsum(a, b);
                }
                ch[idx] = r;
            }
            __syncwarp();
        }

        if ((o & 31) == 0) frozen_word = __ldg(p.frozen_words + (o >> 5));
        const bool frozen = (frozen_word >> (o & 31)) & 1u;
        uint32_t s = 0u, z = 0u;  // this lane's slot decides (sign, zero)
        if (frozen) {
            // polar.py:812-826: pay |L| unless sign(L) == +1, store L + infty, decide +1
            if (lane < np) {
                const float Lv = S.tree[lane];
                S.metric[lane] = S.metric[lane] + ((Lv > 0.0f) ? 0.0f : fabsf(Lv));
                if (want_leaf) S.leaf[(size_t)o * G + lane] = Lv + p.infty;
            }
        } else if (2 * np <= L) {
            // ---- the doubled list still fits: clone slot c into slot np + c in place (polar.py:832-846) ----
            for (int idx = lane; idx < G * (N - 1); idx += 32) {
                const int c = idx & (G - 1);
                if (c >= np && c < 2 * np) S.tree[idx] = S.tree[idx - np];
            }
            if (want_leaf)
                for (int idx = lane; idx < G * o; idx += 32) {
                    const int c = idx & (G - 1);
                    if (c >= np && c < 2 * np) S.leaf[idx] = S.leaf[idx - np];
                }
            for (int idx = lane; idx < 4 * G * NWP; idx += 32) {
                const int c = (idx / NWP) & (G - 1);
                if (c >= np && c < 2 * np) S.PS[idx] = S.PS[idx - np * NWP];
            }
            __syncwarp();
            if (lane < 2 * np) {
                const int parent = lane < np ? lane : lane - np;
                const float Lv = S.tree[parent];
                z = (Lv == 0.0f);
                s = lane < np ? (Lv < 0.0f) : (Lv > 0.0f);
                if (lane >= np) S.metric[lane] = S.metric[parent] + fabsf(Lv);
                if (want_leaf) S.leaf[(size_t)o * G + lane] = Lv;
            }
            np *= 2;
        } else {
            // ---- 2 np candidates, keep the L smallest metrics in ascending candidate order (pruneLists) ----
            // candidate c < np: path c keeping sign(L); candidate np + c: path c taking -sign(L), metric + |L|
            const float Lv = lane < np ? S.tree[lane] : 0.0f;
            const float mA = lane < np ? S.metric[lane] : 0.0f;
            const float mB = mA + fabsf(Lv);
            int rankA = 0, rankB = 0;
            for (int j = 0; j < np; ++j) {
                const float a = __shfl_sync(NPD_FULL, mA, j), b = __shfl_sync(NPD_FULL, mB, j);
                rankA += (a < mA) || (a == mA && j < lane);
                rankA += (b < mA);                          // candidate np + j comes after candidate lane
                rankB += (a < mB) || (a == mB);             // candidate j comes before candidate np + lane
                rankB += (b < mB) || (b == mB && j < lane);
            }
            const bool keepA = lane < np && rankA < L, keepB = lane < np && rankB < L;
            const uint32_t maskA = __ballot_sync(NPD_FULL, keepA), maskB = __ballot_sync(NPD_FULL, keepB);
            const uint32_t below = (1u << lane) - 1u;
            SclState<G> D(base + (cur ^ 1) * SW, N, NWP);
            if (keepA) {
                const int j = __popc(maskA & below);
                s_par[j] = lane; s_typ[j] = 0; s_L[j] = Lv;
                D.metric[j] = mA;
            }
            if (keepB) {
                const int j = __popc(maskA) + __popc(maskB & below);
                s_par[j] = lane; s_typ[j] = 1; s_L[j] = Lv;
                D.metric[j] = mB;
            }
            const int nnew = __popc(maskA) + __popc(maskB);  // = min(L, 2 np)
            __syncwarp();
            // gather the survivors' state into the other buffer
            for (int idx = lane; idx < G * (N - 1); idx += 32) {
                const int j = idx & (G - 1);
                if (j < nnew) D.tree[idx] = S.tree[idx - j + s_par[j]];
            }
            if (want_leaf)
                for (int idx = lane; idx < G * o; idx += 32) {
                    const int j = idx & (G - 1);
                    if (j < nnew) D.leaf[idx] = S.leaf[idx - j + s_par[j]];
                }
            for (int idx = lane; idx < 4 * G * NWP; idx += 32) {
                const int j = (idx / NWP) & (G - 1);
                if (j < nnew) D.PS[idx] = S.PS[idx + (s_par[j] - j) * NWP];
            }
            __syncwarp();
            cur ^= 1;
            np = nnew;
            if (lane < np) {
                const float Lp = s_L[lane];
                z = (Lp == 0.0f);
                s = s_typ[lane] ? (Lp > 0.0f) : (Lp < 0.0f);
                if (want_leaf) D.leaf[(size_t)o * G + lane] = Lp;
            }
            __syncwarp();
        }

        // ---- record the decision and merge the transformed partial sums (same bit algebra as the SC group kernel) ----
        SclState<G> T(base + cur * SW, N, NWP);
        const int w = o >> 5, bpos = o & 31;
        if (lane < np) {
            const int c = lane;
            T.US[c * NWP + w] |= s << bpos;
            T.UZ[c * NWP + w] |= z << bpos;
            uint32_t ps = T.PS[c * NWP + w] | (s << bpos);
            uint32_t pz = T.PZ[c * NWP + w] | (z << bpos);
            const int m = __ffs(~o) - 1;  // trailing ones of o = number of completed merges
            const int mi = min(m, min(n, 5));
#pragma unroll 1
            for (int j = 0; j < mi; ++j) {
                const int hb = 1 << j;
                const uint32_t mask = ((1u << hb) - 1u) << ((o + 1 - 2 * hb) & 31);
                ps ^= (ps >> hb) & mask;
                pz |= (pz >> hb) & mask;
            }
            T.PS[c * NWP + w] = ps;
            T.PZ[c * NWP + w] = pz;
        }
        __syncwarp();
        {
            const int m = min(__ffs(~o) - 1, n);
            for (int j = 5; j < m; ++j) {
                const int nw = 1 << (j - 5);
                const int wl = (o + 1 - 2 * (1 << j)) >> 5;
                for (int idx = lane; idx < G * nw; idx += 32) {
                    const int c = idx >> (j - 5), i = idx & (nw - 1);
                    if (c < np) {
                        T.PS[c * NWP + wl + i] ^= T.PS[c * NWP + wl + nw + i];
                        T.PZ[c * NWP + wl + i] |= T.PZ[c * NWP + wl + nw + i];
                    }
                }
                __syncwarp();
            }
        }
    }

    // ---- ML pick among the list (polar.py:869-874): after the last leaf PS / PZ are the re-encoded codewords ----
    SclState<G> S(base + cur * SW, N, NWP);
    int best = 0;
    double best_d = 0.0;
    for (int c = 0; c < np; ++c) {
        double d = 0.0;
        for (int e = lane; e < N; e += 32) {
            const uint32_t sb = (S.PS[c * NWP + (e >> 5)] >> (e & 31)) & 1u;
            const uint32_t zb = (S.PZ[c * NWP + (e >> 5)] >> (e & 31)) & 1u;
            const float x = zb ? 0.0f : (sb ? -1.0f : 1.0f);
            const float df = x - __ldg(yrow + e);
            d += (double)(df * df);
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) d += __shfl_xor_sync(NPD_FULL, d, sft);
        if (c == 0 || d < best_d) { best = c; best_d = d; }
    }
    float *dst = p.decoded + cw * p.K;
    for (int k = lane; k < p.K; k += 32) {
        const int pos = __ldg(p.info + k);
        const uint32_t sb = (S.US[best * NWP + (pos >> 5)] >> (pos & 31)) & 1u;
        const uint32_t zb = (S.UZ[best * NWP + (pos >> 5)] >> (pos & 31)) & 1u;
        dst[k] = zb ? 0.0f : (sb ? -1.0f : 1.0f);
    }
    if (want_leaf)
        for (int e = lane; e < N; e += 32) p.leaf_llr[cw * N + e] = S.leaf[(size_t)e * G + best];
    __syncwarp();
}

template <int G>
__global__ void __launch_bounds__(128) scl_kernel(const SclParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int N = 1 << p.n;
    float *base = reinterpret_cast<float *>(smem_raw + (size_t)warp * scl_warp_smem_bytes<G>(N));
    for (int64_t cw = (int64_t)blockIdx.x * wpb + warp; cw < p.B; cw += (int64_t)gridDim.x * wpb)
        scl_decode_one<G>(p, base, cw, lane);
}

template <int G>
int launch_scl(const npd_code *code, const SclParams &p, cudaStream_t st)
{
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    const size_t per_warp = scl_warp_smem_bytes<G>(code->N);
    if (per_warp > (size_t)dp.smem_optin) {
        npd_set_error("npd_scl_decode: N=%d with list size %d needs %zu B of shared memory per codeword (limit %d)", code->N,
                      p.L, per_warp, dp.smem_optin);
        return NPD_EUNSUPPORTED;
    }
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > (size_t)dp.smem_optin) wpb >>= 1;
    int blocks_per_sm = (int)((size_t)(220 * 1024) / (per_warp * wpb + 1024));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    if (blocks_per_sm * wpb > 32) blocks_per_sm = 32 / wpb;
    int64_t grid = (int64_t)dp.sm_count * blocks_per_sm;
    const int64_t need = (p.B + wpb - 1) / wpb;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    auto kern = scl_kernel<G>;
    NPD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpb)));
    kern<<<(unsigned)grid, 32 * wpb, per_warp * wpb, st>>>(p);
    NPD_CHECK_CUDA(cudaGetLastError());
    return NPD_OK;
}

}  // namespace

NPD_API int npd_scl_decode(const npd_code_t *code, const float *y, float llr_scale, int list_size, float *leaf_llr,
                           float *decoded, int64_t B, void *stream)
{
    NPD_REQUIRE(code && y && (decoded || code->K == 0), "npd_scl_decode: null argument");
    NPD_REQUIRE(B >= 0, "npd_scl_decode: negative batch");
    NPD_REQUIRE(code->pac_g == 0, "npd_scl_decode: polar code objects only");
    NPD_REQUIRE(list_size >= 1, "npd_scl_decode: list size %d < 1", list_size);
    if (list_size > 32) {
        npd_set_error("npd_scl_decode: list size %d > 32 is outside the implemented envelope", list_size);
        return NPD_EUNSUPPORTED;
    }
    if (B == 0) return NPD_OK;
    SclParams p{};
    p.y = y; p.leaf_llr = leaf_llr; p.decoded = decoded; p.info = code->d_info; p.frozen_words = code->d_frozen_words;
    p.B = B; p.n = code->n; p.K = code->K; p.L = list_size; p.scale = llr_scale; p.infty = code->infty;
    cudaStream_t st = (cudaStream_t)stream;
    if (list_size <= 1) return launch_scl<1>(code, p, st);
    if (list_size <= 2) return launch_scl<2>(code, p, st);
    if (list_size <= 4) return launch_scl<4>(code, p, st);
    if (list_size <= 8) return launch_scl<8>(code, p, st);
    if (list_size <= 16) return launch_scl<16>(code, p, st);
    return launch_scl<32>(code, p, st);
}
