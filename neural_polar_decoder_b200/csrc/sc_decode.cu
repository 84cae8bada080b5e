// sc_decode.cu -- bit-exact fp32 min-sum successive-cancellation decoders (polar + PAC).
//
// Replaces PolarCode.sc_decode_new (reference polar.py:465-484 with 361-463, utils.py:272-275) and
// PAC.pac_sc_decode (pac_code.py:534-573).  The reference re-encodes the partial sums from scratch
// after every bit (O(N^2) torch.cat calls); here they are bit planes that are Plotkin-transformed in
// place, which is exact because products of {-1,0,+1} floats are exact.
//
// Two kernels (DESIGN.md "SC decoder"):
//
//  * sc_lane_kernel -- the fast path.  One warp decodes 32 codewords, lane = codeword (the SC schedule
//    is data independent, so the warp never diverges).  The bottom 32-leaf subtree of the LLR tree is
//    fully unrolled and lives in registers (63 floats); the levels above it live in shared memory as
//    [element][33] (conflict-free for lane = codeword AND for the lane = element transposed stores);
//    the topmost levels are never stored: they are recomputed from y in global memory (coalesced,
//    lane = element) whenever the first stored level is refreshed.  Partial sums and decisions are one
//    sign bit per leaf.  A leaf LLR that is exactly 0 (sign(0) = 0, which the reference propagates as a
//    zero partial sum) cannot be represented in one bit: the codeword is flagged by a NaN sentinel in
//    decoded[cw][0] and re-decoded by
//
//  * sc_group_kernel -- the exact tie path (also the general path for K = 0).  One warp decodes G
//    codewords with the whole tree in shared memory as [element][G] and two bit planes per quantity
//    (sign, is-zero), so û in {-1,0,+1} is carried exactly.  In scan mode it visits only the flagged
//    codewords.  Ties need an exact fp32 cancellation; they do not occur on real-valued noise.
#include "npd_common.cuh"
#include <mutex>

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
    int scan_flagged;      // group kernel: 1 = only codewords whose decoded[cw][0] is the NaN sentinel
    int slog;              // lane kernel: highest stored level
    int vec_out;           // quad kernel: 16-byte output stores (K % 4 == 0, aligned `decoded`)
    float *scratch;        // quad kernel, N >= 2048: level n-2 of every resident warp ([warp][element][8 codewords])
    // fused error counting (the Monte-Carlo sweep): instead of writing decisions, the quad kernel compares its u-domain
    // decision words with the generator's packed u words on the info positions and accumulates the counters
    const uint32_t *ubits;       // [B, N/32] transmitted u (message bits at the info positions, 0 elsewhere) or null
    const uint32_t *info_words;  // [N/32] bit i = position i carries information
    unsigned long long *counts;  // [2] bit errors, block errors (accumulated) -- non-null selects the fused mode
    unsigned char *flags;        // [B] fused / packed mode: 1 = codeword needs the exact re-decode (replaces the NaN sentinel)
    // packed mode (sub-block decodes of the split large-N path): decision words (u domain) and the block's encoded
    // partial sums (x domain) go out as bit words at [cw * out_words + word_off + q]; no float outputs
    uint32_t *us_out, *xs_out;
    int out_words, word_off;
    long long *trace;      // bench-only (NPD_SC_TRACE): cycles of warp 0 / block 0's second group: top, levels, block, merge, output, total
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

__device__ __forceinline__ float trit_value(uint32_t s, uint32_t z)
{
    return z ? 0.0f : (s ? -1.0f : 1.0f);
}

// =================================================================================================
// exact tie path: G codewords per warp, whole tree in shared memory, (sign, zero) planes
// =================================================================================================
template <int G, bool PAC>
__host__ __device__ inline size_t sc_warp_smem_bytes(int N)
{
    const int planes = PAC ? 6 : 4;
    return (size_t)4 * G * (N - 1) + (size_t)4 * planes * G * plane_stride(N);
}

template <int G, bool PAC>
__device__ void sc_group_decode(const ScParams &p, unsigned char *base, const int64_t my_cw, const int lane)
{
    // my_cw: lanes c < G hold the codeword index of slot c (-1 = empty slot)
    constexpr int LG = Log2<G>::v;
    const int n = p.n;
    const int N = 1 << n;
    const int NWP = plane_stride(N);
    float *buf = reinterpret_cast<float *>(base);                            // level lv at G*(2^lv - 1)
    uint32_t *PS = reinterpret_cast<uint32_t *>(buf + (size_t)G * (N - 1));  // transformed sums: sign
    uint32_t *PZ = PS + G * NWP;                                             // transformed sums: zero
    uint32_t *US = PZ + G * NWP;                                             // raw decisions: sign
    uint32_t *UZ = US + G * NWP;                                             // raw decisions: zero
    uint32_t *VS = UZ + G * NWP;                                             // PAC: v_hat sign
    uint32_t *VZ = VS + G * NWP;                                             // PAC: v_hat undecided

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
                        const int64_t cw = __shfl_sync(NPD_FULL, my_cw, c);
                        const bool ok = cw >= 0;
                        const float *row = p.y + (ok ? cw : 0) * N;
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
                    for (int idx0 = 0; idx0 < G * h; idx0 += 32) {
                        const int idx = idx0 + lane;
                        const int c = idx & (G - 1), e = idx >> LG;
                        const int64_t cw = __shfl_sync(NPD_FULL, my_cw, c);
                        if (idx < G * h) {
                            const bool ok = cw >= 0;
                            const float *row = p.y + (ok ? cw : 0) * N;
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
            const bool ok = my_cw >= 0;
            float L = buf[c];
            if (!PAC) L = L + (frozen ? p.infty : 0.0f);
            if (p.leaf_llr && ok) p.leaf_llr[my_cw * N + o] = L;
            uint32_t s = L < 0.0f, z = (L == 0.0f);
            bool have_gt = false;
            if (p.use_gt && ok) {
                float t = p.use_gt[my_cw * N + o];
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
                    const int c = idx >> (j - 5), i = idx & (nw - 1);
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
    for (int c = 0; c < G; ++c) {
        const int64_t cw = __shfl_sync(NPD_FULL, my_cw, c);
        if (cw < 0) continue;
        float *dst = p.decoded + cw * p.K;
        for (int k = lane; k < p.K; k += 32) {
            const int pos = __ldg(p.info + k);
            uint32_t s = (OS[c * NWP + (pos >> 5)] >> (pos & 31)) & 1u;
            uint32_t z = (OZ[c * NWP + (pos >> 5)] >> (pos & 31)) & 1u;
            dst[k] = trit_value(s, z);
        }
        if (PAC && p.u_hat) {
            float *du = p.u_hat + cw * N;
            for (int e = lane; e < N; e += 32) {
                uint32_t s = (US[c * NWP + (e >> 5)] >> (e & 31)) & 1u;
                uint32_t z = (UZ[c * NWP + (e >> 5)] >> (e & 31)) & 1u;
                du[e] = trit_value(s, z);
            }
        }
    }
    __syncwarp();
}

template <int G, bool PAC>
__global__ void __launch_bounds__(64) sc_group_kernel(const ScParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int N = 1 << p.n;
    unsigned char *base = smem_raw + (size_t)warp * sc_warp_smem_bytes<G, PAC>(N);

    if (!p.scan_flagged) {
        const int64_t ngroups = (p.B + G - 1) / G;
        for (int64_t grp = (int64_t)blockIdx.x * wpb + warp; grp < ngroups; grp += (int64_t)gridDim.x * wpb) {
            const int64_t cw = grp * G + lane;
            sc_group_decode<G, PAC>(p, base, (lane < G && cw < p.B) ? cw : -1, lane);
        }
    } else {
        // visit chunks of 32 codewords; re-decode those flagged by the lane kernel's NaN sentinel
        const int64_t nchunks = (p.B + 31) / 32;
        for (int64_t ch = (int64_t)blockIdx.x * wpb + warp; ch < nchunks; ch += (int64_t)gridDim.x * wpb) {
            const int64_t cw = ch * 32 + lane;
            bool flag = false;
            if (cw < p.B) {
                if (p.flags) {
                    flag = p.flags[cw] != 0;
                } else {
                    const float v = p.decoded[cw * p.K];
                    flag = (v != v);
                }
            }
            uint32_t mask = __ballot_sync(NPD_FULL, flag);
            while (mask) {
                // slot c <- c-th flagged lane
                int64_t mine = -1;
                if (lane < G) {
                    const uint32_t src = __fns(mask, 0, lane + 1);
                    if (src < 32u) mine = ch * 32 + src;
                }
                sc_group_decode<G, PAC>(p, base, mine, lane);
                for (int i = 0; i < G && mask; ++i) mask &= mask - 1;  // drop the G lowest set bits
            }
        }
    }
}

// =================================================================================================
// fast path: lane = codeword
// =================================================================================================
template <int BLOG, bool PAC, bool EXTRAS>
struct LaneCtx {
    uint32_t ps;      // transformed partial sums of the current block (bit i = leaf i of the block)
    uint32_t us;      // raw decisions of the block
    uint32_t vs;      // PAC: v_hat sign bits of the block
    uint32_t frozen;  // frozen mask of the block
    uint32_t gt_s;    // EXTRAS: genie bits of the block
    uint32_t tie;     // a leaf LLR (or genie value) was exactly zero somewhere in this codeword
    uint32_t pac_state;
    uint32_t pac_taps, pac_state_mask;
    float infty;
    float *llr_out;   // EXTRAS: leaf_llr + cw*N + block offset, or null
    bool have_gt;
};

template <int O, int BLOG, bool PAC, bool EXTRAS>
__device__ __forceinline__ void lane_leaf(float L, LaneCtx<BLOG, PAC, EXTRAS> &c)
{
    const bool frozen = (c.frozen >> O) & 1u;
    if (!PAC) L = L + (frozen ? c.infty : 0.0f);  // polar.py:399,415,471-472
    if (EXTRAS) {
        if (c.llr_out) c.llr_out[O] = L;
    }
    uint32_t s = L < 0.0f;
    c.tie |= (L == 0.0f);
    if (EXTRAS) {
        if (c.have_gt) s = (c.gt_s >> O) & 1u;
    }
    if (PAC) {
        const uint32_t par_bit = __popc(c.pac_state & c.pac_taps) & 1u;
        if (frozen) {
            if (!(EXTRAS && c.have_gt)) {
                s = par_bit;
                c.pac_state = (c.pac_state << 1) & c.pac_state_mask;
            }
        } else {
            const uint32_t v = s ^ par_bit;
            c.vs |= v << O;
            c.pac_state = ((c.pac_state << 1) | v) & c.pac_state_mask;
        }
    }
    c.us |= s << O;
    c.ps |= s << O;
}

// node covering leaves [O, O+S) of the current block; L = its S LLRs in registers
template <int S, int O, int BLOG, bool PAC, bool EXTRAS>
__device__ __forceinline__ void lane_node(const float (&L)[S], LaneCtx<BLOG, PAC, EXTRAS> &c)
{
    if constexpr (!PAC && !EXTRAS && S >= 4) {
        // simplified SC (see the quad kernel): subtrees that are entirely frozen / entirely information are not walked.
        // The frozen pattern is the same for all 32 codewords of the warp, so the branch is warp-uniform; a codeword
        // that violates a shortcut's condition is flagged (c.tie) and re-decoded by the exact path.
        constexpr uint32_t M = (S >= 32) ? 0xffffffffu : ((1u << S) - 1u);
        const uint32_t fm = (c.frozen >> O) & M;
        if (fm == M) {
            // rate-0: every leaf satisfies |L| <= sum |alpha_i| < infty, so sign(L + infty) = +1 (polar.py:399, 471-472)
            float sm = fabsf(L[0]);
#pragma unroll
            for (int j = 1; j < S; ++j) sm += fabsf(L[j]);
            c.tie |= !(sm < 0.99f * c.infty);
            return;
        }
        if (fm == 0u) {
            // rate-1: partial sums = hard decisions of the node's LLRs, u = x F^(x)s (exact unless an LLR is 0)
            uint32_t x = 0u;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                x |= (__float_as_uint(L[j]) >> 31) << j;
                c.tie |= (L[j] == 0.0f);
            }
            uint32_t u = x;
#pragma unroll
            for (int h = 1; h < S; h <<= 1) {
                uint32_t mk = 0u;
#pragma unroll
                for (int b = 0; b < S; ++b)
                    if (!(b & h)) mk |= 1u << b;
                u ^= (u >> h) & mk;
            }
            c.ps |= x << O;
            c.us |= u << O;
            return;
        }
    }
    if constexpr (S == 1) {
        lane_leaf<O>(L[0], c);
    } else {
        constexpr int H = S / 2;
        float C[H];
#pragma unroll
        for (int j = 0; j < H; ++j) C[j] = npd_f_minsum(L[j], L[j + H]);
        lane_node<H, O>(C, c);
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const uint32_t sg = (c.ps << (31 - (O + j))) & 0x80000000u;
            C[j] = __uint_as_float(__float_as_uint(L[j]) ^ sg) + L[j + H];
        }
        lane_node<H, O + H>(C, c);
        constexpr uint32_t lowmask = (H >= 32) ? 0xffffffffu : ((1u << H) - 1u);
        c.ps ^= (c.ps >> H) & (lowmask << O);
    }
}

// First stored level (level slog, hS = 2^slog elements) of every codeword of the group, recomputed from
// y for the path to leaf block o.  lane = element (coalesced global loads), transposed store into
// dst[element][33].  DEPTH = n-1-slog unstored levels sit between level slog and the root, so one output
// element j folds the T = 2^(DEPTH+1) root values y[j + t*hS]; level L combines the values t and
// t + half (element j + t*hS of that level's node) with g when the node on the path is a right child
// (bit L of o set; partial sums of its left sibling = bits [s-h, s), s = node start) and f otherwise.
// U work items are loaded before any is reduced so that U*T global loads are in flight per lane.
template <int DEPTH, int U>
__device__ __forceinline__ void lane_top_phase(const ScParams &p, float *dst, const uint32_t *PS,
                                               int64_t cw0, int nvalid, int o, int lane)
{
    constexpr int T = 2 << DEPTH;
    const int N = 1 << p.n;
    const int slog = p.slog;
    const int hS = 1 << slog;
    const int ilog = slog - 5;          // log2(iterations per codeword)
    const int total = nvalid << ilog;   // work items: (codeword, 32-element slice)
    for (int it0 = 0; it0 < total; it0 += U) {
        float v[U][T];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int it = min(it0 + u, total - 1);
            const int cc = it >> ilog;
            const int j = ((it & ((1 << ilog) - 1)) << 5) + lane;
            const float *yrow = p.y + (cw0 + cc) * N + j;
#pragma unroll
            for (int t = 0; t < T; ++t) v[u][t] = __ldg(yrow + t * hS);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int it = it0 + u;
            const int cc = min(it, total - 1) >> ilog;
            const int j = ((it & ((1 << ilog) - 1)) << 5) + lane;
#pragma unroll
            for (int t = 0; t < T; ++t) v[u][t] = p.scale * v[u][t];  // polar.py:468-469
#pragma unroll
            for (int d = 0; d <= DEPTH; ++d) {
                const int lv = p.n - 1 - d;        // level being produced
                const int h = 1 << lv;
                constexpr int dummy = 0; (void)dummy;
                const int half = T >> (d + 1);
                if (o & h) {
                    const int bit0 = (o & ~(h - 1)) - h + j;
#pragma unroll
                    for (int t = 0; t < T / 2; ++t) {
                        if (t < half) {
                            const int bit = bit0 + t * hS;
                            const uint32_t sg = ((PS[(bit >> 5) * 32 + cc] >> (bit & 31)) & 1u) << 31;
                            v[u][t] = __uint_as_float(__float_as_uint(v[u][t]) ^ sg) + v[u][t + half];
                        }
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < T / 2; ++t)
                        if (t < half) v[u][t] = npd_f_minsum(v[u][t], v[u][t + half]);
                }
            }
            if (it < total) dst[j * 33 + cc] = v[u][0];
        }
    }
}

__host__ __device__ inline size_t lane_warp_smem_bytes(int n, int slog, bool pac)
{
    if (n <= 5) return 16;
    const int NW = (1 << n) >> 5;
    const size_t tree = (size_t)4 * 33 * ((2u << slog) - 32u);
    return tree + (size_t)4 * (pac ? 3 : 2) * NW * 32;
}

template <int BLOG, bool PAC, bool EXTRAS>
__global__ void __launch_bounds__(128) sc_lane_kernel(const ScParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int BS = 1 << BLOG;  // leaves per bottom block
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int n = p.n;
    const int N = 1 << n;
    const int NW = N >> 5;  // words per plane (only used when n > 5)
    const int slog = p.slog;

    unsigned char *base = smem_raw + (size_t)warp * lane_warp_smem_bytes(n, slog, PAC);
    float *tree = reinterpret_cast<float *>(base);  // level lv (5 <= lv <= slog) at 33*(2^lv - 32)
    uint32_t *PS = reinterpret_cast<uint32_t *>(tree + (size_t)33 * ((2u << slog) - 32u));
    uint32_t *US = PS + NW * 32;
    uint32_t *VS = US + NW * 32;

    const int64_t ngroups = (p.B + 31) / 32;
    for (int64_t grp = (int64_t)blockIdx.x * wpb + warp; grp < ngroups; grp += (int64_t)gridDim.x * wpb) {
        const int64_t cw0 = grp * 32;
        const int64_t cw = cw0 + lane;
        const bool ok = cw < p.B;
        const int nvalid = (int)min((int64_t)32, p.B - cw0);

        LaneCtx<BLOG, PAC, EXTRAS> c;
        c.tie = 0u;
        c.pac_state = 0u;
        c.pac_taps = p.pac_taps;
        c.pac_state_mask = p.pac_state_mask;
        c.infty = p.infty;
        c.have_gt = EXTRAS && (p.use_gt != nullptr);
        c.llr_out = nullptr;
        c.gt_s = 0u;

        const int nblocks = N >> BLOG;
        for (int q = 0; q < nblocks; ++q) {
            const int o = q << BLOG;
            float L[BS];
            if (n > BLOG) {
                // ---- refresh the stored levels on the path to block q (only possible when BLOG == 5) ----
                const int top = (q == 0) ? n - 1 : (BLOG + __ffs(q) - 1);
                if (top >= slog) {
                    // first stored level, from y (coalesced, lane = element), transposed store
                    float *dst = tree + (size_t)33 * ((1u << slog) - 32u);
                    __syncwarp();
                    switch (n - 1 - slog) {
                    case 0: lane_top_phase<0, 8>(p, dst, PS, cw0, nvalid, o, lane); break;
                    case 1: lane_top_phase<1, 8>(p, dst, PS, cw0, nvalid, o, lane); break;
                    case 2: lane_top_phase<2, 4>(p, dst, PS, cw0, nvalid, o, lane); break;
                    case 3: lane_top_phase<3, 2>(p, dst, PS, cw0, nvalid, o, lane); break;
                    default: lane_top_phase<4, 1>(p, dst, PS, cw0, nvalid, o, lane); break;
                    }
                    __syncwarp();
                }
                // stored level -> stored level, lane = codeword
                for (int lv = min(top, slog - 1); lv >= BLOG; --lv) {
                    const int h = 1 << lv;
                    const float *par = tree + (size_t)33 * ((2u << lv) - 32u) + lane;
                    float *ch = tree + (size_t)33 * ((1u << lv) - 32u) + lane;
                    if (lv == top) {  // right child (q > 0): g with bits [o-h, o)
                        const int wbase = (o - h) >> 5;
                        for (int e0 = 0; e0 < h; e0 += 32) {
                            const uint32_t w = PS[(wbase + (e0 >> 5)) * 32 + lane];
#pragma unroll 8
                            for (int e = 0; e < 32; ++e) {
                                const float a = par[(e0 + e) * 33], b = par[(e0 + e + h) * 33];
                                const uint32_t sg = ((w >> e) & 1u) << 31;
                                ch[(e0 + e) * 33] = __uint_as_float(__float_as_uint(a) ^ sg) + b;
                            }
                        }
                    } else {
#pragma unroll 8
                        for (int e = 0; e < h; ++e) ch[e * 33] = npd_f_minsum(par[e * 33], par[(e + h) * 33]);
                    }
                }
                const float *src = tree + lane;  // level BLOG (= 5) sits at offset 0
#pragma unroll
                for (int e = 0; e < BS; ++e) L[e] = src[e * 33];
            } else {
                // the whole code is one block: root straight from global (N <= 32)
#pragma unroll
                for (int e = 0; e < BS; ++e) L[e] = ok ? p.scale * __ldg(p.y + cw * N + e) : 0.0f;
            }

            c.ps = 0u;
            c.us = 0u;
            c.vs = 0u;
            c.frozen = __ldg(p.frozen_words + (o >> 5));
            if (EXTRAS) {
                c.llr_out = (p.leaf_llr && ok) ? p.leaf_llr + cw * N + o : nullptr;
                if (c.have_gt) {
                    uint32_t gs = 0u;
                    if (ok)
                        for (int e = 0; e < BS; ++e) {
                            const float t = p.use_gt[cw * N + o + e];
                            gs |= (uint32_t)(t < 0.0f) << e;
                            c.tie |= (t == 0.0f);
                        }
                    c.gt_s = gs;
                }
            }
            lane_node<BS, 0>(L, c);

            if (n > BLOG) {
                PS[q * 32 + lane] = c.ps;
                US[q * 32 + lane] = c.us;
                if (PAC) VS[q * 32 + lane] = c.vs;
                // word-level merges: every trailing one of q completes a block of 2^(5+j+1) leaves
                const int m = __ffs(~q) - 1;
                for (int j = 0; j < m; ++j) {
                    const int nw = 1 << j;
                    const int wl = q + 1 - 2 * nw;
                    for (int i = 0; i < nw; ++i) PS[(wl + i) * 32 + lane] ^= PS[(wl + nw + i) * 32 + lane];
                }
            }
        }

        // ---- outputs (coalesced: lane = k), then the tie sentinel ----
        if (n > BLOG) {
            __syncwarp();
            const uint32_t *OS = PAC ? VS : US;
            for (int cc = 0; cc < nvalid; ++cc) {
                float *dst = p.decoded + (cw0 + cc) * p.K;
                for (int k = lane; k < p.K; k += 32) {
                    const int pos = __ldg(p.info + k);
                    dst[k] = ((OS[(pos >> 5) * 32 + cc] >> (pos & 31)) & 1u) ? -1.0f : 1.0f;
                }
                if (PAC && p.u_hat) {
                    float *du = p.u_hat + (cw0 + cc) * N;
                    for (int e = lane; e < N; e += 32)
                        du[e] = ((US[(e >> 5) * 32 + cc] >> (e & 31)) & 1u) ? -1.0f : 1.0f;
                }
            }
            __syncwarp();
        } else if (ok) {
            const uint32_t os = PAC ? c.vs : c.us;
            for (int k = 0; k < p.K; ++k) {
                const int pos = __ldg(p.info + k);
                p.decoded[cw * p.K + k] = ((os >> pos) & 1u) ? -1.0f : 1.0f;
            }
            if (PAC && p.u_hat)
                for (int e = 0; e < N; ++e) p.u_hat[cw * N + e] = ((c.us >> e) & 1u) ? -1.0f : 1.0f;
        }
        if (ok && c.tie) p.decoded[cw * p.K] = __int_as_float(0x7fc00000);  // re-decode in the exact path
        __syncwarp();
    }
}

// =================================================================================================
// quad path: 4 lanes per codeword, 8 codewords per warp, simplified SC (rate-0 / rate-1 nodes pruned)
// =================================================================================================
// The lane kernel above keeps 32 codewords per warp, so at N >= 256 shared-memory capacity leaves only a few
// warps per SM and the unstored top levels are recomputed from y again and again.  Here a codeword is spread
// over 4 lanes (element e of every level lives on lane quarter e mod 4), a warp holds 8 codewords, every level
// from 6 to n-2 is stored ([element][8 codewords], XOR-swizzled so that both the lane = element writes of the
// top phase and the lane = (quarter, codeword) accesses are conflict-free) and only level n-1 is folded into
// the computation of level n-2 from y.  The bottom 32-leaf block is unrolled in registers (8+4+2+1 values per
// lane, the last two levels replicated through shuffles).  Inside it, subtrees of >= 4 leaves that are
// entirely frozen or entirely information are not traversed (simplified SC, Alamdar-Yazdi & Kschischang):
//   * rate-1: the node's partial sums are the hard decisions of its LLRs and u = x F^(x)s; with min-sum f / g
//     this is what the reference's leaf-by-leaf walk produces whenever no LLR of the node is exactly 0
//     (sign(f) = sign(a) sign(b), g = sign(b) (|a| + |b|));
//   * rate-0: every leaf satisfies |L| <= sum |alpha_i|, so if that sum is below the frozen prior all decisions
//     are +1 exactly as sign(L + infty) gives (polar.py:399, 471-472).
// A codeword that violates either condition (or hits a zero leaf LLR) is flagged with the NaN sentinel and
// re-decoded by the exact path, so the results stay bit-identical to the reference.
// decisions are written once and never re-read by the kernel: streaming stores keep them from displacing the y rows in L2
#define NPD_SC_ST4(p, v) __stcs((p), (v))
#define NPD_SC_ST1(p, v) __stcs((p), (v))
struct QuadCtx {
    uint32_t ps, us, frozen, flag;
    float thr0, infty;
    int sub;
};

template <int O>
__device__ __forceinline__ void quad_leaf(float L, QuadCtx &c)
{
    L = L + (((c.frozen >> O) & 1u) ? c.infty : 0.0f);  // polar.py:399,415,471-472
    const uint32_t s = L < 0.0f;
    c.flag |= (L == 0.0f);
    c.us |= s << O;
    c.ps |= s << O;
}

// node of 2 leaves; v = element (sub & 1), replicated on the lane quarters sub and sub ^ 2
template <int O>
__device__ __forceinline__ void quad_pair(float v, QuadCtx &c)
{
    const float w = __shfl_xor_sync(NPD_FULL, v, 8);
    const float a = (c.sub & 1) ? w : v, b = (c.sub & 1) ? v : w;
    quad_leaf<O>(npd_f_minsum(a, b), c);
    const uint32_t sg = (c.ps << (31 - O)) & 0x80000000u;
    quad_leaf<O + 1>(__uint_as_float(__float_as_uint(a) ^ sg) + b, c);
    c.ps ^= (c.ps >> 1) & (1u << O);
}

// node covering leaves [O, O+S) of the block, S >= 4; L[i] = element 4 i + sub
template <int S, int O>
__device__ __forceinline__ void quad_node(const float (&L)[S / 4], QuadCtx &c)
{
    constexpr uint32_t M = (S >= 32) ? 0xffffffffu : ((1u << S) - 1u);
    const uint32_t fm = (c.frozen >> O) & M;  // warp-uniform
    if (fm == M) {
        // rate-0: all decisions +1 provided sum |alpha| < infty (checked; else exact re-decode)
        float sm = fabsf(L[0]);
#pragma unroll
        for (int i = 1; i < S / 4; ++i) sm += fabsf(L[i]);
        sm += __shfl_xor_sync(NPD_FULL, sm, 8);
        sm += __shfl_xor_sync(NPD_FULL, sm, 16);
        c.flag |= !(sm < c.thr0);
        return;
    }
    if (fm == 0u) {
        // rate-1: partial sums = hard decisions, u = x F^(x)s
        uint32_t x = 0u;
#pragma unroll
        for (int i = 0; i < S / 4; ++i) {
            x |= (__float_as_uint(L[i]) >> 31) << (4 * i);
            c.flag |= (L[i] == 0.0f);
        }
        x <<= c.sub;
        x |= __shfl_xor_sync(NPD_FULL, x, 8);
        x |= __shfl_xor_sync(NPD_FULL, x, 16);
        uint32_t u = x;
#pragma unroll
        for (int h = 1; h < S; h <<= 1) {
            // positions of the node whose index bit h is 0
            uint32_t mk = 0u;
#pragma unroll
            for (int b = 0; b < S; ++b)
                if (!(b & h)) mk |= 1u << b;
            u ^= (u >> h) & mk;
        }
        c.ps |= x << O;
        c.us |= u << O;
        return;
    }
    if constexpr (S == 4) {
        const float w = __shfl_xor_sync(NPD_FULL, L[0], 16);
        const float a = (c.sub & 2) ? w : L[0], b = (c.sub & 2) ? L[0] : w;
        quad_pair<O>(npd_f_minsum(a, b), c);
        const uint32_t sg = (c.ps << (31 - O - (c.sub & 1))) & 0x80000000u;
        quad_pair<O + 2>(__uint_as_float(__float_as_uint(a) ^ sg) + b, c);
        c.ps ^= (c.ps >> 2) & (3u << O);
    } else {
        constexpr int H = S / 2, HQ = H / 4;
        float C[HQ];
#pragma unroll
        for (int i = 0; i < HQ; ++i) C[i] = npd_f_minsum(L[i], L[i + HQ]);
        quad_node<H, O>(C, c);
#pragma unroll
        for (int i = 0; i < HQ; ++i) {
            const uint32_t sg = ((c.ps >> (O + 4 * i)) >> c.sub) << 31;
            C[i] = __uint_as_float(__float_as_uint(L[i]) ^ sg) + L[i + HQ];
        }
        quad_node<H, O + H>(C, c);
        constexpr uint32_t lowmask = (H >= 32) ? 0xffffffffu : ((1u << H) - 1u);
        c.ps ^= (c.ps >> H) & (lowmask << O);
    }
}

// word index of element e of codeword c inside a stored level: [element][8], codeword XOR-swizzled
__device__ __forceinline__ int quad_idx(int e, int c) { return e * 8 + (c ^ ((e >> 2) & 7)); }

// gl: the top gl stored levels (n-2: half of the stored tree, n-3: a quarter) live in an L2-resident global scratch
// instead of shared memory
__host__ __device__ inline size_t quad_warp_smem_bytes(int n, int gl = 0)
{
    const int slog = n - 2;
    return (size_t)4 * 8 * ((2u << (slog - gl)) - 64u) + (size_t)4 * 2 * ((1u << n) >> 5) * 8;
}
// floats of global scratch per warp: level L <= n-2 starts at quad_scratch_off(n, L)
__host__ __device__ inline size_t quad_scratch_off(int n, int level) { return (size_t)8 * ((2u << (n - 2)) - (2u << level)); }
__host__ __device__ inline size_t quad_scratch_floats(int n, int gl) { return gl ? quad_scratch_off(n, n - 2 - gl) : 0; }

// Level n-2 of the 8 codewords of the group for quarter r of the code, from y: lane = element (coalesced loads
// of y[j], y[j+h], y[j+2h], y[j+3h], h = N/4, 32 loads in flight per lane), level n-1 folded in.  G1 / G0: the
// level n-1 / n-2 node on the path is a right child (g with the partial sums of its left sibling) or not (f).
template <int NLOG, bool G1, bool G0, bool FULL, bool GTOP = false>
__device__ __forceinline__ void quad_top_phase(const ScParams &p, float *dst, const uint32_t *PS, const float *ygrp,
                                               int nvalid, int r, int lane)
{
    constexpr int N = 1 << NLOG, HS = N >> 2, SLICES = HS >> 5;
    const int kx = (lane >> 2) & 7;  // swizzle key of element j = 32 slice + lane
    auto load = [&](float (&v)[8][4], int slice) {
        const float *yj = ygrp + slice * 32 + lane;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const float *row = yj + (FULL ? cc : min(cc, nvalid - 1)) * N;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                // the fourth pass over a codeword's y is the last one: stream it (evict-first) so that the rows of resident
                // groups, which are re-read, keep their place in L2 (with the streaming decision stores: -3.7 % at N = 1024)
                v[cc][t] = (G1 && G0) ? __ldcs(row + t * HS) : __ldg(row + t * HS);
            }
        }
    };
    auto reduce = [&](float (&v)[8][4], int slice) {
        float *drow = dst + (slice * 32 + lane) * 8;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
#pragma unroll
            for (int t = 0; t < 4; ++t) v[cc][t] = p.scale * v[cc][t];  // polar.py:468-469
            float a0, a1;  // level n-1 elements j and j + HS
            if (G1) {
                const uint32_t s0 = ((PS[slice * 8 + cc] >> lane) & 1u) << 31;
                const uint32_t s1 = ((PS[(slice + SLICES) * 8 + cc] >> lane) & 1u) << 31;
                a0 = __uint_as_float(__float_as_uint(v[cc][0]) ^ s0) + v[cc][2];
                a1 = __uint_as_float(__float_as_uint(v[cc][1]) ^ s1) + v[cc][3];
            } else {
                a0 = npd_f_minsum(v[cc][0], v[cc][2]);
                a1 = npd_f_minsum(v[cc][1], v[cc][3]);
            }
            float o;
            if (G0) {
                const uint32_t sg = ((PS[((r - 1) * SLICES + slice) * 8 + cc] >> lane) & 1u) << 31;
                o = __uint_as_float(__float_as_uint(a0) ^ sg) + a1;
            } else {
                o = npd_f_minsum(a0, a1);
            }
            if (GTOP) v[cc][0] = o;  // global scratch: no bank swizzle, the lane's 32-byte row goes out as two vectors
            else drow[cc ^ kx] = o;
        }
        if (GTOP) {
            __stcg(reinterpret_cast<float4 *>(drow), make_float4(v[0][0], v[1][0], v[2][0], v[3][0]));
            __stcg(reinterpret_cast<float4 *>(drow) + 1, make_float4(v[4][0], v[5][0], v[6][0], v[7][0]));
        }
    };
    // explicit two-deep software pipeline: the 32 loads of the next slice are in flight while this one is reduced
    float va[8][4], vb[8][4];
    load(va, 0);
#pragma unroll 1
    for (int slice = 0; slice < SLICES; slice += 2) {
        load(vb, slice + 1);
        reduce(va, slice);
        if (slice + 2 < SLICES) load(va, slice + 2);
        reduce(vb, slice + 1);
    }
}

// stored level LV+1 -> stored level LV; the lane quarter `sub` owns elements 4 i + sub, off[ii] = swizzled word
// offset of element 4 ii + sub inside an aligned group of 32 elements
template <int LV, bool G>
__device__ __forceinline__ void quad_level(float *tree, const uint32_t *PSw, const int (&off)[8], int sub)
{
    constexpr int H = 1 << LV;
    const float *par = tree + 8 * ((2 << LV) - 64);
    float *ch = tree + 8 * ((1 << LV) - 64);
#pragma unroll 2
    for (int i0 = 0; i0 < H / 4; i0 += 8) {
        uint32_t w = 0u;
        if (G) w = PSw[i0] >> sub;  // word (i0 / 8) of the left sibling's partial sums, stride 8 words
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
            const int idx = 32 * i0 + off[ii];
            const float a = par[idx], b = par[idx + 8 * H];
            if (G) {
                const uint32_t sg = ((w >> (4 * ii)) & 1u) << 31;
                ch[idx] = __uint_as_float(__float_as_uint(a) ^ sg) + b;
            } else {
                ch[idx] = npd_f_minsum(a, b);
            }
        }
    }
}

// global scratch level LV+1 ([element][8 codewords], unswizzled) -> stored level LV, lane = element like the top phase:
// a lane reads its two 32-byte rows as four 16-byte vectors (the (quarter, codeword) mapping of quad_level would issue
// sixteen 4-byte loads for the same data) and stores transposed with the level's bank swizzle
template <int LV, bool G, bool GOUT>
__device__ __forceinline__ void quad_level_from_scratch(const float *gpar, float *tree, float *gch, const uint32_t *PSl, int lane)
{
    constexpr int H = 1 << LV;
    float *ch = tree + 8 * ((1 << LV) - 64);
    const int kx = (lane >> 2) & 7;
#pragma unroll 4
    for (int slice = 0; slice < H / 32; ++slice) {
        const int j = slice * 32 + lane;
        const float4 *pa = reinterpret_cast<const float4 *>(gpar + (size_t)j * 8);
        const float4 *pb = reinterpret_cast<const float4 *>(gpar + (size_t)(j + H) * 8);
        const float4 a0 = __ldcg(pa), a1 = __ldcg(pa + 1), b0 = __ldcg(pb), b1 = __ldcg(pb + 1);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float *drow = ch + j * 8;
        float o[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            if (G) {
                const uint32_t sg = ((PSl[slice * 8 + cc] >> lane) & 1u) << 31;
                o[cc] = __uint_as_float(__float_as_uint(a[cc]) ^ sg) + b[cc];
            } else {
                o[cc] = npd_f_minsum(a[cc], b[cc]);
            }
            if (!GOUT) drow[cc ^ kx] = o[cc];
        }
        if (GOUT) {  // the child level is a scratch level too
            __stcg(reinterpret_cast<float4 *>(gch + (size_t)j * 8), make_float4(o[0], o[1], o[2], o[3]));
            __stcg(reinterpret_cast<float4 *>(gch + (size_t)j * 8) + 1, make_float4(o[4], o[5], o[6], o[7]));
        }
    }
}

// GL = how many of the levels LV+1, LV, ... are scratch levels (LV+1 is one when GL >= 1, LV itself when GL >= 2);
// NL = log2 of the code length
template <int LV, int GL = 0, int NL = 0>
struct QuadLevelsDown {
    static __device__ __forceinline__ void run(float *tree, const uint32_t *PS, const int (&off)[8], int sub, int cl, int top, int o,
                                               float *scratch = nullptr)
    {
        if (LV <= top) {  // warp-uniform
            if (GL >= 1) {
                const float *gpar = scratch + quad_scratch_off(NL, LV + 1);
                float *gch = scratch + quad_scratch_off(NL, LV);
                __syncwarp();
                if (LV == top)
                    quad_level_from_scratch<LV, true, (GL >= 2)>(gpar, tree, gch, PS + ((o - (1 << LV)) >> 5) * 8, 8 * sub + cl);
                else
                    quad_level_from_scratch<LV, false, (GL >= 2)>(gpar, tree, gch, nullptr, 8 * sub + cl);
                __syncwarp();
            } else if (LV == top) {
                quad_level<LV, true>(tree, PS + ((o - (1 << LV)) >> 5) * 8 + cl, off, sub);
            } else {
                quad_level<LV, false>(tree, nullptr, off, sub);
            }
        }
        QuadLevelsDown<LV - 1, (GL > 0 ? GL - 1 : 0), NL>::run(tree, PS, off, sub, cl, top, o, scratch);
    }
};
template <int NL>
struct QuadLevelsDown<5, 0, NL> {
    static __device__ __forceinline__ void run(float *, const uint32_t *, const int (&)[8], int, int, int, int, float * = nullptr) {}
};

template <int NLOG, bool TRACE = false, int GL = 0>
__global__ void __launch_bounds__(128) sc_quad_kernel(const ScParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = 1 << NLOG, NW = N >> 5, SLOG = NLOG - 2;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int sub = lane >> 3, cl = lane & 7;

    constexpr bool GTOP = GL > 0;
    unsigned char *base = smem_raw + (size_t)warp * quad_warp_smem_bytes(NLOG, GL);
    float *tree = reinterpret_cast<float *>(base);  // level lv (6 <= lv <= SLOG - GL) at 8 * (2^lv - 64)
    uint32_t *PS = reinterpret_cast<uint32_t *>(tree + 8 * ((2 << (SLOG - GL)) - 64));
    uint32_t *US = PS + NW * 8;
    // level SLOG (the current quarter of the code, half of the stored tree): shared memory, or -- GL > 0, large N --
    // this warp's slice of a global scratch that stays in L2 (written once, read twice as whole 32-byte rows)
    float *scratch = GTOP ? p.scratch + ((size_t)blockIdx.x * wpb + warp) * quad_scratch_floats(NLOG, GL) : nullptr;
    float *top_dst = GTOP ? scratch : tree + 8 * ((1 << SLOG) - 64);

    int off[8];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) off[ii] = 32 * ii + 8 * sub + (cl ^ ii);

    QuadCtx c;
    c.sub = sub;
    c.infty = p.infty;
    c.thr0 = 0.99f * p.infty;

    unsigned long long cnt_bits = 0, cnt_blocks = 0;  // fused counting: this lane's codewords (lanes with sub == 0)
    const int64_t ngroups = (p.B + 7) / 8;
    for (int64_t grp = (int64_t)blockIdx.x * wpb + warp; grp < ngroups; grp += (int64_t)gridDim.x * wpb) {
        const int64_t cw0 = grp * 8;
        const int64_t cw = cw0 + cl;
        const bool ok = cw < p.B;
        const int nvalid = (int)min((int64_t)8, p.B - cw0);
        const float *ygrp = p.y + cw0 * N;
        c.flag = 0u;
        const bool tr = TRACE && p.trace && blockIdx.x == 0 && warp == 0 && grp == (int64_t)gridDim.x * wpb;
        long long t_top = 0, t_lev = 0, t_blk = 0, t_mrg = 0, t_a = 0, t_b, t_start = tr ? clock64() : 0;

#pragma unroll 1
        for (int q = 0; q < NW; ++q) {
            const int o = q << 5;
            const int top = (q == 0) ? NLOG - 1 : (5 + __ffs(q) - 1);
            if (tr) t_a = clock64();
            if (top >= SLOG) {
                const int r = q >> (SLOG - 5);
                __syncwarp();
                if (nvalid == 8) {
                    switch (r) {
                    case 0: quad_top_phase<NLOG, false, false, true, GTOP>(p, top_dst, PS, ygrp, 8, r, lane); break;
                    case 1: quad_top_phase<NLOG, false, true, true, GTOP>(p, top_dst, PS, ygrp, 8, r, lane); break;
                    case 2: quad_top_phase<NLOG, true, false, true, GTOP>(p, top_dst, PS, ygrp, 8, r, lane); break;
                    default: quad_top_phase<NLOG, true, true, true, GTOP>(p, top_dst, PS, ygrp, 8, r, lane); break;
                    }
                } else {
                    switch (r) {
                    case 0: quad_top_phase<NLOG, false, false, false, GTOP>(p, top_dst, PS, ygrp, nvalid, r, lane); break;
                    case 1: quad_top_phase<NLOG, false, true, false, GTOP>(p, top_dst, PS, ygrp, nvalid, r, lane); break;
                    case 2: quad_top_phase<NLOG, true, false, false, GTOP>(p, top_dst, PS, ygrp, nvalid, r, lane); break;
                    default: quad_top_phase<NLOG, true, true, false, GTOP>(p, top_dst, PS, ygrp, nvalid, r, lane); break;
                    }
                }
                __syncwarp();
            }
            if (tr) { t_b = clock64(); t_top += t_b - t_a; t_a = t_b; }
            QuadLevelsDown<SLOG - 1, GL, NLOG>::run(tree, PS, off, sub, cl, top, o, scratch);
            if (tr) { t_b = clock64(); t_lev += t_b - t_a; t_a = t_b; }
            // level 5 straight into registers: L[i] = element 4 i + sub (level 6 sits at offset 0)
            float L[8];
            if (top == 5) {
                const uint32_t w = PS[(q - 1) * 8 + cl] >> sub;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t sg = ((w >> (4 * i)) & 1u) << 31;
                    L[i] = __uint_as_float(__float_as_uint(tree[off[i]]) ^ sg) + tree[off[i] + 256];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) L[i] = npd_f_minsum(tree[off[i]], tree[off[i] + 256]);
            }
            c.ps = 0u;
            c.us = 0u;
            c.frozen = __ldg(p.frozen_words + q);
            quad_node<32, 0>(L, c);

            if (sub == 0) {
                PS[q * 8 + cl] = c.ps;
                US[q * 8 + cl] = c.us;
            }
            __syncwarp();
            if (tr) { t_b = clock64(); t_blk += t_b - t_a; t_a = t_b; }
            // word-level merges: every trailing one of q completes a block of 2^(5+j+1) leaves
            const int m = __ffs(~q) - 1;
            for (int j = 0; j < m; ++j) {
                const int nw = 1 << j;
                const int wl = q + 1 - 2 * nw;
                for (int i = sub; i < nw; i += 4) PS[(wl + i) * 8 + cl] ^= PS[(wl + nw + i) * 8 + cl];
                __syncwarp();
            }
            if (tr) { t_b = clock64(); t_mrg += t_b - t_a; }
        }
        const long long t_loop_end = tr ? clock64() : 0;

        // ---- outputs (coalesced: lane = k), then the sentinel of codewords that need the exact path ----
        uint32_t fl = c.flag;
        fl |= __shfl_xor_sync(NPD_FULL, fl, 8);
        fl |= __shfl_xor_sync(NPD_FULL, fl, 16);
        __syncwarp();
        if (p.us_out) {
            if (ok) {
                uint32_t *uo = p.us_out + cw * p.out_words + p.word_off, *xo = p.xs_out + cw * p.out_words + p.word_off;
                for (int q = sub; q < NW; q += 4) {
                    uo[q] = US[q * 8 + cl];
                    xo[q] = PS[q * 8 + cl];  // after the last merge: the Plotkin encoding of the block's decisions
                }
                if (sub == 0 && fl) p.flags[cw] = 1;
            }
            __syncwarp();
            continue;
        }
        if (p.counts) {
            // fused counting: lane (sub, cl) xors the words q = sub, sub + 4, ... of codeword cl's decisions with the
            // transmitted u words, masked to the info positions; flagged codewords are left to the exact path
            uint32_t e = 0u;
            if (ok) {
                const uint32_t *ub = p.ubits + cw * NW;
                for (int q = sub; q < NW; q += 4) e += __popc((US[q * 8 + cl] ^ __ldg(ub + q)) & __ldg(p.info_words + q));
            }
            e += __shfl_xor_sync(NPD_FULL, e, 8);
            e += __shfl_xor_sync(NPD_FULL, e, 16);
            if (sub == 0 && ok) {
                p.flags[cw] = fl ? 1 : 0;
                if (!fl) {
                    cnt_bits += e;
                    cnt_blocks += e != 0u;
                }
            }
            __syncwarp();
            continue;
        }
        float *dst0 = p.decoded + cw0 * p.K;
        // (measured: +4.5 % at N = 256, where the output phase is a larger share; nothing at N = 1024, so the code is
        // only compiled into the short-code kernels)
        if (NLOG <= 9 && p.vec_out) {
            // four consecutive k per lane: the info positions arrive as one 16-byte load, the 8 codewords' decision
            // words of a position as two, and every codeword's four decisions leave as one 16-byte store
            for (int k0 = 0; k0 < p.K; k0 += 128 * 4) {
                int4 pos[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = k0 + 128 * i + 4 * lane;
                    pos[i] = k < p.K ? __ldg(reinterpret_cast<const int4 *>(p.info + k)) : make_int4(0, 0, 0, 0);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = k0 + 128 * i + 4 * lane;
                    if (k < p.K) {
                        const int ps[4] = {pos[i].x, pos[i].y, pos[i].z, pos[i].w};
                        uint32_t w[4][8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 lo = *reinterpret_cast<const uint4 *>(US + (ps[j] >> 5) * 8);
                            const uint4 hi = *reinterpret_cast<const uint4 *>(US + (ps[j] >> 5) * 8 + 4);
                            const int sh = ps[j] & 31;
                            w[j][0] = lo.x >> sh; w[j][1] = lo.y >> sh; w[j][2] = lo.z >> sh; w[j][3] = lo.w >> sh;
                            w[j][4] = hi.x >> sh; w[j][5] = hi.y >> sh; w[j][6] = hi.z >> sh; w[j][7] = hi.w >> sh;
                        }
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc)
                            if (cc < nvalid)
                                NPD_SC_ST4(reinterpret_cast<float4 *>(dst0 + (size_t)cc * p.K + k),
                                    make_float4((w[0][cc] & 1u) ? -1.0f : 1.0f, (w[1][cc] & 1u) ? -1.0f : 1.0f,
                                                (w[2][cc] & 1u) ? -1.0f : 1.0f, (w[3][cc] & 1u) ? -1.0f : 1.0f));
                    }
                }
            }
        } else
        for (int k0 = 0; k0 < p.K; k0 += 32 * 16) {
            // the info positions of 16 rounds are fetched together (they were 16 exposed L2 round trips per group)
            int pos[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = k0 + 32 * i + lane;
                pos[i] = k < p.K ? __ldg(p.info + k) : 0;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = k0 + 32 * i + lane;
                if (k < p.K) {
                    const uint32_t *w = US + (pos[i] >> 5) * 8;
                    const int sh = pos[i] & 31;
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc)
                        if (cc < nvalid) NPD_SC_ST1(dst0 + (size_t)cc * p.K + k, ((w[cc] >> sh) & 1u) ? -1.0f : 1.0f);
                }
            }
        }
        __syncwarp();
        if (ok && fl && sub == 0) p.decoded[cw * p.K] = __int_as_float(0x7fc00000);
        __syncwarp();
        if (tr && lane == 0) {
            const long long t_end = clock64();
            p.trace[0] = t_top; p.trace[1] = t_lev; p.trace[2] = t_blk; p.trace[3] = t_mrg;
            p.trace[4] = t_end - t_loop_end; p.trace[5] = t_end - t_start;
        }
    }
    if (p.counts) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            cnt_bits += __shfl_xor_sync(NPD_FULL, cnt_bits, o);
            cnt_blocks += __shfl_xor_sync(NPD_FULL, cnt_blocks, o);
        }
        if (lane == 0) {
            if (cnt_bits) atomicAdd(p.counts + 0, cnt_bits);
            if (cnt_blocks) atomicAdd(p.counts + 1, cnt_blocks);
        }
    }
}

int env_int(const char *name, int dflt)
{
    const char *v = npd_knob(name);
    return v ? atoi(v) : dflt;
}

// ---- launchers ----------------------------------------------------------------------------------
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
    int blocks_per_sm = (int)((size_t)(228 * 1024) / (per_warp * wpb + 1024));
    if (blocks_per_sm > 32) blocks_per_sm = 32;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    const int64_t units = p.scan_flagged ? (p.B + 31) / 32 : (p.B + G - 1) / G;
    int64_t grid = (int64_t)dp.sm_count * blocks_per_sm;
    const int64_t need = (units + wpb - 1) / wpb;
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
int dispatch_group(const npd_code *code, const ScParams &p, cudaStream_t st)
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

int default_slog(int n, bool)
{
    // highest stored level (tuned on B200, tools/tune_sc.sh): all levels for N <= 128; above that, trade
    // recomputation of the top levels from y against resident warps per SM (4*33*(2^(slog+1)-32) B each)
    if (n <= 7) return n <= 5 ? 5 : n - 1;
    if (n == 8) return 6;
    if (n <= 10) return 7;
    return 8;
}

template <int BLOG, bool PAC, bool EXTRAS>
int launch_lane(const npd_code *code, ScParams p, cudaStream_t st)
{
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    const int n = code->n;
    int slog = env_int("NPD_SC_SLOG", 0);
    if (slog < 5 || slog > n - 1 || n - 1 - slog > 4) slog = default_slog(n, PAC);
    p.slog = slog;
    const size_t per_warp = lane_warp_smem_bytes(n, slog, PAC);
    const size_t budget = (size_t)dp.smem_optin;
    if (per_warp + 1024 > budget) {
        npd_set_error("SC lane kernel: N=%d needs %zu B of shared memory per warp", code->N, per_warp);
        return NPD_EUNSUPPORTED;
    }
    int warps_per_sm = (int)((size_t)(228 * 1024) / (per_warp + 1024));
    int max_warps = 20;  // register file: 96 regs/thread -> <= 20 warps
    if (n - 1 - slog > 0) {
        // the unstored top levels re-read y (2^(n-1-slog+1) times per codeword): keep the y rows of all
        // resident codewords (warps * 32 * 4N bytes per SM) inside ~80 MB of the 126 MB L2 so that only the
        // first read comes from HBM (ncu: 7x DRAM read traffic at 6 warps/SM for N = 1024, 1.0x at <= 4)
        const size_t per_warp_y = (size_t)32 * 4 * code->N * dp.sm_count;
        const int l2_warps = (int)((size_t)80 * 1024 * 1024 / per_warp_y);
        if (l2_warps < max_warps) max_warps = l2_warps < 2 ? 2 : l2_warps;
    }
    max_warps = env_int("NPD_SC_WARPS", max_warps);
    if (warps_per_sm > max_warps) warps_per_sm = max_warps;
    if (warps_per_sm < 1) warps_per_sm = 1;
    int wpb = 1;
    while (wpb < 4 && warps_per_sm % (wpb * 2) == 0 && per_warp * wpb * 2 + 1024 <= budget) wpb *= 2;
    const int blocks_per_sm = warps_per_sm / wpb;
    const int64_t ngroups = (p.B + 31) / 32;
    int64_t grid = (int64_t)dp.sm_count * blocks_per_sm;
    const int64_t need = (ngroups + wpb - 1) / wpb;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    auto kern = sc_lane_kernel<BLOG, PAC, EXTRAS>;
    NPD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(per_warp * wpb)));
    kern<<<(unsigned)grid, 32 * wpb, per_warp * wpb, st>>>(p);
    NPD_CHECK_CUDA(cudaGetLastError());
    return NPD_OK;
}

// Library-owned memory pool for the quad kernel's scratch: stream-ordered allocations that stay cached across
// synchronisations (the device's default pool hands freed memory back at every sync, and the next launch then pays a
// fresh 70 MB mapping -- measured as sporadic 2x slower launches)
int scratch_pool(cudaMemPool_t *out)
{
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    NPD_CHECK_CUDA(cudaGetDevice(&dev));
    NPD_REQUIRE(dev >= 0 && dev < 64, "SC quad kernel: device index %d not supported", dev);
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        NPD_CHECK_CUDA(cudaMemPoolCreate(&pools[dev], &props));
        uint64_t keep = ~(uint64_t)0;
        NPD_CHECK_CUDA(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    *out = pools[dev];
    return NPD_OK;
}

}  // namespace
int npd_scratch_pool(cudaMemPool_t *out) { return scratch_pool(out); }  // shared with gru_decode.cu (residual state)
namespace {

// resident configuration of sc_quad_kernel<n>: scratch levels, shared memory per warp, warps per block, blocks per SM
struct QuadResidency { int gl; size_t per_warp; int wpb, blocks_per_sm; };
int quad_residency(const npd_code *code, const int n, const DeviceProps &dp, QuadResidency *out)
{
    // N >= 2048: levels n-2 and n-3 (3/4 of the stored tree) go to a global scratch: 22 KB instead of 70 KB of shared
    // memory per warp at N = 4096, 10 warps per SM instead of 3 (1.17e7 -> 1.62e7 cw/s; one scratch level: 1.57e7).
    // At N = 1024 the same change measures slower at every occupancy (best: one scratch level, 16 warps, 0.835 ms per
    // 131072 codewords against 0.775 ms for 12 warps with the whole tree in shared memory), so it stays off there
    // (NPD_SC_GTOP10 = 1 | 2 turns it on for experiments).
    const int gl = n >= 11 ? max(0, min(2, env_int("NPD_SC_GTOP", 2)))  // scratch levels: 0, 1 or 2
                           : n == 10 ? max(0, min(2, env_int("NPD_SC_GTOP10", 0))) : 0;
    const size_t per_warp = quad_warp_smem_bytes(n, gl);
    const size_t budget = (size_t)dp.smem_optin;
    if (per_warp + 1024 > budget) {
        npd_set_error("SC quad kernel: N=%d needs %zu B of shared memory per warp", code->N, per_warp);
        return NPD_EUNSUPPORTED;
    }
    int warps_per_sm = (int)((size_t)(228 * 1024) / (per_warp + 256));
    // resident warps per SM: beyond these the y rows and scratch of the resident codewords thrash L2 (measured at
    // N = 4096: 8 warps 1.97e7 cw/s, 10 warps 1.84e7; at N = 2048: 12 warps 5.5e7, 20 warps 5.0e7, 8 warps 4.9e7)
    int max_warps = env_int("NPD_SC_WARPS", n >= 12 ? 8 : n == 11 ? 12 : 24);
    if (warps_per_sm > max_warps) warps_per_sm = max_warps;
    if (warps_per_sm < 1) warps_per_sm = 1;
    // warps per block: 4 unless smaller blocks pack at least 20 % more warps into the SM's 228 KB (every block also
    // reserves 1 KB).  At N = 4096 a warp's state is 70 KB: blocks of two leave a third of the shared memory empty
    // (2 warps per SM, 8.3e6 cw/s; single-warp blocks: 3 warps, 1.17e7).  Small blocks are a last resort, though: at
    // N = 1024 and the same 12 warps per SM, one- or two-warp blocks measured 14 % slower than four-warp blocks
    // (0.89 vs 0.78 ms), and a 13th warp (thirteen single-warp blocks) ran at 9.8e7 cw/s against 1.44e8.
    int wpb = 1, blocks_per_sm = 1;
    {
        int best = 0;
        for (int w = env_int("NPD_SC_WPB", 4); w >= 1; w >>= 1) {
            if (per_warp * w + 1024 > budget) continue;
            int b = (int)((size_t)(228 * 1024) / (per_warp * w + 1024));
            if (b * w > max_warps) b = max_warps / w;
            if (b >= 1 && 5 * b * w >= 6 * best) { best = b * w; wpb = w; blocks_per_sm = b; }
        }
    }
    out->gl = gl; out->per_warp = per_warp; out->wpb = wpb; out->blocks_per_sm = blocks_per_sm;
    return NPD_OK;
}

int launch_quad_n(const npd_code *code, const int n, ScParams p, cudaStream_t st)
{
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    QuadResidency qr;
    if (int rc = quad_residency(code, n, dp, &qr)) return rc;
    const int gl = qr.gl, wpb = qr.wpb, blocks_per_sm = qr.blocks_per_sm;
    const bool gtop = gl > 0;
    const size_t per_warp = qr.per_warp;
    const int64_t ngroups = (p.B + 7) / 8;
    int64_t grid = (int64_t)dp.sm_count * blocks_per_sm;
    const int64_t need = (ngroups + wpb - 1) / wpb;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    p.vec_out = (p.K & 3) == 0 && (reinterpret_cast<uintptr_t>(p.decoded) & 15) == 0 && env_int("NPD_SC_VECOUT", 1) != 0;
    void (*kern)(const ScParams) = nullptr;
    const char *trace_path = npd_knob("NPD_SC_TRACE");  // bench-only (synchronises!): phase cycles of one group, N = 1024
    if (trace_path && n != 10) trace_path = nullptr;
    switch (n) {
    case 8: kern = sc_quad_kernel<8>; break;
    case 9: kern = sc_quad_kernel<9>; break;
    case 10: kern = gl >= 2 ? sc_quad_kernel<10, false, 2> : gl == 1 ? sc_quad_kernel<10, false, 1>
                  : trace_path ? sc_quad_kernel<10, true> : sc_quad_kernel<10>; break;
    case 11: kern = gl >= 2 ? sc_quad_kernel<11, false, 2> : gl == 1 ? sc_quad_kernel<11, false, 1> : sc_quad_kernel<11>; break;
    case 12: kern = gl >= 2 ? sc_quad_kernel<12, false, 2> : gl == 1 ? sc_quad_kernel<12, false, 1> : sc_quad_kernel<12>; break;
    default: npd_set_error("SC quad kernel: n=%d outside 8..12", n); return NPD_EUNSUPPORTED;
    }
    NPD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpb)));
    if (trace_path) {
        NPD_CHECK_CUDA(cudaMalloc(&p.trace, 6 * sizeof(long long)));
        NPD_CHECK_CUDA(cudaMemsetAsync(p.trace, 0, 6 * sizeof(long long), st));
    }
    if (gtop) {  // stream-ordered: concurrent decodes on other streams get their own scratch
        cudaMemPool_t pool;
        if (int rc = scratch_pool(&pool)) return rc;
        NPD_CHECK_CUDA(cudaMallocFromPoolAsync(&p.scratch, (size_t)grid * wpb * quad_scratch_floats(n, gl) * sizeof(float), pool, st));
    }
    kern<<<(unsigned)grid, 32 * wpb, per_warp * wpb, st>>>(p);
    NPD_CHECK_CUDA(cudaGetLastError());
    if (gtop) NPD_CHECK_CUDA(cudaFreeAsync(p.scratch, st));
    if (trace_path) {
        long long h[6];
        NPD_CHECK_CUDA(cudaStreamSynchronize(st));
        NPD_CHECK_CUDA(cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(p.trace);
        if (h[5] == 0) return NPD_OK;  // this launch was too small to reach the traced group
        if (FILE *f = fopen(trace_path, "w")) {
            fprintf(f, "quad kernel, one group of 8 codewords on one warp (%d warps per SM): top phase %lld, stored levels %lld, "
                       "32-leaf blocks %lld, partial-sum merges %lld, output %lld, total %lld cycles\n",
                    blocks_per_sm * wpb, h[0], h[1], h[2], h[3], h[4], h[5]);
            fclose(f);
        }
    }
    return NPD_OK;
}

// =================================================================================================
// split path (N = 4096; works for 2048 too): the top log2(N/1024) levels as streaming kernels, then 1024-leaf sub-block decodes
// =================================================================================================
// A codeword of N = Q * 1024 leaves is Q sub-blocks that the successive-cancellation schedule visits in order; the input
// LLRs of sub-block r (level 10 of the tree) depend on y and on the ENCODED decisions of sub-blocks < r only.  So:
//   for r in 0..Q-1:  split_top_kernel   level-10 LLRs of sub-block r for every codeword of the chunk (one thread per
//                                        element: Q coalesced y loads, the f / g of the top one or two levels)
//                     sc_quad_kernel<10> decodes the sub-blocks (its y = that LLR buffer, scale 1, its frozen words =
//                                        the sub-block's; packed mode: decision words + encoded partial sums as bits)
//   then one output kernel (float decisions on the info positions + NaN sentinel of flagged codewords, or the fused
//   error count).  Arithmetic and its order are those of the monolithic kernel's top phase (quad_top_phase), so results
//   stay bit-identical.  Why: at N = 4096 the monolithic kernel keeps only 8-10 warps per SM resident (22 KB of state
//   per warp + a global scratch) and re-reads 16 KB of y per codeword four times through a latency-bound warp; here the
//   top levels are bandwidth-bound streaming passes at full occupancy and the sub-block decode runs at the N = 1024
//   kernel's 12 warps per SM with its whole tree in shared memory.
template <int T>  // T = log2(Q): 1 or 2 levels above the sub-blocks
__global__ void __launch_bounds__(256) split_top_kernel(const float *__restrict__ y, const uint32_t *__restrict__ xs,
                                                        float *__restrict__ llr, int64_t C, int r, float scale, int words)
{
    constexpr int Q = 1 << T, S = 1024;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * S) return;
    const int64_t cw = i >> 10;
    const int j = (int)(i & (S - 1));
    const float *row = y + cw * (int64_t)(Q * S);
    const uint32_t *xw = xs + cw * words;  // encoded decisions of this codeword's finished sub-blocks, 32 words each
    auto bit = [&](int blk) -> uint32_t { return ((xw[blk * 32 + (j >> 5)] >> (j & 31)) & 1u) << 31; };
    float o;
    if (T == 1) {
        const float v0 = scale * row[j], v1 = scale * row[j + S];  // polar.py:468-469
        o = (r == 0) ? npd_f_minsum(v0, v1) : __uint_as_float(__float_as_uint(v0) ^ bit(0)) + v1;
    } else {
        const float v0 = scale * row[j], v1 = scale * row[j + S], v2 = scale * row[j + 2 * S], v3 = scale * row[j + 3 * S];
        float a0, a1;  // level n-1 elements j and j + 1024
        if (r >= 2) {  // right half: g with the encoding of the left half's decisions = (x0 ^ x1, x1)
            const uint32_t b0 = bit(0), b1 = bit(1);
            a0 = __uint_as_float(__float_as_uint(v0) ^ (b0 ^ b1)) + v2;
            a1 = __uint_as_float(__float_as_uint(v1) ^ b1) + v3;
        } else {
            a0 = npd_f_minsum(v0, v2);
            a1 = npd_f_minsum(v1, v3);
        }
        o = (r & 1) ? __uint_as_float(__float_as_uint(a0) ^ bit(r - 1)) + a1 : npd_f_minsum(a0, a1);
    }
    llr[i] = o;
}

// decision words -> float decisions on the info positions (one warp per codeword, lane = k) + the NaN sentinel
__global__ void __launch_bounds__(256) split_out_kernel(const uint32_t *__restrict__ us, const unsigned char *__restrict__ flags,
                                                        const int32_t *__restrict__ info, float *__restrict__ decoded,
                                                        int64_t C, int K, int words)
{
    const int lane = threadIdx.x & 31;
    const int64_t cw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cw >= C) return;
    const uint32_t *w = us + cw * words;
    for (int k = lane; k < K; k += 32) {
        const int pos = __ldg(info + k);
        decoded[cw * K + k] = ((w[pos >> 5] >> (pos & 31)) & 1u) ? -1.0f : 1.0f;
    }
    __syncwarp();
    if (lane == 0 && flags[cw]) decoded[cw * K] = __int_as_float(0x7fc00000);  // exact re-decode (dispatch_group, scan mode)
}

// fused counting over the decision words (one warp per codeword); flagged codewords are left to the exact path
__global__ void __launch_bounds__(256) split_count_kernel(const uint32_t *__restrict__ us, const uint32_t *__restrict__ ubits,
                                                          const uint32_t *__restrict__ info_words,
                                                          const unsigned char *__restrict__ flags_in, unsigned char *__restrict__ flags_out,
                                                          int64_t C, int words, unsigned long long *counts)
{
    const int lane = threadIdx.x & 31;
    const int64_t cw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cw >= C) return;
    const bool fl = flags_in[cw] != 0;
    uint32_t e = 0u;
    for (int q = lane; q < words; q += 32) e += __popc((us[cw * words + q] ^ ubits[cw * words + q]) & __ldg(info_words + q));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(NPD_FULL, e, o);
    if (lane == 0) {
        flags_out[cw] = fl ? 1 : 0;
        if (!fl && e) {
            atomicAdd(counts + 0, (unsigned long long)e);
            atomicAdd(counts + 1, 1ull);
        }
    }
}

int launch_quad_split(const npd_code *code, ScParams p, cudaStream_t st)
{
    const int n = code->n, T = n - 10, Q = 1 << T, words = code->N >> 5;
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    cudaMemPool_t pool;
    if (int rc = scratch_pool(&pool)) return rc;
    // one chunk per 131072 codewords (0.5 GB of level-10 LLRs): the sub-block decode runs in rounds of sm_count x 12 warps x
    // 8 codewords, so large chunks keep the last, partly filled round's share small
    const int64_t chunk = p.B < 131072 ? p.B : 131072;
    float *llr = nullptr;
    uint32_t *us = nullptr, *xs = nullptr;
    unsigned char *flags = nullptr;
    NPD_CHECK_CUDA(cudaMallocFromPoolAsync((void **)&llr, (size_t)chunk * 1024 * sizeof(float), pool, st));
    NPD_CHECK_CUDA(cudaMallocFromPoolAsync((void **)&us, (size_t)chunk * words * 4, pool, st));
    NPD_CHECK_CUDA(cudaMallocFromPoolAsync((void **)&xs, (size_t)chunk * words * 4, pool, st));
    NPD_CHECK_CUDA(cudaMallocFromPoolAsync((void **)&flags, (size_t)chunk, pool, st));
    for (int64_t c0 = 0; c0 < p.B; c0 += chunk) {
        const int64_t C = p.B - c0 < chunk ? p.B - c0 : chunk;
        NPD_CHECK_CUDA(cudaMemsetAsync(flags, 0, (size_t)C, st));
        for (int r = 0; r < Q; ++r) {
            const unsigned grid = (unsigned)((C * 1024 + 255) / 256);
            if (T == 1) split_top_kernel<1><<<grid, 256, 0, st>>>(p.y + c0 * code->N, xs, llr, C, r, p.scale, words);
            else split_top_kernel<2><<<grid, 256, 0, st>>>(p.y + c0 * code->N, xs, llr, C, r, p.scale, words);
            NPD_CHECK_CUDA(cudaGetLastError());
            ScParams s{};
            s.y = llr; s.scale = 1.0f; s.infty = p.infty; s.B = C; s.n = 10; s.K = 0;
            s.frozen_words = p.frozen_words + r * 32;
            s.us_out = us; s.xs_out = xs; s.out_words = words; s.word_off = r * 32; s.flags = flags;
            if (int rc = launch_quad_n(code, 10, s, st)) return rc;
        }
        const unsigned wgrid = (unsigned)((C * 32 + 255) / 256);
        if (p.counts)
            split_count_kernel<<<wgrid, 256, 0, st>>>(us, p.ubits + c0 * words, p.info_words, flags, p.flags + c0, C, words, p.counts);
        else
            split_out_kernel<<<wgrid, 256, 0, st>>>(us, flags, p.info, p.decoded + c0 * p.K, C, p.K, words);
        NPD_CHECK_CUDA(cudaGetLastError());
    }
    NPD_CHECK_CUDA(cudaFreeAsync(llr, st));
    NPD_CHECK_CUDA(cudaFreeAsync(us, st));
    NPD_CHECK_CUDA(cudaFreeAsync(xs, st));
    NPD_CHECK_CUDA(cudaFreeAsync(flags, st));
    return NPD_OK;
}

int launch_quad(const npd_code *code, ScParams p, cudaStream_t st)
{
    // N = 4096: split path, 1.37 ms per 32768 codewords against 1.55 ms for the monolithic kernel with its global scratch
    // (NPD_SC_SPLIT=0 in a debug-knob build).  At N = 2048 the monolithic kernel still wins (1.19 vs 1.24 ms per 65536):
    // its two top-phase passes over y are cheap next to the split path's extra launches and round quantisation.
    if (code->n >= env_int("NPD_SC_SPLIT_MIN_N", 12) && env_int("NPD_SC_SPLIT", 1) != 0) return launch_quad_split(code, p, st);
    return launch_quad_n(code, code->n, p, st);
}

template <bool PAC, bool EXTRAS>
int dispatch_lane_blog(const npd_code *code, const ScParams &p, cudaStream_t st)
{
    switch (code->n) {
    case 1: return launch_lane<1, PAC, EXTRAS>(code, p, st);
    case 2: return launch_lane<2, PAC, EXTRAS>(code, p, st);
    case 3: return launch_lane<3, PAC, EXTRAS>(code, p, st);
    case 4: return launch_lane<4, PAC, EXTRAS>(code, p, st);
    default: return launch_lane<5, PAC, EXTRAS>(code, p, st);
    }
}

template <bool PAC>
int dispatch(const npd_code *code, ScParams p, cudaStream_t st)
{
    const bool use_lane = code->K >= 1 && env_int("NPD_SC_IMPL_GROUP", 0) == 0;
    if (!use_lane) {
        p.scan_flagged = 0;
        return dispatch_group<PAC>(code, p, st);
    }
    const bool extras = p.use_gt != nullptr || p.leaf_llr != nullptr;
    // decisions only, plain polar code, N >= 256: 4 lanes per codeword + simplified-SC pruning
    const bool use_quad = !PAC && !extras && code->n >= 8 && env_int("NPD_SC_IMPL_LANE", 0) == 0;
    int rc = use_quad ? launch_quad(code, p, st) : extras ? dispatch_lane_blog<PAC, true>(code, p, st) : dispatch_lane_blog<PAC, false>(code, p, st);
    if (rc) return rc;
    p.scan_flagged = 1;  // exact re-decode of codewords that hit sign(0) = 0
    return dispatch_group<PAC>(code, p, st);
}

// fused mode, after the exact re-decode: count the flagged codewords from their float decisions (a tie decision 0
// differs from +-1, utils.py:23) and add the chunk's frame count
__global__ void __launch_bounds__(256) count_flagged_kernel(const unsigned char *__restrict__ flags, const float *__restrict__ decoded,
                                                            const uint32_t *__restrict__ ubits, const int32_t *__restrict__ info,
                                                            int64_t B, int K, int NW, unsigned long long *counts,
                                                            unsigned long long add_frames)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t ch = warp; ch * 32 < B; ch += nwarps) {
        const int64_t cw = ch * 32 + lane;
        uint32_t mask = __ballot_sync(NPD_FULL, cw < B && flags[cw] != 0);
        while (mask) {
            const int64_t c = ch * 32 + (__ffs(mask) - 1);
            mask &= mask - 1;
            uint32_t e = 0u;
            for (int k = lane; k < K; k += 32) {
                const int pos = __ldg(info + k);
                const float want = ((ubits[c * NW + (pos >> 5)] >> (pos & 31)) & 1u) ? -1.0f : 1.0f;
                e += rintf(decoded[c * K + k]) != want;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(NPD_FULL, e, o);
            if (lane == 0 && e) {
                atomicAdd(counts + 0, (unsigned long long)e);
                atomicAdd(counts + 1, 1ull);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && add_frames) atomicAdd(counts + 2, add_frames);
}

}  // namespace

// internal (count_sweep.cu): SC-decode B codewords of a plain polar code with N >= 256 and accumulate
// counts[0..2] += (bit errors, block errors, B) against the transmitted u words -- no decisions are written except the
// exact re-decodes of flagged codewords into `decoded_scratch` [B,K]
int npd_sc_decode_count(const npd_code *code, const float *y, float llr_scale, const uint32_t *ubits, float *decoded_scratch,
                        unsigned char *flags, int64_t B, uint64_t *counts, cudaStream_t st)
{
    NPD_REQUIRE(code && y && ubits && decoded_scratch && flags && counts, "npd_sc_decode_count: null argument");
    NPD_REQUIRE(code->pac_g == 0 && code->n >= 8 && code->K >= 1, "npd_sc_decode_count: plain polar codes with N >= 256 only");
    if (B == 0) return NPD_OK;
    ScParams p{};
    p.y = y; p.decoded = decoded_scratch; p.info = code->d_info; p.frozen_words = code->d_frozen_words;
    p.B = B; p.n = code->n; p.K = code->K; p.scale = llr_scale; p.infty = code->infty;
    p.ubits = ubits; p.info_words = code->d_info_words; p.counts = (unsigned long long *)counts; p.flags = flags;
    if (int rc = launch_quad(code, p, st)) return rc;
    p.scan_flagged = 1;
    if (int rc = dispatch_group<false>(code, p, st)) return rc;
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;
    int64_t grid = (B + 255) / 256;
    if (grid > dp.sm_count * 4) grid = dp.sm_count * 4;
    count_flagged_kernel<<<(unsigned)grid, 256, 0, st>>>(flags, decoded_scratch, ubits, code->d_info, B, code->K, code->N >> 5,
                                                          (unsigned long long *)counts, (unsigned long long)B);
    NPD_CHECK_CUDA(cudaGetLastError());
    return NPD_OK;
}

// codewords one full round of the persistent decisions-only kernel decodes on this device (0: the code runs on the lane
// kernel, whose blocks are scheduled dynamically)
NPD_API int64_t npd_sc_round_codewords(const npd_code_t *code)
{
    if (!code || code->pac_g != 0 || code->n < 8) return 0;
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return 0;
    QuadResidency qr;
    // N = 4096 decodes as 1024-leaf sub-blocks (launch_quad_split)
    if (quad_residency(code, code->n >= 12 ? 10 : code->n, dp, &qr)) return 0;
    return (int64_t)dp.sm_count * qr.blocks_per_sm * qr.wpb * 8;
}

NPD_API int npd_sc_decode(const npd_code_t *code, const float *y, float llr_scale,
                          const float *use_gt, float *leaf_llr, float *decoded, int64_t B,
                          void *stream)
{
    NPD_REQUIRE(code && y && (decoded || code->K == 0), "npd_sc_decode: null argument");
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
