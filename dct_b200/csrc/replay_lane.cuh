// replay_lane.cuh -- the one-lane-per-block replay of flagged blocks, as device functions.
//
// K1 / K2 hand every block whose fp32 result sits inside the error band of a .5 rounding boundary to this code, which
// repeats what the reference does (dct_forward src/dct.c:57-74, quantize src/quantization.c:113-131, dequantize
// :133-151, dct_inverse src/dct.c:85-102) with the HOST-computed tables and non-contracted fp64 operations.
// The functions are called from two places: the stand-alone replay kernels of replay_f64.cu (worklists of the cp.async /
// one-shot kernels, float tiles) and the TAILS of the bulk-tensor kernels, where every warp replays the blocks of
// its own worklist segment as soon as its tile loop is done -- no separate launch, no worklist round trip through
// another grid, which is most of a single frame's latency and ~4 % of a large batch's time.
// They are not inlined: their registers are their own and do not weigh on the tile loops that call them.
#pragma once
#include "fast_core.cuh"
#include "kernels.cuh"

namespace dctb {

// natural index -> zigzag position (inverse of the scan), for run-time indexing on the device
struct ZigZagInv {
    int pos[64];
    constexpr ZigZagInv() : pos{}
    {
        ZigZag z{};
        for (int p = 0; p < 64; ++p) pos[z.nat[p]] = p;
    }
};
__device__ __constant__ const ZigZagInv cZigZagInv{};

// C99 round(): half away from zero.  (y - trunc(y)) is exact for |y| < 2^52.
__device__ __forceinline__ double round_half_away(double y)
{
    const double t = trunc(y);
    return (fabs(__dsub_rn(y, t)) >= 0.5) ? __dadd_rn(t, copysign(1.0, y)) : t;
}

__device__ __forceinline__ bool near_half(double a)
{
    a = fabs(a);
    const double f = __dsub_rn(a, floor(a));
    return fabs(__dsub_rn(f, 0.5)) <= 1e-9;
}

// src/quantization.c:186: fmin(1.0, fmax(0.1, variance / 1000.0))
__device__ __forceinline__ double norm_variance(double variance)
{
    return fmin(1.0, fmax(0.1, __ddiv_rn(variance, 1000.0)));
}

__device__ __forceinline__ double byte_centered(uint2 raw, int m)
{
    const unsigned w = m < 4 ? raw.x : raw.y;
    return __dsub_rn((double)((w >> (8 * (m & 3))) & 0xFFu), 128.0);   // src/dct.c:115
}

constexpr int kPairsPerRound = 4;                  // (block, value) pairs one lane may append per round

// per-warp scratch of the lane replays (3 712 bytes; the bulk-tensor kernels lend their idle record stage)
struct alignas(16) LaneScratch {
    uint2 px[32][9];                               // [block][row + pad]: the blocks' pixels for the replay phase
    int16_t *rec[32];                              // each block's record,
    const double *Q[32];                           // its plan's quant_matrix
    Counters *ctr[32];                             // and its plan's counters: the blocks of one pass may belong to several planes
    unsigned ties[32];                             // near-ties found in each block by the lanes that replayed its values
    double scale[32];                              // adaptive: 2 - nv (forward) or 1 / (2 - nv) (inverse) of each block
    unsigned short pairs[32 * kPairsPerRound];     // (source lane << 6) | natural index
};

// The lanes' counts `mine` added to their own counters, one atomic per distinct set of counters in the warp
__device__ __forceinline__ void add_near_ties(Counters *ctr, unsigned mine)
{
    const unsigned grp = __match_any_sync(0xffffffffu, (unsigned long long)ctr);
    const unsigned sum = __reduce_add_sync(grp, mine);
    if ((threadIdx.x & 31) == (unsigned)__ffs(grp) - 1u && sum != 0) atomicAdd(&ctr->near_ties, (unsigned long long)sum);
}

// Entry g of a list made of `n_items` runs of cnt[0], cnt[1], ... entries: which run, and where in it (false: past the end)
__device__ __forceinline__ bool locate_entry(const uint32_t *cnt, int n_items, uint32_t g, int &item, uint32_t &e)
{
    uint32_t base = 0;
    for (int it = 0; it < n_items; ++it) {
        const uint32_t c = cnt[it];
        if (g - base < c) {
            item = it, e = g - base;
            return true;
        }
        base += c;
    }
    return false;
}

// where the forward replay finds its tables and its output (any address space behind the pointers)
struct FwdReplayCtx {
    const float *r32, *thr32;                      // K1's multipliers and thresholds, natural index
    const double *D, *Q;                           // dct_matrix, quant_matrix as the host made them
    int adaptive;
    int16_t *coef;
    Counters *ctr;
};

// One warp, up to 32 flagged blocks, one per lane (`active`, block index `b`, its 64 pixels at src + i * src_pitch).
//   1. RE-FLAG: the lane repeats K1's fp32 arithmetic for its whole block in registers (the same fast_core.cuh
//      functions, the same operation order per element, hence the same bits) to find WHICH coefficients sit inside
//      the band; everything else in the block was already written correctly by K1 and is left alone;
//   2. the warp compacts the flagged (block, coefficient) pairs into a list and every lane replays ONE value on its
//      own: the reference's 64 + 8 non-contracted fp64 multiply-adds in its own order (src/dct.c:57-74), true
//      division, half-away rounding (src/quantization.c:122-126), and patches it into the record.
// Adds the near-tie / saturation counts to cx.ctr.  Every lane of the warp must call it.  The context may differ from
// lane to lane (blocks of several planes, each with its own tables, in one pass): what phase 2 needs of another lane's
// block goes through the scratch.  D (the 8x8 DCT matrix) and `adaptive` are the same for all lanes.
template <int LAYOUT>
__device__ __noinline__ void replay_fwd_lanes(const FwdReplayCtx cx, LaneScratch *ws, bool active, unsigned b, const uint8_t *src,
                                              long long src_pitch)
{
    const int lane = threadIdx.x & 31;
    uint2 row[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        row[i] = active ? *reinterpret_cast<const uint2 *>(src + i * src_pitch) : make_uint2(0x80808080u, 0x80808080u);

    // ---- phase 1: K1's fp32 arithmetic for the whole block (rows, then columns), find the flagged coefficients
    float c[64];
#pragma unroll
    for (int i = 0; i < 8; ++i) fdct8_row_from_bytes(&c[8 * i], row[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) fdct8<float, 8>(&c[j]);          // c[8u + j] = scaled coefficient (u, j)
    // adaptive tables: the block's variance from exact integer moments, as in K1 (src/quantization.c:153-190)
    float inv_s = 1.0f;
    double scale = 1.0;
    if (cx.adaptive) {
        int isum = 0, isq = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) row_moments(row[i], isum, isq);
        const double mean = __ddiv_rn((double)isum, 64.0);
        const double var = __dsub_rn(__ddiv_rn((double)isq, 64.0), __dmul_rn(mean, mean));
        scale = __dsub_rn(2.0, norm_variance(var));   // src/quantization.c:190
        inv_s = adaptive_inv_scale(64 * isq - isum * isum);
    }
    unsigned need_lo = 0, need_hi = 0;
#pragma unroll
    for (int k = 0; k < 64; ++k) {
        float t, e;
        const float ck = k != 0 ? __fmul_rn(c[k], inv_s) : c[k];   // as fwd_block; inv_s == 1.0f exactly when the table is not adaptive
        quant_residual(ck, cx.r32[k], t, e);
        if (fabsf(e) >= cx.thr32[k]) (k < 32 ? need_lo : need_hi) |= 1u << (k & 31);
    }
    if (!active) need_lo = need_hi = 0;
    __syncwarp();                                      // the scratch may still be read by the previous call's last round
    if (cx.adaptive) ws->scale[lane] = scale;
#pragma unroll
    for (int i = 0; i < 8; ++i) ws->px[lane][i] = row[i];
    ws->rec[lane] = cx.coef + (size_t)b * 64;
    ws->Q[lane] = cx.Q;
    ws->ctr[lane] = cx.ctr;
    ws->ties[lane] = 0;

    // ---- phase 2: rounds of (compact the pairs, one lane replays one value) until no lane has any left
    while (__any_sync(0xffffffffu, (need_lo | need_hi) != 0)) {
        const int mine = min(__popc(need_lo) + __popc(need_hi), kPairsPerRound);
        int before = mine;                                       // inclusive prefix sum over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, before, d);
            if (lane >= d) before += v;
        }
        const int total = __shfl_sync(0xffffffffu, before, 31);
        before -= mine;
        for (int n = 0; n < mine; ++n) {
            int k;
            if (need_lo) {
                k = __ffs(need_lo) - 1;
                need_lo &= need_lo - 1;
            } else {
                k = 32 + __ffs(need_hi) - 1;
                need_hi &= need_hi - 1;
            }
            ws->pairs[before + n] = (unsigned short)((lane << 6) | k);
        }
        __syncwarp();
        for (int pi = lane; pi < total; pi += 32) {
            const unsigned pr = ws->pairs[pi];
            const int sl = pr >> 6, k = pr & 63, i = k >> 3, j = k & 7;
            const double *Dj = &cx.D[j * 8], *Di = &cx.D[i * 8];
            double dj[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) dj[m] = Dj[m];
            double out = 0.0;    // out[i][j] = sum_kk D[i][kk] * temp[kk][j]  (src/dct.c:67-74), kk ascending
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint2 raw = ws->px[sl][kk];
                double temp = 0.0;   // temp[kk][j] = sum_m X[kk][m] * D[j][m]  (src/dct.c:57-64), m ascending
#pragma unroll
                for (int m = 0; m < 8; ++m) temp = __dadd_rn(temp, __dmul_rn(byte_centered(raw, m), dj[m]));
                out = __dadd_rn(out, __dmul_rn(Di[kk], temp));
            }
            double mq = ws->Q[sl][k];
            if (cx.adaptive && k != 0) {                          // src/quantization.c:196-204
                mq = __dmul_rn(mq, ws->scale[sl]);
                if (mq < 1.0) mq = 1.0;
            }
            const double y = __ddiv_rn(out, mq);                 // src/quantization.c:124
            const double rr = round_half_away(y);
            int q = (int)rr;
            if (rr > 32767.0 || rr < -32768.0) {                 // exotic tables only: counted where it happens
                q = rr > 0.0 ? 32767 : -32768;
                atomicAdd(&ws->ctr[sl]->saturated, 1ull);
            }
            if (near_half(y)) atomicAdd(&ws->ties[sl], 1u);      // integer pixels make exact ties common (DC = sum / 8)
            const int pos = LAYOUT == LAYOUT_ZIGZAG ? cZigZagInv.pos[k] : k;
            ws->rec[sl][pos] = (int16_t)q;
        }
        __syncwarp();
    }
    add_near_ties(cx.ctr, ws->ties[lane]);
}


// ------------------------------------------------------------------------------------------------------------------
// inverse: dequantize (src/quantization.c:133-151) and dct_inverse (src/dct.c:80-105), one lane per flagged block
// ------------------------------------------------------------------------------------------------------------------
struct alignas(16) InvLaneScratch {
    const int16_t *rec[32];                        // each block's record,
    uint8_t *dst[32];                              // its first pixel and the pitch of its plane,
    long long pitch[32];
    const double *R[32];                           // and its plan's dequant_matrix: the blocks of one pass may belong to several planes
    unsigned ties[32];                             // near-ties found in each block by the lanes that replayed its pixels
    double inv_s[32];                              // adaptive: 1 / (2 - nv) of each block (src/quantization.c:193)
    double s[32];                                  // adaptive: 2 - nv
    float bound[32];                               // the block's fp32 bound (sum gain_k |v_k|)
    unsigned short pairs[32 * kPairsPerRound];     // (source lane << 6) | pixel index 8i + j
};

struct InvReplayCtx {
    const float *rs32, *rg32;                      // K2's fp32 tables, natural index
    float band_floor;
    const double *D, *R;                           // dct_matrix, dequant_matrix as the host made them
    const double *mult;                            // adaptive: 1 / R in fp64 (ExactTables::mult64), for the fast settle
    const int16_t *coef;
    const double *var;                             // adaptive: per-block variance (may be null)
    uint8_t *px;
    long long pitch;
    uint32_t bw;
    Counters *ctr;
};

// Sample (i, j) of one block in the reference's own operation order:
// in[m][k] = the reference's dequantised value (src/quantization.c:133-151), temp[i][k] = sum_m D[m][i] * in[m][k] (src/dct.c:85-92),
// out = sum_k temp[i][k] * D[k][j] (:95-102), every sum from 0.0 in ascending order, no contraction.
// The eight sums temp[i][0..7] advance together, m ascending for each of them (the reference's order per sum): eight
// independent chains, and only row m of the coefficients is live at a time.  Not inlined: its registers are its own.
template <int LAYOUT, bool ADAPTIVE>
__device__ __noinline__ double exact_inverse_sample(const uint4 *q4, const double *D, const double *R, double inv_two_minus_nv, int i, int j)
{
    uint32_t w[32];       // the record: one 128-byte line, eight 16-byte loads, indexed statically below
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 t = q4[c];
        w[4 * c] = t.x, w[4 * c + 1] = t.y, w[4 * c + 2] = t.z, w[4 * c + 3] = t.w;
    }
    double di[8], temp[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) di[m] = D[m * 8 + i], temp[m] = 0.0;
    static_for<0, 8>([&](auto M) {
        constexpr int m = decltype(M)::value;
        static_for<0, 8>([&](auto K) {
            constexpr int k = decltype(K)::value;
            constexpr int nat = 8 * m + k;
            constexpr int pos = LAYOUT == LAYOUT_ZIGZAG ? ZigZagInv{}.pos[nat] : nat;
            const int qq = (int)(int16_t)(pos & 1 ? (w[pos >> 1] >> 16) : (w[pos >> 1] & 0xFFFFu));
            // (double)qq without a conversion instruction: (2^52 + 2^31 + qq) - (2^52 + 2^31), exact
            const double qd = __dsub_rn(__hiloint2double(0x43300000, (int)((unsigned)qq ^ 0x80000000u)), 4503601774854144.0);
            double in;
            if constexpr (ADAPTIVE) {      // q * (1.0 / (R * (1/(2-nv)))), DC unscaled (src/quantization.c:137,144,193-201)
                double mm = R[nat];
                if (nat != 0) mm = __dmul_rn(mm, inv_two_minus_nv);
                in = __dmul_rn(qd, __ddiv_rn(1.0, mm));
            } else {
                in = __dmul_rn(qd, R[nat]);    // q * R (src/quantization.c:139,144)
            }
            temp[k] = __dadd_rn(temp[k], __dmul_rn(di[m], in));
        });
    });
    double out = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) out = __dadd_rn(out, __dmul_rn(temp[k], D[k * 8 + j]));
    return out;
}


// Sample (i, j) of one block in plain fp64 (fused multiply-adds, any order): sum_uv q_uv * mult_uv * s_uv * D[u][i] * D[v][j].
// Within ~1e-13 of the reference's value, at an eighth of the cost of its reciprocal chain (src/quantization.c:137,144
// divides once per coefficient): settles every flagged pixel of an adaptive plan that is not a true near-tie.
template <int LAYOUT>
__device__ __noinline__ double fast_inverse_sample(const uint4 *q4, const double *D, const double *mult, double s, int i, int j)
{
    uint32_t w[32];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 t = q4[c];
        w[4 * c] = t.x, w[4 * c + 1] = t.y, w[4 * c + 2] = t.z, w[4 * c + 3] = t.w;
    }
    double di[8], temp[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) di[m] = D[m * 8 + i], temp[m] = 0.0;
    static_for<0, 8>([&](auto M) {
        constexpr int m = decltype(M)::value;
        static_for<0, 8>([&](auto K) {
            constexpr int k = decltype(K)::value;
            constexpr int nat = 8 * m + k;
            constexpr int pos = LAYOUT == LAYOUT_ZIGZAG ? ZigZagInv{}.pos[nat] : nat;
            const int qq = (int)(int16_t)(pos & 1 ? (w[pos >> 1] >> 16) : (w[pos >> 1] & 0xFFFFu));
            const double qd = __hiloint2double(0x43300000, (int)((unsigned)qq ^ 0x80000000u)) - 4503601774854144.0;
            double c = qd * mult[nat];
            if (nat != 0) c *= s;
            temp[k] = fma(di[m], c, temp[k]);
        });
    });
    double out = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) out = fma(temp[k], D[k * 8 + j], out);
    return out;
}

// One warp, up to 32 flagged blocks, one per lane.
//   1. RE-FLAG: the lane repeats K2's fp32 arithmetic for its whole block in registers (the same fast_core.cuh /
//      butterfly.cuh functions, scalar instantiation: bit-identical per element to K2's packed lanes) to find WHICH
//      pixels sit inside the band -- with a wide dynamic band (high-quality tables) that is typically ONE pixel of
//      the block; every other pixel is already right in memory and is not touched;
//   2. the warp compacts the flagged (block, pixel) pairs into a list and every lane replays ONE pixel on its own,
//      in the reference's own operation order (exact_inverse_sample), then the pixel rule; one byte store.
// Adds the near-tie count to cx.ctr.  Every lane of the warp must call it.  As in replay_fwd_lanes the context may
// differ from lane to lane (non-adaptive plans; D is the same for all).
template <int LAYOUT, bool ADAPTIVE>
__device__ __noinline__ void replay_inv_lanes(const InvReplayCtx cx, InvLaneScratch *ws, bool active, unsigned b)
{
    const int lane = threadIdx.x & 31;
    const int16_t *rec = cx.coef + (size_t)b * 64;
    float s32 = 1.0f;
    __syncwarp();                                      // the scratch may still be read by the previous call's last round
    if constexpr (ADAPTIVE) {
        const double var = (cx.var && active) ? cx.var[b] : 0.0;
        const double two_minus_nv = __dsub_rn(2.0, norm_variance(var));
        ws->s[lane] = two_minus_nv;
        ws->inv_s[lane] = __ddiv_rn(1.0, two_minus_nv);
        s32 = adaptive_scale(var);
    }
    {
        const unsigned by = b / cx.bw, bx = b - by * cx.bw;
        ws->rec[lane] = rec;
        ws->dst[lane] = cx.px + (long long)by * 8 * cx.pitch + (long long)bx * 8;
        ws->pitch[lane] = cx.pitch;
        ws->R[lane] = cx.R;
        ws->ties[lane] = 0;
    }

    // ---- phase 1: K2's fp32 arithmetic for the whole block (inv_block of dequant_idct.cu, scalar), flagged pixels
    float v[64];
    float bound = 0.f, bound_dc = 0.f;
    static_for<0, 8>([&](auto J) {
        constexpr int j = decltype(J)::value;
        const uint4 t = active ? reinterpret_cast<const uint4 *>(rec)[j] : make_uint4(0, 0, 0, 0);
        const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
        static_for<0, 4>([&](auto Hh) {
            constexpr int h = decltype(Hh)::value;
            constexpr int k0 = storage_to_natural<LAYOUT>(8 * j + 2 * h), k1 = storage_to_natural<LAYOUT>(8 * j + 2 * h + 1);
            const float f0 = half_to_float<0>(w4[h]), f1 = half_to_float<1>(w4[h]);
            v[k0] = f0, v[k1] = f1;
            if (ADAPTIVE && k0 == 0) bound_dc = __fmul_rn(fabsf(f0), cx.rg32[0]);
            else bound = __fmaf_rn(fabsf(f0), cx.rg32[k0], bound);
            if (ADAPTIVE && k1 == 0) bound_dc = __fmul_rn(fabsf(f1), cx.rg32[0]);
            else bound = __fmaf_rn(fabsf(f1), cx.rg32[k1], bound);
        });
    });
    if constexpr (ADAPTIVE) bound = adaptive_bound(bound, bound_dc, s32);
    {
        // K2's folded first stage (idct8_dequant), scalar: same operations on the same operands; adaptive plans scale
        // the 63 AC values by the block's (2 - nv) first
        constexpr int ra[4] = {0, 2, 5, 1}, rb[4] = {4, 6, 3, 7};
        if constexpr (ADAPTIVE) {
#pragma unroll
            for (int k = 1; k < 64; ++k) v[k] = __fmul_rn(v[k], s32);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float ma[4];
            PosNeg1 mb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ma[j] = cx.rs32[8 * ra[j] + c];
                mb[j].pos = cx.rs32[8 * rb[j] + c];
                mb[j].neg = -mb[j].pos;
            }
            idct8_dequant<float, 8>(&v[c], ma, mb);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) idct8<float, 1>(&v[8 * i]);
    const float thr = pixel_threshold(bound, cx.band_floor);
    if constexpr (ADAPTIVE) ws->bound[lane] = bound;
    unsigned need_lo = 0, need_hi = 0;
    if (!(bound < 1.4e5f)) {
        need_lo = need_hi = 0xffffffffu;               // outside the int16 trick's range: K2 wrote nothing reliable
    } else {
#pragma unroll
        for (int e = 0; e < 64; ++e) {
            float t, r;
            pixel_residual(v[e], t, r);
            if (fabsf(r) >= thr) (e < 32 ? need_lo : need_hi) |= 1u << (e & 31);
        }
    }
    if (!active) need_lo = need_hi = 0;

    // ---- phase 2: rounds of (compact the pairs, one lane replays one pixel) until no lane has any left
    while (__any_sync(0xffffffffu, (need_lo | need_hi) != 0)) {
        const int mine = min(__popc(need_lo) + __popc(need_hi), kPairsPerRound);
        int before = mine;                                       // inclusive prefix sum over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, before, d);
            if (lane >= d) before += x;
        }
        const int total = __shfl_sync(0xffffffffu, before, 31);
        before -= mine;
        for (int n = 0; n < mine; ++n) {
            int e;
            if (need_lo) {
                e = __ffs(need_lo) - 1;
                need_lo &= need_lo - 1;
            } else {
                e = 32 + __ffs(need_hi) - 1;
                need_hi &= need_hi - 1;
            }
            ws->pairs[before + n] = (unsigned short)((lane << 6) | e);
        }
        __syncwarp();
        for (int pi = lane; pi < total; pi += 32) {
            const unsigned pr = ws->pairs[pi];
            const int sl = pr >> 6, e = pr & 63, i = e >> 3, j = e & 7;
            const uint4 *q4 = reinterpret_cast<const uint4 *>(ws->rec[sl]);
            const double inv_two_minus_nv = ADAPTIVE ? ws->inv_s[sl] : 1.0;
            uint8_t *px = ws->dst[sl] + i * ws->pitch[sl] + j;
            if constexpr (ADAPTIVE) {
                // plain fp64 first: |fast - reference| <= 1e-14 * bound (64 products, each within 2^-52 of terms that
                // sum gain_k |v_k| dominates); farther than that (+ the 1e-9 tie-accounting margin) from a boundary
                // the pixel is settled.  Adaptive plans flag two or three pixels per block on busy content.
                const double fast = fast_inverse_sample<LAYOUT>(q4, cx.D, cx.mult, ws->s[sl], i, j) + 128.0;
                const double n = rint(fast);
                if (fabs(fast - n) < 0.5 - (2e-9 + 1e-14 * (double)ws->bound[sl]) && fabs(fast) < 1e9) {
                    *px = (uint8_t)(n < 0.0 ? 0.0 : (n > 255.0 ? 255.0 : n));
                    continue;
                }
            }
            const double out = exact_inverse_sample<LAYOUT, ADAPTIVE>(q4, cx.D, ws->R[sl], inv_two_minus_nv, i, j);
            const double val = __dadd_rn(out, 128.0);
            double rr = round_half_away(val);
            rr = rr < 0.0 ? 0.0 : (rr > 255.0 ? 255.0 : rr);
            if (near_half(val)) atomicAdd(&ws->ties[sl], 1u);
            *px = (uint8_t)rr;
        }
        __syncwarp();
    }
    add_near_ties(cx.ctr, ws->ties[lane]);
}

}  // namespace dctb
