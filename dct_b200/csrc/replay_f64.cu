// replay_f64.cu -- K3 / K4 / K6: the reference's arithmetic, operation for operation, in fp64.
//
// K1/K2 hand every block whose fp32 result sits inside the error band of a .5 rounding
// boundary to these kernels.  They repeat exactly what the reference does:
//   dct_forward   src/dct.c:57-74    temp = X*D^T, out = D*temp, acc = 0.0 then += for k ascending
//   quantize      src/quantization.c:113-131   (int) round(c / M), true division, half away
//   dequantize    src/quantization.c:133-151   q * R   |   q * (1.0 / (R * (1.0/(2-nv))))
//   dct_inverse   src/dct.c:85-102   temp = D^T*in, out = temp*D
//   adjust_matrix_for_block src/quantization.c:171-211
// using the HOST-computed tables (glibc's cos() noise is part of the answer)
// and non-contracted __dmul_rn/__dadd_rn/__ddiv_rn, so the integers are bit-identical to
// the reference's, exact ties included.  The same code, generalised to n x n, backs the
// per-block drop-in entry points dct_forward/dct_inverse/quantize/dequantize.
// The one-lane-per-block replay bodies live in replay_lane.cuh (they are also called from the tails of K1 / K2);
// this file holds the stand-alone kernels around them, the 8-threads-per-block kernels for float tiles and for
// tables outside the fast path's domain (everything replayed), and the per-block kernels.
#include <cstring>

#include "fast_core.cuh"
#include "kernels.cuh"
#include "replay_lane.cuh"

namespace dctb {

namespace {

constexpr int kReplayThreads = 256;                // 8 warps; a warp holds 4 blocks, 8 threads each

// K3 visits the flagged blocks only, 8 threads per block (thread r owns row r, then column r,
// exchanged through a warp-private shared-memory tile), in two phases:
//   1. RE-FLAG: repeat K1's / K2's fp32 arithmetic bit for bit (same functions from fast_core.cuh,
//      same inputs) to find WHICH values of the block sit inside the error band.  Everything else
//      in the block was already written correctly by K1 / K2 and is left alone.
//   2. REPLAY each such value -- typically one per block, a mathematically exact tie (SURVEY.md
//      S6) -- in the reference's own operation order: 64 + 8 non-contracted fp64 multiply-adds
//      (each of the 8 threads does one inner sum, one thread adds the 8 products in ascending
//      order), true division, half-away rounding; patch it into the output.
// With worklist == null (tables outside the fast path's domain; K1/K2 skipped) every value of
// every block is replayed.

__device__ __forceinline__ double shfl_double(double v, int src_lane)
{
    return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src_lane),
                            __shfl_sync(0xffffffffu, __double2loint(v), src_lane));
}

// 64-bit OR across the 8 lanes of a group
__device__ __forceinline__ unsigned long long group_or(unsigned long long m)
{
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)m, d);
        const unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(m >> 32), d);
        m |= ((unsigned long long)hi << 32) | lo;
    }
    return m;
}

struct ReplayShared {
    ExactTables tab;
    alignas(16) double tile[kReplayThreads / 32][4][8][9];  // [warp][block][row][col + pad]; reused as float / int16
};

// one 8-pixel row of a block, as the fused kernel read it: 8 bytes, or 8 floats (float pixel tiles)
template <bool F32> struct PixelRow;
template <> struct PixelRow<false> {
    uint2 raw;
    __device__ __forceinline__ void load(const uint8_t *base, long long pitch, unsigned by, unsigned bx, int r)
    {
        raw = *reinterpret_cast<const uint2 *>(base + ((long long)by * 8 + r) * pitch + (long long)bx * 8);
    }
    __device__ __forceinline__ double exact(int m) const { return byte_centered(raw, m); }
};
template <> struct PixelRow<true> {
    float4 lo, hi;
    __device__ __forceinline__ void load(const uint8_t *base, long long pitch, unsigned by, unsigned bx, int r)
    {
        const float4 *row = reinterpret_cast<const float4 *>(base + ((long long)by * 8 + r) * pitch + (long long)bx * 32);
        lo = row[0], hi = row[1];
    }
    __device__ __forceinline__ float at(int m) const
    {
        return m == 0 ? lo.x : m == 1 ? lo.y : m == 2 ? lo.z : m == 3 ? lo.w : m == 4 ? hi.x : m == 5 ? hi.y : m == 6 ? hi.z : hi.w;
    }
    __device__ __forceinline__ double exact(int m) const { return __dsub_rn((double)at(m), 128.0); }   // src/dct.c:115 on a float
};

template <int LAYOUT, bool F32>
__global__ void __launch_bounds__(kReplayThreads, 4) k_replay_fwd(const ReplayParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ReplayShared &sh = *reinterpret_cast<ReplayShared *>(smem_raw);
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(p.tab);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sh.tab);
        for (int i = threadIdx.x; i < (int)(sizeof(ExactTables) / 4); i += kReplayThreads) dst[i] = src[i];
    }
    __syncthreads();
    const ExactTables &tab = sh.tab;

    unsigned count = p.nblocks;
    if (p.worklist != nullptr) {
        count = p.ctr->wl_count;
        if (count > p.wl_cap) count = p.wl_cap;
    }
    const bool replay_all = p.worklist == nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, r = lane & 7, gbase = lane & 24;
    float(*T)[9] = reinterpret_cast<float(*)[9]>(&sh.tile[warp][g][0][0]);
    unsigned ties = 0, sat = 0;
    const unsigned groups_per_grid = gridDim.x * (kReplayThreads / 8);

    // Warp-uniform trip count so that the shuffles below always see the whole warp.  Two dependent
    // global loads feed an iteration (worklist entry -> pixel row), so the pipeline is two deep: the
    // entry of iteration i+2 and the pixel row of iteration i+1 are requested before the arithmetic
    // of iteration i, and each has a whole iteration to arrive.
    auto fetch_entry = [&](unsigned base_, bool &act_, unsigned &b_) {
        const unsigned slot = base_ + g;
        act_ = slot < count;
        b_ = act_ ? (p.worklist ? p.worklist[slot] : slot) : 0;
    };
    auto fetch_row = [&](unsigned b_, PixelRow<F32> &raw_) {
        const unsigned by = b_ / p.bw, bx = b_ - by * p.bw;
        raw_.load(p.px_in, p.pitch, by, bx, r);
    };
    unsigned base = (blockIdx.x * (kReplayThreads / 32) + warp) * 4;
    bool active = false, active_n = false, active_nn = false;
    unsigned b = 0, b_n = 0, b_nn = 0;
    PixelRow<F32> raw{}, raw_n{};
    if (base < count) {
        fetch_entry(base, active, b);
        fetch_row(b, raw);
    }
    if (base + groups_per_grid < count) fetch_entry(base + groups_per_grid, active_n, b_n);
    for (; base < count; base += groups_per_grid, active = active_n, b = b_n, raw = raw_n, active_n = active_nn, b_n = b_nn) {
        if (base + groups_per_grid < count) fetch_row(b_n, raw_n);
        if (base + 2 * groups_per_grid < count && base + 2 * groups_per_grid > base) fetch_entry(base + 2 * groups_per_grid, active_nn, b_nn);

        // per-block variance (adaptive): exact integers, reduced over the 8 rows
        double scale = 1.0;
        float inv_s = 1.0f;
        if constexpr (!F32) if (p.adaptive) {
            int isum = 0, isq = 0;
            row_moments(raw.raw, isum, isq);
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                isum += __shfl_xor_sync(0xffffffffu, isum, d);
                isq += __shfl_xor_sync(0xffffffffu, isq, d);
            }
            // src/quantization.c:153-169: every partial sum there is an exact integer, so these three
            // fp64 operations see the reference's operands
            const double mean = __ddiv_rn((double)isum, 64.0);
            const double var = __dsub_rn(__ddiv_rn((double)isq, 64.0), __dmul_rn(mean, mean));
            if (r == 0 && active && p.var_out) p.var_out[b] = var;
            scale = __dsub_rn(2.0, norm_variance(var));   // src/quantization.c:190
            inv_s = adaptive_inv_scale(64 * isq - isum * isum);
        }

        // ---- phase 1: K1's fp32 arithmetic again (rows, then columns), to find the flagged coefficients
        unsigned long long need = ~0ull;
        if (!replay_all) {
            float x[8];
            bool whole = false;
            if constexpr (F32) {
                float amax = 0.0f;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    x[m] = centre_float_pixel(raw.at(m));
                    amax = fmaxf(amax, fabsf(x[m]));
                }
                whole = __any_sync(0xffu << gbase, !(amax <= 128.0f));   // K1 flagged the block as out of domain
                fdct8<float, 1>(x);
            } else {
                fdct8_row_from_bytes(x, raw.raw);
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) T[r][m] = x[m];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = T[i][r];
            __syncwarp();
            fdct8<float, 1>(x);                           // x[u] = scaled coefficient (u, r)
            need = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = 8 * u + r;
                const float xu = (p.adaptive && k != 0) ? __fmul_rn(x[u], inv_s) : x[u];   // as fwd_block (fwd_quant.cu)
                float t, e;
                quant_residual(xu, tab.r32[k], t, e);
                if (fabsf(e) >= (F32 ? tab.thr32f[k] : tab.thr32[k])) need |= 1ull << k;
            }
            need = group_or(need);            // full-warp shuffles: every lane must take part
            if (whole) need = ~0ull;
        }

        // ---- phase 2: exact replay of single coefficients, cooperatively by the group ---------
        while (__any_sync(0xffffffffu, need != 0)) {
            const bool mine = need != 0;
            const int k = mine ? __ffsll((long long)need) - 1 : 0;
            need &= need - 1;
            const int i = k >> 3, j = k & 7;
            double temp = 0.0;   // temp[r][j] = sum_m X[r][m] * D[j][m]      (src/dct.c:57-64)
#pragma unroll
            for (int m = 0; m < 8; ++m) temp = __dadd_rn(temp, __dmul_rn(raw.exact(m), tab.D[j * 8 + m]));
            const double prod = __dmul_rn(tab.D[i * 8 + r], temp);
            double out = 0.0;    // out[i][j] = sum_kk D[i][kk] * temp[kk][j]  (src/dct.c:67-74), kk ascending
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) out = __dadd_rn(out, shfl_double(prod, gbase + kk));
            if (mine && r == 0 && active) {
                double mq = tab.Q[k];
                if (p.adaptive && k != 0) {
                    mq = __dmul_rn(mq, scale);
                    if (mq < 1.0) mq = 1.0;
                }
                const double y = __ddiv_rn(out, mq);
                const double rr = round_half_away(y);
                int q = (int)rr;
                if (rr > 32767.0) q = 32767, ++sat;
                if (rr < -32768.0) q = -32768, ++sat;
                ties += near_half(y);
                const int pos = LAYOUT == LAYOUT_ZIGZAG ? cZigZagInv.pos[k] : k;
                p.coef_out[(size_t)b * 64 + pos] = (int16_t)q;
            }
        }
    }

    ties = __reduce_add_sync(0xffffffffu, ties);
    sat = __reduce_add_sync(0xffffffffu, sat);
    if (lane == 0) {
        if (ties) atomicAdd(&p.ctr->near_ties, (unsigned long long)ties);
        if (sat) atomicAdd(&p.ctr->saturated, (unsigned long long)sat);
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&p.ctr->replayed, (unsigned long long)count);
    // the last CTA to finish empties the worklist for the next K1/K2 on this lane (saves a memset launch);
    // every CTA has read wl_count long before it gets here
    __syncthreads();
    if (threadIdx.x == 0 && p.worklist != nullptr) {
        __threadfence();
        if (atomicAdd(&p.ctr->done_ctas, 1u) == gridDim.x - 1) {
            p.ctr->wl_count = 0;
            p.ctr->done_ctas = 0;
        }
    }
}

// ---- segmented worklists (K1 / K2 bulk-tensor kernels): one segment per warp of the producer's grid ------------------
// Exclusive prefix sums of the segments' counts into prefix[0 .. n_segs]; every thread of a kLaneThreads-wide CTA calls it.
constexpr int kLaneThreads = 128;
constexpr int kLaneWarps = kLaneThreads / 32;

__device__ __forceinline__ void build_segment_prefix(unsigned *prefix, const uint32_t *seg_count, unsigned n_segs)
{
    // counts -> shared memory (coalesced), per-thread sums of contiguous chunks, block-wide scan of the 128 sums,
    // then every thread turns its chunk into exclusive prefixes in place
    for (unsigned i = threadIdx.x; i < n_segs; i += kLaneThreads) prefix[i] = seg_count[i];
    __syncthreads();
    const unsigned per = (n_segs + kLaneThreads - 1) / kLaneThreads, first = threadIdx.x * per;
    const unsigned last = first + per < n_segs ? first + per : n_segs;
    unsigned sum = 0;
    for (unsigned i = first; i < last; ++i) sum += prefix[i];
    unsigned incl = sum;                              // inclusive scan over the CTA's 128 threads
    const int ln_ = threadIdx.x & 31, wp_ = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
        if (ln_ >= d) incl += v;
    }
    __shared__ unsigned warp_tot[kLaneWarps];
    if (ln_ == 31) warp_tot[wp_] = incl;
    __syncthreads();
    unsigned base = incl - sum;
    for (int w = 0; w < wp_; ++w) base += warp_tot[w];
    for (unsigned i = first; i < last; ++i) {
        const unsigned c = prefix[i];
        prefix[i] = base;
        base += c;
    }
    if (threadIdx.x == kLaneThreads - 1) prefix[n_segs] = base;   // the last thread's running sum is the total
    __syncthreads();
}

// binary search: the segment with prefix[seg] <= entry < prefix[seg + 1], and the entry's index inside it
__device__ __forceinline__ void find_segment(const unsigned *prefix, unsigned n_segs, unsigned entry, unsigned &seg, unsigned &local)
{
    unsigned lo = 0, hi = n_segs;
    while (hi - lo > 1) {
        const unsigned mid = (lo + hi) >> 1;
        if (prefix[mid] <= entry) lo = mid;
        else hi = mid;
    }
    seg = lo, local = entry - prefix[lo];
}

// ------------------------------------------------------------------------------------------
// K3 forward, common case (uint8 pixels, worklist mode; adaptive tables included): ONE LANE PER FLAGGED BLOCK.
// The 8-lanes-per-block kernel above spends ~125 warp instructions per block, most of them on the
// shared-memory transposes of its re-flagging phase and on a replay phase that keeps 8 lanes busy with
// one value.  Here a lane repeats K1's fp32 arithmetic for its whole block in registers (the same
// fast_core.cuh functions, the same operation order per element, hence the same bits), the warp
// compacts the flagged (block, coefficient) pairs into a list, and then every lane replays ONE
// value on its own: the reference's 64 + 8 non-contracted fp64 multiply-adds in its own order
// (src/dct.c:57-74), true division, half-away rounding (src/quantization.c:122-126).
// ------------------------------------------------------------------------------------------

struct LaneShared {
    ExactTables tab;
    LaneScratch scratch[kLaneWarps];
    unsigned prefix[kMaxWorklistSegments + 1];     // exclusive prefix sums of K1's per-warp worklist counts
};

template <int LAYOUT>
__global__ void __launch_bounds__(kLaneThreads, 4) k_replay_fwd_lane(const ReplayParams p)
{
    __shared__ LaneShared sh;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(p.tab);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sh.tab);
        for (int i = threadIdx.x; i < (int)(sizeof(ExactTables) / 4); i += kLaneThreads) dst[i] = src[i];
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();                 // the tables above are the plan's own; everything below is K1's output
    const ExactTables &tab = sh.tab;
    // K1 left one worklist segment per warp of its grid and the segments' counts; entry e of the concatenation lives
    // in the segment s with prefix[s] <= e < prefix[s + 1].  Every CTA builds the prefix sums for itself.
    const unsigned n_segs = p.seg.n_segs;
    build_segment_prefix(sh.prefix, p.seg_count, n_segs);
    const unsigned count = sh.prefix[n_segs];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned tiles = (count + 31) / 32, warps_per_grid = gridDim.x * kLaneWarps;
    const FwdReplayCtx cx{tab.r32, tab.thr32, tab.D, tab.Q, p.adaptive, p.coef_out, p.ctr};

    for (unsigned tile = blockIdx.x * kLaneWarps + warp; tile < tiles; tile += warps_per_grid) {
        const unsigned entry = tile * 32 + lane;
        const bool active = entry < count;
        unsigned seg = 0, local = 0;
        if (active) find_segment(sh.prefix, n_segs, entry, seg, local);
        const unsigned slot = local < p.seg.side_seg_cap ? seg * p.seg.side_seg_cap + local : 0xffffffffu;   // side slot
        const unsigned b = active ? p.worklist[(size_t)seg * p.seg.seg_cap + local] : 0;
        const unsigned by = b / p.bw, bx = b - by * p.bw;
        // K1 left the block's 64 pixels next to its worklist entry (8-byte rows back to back); later slots read the plane
        const bool side = slot != 0xffffffffu;
        const uint8_t *src = side ? p.side + (size_t)slot * 64 : p.px_in + (long long)by * 8 * p.pitch + (long long)bx * 8;
        replay_fwd_lanes<LAYOUT>(cx, &sh.scratch[warp], active, b, src, side ? 8 : p.pitch);
    }

    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&p.ctr->replayed, (unsigned long long)count);
    // nothing to reset: the next K1 overwrites every segment count
}

// the reference's dequantised value of natural index k (src/quantization.c:133-151)
__device__ __forceinline__ double exact_dequant(const ExactTables &tab, int adaptive, double inv_two_minus_nv, int k, int q)
{
    double m = tab.R[k];
    if (adaptive) {
        if (k != 0) m = __dmul_rn(m, inv_two_minus_nv);
        return __dmul_rn((double)q, __ddiv_rn(1.0, m));
    }
    return __dmul_rn((double)q, m);
}

template <int LAYOUT>
__global__ void __launch_bounds__(kReplayThreads, 4) k_replay_inv(const ReplayParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ReplayShared &sh = *reinterpret_cast<ReplayShared *>(smem_raw);
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(p.tab);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sh.tab);
        for (int i = threadIdx.x; i < (int)(sizeof(ExactTables) / 4); i += kReplayThreads) dst[i] = src[i];
    }
    __syncthreads();
    const ExactTables &tab = sh.tab;

    unsigned count = p.nblocks;
    if (p.worklist != nullptr) {
        count = p.ctr->wl_count;
        if (count > p.wl_cap) count = p.wl_cap;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, r = lane & 7, gbase = lane & 24;
    float(*T)[9] = reinterpret_cast<float(*)[9]>(&sh.tile[warp][g][0][0]);
    double(*Td)[9] = sh.tile[warp][g];
    unsigned ties = 0;
    const unsigned groups_per_grid = gridDim.x * (kReplayThreads / 8);

    for (unsigned base = (blockIdx.x * (kReplayThreads / 32) + warp) * 4; base < count; base += groups_per_grid) {
        const unsigned slot = base + g;
        const bool active = slot < count;
        const unsigned b = active ? (p.worklist ? p.worklist[slot] : slot) : 0;
        const unsigned by = b / p.bw, bx = b - by * p.bw;
        uint8_t *dst = p.px_out + (long long)by * 8 * p.pitch + (long long)bx * 8;

        double inv_two_minus_nv = 1.0, two_minus_nv = 1.0;
        if (p.adaptive) {
            const double var = p.var_in ? p.var_in[b] : 0.0;
            two_minus_nv = __dsub_rn(2.0, norm_variance(var));
            inv_two_minus_nv = __ddiv_rn(1.0, two_minus_nv);   // src/quantization.c:193
        }

        // the record goes through the tile so that thread r can pick column r in any layout
        int16_t *rec16 = reinterpret_cast<int16_t *>(&T[0][0]);
        reinterpret_cast<uint4 *>(rec16)[r] = active ? reinterpret_cast<const uint4 *>(p.coef_in + (size_t)b * 64)[r]
                                                     : make_uint4(0, 0, 0, 0);
        __syncwarp();
        int q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = 8 * i + r;
            q[i] = rec16[LAYOUT == LAYOUT_ZIGZAG ? cZigZagInv.pos[k] : k];
        }
        __syncwarp();

        // This kernel now serves only plans whose tables are outside the fast path's domain (K2 skipped, every pixel
        // of every block replayed); blocks flagged by K2 go through the one-lane-per-block kernel (replay_lane.cuh),
        // which is where K2's fp32 arithmetic is repeated to find the flagged pixels.
        unsigned long long need = ~0ull;

        // ---- phase 2 and 3 -------------------------------------------------------------------
        if (__any_sync(0xffffffffu, need != 0)) {
            // phase 2: the same butterfly in fp64.  Its error (~1e-13 * bound) is far below the fp32
            // band, so it settles every flagged pixel that is not within band64 of a boundary -- with a
            // loose dynamic fp32 bound (large coefficients, adaptive decode) that is nearly all of them.
            // Worth its ~100 fp64 instructions only when a block has several flagged pixels.
            if (__any_sync(0xffffffffu, __popcll(need) >= 3)) {
                double x[8];
                double bound64 = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = 8 * i + r;
                    // q * multiplier * (2-nv) instead of the reference's reciprocal chain: a few ulps apart,
                    // inside the 8 * 2^-53 input error the bound allows for
                    x[i] = (double)q[i] * tab.mp64[k];
                    if (k != 0) x[i] *= two_minus_nv;
                    bound64 = fma(fabs(x[i]), (double)tab.gain32[k], bound64);
                }
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) bound64 += shfl_double(bound64, lane ^ d);
                idct8<double, 1>(x);                      // column r
#pragma unroll
                for (int i = 0; i < 8; ++i) Td[i][r] = x[i];
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = Td[r][j];
                __syncwarp();
                idct8<double, 1>(x);                      // row r
                // 1e-9: the tie-accounting margin; 8 * 2^-53 * bound: fp64 butterfly + the reference's own rounding
                const double band64 = 2e-9 + bound64 * 8.9e-16;
                unsigned long long settled = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned long long bit = 1ull << (8 * r + j);
                    if (need & bit) {
                        const double val = x[j] + 128.0;
                        const double n = rint(val);
                        if (fabs(val - n) < 0.5 - band64 && fabs(val) < 1e9) {
                            if (active) dst[r * p.pitch + j] = (uint8_t)(int)fmin(fmax(n, 0.0), 255.0);
                            settled |= bit;
                        }
                    }
                }
                need &= ~group_or(settled);
            }
        }
        if (__any_sync(0xffffffffu, need != 0)) {
            double in[8];                                 // the reference's dequantised column r, exactly
#pragma unroll
            for (int m = 0; m < 8; ++m) in[m] = exact_dequant(tab, p.adaptive, inv_two_minus_nv, 8 * m + r, q[m]);

            // phase 3: exact replay of what is left (values within ~1e-9 of a .5 boundary)
            while (__any_sync(0xffffffffu, need != 0)) {
                const bool mine = need != 0;
                const int e = mine ? __ffsll((long long)need) - 1 : 0;
                need &= need - 1;
                const int i = e >> 3, j = e & 7;
                double temp = 0.0;   // temp[i][r] = sum_m D[m][i] * in[m][r]     (src/dct.c:85-92)
#pragma unroll
                for (int m = 0; m < 8; ++m) temp = __dadd_rn(temp, __dmul_rn(tab.D[m * 8 + i], in[m]));
                const double prod = __dmul_rn(temp, tab.D[r * 8 + j]);
                double out = 0.0;    // out[i][j] = sum_k temp[i][k] * D[k][j]  (src/dct.c:95-102), k ascending
#pragma unroll
                for (int k = 0; k < 8; ++k) out = __dadd_rn(out, shfl_double(prod, gbase + k));
                if (mine && r == 0 && active) {
                    const double val = __dadd_rn(out, 128.0);
                    double rr = round_half_away(val);
                    rr = rr < 0.0 ? 0.0 : (rr > 255.0 ? 255.0 : rr);
                    ties += near_half(val);
                    dst[i * p.pitch + j] = (uint8_t)rr;
                }
            }
        }
        __syncwarp();
    }

    ties = __reduce_add_sync(0xffffffffu, ties);
    if (lane == 0 && ties) atomicAdd(&p.ctr->near_ties, (unsigned long long)ties);
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&p.ctr->replayed, (unsigned long long)count);
    // the last CTA to finish empties the worklist for the next K1/K2 on this lane (saves a memset launch);
    // every CTA has read wl_count long before it gets here
    __syncthreads();
    if (threadIdx.x == 0 && p.worklist != nullptr) {
        __threadfence();
        if (atomicAdd(&p.ctr->done_ctas, 1u) == gridDim.x - 1) {
            p.ctr->wl_count = 0;
            p.ctr->done_ctas = 0;
        }
    }
}


// ------------------------------------------------------------------------------------------
// K3 inverse, worklist mode: ONE LANE PER FLAGGED BLOCK (the decoder's counterpart of k_replay_fwd_lane).
// The reference path being reproduced: dequantize (src/quantization.c:133-151) and dct_inverse (src/dct.c:80-105).
//   1. RE-FLAG: the lane repeats K2's fp32 arithmetic for its whole block in registers (the same fast_core.cuh /
//      butterfly.cuh functions, scalar instantiation: bit-identical per element to K2's packed lanes) to find WHICH
//      pixels sit inside the band -- with a wide dynamic band (high-quality tables) that is typically ONE pixel of
//      the block; every other pixel is already right in memory and is not touched;
//   2. the warp compacts the flagged (block, pixel) pairs into a list and every lane replays ONE pixel on its own,
//      in the reference's own operation order: 64 + 8 non-contracted multiply-adds on the reference's dequantised
//      values, ascending index, then the pixel rule; one byte store.
// Against the 8-lanes-per-block kernel: no shared-memory transposes, no shuffles, no lanes idling during the replay.
// (A first version redid the whole block with an fp64 butterfly per lane: 3 300 instructions per warp at 204
// registers, latency-bound at two warps per scheduler -- slower than the kernel it replaced.)
// ------------------------------------------------------------------------------------------
// The fp32 tables travel as kernel parameters (no table copy in the prologue); D and R, which the exact replay indexes
// per lane, go to shared memory.
struct InvLaneParams {
    ReplayParams p;
    float rs32[64], rg32[64];                      // K2's fp32 tables (ExactTables)
    float band_floor;
};

struct LaneInvShared {
    double D[64];                                  // dct_matrix: indexed by the replayed pixel's (i, j), per lane
    double R[64];                                  // dequant_matrix
    InvLaneScratch scratch[kLaneWarps];
    unsigned prefix[kMaxWorklistSegments + 1];
};

template <int LAYOUT, bool ADAPTIVE>
__global__ void __launch_bounds__(kLaneThreads, 5) k_replay_inv_lane(const __grid_constant__ InvLaneParams P)
{
    __shared__ LaneInvShared sh;
    const ReplayParams &p = P.p;
    if (threadIdx.x < 64) sh.D[threadIdx.x] = p.tab->D[threadIdx.x], sh.R[threadIdx.x] = p.tab->R[threadIdx.x];
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();                 // the tables above are the plan's own; everything below is K2's output
    // segmented worklist (bulk-tensor K2: one segment per warp of its grid) or a flat one counted in ctr->wl_count
    const bool segmented = p.seg_count != nullptr && p.seg.n_segs != 0;
    const unsigned n_segs = p.seg.n_segs;
    unsigned count;
    if (segmented) {
        build_segment_prefix(sh.prefix, p.seg_count, n_segs);
        count = sh.prefix[n_segs];
    } else {
        count = p.ctr->wl_count;
        if (count > p.wl_cap) count = p.wl_cap;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned tiles = (count + 31) / 32, warps_per_grid = gridDim.x * kLaneWarps;
    const InvReplayCtx cx{P.rs32, P.rg32, P.band_floor, sh.D, sh.R, p.tab->mult64, p.coef_in, p.var_in, p.px_out, p.pitch, p.bw, p.ctr};

    // worklist entry -> block index; the entry of the NEXT tile is looked up, and its record's line requested from
    // L2, one tile ahead (two dependent global loads would otherwise sit at the head of every tile)
    auto lookup = [&](unsigned tile_, bool &active_, unsigned &b_) {
        const unsigned entry = tile_ * 32 + lane;
        active_ = tile_ < tiles && entry < count;
        b_ = 0;
        if (active_) {
            if (segmented) {
                unsigned seg, local;
                find_segment(sh.prefix, n_segs, entry, seg, local);
                b_ = p.worklist[(size_t)seg * p.seg.seg_cap + local];
            } else {
                b_ = p.worklist[entry];
            }
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.coef_in + (size_t)b_ * 64));
        }
    };
    bool active_n;
    unsigned b_n;
    lookup(blockIdx.x * kLaneWarps + warp, active_n, b_n);
    for (unsigned tile = blockIdx.x * kLaneWarps + warp; tile < tiles; tile += warps_per_grid) {
        const bool active = active_n;
        const unsigned b = b_n;
        lookup(tile + warps_per_grid, active_n, b_n);
        replay_inv_lanes<LAYOUT, ADAPTIVE>(cx, &sh.scratch[warp], active, b);
    }

    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&p.ctr->replayed, (unsigned long long)count);
    if (!segmented) {
        // the last CTA to finish empties the flat worklist for the next K2 on this lane (saves a memset launch);
        // every CTA has read wl_count long before it gets here
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&p.ctr->done_ctas, 1u) == gridDim.x - 1) {
                p.ctr->wl_count = 0;
                p.ctr->done_ctas = 0;
            }
        }
    }
}

// ---- generic n x n single-block kernels (n <= 32): the per-block drop-in API -------------

__global__ void k_block_dct_f64(int n, const double *D, const double *in, double *out, int inverse)
{
    extern __shared__ double sm[];
    double *sD = sm, *sI = sm + n * n, *sT = sm + 2 * n * n;
    const int e = threadIdx.x, i = e / n, j = e % n;
    sD[e] = D[e];
    sI[e] = in[e];
    __syncthreads();
    double acc = 0.0;
    if (!inverse) {
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(sI[i * n + k], sD[j * n + k]));
    } else {
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(sD[k * n + i], sI[k * n + j]));
    }
    sT[e] = acc;
    __syncthreads();
    acc = 0.0;
    if (!inverse) {
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(sD[i * n + k], sT[k * n + j]));
    } else {
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(sT[i * n + k], sD[k * n + j]));
    }
    out[e] = acc;
}

__global__ void k_block_quantize_f64(int n, const double *Q, int adaptive, double variance, const double *c,
                                     int *q)
{
    const int e = threadIdx.x;
    double m = Q[e];
    if (adaptive && e != 0) {
        m = __dmul_rn(m, __dsub_rn(2.0, norm_variance(variance)));
        if (m < 1.0) m = 1.0;
    }
    q[e] = (int)round_half_away(__ddiv_rn(c[e], m));
}

__global__ void k_block_dequantize_f64(int n, const double *R, int adaptive, double variance, const int *q,
                                       double *c)
{
    const int e = threadIdx.x;
    double m = R[e];
    if (adaptive) {
        if (e != 0) m = __dmul_rn(m, __ddiv_rn(1.0, __dsub_rn(2.0, norm_variance(variance))));
        c[e] = __dmul_rn((double)q[e], __ddiv_rn(1.0, m));
    } else {
        c[e] = __dmul_rn((double)q[e], m);
    }
}

}  // namespace

static unsigned replay_grid(const ReplayParams &p)
{
    // worklist mode: the count lives on the device; a fixed grid (4 CTAs per SM) strides over it
    if (p.worklist != nullptr) {                         // small planes: no point in 592 CTAs that only set up and leave
        const unsigned ctas = (p.nblocks + 2047) / 2048;
        return ctas < 16u ? 16u : (ctas > 148u * 4u ? 148u * 4u : ctas);
    }
    const unsigned ctas = (p.nblocks + kReplayThreads / 8 - 1) / (kReplayThreads / 8);
    return ctas < 148u * 8u ? (ctas ? ctas : 1u) : 148u * 8u;
}

template <typename K> static cudaError_t launch_replay(K kernel, const ReplayParams &p, cudaStream_t s)
{
    kernel<<<replay_grid(p), kReplayThreads, sizeof(ReplayShared), s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_replay_fwd(const ReplayParams &p, cudaStream_t s)
{
    if (!p.px_is_f32 && p.worklist != nullptr && p.seg_count != nullptr) {      // uint8 planes on the fast path: one lane per flagged block
        // one CTA per ~1024 blocks of the plane (a tenth of them flagged would keep its four warps busy), at most 4 per SM
        unsigned grid = (p.nblocks + 1023) / 1024;
        grid = grid < 16u ? 16u : (grid > 148u * 4u ? 148u * 4u : grid);
        return p.layout == LAYOUT_ZIGZAG ? launch_pdl(k_replay_fwd_lane<LAYOUT_ZIGZAG>, grid, kLaneThreads, 0, s, p)
                                         : launch_pdl(k_replay_fwd_lane<LAYOUT_NATURAL>, grid, kLaneThreads, 0, s, p);
    }
    if (p.px_is_f32)
        return p.layout == LAYOUT_ZIGZAG ? launch_replay(k_replay_fwd<LAYOUT_ZIGZAG, true>, p, s)
                                         : launch_replay(k_replay_fwd<LAYOUT_NATURAL, true>, p, s);
    return p.layout == LAYOUT_ZIGZAG ? launch_replay(k_replay_fwd<LAYOUT_ZIGZAG, false>, p, s)
                                     : launch_replay(k_replay_fwd<LAYOUT_NATURAL, false>, p, s);
}

cudaError_t launch_replay_inv(const ReplayParams &p, cudaStream_t s)
{
    if (p.worklist != nullptr && p.h_tab != nullptr) {   // the fused kernel ran: one lane per flagged block
        // one CTA per ~1024 blocks of the plane (a tenth of them flagged would keep its four warps busy), at most 4 per SM
        unsigned grid = (p.nblocks + 1023) / 1024;
        grid = grid < 16u ? 16u : (grid > 148u * 5u ? 148u * 5u : grid);
        InvLaneParams q;
        q.p = p;
        memcpy(q.rs32, p.h_tab->rs32, sizeof q.rs32);
        memcpy(q.rg32, p.h_tab->rg32, sizeof q.rg32);
        q.band_floor = p.h_tab->band_floor;
        if (p.adaptive)
            return p.layout == LAYOUT_ZIGZAG ? launch_pdl(k_replay_inv_lane<LAYOUT_ZIGZAG, true>, grid, kLaneThreads, 0, s, q)
                                             : launch_pdl(k_replay_inv_lane<LAYOUT_NATURAL, true>, grid, kLaneThreads, 0, s, q);
        return p.layout == LAYOUT_ZIGZAG ? launch_pdl(k_replay_inv_lane<LAYOUT_ZIGZAG, false>, grid, kLaneThreads, 0, s, q)
                                         : launch_pdl(k_replay_inv_lane<LAYOUT_NATURAL, false>, grid, kLaneThreads, 0, s, q);
    }
    return p.layout == LAYOUT_ZIGZAG ? launch_replay(k_replay_inv<LAYOUT_ZIGZAG>, p, s)
                                     : launch_replay(k_replay_inv<LAYOUT_NATURAL>, p, s);
}

cudaError_t launch_block_dct_f64(int n, const double *d_D, const double *d_in, double *d_out, int inverse,
                                 cudaStream_t s)
{
    k_block_dct_f64<<<1, n * n, 3 * n * n * sizeof(double), s>>>(n, d_D, d_in, d_out, inverse);
    return cudaGetLastError();
}

cudaError_t launch_block_quantize_f64(int n, const double *d_Q, int adaptive, double variance,
                                      const double *d_c, int *d_q, cudaStream_t s)
{
    k_block_quantize_f64<<<1, n * n, 0, s>>>(n, d_Q, adaptive, variance, d_c, d_q);
    return cudaGetLastError();
}

cudaError_t launch_block_dequantize_f64(int n, const double *d_R, int adaptive, double variance,
                                        const int *d_q, double *d_c, cudaStream_t s)
{
    k_block_dequantize_f64<<<1, n * n, 0, s>>>(n, d_R, adaptive, variance, d_q, d_c);
    return cudaGetLastError();
}

}  // namespace dctb
