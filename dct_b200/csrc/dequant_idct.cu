// dequant_idct.cu -- K2: int16 records -> dequantise -> 8x8 inverse DCT -> +128, round, clamp -> u8.
//
// Fuses, for every 8x8 block, the reference's
//   zigzag_to_block (src/entropy.c:183-210, ZIGZAG layout only),
//   dequantize (src/quantization.c:133-151: non-adaptive q * (1/Q) -- sic, SURVEY.md S2;
//   adaptive q * (1.0 / ((1/Q) * (1/(2-nv)))), DC unscaled), dct_inverse (src/dct.c:80-105)
// and the pixel rule p = clamp(round(x + 128.0), 0, 255) into one pass: 128 B in, 64 B out.
//
// Mapping mirrors K1: one thread per block, a warp owns a tile of 32 records (4 KB) and writes a 256-pixel x 8-row
// tile with 8 STG.64 per lane (256 contiguous bytes per warp instruction).  Both butterfly passes and the pixel
// residuals run on packed fp32 instructions (FADD2 / FFMA2); the table multiplication is folded into the first
// butterfly stage (inv_block below, shared by the kernels).  The fp32 error bound is dynamic here (inputs are
// arbitrary int16): 2^-24 * sum gain_k |v_k|.
//
//   k_dequant_idct_u8_tma (default)  persistent, ONE CTA of 16 warps per SM; per warp a two-stage bulk-tensor pipeline:
//                         one cp.async.bulk.tensor.2d (UTMALDG) per tile into a 128B-swizzled stage, read back with
//                         conflict-free LDS.128; per-warp worklist segments; small planes replay in the tail (FOLD).
//   k_dequant_idct_u8     (fallback: unaligned records, planes under 256 pixels wide, peer memory)  one-shot grid,
//                         LDG.128 through a padded stage, atomic worklist append.
//   k_dequant_idct_u8_f64 the butterfly in fp64 (adaptive plans with DCT_CUDA_INV_FP64=1).
#include <cstdio>
#include <cstdlib>

#include "fast_core.cuh"
#include "kernels.cuh"
#include "replay_lane.cuh"
#include "tma.cuh"

namespace dctb {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStageWordsPerBlock = 36;
constexpr int kStageWordsPerWarp = 32 * kStageWordsPerBlock;


__device__ __forceinline__ uint4 ldg_stream_u4(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ void stg_stream_u2(void *p, uint32_t a, uint32_t b)
{
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}

// The arithmetic of one block, shared by the one-shot and the bulk-tensor kernels: w[m] = packed int16 pair m of the
// record in storage order; writes the block's 64 pixels (if valid); returns "replay this block in fp64".
template <int LAYOUT, bool ADAPTIVE>
__device__ __forceinline__ bool inv_block(const InvParams &p, const uint32_t (&w)[32], bool valid, uint8_t *dst, double var)
{
    float s = 1.0f;
    // ADAPTIVE: `var` = the block's variance (side information); multiplier 1/((1/Q)*(1/(2-nv))) = Q*(2-nv) up to
    // fp64 rounding (rs holds Q * prescale); exact form in K3
    if constexpr (ADAPTIVE) s = adaptive_scale(var);

    // v holds the converted q themselves: the multiplication by the table happens inside the first butterfly stage
    // (idct8_dequant: one FMA gives v_a +- q_b * rs_b).  ADAPTIVE: the block's factor (2 - nv) goes onto the 63 AC
    // values first (32 packed multiplications per block; the DC entry is not scaled, src/quantization.c:199-201) --
    // the product q_k * s is one more rounding inside the 8u the inputs are allowed.
    float v[64];
    float bound = 0.f, bound_dc = 0.f;   // bound >= sum gain_k * |v_k|
    static_for<0, 32>([&](auto M) {
        constexpr int m = decltype(M)::value;
        constexpr int k0 = storage_to_natural<LAYOUT>(2 * m), k1 = storage_to_natural<LAYOUT>(2 * m + 1);
        const float f0 = half_to_float<0>(w[m]), f1 = half_to_float<1>(w[m]);
        v[k0] = f0, v[k1] = f1;
        if (ADAPTIVE && k0 == 0) bound_dc = __fmul_rn(fabsf(f0), p.rg[0]);
        else bound = __fmaf_rn(fabsf(f0), p.rg[k0], bound);
        if (ADAPTIVE && k1 == 0) bound_dc = __fmul_rn(fabsf(f1), p.rg[0]);
        else bound = __fmaf_rn(fabsf(f1), p.rg[k1], bound);
    });
    if constexpr (ADAPTIVE) bound = adaptive_bound(bound, bound_dc, s);
    // |fp32 pixel - exact pixel| <= 2^-24 * bound (derive_bands.py) ; + floor for the residual's own rounding
    const float thr = pixel_threshold(bound, p.band_floor);

    // Columns (D^T * in), then rows (temp * D): same order as src/dct.c:85-102.  Both passes run on packed
    // pairs (FADD2 / FFMA2: two fp32 lanes per instruction, half the issue slots).  The column pass takes
    // the column pairs (0,4) (2,6) (5,3) (1,7) in its two lanes -- exactly the pairs the first stage of the
    // row pass adds and subtracts, so that stage is 64 scalar FADDs on the two halves of a register pair
    // whose results are written straight into (row 2a, row 2a+1) pairs: no re-pairing instruction.
    constexpr int kPairA[4] = {0, 2, 5, 1}, kPairB[4] = {4, 6, 3, 7};
    float2 cpv[4][8];                        // cpv[c][i] = (T[i][A_c], T[i][B_c]) after the column pass
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 8; ++i) cpv[c][i] = make_float2(v[8 * i + kPairA[c]], v[8 * i + kPairB[c]]);
        if constexpr (ADAPTIVE) {          // the block's (2 - nv) goes onto the 63 AC values; the tables stay constant operands
            const float2 s2 = make_float2(s, s);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (c == 0 && i == 0) cpv[0][0].y = __fmul_rn(cpv[0][0].y, s);      // natural index 0: the DC entry is not scaled
                else cpv[c][i] = Ops<float2>::mul(cpv[c][i], s2);
            }
        }
        idct8_dequant<float2, 1>(cpv[c], p.ma[c], p.mb[c]);
    }

    // Row pass on row pairs (2a, 2a+1), then per pixel: t = x + (1.5*2^23 + 128): the low 16 mantissa bits
    // of t are round(x) + 128 as an int16 (valid for |x| < 2^15 - 128, guaranteed below by the bound test);
    // residual e = x - round(x) is exact.  Clamp to [0, 255] on packed int16 pairs (VIMNMX.S16x2.RELU),
    // then pack four bytes per word.
    float emax = 0.0f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i0 = 2 * a, i1 = 2 * a + 1;
        // first stage of both rows: sums and differences of the two halves of the column-pass pairs
        const float2 t10 = make_float2(__fadd_rn(cpv[0][i0].x, cpv[0][i0].y), __fadd_rn(cpv[0][i1].x, cpv[0][i1].y));
        const float2 t11 = make_float2(__fsub_rn(cpv[0][i0].x, cpv[0][i0].y), __fsub_rn(cpv[0][i1].x, cpv[0][i1].y));
        const float2 t13 = make_float2(__fadd_rn(cpv[1][i0].x, cpv[1][i0].y), __fadd_rn(cpv[1][i1].x, cpv[1][i1].y));
        const float2 d26 = make_float2(__fsub_rn(cpv[1][i0].x, cpv[1][i0].y), __fsub_rn(cpv[1][i1].x, cpv[1][i1].y));
        const float2 z13 = make_float2(__fadd_rn(cpv[2][i0].x, cpv[2][i0].y), __fadd_rn(cpv[2][i1].x, cpv[2][i1].y));
        const float2 z10 = make_float2(__fsub_rn(cpv[2][i0].x, cpv[2][i0].y), __fsub_rn(cpv[2][i1].x, cpv[2][i1].y));
        const float2 z11 = make_float2(__fadd_rn(cpv[3][i0].x, cpv[3][i0].y), __fadd_rn(cpv[3][i1].x, cpv[3][i1].y));
        const float2 z12 = make_float2(__fsub_rn(cpv[3][i0].x, cpv[3][i0].y), __fsub_rn(cpv[3][i1].x, cpv[3][i1].y));
        float2 x2[8];                        // x2[j] = samples (2a, j) and (2a+1, j)
        idct8_tail<float2, 1>(x2, t10, t11, t13, d26, z13, z10, z11, z12);
        float2 t2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float2 e2;
            pixel_residual2(x2[j], t2[j], e2);
            emax = fmaxf(fmaxf(emax, fabsf(e2.x)), fabsf(e2.y));                               // FMNMX3
        }
        uint32_t pr0[4], pr1[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint32_t p0 = __byte_perm(__float_as_uint(t2[2 * h].x), __float_as_uint(t2[2 * h + 1].x), 0x5410);
            const uint32_t p1 = __byte_perm(__float_as_uint(t2[2 * h].y), __float_as_uint(t2[2 * h + 1].y), 0x5410);
            asm("min.s16x2.relu %0, %1, %2;" : "=r"(pr0[h]) : "r"(p0), "r"(0x00ff00ffu));
            asm("min.s16x2.relu %0, %1, %2;" : "=r"(pr1[h]) : "r"(p1), "r"(0x00ff00ffu));
        }
        if (valid) {
            stg_stream_u2(dst + i0 * p.pitch, __byte_perm(pr0[0], pr0[1], 0x6420), __byte_perm(pr0[2], pr0[3], 0x6420));
            stg_stream_u2(dst + i1 * p.pitch, __byte_perm(pr1[0], pr1[1], 0x6420), __byte_perm(pr1[2], pr1[3], 0x6420));
        }
    }
    // |x| <= bound / 5 (every basis product is <= 1/4 and gain_k * prescale_k >= 1.25), so bound < 1.4e5
    // keeps |x| far below 2^15; larger inputs go to the fp64 path
    const bool flag = (emax >= thr) | !(bound < 1.4e5f);

    return flag;
}

template <int LAYOUT, bool ADAPTIVE>
__global__ void __launch_bounds__(kThreads, 3) k_dequant_idct_u8(const __grid_constant__ InvParams p)
{
    __shared__ uint4 stage[kWarps * kStageWordsPerWarp / 4];

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warp_base = blockIdx.x * kThreads + warp * 32;
    const uint32_t b = warp_base + lane;
    const bool valid = b < p.nblocks;

    // 4 KB of records: coalesced 16-byte chunks -> padded stage -> one record per lane
    uint32_t *wstage = reinterpret_cast<uint32_t *>(stage) + warp * kStageWordsPerWarp;
    const uint4 *srcv = reinterpret_cast<const uint4 *>(p.coef) + (size_t)warp_base * 8;
    uint4 chunk[8];
    const uint32_t full = warp_base < p.nblocks ? p.nblocks - warp_base : 0;   // records in this tile
    if (full >= 32) {                                   // warp-uniform: every tile but possibly the last
#pragma unroll
        for (int j = 0; j < 8; ++j) chunk[j] = ldg_stream_u4(srcv + j * 32 + lane);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            chunk[j] = (j * 4 + (lane >> 3) < full) ? ldg_stream_u4(srcv + j * 32 + lane) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t c = j * 32 + lane;
        *reinterpret_cast<uint4 *>(wstage + (c >> 3) * kStageWordsPerBlock + 4 * (c & 7)) = chunk[j];
    }
    __syncwarp();
    uint32_t w[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint4 t = *reinterpret_cast<const uint4 *>(wstage + lane * kStageWordsPerBlock + 4 * j);
        w[4 * j] = t.x, w[4 * j + 1] = t.y, w[4 * j + 2] = t.z, w[4 * j + 3] = t.w;
    }

    const uint32_t bb = valid ? b : p.nblocks - 1;
    const uint32_t by = bb / p.bw, bx = bb - by * p.bw;
    const double var = (ADAPTIVE && p.var_in != nullptr && valid) ? p.var_in[b] : 0.0;
    const bool flag = inv_block<LAYOUT, ADAPTIVE>(p, w, valid, p.px + (long long)by * 8 * p.pitch + (long long)bx * 8, var);

    const unsigned ballot = __ballot_sync(0xffffffffu, flag && valid);
    if (ballot != 0) {
        const int leader = __ffs(ballot) - 1;
        unsigned base = 0;
        if ((int)lane == leader) base = atomicAdd(&p.ctr->wl_count, (unsigned)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (flag && valid) {
            const unsigned pos = base + __popc(ballot & ((1u << lane) - 1u));
            if (pos < p.wl_cap) p.worklist[pos] = b;
        }
    }
}



// ------------------------------------------------------------------------------------------
// K2 with bulk-tensor (TMA) record tiles -- the default whenever the plane allows it (16-byte aligned records,
// block rows of at least 32 blocks).  Same arithmetic as k_dequant_idct_u8 above (inv_block); what changes:
//   * persistent grid, every warp runs its own two-stage pipeline: lane 0 fetches the 4 KB of the next tile's 32 records
//     with ONE cp.async.bulk.tensor.2d (128-byte swizzle: a lane then reads its own record with 8 conflict-free LDS.128)
//     while the warp transforms the current tile; no LDG / STS re-staging pass, no address arithmetic per chunk;
//   * a tile is 32 blocks of ONE block row, so the pixel address of a lane is a warp-uniform base + 8 * lane;
//   * flagged blocks go to the warp's own worklist segment (as in K1): no global atomic, which matters when a
//     high-quality table makes the dynamic band flag every other tile.
// ------------------------------------------------------------------------------------------
struct alignas(64) InvTmaParams {
    CUtensorMap map_rec;       // records, box 128 B x 32, 128-byte swizzle
    InvParams f;
    uint32_t tpr;              // tiles per block row = ceil(bw / 32)
    uint32_t nby;              // block rows
    uint32_t step_ty, step_tx; // divmod(warps in the grid, tpr)
    uint32_t n_segs, rot;      // warps in the grid; the warp that takes the plane's tile 0
    uint32_t *seg_count;       // one entry per warp of the grid
    uint32_t seg_cap;          // worklist entries per segment
};

constexpr int kTmaInBytes = 4096;                                   // per warp and stage: 32 records
constexpr int kTmaCtlBytes = 64;                                    // per warp: up to 8 mbarriers
constexpr int kTmaVarBytes = 256;                                   // per warp and stage: 32 variances (adaptive plans)
constexpr int kTmaMaxPlanes = 3;                                    // planes one launch may cover
constexpr int kTmaCntBytes = 256;                                   // per CTA: flagged blocks per (plane, warp), up to 3 x 16

// geometry of one kernel variant: warps per CTA, CTAs per SM the register budget is cut for, record stages per warp
template <int WARPS, int MIN_CTAS, int STAGES> struct TmaCfg {
    static constexpr int kWarpsT = WARPS, kMinCtas = MIN_CTAS, kStages = STAGES, kThreadsT = 32 * WARPS;
    static constexpr int kSmem = 1024 /* alignment slack */ + WARPS * (STAGES * (kTmaInBytes + kTmaVarBytes) + kTmaCtlBytes) + kTmaCntBytes;
    static_assert(STAGES >= 2 && STAGES <= 8, "stages");
    static_assert(sizeof(InvLaneScratch) <= STAGES * kTmaInBytes, "the replay's scratch is the warp's record stages");
    static_assert(kTmaMaxPlanes * WARPS * 4 <= kTmaCntBytes, "per-CTA counts");
};

// the warp's shared memory: [record stages, 1024-aligned for the swizzle][mbarriers][variance stages], then the CTA's
// counts of flagged blocks; the replay's scratch is the warp's (by then idle) record stages
template <typename CFG> struct InvWarpSmem {
    uint8_t *in_p;
    InvLaneScratch *ws;
    double *var_p;
    uint32_t *cnt_all;     // CTA-wide: blocks flagged per (plane, warp), for the pooled replay
    uint32_t bar_s;
    __device__ __forceinline__ InvWarpSmem(uint8_t *smem_raw, uint32_t warp, uint32_t lane)
    {
        constexpr int kW = CFG::kWarpsT, kS = CFG::kStages;
        uint8_t *sm = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
        in_p = sm + warp * (kS * kTmaInBytes);
        uint32_t *ctl_p = reinterpret_cast<uint32_t *>(sm + kW * kS * kTmaInBytes + warp * kTmaCtlBytes);
        ws = reinterpret_cast<InvLaneScratch *>(in_p);
        var_p = reinterpret_cast<double *>(sm + kW * (kS * kTmaInBytes + kTmaCtlBytes) + warp * (kS * kTmaVarBytes));
        cnt_all = reinterpret_cast<uint32_t *>(sm + kW * (kS * (kTmaInBytes + kTmaVarBytes) + kTmaCtlBytes));
        bar_s = (uint32_t)__cvta_generic_to_shared(ctl_p);           // kS 8-byte mbarriers
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < kS; ++i) tma::mbar_init(bar_s + 8 * i, 1);
            tma::fence_barrier_init();
        }
        __syncwarp();
    }
};

// All the tiles of ONE plane that fall to this warp; returns how many blocks it flagged.  `phases` holds the parity bit
// of each stage's mbarrier and is carried from plane to plane by the multi-plane kernel (as in K1).
template <int LAYOUT, bool ADAPTIVE, typename CFG, bool MULTI>
__device__ __forceinline__ uint32_t inv_plane_tiles(const InvTmaParams &P, const uint32_t lane, const uint32_t gwarp, uint8_t *in_p,
                                                    double *var_p, const uint32_t bar_s, uint32_t &phases)
{
    constexpr int kS = CFG::kStages;
    const InvParams &p = P.f;
    const uint32_t var_s = (uint32_t)__cvta_generic_to_shared(var_p);
    // adaptive plans: the tile's 32 variances (side information) ride the same mbarrier as its records when the
    // array allows a bulk copy (16-byte aligned, an even number of blocks per row); a plain load otherwise
    const bool var_bulk = ADAPTIVE && p.var_in != nullptr && (p.bw % 2 == 0) && ((uintptr_t)p.var_in % 16 == 0);
    const uint32_t in_s = (uint32_t)__cvta_generic_to_shared(in_p);

    // (ty, tx): block row and tile-in-row of the tile being transformed; (fy, fx): of the next tile to fetch, kS - 1 ahead;
    // the plane's tile 0 belongs to warp `rot` of the grid (see K1)
    uint32_t ty, tx;
    {
        const uint32_t t = !MULTI ? gwarp : (gwarp >= P.rot ? gwarp - P.rot : gwarp + P.n_segs - P.rot);
        ty = t / P.tpr;
        tx = t - ty * P.tpr;
    }
    uint32_t fy = ty, fx = tx, fstage = 0;
    auto fetch = [&]() {        // fetch tile (fy, fx) into stage fstage, then advance both
        if (fy < P.nby && lane == 0) {
            const uint32_t vbytes = var_bulk ? min(32u, p.bw - fx * 32) * 8u : 0u;
            tma::mbar_expect_tx(bar_s + fstage * 8, kTmaInBytes + vbytes);
            tma::load_2d(in_s + fstage * kTmaInBytes, &P.map_rec, 0, (int)(fy * p.bw + fx * 32), bar_s + fstage * 8);
            if (var_bulk) tma::load_1d(var_s + fstage * kTmaVarBytes, p.var_in + (size_t)fy * p.bw + fx * 32, vbytes, bar_s + fstage * 8);
        }
        fx += P.step_tx;
        fy += P.step_ty;
        if (fx >= P.tpr) fx -= P.tpr, ++fy;
        fstage = fstage + 1 == kS ? 0 : fstage + 1;
    };
#pragma unroll
    for (int i = 0; i < kS - 1; ++i) fetch();
    const uint32_t swz = (lane & 7) << 4;
    uint32_t wl_n = 0;          // entries this warp has appended to its worklist segment

    uint32_t stage = 0;
    while (ty < P.nby) {
        const uint32_t bx0 = tx * 32;
        const uint32_t warp_base = ty * p.bw + bx0;
        const uint32_t nvalid = min(32u, p.bw - bx0);
        uint8_t *dst = p.px + (long long)ty * 8 * p.pitch + (long long)(bx0 + lane) * 8;
        fetch();                                                  // into the stage the previous iteration consumed
        // one plane per launch: the stages' parities move in step (one bit); planes sharing a launch leave the stages
        // with different parities (a warp may get an odd number of tiles of a plane): one bit per stage
        if constexpr (MULTI) {
            tma::mbar_wait(bar_s + stage * 8, (phases >> stage) & 1u);
            phases ^= 1u << stage;
        } else {
            tma::mbar_wait(bar_s + stage * 8, phases);
        }
        // adaptive plans: the lane's variance (8 bytes of side information per block), from the stage or from memory
        double var = 0.0;
        if (ADAPTIVE && p.var_in != nullptr && lane < nvalid)
            var = var_bulk ? var_p[stage * (kTmaVarBytes / 8) + lane] : p.var_in[warp_base + lane];

        const uint32_t b = warp_base + lane;
        const bool valid = lane < nvalid;
        const uint8_t *rec = in_p + stage * kTmaInBytes + lane * 128;
        uint32_t w[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 t = *reinterpret_cast<const uint4 *>(rec + ((j << 4) ^ swz));
            w[4 * j] = t.x, w[4 * j + 1] = t.y, w[4 * j + 2] = t.z, w[4 * j + 3] = t.w;
        }
        __syncwarp();           // every lane has its record: the next fetch overwrites this stage

        const bool flag = inv_block<LAYOUT, ADAPTIVE>(p, w, valid, dst, var);
        const unsigned ballot = __ballot_sync(0xffffffffu, flag && valid);
        if (ballot != 0) {
            if (flag && valid) p.worklist[(size_t)gwarp * P.seg_cap + wl_n + __popc(ballot & ((1u << lane) - 1u))] = b;
            wl_n += __popc(ballot);
        }
        tx += P.step_tx;
        ty += P.step_ty;
        if (tx >= P.tpr) tx -= P.tpr, ++ty;
        if (++stage == kS) {
            stage = 0;
            if constexpr (!MULTI) phases ^= 1u;
        }
    }
    if (lane == 0) P.seg_count[gwarp] = 0;                    // nothing for K3: replayed below (inv_replay_pooled)
    __syncwarp();
    return wl_n;
}

// Small planes: the flagged blocks are replayed in the kernel's tail and no K3 is launched.  As in K1
// (fwd_replay_pooled, fwd_quant.cu) the CTA pools the entries of all its warps and of all the planes of the launch and
// deals them out in batches of 32, one block per lane.  The barrier also orders the pixels K2 stored before the bytes the
// replay patches into them.  Large planes leave the segments to K3's grid of replay-only warps.
template <int LAYOUT, bool ADAPTIVE, typename CFG, int NPL>
__device__ __forceinline__ void inv_replay_pooled(const InvTmaParams *pl, const uint32_t (&n_mine)[NPL], uint32_t *cnt_all, const uint32_t lane,
                                                  const uint32_t warp, InvLaneScratch *ws)
{
    constexpr int kW = CFG::kWarpsT;
    if (lane == 0) {
        static_for<0, NPL>([&](auto I) {
            constexpr int i = decltype(I)::value;
            cnt_all[i * kW + warp] = n_mine[i];
            if (n_mine[i] != 0) atomicAdd(&pl[i].f.ctr->replayed, (unsigned long long)n_mine[i]);
        });
    }
    __syncthreads();
    uint32_t total = 0;
    for (int it = 0; it < NPL * kW; ++it) total += cnt_all[it];
    for (uint32_t first = warp * 32; first < total; first += kW * 32) {
        int item = 0;
        uint32_t e = 0;
        const bool active = locate_entry(cnt_all, NPL * kW, first + lane, item, e);
        const int plane = item / kW;
        const uint32_t gw = blockIdx.x * kW + (uint32_t)(item - plane * kW);     // the warp that flagged the block
        InvReplayCtx cx{};
        uint32_t b = 0;
        static_for<0, NPL>([&](auto I) {
            constexpr int i = decltype(I)::value;
            if (i == 0 || plane == i) {       // idle lanes run the arithmetic on plane 0's tables
                const InvParams &p = pl[i].f;
                cx = InvReplayCtx{p.rs, p.rg, p.band_floor, p.tab->D, p.tab->R, p.tab->mult64, p.coef, p.var_in, p.px, p.pitch, p.bw, p.ctr};
                if (active && plane == i) b = p.worklist[(size_t)gw * pl[i].seg_cap + e];
            }
        });
        replay_inv_lanes<LAYOUT, ADAPTIVE>(cx, ws, active, b);
    }
}

// The streaming kernel: one large plane, flagged blocks left to K3.  (Kept apart from the frame kernel below, which shares
// its tile loop in spirit but not in text: at 128 registers per thread the schedule of this loop is worth 6 % of the
// kernel, and any code around it moves it.)
template <int LAYOUT, bool ADAPTIVE, typename CFG>
__global__ void __launch_bounds__(CFG::kThreadsT, CFG::kMinCtas) k_dequant_idct_u8_tma(const __grid_constant__ InvTmaParams P)
{
    constexpr int kW = CFG::kWarpsT, kS = CFG::kStages;
    extern __shared__ uint8_t smem_raw[];
    const InvParams &p = P.f;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // tells the compiler it is warp-uniform
    uint8_t *sm = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    uint8_t *in_p = sm + warp * (kS * kTmaInBytes);
    uint32_t *ctl_p = reinterpret_cast<uint32_t *>(sm + kW * kS * kTmaInBytes + warp * kTmaCtlBytes);
    double *var_p = reinterpret_cast<double *>(sm + kW * (kS * kTmaInBytes + kTmaCtlBytes) + warp * (kS * kTmaVarBytes));
    const uint32_t var_s = (uint32_t)__cvta_generic_to_shared(var_p);
    // adaptive plans: the tile's 32 variances (side information) ride the same mbarrier as its records when the
    // array allows a bulk copy (16-byte aligned, an even number of blocks per row); a plain load otherwise
    const bool var_bulk = ADAPTIVE && p.var_in != nullptr && (p.bw % 2 == 0) && ((uintptr_t)p.var_in % 16 == 0);
    const uint32_t in_s = (uint32_t)__cvta_generic_to_shared(in_p);
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(ctl_p);           // kS 8-byte mbarriers

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kS; ++i) tma::mbar_init(bar_s + 8 * i, 1);
        tma::fence_barrier_init();
    }
    __syncwarp();
    pdl_launch_dependents();    // the replay kernel may start its prologue while this grid runs
    pdl_wait();                 // everything below touches memory the previous kernels on the stream wrote or read

    // (ty, tx): block row and tile-in-row of the tile being transformed; (fy, fx): of the next tile to fetch, kS - 1 ahead
    uint32_t ty, tx;
    {
        const uint32_t t = blockIdx.x * kW + warp;
        ty = t / P.tpr;
        tx = t - ty * P.tpr;
    }
    uint32_t fy = ty, fx = tx, fstage = 0;
    auto fetch = [&]() {        // fetch tile (fy, fx) into stage fstage, then advance both
        if (fy < P.nby && lane == 0) {
            const uint32_t vbytes = var_bulk ? min(32u, p.bw - fx * 32) * 8u : 0u;
            tma::mbar_expect_tx(bar_s + fstage * 8, kTmaInBytes + vbytes);
            tma::load_2d(in_s + fstage * kTmaInBytes, &P.map_rec, 0, (int)(fy * p.bw + fx * 32), bar_s + fstage * 8);
            if (var_bulk) tma::load_1d(var_s + fstage * kTmaVarBytes, p.var_in + (size_t)fy * p.bw + fx * 32, vbytes, bar_s + fstage * 8);
        }
        fx += P.step_tx;
        fy += P.step_ty;
        if (fx >= P.tpr) fx -= P.tpr, ++fy;
        fstage = fstage + 1 == kS ? 0 : fstage + 1;
    };
#pragma unroll
    for (int i = 0; i < kS - 1; ++i) fetch();
    const uint32_t swz = (lane & 7) << 4;
    uint32_t wl_n = 0;          // entries this warp has appended to its worklist segment
    const uint32_t gwarp = blockIdx.x * kW + warp;

    uint32_t stage = 0, phase = 0;
    while (ty < P.nby) {
        const uint32_t bx0 = tx * 32;
        const uint32_t warp_base = ty * p.bw + bx0;
        const uint32_t nvalid = min(32u, p.bw - bx0);
        uint8_t *dst = p.px + (long long)ty * 8 * p.pitch + (long long)(bx0 + lane) * 8;
        fetch();                                                  // into the stage the previous iteration consumed
        tma::mbar_wait(bar_s + stage * 8, phase);
        // adaptive plans: the lane's variance (8 bytes of side information per block), from the stage or from memory
        double var = 0.0;
        if (ADAPTIVE && p.var_in != nullptr && lane < nvalid)
            var = var_bulk ? var_p[stage * (kTmaVarBytes / 8) + lane] : p.var_in[warp_base + lane];

        const uint32_t b = warp_base + lane;
        const bool valid = lane < nvalid;
        const uint8_t *rec = in_p + stage * kTmaInBytes + lane * 128;
        uint32_t w[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 t = *reinterpret_cast<const uint4 *>(rec + ((j << 4) ^ swz));
            w[4 * j] = t.x, w[4 * j + 1] = t.y, w[4 * j + 2] = t.z, w[4 * j + 3] = t.w;
        }
        __syncwarp();           // every lane has its record: the next fetch overwrites this stage

        const bool flag = inv_block<LAYOUT, ADAPTIVE>(p, w, valid, dst, var);
        const unsigned ballot = __ballot_sync(0xffffffffu, flag && valid);
        if (ballot != 0) {
            if (flag && valid) p.worklist[(size_t)gwarp * P.seg_cap + wl_n + __popc(ballot & ((1u << lane) - 1u))] = b;
            wl_n += __popc(ballot);
        }
        tx += P.step_tx;
        ty += P.step_ty;
        if (tx >= P.tpr) tx -= P.tpr, ++ty;
        if (++stage == kS) stage = 0, phase ^= 1;
    }
    // the flagged blocks are left to K3's grid of replay-only warps (small planes: k_dequant_idct_u8_tma_frame below)
    if (lane == 0) P.seg_count[gwarp] = wl_n;                 // <= 32 per tile visited, < seg_cap by construction
}

// The frame kernel: one small plane, or the 2 or 3 planes of one frame (Y, Cb, Cr) in ONE launch as in K1 (fwd_quant.cu);
// the flagged blocks of all the planes are replayed in one pooled tail and no K3 is launched.
template <int NPL> struct alignas(64) InvTmaMulti {
    InvTmaParams pl[NPL];
};

template <int LAYOUT, bool ADAPTIVE, typename CFG, int NPL>
__global__ void __launch_bounds__(CFG::kThreadsT, CFG::kMinCtas) k_dequant_idct_u8_tma_frame(const __grid_constant__ InvTmaMulti<NPL> M)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const InvWarpSmem<CFG> sm(smem_raw, warp, lane);
    pdl_launch_dependents();
    pdl_wait();
    uint32_t phases = 0;
    uint32_t n_mine[NPL];
    static_for<0, NPL>([&](auto I) {
        constexpr int i = decltype(I)::value;
        n_mine[i] = inv_plane_tiles<LAYOUT, ADAPTIVE, CFG, (NPL > 1)>(M.pl[i], lane, blockIdx.x * CFG::kWarpsT + warp, sm.in_p, sm.var_p, sm.bar_s, phases);
    });
    inv_replay_pooled<LAYOUT, ADAPTIVE, CFG, NPL>(M.pl, n_mine, sm.cnt_all, lane, warp, sm.ws);
}

// ------------------------------------------------------------------------------------------
// fp64 variant: same mapping, the butterfly in double precision.  Used for ADAPTIVE plans, whose
// dequantised values are full-scale (q * Q * (2-nv)): there the fp32 dynamic bound flags most
// blocks of busy content, while the fp64 band (~1e-9) flags only true near-ties.  fp64 runs at
// half the fp32 rate on B200, so this costs ~1.5x the fp32 kernel but is content-independent.
// ------------------------------------------------------------------------------------------
constexpr int kThreads64 = 128;
constexpr int kWarps64 = kThreads64 / 32;
constexpr double kMagic52_128 = 6755399441055744.0 + 128.0;   // 1.5 * 2^52 + 128

template <int LAYOUT>
__global__ void __launch_bounds__(kThreads64) k_dequant_idct_u8_f64(const __grid_constant__ InvParams p, const ExactTables *__restrict__ tab)
{
    __shared__ uint4 stage[kWarps64 * kStageWordsPerWarp / 4];
    __shared__ double s_mp[64];
    __shared__ float s_gain[64];
    if (threadIdx.x < 64) {
        s_mp[threadIdx.x] = tab->mp64[threadIdx.x];
        s_gain[threadIdx.x] = tab->gain32[threadIdx.x];
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warp_base = blockIdx.x * kThreads64 + warp * 32;
    const uint32_t b = warp_base + lane;
    const bool valid = b < p.nblocks;

    uint32_t *wstage = reinterpret_cast<uint32_t *>(stage) + warp * kStageWordsPerWarp;
    const uint4 *srcv = reinterpret_cast<const uint4 *>(p.coef) + (size_t)warp_base * 8;
    const uint32_t full = warp_base < p.nblocks ? p.nblocks - warp_base : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t c = j * 32 + lane;
        const uint4 chunk = (j * 4 + (lane >> 3) < full) ? ldg_stream_u4(srcv + c) : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(wstage + (c >> 3) * kStageWordsPerBlock + 4 * (c & 7)) = chunk;
    }
    __syncwarp();

    // (2 - nv) in fp64 with the reference's own expression (src/quantization.c:186-190)
    const double var = (p.var_in != nullptr && valid) ? p.var_in[b] : 0.0;
    const double two_minus_nv = __dsub_rn(2.0, fmin(1.0, fmax(0.1, __ddiv_rn(var, 1000.0))));

    double v[64];
    double bound = 0.0;
    static_for<0, 8>([&](auto J) {
        constexpr int j = decltype(J)::value;
        const uint4 t = *reinterpret_cast<const uint4 *>(wstage + lane * kStageWordsPerBlock + 4 * j);
        const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
        static_for<0, 8>([&](auto Hh) {
            constexpr int h = decltype(Hh)::value;
            constexpr int k = storage_to_natural<LAYOUT>(8 * j + h);
            const int q = (int)(int16_t)(h & 1 ? (w4[h >> 1] >> 16) : (w4[h >> 1] & 0xFFFFu));
            // q * (1/R) * (2-nv) * prescale: a few ulps from the reference's reciprocal chain, inside the
            // 8 * 2^-53 relative input error the bound allows for
            double x = (double)q * s_mp[k];
            if (k != 0) x *= two_minus_nv;
            v[k] = x;
            bound = fma(fabs(x), (double)s_gain[k], bound);
        });
    });
#pragma unroll
    for (int j = 0; j < 8; ++j) idct8<double, 8>(&v[j]);
#pragma unroll
    for (int i = 0; i < 8; ++i) idct8<double, 1>(&v[8 * i]);

    const uint32_t bb = valid ? b : p.nblocks - 1;
    const uint32_t by = bb / p.bw, bx = bb - by * p.bw;
    uint8_t *dst = p.px + (long long)by * 8 * p.pitch + (long long)bx * 8;

    // 1e-9 tie-accounting margin + 8 * 2^-53 * bound (fp64 butterfly + the reference's own rounding)
    const double thr = 0.5 - (2e-9 + bound * 8.9e-16);
    bool flag = !(bound < 1e12);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double x = v[8 * i + j];
            const double t = x + kMagic52_128;            // low mantissa bits: round(x) + 128
            const double e = x - (t - kMagic52_128);
            flag |= fabs(e) >= thr;
            int px = __double2loint(t);
            px = px < 0 ? 0 : (px > 255 ? 255 : px);
            if (j < 4) lo |= (uint32_t)px << (8 * j);
            else hi |= (uint32_t)px << (8 * (j - 4));
        }
        if (valid) stg_stream_u2(dst + i * p.pitch, lo, hi);
    }

    const unsigned ballot = __ballot_sync(0xffffffffu, flag && valid);
    if (ballot != 0) {
        const int leader = __ffs(ballot) - 1;
        unsigned base = 0;
        if ((int)lane == leader) base = atomicAdd(&p.ctr->wl_count, (unsigned)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (flag && valid) {
            const unsigned pos = base + __popc(ballot & ((1u << lane) - 1u));
            if (pos < p.wl_cap) p.worklist[pos] = b;
        }
    }
}

}  // namespace

cudaError_t launch_dequant_idct_u8_f64(const InvParams &p, const ExactTables *d_tab, int layout, cudaStream_t s)
{
    if (p.nblocks == 0) return cudaSuccess;
    const unsigned grid = (p.nblocks + kThreads64 - 1) / kThreads64;
    if (layout == LAYOUT_ZIGZAG) k_dequant_idct_u8_f64<LAYOUT_ZIGZAG><<<grid, kThreads64, 0, s>>>(p, d_tab);
    else k_dequant_idct_u8_f64<LAYOUT_NATURAL><<<grid, kThreads64, 0, s>>>(p, d_tab);
    return cudaGetLastError();
}

static int sm_count_k2()
{
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    return n;
}

template <typename CFG, typename K> static unsigned tma_ctas_per_sm(K kernel)
{
    static int per_sm = 0;   // one instance per CFG; the same for every variant of it: identical launch bounds and shared memory
    if (per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, CFG::kThreadsT, CFG::kSmem) != cudaSuccess || n < 1) n = 1;
        per_sm = n < CFG::kMinCtas ? n : CFG::kMinCtas;
        if (getenv("DCT_CUDA_DEBUG"))
            fprintf(stderr, "libdct_cuda: K2 (bulk tensor, %d warps, %d stages): %d CTAs/SM (occupancy %d), %d B smem\n", CFG::kWarpsT,
                    CFG::kStages, per_sm, n, CFG::kSmem);
    }
    return (unsigned)per_sm;
}

static unsigned tma_tiles(const InvParams &p) { return (p.nblocks / p.bw) * ((p.bw + 31) / 32); }

// one plane's share of a launch of `n_segs` warps whose warp `rot` takes the plane's first tile
static cudaError_t fill_tma_plane(InvTmaParams &q, const InvParams &p, unsigned n_segs, unsigned rot, unsigned seg_off = 0)
{
    q.f = p;
    q.nby = p.nblocks / p.bw;
    q.tpr = (p.bw + 31) / 32;
    q.n_segs = n_segs;
    q.rot = rot;
    q.step_ty = n_segs / q.tpr;
    q.step_tx = n_segs - q.step_ty * q.tpr;
    const unsigned tiles_per_warp = (q.nby * q.tpr + n_segs - 1) / n_segs;
    q.seg_cap = p.wl_cap / n_segs;
    q.seg_count = p.seg_count;
    q.f.worklist = p.worklist + seg_off;       // planes of one plan share its worklist: each has its own range of every segment
    if (n_segs > kMaxWorklistSegments - 128 || q.seg_cap < seg_off + tiles_per_warp * 32 || p.seg_count == nullptr) return cudaErrorInvalidValue;
    return make_record_map(&q.map_rec, p.coef, p.nblocks, 32);
}

template <typename CFG, typename K>
static cudaError_t launch_persistent_tma(K kernel, const InvParams &p, cudaStream_t s, WorklistSegments *segments)
{
    constexpr int kW = CFG::kWarpsT;
    cudaError_t e = ensure_smem_attributes(reinterpret_cast<const void *>(kernel), CFG::kSmem);
    if (e != cudaSuccess) return e;
    const unsigned resident = (unsigned)sm_count_k2() * tma_ctas_per_sm<CFG>(kernel);
    const unsigned want = (tma_tiles(p) + kW - 1) / kW;
    const unsigned grid = want < resident ? want : resident;
    InvTmaParams q;
    if ((e = fill_tma_plane(q, p, grid * kW, 0)) != cudaSuccess) return e;
    if (segments) *segments = WorklistSegments{q.n_segs, q.seg_cap, 0};
    return launch_pdl(kernel, grid, CFG::kThreadsT, CFG::kSmem, s, q);
}

// several planes, one launch (k_dequant_idct_u8_tma_multi): each plane starts at the warp where the previous one ended
template <typename CFG, int NPL, typename K>
static cudaError_t launch_multi_tma(K kernel, const InvParams *pl, cudaStream_t s)
{
    constexpr int kW = CFG::kWarpsT;
    cudaError_t e = ensure_smem_attributes(reinterpret_cast<const void *>(kernel), CFG::kSmem);
    if (e != cudaSuccess) return e;
    const unsigned resident = (unsigned)sm_count_k2() * tma_ctas_per_sm<CFG>(kernel);
    unsigned total = 0;
    for (int i = 0; i < NPL; ++i) total += tma_tiles(pl[i]);
    const unsigned want = (total + kW - 1) / kW;
    const unsigned grid = want < resident ? want : resident;
    const unsigned n_segs = grid * kW;
    InvTmaMulti<NPL> m;
    unsigned rot = 0;
    for (int i = 0; i < NPL; ++i) {
        unsigned seg_off = 0;
        for (int j = 0; j < i; ++j)
            if (pl[j].worklist == pl[i].worklist) seg_off += (tma_tiles(pl[j]) + n_segs - 1) / n_segs * 32;
        if ((e = fill_tma_plane(m.pl[i], pl[i], n_segs, rot, seg_off)) != cudaSuccess)
            return NPL == 1 ? e : cudaErrorNotSupported;       // several planes: queued one by one instead
        rot = (rot + tma_tiles(pl[i])) % n_segs;
    }
    return launch_pdl(kernel, grid, CFG::kThreadsT, CFG::kSmem, s, m);
}

// small planes go through the frame kernel (replay folded into its tail: one launch instead of two)
template <typename CFG>
static cudaError_t launch_k2_tma(const InvParams &p, int layout, int adaptive, cudaStream_t s, WorklistSegments *segments, bool *folded)
{
    static const bool no_fold = getenv("DCT_CUDA_NO_FOLD") != nullptr;   // measurement aid
    if (!no_fold && p.tab != nullptr && p.ctr != nullptr && p.nblocks <= kFoldMaxBlocks) {
        if (folded) *folded = true;
        if (layout == LAYOUT_ZIGZAG)
            return adaptive ? launch_multi_tma<CFG, 1>(k_dequant_idct_u8_tma_frame<LAYOUT_ZIGZAG, true, CFG, 1>, &p, s)
                            : launch_multi_tma<CFG, 1>(k_dequant_idct_u8_tma_frame<LAYOUT_ZIGZAG, false, CFG, 1>, &p, s);
        return adaptive ? launch_multi_tma<CFG, 1>(k_dequant_idct_u8_tma_frame<LAYOUT_NATURAL, true, CFG, 1>, &p, s)
                        : launch_multi_tma<CFG, 1>(k_dequant_idct_u8_tma_frame<LAYOUT_NATURAL, false, CFG, 1>, &p, s);
    }
    if (layout == LAYOUT_ZIGZAG)
        return adaptive ? launch_persistent_tma<CFG>(k_dequant_idct_u8_tma<LAYOUT_ZIGZAG, true, CFG>, p, s, segments)
                        : launch_persistent_tma<CFG>(k_dequant_idct_u8_tma<LAYOUT_ZIGZAG, false, CFG>, p, s, segments);
    return adaptive ? launch_persistent_tma<CFG>(k_dequant_idct_u8_tma<LAYOUT_NATURAL, true, CFG>, p, s, segments)
                    : launch_persistent_tma<CFG>(k_dequant_idct_u8_tma<LAYOUT_NATURAL, false, CFG>, p, s, segments);
}

static bool tma_eligible(const InvParams &p)
{
    static const bool disabled = getenv("DCT_CUDA_NO_TMA") != nullptr;   // measurement aid: force the one-shot kernels
    if (disabled || p.no_tma || !tma_available() || p.seg_count == nullptr) return false;
    if (p.bw < 32 || ((uintptr_t)p.coef % 16)) return false;
    const unsigned long long padded = (unsigned long long)(p.nblocks / p.bw) * ((p.bw + 31) / 32) * 32;
    return padded + (unsigned long long)kMaxWorklistSegments * 64 <= p.wl_cap;
}

// 2 or 3 non-adaptive planes in one launch; cudaErrorNotSupported when the planes do not qualify (the caller then
// queues them one by one)
cudaError_t launch_dequant_idct_u8_multi(const InvParams *pl, int n, int layout, cudaStream_t s)
{
    static const bool off = getenv("DCT_CUDA_NO_MULTI") != nullptr || getenv("DCT_CUDA_NO_FOLD") != nullptr;   // measurement aids
    if (off || n < 2 || n > 3) return cudaErrorNotSupported;
    unsigned long long blocks = 0;
    for (int i = 0; i < n; ++i) {
        if (pl[i].nblocks == 0 || !tma_eligible(pl[i]) || pl[i].tab == nullptr || pl[i].ctr == nullptr || pl[i].nblocks > kFoldMaxBlocks)
            return cudaErrorNotSupported;
        blocks += pl[i].nblocks;
    }
    if (blocks > 2ull * kFoldMaxBlocks) return cudaErrorNotSupported;
    using CFG = TmaCfg<16, 1, 2>;
    if (layout == LAYOUT_ZIGZAG)
        return n == 2 ? launch_multi_tma<CFG, 2>(k_dequant_idct_u8_tma_frame<LAYOUT_ZIGZAG, false, CFG, 2>, pl, s)
                      : launch_multi_tma<CFG, 3>(k_dequant_idct_u8_tma_frame<LAYOUT_ZIGZAG, false, CFG, 3>, pl, s);
    return n == 2 ? launch_multi_tma<CFG, 2>(k_dequant_idct_u8_tma_frame<LAYOUT_NATURAL, false, CFG, 2>, pl, s)
                  : launch_multi_tma<CFG, 3>(k_dequant_idct_u8_tma_frame<LAYOUT_NATURAL, false, CFG, 3>, pl, s);
}

cudaError_t launch_dequant_idct_u8(const InvParams &p, int layout, int adaptive, cudaStream_t s, WorklistSegments *segments, bool *folded)
{
    if (folded) *folded = false;
    if (segments) *segments = WorklistSegments{0, 0, 0};
    if (p.nblocks == 0) return cudaSuccess;
    if (tma_eligible(p)) {
        // Geometry (measured on B200, 64 4K frames, profiles/r2_geometry.md): ONE CTA of 16 warps per SM with 128
        // registers per thread beats 3 x 8 warps at 80 registers (0.91 against 0.83 of the copy peak).
        // DCT_CUDA_K2_GEOMETRY=2 (tuning aid): 8 warps x 3 CTAs.
        static const int variant = getenv("DCT_CUDA_K2_GEOMETRY") ? atoi(getenv("DCT_CUDA_K2_GEOMETRY")) : 0;
        if (variant == 2) return launch_k2_tma<TmaCfg<8, 3, 2>>(p, layout, adaptive, s, segments, folded);
        return launch_k2_tma<TmaCfg<16, 1, 2>>(p, layout, adaptive, s, segments, folded);
    }
    const unsigned grid = (p.nblocks + kThreads - 1) / kThreads;
    if (layout == LAYOUT_ZIGZAG) {
        if (adaptive) k_dequant_idct_u8<LAYOUT_ZIGZAG, true><<<grid, kThreads, 0, s>>>(p);
        else          k_dequant_idct_u8<LAYOUT_ZIGZAG, false><<<grid, kThreads, 0, s>>>(p);
    } else {
        if (adaptive) k_dequant_idct_u8<LAYOUT_NATURAL, true><<<grid, kThreads, 0, s>>>(p);
        else          k_dequant_idct_u8<LAYOUT_NATURAL, false><<<grid, kThreads, 0, s>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace dctb
