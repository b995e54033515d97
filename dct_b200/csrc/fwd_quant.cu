// fwd_quant.cu -- K1: u8 pixels -> level shift -> 8x8 forward DCT -> quantise -> int16 records.
//
// Fuses, for every 8x8 block of a plane, the reference's
//   create_block_from_pixels (src/dct.c:109-120), dct_forward (src/dct.c:52-77),
//   quantize (src/quantization.c:113-131; adaptive: calculate_block_variance :153-169 and
//   adjust_matrix_for_block :171-211) and, for the ZIGZAG layout, block_to_zigzag
//   (src/entropy.c:158-178)
// into one pass: each block is read once (64 B) and written once (128 B).
//
// Mapping: one thread owns one block; a warp owns a TILE of 32 blocks, i.e. a 256-pixel wide, 8-row strip.  A lane
// keeps its block in 64 registers through both butterfly passes (no transpose, no shuffles); both passes and the
// quantisation run on sm_100's packed fp32 instructions (FADD2 / FFMA2: two IEEE lanes per instruction, half the
// issue slots of scalar code, bit-identical results): one FFMA2 quantises two coefficients (round-to-nearest via the
// 1.5*2^23 trick) and the fp32 residual is checked against the band (fwd_block below, shared by both kernels).
// Blocks with a coefficient inside the band go to the warp's worklist segment and are replayed in fp64 (K3).
//
// Two kernels move the tiles:
//   k_fwd_quant_u8_tma  (default)  persistent, ONE CTA of 12 warps per SM; per warp a two-stage bulk-tensor pipeline:
//                       cp.async.bulk.tensor.2d (UTMALDG) brings a 256 B x 8-row box of the plane per tile, the 32
//                       records leave through a 128B-swizzled stage with one bulk-tensor store (UTMASTG).
//                       Small planes replay their flagged blocks in the kernel's tail (FOLD, replay_lane.cuh).
//   k_fwd_quant_u8      (fallback: pitch not a multiple of 16, unaligned base, planes under 256 pixels wide, peer
//                       memory)  persistent, 3 CTAs of 8 warps per SM, per-warp two-stage cp.async pipeline, records
//                       through a padded stage so that every STG.128 of the warp covers 512 contiguous bytes.
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
constexpr int kStageWordsPerBlock = 36;   // 128 B record + 16 B pad: conflict-free STS.128 / LDS.128
constexpr int kStageWordsPerWarp = 32 * kStageWordsPerBlock;


constexpr int kInStages = 2;                          // tiles in flight per warp
constexpr int kInWordsPerStage = 8 * 64;              // 8 rows x 256 B
constexpr int kSmemWordsPerWarp = kInStages * kInWordsPerStage + kStageWordsPerWarp;
constexpr int kSmemBytes = kWarps * kSmemWordsPerWarp * 4;   // 69 632 B per CTA -> 3 CTAs per SM

__device__ __forceinline__ void cp_async_8(uint32_t smem_addr, const void *gptr)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void stg_stream_u4(void *p, const uint4 &v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

// The arithmetic of one block, shared by the cp.async and the bulk-tensor kernels: raw[i] = the 8 bytes of pixel
// row i; w[m] = the packed int16 pair m of the record in storage order; returns "replay this block in fp64".
template <int LAYOUT, bool ADAPTIVE, bool UNIFORM>
__device__ __forceinline__ bool fwd_block(const FwdParams &p, const uint2 (&raw)[8], uint32_t (&w)[32], bool valid, uint32_t b)
{
    float inv_s = 1.0f;
    if constexpr (ADAPTIVE) {
        // Sum and sum of squares of the centred samples, as exact integers: packed byte dot products.
        // 4096 * variance = 64 * sum(x^2) - sum(x)^2 is exact, so the side array equals
        // calculate_block_variance (src/quantization.c:153-169) bit for bit.
        int isum = 0, isq = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) row_moments(raw[i], isum, isq);
        const int num = 64 * isq - isum * isum;   // 4096 * variance, exact (< 2^27)
        if (p.var_out != nullptr && valid) p.var_out[b] = (double)num * (1.0 / 4096.0);
        // s = 2 - clamp(var/1000, 0.1, 1)  (src/quantization.c:186-190), fp32 here, exact in K3
        inv_s = adaptive_inv_scale(num);
    }

    // Rows (X * D^T), then columns (D * temp): same order as src/dct.c:57-74.  Both passes run on packed
    // pairs (FADD2 / FFMA2: two fp32 lanes per instruction, half the issue slots).  The row pass takes rows
    // a and 7-a in its two lanes -- exactly the pairs the first butterfly stage of the column pass adds and
    // subtracts, so that stage is 64 scalar FADDs on the two halves of a register pair, whose results are
    // written straight into (column 2b, column 2b+1) pairs: the 2x2 re-pairing costs no instruction.
    float2 rp[4][8];                         // rp[a][k] = (T[a][k], T[7-a][k])
#pragma unroll
    for (int a = 0; a < 4; ++a) fdct8_rowpair_from_bytes(rp[a], raw[a], raw[7 - a]);
    float2 cp[4][8];                         // cp[b][u] = scaled coefficients (u, 2b) and (u, 2b+1)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float2 s[4], d[4];                   // s[a] = T[a] + T[7-a], d[a] = T[a] - T[7-a] for columns 2b, 2b+1
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            s[a] = make_float2(__fadd_rn(rp[a][2 * b].x, rp[a][2 * b].y), __fadd_rn(rp[a][2 * b + 1].x, rp[a][2 * b + 1].y));
            d[a] = make_float2(__fsub_rn(rp[a][2 * b].x, rp[a][2 * b].y), __fsub_rn(rp[a][2 * b + 1].x, rp[a][2 * b + 1].y));
        }
        fdct8_tail<float2, 1>(cp[b], s[0], s[1], s[2], s[3], d[0], d[1], d[2], d[3]);
    }

    // Quantise in natural pairs (k, k+1) = cp[(k%8)/2][k/8]: t = c*r + 1.5*2^23 holds round(c*r) in its low
    // mantissa bits; residual e = c*r - round(c*r) (one rounding); |e| >= 0.5 - band  => replay.
    // t overwrites cp; the layout (natural / zigzag) is applied when the int16 halves are packed.
    bool flag = false;
    float emax = 0.0f;
    static_for<0, 32>([&](auto M) {
        constexpr int m = decltype(M)::value;      // natural pair: coefficients 2m, 2m+1
        constexpr int u = m >> 2, b = m & 3;
        const float2 r2 = reinterpret_cast<const float2 *>(p.r)[m];
        if constexpr (ADAPTIVE) {   // the block's 1/(2 - nv) goes onto the coefficient (in place), the table stays a constant operand
            if (m == 0) cp[b][u].y = __fmul_rn(cp[b][u].y, inv_s);       // DC keeps the unscaled table entry
            else cp[b][u] = Ops<float2>::mul(cp[b][u], make_float2(inv_s, inv_s));
        }
        float2 t2, e2;
        quant_residual2(cp[b][u], r2, t2, e2);
        if constexpr (UNIFORM) emax = fmaxf(fmaxf(emax, fabsf(e2.x)), fabsf(e2.y));   // FMNMX3
        else flag |= (fabsf(e2.x) >= p.thr[2 * m]) | (fabsf(e2.y) >= p.thr[2 * m + 1]);
        cp[b][u] = t2;
    });
    static_for<0, 32>([&](auto M) {
        constexpr int m = decltype(M)::value;      // storage pair
        constexpr int k0 = storage_to_natural<LAYOUT>(2 * m), k1 = storage_to_natural<LAYOUT>(2 * m + 1);
        const float2 a2 = cp[(k0 & 7) >> 1][k0 >> 3], b2 = cp[(k1 & 7) >> 1][k1 >> 3];
        w[m] = __byte_perm(__float_as_uint((k0 & 1) ? a2.y : a2.x), __float_as_uint((k1 & 1) ? b2.y : b2.x), 0x5410);
    });

    if constexpr (UNIFORM) flag = emax >= p.thr_min;

    return flag;
}

// UNIFORM: one band for all 64 coefficients (the widest), tested with 3-input max -- half the
// instructions of the per-coefficient compare; chosen by the host when the widest band is small.
template <int LAYOUT, bool ADAPTIVE, bool UNIFORM>
__global__ void __launch_bounds__(kThreads, 3) k_fwd_quant_u8(const __grid_constant__ FwdParams p)
{
    extern __shared__ uint4 smem_dyn[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wsm = reinterpret_cast<uint32_t *>(smem_dyn) + warp * kSmemWordsPerWarp;
    uint32_t *wstage = wsm + kInStages * kInWordsPerStage;                       // output stage
    const uint32_t in_addr = (uint32_t)__cvta_generic_to_shared(wsm) + lane * 8;   // + stage*2048 + row*256

    // loop bounds straight from the parameter bank (set by the launcher): no registers held for them
#define ntiles p.ntiles
#define tile_stride p.tile_stride
    uint32_t tile = blockIdx.x * kWarps + warp;

    // Each lane copies the 8 rows of its own block (lane-private data: no cross-lane hazard on the
    // input).  The lane's block coordinates advance by a fixed step from tile to tile, so the only
    // integer divisions of the kernel are the two below.
    // (step_q, step_r) = divmod(32 * tiles per grid sweep, blocks per row) comes from the host with the launch.
    uint32_t nby, nbx;                                          // block row / column the next issue() fetches
    {
        const uint32_t nb = tile * 32 + lane;
        nby = nb / p.bw;
        nbx = nb - nby * p.bw;
    }
    auto issue = [&](int stage) {
        if (nby * p.bw + nbx < p.nblocks) {   // lanes past the end keep stale (but valid u8) data and store nothing
            const uint8_t *src = p.px + ((long long)nby * p.pitch + nbx) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) cp_async_8(in_addr + stage * (kInWordsPerStage * 4) + i * 256, src + i * p.pitch);
        }
        cp_async_commit();
        nbx += p.step_r;
        nby += p.step_q;
        if (nbx >= p.bw) nbx -= p.bw, ++nby;
    };

    // entries this warp has appended to its worklist segment: kept in a padding word of the warp's output stage
    // (words 32..35 of a record row are never written), not in a register -- the kernel sits on its 80-register budget
    if (lane == 0) wstage[32] = 0;
    __syncwarp();

    int stage = 0;
    if (tile < ntiles) issue(0);
    for (; tile < ntiles; tile += tile_stride, stage ^= 1) {
    if (tile + tile_stride < ntiles) {
        issue(stage ^ 1);
        cp_async_wait<1>();
    } else {
        cp_async_wait<0>();
    }
    const uint32_t warp_base = tile * 32;
    const uint32_t b = warp_base + lane;
    const bool valid = b < p.nblocks;

    uint2 raw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        raw[i] = *reinterpret_cast<const uint2 *>(wsm + stage * kInWordsPerStage + i * 64 + lane * 2);

    uint32_t w[32];
    const bool flag = fwd_block<LAYOUT, ADAPTIVE, UNIFORM>(p, raw, w, valid, b);

    const unsigned ballot = __ballot_sync(0xffffffffu, flag && valid);

    // stage: lane-major padded records in shared memory, then 512-byte contiguous warp stores
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4 *>(wstage + lane * kStageWordsPerBlock + 4 * j) =
            make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
    __syncwarp();
    uint4 *dst = reinterpret_cast<uint4 *>(p.coef) + (size_t)warp_base * 8 + lane;
    const uint32_t *rd = wstage + (lane >> 3) * kStageWordsPerBlock + 4 * (lane & 7);   // chunk c = j*32 + lane
    const uint32_t full = p.nblocks - warp_base;      // records in this tile
    if (full >= 32) {                                 // warp-uniform: every tile but possibly the last
#pragma unroll
        for (int j = 0; j < 8; ++j)
            stg_stream_u4(dst + j * 32, *reinterpret_cast<const uint4 *>(rd + j * 4 * kStageWordsPerBlock));
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j * 4 + (lane >> 3) < full)
                stg_stream_u4(dst + j * 32, *reinterpret_cast<const uint4 *>(rd + j * 4 * kStageWordsPerBlock));
    }

    if (ballot != 0) {
        // The warp appends to its own segment of the worklist: no global atomic (a same-address atomic per flagged
        // tile was measured to cost ~5 us per tile once a process had several GB of planes in play).  One round per
        // flagged block (usually a single one): lane 8 writes the entry, lanes 0-7 copy one pixel row each from this
        // tile's input stage to the 64 bytes that go with the entry (K3 reads those instead of eight scattered
        // rows of the plane).
        const uint32_t gwarp = blockIdx.x * kWarps + warp;
        uint32_t wl_n = wstage[32];
        __syncwarp();
        if (lane == 0) wstage[32] = wl_n + __popc(ballot);
        for (unsigned todo = ballot; todo != 0; todo &= todo - 1, ++wl_n) {
            const unsigned f = __ffs(todo) - 1;
            if (lane < 8) {
                if (wl_n < p.side_seg_cap)
                    reinterpret_cast<uint2 *>(p.side + ((size_t)gwarp * p.side_seg_cap + wl_n) * 64)[lane] =
                        *reinterpret_cast<const uint2 *>(wsm + stage * kInWordsPerStage + lane * 64 + f * 2);
            } else if (lane == 8) {
                p.worklist[(size_t)gwarp * p.seg_cap + wl_n] = warp_base + f;
            }
        }
    }
    __syncwarp();   // the output stage is rewritten by the next tile
    }   // tile loop
    if (lane == 0) p.seg_count[blockIdx.x * kWarps + warp] = wstage[32];   // < seg_cap by construction (32 per tile)
}


#undef ntiles
#undef tile_stride

// ------------------------------------------------------------------------------------------
// K1 with bulk-tensor (TMA) tile movement -- the default whenever the plane allows it (16-byte aligned base,
// pitch a multiple of 16, at least 256 pixels wide).  Same mapping and arithmetic as k_fwd_quant_u8 above;
// what changes is how the tiles move:
//   in   one cp.async.bulk.tensor.2d per tile, issued by lane 0: a 256-byte x 8-row box of the pixel plane lands in
//        the warp's input stage and signals the stage's mbarrier (two stages: tile t+1 is in flight while tile t is
//        transformed).  Columns past the plane's width arrive as zeros, so a block row whose width is not a multiple
//        of 256 simply ends with a partial tile: tiles never straddle block rows here (tile = 32 blocks of ONE block row).
//   out  each lane writes its 128-byte record into a 128B-swizzled 4 KB stage (8 conflict-free STS.128), lane 0 hands
//        the stage to the copy engine (cp.async.bulk.tensor.2d shared -> global); the store drains while the warp is
//        already transforming the next tile.  Partial tiles use a second tensor map whose box holds bw % 32 records.
// Against the cp.async kernel this removes 8 LDGSTS, 8 LDS.128 and 8 STG.128 per lane and the address arithmetic
// that went with them.
// ------------------------------------------------------------------------------------------
struct alignas(64) FwdTmaParams {
    CUtensorMap map_px;        // uint8 plane, box 256 x 8
    CUtensorMap map_rec;       // records, box 128 B x 32, 128-byte swizzle
    CUtensorMap map_rec_tail;  // records, box 128 B x (bw % 32)
    FwdParams f;
    uint32_t tpr;              // tiles per block row = ceil(bw / 32)
    uint32_t nby;              // block rows
    uint32_t step_ty, step_tx; // divmod(warps in the grid, tpr): how a warp's tile coordinates advance
    uint32_t n_segs, rot;      // warps in the grid; the warp that takes the plane's tile 0
};

constexpr int kTmaOutBytes = 4096;                                  // per warp: 32 records
constexpr int kTmaInBytes = 2048;                                   // per warp and stage: 8 rows x 256 B
constexpr int kTmaCtlBytes = 64;                                    // per warp: up to 7 mbarriers, then the worklist count
constexpr int kTmaMaxPlanes = 3;                                    // planes one launch may cover
constexpr int kTmaCntBytes = 256;                                   // per CTA: flagged blocks per (plane, warp), up to 3 x 16

// geometry of one kernel variant: warps per CTA, CTAs per SM the register budget is cut for, input stages per warp
template <int WARPS, int MIN_CTAS, int STAGES> struct TmaCfg {
    static constexpr int kWarpsT = WARPS, kMinCtas = MIN_CTAS, kStages = STAGES, kThreadsT = 32 * WARPS;
    static constexpr int kSmem = 1024 /* alignment slack */ + WARPS * (kTmaOutBytes + STAGES * kTmaInBytes + kTmaCtlBytes) + kTmaCntBytes;
    static_assert(STAGES >= 2 && STAGES <= 7, "stages");
    static_assert(kTmaMaxPlanes * WARPS * 4 <= kTmaCntBytes, "per-CTA counts");
};

// What a warp carries from one plane of a launch to the next: the parity bit of each stage's mbarrier (bit i: stage i).
struct WarpPipe {
    uint32_t phases;
};

// All the tiles of ONE plane that fall to this warp; returns how many blocks it flagged.  The single-plane
// kernel calls it once; the multi-plane kernel once per plane, the parameters of each plane being compile-time offsets
// into the kernel's parameter block (so the tables stay constant-bank operands).
template <int LAYOUT, bool ADAPTIVE, bool UNIFORM, typename CFG>
__device__ __forceinline__ uint32_t fwd_plane_tiles(const FwdTmaParams &P, const uint32_t lane, const uint32_t gwarp, uint8_t *out_p,
                                                uint8_t *in_p, uint32_t *cnt_p, const uint32_t bar_s, WarpPipe &pipe)
{
    constexpr int kS = CFG::kStages;
    const FwdParams &p = P.f;
    const uint32_t out_s = (uint32_t)__cvta_generic_to_shared(out_p), in_s = (uint32_t)__cvta_generic_to_shared(in_p);

    // (ty, tx): block row and tile-in-row of the tile being transformed; (fy, fx): of the next tile to fetch, kS - 1 ahead.
    // The plane's tile 0 belongs to warp `rot` of the grid (0 unless planes share the launch: each starts where the
    // previous one ended, so the warps' tile counts differ by at most one over the whole launch).
    uint32_t ty, tx;
    {
        const uint32_t t = gwarp >= P.rot ? gwarp - P.rot : gwarp + P.n_segs - P.rot;
        ty = t / P.tpr;
        tx = t - ty * P.tpr;
    }
    uint32_t fy = ty, fx = tx, fstage = 0;
    auto fetch = [&]() {        // fetch tile (fy, fx) into stage fstage, then advance both
        if (fy < P.nby && lane == 0) {
            tma::mbar_expect_tx(bar_s + fstage * 8, kTmaInBytes);
            tma::load_2d(in_s + fstage * kTmaInBytes, &P.map_px, (int)(fx * 256), (int)(fy * 8), bar_s + fstage * 8);
        }
        fx += P.step_tx;
        fy += P.step_ty;
        if (fx >= P.tpr) fx -= P.tpr, ++fy;
        fstage = fstage + 1 == kS ? 0 : fstage + 1;
    };
#pragma unroll
    for (int i = 0; i < kS - 1; ++i) fetch();
    // my chunk j goes to 16-byte slot j ^ (lane & 7) of my 128-byte row
    uint8_t *const my_out = out_p + lane * 128;
    const uint32_t swz = (lane & 7) << 4;

    uint32_t stage = 0;
    while (ty < P.nby) {
        const uint32_t bx0 = tx * 32;
        const uint32_t warp_base = ty * p.bw + bx0;               // first record of this tile
        const uint32_t nvalid = min(32u, p.bw - bx0);
        fetch();                                                  // into the stage the previous iteration consumed
        tma::mbar_wait(bar_s + stage * 8, (pipe.phases >> stage) & 1u);
        pipe.phases ^= 1u << stage;

        const uint32_t b = warp_base + lane;
        const bool valid = lane < nvalid;
        const uint8_t *in_stage = in_p + stage * kTmaInBytes;
        uint2 raw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) raw[i] = *reinterpret_cast<const uint2 *>(in_stage + i * 256 + lane * 8);

        uint32_t w[32];
        const bool flag = fwd_block<LAYOUT, ADAPTIVE, UNIFORM>(p, raw, w, valid, b);
        const unsigned ballot = __ballot_sync(0xffffffffu, flag && valid);

        // the previous tile's store must have finished reading the stage before it is rewritten
        if (lane == 0) tma::store_wait_read();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4 *>(my_out + ((j << 4) ^ swz)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        tma::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma::store_2d(nvalid == 32 ? &P.map_rec : &P.map_rec_tail, 0, (int)warp_base, out_s);
            tma::store_commit();
        }

        if (ballot != 0) {
            // The warp appends to its own segment of the worklist (no global atomic); lane 8 writes the entry,
            // lanes 0-7 copy one pixel row each from the input stage to the 64 bytes that go with it (read by K3).
            uint32_t wl_n = *cnt_p;
            __syncwarp();
            if (lane == 0) *cnt_p = wl_n + __popc(ballot);
            for (unsigned todo = ballot; todo != 0; todo &= todo - 1, ++wl_n) {
                const unsigned f = __ffs(todo) - 1;
                if (lane < 8) {
                    if (wl_n < p.side_seg_lim)
                        reinterpret_cast<uint2 *>(p.side + ((size_t)gwarp * p.side_seg_cap + wl_n) * 64)[lane] =
                            *reinterpret_cast<const uint2 *>(in_stage + lane * 256 + f * 8);
                } else if (lane == 8) {
                    p.worklist[(size_t)gwarp * p.seg_cap + wl_n] = warp_base + f;
                }
            }
        }
        __syncwarp();   // every lane is done with this input stage: the next fetch overwrites it
        tx += P.step_tx;
        ty += P.step_ty;
        if (tx >= P.tpr) tx -= P.tpr, ++ty;
        if (++stage == kS) stage = 0;
    }
    // entries this warp appended to its segment for this plane; the count in shared memory starts the next plane at zero
    __syncwarp();
    const uint32_t n_mine = *cnt_p;
    __syncwarp();
    if (lane == 0) {
        p.seg_count[gwarp] = 0;                       // replayed before the kernel ends (fwd_replay_pooled): nothing left for K3
        *cnt_p = 0;
    }
    __syncwarp();
    return n_mine;
}

// FOLD (small planes, where a launch costs more than the work): the flagged blocks are replayed in the kernel's tail and no
// K3 is launched.  The CTA pools the entries of all its warps (and of all the planes of the launch) and deals them out in
// batches of 32, one block per lane (replay_lane.cuh): a warp of a 4K plane flags three or four blocks, and a pass costs
// the same for 3 lanes as for 32.  The record stage of each warp is the scratch of its passes.
// Large planes leave the segments to K3, whose grid of replay-only warps hides the replay's latency better than a
// tile-loop warp that stops to do it.
template <int LAYOUT, bool ADAPTIVE, typename CFG, int NPL>
__device__ __forceinline__ void fwd_replay_pooled(const FwdTmaParams *pl, const uint32_t (&n_mine)[NPL], uint32_t *cnt_all, const uint32_t lane,
                                                  const uint32_t warp, uint8_t *out_p)
{
    constexpr int kW = CFG::kWarpsT;
    if (lane == 0) {
        // the patches are generic-proxy stores to records that bulk stores wrote: those must be complete, not just read
        tma::store_wait_all();
        asm volatile("fence.proxy.async;" ::: "memory");
        static_for<0, NPL>([&](auto I) {
            constexpr int i = decltype(I)::value;
            cnt_all[i * kW + warp] = n_mine[i];
            if (n_mine[i] != 0) atomicAdd(&pl[i].f.ctr->replayed, (unsigned long long)n_mine[i]);
        });
    }
    __syncthreads();
    uint32_t total = 0;
    for (int it = 0; it < NPL * kW; ++it) total += cnt_all[it];
    LaneScratch *ws = reinterpret_cast<LaneScratch *>(out_p);
    for (uint32_t first = warp * 32; first < total; first += kW * 32) {
        int item = 0;
        uint32_t e = 0;
        const bool active = locate_entry(cnt_all, NPL * kW, first + lane, item, e);
        const int plane = item / kW;
        const uint32_t gw = blockIdx.x * kW + (uint32_t)(item - plane * kW);     // the warp that flagged the block
        FwdReplayCtx cx{};
        uint32_t b = 0;
        const uint8_t *src = nullptr;
        long long src_pitch = 8;
        static_for<0, NPL>([&](auto I) {
            constexpr int i = decltype(I)::value;
            if (i == 0 || plane == i) {       // idle lanes run the arithmetic on plane 0's tables
                const FwdParams &p = pl[i].f;
                cx = FwdReplayCtx{p.r, p.thr, p.tab->D, p.tab->Q, ADAPTIVE ? 1 : 0, p.coef, p.ctr};
                if (active && plane == i) {
                    b = p.worklist[(size_t)gw * p.seg_cap + e];
                    const bool side = e < p.side_seg_lim;      // the block's pixels sit next to its entry (8-byte rows), else in the plane
                    const uint32_t by = b / p.bw, bx = b - by * p.bw;
                    src = side ? p.side + ((size_t)gw * p.side_seg_cap + e) * 64 : p.px + ((long long)by * p.pitch + bx) * 8;
                    src_pitch = side ? 8 : p.pitch;
                }
            }
        });
        replay_fwd_lanes<LAYOUT>(cx, ws, active, b, src, src_pitch);
    }
}

// the warp's shared memory: [record stages, 1024-aligned for the swizzle][pixel stages][per warp: mbarriers + count]
template <typename CFG> struct FwdWarpSmem {
    uint8_t *out_p, *in_p;
    uint32_t *cnt_p;
    uint32_t *cnt_all;     // CTA-wide: blocks flagged per (plane, warp), for the pooled replay
    uint32_t bar_s;
    __device__ __forceinline__ FwdWarpSmem(uint8_t *smem_raw, uint32_t warp, uint32_t lane)
    {
        constexpr int kW = CFG::kWarpsT, kS = CFG::kStages;
        uint8_t *sm = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
        out_p = sm + warp * kTmaOutBytes;
        in_p = sm + kW * kTmaOutBytes + warp * (kS * kTmaInBytes);
        uint32_t *ctl_p = reinterpret_cast<uint32_t *>(sm + kW * (kTmaOutBytes + kS * kTmaInBytes) + warp * kTmaCtlBytes);
        cnt_p = ctl_p + 2 * kS;                                       // after the kS 8-byte mbarriers
        cnt_all = reinterpret_cast<uint32_t *>(sm + kW * (kTmaOutBytes + kS * kTmaInBytes + kTmaCtlBytes));
        bar_s = (uint32_t)__cvta_generic_to_shared(ctl_p);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < kS; ++i) tma::mbar_init(bar_s + 8 * i, 1);
            *cnt_p = 0;                      // entries this warp has appended to its worklist segment
            tma::fence_barrier_init();
        }
        __syncwarp();
    }
};

// The streaming kernel: one large plane, flagged blocks left to K3.  (Kept apart from the frame kernel below, which shares
// its tile loop in spirit but not in text: the schedule of this loop is tuned, and any code around it moves it.)
template <int LAYOUT, bool ADAPTIVE, bool UNIFORM, typename CFG>
__global__ void __launch_bounds__(CFG::kThreadsT, CFG::kMinCtas) k_fwd_quant_u8_tma(const __grid_constant__ FwdTmaParams P)
{
    constexpr int kW = CFG::kWarpsT, kS = CFG::kStages;
    extern __shared__ uint8_t smem_raw[];
    const FwdParams &p = P.f;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // tells the compiler it is warp-uniform
    // carve: [record stages, 1024-aligned for the swizzle][pixel stages][per warp: mbarriers + count]
    uint8_t *sm = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    uint8_t *out_p = sm + warp * kTmaOutBytes;
    uint8_t *in_p = sm + kW * kTmaOutBytes + warp * (kS * kTmaInBytes);
    uint32_t *ctl_p = reinterpret_cast<uint32_t *>(sm + kW * (kTmaOutBytes + kS * kTmaInBytes) + warp * kTmaCtlBytes);
    uint32_t *const cnt_p = ctl_p + 2 * kS;                                       // after the mbarriers
    const uint32_t out_s = (uint32_t)__cvta_generic_to_shared(out_p), in_s = (uint32_t)__cvta_generic_to_shared(in_p);
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(ctl_p);           // kS 8-byte mbarriers

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kS; ++i) tma::mbar_init(bar_s + 8 * i, 1);
        *cnt_p = 0;                          // entries this warp has appended to its worklist segment
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
            tma::mbar_expect_tx(bar_s + fstage * 8, kTmaInBytes);
            tma::load_2d(in_s + fstage * kTmaInBytes, &P.map_px, (int)(fx * 256), (int)(fy * 8), bar_s + fstage * 8);
        }
        fx += P.step_tx;
        fy += P.step_ty;
        if (fx >= P.tpr) fx -= P.tpr, ++fy;
        fstage = fstage + 1 == kS ? 0 : fstage + 1;
    };
#pragma unroll
    for (int i = 0; i < kS - 1; ++i) fetch();
    // my chunk j goes to 16-byte slot j ^ (lane & 7) of my 128-byte row
    uint8_t *const my_out = out_p + lane * 128;
    const uint32_t swz = (lane & 7) << 4;

    uint32_t stage = 0, phase = 0;
    while (ty < P.nby) {
        const uint32_t bx0 = tx * 32;
        const uint32_t warp_base = ty * p.bw + bx0;               // first record of this tile
        const uint32_t nvalid = min(32u, p.bw - bx0);
        fetch();                                                  // into the stage the previous iteration consumed
        tma::mbar_wait(bar_s + stage * 8, phase);

        const uint32_t b = warp_base + lane;
        const bool valid = lane < nvalid;
        const uint8_t *in_stage = in_p + stage * kTmaInBytes;
        uint2 raw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) raw[i] = *reinterpret_cast<const uint2 *>(in_stage + i * 256 + lane * 8);

        uint32_t w[32];
        const bool flag = fwd_block<LAYOUT, ADAPTIVE, UNIFORM>(p, raw, w, valid, b);
        const unsigned ballot = __ballot_sync(0xffffffffu, flag && valid);

        // the previous tile's store must have finished reading the stage before it is rewritten
        if (lane == 0) tma::store_wait_read();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4 *>(my_out + ((j << 4) ^ swz)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        tma::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma::store_2d(nvalid == 32 ? &P.map_rec : &P.map_rec_tail, 0, (int)warp_base, out_s);
            tma::store_commit();
        }

        if (ballot != 0) {
            // The warp appends to its own segment of the worklist (no global atomic); lane 8 writes the entry,
            // lanes 0-7 copy one pixel row each from the input stage to the 64 bytes that go with it (read by K3).
            const uint32_t gwarp = blockIdx.x * kW + warp;
            uint32_t wl_n = *cnt_p;
            __syncwarp();
            if (lane == 0) *cnt_p = wl_n + __popc(ballot);
            for (unsigned todo = ballot; todo != 0; todo &= todo - 1, ++wl_n) {
                const unsigned f = __ffs(todo) - 1;
                if (lane < 8) {
                    if (wl_n < p.side_seg_cap)
                        reinterpret_cast<uint2 *>(p.side + ((size_t)gwarp * p.side_seg_cap + wl_n) * 64)[lane] =
                            *reinterpret_cast<const uint2 *>(in_stage + lane * 256 + f * 8);
                } else if (lane == 8) {
                    p.worklist[(size_t)gwarp * p.seg_cap + wl_n] = warp_base + f;
                }
            }
        }
        __syncwarp();   // every lane is done with this input stage: the next fetch overwrites it
        tx += P.step_tx;
        ty += P.step_ty;
        if (tx >= P.tpr) tx -= P.tpr, ++ty;
        if (++stage == kS) stage = 0, phase ^= 1;
    }
    // the flagged blocks are left to K3, whose grid of replay-only warps hides the replay's latency better than a tile-loop
    // warp that stops to do it (small planes: k_fwd_quant_u8_tma_frame below)
    if (lane == 0) {
        p.seg_count[blockIdx.x * kW + warp] = *cnt_p;
        tma::store_wait_read();               // shared memory must outlive the last store's reads
    }
}

// The frame kernel: one small plane, or the 2 or 3 planes of one frame (Y, Cb, Cr) in ONE launch: a persistent grid's ramp
// and tail cost more than the tiles of a 4K plane, so a frame's planes share them, and the flagged blocks of all of them
// are replayed in one pooled tail.  Every plane has its own worklist range (the host offsets planes of one plan).
template <int NPL> struct alignas(64) FwdTmaMulti {
    FwdTmaParams pl[NPL];
};

template <int LAYOUT, bool ADAPTIVE, bool UNIFORM, typename CFG, int NPL>
__global__ void __launch_bounds__(CFG::kThreadsT, CFG::kMinCtas) k_fwd_quant_u8_tma_frame(const __grid_constant__ FwdTmaMulti<NPL> M)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const FwdWarpSmem<CFG> sm(smem_raw, warp, lane);
    pdl_launch_dependents();
    pdl_wait();
    WarpPipe pipe{0};
    uint32_t n_mine[NPL];
    static_for<0, NPL>([&](auto I) {
        constexpr int i = decltype(I)::value;
        n_mine[i] = fwd_plane_tiles<LAYOUT, ADAPTIVE, UNIFORM, CFG>(M.pl[i], lane, blockIdx.x * CFG::kWarpsT + warp, sm.out_p, sm.in_p,
                                                                   sm.cnt_p, sm.bar_s, pipe);
    });
    fwd_replay_pooled<LAYOUT, ADAPTIVE, CFG, NPL>(M.pl, n_mine, sm.cnt_all, lane, warp, sm.out_p);
}


// ------------------------------------------------------------------------------------------
// K1 from FLOAT pixel tiles (north_star: "8-bit or float pixel tiles"): block = (double)p - 128.0
// for arbitrary float p, then dct_forward + quantize -- what a caller of the reference gets by
// filling the input block by hand (tests/test_dct.c:46-50).  256 B in + 128 B out per block.
// One-shot grid, 16 LDG.128 per lane (a warp instruction covers 1 KB contiguous); the loaded
// registers ARE the block.  The band table assumes p in [0, 255] (|p - 128| <= 128, input rounding
// <= 128 u); blocks with a pixel outside that range, or NaN, are replayed whole in fp64.
// Non-adaptive plans only (the host rejects the others).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const void *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

template <int LAYOUT>
__global__ void __launch_bounds__(kThreads) k_fwd_quant_f32(const __grid_constant__ FwdParams p)
{
    __shared__ uint4 stage[kWarps * kStageWordsPerWarp / 4];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warp_base = blockIdx.x * kThreads + warp * 32;
    const uint32_t b = warp_base + lane;
    const bool valid = b < p.nblocks;
    const uint32_t bb = valid ? b : p.nblocks - 1;
    const uint32_t by = bb / p.bw, bx = bb - by * p.bw;
    const uint8_t *src = p.px + (long long)by * 8 * p.pitch + (long long)bx * 32;

    float v[64];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 lo = ldg_stream_f4(src + i * p.pitch), hi = ldg_stream_f4(src + i * p.pitch + 16);
        v[8 * i + 0] = lo.x, v[8 * i + 1] = lo.y, v[8 * i + 2] = lo.z, v[8 * i + 3] = lo.w;
        v[8 * i + 4] = hi.x, v[8 * i + 5] = hi.y, v[8 * i + 6] = hi.z, v[8 * i + 7] = hi.w;
    }
    float amax = 0.0f;
#pragma unroll
    for (int k = 0; k < 64; k += 2) {
        v[k] = centre_float_pixel(v[k]);
        v[k + 1] = centre_float_pixel(v[k + 1]);
        amax = fmaxf(fmaxf(amax, fabsf(v[k])), fabsf(v[k + 1]));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) fdct8<float, 1>(&v[8 * i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) fdct8<float, 8>(&v[j]);

    uint32_t w[32];
    bool flag = !(amax <= 128.0f);   // outside the band table's domain (or NaN): whole block to fp64
    static_for<0, 32>([&](auto M) {
        constexpr int m = decltype(M)::value;
        constexpr int k0 = storage_to_natural<LAYOUT>(2 * m), k1 = storage_to_natural<LAYOUT>(2 * m + 1);
        float t0, t1, e0, e1;
        quant_residual(v[k0], p.r[k0], t0, e0);
        quant_residual(v[k1], p.r[k1], t1, e1);
        flag |= (fabsf(e0) >= p.thr[k0]) | (fabsf(e1) >= p.thr[k1]);
        w[m] = __byte_perm(__float_as_uint(t0), __float_as_uint(t1), 0x5410);
    });

    uint32_t *wstage = reinterpret_cast<uint32_t *>(stage) + warp * kStageWordsPerWarp;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4 *>(wstage + lane * kStageWordsPerBlock + 4 * j) =
            make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
    __syncwarp();
    uint4 *dst = reinterpret_cast<uint4 *>(p.coef) + (size_t)warp_base * 8 + lane;
    const uint32_t *rd = wstage + (lane >> 3) * kStageWordsPerBlock + 4 * (lane & 7);
    const uint32_t full = warp_base < p.nblocks ? p.nblocks - warp_base : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j * 4 + (lane >> 3) < full)
            stg_stream_u4(dst + j * 32, *reinterpret_cast<const uint4 *>(rd + j * 4 * kStageWordsPerBlock));

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

static int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

template <typename K>
static cudaError_t launch_persistent(K kernel, const FwdParams &p, cudaStream_t s, unsigned *launches, WorklistSegments *segments)
{
    // NB: K is the same function-pointer type for every kernel variant, so this template has ONE instance and its
    // statics are shared by all variants: nothing per-kernel may be cached here (the attributes are set on every launch;
    // per_sm is the same for all variants: identical launch bounds and shared memory).
    cudaError_t e = ensure_smem_attributes(reinterpret_cast<const void *>(kernel), kSmemBytes);
    if (e != cudaSuccess) return e;
    static int per_sm = 0;   // resident CTAs per SM for this instantiation (3 by design: registers and shared memory)
    if (per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, kSmemBytes) != cudaSuccess || n < 1) n = 1;
        per_sm = n;
        if (getenv("DCT_CUDA_DEBUG")) fprintf(stderr, "libdct_cuda: %s: %d CTAs/SM, %d B smem\n", __FILE__, n, kSmemBytes);
    }
    const unsigned resident = (unsigned)sm_count() * (unsigned)per_sm;
    const unsigned ntiles = (p.nblocks + 31) / 32;
    const unsigned want = (ntiles + kWarps - 1) / kWarps;
    const unsigned grid = want < resident ? want : resident;
    FwdParams q = p;
    const unsigned step = grid * kWarps * 32;                  // blocks between consecutive tiles of a warp
    q.step_q = step / p.bw;
    q.step_r = step - q.step_q * p.bw;
    q.ntiles = ntiles;
    q.tile_stride = grid * kWarps;
    // one worklist / side-array segment per warp of the grid
    const unsigned n_segs = grid * kWarps;
    const unsigned tiles_per_warp = (ntiles + n_segs - 1) / n_segs;
    q.seg_cap = p.wl_cap / n_segs;
    q.side_seg_cap = p.side ? p.side_cap / n_segs : 0;
    if (n_segs > kMaxWorklistSegments - 128 || q.seg_cap < tiles_per_warp * 32 || p.seg_count == nullptr) {
        if (getenv("DCT_CUDA_DEBUG"))
            fprintf(stderr, "libdct_cuda: K1 worklist segments: n_segs %u seg_cap %u tiles/warp %u wl_cap %u seg_count %p\n", n_segs, q.seg_cap,
                    tiles_per_warp, p.wl_cap, (void *)p.seg_count);
        return cudaErrorInvalidValue;
    }
    if (segments) *segments = WorklistSegments{n_segs, q.seg_cap, q.side_seg_cap};
    kernel<<<grid, kThreads, kSmemBytes, s>>>(q);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// CTAs of a bulk-tensor variant that fit one SM (one value per CFG: its variants share launch bounds and shared memory)
template <typename CFG, typename K> static unsigned tma_ctas_per_sm(K kernel)
{
    static int per_sm = 0;
    if (per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, CFG::kThreadsT, CFG::kSmem) != cudaSuccess || n < 1) n = 1;
        per_sm = n < CFG::kMinCtas ? n : CFG::kMinCtas;
        if (getenv("DCT_CUDA_DEBUG"))
            fprintf(stderr, "libdct_cuda: K1 (bulk tensor, %d warps, %d stages): %d CTAs/SM (occupancy %d), %d B smem\n", CFG::kWarpsT,
                    CFG::kStages, per_sm, n, CFG::kSmem);
    }
    return (unsigned)per_sm;
}

static unsigned tma_tiles(const FwdParams &p) { return (p.nblocks / p.bw) * ((p.bw + 31) / 32); }

// one plane's share of a launch of `n_segs` warps whose warp `rot` takes the plane's first tile; `seg_off`: entries of
// every worklist segment that earlier planes of the launch (same plan, same worklist) may fill
static cudaError_t fill_tma_plane(FwdTmaParams &q, const FwdParams &p, unsigned n_segs, unsigned rot, unsigned seg_off = 0)
{
    q.f = p;
    q.nby = p.nblocks / p.bw;
    q.tpr = (p.bw + 31) / 32;
    q.n_segs = n_segs;
    q.rot = rot;
    q.step_ty = n_segs / q.tpr;
    q.step_tx = n_segs - q.step_ty * q.tpr;
    const unsigned ntiles = q.nby * q.tpr;
    const unsigned tiles_per_warp = (ntiles + n_segs - 1) / n_segs;
    q.f.seg_cap = p.wl_cap / n_segs;
    q.f.side_seg_cap = p.side ? p.side_cap / n_segs : 0;
    q.f.side_seg_lim = q.f.side_seg_cap > seg_off ? q.f.side_seg_cap - seg_off : 0;
    q.f.worklist = p.worklist + seg_off;
    if (p.side) q.f.side = p.side + (size_t)seg_off * 64;
    if (n_segs > kMaxWorklistSegments - 128 || q.f.seg_cap < seg_off + tiles_per_warp * 32 || p.seg_count == nullptr) return cudaErrorInvalidValue;
    const int W = (int)p.bw * 8, H = (int)q.nby * 8;
    cudaError_t e;
    if ((e = make_pixel_map(&q.map_px, p.px, p.pitch, W, H)) != cudaSuccess) return e;
    if ((e = make_record_map(&q.map_rec, p.coef, p.nblocks, 32)) != cudaSuccess) return e;
    const int tail = (int)(p.bw % 32);
    return make_record_map(&q.map_rec_tail, p.coef, p.nblocks, tail ? tail : 32);
}

// bulk-tensor kernel: persistent grid like the cp.async one, tiles = 32 blocks of one block row
template <typename CFG, typename K>
static cudaError_t launch_persistent_tma(K kernel, const FwdParams &p, cudaStream_t s, unsigned *launches, WorklistSegments *segments)
{
    constexpr int kW = CFG::kWarpsT;
    cudaError_t e = ensure_smem_attributes(reinterpret_cast<const void *>(kernel), CFG::kSmem);
    if (e != cudaSuccess) return e;
    const unsigned resident = (unsigned)sm_count() * tma_ctas_per_sm<CFG>(kernel);
    const unsigned want = (tma_tiles(p) + kW - 1) / kW;
    const unsigned grid = want < resident ? want : resident;
    FwdTmaParams q;
    if ((e = fill_tma_plane(q, p, grid * kW, 0)) != cudaSuccess) return e;
    if (segments) *segments = WorklistSegments{q.n_segs, q.f.seg_cap, q.f.side_seg_cap};
    e = launch_pdl(kernel, grid, CFG::kThreadsT, CFG::kSmem, s, q);
    if (launches) ++*launches;
    return e;
}

// several planes, one launch (k_fwd_quant_u8_tma_multi): each plane starts at the warp where the previous one ended
template <typename CFG, int NPL, typename K>
static cudaError_t launch_multi_tma(K kernel, const FwdParams *pl, cudaStream_t s)
{
    constexpr int kW = CFG::kWarpsT;
    cudaError_t e = ensure_smem_attributes(reinterpret_cast<const void *>(kernel), CFG::kSmem);
    if (e != cudaSuccess) return e;
    const unsigned resident = (unsigned)sm_count() * tma_ctas_per_sm<CFG>(kernel);
    unsigned total = 0;
    for (int i = 0; i < NPL; ++i) total += tma_tiles(pl[i]);
    const unsigned want = (total + kW - 1) / kW;
    const unsigned grid = want < resident ? want : resident;
    const unsigned n_segs = grid * kW;
    FwdTmaMulti<NPL> m;
    unsigned rot = 0;
    for (int i = 0; i < NPL; ++i) {
        unsigned seg_off = 0;        // planes of one plan share its worklist: each gets its own range of every segment
        for (int j = 0; j < i; ++j)
            if (pl[j].worklist == pl[i].worklist) seg_off += (tma_tiles(pl[j]) + n_segs - 1) / n_segs * 32;
        if ((e = fill_tma_plane(m.pl[i], pl[i], n_segs, rot, seg_off)) != cudaSuccess)
            return NPL == 1 ? e : cudaErrorNotSupported;       // several planes: queued one by one instead
        rot = (rot + tma_tiles(pl[i])) % n_segs;
    }
    return launch_pdl(kernel, grid, CFG::kThreadsT, CFG::kSmem, s, m);
}

// Small planes fold the replay into K1's tail (one launch instead of two); see the kernel's tail for why large ones do not.
static bool fold_small_plane(const FwdParams &p)
{
    static const bool no_fold = getenv("DCT_CUDA_NO_FOLD") != nullptr;   // measurement aid
    return !no_fold && p.tab != nullptr && p.ctr != nullptr && p.nblocks <= kFoldMaxBlocks;
}

template <int LAYOUT, bool ADAPTIVE, typename CFG>
static cudaError_t launch_k1_tma(const FwdParams &p, cudaStream_t s, unsigned *launches, WorklistSegments *segments, bool *folded)
{
    if (fold_small_plane(p)) {
        if (folded) *folded = true;
        if (launches) ++*launches;
        return p.uniform_band ? launch_multi_tma<CFG, 1>(k_fwd_quant_u8_tma_frame<LAYOUT, ADAPTIVE, true, CFG, 1>, &p, s)
                              : launch_multi_tma<CFG, 1>(k_fwd_quant_u8_tma_frame<LAYOUT, ADAPTIVE, false, CFG, 1>, &p, s);
    }
    return p.uniform_band ? launch_persistent_tma<CFG>(k_fwd_quant_u8_tma<LAYOUT, ADAPTIVE, true, CFG>, p, s, launches, segments)
                          : launch_persistent_tma<CFG>(k_fwd_quant_u8_tma<LAYOUT, ADAPTIVE, false, CFG>, p, s, launches, segments);
}

// the bulk-tensor kernel needs: the driver's tensor-map encoder, a 16-byte aligned plane whose pitch is a multiple
// of 16, and block rows of at least one whole tile (narrower planes stay on the cp.async kernel, whose tiles wrap)
static bool tma_eligible(const FwdParams &p)
{
    static const bool disabled = getenv("DCT_CUDA_NO_TMA") != nullptr;   // measurement aid: force the cp.async kernels
    if (disabled || p.no_tma || !tma_available()) return false;
    if (p.bw < 32 || (p.pitch % 16) || ((uintptr_t)p.px % 16) || ((uintptr_t)p.coef % 16)) return false;
    const unsigned long long padded = (unsigned long long)(p.nblocks / p.bw) * ((p.bw + 31) / 32) * 32;
    return padded + (unsigned long long)kMaxWorklistSegments * 64 <= p.wl_cap;
}

template <int LAYOUT, bool ADAPTIVE>
static cudaError_t launch_k1(const FwdParams &p, cudaStream_t s, unsigned *launches, WorklistSegments *segments, bool *folded)
{
    if (tma_eligible(p)) {
        // Geometry (measured on B200, 64 4K frames, profiles/r2_geometry.md): ONE CTA of 12 warps per SM with up to 168
        // registers per thread beats 3 x 8 warps at 80 registers (0.93 against 0.85 of the copy peak): fewer, fatter
        // warps keep more of a block's arithmetic in flight per warp and leave the memory system shallower queues.
        // DCT_CUDA_K1_GEOMETRY=2 (tuning aid): 8 warps x 3 CTAs.
        static const int variant = getenv("DCT_CUDA_K1_GEOMETRY") ? atoi(getenv("DCT_CUDA_K1_GEOMETRY")) : 0;
        if (variant == 2) return launch_k1_tma<LAYOUT, ADAPTIVE, TmaCfg<8, 3, 2>>(p, s, launches, segments, folded);
        return launch_k1_tma<LAYOUT, ADAPTIVE, TmaCfg<12, 1, 2>>(p, s, launches, segments, folded);
    }
    return p.uniform_band ? launch_persistent(k_fwd_quant_u8<LAYOUT, ADAPTIVE, true>, p, s, launches, segments)
                          : launch_persistent(k_fwd_quant_u8<LAYOUT, ADAPTIVE, false>, p, s, launches, segments);
}

cudaError_t launch_fwd_quant_f32(const FwdParams &p, int layout, cudaStream_t s)
{
    if (p.nblocks == 0) return cudaSuccess;
    const unsigned grid = (p.nblocks + kThreads - 1) / kThreads;
    if (layout == LAYOUT_ZIGZAG) k_fwd_quant_f32<LAYOUT_ZIGZAG><<<grid, kThreads, 0, s>>>(p);
    else k_fwd_quant_f32<LAYOUT_NATURAL><<<grid, kThreads, 0, s>>>(p);
    return cudaGetLastError();
}

// 2 or 3 non-adaptive planes in one launch; cudaErrorNotSupported when the planes do not qualify (the caller then
// queues them one by one)
cudaError_t launch_fwd_quant_u8_multi(const FwdParams *pl, int n, int layout, cudaStream_t s)
{
    static const bool off = getenv("DCT_CUDA_NO_MULTI") != nullptr;      // measurement aid
    if (off || n < 2 || n > 3) return cudaErrorNotSupported;
    bool uniform = true;
    unsigned long long blocks = 0;
    for (int i = 0; i < n; ++i) {
        if (pl[i].nblocks == 0 || !tma_eligible(pl[i]) || !fold_small_plane(pl[i])) return cudaErrorNotSupported;
        uniform = uniform && pl[i].uniform_band;
        blocks += pl[i].nblocks;
    }
    if (blocks > 2ull * kFoldMaxBlocks) return cudaErrorNotSupported;
    using CFG = TmaCfg<12, 1, 2>;
#define DCTB_MULTI(L, U, N) launch_multi_tma<CFG, N>(k_fwd_quant_u8_tma_frame<L, false, U, CFG, N>, pl, s)
    if (layout == LAYOUT_ZIGZAG) {
        if (n == 2) return uniform ? DCTB_MULTI(LAYOUT_ZIGZAG, true, 2) : DCTB_MULTI(LAYOUT_ZIGZAG, false, 2);
        return uniform ? DCTB_MULTI(LAYOUT_ZIGZAG, true, 3) : DCTB_MULTI(LAYOUT_ZIGZAG, false, 3);
    }
    if (n == 2) return uniform ? DCTB_MULTI(LAYOUT_NATURAL, true, 2) : DCTB_MULTI(LAYOUT_NATURAL, false, 2);
    return uniform ? DCTB_MULTI(LAYOUT_NATURAL, true, 3) : DCTB_MULTI(LAYOUT_NATURAL, false, 3);
#undef DCTB_MULTI
}

cudaError_t launch_fwd_quant_u8(const FwdParams &p, int layout, int adaptive, cudaStream_t s, unsigned *launches,
                                WorklistSegments *segments, bool *folded)
{
    if (folded) *folded = false;
    if (launches) *launches = 0;
    if (segments) *segments = WorklistSegments{0, 0, 0};
    if (p.nblocks == 0) return cudaSuccess;
    if (layout == LAYOUT_ZIGZAG)
        return adaptive ? launch_k1<LAYOUT_ZIGZAG, true>(p, s, launches, segments, folded)
                        : launch_k1<LAYOUT_ZIGZAG, false>(p, s, launches, segments, folded);
    return adaptive ? launch_k1<LAYOUT_NATURAL, true>(p, s, launches, segments, folded)
                    : launch_k1<LAYOUT_NATURAL, false>(p, s, launches, segments, folded);
}

}  // namespace dctb
