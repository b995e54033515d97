/*
 * dct_cuda.h -- plane / batch entry points of libdct_cuda (C ABI, plain pointers and sizes).
 *
 * The reference (erkinov-wtf/dct) has no image-level code: its only composed caller,
 * tests/test_entropy.c:278-405, walks ONE 8x8 block through
 *     create_block_from_pixels  src/dct.c:109      \
 *     dct_forward               src/dct.c:52        > dct_cuda_fwd_quant_u8*      (kernel K1 + K3)
 *     quantize                  src/quantization.c:113 /
 *     [block_to_zigzag          src/entropy.c:158]    layout = DCT_CUDA_ZIGZAG
 *     dequantize                src/quantization.c:133 \
 *     dct_inverse               src/dct.c:80            > dct_cuda_dequant_idct_u8* (kernel K2 + K3)
 *     +128, round, clamp to u8  tests/test_entropy.c:380-384 /
 * These entry points are that loop over every block of a plane, bit-identical in their integer
 * results to calling the reference's functions block by block (exact .5 ties included: blocks
 * whose fp32 result falls inside the proven error band are re-done in the reference's own
 * fp64 operation order; their number is reported in dct_cuda_stats).
 *
 * Data formats
 *   pixels        uint8, row-major, `pitch` bytes between rows (multiple of 8), W and H
 *                 multiples of 8.  A batch of equally sized frames stored back to back is one
 *                 plane of height n_frames*H.
 *   coefficients  int16 records, block-major: coef[(by*(W/8) + bx)*64 + k], 128 bytes per block.
 *                 DCT_CUDA_NATURAL: k = 8*i + j.  DCT_CUDA_ZIGZAG: k = position in
 *                 block_to_zigzag()'s output.  int16 is lossless: |c| <= 1024 and Q >= 1.
 *                 dct_cuda_record_to_block() widens one record into the ragged int** that the
 *                 untouched run_length_encode() (src/entropy.c:216) expects -- NATURAL order,
 *                 because run_length_encode applies the zigzag itself.
 *   variance      adaptive contexts only: one double per block (the side information the
 *                 reference passes as `block_variance`), written by the forward call and read
 *                 by the inverse call.
 *
 * Block sizes other than 8 (any block_size <= 32, both contexts alike) are accepted by the same calls:
 * records then hold n*n coefficients, W and H must be multiples of n, and the arithmetic is the
 * reference's fp64 loops (bit-identical, not a throughput path).  Float tiles, run-length symbols and
 * the multi-GPU helpers are 8x8 only.
 *
 * A dct_cuda_plan binds a (DCTContext, QuantContext) pair to one GPU: it uploads the host-made
 * fp64 tables, derives the fp32 multipliers and error bands, and owns the replay worklist.
 * Plans are cheap; use one per host thread / stream.  All functions return 0 or a negative
 * DCT_CUDA_E* code; dct_cuda_last_error() describes the last failure on the calling thread.
 * There is no CPU fallback anywhere: without a CUDA device every call fails.
 */
#ifndef DCT_CUDA_H
#define DCT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#include <dct.h>
#include <quantization.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCT_CUDA_NATURAL 0
#define DCT_CUDA_ZIGZAG 1

#define DCT_CUDA_OK 0
#define DCT_CUDA_EINVAL (-1)  /* bad shape, alignment, layout or NULL argument */
#define DCT_CUDA_ECUDA (-2)   /* a CUDA call failed; see dct_cuda_last_error() */
#define DCT_CUDA_ENOMEM (-3)
#define DCT_CUDA_ENODEV (-4)  /* no usable CUDA device */

typedef struct dct_cuda_plan dct_cuda_plan;

typedef struct {
    uint64_t blocks;          /* 8x8 blocks processed since the last fetch                          */
    uint64_t replayed_blocks; /* of those, re-done in fp64 because a value sat inside the band       */
    uint64_t near_ties;       /* fp64 values within 1e-9 of a .5 rounding boundary (exact ties)      */
    uint64_t saturated;       /* quantised values outside int16 (only with tables below 1.0)         */
} dct_cuda_stats;

/* one plane of a multi-plane frame (e.g. Y, Cb, Cr of a 4:2:0 frame, each with its own tables) */
typedef struct {
    dct_cuda_plan *plan;
    const void *pixels_in; /* forward: uint8 source;   inverse: unused  */
    void *pixels_out;      /* inverse: uint8 destination                 */
    size_t pitch;
    int width, height;
    void *coef;            /* int16 records (forward: destination, inverse: source) */
    double *variance;      /* adaptive plans only, else NULL */
} dct_cuda_plane;

const char *dct_cuda_last_error(void);
int dct_cuda_device_count(void);

/* Binds the contexts to GPU `device`.  The contexts must outlive the plan. */
dct_cuda_plan *dct_cuda_plan_create(const DCTContext *dct, const QuantContext *quant, int device);
/* Re-reads the tables after the caller edited quant->quant_matrix / dequant_matrix / adaptive. */
int dct_cuda_plan_refresh(dct_cuda_plan *plan);
void dct_cuda_plan_destroy(dct_cuda_plan *plan);
int dct_cuda_plan_device(const dct_cuda_plan *plan);
/* Kernels launched for this plan so far (fused kernels, replays, record conversions). */
uint64_t dct_cuda_plan_kernel_launches(const dct_cuda_plan *plan);

/* ---- device-resident planes: all data pointers are device pointers on the plan's GPU; the work
 * is queued on `stream` (a cudaStream_t, NULL = default stream) and the call returns at once. ---- */
int dct_cuda_fwd_quant_u8_dev(dct_cuda_plan *plan, const uint8_t *d_pixels, size_t pitch, int width,
                              int height, int16_t *d_coef, int layout, double *d_variance, void *stream);
int dct_cuda_dequant_idct_u8_dev(dct_cuda_plan *plan, const int16_t *d_coef, int width, int height,
                                 int layout, const double *d_variance, uint8_t *d_pixels, size_t pitch,
                                 void *stream);
/* The planes of one frame (e.g. Y, Cb, Cr with a luma and a chroma plan).  2 or 3 planes of non-adaptive 8x8 plans
 * on one device, each of at most 600 000 blocks, 16-byte aligned with pitches that are multiples of 16 and at least
 * 256 pixels wide, go through ONE kernel launch per call; anything else is queued plane by plane.  Same results. */
int dct_cuda_fwd_quant_planes_dev(const dct_cuda_plane *planes, int n_planes, int layout, void *stream);
int dct_cuda_dequant_idct_planes_dev(const dct_cuda_plane *planes, int n_planes, int layout, void *stream);

/* ---- host planes: pointers are host memory (pinned memory overlaps best).  The plane is cut into
 * block-row strips that flow through a 3-deep H2D / kernel / D2H pipeline on the plan's own
 * streams; the call returns when the result is in host memory.  `stats` may be NULL. ---- */
int dct_cuda_fwd_quant_u8(dct_cuda_plan *plan, const uint8_t *pixels, size_t pitch, int width, int height,
                          int16_t *coef, int layout, double *variance, dct_cuda_stats *stats);
int dct_cuda_dequant_idct_u8(dct_cuda_plan *plan, const int16_t *coef, int width, int height, int layout,
                             const double *variance, uint8_t *pixels, size_t pitch, dct_cuda_stats *stats);

/* ---- float pixel tiles (forward only, non-adaptive plans): the block is (double)p - 128.0 for any
 * float p, i.e. what a caller of the reference gets by filling dct_forward's input block by hand
 * (tests/test_dct.c:46-50).  Bit-exact like the uint8 path; pixels outside [0, 255] are legal (their
 * blocks are computed in fp64).  `pitch_bytes` must be a multiple of 16 for the device call. ---- */
int dct_cuda_fwd_quant_f32_dev(dct_cuda_plan *plan, const float *d_pixels, size_t pitch_bytes, int width, int height,
                               int16_t *d_coef, int layout, void *stream);
int dct_cuda_fwd_quant_f32(dct_cuda_plan *plan, const float *pixels, size_t pitch_bytes, int width, int height,
                           int16_t *coef, int layout, dct_cuda_stats *stats);

/* Asynchronous forms: queue the whole strip pipeline on the plan's own streams and return at once.
 * The host buffers must be PINNED and stay untouched until dct_cuda_plan_wait(plan, stats) returns.
 * Two plans (e.g. an encoder and a decoder) driven this way overlap each other's H2D and D2H
 * traffic, which the synchronous calls cannot (forward is D2H-heavy, inverse is H2D-heavy). */
int dct_cuda_fwd_quant_u8_async(dct_cuda_plan *plan, const uint8_t *pixels, size_t pitch, int width, int height,
                                int16_t *coef, int layout, double *variance);
int dct_cuda_dequant_idct_u8_async(dct_cuda_plan *plan, const int16_t *coef, int width, int height, int layout,
                                   const double *variance, uint8_t *pixels, size_t pitch);
int dct_cuda_plan_wait(dct_cuda_plan *plan, dct_cuda_stats *stats);

/* ---- several GPUs, one host plane: block-row ranges are dealt to the plans (one per GPU, same
 * tables) and run concurrently, one host thread per GPU; no inter-GPU traffic. ---- */
int dct_cuda_fwd_quant_u8_multi(dct_cuda_plan *const *plans, int n_plans, const uint8_t *pixels, size_t pitch,
                                int width, int height, int16_t *coef, int layout, double *variance,
                                dct_cuda_stats *stats);
int dct_cuda_dequant_idct_u8_multi(dct_cuda_plan *const *plans, int n_plans, const int16_t *coef, int width,
                                   int height, int layout, const double *variance, uint8_t *pixels,
                                   size_t pitch, dct_cuda_stats *stats);

/* ---- int8 records for the host-plane calls ----
 * The host-plane calls are bound by the PCIe link, and two of their three bytes per pixel are records.
 * When every entry of the plan's table is >= block_size * 128 / 127.5 (8.03 for 8x8: quality <= ~56 with the
 * reference's luminance table) no quantised value of an 8-bit plane can leave [-127, 127], so the same
 * record fits 64 BYTES: these calls are dct_cuda_fwd_quant_u8 / dct_cuda_dequant_idct_u8 with int8 records
 * on the host side (same order, same values) and half the record traffic.  The forward calls fail with
 * DCT_CUDA_EINVAL when dct_cuda_plan_records_fit_i8(plan) is 0; decoding int8 records works with any table.
 * On the device the records stay int16. */
int dct_cuda_plan_records_fit_i8(const dct_cuda_plan *plan);
int dct_cuda_fwd_quant_u8_i8(dct_cuda_plan *plan, const uint8_t *pixels, size_t pitch, int width, int height,
                             int8_t *coef8, int layout, double *variance, dct_cuda_stats *stats);
int dct_cuda_dequant_idct_i8_u8(dct_cuda_plan *plan, const int8_t *coef8, int width, int height, int layout,
                                const double *variance, uint8_t *pixels, size_t pitch, dct_cuda_stats *stats);
int dct_cuda_fwd_quant_u8_i8_async(dct_cuda_plan *plan, const uint8_t *pixels, size_t pitch, int width, int height,
                                   int8_t *coef8, int layout, double *variance);
int dct_cuda_dequant_idct_i8_u8_async(dct_cuda_plan *plan, const int8_t *coef8, int width, int height, int layout,
                                      const double *variance, uint8_t *pixels, size_t pitch);
/* widen one 64-byte record into a ragged 8x8 int block in NATURAL order (cf. dct_cuda_record_to_block) */
void dct_cuda_record8_to_block(const int8_t *record, int layout, int **block);

/* ---- several GPUs, data resident on ONE of them (NVLink / NVSwitch peers) ----
 * d_pixels / d_coef / d_variance live on plans[0]'s GPU (the owner); plans[1..] are plans with the same
 * tables on other GPUs that can map the owner's memory.  Block-row ranges are dealt out -- peer g gets
 * share[g] of the rows (share[0] is ignored: the owner keeps the rest; NULL = dct_cuda_peer_default_share
 * each: a split for plans on the fp64 exact path, which are arithmetic-bound, and NOTHING for the
 * fused fp32 path, which local HBM serves faster than an NVLink port can -- see DESIGN.md for the
 * measurement) -- and every GPU runs the ordinary kernels on its range: a peer's loads and stores go straight
 * to the owner's memory over NVLink, inside the kernel, with no staging copy and no collective.  The
 * whole call is ordered on `stream` (an owner-GPU stream): it starts after the work queued there before
 * and work queued there afterwards sees every shard.  Per-plan counters: dct_cuda_stats_fetch on each. */
float dct_cuda_peer_default_share(const dct_cuda_plan *plan, int n_plans, int forward);
int dct_cuda_fwd_quant_u8_peer(dct_cuda_plan *const *plans, int n_plans, const uint8_t *d_pixels, size_t pitch,
                               int width, int height, int16_t *d_coef, int layout, double *d_variance,
                               const float *share, void *stream);
int dct_cuda_dequant_idct_u8_peer(dct_cuda_plan *const *plans, int n_plans, const int16_t *d_coef, int width,
                                  int height, int layout, const double *d_variance, uint8_t *d_pixels, size_t pitch,
                                  const float *share, void *stream);

/* ---- planar front / back end (NOT in the reference, which starts from whole-block grayscale planes;
 * the conventions are ours and are stated in dct_b200/csrc/planar.cu and DESIGN.md) ----
 *   edges    a plane is completed to whole blocks by replicating its last column / row;
 *   colour   JFIF full-range BT.601 in 16-bit fixed point, interleaved R,G,B bytes <-> Y, Cb, Cr planes;
 *   4:2:0    one chroma sample per 2x2 pixels (from the sum of the four pixels), replicated on decode. */

/* Host planes of ANY positive size: like dct_cuda_fwd_quant_u8 / dct_cuda_dequant_idct_u8, but the
 * records cover ceil(width/n) x ceil(height/n) blocks; the inverse writes width x height pixels. */
int dct_cuda_fwd_quant_u8_edge(dct_cuda_plan *plan, const uint8_t *pixels, size_t pitch, int width, int height,
                               int16_t *coef, int layout, double *variance, dct_cuda_stats *stats);
int dct_cuda_dequant_idct_u8_edge(dct_cuda_plan *plan, const int16_t *coef, int width, int height, int layout,
                                  const double *variance, uint8_t *pixels, size_t pitch, dct_cuda_stats *stats);
/* Device plane: fill columns [width, width_padded) and rows [height, height_padded) in place. */
int dct_cuda_pad_edges_dev(int device, uint8_t *d_pixels, size_t pitch, int width, int height, int width_padded,
                           int height_padded, void *stream);

/* Plane sizes of a 4:2:0 frame whose planes are completed to whole 8x8 blocks. */
typedef struct {
    int width, height;       /* the RGB image                                              */
    int y_width, y_height;   /* luma plane: width, height rounded up to 8                   */
    int c_width, c_height;   /* each chroma plane: ceil(width/2), ceil(height/2) rounded up to 8 */
} dct_cuda_frame420;
void dct_cuda_frame420_geometry(int width, int height, dct_cuda_frame420 *geometry);

/* Device frames.  The planes are written at their padded sizes (edges replicated), ready for
 * dct_cuda_fwd_quant_planes_dev; the way back writes the width x height image only.  16-byte aligned
 * pointers and pitches take the vector path (rgb, y: 16; cb, cr: 8), anything else a slow scalar one. */
int dct_cuda_rgb_to_ycbcr420_dev(int device, const uint8_t *d_rgb, size_t rgb_pitch, const dct_cuda_frame420 *geometry,
                                 uint8_t *d_y, size_t y_pitch, uint8_t *d_cb, uint8_t *d_cr, size_t c_pitch,
                                 void *stream);
int dct_cuda_ycbcr420_to_rgb_dev(int device, const uint8_t *d_y, size_t y_pitch, const uint8_t *d_cb,
                                 const uint8_t *d_cr, size_t c_pitch, const dct_cuda_frame420 *geometry, uint8_t *d_rgb,
                                 size_t rgb_pitch, void *stream);

/* Whole RGB frames in host memory: copy, convert, transform all three planes, copy back (and the
 * reverse).  `luma` and `chroma` are two non-adaptive 8x8 plans on the same GPU (they may be the same
 * plan); coef_y holds y_width*y_height int16, coef_cb / coef_cr hold c_width*c_height each. */
int dct_cuda_encode_rgb420(dct_cuda_plan *luma, dct_cuda_plan *chroma, const uint8_t *rgb, size_t rgb_pitch, int width,
                           int height, int16_t *coef_y, int16_t *coef_cb, int16_t *coef_cr, int layout,
                           dct_cuda_stats *stats);
int dct_cuda_decode_rgb420(dct_cuda_plan *luma, dct_cuda_plan *chroma, const int16_t *coef_y, const int16_t *coef_cb,
                           const int16_t *coef_cr, int width, int height, int layout, uint8_t *rgb, size_t rgb_pitch,
                           dct_cuda_stats *stats);

/* Waits for the plan's queued work, returns the counters accumulated since the last fetch and
 * clears them.  `stream` is the stream the *_dev calls were queued on. */
int dct_cuda_stats_fetch(dct_cuda_plan *plan, dct_cuda_stats *stats, void *stream);

/* Optional per-kernel timing for benchmarks: while enabled, every K1 / K2 launch is bracketed by
 * CUDA events on its stream; dct_cuda_profile_fetch() synchronises, returns the summed device
 * time and the number of launches per direction, and clears the lists. */
int dct_cuda_plan_profile(dct_cuda_plan *plan, int enable);
int dct_cuda_profile_fetch(dct_cuda_plan *plan, double *fwd_ms, int *fwd_launches, double *inv_ms,
                           int *inv_launches);

/* ---- run-length symbols for the host entropy coder (device-resident records) ----
 * For every record, the list of symbols the untouched run_length_encode() (src/entropy.c:216-256)
 * would produce for that block: zigzag order, one {value, run_length} per non-zero coefficient plus
 * the closing symbol at position 63.  dct_cuda_rle_symbol has the layout of RLESymbol
 * (include/entropy.h:35-38), so a block's list can be memcpy'd into EntropyContext.symbols
 * (set .count) before build_huffman_codes().
 *   pass 1  fills d_offsets[0..nblocks] (exclusive prefix sums of the per-record symbol counts) and
 *           returns their total (synchronises `stream`); the caller sizes d_symbols from it;
 *   pass 2  writes the symbols of record b to d_symbols[d_offsets[b] .. d_offsets[b+1]).
 * `layout` is how the records are stored; the symbols are in zigzag order either way. */
typedef struct {
    int value;
    int run_length;
} dct_cuda_rle_symbol;
int dct_cuda_rle_count_dev(dct_cuda_plan *plan, const int16_t *d_coef, size_t n_blocks, uint32_t *d_offsets,
                           uint64_t *total_symbols, void *stream);
int dct_cuda_rle_emit_dev(dct_cuda_plan *plan, const int16_t *d_coef, size_t n_blocks, int layout,
                          const uint32_t *d_offsets, dct_cuda_rle_symbol *d_symbols, void *stream);

/* Test hook: with skip != 0 the fp64 replay (K3) is not queued, so the output holds the fused kernels'
 * own fp32-path values.  Used by the tests to prove that every value outside a replayed block is
 * already bit-exact, i.e. that the replay hides nothing.  Never set this in production. */
int dct_cuda_plan_debug_skip_replay(dct_cuda_plan *plan, int skip);

/* ---- adapters for the untouched host consumer (src/entropy.c) ---- */
/* widen one 128-byte record into a ragged 8x8 int block in NATURAL order */
void dct_cuda_record_to_block(const int16_t *record, int layout, int **block);
/* narrow a ragged 8x8 int block (NATURAL order) into a record of the given layout */
void dct_cuda_block_to_record(int **block, int layout, int16_t *record);

/* pinned host memory for the host-plane calls */
void *dct_cuda_host_alloc(size_t bytes);
void dct_cuda_host_free(void *p);

#ifdef __cplusplus
}
#endif

#endif /* DCT_CUDA_H */
