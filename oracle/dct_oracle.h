/*
 * dct_oracle.h -- CPU oracle for the 8x8 DCT + quantization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dct_b200/ (the product) may link,
 * import or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * It is a plain-C restatement (flat arrays instead of ragged ones, same
 * arithmetic in the same order) of the reference's
 *   src/dct.c, src/quantization.c and the zigzag scan of src/entropy.c.
 * Parity is PINNED: tests/test_oracle.py checks every function against the
 * golden vectors captured from the reference's own tests (tests/golden/) and,
 * when oracle/_ref/libdct_ref.so exists (built from the sources under
 * /root/reference by oracle/Makefile), bit-for-bit against the reference
 * itself on random planes.
 *
 * Build: gcc -std=c99 -O2 -ffp-contract=off (ISO mode: no FMA contraction,
 * which is what makes the fp64 results optimisation-independent).
 */
#ifndef DCT_ORACLE_H
#define DCT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#define ORC_LAYOUT_NATURAL 0 /* k = 8*i + j                                 */
#define ORC_LAYOUT_ZIGZAG  1 /* k = position in block_to_zigzag()'s output  */

/* ---- block level (n x n, row-major flat arrays) ------------------------ */
void   orc_dct_matrix(int n, double *D);
void   orc_dct_forward(int n, const double *D, const double *in, double *out);
void   orc_dct_inverse(int n, const double *D, const double *in, double *out);
void   orc_quant_table(int n, int quality, double *Q);
void   orc_dequant_table(int n, const double *Q, double *R);
double orc_block_variance(int n, const double *blk);
void   orc_adjust_table(int n, const double *src, double variance, int is_quantize, double *out);
void   orc_quantize(int n, const double *Q, int adaptive, const double *c, int *q, double variance);
void   orc_dequantize(int n, const double *Q, const double *R, int adaptive, const int *q, double *c,
                      double variance);
void   orc_zigzag_order(int n, int *order); /* order[k] = natural index visited k-th */
void   orc_round_to_int(int n, const double *blk, int *out);

/* ---- plane level (8x8 blocks, W and H multiples of 8) ------------------- */
/* coefficient record: block-major, coef[(by*(W/8)+bx)*64 + k], int16.       */
int orc_fwd_quant_plane(const uint8_t *px, size_t pitch, int W, int H, const double *Q, int adaptive,
                        int layout, int16_t *coef, double *var_out, int nthreads, uint64_t *near_ties);
int orc_fwd_quant_plane_f32(const float *px, size_t pitch_floats, int W, int H, const double *Q, int adaptive,
                            int layout, int16_t *coef, double *var_out, int nthreads, uint64_t *near_ties);
int orc_dequant_idct_plane(const int16_t *coef, int W, int H, const double *Q, const double *R,
                           int adaptive, int layout, const double *var_in, uint8_t *px, size_t pitch,
                           int nthreads, uint64_t *near_ties);

/* ---- the same for any block size n <= 32 (tables n*n, records n*n int16, W and H multiples of n) ---- */
int orc_fwd_quant_plane_n(int n, const uint8_t *px, size_t pitch, int W, int H, const double *Q, int adaptive,
                          int layout, int16_t *coef, double *var_out, int nthreads, uint64_t *near_ties);
int orc_dequant_idct_plane_n(int n, const int16_t *coef, int W, int H, const double *Q, const double *R,
                             int adaptive, int layout, const double *var_in, uint8_t *px, size_t pitch,
                             int nthreads, uint64_t *near_ties);

/* ---- run-length symbols (value, run) per record; returns the total, offsets has nblocks+1 entries ---- */
size_t orc_rle_plane(const int16_t *coef, size_t nblocks, int layout, uint32_t *offsets, int32_t *symbols);

/* ---- planar front / back end: OUR convention, no reference counterpart => parity unpinned (see .c) ---- */
void orc_rgb_to_ycbcr420(const uint8_t *rgb, size_t rgb_pitch, int W, int H, uint8_t *y, size_t y_pitch, int y_w,
                         int y_h, uint8_t *cb, uint8_t *cr, size_t c_pitch, int c_w, int c_h);
void orc_ycbcr420_to_rgb(const uint8_t *y, size_t y_pitch, const uint8_t *cb, const uint8_t *cr, size_t c_pitch, int W,
                         int H, uint8_t *rgb, size_t rgb_pitch);
void orc_pad_edges(uint8_t *px, size_t pitch, int W, int H, int Wp, int Hp, int elem);

/* ---- helpers shared by the tests --------------------------------------- */
void     orc_fill_xorshift(uint8_t *dst, size_t n, uint64_t seed, int dist, int W);
uint64_t orc_fnv_i16(const int16_t *v, size_t n);
uint64_t orc_fnv_u8(const uint8_t *v, size_t n);
uint64_t orc_fnv_u8_blockorder(const uint8_t *px, size_t pitch, int W, int H);

#endif
