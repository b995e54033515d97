/*
 * quantization.h -- drop-in replacement for the reference's include/quantization.h, served by
 * libdct_cuda.
 *
 * Same type and the same eight entry points as erkinov-wtf/dct include/quantization.h:18-24
 * (QuantContext) and :34,41,50,59,69,79,88,98.  Semantics are the reference's, warts included:
 *
 *   - tables are UN-ROUNDED doubles, clamp(lumaTable * scale, 1, 255), scale = (q < 50 ?
 *     5000/q : 200 - 2q) / 100, quality clamped to [1, 100]        (src/quantization.c:26-31, :51-77)
 *   - quantize:   (int) round(c / M), true division, half away from zero            (:124)
 *   - dequantize, adaptive == 0:  q * dequant_matrix = q * (1/Q)    -- sic           (:139, :144)
 *     dequantize, adaptive == 1:  q * (1.0 / ((1/Q) * (1/(2 - nv)))), DC unscaled    (:137, :144, :171-211)
 *
 *   quant_init / quant_free / generate_*_matrix / adjust_matrix_for_block /
 *   calculate_block_variance     host (table set-up and scalar helpers, same fp64 expressions)
 *   quantize / dequantize         one CUDA launch per call (replay_f64.cu), bit-identical ints /
 *                                 doubles.  The throughput path is the plane API in dct_cuda.h.
 *
 * Callers may overwrite ctx->quant_matrix entries after quant_init (that is how a chroma table
 * gets in); the per-block calls read the tables at every call, a dct_cuda_plan re-reads them on
 * dct_cuda_plan_refresh().
 */
#ifndef QUANTIZATION_H
#define QUANTIZATION_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <utils.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int block_size;          /* n */
    int quality;             /* 1..100 after clamping */
    double **quant_matrix;   /* Q, ragged n x n */
    double **dequant_matrix; /* 1.0 / Q */
    int adaptive;            /* 0 = fixed table, 1 = per-block variance scaling */
} QuantContext;

QuantContext *quant_init(int block_size, int quality, int adaptive);
void quant_free(QuantContext *ctx); /* NULL is accepted */

/* both return fresh ragged arrays, released with free_array */
double **generate_quant_matrix(int block_size, int quality);
double **generate_dequant_matrix(double **quant_matrix, int block_size);

void quantize(QuantContext *ctx, double **dct_coeffs, int **quant_coeffs, double block_variance);
void dequantize(QuantContext *ctx, int **quant_coeffs, double **dct_coeffs, double block_variance);

/* sum_sq/n^2 - mean^2 over the block handed in */
double calculate_block_variance(double **block, int block_size);
/* fresh ragged table for this block: source * (2 - nv) (quantize, floor 1.0) or source / (2 - nv) */
double **adjust_matrix_for_block(QuantContext *ctx, double variance, int is_quantize);

#ifdef __cplusplus
}
#endif

#endif /* QUANTIZATION_H */
