/*
 * host_context.c -- the host-side half of the drop-in API: context set-up and marshalling.
 *
 * These are the parts of include/dct.h and include/quantization.h that the reference runs once
 * per context (or that only move data): dct_init/dct_free (src/dct.c:7-49), quant_init/quant_free
 * (src/quantization.c:19-49), generate_quant_matrix (:51-99), generate_dequant_matrix (:101-111),
 * adjust_matrix_for_block (:171-211), calculate_block_variance (:153-169),
 * create_block_from_pixels (src/dct.c:109-120), copy_block_to_coefficients (:123-129).
 * The fp64 expressions are evaluated in the reference's order with the host libm, because the
 * resulting doubles are data the GPU must receive bit-for-bit.  Compiled as
 * ISO C99 (-std=c99 => no FMA contraction), like the reference.
 *
 * The per-block compute calls (dct_forward, dct_inverse, quantize, dequantize) are NOT here:
 * they launch CUDA kernels from shim.cu.
 */
#include <dct.h>
#include <quantization.h>

/* Ragged arrays laid out exactly like the reference's alloc_array (src/utils.c:8-25) so that
 * callers can release what we return with their own free_array / free_int_array. */
static double **ragged_alloc(int rows, int cols)
{
    double **a = (double **)malloc((size_t)rows * sizeof(double *));
    if (!a) {
        fprintf(stderr, "Memory allocation failed, when creating new 2D array\n");
        exit(EXIT_FAILURE);
    }
    for (int i = 0; i < rows; ++i) {
        a[i] = (double *)calloc((size_t)cols, sizeof(double));
        if (!a[i]) {
            fprintf(stderr, "Memory allocation failed, when creating new 2D array\n");
            exit(EXIT_FAILURE);
        }
    }
    return a;
}

static void ragged_free(double **a, int rows)
{
    if (!a) return;
    for (int i = 0; i < rows; ++i) free(a[i]);
    free(a);
}

DCTContext *dct_init(int block_size)
{
    DCTContext *ctx = (DCTContext *)malloc(sizeof(DCTContext));
    if (!ctx) {
        fprintf(stderr, "Memory allocation failed, when creating new context\n");
        exit(EXIT_FAILURE);
    }
    const int n = block_size;
    ctx->block_size = n;
    ctx->dct_matrix = ragged_alloc(n, n);
    ctx->transposed_dct = ragged_alloc(n, n);
    for (int i = 0; i < n; ++i) {
        /* src/dct.c:21-26: alpha_0 = 1/sqrt(n), alpha_i = sqrt(2/n) */
        const double alpha = (i == 0) ? 1.0 / sqrt(n) : sqrt(2.0 / n);
        for (int j = 0; j < n; ++j) {
            const double v = alpha * cos((PI * (2 * j + 1) * i) / (2.0 * n)); /* src/dct.c:28 */
            ctx->dct_matrix[i][j] = v;
            ctx->transposed_dct[j][i] = v;
        }
    }
    return ctx;
}

void dct_free(DCTContext *ctx)
{
    if (!ctx) return;
    ragged_free(ctx->dct_matrix, ctx->block_size);
    ragged_free(ctx->transposed_dct, ctx->block_size);
    free(ctx);
}

double **create_block_from_pixels(unsigned char *pixels, int width, int row_start, int col_start,
                                  int block_size)
{
    double **block = ragged_alloc(block_size, block_size);
    for (int i = 0; i < block_size; ++i) {
        const unsigned char *row = pixels + (row_start + i) * width + col_start; /* int index, as :114 */
        for (int j = 0; j < block_size; ++j) block[i][j] = (double)row[j] - 128.0;
    }
    return block;
}

void copy_block_to_coefficients(double **block, int **coefficients, int block_size)
{
    for (int i = 0; i < block_size; ++i)
        for (int j = 0; j < block_size; ++j) coefficients[i][j] = (int)round(block[i][j]);
}

/* ITU-T T.81 Annex K, Table K.1 (luminance) -- the base table of src/quantization.c:8-17 */
static const unsigned char annex_k_luma[8][8] = {
    {16, 11, 10, 16, 24, 40, 51, 61},     {12, 12, 14, 19, 26, 58, 60, 55},
    {14, 13, 16, 24, 40, 57, 69, 56},     {14, 17, 22, 29, 51, 87, 80, 62},
    {18, 22, 37, 56, 68, 109, 103, 77},   {24, 35, 55, 64, 81, 104, 113, 92},
    {49, 64, 78, 87, 103, 121, 120, 101}, {72, 92, 95, 98, 112, 100, 103, 99}};

double **generate_quant_matrix(int block_size, int quality)
{
    const int n = block_size;
    double **m = ragged_alloc(n, n);
    double scale = (quality < 50) ? 5000.0 / quality : 200.0 - 2 * quality;
    scale /= 100.0;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            double v;
            if (n == 8) {
                v = annex_k_luma[i][j] * scale;
            } else {
                /* src/quantization.c:80-84: grows with the distance from DC */
                const double distance = sqrt((double)(i * i + j * j));
                v = (1.0 + distance) * scale * 8.0;
            }
            m[i][j] = v < 1.0 ? 1.0 : (v > 255.0 ? 255.0 : v);
        }
    }
    return m;
}

double **generate_dequant_matrix(double **quant_matrix, int block_size)
{
    double **r = ragged_alloc(block_size, block_size);
    for (int i = 0; i < block_size; ++i)
        for (int j = 0; j < block_size; ++j) r[i][j] = 1.0 / quant_matrix[i][j];
    return r;
}

QuantContext *quant_init(int block_size, int quality, int adaptive)
{
    QuantContext *ctx = (QuantContext *)malloc(sizeof(QuantContext));
    if (!ctx) {
        fprintf(stderr, "Memory allocation failed when creating quantization context\n");
        exit(EXIT_FAILURE);
    }
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    ctx->block_size = block_size;
    ctx->quality = quality;
    ctx->adaptive = adaptive;
    ctx->quant_matrix = generate_quant_matrix(block_size, quality);
    ctx->dequant_matrix = generate_dequant_matrix(ctx->quant_matrix, block_size);
    return ctx;
}

void quant_free(QuantContext *ctx)
{
    if (!ctx) return;
    ragged_free(ctx->quant_matrix, ctx->block_size);
    ragged_free(ctx->dequant_matrix, ctx->block_size);
    free(ctx);
}

double calculate_block_variance(double **block, int block_size)
{
    double sum = 0.0, sum_sq = 0.0;
    const int count = block_size * block_size;
    for (int i = 0; i < block_size; ++i) {
        for (int j = 0; j < block_size; ++j) {
            sum += block[i][j];
            sum_sq += block[i][j] * block[i][j];
        }
    }
    const double mean = sum / count;
    return (sum_sq / count) - (mean * mean);
}

double **adjust_matrix_for_block(QuantContext *ctx, double variance, int is_quantize)
{
    const int n = ctx->block_size;
    double **m = ragged_alloc(n, n);
    double **source = is_quantize ? ctx->quant_matrix : ctx->dequant_matrix;
    const double nv = fmin(1.0, fmax(0.1, variance / 1000.0));
    const double scale = is_quantize ? 2.0 - nv : 1.0 / (2.0 - nv);
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            if (i == 0 && j == 0) {
                m[i][j] = source[i][j]; /* DC keeps the table entry */
            } else {
                m[i][j] = source[i][j] * scale;
                if (is_quantize && m[i][j] < 1.0) m[i][j] = 1.0;
            }
        }
    }
    return m;
}
