/*
 * ref_harness.c -- drives the UNMODIFIED reference (erkinov-wtf/dct) block by block.
 *
 * TEST INFRASTRUCTURE.  This file contains no algorithm of its own: every number comes
 * from the reference's dct_init / create_block_from_pixels / dct_forward / quantize /
 * dequantize / dct_inverse / block_to_zigzag / zigzag_to_block, compiled from the
 * sources where they lie under /root/reference (see oracle/Makefile, target `ref`).
 * The result, oracle/_ref/libdct_ref.so, validates oracle/dct_oracle.c and is the
 * "reference" CPU baseline of bench.py.  It exports flat-array wrappers (ref_*) with the
 * same signatures as the oracle's orc_* functions.
 */
#define _POSIX_C_SOURCE 199309L
#include <dct.h>
#include <entropy.h>
#include <quantization.h>
#include <utils.h>

#include <pthread.h>
#include <stddef.h>
#include <stdint.h>

static void to_ragged(int n, const double *flat, double **r)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) r[i][j] = flat[i * n + j];
}
static void from_ragged(int n, double **r, double *flat)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) flat[i * n + j] = r[i][j];
}

void ref_dct_matrix(int n, double *D)
{
    DCTContext *c = dct_init(n);
    from_ragged(n, c->dct_matrix, D);
    dct_free(c);
}

void ref_dct_forward(int n, const double *D, const double *in, double *out)
{
    (void)D;
    DCTContext *c = dct_init(n);
    double **a = alloc_array(n, n), **b = alloc_array(n, n);
    to_ragged(n, in, a);
    dct_forward(c, a, b);
    from_ragged(n, b, out);
    free_array(a, n);
    free_array(b, n);
    dct_free(c);
}

void ref_dct_inverse(int n, const double *D, const double *in, double *out)
{
    (void)D;
    DCTContext *c = dct_init(n);
    double **a = alloc_array(n, n), **b = alloc_array(n, n);
    to_ragged(n, in, a);
    dct_inverse(c, a, b);
    from_ragged(n, b, out);
    free_array(a, n);
    free_array(b, n);
    dct_free(c);
}

void ref_quant_table(int n, int quality, double *Q)
{
    QuantContext *c = quant_init(n, quality, 0);
    from_ragged(n, c->quant_matrix, Q);
    quant_free(c);
}

void ref_dequant_table(int n, const double *Q, double *R)
{
    double **q = alloc_array(n, n);
    to_ragged(n, Q, q);
    double **r = generate_dequant_matrix(q, n);
    from_ragged(n, r, R);
    free_array(q, n);
    free_array(r, n);
}

double ref_block_variance(int n, const double *blk)
{
    double **a = alloc_array(n, n);
    to_ragged(n, blk, a);
    double v = calculate_block_variance(a, n);
    free_array(a, n);
    return v;
}

/* a context whose tables are the caller's (how a chroma table gets in) */
static QuantContext *ctx_with_table(int n, const double *Q, int adaptive)
{
    QuantContext *c = quant_init(n, 50, adaptive);
    to_ragged(n, Q, c->quant_matrix);
    free_array(c->dequant_matrix, n);
    c->dequant_matrix = generate_dequant_matrix(c->quant_matrix, n);
    return c;
}

void ref_adjust_table(int n, const double *src, double variance, int is_quantize, double *out)
{
    QuantContext *c = quant_init(n, 50, 1);
    to_ragged(n, src, is_quantize ? c->quant_matrix : c->dequant_matrix);
    double **m = adjust_matrix_for_block(c, variance, is_quantize);
    from_ragged(n, m, out);
    free_array(m, n);
    quant_free(c);
}

void ref_quantize(int n, const double *Q, int adaptive, const double *cf, int *q, double variance)
{
    QuantContext *c = ctx_with_table(n, Q, adaptive);
    double **a = alloc_array(n, n);
    int **b = alloc_int_array(n, n);
    to_ragged(n, cf, a);
    quantize(c, a, b, variance);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) q[i * n + j] = b[i][j];
    free_array(a, n);
    free_int_array(b, n);
    quant_free(c);
}

void ref_dequantize(int n, const double *Q, const double *R, int adaptive, const int *q, double *cf,
                    double variance)
{
    (void)R;
    QuantContext *c = ctx_with_table(n, Q, adaptive);
    int **a = alloc_int_array(n, n);
    double **b = alloc_array(n, n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) a[i][j] = q[i * n + j];
    dequantize(c, a, b, variance);
    from_ragged(n, b, cf);
    free_int_array(a, n);
    free_array(b, n);
    quant_free(c);
}

void ref_zigzag_order(int n, int *order)
{
    int **blk = alloc_int_array(n, n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) blk[i][j] = i * n + j;
    block_to_zigzag(blk, order, n);
    free_int_array(blk, n);
}

void ref_round_to_int(int n, const double *blk, int *out)
{
    double **a = alloc_array(n, n);
    int **b = alloc_int_array(n, n);
    to_ragged(n, blk, a);
    copy_block_to_coefficients(a, b, n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) out[i * n + j] = b[i][j];
    free_array(a, n);
    free_int_array(b, n);
}

/* ---- plane loops over the reference's block functions -------------------- */

typedef struct {
    int dir;
    const uint8_t *px_in;
    uint8_t *px_out;
    size_t pitch;
    int W, H;
    int adaptive, layout;
    int16_t *coef_out;
    const int16_t *coef_in;
    double *var_out;
    const double *var_in;
    int row0, row1;
    DCTContext *dctx;
    QuantContext *qctx;
} rjob_t;

static void *ref_worker(void *arg)
{
    rjob_t *jb = (rjob_t *)arg;
    const int bw = jb->W / 8;
    double **c = alloc_array(8, 8), **o = alloc_array(8, 8);
    int **q = alloc_int_array(8, 8);
    int zz[64];
    for (int by = jb->row0; by < jb->row1; ++by) {
        for (int bx = 0; bx < bw; ++bx) {
            size_t b = (size_t)by * bw + bx;
            if (jb->dir == 0) {
                /* strip-relative indexing keeps the reference's `int pixel_index` in range */
                unsigned char *strip = (unsigned char *)jb->px_in + (size_t)by * 8 * jb->pitch;
                double **blk = create_block_from_pixels(strip, (int)jb->pitch, 0, bx * 8, 8);
                dct_forward(jb->dctx, blk, c);
                double var = 0.0;
                if (jb->adaptive) {
                    var = calculate_block_variance(blk, 8);
                    if (jb->var_out) jb->var_out[b] = var;
                }
                quantize(jb->qctx, c, q, var);
                int16_t *dst = jb->coef_out + b * 64;
                if (jb->layout == 1) {
                    block_to_zigzag(q, zz, 8);
                    for (int k = 0; k < 64; ++k) dst[k] = (int16_t)zz[k];
                } else {
                    for (int k = 0; k < 64; ++k) dst[k] = (int16_t)q[k / 8][k % 8];
                }
                free_array(blk, 8);
            } else {
                const int16_t *src = jb->coef_in + b * 64;
                if (jb->layout == 1) {
                    for (int k = 0; k < 64; ++k) zz[k] = src[k];
                    zigzag_to_block(zz, q, 8);
                } else {
                    for (int k = 0; k < 64; ++k) q[k / 8][k % 8] = src[k];
                }
                double var = (jb->adaptive && jb->var_in) ? jb->var_in[b] : 0.0;
                dequantize(jb->qctx, q, c, var);
                dct_inverse(jb->dctx, c, o);
                for (int i = 0; i < 8; ++i)
                    for (int j = 0; j < 8; ++j) {
                        double r = round(o[i][j] + 128.0);
                        if (r < 0.0) r = 0.0;
                        if (r > 255.0) r = 255.0;
                        jb->px_out[((size_t)by * 8 + i) * jb->pitch + (size_t)bx * 8 + j] = (uint8_t)r;
                    }
            }
        }
    }
    free_array(c, 8);
    free_array(o, 8);
    free_int_array(q, 8);
    return NULL;
}

static int ref_run(rjob_t *proto, const double *Q, int nthreads)
{
    if (proto->W <= 0 || proto->H <= 0 || proto->W % 8 || proto->H % 8) return -1;
    proto->dctx = dct_init(8);
    proto->qctx = ctx_with_table(8, Q, proto->adaptive);
    const int bh = proto->H / 8;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > bh) nthreads = bh;
    if (nthreads > 256) nthreads = 256;
    rjob_t jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto;
        jobs[t].row0 = (int)((long long)bh * t / nthreads);
        jobs[t].row1 = (int)((long long)bh * (t + 1) / nthreads);
    }
    if (nthreads == 1) {
        ref_worker(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, ref_worker, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    }
    dct_free(proto->dctx);
    quant_free(proto->qctx);
    return 0;
}

int ref_fwd_quant_plane(const uint8_t *px, size_t pitch, int W, int H, const double *Q, int adaptive,
                        int layout, int16_t *coef, double *var_out, int nthreads, uint64_t *near_ties)
{
    rjob_t jb = {0};
    jb.dir = 0;
    jb.px_in = px;
    jb.pitch = pitch;
    jb.W = W;
    jb.H = H;
    jb.adaptive = adaptive;
    jb.layout = layout;
    jb.coef_out = coef;
    jb.var_out = var_out;
    if (near_ties) *near_ties = 0; /* the reference has no tie accounting */
    return ref_run(&jb, Q, nthreads);
}

int ref_dequant_idct_plane(const int16_t *coef, int W, int H, const double *Q, const double *R,
                           int adaptive, int layout, const double *var_in, uint8_t *px, size_t pitch,
                           int nthreads, uint64_t *near_ties)
{
    (void)R;
    rjob_t jb = {0};
    jb.dir = 1;
    jb.coef_in = coef;
    jb.px_out = px;
    jb.pitch = pitch;
    jb.W = W;
    jb.H = H;
    jb.adaptive = adaptive;
    jb.layout = layout;
    jb.var_in = var_in;
    if (near_ties) *near_ties = 0;
    return ref_run(&jb, Q, nthreads);
}

/* ---- the reference's run_length_encode over every record (src/entropy.c:216-256) ---- */
size_t ref_rle_plane(const int16_t *coef, size_t nblocks, int layout, uint32_t *offsets, int32_t *symbols)
{
    EntropyContext *e = entropy_init(0);
    int **q = alloc_int_array(8, 8);
    int zz[64];
    size_t total = 0;
    for (size_t b = 0; b < nblocks; ++b) {
        const int16_t *rec = coef + b * 64;
        if (layout == 1) {
            for (int k = 0; k < 64; ++k) zz[k] = rec[k];
            zigzag_to_block(zz, q, 8);
        } else {
            for (int k = 0; k < 64; ++k) q[k / 8][k % 8] = rec[k];
        }
        const int n = run_length_encode(e, q, 8);
        offsets[b] = (uint32_t)total;
        for (int i = 0; i < n; ++i) {
            if (symbols) {
                symbols[2 * (total + i)] = e->symbols[i].value;
                symbols[2 * (total + i) + 1] = e->symbols[i].run_length;
            }
        }
        total += (size_t)n;
    }
    offsets[nblocks] = (uint32_t)total;
    free_int_array(q, 8);
    entropy_free(e);
    return total;
}

/* ---- float pixel tiles: the block is filled by hand as tests/test_dct.c:46-50 does, then the
 * reference's dct_forward / quantize ---- */
int ref_fwd_quant_plane_f32(const float *px, size_t pitch_floats, int W, int H, const double *Q, int adaptive,
                            int layout, int16_t *coef, double *var_out, int nthreads, uint64_t *near_ties)
{
    (void)nthreads;
    if (W <= 0 || H <= 0 || W % 8 || H % 8) return -1;
    DCTContext *d = dct_init(8);
    QuantContext *qc = ctx_with_table(8, Q, adaptive);
    double **blk = alloc_array(8, 8), **c = alloc_array(8, 8);
    int **q = alloc_int_array(8, 8);
    int zz[64];
    const int bw = W / 8;
    for (int by = 0; by < H / 8; ++by)
        for (int bx = 0; bx < bw; ++bx) {
            const size_t b = (size_t)by * bw + bx;
            for (int i = 0; i < 8; ++i)
                for (int j = 0; j < 8; ++j)
                    blk[i][j] = (double)px[((size_t)by * 8 + i) * pitch_floats + (size_t)bx * 8 + j] - 128.0;
            dct_forward(d, blk, c);
            double var = 0.0;
            if (adaptive) {
                var = calculate_block_variance(blk, 8);
                if (var_out) var_out[b] = var;
            }
            quantize(qc, c, q, var);
            int16_t *dst = coef + b * 64;
            if (layout == 1) {
                block_to_zigzag(q, zz, 8);
                for (int k = 0; k < 64; ++k) dst[k] = (int16_t)zz[k];
            } else {
                for (int k = 0; k < 64; ++k) dst[k] = (int16_t)q[k / 8][k % 8];
            }
        }
    if (near_ties) *near_ties = 0;
    free_array(blk, 8);
    free_array(c, 8);
    free_int_array(q, 8);
    dct_free(d);
    quant_free(qc);
    return 0;
}

/* ---- any block size n: the reference's block functions with block_size = n ---- */
int ref_fwd_quant_plane_n(int n, const uint8_t *px, size_t pitch, int W, int H, const double *Q, int adaptive,
                          int layout, int16_t *coef, double *var_out, int nthreads, uint64_t *near_ties)
{
    (void)nthreads;
    if (n < 1 || W <= 0 || H <= 0 || W % n || H % n) return -1;
    DCTContext *d = dct_init(n);
    QuantContext *qc = ctx_with_table(n, Q, adaptive);
    double **c = alloc_array(n, n);
    int **q = alloc_int_array(n, n);
    int *zz = (int *)malloc(sizeof(int) * n * n);
    const int bw = W / n;
    for (int by = 0; by < H / n; ++by)
        for (int bx = 0; bx < bw; ++bx) {
            const size_t b = (size_t)by * bw + bx;
            unsigned char *strip = (unsigned char *)px + (size_t)by * n * pitch;
            double **blk = create_block_from_pixels(strip, (int)pitch, 0, bx * n, n);
            dct_forward(d, blk, c);
            double var = 0.0;
            if (adaptive) {
                var = calculate_block_variance(blk, n);
                if (var_out) var_out[b] = var;
            }
            quantize(qc, c, q, var);
            int16_t *dst = coef + b * n * n;
            if (layout == 1) {
                block_to_zigzag(q, zz, n);
                for (int k = 0; k < n * n; ++k) dst[k] = (int16_t)zz[k];
            } else {
                for (int k = 0; k < n * n; ++k) dst[k] = (int16_t)q[k / n][k % n];
            }
            free_array(blk, n);
        }
    if (near_ties) *near_ties = 0;
    free(zz);
    free_array(c, n);
    free_int_array(q, n);
    dct_free(d);
    quant_free(qc);
    return 0;
}

int ref_dequant_idct_plane_n(int n, const int16_t *coef, int W, int H, const double *Q, const double *R,
                             int adaptive, int layout, const double *var_in, uint8_t *px, size_t pitch,
                             int nthreads, uint64_t *near_ties)
{
    (void)R;
    (void)nthreads;
    if (n < 1 || W <= 0 || H <= 0 || W % n || H % n) return -1;
    DCTContext *d = dct_init(n);
    QuantContext *qc = ctx_with_table(n, Q, adaptive);
    double **c = alloc_array(n, n), **o = alloc_array(n, n);
    int **q = alloc_int_array(n, n);
    int *zz = (int *)malloc(sizeof(int) * n * n);
    const int bw = W / n;
    for (int by = 0; by < H / n; ++by)
        for (int bx = 0; bx < bw; ++bx) {
            const size_t b = (size_t)by * bw + bx;
            const int16_t *src = coef + b * n * n;
            if (layout == 1) {
                for (int k = 0; k < n * n; ++k) zz[k] = src[k];
                zigzag_to_block(zz, q, n);
            } else {
                for (int k = 0; k < n * n; ++k) q[k / n][k % n] = src[k];
            }
            dequantize(qc, q, c, (adaptive && var_in) ? var_in[b] : 0.0);
            dct_inverse(d, c, o);
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    double v = round(o[i][j] + 128.0);
                    v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
                    px[((size_t)by * n + i) * pitch + (size_t)bx * n + j] = (uint8_t)v;
                }
        }
    if (near_ties) *near_ties = 0;
    free(zz);
    free_array(c, n);
    free_array(o, n);
    free_int_array(q, n);
    dct_free(d);
    quant_free(qc);
    return 0;
}
