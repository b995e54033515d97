/*
 * dct_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see dct_oracle.h).
 *
 * Each function names the reference lines it restates.  The arithmetic order is the
 * reference's: accumulators start at 0.0 and add products for k ascending, no fused
 * multiply-add (build with -std=c99 / -ffp-contract=off), true fp64 division in
 * quantize, C99 round() (half away from zero).
 */
#define _POSIX_C_SOURCE 199309L
#include "dct_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323846 /* include/dct.h:15 */

/* src/dct.c:19-30 -- D[i][j] = alpha_i * cos(PI*(2j+1)*i / (2n)) */
void orc_dct_matrix(int n, double *D)
{
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            double alpha = (i == 0) ? 1.0 / sqrt(n) : sqrt(2.0 / n);
            D[i * n + j] = alpha * cos((ORC_PI * (2 * j + 1) * i) / (2.0 * n));
        }
    }
}

/* src/dct.c:52-77 -- temp = X * D^T (rows), out = D * temp (columns) */
void orc_dct_forward(int n, const double *D, const double *in, double *out)
{
    double temp[32 * 32];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k)
                acc += in[i * n + k] * D[j * n + k]; /* transposed_dct[k][j] == D[j][k] */
            temp[i * n + j] = acc;
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k)
                acc += D[i * n + k] * temp[k * n + j];
            out[i * n + j] = acc;
        }
}

/* src/dct.c:80-105 -- temp = D^T * in (columns), out = temp * D (rows) */
void orc_dct_inverse(int n, const double *D, const double *in, double *out)
{
    double temp[32 * 32];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k)
                acc += D[k * n + i] * in[k * n + j]; /* transposed_dct[i][k] == D[k][i] */
            temp[i * n + j] = acc;
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k)
                acc += temp[i * n + k] * D[k * n + j];
            out[i * n + j] = acc;
        }
}

/* src/quantization.c:8-17 (Annex-K luma), :26-31 (quality clamp), :51-99 */
static const int k_luma[64] = {
    16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
    14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
    18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};

void orc_quant_table(int n, int quality, double *Q)
{
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    double scale = (quality < 50) ? 5000.0 / quality : 200.0 - 2 * quality;
    scale /= 100.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double v;
            if (n == 8) {
                v = k_luma[i * 8 + j] * scale;
            } else {
                double distance = sqrt((double)(i * i + j * j));
                v = (1.0 + distance) * scale * 8.0;
            }
            if (v < 1.0) v = 1.0;
            if (v > 255.0) v = 255.0;
            Q[i * n + j] = v;
        }
}

/* src/quantization.c:101-111 */
void orc_dequant_table(int n, const double *Q, double *R)
{
    for (int k = 0; k < n * n; ++k) R[k] = 1.0 / Q[k];
}

/* src/quantization.c:153-169 */
double orc_block_variance(int n, const double *blk)
{
    double sum = 0.0, sum_sq = 0.0;
    int count = n * n;
    for (int k = 0; k < count; ++k) {
        sum += blk[k];
        sum_sq += blk[k] * blk[k];
    }
    double mean = sum / count;
    return (sum_sq / count) - (mean * mean);
}

/* src/quantization.c:171-211 */
void orc_adjust_table(int n, const double *src, double variance, int is_quantize, double *out)
{
    double nv = fmin(1.0, fmax(0.1, variance / 1000.0));
    double scale = is_quantize ? 2.0 - nv : 1.0 / (2.0 - nv);
    for (int k = 0; k < n * n; ++k) {
        if (k == 0) {
            out[k] = src[k];
        } else {
            out[k] = src[k] * scale;
            if (is_quantize && out[k] < 1.0) out[k] = 1.0;
        }
    }
}

/* src/quantization.c:113-131 */
void orc_quantize(int n, const double *Q, int adaptive, const double *c, int *q, double variance)
{
    double adj[32 * 32];
    const double *m = Q;
    if (adaptive) {
        orc_adjust_table(n, Q, variance, 1, adj);
        m = adj;
    }
    for (int k = 0; k < n * n; ++k) q[k] = (int)round(c[k] / m[k]);
}

/* src/quantization.c:133-151 -- note the non-adaptive branch multiplies by R = 1/Q */
void orc_dequantize(int n, const double *Q, const double *R, int adaptive, const int *q, double *c,
                    double variance)
{
    (void)Q;
    double adj[32 * 32];
    const double *m = R;
    if (adaptive) {
        orc_adjust_table(n, R, variance, 0, adj);
        m = adj;
    }
    for (int k = 0; k < n * n; ++k) c[k] = q[k] * (adaptive ? 1.0 / m[k] : m[k]);
}

/* src/entropy.c:158-178 -- anti-diagonals; even ones walk up-right, odd ones down-left */
void orc_zigzag_order(int n, int *order)
{
    int idx = 0;
    for (int s = 0; s <= 2 * (n - 1); ++s) {
        if (s % 2 == 0) {
            for (int i = (s < n) ? s : n - 1; i >= 0 && (s - i) < n; --i) order[idx++] = i * n + (s - i);
        } else {
            for (int i = (s < n) ? 0 : s - n + 1; i < n && (s - i) >= 0; ++i) order[idx++] = i * n + (s - i);
        }
    }
}

/* src/dct.c:123-129 */
void orc_round_to_int(int n, const double *blk, int *out)
{
    for (int k = 0; k < n * n; ++k) out[k] = (int)round(blk[k]);
}

/* ------------------------------------------------------------------------ */
/* Plane loops.  The reference has no image-level code: the   */
/* oracle for a plane is "its block functions looped over the 8x8 grid",     */
/* composed exactly as tests/test_entropy.c:302-316 and :370-384 compose     */
/* them, plus the pixel rule  p = clamp(round(x + 128.0), 0, 255).           */
/* ------------------------------------------------------------------------ */

static int near_half(double a)
{
    a = fabs(a);
    double f = a - floor(a);
    return fabs(f - 0.5) <= 1e-9;
}

typedef struct {
    int n;   /* block size, 0 = 8 */
    int dir; /* 0 fwd, 1 inv */
    const float *pxf_in; /* forward from float pixels when non-NULL (pitch in floats) */
    const uint8_t *px_in;
    uint8_t *px_out;
    size_t pitch;
    int W, H;
    const double *Q, *R;
    int adaptive, layout;
    int16_t *coef_out;
    const int16_t *coef_in;
    double *var_out;
    const double *var_in;
    int row0, row1; /* block rows */
    uint64_t ties;
    const double *D;
    const int *zz;
} job_t;

static void *plane_worker(void *arg)
{
    job_t *jb = (job_t *)arg;
    const int n = jb->n, nn = n * n;
    const int bw = jb->W / n;
    double *blk = malloc(sizeof(double) * nn), *c = malloc(sizeof(double) * nn), *adj = malloc(sizeof(double) * nn);
    int *q = malloc(sizeof(int) * nn);
    uint64_t ties = 0;
    for (int by = jb->row0; by < jb->row1; ++by) {
        for (int bx = 0; bx < bw; ++bx) {
            size_t b = (size_t)by * bw + bx;
            if (jb->dir == 0) {
                /* src/dct.c:109-120 */
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < n; ++j) {
                        const size_t at = ((size_t)by * n + i) * jb->pitch + (size_t)bx * n + j;
                        /* float tiles: the block a caller fills by hand, tests/test_dct.c:46-50 */
                        blk[i * n + j] = (jb->pxf_in ? (double)jb->pxf_in[at] : (double)jb->px_in[at]) - 128.0;
                    }
                orc_dct_forward(n, jb->D, blk, c);
                double var = 0.0;
                const double *m = jb->Q;
                if (jb->adaptive) {
                    var = orc_block_variance(n, blk); /* tests/test_entropy.c:315 */
                    if (jb->var_out) jb->var_out[b] = var;
                    orc_adjust_table(n, jb->Q, var, 1, adj);
                    m = adj;
                }
                orc_quantize(n, jb->Q, jb->adaptive, c, q, var);
                for (int k = 0; k < nn; ++k) ties += near_half(c[k] / m[k]);
                int16_t *dst = jb->coef_out + b * nn;
                if (jb->layout == ORC_LAYOUT_ZIGZAG)
                    for (int k = 0; k < nn; ++k) dst[k] = (int16_t)q[jb->zz[k]];
                else
                    for (int k = 0; k < nn; ++k) dst[k] = (int16_t)q[k];
            } else {
                const int16_t *src = jb->coef_in + b * nn;
                if (jb->layout == ORC_LAYOUT_ZIGZAG) {
                    for (int k = 0; k < nn; ++k) q[jb->zz[k]] = src[k]; /* src/entropy.c:183-210 */
                } else {
                    for (int k = 0; k < nn; ++k) q[k] = src[k];
                }
                double var = (jb->adaptive && jb->var_in) ? jb->var_in[b] : 0.0;
                orc_dequantize(n, jb->Q, jb->R, jb->adaptive, q, c, var);
                orc_dct_inverse(n, jb->D, c, blk);
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < n; ++j) {
                        double v = blk[i * n + j] + 128.0;
                        ties += near_half(v);
                        double r = round(v);
                        if (r < 0.0) r = 0.0;
                        if (r > 255.0) r = 255.0;
                        jb->px_out[((size_t)by * n + i) * jb->pitch + (size_t)bx * n + j] = (uint8_t)r;
                    }
            }
        }
    }
    free(blk), free(c), free(adj), free(q);
    jb->ties = ties;
    return NULL;
}

static int run_plane(job_t *proto, int nthreads, uint64_t *near_ties)
{
    const int n = proto->n ? proto->n : 8;
    if (n < 1 || n > 32 || proto->W <= 0 || proto->H <= 0 || proto->W % n || proto->H % n) return -1;
    double D[32 * 32];
    int zz[32 * 32];
    orc_dct_matrix(n, D);
    orc_zigzag_order(n, zz);
    const int bh = proto->H / n;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > bh) nthreads = bh;
    if (nthreads > 256) nthreads = 256;
    job_t jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto;
        jobs[t].n = n;
        jobs[t].D = D;
        jobs[t].zz = zz;
        jobs[t].row0 = (int)((long long)bh * t / nthreads);
        jobs[t].row1 = (int)((long long)bh * (t + 1) / nthreads);
    }
    if (nthreads == 1) {
        plane_worker(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, plane_worker, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    }
    uint64_t ties = 0;
    for (int t = 0; t < nthreads; ++t) ties += jobs[t].ties;
    if (near_ties) *near_ties = ties;
    return 0;
}

/* generic block size n (the reference's block_size argument, include/dct.h:34): tables are n*n */
int orc_fwd_quant_plane_n(int n, const uint8_t *px, size_t pitch, int W, int H, const double *Q, int adaptive,
                          int layout, int16_t *coef, double *var_out, int nthreads, uint64_t *near_ties)
{
    job_t jb;
    memset(&jb, 0, sizeof jb);
    jb.n = n;
    jb.dir = 0;
    jb.px_in = px;
    jb.pitch = pitch;
    jb.W = W;
    jb.H = H;
    jb.Q = Q;
    jb.adaptive = adaptive;
    jb.layout = layout;
    jb.coef_out = coef;
    jb.var_out = var_out;
    return run_plane(&jb, nthreads, near_ties);
}

int orc_dequant_idct_plane_n(int n, const int16_t *coef, int W, int H, const double *Q, const double *R,
                             int adaptive, int layout, const double *var_in, uint8_t *px, size_t pitch,
                             int nthreads, uint64_t *near_ties)
{
    job_t jb;
    memset(&jb, 0, sizeof jb);
    jb.n = n;
    jb.dir = 1;
    jb.coef_in = coef;
    jb.px_out = px;
    jb.pitch = pitch;
    jb.W = W;
    jb.H = H;
    jb.Q = Q;
    jb.R = R;
    jb.adaptive = adaptive;
    jb.layout = layout;
    jb.var_in = var_in;
    return run_plane(&jb, nthreads, near_ties);
}

int orc_fwd_quant_plane(const uint8_t *px, size_t pitch, int W, int H, const double *Q, int adaptive,
                        int layout, int16_t *coef, double *var_out, int nthreads, uint64_t *near_ties)
{
    job_t jb;
    memset(&jb, 0, sizeof jb);
    jb.dir = 0;
    jb.px_in = px;
    jb.pitch = pitch;
    jb.W = W;
    jb.H = H;
    jb.Q = Q;
    jb.adaptive = adaptive;
    jb.layout = layout;
    jb.coef_out = coef;
    jb.var_out = var_out;
    return run_plane(&jb, nthreads, near_ties);
}

int orc_fwd_quant_plane_f32(const float *px, size_t pitch_floats, int W, int H, const double *Q, int adaptive,
                            int layout, int16_t *coef, double *var_out, int nthreads, uint64_t *near_ties)
{
    job_t jb;
    memset(&jb, 0, sizeof jb);
    jb.dir = 0;
    jb.pxf_in = px;
    jb.pitch = pitch_floats;
    jb.W = W;
    jb.H = H;
    jb.Q = Q;
    jb.adaptive = adaptive;
    jb.layout = layout;
    jb.coef_out = coef;
    jb.var_out = var_out;
    return run_plane(&jb, nthreads, near_ties);
}

int orc_dequant_idct_plane(const int16_t *coef, int W, int H, const double *Q, const double *R,
                           int adaptive, int layout, const double *var_in, uint8_t *px, size_t pitch,
                           int nthreads, uint64_t *near_ties)
{
    job_t jb;
    memset(&jb, 0, sizeof jb);
    jb.dir = 1;
    jb.coef_in = coef;
    jb.px_out = px;
    jb.pitch = pitch;
    jb.W = W;
    jb.H = H;
    jb.Q = Q;
    jb.R = R;
    jb.adaptive = adaptive;
    jb.layout = layout;
    jb.var_in = var_in;
    return run_plane(&jb, nthreads, near_ties);
}

/* ------------------------------------------------------------------------ */
/* Run-length symbols of every record, as run_length_encode (src/entropy.c:  */
/* 216-256) produces them: zigzag order, one (value, zero-run) symbol per     */
/* non-zero coefficient, plus a closing symbol at position 63 whose run       */
/* counts a zero last coefficient too.  symbols == NULL only counts.          */
/* ------------------------------------------------------------------------ */
size_t orc_rle_plane(const int16_t *coef, size_t nblocks, int layout, uint32_t *offsets, int32_t *symbols)
{
    int zz[64];
    orc_zigzag_order(8, zz);
    size_t total = 0;
    for (size_t b = 0; b < nblocks; ++b) {
        const int16_t *rec = coef + b * 64;
        int zero_count = 0;
        offsets[b] = (uint32_t)total;
        for (int i = 0; i < 64; ++i) {
            const int v = layout == ORC_LAYOUT_ZIGZAG ? rec[i] : rec[zz[i]];
            if (v != 0 || i == 63) {
                if (i == 63 && v == 0) zero_count++;
                if (symbols) {
                    symbols[2 * total] = v;
                    symbols[2 * total + 1] = zero_count;
                }
                ++total;
                zero_count = 0;
            } else {
                zero_count++;
            }
        }
    }
    offsets[nblocks] = (uint32_t)total;
    return total;
}

/* ------------------------------------------------------------------------ */
/* Planar front / back end.                              */
/* PARITY UNPINNED for this section: the reference has no colour conversion,    */
/* no subsampling and no edge handling (src/dct.c:109-120 reads out of bounds). */
/* The convention is ours (dct_b200/csrc/planar.cu states it); this is a        */
/* second, independent statement of it in signed 64-bit arithmetic with an       */
/* explicit floor division, anchored only by the published JFIF known answers    */
/* that tests/test_oracle.py checks (grey stays grey, the primaries).            */
/* ------------------------------------------------------------------------ */
static long long floor_div(long long a, long long b) /* b > 0 */
{
    long long q = a / b;
    if ((a % b) != 0 && a < 0) --q;
    return q;
}
static int clamp_u8(long long v) { return v < 0 ? 0 : (v > 255 ? 255 : (int)v); }
static int imin(int a, int b) { return a < b ? a : b; }

void orc_rgb_to_ycbcr420(const uint8_t *rgb, size_t rgb_pitch, int W, int H, uint8_t *y, size_t y_pitch, int y_w,
                         int y_h, uint8_t *cb, uint8_t *cr, size_t c_pitch, int c_w, int c_h)
{
    if (W <= 0 || H <= 0) return;
    for (int j = 0; j < y_h; ++j)
        for (int i = 0; i < y_w; ++i) {
            const uint8_t *s = rgb + (size_t)imin(j, H - 1) * rgb_pitch + (size_t)imin(i, W - 1) * 3;
            y[(size_t)j * y_pitch + i] =
                (uint8_t)floor_div(19595LL * s[0] + 38470LL * s[1] + 7471LL * s[2] + 32768, 65536);
        }
    const int cw_img = (W + 1) / 2, ch_img = (H + 1) / 2;
    for (int j = 0; j < c_h; ++j)
        for (int i = 0; i < c_w; ++i) {
            const int sx = 2 * imin(i, cw_img - 1), sy = 2 * imin(j, ch_img - 1);
            long long sum[3] = {0, 0, 0};
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx) {
                    const uint8_t *s = rgb + (size_t)imin(sy + dy, H - 1) * rgb_pitch + (size_t)imin(sx + dx, W - 1) * 3;
                    for (int c = 0; c < 3; ++c) sum[c] += s[c];
                }
            const long long bias = 4LL * (128LL << 16) + 4LL * 32768 - 1; /* (128<<18) + (1<<17) - 1 */
            cb[(size_t)j * c_pitch + i] =
                (uint8_t)floor_div(-11059LL * sum[0] - 21709LL * sum[1] + 32768LL * sum[2] + bias, 4 * 65536);
            cr[(size_t)j * c_pitch + i] =
                (uint8_t)floor_div(32768LL * sum[0] - 27439LL * sum[1] - 5329LL * sum[2] + bias, 4 * 65536);
        }
}

void orc_ycbcr420_to_rgb(const uint8_t *y, size_t y_pitch, const uint8_t *cb, const uint8_t *cr, size_t c_pitch, int W,
                         int H, uint8_t *rgb, size_t rgb_pitch)
{
    for (int j = 0; j < H; ++j)
        for (int i = 0; i < W; ++i) {
            const long long l = y[(size_t)j * y_pitch + i];
            const long long u = (long long)cb[(size_t)(j / 2) * c_pitch + i / 2] - 128;
            const long long v = (long long)cr[(size_t)(j / 2) * c_pitch + i / 2] - 128;
            uint8_t *o = rgb + (size_t)j * rgb_pitch + (size_t)i * 3;
            o[0] = (uint8_t)clamp_u8(l + floor_div(91881 * v + 32768, 65536));
            o[1] = (uint8_t)clamp_u8(l + floor_div(-22554 * u - 46802 * v + 32768, 65536));
            o[2] = (uint8_t)clamp_u8(l + floor_div(116130 * u + 32768, 65536));
        }
}

/* complete a W x H plane of `elem`-byte elements to Wp x Hp in place: last column / row replicated */
void orc_pad_edges(uint8_t *px, size_t pitch, int W, int H, int Wp, int Hp, int elem)
{
    if (W <= 0 || H <= 0) return;
    for (int j = 0; j < Hp; ++j)
        for (int i = 0; i < Wp; ++i) {
            if (i < W && j < H) continue;
            memcpy(px + (size_t)j * pitch + (size_t)i * elem,
                   px + (size_t)imin(j, H - 1) * pitch + (size_t)imin(i, W - 1) * elem, (size_t)elem);
        }
}

/* ------------------------------------------------------------------------ */
/* Synthetic inputs and hashes                  */
/* ------------------------------------------------------------------------ */

/* dist 0 = U (uniform), 1 = S (smooth + small noise).  Row-major fill; W is only used by S. */
void orc_fill_xorshift(uint8_t *dst, size_t n, uint64_t seed, int dist, int W)
{
    uint64_t s = seed;
    for (size_t i = 0; i < n; ++i) {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        uint32_t out = (uint32_t)(s >> 32);
        if (dist == 0) {
            dst[i] = (uint8_t)(out & 255u);
        } else {
            size_t x = W > 0 ? i % (size_t)W : i, y = W > 0 ? i / (size_t)W : 0;
            int v = 128 + (int)(100.0 * sin(0.05 * (double)x) * cos(0.03 * (double)y)) + (int)(out % 9u) - 4;
            dst[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
}

#define FNV_INIT 1469598103934665603ULL
#define FNV_MUL  1099511628211ULL

uint64_t orc_fnv_i16(const int16_t *v, size_t n)
{
    uint64_t h = FNV_INIT;
    for (size_t i = 0; i < n; ++i) h = (h ^ (uint32_t)(int)v[i]) * FNV_MUL;
    return h;
}

uint64_t orc_fnv_u8(const uint8_t *v, size_t n)
{
    uint64_t h = FNV_INIT;
    for (size_t i = 0; i < n; ++i) h = (h ^ (uint32_t)v[i]) * FNV_MUL;
    return h;
}

/* pixels visited block by block (block raster, then natural order inside the block) */
uint64_t orc_fnv_u8_blockorder(const uint8_t *px, size_t pitch, int W, int H)
{
    uint64_t h = FNV_INIT;
    for (int by = 0; by < H / 8; ++by)
        for (int bx = 0; bx < W / 8; ++bx)
            for (int i = 0; i < 8; ++i)
                for (int j = 0; j < 8; ++j)
                    h = (h ^ (uint32_t)px[((size_t)by * 8 + i) * pitch + (size_t)bx * 8 + j]) * FNV_MUL;
    return h;
}
