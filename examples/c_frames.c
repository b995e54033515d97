/*
 * c_frames.c -- the entry points around the hot path, from plain C99 (-Wall -Wextra -Werror -pedantic):
 *   1. a plane whose sides are not multiples of 8 (dct_cuda_fwd_quant_u8_edge): equals the plain call on the same
 *      plane completed by hand with its last column / row, and decodes to the same pixels (cropped);
 *   2. int8 records (dct_cuda_fwd_quant_u8_i8 / dct_cuda_dequant_idct_i8_u8): the same values as the int16 records;
 *   3. an interleaved RGB frame (dct_cuda_encode_rgb420 / dct_cuda_decode_rgb420): grey frames come back exactly
 *      at quality 100 up to the reference's own (lossy) dequantisation -- checked here only for
 *      self-consistency: decoding the records twice gives the same frame, and a flat grey frame keeps Cb = Cr = 128
 *      (all chroma AC coefficients zero).
 * No reference sources are needed: only libdct_cuda and its headers.  Built by `make -C oracle dropin`, run by
 * tests/test_gpu_parity.py.  Exit code 0 = all checks passed.
 */
#include <dct.h>
#include <dct_cuda.h>
#include <quantization.h>

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define W 203
#define H 77
#define WP 208
#define HP 80

static int fail(const char *what)
{
    fprintf(stderr, "c_frames: %s: %s\n", what, dct_cuda_last_error());
    return 2;
}

int main(void)
{
    static unsigned char px[W * H], padded[WP * HP], rec_a[W * H], rec_b[WP * HP], rec_c[WP * HP];
    static int16_t coef_a[(WP / 8) * (HP / 8) * 64], coef_b[(WP / 8) * (HP / 8) * 64];
    static int8_t coef8[(WP / 8) * (HP / 8) * 64];
    uint32_t s = 88172645u;
    long bad = 0;
    for (int i = 0; i < W * H; ++i) {
        s ^= s << 13, s ^= s >> 17, s ^= s << 5;
        px[i] = (unsigned char)(s >> 9);
    }
    for (int y = 0; y < HP; ++y)
        for (int x = 0; x < WP; ++x) padded[y * WP + x] = px[(y < H ? y : H - 1) * W + (x < W ? x : W - 1)];

    DCTContext *d = dct_init(8);
    QuantContext *q = quant_init(8, 50, 0);
    dct_cuda_plan *plan = dct_cuda_plan_create(d, q, 0);
    if (!plan) return fail("plan");

    /* 1. ragged plane == hand-padded plane */
    if (dct_cuda_fwd_quant_u8_edge(plan, px, W, W, H, coef_a, DCT_CUDA_NATURAL, NULL, NULL)) return fail("fwd edge");
    if (dct_cuda_fwd_quant_u8(plan, padded, WP, WP, HP, coef_b, DCT_CUDA_NATURAL, NULL, NULL)) return fail("fwd padded");
    bad += memcmp(coef_a, coef_b, sizeof coef_a) != 0;
    if (dct_cuda_dequant_idct_u8_edge(plan, coef_a, W, H, DCT_CUDA_NATURAL, NULL, rec_a, W, NULL)) return fail("inv edge");
    if (dct_cuda_dequant_idct_u8(plan, coef_b, WP, HP, DCT_CUDA_NATURAL, NULL, rec_b, WP, NULL)) return fail("inv padded");
    for (int y = 0; y < H; ++y) bad += memcmp(rec_a + y * W, rec_b + y * WP, W) != 0;

    /* 2. int8 records == int16 records (quality 50: every table entry >= 8.03) */
    if (!dct_cuda_plan_records_fit_i8(plan)) return fail("records_fit_i8");
    if (dct_cuda_fwd_quant_u8_i8(plan, padded, WP, WP, HP, coef8, DCT_CUDA_NATURAL, NULL, NULL)) return fail("fwd i8");
    for (size_t k = 0; k < sizeof coef8; ++k) bad += coef8[k] != coef_b[k];
    if (dct_cuda_dequant_idct_i8_u8(plan, coef8, WP, HP, DCT_CUDA_NATURAL, NULL, rec_c, WP, NULL)) return fail("inv i8");
    bad += memcmp(rec_b, rec_c, sizeof rec_b) != 0;

    /* 3. RGB frame: geometry, encode, decode twice, flat grey keeps neutral chroma */
    dct_cuda_frame420 g;
    dct_cuda_frame420_geometry(W, H, &g);
    bad += g.y_width != WP || g.y_height != HP || g.c_width != 104 || g.c_height != 40;
    unsigned char *rgb = malloc((size_t)W * H * 3), *back1 = malloc((size_t)W * H * 3), *back2 = malloc((size_t)W * H * 3);
    int16_t *ky = malloc((size_t)g.y_width * g.y_height * 2), *kcb = malloc((size_t)g.c_width * g.c_height * 2),
            *kcr = malloc((size_t)g.c_width * g.c_height * 2);
    if (!rgb || !back1 || !back2 || !ky || !kcb || !kcr) return 3;
    for (int i = 0; i < W * H; ++i) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = 90;      /* flat grey */
    if (dct_cuda_encode_rgb420(plan, plan, rgb, (size_t)W * 3, W, H, ky, kcb, kcr, DCT_CUDA_ZIGZAG, NULL)) return fail("encode");
    for (int k = 1; k < g.c_width * g.c_height; ++k)
        if (k % 64) bad += kcb[k] != 0 || kcr[k] != 0;   /* AC of a flat plane */
    bad += kcb[0] != 0 || kcr[0] != 0;                   /* DC of (128 - 128) */
    for (int i = 0; i < W * H; ++i) {
        s ^= s << 13, s ^= s >> 17, s ^= s << 5;
        rgb[3 * i] = (unsigned char)s, rgb[3 * i + 1] = (unsigned char)(s >> 8), rgb[3 * i + 2] = (unsigned char)(s >> 16);
    }
    if (dct_cuda_encode_rgb420(plan, plan, rgb, (size_t)W * 3, W, H, ky, kcb, kcr, DCT_CUDA_ZIGZAG, NULL)) return fail("encode 2");
    if (dct_cuda_decode_rgb420(plan, plan, ky, kcb, kcr, W, H, DCT_CUDA_ZIGZAG, back1, (size_t)W * 3, NULL)) return fail("decode");
    if (dct_cuda_decode_rgb420(plan, plan, ky, kcb, kcr, W, H, DCT_CUDA_ZIGZAG, back2, (size_t)W * 3, NULL)) return fail("decode 2");
    bad += memcmp(back1, back2, (size_t)W * H * 3) != 0;

    printf("c_frames: %ld mismatches, %llu kernels launched\n", bad, (unsigned long long)dct_cuda_plan_kernel_launches(plan));
    free(rgb), free(back1), free(back2), free(ky), free(kcb), free(kcr);
    dct_cuda_plan_destroy(plan);
    quant_free(q);
    dct_free(d);
    return bad ? 1 : 0;
}
