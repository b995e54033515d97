/*
 * c_roundtrip.c -- a C99 host program on top of libdct_cuda, the way a maintainer of the reference
 * would write it: the reference's own headers and its UNTOUCHED entropy.c / utils.c, plus the plane
 * calls of dct_cuda.h.  It checks, in plain C and with no test framework (like the reference's tests):
 *   1. the plane call equals the reference-style per-block sequence
 *        create_block_from_pixels -> dct_forward -> quantize          (tests/test_entropy.c:302-316)
 *      for every block (both go through libdct_cuda; the per-block calls are the drop-in symbols);
 *   2. the records feed the untouched run_length_encode / run_length_decode unchanged
 *        (src/entropy.c:216, :333) through dct_cuda_record_to_block;
 *   3. the inverse plane call equals dequantize -> dct_inverse -> +128, round, clamp per block.
 * Built by `make -C oracle dropin` (needs the reference's entropy.c and utils.c), run by
 * tests/test_gpu_parity.py.  Exit code 0 = all equal.
 */
#include <dct.h>
#include <dct_cuda.h>
#include <entropy.h>
#include <quantization.h>

#include <stdint.h>

#define W 136
#define H 72

int main(void)
{
    static unsigned char pixels[W * H], out_plane[W * H];
    static int16_t coef[(W / 8) * (H / 8) * 64];
    uint32_t s = 2463534242u;
    for (int i = 0; i < W * H; ++i) {
        s ^= s << 13, s ^= s >> 17, s ^= s << 5;
        pixels[i] = (unsigned char)(s >> 11);
    }

    DCTContext *d = dct_init(8);
    QuantContext *q = quant_init(8, 75, 0);
    dct_cuda_plan *plan = dct_cuda_plan_create(d, q, 0);
    if (!plan) {
        fprintf(stderr, "c_roundtrip: %s\n", dct_cuda_last_error());
        return 2;
    }
    dct_cuda_stats st;
    if (dct_cuda_fwd_quant_u8(plan, pixels, W, W, H, coef, DCT_CUDA_ZIGZAG, NULL, &st) != DCT_CUDA_OK) {
        fprintf(stderr, "c_roundtrip: %s\n", dct_cuda_last_error());
        return 2;
    }

    EntropyContext *e = entropy_init(0);
    double **c = alloc_array(8, 8), **dq = alloc_array(8, 8), **px = alloc_array(8, 8);
    int **qb = alloc_int_array(8, 8), **rec = alloc_int_array(8, 8), **dec = alloc_int_array(8, 8);
    long mismatches = 0, symbols = 0;
    for (int by = 0; by < H / 8; ++by) {
        for (int bx = 0; bx < W / 8; ++bx) {
            const size_t b = (size_t)by * (W / 8) + bx;
            double **blk = create_block_from_pixels(pixels, W, by * 8, bx * 8, 8);
            dct_forward(d, blk, c);
            quantize(q, c, qb, 0.0);
            dct_cuda_record_to_block(coef + b * 64, DCT_CUDA_ZIGZAG, rec);
            for (int i = 0; i < 8; ++i)
                for (int j = 0; j < 8; ++j) mismatches += rec[i][j] != qb[i][j];
            symbols += run_length_encode(e, rec, 8);      /* untouched host consumer */
            run_length_decode(e, dec, 8);
            for (int i = 0; i < 8; ++i)
                for (int j = 0; j < 8; ++j) mismatches += dec[i][j] != rec[i][j];
            free_array(blk, 8);
        }
    }

    if (dct_cuda_dequant_idct_u8(plan, coef, W, H, DCT_CUDA_ZIGZAG, NULL, out_plane, W, NULL) != DCT_CUDA_OK) {
        fprintf(stderr, "c_roundtrip: %s\n", dct_cuda_last_error());
        return 2;
    }
    for (int by = 0; by < H / 8; ++by) {
        for (int bx = 0; bx < W / 8; ++bx) {
            const size_t b = (size_t)by * (W / 8) + bx;
            dct_cuda_record_to_block(coef + b * 64, DCT_CUDA_ZIGZAG, rec);
            dequantize(q, rec, dq, 0.0);
            dct_inverse(d, dq, px);
            for (int i = 0; i < 8; ++i)
                for (int j = 0; j < 8; ++j) {
                    double v = round(px[i][j] + 128.0);
                    v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
                    mismatches += out_plane[(by * 8 + i) * W + bx * 8 + j] != (unsigned char)v;
                }
        }
    }

    printf("blocks %llu  replayed %llu  exact ties %llu  RLE symbols %ld  mismatches %ld\n",
           (unsigned long long)st.blocks, (unsigned long long)st.replayed_blocks, (unsigned long long)st.near_ties,
           symbols, mismatches);
    printf(mismatches == 0 ? "C ROUNDTRIP PASSED\n" : "C ROUNDTRIP FAILED\n");

    free_array(c, 8), free_array(dq, 8), free_array(px, 8);
    free_int_array(qb, 8), free_int_array(rec, 8), free_int_array(dec, 8);
    entropy_free(e);
    dct_cuda_plan_destroy(plan);
    dct_free(d);
    quant_free(q);
    return mismatches == 0 ? 0 : 1;
}
