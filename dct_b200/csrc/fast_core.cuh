// fast_core.cuh -- the fp32 building blocks shared by the fused kernels (K1/K2) and by K3's
// re-flagging phase.  K3 must reproduce K1's / K2's fp32 values BIT FOR BIT (same operations in
// the same order on the same inputs) to find out which single values fell inside the error band,
// so both sides call these functions and nothing else.
#pragma once
#include <stdint.h>

#include "butterfly.cuh"

namespace dctb {

constexpr float kMagic = 12582912.0f;      // 1.5 * 2^23: x + kMagic rounds x to an integer (RNE)
constexpr float kMagic128 = 12583040.0f;   // 1.5 * 2^23 + 128

// byte `idx` of w -> 2^15 + byte as a float: (0x47000000 | byte << 8), one PRMT, no conversion instruction
// (the byte lands in mantissa bits 8..15, whose weight at exponent 15 is 1)
template <int idx> __device__ __forceinline__ float byte_to_magic(uint32_t w)
{
    return __uint_as_float(__byte_perm(w, 0x47000000u, 0x7504 + (idx << 4)));
}

// bias that the magic-biased bytes leave in output 0 of a row transform, plus the level shift of its 8 samples:
// 8 * 2^15 + 8 * 128
constexpr float kRowBias = 263168.0f;

// Row transform straight from the magic-biased bytes m_j = 2^15 + p_j.  Every value of the first two butterfly stages
// is an integer below 2^19, so those stages are exact whatever the bias; it cancels in every difference and survives
// only in output 0 = 8 * 2^15 + sum p_j, from which kRowBias takes it out together with the level shift:
//   x[0] = sum (p_j - 128)                 (src/dct.c:115's (double)px - 128.0, summed)
//   d = m_a - m_b = p_a - p_b,  s_a + s_b, s_a - s_b: as without the bias
// One addition per sum instead of two (the bias used to be removed sum by sum).
__device__ __forceinline__ void fdct8_row_from_bytes(float *x, uint2 raw)
{
    const float m0 = byte_to_magic<0>(raw.x), m1 = byte_to_magic<1>(raw.x), m2 = byte_to_magic<2>(raw.x),
                m3 = byte_to_magic<3>(raw.x), m4 = byte_to_magic<0>(raw.y), m5 = byte_to_magic<1>(raw.y),
                m6 = byte_to_magic<2>(raw.y), m7 = byte_to_magic<3>(raw.y);
    const float s07 = __fadd_rn(m0, m7), d07 = __fsub_rn(m0, m7);
    const float s16 = __fadd_rn(m1, m6), d16 = __fsub_rn(m1, m6);
    const float s25 = __fadd_rn(m2, m5), d25 = __fsub_rn(m2, m5);
    const float s34 = __fadd_rn(m3, m4), d34 = __fsub_rn(m3, m4);
    fdct8_tail<float, 1>(x, s07, s16, s25, s34, d07, d16, d25, d34);
    x[0] = __fsub_rn(x[0], kRowBias);
}

// float pixel -> centred sample on the fast path (the reference's (double)p - 128.0 is exact; this rounds once)
__device__ __forceinline__ float centre_float_pixel(float p) { return __fsub_rn(p, 128.0f); }

// The same row transform for TWO rows at once, one per lane of the packed operations: x[k] = (row A, row B).
__device__ __forceinline__ void fdct8_rowpair_from_bytes(float2 *x, uint2 a, uint2 b)
{
    using O = Ops<float2>;
    const float2 m0 = make_float2(byte_to_magic<0>(a.x), byte_to_magic<0>(b.x)), m1 = make_float2(byte_to_magic<1>(a.x), byte_to_magic<1>(b.x)),
                 m2 = make_float2(byte_to_magic<2>(a.x), byte_to_magic<2>(b.x)), m3 = make_float2(byte_to_magic<3>(a.x), byte_to_magic<3>(b.x)),
                 m4 = make_float2(byte_to_magic<0>(a.y), byte_to_magic<0>(b.y)), m5 = make_float2(byte_to_magic<1>(a.y), byte_to_magic<1>(b.y)),
                 m6 = make_float2(byte_to_magic<2>(a.y), byte_to_magic<2>(b.y)), m7 = make_float2(byte_to_magic<3>(a.y), byte_to_magic<3>(b.y));
    const float2 s07 = O::add(m0, m7), d07 = O::sub(m0, m7);
    const float2 s16 = O::add(m1, m6), d16 = O::sub(m1, m6);
    const float2 s25 = O::add(m2, m5), d25 = O::sub(m2, m5);
    const float2 s34 = O::add(m3, m4), d34 = O::sub(m3, m4);
    fdct8_tail<float2, 1>(x, s07, s16, s25, s34, d07, d16, d25, d34);
    x[0] = O::add(x[0], make_float2(-kRowBias, -kRowBias));
}

// packed quantise-and-residual of two coefficients: bit-identical per lane to quant_residual
__device__ __forceinline__ void quant_residual2(float2 c, float2 r, float2 &t, float2 &e)
{
    using O = Ops<float2>;
    const float2 magic = make_float2(kMagic, kMagic);
    t = O::fma(c, r, magic);
    e = O::fma(c, r, O::sub(magic, t));     // magic - t == -(t - magic) exactly
}

// sum and sum of squares of the centred samples of one 8-pixel row, exact (packed byte dot products)
__device__ __forceinline__ void row_moments(uint2 raw, int &isum, int &isq)
{
    const uint32_t wa = raw.x ^ 0x80808080u, wb = raw.y ^ 0x80808080u;   // p - 128 as int8
    isum = __dp4a((int)wa, 0x01010101, isum);
    isum = __dp4a((int)wb, 0x01010101, isum);
    isq = __dp4a((int)wa, (int)wa, isq);
    isq = __dp4a((int)wb, (int)wb, isq);
}

// fast-path 1/(2 - nv) from num = 4096 * variance (src/quantization.c:186-190 in fp32; exact in K3)
__device__ __forceinline__ float adaptive_inv_scale(int num)
{
    const float nv = fminf(1.0f, fmaxf(0.1f, __fmul_rn((float)num, 1.0f / 4096000.0f)));
    return __frcp_rn(__fsub_rn(2.0f, nv));
}

// fast-path (2 - nv) for the decoder, from the variance side information
__device__ __forceinline__ float adaptive_scale(double var)
{
    const float nv = fminf(1.0f, fmaxf(0.1f, __fmul_rn((float)var, 1.0f / 1000.0f)));
    return __fsub_rn(2.0f, nv);
}

// K2, adaptive plans: bound >= sum_k gain_k |q_k rs_k s| from the sum over the AC entries taken with the unscaled table
// (`ac`), the DC entry's own term (`dc`) and the block's factor s = 2 - nv; 1.00001 covers the rounding of rs_k * s
__device__ __forceinline__ float adaptive_bound(float ac, float dc, float s)
{
    return __fmaf_rn(ac, __fmul_rn(s, 1.00001f), dc);
}

// quantise one coefficient: t holds round(c*r) in its low mantissa bits, e = c*r - round(c*r)
__device__ __forceinline__ void quant_residual(float c, float r, float &t, float &e)
{
    t = __fmaf_rn(c, r, kMagic);
    e = __fmaf_rn(c, r, -__fsub_rn(t, kMagic));
}

// one int16 half of a packed word -> float, exactly: PRMT with sign replication widens it to int32,
// then the full-width conversion (I2FP.F32.S32 runs at 64 lanes/clk/SM; the 16-bit form I2F.S16 at 16)
template <int hi> __device__ __forceinline__ float half_to_float(uint32_t w)
{
    // PTX prmt (not __byte_perm, which ignores bit 3 of a selector nibble): nibble 8|n replicates the
    // sign bit of byte n, so {b0, b1, sign(b1), sign(b1)} is the sign-extended low half
    uint32_t x;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(x) : "r"(w), "r"(0u), "r"(hi ? 0xBB32u : 0x9910u));
    float f;
    asm("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(x));
    return f;
}

// pixel from an inverse-transform sample: t holds round(x) + 128 in its low 16 mantissa bits, e = x - round(x)
__device__ __forceinline__ void pixel_residual(float x, float &t, float &e)
{
    t = __fadd_rn(x, kMagic128);
    e = __fsub_rn(x, __fsub_rn(t, kMagic128));
}

// packed pixel-and-residual of two samples: bit-identical per lane to pixel_residual
__device__ __forceinline__ void pixel_residual2(float2 x, float2 &t, float2 &e)
{
    using O = Ops<float2>;
    const float2 magic = make_float2(kMagic128, kMagic128);
    t = O::add(x, magic);
    e = O::add(x, O::sub(magic, t));        // x - (t - magic)
}

// K2's threshold from the accumulated bound: |fp32 pixel - exact pixel| <= 2^-24 * bound (derive_bands.py);
// 1.0625 covers the rounding of the bound's own accumulation (any summation order)
__device__ __forceinline__ float pixel_threshold(float bound, float band_floor)
{
    return __fsub_rn(0.5f, __fmaf_rn(bound, 5.9604645e-8f * 1.0625f, band_floor));
}

}  // namespace dctb
