// butterfly.cuh -- scaled 8-point forward / inverse DCT flowgraphs, 30 operations each.
//
// Replaces the inner triple loops of the reference's dct_forward (src/dct.c:57-74) and
// dct_inverse (src/dct.c:85-102) on the fast path.  The arithmetic is NOT the reference's
// (that is replayed bit-for-bit in replay_f64.cu); it is a Arai-Agui-Nakajima style
// factorisation whose per-coefficient output scale is folded into the quantisation
// multiplier.  tools/derive_bands.py executes this exact sequence of operations on
// (functional, error-bound) pairs to derive the band inside which a result is re-done in
// fp64 -- keep the two in step, operation for operation.
//
// Every operation is an explicit round-to-nearest intrinsic so that nvcc can neither
// contract nor re-associate: the error analysis is of this code, not of what a compiler
// might make of it.
#pragma once
#include <cuda_runtime.h>

namespace dctb {

template <typename T> struct Ops;

template <> struct Ops<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float c) { return __fmul_rn(a, c); }
    static __device__ __forceinline__ float fma(float a, float c, float b) { return __fmaf_rn(a, c, b); }
};

template <> struct Ops<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double c) { return __dmul_rn(a, c); }
    static __device__ __forceinline__ double fma(double a, double c, double b) { return __fma_rn(a, c, b); }
};

// Two independent fp32 lanes per instruction (sm_100 FADD2 / FMUL2 / FFMA2): each lane is an ordinary
// IEEE round-to-nearest operation, so a packed butterfly is bit-identical, lane by lane, to the scalar
// one -- it just needs half the issue slots.  Constants are broadcast immediates.
template <> struct Ops<float2> {
    static __device__ __forceinline__ float2 add(float2 a, float2 b)
    {
        float2 r;
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long &>(r)) : "l"(reinterpret_cast<unsigned long long &>(a)), "l"(reinterpret_cast<unsigned long long &>(b)));
        return r;
    }
    static __device__ __forceinline__ float2 sub(float2 a, float2 b)
    {
        float2 r;
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long &>(r)) : "l"(reinterpret_cast<unsigned long long &>(a)), "l"(reinterpret_cast<unsigned long long &>(b)));
        return r;
    }
    static __device__ __forceinline__ float2 mul(float2 a, float2 c)
    {
        float2 r;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long &>(r)) : "l"(reinterpret_cast<unsigned long long &>(a)), "l"(reinterpret_cast<unsigned long long &>(c)));
        return r;
    }
    static __device__ __forceinline__ float2 fma(float2 a, float2 c, float2 b)
    {
        float2 r;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long &>(r)) : "l"(reinterpret_cast<unsigned long long &>(a)), "l"(reinterpret_cast<unsigned long long &>(c)), "l"(reinterpret_cast<unsigned long long &>(b)));
        return r;
    }
};

// a butterfly constant in the arithmetic type: rounded once to fp32 (as numpy.float32(c) in the model),
// broadcast to both lanes of a packed pair
template <typename T> __device__ __forceinline__ T cst(double c);
template <> __device__ __forceinline__ float cst<float>(double c) { return (float)c; }
template <> __device__ __forceinline__ double cst<double>(double c) { return c; }
template <> __device__ __forceinline__ float2 cst<float2>(double c) { return make_float2((float)c, (float)c); }

// exact reals; cst<T>(...) rounds them once, exactly as numpy.float32(c) does in the model
#define DCTB_C4    0.70710678118654752440
#define DCTB_C382  0.38268343236508977173
#define DCTB_C541  0.54119610014619698440
#define DCTB_C1306 1.30656296487637652786
#define DCTB_SQRT2 1.41421356237309504880
#define DCTB_C1847 1.84775906502257351226
#define DCTB_C1082 1.08239220029239396880
#define DCTB_C2613 2.61312592975275305571

// Everything after the first butterfly stage.  Inputs: s_ab = x_a + x_b, d_ab = x_a - x_b.
// out[k*S] = a_k * sqrt(8) * (orthonormal DCT-II)[k],  a_0 = 1, a_k = sqrt(2) cos(k pi / 16).
template <typename T, int S>
__device__ __forceinline__ void fdct8_tail(T *x, T s07, T s16, T s25, T s34, T d07, T d16, T d25, T d34)
{
    using O = Ops<T>;
    const T e0 = O::add(s07, s34), e3 = O::sub(s07, s34);
    const T e1 = O::add(s16, s25), e2 = O::sub(s16, s25);
    x[0 * S] = O::add(e0, e1);
    x[4 * S] = O::sub(e0, e1);
    const T t = O::add(e2, e3);
    x[2 * S] = O::fma(t, cst<T>(DCTB_C4), e3);
    x[6 * S] = O::fma(t, cst<T>(-DCTB_C4), e3);
    const T a = O::add(d34, d25), b = O::add(d25, d16), c = O::add(d16, d07);
    const T z5 = O::mul(O::sub(a, c), cst<T>(DCTB_C382));
    const T z2 = O::fma(a, cst<T>(DCTB_C541), z5), z4 = O::fma(c, cst<T>(DCTB_C1306), z5);
    const T z11 = O::fma(b, cst<T>(DCTB_C4), d07), z13 = O::fma(b, cst<T>(-DCTB_C4), d07);
    x[5 * S] = O::add(z13, z2);
    x[3 * S] = O::sub(z13, z2);
    x[1 * S] = O::add(z11, z4);
    x[7 * S] = O::sub(z11, z4);
}

// x[k*S], k = 0..7, transformed in place (30 operations).
template <typename T, int S>
__device__ __forceinline__ void fdct8(T *x)
{
    using O = Ops<T>;
    const T s07 = O::add(x[0 * S], x[7 * S]), d07 = O::sub(x[0 * S], x[7 * S]);
    const T s16 = O::add(x[1 * S], x[6 * S]), d16 = O::sub(x[1 * S], x[6 * S]);
    const T s25 = O::add(x[2 * S], x[5 * S]), d25 = O::sub(x[2 * S], x[5 * S]);
    const T s34 = O::add(x[3 * S], x[4 * S]), d34 = O::sub(x[3 * S], x[4 * S]);
    fdct8_tail<T, S>(x, s07, s16, s25, s34, d07, d16, d25, d34);
}

// Everything after the first stage of the inverse flowgraph.  Inputs (from v[k], pre-multiplied by
// a_k / sqrt(8)):  t10 = v0+v4, t11 = v0-v4, t13 = v2+v6, d26 = v2-v6, z13 = v5+v3, z10 = v5-v3,
// z11 = v1+v7, z12 = v1-v7.  Writes the 8 spatial samples to out[k*S].
template <typename T, int S>
__device__ __forceinline__ void idct8_tail(T *out, T t10, T t11, T t13, T d26, T z13, T z10, T z11, T z12)
{
    using O = Ops<T>;
    // t12n = -(d26*sqrt2 - t13): the negated form needs no operand negation; round-to-nearest is symmetric,
    // so t11 - t12n == t11 + t12 bit for bit (the model's fms(d, sqrt2, t13))
    const T t12n = O::fma(d26, cst<T>(-DCTB_SQRT2), t13);
    const T e0 = O::add(t10, t13), e3 = O::sub(t10, t13);
    const T e1 = O::sub(t11, t12n), e2 = O::add(t11, t12n);
    const T t7 = O::add(z11, z13);
    const T zd = O::sub(z11, z13);
    const T z5 = O::mul(O::add(z10, z12), cst<T>(DCTB_C1847));
    const T t10o = O::fma(z12, cst<T>(-DCTB_C1082), z5);
    const T t12o = O::fma(z10, cst<T>(-DCTB_C2613), z5);
    const T t6 = O::sub(t12o, t7);
    const T t5n = O::fma(zd, cst<T>(-DCTB_SQRT2), t6);   // -(zd*sqrt2 - t6), same remark
    const T t4 = O::add(t10o, t5n);
    out[0 * S] = O::add(e0, t7);
    out[1 * S] = O::add(e1, t6);
    out[2 * S] = O::sub(e2, t5n);
    out[3 * S] = O::add(e3, t4);
    out[4 * S] = O::sub(e3, t4);
    out[5 * S] = O::add(e2, t5n);
    out[6 * S] = O::sub(e1, t6);
    out[7 * S] = O::sub(e0, t7);
}

// v[k*S] pre-multiplied by a_k / sqrt(8); transformed in place to the 8 spatial samples (30 operations).
template <typename T, int S>
__device__ __forceinline__ void idct8(T *v)
{
    using O = Ops<T>;
    const T t10 = O::add(v[0 * S], v[4 * S]), t11 = O::sub(v[0 * S], v[4 * S]);
    const T t13 = O::add(v[2 * S], v[6 * S]), d26 = O::sub(v[2 * S], v[6 * S]);
    const T z13 = O::add(v[5 * S], v[3 * S]), z10 = O::sub(v[5 * S], v[3 * S]);
    const T z11 = O::add(v[1 * S], v[7 * S]), z12 = O::sub(v[1 * S], v[7 * S]);
    idct8_tail<T, S>(v, t10, t11, t13, d26, z13, z10, z11, z12);
}

// idct8 with the dequantisation folded into its first stage.  q[k*S] are the quantised values (converted, not yet
// multiplied); ma[j] multiplies q[a_j] for a = (0, 2, 5, 1) and mb[j] = (m, -m) multiplies q[b_j] for
// b = (4, 6, 3, 7).  v_a = q_a * m_a is rounded as before; v_a +- q_b * m_b is ONE fused operation, i.e. one
// rounding fewer than idct8 on pre-multiplied inputs, so idct8's error bound covers it.  12 operations
// instead of 8 multiplications + 8 additions.  MB is a pair {T pos, T neg}.
template <typename T, int S, typename MB>
__device__ __forceinline__ void idct8_dequant(T *q, const T (&ma)[4], const MB (&mb)[4])
{
    using O = Ops<T>;
    const T p0 = O::mul(q[0 * S], ma[0]), p2 = O::mul(q[2 * S], ma[1]);
    const T p5 = O::mul(q[5 * S], ma[2]), p1 = O::mul(q[1 * S], ma[3]);
    const T t10 = O::fma(q[4 * S], mb[0].pos, p0), t11 = O::fma(q[4 * S], mb[0].neg, p0);
    const T t13 = O::fma(q[6 * S], mb[1].pos, p2), d26 = O::fma(q[6 * S], mb[1].neg, p2);
    const T z13 = O::fma(q[3 * S], mb[2].pos, p5), z10 = O::fma(q[3 * S], mb[2].neg, p5);
    const T z11 = O::fma(q[7 * S], mb[3].pos, p1), z12 = O::fma(q[7 * S], mb[3].neg, p1);
    idct8_tail<T, S>(q, t10, t11, t13, d26, z13, z10, z11, z12);
}

// zigzag position -> natural index (8i+j); the scan of src/entropy.c:158-178 for N = 8
// (checked against the reference's block_to_zigzag in tests/test_host_logic.py).
struct ZigZag {
    int nat[64];
    constexpr ZigZag() : nat{}
    {
        int idx = 0;
        for (int s = 0; s <= 14; ++s) {
            if (s % 2 == 0) {
                for (int i = (s < 8) ? s : 7; i >= 0 && (s - i) < 8; --i) nat[idx++] = i * 8 + (s - i);
            } else {
                for (int i = (s < 8) ? 0 : s - 7; i < 8 && (s - i) >= 0; ++i) nat[idx++] = i * 8 + (s - i);
            }
        }
    }
};
constexpr ZigZag kZigZag{};

template <int LAYOUT> __host__ __device__ constexpr int storage_to_natural(int p)
{
    return LAYOUT == 1 ? kZigZag.nat[p] : p;
}

}  // namespace dctb
