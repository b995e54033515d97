// fwd_quant.cu -- K1: u8 pixels -> level shift -> 8x8 forward DCT -> quantise -> int16 records.
//
// Fuses, for every 8x8 block of a plane, the reference's
//   create_block_from_pixels (src/dct.c:109-120), dct_forward (src/dct.c:52-77),
//   quantize (src/quantization.c:113-131; adaptive: calculate_block_variance :153-169 and
//   adjust_matrix_for_block :171-211) and, for the ZIGZAG layout, block_to_zigzag
//   (src/entropy.c:158-178)
// into one pass: each block is read once (64 B) and written once (128 B).
//
// Mapping: one thread owns one block; a warp owns 32 consecutive records, i.e. a 256-pixel
// wide, 8-row tile.  A lane reads its 8-byte block rows with 8 independent LDG.64 (a warp
// instruction covers 256 contiguous bytes), keeps the block in 64 registers through both
// butterfly passes (no transpose, no shuffles), quantises with one FFMA per coefficient
// (round-to-nearest via the 1.5*2^23 trick), checks the fp32 residual against the
// per-coefficient band, and stores through a padded per-warp shared-memory stage so that
// every STG.128 of the warp covers 512 contiguous bytes of the record array.
// Blocks with a coefficient inside the band go to the worklist and are replayed in fp64 (K3).
#include "butterfly.cuh"
#include "kernels.cuh"

namespace dctb {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStageWordsPerBlock = 36;   // 128 B record + 16 B pad: conflict-free STS.128 / LDS.128
constexpr int kStageWordsPerWarp = 32 * kStageWordsPerBlock;

constexpr float kMagic = 12582912.0f;     // 1.5 * 2^23: x + kMagic rounds x to an integer (RNE)

__device__ __forceinline__ uint2 ldg_stream_u2(const uint8_t *p)
{
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ void stg_stream_u4(void *p, const uint4 &v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

// byte `idx` of w -> 2^23 + byte as a float: (0x4B000000 | byte), one PRMT, no conversion instruction
template <int idx> __device__ __forceinline__ float byte_to_magic(uint32_t w)
{
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 + idx));
}

// Row transform straight from the magic-biased bytes m_j = 2^23 + p_j.  The level shift and the
// removal of the bias are folded into the first butterfly stage, every step exact:
//   d = m_a - m_b                    = p_a - p_b
//   s = (m_a - (2^24 + 256)) + m_b   = (p_a - 128) + (p_b - 128)      (|m_a - 2^24 - 256| < 2^24: exact)
// i.e. the same integers the reference forms as (double)px - 128.0 (src/dct.c:115) summed pairwise.
__device__ __forceinline__ void fdct8_row_from_bytes(float *x, uint2 raw)
{
    constexpr float kBias2 = 16777472.0f;   // 2 * (2^23 + 128)
    const float m0 = byte_to_magic<0>(raw.x), m1 = byte_to_magic<1>(raw.x), m2 = byte_to_magic<2>(raw.x),
                m3 = byte_to_magic<3>(raw.x), m4 = byte_to_magic<0>(raw.y), m5 = byte_to_magic<1>(raw.y),
                m6 = byte_to_magic<2>(raw.y), m7 = byte_to_magic<3>(raw.y);
    const float s07 = __fadd_rn(__fadd_rn(m0, -kBias2), m7), d07 = __fsub_rn(m0, m7);
    const float s16 = __fadd_rn(__fadd_rn(m1, -kBias2), m6), d16 = __fsub_rn(m1, m6);
    const float s25 = __fadd_rn(__fadd_rn(m2, -kBias2), m5), d25 = __fsub_rn(m2, m5);
    const float s34 = __fadd_rn(__fadd_rn(m3, -kBias2), m4), d34 = __fsub_rn(m3, m4);
    fdct8_tail<float, 1>(x, s07, s16, s25, s34, d07, d16, d25, d34);
}

// UNIFORM: one band for all 64 coefficients (the widest), tested with 3-input max -- half the
// instructions of the per-coefficient compare; chosen by the host when the widest band is small.
template <int LAYOUT, bool ADAPTIVE, bool UNIFORM>
__global__ void __launch_bounds__(kThreads) k_fwd_quant_u8(const __grid_constant__ FwdParams p)
{
    __shared__ uint4 stage[kWarps * kStageWordsPerWarp / 4];

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warp_base = (blockIdx.x * kThreads + warp * 32);   // first record of this warp
    const uint32_t b = warp_base + lane;
    const bool valid = b < p.nblocks;
    const uint32_t bb = valid ? b : p.nblocks - 1;
    const uint32_t by = bb / p.bw, bx = bb - by * p.bw;
    const uint8_t *src = p.px + (long long)by * 8 * p.pitch + (long long)bx * 8;

    uint2 raw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) raw[i] = ldg_stream_u2(src + i * p.pitch);

    float v[64];
    float inv_s = 1.0f;
    if constexpr (ADAPTIVE) {
        // Sum and sum of squares of the centred samples, as exact integers: packed byte dot products.
        // 4096 * variance = 64 * sum(x^2) - sum(x)^2 is exact, so the side array equals
        // calculate_block_variance (src/quantization.c:153-169) bit for bit.
        int isum = 0, isq = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t wa = raw[i].x ^ 0x80808080u, wb = raw[i].y ^ 0x80808080u;   // p - 128 as int8
            isum = __dp4a((int)wa, 0x01010101, isum);
            isum = __dp4a((int)wb, 0x01010101, isum);
            isq = __dp4a((int)wa, (int)wa, isq);
            isq = __dp4a((int)wb, (int)wb, isq);
        }
        const int num = 64 * isq - isum * isum;   // 4096 * variance, exact (< 2^27)
        if (p.var_out != nullptr && valid) p.var_out[b] = (double)num * (1.0 / 4096.0);
        // s = 2 - clamp(var/1000, 0.1, 1)  (src/quantization.c:186-190), fp32 here, exact in K3
        const float nv = fminf(1.0f, fmaxf(0.1f, __fmul_rn((float)num, 1.0f / 4096000.0f)));
        inv_s = __frcp_rn(__fsub_rn(2.0f, nv));
    }

    // rows (X * D^T), then columns (D * temp): same order as src/dct.c:57-74
#pragma unroll
    for (int i = 0; i < 8; ++i) fdct8_row_from_bytes(&v[8 * i], raw[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) fdct8<float, 8>(&v[j]);

    // quantise: t = c*r + 1.5*2^23 holds round(c*r) in its low mantissa bits;
    // residual e = c*r - round(c*r) (one rounding); |e| >= 0.5 - band  => replay
    uint32_t w[32];
    bool flag = false;
    float emax = 0.0f;
    static_for<0, 32>([&](auto M) {
        constexpr int m = decltype(M)::value;
        constexpr int k0 = storage_to_natural<LAYOUT>(2 * m), k1 = storage_to_natural<LAYOUT>(2 * m + 1);
        float r0 = p.r[k0], r1 = p.r[k1];
        if constexpr (ADAPTIVE) {
            if (k0 != 0) r0 = __fmul_rn(r0, inv_s);   // DC keeps the unscaled table entry
            r1 = __fmul_rn(r1, inv_s);
        }
        const float t0 = __fmaf_rn(v[k0], r0, kMagic), t1 = __fmaf_rn(v[k1], r1, kMagic);
        const float e0 = __fmaf_rn(v[k0], r0, -__fsub_rn(t0, kMagic));
        const float e1 = __fmaf_rn(v[k1], r1, -__fsub_rn(t1, kMagic));
        if constexpr (UNIFORM) emax = fmaxf(fmaxf(emax, fabsf(e0)), fabsf(e1));   // FMNMX3
        else flag |= (fabsf(e0) >= p.thr[k0]) | (fabsf(e1) >= p.thr[k1]);
        w[m] = __byte_perm(__float_as_uint(t0), __float_as_uint(t1), 0x5410);   // lo16(t0) | lo16(t1) << 16
    });

    if constexpr (UNIFORM) flag = emax >= p.thr_min;

    // stage: lane-major padded records in shared memory, then 512-byte contiguous warp stores
    uint32_t *wstage = reinterpret_cast<uint32_t *>(stage) + warp * kStageWordsPerWarp;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4 *>(wstage + lane * kStageWordsPerBlock + 4 * j) =
            make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
    __syncwarp();
    uint4 *dst = reinterpret_cast<uint4 *>(p.coef) + (size_t)warp_base * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t c = j * 32 + lane;             // 16-byte chunk of the warp's 4 KB
        const uint32_t blk = c >> 3, part = c & 7;
        const uint4 val = *reinterpret_cast<const uint4 *>(wstage + blk * kStageWordsPerBlock + 4 * part);
        if (warp_base + blk < p.nblocks) stg_stream_u4(dst + c, val);
    }

    // warp-aggregated append to the replay worklist
    const unsigned ballot = __ballot_sync(0xffffffffu, flag && valid);
    if (ballot != 0) {
        const int leader = __ffs(ballot) - 1;
        unsigned base = 0;
        if ((int)lane == leader) base = atomicAdd(&p.ctr->wl_count, (unsigned)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (flag && valid) {
            const unsigned pos = base + __popc(ballot & ((1u << lane) - 1u));
            if (pos < p.wl_cap) p.worklist[pos] = b;
        }
    }
}

}  // namespace

template <int LAYOUT, bool ADAPTIVE> static void launch_k1(const FwdParams &p, unsigned grid, cudaStream_t s)
{
    if (p.uniform_band) k_fwd_quant_u8<LAYOUT, ADAPTIVE, true><<<grid, kThreads, 0, s>>>(p);
    else k_fwd_quant_u8<LAYOUT, ADAPTIVE, false><<<grid, kThreads, 0, s>>>(p);
}

cudaError_t launch_fwd_quant_u8(const FwdParams &p, int layout, int adaptive, cudaStream_t s)
{
    if (p.nblocks == 0) return cudaSuccess;
    const unsigned grid = (p.nblocks + kThreads - 1) / kThreads;
    if (layout == LAYOUT_ZIGZAG) {
        if (adaptive) launch_k1<LAYOUT_ZIGZAG, true>(p, grid, s);
        else          launch_k1<LAYOUT_ZIGZAG, false>(p, grid, s);
    } else {
        if (adaptive) launch_k1<LAYOUT_NATURAL, true>(p, grid, s);
        else          launch_k1<LAYOUT_NATURAL, false>(p, grid, s);
    }
    return cudaGetLastError();
}

}  // namespace dctb
