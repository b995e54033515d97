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

// byte `idx` of w -> float(byte) - 128, exactly: (0x4B000000 | byte) is 2^23 + byte
template <int idx> __device__ __forceinline__ float byte_to_centered(uint32_t w)
{
    const uint32_t m = __byte_perm(w, 0x4B000000u, 0x7440 + idx);
    return __fadd_rn(__uint_as_float(m), -8388736.0f);   // -(2^23 + 128)
}

template <int LAYOUT, bool ADAPTIVE>
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
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[8 * i + 0] = byte_to_centered<0>(raw[i].x);
        v[8 * i + 1] = byte_to_centered<1>(raw[i].x);
        v[8 * i + 2] = byte_to_centered<2>(raw[i].x);
        v[8 * i + 3] = byte_to_centered<3>(raw[i].x);
        v[8 * i + 4] = byte_to_centered<0>(raw[i].y);
        v[8 * i + 5] = byte_to_centered<1>(raw[i].y);
        v[8 * i + 6] = byte_to_centered<2>(raw[i].y);
        v[8 * i + 7] = byte_to_centered<3>(raw[i].y);
    }

    float inv_s = 1.0f;
    if constexpr (ADAPTIVE) {
        // sum and sum of squares of the centred samples: integers below 2^24, exact in fp32
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int k = 0; k < 64; k += 2) {
            s0 = __fadd_rn(s0, v[k]);
            s1 = __fadd_rn(s1, v[k + 1]);
            q0 = __fmaf_rn(v[k], v[k], q0);
            q1 = __fmaf_rn(v[k + 1], v[k + 1], q1);
        }
        const int isum = __float2int_rn(__fadd_rn(s0, s1));
        const int isq = __float2int_rn(__fadd_rn(q0, q1));
        const int num = 64 * isq - isum * isum;   // 4096 * variance, exact (< 2^27)
        if (p.var_out != nullptr && valid) p.var_out[b] = (double)num * (1.0 / 4096.0);
        // s = 2 - clamp(var/1000, 0.1, 1)  (src/quantization.c:186-190), fp32 here, exact in K3
        const float nv = fminf(1.0f, fmaxf(0.1f, __fmul_rn((float)num, 1.0f / 4096000.0f)));
        inv_s = __frcp_rn(__fsub_rn(2.0f, nv));
    }

    // rows (X * D^T), then columns (D * temp): same order as src/dct.c:57-74
#pragma unroll
    for (int i = 0; i < 8; ++i) fdct8<float, 1>(&v[8 * i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) fdct8<float, 8>(&v[j]);

    // quantise: t = c*r + 1.5*2^23 holds round(c*r) in its low mantissa bits;
    // residual e = c*r - round(c*r) (one rounding); |e| >= 0.5 - band  => replay
    uint32_t w[32];
    bool flag = false;
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
        flag |= (fabsf(e0) >= p.thr[k0]) | (fabsf(e1) >= p.thr[k1]);
        w[m] = __byte_perm(__float_as_uint(t0), __float_as_uint(t1), 0x5410);   // lo16(t0) | lo16(t1) << 16
    });

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

cudaError_t launch_fwd_quant_u8(const FwdParams &p, int layout, int adaptive, cudaStream_t s)
{
    if (p.nblocks == 0) return cudaSuccess;
    const unsigned grid = (p.nblocks + kThreads - 1) / kThreads;
    if (layout == LAYOUT_ZIGZAG) {
        if (adaptive) k_fwd_quant_u8<LAYOUT_ZIGZAG, true><<<grid, kThreads, 0, s>>>(p);
        else          k_fwd_quant_u8<LAYOUT_ZIGZAG, false><<<grid, kThreads, 0, s>>>(p);
    } else {
        if (adaptive) k_fwd_quant_u8<LAYOUT_NATURAL, true><<<grid, kThreads, 0, s>>>(p);
        else          k_fwd_quant_u8<LAYOUT_NATURAL, false><<<grid, kThreads, 0, s>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace dctb
