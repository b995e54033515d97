// rle.cu -- K5: run-length symbols of the quantised records, on the device.
//
// First of the "next" rows: the untouched host consumer's run_length_encode
// (src/entropy.c:216-256) walks a block in zigzag order and emits one RLESymbol
// {value, run_length} (include/entropy.h:35-38) per non-zero coefficient -- run_length = the zeros
// skipped since the previous symbol -- plus a closing symbol at position 63 (whose run also
// counts a zero last coefficient).  These kernels produce exactly those lists for every record
// of a plane: symbol counts -> exclusive prefix sums (offsets) -> symbols, compacted.
// Integer work only; parity is bit-exact against run_length_encode.
#include "butterfly.cuh"
#include "kernels.cuh"

namespace dctb {

namespace {

constexpr int kRleThreads = 256;

__device__ __forceinline__ void load_record(const int16_t *coef, uint32_t b, uint32_t (&w)[32])
{
    const uint4 *src = reinterpret_cast<const uint4 *>(coef + (size_t)b * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint4 t = __ldg(src + j);
        w[4 * j] = t.x, w[4 * j + 1] = t.y, w[4 * j + 2] = t.z, w[4 * j + 3] = t.w;
    }
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t &cta_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int i = 0; i < kRleThreads / 32; ++i) {
        const uint32_t s = warp_sums[i];
        if (i < warp) base += s;
        total += s;
    }
    cta_total = total;
    __syncthreads();
    return base + incl - v;
}

// symbols of a record = non-zero coefficients among the first 63 zigzag positions + 1.  Zigzag
// position 63 is natural index 63 in both layouts, so the count does not depend on the layout.
__global__ void __launch_bounds__(kRleThreads) k_rle_count(const int16_t *coef, uint32_t nblocks, uint32_t *offsets,
                                                            uint32_t *cta_sums)
{
    __shared__ uint32_t warp_sums[kRleThreads / 32];
    const uint32_t b = blockIdx.x * kRleThreads + threadIdx.x;
    uint32_t count = 0;
    if (b < nblocks) {
        uint32_t w[32];
        load_record(coef, b, w);
        count = 1;
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            count += (w[m] & 0xFFFFu) != 0;
            if (m != 31) count += (w[m] >> 16) != 0;
        }
    }
    uint32_t total;
    const uint32_t excl = block_exclusive_scan(count, warp_sums, total);
    if (b < nblocks) offsets[b] = excl;               // CTA-local for now
    if (threadIdx.x == 0) cta_sums[blockIdx.x] = total;
}

// exclusive scan of the per-CTA totals, one CTA, chunk by chunk; grand total to *total_out
__global__ void __launch_bounds__(kRleThreads) k_rle_scan_sums(uint32_t *cta_sums, uint32_t n, unsigned long long *total_out)
{
    __shared__ uint32_t warp_sums[kRleThreads / 32];
    unsigned long long running = 0;
    for (uint32_t base = 0; base < n; base += kRleThreads) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? cta_sums[i] : 0;
        uint32_t chunk_total;
        const uint32_t excl = block_exclusive_scan(v, warp_sums, chunk_total);
        if (i < n) cta_sums[i] = (uint32_t)running + excl;
        running += chunk_total;
    }
    if (threadIdx.x == 0) *total_out = running;
}

__global__ void __launch_bounds__(kRleThreads) k_rle_add_base(uint32_t *offsets, uint32_t nblocks, const uint32_t *cta_sums,
                                                               const unsigned long long *total)
{
    const uint32_t b = blockIdx.x * kRleThreads + threadIdx.x;
    if (b < nblocks) offsets[b] += cta_sums[blockIdx.x];
    if (b == 0) offsets[nblocks] = (uint32_t)*total;
}

template <int LAYOUT>
__global__ void __launch_bounds__(kRleThreads) k_rle_emit(const int16_t *coef, uint32_t nblocks, const uint32_t *offsets,
                                                           int2 *symbols)
{
    const uint32_t b = blockIdx.x * kRleThreads + threadIdx.x;
    if (b >= nblocks) return;
    uint32_t w[32];
    load_record(coef, b, w);
    int2 *out = symbols + offsets[b];
    int run = 0;
    static_for<0, 64>([&](auto P) {
        constexpr int p = decltype(P)::value;                 // zigzag position
        constexpr int s = LAYOUT == LAYOUT_ZIGZAG ? p : kZigZag.nat[p];   // where the record stores it
        const int v = (int)(int16_t)(s & 1 ? (w[s >> 1] >> 16) : (w[s >> 1] & 0xFFFFu));
        if (p == 63) {
            *out = make_int2(v, v == 0 ? run + 1 : run);      // closing symbol (src/entropy.c:226-238)
        } else if (v != 0) {
            *out++ = make_int2(v, run);
            run = 0;
        } else {
            ++run;
        }
    });
}

}  // namespace

cudaError_t launch_rle_count(const int16_t *d_coef, uint32_t nblocks, uint32_t *d_offsets, uint32_t *d_cta_sums,
                             unsigned long long *d_total, cudaStream_t s)
{
    const uint32_t ctas = (nblocks + kRleThreads - 1) / kRleThreads;
    k_rle_count<<<ctas, kRleThreads, 0, s>>>(d_coef, nblocks, d_offsets, d_cta_sums);
    k_rle_scan_sums<<<1, kRleThreads, 0, s>>>(d_cta_sums, ctas, d_total);
    k_rle_add_base<<<ctas, kRleThreads, 0, s>>>(d_offsets, nblocks, d_cta_sums, d_total);
    return cudaGetLastError();
}

cudaError_t launch_rle_emit(const int16_t *d_coef, uint32_t nblocks, int layout, const uint32_t *d_offsets, void *d_symbols,
                            cudaStream_t s)
{
    const uint32_t ctas = (nblocks + kRleThreads - 1) / kRleThreads;
    if (layout == LAYOUT_ZIGZAG) k_rle_emit<LAYOUT_ZIGZAG><<<ctas, kRleThreads, 0, s>>>(d_coef, nblocks, d_offsets, (int2 *)d_symbols);
    else k_rle_emit<LAYOUT_NATURAL><<<ctas, kRleThreads, 0, s>>>(d_coef, nblocks, d_offsets, (int2 *)d_symbols);
    return cudaGetLastError();
}

}  // namespace dctb
