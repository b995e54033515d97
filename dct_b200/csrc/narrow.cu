// narrow.cu -- int16 records <-> int8 records for the PCIe-bound host-plane calls.
//
// The reference keeps quantised coefficients in `int` (include/quantization.h:69); libdct_cuda stores
// int16 because |c| <= 1024 and Q >= 1.  When every entry of the plan's table is
// >= 1024 / 127.5 the same argument gives |q| <= 127, so a record also fits 64 BYTES -- and the
// host-plane calls, which are bound by the PCIe link (2 of their 3 bytes per pixel are records),
// can move half as much.  These two streaming kernels do the conversion next to K1 / K2 on the
// strip that is already on the device; the fused kernels and their int16 records stay as they are.
// Narrowing saturates and counts what it had to clamp (nothing, for 8-bit pixels and a table that passed
// the plan's check); widening is exact.
#include "kernels.cuh"

namespace dctb {
namespace {

__device__ __forceinline__ uint32_t clamp_s16x2(uint32_t w)
{
    uint32_t r;
    asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(w), "r"(0xff80ff80u));   // >= -128
    asm("min.s16x2 %0, %0, %1;" : "+r"(r) : "r"(0x007f007fu));           // <=  127
    return r;
}

// one thread: 16 values, 2 x LDG.128 -> 1 x STG.128 (a warp stores 512 contiguous bytes)
__global__ void __launch_bounds__(256) k_narrow_records(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n16,
                                                        Counters *ctr)
{
    unsigned clipped = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 a = __ldcs(in + 2 * i), b = __ldcs(in + 2 * i + 1);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t c[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            c[k] = clamp_s16x2(w[k]);
            clipped += __popc(__vcmpne2(c[k], w[k])) >> 4;      // 0xffff per differing half
        }
        uint4 o;
        o.x = __byte_perm(c[0], c[1], 0x6420), o.y = __byte_perm(c[2], c[3], 0x6420);
        o.z = __byte_perm(c[4], c[5], 0x6420), o.w = __byte_perm(c[6], c[7], 0x6420);
        __stcs(out + i, o);
    }
    if (clipped) atomicAdd(&ctr->saturated, (unsigned long long)clipped);
}

// one thread: 16 values, 1 x LDG.128 -> 2 x STG.128
__global__ void __launch_bounds__(256) k_widen_records(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n16)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(in + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // prmt with bit 3 of a selector nibble set replicates the sign of the selected byte
            asm("prmt.b32 %0, %1, 0, 0x9180;" : "=r"(o[2 * k]) : "r"(w[k]));       // bytes 0, 1 -> two int16
            asm("prmt.b32 %0, %1, 0, 0xb3a2;" : "=r"(o[2 * k + 1]) : "r"(w[k]));   // bytes 2, 3
        }
        __stcs(out + 2 * i, make_uint4(o[0], o[1], o[2], o[3]));
        __stcs(out + 2 * i + 1, make_uint4(o[4], o[5], o[6], o[7]));
    }
}

unsigned stream_grid(size_t n16)
{
    const size_t blocks = (n16 + 255) / 256;
    return (unsigned)(blocks < 148 * 8 ? (blocks ? blocks : 1) : 148 * 8);
}

}  // namespace

// n = number of coefficients, a multiple of 64 (whole records), so always a multiple of 16
cudaError_t launch_narrow_records(const int16_t *d_in, int8_t *d_out, size_t n, Counters *ctr, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    k_narrow_records<<<stream_grid(n / 16), 256, 0, s>>>(reinterpret_cast<const uint4 *>(d_in), reinterpret_cast<uint4 *>(d_out),
                                                         n / 16, ctr);
    return cudaGetLastError();
}

cudaError_t launch_widen_records(const int8_t *d_in, int16_t *d_out, size_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    k_widen_records<<<stream_grid(n / 16), 256, 0, s>>>(reinterpret_cast<const uint4 *>(d_in), reinterpret_cast<uint4 *>(d_out), n / 16);
    return cudaGetLastError();
}

}  // namespace dctb
