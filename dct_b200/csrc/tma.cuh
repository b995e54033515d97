// tma.cuh -- bulk-tensor (TMA) tile movement and mbarrier helpers for the fused kernels.
//
// K1 / K2 move their tiles with cp.async.bulk.tensor (SASS UTMALDG / UTMASTG): one instruction of one
// lane moves a whole 256-pixel x 8-row strip or a 32-record (4 KB) tile between HBM and shared memory,
// instead of 8-16 LDGSTS / LDG / STG instructions per lane plus their address arithmetic.  The record
// tiles use the 128-byte swizzle of the tensor map (16-byte chunk c of record r lives at chunk
// c ^ (r & 7)), so a lane reads or writes its own 128-byte record with conflict-free LDS.128 / STS.128
// and the tile is still one dense box for the copy engine -- no padding, no re-staging pass.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dctb {
namespace tma {

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (the copy engine signals them)
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's generic-proxy shared-memory writes before later async-proxy (bulk copy) reads
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Waits for the phase with the given parity.  try_wait suspends the thread in hardware for a while and returns;
// the loop is bounded so that a copy that never arrives (a bad tensor map) ends in a trap instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spins = 0; spins < (1u << 26); ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

// global -> shared, 2-D box at (c0, c1); completion is counted in bytes on `bar`
__device__ __forceinline__ void load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
// shared -> global, 2-D box at (c0, c1); rows / columns outside the tensor are not written
__device__ __forceinline__ void store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(src)
                 : "memory");
}
// global -> shared, `bytes` contiguous bytes (a multiple of 16, both addresses 16-byte aligned); counted on `bar`
__device__ __forceinline__ void load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread's bulk stores have finished READING shared memory (the stage may be rewritten)
__device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// byte offset of 16-byte chunk `c` (0..7) of record `r` inside a 128B-swizzled tile of 128-byte records
__device__ __forceinline__ uint32_t swz128(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }

}  // namespace tma

// host side (tma_host.cu): tensor maps through the driver entry point, no link-time dependency on libcuda
// pixels: uint8 plane, W x H, `pitch` bytes per row (multiple of 16), box 256 x 8, no swizzle
cudaError_t make_pixel_map(CUtensorMap *map, const void *base, long long pitch, int W, int H);
// records: nblocks x 128 bytes, box 128 x `rows` (1..32), 128-byte swizzle
cudaError_t make_record_map(CUtensorMap *map, const void *base, uint32_t nblocks, int rows);
bool tma_available();

}  // namespace dctb
