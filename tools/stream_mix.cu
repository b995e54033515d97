// stream_mix.cu -- what HBM delivers for the read:write mixes of this repo's kernels (context for roofline.frac).
//   copy 1:1 (the mix MEASURED_PEAKS.json was taken with), 1:2 (K1: 1 B/px in, 2 B/px out; YCbCr->RGB),
//   2:1 (K2: 2 B/px in, 1 B/px out; RGB->YCbCr), read-only, write-only.  16-byte accesses, no arithmetic.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/stream_mix tools/stream_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int R, int W> __global__ void __launch_bounds__(256) k_mix(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n)
{
    // unit i: reads R vectors in[i*R .. i*R+R), writes W vectors out[i*W .. i*W+W); warp-contiguous per vector index
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const size_t warp_base = i - (i & 31), lane = i & 31;
        uint4 acc = make_uint4(lane, 1, 2, 3);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint4 v = __ldcs(in + warp_base * R + r * 32 + lane);
            acc.x ^= v.x, acc.y ^= v.y, acc.z ^= v.z, acc.w ^= v.w;
        }
#pragma unroll
        for (int w = 0; w < W; ++w) {
            acc.x += w;
            __stcs(out + warp_base * W + w * 32 + lane, acc);
        }
        if (W == 0 && acc.x == 0x12345678u && acc.y == 0x9abcdef0u) out[0] = acc;   // keep the loads alive
    }
}

template <int R, int W> static void run(const char *name, const uint4 *in, uint4 *out, size_t bytes_total, uint4 *flush, size_t flush_n)
{
    const size_t n = bytes_total / 16 / (R + W) / 32 * 32;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 8; ++rep) {
        cudaMemsetAsync(flush, rep, flush_n);
        cudaEventRecord(e0);
        k_mix<R, W><<<148 * 8, 256>>>(in, out, n);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double gb = (double)n * 16 * (R + W) / 1e9;
    printf("{\"mix\": \"%s\", \"read_vec\": %d, \"write_vec\": %d, \"gb\": %.3f, \"ms\": %.4f, \"gb_s\": %.1f}\n", name, R, W, gb, best,
           gb / best * 1e3);
}

int main()
{
    const size_t cap = (size_t)3 << 30;   // 3 GiB each side
    uint4 *in, *out, *flush;
    const size_t flush_n = (size_t)256 << 20;
    if (cudaMalloc(&in, cap) || cudaMalloc(&out, cap) || cudaMalloc(&flush, flush_n)) return 1;
    cudaMemset(in, 1, cap), cudaMemset(out, 2, cap);
    const size_t total = (size_t)3 << 30; // bytes moved per launch (read + write)
    run<1, 1>("copy 1:1", in, out, total, flush, flush_n);
    run<1, 2>("1:2 (K1, YCbCr->RGB)", in, out, total, flush, flush_n);
    run<2, 1>("2:1 (K2, RGB->YCbCr)", in, out, total, flush, flush_n);
    run<1, 0>("read only", in, out, total / 2, flush, flush_n);
    run<0, 1>("write only", in, out, total / 2, flush, flush_n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e) { fprintf(stderr, "%s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
