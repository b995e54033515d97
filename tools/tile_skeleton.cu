// tile_skeleton.cu -- what the TILE MOVEMENT of K1 / K2 alone can reach (context for roofline.frac).
// The bulk-tensor kernels of fwd_quant.cu / dequant_idct.cu with the arithmetic taken out: same tiling, same tensor
// maps, same per-warp two-stage mbarrier pipeline, same swizzled stages, the record / pixel bytes derived from the input
// by a handful of XORs.  The gap between these figures and the real kernels is what the arithmetic costs on top of
// the data movement; the gap to the copy peak is what this access pattern (256-byte row segments at the plane's
// pitch in, 4 KB tiles out; or the reverse) costs against a linear copy.
// build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -Idct_b200/csrc -o tools/tile_skeleton tools/tile_skeleton.cu dct_b200/csrc/tma_host.cu
#include <cstdio>
#include <cstdlib>

#include "tma.cuh"

using namespace dctb;

constexpr int kThreads = 256, kWarps = 8;

struct alignas(64) Params {
    CUtensorMap map_px, map_rec;
    uint8_t *px;
    long long pitch;
    uint32_t bw, nby, tpr, step_ty, step_tx;
};

// MODE 0: K1 movement (pixels in through TMA, records out through TMA); MODE 1: K2 movement (records in through TMA,
// pixels out with 8 STG.64 per lane)
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_skel(const __grid_constant__ Params P)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    uint8_t *sm = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    constexpr int kIn = MODE == 0 ? 2048 : 4096;
    uint8_t *out_p = sm + warp * 4096;
    uint8_t *in_p = sm + kWarps * 4096 + warp * (2 * kIn);
    uint32_t *ctl_p = reinterpret_cast<uint32_t *>(sm + kWarps * 4096 + kWarps * 2 * kIn + warp * 32);
    const uint32_t out_s = (uint32_t)__cvta_generic_to_shared(out_p), in_s = (uint32_t)__cvta_generic_to_shared(in_p);
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(ctl_p);
    if (lane == 0) {
        tma::mbar_init(bar_s, 1);
        tma::mbar_init(bar_s + 8, 1);
        tma::fence_barrier_init();
    }
    __syncwarp();
    uint32_t ty, tx;
    {
        const uint32_t t = blockIdx.x * kWarps + warp;
        ty = t / P.tpr;
        tx = t - ty * P.tpr;
    }
    auto issue = [&](uint32_t stage) {
        if (lane == 0) {
            tma::mbar_expect_tx(bar_s + stage * 8, kIn);
            if (MODE == 0) tma::load_2d(in_s + stage * kIn, &P.map_px, (int)(tx * 256), (int)(ty * 8), bar_s + stage * 8);
            else tma::load_2d(in_s + stage * kIn, &P.map_rec, 0, (int)(ty * P.bw + tx * 32), bar_s + stage * 8);
        }
    };
    if (ty < P.nby) issue(0);
    const uint32_t swz = (lane & 7) << 4;
    for (uint32_t it = 0; ty < P.nby; ++it) {
        const uint32_t stage = it & 1;
        const uint32_t bx0 = tx * 32;
        const uint32_t warp_base = ty * P.bw + bx0;
        uint8_t *dst = P.px + (long long)ty * 8 * P.pitch + (long long)(bx0 + lane) * 8;
        tx += P.step_tx;
        ty += P.step_ty;
        if (tx >= P.tpr) tx -= P.tpr, ++ty;
        if (ty < P.nby) issue(stage ^ 1);
        tma::mbar_wait(bar_s + stage * 8, (it >> 1) & 1);
        if (MODE == 0) {
            uint2 raw[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) raw[i] = *reinterpret_cast<const uint2 *>(in_p + stage * kIn + i * 256 + lane * 8);
            if (lane == 0) tma::store_wait_read();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4 *>(out_p + lane * 128 + ((j << 4) ^ swz)) =
                    make_uint4(raw[j].x, raw[j].y ^ j, raw[(j + 1) & 7].x, raw[(j + 3) & 7].y);
            tma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma::store_2d(&P.map_rec, 0, (int)warp_base, out_s);
                tma::store_commit();
            }
        } else {
            uint4 t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = *reinterpret_cast<const uint4 *>(in_p + stage * kIn + lane * 128 + ((j << 4) ^ swz));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" ::"l"(dst + i * P.pitch), "r"(t[i].x ^ t[i].z), "r"(t[i].y ^ t[i].w) : "memory");
        }
    }
    if (MODE == 0 && lane == 0) tma::store_wait_read();
}

template <int MODE> static void run(const char *name, int per_sm, uint8_t *px, int16_t *coef, int W, int H, uint8_t *flush, size_t flush_n)
{
    constexpr int kIn = MODE == 0 ? 2048 : 4096;
    const int smem = 1024 + kWarps * 4096 + kWarps * 2 * kIn + kWarps * 32;
    cudaFuncSetAttribute(k_skel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_skel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_skel<MODE>, kThreads, smem);
    if (per_sm > occ) return;
    Params P;
    P.px = px, P.pitch = W, P.bw = W / 8, P.nby = H / 8, P.tpr = (P.bw + 31) / 32;
    const unsigned grid = 148 * per_sm, n_segs = grid * kWarps;
    P.step_ty = n_segs / P.tpr, P.step_tx = n_segs - P.step_ty * P.tpr;
    if (make_pixel_map(&P.map_px, px, W, W, H) || make_record_map(&P.map_rec, coef, P.bw * P.nby, 32)) {
        printf("{\"error\": \"tensor map\"}\n");
        return;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 8; ++rep) {
        cudaMemsetAsync(flush, rep, flush_n);
        cudaEventRecord(e0);
        k_skel<MODE><<<grid, kThreads, smem>>>(P);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double gb = 3.0 * W * H / 1e9;
    printf("{\"movement\": \"%s\", \"ctas_per_sm\": %d, \"occupancy_limit\": %d, \"gb\": %.3f, \"ms\": %.4f, \"gb_s\": %.1f}\n", name, per_sm, occ, gb,
           best, gb / best * 1e3);
}

int main()
{
    const int W = 3840, H = 2160 * 64;
    uint8_t *px, *flush;
    int16_t *coef;
    const size_t flush_n = (size_t)256 << 20;
    if (cudaMalloc(&px, (size_t)W * H) || cudaMalloc(&coef, (size_t)W * H * 2) || cudaMalloc(&flush, flush_n)) return 1;
    cudaMemset(px, 1, (size_t)W * H), cudaMemset(coef, 2, (size_t)W * H * 2);
    for (int n = 1; n <= 4; ++n) run<0>("K1: pixel boxes in (TMA), record tiles out (TMA)", n, px, coef, W, H, flush, flush_n);
    for (int n = 1; n <= 4; ++n) run<1>("K2: record tiles in (TMA), pixel rows out (STG.64)", n, px, coef, W, H, flush, flush_n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e) {
        fprintf(stderr, "%s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
