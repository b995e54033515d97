// latency.cu -- single-frame latency of the plane calls, from a plain C++ caller of libdct_cuda's C ABI
// (BASELINE configs[1] and [2] as written: ONE 3840x2160 frame; ONE 7680x4320 4:2:0 frame = luma + two chroma planes).
// No Python in the loop: what a host program linking -ldct_cuda sees.  Each measurement is one frame out of a pool
// larger than L2 (so the frame is cold), CUDA events around the calls of that frame; median over the pool.
// build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -Iinclude -o tools/latency tools/latency.cu -Ldct_b200 -ldct_cuda -Xlinker -rpath,'$ORIGIN/../dct_b200'
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>
#include <dct_cuda.h>

#define CK(x)                                                                      \
    do {                                                                           \
        if ((x) != 0) {                                                            \
            fprintf(stderr, "%s failed: %s\n", #x, dct_cuda_last_error());         \
            return 1;                                                              \
        }                                                                          \
    } while (0)

static double median(std::vector<float> v)
{
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}

int main()
{
    DCTContext *d = dct_init(8);
    QuantContext *q = quant_init(8, 50, 0), *qc = quant_init(8, 50, 0);
    for (int i = 0; i < 8; ++i)   // a flatter table for the chroma planes (any second table will do for timing)
        for (int j = 0; j < 8; ++j) qc->quant_matrix[i][j] = 99.0, qc->dequant_matrix[i][j] = 1.0 / 99.0;
    dct_cuda_plan *plan = dct_cuda_plan_create(d, q, 0), *planc = dct_cuda_plan_create(d, qc, 0);
    if (!plan || !planc) {
        fprintf(stderr, "plan: %s\n", dct_cuda_last_error());
        return 1;
    }
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0), cudaEventCreate(&e1), cudaEventCreate(&e2);

    // ---- C2: a pool of 24 4K frames (199 MB of pixels + 398 MB of records > L2) ----
    {
        const int W = 3840, H = 2160, N = 24;
        const size_t px_b = (size_t)W * H, rec_b = px_b * 2;
        uint8_t *px, *out;
        int16_t *coef;
        cudaMalloc(&px, px_b * N), cudaMalloc(&out, px_b * N), cudaMalloc(&coef, rec_b * N);
        std::vector<uint8_t> h(px_b * N);
        uint64_t x = 0x9E3779B97F4A7C15ull;
        for (auto &b : h) x ^= x << 13, x ^= x >> 7, x ^= x << 17, b = (uint8_t)(x >> 32);
        cudaMemcpy(px, h.data(), h.size(), cudaMemcpyHostToDevice);
        std::vector<float> f, i, r;
        for (int rep = 0; rep < 3; ++rep)
            for (int k = 0; k < N; ++k) {
                cudaStreamSynchronize(s);
                cudaEventRecord(e0, s);
                CK(dct_cuda_fwd_quant_u8_dev(plan, px + px_b * k, W, W, H, coef + px_b * k, DCT_CUDA_NATURAL, nullptr, s));
                cudaEventRecord(e1, s);
                CK(dct_cuda_dequant_idct_u8_dev(plan, coef + px_b * k, W, H, DCT_CUDA_NATURAL, nullptr, out + px_b * k, W, s));
                cudaEventRecord(e2, s);
                cudaEventSynchronize(e2);
                float a, b;
                cudaEventElapsedTime(&a, e0, e1), cudaEventElapsedTime(&b, e1, e2);
                if (rep) f.push_back(a * 1e3f), i.push_back(b * 1e3f), r.push_back((a + b) * 1e3f);
            }
        printf("{\"shape\": \"C2 3840x2160\", \"fwd_us\": %.1f, \"inv_us\": %.1f, \"fwd_plus_inv_us\": %.1f, \"min_fwd_plus_inv_us\": %.1f, "
               "\"gpixel_s\": %.1f}\n",
               median(f), median(i), median(r), *std::min_element(r.begin(), r.end()), 2.0 * px_b / median(r) / 1e3);
        cudaFree(px), cudaFree(out), cudaFree(coef);
    }
    // ---- C3: a pool of 6 8K 4:2:0 frames (6 x 49.8 MB of samples + records) ----
    {
        const int W = 7680, H = 4320, N = 6;
        const size_t y_b = (size_t)W * H, c_b = y_b / 4, fr_b = y_b + 2 * c_b;
        uint8_t *px, *out;
        int16_t *coef;
        cudaMalloc(&px, fr_b * N), cudaMalloc(&out, fr_b * N), cudaMalloc(&coef, fr_b * 2 * N);
        std::vector<uint8_t> h(fr_b * N);
        uint64_t x = 0x1234567ull;
        for (auto &b : h) x ^= x << 13, x ^= x >> 7, x ^= x << 17, b = (uint8_t)(x >> 32);
        cudaMemcpy(px, h.data(), h.size(), cudaMemcpyHostToDevice);
        std::vector<float> f, i, r;
        for (int rep = 0; rep < 4; ++rep)
            for (int k = 0; k < N; ++k) {
                dct_cuda_plane pl[3];
                const size_t off[3] = {0, y_b, y_b + c_b};
                for (int c = 0; c < 3; ++c) {
                    pl[c].plan = c ? planc : plan;
                    pl[c].pixels_in = px + fr_b * k + off[c];
                    pl[c].pixels_out = out + fr_b * k + off[c];
                    pl[c].pitch = c ? W / 2 : W;
                    pl[c].width = c ? W / 2 : W, pl[c].height = c ? H / 2 : H;
                    pl[c].coef = coef + fr_b * k + off[c];
                    pl[c].variance = nullptr;
                }
                cudaStreamSynchronize(s);
                cudaEventRecord(e0, s);
                CK(dct_cuda_fwd_quant_planes_dev(pl, 3, DCT_CUDA_NATURAL, s));
                cudaEventRecord(e1, s);
                CK(dct_cuda_dequant_idct_planes_dev(pl, 3, DCT_CUDA_NATURAL, s));
                cudaEventRecord(e2, s);
                cudaEventSynchronize(e2);
                float a, b;
                cudaEventElapsedTime(&a, e0, e1), cudaEventElapsedTime(&b, e1, e2);
                if (rep) f.push_back(a * 1e3f), i.push_back(b * 1e3f), r.push_back((a + b) * 1e3f);
            }
        printf("{\"shape\": \"C3 7680x4320 4:2:0\", \"fwd_us\": %.1f, \"inv_us\": %.1f, \"fwd_plus_inv_us\": %.1f, \"min_fwd_plus_inv_us\": %.1f, "
               "\"gpixel_s\": %.1f}\n",
               median(f), median(i), median(r), *std::min_element(r.begin(), r.end()), 2.0 * fr_b / median(r) / 1e3);
        cudaFree(px), cudaFree(out), cudaFree(coef);
    }
    dct_cuda_plan_destroy(plan), dct_cuda_plan_destroy(planc);
    dct_free(d), quant_free(q), quant_free(qc);
    return 0;
}
