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

__global__ void k_empty() {}

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

    // ---- the floor of this way of timing: two empty kernels between the events ----
    {
        std::vector<float> r;
        for (int k = 0; k < 50; ++k) {
            cudaStreamSynchronize(s);
            cudaEventRecord(e0, s);
            k_empty<<<148, 384, 0, s>>>();
            k_empty<<<148, 512, 0, s>>>();
            cudaEventRecord(e2, s);
            cudaEventSynchronize(e2);
            float a;
            cudaEventElapsedTime(&a, e0, e2);
            if (k > 5) r.push_back(a * 1e3f);
        }
        printf("{\"shape\": \"two empty kernels\", \"fwd_plus_inv_us\": %.1f}\n", median(r));
    }
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
        // the same two calls captured once per frame into a CUDA graph and replayed: what is left when the host's
        // share of a call (argument checks, three tensor maps, two launches) is off the critical path
        std::vector<cudaGraphExec_t> gx(N);
        bool graphs = true;
        for (int k = 0; k < N && graphs; ++k) {
            cudaGraph_t g;
            cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
            int rc = dct_cuda_fwd_quant_u8_dev(plan, px + px_b * k, W, W, H, coef + px_b * k, DCT_CUDA_NATURAL, nullptr, s);
            rc |= dct_cuda_dequant_idct_u8_dev(plan, coef + px_b * k, W, H, DCT_CUDA_NATURAL, nullptr, out + px_b * k, W, s);
            graphs = cudaStreamEndCapture(s, &g) == cudaSuccess && rc == 0 && cudaGraphInstantiate(&gx[k], g, 0) == cudaSuccess;
            if (g) cudaGraphDestroy(g);
        }
        std::vector<float> gr;
        if (graphs)
            for (int rep = 0; rep < 3; ++rep)
                for (int k = 0; k < N; ++k) {
                    cudaStreamSynchronize(s);
                    cudaEventRecord(e0, s);
                    cudaGraphLaunch(gx[k], s);
                    cudaEventRecord(e2, s);
                    cudaEventSynchronize(e2);
                    float a;
                    cudaEventElapsedTime(&a, e0, e2);
                    if (rep) gr.push_back(a * 1e3f);
                }
        else
            cudaGetLastError();
        printf("{\"shape\": \"C2 3840x2160\", \"fwd_us\": %.1f, \"inv_us\": %.1f, \"fwd_plus_inv_us\": %.1f, \"min_fwd_plus_inv_us\": %.1f, "
               "\"gpixel_s\": %.1f, \"graph_fwd_plus_inv_us\": %.1f, \"graph_min_us\": %.1f}\n",
               median(f), median(i), median(r), *std::min_element(r.begin(), r.end()), 2.0 * px_b / median(r) / 1e3,
               graphs ? median(gr) : -1.0, graphs ? *std::min_element(gr.begin(), gr.end()) : -1.0);
        for (int k = 0; k < N; ++k)
            if (graphs) cudaGraphExecDestroy(gx[k]);
        cudaFree(px), cudaFree(out), cudaFree(coef);
    }
    // ---- 4:2:0 frames (luma + two chroma planes, one call each way): C3 = a pool of 6 8K frames (6 x 49.8 MB of samples
    // + records), and 1080p video frames (padded to 1088 rows) from a pool of 96 ----
    struct Yuv { int W, H, N; const char *name; };
    const Yuv yuv[2] = {{7680, 4320, 6, "C3 7680x4320 4:2:0"}, {1920, 1088, 96, "1920x1088 4:2:0"}};
    for (const Yuv &cfg : yuv) {
        const int W = cfg.W, H = cfg.H, N = cfg.N;
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
        std::vector<cudaGraphExec_t> gx(N);
        bool graphs = true;
        for (int k = 0; k < N && graphs; ++k) {
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
            cudaGraph_t g;
            cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
            int rc = dct_cuda_fwd_quant_planes_dev(pl, 3, DCT_CUDA_NATURAL, s);
            rc |= dct_cuda_dequant_idct_planes_dev(pl, 3, DCT_CUDA_NATURAL, s);
            graphs = cudaStreamEndCapture(s, &g) == cudaSuccess && rc == 0 && cudaGraphInstantiate(&gx[k], g, 0) == cudaSuccess;
            if (g) cudaGraphDestroy(g);
        }
        std::vector<float> gr;
        if (graphs)
            for (int rep = 0; rep < 4; ++rep)
                for (int k = 0; k < N; ++k) {
                    cudaStreamSynchronize(s);
                    cudaEventRecord(e0, s);
                    cudaGraphLaunch(gx[k], s);
                    cudaEventRecord(e2, s);
                    cudaEventSynchronize(e2);
                    float a;
                    cudaEventElapsedTime(&a, e0, e2);
                    if (rep) gr.push_back(a * 1e3f);
                }
        else
            cudaGetLastError();
        printf("{\"shape\": \"%s\", \"fwd_us\": %.1f, \"inv_us\": %.1f, \"fwd_plus_inv_us\": %.1f, \"min_fwd_plus_inv_us\": %.1f, "
               "\"gpixel_s\": %.1f, \"graph_fwd_plus_inv_us\": %.1f, \"graph_min_us\": %.1f}\n",
               cfg.name, median(f), median(i), median(r), *std::min_element(r.begin(), r.end()), 2.0 * fr_b / median(r) / 1e3,
               graphs ? median(gr) : -1.0, graphs ? *std::min_element(gr.begin(), gr.end()) : -1.0);
        for (int k = 0; k < N; ++k)
            if (graphs) cudaGraphExecDestroy(gx[k]);
        cudaFree(px), cudaFree(out), cudaFree(coef);
    }
    dct_cuda_plan_destroy(plan), dct_cuda_plan_destroy(planc);
    dct_free(d), quant_free(q), quant_free(qc);
    return 0;
}
