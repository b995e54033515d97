// shim_blocks.cu -- C ABI: record adapters for the host consumer and the per-block drop-in calls of
// include/dct.h:51,61 / include/quantization.h:69,79 (dct_forward, dct_inverse, quantize, dequantize).
#include "plan.cuh"

using namespace dctb;
using namespace dctb::shim;

// ------------------------------------------------------------------------------------------
// adapters for the host consumer
// ------------------------------------------------------------------------------------------
extern "C" void dct_cuda_record_to_block(const int16_t *rec, int layout, int **block)
{
    for (int k = 0; k < 64; ++k) {
        const int nat = layout == DCT_CUDA_ZIGZAG ? kZigZag.nat[k] : k;
        block[nat >> 3][nat & 7] = rec[k];
    }
}

extern "C" void dct_cuda_block_to_record(int **block, int layout, int16_t *rec)
{
    for (int k = 0; k < 64; ++k) {
        const int nat = layout == DCT_CUDA_ZIGZAG ? kZigZag.nat[k] : k;
        rec[k] = (int16_t)block[nat >> 3][nat & 7];
    }
}

extern "C" void *dct_cuda_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        fail(DCT_CUDA_ENOMEM, "cudaMallocHost(%zu) failed", bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void dct_cuda_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------
// per-block drop-in calls: include/dct.h:51,61  include/quantization.h:69,79
// ------------------------------------------------------------------------------------------
namespace {

struct BlockScratch {
    bool ready = false;
    double *d_tab = nullptr, *d_in = nullptr, *d_out = nullptr;
    int *d_int = nullptr;
    double *h = nullptr;      // pinned: 3 * 1024 doubles
    int *h_int = nullptr;     // pinned: 1024 ints
    cudaStream_t stream = nullptr;
    int device = 0;           // the device that was current on first use; every later call switches to it
};
std::mutex g_block_mu;
BlockScratch g_block;

[[noreturn]] void die(const char *what, cudaError_t e)
{
    fprintf(stderr, "libdct_cuda: %s failed: %s (no CPU fallback)\n", what, cudaGetErrorString(e));
    exit(EXIT_FAILURE);
}
#define CU_DIE(expr)                                  \
    do {                                              \
        cudaError_t e_ = (expr);                      \
        if (e_ != cudaSuccess) die(#expr, e_);        \
    } while (0)

BlockScratch &scratch()
{
    if (!g_block.ready) {
        CU_DIE(cudaGetDevice(&g_block.device));
        CU_DIE(cudaMalloc(&g_block.d_tab, 1024 * sizeof(double)));
        CU_DIE(cudaMalloc(&g_block.d_in, 1024 * sizeof(double)));
        CU_DIE(cudaMalloc(&g_block.d_out, 1024 * sizeof(double)));
        CU_DIE(cudaMalloc(&g_block.d_int, 1024 * sizeof(int)));
        CU_DIE(cudaMallocHost(&g_block.h, 3 * 1024 * sizeof(double)));
        CU_DIE(cudaMallocHost(&g_block.h_int, 1024 * sizeof(int)));
        CU_DIE(cudaStreamCreateWithFlags(&g_block.stream, cudaStreamNonBlocking));
        g_block.ready = true;
    }
    return g_block;
}

void check_n(int n)
{
    if (n < 1 || n > 32) {
        fprintf(stderr, "libdct_cuda: block_size %d unsupported (1..32)\n", n);
        exit(EXIT_FAILURE);
    }
}

void block_transform(DCTContext *ctx, double **input, double **output, int inverse)
{
    const int n = ctx->block_size;
    check_n(n);
    std::lock_guard<std::mutex> lk(g_block_mu);
    BlockScratch &s = scratch();
    DeviceGuard dg(s.device);
    double *hD = s.h, *hI = s.h + 1024, *hO = s.h + 2048;
    for (int i = 0; i < n; ++i) {
        memcpy(hD + i * n, ctx->dct_matrix[i], n * sizeof(double));   // rows are separate mallocs
        memcpy(hI + i * n, input[i], n * sizeof(double));
    }
    const size_t bytes = (size_t)n * n * sizeof(double);
    CU_DIE(cudaMemcpyAsync(s.d_tab, hD, bytes, cudaMemcpyHostToDevice, s.stream));
    CU_DIE(cudaMemcpyAsync(s.d_in, hI, bytes, cudaMemcpyHostToDevice, s.stream));
    CU_DIE(launch_block_dct_f64(n, s.d_tab, s.d_in, s.d_out, inverse, s.stream));
    CU_DIE(cudaMemcpyAsync(hO, s.d_out, bytes, cudaMemcpyDeviceToHost, s.stream));
    CU_DIE(cudaStreamSynchronize(s.stream));
    for (int i = 0; i < n; ++i) memcpy(output[i], hO + i * n, n * sizeof(double));
}

}  // namespace

extern "C" void dct_forward(DCTContext *ctx, double **input, double **output) { block_transform(ctx, input, output, 0); }
extern "C" void dct_inverse(DCTContext *ctx, double **input, double **output) { block_transform(ctx, input, output, 1); }

extern "C" void quantize(QuantContext *ctx, double **dct_coeffs, int **quant_coeffs, double block_variance)
{
    const int n = ctx->block_size;
    check_n(n);
    std::lock_guard<std::mutex> lk(g_block_mu);
    BlockScratch &s = scratch();
    DeviceGuard dg(s.device);
    double *hQ = s.h, *hC = s.h + 1024;
    for (int i = 0; i < n; ++i) {
        memcpy(hQ + i * n, ctx->quant_matrix[i], n * sizeof(double));
        memcpy(hC + i * n, dct_coeffs[i], n * sizeof(double));
    }
    const size_t bytes = (size_t)n * n * sizeof(double);
    CU_DIE(cudaMemcpyAsync(s.d_tab, hQ, bytes, cudaMemcpyHostToDevice, s.stream));
    CU_DIE(cudaMemcpyAsync(s.d_in, hC, bytes, cudaMemcpyHostToDevice, s.stream));
    CU_DIE(launch_block_quantize_f64(n, s.d_tab, ctx->adaptive, block_variance, s.d_in, s.d_int, s.stream));
    CU_DIE(cudaMemcpyAsync(s.h_int, s.d_int, (size_t)n * n * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CU_DIE(cudaStreamSynchronize(s.stream));
    for (int i = 0; i < n; ++i) memcpy(quant_coeffs[i], s.h_int + i * n, n * sizeof(int));
}

extern "C" void dequantize(QuantContext *ctx, int **quant_coeffs, double **dct_coeffs, double block_variance)
{
    const int n = ctx->block_size;
    check_n(n);
    std::lock_guard<std::mutex> lk(g_block_mu);
    BlockScratch &s = scratch();
    DeviceGuard dg(s.device);
    double *hR = s.h, *hO = s.h + 2048;
    for (int i = 0; i < n; ++i) {
        memcpy(hR + i * n, ctx->dequant_matrix[i], n * sizeof(double));
        memcpy(s.h_int + i * n, quant_coeffs[i], n * sizeof(int));
    }
    const size_t bytes = (size_t)n * n * sizeof(double);
    CU_DIE(cudaMemcpyAsync(s.d_tab, hR, bytes, cudaMemcpyHostToDevice, s.stream));
    CU_DIE(cudaMemcpyAsync(s.d_int, s.h_int, (size_t)n * n * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    CU_DIE(launch_block_dequantize_f64(n, s.d_tab, ctx->adaptive, block_variance, s.d_int, s.d_out, s.stream));
    CU_DIE(cudaMemcpyAsync(hO, s.d_out, bytes, cudaMemcpyDeviceToHost, s.stream));
    CU_DIE(cudaStreamSynchronize(s.stream));
    for (int i = 0; i < n; ++i) memcpy(dct_coeffs[i], hO + i * n, n * sizeof(double));
}

