// microbench.cu -- issue rate of the SASS instructions the kernels' inner loops are made of (sm_100a).
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench.cu -o gpurun_out/microbench && gpurun_out/microbench
// Each test runs 8 independent dependency chains per thread, 1024 threads per SM resident, and
// reports lane-operations per clock per SM (128 = one warp instruction per SMSP per clock).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ITERS 2048

template <typename Op> __global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed)
{
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) Op::run(r[i]);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

struct OpFFMA  { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+r"(r)); } };
struct OpFADD  { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("add.rn.f32 %0, %0, %0;" : "+r"(r)); } };
struct OpPRMT  { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("prmt.b32 %0, %0, 0x4B000000, 0x7440;" : "+r"(r)); } };
struct OpLOP3  { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("lop3.b32 %0, %0, 0x80808080, %0, 0x96;" : "+r"(r)); } };
struct OpIADD  { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("add.u32 %0, %0, 12345;" : "+r"(r)); } };
struct OpI2FS8 { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("{.reg .s8 t; .reg .b8 a,b,c,d; mov.b32 {a,b,c,d}, %0; cvt.rn.f32.s8 %0, b;}" : "+r"(r)); } };
struct OpI2FS16{ static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("{.reg .b16 lo,hi; mov.b32 {lo,hi}, %0; cvt.rn.f32.s16 %0, hi;}" : "+r"(r)); } };
struct OpI2FS32{ static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r)); } };
struct OpF2I   { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("cvt.rni.s32.f32 %0, %0;" : "+r"(r)); } };
struct OpF2U8  { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("cvt.rni.sat.u8.f32 %0, %0;" : "+r"(r)); } };
struct OpFMNMX3{ static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("max.f32 %0, %0, %0, 0f3F000000;" : "+r"(r)); } };
struct OpFMNMX { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("max.f32 %0, %0, 0f3F000000;" : "+r"(r)); } };
struct OpVIMN  { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("min.s16x2.relu %0, %0, %1;" : "+r"(r) : "r"(0x00ff00ffu)); } };
struct OpH2F   { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("{.reg .b16 lo,hi; mov.b32 {lo,hi}, %0; cvt.f32.f16 %0, hi;}" : "+r"(r)); } };
struct OpHADD2 { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("add.rn.f16x2 %0, %0, %0;" : "+r"(r)); } };
struct OpFSETP { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("{.reg .pred p; setp.ge.f32 p, %0, 0f3F000000; selp.b32 %0, %0, 7, p;}" : "+r"(r)); } };
struct OpSHFL  { static __device__ __forceinline__ void run(uint32_t &r) { asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(r)); } };

template <typename Op> __global__ void __launch_bounds__(256) kd(double *out, double seed)
{
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) Op::run(r[i]);
    }
    double acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += r[i];
    if (acc == 0.12345) out[threadIdx.x] = acc;
}
template <typename Op> __global__ void __launch_bounds__(256) kl(unsigned long long *out, unsigned long long seed)
{
    unsigned long long r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) Op::run(r[i]);
    }
    unsigned long long acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i];
    if (acc == 0x12345678ull) out[threadIdx.x] = acc;
}
struct OpFFMA2 { static __device__ __forceinline__ void run(unsigned long long &r) { asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(r)); } };
struct OpFADD2 { static __device__ __forceinline__ void run(unsigned long long &r) { asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(r)); } };
struct OpFMUL2 { static __device__ __forceinline__ void run(unsigned long long &r) { asm volatile("mul.rn.f32x2 %0, %0, %0;" : "+l"(r)); } };
struct OpDFMA { static __device__ __forceinline__ void run(double &r) { asm volatile("fma.rn.f64 %0, %0, %0, %0;" : "+d"(r)); } };
struct OpDADD { static __device__ __forceinline__ void run(double &r) { asm volatile("add.rn.f64 %0, %0, %0;" : "+d"(r)); } };
struct OpDMUL { static __device__ __forceinline__ void run(double &r) { asm volatile("mul.rn.f64 %0, %0, %0;" : "+d"(r)); } };

// mixed issue: 8 packed chains + S scalar chains per iteration -- do scalar FP32 / integer instructions overlap with
// the packed ones (separate pipe) or queue behind them (same pipe)?
template <typename OpS, int S> __global__ void __launch_bounds__(256) kmix(unsigned long long *out, unsigned long long seed)
{
    unsigned long long r[8];
    uint32_t q[S > 0 ? S : 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;
#pragma unroll
    for (int i = 0; i < S; ++i) q[i] = (uint32_t)seed + threadIdx.x * 3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(r[i]));
#pragma unroll
            for (int j = 0; j < S / 8; ++j) OpS::run(q[i * (S / 8) + j]);
        }
    }
    unsigned long long acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i];
#pragma unroll
    for (int i = 0; i < S; ++i) acc ^= q[i];
    if (acc == 0x12345678ull) out[threadIdx.x] = acc;
}

template <typename F> static void timeit(const char *name, F launch, double ops_per_thread, int sms, double mhz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double threads = (double)sms * 4 * 256;
    const double per_clk_sm = threads * ops_per_thread / (ms * 1e-3) / (mhz * 1e6) / sms;
    printf("%-14s %8.3f ms  %7.1f lane-ops/clk/SM  (%.2f clk per warp instr per SMSP)\n", name, ms, per_clk_sm, 32.0 * 4 / per_clk_sm);
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz (nominal max; rates assume the clock holds)\n", prop.name, sms, mhz);
    uint32_t *out;
    double *outd;
    cudaMalloc(&out, 4096);
    cudaMalloc(&outd, 8192);
    const dim3 grid(sms * 4), block(256);
#define T(OP) timeit(#OP, [&] { k<OP><<<grid, block>>>(out, 1u); }, 8.0 * ITERS, sms, mhz)
    T(OpFFMA); T(OpFADD); T(OpPRMT); T(OpLOP3); T(OpIADD); T(OpI2FS8); T(OpI2FS16); T(OpI2FS32); T(OpF2I); T(OpF2U8);
    T(OpFMNMX3); T(OpFMNMX); T(OpVIMN); T(OpH2F); T(OpHADD2); T(OpFSETP); T(OpSHFL);
#define TD(OP) timeit(#OP, [&] { kd<OP><<<grid, block>>>(outd, 1.0); }, 8.0 * ITERS / 4, sms, mhz)
    TD(OpDFMA); TD(OpDADD); TD(OpDMUL);
    printf("-- packed fp32x2 (each instruction = 2 fp32 operations per lane; rate below counts instructions)\n");
#define TL(OP) timeit(#OP, [&] { kl<OP><<<grid, block>>>((unsigned long long *)outd, 1ull); }, 8.0 * ITERS, sms, mhz)
    TL(OpFFMA2); TL(OpFADD2); TL(OpFMUL2);
    printf("-- 8 FFMA2 + S other instructions per iteration; clk per iteration per SMSP with 8 warps resident (8 FFMA2 alone ~ 134)\n");
#define TM(OP, S)                                                                                                        \
    do {                                                                                                                 \
        cudaEvent_t e0, e1;                                                                                              \
        cudaEventCreate(&e0), cudaEventCreate(&e1);                                                                      \
        kmix<OP, S><<<grid, block>>>((unsigned long long *)outd, 1ull);                                                  \
        cudaDeviceSynchronize();                                                                                         \
        cudaEventRecord(e0);                                                                                             \
        kmix<OP, S><<<grid, block>>>((unsigned long long *)outd, 1ull);                                                  \
        cudaEventRecord(e1);                                                                                             \
        cudaEventSynchronize(e1);                                                                                        \
        float ms;                                                                                                        \
        cudaEventElapsedTime(&ms, e0, e1);                                                                               \
        printf("8 FFMA2 + %2d %-9s %8.3f ms  %7.1f clk per iteration per SMSP\n", S, #OP, ms, ms * 1e-3 * mhz * 1e6 / ITERS); \
    } while (0)
    TM(OpFFMA, 0); TM(OpFFMA, 8); TM(OpFFMA, 16); TM(OpFADD, 8); TM(OpPRMT, 8); TM(OpPRMT, 16); TM(OpIADD, 8); TM(OpIADD, 16);
    TM(OpFMNMX, 8); TM(OpLOP3, 16);
    return 0;
}
