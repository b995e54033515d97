// replay_f64.cu -- K3 / K4 / K6: the reference's arithmetic, operation for operation, in fp64.
//
// K1/K2 hand every block whose fp32 result sits inside the error band of a .5 rounding
// boundary to these kernels.  They repeat exactly what the reference does:
//   dct_forward   src/dct.c:57-74    temp = X*D^T, out = D*temp, acc = 0.0 then += for k ascending
//   quantize      src/quantization.c:113-131   (int) round(c / M), true division, half away
//   dequantize    src/quantization.c:133-151   q * R   |   q * (1.0 / (R * (1.0/(2-nv))))
//   dct_inverse   src/dct.c:85-102   temp = D^T*in, out = temp*D
//   adjust_matrix_for_block src/quantization.c:171-211
// using the HOST-computed tables (glibc's cos() noise is part of the answer)
// and non-contracted __dmul_rn/__dadd_rn/__ddiv_rn, so the integers are bit-identical to
// the reference's, exact ties included.  The same code, generalised to n x n, backs the
// per-block drop-in entry points dct_forward/dct_inverse/quantize/dequantize.
#include "butterfly.cuh"
#include "kernels.cuh"

namespace dctb {

namespace {

__constant__ ZigZag cZigZag{};   // device copy of the scan order for run-time indexing

// C99 round(): half away from zero.  (y - trunc(y)) is exact for |y| < 2^52.
__device__ __forceinline__ double round_half_away(double y)
{
    const double t = trunc(y);
    return (fabs(__dsub_rn(y, t)) >= 0.5) ? __dadd_rn(t, copysign(1.0, y)) : t;
}

__device__ __forceinline__ bool near_half(double a)
{
    a = fabs(a);
    const double f = __dsub_rn(a, floor(a));
    return fabs(__dsub_rn(f, 0.5)) <= 1e-9;
}

// src/quantization.c:186: fmin(1.0, fmax(0.1, variance / 1000.0))
__device__ __forceinline__ double norm_variance(double variance)
{
    return fmin(1.0, fmax(0.1, __ddiv_rn(variance, 1000.0)));
}

constexpr int kBlocksPerCta = 4;   // 64 threads per 8x8 block

template <bool FORWARD>
__global__ void __launch_bounds__(64 * kBlocksPerCta) k_replay(const ReplayParams p)
{
    __shared__ double sD[64];
    __shared__ double sM[64];                     // quant matrix (fwd) / dequant matrix (inv)
    __shared__ double sX[kBlocksPerCta][64];
    __shared__ double sT[kBlocksPerCta][64];
    __shared__ double sVar[kBlocksPerCta];
    __shared__ unsigned long long sTies, sSat;

    const int tid = threadIdx.x, sub = tid >> 6, e = tid & 63, i = e >> 3, j = e & 7;
    if (tid < 64) {
        sD[tid] = p.tab->D[tid];
        sM[tid] = FORWARD ? p.tab->Q[tid] : p.tab->R[tid];
    }
    if (tid == 0) sTies = 0, sSat = 0;
    __syncthreads();

    unsigned count = p.nblocks;
    if (p.worklist != nullptr) {
        count = p.ctr->wl_count;
        if (count > p.wl_cap) count = p.wl_cap;
    }
    const unsigned rounds = (count + kBlocksPerCta - 1) / kBlocksPerCta;
    unsigned ties = 0, sat = 0;

    for (unsigned rnd = blockIdx.x; rnd < rounds; rnd += gridDim.x) {
        const unsigned slot = rnd * kBlocksPerCta + sub;
        const bool active = slot < count;
        const unsigned b = active ? (p.worklist ? p.worklist[slot] : slot) : 0;
        const unsigned by = b / p.bw, bx = b - by * p.bw;
        // position `e` of the record holds natural index nat (zigzag: src/entropy.c:158-178)
        const int nat = p.layout == LAYOUT_ZIGZAG ? cZigZag.nat[e] : e;

        if (FORWARD) {
            // src/dct.c:115  (double)px - 128.0
            const uint8_t px = p.px_in[((long long)by * 8 + i) * p.pitch + (long long)bx * 8 + j];
            sX[sub][e] = __dsub_rn((double)px, 128.0);
        } else {
            const int q = active ? (int)p.coef_in[(size_t)b * 64 + e] : 0;
            double m = sM[nat];
            double val;
            if (p.adaptive) {
                const double var = p.var_in ? p.var_in[b] : 0.0;
                if (nat != 0) m = __dmul_rn(m, __ddiv_rn(1.0, __dsub_rn(2.0, norm_variance(var))));
                val = __dmul_rn((double)q, __ddiv_rn(1.0, m));
            } else {
                val = __dmul_rn((double)q, m);
            }
            sX[sub][nat] = val;
        }
        __syncthreads();

        if (FORWARD) {
            if (p.adaptive && e == 0) {
                // src/quantization.c:153-169, sequential like the reference (all terms are exact integers)
                double sum = 0.0, sum_sq = 0.0;
                for (int k = 0; k < 64; ++k) {
                    sum = __dadd_rn(sum, sX[sub][k]);
                    sum_sq = __dadd_rn(sum_sq, __dmul_rn(sX[sub][k], sX[sub][k]));
                }
                const double mean = __ddiv_rn(sum, 64.0);
                sVar[sub] = __dsub_rn(__ddiv_rn(sum_sq, 64.0), __dmul_rn(mean, mean));
            }
            double acc = 0.0;   // temp[i][j] = sum_k X[i][k] * D[j][k]
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = __dadd_rn(acc, __dmul_rn(sX[sub][i * 8 + k], sD[j * 8 + k]));
            sT[sub][e] = acc;
        } else {
            double acc = 0.0;   // temp[i][j] = sum_k D[k][i] * in[k][j]
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = __dadd_rn(acc, __dmul_rn(sD[k * 8 + i], sX[sub][k * 8 + j]));
            sT[sub][e] = acc;
        }
        __syncthreads();

        if (FORWARD) {
            double acc = 0.0;   // out[i][j] = sum_k D[i][k] * temp[k][j]
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = __dadd_rn(acc, __dmul_rn(sD[i * 8 + k], sT[sub][k * 8 + j]));
            double m = sM[e];
            if (p.adaptive) {
                const double var = sVar[sub];
                if (e != 0) {
                    m = __dmul_rn(m, __dsub_rn(2.0, norm_variance(var)));
                    if (m < 1.0) m = 1.0;
                }
                if (e == 0 && active && p.var_out) p.var_out[b] = var;
            }
            const double y = __ddiv_rn(acc, m);
            const double r = round_half_away(y);
            int q = (int)r;
            if (r > 32767.0) q = 32767, ++sat;
            if (r < -32768.0) q = -32768, ++sat;
            if (active) {
                ties += near_half(y);
                sX[sub][e] = (double)q;   // own element only; re-read below by position
            }
        } else {
            double acc = 0.0;   // out[i][j] = sum_k temp[i][k] * D[k][j]
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = __dadd_rn(acc, __dmul_rn(sT[sub][i * 8 + k], sD[k * 8 + j]));
            const double v = __dadd_rn(acc, 128.0);
            double r = round_half_away(v);
            r = r < 0.0 ? 0.0 : (r > 255.0 ? 255.0 : r);
            if (active) {
                ties += near_half(v);
                p.px_out[((long long)by * 8 + i) * p.pitch + (long long)bx * 8 + j] = (uint8_t)r;
            }
        }
        __syncthreads();
        if (FORWARD && active) p.coef_out[(size_t)b * 64 + e] = (int16_t)(int)sX[sub][nat];
        __syncthreads();
    }

    if (ties) atomicAdd(&sTies, (unsigned long long)ties);
    if (sat) atomicAdd(&sSat, (unsigned long long)sat);
    __syncthreads();
    if (tid == 0) {
        if (sTies) atomicAdd(&p.ctr->near_ties, sTies);
        if (sSat) atomicAdd(&p.ctr->saturated, sSat);
        if (blockIdx.x == 0) atomicAdd(&p.ctr->replayed, (unsigned long long)count);
    }
}

// ---- generic n x n single-block kernels (n <= 32): the per-block drop-in API -------------

__global__ void k_block_dct_f64(int n, const double *D, const double *in, double *out, int inverse)
{
    extern __shared__ double sm[];
    double *sD = sm, *sI = sm + n * n, *sT = sm + 2 * n * n;
    const int e = threadIdx.x, i = e / n, j = e % n;
    sD[e] = D[e];
    sI[e] = in[e];
    __syncthreads();
    double acc = 0.0;
    if (!inverse) {
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(sI[i * n + k], sD[j * n + k]));
    } else {
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(sD[k * n + i], sI[k * n + j]));
    }
    sT[e] = acc;
    __syncthreads();
    acc = 0.0;
    if (!inverse) {
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(sD[i * n + k], sT[k * n + j]));
    } else {
        for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(sT[i * n + k], sD[k * n + j]));
    }
    out[e] = acc;
}

__global__ void k_block_quantize_f64(int n, const double *Q, int adaptive, double variance, const double *c,
                                     int *q)
{
    const int e = threadIdx.x;
    double m = Q[e];
    if (adaptive && e != 0) {
        m = __dmul_rn(m, __dsub_rn(2.0, norm_variance(variance)));
        if (m < 1.0) m = 1.0;
    }
    q[e] = (int)round_half_away(__ddiv_rn(c[e], m));
}

__global__ void k_block_dequantize_f64(int n, const double *R, int adaptive, double variance, const int *q,
                                       double *c)
{
    const int e = threadIdx.x;
    double m = R[e];
    if (adaptive) {
        if (e != 0) m = __dmul_rn(m, __ddiv_rn(1.0, __dsub_rn(2.0, norm_variance(variance))));
        c[e] = __dmul_rn((double)q[e], __ddiv_rn(1.0, m));
    } else {
        c[e] = __dmul_rn((double)q[e], m);
    }
}

}  // namespace

static unsigned replay_grid(const ReplayParams &p)
{
    // worklist mode: the count lives on the device; a fixed grid strides over it
    if (p.worklist != nullptr) return 148 * 4;
    const unsigned rounds = (p.nblocks + kBlocksPerCta - 1) / kBlocksPerCta;
    return rounds < 148u * 8u ? (rounds ? rounds : 1u) : 148u * 8u;
}

cudaError_t launch_replay_fwd(const ReplayParams &p, cudaStream_t s)
{
    k_replay<true><<<replay_grid(p), 64 * kBlocksPerCta, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_replay_inv(const ReplayParams &p, cudaStream_t s)
{
    k_replay<false><<<replay_grid(p), 64 * kBlocksPerCta, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_block_dct_f64(int n, const double *d_D, const double *d_in, double *d_out, int inverse,
                                 cudaStream_t s)
{
    k_block_dct_f64<<<1, n * n, 3 * n * n * sizeof(double), s>>>(n, d_D, d_in, d_out, inverse);
    return cudaGetLastError();
}

cudaError_t launch_block_quantize_f64(int n, const double *d_Q, int adaptive, double variance,
                                      const double *d_c, int *d_q, cudaStream_t s)
{
    k_block_quantize_f64<<<1, n * n, 0, s>>>(n, d_Q, adaptive, variance, d_c, d_q);
    return cudaGetLastError();
}

cudaError_t launch_block_dequantize_f64(int n, const double *d_R, int adaptive, double variance,
                                        const int *d_q, double *d_c, cudaStream_t s)
{
    k_block_dequantize_f64<<<1, n * n, 0, s>>>(n, d_R, adaptive, variance, d_q, d_c);
    return cudaGetLastError();
}

}  // namespace dctb
