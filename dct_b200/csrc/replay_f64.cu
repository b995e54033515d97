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

constexpr int kReplayWarps = 8;   // one warp per 8x8 block, 8 independent warps per CTA

// One warp replays one block: lane l owns elements l and l+32 (rows l/8 and l/8+4, column l%8).
// Only __syncwarp() is needed, so the warps of a CTA never wait for each other and the long
// dependent DMUL/DADD chains of different blocks overlap.
template <bool FORWARD>
__global__ void __launch_bounds__(32 * kReplayWarps) k_replay(const ReplayParams p)
{
    __shared__ double sD[64];
    __shared__ double sM[64];                     // quant matrix (fwd) / dequant matrix (inv)
    __shared__ double sX[kReplayWarps][64];
    __shared__ double sT[kReplayWarps][64];

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if (tid < 64) {
        sD[tid] = p.tab->D[tid];
        sM[tid] = FORWARD ? p.tab->Q[tid] : p.tab->R[tid];
    }
    __syncthreads();

    unsigned count = p.nblocks;
    if (p.worklist != nullptr) {
        count = p.ctr->wl_count;
        if (count > p.wl_cap) count = p.wl_cap;
    }
    unsigned ties = 0, sat = 0;
    double *X = sX[w], *T = sT[w];
    const int j = lane & 7, i0 = lane >> 3;

    for (unsigned slot = blockIdx.x * kReplayWarps + w; slot < count; slot += gridDim.x * kReplayWarps) {
        const unsigned b = p.worklist ? p.worklist[slot] : slot;
        const unsigned by = b / p.bw, bx = b - by * p.bw;
        double var = 0.0;

        if (FORWARD) {
            int isum = 0, isq = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + 4 * h;
                // src/dct.c:115  (double)px - 128.0
                const int px = p.px_in[((long long)by * 8 + i) * p.pitch + (long long)bx * 8 + j];
                X[i * 8 + j] = __dsub_rn((double)px, 128.0);
                isum += px - 128;
                isq += (px - 128) * (px - 128);
            }
            if (p.adaptive) {
                // src/quantization.c:153-169.  Every partial sum there is an exactly representable
                // integer, so the summation order does not matter; the three fp64 ops below see the
                // same operands as the reference's.
                isum = __reduce_add_sync(0xffffffffu, isum);
                isq = __reduce_add_sync(0xffffffffu, isq);
                const double mean = __ddiv_rn((double)isum, 64.0);
                var = __dsub_rn(__ddiv_rn((double)isq, 64.0), __dmul_rn(mean, mean));
                if (lane == 0 && p.var_out) p.var_out[b] = var;
            }
        } else {
            if (p.adaptive) var = p.var_in ? p.var_in[b] : 0.0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = lane + 32 * h;   // record position
                // position e holds natural index nat (zigzag: src/entropy.c:183-210)
                const int nat = p.layout == LAYOUT_ZIGZAG ? cZigZag.nat[e] : e;
                const int q = (int)p.coef_in[(size_t)b * 64 + e];
                double m = sM[nat];
                double val;
                if (p.adaptive) {
                    if (nat != 0) m = __dmul_rn(m, __ddiv_rn(1.0, __dsub_rn(2.0, norm_variance(var))));
                    val = __dmul_rn((double)q, __ddiv_rn(1.0, m));
                } else {
                    val = __dmul_rn((double)q, m);
                }
                X[nat] = val;
            }
        }
        __syncwarp();

#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = i0 + 4 * h;
            double acc = 0.0;
            if (FORWARD) {   // temp[i][j] = sum_k X[i][k] * D[j][k]
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = __dadd_rn(acc, __dmul_rn(X[i * 8 + k], sD[j * 8 + k]));
            } else {         // temp[i][j] = sum_k D[k][i] * in[k][j]
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = __dadd_rn(acc, __dmul_rn(sD[k * 8 + i], X[k * 8 + j]));
            }
            T[i * 8 + j] = acc;
        }
        __syncwarp();

        int qv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = i0 + 4 * h, e = i * 8 + j;
            double acc = 0.0;
            if (FORWARD) {   // out[i][j] = sum_k D[i][k] * temp[k][j]
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = __dadd_rn(acc, __dmul_rn(sD[i * 8 + k], T[k * 8 + j]));
                double m = sM[e];
                if (p.adaptive && e != 0) {
                    m = __dmul_rn(m, __dsub_rn(2.0, norm_variance(var)));
                    if (m < 1.0) m = 1.0;
                }
                const double y = __ddiv_rn(acc, m);
                const double r = round_half_away(y);
                int q = (int)r;
                if (r > 32767.0) q = 32767, ++sat;
                if (r < -32768.0) q = -32768, ++sat;
                ties += near_half(y);
                qv[h] = q;
            } else {         // out[i][j] = sum_k temp[i][k] * D[k][j]
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = __dadd_rn(acc, __dmul_rn(T[i * 8 + k], sD[k * 8 + j]));
                const double v = __dadd_rn(acc, 128.0);
                double r = round_half_away(v);
                r = r < 0.0 ? 0.0 : (r > 255.0 ? 255.0 : r);
                ties += near_half(v);
                p.px_out[((long long)by * 8 + i) * p.pitch + (long long)bx * 8 + j] = (uint8_t)r;
            }
        }
        if (FORWARD) {
            // re-order through shared memory: record position e <- natural index nat
            int *Xi = reinterpret_cast<int *>(X);
            __syncwarp();
            Xi[i0 * 8 + j] = qv[0];
            Xi[(i0 + 4) * 8 + j] = qv[1];
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = lane + 32 * h;
                const int nat = p.layout == LAYOUT_ZIGZAG ? cZigZag.nat[e] : e;
                p.coef_out[(size_t)b * 64 + e] = (int16_t)Xi[nat];
            }
        }
        __syncwarp();
    }

    ties = __reduce_add_sync(0xffffffffu, ties);
    sat = __reduce_add_sync(0xffffffffu, sat);
    if (lane == 0) {
        if (ties) atomicAdd(&p.ctr->near_ties, (unsigned long long)ties);
        if (sat) atomicAdd(&p.ctr->saturated, (unsigned long long)sat);
    }
    if (tid == 0 && blockIdx.x == 0) atomicAdd(&p.ctr->replayed, (unsigned long long)count);
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
    // worklist mode: the count lives on the device; a fixed grid (8 CTAs per SM) strides over it
    if (p.worklist != nullptr) return 148 * 8;
    const unsigned ctas = (p.nblocks + kReplayWarps - 1) / kReplayWarps;
    return ctas < 148u * 8u ? (ctas ? ctas : 1u) : 148u * 8u;
}

cudaError_t launch_replay_fwd(const ReplayParams &p, cudaStream_t s)
{
    k_replay<true><<<replay_grid(p), 32 * kReplayWarps, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_replay_inv(const ReplayParams &p, cudaStream_t s)
{
    k_replay<false><<<replay_grid(p), 32 * kReplayWarps, 0, s>>>(p);
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
