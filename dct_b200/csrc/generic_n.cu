// generic_n.cu -- K6: the plane calls for block sizes other than 8.
//
// The reference accepts any block_size (include/dct.h:34, include/quantization.h:34; its custom
// distance-based quantisation table for n != 8 is src/quantization.c:78-96).  There is no fast
// factorisation here: every block is computed in the reference's own operation order in
// non-contracted fp64 (src/dct.c:57-74 / :85-102, src/quantization.c:124 / :144), one thread per
// block element, so the results are bit-identical by construction.  Signature completeness, not a
// roofline path.  Records are n*n int16, block-major; ZIGZAG follows src/entropy.c:158-178 for n.
#include "kernels.cuh"

namespace dctb {

namespace {

__device__ __forceinline__ double round_half_away_g(double y)
{
    const double t = trunc(y);
    return (fabs(__dsub_rn(y, t)) >= 0.5) ? __dadd_rn(t, copysign(1.0, y)) : t;
}

__device__ __forceinline__ bool near_half_g(double a)
{
    a = fabs(a);
    const double f = __dsub_rn(a, floor(a));
    return fabs(__dsub_rn(f, 0.5)) <= 1e-9;
}

// src/quantization.c:186
__device__ __forceinline__ double norm_variance_g(double variance)
{
    return fmin(1.0, fmax(0.1, __ddiv_rn(variance, 1000.0)));
}

// shared memory: D[nn] M[nn] | pos[nn] (int) | per sub-block: X[nn] T[nn] | isum[bpc] isq[bpc]
template <bool FORWARD>
__global__ void k_generic_plane(const GenericParams p)
{
    extern __shared__ double gsm[];
    const int n = p.n, nn = n * n, bpc = p.blocks_per_cta;
    double *sD = gsm, *sM = gsm + nn;
    double *sX = gsm + 2 * nn, *sT = sX + (size_t)bpc * nn;
    int *sPos = reinterpret_cast<int *>(sT + (size_t)bpc * nn);
    int *sSum = sPos + nn, *sSq = sSum + bpc;

    const int tid = threadIdx.x, sub = tid / nn, e = tid - sub * nn, i = e / n, j = e - i * n;
    for (int t = tid; t < nn; t += blockDim.x) {
        sD[t] = p.D[t];
        sM[t] = FORWARD ? p.Q[t] : p.R[t];
        sPos[t] = p.layout == LAYOUT_ZIGZAG ? p.pos_of_natural[t] : t;
    }
    __syncthreads();
    double *X = sX + (size_t)sub * nn, *T = sT + (size_t)sub * nn;
    unsigned ties = 0, sat = 0;

    for (uint32_t base = blockIdx.x * bpc; base < p.nblocks; base += gridDim.x * bpc) {
        const uint32_t b = base + sub;
        const bool active = b < p.nblocks;
        const uint32_t bb = active ? b : 0;
        const uint32_t by = bb / p.bw, bx = bb - by * p.bw;
        double var = 0.0;
        if (FORWARD) {
            const int px = (int)p.px_in[((long long)by * n + i) * p.pitch + (long long)bx * n + j] - 128;
            X[e] = (double)px;                            // (double)px - 128.0, src/dct.c:115
            if (p.adaptive) {
                if (e == 0) sSum[sub] = 0, sSq[sub] = 0;
                __syncthreads();
                atomicAdd(&sSum[sub], px);                // exact integers: order does not matter
                atomicAdd(&sSq[sub], px * px);
                __syncthreads();
                const double count = (double)nn;          // src/quantization.c:153-169
                const double mean = __ddiv_rn((double)sSum[sub], count);
                var = __dsub_rn(__ddiv_rn((double)sSq[sub], count), __dmul_rn(mean, mean));
                if (e == 0 && active && p.var_out) p.var_out[b] = var;
            }
        } else {
            if (p.adaptive) var = p.var_in ? p.var_in[bb] : 0.0;
            const int q = active ? (int)p.coef_in[(size_t)bb * nn + sPos[e]] : 0;
            double m = sM[e];
            double val;
            if (p.adaptive) {                              // src/quantization.c:133-151, :171-211
                if (e != 0) m = __dmul_rn(m, __ddiv_rn(1.0, __dsub_rn(2.0, norm_variance_g(var))));
                val = __dmul_rn((double)q, __ddiv_rn(1.0, m));
            } else {
                val = __dmul_rn((double)q, m);
            }
            X[e] = val;
        }
        __syncthreads();
        double acc = 0.0;
        if (FORWARD) {       // temp[i][j] = sum_k X[i][k] * D[j][k]        (src/dct.c:57-64)
            for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(X[i * n + k], sD[j * n + k]));
        } else {             // temp[i][j] = sum_k D[k][i] * in[k][j]       (src/dct.c:85-92)
            for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(sD[k * n + i], X[k * n + j]));
        }
        T[e] = acc;
        __syncthreads();
        acc = 0.0;
        if (FORWARD) {       // out[i][j] = sum_k D[i][k] * temp[k][j]      (src/dct.c:67-74)
            for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(sD[i * n + k], T[k * n + j]));
            double m = sM[e];
            if (p.adaptive && e != 0) {
                m = __dmul_rn(m, __dsub_rn(2.0, norm_variance_g(var)));
                if (m < 1.0) m = 1.0;
            }
            const double y = __ddiv_rn(acc, m);
            const double r = round_half_away_g(y);
            int q = (int)r;
            if (r > 32767.0) q = 32767, ++sat;
            if (r < -32768.0) q = -32768, ++sat;
            if (active) {
                ties += near_half_g(y);
                p.coef_out[(size_t)b * nn + sPos[e]] = (int16_t)q;
            }
        } else {             // out[i][j] = sum_k temp[i][k] * D[k][j]      (src/dct.c:95-102)
            for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __dmul_rn(T[i * n + k], sD[k * n + j]));
            const double v = __dadd_rn(acc, 128.0);
            double r = round_half_away_g(v);
            r = r < 0.0 ? 0.0 : (r > 255.0 ? 255.0 : r);
            if (active) {
                ties += near_half_g(v);
                p.px_out[((long long)by * n + i) * p.pitch + (long long)bx * n + j] = (uint8_t)r;
            }
        }
        __syncthreads();
    }
    if (ties) atomicAdd(&p.ctr->near_ties, (unsigned long long)ties);
    if (sat) atomicAdd(&p.ctr->saturated, (unsigned long long)sat);
    if (tid == 0 && blockIdx.x == 0) atomicAdd(&p.ctr->replayed, (unsigned long long)p.nblocks);
}

}  // namespace

cudaError_t launch_generic_plane(const GenericParams &p, int forward, cudaStream_t s)
{
    if (p.nblocks == 0) return cudaSuccess;
    const int nn = p.n * p.n, bpc = p.blocks_per_cta;
    const size_t smem = (size_t)(2 * nn + 2 * bpc * nn) * sizeof(double) + (size_t)(nn + 2 * bpc) * sizeof(int);
    unsigned ctas = (p.nblocks + bpc - 1) / bpc;
    if (ctas > 148u * 16u) ctas = 148u * 16u;
    if (forward) k_generic_plane<true><<<ctas, bpc * nn, smem, s>>>(p);
    else k_generic_plane<false><<<ctas, bpc * nn, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace dctb
