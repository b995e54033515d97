// planar.cu -- the planar front / back end around K1 / K2.
//
// NONE of this exists in the reference: erkinov-wtf/dct starts from an 8-bit grayscale plane whose
// sides are multiples of the block size (create_block_from_pixels, src/dct.c:109-120, reads out of
// bounds otherwise) and knows nothing of colour.  BASELINE config 3 presupposes Y / Cb / Cr planes,
// so the conventions are fixed HERE (the tests hold a second, independent CPU statement of them):
//
//   colour     JFIF (ITU-T T.871) full-range BT.601 in 16-bit fixed point, integer arithmetic only:
//                Y  = ( 19595 R + 38470 G +  7471 B + 32768) >> 16
//                Cb = (-11059 R - 21709 G + 32768 B + (128 << 16) + 32767) >> 16
//                Cr = ( 32768 R - 27439 G -  5329 B + (128 << 16) + 32767) >> 16
//                R  = clamp(Y + (( 91881 (Cr-128)                  + 32768) >> 16))
//                G  = clamp(Y + ((-22554 (Cb-128) - 46802 (Cr-128) + 32768) >> 16))   (>> = floor)
//                B  = clamp(Y + ((116130 (Cb-128)                  + 32768) >> 16))
//   4:2:0      one chroma sample per 2x2 pixels, computed from the SUM of the four pixels' R, G, B
//              (one rounding: the Cb / Cr formulas with the bias and the shift scaled by 4, i.e.
//              `+ (128 << 18) + (1 << 17) - 1) >> 18`); decoding replicates each chroma sample 2x2.
//   edges      a plane is completed to whole blocks by replicating its last column / row (the JPEG
//              convention); odd widths / heights replicate the last pixel into the 2x2 cell.
//
// All three kernels are streaming byte kernels: every input byte is read once and every output byte
// written once with 16-byte accesses (a warp covers 1536 contiguous bytes of an RGB row), so the
// bound is HBM: 4.5 B per pixel for the colour conversions (3 in + 1.5 out, or the reverse).
#include "kernels.cuh"

namespace dctb {
namespace {

constexpr int kTX = 32, kTY = 8;   // CTA = 32 x 8 threads; a thread owns 16 pixels x 2 rows

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return __byte_perm(w, 0u, 0x4440 + i); }

// ---- one pixel / one 2x2 cell, integer arithmetic as specified above --------------------------------
__device__ __forceinline__ uint32_t luma_q16(uint32_t r, uint32_t g, uint32_t b)
{
    return r * 19595u + g * 38470u + b * 7471u + 32768u;   // < 2^24: Y is byte 2
}
// sums of a 2x2 cell (each <= 1020).  The result is scaled by 2^6 on top of the 2^18 so that the
// sample is byte 3 of the 32-bit word; unsigned wrap-around is harmless, the final value is < 2^32.
__device__ __forceinline__ uint32_t cb_q24(uint32_t sr, uint32_t sg, uint32_t sb)
{
    return sb * (32768u << 6) - sr * (11059u << 6) - sg * (21709u << 6) + (((128u << 18) + (1u << 17) - 1u) << 6);
}
__device__ __forceinline__ uint32_t cr_q24(uint32_t sr, uint32_t sg, uint32_t sb)
{
    return sr * (32768u << 6) - sg * (27439u << 6) - sb * (5329u << 6) + (((128u << 18) + (1u << 17) - 1u) << 6);
}
__device__ __forceinline__ int clamp255(int v) { return max(0, min(v, 255)); }

// bytes 2 of four words -> one word / bytes 3 of four words -> one word
__device__ __forceinline__ uint32_t pack_b2(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return __byte_perm(__byte_perm(a, b, 0x0062), __byte_perm(c, d, 0x0062), 0x5410);
}
__device__ __forceinline__ uint32_t pack_b3(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return __byte_perm(__byte_perm(a, b, 0x0073), __byte_perm(c, d, 0x0073), 0x5410);
}
// channel c of pixel p (0..15) of a 48-byte group held in 12 words
template <int P, int C> __device__ __forceinline__ uint32_t channel(const uint32_t (&w)[12])
{
    constexpr int b = 3 * P + C;
    return byte_of(w[b >> 2], b & 3);
}

// =====================================================================================================
// RGB (interleaved) -> Y, Cb, Cr planes (4:2:0), planes completed to their padded sizes on the fly
// =====================================================================================================
__global__ void __launch_bounds__(kTX * kTY) k_rgb_to_ycbcr420(PlanarParams p)
{
    const int x0 = (blockIdx.x * kTX + threadIdx.x) * 16;
    const int y0 = (blockIdx.y * kTY + threadIdx.y) * 2;
    const bool fast = p.vec_ok && x0 + 16 <= p.W && y0 + 2 <= p.H;
    if (fast) {
        uint32_t a[12], b[12];
        const uint4 *ra = reinterpret_cast<const uint4 *>(p.rgb + (size_t)y0 * p.rgb_pitch + (size_t)x0 * 3);
        const uint4 *rb = reinterpret_cast<const uint4 *>(p.rgb + (size_t)(y0 + 1) * p.rgb_pitch + (size_t)x0 * 3);
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            const uint4 ua = __ldg(ra + v), ub = __ldg(rb + v);
            a[4 * v] = ua.x, a[4 * v + 1] = ua.y, a[4 * v + 2] = ua.z, a[4 * v + 3] = ua.w;
            b[4 * v] = ub.x, b[4 * v + 1] = ub.y, b[4 * v + 2] = ub.z, b[4 * v + 3] = ub.w;
        }
        uint32_t ya[16], yb[16], cb[8], cr[8];
        static_for<0, 8>([&](auto C) {
            constexpr int c = decltype(C)::value, p0 = 2 * c, p1 = 2 * c + 1;
            const uint32_t r00 = channel<p0, 0>(a), g00 = channel<p0, 1>(a), b00 = channel<p0, 2>(a);
            const uint32_t r01 = channel<p1, 0>(a), g01 = channel<p1, 1>(a), b01 = channel<p1, 2>(a);
            const uint32_t r10 = channel<p0, 0>(b), g10 = channel<p0, 1>(b), b10 = channel<p0, 2>(b);
            const uint32_t r11 = channel<p1, 0>(b), g11 = channel<p1, 1>(b), b11 = channel<p1, 2>(b);
            ya[p0] = luma_q16(r00, g00, b00), ya[p1] = luma_q16(r01, g01, b01);
            yb[p0] = luma_q16(r10, g10, b10), yb[p1] = luma_q16(r11, g11, b11);
            const uint32_t sr = (r00 + r01) + (r10 + r11), sg = (g00 + g01) + (g10 + g11), sb = (b00 + b01) + (b10 + b11);
            cb[c] = cb_q24(sr, sg, sb), cr[c] = cr_q24(sr, sg, sb);
        });
        uint4 oa, ob;
        oa.x = pack_b2(ya[0], ya[1], ya[2], ya[3]), oa.y = pack_b2(ya[4], ya[5], ya[6], ya[7]);
        oa.z = pack_b2(ya[8], ya[9], ya[10], ya[11]), oa.w = pack_b2(ya[12], ya[13], ya[14], ya[15]);
        ob.x = pack_b2(yb[0], yb[1], yb[2], yb[3]), ob.y = pack_b2(yb[4], yb[5], yb[6], yb[7]);
        ob.z = pack_b2(yb[8], yb[9], yb[10], yb[11]), ob.w = pack_b2(yb[12], yb[13], yb[14], yb[15]);
        *reinterpret_cast<uint4 *>(p.y + (size_t)y0 * p.y_pitch + x0) = oa;
        *reinterpret_cast<uint4 *>(p.y + (size_t)(y0 + 1) * p.y_pitch + x0) = ob;
        const size_t co = (size_t)(y0 >> 1) * p.c_pitch + (x0 >> 1);
        *reinterpret_cast<uint2 *>(p.cb + co) = make_uint2(pack_b3(cb[0], cb[1], cb[2], cb[3]), pack_b3(cb[4], cb[5], cb[6], cb[7]));
        *reinterpret_cast<uint2 *>(p.cr + co) = make_uint2(pack_b3(cr[0], cr[1], cr[2], cr[3]), pack_b3(cr[4], cr[5], cr[6], cr[7]));
        return;
    }
    // edge threads (and unaligned planes): one sample at a time, source coordinates clamped
    auto src = [&](int x, int y, int c) -> uint32_t {
        return p.rgb[(size_t)min(y, p.H - 1) * p.rgb_pitch + (size_t)min(x, p.W - 1) * 3 + c];
    };
    for (int dy = 0; dy < 2; ++dy) {
        const int y = y0 + dy;
        if (y >= p.y_h) break;
        for (int dx = 0; dx < 16; ++dx) {
            const int x = x0 + dx;
            if (x >= p.y_w) break;
            p.y[(size_t)y * p.y_pitch + x] = (uint8_t)(luma_q16(src(x, y, 0), src(x, y, 1), src(x, y, 2)) >> 16);
        }
    }
    const int cy = y0 >> 1, cw_img = (p.W + 1) >> 1, ch_img = (p.H + 1) >> 1;
    if (cy < p.c_h) {
        const int sy = 2 * min(cy, ch_img - 1);
        for (int dx = 0; dx < 8; ++dx) {
            const int cx = (x0 >> 1) + dx;
            if (cx >= p.c_w) break;
            const int sx = 2 * min(cx, cw_img - 1);
            uint32_t s[3];
            for (int c = 0; c < 3; ++c) s[c] = src(sx, sy, c) + src(sx + 1, sy, c) + src(sx, sy + 1, c) + src(sx + 1, sy + 1, c);
            p.cb[(size_t)cy * p.c_pitch + cx] = (uint8_t)(cb_q24(s[0], s[1], s[2]) >> 24);
            p.cr[(size_t)cy * p.c_pitch + cx] = (uint8_t)(cr_q24(s[0], s[1], s[2]) >> 24);
        }
    }
}

// =====================================================================================================
// Y, Cb, Cr planes (4:2:0) -> RGB (interleaved); only the W x H image is written
// =====================================================================================================
struct ChromaTerms {
    int r, g, b;
};
__device__ __forceinline__ ChromaTerms chroma_terms(int cb, int cr)
{
    cb -= 128, cr -= 128;
    ChromaTerms t;
    t.r = (91881 * cr + 32768) >> 16;
    t.g = (-22554 * cb - 46802 * cr + 32768) >> 16;
    t.b = (116130 * cb + 32768) >> 16;
    return t;
}

__global__ void __launch_bounds__(kTX * kTY) k_ycbcr420_to_rgb(PlanarParams p)
{
    __shared__ uint4 stage[kTY][2 * 96];                        // per warp: two rows of 32 x 48 bytes
    const int x0 = (blockIdx.x * kTX + threadIdx.x) * 16;
    const int y0 = (blockIdx.y * kTY + threadIdx.y) * 2;
    const bool fast = p.vec_ok && x0 + 16 <= p.W && y0 + 2 <= p.H;
    const bool warp_fast = __all_sync(0xffffffffu, fast);       // before any thread leaves
    if (x0 >= p.W || y0 >= p.H) return;
    if (fast) {
        const uint4 ya = __ldg(reinterpret_cast<const uint4 *>(p.y + (size_t)y0 * p.y_pitch + x0));
        const uint4 yb = __ldg(reinterpret_cast<const uint4 *>(p.y + (size_t)(y0 + 1) * p.y_pitch + x0));
        const size_t co = (size_t)(y0 >> 1) * p.c_pitch + (x0 >> 1);
        const uint2 cbv = __ldg(reinterpret_cast<const uint2 *>(p.cb + co));
        const uint2 crv = __ldg(reinterpret_cast<const uint2 *>(p.cr + co));
        const uint32_t yaw[4] = {ya.x, ya.y, ya.z, ya.w}, ybw[4] = {yb.x, yb.y, yb.z, yb.w};
        const uint32_t cbw[2] = {cbv.x, cbv.y}, crw[2] = {crv.x, crv.y};
        // Packed int16 arithmetic: the two pixels of a 2x2 cell's row share their chroma terms, so one VIADD.16x2 adds
        // the term to both luma samples and one VIMNMX.S16x2.RELU clamps both to [0, 255] (three scalar operations per
        // component before).  Per cell and row: r2 = (R0, R1), g2, b2 as int16 pairs -> the six bytes R0 G0 B0 R1 G1 B1.
        uint32_t wa[12], wb[12];                   // the 48 output bytes of each row
        static_for<0, 4>([&](auto Q4) {
            constexpr int qd = decltype(Q4)::value;            // four pixels = two cells = three output words
            uint32_t rga[2], ba[2], rgb_[2], bb[2];            // per cell: (R0 G0 R1 G1) and (B0 . B1 .) bytes, rows a and b
            static_for<0, 2>([&](auto H2) {
                constexpr int h = decltype(H2)::value, c = 2 * qd + h;                 // chroma sample / cell index
                const ChromaTerms t = chroma_terms((int)byte_of(cbw[c >> 2], c & 3), (int)byte_of(crw[c >> 2], c & 3));
                const uint32_t tr = __byte_perm((uint32_t)t.r, 0u, 0x1010), tg = __byte_perm((uint32_t)t.g, 0u, 0x1010),
                               tb = __byte_perm((uint32_t)t.b, 0u, 0x1010);           // (term, term) as int16 pairs
                // luma pair of this cell: bytes (2c, 2c+1) of the row -> zero-extended int16 pair
                constexpr int sel = ((2 * c) & 3) == 0 ? 0x4140 : 0x4342;
                const uint32_t la = __byte_perm(yaw[(2 * c) >> 2], 0u, sel), lb = __byte_perm(ybw[(2 * c) >> 2], 0u, sel);
                auto add_clamp = [](uint32_t l2, uint32_t t2) {
                    uint32_t r;
                    asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(l2), "r"(t2));
                    asm("min.s16x2.relu %0, %0, %1;" : "+r"(r) : "r"(0x00ff00ffu));
                    return r;
                };
                const uint32_t ra2 = add_clamp(la, tr), ga2 = add_clamp(la, tg), ba2 = add_clamp(la, tb);
                const uint32_t rb2 = add_clamp(lb, tr), gb2 = add_clamp(lb, tg), bb2 = add_clamp(lb, tb);
                rga[h] = __byte_perm(ra2, ga2, 0x6240);        // R0 G0 R1 G1
                rgb_[h] = __byte_perm(rb2, gb2, 0x6240);
                ba[h] = ba2, bb[h] = bb2;                      // B0 at byte 0, B1 at byte 2
            });
            // bytes: R0 G0 B0 R1 | G1 B1 R0' G0' | B0' R1' G1' B1'
            wa[3 * qd] = __byte_perm(rga[0], ba[0], 0x2410);
            wa[3 * qd + 1] = __byte_perm(__byte_perm(rga[0], ba[0], 0x0063), rga[1], 0x5410);
            wa[3 * qd + 2] = __byte_perm(rga[1], ba[1], 0x6324);
            wb[3 * qd] = __byte_perm(rgb_[0], bb[0], 0x2410);
            wb[3 * qd + 1] = __byte_perm(__byte_perm(rgb_[0], bb[0], 0x0063), rgb_[1], 0x5410);
            wb[3 * qd + 2] = __byte_perm(rgb_[1], bb[1], 0x6324);
        });
        uint4 ua[3], ub[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            ua[v] = make_uint4(wa[4 * v], wa[4 * v + 1], wa[4 * v + 2], wa[4 * v + 3]);
            ub[v] = make_uint4(wb[4 * v], wb[4 * v + 1], wb[4 * v + 2], wb[4 * v + 3]);
        }
        // A thread's 48 bytes are 48 bytes apart from its neighbour's: written directly, every STG.128 of the
        // warp would touch 32 half-filled sectors.  When the whole warp is on this path the two rows go through
        // a per-warp stage instead (48-byte lane stride = conflict-free STS.128) and leave as 512 contiguous
        // bytes per store instruction.
        uint4 *da = reinterpret_cast<uint4 *>(p.rgb_out + (size_t)y0 * p.rgb_pitch + (size_t)x0 * 3);
        uint4 *db = reinterpret_cast<uint4 *>(p.rgb_out + (size_t)(y0 + 1) * p.rgb_pitch + (size_t)x0 * 3);
        if (warp_fast) {
            uint4 *st = stage[threadIdx.y];
            const int lane = threadIdx.x;
#pragma unroll
            for (int v = 0; v < 3; ++v) st[lane * 3 + v] = ua[v], st[96 + lane * 3 + v] = ub[v];
            __syncwarp();
            uint4 *wa = da - lane * 3, *wb = db - lane * 3;       // the warp's first byte of each row
#pragma unroll
            for (int v = 0; v < 3; ++v) wa[v * 32 + lane] = st[v * 32 + lane], wb[v * 32 + lane] = st[96 + v * 32 + lane];
        } else {
#pragma unroll
            for (int v = 0; v < 3; ++v) da[v] = ua[v], db[v] = ub[v];
        }
        return;
    }
    for (int dy = 0; dy < 2; ++dy) {
        const int y = y0 + dy;
        if (y >= p.H) break;
        for (int dx = 0; dx < 16; ++dx) {
            const int x = x0 + dx;
            if (x >= p.W) break;
            const size_t co = (size_t)(y >> 1) * p.c_pitch + (x >> 1);
            const ChromaTerms t = chroma_terms(p.cb[co], p.cr[co]);
            const int l = p.y[(size_t)y * p.y_pitch + x];
            uint8_t *o = p.rgb_out + (size_t)y * p.rgb_pitch + (size_t)x * 3;
            o[0] = (uint8_t)clamp255(l + t.r), o[1] = (uint8_t)clamp255(l + t.g), o[2] = (uint8_t)clamp255(l + t.b);
        }
    }
}

// =====================================================================================================
// complete a W x H plane (elements of `elem` bytes) to W_pad x H_pad in place by edge replication.
// Only the padding is touched: one thread per padded element.
// =====================================================================================================
__global__ void k_pad_edges(uint8_t *px, long long pitch, int W, int H, int Wp, int Hp, int elem)
{
    const int right = Wp - W;                                   // columns to add to each of the H image rows
    const long long n_right = (long long)right * H, n_bottom = (long long)Wp * (Hp - H);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_right + n_bottom;
         i += (long long)gridDim.x * blockDim.x) {
        int x, y;
        if (i < n_right) {
            y = (int)(i / right), x = W + (int)(i % right);
        } else {
            const long long j = i - n_right;
            y = H + (int)(j / Wp), x = (int)(j % Wp);
        }
        const uint8_t *s = px + (size_t)min(y, H - 1) * pitch + (size_t)min(x, W - 1) * elem;
        uint8_t *d = px + (size_t)y * pitch + (size_t)x * elem;
        for (int b = 0; b < elem; ++b) d[b] = s[b];
    }
}

}  // namespace

static dim3 planar_grid(int cover_w, int cover_h)
{
    return dim3((unsigned)((cover_w + 16 * kTX - 1) / (16 * kTX)), (unsigned)((cover_h + 2 * kTY - 1) / (2 * kTY)));
}

cudaError_t launch_rgb_to_ycbcr420(const PlanarParams &p, cudaStream_t s)
{
    const int cover_w = max(p.y_w, 2 * p.c_w), cover_h = max(p.y_h, 2 * p.c_h);
    if (cover_w <= 0 || cover_h <= 0) return cudaSuccess;
    k_rgb_to_ycbcr420<<<planar_grid(cover_w, cover_h), dim3(kTX, kTY), 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_ycbcr420_to_rgb(const PlanarParams &p, cudaStream_t s)
{
    if (p.W <= 0 || p.H <= 0) return cudaSuccess;
    k_ycbcr420_to_rgb<<<planar_grid(p.W, p.H), dim3(kTX, kTY), 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_pad_edges(uint8_t *px, long long pitch, int W, int H, int Wp, int Hp, int elem, cudaStream_t s)
{
    const long long n = (long long)(Wp - W) * H + (long long)Wp * (Hp - H);
    if (n <= 0) return cudaSuccess;
    const int threads = 256;
    const long long blocks = (n + threads - 1) / threads;
    k_pad_edges<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), threads, 0, s>>>(px, pitch, W, H, Wp, Hp, elem);
    return cudaGetLastError();
}

}  // namespace dctb
