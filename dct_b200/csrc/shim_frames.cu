// shim_frames.cu -- C ABI: planar front / back end (colour conversion, 4:2:0, edge completion, whole RGB frames).
#include "plan.cuh"

using namespace dctb;
using namespace dctb::shim;

// ------------------------------------------------------------------------------------------
// planar front / back end (planar.cu): colour conversion + 4:2:0, edge completion.  Not in the reference.
// ------------------------------------------------------------------------------------------
static int round_up(int v, int m) { return (v + m - 1) / m * m; }

extern "C" void dct_cuda_frame420_geometry(int width, int height, dct_cuda_frame420 *g)
{
    if (!g) return;
    g->width = width, g->height = height;
    g->y_width = round_up(width, 8), g->y_height = round_up(height, 8);
    g->c_width = round_up((width + 1) / 2, 8), g->c_height = round_up((height + 1) / 2, 8);
}

static int check_frame(const dct_cuda_frame420 *g, int device)
{
    if (!g) return fail(DCT_CUDA_EINVAL, "NULL geometry");
    dct_cuda_frame420 want;
    dct_cuda_frame420_geometry(g->width, g->height, &want);
    if (g->width <= 0 || g->height <= 0 || memcmp(&want, g, sizeof want))
        return fail(DCT_CUDA_EINVAL, "geometry does not come from dct_cuda_frame420_geometry (%dx%d)", g->width, g->height);
    const int ndev = dct_cuda_device_count();
    if (ndev <= 0) return fail(DCT_CUDA_ENODEV, "no CUDA device available (libdct_cuda has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(DCT_CUDA_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
    return DCT_CUDA_OK;
}

static PlanarParams planar_params(const dct_cuda_frame420 *g, const uint8_t *rgb_in, uint8_t *rgb_out, size_t rgb_pitch,
                                  const uint8_t *y, size_t y_pitch, const uint8_t *cb, const uint8_t *cr, size_t c_pitch)
{
    PlanarParams pp{};
    pp.rgb = rgb_in, pp.rgb_out = rgb_out, pp.rgb_pitch = (long long)rgb_pitch;
    pp.W = g->width, pp.H = g->height;
    pp.y = const_cast<uint8_t *>(y), pp.cb = const_cast<uint8_t *>(cb), pp.cr = const_cast<uint8_t *>(cr);
    pp.y_pitch = (long long)y_pitch, pp.c_pitch = (long long)c_pitch;
    pp.y_w = g->y_width, pp.y_h = g->y_height, pp.c_w = g->c_width, pp.c_h = g->c_height;
    const uintptr_t rgbp = (uintptr_t)(rgb_in ? rgb_in : rgb_out);
    pp.vec_ok = !(rgbp % 16) && !(rgb_pitch % 16) && !((uintptr_t)y % 16) && !(y_pitch % 16) && !((uintptr_t)cb % 8) &&
                !((uintptr_t)cr % 8) && !(c_pitch % 8);
    return pp;
}

extern "C" int dct_cuda_rgb_to_ycbcr420_dev(int device, const uint8_t *d_rgb, size_t rgb_pitch, const dct_cuda_frame420 *g,
                                            uint8_t *d_y, size_t y_pitch, uint8_t *d_cb, uint8_t *d_cr, size_t c_pitch,
                                            void *stream)
{
    int rc = check_frame(g, device);
    if (rc) return rc;
    if (!d_rgb || !d_y || !d_cb || !d_cr) return fail(DCT_CUDA_EINVAL, "NULL data pointer");
    if (rgb_pitch < (size_t)g->width * 3 || y_pitch < (size_t)g->y_width || c_pitch < (size_t)g->c_width)
        return fail(DCT_CUDA_EINVAL, "pitch smaller than a row");
    DeviceGuard dg(device);
    CU_TRY(launch_rgb_to_ycbcr420(planar_params(g, d_rgb, nullptr, rgb_pitch, d_y, y_pitch, d_cb, d_cr, c_pitch),
                                  (cudaStream_t)stream));
    return DCT_CUDA_OK;
}

extern "C" int dct_cuda_ycbcr420_to_rgb_dev(int device, const uint8_t *d_y, size_t y_pitch, const uint8_t *d_cb,
                                            const uint8_t *d_cr, size_t c_pitch, const dct_cuda_frame420 *g, uint8_t *d_rgb,
                                            size_t rgb_pitch, void *stream)
{
    int rc = check_frame(g, device);
    if (rc) return rc;
    if (!d_rgb || !d_y || !d_cb || !d_cr) return fail(DCT_CUDA_EINVAL, "NULL data pointer");
    if (rgb_pitch < (size_t)g->width * 3 || y_pitch < (size_t)g->y_width || c_pitch < (size_t)g->c_width)
        return fail(DCT_CUDA_EINVAL, "pitch smaller than a row");
    DeviceGuard dg(device);
    CU_TRY(launch_ycbcr420_to_rgb(planar_params(g, nullptr, d_rgb, rgb_pitch, d_y, y_pitch, d_cb, d_cr, c_pitch),
                                  (cudaStream_t)stream));
    return DCT_CUDA_OK;
}

extern "C" int dct_cuda_pad_edges_dev(int device, uint8_t *d_px, size_t pitch, int W, int H, int W_pad, int H_pad,
                                      void *stream)
{
    if (!d_px) return fail(DCT_CUDA_EINVAL, "NULL data pointer");
    if (W <= 0 || H <= 0 || W_pad < W || H_pad < H || pitch < (size_t)W_pad)
        return fail(DCT_CUDA_EINVAL, "bad sizes %dx%d -> %dx%d, pitch %zu", W, H, W_pad, H_pad, pitch);
    const int ndev = dct_cuda_device_count();
    if (ndev <= 0) return fail(DCT_CUDA_ENODEV, "no CUDA device available (libdct_cuda has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(DCT_CUDA_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
    DeviceGuard dg(device);
    CU_TRY(launch_pad_edges(d_px, (long long)pitch, W, H, W_pad, H_pad, 1, (cudaStream_t)stream));
    return DCT_CUDA_OK;
}

// whole RGB frames from / to host memory: H2D, colour conversion, K1 on the three planes, D2H (and back)
namespace {
struct FrameLayout {
    dct_cuda_frame420 g;
    size_t rgb_pitch, y_pitch, c_pitch;
    size_t off_y, off_cb, off_cr, off_ky, off_kcb, off_kcr, bytes;   // offsets into plan->d_frame
    size_t ny, nc;                                                    // samples per luma / chroma plane
};

FrameLayout frame_layout(int W, int H)
{
    FrameLayout f{};
    dct_cuda_frame420_geometry(W, H, &f.g);
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    f.rgb_pitch = ((size_t)W * 3 + 15) / 16 * 16;
    f.y_pitch = ((size_t)f.g.y_width + 15) / 16 * 16;
    f.c_pitch = (size_t)f.g.c_width;
    f.ny = (size_t)f.g.y_width * f.g.y_height, f.nc = (size_t)f.g.c_width * f.g.c_height;
    size_t o = up(f.rgb_pitch * H);
    f.off_y = o, o += up(f.y_pitch * f.g.y_height);
    f.off_cb = o, o += up(f.c_pitch * f.g.c_height);
    f.off_cr = o, o += up(f.c_pitch * f.g.c_height);
    f.off_ky = o, o += up(f.ny * 2);
    f.off_kcb = o, o += up(f.nc * 2);
    f.off_kcr = o, o += up(f.nc * 2);
    f.bytes = o;
    return f;
}

int frame_prepare(dct_cuda_plan *luma, dct_cuda_plan *chroma, int W, int H, int layout, FrameLayout *f)
{
    if (!luma || !chroma) return fail(DCT_CUDA_EINVAL, "NULL plan");
    if (luma->n != 8 || chroma->n != 8 || luma->adaptive || chroma->adaptive)
        return fail(DCT_CUDA_EINVAL, "the RGB 4:2:0 frame calls need two non-adaptive 8x8 plans");
    if (luma->device != chroma->device) return fail(DCT_CUDA_EINVAL, "both plans must live on the same GPU");
    if (W <= 0 || H <= 0) return fail(DCT_CUDA_EINVAL, "width and height must be positive (got %dx%d)", W, H);
    if (layout != DCT_CUDA_NATURAL && layout != DCT_CUDA_ZIGZAG) return fail(DCT_CUDA_EINVAL, "bad layout %d", layout);
    *f = frame_layout(W, H);
    std::lock_guard<std::mutex> plan_lock(luma->mu);
    if (luma->frame_cap < f->bytes) {
        CU_TRY(cudaDeviceSynchronize());
        if (luma->d_frame) cudaFree(luma->d_frame);
        luma->d_frame = nullptr, luma->frame_cap = 0;
        CU_TRY(cudaMalloc(&luma->d_frame, f->bytes));
        luma->frame_cap = f->bytes;
    }
    return DCT_CUDA_OK;
}

int frame_stats(dct_cuda_plan *luma, dct_cuda_plan *chroma, cudaStream_t s, dct_cuda_stats *stats)
{
    dct_cuda_stats a{}, b{};
    int rc = collect_stats(luma, &a, s);
    if (rc) return rc;
    if (chroma != luma && (rc = collect_stats(chroma, &b, s))) return rc;
    if (stats) {
        stats->blocks = a.blocks + b.blocks, stats->replayed_blocks = a.replayed_blocks + b.replayed_blocks;
        stats->near_ties = a.near_ties + b.near_ties, stats->saturated = a.saturated + b.saturated;
    }
    return DCT_CUDA_OK;
}
}  // namespace

extern "C" int dct_cuda_encode_rgb420(dct_cuda_plan *luma, dct_cuda_plan *chroma, const uint8_t *rgb, size_t rgb_pitch,
                                      int W, int H, int16_t *coef_y, int16_t *coef_cb, int16_t *coef_cr, int layout,
                                      dct_cuda_stats *stats)
{
    FrameLayout f;
    int rc = frame_prepare(luma, chroma, W, H, layout, &f);
    if (rc) return rc;
    if (!rgb || !coef_y || !coef_cb || !coef_cr) return fail(DCT_CUDA_EINVAL, "NULL data pointer");
    if (rgb_pitch < (size_t)W * 3) return fail(DCT_CUDA_EINVAL, "pitch %zu must be >= 3 * width", rgb_pitch);
    DeviceGuard dg(luma->device);
    std::unique_lock<std::mutex> lock_l(luma->mu, std::defer_lock), lock_c(chroma->mu, std::defer_lock);
    if (chroma != luma) std::lock(lock_l, lock_c);
    else lock_l.lock();
    uint8_t *base = luma->d_frame;
    cudaStream_t s = luma->lane[0].stream;
    CU_TRY(cudaMemcpy2DAsync(base, f.rgb_pitch, rgb, rgb_pitch, (size_t)W * 3, (size_t)H, cudaMemcpyHostToDevice, s));
    CU_TRY(launch_rgb_to_ycbcr420(planar_params(&f.g, base, nullptr, f.rgb_pitch, base + f.off_y, f.y_pitch, base + f.off_cb,
                                                base + f.off_cr, f.c_pitch), s));
    int16_t *ky = (int16_t *)(base + f.off_ky), *kcb = (int16_t *)(base + f.off_kcb), *kcr = (int16_t *)(base + f.off_kcr);
    if ((rc = queue_fwd(luma, luma->lane[0], base + f.off_y, f.y_pitch, f.g.y_width, f.g.y_height, ky, layout, nullptr, s))) return rc;
    if ((rc = queue_fwd(chroma, chroma->lane[0], base + f.off_cb, f.c_pitch, f.g.c_width, f.g.c_height, kcb, layout, nullptr, s))) return rc;
    if ((rc = queue_fwd(chroma, chroma->lane[0], base + f.off_cr, f.c_pitch, f.g.c_width, f.g.c_height, kcr, layout, nullptr, s))) return rc;
    CU_TRY(cudaMemcpyAsync(coef_y, ky, f.ny * 2, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(coef_cb, kcb, f.nc * 2, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(coef_cr, kcr, f.nc * 2, cudaMemcpyDeviceToHost, s));
    return frame_stats(luma, chroma, s, stats);
}

extern "C" int dct_cuda_decode_rgb420(dct_cuda_plan *luma, dct_cuda_plan *chroma, const int16_t *coef_y,
                                      const int16_t *coef_cb, const int16_t *coef_cr, int W, int H, int layout,
                                      uint8_t *rgb, size_t rgb_pitch, dct_cuda_stats *stats)
{
    FrameLayout f;
    int rc = frame_prepare(luma, chroma, W, H, layout, &f);
    if (rc) return rc;
    if (!rgb || !coef_y || !coef_cb || !coef_cr) return fail(DCT_CUDA_EINVAL, "NULL data pointer");
    if (rgb_pitch < (size_t)W * 3) return fail(DCT_CUDA_EINVAL, "pitch %zu must be >= 3 * width", rgb_pitch);
    DeviceGuard dg(luma->device);
    std::unique_lock<std::mutex> lock_l(luma->mu, std::defer_lock), lock_c(chroma->mu, std::defer_lock);
    if (chroma != luma) std::lock(lock_l, lock_c);
    else lock_l.lock();
    uint8_t *base = luma->d_frame;
    cudaStream_t s = luma->lane[0].stream;
    int16_t *ky = (int16_t *)(base + f.off_ky), *kcb = (int16_t *)(base + f.off_kcb), *kcr = (int16_t *)(base + f.off_kcr);
    CU_TRY(cudaMemcpyAsync(ky, coef_y, f.ny * 2, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(kcb, coef_cb, f.nc * 2, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(kcr, coef_cr, f.nc * 2, cudaMemcpyHostToDevice, s));
    if ((rc = queue_inv(luma, luma->lane[0], ky, f.g.y_width, f.g.y_height, layout, nullptr, base + f.off_y, f.y_pitch, s))) return rc;
    if ((rc = queue_inv(chroma, chroma->lane[0], kcb, f.g.c_width, f.g.c_height, layout, nullptr, base + f.off_cb, f.c_pitch, s))) return rc;
    if ((rc = queue_inv(chroma, chroma->lane[0], kcr, f.g.c_width, f.g.c_height, layout, nullptr, base + f.off_cr, f.c_pitch, s))) return rc;
    CU_TRY(launch_ycbcr420_to_rgb(planar_params(&f.g, nullptr, base, f.rgb_pitch, base + f.off_y, f.y_pitch, base + f.off_cb,
                                                base + f.off_cr, f.c_pitch), s));
    CU_TRY(cudaMemcpy2DAsync(rgb, rgb_pitch, base, f.rgb_pitch, (size_t)W * 3, (size_t)H, cudaMemcpyDeviceToHost, s));
    return frame_stats(luma, chroma, s, stats);
}

