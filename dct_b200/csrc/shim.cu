// shim.cu -- the C ABI of libdct_cuda, part 1: plans, device-plane and host-plane calls, statistics,
// the multi-GPU host helpers and the run-length entry points.  (shim_blocks.cu: the per-block calls of
// include/dct.h / include/quantization.h; shim_frames.cu: colour / edges / RGB frames; shim_peer.cu:
// NVLink peers.)
//
// Reference interface replaced here: the block loop a caller of the reference's per-block functions
// writes (tests/test_entropy.c:302-316, :370-384).  No CPU arithmetic happens on these paths; if CUDA
// is not usable they fail with an error code.
#include "plan.cuh"

using namespace dctb;
using namespace dctb::shim;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int dctb::shim::fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *dct_cuda_last_error(void) { return g_err; }

extern "C" int dct_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

namespace dctb {
namespace shim {


int read_tables(dct_cuda_plan *p)
{
    const DCTContext *d = p->dct;
    const QuantContext *q = p->quant;
    if (d->block_size != q->block_size || d->block_size < 1 || d->block_size > 32)
        return fail(DCT_CUDA_EINVAL, "block sizes %d / %d: both contexts must use the same size in 1..32",
                    d->block_size, q->block_size);
    p->n = d->block_size;
    p->adaptive = q->adaptive ? 1 : 0;
    // |c| <= n * 128 for centred 8-bit pixels (orthonormal transform), so |round(c / Q)| <= 127 whenever every
    // Q >= n * 128 / 127.5; the adaptive table only grows (Q * (2 - nv) >= Q, DC untouched)
    p->fits_i8 = true;
    for (int i = 0; i < p->n; ++i)
        for (int j = 0; j < p->n; ++j)
            if (!(q->quant_matrix[i][j] >= p->n * 128.0 / 127.5)) p->fits_i8 = false;
    if (p->n != 8) return DCT_CUDA_OK;   // K6 reads the contexts' tables as they are (upload_generic_tables)
    bool exotic = false;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            const int k = 8 * i + j;
            p->h_tab.D[k] = d->dct_matrix[i][j];
            p->h_tab.Q[k] = q->quant_matrix[i][j];
            p->h_tab.R[k] = q->dequant_matrix[i][j];
            const double Q = p->h_tab.Q[k], R = p->h_tab.R[k];
            // the fp32 path's proof assumes Q in [1, 1e6] (|c/Q| < 2^15, clamp of adjust_matrix inert)
            if (!(Q >= 1.0 && Q <= 1e6) || !std::isfinite(R) || R == 0.0) exotic = true;
        }
    p->exotic = exotic;

    const double u = std::ldexp(1.0, -24);
    for (int k = 0; k < 64; ++k) {
        const double Q = p->h_tab.Q[k], R = p->h_tab.R[k];
        // K1: y = c_scaled * r, r = 1/(Q * scale_k).  band = (beta_k + extra) / Q + floor:
        //   beta_k   fp32 butterfly error + rounding of r itself (derive_bands.py)
        //   extra    adaptive only: the fp32 1/(2-nv) and its product with r, <= 12u relative
        const double extra = p->adaptive ? 12.0 * u * kFwdCmax[k] : 0.0;
        const double band = ((double)kFwdBeta[k] * 1.02 + extra) / Q + std::ldexp(1.0, -22);
        p->r[k] = (float)(1.0 / (Q * kFwdScale[k]));
        double thr = 0.5 - band;
        if (exotic || !(thr > 0.0)) thr = -1.0;   // everything replays
        p->thr[k] = (float)thr;
        if ((double)p->thr[k] > thr) p->thr[k] = std::nextafterf(p->thr[k], -1.0f);   // round down
        // float pixel tiles: same multiplier, wider band (the inputs fl32(p - 128) are already rounded)
        double thr_f = 0.5 - (((double)kFwdBetaF32[k] * 1.02) / Q + std::ldexp(1.0, -22));
        if (exotic || !(thr_f > 0.0)) thr_f = -1.0;
        p->thr_f32[k] = (float)thr_f;
        if ((double)p->thr_f32[k] > thr_f) p->thr_f32[k] = std::nextafterf(p->thr_f32[k], -1.0f);
        // K2: v = q * rs.  non-adaptive: rs = R * prescale (sic: the reference multiplies by 1/Q);
        //     adaptive: rs = (1/R) * prescale, times (2-nv) per block in the kernel.
        const double mult = p->adaptive ? 1.0 / R : R;
        p->rs[k] = (float)(mult * kInvPrescale[k]);
        p->gain[k] = exotic ? 1e30f : kInvGain[k] * 1.02f;
        // bound on the raw quantised value: |q| * rg >= gain * |q * rs| with room for the fp32 product's rounding
        p->rg[k] = exotic ? 1e30f : std::nextafterf((float)((double)std::fabs(p->rs[k]) * (double)p->gain[k] * (1.0 + 1e-6)), INFINITY);
        p->h_tab.mp64[k] = mult * kInvPrescale[k];
        p->h_tab.mult64[k] = mult;
    }
    p->band_floor = 1.0e-6f;
    // One band for all coefficients costs half the instructions of 64 separate compares but replays
    // a block whenever ANY coefficient is within the WIDEST band of a boundary: about
    // 2 * 64 * max_band of all blocks.  Worth it while that stays below ~1 %.
    p->thr_min = p->thr[0];
    for (int k = 1; k < 64; ++k) p->thr_min = std::min(p->thr_min, p->thr[k]);
    p->uniform_band = (!exotic && (0.5 - (double)p->thr_min) * 128.0 < 0.02) ? 1 : 0;
    memcpy(p->h_tab.r32, p->r, sizeof p->r);
    memcpy(p->h_tab.thr32, p->thr, sizeof p->thr);
    memcpy(p->h_tab.thr32f, p->thr_f32, sizeof p->thr_f32);
    memcpy(p->h_tab.rs32, p->rs, sizeof p->rs);
    memcpy(p->h_tab.gain32, p->gain, sizeof p->gain);
    memcpy(p->h_tab.rg32, p->rg, sizeof p->rg);
    p->h_tab.band_floor = p->band_floor;
    p->h_tab.pad_[0] = p->h_tab.pad_[1] = p->h_tab.pad_[2] = 0.f;
    double per_q = 0.0;
    for (int k = 0; k < 64; ++k) per_q += (double)p->gain[k] * std::fabs(p->h_tab.mp64[k]);
    p->h_tab.bound_per_q = per_q * (p->adaptive ? 1.9 : 1.0);      // 2 - nv <= 1.9 (src/quantization.c:186-190)
    return DCT_CUDA_OK;
}

// K6: n*n tables exactly as the host contexts hold them + the zigzag position of every natural index
// (the scan of src/entropy.c:158-178 for block size n)
int upload_generic_tables(dct_cuda_plan *p)
{
    const int n = p->n, nn = n * n;
    std::vector<double> tab(3 * (size_t)nn);
    std::vector<int> pos(nn);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            tab[i * n + j] = p->dct->dct_matrix[i][j];
            tab[nn + i * n + j] = p->quant->quant_matrix[i][j];
            tab[2 * nn + i * n + j] = p->quant->dequant_matrix[i][j];
        }
    int idx = 0;
    for (int sum = 0; sum <= 2 * (n - 1); ++sum) {
        if (sum % 2 == 0) {
            for (int i = (sum < n) ? sum : n - 1; i >= 0 && (sum - i) < n; --i) pos[i * n + (sum - i)] = idx++;
        } else {
            for (int i = (sum < n) ? 0 : sum - n + 1; i < n && (sum - i) >= 0; ++i) pos[i * n + (sum - i)] = idx++;
        }
    }
    if (!p->d_gen) CU_TRY(cudaMalloc(&p->d_gen, 3 * (size_t)nn * sizeof(double)));
    if (!p->d_gen_pos) CU_TRY(cudaMalloc(&p->d_gen_pos, (size_t)nn * sizeof(int)));
    CU_TRY(cudaMemcpy(p->d_gen, tab.data(), 3 * (size_t)nn * sizeof(double), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(p->d_gen_pos, pos.data(), (size_t)nn * sizeof(int), cudaMemcpyHostToDevice));
    return DCT_CUDA_OK;
}

int queue_generic(dct_cuda_plan *p, Lane &ln, int forward, const uint8_t *px_in, uint8_t *px_out, size_t pitch, int W, int H,
                  const int16_t *coef_in, int16_t *coef_out, int layout, const double *var_in, double *var_out, cudaStream_t s)
{
    const int n = p->n;
    GenericParams gp{};
    gp.n = n;
    gp.blocks_per_cta = std::max(1, 256 / (n * n));
    gp.adaptive = p->adaptive;
    gp.layout = layout;
    gp.bw = (uint32_t)(W / n);
    gp.nblocks = gp.bw * (uint32_t)(H / n);
    gp.pitch = (long long)pitch;
    gp.D = p->d_gen;
    gp.Q = p->d_gen + n * n;
    gp.R = p->d_gen + 2 * n * n;
    gp.pos_of_natural = p->d_gen_pos;
    gp.ctr = ln.d_ctr;
    gp.px_in = px_in;
    gp.px_out = px_out;
    gp.coef_in = coef_in;
    gp.coef_out = coef_out;
    gp.var_in = var_in;
    gp.var_out = var_out;
    if (layout != DCT_CUDA_NATURAL && layout != DCT_CUDA_ZIGZAG) return fail(DCT_CUDA_EINVAL, "bad layout %d", layout);
    if (!forward && p->adaptive && !var_in) return fail(DCT_CUDA_EINVAL, "adaptive plan needs the per-block variance array");
    CU_TRY(launch_generic_plane(gp, forward, s));
    ln.blocks += gp.nblocks;
    return DCT_CUDA_OK;
}

// worklist entries K1 / K2 may need for a plane of bw x nby blocks: their bulk-tensor kernels cut every block row into
// tiles of 32 blocks (the last one of a row may be partial) and size a warp's segment for 32 entries per tile it visits
static size_t worklist_entries(uint32_t bw, uint32_t nby) { return (size_t)nby * ((bw + 31) / 32) * 32; }

int ensure_worklist(Lane &ln, size_t nblocks)
{
    // K1 / K2 append to one segment per warp of their grid, each sized for all the blocks that warp visits, so the
    // worklist holds `nblocks` entries (already rounded up to whole tiles by the caller) plus up to 64 of slack per segment
    const size_t want = nblocks + (size_t)kMaxWorklistSegments * 64;
    if (ln.wl_cap >= want) return DCT_CUDA_OK;
    if (ln.d_wl) {
        CU_TRY(cudaStreamSynchronize(ln.stream));
        CU_TRY(cudaDeviceSynchronize());
        CU_TRY(cudaFree(ln.d_wl));
        if (ln.d_side) CU_TRY(cudaFree(ln.d_side));
        ln.d_wl = nullptr, ln.d_side = nullptr;
        ln.wl_cap = 0, ln.side_cap = 0;
    }
    if (!ln.d_seg_count) {
        CU_TRY(cudaMalloc(&ln.d_seg_count, kMaxWorklistSegments * sizeof(uint32_t)));
        CU_TRY(cudaMemset(ln.d_seg_count, 0, kMaxWorklistSegments * sizeof(uint32_t)));
    }
    CU_TRY(cudaMalloc(&ln.d_wl, want * sizeof(uint32_t)));
    ln.wl_cap = (uint32_t)want;
    // room for the pixels of one block in eight (uniform noise at q50 flags 2.7 %), divided among the segments
    // like the worklist; entries beyond a segment's share fall back to the plane
    const size_t side = nblocks / 8 + (size_t)kMaxWorklistSegments * 32;
    CU_TRY(cudaMalloc(&ln.d_side, side * 64));
    ln.side_cap = (uint32_t)side;
    return DCT_CUDA_OK;
}

// `dev`: the pitch is used by the kernels directly (8-byte rows); host planes are re-packed by the copy
int check_plane(const void *a, const void *b, size_t pitch, int W, int H, bool dev, int n)
{
    if (!a || !b) return fail(DCT_CUDA_EINVAL, "NULL data pointer");
    if (W < 0 || H < 0 || (W % n) || (H % n))
        return fail(DCT_CUDA_EINVAL, "width and height must be non-negative multiples of %d (got %dx%d)", n, W, H);
    if (pitch < (size_t)W || (dev && n == 8 && (pitch % 8)))
        return fail(DCT_CUDA_EINVAL, "pitch %zu must be >= width%s", pitch, dev ? " and a multiple of 8" : "");
    if ((uint64_t)(W / n) * (uint64_t)(H / n) > 0xFFFFFFF0ull) return fail(DCT_CUDA_EINVAL, "too many blocks");
    return DCT_CUDA_OK;
}

// host planes whose sides are not multiples of the block size (the *_edge calls)
int check_ragged(const void *a, const void *b, size_t pitch, int W, int H, int n)
{
    if (W <= 0 || H <= 0) return fail(DCT_CUDA_EINVAL, "width and height must be positive (got %dx%d)", W, H);
    if (!a || !b) return fail(DCT_CUDA_EINVAL, "NULL data pointer");
    if (pitch < (size_t)W) return fail(DCT_CUDA_EINVAL, "pitch %zu must be >= width", pitch);
    if ((uint64_t)((W + n - 1) / n) * (uint64_t)((H + n - 1) / n) > 0xFFFFFFF0ull) return fail(DCT_CUDA_EINVAL, "too many blocks");
    return DCT_CUDA_OK;
}

// K1's parameters for one device-resident 8x8 plane on lane `ln`
FwdParams fwd_params(dct_cuda_plan *p, Lane &ln, const uint8_t *d_px, size_t pitch, uint32_t bw, uint32_t nblocks, int16_t *d_coef,
                     double *d_var, bool use_side, int elem)
{
    FwdParams fp{};
    fp.px = d_px;
    fp.pitch = (long long)pitch;
    fp.bw = bw;
    fp.nblocks = nblocks;
    fp.coef = d_coef;
    fp.var_out = p->adaptive ? d_var : nullptr;
    fp.worklist = ln.d_wl;
    fp.wl_cap = ln.wl_cap;
    fp.ctr = ln.d_ctr;
    memcpy(fp.r, p->r, sizeof fp.r);
    memcpy(fp.thr, elem == 4 ? p->thr_f32 : p->thr, sizeof fp.thr);
    fp.thr_min = p->thr_min;
    fp.uniform_band = p->uniform_band;
    fp.side = use_side ? ln.d_side : nullptr;
    fp.side_cap = use_side ? ln.side_cap : 0;
    fp.seg_count = ln.d_seg_count;
    fp.no_tma = p->no_tma ? 1 : 0;
    fp.tab = p->skip_replay ? nullptr : p->d_tab;     // lets the bulk-tensor kernel replay its own flagged blocks
    return fp;
}

// K2's parameters for one device-resident 8x8 plane on lane `ln`
InvParams inv_params(dct_cuda_plan *p, Lane &ln, const int16_t *d_coef, const double *d_var, uint8_t *d_px, size_t pitch, uint32_t bw,
                     uint32_t nblocks)
{
    InvParams ip{};
    ip.coef = d_coef;
    ip.var_in = p->adaptive ? d_var : nullptr;
    ip.px = d_px;
    ip.pitch = (long long)pitch;
    ip.bw = bw;
    ip.nblocks = nblocks;
    ip.worklist = ln.d_wl;
    ip.wl_cap = ln.wl_cap;
    ip.ctr = ln.d_ctr;
    memcpy(ip.rs, p->rs, sizeof ip.rs);
    memcpy(ip.gain, p->gain, sizeof ip.gain);
    memcpy(ip.rg, p->rg, sizeof ip.rg);
    ip.band_floor = p->band_floor;
    ip.seg_count = ln.d_seg_count;
    ip.no_tma = p->no_tma ? 1 : 0;
    ip.tab = p->skip_replay ? nullptr : p->d_tab;     // lets the bulk-tensor kernel replay its own flagged blocks
    // multipliers of the folded first stage, in the kernel's pair order
    static const int colA[4] = {0, 2, 5, 1}, colB[4] = {4, 6, 3, 7};
    for (int c = 0; c < 4; ++c)
        for (int j = 0; j < 4; ++j) {
            const int ra = colA[j], rb = colB[j];                  // first-stage row pair (a_j, b_j)
            ip.ma[c][j] = make_float2(p->rs[8 * ra + colA[c]], p->rs[8 * ra + colB[c]]);
            ip.mb[c][j].pos = make_float2(p->rs[8 * rb + colA[c]], p->rs[8 * rb + colB[c]]);
            ip.mb[c][j].neg = make_float2(-ip.mb[c][j].pos.x, -ip.mb[c][j].pos.y);
        }
    return ip;
}

// queue K1 (+K3) for one device-resident plane on lane `ln`, stream `s`
// elem: bytes per pixel of the source plane -- 1 (uint8) or 4 (float tiles); pitch is in bytes
int queue_fwd(dct_cuda_plan *p, Lane &ln, const uint8_t *d_px, size_t pitch, int W, int H, int16_t *d_coef,
              int layout, double *d_var, cudaStream_t s, int elem)
{
    if (p->n != 8) {
        if (elem != 1) return fail(DCT_CUDA_EINVAL, "float pixel tiles are 8x8 only");
        return queue_generic(p, ln, 1, d_px, nullptr, pitch, W, H, nullptr, d_coef, layout, nullptr, p->adaptive ? d_var : nullptr, s);
    }
    const uint32_t bw = W / 8, nblocks = bw * (uint32_t)(H / 8);
    if (nblocks == 0) return DCT_CUDA_OK;
    if (((uintptr_t)d_px % (elem == 4 ? 16 : 8)) || ((uintptr_t)d_coef % 16))
        return fail(DCT_CUDA_EINVAL, "pixels must be %d-byte and coefficients 16-byte aligned", elem == 4 ? 16 : 8);
    if (elem == 4 && p->adaptive)
        return fail(DCT_CUDA_EINVAL, "float pixel tiles are supported for non-adaptive plans only");
    if (layout != DCT_CUDA_NATURAL && layout != DCT_CUDA_ZIGZAG) return fail(DCT_CUDA_EINVAL, "bad layout %d", layout);
    int rc = ensure_worklist(ln, worklist_entries(bw, (uint32_t)(H / 8)));
    if (rc) return rc;
    // wl_count is zero here: plan creation zeroes it and K3's last CTA re-zeroes it after every replay
    if (p->skip_replay) CU_TRY(cudaMemsetAsync(&ln.d_ctr->wl_count, 0, sizeof(unsigned), s));

    ReplayParams rp{};
    rp.tab = p->d_tab;
    rp.h_tab = &p->h_tab;
    rp.ctr = ln.d_ctr;
    rp.nblocks = nblocks;
    rp.wl_cap = ln.wl_cap;
    rp.bw = bw;
    rp.adaptive = p->adaptive;
    rp.layout = layout;
    rp.pitch = (long long)pitch;
    rp.px_in = d_px;
    rp.px_is_f32 = elem == 4;
    static const bool no_side = getenv("DCT_CUDA_NO_SIDE") != nullptr;   // debugging aid: K3 reads the plane instead
    const bool use_side = elem == 1 && !no_side;          // the one-lane-per-block replay kernel reads it
    rp.side = use_side ? ln.d_side : nullptr;
    rp.side_cap = use_side ? ln.side_cap : 0;
    rp.coef_out = d_coef;
    rp.var_out = p->adaptive ? d_var : nullptr;
    bool folded = false;                                  // K1 replayed its flagged blocks itself: no K3 launch
    if (!p->exotic) {
        const FwdParams fp = fwd_params(p, ln, d_px, pitch, bw, nblocks, d_coef, d_var, use_side, elem);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (p->profile) {
            CU_TRY(cudaEventCreate(&e0));
            CU_TRY(cudaEventCreate(&e1));
            CU_TRY(cudaEventRecord(e0, s));
        }
        unsigned k1_launches = 1;
        if (elem == 4) {
            CU_TRY(launch_fwd_quant_f32(fp, layout, s));
        } else {
            CU_TRY(launch_fwd_quant_u8(fp, layout, p->adaptive, s, &k1_launches, &rp.seg, &folded));
            rp.seg_count = ln.d_seg_count;
        }
        p->launches += k1_launches;
        if (p->profile) {
            CU_TRY(cudaEventRecord(e1, s));
            ln.ev_fwd.emplace_back(e0, e1);
        }
        rp.worklist = ln.d_wl;
    }
    if ((!p->skip_replay || p->exotic) && !folded) {
        CU_TRY(launch_replay_fwd(rp, s));
        ++p->launches;
    }
    ln.blocks += nblocks;
    return DCT_CUDA_OK;
}

int queue_inv(dct_cuda_plan *p, Lane &ln, const int16_t *d_coef, int W, int H, int layout, const double *d_var,
              uint8_t *d_px, size_t pitch, cudaStream_t s)
{
    if (p->n != 8)
        return queue_generic(p, ln, 0, nullptr, d_px, pitch, W, H, d_coef, nullptr, layout, p->adaptive ? d_var : nullptr, nullptr, s);
    const uint32_t bw = W / 8, nblocks = bw * (uint32_t)(H / 8);
    if (nblocks == 0) return DCT_CUDA_OK;
    if (((uintptr_t)d_px % 8) || ((uintptr_t)d_coef % 16))
        return fail(DCT_CUDA_EINVAL, "pixels must be 8-byte and coefficients 16-byte aligned");
    if (layout != DCT_CUDA_NATURAL && layout != DCT_CUDA_ZIGZAG) return fail(DCT_CUDA_EINVAL, "bad layout %d", layout);
    if (p->adaptive && !d_var) return fail(DCT_CUDA_EINVAL, "adaptive plan needs the per-block variance array");
    int rc = ensure_worklist(ln, worklist_entries(bw, (uint32_t)(H / 8)));
    if (rc) return rc;
    // wl_count is zero here: plan creation zeroes it and K3's last CTA re-zeroes it after every replay
    if (p->skip_replay) CU_TRY(cudaMemsetAsync(&ln.d_ctr->wl_count, 0, sizeof(unsigned), s));

    ReplayParams rp{};
    rp.tab = p->d_tab;
    rp.h_tab = &p->h_tab;
    rp.ctr = ln.d_ctr;
    rp.nblocks = nblocks;
    rp.wl_cap = ln.wl_cap;
    rp.bw = bw;
    rp.adaptive = p->adaptive;
    rp.layout = layout;
    rp.pitch = (long long)pitch;
    rp.coef_in = d_coef;
    rp.var_in = p->adaptive ? d_var : nullptr;
    rp.px_out = d_px;
    bool folded = false;                                  // K2 replayed its flagged blocks itself: no K3 launch
    if (!p->exotic) {
        const InvParams ip = inv_params(p, ln, d_coef, d_var, d_px, pitch, bw, nblocks);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (p->profile) {
            CU_TRY(cudaEventCreate(&e0));
            CU_TRY(cudaEventCreate(&e1));
            CU_TRY(cudaEventRecord(e0, s));
        }
        // DCT_CUDA_INV_FP64=1: adaptive plans through the fp64 butterfly (content-independent 0.40 of peak; the fp32
        // kernel with its replay is faster even on uniform noise since the bands were re-derived)
        if (p->adaptive && p->fp64_inverse) {
            CU_TRY(launch_dequant_idct_u8_f64(ip, p->d_tab, layout, s));
        } else {
            CU_TRY(launch_dequant_idct_u8(ip, layout, p->adaptive, s, &rp.seg, &folded));
            if (rp.seg.n_segs) rp.seg_count = ln.d_seg_count;     // segmented worklist (bulk-tensor kernel)
        }
        ++p->launches;
        if (p->profile) {
            CU_TRY(cudaEventRecord(e1, s));
            ln.ev_inv.emplace_back(e0, e1);
        }
        rp.worklist = ln.d_wl;
    }
    if ((!p->skip_replay || p->exotic) && !folded) {
        CU_TRY(launch_replay_inv(rp, s));
        ++p->launches;
    }
    ln.blocks += nblocks;
    return DCT_CUDA_OK;
}

// device strip buffers of one lane, sized in pixels: `pixels * elem` bytes of pixels, 2 + 1 bytes of record per
// pixel (int16 and int8 form), one variance per block
int ensure_strip_buffers(dct_cuda_plan *p, Lane &ln, size_t pixels, int elem)
{
    const size_t need = (pixels * (size_t)elem + 255) & ~(size_t)255;   // whole 256-byte units: the int8 records behind the int16 ones stay aligned
    if (ln.cap_blocks >= need) return DCT_CUDA_OK;
    CU_TRY(cudaStreamSynchronize(ln.stream));
    if (ln.d_px) cudaFree(ln.d_px);
    if (ln.d_coef) cudaFree(ln.d_coef);
    if (ln.d_var) cudaFree(ln.d_var);
    ln.d_px = nullptr, ln.d_coef = nullptr, ln.d_var = nullptr, ln.cap_blocks = 0;
    CU_TRY(cudaMalloc(&ln.d_px, need));
    CU_TRY(cudaMalloc(&ln.d_coef, need * 3));      // int16 records, then their int8 form (narrow.cu)
    if (p->adaptive) CU_TRY(cudaMalloc(&ln.d_var, (need / ((size_t)p->n * p->n) + 1) * sizeof(double)));
    ln.cap_blocks = need;
    return DCT_CUDA_OK;
}

int collect_stats(dct_cuda_plan *p, dct_cuda_stats *out, cudaStream_t user_stream)
{
    // lane 0 may have been driven on a caller's stream
    CU_TRY(cudaStreamSynchronize(user_stream));
    dct_cuda_stats st{};
    for (int l = 0; l < kLanes; ++l) {
        Lane &ln = p->lane[l];
        CU_TRY(cudaStreamSynchronize(ln.stream));
        CU_TRY(cudaMemcpyAsync(&p->h_ctr[l], ln.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ln.stream));
        // reset the statistics only (wl_count / done_ctas belong to K3, which re-zeroes them itself), on the lane's own
        // stream: a memset on the legacy default stream is not ordered against work queued on non-blocking streams
        CU_TRY(cudaMemsetAsync(&ln.d_ctr->replayed, 0, sizeof(Counters) - offsetof(Counters, replayed), ln.stream));
        CU_TRY(cudaStreamSynchronize(ln.stream));
        st.blocks += ln.blocks;
        st.replayed_blocks += p->h_ctr[l].replayed;
        st.near_ties += p->h_ctr[l].near_ties;
        st.saturated += p->h_ctr[l].saturated;
        ln.blocks = 0;
    }
    if (out) *out = st;
    return DCT_CUDA_OK;
}

}  // namespace shim
}  // namespace dctb

extern "C" dct_cuda_plan *dct_cuda_plan_create(const DCTContext *dct, const QuantContext *quant, int device)
{
    if (!dct || !quant) {
        fail(DCT_CUDA_EINVAL, "NULL context");
        return nullptr;
    }
    const int ndev = dct_cuda_device_count();
    if (ndev <= 0) {
        fail(DCT_CUDA_ENODEV, "no CUDA device available (libdct_cuda has no CPU fallback)");
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        fail(DCT_CUDA_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
        return nullptr;
    }
    dct_cuda_plan *p = new (std::nothrow) dct_cuda_plan();
    if (!p) {
        fail(DCT_CUDA_ENOMEM, "out of host memory");
        return nullptr;
    }
    p->device = device;
    p->fp64_inverse = getenv("DCT_CUDA_INV_FP64") != nullptr;
    p->dct = dct;
    p->quant = quant;
    DeviceGuard g(device);
    auto init = [&]() -> int {
        int rc = read_tables(p);
        if (rc) return rc;
        if (p->n != 8 && (rc = upload_generic_tables(p))) return rc;
        CU_TRY(cudaMalloc(&p->d_tab, sizeof(ExactTables)));
        CU_TRY(cudaMemcpy(p->d_tab, &p->h_tab, sizeof(ExactTables), cudaMemcpyHostToDevice));
        CU_TRY(cudaMallocHost(&p->h_ctr, kLanes * sizeof(Counters)));
        for (int l = 0; l < kLanes; ++l) {
            CU_TRY(cudaStreamCreateWithFlags(&p->lane[l].stream, cudaStreamNonBlocking));
            CU_TRY(cudaMalloc(&p->lane[l].d_ctr, sizeof(Counters)));
            CU_TRY(cudaMemset(p->lane[l].d_ctr, 0, sizeof(Counters)));
        }
        return DCT_CUDA_OK;
    };
    if (init() != DCT_CUDA_OK) {
        dct_cuda_plan_destroy(p);
        return nullptr;
    }
    return p;
}

extern "C" int dct_cuda_plan_refresh(dct_cuda_plan *p)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    CU_TRY(cudaDeviceSynchronize());
    int rc = read_tables(p);
    if (rc) return rc;
    if (p->n != 8) return upload_generic_tables(p);
    CU_TRY(cudaMemcpy(p->d_tab, &p->h_tab, sizeof(ExactTables), cudaMemcpyHostToDevice));
    return DCT_CUDA_OK;
}

extern "C" void dct_cuda_plan_destroy(dct_cuda_plan *p)
{
    if (!p) return;
    DeviceGuard g(p->device);
    cudaDeviceSynchronize();
    for (int l = 0; l < kLanes; ++l) {
        Lane &ln = p->lane[l];
        if (ln.d_ctr) cudaFree(ln.d_ctr);
        if (ln.d_wl) cudaFree(ln.d_wl);
        if (ln.d_side) cudaFree(ln.d_side);
        if (ln.d_seg_count) cudaFree(ln.d_seg_count);
        if (ln.d_px) cudaFree(ln.d_px);
        if (ln.d_coef) cudaFree(ln.d_coef);
        if (ln.d_var) cudaFree(ln.d_var);
        if (ln.stream) cudaStreamDestroy(ln.stream);
    }
    if (p->d_gen) cudaFree(p->d_gen);
    if (p->d_gen_pos) cudaFree(p->d_gen_pos);
    if (p->d_rle_sums) cudaFree(p->d_rle_sums);
    if (p->d_rle_total) cudaFree(p->d_rle_total);
    if (p->d_tab) cudaFree(p->d_tab);
    if (p->d_frame) cudaFree(p->d_frame);
    if (p->ev_peer) cudaEventDestroy(p->ev_peer);
    if (p->h_ctr) cudaFreeHost(p->h_ctr);
    delete p;
}

extern "C" int dct_cuda_plan_device(const dct_cuda_plan *p) { return p ? p->device : -1; }

extern "C" uint64_t dct_cuda_plan_kernel_launches(const dct_cuda_plan *p) { return p ? p->launches : 0; }

// ------------------------------------------------------------------------------------------
// device-resident planes
// ------------------------------------------------------------------------------------------
extern "C" int dct_cuda_fwd_quant_u8_dev(dct_cuda_plan *p, const uint8_t *d_px, size_t pitch, int W, int H,
                                         int16_t *d_coef, int layout, double *d_var, void *stream)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    int rc = check_plane(d_px, d_coef, pitch, W, H, true, p->n);
    if (rc) return rc;
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    return queue_fwd(p, p->lane[0], d_px, pitch, W, H, d_coef, layout, d_var, (cudaStream_t)stream);
}

extern "C" int dct_cuda_fwd_quant_f32_dev(dct_cuda_plan *p, const float *d_px, size_t pitch_bytes, int W, int H,
                                          int16_t *d_coef, int layout, void *stream)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    if (p->n != 8) return fail(DCT_CUDA_EINVAL, "float pixel tiles are 8x8 only");
    int rc = check_plane(d_px, d_coef, pitch_bytes / 4, W, H, false);
    if (rc) return rc;
    if (pitch_bytes % 16) return fail(DCT_CUDA_EINVAL, "float planes need a pitch that is a multiple of 16 bytes");
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    return queue_fwd(p, p->lane[0], (const uint8_t *)d_px, pitch_bytes, W, H, d_coef, layout, nullptr, (cudaStream_t)stream, 4);
}

extern "C" int dct_cuda_dequant_idct_u8_dev(dct_cuda_plan *p, const int16_t *d_coef, int W, int H, int layout,
                                            const double *d_var, uint8_t *d_px, size_t pitch, void *stream)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    int rc = check_plane(d_px, d_coef, pitch, W, H, true, p->n);
    if (rc) return rc;
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    return queue_inv(p, p->lane[0], d_coef, W, H, layout, d_var, d_px, pitch, (cudaStream_t)stream);
}

// The 2 or 3 planes of one frame (Y, Cb, Cr) in ONE launch each way.  A persistent grid's ramp and tail cost more than
// the tiles of a 4K plane, so planes queued one by one pay them three times: 8K 4:2:0 took 107 us per frame that way.
// Qualifying planes: 8x8 non-adaptive plans on one device, each plane small enough to replay its flagged blocks in the
// kernel's tail (kFoldMaxBlocks) and laid out for the bulk-tensor kernels.  Returns false (nothing queued) otherwise:
// the caller then queues the planes one by one, which also reports any argument error.
static bool queue_planes_together(const dct_cuda_plane *pl, int n, int layout, cudaStream_t s, bool forward, int *rc_out)
{
    if (n < 2 || n > 3 || (layout != DCT_CUDA_NATURAL && layout != DCT_CUDA_ZIGZAG)) return false;
    dct_cuda_plan *plans[3];
    for (int i = 0; i < n; ++i) {
        dct_cuda_plan *p = pl[i].plan;
        if (!p || p->n != 8 || p->adaptive || p->exotic || p->skip_replay || p->profile || p->no_tma || p->device != pl[0].plan->device)
            return false;
        const void *px = forward ? pl[i].pixels_in : (const void *)pl[i].pixels_out;
        if (!px || !pl[i].coef || pl[i].width <= 0 || pl[i].height <= 0 || (pl[i].width % 8) || (pl[i].height % 8)) return false;
        if (pl[i].pitch < (size_t)pl[i].width || (pl[i].pitch % 8) || ((uintptr_t)px % 8) || ((uintptr_t)pl[i].coef % 16)) return false;
        if ((uint64_t)(pl[i].width / 8) * (uint64_t)(pl[i].height / 8) > kFoldMaxBlocks) return false;
        plans[i] = p;
    }
    // every distinct plan once, in address order (two threads queueing the same plans cannot deadlock)
    dct_cuda_plan *uniq[3];
    int nu = 0;
    for (int i = 0; i < n; ++i) {
        bool seen = false;
        for (int j = 0; j < nu; ++j) seen = seen || uniq[j] == plans[i];
        if (!seen) uniq[nu++] = plans[i];
    }
    std::sort(uniq, uniq + nu);
    std::unique_lock<std::mutex> locks[3];
    for (int j = 0; j < nu; ++j) locks[j] = std::unique_lock<std::mutex>(uniq[j]->mu);
    DeviceGuard g(plans[0]->device);

    FwdParams fp[3];
    InvParams ip[3];
    for (int j = 0; j < nu; ++j) {          // one worklist per plan: its planes take consecutive ranges of every warp's segment
        size_t want = 0;
        for (int i = 0; i < n; ++i)
            if (plans[i] == uniq[j]) want += worklist_entries(pl[i].width / 8, pl[i].height / 8);
        if (ensure_worklist(uniq[j]->lane[0], want) != DCT_CUDA_OK) return false;
    }
    static const bool no_side = getenv("DCT_CUDA_NO_SIDE") != nullptr;
    for (int i = 0; i < n; ++i) {
        Lane &ln = plans[i]->lane[0];
        const uint32_t bw = pl[i].width / 8, nblocks = bw * (uint32_t)(pl[i].height / 8);
        if (forward)
            fp[i] = fwd_params(plans[i], ln, (const uint8_t *)pl[i].pixels_in, pl[i].pitch, bw, nblocks, (int16_t *)pl[i].coef, nullptr,
                               !no_side, 1);
        else
            ip[i] = inv_params(plans[i], ln, (const int16_t *)pl[i].coef, nullptr, (uint8_t *)pl[i].pixels_out, pl[i].pitch, bw, nblocks);
    }
    const cudaError_t e = forward ? launch_fwd_quant_u8_multi(fp, n, layout, s) : launch_dequant_idct_u8_multi(ip, n, layout, s);
    if (e == cudaErrorNotSupported) return false;
    if (e != cudaSuccess) {
        *rc_out = fail(DCT_CUDA_ECUDA, "%s planes in one launch: %s", forward ? "forward" : "inverse", cudaGetErrorString(e));
        return true;
    }
    ++plans[0]->launches;
    for (int i = 0; i < n; ++i) plans[i]->lane[0].blocks += (uint64_t)(pl[i].width / 8) * (uint64_t)(pl[i].height / 8);
    *rc_out = DCT_CUDA_OK;
    return true;
}

extern "C" int dct_cuda_fwd_quant_planes_dev(const dct_cuda_plane *pl, int n, int layout, void *stream)
{
    if (!pl || n < 0) return fail(DCT_CUDA_EINVAL, "bad plane list");
    int together = DCT_CUDA_OK;
    if (queue_planes_together(pl, n, layout, (cudaStream_t)stream, true, &together)) return together;
    for (int i = 0; i < n; ++i) {
        int rc = dct_cuda_fwd_quant_u8_dev(pl[i].plan, (const uint8_t *)pl[i].pixels_in, pl[i].pitch, pl[i].width,
                                           pl[i].height, (int16_t *)pl[i].coef, layout, pl[i].variance, stream);
        if (rc) return rc;
    }
    return DCT_CUDA_OK;
}

extern "C" int dct_cuda_dequant_idct_planes_dev(const dct_cuda_plane *pl, int n, int layout, void *stream)
{
    if (!pl || n < 0) return fail(DCT_CUDA_EINVAL, "bad plane list");
    int together = DCT_CUDA_OK;
    if (queue_planes_together(pl, n, layout, (cudaStream_t)stream, false, &together)) return together;
    for (int i = 0; i < n; ++i) {
        int rc = dct_cuda_dequant_idct_u8_dev(pl[i].plan, (const int16_t *)pl[i].coef, pl[i].width, pl[i].height,
                                              layout, pl[i].variance, (uint8_t *)pl[i].pixels_out, pl[i].pitch,
                                              stream);
        if (rc) return rc;
    }
    return DCT_CUDA_OK;
}

extern "C" int dct_cuda_stats_fetch(dct_cuda_plan *p, dct_cuda_stats *stats, void *stream)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    return collect_stats(p, stats, (cudaStream_t)stream);
}

extern "C" int dct_cuda_plan_debug_skip_replay(dct_cuda_plan *p, int skip)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    CU_TRY(cudaDeviceSynchronize());
    for (int l = 0; l < kLanes; ++l)   // skipped replays leave their worklist count behind
        CU_TRY(cudaMemset(&p->lane[l].d_ctr->wl_count, 0, 2 * sizeof(unsigned)));
    p->skip_replay = skip != 0;
    return DCT_CUDA_OK;
}

extern "C" int dct_cuda_plan_profile(dct_cuda_plan *p, int enable)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    p->profile = enable != 0;
    return DCT_CUDA_OK;
}

extern "C" int dct_cuda_profile_fetch(dct_cuda_plan *p, double *fwd_ms, int *fwd_launches, double *inv_ms,
                                      int *inv_launches)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    CU_TRY(cudaDeviceSynchronize());
    double ms[2] = {0.0, 0.0};
    int n[2] = {0, 0};
    for (int l = 0; l < kLanes; ++l) {
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> *lists[2] = {&p->lane[l].ev_fwd, &p->lane[l].ev_inv};
        for (int d = 0; d < 2; ++d) {
            for (auto &pr : *lists[d]) {
                float t = 0.f;
                CU_TRY(cudaEventElapsedTime(&t, pr.first, pr.second));
                ms[d] += t;
                ++n[d];
                cudaEventDestroy(pr.first);
                cudaEventDestroy(pr.second);
            }
            lists[d]->clear();
        }
    }
    if (fwd_ms) *fwd_ms = ms[0];
    if (fwd_launches) *fwd_launches = n[0];
    if (inv_ms) *inv_ms = ms[1];
    if (inv_launches) *inv_launches = n[1];
    return DCT_CUDA_OK;
}

// ------------------------------------------------------------------------------------------
// host planes: strips of block rows through a kLanes-deep H2D / kernel / D2H pipeline
// ------------------------------------------------------------------------------------------
static int strip_rows(int W, int H, int n = 8)
{
    const size_t bw = (size_t)W / n;
    if (bw == 0 || H == 0) return 0;
    static const size_t strip_px = getenv("DCT_CUDA_STRIP_MPX") ? (size_t)atol(getenv("DCT_CUDA_STRIP_MPX")) << 20 : kStripPixels;   // tuning aid
    size_t rows = std::max<size_t>(1, strip_px / (bw * n * n));             // block rows per strip
    const size_t total = (size_t)H / n;
    // at least kLanes strips when the plane is big enough to be worth overlapping
    if (total >= (size_t)kLanes * 4) rows = std::min(rows, (total + kLanes - 1) / kLanes);
    return (int)std::min(rows, total);
}

// `ragged`: W and H are any positive sizes; the strips are completed to whole blocks on the device by
// replicating the last column / row (planar.cu), so the records cover ceil(W/n) x ceil(H/n) blocks
// waits for everything queued on the plan's lanes (after a failure half way through a plane: the strips already
// queued still read / write the caller's buffers)
static void drain_lanes(dct_cuda_plan *p)
{
    if (!p) return;
    DeviceGuard g(p->device);
    for (int l = 0; l < kLanes; ++l)
        if (p->lane[l].stream) cudaStreamSynchronize(p->lane[l].stream);
}

static int fwd_host_queue(dct_cuda_plan *p, const uint8_t *px, size_t pitch, int W, int H, void *coef, int layout,
                          double *var, int elem, bool ragged, bool rec8);

static int fwd_host_async(dct_cuda_plan *p, const uint8_t *px, size_t pitch, int W, int H, void *coef, int layout,
                          double *var, int elem, bool ragged = false, bool rec8 = false)
{
    const int rc = fwd_host_queue(p, px, pitch, W, H, coef, layout, var, elem, ragged, rec8);
    if (rc) drain_lanes(p);
    return rc;
}

static int fwd_host_queue(dct_cuda_plan *p, const uint8_t *px, size_t pitch, int W, int H, void *coef, int layout,
                          double *var, int elem, bool ragged, bool rec8)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    if (rec8 && !p->fits_i8)
        return fail(DCT_CUDA_EINVAL, "int8 records need every table entry >= %.3f (dct_cuda_plan_records_fit_i8)",
                    p->n * 128.0 / 127.5);
    const int n = p->n, nn = n * n;
    if (rec8 && (nn % 16))
        return fail(DCT_CUDA_EINVAL, "int8 records need a block size whose square is a multiple of 16 (got %d)", n);
    int rc = ragged ? check_ragged(px, coef, pitch / elem, W, H, n) : check_plane(px, coef, pitch / elem, W, H, false, n);
    if (rc) return rc;
    if (elem == 4 && p->adaptive) return fail(DCT_CUDA_EINVAL, "float pixel tiles are supported for non-adaptive plans only");
    if (layout != DCT_CUDA_NATURAL && layout != DCT_CUDA_ZIGZAG) return fail(DCT_CUDA_EINVAL, "bad layout %d", layout);
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    const int Wp = (W + n - 1) / n * n, Hp = (H + n - 1) / n * n;   // == W, H unless ragged
    const int bw = Wp / n, total_rows = Hp / n, rows = strip_rows(Wp, Hp, n);
    if (rows > 0) {
        // rows of the device strip start on 16-byte boundaries (the bulk-tensor kernels need that)
        const size_t dev_pitch = ((size_t)Wp * elem + 15) & ~(size_t)15;
        for (int l = 0; l < kLanes; ++l)
            if ((rc = ensure_strip_buffers(p, p->lane[l], (size_t)rows * n * dev_pitch, 1))) return rc;
        int idx = 0;
        for (int r0 = 0; r0 < total_rows; r0 += rows, ++idx) {
            Lane &ln = p->lane[idx % kLanes];
            const int nr = std::min(rows, total_rows - r0);
            const int have = std::min(nr * n, H - r0 * n);               // image rows in this strip
            const size_t nb = (size_t)nr * bw, b0 = (size_t)r0 * bw;
            if (pitch == dev_pitch && (size_t)W * elem == pitch)         // dense rows on both sides: one linear copy
                CU_TRY(cudaMemcpyAsync(ln.d_px, px + (size_t)r0 * n * pitch, pitch * (size_t)have, cudaMemcpyHostToDevice, ln.stream));
            else
                CU_TRY(cudaMemcpy2DAsync(ln.d_px, dev_pitch, px + (size_t)r0 * n * pitch, pitch, (size_t)W * elem, (size_t)have,
                                         cudaMemcpyHostToDevice, ln.stream));
            if (Wp != W || have != nr * n)
                CU_TRY(launch_pad_edges(ln.d_px, (long long)dev_pitch, W, have, Wp, nr * n, elem, ln.stream));
            if ((rc = queue_fwd(p, ln, ln.d_px, dev_pitch, Wp, nr * n, ln.d_coef, layout, ln.d_var, ln.stream, elem))) return rc;
            if (rec8) {
                int8_t *d8 = reinterpret_cast<int8_t *>(ln.d_coef) + ln.cap_blocks * 2;
                CU_TRY(launch_narrow_records(ln.d_coef, d8, nb * nn, ln.d_ctr, ln.stream));
                ++p->launches;
                CU_TRY(cudaMemcpyAsync((int8_t *)coef + b0 * nn, d8, nb * nn, cudaMemcpyDeviceToHost, ln.stream));
            } else {
                CU_TRY(cudaMemcpyAsync((int16_t *)coef + b0 * nn, ln.d_coef, nb * nn * 2, cudaMemcpyDeviceToHost, ln.stream));
            }
            if (p->adaptive && var)
                CU_TRY(cudaMemcpyAsync(var + b0, ln.d_var, nb * sizeof(double), cudaMemcpyDeviceToHost, ln.stream));
        }
    }
    return DCT_CUDA_OK;
}

extern "C" int dct_cuda_fwd_quant_u8_async(dct_cuda_plan *p, const uint8_t *px, size_t pitch, int W, int H,
                                           int16_t *coef, int layout, double *var)
{
    return fwd_host_async(p, px, pitch, W, H, coef, layout, var, 1);
}

extern "C" int dct_cuda_plan_wait(dct_cuda_plan *p, dct_cuda_stats *stats);

extern "C" int dct_cuda_fwd_quant_f32(dct_cuda_plan *p, const float *px, size_t pitch_bytes, int W, int H, int16_t *coef,
                                      int layout, dct_cuda_stats *stats)
{
    int rc = fwd_host_async(p, (const uint8_t *)px, pitch_bytes, W, H, coef, layout, nullptr, 4);
    if (rc) return rc;
    return dct_cuda_plan_wait(p, stats);
}

extern "C" int dct_cuda_plan_wait(dct_cuda_plan *p, dct_cuda_stats *stats)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    for (int l = 0; l < kLanes; ++l) CU_TRY(cudaStreamSynchronize(p->lane[l].stream));
    if (stats) return collect_stats(p, stats, nullptr);
    return DCT_CUDA_OK;
}

extern "C" int dct_cuda_fwd_quant_u8(dct_cuda_plan *p, const uint8_t *px, size_t pitch, int W, int H,
                                     int16_t *coef, int layout, double *var, dct_cuda_stats *stats)
{
    int rc = dct_cuda_fwd_quant_u8_async(p, px, pitch, W, H, coef, layout, var);
    if (rc) return rc;
    return dct_cuda_plan_wait(p, stats);
}

static int inv_host_queue(dct_cuda_plan *p, const void *coef, int W, int H, int layout, const double *var, uint8_t *px,
                          size_t pitch, bool ragged, bool rec8);

static int inv_host_async(dct_cuda_plan *p, const void *coef, int W, int H, int layout, const double *var, uint8_t *px,
                          size_t pitch, bool ragged = false, bool rec8 = false)
{
    const int rc = inv_host_queue(p, coef, W, H, layout, var, px, pitch, ragged, rec8);
    if (rc) drain_lanes(p);
    return rc;
}

static int inv_host_queue(dct_cuda_plan *p, const void *coef, int W, int H, int layout, const double *var, uint8_t *px,
                          size_t pitch, bool ragged, bool rec8)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    const int n = p->n, nn = n * n;
    if (rec8 && (nn % 16))
        return fail(DCT_CUDA_EINVAL, "int8 records need a block size whose square is a multiple of 16 (got %d)", n);
    int rc = ragged ? check_ragged(px, coef, pitch, W, H, n) : check_plane(px, coef, pitch, W, H, false, n);
    if (rc) return rc;
    if (layout != DCT_CUDA_NATURAL && layout != DCT_CUDA_ZIGZAG) return fail(DCT_CUDA_EINVAL, "bad layout %d", layout);
    if (p->adaptive && !var) return fail(DCT_CUDA_EINVAL, "adaptive plan needs the per-block variance array");
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    const int Wp = (W + n - 1) / n * n, Hp = (H + n - 1) / n * n;   // == W, H unless ragged
    const int bw = Wp / n, total_rows = Hp / n, rows = strip_rows(Wp, Hp, n);
    if (rows > 0) {
        const size_t dev_pitch = ((size_t)Wp + 15) & ~(size_t)15;
        for (int l = 0; l < kLanes; ++l)
            if ((rc = ensure_strip_buffers(p, p->lane[l], (size_t)rows * n * dev_pitch))) return rc;
        int idx = 0;
        for (int r0 = 0; r0 < total_rows; r0 += rows, ++idx) {
            Lane &ln = p->lane[idx % kLanes];
            const int nr = std::min(rows, total_rows - r0);
            const int have = std::min(nr * n, H - r0 * n);               // image rows in this strip
            const size_t nb = (size_t)nr * bw, b0 = (size_t)r0 * bw;
            if (rec8) {
                int8_t *d8 = reinterpret_cast<int8_t *>(ln.d_coef) + ln.cap_blocks * 2;
                CU_TRY(cudaMemcpyAsync(d8, (const int8_t *)coef + b0 * nn, nb * nn, cudaMemcpyHostToDevice, ln.stream));
                CU_TRY(launch_widen_records(d8, ln.d_coef, nb * nn, ln.stream));
                ++p->launches;
            } else {
                CU_TRY(cudaMemcpyAsync(ln.d_coef, (const int16_t *)coef + b0 * nn, nb * nn * 2, cudaMemcpyHostToDevice, ln.stream));
            }
            if (p->adaptive)
                CU_TRY(cudaMemcpyAsync(ln.d_var, var + b0, nb * sizeof(double), cudaMemcpyHostToDevice, ln.stream));
            if ((rc = queue_inv(p, ln, ln.d_coef, Wp, nr * n, layout, ln.d_var, ln.d_px, dev_pitch, ln.stream))) return rc;
            if (pitch == dev_pitch && (size_t)W == pitch)                // dense rows on both sides: one linear copy
                CU_TRY(cudaMemcpyAsync(px + (size_t)r0 * n * pitch, ln.d_px, pitch * (size_t)have, cudaMemcpyDeviceToHost, ln.stream));
            else
                CU_TRY(cudaMemcpy2DAsync(px + (size_t)r0 * n * pitch, pitch, ln.d_px, dev_pitch, (size_t)W, (size_t)have,
                                         cudaMemcpyDeviceToHost, ln.stream));
        }
    }
    return DCT_CUDA_OK;
}

extern "C" int dct_cuda_dequant_idct_u8_async(dct_cuda_plan *p, const int16_t *coef, int W, int H, int layout,
                                              const double *var, uint8_t *px, size_t pitch)
{
    return inv_host_async(p, coef, W, H, layout, var, px, pitch);
}

// ---- planes of any size: edge blocks completed by replication (not in the reference: src/dct.c:109-120
// reads out of bounds there) ----
extern "C" int dct_cuda_fwd_quant_u8_edge(dct_cuda_plan *p, const uint8_t *px, size_t pitch, int W, int H, int16_t *coef,
                                          int layout, double *var, dct_cuda_stats *stats)
{
    int rc = fwd_host_async(p, px, pitch, W, H, coef, layout, var, 1, true);
    if (rc) return rc;
    return dct_cuda_plan_wait(p, stats);
}

extern "C" int dct_cuda_dequant_idct_u8_edge(dct_cuda_plan *p, const int16_t *coef, int W, int H, int layout,
                                             const double *var, uint8_t *px, size_t pitch, dct_cuda_stats *stats)
{
    int rc = inv_host_async(p, coef, W, H, layout, var, px, pitch, true);
    if (rc) return rc;
    return dct_cuda_plan_wait(p, stats);
}

extern "C" int dct_cuda_dequant_idct_u8(dct_cuda_plan *p, const int16_t *coef, int W, int H, int layout,
                                        const double *var, uint8_t *px, size_t pitch, dct_cuda_stats *stats)
{
    int rc = dct_cuda_dequant_idct_u8_async(p, coef, W, H, layout, var, px, pitch);
    if (rc) return rc;
    return dct_cuda_plan_wait(p, stats);
}

// ---- int8 records over PCIe (narrow.cu): same calls, half the record bytes, when the table allows it ----
extern "C" int dct_cuda_plan_records_fit_i8(const dct_cuda_plan *p) { return p && p->fits_i8 ? 1 : 0; }

extern "C" int dct_cuda_fwd_quant_u8_i8_async(dct_cuda_plan *p, const uint8_t *px, size_t pitch, int W, int H,
                                              int8_t *coef8, int layout, double *var)
{
    return fwd_host_async(p, px, pitch, W, H, coef8, layout, var, 1, false, true);
}

extern "C" int dct_cuda_dequant_idct_i8_u8_async(dct_cuda_plan *p, const int8_t *coef8, int W, int H, int layout,
                                                 const double *var, uint8_t *px, size_t pitch)
{
    return inv_host_async(p, coef8, W, H, layout, var, px, pitch, false, true);
}

extern "C" int dct_cuda_fwd_quant_u8_i8(dct_cuda_plan *p, const uint8_t *px, size_t pitch, int W, int H, int8_t *coef8,
                                        int layout, double *var, dct_cuda_stats *stats)
{
    int rc = dct_cuda_fwd_quant_u8_i8_async(p, px, pitch, W, H, coef8, layout, var);
    if (rc) return rc;
    return dct_cuda_plan_wait(p, stats);
}

extern "C" int dct_cuda_dequant_idct_i8_u8(dct_cuda_plan *p, const int8_t *coef8, int W, int H, int layout,
                                           const double *var, uint8_t *px, size_t pitch, dct_cuda_stats *stats)
{
    int rc = dct_cuda_dequant_idct_i8_u8_async(p, coef8, W, H, layout, var, px, pitch);
    if (rc) return rc;
    return dct_cuda_plan_wait(p, stats);
}

extern "C" void dct_cuda_record8_to_block(const int8_t *rec, int layout, int **block)
{
    for (int k = 0; k < 64; ++k) {
        const int nat = layout == DCT_CUDA_ZIGZAG ? kZigZag.nat[k] : k;
        block[nat >> 3][nat & 7] = rec[k];
    }
}

// ------------------------------------------------------------------------------------------
// several GPUs: block-row ranges, one host thread per plan, no inter-GPU traffic
// ------------------------------------------------------------------------------------------
template <typename F> static int run_sharded(dct_cuda_plan *const *plans, int n, int H, dct_cuda_stats *stats, F &&body)
{
    if (!plans || n <= 0) return fail(DCT_CUDA_EINVAL, "no plans");
    const int total_rows = H / 8;
    std::vector<int> rcs(n, 0);
    std::vector<dct_cuda_stats> sts(n);
    std::vector<std::string> errs(n);
    std::vector<std::thread> th;
    for (int g = 0; g < n; ++g) {
        const int r0 = (int)((long long)total_rows * g / n), r1 = (int)((long long)total_rows * (g + 1) / n);
        th.emplace_back([&, g, r0, r1]() {
            memset(&sts[g], 0, sizeof sts[g]);
            rcs[g] = (r1 > r0) ? body(plans[g], r0, r1, &sts[g]) : 0;
            if (rcs[g]) errs[g] = dct_cuda_last_error();
        });
    }
    for (auto &t : th) t.join();
    dct_cuda_stats sum{};
    for (int g = 0; g < n; ++g) {
        if (rcs[g]) return fail(rcs[g], "GPU shard %d: %s", g, errs[g].c_str());
        sum.blocks += sts[g].blocks, sum.replayed_blocks += sts[g].replayed_blocks;
        sum.near_ties += sts[g].near_ties, sum.saturated += sts[g].saturated;
    }
    if (stats) *stats = sum;
    return DCT_CUDA_OK;
}

extern "C" int dct_cuda_fwd_quant_u8_multi(dct_cuda_plan *const *plans, int n, const uint8_t *px, size_t pitch, int W,
                                           int H, int16_t *coef, int layout, double *var, dct_cuda_stats *stats)
{
    int rc = check_plane(px, coef, pitch, W, H, false);
    if (rc) return rc;
    for (int g = 0; plans && g < n; ++g)
        if (plans[g] && plans[g]->n != 8) return fail(DCT_CUDA_EINVAL, "the multi-GPU helpers are 8x8 only");
    const size_t bw = (size_t)W / 8;
    return run_sharded(plans, n, H, stats, [&](dct_cuda_plan *p, int r0, int r1, dct_cuda_stats *st) {
        return dct_cuda_fwd_quant_u8(p, px + (size_t)r0 * 8 * pitch, pitch, W, (r1 - r0) * 8, coef + (size_t)r0 * bw * 64,
                                     layout, var ? var + (size_t)r0 * bw : nullptr, st);
    });
}

extern "C" int dct_cuda_dequant_idct_u8_multi(dct_cuda_plan *const *plans, int n, const int16_t *coef, int W, int H,
                                              int layout, const double *var, uint8_t *px, size_t pitch,
                                              dct_cuda_stats *stats)
{
    int rc = check_plane(px, coef, pitch, W, H, false);
    if (rc) return rc;
    for (int g = 0; plans && g < n; ++g)
        if (plans[g] && plans[g]->n != 8) return fail(DCT_CUDA_EINVAL, "the multi-GPU helpers are 8x8 only");
    const size_t bw = (size_t)W / 8;
    return run_sharded(plans, n, H, stats, [&](dct_cuda_plan *p, int r0, int r1, dct_cuda_stats *st) {
        return dct_cuda_dequant_idct_u8(p, coef + (size_t)r0 * bw * 64, W, (r1 - r0) * 8, layout,
                                        var ? var + (size_t)r0 * bw : nullptr, px + (size_t)r0 * 8 * pitch, pitch, st);
    });
}

// ------------------------------------------------------------------------------------------
// K5: run-length symbols for the host entropy coder (src/entropy.c:216-256)
// ------------------------------------------------------------------------------------------
extern "C" int dct_cuda_rle_count_dev(dct_cuda_plan *p, const int16_t *d_coef, size_t nblocks, uint32_t *d_offsets,
                                      uint64_t *total_symbols, void *stream)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    if (p->n != 8) return fail(DCT_CUDA_EINVAL, "run-length symbols are implemented for 8x8 records only");
    if ((nblocks && !d_coef) || !d_offsets) return fail(DCT_CUDA_EINVAL, "NULL data pointer");
    if (nblocks >= (1u << 26)) return fail(DCT_CUDA_EINVAL, "fewer than 2^26 records per call (32-bit symbol offsets: 64 symbols per record)");
    if ((uintptr_t)d_coef % 16) return fail(DCT_CUDA_EINVAL, "coefficients must be 16-byte aligned");
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    cudaStream_t s = (cudaStream_t)stream;
    if (!p->d_rle_total) CU_TRY(cudaMalloc(&p->d_rle_total, sizeof(unsigned long long)));
    const size_t ctas = (nblocks + 255) / 256;
    if (p->rle_sums_cap < ctas + 1) {
        CU_TRY(cudaStreamSynchronize(s));
        if (p->d_rle_sums) CU_TRY(cudaFree(p->d_rle_sums));
        p->d_rle_sums = nullptr;
        CU_TRY(cudaMalloc(&p->d_rle_sums, (ctas + 1) * sizeof(uint32_t)));
        p->rle_sums_cap = ctas + 1;
    }
    unsigned long long total = 0;
    if (nblocks == 0) {
        CU_TRY(cudaMemsetAsync(d_offsets, 0, sizeof(uint32_t), s));
    } else {
        CU_TRY(launch_rle_count(d_coef, (uint32_t)nblocks, d_offsets, p->d_rle_sums, p->d_rle_total, s));
        CU_TRY(cudaMemcpyAsync(&total, p->d_rle_total, sizeof total, cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (total_symbols) *total_symbols = total;
    return DCT_CUDA_OK;
}

extern "C" int dct_cuda_rle_emit_dev(dct_cuda_plan *p, const int16_t *d_coef, size_t nblocks, int layout,
                                     const uint32_t *d_offsets, dct_cuda_rle_symbol *d_symbols, void *stream)
{
    if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan");
    if (nblocks == 0) return DCT_CUDA_OK;
    if (!d_coef || !d_offsets || !d_symbols) return fail(DCT_CUDA_EINVAL, "NULL data pointer");
    if (nblocks >= (1u << 26)) return fail(DCT_CUDA_EINVAL, "fewer than 2^26 records per call (32-bit symbol offsets: 64 symbols per record)");
    if (layout != DCT_CUDA_NATURAL && layout != DCT_CUDA_ZIGZAG) return fail(DCT_CUDA_EINVAL, "bad layout %d", layout);
    if (((uintptr_t)d_coef % 16) || ((uintptr_t)d_symbols % 8)) return fail(DCT_CUDA_EINVAL, "misaligned buffer");
    if (nblocks == 0) return DCT_CUDA_OK;
    DeviceGuard g(p->device);
    std::lock_guard<std::mutex> plan_lock(p->mu);
    CU_TRY(launch_rle_emit(d_coef, (uint32_t)nblocks, layout, d_offsets, d_symbols, (cudaStream_t)stream));
    return DCT_CUDA_OK;
}

