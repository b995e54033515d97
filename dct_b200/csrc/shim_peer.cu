// shim_peer.cu -- C ABI: NVLink peers working in place on one GPU's memory (dct_cuda_*_peer).
#include "plan.cuh"

using namespace dctb;
using namespace dctb::shim;

// ------------------------------------------------------------------------------------------
// several GPUs, data resident on ONE of them: peers work on the owner's memory through NVLink.
// There is no staging copy and no collective: a peer's K1 / K2 / K3 load their shard straight from the
// owner's HBM and store their results straight into it (peer-mapped addresses), so the transfer is
// part of the kernel's own load / store stream and overlaps its arithmetic tile by tile.
// ------------------------------------------------------------------------------------------
namespace {

int peer_prepare(dct_cuda_plan *const *plans, int n, int H, const float *share, int forward, std::vector<int> *row_end)
{
    if (!plans || n <= 0 || !plans[0]) return fail(DCT_CUDA_EINVAL, "no plans");
    const int owner = plans[0]->device;
    for (int g = 0; g < n; ++g) {
        dct_cuda_plan *p = plans[g];
        if (!p) return fail(DCT_CUDA_EINVAL, "NULL plan %d", g);
        if (p->n != 8) return fail(DCT_CUDA_EINVAL, "the peer calls are 8x8 only");
        if (p->adaptive != plans[0]->adaptive) return fail(DCT_CUDA_EINVAL, "plans differ in their adaptive flag");
        for (int h = 0; h < g; ++h)
            if (plans[h]->device == p->device) return fail(DCT_CUDA_EINVAL, "plans %d and %d share GPU %d", h, g, p->device);
        DeviceGuard dg(p->device);
        if (!p->ev_peer) CU_TRY(cudaEventCreateWithFlags(&p->ev_peer, cudaEventDisableTiming));
        if (g == 0) continue;
        int ok = 0;
        CU_TRY(cudaDeviceCanAccessPeer(&ok, p->device, owner));
        if (!ok) return fail(DCT_CUDA_EINVAL, "GPU %d cannot map the memory of GPU %d", p->device, owner);
        cudaError_t e = cudaDeviceEnablePeerAccess(owner, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return fail(DCT_CUDA_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
    }
    // block-row ranges: peers get `share[g]` of the rows each (default: dct_cuda_peer_default_share), the owner the rest
    const int total = H / 8;
    row_end->assign(n, total);
    double acc = 0.0;
    for (int g = 1; g < n; ++g) {
        const double sh = share ? (double)share[g] : (double)dct_cuda_peer_default_share(plans[0], n, forward);
        if (!(sh >= 0.0) || sh > 1.0) return fail(DCT_CUDA_EINVAL, "share[%d] = %g outside [0, 1]", g, sh);
        acc += sh;
    }
    if (acc > 1.0 + 1e-6) return fail(DCT_CUDA_EINVAL, "the peers' shares add up to %g > 1", acc);
    // the owner's rows come first, then peer 1, 2, ...
    double edge = 1.0 - acc;
    int prev = 0;
    for (int g = 0; g < n; ++g) {
        int end = g == n - 1 ? total : (int)llround(edge * total);
        end = std::max(prev, std::min(end, total));
        (*row_end)[g] = end;
        prev = end;
        if (g + 1 < n) edge += share ? (double)share[g + 1] : (double)dct_cuda_peer_default_share(plans[0], n, forward);
    }
    return DCT_CUDA_OK;
}

template <typename F> int peer_run(dct_cuda_plan *const *plans, int n, const std::vector<int> &row_end, cudaStream_t stream, F &&queue)
{
    dct_cuda_plan *own = plans[0];
    {
        DeviceGuard dg(own->device);
        CU_TRY(cudaEventRecord(own->ev_peer, stream));            // the data the caller queued before us is ready
    }
    for (int g = 1; g < n; ++g) {
        const int r0 = row_end[g - 1], r1 = row_end[g];
        if (r1 <= r0) continue;
        dct_cuda_plan *p = plans[g];
        DeviceGuard dg(p->device);
        std::lock_guard<std::mutex> plan_lock(p->mu);
        cudaStream_t s = p->lane[0].stream;
        CU_TRY(cudaStreamWaitEvent(s, own->ev_peer, 0));
        p->no_tma = true;                    // the shard lives in the owner's memory: plain loads / stores through NVLink
        int rc = queue(p, r0, r1, s);
        p->no_tma = false;
        if (rc) return rc;
        CU_TRY(cudaEventRecord(p->ev_peer, s));
    }
    DeviceGuard dg(own->device);
    std::lock_guard<std::mutex> plan_lock(own->mu);
    if (row_end[0] > 0) {
        int rc = queue(own, 0, row_end[0], stream);
        if (rc) return rc;
    }
    for (int g = 1; g < n; ++g)
        if (row_end[g] > row_end[g - 1]) CU_TRY(cudaStreamWaitEvent(stream, plans[g]->ev_peer, 0));
    return DCT_CUDA_OK;
}

}  // namespace

// Default share of each peer.  Measured on 2 x B200 over NV18 (profiles/peer_r1f.json): through the owner's
// NVLink port a peer's K1 sustains ~270 Gpixel/s and its K2 ~370 Gpixel/s (0.8 - 1.1 TB/s of port traffic, both
// directions together) while the owner alone runs 1500 - 1700 Gpixel/s out of local HBM, and every row handed to a
// peer also costs the owner HBM bandwidth -- so for the fused fp32 path the best split is "all rows to the owner".
// Plans on the exact path (tables outside the fast path's domain: every block goes through the fp64 replay) are
// arithmetic- and latency-bound far below the port's rate (owner alone: ~21 Gpixel/s); there the rows are dealt
// out: evenly for the inverse (measured 1.96x on 2 GPUs), with half weight per peer for the forward, whose
// 8-byte row loads feel the NVLink latency (a peer runs it at about half the owner's rate).
extern "C" float dct_cuda_peer_default_share(const dct_cuda_plan *plan, int n_plans, int forward)
{
    if (!plan || n_plans <= 1 || !plan->exotic) return 0.0f;
    const float w = forward ? 0.5f : 1.0f;
    return w / (1.0f + w * (float)(n_plans - 1));
}

extern "C" int dct_cuda_fwd_quant_u8_peer(dct_cuda_plan *const *plans, int n, const uint8_t *d_px, size_t pitch, int W,
                                          int H, int16_t *d_coef, int layout, double *d_var, const float *share,
                                          void *stream)
{
    int rc = check_plane(d_px, d_coef, pitch, W, H, true);
    if (rc) return rc;
    std::vector<int> row_end;
    if ((rc = peer_prepare(plans, n, H, share, 1, &row_end))) return rc;
    const size_t bw = (size_t)W / 8;
    return peer_run(plans, n, row_end, (cudaStream_t)stream, [&](dct_cuda_plan *p, int r0, int r1, cudaStream_t s) {
        return queue_fwd(p, p->lane[0], d_px + (size_t)r0 * 8 * pitch, pitch, W, (r1 - r0) * 8, d_coef + (size_t)r0 * bw * 64,
                         layout, d_var ? d_var + (size_t)r0 * bw : nullptr, s);
    });
}

extern "C" int dct_cuda_dequant_idct_u8_peer(dct_cuda_plan *const *plans, int n, const int16_t *d_coef, int W, int H,
                                             int layout, const double *d_var, uint8_t *d_px, size_t pitch,
                                             const float *share, void *stream)
{
    int rc = check_plane(d_px, d_coef, pitch, W, H, true);
    if (rc) return rc;
    if (plans && n > 0 && plans[0] && plans[0]->adaptive && !d_var)
        return fail(DCT_CUDA_EINVAL, "adaptive plan needs the per-block variance array");
    std::vector<int> row_end;
    if ((rc = peer_prepare(plans, n, H, share, 0, &row_end))) return rc;
    const size_t bw = (size_t)W / 8;
    return peer_run(plans, n, row_end, (cudaStream_t)stream, [&](dct_cuda_plan *p, int r0, int r1, cudaStream_t s) {
        return queue_inv(p, p->lane[0], d_coef + (size_t)r0 * bw * 64, W, (r1 - r0) * 8, layout,
                         d_var ? d_var + (size_t)r0 * bw : nullptr, d_px + (size_t)r0 * 8 * pitch, pitch, s);
    });
}

