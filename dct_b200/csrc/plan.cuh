// plan.cuh -- internals shared by the shim*.cu files: the plan object, error reporting, and the
// helpers that queue K1 / K2 / K3 for one device-resident plane.  Not installed; the public
// interface is include/dct_cuda.h.
#pragma once
#include <dct_cuda.h>

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "band_tables.h"
#include "butterfly.cuh"
#include "kernels.cuh"

namespace dctb {
namespace shim {

// records the message for dct_cuda_last_error() (thread-local) and returns `code`
int fail(int code, const char *fmt, ...);

#define CU_TRY(expr)                                                                                 \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return ::dctb::shim::fail(DCT_CUDA_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                      __FILE__, __LINE__);                                           \
    } while (0)

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
constexpr int kLanes = 3;                       // depth of the host-plane pipeline
constexpr size_t kStripPixels = 32u << 20;      // ~32 Mpx per strip (measured: 16 -> 32 Mpx is +4 % on two overlapped plans, -1 % on one)

struct Lane {
    cudaStream_t stream = nullptr;
    Counters *d_ctr = nullptr;
    uint32_t *d_wl = nullptr;
    uint32_t wl_cap = 0;
    uint8_t *d_side = nullptr;                  // pixels of the flagged blocks, 64 B per worklist slot (first side_cap slots)
    uint32_t side_cap = 0;
    uint32_t *d_seg_count = nullptr;            // K1's per-warp worklist counts (kMaxWorklistSegments entries)
    // strip buffers of the host-plane pipeline
    uint8_t *d_px = nullptr;
    int16_t *d_coef = nullptr;
    double *d_var = nullptr;
    size_t cap_blocks = 0;
    uint64_t blocks = 0;                        // blocks queued since the last stats fetch
    // optional per-kernel timing (dct_cuda_plan_profile): event pairs around K1 / K2 launches
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_fwd, ev_inv;
};

}  // namespace shim
}  // namespace dctb

struct dct_cuda_plan {
    using Lane = dctb::shim::Lane;
    using ExactTables = dctb::ExactTables;
    using Counters = dctb::Counters;
    static constexpr int kLanes = dctb::shim::kLanes;
    int device = 0;
    const DCTContext *dct = nullptr;
    const QuantContext *quant = nullptr;
    int adaptive = 0;
    int n = 8;                                  // block size; != 8 routes every plane call to K6 (generic_n.cu)
    double *d_gen = nullptr;                    // K6 tables: D, Q, R (n*n doubles each)
    int *d_gen_pos = nullptr;                   // K6: zigzag position of each natural index
    bool exotic = false;                        // tables outside the fast path's proven domain
    ExactTables h_tab;
    ExactTables *d_tab = nullptr;
    float r[64], thr[64], thr_min;              // K1
    float thr_f32[64];                          // K1 from float pixel tiles
    int uniform_band;
    float rs[64], gain[64], band_floor;         // K2
    float rg[64];                               // K2: rs * gain rounded up
    Lane lane[kLanes];
    Counters *h_ctr = nullptr;                  // pinned, kLanes entries
    bool profile = false;
    uint32_t *d_rle_sums = nullptr;             // K5 workspace: per-CTA symbol totals + the grand total
    size_t rle_sums_cap = 0;
    unsigned long long *d_rle_total = nullptr;
    bool fp64_inverse = false;                  // DCT_CUDA_INV_FP64=1: adaptive plans decode through the fp64 K2
    bool skip_replay = false;                   // test hook: leave K1/K2's fast-path values unpatched
    // whole-frame RGB 4:2:0 calls (luma plan only): device copy of the frame, its planes and records
    uint8_t *d_frame = nullptr;
    size_t frame_cap = 0;
    cudaEvent_t ev_peer = nullptr;              // *_peer calls: "input ready" (owner) / "shard done" (peers)
    uint64_t launches = 0;                      // kernels launched for this plan (dct_cuda_plan_kernel_launches)
    bool no_tma = false;                        // set around the peer calls: planes mapped from another GPU keep the cp.async / LDG kernels
    bool fits_i8 = false;                       // every quantised value of a uint8 plane fits int8 (narrow.cu)
    std::mutex mu;                              // serialises the entry points on one plan (lanes and buffers are state)
};

namespace dctb {
namespace shim {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int read_tables(dct_cuda_plan *p);
int upload_generic_tables(dct_cuda_plan *p);
int ensure_worklist(Lane &ln, size_t nblocks);
// `dev`: the pitch is used by the kernels directly (8-byte rows); host planes are re-packed by the copy
int check_plane(const void *a, const void *b, size_t pitch, int W, int H, bool dev, int n = 8);
int check_ragged(const void *a, const void *b, size_t pitch, int W, int H, int n);
// queue K1 (+K3) / K2 (+K3) for one device-resident plane on lane `ln`, stream `s`
int queue_fwd(dct_cuda_plan *p, Lane &ln, const uint8_t *d_px, size_t pitch, int W, int H, int16_t *d_coef,
              int layout, double *d_var, cudaStream_t s, int elem = 1);
int queue_inv(dct_cuda_plan *p, Lane &ln, const int16_t *d_coef, int W, int H, int layout, const double *d_var,
              uint8_t *d_px, size_t pitch, cudaStream_t s);
int ensure_strip_buffers(dct_cuda_plan *p, Lane &ln, size_t pixels, int elem = 1);
int collect_stats(dct_cuda_plan *p, dct_cuda_stats *out, cudaStream_t user_stream);

}  // namespace shim
}  // namespace dctb
