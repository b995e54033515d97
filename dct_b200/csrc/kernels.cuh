// kernels.cuh -- parameter blocks and launch entry points shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include <utility>

namespace dctb {

enum { LAYOUT_NATURAL = 0, LAYOUT_ZIGZAG = 1 };

// Device-side counters of one plan (zeroed by the host before each plane call).
struct Counters {
    unsigned int wl_count;         // blocks appended to the replay worklist by K1/K2; zeroed again by K3's last CTA
    unsigned int done_ctas;        // K3 CTAs that have finished (for that reset)
    unsigned long long replayed;   // blocks re-done in fp64 by K3 (== min(wl_count, capacity))
    unsigned long long near_ties;  // fp64 values within 1e-9 of a .5 boundary seen by K3
    unsigned long long saturated;  // fp64 quantised values outside int16 (exotic tables only)
};

// fp64 tables exactly as the host contexts hold them (src/dct.c:19-30, src/quantization.c:51-111)
struct ExactTables {
    double D[64];   // dct_matrix, row-major
    double Q[64];   // quant_matrix
    double R[64];   // dequant_matrix
    double mp64[64];    // K3's fp64 pass: dequantisation multiplier (R, or 1/R when adaptive) * a_u a_v / 8
    // the fused kernels' fp32 tables, so that K3 can repeat their arithmetic bit for bit
    float r32[64];      // K1: 1 / (Q_k * 8 a_u a_v)
    float thr32[64];    // K1: 0.5 - band_k
    float thr32f[64];   // K1 from float pixel tiles: 0.5 - band_k (wider: inputs carry a rounding error)
    float rs32[64];     // K2: dequantisation multiplier * a_u a_v / 8
    float gain32[64];   // K2: error gain per unit |input|
    float rg32[64];     // K2 (non-adaptive): rs32 * gain32 rounded up
    float band_floor;   // K2
    float pad_[3];
    // K3 inverse (one lane per block): sum_k gain_k * |mp64_k| * max(2 - nv), so that bound_per_q * max_k |q_k| bounds
    // sum_k gain_k * |dequantised value k| without a pass over the 64 values
    double bound_per_q;
    double mult64[64];  // dequantisation multiplier in fp64: R (non-adaptive, sic) or 1/R (adaptive); for the fast fp64 settle in K3
};

// K1: forward DCT + quantise.  One thread per 8x8 block.
struct FwdParams {
    const uint8_t *px;     // uint8 pixels, or float pixels (k_fwd_quant_f32) reinterpreted
    long long pitch;       // bytes between pixel rows (multiple of 8)
    uint32_t bw;           // blocks per block-row (W/8)
    uint32_t nblocks;
    int16_t *coef;         // block-major records, 64 x int16 = 128 B per block
    double *var_out;       // adaptive only: per-block spatial variance (may be null)
    uint32_t *worklist;
    uint32_t wl_cap;
    Counters *ctr;
    float r[64];           // natural index: 1 / (Q_k * 8 a_u a_v)
    float thr[64];         // natural index: 0.5 - band_k ; |residual| >= thr  => replay in fp64
    uint32_t step_q, step_r;   // divmod(blocks between a warp's consecutive tiles, bw): set by the launcher
    uint32_t ntiles, tile_stride;   // ceil(nblocks / 32); warps in the grid: set by the launcher
    float thr_min;         // min_k thr[k]: the single threshold of the uniform-band variant
    int uniform_band;      // 1: test max_k |residual| >= thr_min (cheaper, slightly more replays)
    // Optional: the pixels of the flagged blocks, copied next to their worklist entry (64 bytes at side + 64 * slot
    // for slot < side_cap), so that K3 reads a compact, L2-resident array instead of 8 scattered rows per block.
    uint8_t *side;
    uint32_t side_cap;
    // k_fwd_quant_u8 only: every warp of the persistent grid appends to its OWN segment of the worklist (and of the
    // side array) and leaves its count in seg_count[warp] -- no global atomic on the path.  Set by the launcher.
    uint32_t *seg_count;       // one entry per warp of the grid
    uint32_t seg_cap;          // worklist entries per segment (>= 32 * tiles per warp)
    uint32_t side_seg_cap;     // side slots per segment (entries beyond it have no pixel copy)
    uint32_t side_seg_lim;     // bulk-tensor kernel: entries of this plane with a pixel copy (= side_seg_cap unless planes share the segment)
    int no_tma;                // 1: keep the cp.async kernel (planes mapped from a peer GPU)
    // bulk-tensor kernel only: every warp replays the blocks of its own worklist segment at the end of its tile loop
    // (replay_lane.cuh) instead of leaving them to a K3 launch; needs the exact tables
    const ExactTables *tab;
};

// planes up to this many blocks (a 7680x4320 luma plane) replay their flagged blocks in the tail of K1 / K2 instead of a K3 launch
constexpr uint32_t kFoldMaxBlocks = 600000;

// geometry of the segmented worklist a k_fwd_quant_u8 launch produced (consumed by k_replay_fwd_lane)
struct WorklistSegments {
    uint32_t n_segs, seg_cap, side_seg_cap;
};

struct alignas(16) PosNeg2 {
    float2 pos, neg;
};
struct PosNeg1 {
    float pos, neg;
};

// K2: dequantise + inverse DCT.  One thread per 8x8 block.
struct InvParams {
    const int16_t *coef;
    const double *var_in;  // adaptive only
    uint8_t *px;
    long long pitch;
    uint32_t bw;
    uint32_t nblocks;
    uint32_t *worklist;
    uint32_t wl_cap;
    Counters *ctr;
    float rs[64];          // natural index: dequant multiplier * a_u a_v / 8
    float gain[64];        // natural index: error gain of the butterfly per unit |input|
    float band_floor;
    // the non-adaptive kernel's folded first stage (butterfly.cuh idct8_dequant): for column pair c = (A_c, B_c) of
    // (0,4) (2,6) (5,3) (1,7) and first-stage pair j of rows (0,4) (2,6) (5,3) (1,7):
    //   ma[c][j] = (rs[8 a_j + A_c], rs[8 a_j + B_c]),  mb[c][j] = (same for row b_j, and its negation)
    float2 ma[4][4];
    PosNeg2 mb[4][4];
    float rg[64];          // natural index: rs * gain, rounded up (bound on the raw quantised values)
    // bulk-tensor kernel only: every warp of the persistent grid appends to its OWN worklist segment (see FwdParams)
    uint32_t *seg_count;   // one entry per warp of the grid, or null (keeps the one-shot kernel and its atomic append)
    int no_tma;            // 1: keep the one-shot kernel (planes mapped from a peer GPU)
    // bulk-tensor kernel only: every warp replays the blocks of its own worklist segment itself (replay_lane.cuh),
    // in batches of 32 between tiles, instead of leaving them to a K3 launch; needs the exact tables
    const ExactTables *tab;
};

// K3: exact fp64 replay of the blocks on the worklist (or of every block when wl == null).
struct ReplayParams {
    const ExactTables *tab;
    const ExactTables *h_tab;   // the same tables in host memory (launchers copy what a kernel takes as parameters)
    const uint32_t *worklist;   // null => replay all nblocks
    Counters *ctr;
    uint32_t nblocks;
    uint32_t wl_cap;
    uint32_t bw;
    int adaptive;
    int layout;
    long long pitch;
    const uint8_t *px_in;       // forward (uint8 pixels; float pixels when px_is_f32)
    int px_is_f32;
    const uint8_t *side;        // forward: K1's copy of the flagged blocks' pixels (see FwdParams), or null
    uint32_t side_cap;
    const uint32_t *seg_count;  // forward, segmented worklist (see FwdParams): counts per segment, or null
    WorklistSegments seg;
    int16_t *coef_out;
    double *var_out;
    const int16_t *coef_in;     // inverse
    const double *var_in;
    uint8_t *px_out;
};

// `launches` (optional) is incremented by the number of kernel launches made
// `*folded` (optional) is set when the kernel replayed its flagged blocks itself (no K3 launch needed)
cudaError_t launch_fwd_quant_u8(const FwdParams &p, int layout, int adaptive, cudaStream_t s, unsigned *launches = nullptr,
                                WorklistSegments *segments = nullptr, bool *folded = nullptr);
constexpr uint32_t kMaxWorklistSegments = 4096;   // warps of the largest persistent grid this library launches
cudaError_t launch_fwd_quant_f32(const FwdParams &p, int layout, cudaStream_t s);
// `segments`: geometry of the segmented worklist the launch produced (n_segs == 0: flat worklist counted in ctr->wl_count)
cudaError_t launch_dequant_idct_u8(const InvParams &p, int layout, int adaptive, cudaStream_t s, WorklistSegments *segments = nullptr,
                                   bool *folded = nullptr);
// 2 or 3 small non-adaptive planes (one frame's Y, Cb, Cr) in ONE launch each way, every plane replaying its own flagged
// blocks; cudaErrorNotSupported (nothing launched) when the planes do not qualify
cudaError_t launch_fwd_quant_u8_multi(const FwdParams *planes, int n, int layout, cudaStream_t s);
cudaError_t launch_dequant_idct_u8_multi(const InvParams *planes, int n, int layout, cudaStream_t s);
cudaError_t launch_dequant_idct_u8_f64(const InvParams &p, const ExactTables *d_tab, int layout, cudaStream_t s);
cudaError_t launch_replay_fwd(const ReplayParams &p, cudaStream_t s);
cudaError_t launch_replay_inv(const ReplayParams &p, cudaStream_t s);

// K6: plane calls for block sizes other than 8 (generic_n.cu): the reference's arithmetic in fp64
struct GenericParams {
    int n;                  // block size, 1..32
    int blocks_per_cta;     // max(1, 256 / (n*n))
    int adaptive, layout;
    uint32_t bw, nblocks;
    long long pitch;
    const double *D, *Q, *R;         // n*n each, as the host contexts hold them
    const int *pos_of_natural;       // n*n: zigzag position of natural index (src/entropy.c:158-178)
    Counters *ctr;
    const uint8_t *px_in;
    int16_t *coef_out;
    double *var_out;
    const int16_t *coef_in;
    const double *var_in;
    uint8_t *px_out;
};
cudaError_t launch_generic_plane(const GenericParams &p, int forward, cudaStream_t s);

// K5: run-length symbols of the records (rle.cu)
cudaError_t launch_rle_count(const int16_t *d_coef, uint32_t nblocks, uint32_t *d_offsets, uint32_t *d_cta_sums,
                             unsigned long long *d_total, cudaStream_t s);
cudaError_t launch_rle_emit(const int16_t *d_coef, uint32_t nblocks, int layout, const uint32_t *d_offsets, void *d_symbols,
                            cudaStream_t s);

// planar front / back end (planar.cu): colour conversion with 4:2:0 subsampling, edge completion
struct PlanarParams {
    const uint8_t *rgb;        // forward: interleaved R,G,B source
    uint8_t *rgb_out;          // inverse: destination
    long long rgb_pitch;
    int W, H;                  // the image
    uint8_t *y, *cb, *cr;      // planes (written by the forward kernel, read by the inverse one)
    long long y_pitch, c_pitch;
    int y_w, y_h, c_w, c_h;    // padded plane sizes (multiples of the block size)
    int vec_ok;                // every base pointer / pitch allows the 16-byte path
};
cudaError_t launch_rgb_to_ycbcr420(const PlanarParams &p, cudaStream_t s);
cudaError_t launch_ycbcr420_to_rgb(const PlanarParams &p, cudaStream_t s);
cudaError_t launch_pad_edges(uint8_t *px, long long pitch, int W, int H, int Wp, int Hp, int elem, cudaStream_t s);

// int16 <-> int8 records for the PCIe-bound host-plane calls (narrow.cu); n = coefficients, a multiple of 64
cudaError_t launch_narrow_records(const int16_t *d_in, int8_t *d_out, size_t n, Counters *ctr, cudaStream_t s);
cudaError_t launch_widen_records(const int8_t *d_in, int16_t *d_out, size_t n, cudaStream_t s);

// generic-N single block kernels behind the per-block drop-in API (K4/K6)
cudaError_t launch_block_dct_f64(int n, const double *d_D, const double *d_in, double *d_out, int inverse,
                                 cudaStream_t s);
cudaError_t launch_block_quantize_f64(int n, const double *d_Q, int adaptive, double variance,
                                      const double *d_c, int *d_q, cudaStream_t s);
cudaError_t launch_block_dequantize_f64(int n, const double *d_R, int adaptive, double variance,
                                        const int *d_q, double *d_c, cudaStream_t s);

// cudaFuncSetAttribute(max dynamic shared memory, carveout) once per (device, kernel): the two calls cost more host
// time than the launch itself, which matters for single frames (launch-bound).  Thread-safe.
cudaError_t ensure_smem_attributes(const void *kernel, int smem_bytes);

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// A step is four short launches on one stream (K1, K3, K2, K3).  Launched with programmatic stream serialisation, a
// kernel's CTAs are scheduled while its predecessor drains, run their prologue (tables to shared memory, barrier
// set-up) and block in pdl_wait() until the predecessor's results are visible -- the launch gap and the prologue
// disappear from the critical path, which is most of a single frame's latency.  Every such kernel calls pdl_wait()
// BEFORE its first access to memory another kernel may have written or may still read, so the results are those of
// plain stream order.  DCT_CUDA_NO_PDL=1 (measurement aid) launches them the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename P>
cudaError_t launch_pdl(void (*kernel)(P), unsigned grid, unsigned block, size_t smem, cudaStream_t s, const P &params)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = smem, cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, params);
}

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E)
template <int B, int E, typename F> __device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

}  // namespace dctb
