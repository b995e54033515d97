// tma_host.cu -- tensor maps for the bulk-tensor tile copies of K1 / K2 (see tma.cuh).
// cuTensorMapEncodeTiled is reached through cudaGetDriverEntryPoint, so libdct_cuda keeps its single
// link-time dependency (the static CUDA runtime) and still loads on a box without a driver.
#include <cstdlib>
#include <mutex>
#include <set>
#include <utility>

#include "kernels.cuh"
#include "tma.cuh"

namespace dctb {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    });
    return fn;
}

}  // namespace

// L2 fetch granularity of the tensor maps (tuning aid: DCT_CUDA_L2PROMO = 0 none, 64, 128, 256; default 256)
static CUtensorMapL2promotion l2_promotion()
{
    static const int v = getenv("DCT_CUDA_L2PROMO") ? atoi(getenv("DCT_CUDA_L2PROMO")) : 256;
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                  : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

bool tma_available() { return encode_fn() != nullptr; }

bool pdl_enabled()
{
    static const bool on = getenv("DCT_CUDA_NO_PDL") == nullptr;
    return on;
}

cudaError_t ensure_smem_attributes(const void *kernel, int smem_bytes)
{
    static std::mutex mu;
    static std::set<std::pair<int, const void *>> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({dev, kernel})) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    done.insert({dev, kernel});
    return cudaSuccess;
}

cudaError_t make_pixel_map(CUtensorMap *map, const void *base, long long pitch, int W, int H)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {256, 8}, estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2_promotion(),
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t make_record_map(CUtensorMap *map, const void *base, uint32_t nblocks, int rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    const cuuint64_t dims[2] = {128, (cuuint64_t)nblocks};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {128, (cuuint32_t)rows}, estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2_promotion(),
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace dctb
