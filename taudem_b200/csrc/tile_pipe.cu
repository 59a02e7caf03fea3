// Host side of the stencil tile pipeline (tile_pipe.cuh): tensor maps and the persistent grid size.
#include <cuda.h>
#include <cuda_runtime.h>

#include <map>
#include <mutex>

#include "ctx.h"
#include "tile_pipe.cuh"

namespace td {
namespace {
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiled encode_fn() {
  static EncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiled>(p);
  });
  return fn;
}
}  // namespace

int make_tile_map(TileMap* tm, const void* base, int elem, int pitch, int rows, int box_w, int box_h) {
  EncodeTiled enc = encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return TD_ERR_CUDA; }
  const CUtensorMapDataType dt = elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch * (cuuint64_t)elem};      // bytes between rows (pitch % 32 == 0: a multiple of 16)
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tm->m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")"); return TD_ERR_CUDA; }
  return TD_OK;
}

int stencil_grid(const void* kernel, int threads, size_t smem, long long ntiles, int* grid) {
  struct Key { const void* k; int dev; bool operator<(const Key& o) const { return k < o.k || (k == o.k && dev < o.dev); } };
  static std::map<Key, int> cache;
  static std::mutex mu;
  int dev = 0;
  TD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(Key{kernel, dev});
  int per_dev;
  if (it == cache.end()) {
    int sms = 0, occ = 0;
    TD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    if (occ < 1) { set_error("stencil kernel does not fit on an SM"); return TD_ERR_CUDA; }
    per_dev = sms * occ;
    cache[Key{kernel, dev}] = per_dev;
  } else per_dev = it->second;
  *grid = (int)(ntiles < per_dev ? ntiles : per_dev);
  return TD_OK;
}
}  // namespace td
