// Tensor-map encoder shared by the tcgen05 kernels.
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tmap.cuh"

namespace dmme {

// ------------------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry point fetched at run time; no -lcuda link)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t dims[5];
  uint64_t strides[4];
  uint32_t box[5];
  uint32_t rank;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return static_cast<size_t>(h);
  }
};

static int encode_map_kind(CUtensorMap* out, const void* ptr, uint32_t rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, bool f32_plain);

int encode_map(CUtensorMap* out, const void* ptr, uint32_t rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_map_kind(out, ptr, rank, dims, strides_bytes, box, false);
}
int encode_map_f32(CUtensorMap* out, const void* ptr, uint32_t rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_map_kind(out, ptr, rank, dims, strides_bytes, box, true);
}

static int encode_map_kind(CUtensorMap* out, const void* ptr, uint32_t rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, bool f32_plain) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr;
  key.rank = rank | (f32_plain ? 0x100u : 0u);
  for (uint32_t i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  for (uint32_t i = 0; i + 1 < rank; ++i) key.strides[i] = strides_bytes[i];
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encoder();
  DMME_REQUIRE(enc != nullptr, DMME_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(out, f32_plain ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                   const_cast<void*>(ptr), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   f32_plain ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DMME_REQUIRE(r == CUDA_SUCCESS, DMME_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
  }
  return 0;
}


}  // namespace dmme
