// Host-side TMA tensor-map construction shared by the tcgen05 kernels.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace dmme {
// bf16 tiled map with 128-byte swizzle and zero OOB fill; cached by (pointer, geometry).
// dims/box innermost first; strides_bytes has rank-1 entries (dims 1..rank-1).
int encode_map(CUtensorMap* out, const void* ptr, uint32_t rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box);
// fp32 tiled map without swizzle (zero OOB fill): the image patches of the input conv
int encode_map_f32(CUtensorMap* out, const void* ptr, uint32_t rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);
}  // namespace dmme
