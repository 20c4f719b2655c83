// Host helper shared by the kernels that build their own TMA tensor maps (implemented in gemm.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace st5 {
// rank-4 bf16 tensor map, 128B swizzle, zero fill for out-of-bounds elements. dims / box in elements (innermost
// first), strides in BYTES for dims 1..3 (multiples of 16). Returns 0 on success, a negative library code otherwise.
int encode_bf16_map_4d(CUtensorMap* map, const void* ptr, const uint64_t dims[4], const uint64_t strides_bytes[3],
                       const uint32_t box[4]);
int encode_map_4d(CUtensorMap* map, const void* ptr, int is_f32, const uint64_t dims[4],
                  const uint64_t strides_bytes[3], const uint32_t box[4], int swizzle_bytes);
int device_sm_count();
}  // namespace st5
