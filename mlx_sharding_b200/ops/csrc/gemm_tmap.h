// Cached TMA descriptor factory shared by the GEMM kernels (defined in gemm_tcgen05.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace b200 {
// bf16 row-major [rows, cols] (row stride ld elements) -> tiles of box_rows x 64 elements, 128B swizzle
bool gemm_make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
// 1-byte elements (e4m3) row-major [rows, cols] -> tiles of box_rows x 128 elements (one 128 B swizzle row)
bool gemm_make_tmap_u8(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
}  // namespace b200
