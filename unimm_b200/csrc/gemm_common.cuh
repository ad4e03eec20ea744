// Host helpers shared by the tcgen05 GEMM translation units (gemm_umma.cu, gemm_umma_ln.cu).
#pragma once

#include <cuda.h>

namespace unimm {

// 2-D TMA descriptor of a row-major 16-bit matrix [rows, cols] (leading dimension ld elements) with a
// 64-column x box_rows box and the 128-byte swizzle the UMMA shared-memory descriptors expect; cached.
int gemm_make_map(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out);
int gemm_num_sms();

}  // namespace unimm
