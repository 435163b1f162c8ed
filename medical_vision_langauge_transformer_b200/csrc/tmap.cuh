// Host-side TMA descriptor helper shared by the tcgen05 kernels (defined in gemm_tc.cu; mvlt_gemm_tc_init() must have
// resolved cuTensorMapEncodeTiled first).
#pragma once
#include <cuda.h>

namespace mvlt {
// 2-D row-major tensor map: dims {cols, rows}, row stride ld_elems, box {box_cols, box_rows}
int make_tmap(CUtensorMap* map, CUtensorMapDataType dt, int elt_bytes, const void* ptr, long long rows, long long cols,
              long long ld_elems, int box_cols, int box_rows, CUtensorMapSwizzle swz, CUtensorMapL2promotion promo);
}  // namespace mvlt
