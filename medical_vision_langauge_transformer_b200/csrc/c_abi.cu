// Library-level entry points of libmvlt_b200.so (see include/mvlt_b200.h).
#include "common.cuh"

extern "C" int mvlt_gemm_tc_init(void);
extern "C" int mvlt_attn_init(void);

extern "C" int mvlt_abi_version(void) { return 1; }

extern "C" int mvlt_init(void) {
  int rc = mvlt_gemm_tc_init();
  if (rc != MVLT_OK) return rc;
  return mvlt_attn_init();
}
