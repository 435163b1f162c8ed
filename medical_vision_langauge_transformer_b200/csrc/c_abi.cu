// Library-level entry points of libmvlt_b200.so (see include/mvlt_b200.h).
#include "common.cuh"

extern "C" int mvlt_gemm_tc_init(void);
extern "C" int mvlt_attn_init(void);

#include <stdlib.h>

static int g_pdl = -1;
int mvlt_pdl_enabled(void) {
  if (g_pdl < 0) {
    const char* e = getenv("MVLT_PDL");
    g_pdl = (e && atoi(e) == 0) ? 0 : 1;
  }
  return g_pdl;
}

extern "C" int mvlt_abi_version(void) { return 1; }

extern "C" int mvlt_init(void) {
  int rc = mvlt_gemm_tc_init();
  if (rc != MVLT_OK) return rc;
  return mvlt_attn_init();
}
