#include "fe_gemm_tables.h"

int64_t fe_gemm_plan_layout(const b200fe_params* p, fe_blob_header* h, int64_t off) {
  (void)p;
  h->gemm_ok = 0;
  return off;
}

int32_t fe_gemm_pack(const b200fe_params* p, const fe_blob_header* h, const float* window,
                     const float* fbank, char* base) {
  (void)p; (void)h; (void)window; (void)fbank; (void)base;
  return B200FE_OK;
}
