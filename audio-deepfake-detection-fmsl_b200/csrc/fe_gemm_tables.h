// Host-side layout / packing of the DFT-GEMM variant's constant tables (see fe_gemm.cuh).
#ifndef FE_GEMM_TABLES_H_
#define FE_GEMM_TABLES_H_
#include "fe_common.h"

// Appends the GEMM tables to the layout starting at byte offset `off`; returns the new end offset.
// Leaves h->gemm_ok == 0 (and returns `off`) when the configuration is not supported by the variant.
int64_t fe_gemm_plan_layout(const b200fe_params* p, fe_blob_header* h, int64_t off);
// Fills the GEMM tables of a blob whose layout was planned by fe_gemm_plan_layout.
int32_t fe_gemm_pack(const b200fe_params* p, fe_blob_header* h, const float* window,
                     const float* fbank, char* base);
// 1 when fe_stream.cu carries the tcgen05 kernel; AUTO's measured preference (variant selection)
bool fe_gemm_variant_built(void);
bool fe_gemm_auto_prefers(const b200fe_params* p);
#endif
