// tcgen05 DFT-GEMM variant — placeholder until the kernel lands (reports "not compiled").
#include "fe_gemm.h"

int32_t fe_gemm_compiled(void) { return 0; }
bool fe_gemm_supported(const b200fe_params*) { return false; }
bool fe_gemm_preferred(const b200fe_params*) { return false; }
int64_t fe_gemm_workspace_bytes(const b200fe_params*, int64_t, int64_t) { return 0; }
cudaError_t fe_gemm_launch(const b200fe_params*, const fe_fft_args&, int64_t, int64_t, void*, cudaStream_t, int*) {
  return cudaErrorNotSupported;
}
