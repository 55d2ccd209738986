// pp_scs.cu -- Sell-C-sigma construction (placeholder until the device build lands).
#include "pp_internal.cuh"

pp_status pp_scs_build(pp_ps*, const int*, const int*, const void* const*, int, cudaStream_t) {
  pp_set_error("Sell-C-sigma construction is not implemented yet");
  return PP_ERR_UNSUPPORTED;
}
