// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace sdfr {
int build_tc_tables(sdfr_decoder* dec, const sdfr_decoder_spec*, const float* const*) { dec->tc.ok = 0; return SDFR_OK; }
int launch_mlp_tc(const sdfr_decoder*, const MlpInputs&, float*, float*, cudaStream_t) {
  set_error("tcgen05 MLP kernel not available for this decoder");
  return SDFR_E_UNSUPPORTED;
}
}
