// jp_glm_tc.cu -- GLM tensor-core log-density path (tcgen05 3xTF32).  Placeholder until the kernel lands.
#include "jp_common.cuh"
bool jp_fit_tc_supported(const jp_posterior*, const jp_fit_args*) {
  jp_set_error("tensor-core path not built");
  return false;
}
int jp_fit_tc_launch(jp_posterior*, const jp_fit_args*) {
  jp_set_error("tensor-core path not built");
  return JP_ERR_UNSUPPORTED;
}
void jp_tc_data_free(jp_data*) {}
