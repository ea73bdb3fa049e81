// placeholder until the tcgen05 projection lands: reports "unsupported" so tg_proj runs the exact path.
#include "common.cuh"
#include "kernels.h"
int tg_proj_tc_impl(cudaStream_t, const float*, int, const float*, int, const float*, float*, int, int, int, int, int) {
  tg_set_error("proj_tc: tensor-core projection not built");
  return TG_ERR_UNSUPPORTED;
}
