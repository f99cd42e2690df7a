// F1 (tcgen05 flavour) -- placeholder translation unit until the tensor-core kernel lands.
#include "mp_common.cuh"

int mp_tc_prepare(mp_mlp* mlp) {
  mlp->tc = nullptr;
  mlp->tc_ok = 0;
  return MP_OK;
}
void mp_tc_release(mp_mlp* mlp) { (void)mlp; }
int mp_launch_query_tc(const mp_mlp*, const mp_feat*, const MpPointSrc&, const MpCalib&, const MpOutDst&, cudaStream_t) {
  mp_set_error("tcgen05 kernel not built");
  return MP_E_UNSUPPORTED;
}
