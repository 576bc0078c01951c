// placeholder until the tcgen05 attention kernel lands (next commit)
#include "common.cuh"
extern "C" int fdm_attn_fwd(const void*, const void*, const void*, void*, const int8_t*, int64_t,
                            int64_t, int64_t, int, int, int64_t, int64_t, int64_t, int64_t, int64_t,
                            int64_t, int, int, float, int, void*) {
  fdm::set_error("attention kernel not built yet");
  return FDM_ERR_UNSUPPORTED;
}
