// placeholder until the tcgen05 attention kernel lands (next commit)
#include "common.cuh"
extern "C" int samq_attn_relpos_fwd(const void* qkv, const void* rel_pos_h, const void* rel_pos_w,
                                    void* out, int B, int H, int W, int heads, int hd, float scale,
                                    int relw_mode, void* stream) {
  samq::set_error("samq_attn_relpos_fwd: not built yet");
  return SAMQ_ERR_LAUNCH;
}
