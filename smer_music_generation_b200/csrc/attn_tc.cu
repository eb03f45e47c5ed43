// tcgen05 flash attention (bf16, dh=64) -- placeholder entry points until the kernels land.
#include "common.cuh"
#include "../../include/smer_b200.h"

extern "C" int smer_attn_fwd_tc(const smer_attn_args* a, void* stream) {
  (void)a; (void)stream;
  smer_set_error("smer_attn_fwd_tc: not implemented in this build");
  return SMER_ERR_UNSUPPORTED;
}
extern "C" int smer_attn_bwd_tc(const smer_attn_args* a, void* stream) {
  (void)a; (void)stream;
  smer_set_error("smer_attn_bwd_tc: not implemented in this build");
  return SMER_ERR_UNSUPPORTED;
}
