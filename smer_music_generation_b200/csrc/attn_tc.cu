// tcgen05 attention entry points (bf16, dh = 64): argument checks and dispatch to the kernels in attn_fwd2.cu (forward)
// and attn_bwd2.cu (fused backward).  Arithmetic of F.multi_head_attention_forward's need_weights branch as the
// reference reaches it (transformer.py:389,459,463): softmax(q k^T / sqrt(dh) + masks) v with dropout on P; masks =
// causal flag (the nopeek tgt_mask) and per-key padding.  Scores are never materialised in HBM.
// (The first-generation kernels that lived here -- one query tile per CTA with P staged through shared memory, and a
// dQ + dK/dV kernel pair that each recomputed S, dP and the exponentials -- were replaced in round 2; see DESIGN.md.)
#include "common.cuh"
#include "../../include/smer_b200.h"

int smer_attn_fwd2_launch(const smer_attn_args* a, void* stream);      // attn_fwd2.cu
int smer_attn_bwd2_launch(const smer_attn_args* a, void* stream);      // attn_bwd2.cu

namespace {

constexpr int DH = 64;
constexpr int MASK_WORDS = 512;                    // key-mask bitmap of one batch row in shared memory: Lk <= 16384

int check_common(const smer_attn_args* a, const char* who) {
  if (!a) { smer_set_error("%s: null args", who); return SMER_ERR_ARG; }
  if (a->dtype != SMER_DT_BF16 || a->dh != DH) {
    smer_set_error("%s: bf16 with head dim 64 only (dtype=%d dh=%d)", who, a->dtype, a->dh);
    return SMER_ERR_UNSUPPORTED;
  }
  if (a->add_mask || a->q_pos0 != 0) {
    smer_set_error("%s: additive masks / query offsets are served by the simt kernels", who);
    return SMER_ERR_UNSUPPORTED;
  }
  if (a->causal && a->Lq != a->Lk) { smer_set_error("%s: causal needs Lq == Lk", who); return SMER_ERR_UNSUPPORTED; }
  if (a->Lk > MASK_WORDS * 32) { smer_set_error("%s: Lk <= %d (key-mask bitmap in shared memory)", who, MASK_WORDS * 32); return SMER_ERR_UNSUPPORTED; }
  if (a->dropout_p > 0.45f) { smer_set_error("%s: dropout_p <= 0.45 (packed half-compare keep rule, common.cuh)", who); return SMER_ERR_UNSUPPORTED; }
  if ((a->cu_q == nullptr) != (a->cu_k == nullptr)) { smer_set_error("%s: cu_q and cu_k come together", who); return SMER_ERR_ARG; }
  if (a->cu_q && (a->kv_len || a->key_pad || a->q_rows <= 0 || a->k_rows <= 0)) {
    smer_set_error("%s: packed rows (cu_q/cu_k) exclude kv_len / key_pad and need q_rows / k_rows", who);
    return SMER_ERR_ARG;
  }
  if (a->B <= 0 || a->H <= 0 || a->Lq <= 0 || a->Lk <= 0) { smer_set_error("%s: empty problem", who); return SMER_ERR_ARG; }
  return SMER_OK;
}

}  // namespace

extern "C" int smer_attn_fwd_tc(const smer_attn_args* a, void* stream) {
  int rc = check_common(a, "smer_attn_fwd_tc");
  if (rc) return rc;
  SMER_CHECK_ARG(a->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(a->o) & 15) == 0, "smer_attn_fwd_tc: o must be 16-byte aligned rows");
  return smer_attn_fwd2_launch(a, stream);
}

extern "C" int smer_attn_bwd_tc(const smer_attn_args* a, void* stream) {
  int rc = check_common(a, "smer_attn_bwd_tc");
  if (rc) return rc;
  SMER_CHECK_ARG(a->dout && a->dq && a->dk && a->dv && a->lse && a->dsum && a->o, "smer_attn_bwd_tc: missing buffers");
  SMER_CHECK_ARG(a->lddq % 8 == 0 && a->lddk % 8 == 0 && a->lddv % 8 == 0 && a->ldo % 8 == 0 && a->lddo % 8 == 0,
                 "smer_attn_bwd_tc: row pitches must be multiples of 8 elements");
  return smer_attn_bwd2_launch(a, stream);
}
