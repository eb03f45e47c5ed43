/* smer_b200.h -- C ABI of libsmer_b200.so: hand-written sm_100a CUDA kernels for the SMER
 * transformer compute path (reference: ruiguo-bio/smer_music_generation).
 *
 * The reference has no FFI/plugin layer of its own (it is pure Python over torch ops, SURVEY.md
 * §2.2); the drop-in boundary is the Python module API of model.py.  This header is what the
 * host-side mirror of that API (smer_music_generation_b200/model.py) binds through ctypes, and
 * what any other host (C++, a Flask worker, ...) would bind.  Each entry point names the
 * reference call site whose arithmetic it replaces.
 *
 * Conventions: plain pointers to DEVICE memory (unless stated), explicit sizes/pitches in
 * ELEMENTS, `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success
 * or a negative SMER_ERR_* code; smer_last_error() returns the message of the calling thread's
 * last failure.  No function allocates, frees or synchronises; workspaces are caller-owned.
 * Activations are token-major: row = b*L + l, features contiguous ("batch-first").
 */
#ifndef SMER_B200_H
#define SMER_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMER_B200_VERSION 100

/* dtype codes */
#define SMER_DT_F32 0
#define SMER_DT_BF16 1

/* GEMM epilogue flags */
#define SMER_EPI_RELU 1    /* max(x,0) after bias                                            */
#define SMER_EPI_ACCUM 2   /* C += result (read-modify-write by the owning thread)           */
#define SMER_EPI_ATOMIC 4  /* fp32 atomics into C (split-K partial sums; C pre-zeroed)        */
#define SMER_EPI_GATE 8    /* C = aux>0 ? result/(1-p) : 0 -- backward of dropout(relu(.))   */

/* grammar state bits (generation.py:538-541, 654-671) */
#define SMER_ST_PITCH 1
#define SMER_ST_REST 2
#define SMER_ST_SEP 4
#define SMER_ST_CONTINUE 8

/* sampling modes */
#define SMER_SAMPLE_GREEDY 0
#define SMER_SAMPLE_MULTINOMIAL 1  /* generation.weighted_sampling                           */
#define SMER_SAMPLE_TOP_P 2        /* generation.nucleus                                     */
#define SMER_SAMPLE_TOP_K 3

#define SMER_XENT_MAX_SUMS 16

int smer_version(void);
/* Data-parallel runs: SMs (even count) the persistent GEMM / attention-forward grids leave to the collective kernels that
 * overlap them (reference: DistributedDataParallel's bucketed all-reduce, train.py:629-642).  0 by default; also read
 * from SMER_RESERVED_SMS at load time. */
int smer_set_reserved_sms(int n);
/* Programmatic dependent launch for the decode-step kernels (smer_decode_embed / _attn / _linear, smer_sample_masked): while
 * on, each of them may be scheduled before its predecessor in the stream has drained (it waits with griddepcontrol.wait
 * before touching activations).  Only valid when EVERY kernel between two of them is one of them: the small-batch decode
 * step (generation.py:528-687 at <= 256 pieces) turns it on around its launches; off by default. */
int smer_set_pdl(int on);
const char* smer_last_error(void);
/* 1 when the running device is compute capability 10.x (sm_100a cubins loadable) */
int smer_device_ok(void);
/* Device uint64 added to every dropout seed at kernel run time (NULL = off).  A training step
 * captured into a CUDA graph bumps this counter inside the graph, so replays draw new masks. */
int smer_set_seed_device_ptr(const uint64_t* dev_counter);

/* ---- K1: embedding * sqrt(d) + sinusoidal PE (+dropout).  model.py:91-92, 123-125 ---------- */
int smer_embed_pe_fwd(const int64_t* ids, const float* emb, const float* pe, void* out, int out_dtype,
                      int B, int L, int d, int V, int pos0, float scale, float dropout_p, uint64_t seed,
                      uint64_t site, void* stream);
/* the same for packed rows with explicit positions: out[r] = dropout(emb[ids[r]] * scale + pe[pos[r]]) */
int smer_embed_pe_packed(const int64_t* ids, const int* pos, const float* emb, const float* pe, void* out, int out_dtype,
                         long long rows, int d, int V, float scale, float dropout_p, uint64_t seed, uint64_t site,
                         void* stream);
/* On-GPU collate of a padded batch into packed rows (dataset.py:802-925 pads; this removes the pads again):
 * out_ids[cu[b] + l] = ids[b, l], out_pos[cu[b] + l] = l for l < cu[b+1] - cu[b]; rows in [cu[B], rows_alloc) get id 0. */
int smer_pack_rows(const int64_t* ids, const int* cu, int B, int L, long long rows_alloc, int64_t* out_ids, int* out_pos,
                   void* stream);
/* rows [first_row_dev[0], rows) of a contiguous row-major buffer := 0: the ghost rows past the last sequence of a packed
 * batch, which no attention kernel writes (first_row_dev = &cu[B], DEVICE memory, so a captured step needs no host value) */
int smer_zero_tail_rows(void* buf, long long row_bytes, long long rows, const int* first_row_dev, void* stream);
/* backward of the nn.Embedding gather: demb[ids] += dout * scale * dropmask */
int smer_embed_bwd(const int64_t* ids, const void* dout, int dtype, float* demb, int B, int L, int d, int V,
                   float scale, float dropout_p, uint64_t seed, uint64_t site, void* stream);

/* ---- K6/K7/K8: y = LayerNorm(resid + dropout(branch)).  transformer.py:391-395,461-469 ------ */
int smer_layernorm_fwd(const void* branch, const void* resid, const float* gamma, const float* beta,
                       void* z_out, void* y, float* mean, float* rstd, int dtype, long long rows, int d,
                       float eps, float dropout_p, uint64_t seed, uint64_t site, void* stream);
/* dbias (nullable): += column sums of the branch gradient = bias gradient of the Linear feeding the branch */
int smer_layernorm_bwd(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma,
                       void* dz, void* dbranch, float* dgamma, float* dbeta, float* dbias, int dtype, long long rows,
                       int d, float dropout_p, uint64_t seed, uint64_t site, void* stream);

/* ---- K2/K3/K6/K7/K9: nn.Linear products (transformer.py:362-364, model.py:82) --------------- */
/* CUDA-core fp32-accumulate GEMM, arbitrary strides: C[m,n] = epi(sum_k A[m*sam+k*sak]*B[n*sbn+k*sbk]) */
int smer_gemm_simt(const void* A, long long sam, long long sak, const void* B, long long sbn, long long sbk,
                   void* C, long long ldc, int in_dtype, int out_dtype, int M, int N, int K, const float* bias,
                   const void* resid, long long ldr, int flags, float dropout_p, uint64_t seed, uint64_t site,
                   int split_k, void* stream);
/* tcgen05/TMEM/TMA bf16 GEMM.  a_kmajor: A is [M,K] with K contiguous (else [K,M] with M
 * contiguous); b_kmajor: B is [N,K] with K contiguous (else [K,N] with N contiguous).
 * lda/ldb are the pitches of the stored matrices.  C is [M,N] row-major, bf16 or fp32.  */
int smer_gemm_bf16_tc(const void* A, long long lda, int a_kmajor, const void* B, long long ldb, int b_kmajor,
                      void* C, long long ldc, int out_dtype, int M, int N, int K, const float* bias,
                      const void* resid, long long ldr, int flags, float dropout_p, uint64_t seed,
                      uint64_t site, int split_k, float* colsum /* nullable: [N] fp32, += column sums of C
                      (GATE epilogue only: the bias gradient of the Linear whose input gradient C is) */,
                      void* stream);
/* out[c] += sum_r x[r,c]  (bias gradients) */
int smer_colsum(const void* x, int dtype, long long ld, float* out, long long rows, int cols, void* stream);

/* ---- K4/K5: attention core.  transformer.py:389,459,463 -> F.multi_head_attention_forward ---- */
typedef struct smer_attn_args {
  const void *q, *k, *v;      /* token-major, head h at column h*dh; pitches ldq/ldk/ldv      */
  void* o;                    /* forward output / backward input                              */
  const void* dout;           /* backward: gradient of o                                      */
  void *dq, *dk, *dv;         /* backward outputs                                             */
  float *dbq, *dbk, *dbv;     /* backward (tc kernels), nullable: [H*dh] fp32, += column sums of dq/dk/dv over
                                 all tokens = the in-projection's bias gradient (transformer.py MHA in_proj_bias) */
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  float* lse;                 /* [B,H,Lq] log-sum-exp of the masked scaled scores             */
  float* dsum;                /* [B,H,Lq] backward scratch: rowsum(dO*O)                      */
  const uint8_t* key_pad;     /* [B,Lk], 1 = key masked (key_padding_mask) or NULL            */
  const int* kv_len;          /* [B] |v| = 1+last unmasked key (loop bound) or NULL; v >= 0 asserts that the mask is a
                                 pure suffix (key j masked <=> j >= v: key_pad is not read), v < 0: it has holes */
  const float* add_mask;      /* [Lq,Lk] additive float mask or NULL (simt kernels only)      */
  long long ld_mask;
  int B, H, Lq, Lk, dh;
  int dtype;
  int causal;                 /* tgt_mask == nopeek mask: key j visible iff j <= i + q_pos0    */
  int q_pos0;
  float scale;                /* 1/sqrt(dh)                                                   */
  float dropout_p;
  uint64_t seed, site;
  void* dq_accum;             /* backward (tc kernels): fp32 [B*Lq, H*dh] workspace; the fused backward accumulates the
                                 unscaled dQ of all key tiles here (it is zeroed by the call) before converting to dq */
  /* Padding-free ("varlen") layout, tc kernels only: sequence b owns query rows [cu_q[b], cu_q[b+1]) and key rows
   * [cu_k[b], cu_k[b+1]) of the packed token-major buffers (int32 prefix sums, B+1 entries each, DEVICE memory).  Lq / Lk
   * are then the LONGEST query / key sequence (they size the grid), kv_len / key_pad must be NULL (every key of a
   * sequence is visible), and lse / dsum are head-major: [H, cu_q[B]].  q_rows / k_rows = the packed buffers' row counts. */
  const int *cu_q, *cu_k;
  long long q_rows, k_rows;
} smer_attn_args;
int smer_attn_fwd_simt(const smer_attn_args* a, void* stream);
int smer_attn_bwd_simt(const smer_attn_args* a, void* stream);
int smer_attn_fwd_tc(const smer_attn_args* a, void* stream);   /* bf16, dh=64, tcgen05 */
int smer_attn_bwd_tc(const smer_attn_args* a, void* stream);
/* head-averaged probabilities (B,Lq,Lk): the reference decoder's second return value */
int smer_attn_weights(const smer_attn_args* a, float* weights, long long ldw, void* stream);

/* ---- K10: class-weighted softmax cross-entropy.  train.py:555-642, 726-780 ------------------ */
/* sums[0]=sum W[y]*nll, sums[1]=sum C[y], sums[2+k]=category k numerator (DEVICE doubles) */
int smer_xent_fwd(const float* logits, long long ld, const int64_t* targets, const float* W, const float* C,
                  const int* category, int ncat, float* lse, double* sums, long long rows, int V, void* stream);
/* sums[1] = sum_i C[y_i] only (the other SMER_XENT_MAX_SUMS entries are zeroed): the normaliser depends on the
 * targets alone, so data-parallel steps reduce it across ranks before the forward pass has finished */
int smer_xent_denominator(const int64_t* targets, const float* C, double* sums, long long rows, int V, void* stream);
int smer_xent_bwd(const float* logits, long long ld, const int64_t* targets, const float* W, const float* lse,
                  const double* sums, void* dlogits, int out_dtype, long long ldo, long long rows, int V,
                  int Vpad, float grad_scale, const float* grad_scale_dev /* device scalar or NULL */,
                  void* stream);

/* Token accuracy by target class -- the metric loop of train.py:988-1034 (`accuracy()`: argmax of every row,
 * pad targets skipped, one counter pair per vocab.token_class_ranges class plus the total).
 * class_of[V]: class index of a token id or -1; counts (DEVICE, caller-zeroed, +=): [0..ncls) correct per
 * class, [ncls] correct total, [ncls+1 .. 2*ncls+1) tokens per class, [2*ncls+1] tokens total.
 * argmax_out (nullable): [rows] first-maximum index of every row (torch.argmax). */
#define SMER_ACC_MAX_CLASSES 30
int smer_token_accuracy(const float* logits, long long ld, const int64_t* targets, const int* class_of, int ncls,
                        unsigned long long* counts, int64_t* argmax_out, long long rows, int V, void* stream);

/* ---- K11: Adam.  train.py:264,786 ------------------------------------------------------------ */
int smer_adam_step(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, int step,
                   float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);
/* the same update for a table of separate tensors in ONE launch (nn.Parameters of the module path):
 * table_dev is DEVICE memory, any element counts, max_n = the largest n in the table */
typedef struct smer_adam_tensor {
  float* p;
  const float* g;
  float *m, *v;
  long long n;
} smer_adam_tensor;
int smer_adam_multi(const smer_adam_tensor* table_dev, int n_tensors, long long max_n, int step, float lr, float beta1,
                    float beta2, float eps, float grad_scale, void* stream);
/* same update with the step number read from device memory (CUDA-graph replays) */
int smer_adam_step_dev(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n,
                       const uint64_t* step_dev, float lr, float beta1, float beta2, float eps, float grad_scale,
                       void* stream);

/* ---- K12: decode step.  generation.py:209-225 (model call), 41-95 + 538-686 (sampling) ------- */
typedef struct smer_decode_attn_args {
  const void* q;              /* [n_seq, H*dh] one query row per piece, pitch ldq             */
  const void *new_k, *new_v;  /* [n_seq, H*dh] K/V of the token being fed (appended) or NULL  */
  void *k_cache, *v_cache;    /* [n_seq, cache_len, H*dh]                                     */
  void* out;                  /* [n_seq, H*dh]                                                */
  const int* kv_len;          /* [n_seq] keys already cached (self) / memory length (cross)   */
  const uint8_t* key_pad;     /* [n_seq, ld_pad] or NULL                                      */
  float* workspace;           /* smer_decode_attn_workspace_bytes() when splits > 1           */
  long long ldq, ld_new, ldo, ld_cache, cache_stride, ld_pad;
  int n_seq, H, dh, cache_len, splits, dtype;
  float scale;
  const int* done;            /* [n_seq] or NULL: pieces whose flag is set are skipped (their K/V is not streamed) */
} smer_decode_attn_args;
long long smer_decode_attn_workspace_bytes(int n_seq, int H, int dh, int splits);
int smer_decode_attn(const smer_decode_attn_args* a, void* stream);
/* ids[s] = tok_buf[s, p], pos[s] = p with p = min(fed_len[s], cur_len[s]-1) (fed_len NULL: cur_len-1) */
int smer_decode_gather(const int64_t* tok_buf, const int* cur_len, const int* fed_len, int64_t* ids, int* pos,
                       int n_seq, int max_len, void* stream);
/* the two steps above in one launch: out[s] = emb[tok_buf[s, p]] * scale + pe[p], pos[s] = p, p as in smer_decode_gather;
 * rows of finished pieces (done[s] != 0, done nullable) are left untouched */
int smer_decode_embed(const int64_t* tok_buf, const int* cur_len, const int* fed_len, const int* done, int* pos,
                      const float* emb, const float* pe, void* out, int out_dtype, int n_seq, int max_len, int d, int V,
                      float scale, void* stream);
int smer_embed_step(const int64_t* ids, const int* pos, const float* emb, const float* pe, void* out,
                    int out_dtype, int n_seq, int d, int V, float scale, void* stream);

/* Small-M linear layer of the decode step (bf16 operands, M = pieces being decoded): out = epi(LN?(a) . w^T + bias),
 * w is [N, K] row-major like nn.Linear.weight (transformer.py:362-364).  Epilogues: relu, or + resid (both bf16 out), or
 * fp32 out (the vocabulary projection, model.py:106).  ln_gamma / ln_beta non-NULL: `a` holds pre-normalisation sums and
 * the rows are LayerNorm-ed on the fly (eps as nn.LayerNorm); ln_out (nullable) receives the normalised rows.
 * ln2_gamma / ln2_beta non-NULL: a second LayerNorm over the (bf16-rounded) result of the first -- the last decoder
 * layer's norm3 followed by the decoder's final norm (transformer.py:469, 286-287) ahead of the vocabulary projection.
 * Launched with programmatic stream serialization: the weight loads start while the previous kernel of the stream
 * drains (see smer_launch_pdl in csrc/common.cuh; SMER_PDL=0 disables it). */
int smer_decode_linear(const void* a, long long lda, const void* w, long long ldw, const float* bias, const void* resid,
                       long long ldr, void* out, long long ldo, int out_dtype, int M, int N, int K, int relu,
                       const float* ln_gamma, const float* ln_beta, void* ln_out, long long ld_ln_out,
                       const float* ln2_gamma, const float* ln2_beta, float eps, void* stream);

typedef struct smer_sample_args {
  const float* logits;        /* [n_seq, ld] last-position logits                             */
  long long ld;
  int n_seq, V;
  int mode;                   /* SMER_SAMPLE_*                                                */
  double temperature, top_p;  /* doubles: the reference divides float64 logits by a Python float */
  int top_k;
  uint64_t seed;              /* Philox key; counter = (seq_base+s, step, draw)               */
  long long seq_base;
  uint64_t step_base;
  /* grammar state, one entry per piece (all optional for stateless calls) */
  int* state;                 /* SMER_ST_* bits                                               */
  const int8_t* targets;      /* [n_seq, max_spans] 0='r' 1='d' 2='o' 3='p' 4='t'             */
  const uint8_t* nwd;         /* no_whole_duration per piece (generation.py:504-507)          */
  int max_spans;
  /* stateless override used by unit tests: explicit flag bits + "is_<class>" range */
  const int* raw_flags;       /* 1 no_pitch 2 no_duration 4 no_rest 8 no_whole 16 no_eos 32 no_continue 64 no_sep */
  const int *raw_only_lo, *raw_only_hi;
  /* decoder input stream bookkeeping (generation.py:673-686) */
  int64_t* tok_buf;           /* [n_seq, max_len]                                             */
  int *cur_len, *span_start, *span_idx;
  int* fed_len;               /* tokens whose K/V are cached (next position to feed), or NULL */
  const int* n_spans;
  int *done, *gen_count;
  const uint32_t* control_bitmap; /* bit i set: id i is in the caller's all_controls list      */
  int max_len, max_span;      /* max_span = 100 in the reference                              */
  /* outputs */
  int64_t* out_token;         /* [n_seq] or NULL                                              */
  double* out_probs;          /* [n_seq, V] final sampling distribution, or NULL              */
  double* trace_masked;       /* [n_seq, max_len, V] or NULL: the masked softmax (before nucleus / top-k) of every
                                 sampling step, stored at the stream position of the token just fed -- parity
                                 instrumentation for "sampled decode matches the per-step masked distributions" */
  int* trace_span;            /* [n_seq, max_len] or NULL: the span index that sampling step belonged to        */
} smer_sample_args;
int smer_sample_masked(const smer_sample_args* a, void* stream);

/* ---- helpers ------------------------------------------------------------------------------- */
int smer_cast2d(const void* src, int src_dtype, long long src_ld, void* dst, int dst_dtype, long long dst_ld,
                long long rows, int cols, int dst_cols, void* stream);
int smer_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
/* kv_len[b] = +(1+last unmasked key) when the masked keys of row b are a pure suffix, -(1+last unmasked key) otherwise */
int smer_kv_len_from_pad(const uint8_t* pad, int* kv_len, int B, int L, void* stream);
/* flags3 (device int[3]): [0] bad lower triangle, [1] some upper entry != -inf, [2] some upper entry != 0 */
int smer_classify_mask(const float* mask, long long ld, int T, int* flags3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMER_B200_H */
