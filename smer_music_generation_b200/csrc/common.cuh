// Shared device/host helpers for the SMER B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define SMER_DT_F32 0
#define SMER_DT_BF16 1

#define SMER_OK 0
#define SMER_ERR_ARG (-1)
#define SMER_ERR_CUDA (-2)
#define SMER_ERR_UNSUPPORTED (-3)

typedef __nv_bfloat16 bf16;

void smer_set_error(const char* fmt, ...);

#define SMER_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      smer_set_error(__VA_ARGS__);                \
      return SMER_ERR_ARG;                        \
    }                                             \
  } while (0)

// Programmatic dependent launch for the kernels of the decode step (a chain of ~36 short launches per token inside a CUDA
// graph): a kernel launched through smer_launch_pdl may be scheduled as soon as every CTA of its predecessor has executed
// pdl_trigger(), and must execute pdl_wait() before it reads or writes anything a predecessor touches -- until then it may
// only set itself up and prefetch data no kernel of the chain writes (weights).  pdl_wait() returns once the predecessor
// grid has completed and its writes are visible, so the chain stays strictly ordered; only launch latency and prologues
// overlap.  The attribute is only set while smer_set_pdl(1) is in effect: the small-batch decode step (every launch of
// the chain is one of these kernels) enables it around its launches.  Mixed into the large-batch step -- the same
// kernels between tcgen05 GEMM / LayerNorm launches that know nothing of it -- the 1024-piece run faulted (illegal
// address; with the attribute off on any one of embed / attention / sampler it did not), so that path stays in plain
// stream order.  SMER_PDL=0 in the environment turns the attribute off everywhere.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
int smer_pdl_flag();                 // capi.cu: set by smer_set_pdl (the small-batch decode step turns it on around its launches)
inline bool smer_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("SMER_PDL"); return !(e && e[0] == '0'); }();
  return on && smer_pdl_flag() != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t smer_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = smer_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

#define SMER_CHECK_LAUNCH(name)                                              \
  do {                                                                       \
    cudaError_t e__ = cudaPeekAtLastError();                                 \
    if (e__ != cudaSuccess) {                                                \
      cudaGetLastError();                                                    \
      smer_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return SMER_ERR_CUDA;                                                  \
    }                                                                        \
  } while (0)

#define SMER_CUDA(call)                                                          \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      smer_set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
      return SMER_ERR_CUDA;                                                      \
    }                                                                            \
  } while (0)

int smer_num_sms();
// Optional device-resident addend for every dropout seed (smer_set_seed_device_ptr): lets a
// captured CUDA graph draw fresh masks on every replay.  NULL when unset.
const unsigned long long* smer_seed_dev();
__device__ __forceinline__ uint64_t eff_seed(uint64_t seed, const unsigned long long* seed_dev) {
  return seed_dev ? seed + *seed_dev : seed;
}

// ---------------------------------------------------------------------------------------
// element conversion
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4-wide vector access (float4 for fp32, 8-byte for bf16)
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 a = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&a);
}

// Blackwell packed fp32 math (FFMA2 / FADD2 / FMUL2) and 3-input max (FMNMX3): the softmax threads of
// the attention kernels and the LayerNorm kernels are instruction-issue-bound, so every halved instruction counts.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

// bf16x2 (one 32-bit word) <-> two fp32: the low half is element 0
__device__ __forceinline__ f32x2 bf2_to_f2(uint32_t u) { return pack2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u)); }

// ---------------------------------------------------------------------------------------
// warp reductions
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (used by the sampler, where stream quality matters).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32);
  uint32_t c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// keep-mask for 4 consecutive elements starting at element index `idx4*4` of dropout site
// `site`: per-element multiplier (0 or 1/(1-p)).  Counter hash, not Philox: Philox4x32-10 is
// ~100 instructions per 4 elements, which made the GEMM/LayerNorm epilogues issue-bound.  One
// two-round xorshift-multiply mix of (key + index) gives the first element's 32-bit uniform, the
// other three are multiply-add steps of it; keep iff word >= thr, thr = p * 2^32.
__device__ __forceinline__ uint32_t lowbias32_(uint32_t x) {
  x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
  return x;
}
__device__ __forceinline__ uint32_t dropout_key(uint64_t seed, uint64_t site) {
  return lowbias32_((uint32_t)seed ^ ((uint32_t)site * 0x9E3779B9u)) + (uint32_t)(seed >> 32);
}
__device__ __forceinline__ void dropout4k(uint32_t key, uint64_t idx4, uint32_t thr, float inv_keep, float (&m)[4]) {
  uint32_t x = key + (uint32_t)idx4 * 0x9E3779B9u + (uint32_t)(idx4 >> 32) * 0x85EBCA6Bu;
  x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u;
  m[0] = x >= thr ? inv_keep : 0.f;
  x = x * 0x297A2D39u + 0x7F4A7C15u;
  m[1] = x >= thr ? inv_keep : 0.f;
  x = x * 0x297A2D39u + 0x7F4A7C15u;
  m[2] = x >= thr ? inv_keep : 0.f;
  x = x * 0x297A2D39u + 0x7F4A7C15u;
  m[3] = x >= thr ? inv_keep : 0.f;
}
__device__ __forceinline__ void dropout4(uint64_t seed, uint64_t site, uint64_t idx4, uint32_t thr,
                                         float inv_keep, float (&m)[4]) {
  dropout4k(dropout_key(seed, site), idx4, thr, inv_keep, m);
}

// v[c] (c = 0..31) in every lane -> the sum over the 32 lanes of column `lane` (recursive halving:
// 31 shuffles; lane bit `off` decides which half of the remaining columns a lane keeps)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ---------------------------------------------------------------------------------------
// Attention-probability dropout.  Philox costs ~100 instructions per 4 elements, which would
// make the softmax threads of the tcgen05 attention kernels the bottleneck several times over
// (they are instruction-issue bound); the (B,H,Lq,Lk) keep-mask is a counter hash instead:
//   site key  = two avalanche rounds of (seed, site)                          -- once per launch
//   row key   = one avalanche round of (site key, row id (b*H+h)*Lq + i)      -- once per row
//   block     = one xorshift-multiply-xorshift round of (row key + (j >> 4) * golden ratio): a
//               32-bit word w per 16 consecutive keys of the row
//   pair word = x_k = w advanced by k multiply-add (LCG) steps, k = (j & 15) >> 1; key j uses
//               the low half of x_k when it is even, the high half when it is odd
//   keep      iff the 16-bit field f, READ AS AN fp16 BIT PATTERN, is >= the (negative) fp16 whose
//               pattern is thr = 0xFC00 - round(p * 65536), unordered compare: as integers,
//               keep <=> f <= thr or f >= 0xFC01.  A drop probability of round(p*65536)/65536.
// The point of that last rule: the tcgen05 kernels decide TWO keys with ONE packed half compare
// on the pair word (HSET2 gives the 0xFFFF / 0 masks the bf16x2 probabilities are ANDed with,
// HSETP2 gives two predicates), without extracting or masking fields -- 1 multiply-add + 1
// compare per two keys plus the block round per 16.  The SIMT kernels evaluate the same rule with
// integer compares, so every kernel (forward, backward, attention weights) agrees bit for bit.
// (Offline check of the scheme at p = 0.1: drop rate 0.10009, per-position rates 0.0988..0.1013,
// pair correlations inside a block <= 0.006, key / row lag correlations ~1e-3, 16-key and 32x32
// block-sum variances within 1 % / 2 % of the binomial values.)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
  return x;
}
constexpr uint32_t ATTN_GOLD = 0x9E3779B9u;
constexpr uint32_t ATTN_A = 0x297A2D39u, ATTN_C = 0x7F4A7C15u;
constexpr uint32_t lcg_mul_n(int n) { uint32_t m = 1u; for (int i = 0; i < n; ++i) m *= ATTN_A; return m; }
constexpr uint32_t lcg_add_n(int n) { uint32_t a = 0u; for (int i = 0; i < n; ++i) a = a * ATTN_A + ATTN_C; return a; }
__device__ __forceinline__ uint32_t attn_site_key(uint64_t seed, uint64_t site) {
  uint32_t a = lowbias32((uint32_t)seed ^ ((uint32_t)site * 0x9E3779B9u) ^ 0x85EBCA6Bu);
  return lowbias32(a + (uint32_t)(seed >> 32));
}
__device__ __forceinline__ uint32_t attn_row_key(uint32_t sitekey, long long rowid) {
  return lowbias32(sitekey ^ ((uint32_t)rowid * 0x9E3779B1u) ^ ((uint32_t)((unsigned long long)rowid >> 32) * 0x85EBCA77u));
}
// the 32-bit word of the 16-key block `jb` = j >> 4 of a row
__device__ __forceinline__ uint32_t attn_block_word(uint32_t rowkey, uint32_t jb) {
  uint32_t x = rowkey + jb * ATTN_GOLD;
  x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12;
  return x;
}
// x[k] = w advanced by k LCG steps, k = 0..7: the pair words of the block's 16 keys, as a depth-3 tree of
// multiply-adds with only three distinct multipliers (registers are what the softmax threads are short of)
__device__ __forceinline__ void attn_pair_words(uint32_t w, uint32_t (&x)[8]) {
  constexpr uint32_t M1 = lcg_mul_n(1), A1 = lcg_add_n(1), M2 = lcg_mul_n(2), A2 = lcg_add_n(2), M4 = lcg_mul_n(4), A4 = lcg_add_n(4);
  x[0] = w;
  x[1] = w * M1 + A1;
  x[2] = w * M2 + A2;
  x[4] = w * M4 + A4;
  x[3] = x[2] * M1 + A1;
  x[5] = x[4] * M1 + A1;
  x[6] = x[4] * M2 + A2;
  x[7] = x[6] * M1 + A1;
}
// packed rule: 0xFFFF in every half of `pairword` that is kept (thr2 = thr | thr << 16)
__device__ __forceinline__ uint32_t attn_keep_mask2(uint32_t pairword, uint32_t thr2) {
  uint32_t m;
  asm("set.geu.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(pairword), "r"(thr2));
  return m;
}
__device__ __forceinline__ void attn_keep_pred2(uint32_t pairword, uint32_t thr2, bool& k0, bool& k1) {
  uint32_t a, b;
  asm("{\n\t.reg .pred p, q;\n\tsetp.geu.f16x2 p|q, %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\tselp.u32 %1, 1, 0, q;\n\t}"
      : "=r"(a), "=r"(b) : "r"(pairword), "r"(thr2));
  k0 = a != 0u; k1 = b != 0u;
}
// integer form of the same rule for one 16-bit field
__device__ __forceinline__ bool attn_keep_field(uint32_t f, uint32_t thr) { return f <= thr || f >= 0xFC01u; }
// generic (slow) form: rowkey = attn_row_key(attn_site_key(seed, site), rowid)
__device__ __forceinline__ bool attn_keep(uint32_t rowkey, int j, uint32_t thr) {
  uint32_t w = attn_block_word(rowkey, (uint32_t)j >> 4);
  uint32_t mul = 1u, add = 0u;
  for (int k = 0; k < ((j & 15) >> 1); ++k) { mul *= ATTN_A; add = add * ATTN_A + ATTN_C; }
  const uint32_t x = w * mul + add;
  return attn_keep_field((j & 1) ? x >> 16 : x & 0xFFFFu, thr);
}
// threshold pattern of the rule above; valid for p <= 0.45 (the pattern must stay a negative finite fp16)
static inline uint32_t attn_dropout_threshold(float p) {
  double n = (double)p * 65536.0 + 0.5;
  if (n < 0) n = 0;
  if (n > 29491.0) n = 29491.0;
  return 0xFC00u - (uint32_t)n;
}

static inline uint32_t dropout_threshold(float p) {
  double t = (double)p * 4294967296.0;
  if (t < 0) t = 0;
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}
