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
// make the softmax warps of the tcgen05 attention kernels the bottleneck several times over;
// the (B,H,Lq,Lk) keep-mask is instead a counter hash: a 32-bit key per (seed, site, b, h, i)
// row, then one avalanche mix per PAIR of keys giving two 16-bit uniforms.  The SIMT and
// tcgen05 kernels (forward, backward, attention-weights) all call these, so they agree bit
// for bit.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
  return x;
}
__device__ __forceinline__ uint32_t attn_row_key(uint64_t seed, uint64_t site, long long rowid) {
  uint32_t a = lowbias32((uint32_t)seed ^ ((uint32_t)site * 0x9E3779B9u) ^ 0x85EBCA6Bu);
  a = lowbias32(a + (uint32_t)(seed >> 32));
  a = lowbias32(a ^ (uint32_t)rowid);
  a = lowbias32(a + (uint32_t)((unsigned long long)rowid >> 32) + 0x632BE5ABu);
  return a;
}
// One 32-bit word per 8x8 BLOCK of (8 query rows, 8 keys): two xorshift-multiply rounds of
// (block-row key + key-block index * golden ratio).  Element (i & 7, j & 7) of the block uses
// that word advanced by 8*(i&7) + (j&7) multiply-add (LCG) steps; keep iff word >= thr32 =
// p * 2^32 (full word compare, no field extraction).  Kernels that walk along keys (forward,
// dQ) pay the mix once per 8 keys and one multiply-add per further key; the dK/dV kernel (one
// thread per key, walking along queries) pays it once per 8 queries and steps by 8 at a time --
// instruction issue on the ALU pipe is what bounds the tcgen05 attention kernels, and the
// multiply-adds run on the FMA pipe.  The row key is taken for the FIRST row of the block:
// attn_row_key(seed, site, rowbase + (i & ~7)).  (Offline check of the scheme: keep rate,
// lag correlations up to 16x16 and 8x8 block-sum variance all at the binomial values.)
constexpr int ATTN_BLK = 8;
constexpr uint32_t ATTN_GOLD = 0x9E3779B9u;
constexpr uint32_t ATTN_A = 0x297A2D39u, ATTN_C = 0x7F4A7C15u;
constexpr uint32_t lcg_mul_n(int n) { uint32_t m = 1u; for (int i = 0; i < n; ++i) m *= ATTN_A; return m; }
constexpr uint32_t lcg_add_n(int n) { uint32_t a = 0u; for (int i = 0; i < n; ++i) a = a * ATTN_A + ATTN_C; return a; }
constexpr uint32_t ATTN_A4 = lcg_mul_n(4), ATTN_C4 = lcg_add_n(4);                  // four steps at once
constexpr uint32_t ATTN_A8 = lcg_mul_n(8), ATTN_C8 = lcg_add_n(8);                  // eight steps: next row of a block
__device__ __forceinline__ uint32_t attn_mix(uint32_t x) {
  x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u;
  return x;
}
__device__ __forceinline__ uint32_t attn_blk_x(uint32_t blockkey, int j) {
  return attn_mix(blockkey + (uint32_t)(j >> 3) * ATTN_GOLD);
}
__device__ __forceinline__ uint32_t attn_step(uint32_t x) { return x * ATTN_A + ATTN_C; }
__device__ __forceinline__ uint32_t attn_step4(uint32_t x) { return x * ATTN_A4 + ATTN_C4; }
__device__ __forceinline__ uint32_t attn_step8(uint32_t x) { return x * ATTN_A8 + ATTN_C8; }
// x[k] = x0 advanced by k*STEP steps, k = 0..7, as a depth-3 tree of multiply-adds (7 IMADs like the serial
// chain, but a dependency depth of 3 instead of 7)
template <int STEP>
__device__ __forceinline__ void attn_block8(uint32_t x0, uint32_t (&x)[8]) {
  constexpr uint32_t M1 = lcg_mul_n(STEP), A1 = lcg_add_n(STEP), M2 = lcg_mul_n(2 * STEP), A2 = lcg_add_n(2 * STEP),
                     M4 = lcg_mul_n(4 * STEP), A4 = lcg_add_n(4 * STEP);
  x[0] = x0;
  x[1] = x0 * M1 + A1;
  x[2] = x0 * M2 + A2;
  x[4] = x0 * M4 + A4;
  x[3] = x[2] * M1 + A1;
  x[5] = x[4] * M1 + A1;
  x[6] = x[4] * M2 + A2;
  x[7] = x[6] * M1 + A1;
}
// n LCG steps as one multiply-add: x -> x * mul + add
__device__ __forceinline__ void attn_advance(int n, uint32_t& mul, uint32_t& add) {
  mul = 1u; add = 0u;
  for (int k = 0; k < n; ++k) { mul *= ATTN_A; add = add * ATTN_A + ATTN_C; }
}
// generic (slow) form: blockkey = attn_row_key(seed, site, rowbase + (i & ~7))
__device__ __forceinline__ bool attn_keep(uint32_t blockkey, int i, int j, uint32_t thr32) {
  uint32_t mul, add;
  attn_advance(8 * (i & 7) + (j & 7), mul, add);
  return attn_blk_x(blockkey, j) * mul + add >= thr32;
}

static inline uint32_t dropout_threshold(float p) {
  double t = (double)p * 4294967296.0;
  if (t < 0) t = 0;
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}
