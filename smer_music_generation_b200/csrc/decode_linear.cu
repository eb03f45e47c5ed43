// Small-M linear layers of the decode step: out = epi(LN?(a) . W^T + b) for M <= a few hundred rows (one row per
// piece being decoded, generation.py:209-225 at batch n).  The tcgen05 GEMM tiles 128 x 128/256 outputs, so a
// [128, 512] x [512, 512] product runs on 4 of the 148 SMs; here a CTA owns a 32-row x 8-column slab, which spreads even
// the smallest projection over >= 256 CTAs, and the weights (29.7 MB for all of a step's layers) stream out of L2 once
// per 32 rows.  The products are latency / L2-bound, far below the tensor-core roofline, so they use the warp-level
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32) with operand fragments loaded straight from global memory: every thread
// fetches 8 consecutive k of its row (one 16-byte load), which feeds two k16 MMAs -- A and B use the same placement of
// k inside a fragment, and a dot product does not care in which order k is summed.
//
// Optional fused LayerNorm prologue (transformer.py:391-395, 461-469: x = LN(resid + branch)): `a` then holds the
// pre-normalisation sums z (written by the previous product's residual epilogue); each warp first reduces mean / rstd of
// its 16 rows, normalises the fragments on the fly, and the CTAs of the first column slab also store y = LN(z), the
// residual input of the next sub-layer.  That removes every stand-alone LayerNorm launch from the decode step.
#include "common.cuh"
#include "../../include/smer_b200.h"

namespace {

constexpr int DL_BM = 32, DL_BN = 16, DL_KG = 4, DL_THREADS = 256;    // 2 row groups x 4 k-groups of warps; 2 n8 tiles per warp
constexpr int DL_CHUNK = 8;                                            // 32-k steps whose loads are in flight together

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct Args {
  const bf16* a; long long lda;
  const bf16* w; long long ldw;
  const float* bias;
  const bf16* resid; long long ldr;
  void* out; long long ldo;
  const float *gamma, *beta;      // LN prologue
  const float *gamma2, *beta2;    // a second LayerNorm applied to the result of the first (last layer's LN3, then the decoder's final norm)
  bf16* y; long long ldy;         // LN prologue: normalised rows (nullable)
  int M, N, K;
  float eps;
};

__device__ __forceinline__ void sum_sq8(uint4 v, float& s, float& q) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float x0 = __uint_as_float(u[i] << 16), x1 = __uint_as_float(u[i] & 0xFFFF0000u);
    s += x0 + x1;
    q = fmaf(x0, x0, fmaf(x1, x1, q));
  }
}
// normalise 8 consecutive k of one row held as 4 bf16x2 words
__device__ __forceinline__ uint4 ln_apply(uint4 v, float mean, float rstd, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, int k0) {
  const float4 g0 = *reinterpret_cast<const float4*>(gamma + k0), g1 = *reinterpret_cast<const float4*>(gamma + k0 + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(beta + k0), b1 = *reinterpret_cast<const float4*>(beta + k0 + 4);
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
  const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float x0 = __uint_as_float(u[i] << 16), x1 = __uint_as_float(u[i] & 0xFFFF0000u);
    o[i] = pack_bf16x2((x0 - mean) * rstd * g[2 * i] + b[2 * i], (x1 - mean) * rstd * g[2 * i + 1] + b[2 * i + 1]);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// CTA = 32 rows x 8 columns; warp = (row group of 16 rows, quarter of K).  The few products of a decode step are bound by
// the latency of their dependent L2 loads, so a warp first issues ALL the 16-byte loads of (up to) 8 k-steps, then runs the
// MMAs; the four K quarters are summed through shared memory.
template <bool LN, bool RELU, bool RESID, bool OUT_F32>
__global__ void __launch_bounds__(DL_THREADS) decode_linear_kernel(Args p) {
  __shared__ float red[DL_KG][DL_BM][DL_BN];
  __shared__ float stat[DL_KG][DL_BM][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = warp & 1, kg = warp >> 1;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * DL_BN;
  const int r0 = blockIdx.y * DL_BM + rg * 16;              // this warp's 16 rows
  const int ra = min(r0 + g, p.M - 1), rb = min(r0 + g + 8, p.M - 1);   // (clamped: out-of-range rows are computed, not stored)
  const int kq = p.K / DL_KG, k_lo = kg * kq, nsteps = kq / 32;
  const bf16* arow = p.a + (long long)ra * p.lda + k_lo + 8 * t;
  const bf16* brow = p.a + (long long)rb * p.lda + k_lo + 8 * t;
  const int wn = min(n0 + g, p.N - 1), wn2 = min(n0 + 8 + g, p.N - 1);
  const bf16* wrow = p.w + (long long)wn * p.ldw + k_lo + 8 * t;
  const bf16* wrow2 = p.w + (long long)wn2 * p.ldw + k_lo + 8 * t;
  const bool write_y = LN && p.y != nullptr && blockIdx.x == 0;

  float c[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
  pdl_trigger();
  for (int s0 = 0; s0 < nsteps; s0 += DL_CHUNK) {
    uint4 va[DL_CHUNK], vb[DL_CHUNK], vw[DL_CHUNK], vx[DL_CHUNK];
#pragma unroll
    for (int u = 0; u < DL_CHUNK; ++u) {
      if (s0 + u < nsteps) {
        vw[u] = *reinterpret_cast<const uint4*>(wrow + (s0 + u) * 32);
        vx[u] = *reinterpret_cast<const uint4*>(wrow2 + (s0 + u) * 32);
      }
    }
    // the weights above are written by no kernel of the decode chain, so their loads are already in flight while the
    // previous launch drains; the activations are its output
    if (s0 == 0) pdl_wait();
#pragma unroll
    for (int u = 0; u < DL_CHUNK; ++u) {
      if (s0 + u < nsteps) {
        va[u] = *reinterpret_cast<const uint4*>(arow + (s0 + u) * 32);
        vb[u] = *reinterpret_cast<const uint4*>(brow + (s0 + u) * 32);
      }
    }
    // (the launcher guarantees nsteps <= DL_CHUNK with a LayerNorm prologue: the whole K quarter is in registers)
    for (int pass = 0; LN && pass < (p.gamma2 ? 2 : 1); ++pass) {
      const float* gamma = pass ? p.gamma2 : p.gamma;
      const float* beta = pass ? p.beta2 : p.beta;
      if (pass) __syncthreads();                      // every warp has read the first pass's partial sums
      float sa = 0.f, qa = 0.f, sb = 0.f, qb = 0.f;
#pragma unroll
      for (int u = 0; u < DL_CHUNK; ++u)
        if (u < nsteps) { sum_sq8(va[u], sa, qa); sum_sq8(vb[u], sb, qb); }
      sa += __shfl_xor_sync(0xffffffffu, sa, 1); sa += __shfl_xor_sync(0xffffffffu, sa, 2);
      qa += __shfl_xor_sync(0xffffffffu, qa, 1); qa += __shfl_xor_sync(0xffffffffu, qa, 2);
      sb += __shfl_xor_sync(0xffffffffu, sb, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
      qb += __shfl_xor_sync(0xffffffffu, qb, 1); qb += __shfl_xor_sync(0xffffffffu, qb, 2);
      if (t == 0) {
        stat[kg][rg * 16 + g][0] = sa; stat[kg][rg * 16 + g][1] = qa;
        stat[kg][rg * 16 + g + 8][0] = sb; stat[kg][rg * 16 + g + 8][1] = qb;
      }
      __syncthreads();
      float ta = 0.f, ua = 0.f, tb = 0.f, ub = 0.f;
#pragma unroll
      for (int q = 0; q < DL_KG; ++q) {
        ta += stat[q][rg * 16 + g][0]; ua += stat[q][rg * 16 + g][1];
        tb += stat[q][rg * 16 + g + 8][0]; ub += stat[q][rg * 16 + g + 8][1];
      }
      const float inv_k = 1.f / p.K;
      const float mean_a = ta * inv_k, mean_b = tb * inv_k;
      const float rstd_a = rsqrtf(fmaxf(ua * inv_k - mean_a * mean_a, 0.f) + p.eps);
      const float rstd_b = rsqrtf(fmaxf(ub * inv_k - mean_b * mean_b, 0.f) + p.eps);
#pragma unroll
      for (int u = 0; u < DL_CHUNK; ++u) {
        if (u < nsteps) {
          const int k = k_lo + u * 32 + 8 * t;
          va[u] = ln_apply(va[u], mean_a, rstd_a, gamma, beta, k);
          vb[u] = ln_apply(vb[u], mean_b, rstd_b, gamma, beta, k);
          if (write_y && pass == 0) {
            if (r0 + g < p.M) *reinterpret_cast<uint4*>(p.y + (long long)(r0 + g) * p.ldy + k) = va[u];
            if (r0 + g + 8 < p.M) *reinterpret_cast<uint4*>(p.y + (long long)(r0 + g + 8) * p.ldy + k) = vb[u];
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < DL_CHUNK; ++u) {
      if (s0 + u < nsteps) {
        mma_bf16_16816(c, va[u].x, vb[u].x, va[u].y, vb[u].y, vw[u].x, vw[u].y);
        mma_bf16_16816(c, va[u].z, vb[u].z, va[u].w, vb[u].w, vw[u].z, vw[u].w);
        mma_bf16_16816(c2, va[u].x, vb[u].x, va[u].y, vb[u].y, vx[u].x, vx[u].y);
        mma_bf16_16816(c2, va[u].z, vb[u].z, va[u].w, vb[u].w, vx[u].z, vx[u].w);
      }
    }
  }
  // sum of the four K quarters; c0,c1 = (row g, cols 2t, 2t+1), c2,c3 = (row g+8, same cols)
  red[kg][rg * 16 + g][2 * t] = c[0]; red[kg][rg * 16 + g][2 * t + 1] = c[1];
  red[kg][rg * 16 + g + 8][2 * t] = c[2]; red[kg][rg * 16 + g + 8][2 * t + 1] = c[3];
  red[kg][rg * 16 + g][8 + 2 * t] = c2[0]; red[kg][rg * 16 + g][8 + 2 * t + 1] = c2[1];
  red[kg][rg * 16 + g + 8][8 + 2 * t] = c2[2]; red[kg][rg * 16 + g + 8][8 + 2 * t + 1] = c2[3];
  __syncthreads();
  // epilogue: the 32 x 16 outputs as 256 column pairs, one per thread
  const int lr = threadIdx.x >> 3, cp = (threadIdx.x & 7) * 2;
  const int col = n0 + cp, row = blockIdx.y * DL_BM + lr;
  if (col >= p.N || row >= p.M) return;
  const float b0 = p.bias ? p.bias[col] : 0.f, b1 = (p.bias && col + 1 < p.N) ? p.bias[col + 1] : 0.f;
  {
    float v0 = b0, v1 = b1;
#pragma unroll
    for (int q = 0; q < DL_KG; ++q) { v0 += red[q][lr][cp]; v1 += red[q][lr][cp + 1]; }
    if (RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
    if (RESID) {
      const uint32_t rr = *reinterpret_cast<const uint32_t*>(p.resid + (long long)row * p.ldr + col);
      v0 += __uint_as_float(rr << 16);
      v1 += __uint_as_float(rr & 0xFFFF0000u);
    }
    if (OUT_F32) {
      float* o = (float*)p.out + (long long)row * p.ldo + col;
      o[0] = v0;
      if (col + 1 < p.N) o[1] = v1;
    } else {
      *reinterpret_cast<uint32_t*>((bf16*)p.out + (long long)row * p.ldo + col) = pack_bf16x2(v0, v1);
    }
  }
}

}  // namespace

extern "C" int smer_decode_linear(const void* a, long long lda, const void* w, long long ldw, const float* bias,
                                  const void* resid, long long ldr, void* out, long long ldo, int out_dtype, int M, int N,
                                  int K, int relu, const float* ln_gamma, const float* ln_beta, void* ln_out,
                                  long long ld_ln_out, const float* ln2_gamma, const float* ln2_beta, float eps,
                                  void* stream) {
  SMER_CHECK_ARG(a && w && out && M > 0 && N > 0 && K > 0, "smer_decode_linear: null args");
  SMER_CHECK_ARG(K % 128 == 0 && lda % 8 == 0 && ldw % 8 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                 "smer_decode_linear: K must be a multiple of 128 and the operand rows 16-byte aligned");
  SMER_CHECK_ARG(!ln_gamma || K <= 128 * DL_CHUNK, "smer_decode_linear: the LayerNorm prologue keeps a row quarter in registers (K <= 1024)");
  SMER_CHECK_ARG(N % 2 == 0 && ldo % 2 == 0 && (!resid || ldr % 2 == 0), "smer_decode_linear: N and the output pitch must be even");
  SMER_CHECK_ARG(!(relu && resid), "smer_decode_linear: ReLU and residual epilogues are exclusive");
  SMER_CHECK_ARG((ln_gamma == nullptr) == (ln_beta == nullptr), "smer_decode_linear: gamma and beta come together");
  SMER_CHECK_ARG((ln2_gamma == nullptr) == (ln2_beta == nullptr) && (!ln2_gamma || ln_gamma),
                 "smer_decode_linear: the second LayerNorm needs both its parameters and a first LayerNorm");
  SMER_CHECK_ARG(!ln_out || (ld_ln_out % 8 == 0 && (reinterpret_cast<uintptr_t>(ln_out) & 15) == 0), "smer_decode_linear: ln_out rows must be 16-byte aligned");
  Args p;
  p.a = (const bf16*)a; p.lda = lda; p.w = (const bf16*)w; p.ldw = ldw; p.bias = bias;
  p.resid = (const bf16*)resid; p.ldr = ldr; p.out = out; p.ldo = ldo;
  p.gamma = ln_gamma; p.beta = ln_beta; p.gamma2 = ln2_gamma; p.beta2 = ln2_beta; p.y = (bf16*)ln_out; p.ldy = ld_ln_out;
  p.M = M; p.N = N; p.K = K; p.eps = eps;
  dim3 grid((N + DL_BN - 1) / DL_BN, (M + DL_BM - 1) / DL_BM);
  cudaStream_t st = (cudaStream_t)stream;
  const bool ln = ln_gamma != nullptr, f32 = out_dtype == SMER_DT_F32, rs = resid != nullptr;
#define DL_LAUNCH(LN_, RELU_, RESID_, F32_) smer_launch_pdl(decode_linear_kernel<LN_, RELU_, RESID_, F32_>, grid, dim3(DL_THREADS), 0, st, p)
  if (ln) {
    if (f32) DL_LAUNCH(true, false, false, true);
    else if (relu) DL_LAUNCH(true, true, false, false);
    else if (rs) DL_LAUNCH(true, false, true, false);
    else DL_LAUNCH(true, false, false, false);
  } else {
    if (f32) DL_LAUNCH(false, false, false, true);
    else if (relu) DL_LAUNCH(false, true, false, false);
    else if (rs) DL_LAUNCH(false, false, true, false);
    else DL_LAUNCH(false, false, false, false);
  }
#undef DL_LAUNCH
  SMER_CHECK_LAUNCH("smer_decode_linear");
  return SMER_OK;
}
