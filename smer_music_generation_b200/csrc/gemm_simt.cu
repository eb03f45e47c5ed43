// FP32-accumulate CUDA-core GEMM with arbitrary operand strides:
//     C[m,n] = epilogue( sum_k A(m,k) * B(n,k) ),  A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk]
// This is the arithmetic of nn.Linear (transformer.py:362-364, model.py:82) and of its two
// backward products.  It is the fp32 parity path (fp32 rel 1e-4 needs true FP32 FMAs: TF32
// tensor-core products carry a 10-bit mantissa) and serves shapes that the tcgen05 kernel
// does not tile (tiny test models, single-row decode).  The bf16 hot path is gemm_tc.cu.
#include "common.cuh"
#include "../../include/smer_b200.h"

constexpr int TM = 64, TN = 64, TK = 16;

template <typename TA, typename TC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, long long sam, long long sak, const TA* __restrict__ B,
                 long long sbn, long long sbk, TC* __restrict__ C, long long ldc, int M, int N, int K,
                 const float* __restrict__ bias, const TC* __restrict__ resid, long long ldr, int flags,
                 uint32_t thr, float inv_keep, uint64_t seed, uint64_t site, int ksplit_len,
                 const unsigned long long* seed_dev) {
  seed = eff_seed(seed, seed_dev);
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  int tid = threadIdx.x;
  int tx = tid & 15, ty = tid >> 4;
  int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  int kbeg = blockIdx.z * ksplit_len;
  int kend = min(K, kbeg + ksplit_len);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  bool a_kfast = (sak == 1), b_kfast = (sbk == 1);
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int mm, kk;
      if (a_kfast) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < kend) ? to_f32(A[gm * sam + gk * sak]) : 0.f;
      int nn;
      if (b_kfast) { kk = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kk = idx >> 6; }
      int gn = n0 + nn;
      gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < kend) ? to_f32(B[gn * sbn + gk * sbk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (flags & SMER_EPI_ATOMIC) {           // split-K partial (fp32 output only)
        if (bias && blockIdx.z == 0) v += bias[gn];
        atomicAdd(reinterpret_cast<float*>(C) + (long long)gm * ldc + gn, v);
        continue;
      }
      if (bias) v += bias[gn];
      if (flags & SMER_EPI_RELU) v = fmaxf(v, 0.f);
      if (flags & SMER_EPI_GATE) {             // backward of dropout(relu(.)): gate = stored activation
        v = to_f32(resid[(long long)gm * ldr + gn]) > 0.f ? v * inv_keep : 0.f;
        C[(long long)gm * ldc + gn] = from_f32<TC>(v);
        continue;
      }
      if (thr) {
        // element-indexed keep mask: lane (e & 3) of dropout4 at counter (row*ldc + col)/4
        long long e = (long long)gm * ldc + gn;
        float m4[4];
        dropout4(seed, site, (uint64_t)(e >> 2), thr, inv_keep, m4);
        v *= m4[e & 3];
      }
      if (resid) v += to_f32(resid[(long long)gm * ldr + gn]);
      if (flags & SMER_EPI_ACCUM) v += to_f32(C[(long long)gm * ldc + gn]);
      C[(long long)gm * ldc + gn] = from_f32<TC>(v);
    }
  }
}

extern "C" int smer_gemm_simt(const void* A, long long sam, long long sak, const void* B, long long sbn,
                              long long sbk, void* C, long long ldc, int in_dtype, int out_dtype, int M, int N,
                              int K, const float* bias, const void* resid, long long ldr, int flags,
                              float dropout_p, uint64_t seed, uint64_t site, int split_k, void* stream) {
  SMER_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "smer_gemm_simt: negative dims");
  if (M == 0 || N == 0) return SMER_OK;
  if (split_k < 1) split_k = 1;
  SMER_CHECK_ARG(split_k == 1 || ((flags & SMER_EPI_ATOMIC) && out_dtype == SMER_DT_F32),
                 "smer_gemm_simt: split_k needs SMER_EPI_ATOMIC and fp32 output");
  SMER_CHECK_ARG(!(flags & SMER_EPI_ATOMIC) || out_dtype == SMER_DT_F32, "smer_gemm_simt: atomic epilogue needs fp32 output");
  uint32_t thr = (dropout_p > 0.f && !(flags & SMER_EPI_GATE)) ? dropout_threshold(dropout_p) : 0u;
  float inv_keep = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  int klen = (K + split_k - 1) / split_k;
  klen = (klen + TK - 1) / TK * TK;
  if (klen == 0) klen = TK;
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, split_k);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(TA, TC)                                                                                              \
  gemm_simt_kernel<TA, TC><<<grid, 256, 0, st>>>((const TA*)A, sam, sak, (const TA*)B, sbn, sbk, (TC*)C, ldc, M, N, K, \
                                                 bias, (const TC*)resid, ldr, flags, thr, inv_keep, seed, site, klen, smer_seed_dev())
  if (in_dtype == SMER_DT_F32 && out_dtype == SMER_DT_F32) LAUNCH(float, float);
  else if (in_dtype == SMER_DT_BF16 && out_dtype == SMER_DT_BF16) LAUNCH(bf16, bf16);
  else if (in_dtype == SMER_DT_BF16 && out_dtype == SMER_DT_F32) LAUNCH(bf16, float);
  else {
    smer_set_error("smer_gemm_simt: unsupported dtype combination %d -> %d", in_dtype, out_dtype);
    return SMER_ERR_UNSUPPORTED;
  }
#undef LAUNCH
  SMER_CHECK_LAUNCH("smer_gemm_simt");
  return SMER_OK;
}
