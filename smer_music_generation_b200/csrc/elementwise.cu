// Bandwidth-bound element-wise kernels: embedding*sqrt(d)+positional (+dropout) forward and
// its scatter-add backward, dtype casts, bias-gradient column sums, mask inspection, Adam.
// Reference arithmetic: model.py:91-92,110-125 (embed/PE), train.py:264,786 (Adam).
#include "common.cuh"
#include "../../include/smer_b200.h"

// ---------------------------------------------------------------------------------------
// K1: out[b*L+l, :] = dropout(emb[ids[b,l], :] * scale + pe[l, :])
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void embed_pe_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ emb,
                                    const float* __restrict__ pe, T* __restrict__ out, long long n4,
                                    int L, int d4, int V, int pos0, float scale, uint32_t thr, float inv_keep,
                                    uint64_t seed, uint64_t site, const unsigned long long* seed_dev,
                                    const int* __restrict__ pos = nullptr) {
  seed = eff_seed(seed, seed_dev);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    long long row = i / d4;
    int c4 = (int)(i - row * d4);
    int l = pos ? pos[row] : (int)(row % L) + pos0;
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    float e[4], p[4], o[4];
    load4(emb + id * (long long)d4 * 4 + c4 * 4, e);
    load4(pe + (long long)l * d4 * 4 + c4 * 4, p);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = e[k] * scale + p[k];
    if (thr) {
      float m[4];
      dropout4(seed, site, (uint64_t)i, thr, inv_keep, m);
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] *= m[k];
    }
    store4(out + i * 4, o);
  }
}

// demb[ids[row], :] += dout[row, :] * scale * dropmask   (fp32 vector reductions in L2)
template <typename T>
__global__ void embed_bwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ dout,
                                 float* __restrict__ demb, long long n4, int d4, int V, float scale,
                                 uint32_t thr, float inv_keep, uint64_t seed, uint64_t site,
                                 const unsigned long long* seed_dev) {
  seed = eff_seed(seed, seed_dev);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    long long row = i / d4;
    int c4 = (int)(i - row * d4);
    long long id = ids[row];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    float g[4];
    load4(dout + i * 4, g);
    if (thr) {
      float m[4];
      dropout4(seed, site, (uint64_t)i, thr, inv_keep, m);
#pragma unroll
      for (int k = 0; k < 4; ++k) g[k] *= m[k];
    }
    float* dst = demb + id * (long long)d4 * 4 + c4 * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g[0] * scale),
                 "f"(g[1] * scale), "f"(g[2] * scale), "f"(g[3] * scale)
                 : "memory");
  }
}

static inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  long long cap = (long long)smer_num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" int smer_embed_pe_fwd(const int64_t* ids, const float* emb, const float* pe, void* out,
                                 int out_dtype, int B, int L, int d, int V, int pos0, float scale,
                                 float dropout_p, uint64_t seed, uint64_t site, void* stream) {
  SMER_CHECK_ARG(d % 4 == 0 && B > 0 && L > 0, "smer_embed_pe_fwd: d must be a multiple of 4");
  long long n4 = (long long)B * L * (d / 4);
  uint32_t thr = dropout_p > 0.f ? dropout_threshold(dropout_p) : 0u;
  float inv_keep = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(n4, 256);
  if (out_dtype == SMER_DT_F32)
    embed_pe_fwd_kernel<float><<<grid, 256, 0, st>>>(ids, emb, pe, (float*)out, n4, L, d / 4, V, pos0, scale, thr, inv_keep, seed, site, smer_seed_dev());
  else
    embed_pe_fwd_kernel<bf16><<<grid, 256, 0, st>>>(ids, emb, pe, (bf16*)out, n4, L, d / 4, V, pos0, scale, thr, inv_keep, seed, site, smer_seed_dev());
  SMER_CHECK_LAUNCH("smer_embed_pe_fwd");
  return SMER_OK;
}

extern "C" int smer_embed_pe_packed(const int64_t* ids, const int* pos, const float* emb, const float* pe, void* out,
                                    int out_dtype, long long rows, int d, int V, float scale, float dropout_p,
                                    uint64_t seed, uint64_t site, void* stream) {
  SMER_CHECK_ARG(d % 4 == 0 && rows > 0 && pos, "smer_embed_pe_packed: d must be a multiple of 4, pos required");
  long long n4 = rows * (d / 4);
  uint32_t thr = dropout_p > 0.f ? dropout_threshold(dropout_p) : 0u;
  float inv_keep = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(n4, 256);
  if (out_dtype == SMER_DT_F32)
    embed_pe_fwd_kernel<float><<<grid, 256, 0, st>>>(ids, emb, pe, (float*)out, n4, 1, d / 4, V, 0, scale, thr, inv_keep, seed, site, smer_seed_dev(), pos);
  else
    embed_pe_fwd_kernel<bf16><<<grid, 256, 0, st>>>(ids, emb, pe, (bf16*)out, n4, 1, d / 4, V, 0, scale, thr, inv_keep, seed, site, smer_seed_dev(), pos);
  SMER_CHECK_LAUNCH("smer_embed_pe_packed");
  return SMER_OK;
}

__global__ void pack_rows_kernel(const int64_t* __restrict__ ids, const int* __restrict__ cu, int B, int L,
                                 long long rows_alloc, int64_t* __restrict__ out_ids, int* __restrict__ out_pos) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < (long long)B * L) {
    const int b = (int)(i / L), l = (int)(i - (long long)b * L);
    const int base = cu[b], len = cu[b + 1] - base;
    if (l < len) {
      out_ids[base + l] = ids[i];
      out_pos[base + l] = l;
    }
  }
  // the tail [cu[B], rows_alloc): ghost rows that belong to no sequence
  const long long tail0 = cu[B];
  if (i < rows_alloc - tail0) {
    out_ids[tail0 + i] = 0;
    out_pos[tail0 + i] = 0;
  }
}

extern "C" int smer_pack_rows(const int64_t* ids, const int* cu, int B, int L, long long rows_alloc, int64_t* out_ids,
                              int* out_pos, void* stream) {
  SMER_CHECK_ARG(B > 0 && L > 0 && rows_alloc > 0, "smer_pack_rows: bad shape");
  long long n = (long long)B * L;
  if (n < rows_alloc) n = rows_alloc;
  pack_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ids, cu, B, L, rows_alloc, out_ids, out_pos);
  SMER_CHECK_LAUNCH("smer_pack_rows");
  return SMER_OK;
}

// rows [first_row[0], rows) of a row-major buffer := 0 (the ghost rows of a packed batch: no kernel of a sequence writes them)
__global__ void zero_tail_rows_kernel(uint4* __restrict__ buf, long long row_vec, long long rows, const int* __restrict__ first_row) {
  const long long r0 = first_row[0];
  const long long n = (rows - r0) * row_vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / row_vec, c = i - r * row_vec;
    buf[(r0 + r) * row_vec + c] = make_uint4(0u, 0u, 0u, 0u);
  }
}

extern "C" int smer_zero_tail_rows(void* buf, long long row_bytes, long long rows, const int* first_row_dev, void* stream) {
  SMER_CHECK_ARG(row_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(buf) & 15) == 0 && first_row_dev,
                 "smer_zero_tail_rows: rows must be contiguous multiples of 16 bytes");
  zero_tail_rows_kernel<<<64, 256, 0, (cudaStream_t)stream>>>((uint4*)buf, row_bytes / 16, rows, first_row_dev);
  SMER_CHECK_LAUNCH("smer_zero_tail_rows");
  return SMER_OK;
}

extern "C" int smer_embed_bwd(const int64_t* ids, const void* dout, int dtype, float* demb, int B, int L,
                              int d, int V, float scale, float dropout_p, uint64_t seed, uint64_t site,
                              void* stream) {
  SMER_CHECK_ARG(d % 4 == 0, "smer_embed_bwd: d must be a multiple of 4");
  long long n4 = (long long)B * L * (d / 4);
  uint32_t thr = dropout_p > 0.f ? dropout_threshold(dropout_p) : 0u;
  float inv_keep = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(n4, 256);
  if (dtype == SMER_DT_F32)
    embed_bwd_kernel<float><<<grid, 256, 0, st>>>(ids, (const float*)dout, demb, n4, d / 4, V, scale, thr, inv_keep, seed, site, smer_seed_dev());
  else
    embed_bwd_kernel<bf16><<<grid, 256, 0, st>>>(ids, (const bf16*)dout, demb, n4, d / 4, V, scale, thr, inv_keep, seed, site, smer_seed_dev());
  SMER_CHECK_LAUNCH("smer_embed_bwd");
  return SMER_OK;
}

// ---------------------------------------------------------------------------------------
// casts: 2-D copy with independent row strides and dtype; columns beyond `cols` up to
// `dst_cols` are zero-filled (pads V=309 to an aligned row pitch for the GEMMs).
// ---------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void cast2d_kernel(const TS* __restrict__ src, long long src_ld, TD* __restrict__ dst,
                              long long dst_ld, long long rows, int cols, int dst_cols) {
  long long n = rows * dst_cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i / dst_cols;
    int c = (int)(i - r * dst_cols);
    float v = c < cols ? to_f32(src[r * src_ld + c]) : 0.f;
    dst[r * dst_ld + c] = from_f32<TD>(v);
  }
}

extern "C" int smer_cast2d(const void* src, int src_dtype, long long src_ld, void* dst, int dst_dtype,
                           long long dst_ld, long long rows, int cols, int dst_cols, void* stream) {
  SMER_CHECK_ARG(dst_cols >= cols && rows >= 0, "smer_cast2d: bad shape");
  if (rows == 0) return SMER_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(rows * dst_cols, 256);
  if (src_dtype == SMER_DT_F32 && dst_dtype == SMER_DT_BF16)
    cast2d_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)src, src_ld, (bf16*)dst, dst_ld, rows, cols, dst_cols);
  else if (src_dtype == SMER_DT_BF16 && dst_dtype == SMER_DT_F32)
    cast2d_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)src, src_ld, (float*)dst, dst_ld, rows, cols, dst_cols);
  else if (src_dtype == SMER_DT_F32 && dst_dtype == SMER_DT_F32)
    cast2d_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, src_ld, (float*)dst, dst_ld, rows, cols, dst_cols);
  else
    cast2d_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)src, src_ld, (bf16*)dst, dst_ld, rows, cols, dst_cols);
  SMER_CHECK_LAUNCH("smer_cast2d");
  return SMER_OK;
}

// flat fp32 -> bf16 shadow of the parameter arena (one launch per step)
__global__ void cast_flat_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float v[4];
    load4(src + i * 4, v);
    store4(dst + i * 4, v);
  }
}

extern "C" int smer_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  SMER_CHECK_ARG(n % 4 == 0, "smer_cast_f32_to_bf16: n must be a multiple of 4");
  if (n == 0) return SMER_OK;
  cast_flat_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n / 4);
  SMER_CHECK_LAUNCH("smer_cast_f32_to_bf16");
  return SMER_OK;
}

// ---------------------------------------------------------------------------------------
// bias gradient: out[c] += sum_r x[r, c].  Block = 32 column-groups x 8 row-lanes; each thread
// owns 8 consecutive columns (one 16-byte load per row for bf16, two for fp32) and strides over
// a slab of rows with 4 loads in flight, then the 8 row-lanes are reduced through shared memory
// and one atomic per column per block is issued.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(h[k]);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long ld, float* __restrict__ out, long long rows, int cols,
              long long rows_per_block) {
  __shared__ float sm[8][32 * 8 + 8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + cg) * 8;
  long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (c0 < cols) {
    if (VEC) {
      long long r = r0 + rl;
      for (; r + 24 < r1; r += 32) {                 // 4 independent 16-byte loads in flight
        float v0[8], v1[8], v2[8], v3[8];
        ld8(x + r * ld + c0, v0);
        ld8(x + (r + 8) * ld + c0, v1);
        ld8(x + (r + 16) * ld + c0, v2);
        ld8(x + (r + 24) * ld + c0, v3);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += (v0[k] + v1[k]) + (v2[k] + v3[k]);
      }
      for (; r < r1; r += 8) {
        float v0[8];
        ld8(x + r * ld + c0, v0);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v0[k];
      }
    } else {
      for (long long r = r0 + rl; r < r1; r += 8)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (c0 + k < cols) acc[k] += to_f32(x[r * ld + c0 + k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sm[rl][cg * 8 + k] = acc[k];
  __syncthreads();
  const int c = threadIdx.x;                       // 256 columns per block
  const int gc = blockIdx.x * 256 + c;
  if (gc < cols) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k][c];
    atomicAdd(out + gc, s);
  }
}

extern "C" int smer_colsum(const void* x, int dtype, long long ld, float* out, long long rows, int cols,
                           void* stream) {
  if (rows == 0 || cols == 0) return SMER_OK;
  int gx = (cols + 255) / 256;
  long long target = (long long)smer_num_sms() * 8 / gx;
  if (target < 1) target = 1;
  long long rpb = (rows + target - 1) / target;
  if (rpb < 64) rpb = 64;
  int gy = (int)((rows + rpb - 1) / rpb);
  dim3 grid(gx, gy);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = cols % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (dtype == SMER_DT_F32) {
    if (vec) colsum_kernel<float, true><<<grid, 256, 0, st>>>((const float*)x, ld, out, rows, cols, rpb);
    else colsum_kernel<float, false><<<grid, 256, 0, st>>>((const float*)x, ld, out, rows, cols, rpb);
  } else {
    if (vec) colsum_kernel<bf16, true><<<grid, 256, 0, st>>>((const bf16*)x, ld, out, rows, cols, rpb);
    else colsum_kernel<bf16, false><<<grid, 256, 0, st>>>((const bf16*)x, ld, out, rows, cols, rpb);
  }
  SMER_CHECK_LAUNCH("smer_colsum");
  return SMER_OK;
}

// ---------------------------------------------------------------------------------------
// Mask inspection.  |kv_len[b]| = 1 + index of the last un-padded key; it bounds the key loop of the attention
// kernels.  The SIGN tells them whether the mask is a pure suffix (>= 0: key j masked <=> j >= kv_len[b], which is how
// the reference's collate functions pad, dataset.py:783-784,828 -- no per-key mask reads needed) or has masked keys
// before its last visible one (< 0: the kernels consult key_pad).
// ---------------------------------------------------------------------------------------
__global__ void kv_len_kernel(const uint8_t* __restrict__ pad, int* __restrict__ kv_len, int L) {
  int b = blockIdx.x;
  int best = 0, cnt = 0;
  for (int j = threadIdx.x; j < L; j += blockDim.x)
    if (!pad[(long long)b * L + j]) { best = j + 1; ++cnt; }
  __shared__ int sm[32], sc[32];
  for (int o = 16; o > 0; o >>= 1) {
    best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = best; sc[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = 0, c = 0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) { m = max(m, sm[w]); c += sc[w]; }
    kv_len[b] = c == m ? m : -m;
  }
}

extern "C" int smer_kv_len_from_pad(const uint8_t* pad, int* kv_len, int B, int L, void* stream) {
  kv_len_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(pad, kv_len, L);
  SMER_CHECK_LAUNCH("smer_kv_len_from_pad");
  return SMER_OK;
}

// Classifies a (T,T) additive float mask: flag = 0 all-zero, 1 exactly the nopeek mask
// (0 on/below the diagonal, -inf above; generation.py:193-206), 2 anything else.
__global__ void classify_mask_kernel(const float* __restrict__ m, long long ld, int T, int* __restrict__ flags) {
  // flags[0]: any non-zero on/below diagonal or any value other than 0/-inf above; flags[1]: any non -inf above
  // flags[2]: any non-zero above
  int bad = 0, not_inf_above = 0, nonzero_above = 0;
  long long n = (long long)T * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / T), c = (int)(i - (long long)r * T);
    float v = m[r * ld + c];
    if (c <= r) {
      if (v != 0.f) bad = 1;
    } else {
      if (v != 0.f) nonzero_above = 1;
      if (!(isinf(v) && v < 0.f)) not_inf_above = 1;
    }
  }
  if (bad) atomicOr(flags + 0, 1);
  if (not_inf_above) atomicOr(flags + 1, 1);
  if (nonzero_above) atomicOr(flags + 2, 1);
}

extern "C" int smer_classify_mask(const float* mask, long long ld, int T, int* flags3, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SMER_CUDA(cudaMemsetAsync(flags3, 0, 3 * sizeof(int), st));
  classify_mask_kernel<<<grid_for((long long)T * T, 256), 256, 0, st>>>(mask, ld, T, flags3);
  SMER_CHECK_LAUNCH("smer_classify_mask");
  return SMER_OK;
}

// ---------------------------------------------------------------------------------------
// Adam (train.py:264: torch.optim.Adam defaults -- no weight decay, no amsgrad), one launch
// over the flat parameter arena; optionally refreshes the bf16 shadow in the same pass.
// ---------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, bf16* __restrict__ shadow, long long n4, float lr, float b1,
                            float b2, float eps, float bc1, float bc2_sqrt, float gscale,
                            const unsigned long long* __restrict__ step_dev) {
  if (step_dev) {                 // step number lives on the device (CUDA-graph replays advance it)
    const float st = (float)(*step_dev);
    bc1 = 1.f - powf(b1, st);
    bc2_sqrt = sqrtf(1.f - powf(b2, st));
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float pp[4], gg[4], mm[4], vv[4];
    load4(p + i * 4, pp);
    load4(g + i * 4, gg);
    load4(m + i * 4, mm);
    load4(v + i * 4, vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = gg[k] * gscale;
      mm[k] = b1 * mm[k] + (1.f - b1) * gk;
      vv[k] = b2 * vv[k] + (1.f - b2) * gk * gk;
      float denom = sqrtf(vv[k]) / bc2_sqrt + eps;
      pp[k] -= (lr / bc1) * (mm[k] / denom);
    }
    store4(p + i * 4, pp);
    store4(m + i * 4, mm);
    store4(v + i * 4, vv);
    if (shadow) store4(shadow + i * 4, pp);
  }
}

static int adam_launch(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, int step,
                       const uint64_t* step_dev, float lr, float beta1, float beta2, float eps, float grad_scale,
                       void* stream) {
  SMER_CHECK_ARG(n % 4 == 0 && (step >= 1 || step_dev), "smer_adam_step: n must be a multiple of 4 and step >= 1");
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, (bf16*)bf16_shadow, n / 4, lr, beta1, beta2, eps, bc1, sqrtf(bc2), grad_scale,
      reinterpret_cast<const unsigned long long*>(step_dev));
  SMER_CHECK_LAUNCH("smer_adam_step");
  return SMER_OK;
}

extern "C" int smer_adam_step(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n,
                              int step, float lr, float beta1, float beta2, float eps, float grad_scale,
                              void* stream) {
  return adam_launch(p, g, m, v, bf16_shadow, n, step, nullptr, lr, beta1, beta2, eps, grad_scale, stream);
}

// Multi-tensor form for ordinary (non-contiguous-arena) nn.Parameters: one launch over a device table of
// tensors, any element counts (scalar head/tail around the 16-byte-aligned body).  All tensors share `step`.
__global__ void __launch_bounds__(256)
adam_multi_kernel(const smer_adam_tensor* __restrict__ table, float lr, float b1, float b2, float eps, float bc1,
                  float bc2_sqrt, float gscale) {
  const smer_adam_tensor t = table[blockIdx.y];
  float* __restrict__ p = t.p;
  const float* __restrict__ g = t.g;
  float* __restrict__ m = t.m;
  float* __restrict__ v = t.v;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    const float gk = gg * gscale;
    mm = b1 * mm + (1.f - b1) * gk;
    vv = b2 * vv + (1.f - b2) * gk * gk;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= (lr / bc1) * (mm / denom);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long long n4 = vec ? t.n / 4 : 0;
  const long long stride = (long long)gridDim.x * blockDim.x, tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long i = tid; i < n4; i += stride) {
    float pp[4], gg[4], mm[4], vv[4];
    load4(p + i * 4, pp);
    load4(g + i * 4, gg);
    load4(m + i * 4, mm);
    load4(v + i * 4, vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) upd(pp[k], gg[k], mm[k], vv[k]);
    store4(p + i * 4, pp);
    store4(m + i * 4, mm);
    store4(v + i * 4, vv);
  }
  for (long long i = n4 * 4 + tid; i < t.n; i += stride) upd(p[i], g[i], m[i], v[i]);
}

extern "C" int smer_adam_multi(const smer_adam_tensor* table_dev, int n_tensors, long long max_n, int step, float lr,
                               float beta1, float beta2, float eps, float grad_scale, void* stream) {
  SMER_CHECK_ARG(table_dev && n_tensors >= 0 && step >= 1, "smer_adam_multi: bad arguments");
  if (n_tensors == 0) return SMER_OK;
  SMER_CHECK_ARG(n_tensors <= 65535, "smer_adam_multi: at most 65535 tensors per launch");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  long long bx = (max_n / 4 + 255) / 256;
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;                        // the largest tensor grid-strides; small ones finish in their first block
  dim3 grid((unsigned)bx, (unsigned)n_tensors);
  adam_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table_dev, lr, beta1, beta2, eps, bc1, sqrtf(bc2), grad_scale);
  SMER_CHECK_LAUNCH("smer_adam_multi");
  return SMER_OK;
}

extern "C" int smer_adam_step_dev(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n,
                                  const uint64_t* step_dev, float lr, float beta1, float beta2, float eps,
                                  float grad_scale, void* stream) {
  SMER_CHECK_ARG(step_dev != nullptr, "smer_adam_step_dev: null step pointer");
  return adam_launch(p, g, m, v, bf16_shadow, n, 0, step_dev, lr, beta1, beta2, eps, grad_scale, stream);
}
