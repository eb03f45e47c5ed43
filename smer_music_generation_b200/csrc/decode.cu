// KV-cached decode-step attention: append the new token's K/V to the self-attention cache and
// attend one query per (piece, head) over the cache (or over the fixed cross-attention K/V of
// the encoder memory).  The reference has no cache -- generation.py:209-225 re-runs the whole
// model per token; the values computed here equal the last row of that
// recompute (decoder self-attention is causal, so earlier rows never change).
//
// HBM-bound (2*L*d*sizeof(T) bytes per piece per layer): 8 lanes own one key
// row (16-byte loads for bf16), 4 key rows per warp per iteration, online softmax per lane
// group, merged through shuffles/shared memory; optional split over the key range with a
// second merge kernel so that a single piece still fills the GPU.
#include "common.cuh"
#include "../../include/smer_b200.h"

constexpr int DEC_THREADS = 128;          // long key ranges (cross-attention K/V): 16 key rows per CTA iteration
constexpr int DEC_THREADS_SHORT = 32;     // short ranges (the growing self-attention cache): one warp per (piece, head)

template <typename T, int EPL>
__device__ __forceinline__ void load_row(const T* p, float (&v)[EPL]) {
#pragma unroll
  for (int c = 0; c < EPL; c += 4) {
    float t[4];
    load4(p + c, t);
    v[c] = t[0]; v[c + 1] = t[1]; v[c + 2] = t[2]; v[c + 3] = t[3];
  }
}
// 8 bf16 = one 16-byte streaming load (the K/V caches are read once per step: do not allocate in L1)
template <>
__device__ __forceinline__ void load_row<bf16, 8>(const bf16* p, float (&v)[8]) {
  uint4 t;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(h[k]);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
}
template <typename T, int DH, int THREADS>
__global__ void __launch_bounds__(THREADS) decode_attn_kernel(smer_decode_attn_args a) {
  constexpr int EPL = DH >= 64 ? 8 : 4;          // elements per lane
  constexpr int LPK = DH / EPL;                  // lanes per key row (8, or 4 for dh=16)
  constexpr int DEC_GROUPS = THREADS / LPK;      // key rows in flight per CTA iteration
  __shared__ float sm_m[DEC_GROUPS], sm_l[DEC_GROUPS];
  __shared__ float sm_acc[DEC_GROUPS][DH];
  int h = blockIdx.x, s = blockIdx.y, sp = blockIdx.z;
  pdl_trigger();
  pdl_wait();
  if (a.done && a.done[s]) return;                // a finished piece: nothing is appended, its K/V is not streamed
  int tid = threadIdx.x;
  int grp = tid / LPK, lig = tid % LPK;
  int npast = a.kv_len[s];                        // keys already in the cache
  int nkeys = npast + (a.new_k ? 1 : 0);
  const T* q = (const T*)a.q + (long long)s * a.ldq + h * DH + lig * EPL;
  float qr[EPL];
  load_row<T, EPL>(q, qr);
#pragma unroll
  for (int c = 0; c < EPL; ++c) qr[c] *= a.scale;
  T* kc = (T*)a.k_cache + (long long)s * a.cache_stride + h * DH + lig * EPL;
  T* vc = (T*)a.v_cache + (long long)s * a.cache_stride + h * DH + lig * EPL;
  const T* nk = a.new_k ? (const T*)a.new_k + (long long)s * a.ld_new + h * DH + lig * EPL : nullptr;
  const T* nv = a.new_v ? (const T*)a.new_v + (long long)s * a.ld_new + h * DH + lig * EPL : nullptr;
  if (nk && sp == 0 && grp == 0 && npast < a.cache_len) {   // KV-cache append (one writer)
    float t[EPL];
    load_row<T, EPL>(nk, t);
#pragma unroll
    for (int c = 0; c < EPL; c += 4) {
      float u[4] = {t[c], t[c + 1], t[c + 2], t[c + 3]};
      store4(kc + (long long)npast * a.ld_cache + c, u);
    }
    load_row<T, EPL>(nv, t);
#pragma unroll
    for (int c = 0; c < EPL; c += 4) {
      float u[4] = {t[c], t[c + 1], t[c + 2], t[c + 3]};
      store4(vc + (long long)npast * a.ld_cache + c, u);
    }
  }
  int per = (nkeys + a.splits - 1) / a.splits;
  int j0 = sp * per, j1 = min(nkeys, j0 + per);
  float m = -INFINITY, l = 0.f, acc[EPL];
#pragma unroll
  for (int c = 0; c < EPL; ++c) acc[c] = 0.f;
  const uint8_t* pad = a.key_pad ? a.key_pad + (long long)s * a.ld_pad : nullptr;
  // HBM-bound loop: every thread first issues the K and V loads of U key rows (2*U independent
  // 16-byte requests in flight), then folds them into its online softmax.
  constexpr int U = 4;
  for (int jb = j0; jb < j1; jb += DEC_GROUPS * U) {    // trip count is uniform across the CTA
    float kr[U][EPL], vr[U][EPL];
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * DEC_GROUPS + grp;
      valid[u] = j < j1;
      const int jc = valid[u] ? j : j1 - 1;
      const bool fresh = nk && jc == npast;
      const T* kp = fresh ? nk : kc + (long long)jc * a.ld_cache;
      const T* vp = fresh ? nv : vc + (long long)jc * a.ld_cache;
      load_row<T, EPL>(kp, kr[u]);
      load_row<T, EPL>(vp, vr[u]);
      if (pad && valid[u] && pad[jc]) valid[u] = false;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < EPL; ++c) d = fmaf(qr[c], kr[u][c], d);
#pragma unroll
      for (int o = 1; o < LPK; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (valid[u]) {
        float mn = fmaxf(m, d);
        float alpha = __expf(m - mn), pr = __expf(d - mn);
        l = l * alpha + pr;
#pragma unroll
        for (int c = 0; c < EPL; ++c) acc[c] = fmaf(pr, vr[u][c], acc[c] * alpha);
        m = mn;
      }
    }
  }
  if (lig == 0) { sm_m[grp] = m; sm_l[grp] = l; }
#pragma unroll
  for (int c = 0; c < EPL; ++c) sm_acc[grp][lig * EPL + c] = acc[c];
  __syncthreads();
  for (int c = tid; c < DH; c += THREADS) {
    float M = -INFINITY;
#pragma unroll
    for (int g = 0; g < DEC_GROUPS; ++g) M = fmaxf(M, sm_m[g]);
    float L = 0.f, o = 0.f;
    if (M > -INFINITY) {
#pragma unroll
      for (int g = 0; g < DEC_GROUPS; ++g) {
        float w = expf(sm_m[g] - M);
        L += w * sm_l[g];
        o += w * sm_acc[g][c];
      }
    }
    if (a.splits == 1) {
      T* out = (T*)a.out + (long long)s * a.ldo + h * DH;
      out[c] = from_f32<T>(L > 0.f ? o / L : 0.f);
    } else {
      float* w = a.workspace + (((long long)s * gridDim.x + h) * a.splits + sp) * (DH + 2);
      w[2 + c] = o;
      if (c == 0) { w[0] = M; w[1] = L; }
    }
  }
}

template <typename T, int DH>
__global__ void decode_merge_kernel(smer_decode_attn_args a, int H) {
  int h = blockIdx.x, s = blockIdx.y, c = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (a.done && a.done[s]) return;
  const float* w = a.workspace + ((long long)s * H + h) * a.splits * (DH + 2);
  float M = -INFINITY;
  for (int sp = 0; sp < a.splits; ++sp) M = fmaxf(M, w[sp * (DH + 2)]);
  float L = 0.f, o = 0.f;
  if (M > -INFINITY)
    for (int sp = 0; sp < a.splits; ++sp) {
      float e = expf(w[sp * (DH + 2)] - M);
      L += e * w[sp * (DH + 2) + 1];
      o += e * w[sp * (DH + 2) + 2 + c];
    }
  T* out = (T*)a.out + (long long)s * a.ldo + h * DH;
  out[c] = from_f32<T>(L > 0.f ? o / L : 0.f);
}

extern "C" long long smer_decode_attn_workspace_bytes(int n_seq, int H, int dh, int splits) {
  return splits <= 1 ? 0 : (long long)n_seq * H * splits * (dh + 2) * (long long)sizeof(float);
}

template <typename T>
static int decode_launch(const smer_decode_attn_args& a, cudaStream_t st) {
  dim3 grid(a.H, a.n_seq, a.splits);
  switch (a.dh) {
    case 16: smer_launch_pdl(decode_attn_kernel<T, 16, DEC_THREADS>, grid, dim3(DEC_THREADS), 0, st, a); break;
    case 32: smer_launch_pdl(decode_attn_kernel<T, 32, DEC_THREADS>, grid, dim3(DEC_THREADS), 0, st, a); break;
    case 64:
      // the self-attention cache (new_k given) holds at most cache_len keys, usually a few hundred:
      // one warp per (piece, head) avoids paying a 128-thread CTA's fixed costs for a handful of keys -- as long as
      // there are enough (piece, head) pairs to fill the GPU with warps; with few pieces a single warp walking a
      // several-hundred-key cache is latency-bound (128 pieces: +30 us per layer at 500 cached tokens), so those
      // take the 128-thread CTA (16 key rows in flight)
      if (a.new_k && a.cache_len <= 2048 && a.splits == 1 && (long long)a.n_seq * a.H > 4096)
        smer_launch_pdl(decode_attn_kernel<T, 64, DEC_THREADS_SHORT>, grid, dim3(DEC_THREADS_SHORT), 0, st, a);
      else
        smer_launch_pdl(decode_attn_kernel<T, 64, DEC_THREADS>, grid, dim3(DEC_THREADS), 0, st, a);
      break;
    default: smer_set_error("smer_decode_attn: head dim %d unsupported (16/32/64)", a.dh); return SMER_ERR_UNSUPPORTED;
  }
  if (a.splits > 1) {
    dim3 g2(a.H, a.n_seq);
    switch (a.dh) {
      case 16: smer_launch_pdl(decode_merge_kernel<T, 16>, g2, dim3(16), 0, st, a, a.H); break;
      case 32: smer_launch_pdl(decode_merge_kernel<T, 32>, g2, dim3(32), 0, st, a, a.H); break;
      case 64: smer_launch_pdl(decode_merge_kernel<T, 64>, g2, dim3(64), 0, st, a, a.H); break;
    }
  }
  return SMER_OK;
}

extern "C" int smer_decode_attn(const smer_decode_attn_args* a, void* stream) {
  SMER_CHECK_ARG(a && a->q && a->k_cache && a->v_cache && a->out && a->kv_len, "smer_decode_attn: null args");
  SMER_CHECK_ARG(a->n_seq > 0 && a->H > 0 && a->splits >= 1, "smer_decode_attn: bad sizes");
  SMER_CHECK_ARG(a->splits == 1 || a->workspace, "smer_decode_attn: splits > 1 needs a workspace");
  SMER_CHECK_ARG((a->new_k == nullptr) == (a->new_v == nullptr), "smer_decode_attn: new_k/new_v must come together");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = a->dtype == SMER_DT_F32 ? decode_launch<float>(*a, st) : decode_launch<bf16>(*a, st);
  if (rc) return rc;
  SMER_CHECK_LAUNCH("smer_decode_attn");
  return SMER_OK;
}

// ---------------------------------------------------------------------------------------
// Gather the next input token of every piece and bump the cache length: ids[s] =
// tok_buf[s, cur_len[s]-1], pos[s] = cur_len[s]-1 (finished pieces keep re-feeding their last
// token; their results are ignored by the sampler).
// ---------------------------------------------------------------------------------------
__global__ void decode_gather_kernel(const long long* __restrict__ tok_buf, const int* __restrict__ cur_len,
                                     const int* __restrict__ fed_len, long long* __restrict__ ids,
                                     int* __restrict__ pos, int n, int max_len) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int p = cur_len[s] - 1;
  if (fed_len) p = min(p, fed_len[s]);
  p = p < 0 ? 0 : (p >= max_len ? max_len - 1 : p);
  ids[s] = tok_buf[(long long)s * max_len + p];
  pos[s] = p;
}

extern "C" int smer_decode_gather(const int64_t* tok_buf, const int* cur_len, const int* fed_len, int64_t* ids,
                                  int* pos, int n_seq, int max_len, void* stream) {
  decode_gather_kernel<<<(n_seq + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const long long*)tok_buf, cur_len,
                                                                               fed_len, (long long*)ids, pos, n_seq,
                                                                               max_len);
  SMER_CHECK_LAUNCH("smer_decode_gather");
  return SMER_OK;
}

// embed + positional for one token per piece at per-piece positions (decode step input)
template <typename T>
__global__ void embed_step_kernel(const long long* __restrict__ ids, const int* __restrict__ pos,
                                  const float* __restrict__ emb, const float* __restrict__ pe, T* __restrict__ out,
                                  int n, int d4, int V, float scale) {
  long long tot = (long long)n * d4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    int s = (int)(i / d4), c4 = (int)(i % d4);
    long long id = ids[s];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    float e[4], p[4], o[4];
    load4(emb + id * d4 * 4 + c4 * 4, e);
    load4(pe + (long long)pos[s] * d4 * 4 + c4 * 4, p);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = e[k] * scale + p[k];
    store4(out + i * 4, o);
  }
}

extern "C" int smer_embed_step(const int64_t* ids, const int* pos, const float* emb, const float* pe, void* out,
                               int out_dtype, int n_seq, int d, int V, float scale, void* stream) {
  SMER_CHECK_ARG(d % 4 == 0, "smer_embed_step: d must be a multiple of 4");
  long long tot = (long long)n_seq * (d / 4);
  int grid = (int)((tot + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == SMER_DT_F32)
    embed_step_kernel<float><<<grid, 256, 0, st>>>((const long long*)ids, pos, emb, pe, (float*)out, n_seq, d / 4, V, scale);
  else
    embed_step_kernel<bf16><<<grid, 256, 0, st>>>((const long long*)ids, pos, emb, pe, (bf16*)out, n_seq, d / 4, V, scale);
  SMER_CHECK_LAUNCH("smer_embed_step");
  return SMER_OK;
}

// gather + embed in one launch: one CTA of d/4 threads per piece
template <typename T>
__global__ void decode_embed_kernel(const long long* __restrict__ tok_buf, const int* __restrict__ cur_len,
                                    const int* __restrict__ fed_len, const int* __restrict__ done, int* __restrict__ pos,
                                    const float* __restrict__ emb, const float* __restrict__ pe, T* __restrict__ out,
                                    int max_len, int d4, int V, float scale) {
  const int s = blockIdx.x;
  pdl_trigger();
  pdl_wait();
  if (done && done[s]) return;
  int p = cur_len[s] - 1;
  if (fed_len) p = min(p, fed_len[s]);
  p = p < 0 ? 0 : (p >= max_len ? max_len - 1 : p);
  long long id = tok_buf[(long long)s * max_len + p];
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  if (threadIdx.x == 0) pos[s] = p;
  for (int c4 = threadIdx.x; c4 < d4; c4 += blockDim.x) {
    float e[4], q[4], o[4];
    load4(emb + id * d4 * 4 + c4 * 4, e);
    load4(pe + (long long)p * d4 * 4 + c4 * 4, q);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = e[k] * scale + q[k];
    store4(out + ((long long)s * d4 + c4) * 4, o);
  }
}

extern "C" int smer_decode_embed(const int64_t* tok_buf, const int* cur_len, const int* fed_len, const int* done, int* pos,
                                 const float* emb, const float* pe, void* out, int out_dtype, int n_seq, int max_len, int d,
                                 int V, float scale, void* stream) {
  SMER_CHECK_ARG(d % 4 == 0 && n_seq > 0, "smer_decode_embed: d must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = d / 4 < 128 ? (d / 4 + 31) / 32 * 32 : 128;
  if (out_dtype == SMER_DT_F32)
    smer_launch_pdl(decode_embed_kernel<float>, dim3(n_seq), dim3(threads), 0, st, (const long long*)tok_buf, cur_len, fed_len, done, pos, emb, pe, (float*)out, max_len, d / 4, V, scale);
  else
    smer_launch_pdl(decode_embed_kernel<bf16>, dim3(n_seq), dim3(threads), 0, st, (const long long*)tok_buf, cur_len, fed_len, done, pos, emb, pe, (bf16*)out, max_len, d / 4, V, scale);
  SMER_CHECK_LAUNCH("smer_decode_embed");
  return SMER_OK;
}
