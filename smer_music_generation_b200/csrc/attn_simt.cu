// Flash-style (online-softmax, no score materialisation) attention on CUDA cores, FP32 math.
// Arithmetic of F.multi_head_attention_forward as the reference reaches it through
// nn.MultiheadAttention (transformer.py:389,459,463): softmax(q k^T / sqrt(dh) + masks) v with
// dropout on the probabilities; masks = causal (tgt_mask), per-key padding (key_padding_mask),
// optional arbitrary additive (Lq,Lk) mask.  Used for the fp32 parity path and for head sizes
// the tcgen05 kernel (attn_tc.cu) does not cover.
//
// Layout: token-major rows.  q[(b*Lq+i)*ldq + h*DH + c], same for k/v/o with their own pitches,
// so the packed QKV projection output is consumed in place.  lse[(b*H+h)*Lq + i].
// Dropout element (b,h,i,j): attn_keep(attn_row_key(attn_site_key(seed, site), (b*H+h)*Lq + i), j, thr) -- common.cuh.
#include "common.cuh"
#include "../../include/smer_b200.h"

constexpr int KT = 32;          // keys (fwd, dq) or queries (dkv) staged per tile
constexpr int NTHREADS = 128;

struct AttnParams {
  const void *q, *k, *v, *o, *dout;
  void *out, *dq, *dk, *dv;
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  float* lse;
  float* dsum;
  const uint8_t* pad;       // [B, Lk] 1 = masked key, or null
  const int* kv_len;        // [B] or null
  const float* addmask;     // [Lq, Lk] additive or null
  long long ldmask;
  int B, H, Lq, Lk;
  float scale;
  int causal;
  int q_pos0;             // causal test is j > i + q_pos0 (incremental decode appends queries)
  uint32_t thr;
  float inv_keep;
  uint64_t seed, site;
  const unsigned long long* seed_dev;
};

template <int DPT, int TPR>
__device__ __forceinline__ float row_dot(const float (&a)[DPT], const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < DPT; c += 4) {
    float4 t = *reinterpret_cast<const float4*>(b + c);
    s = fmaf(a[c], t.x, s);
    s = fmaf(a[c + 1], t.y, s);
    s = fmaf(a[c + 2], t.z, s);
    s = fmaf(a[c + 3], t.w, s);
  }
  if (TPR == 2) s += __shfl_xor_sync(0xffffffffu, s, 1);
  return s;
}

// keys 4*j4 .. 4*j4+3 of one row: two pair words of the 16-key block (4*j4) >> 4
__device__ __forceinline__ void drop_lanes(const AttnParams& p, uint32_t rowkey, int j4, float (&m)[4]) {
  const uint32_t w = attn_block_word(rowkey, (uint32_t)j4 >> 2);
  uint32_t mul = 1u, add = 0u;
  for (int k = 0; k < (j4 & 3) * 2; ++k) { mul *= ATTN_A; add = add * ATTN_A + ATTN_C; }
  const uint32_t x0 = w * mul + add, x1 = x0 * ATTN_A + ATTN_C;
  m[0] = attn_keep_field(x0 & 0xFFFFu, p.thr) ? p.inv_keep : 0.f;
  m[1] = attn_keep_field(x0 >> 16, p.thr) ? p.inv_keep : 0.f;
  m[2] = attn_keep_field(x1 & 0xFFFFu, p.thr) ? p.inv_keep : 0.f;
  m[3] = attn_keep_field(x1 >> 16, p.thr) ? p.inv_keep : 0.f;
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
template <typename T, int DH, int TPR>
__global__ void __launch_bounds__(NTHREADS) attn_fwd_simt_kernel(AttnParams p) {
  constexpr int DPT = DH / TPR;
  constexpr int ROWS = NTHREADS / TPR;
  __shared__ __align__(16) float Ks[KT][DH];
  __shared__ __align__(16) float Vs[KT][DH];
  __shared__ uint8_t Ms[KT];
  int b = blockIdx.z, h = blockIdx.y;
  int i = blockIdx.x * ROWS + threadIdx.x / TPR;
  int part = threadIdx.x % TPR;
  bool row_ok = i < p.Lq;
  int ic = row_ok ? i : p.Lq - 1;
  const T* q = (const T*)p.q + ((long long)b * p.Lq + ic) * p.ldq + h * DH + part * DPT;
  float qr[DPT], acc[DPT];
#pragma unroll
  for (int c = 0; c < DPT; ++c) {
    qr[c] = to_f32(q[c]) * p.scale;
    acc[c] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  int kend = p.kv_len ? min(abs(p.kv_len[b]), p.Lk) : p.Lk;
  if (p.causal) kend = min(kend, min(p.Lq, (int)(blockIdx.x + 1) * ROWS) + p.q_pos0);
  long long rowid = ((long long)b * p.H + h) * p.Lq + ic;
  const uint32_t rowkey = p.thr ? attn_row_key(attn_site_key(eff_seed(p.seed, p.seed_dev), p.site), rowid) : 0u;
  const T* kb = (const T*)p.k + (long long)b * p.Lk * p.ldk + h * DH;
  const T* vb = (const T*)p.v + (long long)b * p.Lk * p.ldv + h * DH;

  for (int j0 = 0; j0 < kend; j0 += KT) {
    __syncthreads();
    for (int e = threadIdx.x; e < KT * DH; e += NTHREADS) {
      int jj = e / DH, c = e % DH;
      int j = j0 + jj;
      float kv = 0.f, vv = 0.f;
      if (j < p.Lk) {
        kv = to_f32(kb[(long long)j * p.ldk + c]);
        vv = to_f32(vb[(long long)j * p.ldv + c]);
      }
      Ks[jj][c] = kv;
      Vs[jj][c] = vv;
    }
    if (threadIdx.x < KT) {
      int j = j0 + threadIdx.x;
      Ms[threadIdx.x] = (j >= kend) || (p.pad && p.pad[(long long)b * p.Lk + j]);
    }
    __syncthreads();
    float s[KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < KT; ++jj) {
      float v = row_dot<DPT, TPR>(qr, &Ks[jj][part * DPT]);
      int j = j0 + jj;
      if (p.addmask && j < p.Lk) v += p.addmask[(long long)ic * p.ldmask + j];
      bool masked = Ms[jj] || (p.causal && j > i + p.q_pos0);
      v = masked ? -INFINITY : v;
      s[jj] = v;
      tmax = fmaxf(tmax, v);
    }
    float mnew = fmaxf(m, tmax);
    if (mnew == -INFINITY) continue;        // nothing visible yet for this row
    float alpha = expf(m - mnew);
    l *= alpha;
#pragma unroll
    for (int c = 0; c < DPT; ++c) acc[c] *= alpha;
    m = mnew;
#pragma unroll
    for (int j4 = 0; j4 < KT / 4; ++j4) {
      float dm[4] = {1.f, 1.f, 1.f, 1.f};
      if (p.thr) drop_lanes(p, rowkey, (j0 >> 2) + j4, dm);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int jj = j4 * 4 + u;
        float pr = expf(s[jj] - m);            // exp(-inf) = 0 for masked keys
        l += pr;
        pr *= dm[u];
#pragma unroll
        for (int c = 0; c < DPT; c += 4) {
          float4 t = *reinterpret_cast<const float4*>(&Vs[jj][part * DPT + c]);
          acc[c] = fmaf(pr, t.x, acc[c]);
          acc[c + 1] = fmaf(pr, t.y, acc[c + 1]);
          acc[c + 2] = fmaf(pr, t.z, acc[c + 2]);
          acc[c + 3] = fmaf(pr, t.w, acc[c + 3]);
        }
      }
    }
  }
  if (row_ok) {
    float inv = l > 0.f ? 1.f / l : 0.f;
    T* o = (T*)p.out + ((long long)b * p.Lq + i) * p.ldo + h * DH + part * DPT;
#pragma unroll
    for (int c = 0; c < DPT; ++c) o[c] = from_f32<T>(acc[c] * inv);
    if (part == 0 && p.lse) p.lse[rowid] = l > 0.f ? m + logf(l) : -INFINITY;
  }
}

// ---------------------------------------------------------------------------------------
// backward preprocessing: dsum[b,h,i] = sum_c dO[i,c] * O[i,c]
// ---------------------------------------------------------------------------------------
template <typename T, int DH>
__global__ void attn_bwd_dot_kernel(AttnParams p) {
  long long n = (long long)p.B * p.H * p.Lq;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    int i = (int)(t % p.Lq);
    int h = (int)((t / p.Lq) % p.H);
    int b = (int)(t / ((long long)p.Lq * p.H));
    const T* o = (const T*)p.o + ((long long)b * p.Lq + i) * p.ldo + h * DH;
    const T* g = (const T*)p.dout + ((long long)b * p.Lq + i) * p.lddo + h * DH;
    float s = 0.f;
#pragma unroll 8
    for (int c = 0; c < DH; ++c) s = fmaf(to_f32(o[c]), to_f32(g[c]), s);
    p.dsum[t] = s;
  }
}

// ---------------------------------------------------------------------------------------
// backward dQ: one (pair of) thread(s) per query row, loop over keys
// ---------------------------------------------------------------------------------------
template <typename T, int DH, int TPR>
__global__ void __launch_bounds__(NTHREADS) attn_bwd_dq_kernel(AttnParams p) {
  constexpr int DPT = DH / TPR;
  constexpr int ROWS = NTHREADS / TPR;
  __shared__ __align__(16) float Ks[KT][DH];
  __shared__ __align__(16) float Vs[KT][DH];
  __shared__ uint8_t Ms[KT];
  int b = blockIdx.z, h = blockIdx.y;
  int i = blockIdx.x * ROWS + threadIdx.x / TPR;
  int part = threadIdx.x % TPR;
  bool row_ok = i < p.Lq;
  int ic = row_ok ? i : p.Lq - 1;
  const T* q = (const T*)p.q + ((long long)b * p.Lq + ic) * p.ldq + h * DH + part * DPT;
  const T* g = (const T*)p.dout + ((long long)b * p.Lq + ic) * p.lddo + h * DH + part * DPT;
  float qr[DPT], gr[DPT], acc[DPT];
#pragma unroll
  for (int c = 0; c < DPT; ++c) {
    qr[c] = to_f32(q[c]) * p.scale;
    gr[c] = to_f32(g[c]);
    acc[c] = 0.f;
  }
  long long rowid = ((long long)b * p.H + h) * p.Lq + ic;
  float lse = p.lse[rowid], dsum = p.dsum[rowid];
  const uint32_t rowkey = p.thr ? attn_row_key(attn_site_key(eff_seed(p.seed, p.seed_dev), p.site), rowid) : 0u;
  int kend = p.kv_len ? min(abs(p.kv_len[b]), p.Lk) : p.Lk;
  if (p.causal) kend = min(kend, min(p.Lq, (int)(blockIdx.x + 1) * ROWS) + p.q_pos0);
  const T* kb = (const T*)p.k + (long long)b * p.Lk * p.ldk + h * DH;
  const T* vb = (const T*)p.v + (long long)b * p.Lk * p.ldv + h * DH;
  for (int j0 = 0; j0 < kend; j0 += KT) {
    __syncthreads();
    for (int e = threadIdx.x; e < KT * DH; e += NTHREADS) {
      int jj = e / DH, c = e % DH;
      int j = j0 + jj;
      float kv = 0.f, vv = 0.f;
      if (j < p.Lk) {
        kv = to_f32(kb[(long long)j * p.ldk + c]);
        vv = to_f32(vb[(long long)j * p.ldv + c]);
      }
      Ks[jj][c] = kv;
      Vs[jj][c] = vv;
    }
    if (threadIdx.x < KT) {
      int j = j0 + threadIdx.x;
      Ms[threadIdx.x] = (j >= kend) || (p.pad && p.pad[(long long)b * p.Lk + j]);
    }
    __syncthreads();
#pragma unroll 2
    for (int j4 = 0; j4 < KT / 4; ++j4) {
      float dm[4] = {1.f, 1.f, 1.f, 1.f};
      if (p.thr) drop_lanes(p, rowkey, (j0 >> 2) + j4, dm);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int jj = j4 * 4 + u;
        int j = j0 + jj;
        float s = row_dot<DPT, TPR>(qr, &Ks[jj][part * DPT]);
        if (p.addmask && j < p.Lk) s += p.addmask[(long long)ic * p.ldmask + j];
        bool masked = Ms[jj] || (p.causal && j > i + p.q_pos0);
        float pr = masked ? 0.f : expf(s - lse);
        float dp = row_dot<DPT, TPR>(gr, &Vs[jj][part * DPT]) * dm[u];
        float ds = pr * (dp - dsum);
#pragma unroll
        for (int c = 0; c < DPT; c += 4) {
          float4 t = *reinterpret_cast<const float4*>(&Ks[jj][part * DPT + c]);
          acc[c] = fmaf(ds, t.x, acc[c]);
          acc[c + 1] = fmaf(ds, t.y, acc[c + 1]);
          acc[c + 2] = fmaf(ds, t.z, acc[c + 2]);
          acc[c + 3] = fmaf(ds, t.w, acc[c + 3]);
        }
      }
    }
  }
  if (row_ok) {
    T* o = (T*)p.dq + ((long long)b * p.Lq + i) * p.lddq + h * DH + part * DPT;
#pragma unroll
    for (int c = 0; c < DPT; ++c) o[c] = from_f32<T>(acc[c] * p.scale);
  }
}

// ---------------------------------------------------------------------------------------
// backward dK/dV: one (pair of) thread(s) per key row, loop over queries
// ---------------------------------------------------------------------------------------
template <typename T, int DH, int TPR>
__global__ void __launch_bounds__(NTHREADS) attn_bwd_dkv_kernel(AttnParams p) {
  constexpr int DPT = DH / TPR;
  constexpr int ROWS = NTHREADS / TPR;
  __shared__ __align__(16) float Qs[KT][DH];
  __shared__ __align__(16) float Gs[KT][DH];
  __shared__ float Ls[KT], Ds[KT];
  int b = blockIdx.z, h = blockIdx.y;
  int j = blockIdx.x * ROWS + threadIdx.x / TPR;
  int part = threadIdx.x % TPR;
  bool row_ok = j < p.Lk;
  int jc = row_ok ? j : p.Lk - 1;
  const T* kp = (const T*)p.k + ((long long)b * p.Lk + jc) * p.ldk + h * DH + part * DPT;
  const T* vp = (const T*)p.v + ((long long)b * p.Lk + jc) * p.ldv + h * DH + part * DPT;
  float kr[DPT], vr[DPT], dk[DPT], dv[DPT];
#pragma unroll
  for (int c = 0; c < DPT; ++c) {
    kr[c] = to_f32(kp[c]);
    vr[c] = to_f32(vp[c]);
    dk[c] = dv[c] = 0.f;
  }
  int kend = p.kv_len ? min(abs(p.kv_len[b]), p.Lk) : p.Lk;
  bool key_masked = !row_ok || j >= kend || (p.pad && p.pad[(long long)b * p.Lk + jc]);
  int istart = p.causal ? max(0, (int)(blockIdx.x * ROWS) - p.q_pos0) / KT * KT : 0;     // queries i >= first key of the block
  const T* qb = (const T*)p.q + (long long)b * p.Lq * p.ldq + h * DH;
  const T* gb = (const T*)p.dout + (long long)b * p.Lq * p.lddo + h * DH;
  long long rowbase = ((long long)b * p.H + h) * p.Lq;
  for (int i0 = istart; i0 < p.Lq; i0 += KT) {
    __syncthreads();
    for (int e = threadIdx.x; e < KT * DH; e += NTHREADS) {
      int ii = e / DH, c = e % DH;
      int i = i0 + ii;
      float qv = 0.f, gv = 0.f;
      if (i < p.Lq) {
        qv = to_f32(qb[(long long)i * p.ldq + c]) * p.scale;
        gv = to_f32(gb[(long long)i * p.lddo + c]);
      }
      Qs[ii][c] = qv;
      Gs[ii][c] = gv;
    }
    if (threadIdx.x < KT) {
      int i = i0 + threadIdx.x;
      Ls[threadIdx.x] = i < p.Lq ? p.lse[rowbase + i] : INFINITY;    // exp(s - inf) = 0
      Ds[threadIdx.x] = i < p.Lq ? p.dsum[rowbase + i] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ii = 0; ii < KT; ++ii) {
      int i = i0 + ii;
      float s = row_dot<DPT, TPR>(kr, &Qs[ii][part * DPT]);
      if (p.addmask && i < p.Lq && row_ok) s += p.addmask[(long long)i * p.ldmask + j];
      bool masked = key_masked || (p.causal && j > i + p.q_pos0) || i >= p.Lq;
      float pr = masked ? 0.f : expf(s - Ls[ii]);
      float dmv = 1.f;
      if (p.thr)
        dmv = attn_keep(attn_row_key(attn_site_key(eff_seed(p.seed, p.seed_dev), p.site), rowbase + min(i, p.Lq - 1)), jc, p.thr) ? p.inv_keep : 0.f;
      float dp = row_dot<DPT, TPR>(vr, &Gs[ii][part * DPT]) * dmv;
      float ds = pr * (dp - Ds[ii]);
      float pd = pr * dmv;
#pragma unroll
      for (int c = 0; c < DPT; c += 4) {
        float4 tq = *reinterpret_cast<const float4*>(&Qs[ii][part * DPT + c]);
        float4 tg = *reinterpret_cast<const float4*>(&Gs[ii][part * DPT + c]);
        dk[c] = fmaf(ds, tq.x, dk[c]);
        dk[c + 1] = fmaf(ds, tq.y, dk[c + 1]);
        dk[c + 2] = fmaf(ds, tq.z, dk[c + 2]);
        dk[c + 3] = fmaf(ds, tq.w, dk[c + 3]);
        dv[c] = fmaf(pd, tg.x, dv[c]);
        dv[c + 1] = fmaf(pd, tg.y, dv[c + 1]);
        dv[c + 2] = fmaf(pd, tg.z, dv[c + 2]);
        dv[c + 3] = fmaf(pd, tg.w, dv[c + 3]);
      }
    }
  }
  if (row_ok) {
    T* ok = (T*)p.dk + ((long long)b * p.Lk + j) * p.lddk + h * DH + part * DPT;
    T* ov = (T*)p.dv + ((long long)b * p.Lk + j) * p.lddv + h * DH + part * DPT;
#pragma unroll
    for (int c = 0; c < DPT; ++c) {
      ok[c] = from_f32<T>(dk[c]);       // Qs already carries the 1/sqrt(dh) scale
      ov[c] = from_f32<T>(dv[c]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// head-averaged probabilities (the second return value of the reference's decoder layers,
// transformer.py:463 / model.py:101): w[b,i,j] = mean_h dropout(softmax(s))[b,h,i,j].  Needs lse.
// ---------------------------------------------------------------------------------------
template <typename T, int DH>
__global__ void __launch_bounds__(NTHREADS) attn_weights_kernel(AttnParams p, float* __restrict__ w, long long ldw) {
  int b = blockIdx.z;
  int i = blockIdx.y;
  int kend = p.kv_len ? min(abs(p.kv_len[b]), p.Lk) : p.Lk;
  for (int j = blockIdx.x * NTHREADS + threadIdx.x; j < p.Lk; j += gridDim.x * NTHREADS) {
    bool masked = j >= kend || (p.pad && p.pad[(long long)b * p.Lk + j]) || (p.causal && j > i + p.q_pos0);
    float acc = 0.f;
    if (!masked) {
      for (int h = 0; h < p.H; ++h) {
        const T* q = (const T*)p.q + ((long long)b * p.Lq + i) * p.ldq + h * DH;
        const T* k = (const T*)p.k + ((long long)b * p.Lk + j) * p.ldk + h * DH;
        float s = 0.f;
#pragma unroll 8
        for (int c = 0; c < DH; ++c) s = fmaf(to_f32(q[c]) * p.scale, to_f32(k[c]), s);
        if (p.addmask) s += p.addmask[(long long)i * p.ldmask + j];
        long long rowid = ((long long)b * p.H + h) * p.Lq + i;
        float pr = expf(s - p.lse[rowid]);
        if (p.thr) pr = attn_keep(attn_row_key(attn_site_key(eff_seed(p.seed, p.seed_dev), p.site), rowid), j, p.thr) ? pr * p.inv_keep : 0.f;
        acc += pr;
      }
    }
    w[((long long)b * p.Lq + i) * ldw + j] = acc / p.H;
  }
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
static int fill_params(AttnParams& p, const smer_attn_args* a, const char* who) {
  if (!a) { smer_set_error("%s: null args", who); return SMER_ERR_ARG; }
  memset(&p, 0, sizeof(p));
  p.q = a->q; p.k = a->k; p.v = a->v; p.o = a->o; p.out = a->o; p.dout = a->dout;
  p.dq = a->dq; p.dk = a->dk; p.dv = a->dv;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo; p.lddo = a->lddo;
  p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
  p.lse = a->lse; p.dsum = a->dsum; p.pad = a->key_pad; p.kv_len = a->kv_len;
  p.addmask = a->add_mask; p.ldmask = a->ld_mask;
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk;
  p.scale = a->scale; p.causal = a->causal; p.q_pos0 = a->q_pos0;
  p.thr = a->dropout_p > 0.f ? attn_dropout_threshold(a->dropout_p) : 0u;  // fp16 pattern rule of common.cuh
  p.inv_keep = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f;
  p.seed = a->seed; p.site = a->site; p.seed_dev = smer_seed_dev();
  if (p.B <= 0 || p.H <= 0 || p.Lq <= 0 || p.Lk <= 0) { smer_set_error("%s: empty problem", who); return SMER_ERR_ARG; }
  return SMER_OK;
}

#define DISPATCH_DH(T, DHV, FN, ...)                                         \
  switch (DHV) {                                                             \
    case 16: FN<T, 16, 1> __VA_ARGS__; break;                                \
    case 32: FN<T, 32, 1> __VA_ARGS__; break;                                \
    case 64: FN<T, 64, 2> __VA_ARGS__; break;                                \
    default:                                                                 \
      smer_set_error("attention (simt): head dim %d unsupported (16/32/64)", DHV); \
      return SMER_ERR_UNSUPPORTED;                                           \
  }

extern "C" int smer_attn_fwd_simt(const smer_attn_args* a, void* stream) {
  AttnParams p;
  int rc = fill_params(p, a, "smer_attn_fwd_simt");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  int tpr = a->dh == 64 ? 2 : 1;
  dim3 grid((p.Lq + NTHREADS / tpr - 1) / (NTHREADS / tpr), p.H, p.B);
  if (a->dtype == SMER_DT_F32) { DISPATCH_DH(float, a->dh, attn_fwd_simt_kernel, <<<grid, NTHREADS, 0, st>>>(p)) }
  else { DISPATCH_DH(bf16, a->dh, attn_fwd_simt_kernel, <<<grid, NTHREADS, 0, st>>>(p)) }
  SMER_CHECK_LAUNCH("smer_attn_fwd_simt");
  return SMER_OK;
}

template <typename T>
static int bwd_launch(AttnParams& p, int dh, cudaStream_t st) {
  long long n = (long long)p.B * p.H * p.Lq;
  int g = (int)((n + 255) / 256);
  switch (dh) {
    case 16: attn_bwd_dot_kernel<T, 16><<<g, 256, 0, st>>>(p); break;
    case 32: attn_bwd_dot_kernel<T, 32><<<g, 256, 0, st>>>(p); break;
    case 64: attn_bwd_dot_kernel<T, 64><<<g, 256, 0, st>>>(p); break;
    default: smer_set_error("attention (simt): head dim %d unsupported", dh); return SMER_ERR_UNSUPPORTED;
  }
  int tpr = dh == 64 ? 2 : 1;
  int rows = NTHREADS / tpr;
  dim3 gq((p.Lq + rows - 1) / rows, p.H, p.B), gk((p.Lk + rows - 1) / rows, p.H, p.B);
  DISPATCH_DH(T, dh, attn_bwd_dq_kernel, <<<gq, NTHREADS, 0, st>>>(p))
  DISPATCH_DH(T, dh, attn_bwd_dkv_kernel, <<<gk, NTHREADS, 0, st>>>(p))
  return SMER_OK;
}

extern "C" int smer_attn_bwd_simt(const smer_attn_args* a, void* stream) {
  AttnParams p;
  int rc = fill_params(p, a, "smer_attn_bwd_simt");
  if (rc) return rc;
  SMER_CHECK_ARG(p.dout && p.dq && p.dk && p.dv && p.lse && p.dsum, "smer_attn_bwd_simt: missing gradient buffers");
  cudaStream_t st = (cudaStream_t)stream;
  rc = a->dtype == SMER_DT_F32 ? bwd_launch<float>(p, a->dh, st) : bwd_launch<bf16>(p, a->dh, st);
  if (rc) return rc;
  SMER_CHECK_LAUNCH("smer_attn_bwd_simt");
  return SMER_OK;
}

extern "C" int smer_attn_weights(const smer_attn_args* a, float* weights, long long ldw, void* stream) {
  AttnParams p;
  int rc = fill_params(p, a, "smer_attn_weights");
  if (rc) return rc;
  SMER_CHECK_ARG(p.lse && weights, "smer_attn_weights: lse and output required");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((p.Lk + NTHREADS - 1) / NTHREADS, p.Lq, p.B);
#define W_CASE(T)                                                                      \
  switch (a->dh) {                                                                     \
    case 16: attn_weights_kernel<T, 16><<<grid, NTHREADS, 0, st>>>(p, weights, ldw); break; \
    case 32: attn_weights_kernel<T, 32><<<grid, NTHREADS, 0, st>>>(p, weights, ldw); break; \
    case 64: attn_weights_kernel<T, 64><<<grid, NTHREADS, 0, st>>>(p, weights, ldw); break; \
    default: smer_set_error("smer_attn_weights: head dim %d unsupported", a->dh); return SMER_ERR_UNSUPPORTED; \
  }
  if (a->dtype == SMER_DT_F32) { W_CASE(float) } else { W_CASE(bf16) }
#undef W_CASE
  SMER_CHECK_LAUNCH("smer_attn_weights");
  return SMER_OK;
}
