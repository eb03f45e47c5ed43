// Attention forward, second generation (bf16, dh = 64): the FlashAttention-4 work split on tcgen05 / TMEM / TMA.
// Arithmetic: softmax(q k^T / sqrt(dh) + masks) v with dropout on P, as F.multi_head_attention_forward's
// need_weights branch computes it for the reference (transformer.py:389,459,463).
//
// One persistent CTA per SM walks (batch, head, PAIR of 128-query tiles) work items.  Four independent softmax
// CHAINS run per CTA: chain q = (query tile c = q / 2, key half hf = q % 2) owns the 64-key half `hf` of every
// 128-key K/V tile for query tile `c`, with its own running maximum, row sum and output accumulator (split-KV, as in
// flash decoding); the two halves of a query tile are merged once, at the end of the item.  No per-tile exchange
// between threads, and sixteen softmax warps (four per scheduler) to hide TMEM / MUFU / barrier latency.
//   warps 0..15  softmax, warpgroup q = chain q, one thread per query row (= TMEM lane): the row's 64 scores of a tile
//                are read once into registers; row maximum; exp2 / row sum / dropout / bf16 pack; P goes back to
//                TENSOR MEMORY over the first 32 columns of S_q.  The reference maximum is LAZY: O_q and the row sum
//                are rescaled only when the maximum grew by more than 2^8, by the row's own thread between
//                "S ready" and "P ready", when no MMA touches O_q.
//   warp 16      TMA producer: the pair's two Q tiles once, then K_t / V_t through 3-stage rings (each K / V tile is
//                loaded once and serves all four chains)
//   warp 17      tcgen05.mma issuer + TMEM owner.  TMEM (512 columns): S_q at 64 q (fp32), O_q at 256 + 64 q.
//                S_q = Q_c K_hf^T (SS), O_q += P_q V_hf with the A operand P_q read from tensor memory (TS-MMA).
//   warps 18,19  idle: roles are warpgroup-aligned so that setmaxnreg can move the producer warpgroup's registers
//                to the softmax warpgroups.
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "../../include/smer_b200.h"

int smer_make_tmap_bf16(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long pitch,
                        int box_inner, int box_outer);

namespace {

constexpr int BM = 128, BN = 128, BH = 64, DH = 64;      // BH: keys per chain and tile
constexpr int TILE = BM * DH * 2;                 // 16 KB: one Q / K / V tile
constexpr int KS = 3;                             // K and V ring depth
constexpr int NCHAIN = 4;
constexpr int THREADS = 640;                      // 4 softmax warpgroups + 1 producer warpgroup (TMA warp, MMA warp, 2 idle)
constexpr int MASK_WORDS = 512;                   // key-mask bitmap of one batch row: Lk <= 16384
constexpr int OFF_K = 2 * TILE, OFF_V = OFF_K + KS * TILE, OFF_BAR = OFF_V + KS * TILE;
constexpr int OFF_MASK = OFF_BAR + 256;
constexpr int OFF_ML = OFF_MASK + 2 * MASK_WORDS * 4;                // [2 query tiles][2 halves][128][2] fp32: (m, l) of every chain row
constexpr int SMEM_BYTES = OFF_ML + 2 * 2 * BM * 2 * 4 + 1024 /*alignment slack*/;
constexpr int TMEM_COLS = 512;
constexpr uint32_t COL_S = 0, COL_O = 256;        // S_q at COL_S + 64 q, O_q at COL_O + 64 q, P_q over S_q's first 32 columns
constexpr float RESCALE_LOG2 = 8.f;               // lazy rescale: only when the maximum grew by more than 2^8

struct Params {
  bf16* o;
  long long ldo;
  float* lse;
  const int* kv_len;
  const uint8_t* pad;
  int B, H, Lq, Lk;
  float c_log2;            // scale * log2(e)
  int causal;
  uint32_t thr2;           // dropout threshold pattern in both halves (common.cuh: attn_dropout_threshold); 0 = dropout off
  float inv_keep;
  uint64_t seed, site;
  const unsigned long long* seed_dev;
  int npair, items;        // query-tile pairs per (b,h); npair * H * B work items
  const int *cu_q, *cu_k;  // padding-free layout (smer_b200.h): per-sequence row ranges of the packed buffers, or NULL
  long long q_rows;        // rows of the packed Q (and O) buffers
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t lds_u(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
// the 256 threads of one query tile (its two key-half chains)
__device__ __forceinline__ void bar_sync_qtile(int c) { asm volatile("bar.sync %0, 256;" ::"r"(c + 1) : "memory"); }

struct Item {
  int b, h, i0;            // first query row of the pair
  int nt0, nt1;            // 128-key K/V tiles of query tile 0 / 1 (0: tile inactive)
  int kend0, kend1;
  int ntk;                 // K / V tiles to load = max
  int klen;                // kv_len[b] clipped to Lk (before the causal bound)
  bool holes;              // the key mask is not a pure suffix: key_pad has to be read
  int q_row0, k_row0;      // first row of sequence b in the Q / K,V buffers
  int lq;                  // query rows of sequence b
  long long stat0;         // index of (b, h, row 0) in lse (and the dropout row id)
};

__device__ __forceinline__ Item item_of(const Params& p, int item) {
  Item it;
  const int HB = p.H * p.B;
  // causal: all (b,h) of the heaviest (last) query pair first; full: the pairs of one (b,h) next to each other
  const int qp = p.causal ? p.npair - 1 - item / HB : item % p.npair;
  const int rem = p.causal ? item % HB : item / p.npair;
  it.h = rem % p.H;
  it.b = rem / p.H;
  it.i0 = qp * 2 * BM;
  int kraw;
  if (p.cu_q) {                                    // packed rows: every key of the sequence is visible
    it.q_row0 = p.cu_q[it.b];
    it.lq = p.cu_q[it.b + 1] - it.q_row0;
    it.k_row0 = p.cu_k[it.b];
    kraw = p.cu_k[it.b + 1] - it.k_row0;
    it.stat0 = (long long)it.h * p.cu_q[p.B] + it.q_row0;
  } else {
    it.q_row0 = it.b * p.Lq;
    it.lq = p.Lq;
    it.k_row0 = it.b * p.Lk;
    kraw = p.kv_len ? p.kv_len[it.b] : p.Lk;
    it.stat0 = ((long long)it.b * p.H + it.h) * p.Lq;
  }
  const int kl = min(abs(kraw), p.Lk);
  it.klen = kl;
  it.holes = p.pad != nullptr && (kraw < 0 || p.kv_len == nullptr);
  int ke0 = it.i0 < it.lq ? kl : 0, ke1 = it.i0 + BM < it.lq ? kl : 0;
  if (p.causal) { ke0 = min(ke0, it.i0 + BM); ke1 = min(ke1, it.i0 + 2 * BM); }
  it.kend0 = ke0; it.kend1 = ke1;
  it.nt0 = (ke0 + BN - 1) / BN; it.nt1 = (ke1 + BN - 1) / BN;
  it.ntk = max(it.nt0, it.nt1);
  return it;
}

template <bool DROP>          // dropout compiled in or out: a run-time test per 8 keys would cut the softmax loop into basic blocks
__global__ void __launch_bounds__(THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [2] x 16 KB
  uint8_t* sK = smem + OFF_K;                      // [KS] x 16 KB
  uint8_t* sV = smem + OFF_V;                      // [KS] x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t *q_full = bars, *q_empty = bars + 1, *k_full = bars + 2 /*[KS]*/, *k_empty = bars + 2 + KS /*[KS]*/,
           *v_full = bars + 2 + 2 * KS, *v_empty = bars + 2 + 3 * KS, *s_full = bars + 2 + 4 * KS /*[4]*/,
           *p_full = s_full + NCHAIN /*[4]*/, *o_full = s_full + 2 * NCHAIN /*[4]*/;
  constexpr int NBARS = 2 + 4 * KS + 3 * NCHAIN;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NBARS);
  static_assert((NBARS + 1) * 8 <= 256, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 16 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    for (int i = 0; i < NBARS; ++i) ptx::mbar_init(bars + i, (bars + i >= p_full && bars + i < p_full + NCHAIN) ? 4 : 1);
    ptx::fence_barrier_init();
  }
  if (warp == 17) ptx::tmem_alloc<TMEM_COLS>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= 16) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 16) {
    // ------------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      uint32_t g = 0, qi = 0;                       // K/V tiles and items (with work) so far: ring slots and phases
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const Item it = item_of(p, item);
        if (it.ntk == 0) continue;
        if (qi > 0) ptx::mbar_wait(q_empty, (qi - 1) & 1);          // the previous item's S MMAs have read sQ
        ptx::mbar_expect_tx(q_full, it.nt1 > 0 ? 2 * TILE : TILE);
        ptx::tma_load_2d(sQ, &tmQ, q_full, it.h * DH, it.q_row0 + it.i0);
        if (it.nt1 > 0) ptx::tma_load_2d(sQ + TILE, &tmQ, q_full, it.h * DH, it.q_row0 + it.i0 + BM);
        for (int t = 0; t < it.ntk; ++t, ++g) {
          const uint32_t st = g % KS, ph = ((g / KS) - 1) & 1;
          if (g >= KS) ptx::mbar_wait(k_empty + st, ph);
          ptx::mbar_expect_tx(k_full + st, TILE);
          ptx::tma_load_2d(sK + st * TILE, &tmK, k_full + st, it.h * DH, it.k_row0 + t * BN);
          if (g >= KS) ptx::mbar_wait(v_empty + st, ph);
          ptx::mbar_expect_tx(v_full + st, TILE);
          ptx::tma_load_2d(sV + st * TILE, &tmV, v_full + st, it.h * DH, it.k_row0 + t * BN);
        }
        ++qi;
      }
    }
    __syncwarp();
  } else if (warp == 17) {
    // ------------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BM, BH, 0, 0);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(BM, DH, 0, 1);
      const uint32_t aQ = ptx::smem_u32(sQ), aK = ptx::smem_u32(sK), aV = ptx::smem_u32(sV);
      uint32_t g = 0, qi = 0, np = 0;               // np: P tiles consumed so far PER CHAIN (chains of a query tile advance together; see below)
      uint32_t npq[NCHAIN] = {0u, 0u, 0u, 0u};
      auto issue_s = [&](int q, uint32_t st) {       // S_q = Q_c K_hf^T   (64 keys: rows hf*64.. of the K tile)
        const int c = q >> 1, hf = q & 1;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          ptx::umma_bf16_ss(tmem_base + COL_S + q * 64, ptx::make_smem_desc(aQ + c * TILE + k * 32, 16, 1024),
                            ptx::make_smem_desc(aK + st * TILE + hf * (TILE / 2) + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        ptx::umma_commit(s_full + q);
      };
      auto issue_pv = [&](int q, uint32_t st, bool acc) {   // O_q (+)= P_q V_hf, P_q (128 x 64 keys, bf16) from tensor memory
        const int hf = q & 1;
#pragma unroll
        for (int k = 0; k < BH / 16; ++k)
          ptx::umma_bf16_ts(tmem_base + COL_O + q * 64, tmem_base + COL_S + q * 64 + k * 8,
                            ptx::make_smem_desc(aV + st * TILE + hf * (TILE / 2) + k * 2048, 8192, 1024), idesc_o,
                            (acc || k > 0) ? 1u : 0u);
      };
      (void)np;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const Item it = item_of(p, item);
        if (it.ntk == 0) continue;
        ptx::mbar_wait(q_full, qi & 1);
        {
          const uint32_t st = g % KS;
          ptx::mbar_wait(k_full + st, (g / KS) & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int q = 0; q < NCHAIN; ++q)
            if ((q < 2 ? it.nt0 : it.nt1) > 0) issue_s(q, st);
          ptx::umma_commit(k_empty + st);
          if (it.ntk == 1) ptx::umma_commit(q_empty);
        }
        for (int t = 0; t < it.ntk; ++t) {
          const uint32_t gc = g + t, st = gc % KS;
          const bool more = t + 1 < it.ntk;
          const uint32_t stn = (gc + 1) % KS;
          if (more) ptx::mbar_wait(k_full + stn, ((gc + 1) / KS) & 1);
          ptx::mbar_wait(v_full + st, (gc / KS) & 1);
#pragma unroll
          for (int q = 0; q < NCHAIN; ++q) {
            const int nt = q < 2 ? it.nt0 : it.nt1;
            if (t < nt) {
              ptx::mbar_wait(p_full + q, npq[q] & 1);
              ++npq[q];
              ptx::tc_fence_after();
              issue_pv(q, st, t > 0);
              if (t == nt - 1) ptx::umma_commit(o_full + q);
            }
            if (q == NCHAIN - 1) ptx::umma_commit(v_empty + st);
            if (more && t + 1 < nt) issue_s(q, stn);              // S_q's columns are free: PV_q(t) was issued before
          }
          if (more) {
            ptx::umma_commit(k_empty + stn);
            if (t + 2 == it.ntk) ptx::umma_commit(q_empty);         // last S of the item: sQ may be refilled
          }
        }
        g += it.ntk;
        ++qi;
      }
    }
    __syncwarp();
  } else if (p.cu_q && blockIdx.x == 0) {
    // warps 18 / 19 of the first CTA: the ghost rows past the last sequence of a packed batch are written by no item; the
    // projections that follow read every row, so they must hold finite values (zeros)
    const int first = p.cu_q[p.B];
    const int cols8 = p.H * DH / 8;
    for (long long i = (long long)first * cols8 + (warp - 18) * 32 + lane; i < p.q_rows * cols8; i += 64) {
      const long long row = i / cols8;
      *reinterpret_cast<uint4*>(p.o + row * p.ldo + (i - row * cols8) * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  } else {
    // ------------------------------------------------------------------ softmax: chain q, one thread per query row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int q = warp >> 2, c = q >> 1, hf = q & 1;
    const int quarter = warp & 3;                   // TMEM lanes this warp may touch: 32 * (warp % 4)
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t tS = lane_base + COL_S + q * 64, tO = lane_base + COL_O + q * 64;
    const uint32_t a_mask = ptx::smem_u32(smem + OFF_MASK) + c * MASK_WORDS * 4;
    const uint32_t a_ml_mine = ptx::smem_u32(smem + OFF_ML) + ((c * 2 + hf) * BM + r) * 8;
    const uint32_t a_ml_peer = ptx::smem_u32(smem + OFF_ML) + ((c * 2 + (hf ^ 1)) * BM + r) * 8;
    const float c2 = p.c_log2;
    const uint32_t sitekey = DROP ? attn_site_key(eff_seed(p.seed, p.seed_dev), p.site) : 0u;
    uint32_t ns = 0, no = 0;                        // S tiles / items of this chain so far (barrier phases)
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const Item it = item_of(p, item);
      const int nt = c ? it.nt1 : it.nt0, kend = c ? it.kend1 : it.kend0;
      const int r0 = it.i0 + c * BM;
      const int i = r0 + r;
      const bool row_ok = i < it.lq;
      const int ii = row_ok ? i : it.lq - 1;
      const long long rowid = it.stat0 + ii;
      const uint32_t rowkey = DROP ? attn_row_key(sitekey, rowid) : 0u;
      // Masked keys: a pure suffix (the reference's padding, and every kv_len bound) is handled arithmetically below;
      // only a mask with holes needs the per-key bytes: its bitmap (bit j: key j masked) is built once per item by the
      // query tile's 256 threads.
      const bool use_bitmap = nt > 0 && it.holes;
      const bool tail_mask = nt > 0 && (kend & (BN - 1)) != 0 && kend == it.klen;     // the last tile crosses kv_len
      bar_sync_qtile(c);                             // every thread of the query tile has left the previous item (O, bitmap, m/l)
      if (use_bitmap) {
        for (int j = (hf * 4 + quarter) * 32 + lane; j < nt * BN; j += 256) {
          const bool msk = j >= kend || p.pad[(long long)it.b * p.Lk + j];      // (never with packed rows: no key_pad there)
          const uint32_t bal = __ballot_sync(0xffffffffu, msk);
          if (lane == 0) sts_u(a_mask + (j >> 5) * 4, bal);
        }
        bar_sync_qtile(c);                           // publishes the bitmap
      }
      float m = -INFINITY, l = 0.f;                  // reference maximum (log2 units, may lag) and row sum of this key half
      for (int t = 0; t < nt; ++t, ++ns) {
        const int j0 = t * BN + hf * BH;             // first key of this chain's half of the tile
        uint32_t mw0 = 0u, mw1 = 0u;
        if (use_bitmap) {
          mw0 = lds_u(a_mask + ((j0 >> 5)) * 4);
          mw1 = lds_u(a_mask + ((j0 >> 5) + 1) * 4);
        } else if (tail_mask && j0 + BH > kend) {
          const int nv0 = kend - j0, nv1 = kend - (j0 + 32);
          mw0 = nv0 <= 0 ? 0xffffffffu : (nv0 >= 32 ? 0u : (0xffffffffu << nv0));
          mw1 = nv1 <= 0 ? 0xffffffffu : (nv1 >= 32 ? 0u : (0xffffffffu << nv1));
        }
        if (p.causal && j0 + BH - 1 > r0) {
          const int nv0 = i - j0 + 1, nv1 = i - (j0 + 32) + 1;
          mw0 |= nv0 <= 0 ? 0xffffffffu : (nv0 >= 32 ? 0u : (0xffffffffu << nv0));
          mw1 |= nv1 <= 0 ? 0xffffffffu : (nv1 >= 32 ? 0u : (0xffffffffu << nv1));
        }
        const bool masked = __any_sync(0xffffffffu, (mw0 | mw1) != 0u);
        ptx::mbar_wait(s_full + q, ns & 1);
        ptx::tc_fence_after();
        // ---- the row's 64 scores, read once
        uint32_t s[64];
        {
          uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
          uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
          ptx::tmem_ld_32x32(tS, s0);
          ptx::tmem_ld_32x32(tS + 32, s1);
          ptx::tmem_ld_wait();
        }
        if (masked) {
#pragma unroll
          for (int k = 0; k < 64; ++k)
            if (((k < 32 ? mw0 : mw1) >> (k & 31)) & 1u) s[k] = 0xff800000u;          // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 64; k += 8) {
          mx0 = max3(mx0, __uint_as_float(s[k]), __uint_as_float(s[k + 1]));
          mx1 = max3(mx1, __uint_as_float(s[k + 2]), __uint_as_float(s[k + 3]));
          mx2 = max3(mx2, __uint_as_float(s[k + 4]), __uint_as_float(s[k + 5]));
          mx3 = max3(mx3, __uint_as_float(s[k + 6]), __uint_as_float(s[k + 7]));
        }
        const float mxs = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c2;      // c2 > 0
        // ---- lazy maximum: rescale O_q and the row sum only when the maximum grew by more than 2^8 (or is new)
        if (__any_sync(0xffffffffu, mxs > m + RESCALE_LOG2 || (m == -INFINITY && mxs > -INFINITY))) {
          const float m_new = fmaxf(m, mxs);
          const float alpha = m_new == -INFINITY ? 1.f : ex2(m - m_new);   // m = -inf: alpha = 0, nothing accumulated yet
          l *= alpha;
          m = m_new;
          if (t > 0) {                               // O_q holds tiles 0..t-1 and no MMA touches it now
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint32_t v[16];
              ptx::tmem_ld_32x16(tO + q4 * 16, v);
              ptx::tmem_ld_wait(v);
#pragma unroll
              for (int k = 0; k < 16; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) * alpha);
              ptx::tmem_st_32x16(tO + q4 * 16, v);
            }
          }
        }
        const float m_use = m == -INFINITY ? 0.f : m;
        const f32x2 c2p = pack2(c2, c2), nm2 = pack2(-m_use, -m_use);
        f32x2 ls0 = pack2(0.f, 0.f), ls1 = pack2(0.f, 0.f);
        // ---- P = exp2(S c - m), row sum, dropout (P stays unscaled: 1/(1-p) is applied once per row at the end), bf16
        uint32_t pk[32];
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          uint32_t x[8];
          if (DROP) attn_pair_words(attn_block_word(rowkey, (uint32_t)(j0 >> 4) + c16), x);     // one word per two keys
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int e = c16 * 16 + 2 * k;
            float e0, e1;
            unpack2(fma2(pack2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), c2p, nm2), e0, e1);
            const float p0 = ex2(e0), p1 = ex2(e1);
            if (k & 1) ls1 = add2(ls1, pack2(p0, p1));
            else ls0 = add2(ls0, pack2(p0, p1));
            uint32_t pb = pack_bf16x2(p0, p1);
            if (DROP) pb &= attn_keep_mask2(x[k], p.thr2);          // both keys of the pair with one packed compare
            pk[e >> 1] = pb;
          }
        }
        ptx::tmem_st_32x32(tS, pk);
        ptx::tmem_st_wait();
        {
          float a0, a1, b0, b1;
          unpack2(ls0, a0, a1);
          unpack2(ls1, b0, b1);
          l += (a0 + a1) + (b0 + b1);
        }
        ptx::tc_fence_before();                      // orders this thread's tcgen05.ld / st before the MMAs that follow the arrive
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_relaxed(p_full + q);     // P went through tensor memory: the tcgen05 fences order it
      }
      // ---- epilogue of the item: merge the two key halves of the row, O / l -> global.  Both halves' accumulators sit
      // in this thread's TMEM lane (O_{2c} and O_{2c+1}), so only (m, l) travel through shared memory; this thread
      // finishes output columns [32 hf, 32 hf + 32).
      sts_f(a_ml_mine, m);
      sts_f(a_ml_mine + 4, l);
      if (nt > 0) {                                  // uniform over the query tile's threads
        ptx::mbar_wait(o_full + 2 * c, no & 1);      // both chains' last PV MMAs have completed
        ptx::mbar_wait(o_full + 2 * c + 1, no & 1);
        ++no;
        ptx::tc_fence_after();
      }
      bar_sync_qtile(c);                             // (m, l) of both halves are published
      if (nt > 0) {
        const uint32_t tOa = lane_base + COL_O + (2 * c) * 64 + hf * 32, tOb = tOa + 64;
        uint32_t va[32], vb[32];
        ptx::tmem_ld_32x32(tOa, va);
        ptx::tmem_ld_32x32(tOb, vb);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        const float mp = lds_f(a_ml_peer), lp = lds_f(a_ml_peer + 4);
        const float ma = hf ? mp : m, mb = hf ? m : mp, la = hf ? lp : l, lb = hf ? l : lp;
        const float mm = fmaxf(ma, mb);
        const float fa = ma == -INFINITY ? 0.f : ex2(ma - mm), fb = mb == -INFINITY ? 0.f : ex2(mb - mm);
        const float lt = la * fa + lb * fb;
        const float inv = lt > 0.f ? p.inv_keep / lt : 0.f;      // dropout's 1/(1-p) applied once per row
        const float wa = fa * inv, wb = fb * inv;
        if (row_ok) {
          bf16* orow = p.o + ((long long)it.q_row0 + i) * p.ldo + it.h * DH + hf * 32;
#pragma unroll
          for (int k = 0; k < 32; k += 8) {
            float f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(va[k + u]) * wa + __uint_as_float(vb[k + u]) * wb;
            uint4 u4;
            u4.x = pack_bf16x2(f[0], f[1]);
            u4.y = pack_bf16x2(f[2], f[3]);
            u4.z = pack_bf16x2(f[4], f[5]);
            u4.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(orow + k) = u4;
          }
          if (p.lse && hf == 0) p.lse[rowid] = lt > 0.f ? (mm + log2f(lt)) * 0.6931471805599453f : -INFINITY;
        }
      } else if (row_ok) {                           // no visible key at all: zeros (uniform branch per query tile)
        bf16* orow = p.o + ((long long)it.q_row0 + i) * p.ldo + it.h * DH + hf * 32;
#pragma unroll
        for (int k = 0; k < 32; k += 8) *reinterpret_cast<uint4*>(orow + k) = make_uint4(0u, 0u, 0u, 0u);
        if (p.lse && hf == 0) p.lse[rowid] = -INFINITY;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 17) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace

// smer_attn_fwd_tc dispatches here (attn_tc.cu); same argument checks have been done there
int smer_attn_fwd2_launch(const smer_attn_args* a, void* stream) {
  int rc;
  CUtensorMap tq, tk, tv;
  const long long dcols = (long long)a->H * DH;
  const long long rq = a->cu_q ? a->q_rows : (long long)a->B * a->Lq, rk = a->cu_q ? a->k_rows : (long long)a->B * a->Lk;
  if ((rc = smer_make_tmap_bf16(&tq, a->q, dcols, rq, a->ldq, DH, BM))) return rc;
  if ((rc = smer_make_tmap_bf16(&tk, a->k, dcols, rk, a->ldk, DH, BN))) return rc;
  if ((rc = smer_make_tmap_bf16(&tv, a->v, dcols, rk, a->ldv, DH, BN))) return rc;
  Params p;
  p.o = (bf16*)a->o; p.ldo = a->ldo; p.lse = a->lse; p.kv_len = a->kv_len; p.pad = a->key_pad;
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk;
  p.c_log2 = a->scale * 1.4426950408889634f;
  p.causal = a->causal;
  p.thr2 = a->dropout_p > 0.f ? attn_dropout_threshold(a->dropout_p) * 0x10001u : 0u;
  p.inv_keep = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f;
  p.seed = a->seed; p.site = a->site; p.seed_dev = smer_seed_dev();
  p.cu_q = a->cu_q; p.cu_k = a->cu_k; p.q_rows = a->q_rows;
  p.npair = (a->Lq + 2 * BM - 1) / (2 * BM);
  const long long items = (long long)p.npair * a->H * a->B;
  SMER_CHECK_ARG(items < (1ll << 31), "smer_attn_fwd_tc: too many work items");
  p.items = (int)items;
  static int attr_dev_mask = 0;                      // cudaFuncSetAttribute is per device
  int dev = 0;
  SMER_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask & (1 << dev))) {
    SMER_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SMER_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_dev_mask |= 1 << dev;
  }
  const long long grid = items < smer_num_sms() ? items : smer_num_sms();
  if (p.thr2) attn_fwd2_kernel<true><<<dim3((unsigned)grid), THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, p);
  else attn_fwd2_kernel<false><<<dim3((unsigned)grid), THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, p);
  SMER_CHECK_LAUNCH("smer_attn_fwd_tc(v2)");
  return SMER_OK;
}
