// Attention forward, second generation (bf16, dh = 64): the FlashAttention-4 work split on tcgen05 / TMEM / TMA.
// Arithmetic: softmax(q k^T / sqrt(dh) + masks) v with dropout on P, as F.multi_head_attention_forward's
// need_weights branch computes it for the reference (transformer.py:389,459,463).
//
// One persistent CTA per SM walks (batch, head, PAIR of 128-query tiles) work items:
//   warp 8       TMA producer: the pair's two Q tiles once, then K_t / V_t through 3-stage rings (each K / V tile is
//                loaded once and serves both query tiles)
//   warp 9       tcgen05.mma issuer + TMEM owner.  TMEM (512 columns): S0 | S1 (128 fp32 columns each) and
//                O0 | O1 (64 each).  Per tile and query tile c: S_c = Q_c K_t^T (SS), then O_c += P_c V_t with the A
//                operand P_c read FROM TENSOR MEMORY (TS-MMA): the softmax threads write P (bf16, two keys per
//                32-bit column) over the first 64 columns of S_c, no shared-memory round trip.  O_c stays resident
//                in TMEM for the whole item.
//   warps 0..3   softmax of query tile 0, one thread per query row (= TMEM lane): the row's 128 scores of a tile are
//   warps 4..7   softmax of query tile 1      read once into registers; row maximum; exp2 / row sum / dropout / bf16
//                pack; P back to TMEM.  The running maximum is LAZY: O_c (and the row sum) are rescaled only when the
//                maximum grew by more than 2^8, done by the row's own thread between "S ready" and "P ready", when
//                no MMA touches O_c.  While one query tile's threads work, the tensor core runs the other tile's MMAs.
//   Roles are warpgroup-aligned so that setmaxnreg can move registers from the producer warpgroup (warps 8..11, two of
//   them idle) to the softmax warpgroups: a row of 128 fp32 scores plus its packed P needs more than the 168 registers
//   a 384-thread CTA starts with.
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "../../include/smer_b200.h"

int smer_make_tmap_bf16(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long pitch,
                        int box_inner, int box_outer);

namespace {

constexpr int BM = 128, BN = 128, DH = 64;
constexpr int TILE = BM * DH * 2;                 // 16 KB: one Q / K / V tile
constexpr int KS = 3;                             // K and V ring depth
constexpr int THREADS = 384;                      // 2 softmax warpgroups + 1 producer warpgroup (TMA warp, MMA warp, 2 idle)
constexpr int MASK_WORDS = 512;                   // key-mask bitmap of one batch row: Lk <= 16384
constexpr int OFF_K = 2 * TILE, OFF_V = OFF_K + KS * TILE, OFF_BAR = OFF_V + KS * TILE;
constexpr int OFF_MASK = OFF_BAR + 256;
constexpr int SMEM_BYTES = OFF_MASK + 2 * MASK_WORDS * 4 + 1024 /*alignment slack*/;
constexpr int TMEM_COLS = 512;
constexpr uint32_t COL_S = 0, COL_O = 256;        // S_c at COL_S + 128 c, O_c at COL_O + 64 c, P_c over S_c's first 64 columns
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
  uint32_t thr;            // dropout threshold p * 2^32 (0 = dropout off)
  float inv_keep;
  uint64_t seed, site;
  const unsigned long long* seed_dev;
  int npair, items;        // query-tile pairs per (b,h); npair * H * B work items
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t lds_u(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void bar_sync_wg(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

struct Item {
  int b, h, i0;            // first query row of the pair
  int nt[2];               // KV tiles of query tile 0 / 1 (0: tile inactive)
  int kend[2];
  int ntk;                 // K / V tiles to load = max
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
  const int kl = p.kv_len ? min(p.kv_len[it.b], p.Lk) : p.Lk;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int r0 = it.i0 + c * BM;
    int ke = r0 < p.Lq ? kl : 0;
    if (p.causal) ke = min(ke, r0 + BM);
    it.kend[c] = ke;
    it.nt[c] = (ke + BN - 1) / BN;
  }
  it.ntk = max(it.nt[0], it.nt[1]);
  return it;
}

template <bool DROP>          // dropout compiled in or out: a run-time test per 8 keys would cut the softmax loop into 16 basic blocks
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
           *v_full = bars + 2 + 2 * KS, *v_empty = bars + 2 + 3 * KS, *s_full = bars + 2 + 4 * KS /*[2]*/,
           *p_full = s_full + 2 /*[2]*/, *o_full = s_full + 4 /*[2]*/;
  constexpr int NBARS = 2 + 4 * KS + 6;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NBARS);
  static_assert((NBARS + 1) * 8 <= 256, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    for (int i = 0; i < NBARS; ++i) ptx::mbar_init(bars + i, (bars + i == p_full || bars + i == p_full + 1) ? 4 : 1);
    ptx::fence_barrier_init();
  }
  if (warp == 9) ptx::tmem_alloc<TMEM_COLS>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= 8) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      uint32_t g = 0, qi = 0;                       // K/V tiles and items (with work) so far: ring slots and phases
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const Item it = item_of(p, item);
        if (it.ntk == 0) continue;
        if (qi > 0) ptx::mbar_wait(q_empty, (qi - 1) & 1);          // the previous item's S MMAs have read sQ
        ptx::mbar_expect_tx(q_full, it.nt[1] > 0 ? 2 * TILE : TILE);
        ptx::tma_load_2d(sQ, &tmQ, q_full, it.h * DH, it.b * p.Lq + it.i0);
        if (it.nt[1] > 0) ptx::tma_load_2d(sQ + TILE, &tmQ, q_full, it.h * DH, it.b * p.Lq + it.i0 + BM);
        for (int t = 0; t < it.ntk; ++t, ++g) {
          const uint32_t st = g % KS, ph = ((g / KS) - 1) & 1;
          if (g >= KS) ptx::mbar_wait(k_empty + st, ph);
          ptx::mbar_expect_tx(k_full + st, TILE);
          ptx::tma_load_2d(sK + st * TILE, &tmK, k_full + st, it.h * DH, it.b * p.Lk + t * BN);
          if (g >= KS) ptx::mbar_wait(v_empty + st, ph);
          ptx::mbar_expect_tx(v_full + st, TILE);
          ptx::tma_load_2d(sV + st * TILE, &tmV, v_full + st, it.h * DH, it.b * p.Lk + t * BN);
        }
        ++qi;
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BM, BN, 0, 0);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(BM, DH, 0, 1);
      const uint32_t aQ = ptx::smem_u32(sQ), aK = ptx::smem_u32(sK), aV = ptx::smem_u32(sV);
      uint32_t g = 0, qi = 0, np[2] = {0u, 0u};      // np[c]: P tiles of chain c consumed so far (p_full phase)
      auto issue_s = [&](int c, uint32_t st) {       // S_c = Q_c K^T
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          ptx::umma_bf16_ss(tmem_base + COL_S + c * 128, ptx::make_smem_desc(aQ + c * TILE + k * 32, 16, 1024),
                            ptx::make_smem_desc(aK + st * TILE + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        ptx::umma_commit(s_full + c);
      };
      auto issue_pv = [&](int c, uint32_t st, bool acc) {   // O_c (+)= P_c V, P_c from tensor memory
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)
          ptx::umma_bf16_ts(tmem_base + COL_O + c * 64, tmem_base + COL_S + c * 128 + k * 8,
                            ptx::make_smem_desc(aV + st * TILE + k * 2048, 8192, 1024), idesc_o, (acc || k > 0) ? 1u : 0u);
      };
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const Item it = item_of(p, item);
        if (it.ntk == 0) continue;
        ptx::mbar_wait(q_full, qi & 1);
        {
          const uint32_t st = g % KS;
          ptx::mbar_wait(k_full + st, (g / KS) & 1);
          ptx::tc_fence_after();
          if (it.nt[0] > 0) issue_s(0, st);
          if (it.nt[1] > 0) issue_s(1, st);
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
          for (int c = 0; c < 2; ++c) {
            if (t < it.nt[c]) {
              ptx::mbar_wait(p_full + c, np[c] & 1);
              ++np[c];
              ptx::tc_fence_after();
              issue_pv(c, st, t > 0);
              if (t == it.nt[c] - 1) ptx::umma_commit(o_full + c);
            }
            if (c == 1) ptx::umma_commit(v_empty + st);
            if (more && t + 1 < it.nt[c]) issue_s(c, stn);        // S_c's columns are free: PV_c(t) was issued before
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
  }
  } else {
    // ------------------------------------------------------------------ softmax: chain c, one thread per query row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int c = warp >> 2;
    const int quarter = warp & 3;                   // TMEM lanes this warp may touch: 32 * (warp % 4)
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t tS = lane_base + COL_S + c * 128, tO = lane_base + COL_O + c * 64;
    const uint32_t a_mask = ptx::smem_u32(smem + OFF_MASK) + c * MASK_WORDS * 4;
    const float c2 = p.c_log2;
    uint32_t ns = 0, no = 0;                        // S tiles / items of this chain so far (barrier phases)
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const Item it = item_of(p, item);
      const int nt = it.nt[c], kend = it.kend[c];
      const int i = it.i0 + c * BM + r;
      const bool row_ok = i < p.Lq;
      const int ii = row_ok ? i : p.Lq - 1;
      const long long rowid = ((long long)it.b * p.H + it.h) * p.Lq + ii;
      const uint32_t rowkey = DROP ? attn_row_key(eff_seed(p.seed, p.seed_dev), p.site, rowid - (ii & 7)) : 0u;
      uint32_t pm = 1u, pa = 0u;                    // row (ii & 7) of the 8x8 dropout block: 8 LCG steps per row
      if (DROP) attn_advance(8 * (ii & 7), pm, pa);
      // key-mask bitmap of this item's tiles (bit j: key j masked), built once per item by the chain's 128 threads
      const bool use_mask = nt > 0 && (p.pad != nullptr || (kend & (BN - 1)) != 0);
      bar_sync_wg(c);                                // every thread of the chain has left the previous item's bitmap
      if (use_mask) {
        for (int j = (warp & 3) * 32 + lane; j < nt * BN; j += 128) {     // (any permutation of the 4 warps)
          const bool msk = j >= kend || (p.pad && p.pad[(long long)it.b * p.Lk + j]);
          const uint32_t bal = __ballot_sync(0xffffffffu, msk);
          if (lane == 0) sts_u(a_mask + (j >> 5) * 4, bal);
        }
      }
      bar_sync_wg(c);                                // publishes the bitmap; separates it from the previous item's reads
      float m = -INFINITY, l = 0.f;                  // reference maximum (log2 units, may lag) and row sum
      for (int t = 0; t < nt; ++t, ++ns) {
        const int j0 = t * BN;
        uint32_t mw[4] = {0u, 0u, 0u, 0u};
        if (use_mask) {
#pragma unroll
          for (int w = 0; w < 4; ++w) mw[w] = lds_u(a_mask + ((j0 >> 5) + w) * 4);
        }
        if (p.causal && j0 + BN - 1 > it.i0 + c * BM) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int nvis = i - (j0 + w * 32) + 1;
            mw[w] |= nvis <= 0 ? 0xffffffffu : (nvis >= 32 ? 0u : (0xffffffffu << nvis));
          }
        }
        const bool masked = __any_sync(0xffffffffu, (mw[0] | mw[1] | mw[2] | mw[3]) != 0u);
        ptx::mbar_wait(s_full + c, ns & 1);
        ptx::tc_fence_after();
        // ---- the row's 128 scores, read once
        uint32_t s[128];
        {
          uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
          uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
          uint32_t (&s2)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[64]);
          uint32_t (&s3)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[96]);
          ptx::tmem_ld_32x32(tS, s0);
          ptx::tmem_ld_32x32(tS + 32, s1);
          ptx::tmem_ld_32x32(tS + 64, s2);
          ptx::tmem_ld_32x32(tS + 96, s3);
          ptx::tmem_ld_wait();
        }
        if (masked) {
#pragma unroll
          for (int k = 0; k < 128; ++k)
            if ((mw[k >> 5] >> (k & 31)) & 1u) s[k] = 0xff800000u;          // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 128; k += 8) {
          mx0 = max3(mx0, __uint_as_float(s[k]), __uint_as_float(s[k + 1]));
          mx1 = max3(mx1, __uint_as_float(s[k + 2]), __uint_as_float(s[k + 3]));
          mx2 = max3(mx2, __uint_as_float(s[k + 4]), __uint_as_float(s[k + 5]));
          mx3 = max3(mx3, __uint_as_float(s[k + 6]), __uint_as_float(s[k + 7]));
        }
        const float mxs = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c2;      // c2 > 0
        // ---- lazy maximum: rescale O_c and the row sum only when the maximum grew by more than 2^8 (or is new)
        if (__any_sync(0xffffffffu, mxs > m + RESCALE_LOG2 || (m == -INFINITY && mxs > -INFINITY))) {
          const float m_new = fmaxf(m, mxs);
          const float alpha = m_new == -INFINITY ? 1.f : ex2(m - m_new);   // m = -inf: alpha = 0, nothing accumulated yet
          l *= alpha;
          m = m_new;
          if (t > 0) {                               // O_c holds tiles 0..t-1 and no MMA touches it now
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
        uint32_t pk[64];
        const uint32_t kbase = rowkey + (uint32_t)(j0 >> 3) * ATTN_GOLD;    // dropout block index of the tile's first key
#pragma unroll
        for (int c8 = 0; c8 < 16; ++c8) {
          float pv[8];
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            float e0, e1;
            unpack2(fma2(pack2(__uint_as_float(s[c8 * 8 + k]), __uint_as_float(s[c8 * 8 + k + 1])), c2p, nm2), e0, e1);
            pv[k] = ex2(e0);
            pv[k + 1] = ex2(e1);
            if (k & 2) ls1 = add2(ls1, pack2(pv[k], pv[k + 1]));
            else ls0 = add2(ls0, pack2(pv[k], pv[k + 1]));
          }
          if (DROP) {                                 // one mixed word per 8 keys, then one multiply-add per key
            uint32_t x[8];
            attn_block8<1>(attn_mix(kbase + (uint32_t)c8 * ATTN_GOLD) * pm + pa, x);
#pragma unroll
            for (int k = 0; k < 8; ++k) pv[k] = x[k] >= p.thr ? pv[k] : 0.f;
          }
#pragma unroll
          for (int k = 0; k < 8; k += 2) pk[c8 * 4 + (k >> 1)] = pack_bf16x2(pv[k], pv[k + 1]);
        }
        {
          uint32_t (&p0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&pk[0]);
          uint32_t (&p1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&pk[32]);
          ptx::tmem_st_32x32(tS, p0);
          ptx::tmem_st_32x32(tS + 32, p1);
          ptx::tmem_st_wait();
        }
        {
          float a0, a1, b0, b1;
          unpack2(ls0, a0, a1);
          unpack2(ls1, b0, b1);
          l += (a0 + a1) + (b0 + b1);
        }
        ptx::tc_fence_before();                      // orders this thread's tcgen05.ld / st before the MMAs that follow the arrive
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(p_full + c);
      }
      // ---- epilogue of the item: O_c / l -> global
      uint32_t v0[32], v1[32];
      if (nt > 0) {                                  // uniform over the chain's threads
        ptx::mbar_wait(o_full + c, no & 1);
        ++no;
        ptx::tc_fence_after();
        ptx::tmem_ld_32x32(tO, v0);
        ptx::tmem_ld_32x32(tO + 32, v1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) v0[k] = v1[k] = 0u;
      }
      if (row_ok) {
        const float inv = l > 0.f ? p.inv_keep / l : 0.f;      // dropout's 1/(1-p) applied once per row
        bf16* orow = p.o + ((long long)it.b * p.Lq + i) * p.ldo + it.h * DH;
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v0[k]) * inv, __uint_as_float(v0[k + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(v0[k + 2]) * inv, __uint_as_float(v0[k + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(v0[k + 4]) * inv, __uint_as_float(v0[k + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(v0[k + 6]) * inv, __uint_as_float(v0[k + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + k) = u;
        }
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v1[k]) * inv, __uint_as_float(v1[k + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(v1[k + 2]) * inv, __uint_as_float(v1[k + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(v1[k + 4]) * inv, __uint_as_float(v1[k + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(v1[k + 6]) * inv, __uint_as_float(v1[k + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + 32 + k) = u;
        }
        if (p.lse) p.lse[rowid] = l > 0.f ? (m + log2f(l)) * 0.6931471805599453f : -INFINITY;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace

// smer_attn_fwd_tc dispatches here (attn_tc.cu); same argument checks have been done there
int smer_attn_fwd2_launch(const smer_attn_args* a, void* stream) {
  int rc;
  CUtensorMap tq, tk, tv;
  const long long dcols = (long long)a->H * DH;
  if ((rc = smer_make_tmap_bf16(&tq, a->q, dcols, (long long)a->B * a->Lq, a->ldq, DH, BM))) return rc;
  if ((rc = smer_make_tmap_bf16(&tk, a->k, dcols, (long long)a->B * a->Lk, a->ldk, DH, BN))) return rc;
  if ((rc = smer_make_tmap_bf16(&tv, a->v, dcols, (long long)a->B * a->Lk, a->ldv, DH, BN))) return rc;
  Params p;
  p.o = (bf16*)a->o; p.ldo = a->ldo; p.lse = a->lse; p.kv_len = a->kv_len; p.pad = a->key_pad;
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk;
  p.c_log2 = a->scale * 1.4426950408889634f;
  p.causal = a->causal;
  p.thr = a->dropout_p > 0.f ? dropout_threshold(a->dropout_p) : 0u;
  p.inv_keep = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f;
  p.seed = a->seed; p.site = a->site; p.seed_dev = smer_seed_dev();
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
  if (p.thr) attn_fwd2_kernel<true><<<dim3((unsigned)grid), THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, p);
  else attn_fwd2_kernel<false><<<dim3((unsigned)grid), THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, p);
  SMER_CHECK_LAUNCH("smer_attn_fwd_tc(v2)");
  return SMER_OK;
}
