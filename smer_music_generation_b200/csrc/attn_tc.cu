// Flash-style attention on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), bf16, dh = 64.
// Arithmetic of F.multi_head_attention_forward's need_weights branch as the reference reaches it
// (transformer.py:389,459,463): softmax(q k^T / sqrt(dh) + masks) v with dropout on P; masks =
// causal flag (the nopeek tgt_mask) and per-key padding.  Scores are never materialised in HBM.
//
// FORWARD.  CTA = 128 query rows of one (batch, head); KV tiles of 128 keys.
//   warp 0      TMA producer: Q once, then K_t / V_t into single smem slots (K's slot is free as
//               soon as S_t = Q K_t^T has been read by the tensor core, V's after O_t = P_t V_t)
//   warp 1      tcgen05.mma issuer (one elected thread) + TMEM owner (256 columns: S 128, O_t 64)
//   warps 2..9  softmax: TWO threads per query row (TMEM lane), each owning 64 of the tile's 128
//               keys and 32 of the 64 output columns -- 8 warps per CTA are what keeps the SM's
//               issue slots busy (ncu: with 4 the kernel was latency-bound at 52 % issue).  Two
//               passes over S in TMEM (max -> exchanged through smem, then exp2 / row-sum / dropout
//               / bf16 pack), P written to smem in the SWIZZLE_128B K-major layout the PV MMA
//               reads; O accumulated in fp32 registers with the online-softmax rescale.
//   Two CTAs share an SM (80 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the
//   other's MMAs.
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "../../include/smer_b200.h"

int smer_make_tmap_bf16(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long pitch,
                        int box_inner, int box_outer);

namespace {

constexpr int BM = 128, BN = 128, DH = 64;
constexpr int TILE_QKV = BM * DH * 2;            // 16 KB
constexpr int TILE_P = BM * BN * 2;              // 32 KB (two 64-key halves of 16 KB)
constexpr int FWD_THREADS = 320;          // TMA warp, MMA warp, 8 softmax warps (2 threads per query row)
constexpr int MASK_WORDS = 512;                    // key-mask bitmap of one batch row: Lk <= 16384
constexpr int FWD_SMEM = 3 * TILE_QKV + TILE_P + 1024 /*align*/ + 2560 /*barriers + row exchange*/ + MASK_WORDS * 4;
constexpr int TMEM_COLS = 256;
constexpr int O_COL = 128;

struct FwdParams {
  bf16* o;
  long long ldo;
  float* lse;
  const int* kv_len;
  const uint8_t* pad;
  int B, H, Lq, Lk;
  float c_log2;            // scale * log2(e)
  int causal;
  uint32_t thr16;          // dropout threshold p * 2^32 (0 = dropout off)
  float inv_keep;
  uint64_t seed, site;
  const unsigned long long* seed_dev;
  int nq, items;           // persistent forward: query tiles per (b,h), nq * H * B work items
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 8 consecutive bf16 of a row (16-byte chunk c16 of a 128-byte SWIZZLE_128B row)
__device__ __forceinline__ void st_row_chunk_fwd(uint32_t row_addr, uint32_t rx, uint32_t c16, const float* v) {
  st_shared_v4(row_addr + ((c16 ^ rx) << 4), pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
               pack_bf16x2(v[6], v[7]));
}
// Explicit shared-state-space accesses for the small exchange / constant arrays: through a generic pointer
// the compiler emits LD.E / ST.E (generic path, long-scoreboard latency) on the per-tile critical path.
__device__ __forceinline__ float lds_f(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds_v4f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void bar_sync_softmax() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_bwd() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <bool PERSIST>      // PERSIST: CTAs walk several work items; otherwise one item per CTA (the loops below run once)
__global__ void __launch_bounds__(FWD_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_QKV;
  uint8_t* sV = smem + 2 * TILE_QKV;
  uint8_t* sP = smem + 3 * TILE_QKV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * TILE_QKV + TILE_P);
  uint64_t *q_full = bars, *k_full = bars + 1, *v_full = bars + 2, *k_empty = bars + 3, *v_empty = bars + 4,
           *s_full = bars + 5, *p_full = bars + 6, *o_full = bars + 7, *q_empty = bars + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);
  const uint32_t a_xch = ptx::smem_u32(bars) + 256;       // float [2 parity][2 halves][128 rows]
  const uint32_t a_mask = ptx::smem_u32(bars) + 2560;     // uint32 [MASK_WORDS], bit j: key j masked

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Persistent CTA: work items (query tile, head, batch row), heaviest (late-causal) query tiles first, dealt
  // round-robin to the resident CTAs.  Barrier phases run on across items: `gt` counts KV tiles, `qi` counts
  // items that had any, identically in every role.
  const int HB = p.H * p.B;
  auto item_of = [&](int item, int& b, int& h, int& i0, int& kend, int& ntiles) {
    // causal: all (b,h) of the heaviest query tile first; full: the query tiles of one (b,h) next to each other
    // (they share K/V through L2 while they run together)
    const int qt = p.causal ? p.nq - 1 - item / HB : item % p.nq;
    const int rem = p.causal ? item % HB : item / p.nq;
    h = rem % p.H;
    b = rem / p.H;
    i0 = qt * BM;
    kend = p.kv_len ? min(p.kv_len[b], p.Lk) : p.Lk;
    if (p.causal) kend = min(kend, i0 + BM);
    ntiles = (kend + BN - 1) / BN;
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    for (int i = 0; i < 9; ++i) ptx::mbar_init(bars + i, i == 6 ? 256 : 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (ptx::elect_one()) {
      uint32_t gt = 0, qi = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int b, h, i0, kend, ntiles;
        item_of(item, b, h, i0, kend, ntiles);
        if (ntiles == 0) continue;
        if (qi > 0) ptx::mbar_wait(q_empty, (qi - 1) & 1);        // the previous item's S MMAs have read sQ
        ptx::mbar_expect_tx(q_full, TILE_QKV);
        ptx::tma_load_2d(sQ, &tmQ, q_full, h * DH, b * p.Lq + i0);
        for (int t = 0; t < ntiles; ++t, ++gt) {
          if (gt > 0) ptx::mbar_wait(k_empty, (gt - 1) & 1);
          ptx::mbar_expect_tx(k_full, TILE_QKV);
          ptx::tma_load_2d(sK, &tmK, k_full, h * DH, b * p.Lk + t * BN);
          if (gt > 0) ptx::mbar_wait(v_empty, (gt - 1) & 1);
          ptx::mbar_expect_tx(v_full, TILE_QKV);
          ptx::tma_load_2d(sV, &tmV, v_full, h * DH, b * p.Lk + t * BN);
        }
        ++qi;
        if (!PERSIST) break;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BM, BN, 0, 0);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(BM, DH, 0, 1);
      const uint32_t aQ = ptx::smem_u32(sQ), aK = ptx::smem_u32(sK), aV = ptx::smem_u32(sV), aP = ptx::smem_u32(sP);
      uint32_t gt = 0, qi = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int b, h, i0, kend, ntiles;
      item_of(item, b, h, i0, kend, ntiles);
      if (ntiles == 0) continue;
      ptx::mbar_wait(q_full, qi & 1);
      for (int t = 0; t < ntiles; ++t, ++gt) {
        ptx::mbar_wait(k_full, gt & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          ptx::umma_bf16_ss(tmem_base, ptx::make_smem_desc(aQ + k * 32, 16, 1024), ptx::make_smem_desc(aK + k * 32, 16, 1024),
                            idesc_s, k > 0 ? 1u : 0u);
        ptx::umma_commit(k_empty);
        ptx::umma_commit(s_full);
        if (t == ntiles - 1) ptx::umma_commit(q_empty);          // last S of the item: sQ may be refilled
        ptx::mbar_wait(p_full, gt & 1);
        ptx::mbar_wait(v_full, gt & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          const uint32_t pa = aP + (k >> 2) * (TILE_P / 2) + (k & 3) * 32;       // K-major, two 64-key halves
          ptx::umma_bf16_ss(tmem_base + O_COL, ptx::make_smem_desc(pa, 16, 1024),
                            ptx::make_smem_desc(aV + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit(v_empty);
        ptx::umma_commit(o_full);
      }
      ++qi;
      if (!PERSIST) break;
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int hf = (warp - 2) >> 2;              // keys [hf*64, hf*64+64) of every tile, output columns [hf*32, hf*32+32)
    const int r = quarter * 32 + lane;
    uint32_t gt = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    int b, h, i0, kend, ntiles;
    item_of(item, b, h, i0, kend, ntiles);
    const int i = i0 + r;
    const bool row_ok = i < p.Lq;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t sP_row = ptx::smem_u32(sP) + hf * (TILE_P / 2) + r * 128;
    const uint32_t rx = (uint32_t)(r & 7);
    const int ii = row_ok ? i : p.Lq - 1;
    const long long rowid = ((long long)b * p.H + h) * p.Lq + ii;
    const uint32_t rowkey = p.thr16 ? attn_row_key(eff_seed(p.seed, p.seed_dev), p.site, rowid - (ii & 7)) : 0u;
    uint32_t pm = 1u, pa = 0u;                     // row (ii & 7) of the 8x8 dropout block: 8 steps per row
    if (p.thr16) attn_advance(8 * (ii & 7), pm, pa);
    const float c2 = p.c_log2;
    float m = -INFINITY, l = 0.f;               // l: this thread's half of the row sum
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    // key-mask bitmap of all the tiles of this item, built once (one barrier per item instead of one per tile;
    // the barrier also separates this item's exchange buffers from the previous item's last read)
    const bool use_mask = p.pad != nullptr || (kend & (BN - 1)) != 0;
    if (use_mask) {
      for (int j = (warp - 2) * 32 + lane; j < ntiles * BN; j += 256) {
        const bool msk = j >= kend || (p.pad && p.pad[(long long)b * p.Lk + j]);
        const uint32_t bal = __ballot_sync(0xffffffffu, msk);
        if (lane == 0) sts_u(a_mask + (j >> 5) * 4, bal);
      }
    }
    bar_sync_bwd();

    for (int t = 0; t < ntiles; ++t, ++gt) {
      const int j0 = t * BN;
      uint32_t mw[2] = {0u, 0u};
      if (use_mask) {
        mw[0] = lds_u(a_mask + ((j0 >> 5) + hf * 2) * 4);
        mw[1] = lds_u(a_mask + ((j0 >> 5) + hf * 2 + 1) * 4);
      }
      if (p.causal && j0 + BN - 1 > i0) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int nvis = i - (j0 + hf * 64 + c * 32) + 1;
          mw[c] |= nvis <= 0 ? 0xffffffffu : (nvis >= 32 ? 0u : (0xffffffffu << nvis));
        }
      }
      ptx::mbar_wait(s_full, gt & 1);
      ptx::tc_fence_after();
      // ---- pass 1: maximum of this thread's 64 visible scores, then the row maximum via smem
      float mx = -INFINITY;
      {
        // four chunks of 16 columns, the next one in flight while the current one is reduced
        uint32_t v[2][16];
        ptx::tmem_ld_32x16(lane_addr + hf * 64, v[0]);
        ptx::tmem_ld_wait(v[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cur = c & 1, nxt = cur ^ 1;
          if (c < 3) {
            __syncwarp();
            ptx::tmem_ld_32x16(lane_addr + hf * 64 + (c + 1) * 16, v[nxt]);
          }
          const uint32_t w = (mw[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu;
          if (w == 0u) {
#pragma unroll
            for (int k = 0; k < 16; k += 2) mx = max3(mx, __uint_as_float(v[cur][k]), __uint_as_float(v[cur][k + 1]));
          } else if (w != 0xFFFFu) {
#pragma unroll
            for (int k = 0; k < 16; ++k) mx = fmaxf(mx, ((w >> k) & 1u) ? -INFINITY : __uint_as_float(v[cur][k]));
          }
          if (c < 3) {
            __syncwarp();
            ptx::tmem_ld_wait(v[nxt]);
          }
        }
      }
      const uint32_t xch = a_xch + (gt & 1) * 1024;
      sts_f(xch + (hf * 128 + r) * 4, mx);
      bar_sync_bwd();
      mx = fmaxf(mx, lds_f(xch + ((hf ^ 1) * 128 + r) * 4));
      const float m_new = fmaxf(m, mx * c2);
      const float m_use = m_new == -INFINITY ? 0.f : m_new;
      const float alpha = ex2(m - m_use);
      f32x2 lsum2 = pack2(0.f, 0.f);
      const f32x2 c2p = pack2(c2, c2), nm2 = pack2(-m_use, -m_use);
      const uint32_t kbase = rowkey + (uint32_t)((j0 + hf * 64) >> 3) * ATTN_GOLD;     // dropout block index of this thread's keys
      // ---- pass 2: P = exp2(S*c - m), row sum, dropout, bf16 pack into the swizzled A-operand tile
      // (dropout keeps P unscaled here: the 1/(1-p) factor is folded into the final O normalisation)
      // The masked variant is a separate instantiation behind a warp-uniform branch (tcgen05.ld is warp-collective):
      // inside one instantiation ptxas would predicate the per-score mask tests and issue them on every tile.
      auto pass2 = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        // eight chunks of 8 columns (= one dropout block each), the next one in flight while the current is consumed
        uint32_t v[2][8];
        if (MASKED) __syncwarp();
        ptx::tmem_ld_32x8(lane_addr + hf * 64, v[0]);
        ptx::tmem_ld_wait(v[0]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int cur = c & 1, nxt = cur ^ 1;
          if (c < 7) {
            if (MASKED) __syncwarp();                    // reconverge after the divergent skip below: tcgen05.ld is .aligned
            ptx::tmem_ld_32x8(lane_addr + hf * 64 + (c + 1) * 8, v[nxt]);
          }
          const uint32_t w = MASKED ? (mw[c >> 2] >> ((c & 3) * 8)) & 0xFFu : 0u;
          const uint32_t dst = sP_row + (((uint32_t)c ^ rx) << 4);
          if (MASKED && w == 0xFFu) {                    // nothing visible in this group: P = 0
            st_shared_v4(dst, 0u, 0u, 0u, 0u);
          } else {
            float pv[8];
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
              float s0 = __uint_as_float(v[cur][k]), s1 = __uint_as_float(v[cur][k + 1]);
              if (MASKED) {
                s0 = ((w >> k) & 1u) ? -INFINITY : s0;
                s1 = ((w >> (k + 1)) & 1u) ? -INFINITY : s1;
              }
              float e0, e1;
              unpack2(fma2(pack2(s0, s1), c2p, nm2), e0, e1);
              pv[k] = ex2(e0);
              pv[k + 1] = ex2(e1);
              lsum2 = add2(lsum2, pack2(pv[k], pv[k + 1]));
            }
            if (p.thr16) {                               // one mixed word per 8 keys, then one multiply-add per key
              uint32_t x[8];
              attn_block8<1>(attn_mix(kbase + (uint32_t)c * ATTN_GOLD) * pm + pa, x);
#pragma unroll
              for (int k = 0; k < 8; ++k) pv[k] = x[k] >= p.thr16 ? pv[k] : 0.f;
            }
            st_shared_v4(dst, pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]),
                         pack_bf16x2(pv[6], pv[7]));
          }
          if (c < 7) {
            if (MASKED) __syncwarp();
            ptx::tmem_ld_wait(v[nxt]);
          }
        }
      };
      if (__any_sync(0xffffffffu, (mw[0] | mw[1]) != 0u)) pass2(std::true_type{});
      else pass2(std::false_type{});
      {
        float ls0, ls1;
        unpack2(lsum2, ls0, ls1);
        l = l * alpha + (ls0 + ls1);
      }
      m = m_new;
      ptx::fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core
      ptx::tc_fence_before();            // orders this thread's tcgen05.ld before the next MMAs
      ptx::mbar_arrive(p_full);
      // ---- O += P V  (rescale the running accumulator, add the tile result): this thread's 32 columns
      ptx::mbar_wait(o_full, gt & 1);
      ptx::tc_fence_after();
      {
        uint32_t v[32];
        ptx::tmem_ld_32x32(lane_addr + O_COL + hf * 32, v);
        ptx::tmem_ld_wait();
        const f32x2 al2 = pack2(alpha, alpha);
#pragma unroll
        for (int k = 0; k < 32; k += 2)
          unpack2(fma2(pack2(acc[k], acc[k + 1]), al2, pack2(__uint_as_float(v[k]), __uint_as_float(v[k + 1]))), acc[k], acc[k + 1]);
      }
    }
    // row sum = both halves
    const uint32_t xch = a_xch + (gt & 1) * 1024;
    sts_f(xch + (hf * 128 + r) * 4, l);
    bar_sync_bwd();
    l += lds_f(xch + ((hf ^ 1) * 128 + r) * 4);
    if (row_ok) {
      const float inv = l > 0.f ? p.inv_keep / l : 0.f;      // dropout's 1/(1-p) applied once per row
      bf16* orow = p.o + ((long long)b * p.Lq + i) * p.ldo + h * DH + hf * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        uint4 u;
        u.x = pack_bf16x2(acc[c] * inv, acc[c + 1] * inv);
        u.y = pack_bf16x2(acc[c + 2] * inv, acc[c + 3] * inv);
        u.z = pack_bf16x2(acc[c + 4] * inv, acc[c + 5] * inv);
        u.w = pack_bf16x2(acc[c + 6] * inv, acc[c + 7] * inv);
        *reinterpret_cast<uint4*>(orow + c) = u;
      }
      if (p.lse && hf == 0)
        p.lse[((long long)b * p.H + h) * p.Lq + i] = l > 0.f ? (m + log2f(l)) * 0.6931471805599453f : -INFINITY;
    }
    if (!PERSIST) break;
    }   // items
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}


// =========================================================================================
// BACKWARD.  Two kernels, both with the forward's warp roles (TMA producer / MMA issuer /
// 128 softmax threads, thread <-> TMEM lane) and two CTAs per SM:
//   dQ kernel  : CTA = 128 query rows, 64-key K/V tiles processed as two 32-key sub-tiles with their own
//                S = Q K^T / dP = dO V^T buffers in TMEM, so the MMAs of the next sub-tile run while the
//                threads form dS = P o (dP*keep/(1-p) - D) in bf16 (smem, A operand) of the current one;
//                dQ += dS K accumulates in TMEM over the whole loop (no rescale in backward).
//                (The same two-buffer split of the dK/dV kernel measured 8 % slower and was dropped.)
//   dKV kernel : CTA = 128 keys, loop over 64-query tiles.  S^T = K Q^T, dP^T = V dO^T; threads
//                (<-> key row) write P^T*keep and dS^T; dV += P^T dO, dK += dS^T Q in TMEM.
// P is recomputed from the saved log-sum-exp; D = rowsum(dO o O) is formed by the dQ kernel's prologue.
// =========================================================================================
constexpr int BKV = 64;                         // second tile dimension of both backward kernels
constexpr int TILE_HALF = BKV * DH * 2;         // 8 KB
constexpr int BWD_THREADS = 320;          // TMA warp, MMA warp, 8 softmax warps (2 threads per TMEM lane)
constexpr int DQ_SMEM = 2 * TILE_QKV + 4 * TILE_HALF + TILE_QKV + 1024 + 1536 + MASK_WORDS * 4;    // + barriers, dsum exchange, key-mask bitmap
constexpr int DKV_SMEM = 2 * TILE_QKV + 4 * TILE_HALF + 2 * TILE_QKV + 1024 + 2048;

struct BwdParams {
  bf16 *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  const float* lse;
  float *dbq, *dbk, *dbv;    // nullable: [H*DH] += column sums of dQ / dK / dV (in-projection bias gradient)
  float* dsum;               // [B,H,Lq] rowsum(dO o O): written by the dQ kernel, read by the dK/dV kernel
  const bf16 *o, *dout;      // forward output and its gradient (for dsum)
  long long ldo, lddo;
  const int* kv_len;
  const uint8_t* pad;
  int B, H, Lq, Lk;
  float c_log2, scale;
  int causal;
  uint32_t thr16;          // dropout threshold p * 2^32 (0 = dropout off)
  float inv_keep;
  uint64_t seed, site;
  const unsigned long long* seed_dev;
};

// 8 consecutive bf16 of row `r` (16-byte chunk c16 of a 128-byte SWIZZLE_128B row)
__device__ __forceinline__ void st_row_chunk(uint32_t row_addr, uint32_t rx, uint32_t c16, const float* v) {
  st_shared_v4(row_addr + ((c16 ^ rx) << 4), pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
               pack_bf16x2(v[6], v[7]));
}


__global__ void __launch_bounds__(BWD_THREADS, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + TILE_QKV;
  uint8_t* sK = smem + 2 * TILE_QKV;                 // [2] x 8 KB
  uint8_t* sV = sK + 2 * TILE_HALF;                  // [2] x 8 KB
  uint8_t* sdS = sV + 2 * TILE_HALF;                 // 16 KB: [128 q][64 keys] K-major
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + TILE_QKV);
  uint64_t *qdo_full = bars, *kv_full = bars + 1 /*[2]*/, *kv_empty = bars + 3 /*[2]*/, *sp_full = bars + 5 /*[2]*/,
           *ds_full = bars + 7 /*[2]*/, *ds_free = bars + 9 /*[2]*/, *dq_done = bars + 11;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);
  const uint32_t a_mask = ptx::smem_u32(bars) + 1536;     // uint32 [MASK_WORDS], bit j: key j masked

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = gridDim.x - 1 - blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int i0 = qt * BM;
  int kend = p.kv_len ? min(p.kv_len[b], p.Lk) : p.Lk;
  if (p.causal) kend = min(kend, i0 + BM);
  const int ntiles = (kend + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmdO); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
    for (int i = 0; i < 12; ++i) ptx::mbar_init(bars + i, (i == 7 || i == 8) ? 256 : 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (ptx::elect_one() && ntiles > 0) {
      ptx::mbar_expect_tx(qdo_full, 2 * TILE_QKV);
      ptx::tma_load_2d(sQ, &tmQ, qdo_full, h * DH, b * p.Lq + i0);
      ptx::tma_load_2d(sdO, &tmdO, qdo_full, h * DH, b * p.Lq + i0);
      for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        if (t >= 2) ptx::mbar_wait(kv_empty + s, ((t - 2) >> 1) & 1);
        ptx::mbar_expect_tx(kv_full + s, 2 * TILE_HALF);
        ptx::tma_load_2d(sK + s * TILE_HALF, &tmK, kv_full + s, h * DH, b * p.Lk + t * BKV);
        ptx::tma_load_2d(sV + s * TILE_HALF, &tmV, kv_full + s, h * DH, b * p.Lk + t * BKV);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (ptx::elect_one() && ntiles > 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BM, 32, 0, 0);       // 32-key sub-tile, both operands K-major
      constexpr uint32_t idesc_kn = ptx::make_idesc_bf16(BM, DH, 0, 1);      // B MN-major
      const uint32_t aQ = ptx::smem_u32(sQ), adO = ptx::smem_u32(sdO), adS = ptx::smem_u32(sdS);
      const int nsub = 2 * ntiles;
      ptx::mbar_wait(qdo_full, 0);
      // S and dP of sub-tile u (keys [32u, 32u+32)) into TMEM buffer u & 1: columns [64b, 64b+32) and [64b+32, 64b+64)
      auto issue_sdp = [&](int u) {
        const int t = u >> 1, hh = u & 1, s = t & 1;
        if (hh == 0) {
          ptx::mbar_wait(kv_full + s, (t >> 1) & 1);
          ptx::tc_fence_after();
        }
        const uint32_t aK = ptx::smem_u32(sK + s * TILE_HALF) + hh * 4096, aV = ptx::smem_u32(sV + s * TILE_HALF) + hh * 4096;
        const uint32_t col = tmem_base + hh * 64;
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          ptx::umma_bf16_ss(col, ptx::make_smem_desc(aQ + kk * 32, 16, 1024), ptx::make_smem_desc(aK + kk * 32, 16, 1024),
                            idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          ptx::umma_bf16_ss(col + 32, ptx::make_smem_desc(adO + kk * 32, 16, 1024),
                            ptx::make_smem_desc(aV + kk * 32, 16, 1024), idesc_s, kk > 0 ? 1u : 0u);
        ptx::umma_commit(sp_full + hh);
      };
      issue_sdp(0);
      issue_sdp(1);
      for (int u = 0; u < nsub; ++u) {
        const int t = u >> 1, hh = u & 1, s = t & 1;
        const uint32_t aK = ptx::smem_u32(sK + s * TILE_HALF);
        ptx::mbar_wait(ds_full + hh, t & 1);       // S/dP buffer hh consumed, dS columns [32hh, 32hh+32) stored
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {           // dQ += dS[:, 32hh : 32hh+32] K[32hh : 32hh+32, :]
          const int ks = hh * 2 + kk;
          ptx::umma_bf16_ss(tmem_base + 128, ptx::make_smem_desc(adS + ks * 32, 16, 1024),
                            ptx::make_smem_desc(aK + ks * 2048, 8192, 1024), idesc_kn, (u > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(ds_free + hh);            // that half of the dS tile may be rewritten
        if (hh == 1) ptx::umma_commit(kv_empty + s);
        if (u + 2 < nsub) issue_sdp(u + 2);        // refill the buffer the softmax threads have just left
      }
      ptx::umma_commit(dq_done);
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int hf = (warp - 2) >> 2;                    // this thread's 32 of the tile's 64 key columns / dQ columns
    const int r = quarter * 32 + lane;
    const int i = i0 + r;
    const bool row_ok = i < p.Lq;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t row_addr = ptx::smem_u32(sdS) + r * 128;
    const uint32_t rx = (uint32_t)(r & 7);
    const int ii = row_ok ? i : p.Lq - 1;
    const long long rowid = ((long long)b * p.H + h) * p.Lq + ii;
    const uint32_t rowkey = p.thr16 ? attn_row_key(eff_seed(p.seed, p.seed_dev), p.site, rowid - (ii & 7)) : 0u;
    uint32_t pm = 1u, pa = 0u;                     // row (ii & 7) of the 8x8 dropout block: 8 steps per row
    if (p.thr16) attn_advance(8 * (ii & 7), pm, pa);
    float lse2 = INFINITY, dsum = 0.f;
    if (row_ok) {
      const float l = p.lse[rowid];
      lse2 = l == -INFINITY ? INFINITY : l * 1.4426950408889634f;
      // D_i = sum_c dO[i,c] * O[i,c]: this thread's 32 of the 64 columns (64 B of each row), then the pair's sum
      const uint4* orow = reinterpret_cast<const uint4*>(p.o + ((long long)b * p.Lq + i) * p.ldo + h * DH + hf * 32);
      const uint4* grow = reinterpret_cast<const uint4*>(p.dout + ((long long)b * p.Lq + i) * p.lddo + h * DH + hf * 32);
      f32x2 acc2 = pack2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 ov = orow[c], gv = grow[c];
        acc2 = fma2(bf2_to_f2(ov.x), bf2_to_f2(gv.x), acc2);
        acc2 = fma2(bf2_to_f2(ov.y), bf2_to_f2(gv.y), acc2);
        acc2 = fma2(bf2_to_f2(ov.z), bf2_to_f2(gv.z), acc2);
        acc2 = fma2(bf2_to_f2(ov.w), bf2_to_f2(gv.w), acc2);
      }
      float d0, d1;
      unpack2(acc2, d0, d1);
      dsum = d0 + d1;
    }
    // key-mask bitmap of all the tiles this CTA visits, built once (published by the barrier below)
    const bool use_mask = p.pad != nullptr || (kend & (BKV - 1)) != 0;
    if (use_mask) {
      for (int j = (warp - 2) * 32 + lane; j < ntiles * BKV; j += 256) {
        const bool msk = j >= kend || (p.pad && p.pad[(long long)b * p.Lk + j]);
        const uint32_t bal = __ballot_sync(0xffffffffu, msk);
        if (lane == 0) sts_u(a_mask + (j >> 5) * 4, bal);
      }
    }
    {
      const uint32_t xch = ptx::smem_u32(bars) + 256;      // float [2][128]
      sts_f(xch + (hf * 128 + r) * 4, dsum);
      bar_sync_bwd();
      dsum += lds_f(xch + ((hf ^ 1) * 128 + r) * 4);
      if (row_ok && hf == 0) p.dsum[rowid] = dsum;             // for the dK/dV kernel that follows on the stream
    }
    const float c2 = p.c_log2;
    const f32x2 c2p = pack2(c2, c2), nl2 = pack2(-lse2, -lse2), ik2 = pack2(p.inv_keep, p.inv_keep), nd2 = pack2(-dsum, -dsum);
    for (int u = 0; u < 2 * ntiles; ++u) {
      const int t = u >> 1, hh = u & 1;
      const int j0 = t * BKV + hh * 32;                // first key of the 32-key sub-tile; this thread: keys j0 + 16hf ...
      uint32_t w32 = use_mask ? lds_u(a_mask + (j0 >> 5) * 4) : 0u;
      if (p.causal && j0 + 31 > i0) {
        const int nvis = i - j0 + 1;
        w32 |= nvis <= 0 ? 0xffffffffu : (nvis >= 32 ? 0u : (0xffffffffu << nvis));
      }
      const uint32_t w = (w32 >> (hf * 16)) & 0xFFFFu;
      ptx::mbar_wait(sp_full + hh, t & 1);
      ptx::tc_fence_after();
      if (t > 0) ptx::mbar_wait(ds_free + hh, (t - 1) & 1);      // dQ of the previous tile has read this half of the dS tile
      {
        const uint32_t scol = lane_addr + hh * 64 + hf * 16;     // S columns of this thread; dP is 32 columns further
        // masked sub-tiles take a separate instantiation behind a warp-uniform branch (see the forward kernel)
        auto chunks = [&](auto masked_tag) {
          constexpr bool MASKED = decltype(masked_tag)::value;
          // two chunks of 8 columns (= one dropout block each); the second chunk's TMEM loads are in flight while
          // the first is consumed
          uint32_t sv[2][8], dv[2][8];
          if (MASKED) __syncwarp();
          ptx::tmem_ld_32x8(scol, sv[0]);
          ptx::tmem_ld_32x8(scol + 32, dv[0]);
          ptx::tmem_ld_32x8(scol + 8, sv[1]);
          ptx::tmem_ld_32x8(scol + 40, dv[1]);
          ptx::tmem_ld_wait(sv[0], dv[0]);
          ptx::tmem_ld_wait(sv[1], dv[1]);
          const uint32_t kbase = rowkey + (uint32_t)((j0 + hf * 16) >> 3) * ATTN_GOLD;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t w8 = MASKED ? (w >> (c * 8)) & 0xFFu : 0u;
            const uint32_t dst = row_addr + (((uint32_t)(hh * 4 + hf * 2 + c) ^ rx) << 4);
            if (MASKED && w8 == 0xFFu) {                 // nothing visible in this group: dS = 0
              st_shared_v4(dst, 0u, 0u, 0u, 0u);
            } else {
              float ds[8];
              uint32_t x[8];
              attn_block8<1>(p.thr16 ? attn_mix(kbase + (uint32_t)c * ATTN_GOLD) * pm + pa : 0u, x);
#pragma unroll
              for (int k = 0; k < 8; k += 2) {
                float e0, e1;
                unpack2(fma2(pack2(__uint_as_float(sv[c][k]), __uint_as_float(sv[c][k + 1])), c2p, nl2), e0, e1);
                float p0 = ex2(e0), p1 = ex2(e1);
                if (MASKED) {
                  p0 = ((w8 >> k) & 1u) ? 0.f : p0;
                  p1 = ((w8 >> (k + 1)) & 1u) ? 0.f : p1;
                }
                float d0 = __uint_as_float(dv[c][k]), d1 = __uint_as_float(dv[c][k + 1]);
                if (p.thr16) {
                  d0 = x[k] >= p.thr16 ? d0 : 0.f;
                  d1 = x[k + 1] >= p.thr16 ? d1 : 0.f;
                }
                // dS = P * (dP * keep/(1-p) - D)
                unpack2(mul2(pack2(p0, p1), fma2(pack2(d0, d1), ik2, nd2)), ds[k], ds[k + 1]);
              }
              st_shared_v4(dst, pack_bf16x2(ds[0], ds[1]), pack_bf16x2(ds[2], ds[3]), pack_bf16x2(ds[4], ds[5]),
                           pack_bf16x2(ds[6], ds[7]));
            }
          }
        };
        if (__any_sync(0xffffffffu, w != 0u)) chunks(std::true_type{});
        else chunks(std::false_type{});
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(ds_full + hh);
    }
    if (ntiles > 0) {
      ptx::mbar_wait(dq_done, 0);
      ptx::tc_fence_after();
    }
    bf16* drow = p.dq + ((long long)b * p.Lq + i) * p.lddq + h * DH + hf * 32;
    {
      uint32_t v[32];
      if (ntiles > 0) {                    // warp-uniform: tcgen05.ld is .sync.aligned
        ptx::tmem_ld_32x32(lane_addr + 128 + hf * 32, v);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0u;
      }
      if (row_ok) {
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[k]) * p.scale, __uint_as_float(v[k + 1]) * p.scale);
          u.y = pack_bf16x2(__uint_as_float(v[k + 2]) * p.scale, __uint_as_float(v[k + 3]) * p.scale);
          u.z = pack_bf16x2(__uint_as_float(v[k + 4]) * p.scale, __uint_as_float(v[k + 5]) * p.scale);
          u.w = pack_bf16x2(__uint_as_float(v[k + 6]) * p.scale, __uint_as_float(v[k + 7]) * p.scale);
          *reinterpret_cast<uint4*>(drow + k) = u;
        }
      }
      if (p.dbq) {                       // bias gradient of the Q projection: column sums over this CTA's 128 rows
        float f[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) f[k] = row_ok ? __uint_as_float(v[k]) * p.scale : 0.f;
        const float cs = warp_colsum32(f, lane);
        atomicAdd(p.dbq + h * DH + hf * 32 + lane, cs);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}


__global__ void __launch_bounds__(BWD_THREADS, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                       const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + TILE_QKV;
  uint8_t* sQ = smem + 2 * TILE_QKV;                 // [2] x 8 KB  (64 query rows each)
  uint8_t* sdO = sQ + 2 * TILE_HALF;                 // [2] x 8 KB
  uint8_t* sPd = sdO + 2 * TILE_HALF;                // 16 KB: [128 keys][64 q] K-major, P^T * keep/(1-p)
  uint8_t* sdS = sPd + TILE_QKV;                     // 16 KB: [128 keys][64 q] K-major, dS^T
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + TILE_QKV);
  uint64_t *kv_full = bars, *qdo_full = bars + 1 /*[2]*/, *qdo_empty = bars + 3 /*[2]*/, *sp_full = bars + 5,
           *pds_full = bars + 6, *done = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t a_lse = ptx::smem_u32(bars) + 128;        // float [2][64], 16-byte aligned
  const uint32_t a_dsum = a_lse + 2 * BKV * 4;             // float [2][64]
  const uint32_t a_key = a_dsum + 2 * BKV * 4;             // uint32 [2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int j0 = blockIdx.x * BM;
  const int kend = p.kv_len ? min(p.kv_len[b], p.Lk) : p.Lk;
  const int nq = (p.Lq + BKV - 1) / BKV;
  const int it0 = p.causal ? j0 / BKV : 0;           // queries i >= j0 only
  const int ntiles = j0 < kend ? max(0, nq - it0) : 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmdO); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
    for (int i = 0; i < 8; ++i) ptx::mbar_init(bars + i, i == 6 ? 256 : 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (ptx::elect_one() && ntiles > 0) {
      ptx::mbar_expect_tx(kv_full, 2 * TILE_QKV);
      ptx::tma_load_2d(sK, &tmK, kv_full, h * DH, b * p.Lk + j0);
      ptx::tma_load_2d(sV, &tmV, kv_full, h * DH, b * p.Lk + j0);
      for (int n = 0; n < ntiles; ++n) {
        const int s = n & 1;
        if (n >= 2) ptx::mbar_wait(qdo_empty + s, ((n - 2) >> 1) & 1);
        ptx::mbar_expect_tx(qdo_full + s, 2 * TILE_HALF);
        ptx::tma_load_2d(sQ + s * TILE_HALF, &tmQ, qdo_full + s, h * DH, b * p.Lq + (it0 + n) * BKV);
        ptx::tma_load_2d(sdO + s * TILE_HALF, &tmdO, qdo_full + s, h * DH, b * p.Lq + (it0 + n) * BKV);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (ptx::elect_one() && ntiles > 0) {
      constexpr uint32_t idesc_kk = ptx::make_idesc_bf16(BM, BKV, 0, 0);
      constexpr uint32_t idesc_kn = ptx::make_idesc_bf16(BM, DH, 0, 1);
      const uint32_t aK = ptx::smem_u32(sK), aV = ptx::smem_u32(sV), aPd = ptx::smem_u32(sPd), adS = ptx::smem_u32(sdS);
      ptx::mbar_wait(kv_full, 0);
      for (int n = 0; n < ntiles; ++n) {
        const int s = n & 1;
        const uint32_t aQ = ptx::smem_u32(sQ + s * TILE_HALF), adO = ptx::smem_u32(sdO + s * TILE_HALF);
        ptx::mbar_wait(qdo_full + s, (n >> 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)          // S^T = K Q^T
          ptx::umma_bf16_ss(tmem_base, ptx::make_smem_desc(aK + k * 32, 16, 1024), ptx::make_smem_desc(aQ + k * 32, 16, 1024),
                            idesc_kk, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)          // dP^T = V dO^T
          ptx::umma_bf16_ss(tmem_base + 64, ptx::make_smem_desc(aV + k * 32, 16, 1024),
                            ptx::make_smem_desc(adO + k * 32, 16, 1024), idesc_kk, k > 0 ? 1u : 0u);
        ptx::umma_commit(sp_full);
        ptx::mbar_wait(pds_full, n & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)         // dV += (P^T keep) dO
          ptx::umma_bf16_ss(tmem_base + 192, ptx::make_smem_desc(aPd + k * 32, 16, 1024),
                            ptx::make_smem_desc(adO + k * 2048, 8192, 1024), idesc_kn, (n > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)         // dK += dS^T Q
          ptx::umma_bf16_ss(tmem_base + 128, ptx::make_smem_desc(adS + k * 32, 16, 1024),
                            ptx::make_smem_desc(aQ + k * 2048, 8192, 1024), idesc_kn, (n > 0 || k > 0) ? 1u : 0u);
        ptx::umma_commit(qdo_empty + s);
      }
      ptx::umma_commit(done);
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int hf = (warp - 2) >> 2;                    // this thread's 32 of the tile's 64 query columns / output columns
    const int r = quarter * 32 + lane;
    const int j = j0 + r;
    const bool row_ok = j < p.Lk;
    const bool key_masked = j >= kend || (p.pad && row_ok && p.pad[(long long)b * p.Lk + j]);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t pd_row = ptx::smem_u32(sPd) + r * 128;
    const uint32_t ds_row = ptx::smem_u32(sdS) + r * 128;
    const uint32_t rx = (uint32_t)(r & 7);
    const long long rowbase = ((long long)b * p.H + h) * p.Lq;
    const float c2 = p.c_log2;
    uint32_t jm = 1u, ja = 0u;                         // key (j & 7) of the 8x8 dropout block: one step per key
    if (p.thr16) attn_advance(j & 7, jm, ja);
    const uint32_t jg = (uint32_t)(j >> 3) * ATTN_GOLD;
    for (int n = 0; n < ntiles; ++n) {
      const int iq0 = (it0 + n) * BKV;
      const int slot = (n & 1) * BKV;
      if (r < BKV && hf == 0) {
        const int i = iq0 + r;
        float l2 = INFINITY, ds_ = 0.f;
        uint32_t rk = 0u;
        if (i < p.Lq) {
          const float l = p.lse[rowbase + i];
          l2 = l == -INFINITY ? INFINITY : l * 1.4426950408889634f;
          ds_ = p.dsum[rowbase + i];
          if (p.thr16) rk = attn_row_key(eff_seed(p.seed, p.seed_dev), p.site, rowbase + (i & ~7));   // key of the 8-row block
        }
        sts_f(a_lse + (slot + r) * 4, l2);
        sts_f(a_dsum + (slot + r) * 4, ds_);
        sts_u(a_key + (slot + r) * 4, rk);
      }
      bar_sync_bwd();
      const int cm = (p.causal && iq0 < j0 + BM) ? (j - iq0) : 0;       // query columns < cm cannot see key j
      ptx::mbar_wait(sp_full, n & 1);
      ptx::tc_fence_after();
      // masked tiles (padded keys, causal diagonal) take a separate instantiation behind a warp-uniform branch
      auto chunks = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        // four chunks of 8 queries (= one dropout block each); the next chunk's TMEM loads are in flight while
        // the current one is consumed
        uint32_t sv[2][8], dv[2][8];
        ptx::tmem_ld_32x8(lane_addr + hf * 32, sv[0]);
        ptx::tmem_ld_32x8(lane_addr + 64 + hf * 32, dv[0]);
        ptx::tmem_ld_wait(sv[0], dv[0]);
        const f32x2 c2p = pack2(c2, c2);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cur = c & 1, nxt = cur ^ 1;
          const int cb = hf * 32 + c * 8;
          if (c < 3) {
            ptx::tmem_ld_32x8(lane_addr + cb + 8, sv[nxt]);
            ptx::tmem_ld_32x8(lane_addr + 64 + cb + 8, dv[nxt]);
          }
          float pd[8], ds[8];
          // 8 consecutive queries (tile starts are multiples of 8) share one mixed word; each next query is 8 steps on
          uint32_t x[8];
          attn_block8<8>(p.thr16 ? attn_mix(lds_u(a_key + (slot + cb) * 4) + jg) * jm + ja : 0u, x);
#pragma unroll
          for (int k4 = 0; k4 < 2; ++k4) {             // per-query-row constants come as 16-byte broadcast loads
            const int colb = cb + k4 * 4;
            const float4 l4 = lds_v4f(a_lse + (slot + colb) * 4);
            const float4 d4 = lds_v4f(a_dsum + (slot + colb) * 4);
            const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
              const int k = k4 * 4 + u;
              const int col = colb + u;
              float e0, e1;
              unpack2(fma2(pack2(__uint_as_float(sv[cur][k]), __uint_as_float(sv[cur][k + 1])), c2p, pack2(-lv[u], -lv[u + 1])), e0, e1);
              float p0 = ex2(e0), p1 = ex2(e1);
              if (MASKED) {
                if (key_masked || col < cm) p0 = 0.f;
                if (key_masked || col + 1 < cm) p1 = 0.f;
              }
              const f32x2 pp = pack2(p0, p1);
              const f32x2 dd2 = pack2(__uint_as_float(dv[cur][k]), __uint_as_float(dv[cur][k + 1]));
              if (p.thr16) {
                // kf = keep/(1-p): one select per element serves both P^T*kf (dV operand) and dP*kf
                const f32x2 kf = pack2(x[k] >= p.thr16 ? p.inv_keep : 0.f, x[k + 1] >= p.thr16 ? p.inv_keep : 0.f);
                unpack2(mul2(pp, kf), pd[k], pd[k + 1]);
                // dS^T = P * (dP * keep/(1-p) - D)
                unpack2(mul2(pp, fma2(dd2, kf, pack2(-dd[u], -dd[u + 1]))), ds[k], ds[k + 1]);
              } else {
                pd[k] = p0;
                pd[k + 1] = p1;
                unpack2(mul2(pp, add2(dd2, pack2(-dd[u], -dd[u + 1]))), ds[k], ds[k + 1]);
              }
            }
          }
          st_row_chunk(pd_row, rx, (uint32_t)(hf * 4 + c), pd);
          st_row_chunk(ds_row, rx, (uint32_t)(hf * 4 + c), ds);
          if (c < 3) ptx::tmem_ld_wait(sv[nxt], dv[nxt]);
        }
      };
      if (__any_sync(0xffffffffu, key_masked || cm > hf * 32)) chunks(std::true_type{});
      else chunks(std::false_type{});
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(pds_full);
    }
    if (ntiles > 0) {
      ptx::mbar_wait(done, 0);
      ptx::tc_fence_after();
    }
#pragma unroll
    for (int part = 0; part < 2; ++part) {           // 0: dK (scaled), 1: dV; this thread's 32 columns
      bf16* drow = (part == 0 ? p.dk + ((long long)b * p.Lk + j) * p.lddk : p.dv + ((long long)b * p.Lk + j) * p.lddv) + h * DH + hf * 32;
      const float sc = part == 0 ? p.scale : 1.f;             // dV's operand P^T already carries keep/(1-p)
      uint32_t v[32];
      if (ntiles > 0) {
        ptx::tmem_ld_32x32(lane_addr + 128 + part * 64 + hf * 32, v);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0u;
      }
      if (row_ok) {
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[k]) * sc, __uint_as_float(v[k + 1]) * sc);
          u.y = pack_bf16x2(__uint_as_float(v[k + 2]) * sc, __uint_as_float(v[k + 3]) * sc);
          u.z = pack_bf16x2(__uint_as_float(v[k + 4]) * sc, __uint_as_float(v[k + 5]) * sc);
          u.w = pack_bf16x2(__uint_as_float(v[k + 6]) * sc, __uint_as_float(v[k + 7]) * sc);
          *reinterpret_cast<uint4*>(drow + k) = u;
        }
      }
      float* db = part == 0 ? p.dbk : p.dbv;
      if (db) {                          // bias gradient of the K / V projection
        float f[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) f[k] = row_ok ? __uint_as_float(v[k]) * sc : 0.f;
        const float cs = warp_colsum32(f, lane);
        atomicAdd(db + h * DH + hf * 32 + lane, cs);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

int check_common(const smer_attn_args* a, const char* who) {
  if (!a) { smer_set_error("%s: null args", who); return SMER_ERR_ARG; }
  if (a->dtype != SMER_DT_BF16 || a->dh != DH) {
    smer_set_error("%s: bf16 with head dim 64 only (dtype=%d dh=%d)", who, a->dtype, a->dh);
    return SMER_ERR_UNSUPPORTED;
  }
  if (a->add_mask || a->q_pos0 != 0) {
    smer_set_error("%s: additive masks / query offsets are served by the simt kernels", who);
    return SMER_ERR_UNSUPPORTED;
  }
  if (a->causal && a->Lq != a->Lk) { smer_set_error("%s: causal needs Lq == Lk", who); return SMER_ERR_UNSUPPORTED; }
  if (a->Lk > MASK_WORDS * 32) { smer_set_error("%s: Lk <= %d (key-mask bitmap in shared memory)", who, MASK_WORDS * 32); return SMER_ERR_UNSUPPORTED; }
  if (a->B <= 0 || a->H <= 0 || a->Lq <= 0 || a->Lk <= 0) { smer_set_error("%s: empty problem", who); return SMER_ERR_ARG; }
  return SMER_OK;
}

}  // namespace

int smer_attn_fwd2_launch(const smer_attn_args* a, void* stream);      // attn_fwd2.cu

extern "C" int smer_attn_fwd_tc(const smer_attn_args* a, void* stream) {
  int rc = check_common(a, "smer_attn_fwd_tc");
  if (rc) return rc;
  SMER_CHECK_ARG(a->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(a->o) & 15) == 0, "smer_attn_fwd_tc: o must be 16-byte aligned rows");
  static const bool use_v1 = getenv("SMER_ATTN_FWD_V1") != nullptr;      // A/B aid: the first-generation kernel below
  if (!use_v1) return smer_attn_fwd2_launch(a, stream);
  CUtensorMap tq, tk, tv;
  const long long dcols = (long long)a->H * DH;
  if ((rc = smer_make_tmap_bf16(&tq, a->q, dcols, (long long)a->B * a->Lq, a->ldq, DH, BM))) return rc;
  if ((rc = smer_make_tmap_bf16(&tk, a->k, dcols, (long long)a->B * a->Lk, a->ldk, DH, BN))) return rc;
  if ((rc = smer_make_tmap_bf16(&tv, a->v, dcols, (long long)a->B * a->Lk, a->ldv, DH, BN))) return rc;
  FwdParams p;
  p.o = (bf16*)a->o; p.ldo = a->ldo; p.lse = a->lse; p.kv_len = a->kv_len; p.pad = a->key_pad;
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk;
  p.c_log2 = a->scale * 1.4426950408889634f;
  p.causal = a->causal;
  p.thr16 = a->dropout_p > 0.f ? dropout_threshold(a->dropout_p) : 0u;      // p * 2^32 (full-word compare)
  p.inv_keep = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f;
  p.seed = a->seed; p.site = a->site; p.seed_dev = smer_seed_dev();
  static bool attr_set = false;
  if (!attr_set) {
    SMER_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    SMER_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    attr_set = true;
  }
  p.nq = (a->Lq + BM - 1) / BM;
  const long long items = (long long)p.nq * a->H * a->B;
  SMER_CHECK_ARG(items < (1ll << 31), "smer_attn_fwd_tc: too many work items");
  p.items = (int)items;
  // Causal: persistent CTAs (two per SM) walking the items heaviest-first -- the per-CTA set-up (TMEM allocation,
  // barrier init, first loads) is paid once instead of once per 1..8-tile item: 136 -> 127 us at B32 x 1024^2.
  // Full attention: one item per CTA; there the hardware's dynamic block scheduling balances the kv_len-dependent
  // item lengths better than a static round-robin (184 vs 196 us).
  const long long slots = 2ll * smer_num_sms();
  if (a->causal && items > slots)
    attn_fwd_tc_kernel<true><<<dim3((unsigned)slots), FWD_THREADS, FWD_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, p);
  else
    attn_fwd_tc_kernel<false><<<dim3((unsigned)items), FWD_THREADS, FWD_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, p);
  SMER_CHECK_LAUNCH("smer_attn_fwd_tc");
  return SMER_OK;
}

extern "C" int smer_attn_bwd_tc(const smer_attn_args* a, void* stream) {
  int rc = check_common(a, "smer_attn_bwd_tc");
  if (rc) return rc;
  SMER_CHECK_ARG(a->dout && a->dq && a->dk && a->dv && a->lse && a->dsum && a->o, "smer_attn_bwd_tc: missing buffers");
  SMER_CHECK_ARG(a->lddq % 8 == 0 && a->lddk % 8 == 0 && a->lddv % 8 == 0 && a->ldo % 8 == 0 && a->lddo % 8 == 0,
                 "smer_attn_bwd_tc: row pitches must be multiples of 8 elements");
  cudaStream_t st = (cudaStream_t)stream;
  const long long dcols = (long long)a->H * DH;
  const long long rq = (long long)a->B * a->Lq, rk = (long long)a->B * a->Lk;
  BwdParams p;
  p.dq = (bf16*)a->dq; p.dk = (bf16*)a->dk; p.dv = (bf16*)a->dv;
  p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
  p.lse = a->lse; p.dsum = a->dsum; p.kv_len = a->kv_len; p.pad = a->key_pad;
  p.dbq = a->dbq; p.dbk = a->dbk; p.dbv = a->dbv;
  p.o = (const bf16*)a->o; p.dout = (const bf16*)a->dout; p.ldo = a->ldo; p.lddo = a->lddo;
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk;
  p.c_log2 = a->scale * 1.4426950408889634f; p.scale = a->scale; p.causal = a->causal;
  p.thr16 = a->dropout_p > 0.f ? dropout_threshold(a->dropout_p) : 0u;      // p * 2^32 (full-word compare)
  p.inv_keep = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f;
  p.seed = a->seed; p.site = a->site; p.seed_dev = smer_seed_dev();
  static bool attr_set = false;
  if (!attr_set) {
    SMER_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    SMER_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM));
    attr_set = true;
  }
  CUtensorMap tq, tdo, tk, tv;
  // dQ kernel: 128-row Q / dO boxes, 64-row K / V boxes
  if ((rc = smer_make_tmap_bf16(&tq, a->q, dcols, rq, a->ldq, DH, BM))) return rc;
  if ((rc = smer_make_tmap_bf16(&tdo, a->dout, dcols, rq, a->lddo, DH, BM))) return rc;
  if ((rc = smer_make_tmap_bf16(&tk, a->k, dcols, rk, a->ldk, DH, BKV))) return rc;
  if ((rc = smer_make_tmap_bf16(&tv, a->v, dcols, rk, a->ldv, DH, BKV))) return rc;
  dim3 gq((a->Lq + BM - 1) / BM, a->H, a->B);
  attn_bwd_dq_tc_kernel<<<gq, BWD_THREADS, DQ_SMEM, st>>>(tq, tdo, tk, tv, p);
  // dKV kernel: 64-row Q / dO boxes, 128-row K / V boxes
  if ((rc = smer_make_tmap_bf16(&tq, a->q, dcols, rq, a->ldq, DH, BKV))) return rc;
  if ((rc = smer_make_tmap_bf16(&tdo, a->dout, dcols, rq, a->lddo, DH, BKV))) return rc;
  if ((rc = smer_make_tmap_bf16(&tk, a->k, dcols, rk, a->ldk, DH, BM))) return rc;
  if ((rc = smer_make_tmap_bf16(&tv, a->v, dcols, rk, a->ldv, DH, BM))) return rc;
  dim3 gk((a->Lk + BM - 1) / BM, a->H, a->B);
  attn_bwd_dkv_tc_kernel<<<gk, BWD_THREADS, DKV_SMEM, st>>>(tq, tdo, tk, tv, p);
  SMER_CHECK_LAUNCH("smer_attn_bwd_tc");
  return SMER_OK;
}
