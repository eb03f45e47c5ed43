// Attention backward, second generation (bf16, dh = 64): ONE fused kernel on tcgen05 / TMEM / TMA.
// Gradients of softmax(q k^T / sqrt(dh) + masks) v with dropout on P -- the need_weights branch of
// F.multi_head_attention_forward as the reference reaches it (transformer.py:389,459,463).
//
// CTA = one 128-key K/V tile of one (batch, head); it loops over the 128-query tiles that see those keys.
// S = Q K^T and dP = dO V^T are computed ONCE per (query tile, key tile) pair (the first generation's two kernels
// each recomputed both, and the exponentials):
//   warp 16      TMA producer: K, V once; Q_n / dO_n through a 2-stage ring
//   warp 17      tcgen05.mma issuer + TMEM owner.  TMEM (512 columns allocated, 448 used):
//                S [0,128) | dP [128,256) | dV [256,320) | dK [320,384) | dQ_n [384,448)
//                per query tile n:  S = Q_n K^T, dP = dO_n V^T            (issued while the threads still work on tile n-1)
//                                   dV += Pd^T dO_n, dK += dS^T Q_n, dQ_n = dS K   (A operands Pd / dS from shared memory,
//                                   written once by the threads: MN-major view for dV / dK, K-major view for dQ)
//   warps 0..15  element-wise work, TMEM lane = query row; warpgroup g owns key columns [32g, 32g+32) of the tile:
//                P = exp2(S c - lse), keep mask, Pd = P keep/(1-p) (bf16), dS = P (dP keep/(1-p) - D) (bf16) -> registers;
//                then (once the previous tile's MMAs are done) move its share of dQ_{n-1} from TMEM to a shared-memory
//                staging tile, store Pd / dS to shared memory and hand them to the MMA warp.
//   warp 18      dQ reducer: when the staging tile is complete, ONE bulk reduction (cp.reduce.async.bulk.tensor .add,
//                fp32) adds it to the accumulation buffer in L2 -- per-thread red.global.add.v4 of the same data kept
//                the LSU busy for ~40 % of the kernel.
//   warp 19      idle (roles are warpgroup-aligned for setmaxnreg).
// D = rowsum(dO o O) comes from a small pre-kernel; dQ accumulates across the key-tile CTAs in an fp32 buffer and a
// post-kernel scales / converts it to bf16 and reduces the in-projection's Q-bias gradient.
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "../../include/smer_b200.h"

int smer_make_tmap_bf16(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long pitch,
                        int box_inner, int box_outer);
int smer_make_tmap_f32(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long pitch,
                       int box_inner, int box_outer);

namespace {

constexpr int BM = 128, BN = 128, DH = 64;
constexpr int TILE = BM * DH * 2;                 // 16 KB
constexpr int THREADS = 640;
constexpr int OFF_K = 0, OFF_V = TILE, OFF_Q = 2 * TILE /*[2]*/, OFF_DO = 4 * TILE /*[2]*/, OFF_P = 6 * TILE /*32 KB*/,
              OFF_DS = 8 * TILE /*32 KB*/, OFF_DQ = 10 * TILE /*32 KB: [2 column halves][128 rows][32 fp32], swizzled*/,
              OFF_BAR = 12 * TILE;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024 /*alignment slack*/;
constexpr int TMEM_COLS = 512;
constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;

struct Params {
  bf16 *dk, *dv;
  long long lddk, lddv;
  float* dq_acc;             // [B*Lq, H*DH] fp32, zeroed: unscaled dQ accumulates here
  const float* lse;
  const float* dsum;         // [B,H,Lq] rowsum(dO o O)
  float *dbk, *dbv;          // nullable: [H*DH] += column sums of dK / dV (in-projection bias gradient)
  const int* kv_len;
  const uint8_t* pad;
  int B, H, Lq, Lk;
  float c_log2, scale;
  int causal;
  uint32_t thr2;             // dropout threshold pattern in both halves (common.cuh: attn_dropout_threshold); 0 = dropout off
  float inv_keep;
  uint64_t seed, site;
  const unsigned long long* seed_dev;
  const int *cu_q, *cu_k;    // padding-free layout (smer_b200.h) or NULL
  long long k_rows;          // rows of the packed K / V (and dK / dV) buffers
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <bool DROP>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                 const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ CUtensorMap tmDQ, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t *kv_full = bars, *qdo_full = bars + 1 /*[2]*/, *qdo_empty = bars + 3 /*[2]*/, *sdp_full = bars + 5,
           *sdp_free = bars + 6, *pds_full = bars + 7, *mma3_done = bars + 8, *dqs_full = bars + 9, *dqs_free = bars + 10;
  constexpr int NBARS = 11;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int j0 = blockIdx.x * BN;                    // first key of this CTA
  // rows of sequence b in the Q / dO (and dq_acc) and K / V buffers, and its lse / dsum / dropout row ids
  const int q_row0 = p.cu_q ? p.cu_q[b] : b * p.Lq, lq = p.cu_q ? p.cu_q[b + 1] - q_row0 : p.Lq;
  const int k_row0 = p.cu_q ? p.cu_k[b] : b * p.Lk, lk = p.cu_q ? p.cu_k[b + 1] - k_row0 : p.Lk;
  const long long rowbase = p.cu_q ? (long long)h * p.cu_q[p.B] + q_row0 : ((long long)b * p.H + h) * p.Lq;
  const int klen = p.cu_q ? lk : (p.kv_len ? p.kv_len[b] : p.Lk);    // < 0: the key mask has holes (smer_b200.h)
  const int kend = min(abs(klen), lk);
  const bool read_pad = p.pad != nullptr && (klen < 0 || p.kv_len == nullptr);
  const int nqt = (lq + BM - 1) / BM;
  const int it0 = p.causal ? j0 / BM : 0;            // causal: queries i >= j0 only
  const int ntiles = j0 < kend ? max(0, nqt - it0) : 0;

  if (warp == 16 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmdO); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
    for (int i = 0; i < NBARS; ++i) ptx::mbar_init(bars + i, (i == 6 || i == 7 || i == 9) ? 16 : 1);
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
      // ---------------------------------------------------------------- TMA producer
      if (ptx::elect_one() && ntiles > 0) {
        ptx::mbar_expect_tx(kv_full, 2 * TILE);
        ptx::tma_load_2d(smem + OFF_K, &tmK, kv_full, h * DH, k_row0 + j0);
        ptx::tma_load_2d(smem + OFF_V, &tmV, kv_full, h * DH, k_row0 + j0);
        for (int n = 0; n < ntiles; ++n) {
          const int s = n & 1;
          if (n >= 2) ptx::mbar_wait(qdo_empty + s, ((n - 2) >> 1) & 1);
          ptx::mbar_expect_tx(qdo_full + s, 2 * TILE);
          ptx::tma_load_2d(smem + OFF_Q + s * TILE, &tmQ, qdo_full + s, h * DH, q_row0 + (it0 + n) * BM);
          ptx::tma_load_2d(smem + OFF_DO + s * TILE, &tmdO, qdo_full + s, h * DH, q_row0 + (it0 + n) * BM);
        }
      }
      __syncwarp();
    } else if (warp == 17) {
      // ---------------------------------------------------------------- MMA issuer
      if (ptx::elect_one() && ntiles > 0) {
        constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BM, BN, 0, 0);        // S, dP: both operands K-major
        constexpr uint32_t idesc_t = ptx::make_idesc_bf16(BM, DH, 1, 1);        // dV, dK: A (Pd^T / dS^T) and B MN-major
        constexpr uint32_t idesc_q = ptx::make_idesc_bf16(BM, DH, 0, 1);        // dQ: A K-major, B MN-major
        const uint32_t aK = ptx::smem_u32(smem + OFF_K), aV = ptx::smem_u32(smem + OFF_V);
        const uint32_t aP = ptx::smem_u32(smem + OFF_P), aDS = ptx::smem_u32(smem + OFF_DS);
        auto issue_sdp = [&](int n) {
          const int s = n & 1;
          const uint32_t aQ = ptx::smem_u32(smem + OFF_Q + s * TILE), aDO = ptx::smem_u32(smem + OFF_DO + s * TILE);
          ptx::mbar_wait(qdo_full + s, (n >> 1) & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)            // S = Q_n K^T
            ptx::umma_bf16_ss(tmem_base + COL_S, ptx::make_smem_desc(aQ + k * 32, 16, 1024),
                              ptx::make_smem_desc(aK + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)            // dP = dO_n V^T
            ptx::umma_bf16_ss(tmem_base + COL_DP, ptx::make_smem_desc(aDO + k * 32, 16, 1024),
                              ptx::make_smem_desc(aV + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
          ptx::umma_commit(sdp_full);
        };
        ptx::mbar_wait(kv_full, 0);
        issue_sdp(0);
        for (int n = 0; n < ntiles; ++n) {
          const int s = n & 1;
          const uint32_t aQ = ptx::smem_u32(smem + OFF_Q + s * TILE), aDO = ptx::smem_u32(smem + OFF_DO + s * TILE);
          if (n + 1 < ntiles) {
            ptx::mbar_wait(sdp_free, n & 1);            // S / dP of tile n are in the threads' registers
            ptx::tc_fence_after();
            issue_sdp(n + 1);
          }
          ptx::mbar_wait(pds_full, n & 1);              // Pd / dS of tile n are in shared memory, dQ_{n-1} has been drained
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < BM / 16; ++k)             // dV += Pd^T dO_n   (K = 128 queries)
            ptx::umma_bf16_ss(tmem_base + COL_DV, ptx::make_smem_desc(aP + k * 2048, 16384, 1024),
                              ptx::make_smem_desc(aDO + k * 2048, 8192, 1024), idesc_t, (n > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < BM / 16; ++k)             // dK += dS^T Q_n
            ptx::umma_bf16_ss(tmem_base + COL_DK, ptx::make_smem_desc(aDS + k * 2048, 16384, 1024),
                              ptx::make_smem_desc(aQ + k * 2048, 8192, 1024), idesc_t, (n > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < BN / 16; ++k)             // dQ_n = dS K        (K = 128 keys: two 64-key halves of the dS tile)
            ptx::umma_bf16_ss(tmem_base + COL_DQ, ptx::make_smem_desc(aDS + (k >> 2) * (2 * TILE / 2) + (k & 3) * 32, 16, 1024),
                              ptx::make_smem_desc(aK + k * 2048, 8192, 1024), idesc_q, k > 0 ? 1u : 0u);
          ptx::umma_commit(mma3_done);
          ptx::umma_commit(qdo_empty + s);
        }
      }
      __syncwarp();
    } else if (warp == 18) {
      // ---------------------------------------------------------------- dQ reducer
      if (ptx::elect_one()) {
        for (int n = 0; n < ntiles; ++n) {
          ptx::mbar_wait(dqs_full, n & 1);              // all 16 warps have written dQ_n to the staging tile
          const int row0 = q_row0 + (it0 + n) * BM;
          ptx::tma_reduce_add_2d(&tmDQ, smem + OFF_DQ, h * DH, row0);
          ptx::tma_reduce_add_2d(&tmDQ, smem + OFF_DQ + TILE, h * DH + 32, row0);
          ptx::tma_commit_group();
          ptx::tma_wait_group_read0();                  // the staging tile has been read
          ptx::mbar_arrive(dqs_free);
        }
        ptx::tma_wait_group0();                         // reductions performed before the CTA retires
      }
      __syncwarp();
    } else if (p.cu_q && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
      // warp 19 of the first CTA: the ghost rows past the last sequence of a packed batch belong to no key tile; the
      // weight-gradient GEMM sums over all rows, so their dK / dV must be zero
      const int first = p.cu_k[p.B];
      for (long long i = (long long)first * (p.H * DH / 8) + lane; i < p.k_rows * (p.H * DH / 8); i += 32) {
        const long long row = i / (p.H * DH / 8);
        const int c8 = (int)(i - row * (p.H * DH / 8));
        *reinterpret_cast<uint4*>(p.dk + row * p.lddk + c8 * 8) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(p.dv + row * p.lddv + c8 * 8) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  } else {
    // ------------------------------------------------------------------ element-wise threads
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int g = warp >> 2;                          // key columns [32g, 32g+32) of the tile
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                // query row within the tile = TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t rx = (uint32_t)(r & 7);
    // this thread's 32 keys = four 16-byte chunks of row r in the 64-key half (g >> 1) of the Pd / dS tiles
    const uint32_t row_off = (uint32_t)(g >> 1) * TILE + (uint32_t)r * 128;
    const uint32_t aP_row = ptx::smem_u32(smem + OFF_P) + row_off, aDS_row = ptx::smem_u32(smem + OFF_DS) + row_off;
    const uint32_t c16_0 = (uint32_t)(g & 1) * 4;
    // dQ staging: column half (g >> 1) holds dQ columns [32 (g >> 1), +32) as 128-byte rows; this thread's 16 columns are
    // the 16-byte chunks 4 (g & 1) .. +3 of row r
    const uint32_t aDQ_row = ptx::smem_u32(smem + OFF_DQ) + (uint32_t)(g >> 1) * TILE + (uint32_t)r * 128;
    // key mask of this thread's 32 keys: padding / beyond kend (constant over the query tiles)
    uint32_t kmask = 0u;
    {
      const int jj = j0 + g * 32 + lane;
      const bool msk = jj >= kend || (read_pad && jj < p.Lk && p.pad[(long long)b * p.Lk + jj]);
      kmask = __ballot_sync(0xffffffffu, msk);
    }
    const float c2 = p.c_log2;
    const uint32_t sitekey = DROP ? attn_site_key(eff_seed(p.seed, p.seed_dev), p.site) : 0u;
    // per-row constants of a query tile are fetched one tile ahead (their global-load latency is off the critical path)
    float lse_nx = -INFINITY, dsum_nx = 0.f;
    if (ntiles > 0 && it0 * BM + r < lq) {
      lse_nx = p.lse[rowbase + it0 * BM + r];
      dsum_nx = p.dsum[rowbase + it0 * BM + r];
    }
    for (int n = 0; n < ntiles; ++n) {
      const int iq0 = (it0 + n) * BM;
      const int i = iq0 + r;
      const bool row_ok = i < lq;
      // invalid rows: lse = -inf -> P = exp2(-inf) = 0
      const float lse2 = (!row_ok || lse_nx == -INFINITY) ? INFINITY : lse_nx * 1.4426950408889634f;
      const float dsum = row_ok ? dsum_nx : 0.f;
      if (n + 1 < ntiles && i + BM < lq) {
        lse_nx = p.lse[rowbase + i + BM];
        dsum_nx = p.dsum[rowbase + i + BM];
      }
      const int ii = row_ok ? i : lq - 1;
      const uint32_t rowkey = DROP ? attn_row_key(sitekey, rowbase + ii) : 0u;
      uint32_t mw = kmask;
      if (p.causal && j0 + g * 32 + 31 > iq0) {
        const int nvis = i - (j0 + g * 32) + 1;
        mw |= nvis <= 0 ? 0xffffffffu : (nvis >= 32 ? 0u : (0xffffffffu << nvis));
      }
      const bool masked = __any_sync(0xffffffffu, mw != 0u);
      ptx::mbar_wait(sdp_full, n & 1);
      ptx::tc_fence_after();
      uint32_t sv[32], dv[32];
      ptx::tmem_ld_32x32(lane_base + COL_S + g * 32, sv);
      ptx::tmem_ld_32x32(lane_base + COL_DP + g * 32, dv);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_relaxed(sdp_free);   // S / dP may be overwritten by the next tile's MMAs
      if (masked) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if ((mw >> k) & 1u) sv[k] = 0xff800000u;     // -inf -> P = 0
      }
      const f32x2 c2p = pack2(c2, c2), nl2 = pack2(-lse2, -lse2), ik2 = pack2(p.inv_keep, p.inv_keep), nd2 = pack2(-dsum, -dsum);
      uint32_t pd[16], ds[16];
#pragma unroll
      for (int c16 = 0; c16 < 2; ++c16) {
        uint32_t x[8];
        if (DROP) attn_pair_words(attn_block_word(rowkey, (uint32_t)((j0 + g * 32) >> 4) + c16), x);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = c16 * 16 + 2 * k;
          float e0, e1;
          unpack2(fma2(pack2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), c2p, nl2), e0, e1);
          const f32x2 pp = pack2(ex2(e0), ex2(e1));
          const f32x2 dd = pack2(__uint_as_float(dv[e]), __uint_as_float(dv[e + 1]));
          float a0, a1, b0, b1;
          if (DROP) {
            // kf = keep/(1-p): one packed compare decides both keys; one select per key serves Pd (dV operand) and dP*kf
            bool k0, k1;
            attn_keep_pred2(x[k], p.thr2, k0, k1);
            const f32x2 kf = pack2(k0 ? p.inv_keep : 0.f, k1 ? p.inv_keep : 0.f);
            unpack2(mul2(pp, kf), a0, a1);
            unpack2(mul2(pp, fma2(dd, kf, nd2)), b0, b1);           // dS = P (dP keep/(1-p) - D)
          } else {
            unpack2(pp, a0, a1);
            unpack2(mul2(pp, add2(dd, nd2)), b0, b1);
          }
          pd[e >> 1] = pack_bf16x2(a0, a1);
          ds[e >> 1] = pack_bf16x2(b0, b1);
        }
      }
      (void)ik2;
      // ---- the previous tile's MMAs are done: its dQ can be drained, and Pd / dS shared memory may be rewritten
      if (n > 0) {
        // dQ_{n-1}: TMEM -> swizzled fp32 staging tile (this warpgroup's 16 columns of row r), reduced into the
        // accumulation buffer by warp 18's bulk reduction
        ptx::mbar_wait(mma3_done, (n - 1) & 1);
        ptx::tc_fence_after();
        uint32_t dqv[16];
        ptx::tmem_ld_32x16(lane_base + COL_DQ + g * 16, dqv);
        ptx::tmem_ld_wait(dqv);
        if (n > 1) ptx::mbar_wait(dqs_free, (n - 2) & 1);          // the reducer has read dQ_{n-2} out of the staging tile
#pragma unroll
        for (int c = 0; c < 4; ++c)
          st_shared_v4(aDQ_row + ((((uint32_t)(g & 1) * 4 + c) ^ rx) << 4), dqv[c * 4], dqv[c * 4 + 1], dqv[c * 4 + 2], dqv[c * 4 + 3]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t off = ((c16_0 + (uint32_t)c) ^ rx) << 4;
        st_shared_v4(aP_row + off, pd[c * 4], pd[c * 4 + 1], pd[c * 4 + 2], pd[c * 4 + 3]);
        st_shared_v4(aDS_row + off, ds[c * 4], ds[c * 4 + 1], ds[c * 4 + 2], ds[c * 4 + 3]);
      }
      ptx::fence_proxy_async();                        // generic-proxy smem writes -> visible to the tensor core / TMA
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (n > 0) ptx::mbar_arrive(dqs_full);
        ptx::mbar_arrive(pds_full);
      }
    }
    // ---- last tile's dQ, then dK / dV of this CTA's 128 keys
    if (ntiles > 0) {
      ptx::mbar_wait(mma3_done, (ntiles - 1) & 1);
      ptx::tc_fence_after();
      uint32_t dqv[16];
      ptx::tmem_ld_32x16(lane_base + COL_DQ + g * 16, dqv);
      ptx::tmem_ld_wait(dqv);
      if (ntiles > 1) ptx::mbar_wait(dqs_free, (ntiles - 2) & 1);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        st_shared_v4(aDQ_row + ((((uint32_t)(g & 1) * 4 + c) ^ rx) << 4), dqv[c * 4], dqv[c * 4 + 1], dqv[c * 4 + 2], dqv[c * 4 + 3]);
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(dqs_full);
    }
    const int j = j0 + r;                              // TMEM lane = key row for dK / dV
    const bool key_ok = j < lk;
#pragma unroll
    for (int part = 0; part < 2; ++part) {             // 0: dK (scaled), 1: dV; this warpgroup's 16 columns
      bf16* drow = (part == 0 ? p.dk + ((long long)k_row0 + j) * p.lddk : p.dv + ((long long)k_row0 + j) * p.lddv) + h * DH + g * 16;
      const float sc = part == 0 ? p.scale : 1.f;       // dV's operand Pd already carries keep/(1-p)
      uint32_t v[16];
      if (ntiles > 0) {                                  // uniform over the CTA
        ptx::tmem_ld_32x16(lane_base + (part == 0 ? COL_DK : COL_DV) + g * 16, v);
        ptx::tmem_ld_wait(v);
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = 0u;
      }
      float f[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) f[k] = key_ok ? __uint_as_float(v[k]) * sc : 0.f;
      if (key_ok) {
#pragma unroll
        for (int k = 0; k < 16; k += 8) {
          uint4 u;
          u.x = pack_bf16x2(f[k], f[k + 1]);
          u.y = pack_bf16x2(f[k + 2], f[k + 3]);
          u.z = pack_bf16x2(f[k + 4], f[k + 5]);
          u.w = pack_bf16x2(f[k + 6], f[k + 7]);
          *reinterpret_cast<uint4*>(drow + k) = u;
        }
      }
      float* db = part == 0 ? p.dbk : p.dbv;
      if (db) {                                          // bias gradient of the K / V projection: column sums over the 32 rows of this warp
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          float sum = warp_sum(f[k]);
          if (lane == 0) atomicAdd(db + h * DH + g * 16 + k, sum);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 17) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

// D[b,h,i] = sum_c dO[i, h*64 + c] * O[i, h*64 + c]: 8 threads per (row, head), 16-byte loads
__global__ void __launch_bounds__(256)
attn_dsum_kernel(const bf16* __restrict__ o, long long ldo, const bf16* __restrict__ dout, long long lddo,
                 float* __restrict__ dsum, long long rows, int H, int Lq, const int* __restrict__ cu_q, int B,
                 float* __restrict__ dq_acc) {
  const long long t = blockIdx.x * 256ll + threadIdx.x;
  const long long grp = t >> 3;                       // (row, head)
  const int sub = (int)(t & 7);
  const long long total = rows * H;
  float acc = 0.f;
  if (grp < total) {
    const long long row = grp / H;
    const int hh = (int)(grp % H);
    // this thread's 8 columns of the fp32 dQ accumulation buffer start the backward at zero (saves a separate memset pass)
    float4* z = reinterpret_cast<float4*>(dq_acc + row * ((long long)H * DH) + hh * DH + sub * 8);
    z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint4 ov = *reinterpret_cast<const uint4*>(o + row * ldo + hh * DH + sub * 8);
    const uint4 gv = *reinterpret_cast<const uint4*>(dout + row * lddo + hh * DH + sub * 8);
    f32x2 a2 = pack2(0.f, 0.f);
    a2 = fma2(bf2_to_f2(ov.x), bf2_to_f2(gv.x), a2);
    a2 = fma2(bf2_to_f2(ov.y), bf2_to_f2(gv.y), a2);
    a2 = fma2(bf2_to_f2(ov.z), bf2_to_f2(gv.z), a2);
    a2 = fma2(bf2_to_f2(ov.w), bf2_to_f2(gv.w), a2);
    float x0, x1;
    unpack2(a2, x0, x1);
    acc = x0 + x1;
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (grp < total && sub == 0) {
    const long long row = grp / H;
    const int hh = (int)(grp % H);
    if (cu_q) {                                       // packed rows: head-major [H, cu_q[B]] (rows past cu_q[B] belong to nobody)
      const long long tq = cu_q[B];
      if (row < tq) dsum[hh * tq + row] = acc;
    } else {
      const long long bb = row / Lq, i = row % Lq;
      dsum[(bb * H + hh) * Lq + i] = acc;
    }
  }
}

// dq (bf16) = scale * dq_acc (fp32), plus the column sums of dq (the Q-projection's bias gradient).
// Block = 64 rows x 256 columns slab; thread = 8 consecutive columns, walking the rows.
__global__ void __launch_bounds__(256)
attn_dq_finish_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long lddq, float* __restrict__ dbq,
                      long long rows, int cols, float scale) {
  const int cg = threadIdx.x & 31;                    // column group (8 columns) within the 256-column slab
  const int rsub = threadIdx.x >> 5;                  // 8 row lanes
  const int c0 = blockIdx.y * 256 + cg * 8;
  const long long r0 = blockIdx.x * 64ll;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < cols) {
    for (int rr = rsub; rr < 64; rr += 8) {
      const long long row = r0 + rr;
      if (row >= rows) break;
      const float4 a = *reinterpret_cast<const float4*>(acc + row * cols + c0);
      const float4 b2 = *reinterpret_cast<const float4*>(acc + row * cols + c0 + 4);
      const float f[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b2.x * scale, b2.y * scale, b2.z * scale, b2.w * scale};
      uint4 u;
      u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(dq + row * lddq + c0) = u;
#pragma unroll
      for (int k = 0; k < 8; ++k) cs[k] += f[k];
    }
  }
  if (dbq) {
    __shared__ float red[8][256];
#pragma unroll
    for (int k = 0; k < 8; ++k) red[rsub][cg * 8 + k] = cs[k];
    __syncthreads();
    const int c = threadIdx.x;                        // one column per thread
    if (blockIdx.y * 256 + c < cols) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += red[q][c];
      atomicAdd(dbq + blockIdx.y * 256 + c, s);
    }
  }
}

}  // namespace

// smer_attn_bwd_tc dispatches here (attn_tc.cu) after its argument checks
int smer_attn_bwd2_launch(const smer_attn_args* a, void* stream) {
  int rc;
  cudaStream_t st = (cudaStream_t)stream;
  SMER_CHECK_ARG(a->dq_accum != nullptr, "smer_attn_bwd_tc: dq_accum (fp32 [B*Lq, H*64] workspace) is required");
  const long long dcols = (long long)a->H * DH;
  const long long rq = a->cu_q ? a->q_rows : (long long)a->B * a->Lq, rk = a->cu_q ? a->k_rows : (long long)a->B * a->Lk;
  Params p;
  p.dk = (bf16*)a->dk; p.dv = (bf16*)a->dv; p.lddk = a->lddk; p.lddv = a->lddv;
  p.dq_acc = (float*)a->dq_accum;
  p.lse = a->lse; p.dsum = a->dsum; p.dbk = a->dbk; p.dbv = a->dbv;
  p.kv_len = a->kv_len; p.pad = a->key_pad;
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk;
  p.c_log2 = a->scale * 1.4426950408889634f; p.scale = a->scale; p.causal = a->causal;
  p.thr2 = a->dropout_p > 0.f ? attn_dropout_threshold(a->dropout_p) * 0x10001u : 0u;
  p.inv_keep = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f;
  p.seed = a->seed; p.site = a->site; p.seed_dev = smer_seed_dev();
  p.cu_q = a->cu_q; p.cu_k = a->cu_k; p.k_rows = a->k_rows;
  static int attr_dev_mask = 0;
  int dev = 0;
  SMER_CUDA(cudaGetDevice(&dev));
  if (!(attr_dev_mask & (1 << dev))) {
    SMER_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SMER_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_dev_mask |= 1 << dev;
  }
  CUtensorMap tq, tdo, tk, tv, tdq;
  if ((rc = smer_make_tmap_f32(&tdq, a->dq_accum, dcols, rq, dcols, 32, BM))) return rc;
  if ((rc = smer_make_tmap_bf16(&tq, a->q, dcols, rq, a->ldq, DH, BM))) return rc;
  if ((rc = smer_make_tmap_bf16(&tdo, a->dout, dcols, rq, a->lddo, DH, BM))) return rc;
  if ((rc = smer_make_tmap_bf16(&tk, a->k, dcols, rk, a->ldk, DH, BN))) return rc;
  if ((rc = smer_make_tmap_bf16(&tv, a->v, dcols, rk, a->ldv, DH, BN))) return rc;
  {
    const long long threads = rq * a->H * 8;
    attn_dsum_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>((const bf16*)a->o, a->ldo, (const bf16*)a->dout, a->lddo,
                                                                        a->dsum, rq, a->H, a->Lq, a->cu_q, a->B, (float*)a->dq_accum);
  }
  dim3 grid((a->Lk + BN - 1) / BN, a->H, a->B);
  if (p.thr2) attn_bwd2_kernel<true><<<grid, THREADS, SMEM_BYTES, st>>>(tq, tdo, tk, tv, tdq, p);
  else attn_bwd2_kernel<false><<<grid, THREADS, SMEM_BYTES, st>>>(tq, tdo, tk, tv, tdq, p);
  {
    dim3 g2((unsigned)((rq + 63) / 64), (unsigned)((dcols + 255) / 256));
    attn_dq_finish_kernel<<<g2, 256, 0, st>>>((const float*)a->dq_accum, (bf16*)a->dq, a->lddq, a->dbq, rq, (int)dcols, a->scale);
  }
  SMER_CHECK_LAUNCH("smer_attn_bwd_tc(v2)");
  return SMER_OK;
}
