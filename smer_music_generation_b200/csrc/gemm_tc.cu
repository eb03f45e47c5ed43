// bf16 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory
// ring -> tcgen05.mma (single issuing thread, fp32 accumulators in TMEM) -> tcgen05.ld epilogue.
//     C[M,N] = epilogue( A . B^T ),   A: M x K,  B: N x K   (either may be stored transposed)
// Serves every nn.Linear product of the reference's stack (transformer.py:362-364 linear1/2,
// MultiheadAttention in/out projections, model.py:82 fc) and both backward products
// (dX = dY.W uses B MN-major; dW = dY^T.X uses A and B MN-major with split-K fp32 atomics).
//
// Persistent: one CTA per SM walks a static list of (128x128 tile, K-split) work items, N fastest
// so that concurrently running CTAs share A rows in L2.  K step 64, 6-stage smem ring (192 KB)
// that keeps filling across work items; TWO 128-column fp32 accumulators in TMEM so the epilogue
// of item i overlaps the MMAs of item i+1.  Epilogue (8 warps), two paths (row_epilogue<> below):
//   bias / ReLU / dropout / gate with bf16 output: math in the accumulator's layout (lane = row), bf16
//   pack, swizzled 32 x 64 staging box, one cp.async.bulk.tensor store per warp and round;
//   residual / fp32 split-K / generic: TMEM -> registers -> per-warp fp32 smem transpose (so that one
//   store instruction covers 4 full 128-byte row segments instead of 32 scattered rows) -> 16-byte
//   global stores / fp32 reductions; the residual operand is read with the same coalesced mapping.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..9 = epilogue (TMEM lane quarter
// = warp_idx % 4, column half = (warp_idx - 2) / 4).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/smer_b200.h"
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <string>

namespace {

constexpr int BM = 128, BK = 64;
constexpr int TILE_BYTES = BM * BK * 2;                      // 16 KB: the A tile of a stage (and 128 rows of B)
constexpr int GEMM_THREADS = 320;
constexpr int EPI_WARPS = 8;
constexpr int STAGE_EPI = 32 * 32 * 4;                        // per epilogue warp: 32 rows x 32 fp32 columns per round
// Tile width BN is 256 when the problem has enough 128x256 tiles to fill the GPU (each A byte
// fetched from L2 then feeds twice the MACs -- 128x128 tiles are L2-bandwidth-bound at ~1/3 of the
// tensor peak), else 128.  Same smem budget either way: 6 x 32 KB or 4 x 48 KB stages.
// CTAS = 2: a CTA pair (cluster of two, one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2:
// each CTA stages its own 128 rows of A and its own BN/2 columns of B, so a stage is again 32 KB
// (6 stages) while every byte fetched from L2 feeds twice the MACs of the single-CTA 128 x 256 tile.
template <int BN, int CTAS = 1> struct Cfg {
  static constexpr int BNC = BN / CTAS;                 // B columns staged by one CTA
  static constexpr int B_BYTES = BNC * BK * 2;
  static constexpr int STAGES = B_BYTES > TILE_BYTES ? 4 : 6;
  static constexpr int STAGE_BYTES = TILE_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * STAGE_EPI + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;
};

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(h[k]);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
}
__device__ __forceinline__ void cvt8(const uint4& t, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(h[k]);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 t;
  t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]); t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = t;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, float* v) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr) : "memory");
}

struct EpiParams {
  void* C;
  long long ldc;
  const float* bias;
  const void* resid;
  long long ldr;
  int M, N;
  int flags;
  uint32_t thr;
  float inv_keep;
  uint64_t seed, site;
  const unsigned long long* seed_dev;
  float* colsum;                         // nullable (GATE / generic epilogue): [N] += column sums of the stored C
  int kb_per_split, num_kb;
  int tiles_m, tiles_n, splits;          // work item w -> (split, m tile, n tile), n fastest
};

// Epilogue variants.  EPI_GENERIC reads the flags at run time; the others fix them at compile time
// (the per-element flag branches and dead operand loads were ~1/3 of the epilogue's instructions).
enum { EPI_GENERIC = 0, EPI_BIAS = 1, EPI_BIAS_RELU = 2, EPI_BIAS_RELU_DROP = 3, EPI_RESID = 4, EPI_GATE = 5, EPI_ATOMIC = 6 };

// The bias / ReLU / dropout epilogues with bf16 output need no second operand (and the gate epilogue of the FFN
// backward reads its activation row-wise, 128 contiguous bytes per lane and round), so they run in the accumulator's own
// layout (lane = row, 64 consecutive columns per round) and leave through the TMA: pack to bf16, 8 x 16-byte shared
// stores into a 128-byte-swizzled 32 x 64 box, one cp.async.bulk.tensor store per warp and round.  The generic path's
// fp32 staging + read-back + per-lane global stores kept the L1 data pipe at 67 % (+26 % for the operand fill), which
// is what held the K = 512 products at 61 % tensor-pipe activity (profiles/r02_e_gemm_qkv_ncu.txt).
template <typename TC, int EPI>
__host__ __device__ constexpr bool row_epilogue() {
  return sizeof(TC) == 2 && (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU || EPI == EPI_BIAS_RELU_DROP || EPI == EPI_GATE);
}

template <bool A_MN, bool B_MN, typename TC, int BN, int EPI, int CTAS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, EpiParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using C = Cfg<BN, CTAS>;
  constexpr int STAGES = C::STAGES;
  constexpr int B_BYTES = C::B_BYTES;
  constexpr int BNC = C::BNC;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * TILE_BYTES;
  uint8_t* sEpi = smem + STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sEpi + EPI_WARPS * STAGE_EPI);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int total = tiles_mn * p.splits;
  const int rank = CTAS == 2 ? (int)ptx::cluster_ctarank() : 0;     // 0 = leader (issues the MMAs of the pair)
  const int cid = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // work-list position of this CTA / pair
  const int nworkers = CTAS == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if (row_epilogue<TC, EPI>()) ptx::prefetch_tmap(&tmap_c);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar + s, 1);
      ptx::mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full_bar + a, 1);
      ptx::mbar_init(tmem_empty_bar + a, EPI_WARPS * CTAS);      // (leader's) drained by the epilogue warps of every CTA
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (CTAS == 2) ptx::tmem_alloc_2sm<C::TMEM_COLS>(tmem_ptr);
    else ptx::tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CTAS == 2) ptx::cluster_sync();            // the peer's barriers are initialised before anyone signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (ptx::elect_one()) {
      int it = 0;                                            // global k-block counter -> ring slot / parity
      auto tma = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
        if (CTAS == 2) ptx::tma_load_2d_2sm(dst, m, bar, c0, c1);       // completes on the leader's barrier
        else ptx::tma_load_2d(dst, m, bar, c0, c1);
      };
      for (int w = cid; w < total; w += nworkers) {
        const int sp = w / tiles_mn, rem = w - sp * tiles_mn;
        const int m0 = (rem / p.tiles_n) * (BM * CTAS) + rank * BM;        // this CTA's 128 rows of A
        const int n0 = (rem % p.tiles_n) * BN + rank * BNC;                // this CTA's BN/CTAS columns of B
        const int kb0 = sp * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          ptx::mbar_wait(empty_bar + s, ((it / STAGES) & 1) ^ 1);
          if (rank == 0) ptx::mbar_expect_tx(full_bar + s, CTAS * C::STAGE_BYTES);   // bytes of the whole pair
          const int k = kb * BK;
          uint8_t* a = sA + s * TILE_BYTES;
          uint8_t* b = sB + s * B_BYTES;
          if (!A_MN) {
            tma(a, &tmap_a, full_bar + s, k, m0);
          } else {
            tma(a, &tmap_a, full_bar + s, m0, k);
            tma(a + TILE_BYTES / 2, &tmap_a, full_bar + s, m0 + 64, k);
          }
          if (!B_MN) {
            tma(b, &tmap_b, full_bar + s, k, n0);                               // one box: BNC rows x 64 k
          } else {
#pragma unroll
            for (int c = 0; c < BNC / 64; ++c)                                  // 64-column chunks, 8 KB apart
              tma(b + c * (TILE_BYTES / 2), &tmap_b, full_bar + s, n0 + c * 64, k);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM * CTAS, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int it = 0, item = 0;
      for (int w = cid; w < total; w += nworkers, ++item) {
        const int sp = w / tiles_mn;
        const int kb0 = sp * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int acc = item & 1;
        ptx::mbar_wait(tmem_empty_bar + acc, ((item >> 1) & 1) ^ 1);     // epilogue drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          ptx::mbar_wait(full_bar + s, (it / STAGES) & 1);
          ptx::tc_fence_after();
          const uint32_t a = ptx::smem_u32(sA + s * TILE_BYTES);
          const uint32_t b = ptx::smem_u32(sB + s * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 elements = 32 B inside the 128 B swizzle row; rows 8 apart are 1024 B apart.
            // MN-major: 16 k-rows = 2048 B; 64-element MN chunks are 8192 B apart, 8-row groups 1024 B.
            const uint64_t ad = A_MN ? ptx::make_smem_desc(a + k * 2048, 8192, 1024) : ptx::make_smem_desc(a + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? ptx::make_smem_desc(b + k * 2048, 8192, 1024) : ptx::make_smem_desc(b + k * 32, 16, 1024);
            if (CTAS == 2) ptx::umma_bf16_ss_2sm(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else ptx::umma_bf16_ss(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
          if (CTAS == 2) ptx::umma_commit_2sm(empty_bar + s); else ptx::umma_commit(empty_bar + s);
        }
        // accumulator complete (signalled to the epilogue warps of both CTAs of a pair)
        if (CTAS == 2) ptx::umma_commit_2sm(tmem_full_bar + acc); else ptx::umma_commit(tmem_full_bar + acc);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    const int quarter = warp & 3;
    const bool f_atomic = EPI == EPI_GENERIC ? (p.flags & SMER_EPI_ATOMIC) != 0 : EPI == EPI_ATOMIC;
    const bool f_gate = EPI == EPI_GENERIC ? (p.flags & SMER_EPI_GATE) != 0 : EPI == EPI_GATE;
    const bool f_relu = EPI == EPI_GENERIC ? (p.flags & SMER_EPI_RELU) != 0 : (EPI == EPI_BIAS_RELU || EPI == EPI_BIAS_RELU_DROP);
    const bool f_drop = EPI == EPI_GENERIC ? (p.thr != 0u && !f_gate) : EPI == EPI_BIAS_RELU_DROP;
    const bool f_bias = EPI == EPI_GENERIC ? p.bias != nullptr : (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU || EPI == EPI_BIAS_RELU_DROP);
    const bool f_accum = EPI == EPI_GENERIC ? (p.flags & SMER_EPI_ACCUM) != 0 : false;
    const bool has_r = EPI == EPI_GENERIC ? (p.resid != nullptr && !f_atomic) : (EPI == EPI_RESID || EPI == EPI_GATE);
    const uint32_t dkey = f_drop ? dropout_key(eff_seed(p.seed, p.seed_dev), p.site) : 0u;
    const int chalf = (warp - 2) >> 2;              // which 64 of the tile's 128 columns
    const uint32_t stage_addr = ptx::smem_u32(sEpi + (warp - 2) * STAGE_EPI);
    int item = 0;
    if constexpr (row_epilogue<TC, EPI>()) {
      for (int w = cid; w < total; w += nworkers, ++item) {
        const int rem = w % tiles_mn;
        const int m0 = (rem / p.tiles_n) * (BM * CTAS) + rank * BM, n0 = (rem % p.tiles_n) * BN;
        const int acc = item & 1;
        const int row0 = m0 + quarter * 32;             // the 32 rows of this warp's TMEM lane quarter
        // gate operand (the stored activation h = dropout(relu(.)): h > 0 <=> kept and active): this lane's row, both
        // rounds, issued before the accumulator wait
        uint4 graw[EPI == EPI_GATE ? BN / 128 : 1][EPI == EPI_GATE ? 8 : 1];
        if constexpr (EPI == EPI_GATE) {
          const bool rv = row0 + lane < p.M;
          const bf16* grow = reinterpret_cast<const bf16*>(p.resid) + (long long)(row0 + lane) * p.ldr;
#pragma unroll
          for (int hp = 0; hp < BN / 128; ++hp)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int col = n0 + chalf * (BN / 2) + hp * 64 + c * 8;
              graw[hp][c] = rv && col < p.N ? __ldg(reinterpret_cast<const uint4*>(grow + col)) : make_uint4(0u, 0u, 0u, 0u);
            }
        }
        ptx::mbar_wait(tmem_full_bar + acc, (item >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t t_addr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16) + chalf * (BN / 2);
        // rounds of 64 columns = one 128-byte row per lane = one 32 x 64 TMA box (4 KB staging, SWIZZLE_128B)
#pragma unroll
        for (int hp = 0; hp < BN / 128; ++hp) {
          const int col0 = n0 + chalf * (BN / 2) + hp * 64;
          uint32_t pk[32];
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(t_addr + hp * 64 + sub * 32, r);
            float4 bv[8];                               // the same 32 bias values in every lane (uniform 16-byte loads)
#pragma unroll
            for (int u = 0; u < 8; ++u)
              bv[u] = f_bias && col0 + sub * 32 + 4 * u < p.N ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + sub * 32) + u)
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            ptx::tmem_ld_wait();
            if (hp == BN / 128 - 1 && sub == 1) {       // the whole accumulator part has left TMEM
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CTAS == 2) ptx::mbar_arrive_leader(tmem_empty_bar + acc);
                else ptx::mbar_arrive(tmem_empty_bar + acc);
              }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float v[4] = {__uint_as_float(r[4 * u]) + bv[u].x, __uint_as_float(r[4 * u + 1]) + bv[u].y,
                            __uint_as_float(r[4 * u + 2]) + bv[u].z, __uint_as_float(r[4 * u + 3]) + bv[u].w};
              if (f_relu) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
              }
              if constexpr (EPI == EPI_GATE) {
                const uint4 gq = graw[hp][sub * 4 + (u >> 1)];
                const uint32_t g0 = (u & 1) ? gq.z : gq.x, g1 = (u & 1) ? gq.w : gq.y;     // 4 bf16 of h
                v[0] = __uint_as_float(g0 << 16) > 0.f ? v[0] * p.inv_keep : 0.f;
                v[1] = __uint_as_float(g0 & 0xFFFF0000u) > 0.f ? v[1] * p.inv_keep : 0.f;
                v[2] = __uint_as_float(g1 << 16) > 0.f ? v[2] * p.inv_keep : 0.f;
                v[3] = __uint_as_float(g1 & 0xFFFF0000u) > 0.f ? v[3] * p.inv_keep : 0.f;
              }
              if (f_drop) {
                float mk[4];
                dropout4k(dkey, (uint64_t)(((long long)(row0 + lane) * p.ldc + col0 + sub * 32 + 4 * u) >> 2), p.thr,
                          p.inv_keep, mk);
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] *= mk[e];
              }
              pk[sub * 16 + 2 * u] = pack_bf16x2(v[0], v[1]);
              pk[sub * 16 + 2 * u + 1] = pack_bf16x2(v[2], v[3]);
            }
          }
          if (col0 >= p.N || row0 >= p.M) continue;     // box entirely outside C (warp-uniform)
          // the previous store of this warp has finished reading the staging buffer
          if (lane == 0) ptx::tma_wait_group_read<0>();
          __syncwarp();
          const uint32_t rowbase = stage_addr + lane * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c)                   // SWIZZLE_128B: chunk ^= row % 8
            st_shared_v4(rowbase + (((uint32_t)c ^ (lane & 7)) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmap_c, stage_addr, col0, row0);
            ptx::tma_commit_group();
          }
          if constexpr (EPI == EPI_GATE) {
            // column sums of the stored tile (the bias gradient of linear1): lane l adds up columns 2l, 2l+1 over the
            // 32 staged rows (one conflict-free 128-byte row per load; rows >= M and columns >= N hold zeros)
            if (p.colsum != nullptr) {
              float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
              for (int rr = 0; rr < 32; ++rr) {
                const uint32_t wv = ld_shared_u32(stage_addr + rr * 128 + ((((uint32_t)lane >> 2) ^ ((uint32_t)rr & 7u)) << 4) + (lane & 3) * 4);
                s0 += __uint_as_float(wv << 16);
                s1 += __uint_as_float(wv & 0xFFFF0000u);
              }
              if (col0 + 2 * lane < p.N)
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p.colsum + col0 + 2 * lane), "f"(s0), "f"(s1) : "memory");
            }
          }
        }
      }
      if (lane == 0) ptx::tma_wait_group0();            // every store has landed before the CTA's smem goes away
      __syncwarp();
    } else
    for (int w = cid; w < total; w += nworkers, ++item) {
      const int sp = w / tiles_mn, rem = w - sp * tiles_mn;
      const int m0 = (rem / p.tiles_n) * (BM * CTAS) + rank * BM, n0 = (rem % p.tiles_n) * BN;
      const int acc = item & 1;
      // bias of this warp's columns: issued before the accumulator wait so its latency is hidden
      const int lc = (lane & 3) * 8;
      float biasr[BN / 64][8];
#pragma unroll
      for (int hh = 0; hh < BN / 64; ++hh) {
        const int colb = n0 + chalf * (BN / 2) + hh * 32 + lc;
#pragma unroll
        for (int u = 0; u < 8; ++u) biasr[hh][u] = 0.f;
        if (f_bias && (!f_atomic || sp == 0) && colb < p.N) load8(p.bias + colb, biasr[hh]);
      }
      // residual / gate operands (bf16): raw 16-byte chunks of all rounds, also issued before the wait
      constexpr bool RAW_R = sizeof(TC) == 2;
      uint4 rraw[RAW_R ? BN / 64 : 1][4];
      if (RAW_R && has_r) {
#pragma unroll
        for (int hh = 0; hh < BN / 64; ++hh) {
          const int colb = n0 + chalf * (BN / 2) + hh * 32 + lc;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int row = m0 + quarter * 32 + it * 8 + (lane >> 2);
            if (row < p.M && colb < p.N)
              rraw[RAW_R ? hh : 0][it] = *reinterpret_cast<const uint4*>(reinterpret_cast<const TC*>(p.resid) + (long long)row * p.ldr + colb);
          }
        }
      }
      ptx::mbar_wait(tmem_full_bar + acc, (item >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16) + chalf * (BN / 2);
      // Rounds of 32 columns: TMEM -> registers, stage (lane == tile row; 16-byte chunk c of the
      // 128-byte row goes to chunk c ^ (row & 7)), then drain with lane -> (row = it*8 + lane/4,
      // 8 columns at (lane%4)*8): every store instruction covers 8 rows x 64 B (bf16).
#pragma unroll
      for (int hh = 0; hh < BN / 64; ++hh) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_addr + hh * 32, r);
        ptx::tmem_ld_wait();
        if (hh == BN / 64 - 1) {                               // whole accumulator part is in registers/smem
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTAS == 2) ptx::mbar_arrive_leader(tmem_empty_bar + acc);   // the MMA issuer lives in the leader CTA
            else ptx::mbar_arrive(tmem_empty_bar + acc);
          }
        }
        __syncwarp();                                          // previous round's staging reads are done
        {
          const uint32_t rowbase = stage_addr + lane * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            st_shared_v4(rowbase + (((uint32_t)c ^ (lane & 7)) << 4), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
        }
        __syncwarp();
        const int col = n0 + chalf * (BN / 2) + hh * 32 + lc;
        const uint32_t colmask = __ballot_sync(0xffffffffu, col < p.N);      // lanes that stay (whole (lane & 3) groups)
        if (col >= p.N) continue;                              // N % 8 == 0
        // residual / gate operands of the round's 4 row-iterations: issue all loads up front, otherwise
        // their latency serialises the epilogue past the MMA time of a K=512 tile
        constexpr bool CAN_CSUM = EPI == EPI_GATE || EPI == EPI_GENERIC;
        const bool f_csum = CAN_CSUM && p.colsum != nullptr;
        float csum[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) csum[u] = 0.f;
        float rsv[4][8];
        if (has_r) {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int row = m0 + quarter * 32 + it * 8 + (lane >> 2);
            if (RAW_R) cvt8(rraw[RAW_R ? hh : 0][it], rsv[it]);
            else if (row < p.M) load8(reinterpret_cast<const TC*>(p.resid) + (long long)row * p.ldr + col, rsv[it]);
          }
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rl = it * 8 + (lane >> 2);
          const int row = m0 + quarter * 32 + rl;
          if (row >= p.M) continue;
          float v[8];
          {
            const uint32_t rowbase = stage_addr + rl * 128;
            const uint32_t c0 = (uint32_t)(lc >> 2);
            ld_shared_v4(rowbase + ((c0 ^ (rl & 7)) << 4), v);
            ld_shared_v4(rowbase + (((c0 + 1) ^ (rl & 7)) << 4), v + 4);
          }
          if (f_bias) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] += biasr[hh][u];
          }
          TC* crow = reinterpret_cast<TC*>(p.C) + (long long)row * p.ldc;
          if (f_atomic) {
            float* dst = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                         "f"(v[3])
                         : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]),
                         "f"(v[7])
                         : "memory");
            continue;
          }
          if (f_relu) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = fmaxf(v[u], 0.f);
          }
          if (f_gate) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = rsv[it][u] > 0.f ? v[u] * p.inv_keep : 0.f;
          } else {
            if (f_drop) {
              float m0_[4], m1_[4];
              const uint64_t e4 = (uint64_t)(((long long)row * p.ldc + col) >> 2);
              dropout4k(dkey, e4, p.thr, p.inv_keep, m0_);
              dropout4k(dkey, e4 + 1, p.thr, p.inv_keep, m1_);
#pragma unroll
              for (int u = 0; u < 4; ++u) { v[u] *= m0_[u]; v[4 + u] *= m1_[u]; }
            }
            if (has_r) {
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] += rsv[it][u];
            }
            if (f_accum) {
              float old[8];
              load8(crow + col, old);
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] += old[u];
            }
          }
          store8(crow + col, v);
          if (f_csum) {
#pragma unroll
            for (int u = 0; u < 8; ++u) csum[u] += v[u];
          }
        }
        if (f_csum) {
          // this round's 32 rows x 8 columns per lane group: add the 8 lanes that share (lane & 3), then one
          // 32-byte reduction per group into the [N] fp32 vector
#pragma unroll
          for (int off = 4; off <= 16; off <<= 1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) csum[u] += __shfl_xor_sync(colmask, csum[u], off);
          }
          if ((lane >> 2) == 0) {
            float* dst = p.colsum + col;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(csum[0]), "f"(csum[1]),
                         "f"(csum[2]), "f"(csum[3]) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(csum[4]), "f"(csum[5]),
                         "f"(csum[6]), "f"(csum[7]) : "memory");
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CTAS == 2) ptx::cluster_sync();            // the leader's MMAs read the peer's smem until the very end
  if (warp == 1) {
    if (CTAS == 2) ptx::tmem_dealloc_2sm<C::TMEM_COLS>(tmem_base);
    else ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry point fetched at run time, no -lcuda)
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  long long inner, outer, pitch;
  int box_inner, box_outer, swizzle;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && pitch == o.pitch && box_inner == o.box_inner &&
           box_outer == o.box_outer && swizzle == o.swizzle;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&h](long long v) { h ^= std::hash<long long>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.pitch); mix(k.box_inner); mix(k.box_outer); mix(k.swizzle);
    return h;
  }
};

}  // namespace

// bf16 2-D tensor map: `inner` contiguous elements per row, `outer` rows, row pitch in elements.
int smer_make_tmap_bf16_sw(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long pitch,
                           int box_inner, int box_outer, int swizzle_bytes) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, inner, outer, pitch, box_inner, box_outer, swizzle_bytes};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return SMER_OK; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { smer_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return SMER_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (pitch * 2) % 16) {
    smer_set_error("TMA operand needs a 16-byte aligned base and pitch (ptr=%p pitch=%lld elements)", ptr, pitch);
    return SMER_ERR_ARG;
  }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    smer_set_error("cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld pitch=%lld box=%dx%d", (int)r, inner, outer,
                   pitch, box_inner, box_outer);
    return SMER_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return SMER_OK;
}

int smer_make_tmap_bf16(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long pitch,
                        int box_inner, int box_outer) {
  return smer_make_tmap_bf16_sw(out, ptr, inner, outer, pitch, box_inner, box_outer, 128);
}

// fp32 tensor map (SWIZZLE_128B, 32 floats = 128 bytes inner box): the dQ accumulation buffer of the fused attention
// backward, the target of cp.reduce.async.bulk.tensor (not cached: one map per launch)
int smer_make_tmap_f32(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long pitch,
                       int box_inner, int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { smer_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return SMER_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (pitch * 4) % 16) {
    smer_set_error("TMA operand needs a 16-byte aligned base and pitch (ptr=%p pitch=%lld elements)", ptr, pitch);
    return SMER_ERR_ARG;
  }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    smer_set_error("cuTensorMapEncodeTiled(f32) failed (%d) inner=%lld outer=%lld pitch=%lld box=%dx%d", (int)r, inner, outer,
                   pitch, box_inner, box_outer);
    return SMER_ERR_CUDA;
  }
  return SMER_OK;
}

template <bool A_MN, bool B_MN, typename TC, int BN, int EPI, int CTAS = 1>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p, dim3 grid, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_set{0ull};         // one bit per device: the attribute is per (function, device)
  auto kern = gemm_tc_kernel<A_MN, B_MN, TC, BN, EPI, CTAS>;
  int dev = 0;
  SMER_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(attr_set.load(std::memory_order_acquire) & bit)) {
    SMER_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, CTAS>::SMEM_BYTES));
    attr_set.fetch_or(bit, std::memory_order_release);
  }
  CUtensorMap tc = ta;                               // (unused by the epilogues that store through the LSU)
  if (row_epilogue<TC, EPI>()) {
    int rc = smer_make_tmap_bf16_sw(&tc, p.C, p.N, p.M, p.ldc, 64, 32, 128);
    if (rc) return rc;
  }
  if (CTAS == 1) {
    kern<<<grid, GEMM_THREADS, Cfg<BN, CTAS>::SMEM_BYTES, st>>>(ta, tb, tc, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;                       // 2 x number of CTA pairs
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg<BN, CTAS>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SMER_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, p));
  }
  return SMER_OK;
}

extern "C" int smer_gemm_bf16_tc(const void* A, long long lda, int a_kmajor, const void* B, long long ldb, int b_kmajor,
                                 void* C, long long ldc, int out_dtype, int M, int N, int K, const float* bias,
                                 const void* resid, long long ldr, int flags, float dropout_p, uint64_t seed,
                                 uint64_t site, int split_k, float* colsum, void* stream) {
  SMER_CHECK_ARG(M > 0 && N > 0 && K > 0, "smer_gemm_bf16_tc: empty problem %dx%dx%d", M, N, K);
  SMER_CHECK_ARG(!colsum || ((flags & SMER_EPI_GATE) && out_dtype == SMER_DT_BF16 && split_k <= 1),
                 "smer_gemm_bf16_tc: colsum is produced by the bf16 gate epilogue only");
  SMER_CHECK_ARG(N % 8 == 0 && ldc % 8 == 0 && (!resid || ldr % 8 == 0),
                 "smer_gemm_bf16_tc: need N%%8==0 and 8-element-aligned C/resid pitches (N=%d ldc=%lld ldr=%lld)", N, ldc, ldr);
  if (split_k < 1) split_k = 1;
  SMER_CHECK_ARG(split_k == 1 || ((flags & SMER_EPI_ATOMIC) && out_dtype == SMER_DT_F32),
                 "smer_gemm_bf16_tc: split_k needs SMER_EPI_ATOMIC and fp32 output");
  SMER_CHECK_ARG(!(flags & SMER_EPI_ATOMIC) || out_dtype == SMER_DT_F32, "smer_gemm_bf16_tc: atomic epilogue needs fp32 output");
  SMER_CHECK_ARG(!(flags & SMER_EPI_GATE) || resid, "smer_gemm_bf16_tc: gate epilogue needs the activation in `resid`");
  CUtensorMap ta, tb;
  int rc;
  const int sms = smer_num_sms();
  const int num_kb = (K + BK - 1) / BK;
  if (split_k > num_kb) split_k = num_kb;
  const int kb_per_split = (num_kb + split_k - 1) / split_k;
  split_k = (num_kb + kb_per_split - 1) / kb_per_split;             // no empty splits
  const int tiles_m = (M + BM - 1) / BM;
  // 128x256 tiles when they still fill the GPU (and N is wide enough to use them); CTA pairs
  // (256x256 tiles, cta_group::2) when there are enough of those for every pair of SMs
  static const bool allow_2sm = [] { const char* e = getenv("SMER_GEMM_2SM"); return !(e && e[0] == '0'); }();
  const bool wide = N >= 256 && (long long)tiles_m * ((N + 255) / 256) * split_k * 20 >= sms * 19;      // >= 95 % of the SMs
  // (measured on B200: a CTA of a pair stages half of B, one third less smem fill + operand traffic per MMA:
  //  1243 vs 1083 TFLOP/s at K=2048, 987 vs 915-932 at K=512 with N >= 1536; the N=512, K=512 out-projection
  //  has too few items per pair and stays on single CTAs: 700 vs 648)
  static const bool force_2sm = [] { const char* e = getenv("SMER_GEMM_2SM"); return e && e[0] == '2'; }();
  const bool pair = allow_2sm && wide && (kb_per_split >= 16 || N >= 1024 || force_2sm) &&
                    (long long)((M + 255) / 256) * ((N + 255) / 256) * split_k * 20 >= (sms / 2) * 19;
  const int bn = wide ? 256 : 128;
  if (a_kmajor) rc = smer_make_tmap_bf16(&ta, A, K, M, lda, BK, BM);
  else rc = smer_make_tmap_bf16(&ta, A, M, K, lda, 64, BK);
  if (rc) return rc;
  if (b_kmajor) rc = smer_make_tmap_bf16(&tb, B, K, N, ldb, BK, pair ? bn / 2 : bn);      // a CTA of a pair stages half of B
  else rc = smer_make_tmap_bf16(&tb, B, N, K, ldb, 64, BK);
  if (rc) return rc;
  EpiParams p;
  p.C = C; p.ldc = ldc; p.bias = bias; p.resid = resid; p.ldr = ldr; p.M = M; p.N = N; p.flags = flags;
  p.thr = (dropout_p > 0.f && !(flags & SMER_EPI_GATE)) ? dropout_threshold(dropout_p) : 0u;
  p.inv_keep = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  p.seed = seed; p.site = site; p.seed_dev = smer_seed_dev();
  p.colsum = colsum;
  p.num_kb = num_kb;
  p.kb_per_split = kb_per_split;
  p.tiles_n = (N + bn - 1) / bn;
  p.tiles_m = pair ? (M + 2 * BM - 1) / (2 * BM) : tiles_m;
  p.splits = split_k;
  const long long total = (long long)p.tiles_n * p.tiles_m * split_k;
  const long long workers = pair ? sms / 2 : sms;
  dim3 grid((unsigned)((total < workers ? total : workers) * (pair ? 2 : 1)));
  cudaStream_t st = (cudaStream_t)stream;
  const bool amn = !a_kmajor, bmn = !b_kmajor, f32 = out_dtype == SMER_DT_F32;
  // epilogue variant: the specialised ones cover the shapes the transformer stack launches
  int epi = EPI_GENERIC;
  if ((flags & SMER_EPI_ATOMIC) && !bias && flags == SMER_EPI_ATOMIC) epi = EPI_ATOMIC;
  else if ((flags & SMER_EPI_GATE) && !bias && flags == SMER_EPI_GATE) epi = EPI_GATE;
  else if (flags == 0 && resid && !bias && !p.thr) epi = EPI_RESID;
  else if (flags == SMER_EPI_RELU && bias && !resid) epi = p.thr ? EPI_BIAS_RELU_DROP : EPI_BIAS_RELU;
  else if (flags == 0 && bias && !resid && !p.thr) epi = EPI_BIAS;
  // SMER_GEMM_ROW_EPI=0: every product through the generic (LSU-store) epilogue -- an A/B and fault-isolation switch
  static const bool row_epi_on = [] { const char* e = getenv("SMER_GEMM_ROW_EPI"); return !(e && e[0] == '0'); }();
  if (!row_epi_on && (epi == EPI_BIAS || epi == EPI_BIAS_RELU || epi == EPI_BIAS_RELU_DROP || epi == EPI_GATE)) epi = EPI_GENERIC;
#define GO3(AM, BMN, T, E)                                                                          \
  (pair ? launch_gemm<AM, BMN, T, 256, E, 2>(ta, tb, p, grid, st)                                   \
        : wide ? launch_gemm<AM, BMN, T, 256, E>(ta, tb, p, grid, st) : launch_gemm<AM, BMN, T, 128, E>(ta, tb, p, grid, st))
  if (!amn && !bmn) {                       // nn.Linear forward
    if (f32) rc = epi == EPI_BIAS ? GO3(false, false, float, EPI_BIAS) : GO3(false, false, float, EPI_GENERIC);
    else if (epi == EPI_BIAS) rc = GO3(false, false, bf16, EPI_BIAS);
    else if (epi == EPI_BIAS_RELU) rc = GO3(false, false, bf16, EPI_BIAS_RELU);
    else if (epi == EPI_BIAS_RELU_DROP) rc = GO3(false, false, bf16, EPI_BIAS_RELU_DROP);
    else rc = GO3(false, false, bf16, EPI_GENERIC);
  } else if (!amn && bmn) {                 // input gradient
    if (f32) rc = GO3(false, true, float, EPI_GENERIC);
    else if (epi == EPI_RESID) rc = GO3(false, true, bf16, EPI_RESID);
    else if (epi == EPI_GATE) rc = GO3(false, true, bf16, EPI_GATE);
    else rc = GO3(false, true, bf16, EPI_GENERIC);
  } else if (amn && bmn) {                  // weight gradient
    if (f32) rc = epi == EPI_ATOMIC ? GO3(true, true, float, EPI_ATOMIC) : GO3(true, true, float, EPI_GENERIC);
    else rc = GO3(true, true, bf16, EPI_GENERIC);
  } else {
    rc = f32 ? GO3(true, false, float, EPI_GENERIC) : GO3(true, false, bf16, EPI_GENERIC);
  }
#undef GO3
  if (rc) return rc;
  SMER_CHECK_LAUNCH("smer_gemm_bf16_tc");
  return SMER_OK;
}
