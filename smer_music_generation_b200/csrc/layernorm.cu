// Fused residual-add (+dropout on the branch) + LayerNorm, forward and backward.
// Reference: transformer.py:391-392,394-395 (encoder), 461-462,465-466,468-469 (decoder):
//     x = norm(x + dropout(branch)),  nn.LayerNorm(d), eps 1e-5, affine.
// HBM-bound: one warp per row, 16-byte (fp32) / 8-byte (bf16) vector accesses, the row is
// held in registers between the statistics pass and the normalisation pass.
#include "common.cuh"
#include "../../include/smer_b200.h"

constexpr int LN_MAX_ITERS = 8;   // d <= 8 * 128 = 1024

// VEC consecutive elements per lane per iteration: 4 for fp32 (16-byte accesses), 8 for bf16 (16 bytes)
// raw 16-byte row chunks (8 bf16 or 4 fp32): loaded one row ahead of the arithmetic
__device__ __forceinline__ uint4 ld_raw16(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void cvt_raw(const uint4& t, float (&v)[8]) {        // bf16 x 8
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(h[k]);
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
}
__device__ __forceinline__ void cvt_raw(const uint4& t, float (&v)[4]) {        // fp32 x 4
  v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
}

template <int VEC> struct VecIO;
template <> struct VecIO<4> {
  template <typename T> static __device__ __forceinline__ void ld(const T* p, float (&v)[4]) { load4(p, v); }
  template <typename T> static __device__ __forceinline__ void st(T* p, const float (&v)[4]) { store4(p, v); }
  static __device__ __forceinline__ void drop(uint64_t seed, uint64_t site, uint64_t idx, uint32_t thr, float ik, float (&m)[4]) {
    dropout4(seed, site, idx, thr, ik, m);
  }
};
template <> struct VecIO<8> {
  static __device__ __forceinline__ void ld(const bf16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(h[k]);
      v[2 * k] = f.x;
      v[2 * k + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void ld(const float* p, float (&v)[8]) {
    float a[4], b[4];
    load4(p, a); load4(p + 4, b);
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = a[k]; v[4 + k] = b[k]; }
  }
  static __device__ __forceinline__ void st(bf16* p, const float (&v)[8]) {
    uint4 t;
    t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]); t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  // same element -> mask mapping as two consecutive 4-wide groups, so VEC does not change results
  static __device__ __forceinline__ void drop(uint64_t seed, uint64_t site, uint64_t idx8, uint32_t thr, float ik, float (&m)[8]) {
    float a[4], b[4];
    dropout4(seed, site, 2 * idx8, thr, ik, a);
    dropout4(seed, site, 2 * idx8 + 1, thr, ik, b);
#pragma unroll
    for (int k = 0; k < 4; ++k) { m[k] = a[k]; m[4 + k] = b[k]; }
  }
};

template <typename T, int ITERS, int VEC>
__global__ void __launch_bounds__(256, 3)
ln_fwd_kernel(const T* __restrict__ branch, const T* __restrict__ resid, const float* __restrict__ gamma,
              const float* __restrict__ beta, T* __restrict__ z_out, T* __restrict__ y, float* __restrict__ mean,
              float* __restrict__ rstd, long long rows, int d, float eps, uint32_t thr, float inv_keep,
              uint64_t seed, uint64_t site, const unsigned long long* seed_dev) {
  seed = eff_seed(seed, seed_dev);
  int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  int d4 = d / VEC;
  // software pipeline: the 16-byte chunks of row r + nwarps are in flight while row r is reduced,
  // normalised and stored (one warp would otherwise alternate between "all loads" and "no loads")
  uint4 rb[ITERS], rr[ITERS];
  if (warp < rows) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      int c4 = lane + it * 32;
      if (c4 < d4) {
        rb[it] = ld_raw16(branch + warp * d + c4 * VEC);
        if (resid) rr[it] = ld_raw16(resid + warp * d + c4 * VEC);
      }
    }
  }
  for (long long row = warp; row < rows; row += nwarps) {
    uint4 nb[ITERS], nr[ITERS];
    const long long nxt = row + nwarps;
    if (nxt < rows) {
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        int c4 = lane + it * 32;
        if (c4 < d4) {
          nb[it] = ld_raw16(branch + nxt * d + c4 * VEC);
          if (resid) nr[it] = ld_raw16(resid + nxt * d + c4 * VEC);
        }
      }
    }
    float v[ITERS][VEC];
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      int c4 = lane + it * 32;
      if (c4 < d4) {
        long long off = row * d + c4 * VEC;
        cvt_raw(rb[it], v[it]);
        if (thr) {
          float m[VEC];
          VecIO<VEC>::drop(seed, site, (uint64_t)(row * d4 + c4), thr, inv_keep, m);
#pragma unroll
          for (int k = 0; k < VEC; ++k) v[it][k] *= m[k];
        }
        if (resid) {
          float r[VEC];
          cvt_raw(rr[it], r);
#pragma unroll
          for (int k = 0; k < VEC; ++k) v[it][k] += r[k];
        }
        if (z_out) {
          VecIO<VEC>::st(z_out + off, v[it]);
          // statistics are taken on the values as stored so that backward (which re-reads z)
          // sees exactly the same normalisation input
#pragma unroll
          for (int k = 0; k < VEC; ++k) v[it][k] = to_f32(from_f32<T>(v[it][k]));
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) s += v[it][k];
      }
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      rb[it] = nb[it];
      rr[it] = nr[it];
    }
    float mu = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      int c4 = lane + it * 32;
      if (c4 < d4) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          float t = v[it][k] - mu;
          q += t * t;
        }
      }
    }
    float rs = rsqrtf(warp_sum(q) / d + eps);
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      int c4 = lane + it * 32;
      if (c4 < d4) {
        float g[VEC], b[VEC], o[VEC];
        VecIO<VEC>::ld(gamma + c4 * VEC, g);
        VecIO<VEC>::ld(beta + c4 * VEC, b);
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = (v[it][k] - mu) * rs * g[k] + b[k];
        VecIO<VEC>::st(y + row * d + c4 * VEC, o);
      }
    }
  }
}

// dz = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dgamma += dy*xhat; dbeta += dy.
// d_branch = dz * dropmask (written only when dropout is on; otherwise the caller aliases dz).
// dbias (optional) += column sums of d_branch: the bias gradient of the nn.Linear that produced
// the branch (out_proj / linear2), so no separate pass over d_branch is needed for it.
template <typename T, int ITERS, int VEC>
__global__ void __launch_bounds__(256, 2)
ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ z, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, T* __restrict__ dz,
              T* __restrict__ dbranch, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dbias, long long rows, int d,
              uint32_t thr, float inv_keep, uint64_t seed, uint64_t site, const unsigned long long* seed_dev) {
  seed = eff_seed(seed, seed_dev);
  extern __shared__ float sm[];   // [3][warps][d]
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  long long warp = (long long)blockIdx.x * nw + wib;
  long long nwarps = (long long)gridDim.x * nw;
  int d4 = d / VEC;
  float ag[ITERS][VEC], ab[ITERS][VEC], ad[ITERS][VEC], g[ITERS][VEC];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    int c4 = lane + it * 32;
#pragma unroll
    for (int k = 0; k < VEC; ++k) ag[it][k] = ab[it][k] = ad[it][k] = 0.f;
    if (c4 < d4) VecIO<VEC>::ld(gamma + c4 * VEC, g[it]);
  }
  for (long long row = warp; row < rows; row += nwarps) {
    float mu = mean[row], rs = rstd[row];
    float xh[ITERS][VEC], gy[ITERS][VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      int c4 = lane + it * 32;
      if (c4 < d4) {
        float a[VEC], zz[VEC];
        VecIO<VEC>::ld(dy + row * d + c4 * VEC, a);
        VecIO<VEC>::ld(z + row * d + c4 * VEC, zz);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          xh[it][k] = (zz[k] - mu) * rs;
          ab[it][k] += a[k];
          ag[it][k] += a[k] * xh[it][k];
          gy[it][k] = a[k] * g[it][k];
          s1 += gy[it][k];
          s2 += gy[it][k] * xh[it][k];
        }
      }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      int c4 = lane + it * 32;
      if (c4 < d4) {
        float o[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = rs * (gy[it][k] - s1 - xh[it][k] * s2);
        VecIO<VEC>::st(dz + row * d + c4 * VEC, o);
        if (dbranch) {
          if (thr) {
            float m[VEC];
            VecIO<VEC>::drop(seed, site, (uint64_t)(row * d4 + c4), thr, inv_keep, m);
#pragma unroll
            for (int k = 0; k < VEC; ++k) o[k] *= m[k];
          }
          VecIO<VEC>::st(dbranch + row * d + c4 * VEC, o);
        }
        if (dbias) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) ad[it][k] += o[k];
        }
      }
    }
  }
  // block reduction of the column sums, then one atomic per column per block
  float* sg = sm;
  float* sb = sm + nw * d;
  float* sd = sm + 2 * nw * d;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    int c4 = lane + it * 32;
    if (c4 < d4) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        sg[wib * d + c4 * VEC + k] = ag[it][k];
        sb[wib * d + c4 * VEC + k] = ab[it][k];
        if (dbias) sd[wib * d + c4 * VEC + k] = ad[it][k];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float tg = 0.f, tb = 0.f, td = 0.f;
    for (int w = 0; w < nw; ++w) {
      tg += sg[w * d + c];
      tb += sb[w * d + c];
      if (dbias) td += sd[w * d + c];
    }
    atomicAdd(dgamma + c, tg);
    atomicAdd(dbeta + c, tb);
    if (dbias) atomicAdd(dbias + c, td);
  }
}


// ---------------------------------------------------------------------------------------
// Fast path: bf16, d == ITERS * 256 (512, 768, 1024).  No per-chunk predicates, gamma/beta and the
// dropout key hoisted out of the row loop, packed fp32 math (FFMA2/FADD2/FMUL2), rows software-
// pipelined one ahead.  ncu on the generic kernel showed ~35 instructions per element with 45 %
// issue utilisation at 3 TB/s: the LayerNorm kernels were instruction-bound, not HBM-bound.
// Results are bit-identical to the generic kernel's element -> dropout-mask mapping.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& t, f32x2 (&v)[4]) {
  v[0] = bf2_to_f2(t.x); v[1] = bf2_to_f2(t.y); v[2] = bf2_to_f2(t.z); v[3] = bf2_to_f2(t.w);
}
__device__ __forceinline__ uint32_t f2_to_bf2(f32x2 v) {
  float a, b;
  unpack2(v, a, b);
  return pack_bf16x2(a, b);
}
__device__ __forceinline__ uint4 pack8(const f32x2 (&v)[4]) {
  uint4 t;
  t.x = f2_to_bf2(v[0]); t.y = f2_to_bf2(v[1]); t.z = f2_to_bf2(v[2]); t.w = f2_to_bf2(v[3]);
  return t;
}
__device__ __forceinline__ void mask8(uint32_t key, uint64_t idx8, uint32_t thr, float inv_keep, f32x2 (&m)[4]) {
  float a[4], b[4];
  dropout4k(key, 2 * idx8, thr, inv_keep, a);
  dropout4k(key, 2 * idx8 + 1, thr, inv_keep, b);
  m[0] = pack2(a[0], a[1]); m[1] = pack2(a[2], a[3]); m[2] = pack2(b[0], b[1]); m[3] = pack2(b[2], b[3]);
}

template <int ITERS>
__global__ void __launch_bounds__(256, 2)
ln_fwd_bf16_kernel(const bf16* __restrict__ branch, const bf16* __restrict__ resid, const float* __restrict__ gamma,
                   const float* __restrict__ beta, bf16* __restrict__ z_out, bf16* __restrict__ y,
                   float* __restrict__ mean, float* __restrict__ rstd, long long rows, float eps, uint32_t thr,
                   float inv_keep, uint64_t seed, uint64_t site, const unsigned long long* seed_dev) {
  constexpr int D = ITERS * 256;
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const uint32_t key = thr ? dropout_key(eff_seed(seed, seed_dev), site) : 0u;
  f32x2 g[ITERS][4], bt[ITERS][4];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = (lane + it * 32) * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + c), b1 = *reinterpret_cast<const float4*>(beta + c + 4);
    g[it][0] = pack2(g0.x, g0.y); g[it][1] = pack2(g0.z, g0.w); g[it][2] = pack2(g1.x, g1.y); g[it][3] = pack2(g1.z, g1.w);
    bt[it][0] = pack2(b0.x, b0.y); bt[it][1] = pack2(b0.z, b0.w); bt[it][2] = pack2(b1.x, b1.y); bt[it][3] = pack2(b1.z, b1.w);
  }
  uint4 rb[ITERS], rr[ITERS];
  if (warp < rows) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const long long off = warp * D + (lane + it * 32) * 8;
      rb[it] = ld_raw16(branch + off);
      if (resid) rr[it] = ld_raw16(resid + off);
    }
  }
  for (long long row = warp; row < rows; row += nwarps) {
    uint4 nb[ITERS], nr[ITERS];
    const long long nxt = row + nwarps;
    if (nxt < rows) {
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const long long off = nxt * D + (lane + it * 32) * 8;
        nb[it] = ld_raw16(branch + off);
        if (resid) nr[it] = ld_raw16(resid + off);
      }
    }
    f32x2 v[ITERS][4];
    f32x2 s2 = pack2(0.f, 0.f);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      unpack8(rb[it], v[it]);
      if (resid) {
        f32x2 r[4];
        unpack8(rr[it], r);
        if (thr) {
          f32x2 m[4];
          mask8(key, (uint64_t)(row * (D / 8) + lane + it * 32), thr, inv_keep, m);
#pragma unroll
          for (int k = 0; k < 4; ++k) v[it][k] = fma2(v[it][k], m[k], r[k]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) v[it][k] = add2(v[it][k], r[k]);
        }
      }
      if (z_out) {
        // statistics are taken on the values as stored (bf16), exactly what backward re-reads
        const uint4 zp = pack8(v[it]);
        *reinterpret_cast<uint4*>(z_out + row * D + (lane + it * 32) * 8) = zp;
        unpack8(zp, v[it]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) s2 = add2(s2, v[it][k]);
    }
    float sa, sb;
    unpack2(s2, sa, sb);
    const float mu = warp_sum(sa + sb) * (1.f / D);
    const f32x2 nmu = pack2(-mu, -mu);
    f32x2 q2 = pack2(0.f, 0.f);
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[it][k] = add2(v[it][k], nmu);                   // centred
        q2 = fma2(v[it][k], v[it][k], q2);
      }
    unpack2(q2, sa, sb);
    const float rs = rsqrtf(warp_sum(sa + sb) * (1.f / D) + eps);
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
    const f32x2 rs2 = pack2(rs, rs);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      f32x2 o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = fma2(mul2(v[it][k], rs2), g[it][k], bt[it][k]);
      *reinterpret_cast<uint4*>(y + row * D + (lane + it * 32) * 8) = pack8(o);
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      rb[it] = nb[it];
      rr[it] = nr[it];
    }
  }
}

template <int ITERS>
__global__ void __launch_bounds__(256, 2)
ln_bwd_bf16_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ z, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma, bf16* __restrict__ dz,
                   bf16* __restrict__ dbranch, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   float* __restrict__ dbias, long long rows, uint32_t thr, float inv_keep, uint64_t seed,
                   uint64_t site, const unsigned long long* seed_dev) {
  constexpr int D = ITERS * 256;
  extern __shared__ float sm[];   // [3][warps][D]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long warp = (long long)blockIdx.x * nw + wib;
  const long long nwarps = (long long)gridDim.x * nw;
  const uint32_t key = thr ? dropout_key(eff_seed(seed, seed_dev), site) : 0u;
  f32x2 ag[ITERS][4], ab[ITERS][4], ad[ITERS][4];
#pragma unroll
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int k = 0; k < 4; ++k) ag[it][k] = ab[it][k] = ad[it][k] = pack2(0.f, 0.f);
  uint4 ra[ITERS];                               // dy of the current row, loaded one row ahead
  if (warp < rows) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) ra[it] = ld_raw16(dy + warp * D + (lane + it * 32) * 8);
  }
  for (long long row = warp; row < rows; row += nwarps) {
    uint4 na[ITERS], rz[ITERS];
    const long long nxt = row + nwarps;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) rz[it] = ld_raw16(z + row * D + (lane + it * 32) * 8);
    if (nxt < rows) {
#pragma unroll
      for (int it = 0; it < ITERS; ++it) na[it] = ld_raw16(dy + nxt * D + (lane + it * 32) * 8);
    }
    const float mu = mean[row], rs = rstd[row];
    const f32x2 rs2 = pack2(rs, rs), nmr = pack2(-mu * rs, -mu * rs);
    f32x2 xh[ITERS][4], gy[ITERS][4];
    f32x2 s1 = pack2(0.f, 0.f), s2 = pack2(0.f, 0.f);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      f32x2 a[4], g[4];
      unpack8(ra[it], a);
      unpack8(rz[it], xh[it]);
      {
        const int c = (lane + it * 32) * 8;                // gamma stays L1-resident: cheaper than 16 live registers
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
        g[0] = pack2(g0.x, g0.y); g[1] = pack2(g0.z, g0.w); g[2] = pack2(g1.x, g1.y); g[3] = pack2(g1.z, g1.w);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xh[it][k] = fma2(xh[it][k], rs2, nmr);            // (z - mu) * rstd
        ab[it][k] = add2(ab[it][k], a[k]);
        ag[it][k] = fma2(a[k], xh[it][k], ag[it][k]);
        gy[it][k] = mul2(a[k], g[k]);
        s1 = add2(s1, gy[it][k]);
        s2 = fma2(gy[it][k], xh[it][k], s2);
      }
    }
    float u0, u1;
    unpack2(s1, u0, u1);
    const float m1 = warp_sum(u0 + u1) * (1.f / D);
    unpack2(s2, u0, u1);
    const float m2 = warp_sum(u0 + u1) * (1.f / D);
    const f32x2 nm1 = pack2(-m1, -m1), nm2r = pack2(-m2 * rs, -m2 * rs);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      f32x2 o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = fma2(xh[it][k], nm2r, mul2(add2(gy[it][k], nm1), rs2));   // rstd*(gy - m1 - xhat*m2)
      const long long off = row * D + (lane + it * 32) * 8;
      *reinterpret_cast<uint4*>(dz + off) = pack8(o);
      if (dbranch) {
        if (thr) {
          f32x2 m[4];
          mask8(key, (uint64_t)(row * (D / 8) + lane + it * 32), thr, inv_keep, m);
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = mul2(o[k], m[k]);
        }
        *reinterpret_cast<uint4*>(dbranch + off) = pack8(o);
      }
      if (dbias) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ad[it][k] = add2(ad[it][k], o[k]);
      }
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) ra[it] = na[it];
  }
  // block reduction of the column sums, then one atomic per column per block
  float* sg = sm;
  float* sb = sm + nw * D;
  float* sd = sm + 2 * nw * D;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = (lane + it * 32) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unpack2(ag[it][k], sg[wib * D + c + 2 * k], sg[wib * D + c + 2 * k + 1]);
      unpack2(ab[it][k], sb[wib * D + c + 2 * k], sb[wib * D + c + 2 * k + 1]);
      if (dbias) unpack2(ad[it][k], sd[wib * D + c + 2 * k], sd[wib * D + c + 2 * k + 1]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float tg = 0.f, tb = 0.f, td = 0.f;
    for (int w = 0; w < nw; ++w) {
      tg += sg[w * D + c];
      tb += sb[w * D + c];
      if (dbias) td += sd[w * D + c];
    }
    atomicAdd(dgamma + c, tg);
    atomicAdd(dbeta + c, tb);
    if (dbias) atomicAdd(dbias + c, td);
  }
}

template <typename T>
static int ln_fwd_launch(const void* branch, const void* resid, const float* gamma, const float* beta, void* z,
                         void* y, float* mean, float* rstd, long long rows, int d, float eps, uint32_t thr,
                         float inv_keep, uint64_t seed, uint64_t site, cudaStream_t st) {
  constexpr int VEC = sizeof(T) == 2 ? 8 : 4;
  if (d % VEC) { smer_set_error("smer_layernorm_fwd: d=%d must be a multiple of %d for this dtype", d, VEC); return SMER_ERR_ARG; }
  int iters = (d / VEC + 31) / 32;
  long long blocks = (rows + 7) / 8;
  long long cap = (long long)smer_num_sms() * 8;
  int grid = (int)(blocks < cap ? blocks : cap);
  if (grid < 1) grid = 1;
  if (sizeof(T) == 2 && d % 256 == 0 && d <= 1024) {            // bf16 fast path
#define LNF(I) case I: ln_fwd_bf16_kernel<I><<<grid, 256, 0, st>>>((const bf16*)branch, (const bf16*)resid, gamma, beta, (bf16*)z, (bf16*)y, mean, rstd, rows, eps, thr, inv_keep, seed, site, smer_seed_dev()); return SMER_OK;
    switch (d / 256) { LNF(1) LNF(2) LNF(3) LNF(4) }
#undef LNF
  }
#define LN_CASE(I)                                                                                          \
  case I:                                                                                                   \
    ln_fwd_kernel<T, I, VEC><<<grid, 256, 0, st>>>((const T*)branch, (const T*)resid, gamma, beta, (T*)z, (T*)y, mean, \
                                              rstd, rows, d, eps, thr, inv_keep, seed, site, smer_seed_dev()); \
    break;
  switch (iters) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
    default:
      smer_set_error("smer_layernorm_fwd: d=%d unsupported (max 1024)", d);
      return SMER_ERR_UNSUPPORTED;
  }
#undef LN_CASE
  return SMER_OK;
}

extern "C" int smer_layernorm_fwd(const void* branch, const void* resid, const float* gamma, const float* beta,
                                  void* z_out, void* y, float* mean, float* rstd, int dtype, long long rows,
                                  int d, float eps, float dropout_p, uint64_t seed, uint64_t site, void* stream) {
  SMER_CHECK_ARG(d % 4 == 0 && d > 0 && d <= 128 * LN_MAX_ITERS, "smer_layernorm_fwd: need d%%4==0 and d<=1024 (got %d)", d);
  if (rows == 0) return SMER_OK;
  uint32_t thr = dropout_p > 0.f ? dropout_threshold(dropout_p) : 0u;
  float inv_keep = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = dtype == SMER_DT_F32
               ? ln_fwd_launch<float>(branch, resid, gamma, beta, z_out, y, mean, rstd, rows, d, eps, thr, inv_keep, seed, site, st)
               : ln_fwd_launch<bf16>(branch, resid, gamma, beta, z_out, y, mean, rstd, rows, d, eps, thr, inv_keep, seed, site, st);
  if (rc) return rc;
  SMER_CHECK_LAUNCH("smer_layernorm_fwd");
  return SMER_OK;
}

template <typename T>
static int ln_bwd_launch(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma,
                         void* dz, void* dbranch, float* dgamma, float* dbeta, float* dbias, long long rows, int d, uint32_t thr,
                         float inv_keep, uint64_t seed, uint64_t site, cudaStream_t st) {
  constexpr int VEC = sizeof(T) == 2 ? 8 : 4;
  if (d % VEC) { smer_set_error("smer_layernorm_bwd: d=%d must be a multiple of %d for this dtype", d, VEC); return SMER_ERR_ARG; }
  int iters = (d / VEC + 31) / 32;
  long long blocks = (rows + 7) / 8;
  long long cap = (long long)smer_num_sms() * 4;
  int grid = (int)(blocks < cap ? blocks : cap);
  if (grid < 1) grid = 1;
  size_t smem = 3 * 8 * (size_t)d * sizeof(float);
  if (sizeof(T) == 2 && d % 256 == 0 && d <= 1024) {            // bf16 fast path
#define LNB(I)                                                                                                       \
  case I:                                                                                                            \
    if (smem > 48 * 1024) cudaFuncSetAttribute(ln_bwd_bf16_kernel<I>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    ln_bwd_bf16_kernel<I><<<grid, 256, smem, st>>>((const bf16*)dy, (const bf16*)z, mean, rstd, gamma, (bf16*)dz, (bf16*)dbranch, \
                                                  dgamma, dbeta, dbias, rows, thr, inv_keep, seed, site, smer_seed_dev());       \
    return SMER_OK;
    switch (d / 256) { LNB(1) LNB(2) LNB(3) LNB(4) }
#undef LNB
  }
#define LN_CASE(I)                                                                                              \
  case I:                                                                                                       \
    if (smem > 48 * 1024)                                                                                       \
      cudaFuncSetAttribute(ln_bwd_kernel<T, I, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    ln_bwd_kernel<T, I, VEC><<<grid, 256, smem, st>>>((const T*)dy, (const T*)z, mean, rstd, gamma, (T*)dz, (T*)dbranch, \
                                                 dgamma, dbeta, dbias, rows, d, thr, inv_keep, seed, site, smer_seed_dev()); \
    break;
  switch (iters) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
    default:
      smer_set_error("smer_layernorm_bwd: d=%d unsupported (max 1024)", d);
      return SMER_ERR_UNSUPPORTED;
  }
#undef LN_CASE
  return SMER_OK;
}

extern "C" int smer_layernorm_bwd(const void* dy, const void* z, const float* mean, const float* rstd,
                                  const float* gamma, void* dz, void* dbranch, float* dgamma, float* dbeta,
                                  float* dbias, int dtype, long long rows, int d, float dropout_p, uint64_t seed,
                                  uint64_t site, void* stream) {
  SMER_CHECK_ARG(d % 4 == 0 && d > 0 && d <= 128 * LN_MAX_ITERS, "smer_layernorm_bwd: need d%%4==0 and d<=1024 (got %d)", d);
  if (rows == 0) return SMER_OK;
  uint32_t thr = dropout_p > 0.f ? dropout_threshold(dropout_p) : 0u;
  float inv_keep = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = dtype == SMER_DT_F32
               ? ln_bwd_launch<float>(dy, z, mean, rstd, gamma, dz, dbranch, dgamma, dbeta, dbias, rows, d, thr, inv_keep, seed, site, st)
               : ln_bwd_launch<bf16>(dy, z, mean, rstd, gamma, dz, dbranch, dgamma, dbeta, dbias, rows, d, thr, inv_keep, seed, site, st);
  if (rc) return rc;
  SMER_CHECK_LAUNCH("smer_layernorm_bwd");
  return SMER_OK;
}
