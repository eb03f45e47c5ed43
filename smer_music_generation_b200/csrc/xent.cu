// Class-weighted softmax cross-entropy over the SMER vocabulary, one pass instead of the
// reference's 7-12 nn.CrossEntropyLoss passes (train.py:555-642 definition, 726-780 use):
//   loss = sum_i W[y_i] * (lse(x_i) - x_i[y_i]) / sum_i C[y_i],   rows with y_i == 0 ignored.
// HBM-bound: one warp per row of V=309 logits, warp-shuffle reductions.
#include "common.cuh"
#include "../../include/smer_b200.h"

// sums[0] = sum W[y]*nll, sums[1] = sum C[y], sums[2+k] = sum over category k of W[y]*nll
__global__ void __launch_bounds__(256)
xent_fwd_kernel(const float* __restrict__ logits, long long ld, const int64_t* __restrict__ tgt,
                const float* __restrict__ W, const float* __restrict__ C, const int* __restrict__ cat,
                float* __restrict__ lse_out, double* __restrict__ sums, long long rows, int V, int ncat) {
  __shared__ double sm[SMER_XENT_MAX_SUMS];
  if (threadIdx.x < SMER_XENT_MAX_SUMS) sm[threadIdx.x] = 0.0;
  __syncthreads();
  int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* x = logits + r * ld;
    float mx = -INFINITY;
    for (int c = lane; c < V; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s += expf(x[c] - mx);
    s = warp_sum(s);
    float lse = mx + logf(s);
    if (lane == 0) {
      lse_out[r] = lse;
      long long y = tgt[r];
      if (y > 0 && y < V) {
        float nll = lse - x[y];
        float w = W[y];
        atomicAdd(&sm[0], (double)(w * nll));
        atomicAdd(&sm[1], (double)C[y]);
        int k = cat[y];
        if (k >= 0 && k < ncat) atomicAdd(&sm[2 + k], (double)(w * nll));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 + ncat && sm[threadIdx.x] != 0.0) atomicAdd(sums + threadIdx.x, sm[threadIdx.x]);
}

template <typename T>
__global__ void __launch_bounds__(256)
xent_bwd_kernel(const float* __restrict__ logits, long long ld, const int64_t* __restrict__ tgt,
                const float* __restrict__ W, const float* __restrict__ lse, const double* __restrict__ sums,
                T* __restrict__ dlogits, long long ldo, long long rows, int V, int Vpad, float gscale,
                const float* __restrict__ gscale_dev) {
  int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  if (gscale_dev) gscale *= *gscale_dev;      // upstream d(loss) kept on the device (no host sync)
  float inv_denom = (float)(1.0 / sums[1]);
  for (long long r = warp; r < rows; r += nwarps) {
    long long y = tgt[r];
    T* o = dlogits + r * ldo;
    float coef = (y > 0 && y < V) ? W[y] * inv_denom * gscale : 0.f;
    if (coef == 0.f) {
      for (int c = lane; c < Vpad; c += 32) o[c] = from_f32<T>(0.f);
      continue;
    }
    const float* x = logits + r * ld;
    float l = lse[r];
    for (int c = lane; c < Vpad; c += 32) {
      float g = 0.f;
      if (c < V) g = coef * (expf(x[c] - l) - (c == y ? 1.f : 0.f));
      o[c] = from_f32<T>(g);
    }
  }
}

extern "C" int smer_xent_fwd(const float* logits, long long ld, const int64_t* targets, const float* W,
                             const float* C, const int* category, int ncat, float* lse, double* sums,
                             long long rows, int V, void* stream) {
  SMER_CHECK_ARG(ncat >= 0 && ncat + 2 <= SMER_XENT_MAX_SUMS, "smer_xent_fwd: too many categories");
  cudaStream_t st = (cudaStream_t)stream;
  SMER_CUDA(cudaMemsetAsync(sums, 0, SMER_XENT_MAX_SUMS * sizeof(double), st));
  if (rows == 0) return SMER_OK;
  long long blocks = (rows + 7) / 8;
  long long cap = (long long)smer_num_sms() * 8;
  int grid = (int)(blocks < cap ? blocks : cap);
  xent_fwd_kernel<<<grid, 256, 0, st>>>(logits, ld, targets, W, C, category, lse, sums, rows, V, ncat);
  SMER_CHECK_LAUNCH("smer_xent_fwd");
  return SMER_OK;
}

// Loss normaliser alone: sums[1] = sum_i C[y_i] (train.py:736).  It depends on the targets only, so under data
// parallelism it is computed and all-reduced at the START of the step, off the critical path, and the loss backward
// reads the batch-global value without waiting for a collective after the forward pass.
__global__ void __launch_bounds__(256)
xent_denominator_kernel(const int64_t* __restrict__ tgt, const float* __restrict__ C, double* __restrict__ sums,
                        long long rows, int V) {
  double acc = 0.0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const long long y = tgt[r];
    if (y > 0 && y < V) acc += (double)C[y];
  }
  acc = warp_sum(acc);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sm[i];
    if (t != 0.0) atomicAdd(sums + 1, t);
  }
}

extern "C" int smer_xent_denominator(const int64_t* targets, const float* C, double* sums, long long rows, int V,
                                     void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SMER_CUDA(cudaMemsetAsync(sums, 0, SMER_XENT_MAX_SUMS * sizeof(double), st));
  if (rows == 0) return SMER_OK;
  long long blocks = (rows + 1023) / 1024;
  int grid = (int)(blocks < 64 ? blocks : 64);
  xent_denominator_kernel<<<grid, 256, 0, st>>>(targets, C, sums, rows, V);
  SMER_CHECK_LAUNCH("smer_xent_denominator");
  return SMER_OK;
}

// Token accuracy per target class (train.py:988-1034): warp per row, argmax = FIRST maximum like
// torch.argmax; counts[c] / counts[ncls + 1 + c] = correct / seen tokens of class c, index ncls = total.
__global__ void __launch_bounds__(256)
token_accuracy_kernel(const float* __restrict__ logits, long long ld, const int64_t* __restrict__ tgt,
                      const int* __restrict__ class_of, int ncls, unsigned long long* __restrict__ counts,
                      int64_t* __restrict__ argmax_out, long long rows, int V) {
  __shared__ unsigned int sm[2 * (SMER_ACC_MAX_CLASSES + 1)];
  const int nslots = 2 * (ncls + 1);
  for (int i = threadIdx.x; i < nslots; i += blockDim.x) sm[i] = 0u;
  __syncthreads();
  int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* x = logits + r * ld;
    float best = -INFINITY;
    int bi = V;                                    // NaN-free logits: some index < V always wins
    for (int c = lane; c < V; c += 32) {
      float v = x[c];
      if (v > best) { best = v; bi = c; }          // strictly greater: the lane keeps its first maximum
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, off);
      int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
      if (argmax_out) argmax_out[r] = bi;
      long long y = tgt[r];
      if (y != 0 && y > 0 && y < V) {              // pad targets are skipped (train.py:1016-1017)
        unsigned int hit = bi == (int)y ? 1u : 0u;
        int k = class_of[y];
        if (k >= 0 && k < ncls) {
          atomicAdd(&sm[k], hit);
          atomicAdd(&sm[ncls + 1 + k], 1u);
        }
        atomicAdd(&sm[ncls], hit);
        atomicAdd(&sm[2 * ncls + 1], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nslots; i += blockDim.x)
    if (sm[i]) atomicAdd(counts + i, (unsigned long long)sm[i]);
}

extern "C" int smer_token_accuracy(const float* logits, long long ld, const int64_t* targets, const int* class_of,
                                   int ncls, unsigned long long* counts, int64_t* argmax_out, long long rows, int V,
                                   void* stream) {
  SMER_CHECK_ARG(ncls >= 0 && ncls <= SMER_ACC_MAX_CLASSES, "smer_token_accuracy: at most %d classes", SMER_ACC_MAX_CLASSES);
  if (rows == 0) return SMER_OK;
  long long blocks = (rows + 7) / 8;
  long long cap = (long long)smer_num_sms() * 8;
  int grid = (int)(blocks < cap ? blocks : cap);
  token_accuracy_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, ld, targets, class_of, ncls, counts, argmax_out, rows, V);
  SMER_CHECK_LAUNCH("smer_token_accuracy");
  return SMER_OK;
}

extern "C" int smer_xent_bwd(const float* logits, long long ld, const int64_t* targets, const float* W,
                             const float* lse, const double* sums, void* dlogits, int out_dtype, long long ldo,
                             long long rows, int V, int Vpad, float grad_scale, const float* grad_scale_dev,
                             void* stream) {
  SMER_CHECK_ARG(Vpad >= V && ldo >= Vpad, "smer_xent_bwd: bad padding");
  if (rows == 0) return SMER_OK;
  cudaStream_t st = (cudaStream_t)stream;
  long long blocks = (rows + 7) / 8;
  long long cap = (long long)smer_num_sms() * 8;
  int grid = (int)(blocks < cap ? blocks : cap);
  if (out_dtype == SMER_DT_F32)
    xent_bwd_kernel<float><<<grid, 256, 0, st>>>(logits, ld, targets, W, lse, sums, (float*)dlogits, ldo, rows, V, Vpad, grad_scale, grad_scale_dev);
  else
    xent_bwd_kernel<bf16><<<grid, 256, 0, st>>>(logits, ld, targets, W, lse, sums, (bf16*)dlogits, ldo, rows, V, Vpad, grad_scale, grad_scale_dev);
  SMER_CHECK_LAUNCH("smer_xent_bwd");
  return SMER_OK;
}
