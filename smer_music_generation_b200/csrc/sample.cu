// Grammar-masked sampling and span bookkeeping for batched infilling decode, on device.
// Restates generation.py:41-95 (sampling: -100 masks in float64, softmax without max shift,
// multinomial / nucleus), generation.py:547-652 (which flag set each grammar state uses and
// the <=12-draw rejection loop), 654-671 (state transitions), 673-686 (forced <eos> after a
// control token, the 100-token span cap and the dropped last element).  One CTA per piece.
// Greedy and top-k are additions (north_star): greedy = argmax of the masked distribution
// with lowest-id tie break and no rejection loop; top-k keeps the k largest and renormalises.
#include "common.cuh"
#include "../../include/smer_b200.h"

#define VMAX 320

__device__ __forceinline__ bool in_range(int i, int lo, int hi) { return i >= lo && i <= hi; }

struct FlagSet {
  bool no_pitch, no_duration, no_rest, no_whole, no_eos, no_continue, no_sep;
  int only_lo, only_hi;        // "is_<class>" range or -1
  int accept;                  // 0 none, 1 in_sep, 2 pitch, 3 pitch+duration, 4 duration, 5 not-duration
};

// generation.py:547-652, priority in_sep > in_continue > in_pitch > in_rest > first token > free
__device__ FlagSet flags_for(int st, bool first, int target, bool nwd) {
  FlagSet f = {false, false, false, false, false, false, false, -1, -1, 0};
  if (st & SMER_ST_SEP) {
    f.no_rest = f.no_sep = f.no_eos = f.no_whole = true; f.accept = 1;
  } else if (st & SMER_ST_CONTINUE) {
    f.no_rest = f.no_sep = f.no_duration = f.no_continue = f.no_eos = true; f.accept = 2;
  } else if (st & SMER_ST_PITCH) {
    f.no_rest = f.no_sep = f.no_continue = f.no_eos = true; f.no_whole = nwd; f.accept = 3;
  } else if (st & SMER_ST_REST) {
    f.no_pitch = f.no_rest = f.no_sep = f.no_continue = f.no_eos = true; f.no_whole = nwd; f.accept = 4;
  } else if (first) {
    if (target == 1) { f.only_lo = 242; f.only_hi = 251; }        // 'd' density
    else if (target == 2) { f.only_lo = 262; f.only_hi = 271; }   // 'o' occupation
    else if (target == 3) { f.only_lo = 252; f.only_hi = 261; }   // 'p' polyphony
    else if (target == 4) { f.only_lo = 296; f.only_hi = 307; }   // 't' tensile
    else { f.no_duration = true; f.accept = 5; }                  // 'r' content
  } else {
    f.no_whole = nwd;
  }
  return f;
}

__device__ __forceinline__ bool allowed(const FlagSet& f, int i) {
  if (in_range(i, 3, 145)) return false;                          // generation.py:82-84, always
  if (f.no_pitch && in_range(i, 146, 233)) return false;
  if (f.no_duration && in_range(i, 234, 238)) return false;
  if (f.no_continue && i == 241) return false;
  if (f.no_rest && i == 239) return false;
  if (f.no_sep && i == 240) return false;
  if (f.no_whole && i == 234) return false;
  if (f.no_eos && i == 1) return false;
  if (f.only_lo >= 0 && !in_range(i, f.only_lo, f.only_hi)) return false;
  return true;
}

__device__ __forceinline__ bool accepted(int mode, int i) {
  switch (mode) {
    case 1: return !(i == 239 || i == 1 || i == 234);
    case 2: return in_range(i, 146, 233);
    case 3: return in_range(i, 146, 238);
    case 4: return in_range(i, 234, 238);
    case 5: return !in_range(i, 234, 238);
    default: return true;
  }
}

__device__ __forceinline__ double uniform01(uint64_t seed, uint64_t seq, uint64_t step, uint32_t draw) {
  uint4 r = philox4x32(seed, seq, (step << 8) | draw);
  uint64_t bits = ((uint64_t)r.x << 32) | r.y;
  return ((double)(bits >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// In-place inclusive prefix sum of x[0..VMAX) by warp 0: each lane sums a block of VMAX/32 entries,
// the block totals are scanned with shuffles, then the prefixes are written back.
__device__ __forceinline__ void warp0_inclusive_scan(double* x, int tid) {
  if (tid >= 32) return;
  constexpr int PER = VMAX / 32;
  double loc[PER];
  double run = 0.0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    run += x[tid * PER + k];
    loc[k] = run;
  }
  double incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, incl, o);
    if (tid >= o) incl += up;
  }
  const double base = incl - run;
#pragma unroll
  for (int k = 0; k < PER; ++k) x[tid * PER + k] = base + loc[k];
}

__global__ void __launch_bounds__(128) sample_kernel(smer_sample_args a) {
  __shared__ double q[VMAX];
  __shared__ double red[128];
  __shared__ int order[VMAX];
  __shared__ double cdf[VMAX];
  __shared__ int chosen, shared_keep;
  int s = blockIdx.x;
  int tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (a.done && a.done[s]) return;
  const int V = a.V;
  int st = a.state ? a.state[s] : 0;
  int pos = a.cur_len ? a.cur_len[s] - 1 : 0;                 // position of the token just fed
  if (a.fed_len) {
    // catch-up step: after a control token the stream grows by two (the token and the next span's
    // m_0, generation.py:673-686); the first of them is fed without sampling.
    int fed = a.fed_len[s];
    if (fed < pos) {
      if (tid == 0) a.fed_len[s] = fed + 1;
      return;
    }
  }
  bool first = a.span_start ? (pos == a.span_start[s]) : false;
  int target = 0;
  if (a.targets) target = a.targets[(long long)s * a.max_spans + a.span_idx[s]];
  bool nwd = a.nwd ? a.nwd[s] != 0 : false;
  FlagSet f = a.raw_flags ? FlagSet{(a.raw_flags[s] & 1) != 0, (a.raw_flags[s] & 2) != 0, (a.raw_flags[s] & 4) != 0,
                                    (a.raw_flags[s] & 8) != 0, (a.raw_flags[s] & 16) != 0, (a.raw_flags[s] & 32) != 0,
                                    (a.raw_flags[s] & 64) != 0, a.raw_only_lo ? a.raw_only_lo[s] : -1,
                                    a.raw_only_hi ? a.raw_only_hi[s] : -1, 0}
                          : flags_for(st, first, target, nwd);
  const float* x = a.logits + (long long)s * a.ld;
  double t = a.temperature;
  // masked softmax in float64, no max shift (generation.py:28-30)
  double part = 0.0;
  for (int i = tid; i < V; i += 128) {
    double l = allowed(f, i) ? (double)x[i] : -100.0;
    double e = exp(l / t);
    q[i] = e;
    part += e;
  }
  red[tid] = part;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  double tot = red[0];
  __syncthreads();
  for (int i = tid; i < V; i += 128) q[i] = q[i] / tot;
  __syncthreads();
  if (a.trace_masked && pos < a.max_len)
    for (int i = tid; i < V; i += 128) a.trace_masked[((long long)s * a.max_len + pos) * V + i] = q[i];
  if (a.trace_span && pos < a.max_len && tid == 0) a.trace_span[(long long)s * a.max_len + pos] = a.span_idx ? a.span_idx[s] : 0;

  if (a.mode == SMER_SAMPLE_TOP_P || a.mode == SMER_SAMPLE_TOP_K) {
    // rank by probability, descending; ties resolved towards the higher id (what reversing an
    // ascending argsort does)
    if (a.mode == SMER_SAMPLE_TOP_P) {
      // nucleus(): probs /= (sum + 1e-5)   (generation.py:12)
      double sum = 0.0;
      for (int i = tid; i < V; i += 128) sum += q[i];
      red[tid] = sum;
      __syncthreads();
      for (int o = 64; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
      }
      double ssum = red[0] + 1e-5;
      __syncthreads();
      for (int i = tid; i < V; i += 128) q[i] = q[i] / ssum;
      __syncthreads();
    }
    for (int i = tid; i < V; i += 128) {
      double qi = q[i];
      int rank = 0;
      for (int u = 0; u < V; ++u) rank += (q[u] > qi) || (q[u] == qi && u > i);
      order[rank] = i;
    }
    __syncthreads();
    // cumulative probability in rank order (warp 0 scans), then the cut-off and renormalisation
    for (int r = tid; r < VMAX; r += 128) cdf[r] = r < V ? q[order[r]] : 0.0;
    __syncthreads();
    warp0_inclusive_scan(cdf, tid);
    __syncthreads();
    int keep;
    if (a.mode == SMER_SAMPLE_TOP_P) {
      // first rank whose cumulative sum exceeds p, inclusive (generation.py:17-22)
      if (tid == 0) shared_keep = V;
      __syncthreads();
      for (int r = tid; r < V; r += 128)
        if (cdf[r] > a.top_p && (r == 0 || !(cdf[r - 1] > a.top_p))) shared_keep = r + 1;
      __syncthreads();
      keep = shared_keep;
    } else {
      keep = a.top_k < 1 ? 1 : (a.top_k > V ? V : a.top_k);
    }
    const double ks = cdf[keep - 1];
    __syncthreads();
    for (int r = tid; r < V; r += 128) {
      const int i = order[r];
      q[i] = r < keep ? q[i] / ks : 0.0;
    }
    __syncthreads();
  }
  if (a.out_probs)
    for (int i = tid; i < V; i += 128) a.out_probs[(long long)s * V + i] = q[i];

  if (a.mode == SMER_SAMPLE_GREEDY) {
    // argmax, lowest id on ties
    double best = -1.0;
    int bi = 0;
    for (int i = tid; i < V; i += 128)
      if (q[i] > best) { best = q[i]; bi = i; }
    red[tid] = best;
    order[tid] = bi;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if (tid < o && (red[tid + o] > red[tid] || (red[tid + o] == red[tid] && order[tid + o] < order[tid]))) {
        red[tid] = red[tid + o];
        order[tid] = order[tid + o];
      }
      __syncthreads();
    }
  } else {
    for (int i = tid; i < VMAX; i += 128) cdf[i] = i < V ? q[i] : 0.0;
    __syncthreads();
    warp0_inclusive_scan(cdf, tid);
    __syncthreads();
  }
  if (tid == 0) {
    int idx = 0;
    if (a.mode == SMER_SAMPLE_GREEDY) {
      idx = order[0];
    } else {
      uint64_t step = a.step_base + (a.gen_count ? (uint64_t)a.gen_count[s] : 0);
      for (uint32_t draw = 0; draw < 12; ++draw) {
        double u = uniform01(a.seed, (uint64_t)(a.seq_base + s), step, draw);
        int lo = 0, hi = V - 1;                  // first i with u < cdf[i]  (cdf is non-decreasing)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (u < cdf[mid]) hi = mid; else lo = mid + 1;
        }
        idx = lo;
        if (a.raw_flags || accepted(f.accept, idx)) break;     // draws 0..10 must be accepted, draw 11 is kept
      }
    }
    chosen = idx;
    if (a.out_token) a.out_token[s] = idx;
    if (a.state) {
      // generation.py:654-671
      if (idx == 241) { st |= SMER_ST_CONTINUE; st &= ~SMER_ST_SEP; }
      if (in_range(idx, 146, 233)) { st |= SMER_ST_PITCH; st &= ~(SMER_ST_SEP | SMER_ST_CONTINUE); }
      if (in_range(idx, 234, 238)) { st &= ~(SMER_ST_REST | SMER_ST_PITCH); }
      if (idx == 240) st |= SMER_ST_SEP;
      if (idx == 239) st |= SMER_ST_REST;
    }
    if (a.tok_buf) {
      // generation.py:673-686 span bookkeeping on the decoder input stream
      int64_t* buf = a.tok_buf + (long long)s * a.max_len;
      int len = a.cur_len[s];
      int gen = a.gen_count[s] + 1;
      bool end_span = false;
      bool is_ctrl = a.control_bitmap && ((a.control_bitmap[idx >> 5] >> (idx & 31)) & 1u);
      if (len >= a.max_len) {
        // the stream buffer is full: the token is dropped and the piece ends here (never write past the row)
        a.done[s] = 1;
        gen -= 1;
      } else if (is_ctrl) {
        buf[len++] = idx;            // kept; the forced <eos> is the element that gets dropped
        gen += 1;
        end_span = true;
      } else if (idx == 1) {
        end_span = true;             // sampled <eos> is dropped
      } else if (len - a.span_start[s] + 1 >= a.max_span) {
        end_span = true;             // cap reached: the last sampled token is dropped (line 686)
      } else {
        buf[len++] = idx;
      }
      if (end_span && !a.done[s]) {
        int si = a.span_idx[s] + 1;
        st = 0;
        if (si < a.n_spans[s] && len + 1 < a.max_len) {
          a.span_idx[s] = si;
          a.span_start[s] = len;
          buf[len++] = 2;            // next span opens with m_0
        } else {
          a.done[s] = 1;
        }
      }
      a.cur_len[s] = len;
      a.gen_count[s] = gen;
      if (a.fed_len && !a.done[s]) a.fed_len[s] = pos + 1;
    }
    if (a.state) a.state[s] = st;
  }
}

extern "C" int smer_sample_masked(const smer_sample_args* a, void* stream) {
  SMER_CHECK_ARG(a && a->logits && a->n_seq > 0, "smer_sample_masked: null args");
  SMER_CHECK_ARG(a->V > 0 && a->V <= VMAX, "smer_sample_masked: V=%d exceeds %d", a->V, VMAX);
  SMER_CHECK_ARG(a->temperature > 0.0, "smer_sample_masked: temperature must be positive");
  SMER_CHECK_ARG(!a->tok_buf || (a->cur_len && a->span_start && a->span_idx && a->n_spans && a->done && a->gen_count),
                 "smer_sample_masked: stream bookkeeping needs cur_len/span_start/span_idx/n_spans/done/gen_count");
  smer_launch_pdl(sample_kernel, dim3(a->n_seq), dim3(128), 0, (cudaStream_t)stream, *a);
  SMER_CHECK_LAUNCH("smer_sample_masked");
  return SMER_OK;
}
