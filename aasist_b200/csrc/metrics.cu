// Detection metrics, the step after the hot path: the reference turns the 71 237 scores of
// an evaluation run into EER and min t-DCF on the host with numpy (evaluation.py:120-154 compute_det_curve /
// compute_eer, :266-282 the t-DCF curve).  Here the whole computation is one single-CTA kernel over scores that
// are already in device memory:
//
//   keys  = order-preserving 64-bit image of the float64 scores, concatenation [targets, nontargets]
//   sort  = 16 passes of a STABLE 4-bit LSD radix sort (np.argsort(kind='mergesort') is stable: equal scores
//           keep the concatenation order, targets first) -- per-thread contiguous chunks, per-thread digit
//           counters in shared memory, one block-wide scan per pass, in-order scatter
//   curve = running count of targets -> frr = cum/n_t, far = (n_n - (i - cum))/n_n in IEEE float64 with the
//           reference's operation order (no FMA contraction), so every value is bit-identical to numpy's
//   pick  = first index of min |frr - far| (EER) and of min (C1*frr + C2*far)/min(C1,C2) (t-DCF), np.argmin order
//
// 71 237 scores are 0.57 MB of keys: the job is latency-, not bandwidth-bound, and one CTA avoids every grid-wide
// synchronisation.  Larger inputs work (the per-thread chunk just grows).
#include <math.h>

#include "common.cuh"

namespace aasist {

constexpr int kDetThreads = 1024;
constexpr int kDetDigits = 16;             // 4-bit digits

struct DetArgs {
  const double* target;
  const double* nontarget;
  long long n_t, n_n;
  double c1, c2;                           // t-DCF weights; c1 < 0: no t-DCF
  unsigned long long *k0, *k1;             // key ping-pong
  unsigned int *v0, *v1;                   // original index ping-pong
  double* results;                         // [8] device
  double *frr, *far, *thr, *tdcf;          // optional curves (n+1 each)
};

__device__ __forceinline__ unsigned long long det_key(double x) {
  if (x == 0.0) x = 0.0;                   // -0.0 and +0.0 compare equal in numpy: same key
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// exclusive block scan of one unsigned value per thread (1024 threads); returns the exclusive prefix, *total = sum
__device__ __forceinline__ unsigned int block_exscan(unsigned int v, unsigned int* warp_tot, unsigned int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned int w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_tot[lane] = w;                    // inclusive over warps
  }
  __syncthreads();
  const unsigned int base = warp == 0 ? 0u : warp_tot[warp - 1];
  if (total) *total = warp_tot[31];
  __syncthreads();
  return base + inc - v;
}

struct Best {                               // an operating point of the DET curve, ordered by (val, idx)
  double val, frr, far, thr;
  long long idx;
};
__device__ __forceinline__ void best_take(Best& b, const Best& o) {
  if (o.val < b.val || (o.val == b.val && o.idx < b.idx)) b = o;
}
__device__ __forceinline__ Best best_shfl_down(const Best& b, int o) {
  Best r;
  r.val = __shfl_down_sync(0xffffffffu, b.val, o);
  r.frr = __shfl_down_sync(0xffffffffu, b.frr, o);
  r.far = __shfl_down_sync(0xffffffffu, b.far, o);
  r.thr = __shfl_down_sync(0xffffffffu, b.thr, o);
  r.idx = __shfl_down_sync(0xffffffffu, b.idx, o);
  return r;
}

__global__ void __launch_bounds__(kDetThreads, 1)
det_metrics_kernel(const DetArgs a) {
  extern __shared__ unsigned int s_hist[];                 // [16][1024] digit counters, then scan scratch
  __shared__ unsigned int s_warp[32];
  __shared__ Best s_best[2][32];
  __shared__ unsigned int s_flags;
  const long long n = a.n_t + a.n_n;
  const int t = threadIdx.x;
  const long long chunk = (n + kDetThreads - 1) / kDetThreads;
  const long long lo = min(n, (long long)t * chunk), hi = min(n, lo + chunk);
  if (t == 0) s_flags = 0;
  __syncthreads();

  // ---- keys
  bool bad = false;
  for (long long i = lo; i < hi; ++i) {
    const double x = i < a.n_t ? a.target[i] : a.nontarget[i - a.n_t];
    if (isnan(x) || isinf(x)) bad = true;
    a.k0[i] = det_key(x);
    a.v0[i] = (unsigned int)i;
  }
  if (bad) atomicOr(&s_flags, 1u);
  __syncthreads();

  // ---- stable LSD radix sort, 4 bits per pass
  unsigned long long *ks = a.k0, *kd = a.k1;
  unsigned int *vs = a.v0, *vd = a.v1;
  for (int pass = 0; pass < 64 / 4; ++pass) {
    const int shift = 4 * pass;
    {   // a digit shared by every key leaves the order unchanged: skip the pass (fp32-born scores have 29 zero
        // mantissa bits as float64, i.e. 7 of the 16 passes)
      const unsigned int d0 = (unsigned int)((ks[0] >> shift) & 15);
      bool same = true;
      for (long long i = lo; i < hi; ++i) same = same && ((unsigned int)((ks[i] >> shift) & 15) == d0);
      if (__syncthreads_and(same)) continue;
    }
#pragma unroll
    for (int d = 0; d < kDetDigits; ++d) s_hist[d * kDetThreads + t] = 0;
    __syncthreads();                                       // previous pass's scatter is complete, too
    for (long long i = lo; i < hi; ++i) s_hist[(int)((ks[i] >> shift) & 15) * kDetThreads + t]++;
    __syncthreads();
    // exclusive scan of the flattened [digit][thread] table: thread t owns entries [16t, 16t+16)
    unsigned int loc[16], sum = 0;
#pragma unroll
    for (int e = 0; e < 16; ++e) { loc[e] = s_hist[16 * t + e]; sum += loc[e]; }
    unsigned int run = block_exscan(sum, s_warp, nullptr);
#pragma unroll
    for (int e = 0; e < 16; ++e) { s_hist[16 * t + e] = run; run += loc[e]; }
    __syncthreads();
    for (long long i = lo; i < hi; ++i) {
      const unsigned long long k = ks[i];
      const unsigned int pos = s_hist[(int)((k >> shift) & 15) * kDetThreads + t]++;
      kd[pos] = k;
      vd[pos] = vs[i];
    }
    __syncthreads();
    unsigned long long* tk = ks; ks = kd; kd = tk;
    unsigned int* tv = vs; vs = vd; vd = tv;
  }
  // the sorted data are in (ks, vs)

  // ---- DET curve, EER and min t-DCF
  unsigned int cnt = 0, uniq = 0;
  for (long long i = lo; i < hi; ++i) {
    cnt += vs[i] < a.n_t ? 1u : 0u;
    uniq += (i == 0 || ks[i] != ks[i - 1]) ? 1u : 0u;
  }
  unsigned int n_uniq = 0;
  const unsigned int before = block_exscan(cnt, s_warp, nullptr);
  (void)block_exscan(uniq, s_warp, &n_uniq);
  const double nt = (double)a.n_t, nn = (double)a.n_n;
  const bool want_tdcf = a.c1 >= 0.0;
  const double cmin = fmin(a.c1, a.c2);
  Best be = {INFINITY, 0.0, 0.0, 0.0, 0}, bt = be;
  auto point = [&](long long idx, double frr, double far, double thr) {
    best_take(be, Best{fabs(__dsub_rn(frr, far)), frr, far, thr, idx});
    double td = NAN;
    if (want_tdcf) {
      td = __ddiv_rn(__dadd_rn(__dmul_rn(a.c1, frr), __dmul_rn(a.c2, far)), cmin);
      best_take(bt, Best{td, frr, far, thr, idx});
    }
    if (a.frr) a.frr[idx] = frr;
    if (a.far) a.far[idx] = far;
    if (a.thr) a.thr[idx] = thr;
    if (a.tdcf) a.tdcf[idx] = td;
  };
  auto score_of = [&](long long i) {
    const unsigned int v = vs[i];
    return v < a.n_t ? a.target[v] : a.nontarget[v - a.n_t];
  };
  if (t == 0 && n > 0) point(0, 0.0, 1.0, __dsub_rn(score_of(0), 0.001));
  unsigned int cum = before;
  for (long long i = lo; i < hi; ++i) {
    cum += vs[i] < a.n_t ? 1u : 0u;
    const double tsum = (double)cum;
    const double frr = __ddiv_rn(tsum, nt);
    // nontarget_scores.size - (arange(1, n+1) - tar_trial_sums), all exact integers in float64
    const double far = __ddiv_rn(__dsub_rn(nn, __dsub_rn((double)(i + 1), tsum)), nn);
    point(i + 1, frr, far, score_of(i));
  }
  // block argmin (first index on ties)
  const int lane = t & 31, warp = t >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    best_take(be, best_shfl_down(be, o));
    best_take(bt, best_shfl_down(bt, o));
  }
  if (lane == 0) { s_best[0][warp] = be; s_best[1][warp] = bt; }
  __syncthreads();
  if (t == 0) {
    for (int w = 1; w < 32; ++w) {
      best_take(be, s_best[0][w]);
      best_take(bt, s_best[1][w]);
    }
    a.results[0] = __ddiv_rn(__dadd_rn(be.frr, be.far), 2.0);   // np.mean((frr[i], far[i]))
    a.results[1] = be.thr;
    a.results[2] = (double)be.idx;
    a.results[3] = want_tdcf ? bt.val : NAN;
    a.results[4] = want_tdcf ? bt.thr : NAN;
    a.results[5] = want_tdcf ? (double)bt.idx : -1.0;
    a.results[6] = (double)n_uniq;
    a.results[7] = (double)s_flags;
  }
}

}  // namespace aasist

using namespace aasist;

#pragma GCC visibility push(default)
extern "C" int64_t aasist_det_workspace_bytes(int64_t n_total) {
  if (n_total < 0) return 0;
  return 2 * (int64_t)sizeof(unsigned long long) * n_total + 2 * (int64_t)sizeof(unsigned int) * n_total + 1024;
}

extern "C" int aasist_det_metrics(const double* target_dev, int64_t n_target, const double* nontarget_dev,
                                  int64_t n_nontarget, double c1, double c2, double* results_host,
                                  double* frr_dev, double* far_dev, double* thr_dev, double* tdcf_dev,
                                  void* workspace_dev, int64_t workspace_bytes, void* stream) {
  const int64_t n = n_target + n_nontarget;
  if (!target_dev || !nontarget_dev || !results_host || n_target < 1 || n_nontarget < 1 || n > 0x7fffffffll) {
    // the reference divides by target_scores.size / nontarget_scores.size (evaluation.py:136-139)
    set_error("aasist_det_metrics: need at least one target and one nontarget score (got %lld, %lld)",
              (long long)n_target, (long long)n_nontarget);
    return AASIST_E_INVALID;
  }
  if (!workspace_dev || workspace_bytes < aasist_det_workspace_bytes(n)) {
    set_error("aasist_det_metrics: workspace needs %lld bytes", (long long)aasist_det_workspace_bytes(n));
    return AASIST_E_WORKSPACE;
  }
  if ((c1 >= 0.0) != (c2 >= 0.0)) {
    set_error("aasist_det_metrics: t-DCF weights must both be >= 0 (or both negative to skip the t-DCF)");
    return AASIST_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  DetArgs a;
  a.target = target_dev; a.nontarget = nontarget_dev; a.n_t = n_target; a.n_n = n_nontarget;
  a.c1 = c1; a.c2 = c2;
  char* w = (char*)(((uintptr_t)workspace_dev + 255) & ~(uintptr_t)255);
  a.results = (double*)w;                 w += 256;
  a.k0 = (unsigned long long*)w;          w += sizeof(unsigned long long) * n;
  a.k1 = (unsigned long long*)w;          w += sizeof(unsigned long long) * n;
  a.v0 = (unsigned int*)w;                w += sizeof(unsigned int) * n;
  a.v1 = (unsigned int*)w;
  a.frr = frr_dev; a.far = far_dev; a.thr = thr_dev; a.tdcf = tdcf_dev;
  const size_t smem = sizeof(unsigned int) * kDetDigits * kDetThreads;
  AASIST_CUDA(cudaFuncSetAttribute(det_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  det_metrics_kernel<<<1, kDetThreads, smem, st>>>(a);
  AASIST_CUDA(cudaGetLastError());
  AASIST_CUDA(cudaMemcpyAsync(results_host, a.results, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
  AASIST_CUDA(cudaStreamSynchronize(st));
  return AASIST_OK;
}
#pragma GCC visibility pop
