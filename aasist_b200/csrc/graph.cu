// Graph module: everything after the encoder, one CTA per utterance, node features in shared
// memory, fp32 throughout (accurate tanhf/expf: GraphPool indices must match the reference).
//   a5  node extraction                 models/AASIST.py:841-842,848-849
//   a6  GraphAttentionLayer             models/AASIST.py:43-110
//   a7  GraphPool (sigmoid, top-k)      models/AASIST.py:294-322
//   a8  HtrgGraphAttentionLayer         models/AASIST.py:150-282
//   a9  branch fusion, readout, head    models/AASIST.py:865-921
//   a11 RawGAT-ST graph tail            models/RawNetGatSpoofST.py:338-356
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace aasist {

using namespace ptx;

constexpr int kGraphThreads = 256;
constexpr int kWarps = kGraphThreads / 32;
constexpr int kMaxDim = 64;          // max feature width of any graph layer

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#ifdef AASIST_KERNEL_STATS
__device__ long long g_graph_stats[32];
#define LSTAMP(k) do { if (threadIdx.x == 0) { long long now_ = clock64(); atomicAdd((unsigned long long*)&g_graph_stats[k], (unsigned long long)(now_ - *S.tlast)); *S.tlast = now_; } } while (0)
#else
#define LSTAMP(k) do { } while (0)
#endif

struct Scratch {
  float* Wst;    // [kMaxDim][kMaxDim] staged attention projection (transposed, padded to Dop).  ALIASES the layer
                 // output buffer and AGG (carve_layer_buffers): neither is live while the attention map is built,
                 // and without 16 KB of its own the kernel fits four CTAs per SM instead of three
  float* vb;     // staged attB
  float* va;     // staged att weight (type 11 / plain)
  float* vb2;    // staged att weight 22
  float* vc;     // staged att weight 12
  float* A;      // [nmax][lda] attention logits / map
  int lda;
  float* AGG;    // [nmax][ld]
  float* sc;     // [nmax] pool scores (sigmoid)
  float* wts;    // [nmax] pool weights (pre-sigmoid) / master logits
  int* idx;      // [nmax]
  float* aggM;   // [kMaxDim]
  // tensor-core attention maps (att_logits_tc): on/off, TMEM base, completion mbarrier and its phase,
  // and the partial logits of the second column half of a 128-pair tile
  int use_tc;
  uint32_t tmem;
  uint64_t* mma_bar;
  uint32_t* mma_phase;
  float* part;   // [128]
  long long* tlast;   // stats build only: time of the previous stamp (thread 0)
};

// ---- attention logits over all unordered node pairs (the map is symmetric) --------------
// A[i][j] = A[j][i] = (w_type . tanh(W_att (x_i * x_j) + b_att)) / temp
// one lane per pair, 8 output dims at a time; weights broadcast from shared memory.
__device__ void att_logits(const float* X, int N, int ld, int D, int Do, const float* attWt,
                           const float* attB, const float* w11, const float* w22,
                           const float* w12, int n1, float temp, const Scratch& S) {
  const int Dop = (Do + 7) & ~7;
  for (int i = threadIdx.x; i < D * Dop; i += kGraphThreads) {
    int d = i / Dop, k = i % Dop;
    S.Wst[i] = k < Do ? attWt[d * Do + k] : 0.f;
  }
  for (int k = threadIdx.x; k < Dop; k += kGraphThreads) {
    bool ok = k < Do;
    S.vb[k] = ok ? attB[k] : 0.f;
    S.va[k] = ok ? w11[k] : 0.f;
    S.vb2[k] = ok ? w22[k] : 0.f;
    S.vc[k] = ok ? w12[k] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pairs = N * (N + 1) / 2;
  for (int base = warp * 32; base < pairs; base += kWarps * 32) {
    int p = base + lane;
    bool valid = p < pairs;
    if (!valid) p = 0;
    // decode p -> (i, j), i <= j, row-major upper triangle; start(i) = i*(2N-i+1)/2
    float fn = (float)(2 * N + 1);
    int i = (int)((fn - sqrtf(fn * fn - 8.f * (float)p)) * 0.5f);
    i = max(0, min(i, N - 1));
    while (i + 1 < N && (i + 1) * (2 * N - i) / 2 <= p) ++i;
    while (i > 0 && i * (2 * N - i + 1) / 2 > p) --i;
    int j = i + (p - i * (2 * N - i + 1) / 2);
    const float* xi = X + i * ld;
    const float* xj = X + j * ld;
    const float* wsel = (j < n1) ? S.va : ((i >= n1) ? S.vb2 : S.vc);
    float logit = 0.f;
    for (int kg = 0; kg < Dop; kg += 8) {
      float pre[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) pre[q] = S.vb[kg + q];
#pragma unroll 4
      for (int d = 0; d < D; ++d) {
        float pr = xi[d] * xj[d];
        const float4 wa = *reinterpret_cast<const float4*>(S.Wst + d * Dop + kg);
        const float4 wb = *reinterpret_cast<const float4*>(S.Wst + d * Dop + kg + 4);
        pre[0] = fmaf(pr, wa.x, pre[0]); pre[1] = fmaf(pr, wa.y, pre[1]);
        pre[2] = fmaf(pr, wa.z, pre[2]); pre[3] = fmaf(pr, wa.w, pre[3]);
        pre[4] = fmaf(pr, wb.x, pre[4]); pre[5] = fmaf(pr, wb.y, pre[5]);
        pre[6] = fmaf(pr, wb.z, pre[6]); pre[7] = fmaf(pr, wb.w, pre[7]);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) logit = fmaf(wsel[kg + q], tanhf(pre[q]), logit);
    }
    logit = logit / temp;
    if (valid) {
      S.A[i * S.lda + j] = logit;
      S.A[j * S.lda + i] = logit;
    }
  }
  __syncthreads();
}

// ---- the same attention map on the tensor cores ------------------------------------------------------------
// pre[p][k] = sum_d (x_i[d] x_j[d]) W[k][d] is a GEMM: M = node pairs (tiles of 128), K = D, N = Do.  A thread owns
// a pair = a TMEM lane: it forms the D products in fp32, splits them into fp16 (hi, lo) and writes them straight
// into tensor memory as the A operand (tcgen05.st); B = att_proj.weight as an fp16 (hi, lo) image staged in shared
// memory (where the fp32 path keeps its transposed copy); three products per K chunk (hi*hi + lo*hi + hi*lo, fp32
// accumulation -- the scheme of the encoder, ~2^-22 relative); the accumulator row comes back to the same thread
// for bias, tanh, the dot product with the attention weight and the symmetric store.  Warps w and w+4 share a lane
// quadrant: each builds half of the K chunks and reduces half of the output columns.
constexpr int kAttTmemCols = 128;      // [0,64) A operand: 16 columns per K chunk [hi 8 | lo 8];  [64,128) accumulator
__host__ __device__ inline int att_pad16(int v) { return (v + 15) & ~15; }
__device__ __forceinline__ uint64_t att_desc_noswz(uint32_t smem_addr) {   // K-major, no swizzle: LBO 128 B, SBO 256 B
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}

__device__ void att_logits_tc(const float* X, int N, int ld, int D, int Do, const float* attImg,
                              const float* attB, const float* w11, const float* w22, const float* w12, int n1,
                              float temp, const Scratch& S) {
  const int Dp = att_pad16(D), Dop = att_pad16(Do), KC = Dp >> 4;
  const int chunk_bytes = Dop * 32;
  // operand image [hi: KC chunks][lo: KC chunks] of [Dop rows x 32 B] -> shared memory (aliases the layer buffers)
  {
    const uint4* src = reinterpret_cast<const uint4*>(attImg);
    uint4* dst = reinterpret_cast<uint4*>(S.Wst);
    for (int i = threadIdx.x; i < 2 * KC * chunk_bytes / 16; i += kGraphThreads) dst[i] = __ldg(src + i);
  }
  for (int k = threadIdx.x; k < Dop; k += kGraphThreads) {
    const bool ok = k < Do;
    S.vb[k] = ok ? attB[k] : 0.f;
    S.va[k] = ok ? w11[k] : 0.f;
    S.vb2[k] = ok ? w22[k] : 0.f;
    S.vc[k] = ok ? w12[k] : 0.f;
  }
  fence_proxy_async_smem();
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int quad = warp & 3, half = warp >> 2;
  const int pairs = N * (N + 1) / 2;
  const uint32_t t_lane = S.tmem + ((uint32_t)(quad * 32) << 16);
  const uint32_t img = smem_u32(S.Wst);
  // K chunks of this half: [kc0, kc1)
  const int kc0 = half == 0 ? 0 : (KC + 1) / 2, kc1 = half == 0 ? (KC + 1) / 2 : KC;
  const int nc = Dop >> 1;                     // accumulator columns of this half: 8, 16, 24 or 32
  for (int base = 0; base < pairs; base += 128) {
    int p = base + quad * 32 + lane;
    const bool valid = p < pairs;
    if (!valid) p = 0;
    // decode p -> (i, j), i <= j, row-major upper triangle; start(i) = i*(2N-i+1)/2
    const float fn = (float)(2 * N + 1);
    int i = (int)((fn - sqrtf(fn * fn - 8.f * (float)p)) * 0.5f);
    i = max(0, min(i, N - 1));
    while (i + 1 < N && (i + 1) * (2 * N - i) / 2 <= p) ++i;
    while (i > 0 && i * (2 * N - i + 1) / 2 > p) --i;
    const int j = i + (p - i * (2 * N - i + 1) / 2);
    const float* xi = X + i * ld;
    const float* xj = X + j * ld;
    for (int kc = kc0; kc < kc1; ++kc) {
      uint32_t w[16];                          // [hi: K 0..15 | lo: K 0..15] of this pair's chunk
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int d = 16 * kc + 2 * q;
        const float a0 = d < D ? xi[d] * xj[d] : 0.f;
        const float a1 = d + 1 < D ? xi[d + 1] * xj[d + 1] : 0.f;
        split2_sat(a0, a1, w[q], w[8 + q]);
      }
      tmem_st16(t_lane + (uint32_t)(16 * kc), w);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after_sync();
      const uint32_t idesc = umma_idesc_f16(128, Dop);
      for (int kc = 0; kc < KC; ++kc) {
        const uint64_t b_hi = att_desc_noswz(img + (uint32_t)(kc * chunk_bytes));
        const uint64_t b_lo = att_desc_noswz(img + (uint32_t)((KC + kc) * chunk_bytes));
        const uint32_t a_hi = S.tmem + (uint32_t)(16 * kc), a_lo = a_hi + 8;
        umma_f16_ts(S.tmem + 64u, a_hi, b_hi, idesc, kc > 0 ? 1u : 0u);
        umma_f16_ts(S.tmem + 64u, a_lo, b_hi, idesc, 1);
        umma_f16_ts(S.tmem + 64u, a_hi, b_lo, idesc, 1);
      }
      umma_commit(S.mma_bar);
    }
    mbar_wait(S.mma_bar, *S.mma_phase);
    tc_fence_after_sync();
    // this half's accumulator columns [half*nc, half*nc + nc): bias, tanh, dot with the attention weight
    const float* wsel = (j < n1) ? S.va : ((i >= n1) ? S.vb2 : S.vc);
    float logit = 0.f;
    for (int c = 0; c < nc; c += 8) {
      uint32_t acc[8];
      tmem_ld8_async(t_lane + 64u + (uint32_t)(half * nc + c), acc);
      tmem_ld_wait8(acc);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = half * nc + c + q;
        logit = fmaf(wsel[k], tanhf(__uint_as_float(acc[q]) + S.vb[k]), logit);
      }
    }
    if (half == 1) S.part[quad * 32 + lane] = logit;
    tc_fence_before_sync();
    __syncthreads();                           // partial sums visible; accumulator and A columns free for the next tile
    if (threadIdx.x == 0) *S.mma_phase ^= 1u;
    if (half == 0 && valid) {
      logit = (logit + S.part[quad * 32 + lane]) / temp;
      S.A[i * S.lda + j] = logit;
      S.A[j * S.lda + i] = logit;
    }
    __syncthreads();                           // phase flip and S.part reads ordered before the next tile
  }
}

// softmax over j of every row i (F.softmax(att_map, dim=-2), AASIST.py:89)
__device__ void softmax_rows(int N, const Scratch& S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < N; i += kWarps) {
    float* row = S.A + i * S.lda;
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < N; j += 32) {
      float e = expf(row[j] - m);
      row[j] = e;
      s += e;
    }
    s = warp_sum(s);
    for (int j = lane; j < N; j += 32) row[j] = row[j] / s;
  }
  __syncthreads();
}

// AGG[i][:] = sum_j A[i][j] X[j][:]     (torch.matmul(att_map.squeeze(-1), x), AASIST.py:94)
__device__ void aggregate(const float* X, int N, int ld, int D, const Scratch& S) {
  for (int t = threadIdx.x; t < N * D; t += kGraphThreads) {
    int i = t / D, d = t % D;
    const float* a = S.A + i * S.lda;
    float s = 0.f;
    for (int j = 0; j < N; ++j) s = fmaf(a[j], X[j * ld + d], s);
    S.AGG[i * ld + d] = s;
  }
  __syncthreads();
}

// OUT[i][k] = act(bias[k] + sum_d Wt[d][k] IN[i][d] (+ sum_d W2t[d][k] IN2[i][d]))
// rows [r0, r0+n) of IN/IN2/OUT; weights in global memory (coalesced over k, L1/L2 resident).
template <bool SELU>
__device__ void linear_rows(const float* IN, const float* IN2, int n, int ldin, int D,
                            const float* Wt, const float* W2t, const float* bias, int Do,
                            float* OUT, int ldout) {
  const int nblk = (n + 3) / 4;
  for (int t = threadIdx.x; t < nblk * Do; t += kGraphThreads) {
    int k = t % Do, i0 = (t / Do) * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) r[q] = min(i0 + q, n - 1) * ldin;
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      float w = __ldg(Wt + d * Do + k);
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(w, IN[r[q] + d], acc[q]);
      if (IN2) {
        float w2 = __ldg(W2t + d * Do + k);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = fmaf(w2, IN2[r[q] + d], acc[q]);
      }
    }
    float bk = __ldg(bias + k);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (i0 + q < n) {
        float v = acc[q] + bk;
        OUT[(i0 + q) * ldout + k] = SELU ? selu(v) : v;
      }
  }
  __syncthreads();
}

// GraphAttentionLayer.forward (AASIST.py:43-59): X (N,D) -> OUT (N,Do)
__device__ void gat_layer(const float* X, int N, int ld, const GatParams& P, float* OUT,
                          const Scratch& S) {
  LSTAMP(12);
  if (S.use_tc) att_logits_tc(X, N, ld, P.D, P.Do, P.attImg, P.attB, P.attW, P.attW, P.attW, N, P.temp, S);
  else att_logits(X, N, ld, P.D, P.Do, P.attWt, P.attB, P.attW, P.attW, P.attW, N, P.temp, S);
  LSTAMP(13);
  softmax_rows(N, S);
  LSTAMP(14);
  aggregate(X, N, ld, P.D, S);
  LSTAMP(15);
  linear_rows<true>(S.AGG, X, N, ld, P.D, P.pWt, P.qWt, P.bias, P.Do, OUT, ld);
  LSTAMP(16);
}

// HtrgGraphAttentionLayer.forward (AASIST.py:150-185).
// X1 (n1,D) temporal, X2 (n2,D) spectral, m_in (D)  ->  OUT rows [0,n1) / [n1,n1+n2), m_out (Do)
// HX: scratch (n1+n2, D) for the type-projected nodes.
__device__ void htrg_layer(const float* X1, int n1, const float* X2, int n2, int ld,
                           const float* m_in, const HtrgParams& P, float* HX, float* OUT,
                           float* m_out, const Scratch& S) {
  const int N = n1 + n2, D = P.D, Do = P.Do;
  LSTAMP(12);
  linear_rows<false>(X1, nullptr, n1, ld, D, P.t1Wt, nullptr, P.t1B, D, HX, ld);              // :158
  linear_rows<false>(X2, nullptr, n2, ld, D, P.t2Wt, nullptr, P.t2B, D, HX + n1 * ld, ld);    // :159
  LSTAMP(17);
  if (S.use_tc) att_logits_tc(HX, N, ld, D, Do, P.attImg, P.attB, P.w11, P.w22, P.w12, n1, P.temp, S);
  else att_logits(HX, N, ld, D, Do, P.attWt, P.attB, P.w11, P.w22, P.w12, n1, P.temp, S);     // :225-251
  LSTAMP(18);
  softmax_rows(N, S);                                                                        // :253
  LSTAMP(19);
  // master attention (AASIST.py:208-223): lm[j] = wM . tanh(W_M (x_j * m) + b_M) / temp
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = warp; j < N; j += kWarps) {
      float part = 0.f;
      for (int k = lane; k < Do; k += 32) {
        float pre = __ldg(P.attMB + k);
        for (int d = 0; d < D; ++d)
          pre = fmaf(HX[j * ld + d] * m_in[d], __ldg(P.attMWt + d * Do + k), pre);
        part = fmaf(__ldg(P.wM + k), tanhf(pre), part);
      }
      part = warp_sum(part);
      if (lane == 0) S.wts[j] = part / P.temp;
    }
    __syncthreads();
    if (warp == 0) {  // softmax over nodes
      float m = -INFINITY;
      for (int j = lane; j < N; j += 32) m = fmaxf(m, S.wts[j]);
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < N; j += 32) {
        float e = expf(S.wts[j] - m);
        S.wts[j] = e;
        s += e;
      }
      s = warp_sum(s);
      for (int j = lane; j < N; j += 32) S.wts[j] = S.wts[j] / s;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += kGraphThreads) {
      float s = 0.f;
      for (int j = 0; j < N; ++j) s = fmaf(S.wts[j], HX[j * ld + d], s);
      S.aggM[d] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Do; k += kGraphThreads) {                                   // :263-269
      float a = 0.f, b = 0.f;
      for (int d = 0; d < D; ++d) {
        a = fmaf(__ldg(P.pMWt + d * Do + k), S.aggM[d], a);
        b = fmaf(__ldg(P.qMWt + d * Do + k), m_in[d], b);
      }
      m_out[k] = a + b + __ldg(P.biasM + k);
    }
    __syncthreads();
  }
  LSTAMP(20);
  aggregate(HX, N, ld, D, S);                                                                 // :258
  LSTAMP(21);
  linear_rows<true>(S.AGG, HX, N, ld, D, P.pWt, P.qWt, P.bias, Do, OUT, ld);                 // :257-261,179-180
  LSTAMP(22);
}

// GraphPool.forward (AASIST.py:294-322): H (N,D) -> OUT (k,D) in descending score order.
// Nodes are ranked on the PRE-sigmoid weight w: sigmoid is monotone, so wherever the reference's fp32
// sigmoid values differ strictly the order is the same as torch.topk's on the scores, and it does not
// depend on the last ulp of any sigmoid implementation (a sigmoid collision of two different w is an
// exact tie in the reference, whose order torch leaves unspecified).  Equal w go to the lower node index.
__device__ void graph_pool(const float* H, int N, int ld, const PoolParams& P, int k, float* OUT,
                           int32_t* g_idx, float* g_wts, const Scratch& S) {
  for (int i = threadIdx.x; i < N; i += kGraphThreads) {
    float w = 0.f;
    for (int d = 0; d < P.D; ++d) w = fmaf(__ldg(P.w + d), H[i * ld + d], w);
    w += P.b;
    S.wts[i] = w;
    S.sc[i] = 1.f / (1.f + expf(-w));
    if (g_wts) g_wts[i] = w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += kGraphThreads) {
    float wi = S.wts[i];
    int r = 0;
    for (int j = 0; j < N; ++j) {
      float wj = S.wts[j];
      r += (wj > wi) || (wj == wi && j < i);
    }
    if (r < k) {
      S.idx[r] = i;
      if (g_idx) g_idx[r] = i;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < k * P.D; t += kGraphThreads) {
    int r = t / P.D, d = t % P.D;
    int i = S.idx[r];
    OUT[r * ld + d] = H[i * ld + d] * S.sc[i];
  }
  __syncthreads();
}

// spectral nodes: XS[f][c] = max_t |e[c][f][t]| (+ pos[f][c]);  temporal: XT[t][c] = max_f |e[c][f][t]|
// ONE pass over e: a warp owns a channel, lane = time step, the 23 loads of a lane are independent (one L2 latency
// per channel, not one per row), the column maximum stays in the lane and the 23 row maxima are warp reductions.
// Either output may be null (RawGAT-ST takes its two node sets from two different encoders).
__device__ void nodes_from_encoder(const float* e, int C, int NT, const float* pos, float* XS, float* XT, int ld) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < C; c += kWarps) {
    const float* ec = e + (size_t)c * kSpecNodes * NT;
    float rm[kSpecNodes];
#pragma unroll
    for (int f = 0; f < kSpecNodes; ++f) rm[f] = 0.f;
    for (int t0 = 0; t0 < NT; t0 += 32) {
      const int t = t0 + lane;
      const bool ok = t < NT;
      float v[kSpecNodes];
#pragma unroll
      for (int f = 0; f < kSpecNodes; ++f) v[f] = ok ? fabsf(__ldg(ec + f * NT + t)) : 0.f;
      float cm = 0.f;
#pragma unroll
      for (int f = 0; f < kSpecNodes; ++f) {
        cm = fmaxf(cm, v[f]);
        rm[f] = fmaxf(rm[f], v[f]);
      }
      if (XT && ok) XT[t * ld + c] = cm;
    }
    if (XS) {
      float mine = 0.f;
#pragma unroll
      for (int f = 0; f < kSpecNodes; ++f) {
        const float m = warp_max(rm[f]);
        if (lane == f) mine = m;
      }
      if (lane < kSpecNodes) XS[lane * ld + c] = mine + (pos ? __ldg(pos + lane * C + c) : 0.f);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float* bump(float*& p, int n) {
  float* r = p;
  p += (n + 3) & ~3;
  return r;
}

__host__ __device__ inline int scratch_floats(int nmax) {
  int lda = nmax + 1;
  return 4 * kMaxDim + ((nmax * lda + 3) & ~3) + 3 * ((nmax + 3) & ~3) + kMaxDim + 128 + 8;
}

__device__ void carve_scratch(float*& p, int nmax, int ld, Scratch& S) {
  S.Wst = nullptr;
  S.vb = bump(p, kMaxDim);
  S.va = bump(p, kMaxDim);
  S.vb2 = bump(p, kMaxDim);
  S.vc = bump(p, kMaxDim);
  S.lda = nmax + 1;
  S.A = bump(p, nmax * S.lda);
  S.sc = bump(p, nmax);
  S.wts = bump(p, nmax);
  S.idx = reinterpret_cast<int*>(bump(p, nmax));
  S.aggM = bump(p, kMaxDim);
  S.part = bump(p, 128);
  float* ctl = bump(p, 8);                   // mbarrier (8 bytes, 8-byte aligned: every bump is a multiple of 16 B), phase, TMEM base
  S.mma_bar = reinterpret_cast<uint64_t*>(ctl);
  S.mma_phase = reinterpret_cast<uint32_t*>(ctl + 2);
  S.tmem = 0;
  S.use_tc = 0;
  S.AGG = nullptr;
}

// layer output buffer B1 and AGG, contiguous, padded so that the pair also holds the staged attention projection
__host__ __device__ inline int layer_buffer_floats(int nmax, int ld) {
  return max(2 * nmax * ld, kMaxDim * kMaxDim);
}
__device__ float* carve_layer_buffers(float*& p, int nmax, int ld, Scratch& S) {
  float* region = bump(p, layer_buffer_floats(nmax, ld));
  S.Wst = region;
  S.AGG = region + nmax * ld;
  return region;                          // B1
}

// tensor memory + completion barrier for att_logits_tc (all threads; one allocation per CTA for its lifetime)
__device__ void att_tc_begin(Scratch& S) {
  uint32_t* slot = S.mma_phase + 1;
  if (threadIdx.x == 0) {
    mbar_init(S.mma_bar, 1);
    *S.mma_phase = 0;
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc<kAttTmemCols>(slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  S.tmem = *slot;
  S.use_tc = 1;
}
__device__ void att_tc_end(const Scratch& S) {
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after_sync();
    tmem_dealloc<kAttTmemCols>(S.tmem);
  }
}

// ---------------------------------------------------------------------------------------
// AASIST graph tail (AASIST.py:841-921)
// ---------------------------------------------------------------------------------------
// rows of the region [B3 | OT | R]; nmax >= NT, so max(..., nmax) always has room for the parked temporal nodes
__host__ __device__ inline int aasist_tail_rows(int nmax, int nS, int nT, int nS2, int nT2) {
  return max((nT2 + nS2) + nT + (nT2 + nS2), nmax);
}
__host__ __device__ inline int aasist_graph_smem_floats(int nmax, int ld, int nS, int nT, int nS2,
                                                        int nT2) {
  // B0, OS, then one region [B3 | OT | R] that also holds the temporal nodes (NT rows) until gat_T has consumed them
  int rows = nmax + nS + aasist_tail_rows(nmax, nS, nT, nS2, nT2);
  return scratch_floats(nmax) + layer_buffer_floats(nmax, ld) + rows * ld + 7 * kMaxDim + 64;
}

// SpeakerConditioningModule.forward, frame level (models/AASIST.py:384-403), in place on `n` rows of R
// (all threads of the CTA).  sp = proj(emb) and u = fusion.weight[:, g1:] @ sp are precomputed per utterance.
//   attention: a = softmax_rows( att2 . tanh(att0 [f_r ; sp]) ),  out_r = relu(fus [f_r ; a_r * sp])
//   plain    :                                                     out_r = relu(fus [f_r ; sp])
__device__ void speaker_condition_rows(float* R, int n, int ld, int g1, const SpkParams& P, const float* sp,
                                       const float* u, float* tmp, const Scratch& S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (P.use_attention) {
    for (int r = warp; r < n; r += kWarps) {
      float part = 0.f;
      for (int k = lane; k < g1; k += 32) {
        float pre = __ldg(P.att0B + k);
        for (int d = 0; d < g1; ++d) pre = fmaf(R[r * ld + d], __ldg(P.att0Wt + d * g1 + k), pre);
        for (int d = 0; d < g1; ++d) pre = fmaf(sp[d], __ldg(P.att0Wt + (g1 + d) * g1 + k), pre);
        part = fmaf(__ldg(P.att2W + k), tanhf(pre), part);
      }
      part = warp_sum(part);
      if (lane == 0) S.wts[r] = part + P.att2B;
    }
    __syncthreads();
    if (warp == 0) {                                 // nn.Softmax(dim=1): over the rows (frames) of this set
      float m = -INFINITY;
      for (int j = lane; j < n; j += 32) m = fmaxf(m, S.wts[j]);
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < n; j += 32) {
        float e = expf(S.wts[j] - m);
        S.wts[j] = e;
        s += e;
      }
      s = warp_sum(s);
      for (int j = lane; j < n; j += 32) S.wts[j] = S.wts[j] / s;
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < n * g1; t += kGraphThreads) {
    const int r = t / g1, k = t % g1;
    float acc = __ldg(P.fusB + k);
    for (int d = 0; d < g1; ++d) acc = fmaf(R[r * ld + d], __ldg(P.fusWt + d * g1 + k), acc);
    acc += (P.use_attention ? S.wts[r] : 1.f) * u[k];
    tmp[r * ld + k] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n * g1; t += kGraphThreads) {
    const int r = t / g1, k = t % g1;
    R[r * ld + k] = tmp[r * ld + k];
  }
  __syncthreads();
}

#ifdef AASIST_KERNEL_STATS
#define GSTAMP(k) do { if (threadIdx.x == 0) { long long now_ = clock64(); atomicAdd((unsigned long long*)&g_graph_stats[k], (unsigned long long)(now_ - last_)); last_ = now_; } } while (0)
#else
#define GSTAMP(k) do { } while (0)
#endif

__global__ void __launch_bounds__(kGraphThreads, 4)
aasist_graph_kernel(const GraphArgsAasist a) {
  extern __shared__ __align__(16) float smem[];
#ifdef AASIST_KERNEL_STATS
  long long last_ = clock64();
#endif
  float* p = smem;
  Scratch S;
  carve_scratch(p, a.nmax, a.ld, S);
#ifdef AASIST_KERNEL_STATS
  S.tlast = &last_;
#endif
  if (a.tc) att_tc_begin(S);
  const int ld = a.ld;
  float* B0 = bump(p, a.nmax * ld);
  float* B1 = carve_layer_buffers(p, a.nmax, ld, S);
  float* OS = bump(p, a.nS * ld);
  float* tail = bump(p, aasist_tail_rows(a.nmax, a.nS, a.nT, a.nS2, a.nT2) * ld);
  float* B3 = tail;                              // pooled hetero nodes: T rows then S rows
  float* OT = B3 + (a.nT2 + a.nS2) * ld;
  float* R = OT + a.nT * ld;                     // branch-1 result: T rows then S rows
  float* XT = tail;                              // temporal nodes, parked here until gat_T has read them
  float* Rm = bump(p, kMaxDim);
  float* m1 = bump(p, kMaxDim);
  float* m2 = bump(p, kMaxDim);
  float* m0 = bump(p, kMaxDim);
  float* spv = bump(p, kMaxDim);                 // speaker projection
  float* spu = bump(p, kMaxDim);                 // fusion.weight[:, g1:] @ speaker projection
  float* emean = bump(p, kMaxDim);               // robust: mean of e over (freq, time) per channel

  // the grid is sized so that every CTA walks the same number of utterances (no partial last wave)
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
  const float* e = a.e + (size_t)b * a.C * kSpecNodes * a.NT;
  int32_t* gi = a.topk_idx ? a.topk_idx + (size_t)b * a.topk_total : nullptr;
  float* gw = a.pool_scores ? a.pool_scores + (size_t)b * a.score_total : nullptr;

  if (a.robust) {                                                         // AASIST_Robust.py:227 e.mean(dim=(2,3))
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cnt = kSpecNodes * a.NT;
    for (int c = warp; c < a.C; c += kWarps) {
      float s = 0.f;
      for (int i = lane; i < cnt; i += 32) s += e[(size_t)c * cnt + i];
      s = warp_sum(s);
      if (lane == 0) emean[c] = s / (float)cnt;
    }
    __syncthreads();
  }
  // spectral graph                                                       (AASIST.py:841-845)
  GSTAMP(0);
  nodes_from_encoder(e, a.C, a.NT, a.posS, B0, XT, ld);                  // both node sets in one pass over e
  GSTAMP(1);
  gat_layer(B0, kSpecNodes, ld, a.gatS, B1, S);
  GSTAMP(2);
  graph_pool(B1, kSpecNodes, ld, a.poolS, a.nS, OS, gi, gw, S);
  GSTAMP(3);
  if (gi) gi += a.nS;
  if (gw) gw += kSpecNodes;
  // temporal graph                                                       (AASIST.py:848-852)
  GSTAMP(4);
  gat_layer(XT, a.NT, ld, a.gatT, B1, S);
  GSTAMP(5);
  graph_pool(B1, a.NT, ld, a.poolT, a.nT, OT, gi, gw, S);
  GSTAMP(6);
  if (gi) gi += a.nT;
  if (gw) gw += a.NT;

  float* PT = B3;
  float* PS = B3 + a.nT2 * ld;
  for (int br = 0; br < (a.robust ? 1 : 2); ++br) {                      // :859-869 / :872-881
    const HtrgParams& L1 = br == 0 ? a.st11 : a.st21;
    const HtrgParams& L2 = br == 0 ? a.st12 : a.st22;
    const PoolParams& pS = br == 0 ? a.poolhS1 : a.poolhS2;
    const PoolParams& pT = br == 0 ? a.poolhT1 : a.poolhT2;
    for (int d = threadIdx.x; d < a.g0; d += kGraphThreads)
      m0[d] = __ldg((br == 0 ? a.master1 : a.master2) + d);
    __syncthreads();
    GSTAMP(7);
    htrg_layer(OT, a.nT, OS, a.nS, ld, m0, L1, B0, B1, m1, S);
    GSTAMP(8);
    graph_pool(B1 + a.nT * ld, a.nS, ld, pS, a.nS2, PS, gi, gw, S);      // pool_hS first (:862)
    if (gi) gi += a.nS2;
    if (gw) gw += a.nS;
    graph_pool(B1, a.nT, ld, pT, a.nT2, PT, gi, gw, S);
    if (gi) gi += a.nT2;
    if (gw) gw += a.nT;
    GSTAMP(9);
    htrg_layer(PT, a.nT2, PS, a.nS2, ld, m1, L2, B0, B1, m2, S);
    GSTAMP(10);
    // residual adds (:867-869) and branch-wise max (:890-892)
    const int n2 = a.nT2 + a.nS2;
    for (int t = threadIdx.x; t < n2 * a.g1; t += kGraphThreads) {
      int r = t / a.g1, k = t % a.g1;
      float v = B3[r * ld + k] + B1[r * ld + k];
      R[r * ld + k] = br == 0 ? v : fmaxf(R[r * ld + k], v);
    }
    for (int k = threadIdx.x; k < a.g1; k += kGraphThreads) {
      float v = m1[k] + m2[k];
      Rm[k] = br == 0 ? v : fmaxf(Rm[k], v);
    }
    __syncthreads();
  }
  const int g1 = a.g1;
  if (a.spk_emb) {                                                       // :895-900, frame-level conditioning
    const float* emb = a.spk_emb + (size_t)b * a.spk.emb_dim;
    for (int k = threadIdx.x; k < g1; k += kGraphThreads) {
      float s = __ldg(a.spk.projB + k);
      for (int q = 0; q < a.spk.emb_dim; ++q) s = fmaf(__ldg(a.spk.projW + (size_t)k * a.spk.emb_dim + q), emb[q], s);
      spv[k] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < g1; k += kGraphThreads) {
      float s = 0.f;
      for (int d = 0; d < g1; ++d) s = fmaf(spv[d], __ldg(a.spk.fusWt + (g1 + d) * g1 + k), s);
      spu[k] = s;
    }
    __syncthreads();
    speaker_condition_rows(R, a.nT2, ld, g1, a.spk, spv, spu, B1, S);                 // out_T
    speaker_condition_rows(R + a.nT2 * ld, a.nS2, ld, g1, a.spk, spv, spu, B1, S);    // out_S
  }
  // readout (:903-910) + output layer (:919)
  float* lh = S.AGG;  // reuse: 5*g1 floats
  for (int k = threadIdx.x; k < g1; k += kGraphThreads) {
    float tmax = 0.f, tsum = 0.f, smax = 0.f, ssum = 0.f;
    for (int r = 0; r < a.nT2; ++r) {
      float v = R[r * ld + k];
      tmax = fmaxf(tmax, fabsf(v));
      tsum += v;
    }
    for (int r = 0; r < a.nS2; ++r) {
      float v = R[(a.nT2 + r) * ld + k];
      smax = fmaxf(smax, fabsf(v));
      ssum += v;
    }
    lh[k] = tmax;
    lh[g1 + k] = tsum / (float)a.nT2;
    lh[2 * g1 + k] = smax;
    lh[3 * g1 + k] = ssum / (float)a.nS2;
    lh[4 * g1 + k] = Rm[k];
  }
  __syncthreads();
  const int nh = (a.robust ? 4 : 5) * g1;       // AASIST_Robust.py:283 drops the master node from the readout
  if (!a.robust)
    for (int k = threadIdx.x; k < nh; k += kGraphThreads) a.last_hidden[(size_t)b * nh + k] = lh[k];
  if (threadIdx.x < 64) {
    const int lane = threadIdx.x & 31, o = threadIdx.x >> 5;
    float s = 0.f;
    for (int k = lane; k < nh; k += 32) s = fmaf(lh[k], __ldg(a.outWt + k * 2 + o), s);
    s = warp_sum(s);
    const float logit = s + (o == 0 ? a.outB0 : a.outB1);
    if (a.robust) {                             // aux head on mean(e) and the ensemble (:290-301)
      float x = 0.f;
      for (int c = lane; c < a.C; c += 32) x = fmaf(emean[c], __ldg(a.auxWt + c * 2 + o), x);
      x = warp_sum(x) + (o == 0 ? a.auxB0 : a.auxB1);
      if (lane == 0) a.last_hidden[(size_t)b * 2 + o] = a.ens0 * logit + a.ens1 * x;
    }
    if (lane == 0) a.logits[(size_t)b * 2 + o] = logit;
  }
  __syncthreads();   // shared buffers are reused by the next utterance
  GSTAMP(11);
  }
  if (a.tc) att_tc_end(S);
}

// att_proj.weight (Do, D) -> the B operand of att_logits_tc: [hi: KC chunks][lo: KC chunks] of [Dop rows x 32 B] in
// the no-swizzle K-major canonical layout (8-row x 16-byte core matrices), D and Do padded to multiples of 16
std::vector<float> att_image_floats(const std::vector<float>& w, int D, int Do) {
  const int Dp = att_pad16(D), Dop = att_pad16(Do), KC = Dp / 16;
  const size_t chunk = (size_t)Dop * 32;
  std::vector<uint8_t> img(2 * KC * chunk, 0);
  for (int n = 0; n < Do; ++n)
    for (int k = 0; k < D; ++k) {
      const float v = w[(size_t)n * D + k];
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      const int kc = k / 16, kb = (k % 16) * 2;
      const size_t off = (size_t)kc * chunk + (size_t)(n / 8) * 256 + (size_t)(kb / 16) * 128 + (size_t)(n % 8) * 16 +
                         (size_t)(kb % 16);
      memcpy(&img[off], &hi, 2);
      memcpy(&img[(size_t)KC * chunk + off], &lo, 2);
    }
  std::vector<float> out(img.size() / 4);
  memcpy(out.data(), img.data(), img.size());
  return out;
}

// grid for a per-utterance kernel: as many CTAs as utterances up to the resident capacity, then the smallest
// grid that gives every CTA the same number of utterances (512 utterances on 444 slots: 256 CTAs x 2, not
// 444 + a 68-CTA tail wave)
template <typename Kernel>
static int balanced_grid(Kernel kern, int B, size_t smem, int device) {
  int per_sm = 1, sms = 148;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kGraphThreads, smem);
  {
    // a kernel that contains tcgen05.alloc is reported as ONE block per SM by the occupancy API, whatever it
    // allocates; the hardware co-schedules blocks as long as their allocations fit the 512 columns (a block whose
    // allocation does not fit waits in tcgen05.alloc).  Resident blocks: shared memory (+1 KB reserved per block),
    // registers (64 per thread), tensor memory (kAttTmemCols per block)
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess) {
      const int by_smem = (int)((size_t)227 * 1024 / (smem + 1024));
      const int by_regs = 65536 / (std::max(fa.numRegs, 1) * kGraphThreads);
      const int by_tmem = 512 / kAttTmemCols;
      per_sm = std::max(per_sm, std::max(1, std::min(std::min(by_smem, by_regs), std::min(by_tmem, 2048 / kGraphThreads))));
    }
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int slots = std::max(1, per_sm * sms);
  const int waves = (B + slots - 1) / slots;
  return (B + waves - 1) / waves;
}

int launch_graph_aasist(aasist_handle* h, const float* e, int B, int NT, const float* spk_emb, float* last_hidden,
                        float* logits, int32_t* topk, float* scores, cudaStream_t st) {
  GraphArgsAasist a = h->ga;
  const aasist_config& c = h->cfg;
  a.NT = NT;
  a.tc = h->cfg.precision != AASIST_PREC_FP32;   // tensor-core attention maps with the split-fp16 precisions
  a.spk_emb = spk_emb;
  a.nS = pooled_count(kSpecNodes, c.pool_ratios[0], 1);
  a.nT = pooled_count(NT, c.pool_ratios[1], 1);
  a.nS2 = pooled_count(a.nS, c.pool_ratios[2], 1);
  // AASIST: both hetero pools use pool_ratios[2] (AASIST.py:796-802); Robust: pool_hT uses [3] (AASIST_Robust.py:179-183)
  a.nT2 = pooled_count(a.nT, c.pool_ratios[a.robust ? 3 : 2], 1);
  a.nmax = max(max(NT, kSpecNodes), a.nT + a.nS);
  int dmax = max(max(a.C, a.g0), a.g1);
  a.ld = dmax | 1;
  a.e = e;
  a.last_hidden = last_hidden;
  a.logits = logits;
  a.topk_idx = topk;
  a.pool_scores = scores;
  const int nbr = a.robust ? 1 : 2;
  a.topk_total = a.nS + a.nT + nbr * (a.nS2 + a.nT2);
  a.score_total = kSpecNodes + NT + nbr * (a.nS + a.nT);
  size_t smem = sizeof(float) * (size_t)aasist_graph_smem_floats(a.nmax, a.ld, a.nS, a.nT, a.nS2, a.nT2);
  if (smem > 227 * 1024) {
    set_error("utterance too long for the on-chip graph stage: %d temporal nodes need %zu bytes "
              "of shared memory (max 232448)", NT, smem);
    return AASIST_E_INVALID;
  }
  AASIST_CUDA(cudaFuncSetAttribute(aasist_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  {
    LaunchSpan span(h, "aasist_graph", st);
    a.B = B;
    aasist_graph_kernel<<<balanced_grid(aasist_graph_kernel, B, smem, h->device), kGraphThreads, smem, st>>>(a);
  }
  AASIST_CUDA(cudaGetLastError());
#ifdef AASIST_KERNEL_STATS
  if (getenv("AASIST_GRAPH_STATS")) {
    long long hst[32];
    AASIST_CUDA(cudaStreamSynchronize(st));
    AASIST_CUDA(cudaMemcpyFromSymbol(hst, g_graph_stats, sizeof(hst)));
    static const char* names[12] = {"setup", "nodes_S", "gat_S", "pool_S", "nodes_T", "gat_T", "pool_T", "branch setup",
                                    "htrg1", "pools_h", "htrg2", "fuse+readout"};
    fprintf(stderr, "[graph stats] cycles per utterance:");
    for (int k = 0; k < 12; ++k) fprintf(stderr, " %s %.0f |", names[k], (double)hst[k] / B);
    static const char* sub[11] = {"(pre)", "gat.att", "gat.softmax", "gat.agg", "gat.linear", "htrg.typeproj", "htrg.att",
                                  "htrg.softmax", "htrg.master", "htrg.agg", "htrg.linear"};
    fprintf(stderr, "\n[graph stats] inside the layers:");
    for (int k = 0; k < 11; ++k) fprintf(stderr, " %s %.0f |", sub[k], (double)hst[12 + k] / B);
    fprintf(stderr, "\n");
    long long z[32] = {0};
    AASIST_CUDA(cudaMemcpyToSymbol(g_graph_stats, z, sizeof(z)));
  }
#endif
  return 0;
}

// ---------------------------------------------------------------------------------------
// RawGAT-ST graph tail (RawNetGatSpoofST.py:338-356)
// ---------------------------------------------------------------------------------------
__host__ __device__ inline int rawgat_graph_smem_floats(int nmax, int ld) {
  return scratch_floats(nmax) + layer_buffer_floats(nmax, ld) + (nmax + 2 * 12 + 12 + 12) * ld + 64;
}

__global__ void __launch_bounds__(kGraphThreads, 4)
rawgat_graph_kernel(const GraphArgsRawGat a) {
  extern __shared__ __align__(16) float smem[];
  float* p = smem;
  Scratch S;
  carve_scratch(p, a.nmax, a.ld, S);
  if (a.tc) att_tc_begin(S);
  const int ld = a.ld;
  float* B0 = bump(p, a.nmax * ld);
  float* B1 = carve_layer_buffers(p, a.nmax, ld, S);
  float* PT = bump(p, 12 * ld);   // proj_T output as 12 nodes x 32 features
  float* PS = bump(p, 12 * ld);
  float* G = bump(p, 12 * ld);
  float* Q = bump(p, 12 * ld);
  const int b = blockIdx.x;
  const float* eT = a.eT + (size_t)b * 64 * kSpecNodes * a.NT;
  const float* eS = a.eS + (size_t)b * 64 * kSpecNodes * a.NT;
  int32_t* gi = a.topk_idx ? a.topk_idx + (size_t)b * a.topk_total : nullptr;
  float* gw = a.pool_scores ? a.pool_scores + (size_t)b * a.score_total : nullptr;
  const int D1 = a.gatT.Do;  // 32

  // "T" branch: max over time -> 23 nodes (:338-341)
  nodes_from_encoder(eT, 64, a.NT, nullptr, B0, nullptr, ld);
  gat_layer(B0, kSpecNodes, ld, a.gatT, B1, S);
  graph_pool(B1, kSpecNodes, ld, a.poolT, a.nT, B0, gi, gw, S);
  if (gi) gi += a.nT;
  if (gw) gw += kSpecNodes;
  // proj_T: Linear(nT,12) over the node axis of pool_T^T -> out_T (D1,12); kept as PT[m][d]
  for (int t = threadIdx.x; t < 12 * D1; t += kGraphThreads) {
    int m = t / D1, d = t % D1;
    float s = 0.f;
    for (int n = 0; n < a.nT; ++n) s = fmaf(__ldg(a.projTW + m * a.nT + n), B0[n * ld + d], s);
    PT[m * ld + d] = s + __ldg(a.projTB + m);
  }
  __syncthreads();
  // "S" branch: max over freq -> NT nodes (:343-347)
  nodes_from_encoder(eS, 64, a.NT, nullptr, nullptr, B0, ld);
  gat_layer(B0, a.NT, ld, a.gatS, B1, S);
  graph_pool(B1, a.NT, ld, a.poolS, a.nS, B0, gi, gw, S);
  if (gi) gi += a.nS;
  if (gw) gw += a.NT;
  for (int t = threadIdx.x; t < 12 * D1; t += kGraphThreads) {
    int m = t / D1, d = t % D1;
    float s = 0.f;
    for (int n = 0; n < a.nS; ++n) s = fmaf(__ldg(a.projSW + m * a.nS + n), B0[n * ld + d], s);
    PS[m * ld + d] = s + __ldg(a.projSB + m);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 12 * D1; t += kGraphThreads) {           // :349
    int m = t / D1, d = t % D1;
    G[m * ld + d] = PT[m * ld + d] * PS[m * ld + d];
  }
  __syncthreads();
  gat_layer(G, 12, ld, a.gatST, Q, S);                                   // :351
  graph_pool(Q, 12, ld, a.poolST, a.nST, B0, gi, gw, S);                 // :352
  // proj_ST Linear(16,1) per node, then out_layer Linear(nST,2)         // :353-354
  float* pr = S.wts;
  for (int n = threadIdx.x; n < a.nST; n += kGraphThreads) {
    float s = 0.f;
    for (int d = 0; d < a.gatST.Do; ++d) s = fmaf(__ldg(a.projSTW + d), B0[n * ld + d], s);
    s += a.projSTB;
    pr[n] = s;
    a.last_hidden[(size_t)b * a.nST + n] = s;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int n = 0; n < a.nST; ++n) s = fmaf(__ldg(a.outW + threadIdx.x * a.nST + n), pr[n], s);
    a.logits[(size_t)b * 2 + threadIdx.x] = s + __ldg(a.outB + threadIdx.x);
  }
  if (a.tc) att_tc_end(S);
}

int launch_graph_rawgat(aasist_handle* h, const float* eT, const float* eS, int B, int NT,
                        float* last_hidden, float* logits, int32_t* topk, float* scores,
                        cudaStream_t st) {
  GraphArgsRawGat a = h->gr;
  a.NT = NT;
  a.tc = h->cfg.precision != AASIST_PREC_FP32;
  a.nT = pooled_count(kSpecNodes, 0.64, 2);
  a.nS = pooled_count(NT, 0.81, 2);
  a.nST = pooled_count(12, 0.64, 2);
  if (a.nT != 14 || a.nS != 23 || a.nST != 7) {
    set_error("RawGAT-ST is hard-wired to 64600-sample inputs (Linear(14,12)/Linear(23,12), "
              "RawNetGatSpoofST.py:319-320); got %d temporal nodes", NT);
    return AASIST_E_INVALID;
  }
  a.nmax = max(NT, kSpecNodes);
  a.ld = 65;
  a.eT = eT;
  a.eS = eS;
  a.last_hidden = last_hidden;
  a.logits = logits;
  a.topk_idx = topk;
  a.pool_scores = scores;
  a.topk_total = a.nT + a.nS + a.nST;
  a.score_total = kSpecNodes + NT + 12;
  size_t smem = sizeof(float) * (size_t)rawgat_graph_smem_floats(a.nmax, a.ld);
  AASIST_CUDA(cudaFuncSetAttribute(rawgat_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  {
    LaunchSpan span(h, "rawgat_graph", st);
    rawgat_graph_kernel<<<B, kGraphThreads, smem, st>>>(a);
  }
  AASIST_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace aasist
