// Fork-only encoders, fp32 CUDA-core path:
//   * Res2NetBlock + SELayer                      reference models/AASIST.py:506-669
//       x -> [bn1 -> SELU] -> split -> per-split 3x3 convs (split i with i % scale == 0, i > 0 also adds the RAW
//       output of split i-1) -> cat -> bn2 -> SELU -> conv_cat 3x3 -> SE gate (global mean -> FC -> ReLU -> FC ->
//       sigmoid) -> + identity | conv_downsample k(1,3) -> MaxPool2d((1,3))
//     Unlike the (2,3) Residual_block, bn1 + SELU on the input is LIVE here (:611-613).
//   * the fork's 3x3 Residual_block               reference models/AASIST.py:672-725 (AASIST-Robust encoder)
// NCHW fp32 throughout.  The split convs have 1..12 channels, so these are direct convolutions; conv_cat (the only
// large contraction, K = 9*Ci) is the shared-memory tiled kernel below.
// The SE gate needs the global mean of conv_cat's OUTPUT before the residual add.  That mean is linear in
// conv_cat's input: mean(y[co]) = b[co] + (1/HW) sum_{ci,dh,dw} W[co][ci][dh][dw] * S[ci][dh][dw], where S is the sum
// of the input plane over the window a tap sees (the whole plane minus one border row / column).  So the gate is
// computed from per-channel plane / border sums of conv_cat's INPUT (emitted by the split-conv kernel), BEFORE
// conv_cat runs, and conv_cat's epilogue applies gate, residual and max-pool directly: its un-pooled output
// (110 MB per utterance at 4 s) never exists.  Every reduction has a fixed order: results are deterministic.
#include "common.cuh"

namespace aasist {

constexpr int kTW3 = 96;     // output columns per CTA (multiple of 3)
constexpr int kCK3 = 8;      // input channels per shared-memory chunk
constexpr int kInLd3 = kTW3 + 2;

enum { M33_SELU = 0, M33_RES_ID = 2, M33_RES_DS = 3, M33_GATE_ID = 4, M33_GATE_DS = 5 };

template <int CO_T>
__device__ __forceinline__ void load_w3(float (&w)[CO_T], const float* p) {
#pragma unroll
  for (int q = 0; q < CO_T; q += 4) {
    float4 t = *reinterpret_cast<const float4*>(p + q);
    w[q] = t.x; w[q + 1] = t.y; w[q + 2] = t.z; w[q + 3] = t.w;
  }
}

// 3x3 convolution, padding (1,1), H = 23 rows.
//   in    (B, Ci, 23, W)      wmain [Ci][3][3][Cop]   bias [Cop]   Cop = 8*CO_T >= Co
//   M33_SELU   : out (B,Co,23,W)   = selu(conv + bias)                          (conv1 of the 3x3 block, bn2 folded)
//   M33_RES_ID : out (B,Co,23,W/3) = maxpool3(conv + bias + side)               side = block input (B,Co,23,W)
//   M33_RES_DS : out (B,Co,23,W/3) = maxpool3(conv + bias + conv_downsample(side)),  wside [Cs][3][Cop]
//   M33_GATE_ID: out (B,Co,23,W/3) = maxpool3(gate[b][co] * (conv + bias) + side)          (Res2Net conv_cat + SE)
//   M33_GATE_DS: out (B,Co,23,W/3) = maxpool3(gate[b][co] * (conv + bias) + conv_downsample(side) + bias2)
template <int CO_T, int MODE>
__global__ void __launch_bounds__(256)
conv3x3_f32_kernel(const float* __restrict__ in, const float* __restrict__ side,
                   const float* __restrict__ wmain, const float* __restrict__ wside,
                   const float* __restrict__ bias, float* __restrict__ out, const float* __restrict__ gate,
                   const float* __restrict__ bias2, int Ci, int Cs, int Co, int W) {
  constexpr int Cop = 8 * CO_T;
  constexpr int H = kSpecNodes;
  __shared__ __align__(16) float s_in[kCK3 * 3 * kInLd3];
  __shared__ __align__(16) float s_w[kCK3 * 9 * Cop];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int w0 = blockIdx.x * kTW3;
  const int row = blockIdx.y;
  const int b = blockIdx.z;

  float acc[3][CO_T];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int c = 0; c < CO_T; ++c) acc[j][c] = 0.f;

  const float* inb = in + (size_t)b * Ci * H * W;
  for (int c0 = 0; c0 < Ci; c0 += kCK3) {
    const int nc = min(kCK3, Ci - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < nc * 3 * kInLd3; i += 256) {
      int col = i % kInLd3, r = (i / kInLd3) % 3, c = i / (3 * kInLd3);
      int gr = row - 1 + r, gw = w0 - 1 + col;
      float v = 0.f;
      if (gr >= 0 && gr < H && gw >= 0 && gw < W) v = inb[((size_t)(c0 + c) * H + gr) * W + gw];
      s_in[i] = v;
    }
    for (int i = threadIdx.x; i < nc * 9 * Cop / 4; i += 256)
      reinterpret_cast<float4*>(s_w)[i] = reinterpret_cast<const float4*>(wmain + (size_t)c0 * 9 * Cop)[i];
    __syncthreads();
    for (int c = 0; c < nc; ++c) {
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const float* ip = s_in + (c * 3 + dh) * kInLd3 + 3 * tx;
        float v[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) v[q] = ip[q];
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          float w[CO_T];
          load_w3<CO_T>(w, s_w + ((c * 3 + dh) * 3 + dw) * Cop + ty * CO_T);
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int q = 0; q < CO_T; ++q) acc[j][q] = fmaf(v[j + dw], w[q], acc[j][q]);
        }
      }
    }
  }
  float acc2[3][CO_T];                   // M33_GATE_DS: the downsample conv is NOT gated: its own accumulators
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int c = 0; c < CO_T; ++c) acc2[j][c] = 0.f;
  if (MODE == M33_RES_DS || MODE == M33_GATE_DS) {
    const float* sb = side + (size_t)b * Cs * H * W;
    for (int c0 = 0; c0 < Cs; c0 += kCK3) {
      const int nc = min(kCK3, Cs - c0);
      __syncthreads();
      for (int i = threadIdx.x; i < nc * kInLd3; i += 256) {
        int col = i % kInLd3, c = i / kInLd3;
        int gw = w0 - 1 + col;
        float v = 0.f;
        if (gw >= 0 && gw < W) v = sb[((size_t)(c0 + c) * H + row) * W + gw];
        s_in[i] = v;
      }
      for (int i = threadIdx.x; i < nc * 3 * Cop / 4; i += 256)
        reinterpret_cast<float4*>(s_w)[i] = reinterpret_cast<const float4*>(wside + (size_t)c0 * 3 * Cop)[i];
      __syncthreads();
      for (int c = 0; c < nc; ++c) {
        const float* ip = s_in + c * kInLd3 + 3 * tx;
        float v[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) v[q] = ip[q];
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          float w[CO_T];
          load_w3<CO_T>(w, s_w + (c * 3 + dw) * Cop + ty * CO_T);
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int q = 0; q < CO_T; ++q) {
              if (MODE == M33_GATE_DS) acc2[j][q] = fmaf(v[j + dw], w[q], acc2[j][q]);
              else acc[j][q] = fmaf(v[j + dw], w[q], acc[j][q]);
            }
        }
      }
    }
  }

  const int wbase = w0 + 3 * tx;
  if (MODE == M33_SELU) {
    float* ob = out + (size_t)b * Co * H * W;
#pragma unroll
    for (int q = 0; q < CO_T; ++q) {
      const int co = ty * CO_T + q;              // warp-uniform
      if (co < Co) {
        const float bq = bias[co];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (wbase + j < W) ob[((size_t)co * H + row) * W + wbase + j] = selu(acc[j][q] + bq);
      }
    }
  } else {
    const int Wo = W / 3;
    const int po = blockIdx.x * (kTW3 / 3) + tx;
    if (po >= Wo) return;
    float* ob = out + (size_t)b * Co * H * Wo;
    const float* sb = side + (size_t)b * Cs * H * W;
#pragma unroll
    for (int q = 0; q < CO_T; ++q) {
      const int co = ty * CO_T + q;
      if (co >= Co) continue;
      const float bq = bias[co];
      const float g = (MODE == M33_GATE_ID || MODE == M33_GATE_DS) ? gate[(size_t)b * Co + co] : 1.f;
      const float b2 = MODE == M33_GATE_DS ? bias2[co] : 0.f;
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float v = acc[j][q] + bq;
        if (MODE == M33_RES_ID) v += sb[((size_t)co * H + row) * W + wbase + j];
        if (MODE == M33_GATE_ID) v = fmaf(v, g, sb[((size_t)co * H + row) * W + wbase + j]);
        if (MODE == M33_GATE_DS) v = fmaf(v, g, acc2[j][q] + b2);
        m = fmaxf(m, v);
      }
      ob[((size_t)co * H + row) * Wo + po] = m;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Res2Net split convolutions (models/AASIST.py:627-643) + bn2 + SELU (:653-654)
// CTA = 128 columns of one row of one (utterance, split); the split's input tile (n channels x 3 rows,
// after bn1/SELU and the scale-group addend) is staged in shared memory, every thread produces the n output
// channels of its pixel.  Only splits of level `level` run in a launch (their addend comes from level-1).
// Also emits msum[b][c][row][tile] = sum of the tile's `mid` values (fixed-order block reduction): the plane and
// border-row sums the SE gate is computed from.
// ---------------------------------------------------------------------------------------
constexpr int kGW = 128;        // threads = staged columns per CTA: 126 output columns + one halo column each side
constexpr int kGOut = kGW - 2;
constexpr int kGLd = kGW;

// SELU through MUFU.EX2 (|abs error| <= ~2.5e-7, the same formulation as the tensor-core epilogues): the split convs
// apply SELU four times per element (bn1 on each of the three row tiles that stage an input row, bn2 on the output)
__device__ __forceinline__ float selu_fast(float v) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
  const float n = fminf(fmaf(e, kSeluScale * kSeluAlpha, -(kSeluScale * kSeluAlpha)), 0.f);
  return fmaf(fmaxf(v, 0.f), kSeluScale, n);
}

// NMAX >= n: accumulators of all n output channels live in registers (each staged input value is read once and
// used for every output channel; weights are staged transposed, [ic][tap][NMAX], and read as broadcasts).
// The CTA owns one (utterance, row, 128-column tile) and walks the splits of this launch (`lvl_groups`: the splits
// of one dependency level and accumulator width), one small tile (n x 3 rows x 130 columns) at a time: the kernel
// is bound by global-load latency, so it keeps its footprint small (16 CTAs per SM) and issues all loads of a tile
// before consuming any.  [measured alternatives, profiles/README.md: one CTA per split = 0.5 M two-channel CTAs per
// launch, same speed; all splits' tiles staged at once = 39 KB per CTA, 31 % occupancy, 1.5x slower]
template <int NMAX>
__global__ void __launch_bounds__(kGW)
res2_group_conv_kernel(const float* __restrict__ x, float* __restrict__ raw, float* __restrict__ mid,
                       float* __restrict__ msum, const Res2Group* __restrict__ groups,
                       const int* __restrict__ lvl_groups, int n_lvl, const float* __restrict__ bn1,
                       const float* __restrict__ gw, const int* __restrict__ gw_off, const float* __restrict__ gb,
                       const float* __restrict__ bn2, int Ci, int W) {
  extern __shared__ float s_dyn[];
  __shared__ float s_red[kGW / 32][NMAX];
  constexpr int H = kSpecNodes;
  const int b = blockIdx.z;
  const int row = blockIdx.y, w0 = blockIdx.x * kGOut;
  const int planei = H * W;                      // < 2^19; Ci * planei < 2^25: 32-bit offsets inside an utterance
  const size_t plane = (size_t)planei;
  const float* xb = x + (size_t)b * Ci * plane;
  const float* rb = raw + (size_t)b * Ci * plane;
  float* rawb = raw + (size_t)b * Ci * plane;
  float* midb = mid + (size_t)b * Ci * plane;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread t stages tile column t = input column w0 - 1 + t (clamped for the load, zeroed when outside) and, for
  // 1 <= t <= 126, produces output column w0 + t - 1
  const int gc = w0 - 1 + (int)threadIdx.x;
  const bool c_ok = gc >= 0 && gc < W;
  const int gcc = min(max(gc, 0), W - 1);
  int roff[3];
  bool rok[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int gr = row - 1 + r;
    rok[r] = c_ok && gr >= 0 && gr < H;
    roff[r] = min(max(gr, 0), H - 1) * W + gcc;
  }
  const int col = gc;
  const bool live = threadIdx.x >= 1 && threadIdx.x <= kGOut && col < W;
  constexpr bool STAGE_W = NMAX <= 32;           // wider splits (res2net_width 1 or 2): weights stay in L1

  for (int gi = 0; gi < n_lvl; ++gi) {
    const int g = lvl_groups[gi];
    const Res2Group G = groups[g];
    const int n = G.n;
    float* s_t = s_dyn;                          // [n][3][kGLd]
    float* s_w = s_dyn + n * 3 * kGLd;           // [n][9][NMAX]
    const int cprev = G.level > 0 ? groups[g - 1].c0 : -1;
    const float* wg = gw + gw_off[g];            // [oc][ic][9]
    // ---- stage the input tile: bn1 + SELU, scale-group addend, zero padding of the conv INPUT.
    //      Two channels (six tile rows) per batch: all loads are issued before any is consumed.
    for (int ic0 = 0; ic0 < n; ic0 += 2) {
      float v[6], a[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int ic = min(ic0 + u / 3, n - 1);
        v[u] = xb[(G.c0 + ic) * planei + roff[u % 3]];
        a[u] = cprev >= 0 ? rb[(cprev + ic) * planei + roff[u % 3]] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int ic = ic0 + u / 3;
        if (ic < n) {
          float xv = v[u];
          if (bn1) xv = selu_fast(fmaf(xv, bn1[G.c0 + ic], bn1[Ci + G.c0 + ic]));
          s_t[(ic * 3 + u % 3) * kGLd + threadIdx.x] = rok[u % 3] ? xv + a[u] : 0.f;
        }
      }
    }
    if (STAGE_W)
      for (int i = threadIdx.x; i < n * 9 * NMAX; i += kGW) {
        const int oc = i % NMAX, k = (i / NMAX) % 9, ic = i / (9 * NMAX);
        s_w[i] = oc < n ? __ldg(wg + ((size_t)oc * n + ic) * 9 + k) : 0.f;
      }
    __syncthreads();
    // ---- n x n x 3 x 3 MACs per pixel
    float acc[NMAX];
#pragma unroll
    for (int oc = 0; oc < NMAX; ++oc) acc[oc] = 0.f;
    for (int ic = 0; ic < n; ++ic) {
      const float* t = s_t + ic * 3 * kGLd + min(max((int)threadIdx.x - 1, 0), kGLd - 3);
      const float* wk = s_w + ic * 9 * NMAX;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh)
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          const float v = t[dh * kGLd + dw];
#pragma unroll
          for (int oc = 0; oc < NMAX; ++oc) {
            const float w = STAGE_W ? wk[(dh * 3 + dw) * NMAX + oc]
                                    : (oc < n ? __ldg(wg + ((size_t)oc * n + ic) * 9 + dh * 3 + dw) : 0.f);
            acc[oc] = fmaf(v, w, acc[oc]);
          }
        }
    }
    // ---- raw output for the dependent split, bn2 + SELU -> mid, tile sums for the SE gate
#pragma unroll
    for (int oc = 0; oc < NMAX; ++oc) {
      if (oc < n) {                              // CTA-uniform
        const int c = G.c0 + oc;
        const float r = acc[oc] + __ldg(gb + c);
        const float m = live ? selu_fast(fmaf(r, __ldg(bn2 + c), __ldg(bn2 + Ci + c))) : 0.f;
        if (live) {
          const int o = c * planei + row * W + col;
          if (G.feeds_next) rawb[o] = r;
          midb[o] = m;
        }
        // tile sum of mid[c][row][w0 .. w0+126): butterfly inside each warp, then the four warp sums in fixed order
        float sum = m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) s_red[warp][oc] = sum;
      }
    }
    __syncthreads();                             // s_red complete; s_t / s_w free for the next split
    if (threadIdx.x < n) {
      const int oc = threadIdx.x;
      msum[(((size_t)b * Ci + G.c0 + oc) * H + row) * gridDim.x + blockIdx.x] =
          (s_red[0][oc] + s_red[1][oc]) + (s_red[2][oc] + s_red[3][oc]);
    }
  }
}

// SELayer gate (models/AASIST.py:518-522) WITHOUT materialising conv_cat's output: for every input channel the
// window sums S[dh][dw] (plane sum minus the border row / column a tap cannot reach, plus the corner both exclude)
// come from msum and the border columns of `mid`; mean(y[co]) = b[co] + (1/HW) sum W[co][ci][dh][dw] S[ci][dh][dw];
// then FC (co/16 x co, no bias) -> ReLU -> FC (co x co/16) -> sigmoid.  One CTA per utterance.
__global__ void __launch_bounds__(256)
se_gate_kernel(const float* __restrict__ mid, const float* __restrict__ msum, int ntile,
               const float* __restrict__ wcat, const float* __restrict__ bcat, int Cop, const float* __restrict__ w0,
               const float* __restrict__ w2, int Ci, int Co, int W, int hidden, float* __restrict__ gate) {
  constexpr int H = kSpecNodes;
  __shared__ float s_S[64 * 9];                   // [ci][dh][dw]
  __shared__ float s_mean[64];
  __shared__ float s_hid[8];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t plane = (size_t)H * W;
  for (int c = warp; c < Ci; c += 8) {
    const float* ps = msum + ((size_t)b * Ci + c) * H * ntile;
    const float* pm = mid + ((size_t)b * Ci + c) * plane;
    // plane sum T and the two border-row sums, fixed order: lanes stride the (row, tile) grid, then a butterfly
    float T = 0.f, R0 = 0.f, R22 = 0.f, C0 = 0.f, CW = 0.f;
    for (int i = lane; i < H * ntile; i += 32) {
      const float v = ps[i];
      T += v;
      if (i < ntile) R0 += v;
      if (i >= (H - 1) * ntile) R22 += v;
    }
    for (int r = lane; r < H; r += 32) {
      C0 += pm[(size_t)r * W];
      CW += pm[(size_t)r * W + W - 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      T += __shfl_xor_sync(0xffffffffu, T, o);
      R0 += __shfl_xor_sync(0xffffffffu, R0, o);
      R22 += __shfl_xor_sync(0xffffffffu, R22, o);
      C0 += __shfl_xor_sync(0xffffffffu, C0, o);
      CW += __shfl_xor_sync(0xffffffffu, CW, o);
    }
    if (lane < 9) {
      const int dh = lane / 3, dw = lane % 3;
      // tap (dh,dw) reads input row h+dh-1 / column w+dw-1: dh = 0 never reaches row 22, dh = 2 never row 0
      const float rex = dh == 0 ? R22 : (dh == 2 ? R0 : 0.f);
      const float cex = dw == 0 ? CW : (dw == 2 ? C0 : 0.f);
      float corner = 0.f;
      if (dh != 1 && dw != 1)
        corner = pm[(size_t)(dh == 0 ? H - 1 : 0) * W + (dw == 0 ? W - 1 : 0)];
      s_S[c * 9 + lane] = T - rex - cex + corner;
    }
  }
  __syncthreads();
  const float inv = 1.f / (float)plane;
  for (int co = warp; co < Co; co += 8) {
    float s = 0.f;
    for (int i = lane; i < Ci * 9; i += 32) s = fmaf(__ldg(wcat + (size_t)i * Cop + co), s_S[i], s);   // wcat [ci][3][3][cop]
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_mean[co] = fmaf(s, inv, __ldg(bcat + co));
  }
  __syncthreads();
  if (threadIdx.x < hidden) {
    float s = 0.f;
    for (int c = 0; c < Co; ++c) s = fmaf(__ldg(w0 + threadIdx.x * Co + c), s_mean[c], s);
    s_hid[threadIdx.x] = fmaxf(s, 0.f);
  }
  __syncthreads();
  if (threadIdx.x < Co) {
    float s = 0.f;
    for (int j = 0; j < hidden; ++j) s = fmaf(__ldg(w2 + threadIdx.x * hidden + j), s_hid[j], s);
    gate[(size_t)b * Co + threadIdx.x] = 1.f / (1.f + expf(-s));
  }
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
static inline size_t al64(size_t v) { return (v + 63) & ~(size_t)63; }

// scratch of one Res2Net block for B utterances: mid, raw (B,Ci,23,W), tile sums of mid, gates
size_t res2_block_scratch_floats(const Res2BlockF32& blk, int B, int W) {
  const size_t plane = (size_t)kSpecNodes * W;
  const size_t ntile = (W + kGOut - 1) / kGOut;
  return 2 * al64((size_t)B * blk.ci * plane) + al64((size_t)B * blk.ci * kSpecNodes * ntile) +
         al64((size_t)B * blk.co);
}

template <int CO_T, int MODE>
static int launch_conv33(aasist_handle* h, const char* name, const float* in, const float* side, const float* wmain,
                         const float* wside, const float* bias, float* out, const float* gate, const float* bias2,
                         int Ci, int Cs, int Co, int W, int B, cudaStream_t st) {
  dim3 grid((W + kTW3 - 1) / kTW3, kSpecNodes, B);
  {
    LaunchSpan span(h, name, st);
    conv3x3_f32_kernel<CO_T, MODE><<<grid, 256, 0, st>>>(in, side, wmain, wside, bias, out, gate, bias2, Ci, Cs, Co, W);
  }
  AASIST_CUDA(cudaGetLastError());
  return 0;
}

static const char* kSplitNames[6] = {"enc0.res2_split_convs_f32", "enc1.res2_split_convs_f32", "enc2.res2_split_convs_f32",
                                     "enc3.res2_split_convs_f32", "enc4.res2_split_convs_f32", "enc5.res2_split_convs_f32"};
static const char* kCatNames[6] = {"enc0.res2_conv_cat_se_res_pool_f32", "enc1.res2_conv_cat_se_res_pool_f32",
                                   "enc2.res2_conv_cat_se_res_pool_f32", "enc3.res2_conv_cat_se_res_pool_f32",
                                   "enc4.res2_conv_cat_se_res_pool_f32", "enc5.res2_conv_cat_se_res_pool_f32"};

template <int CO_T>
static int launch_res2_cat(aasist_handle* h, const Res2BlockF32& blk, const float* mid, const float* in,
                           const float* gate, float* out, int W, int B, cudaStream_t st) {
  if (blk.downsample)
    return launch_conv33<CO_T, M33_GATE_DS>(h, kCatNames[blk.index % 6], mid, in, blk.wcat, blk.wd, blk.bcat, out,
                                            gate, blk.bd, blk.ci, blk.ci, blk.co, W, B, st);
  return launch_conv33<CO_T, M33_GATE_ID>(h, kCatNames[blk.index % 6], mid, in, blk.wcat, nullptr, blk.bcat, out,
                                          gate, nullptr, blk.ci, blk.co, blk.co, W, B, st);
}

int launch_res2_block(aasist_handle* h, const Res2BlockF32& blk, const float* in, int B, int W, float* ws,
                      float* out, cudaStream_t st) {
  if (W < 3) {
    set_error("encoder block input width %d < 3 (MaxPool2d((1,3)) would be empty)", W);
    return AASIST_E_INVALID;
  }
  const size_t plane = (size_t)kSpecNodes * W;
  const int ntile = (W + kGOut - 1) / kGOut;
  float* mid = ws;
  float* raw = mid + al64((size_t)B * blk.ci * plane);
  float* msum = raw + al64((size_t)B * blk.ci * plane);
  float* gate = msum + al64((size_t)B * blk.ci * kSpecNodes * ntile);
  // one launch per (dependency level, accumulator width): only that level's splits, grouped by how many output
  // channels their threads keep in registers
  for (const Res2Launch& L : blk.launches) {
    dim3 grid(ntile, kSpecNodes, B);
    const size_t smem = sizeof(float) * ((size_t)L.n_max * 3 * kGLd + (L.nreg <= 32 ? (size_t)L.n_max * 9 * L.nreg : 0));
    const int* lg = blk.lvl_groups_dev + L.first;
#define AASIST_RES2_LAUNCH(NR)                                                                                       \
  {                                                                                                                  \
    AASIST_CUDA(cudaFuncSetAttribute(res2_group_conv_kernel<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                     (int)smem));                                                                    \
    LaunchSpan span(h, kSplitNames[blk.index % 6], st);                                                              \
    res2_group_conv_kernel<NR><<<grid, kGW, smem, st>>>(in, raw, mid, msum, blk.groups_dev, lg, L.count, blk.bn1,    \
                                                       blk.gw, blk.gw_off, blk.gb, blk.bn2, blk.ci, W);             \
  }
    switch (L.nreg) {
      case 1: AASIST_RES2_LAUNCH(1) break;
      case 2: AASIST_RES2_LAUNCH(2) break;
      case 4: AASIST_RES2_LAUNCH(4) break;
      case 6: AASIST_RES2_LAUNCH(6) break;
      case 8: AASIST_RES2_LAUNCH(8) break;
      case 12: AASIST_RES2_LAUNCH(12) break;
      case 16: AASIST_RES2_LAUNCH(16) break;
      case 32: AASIST_RES2_LAUNCH(32) break;
      default: AASIST_RES2_LAUNCH(64) break;
    }
#undef AASIST_RES2_LAUNCH
  }
  AASIST_CUDA(cudaGetLastError());
  const int cop = blk.co <= 32 ? 32 : 64;
  {
    LaunchSpan span(h, "se_gate", st);
    se_gate_kernel<<<B, 256, 0, st>>>(mid, msum, ntile, blk.wcat, blk.bcat, cop, blk.se0, blk.se2, blk.ci, blk.co, W,
                                     blk.se_hidden, gate);
  }
  AASIST_CUDA(cudaGetLastError());
  if (blk.co <= 32) return launch_res2_cat<4>(h, blk, mid, in, gate, out, W, B, st);
  return launch_res2_cat<8>(h, blk, mid, in, gate, out, W, B, st);
}

template <int CO_T>
static int launch_block33_t(aasist_handle* h, const ConvBlock33F32& blk, const float* in, int B, int W, float* mid,
                            float* out, cudaStream_t st) {
  int rc = launch_conv33<CO_T, M33_SELU>(h, "conv1_3x3_f32", in, nullptr, blk.w1, nullptr, blk.b1, mid, nullptr,
                                         nullptr, blk.ci, 0, blk.co, W, B, st);
  if (rc) return rc;
  if (blk.downsample)
    return launch_conv33<CO_T, M33_RES_DS>(h, "conv2_3x3_res_pool_f32", mid, in, blk.w2, blk.wd, blk.b2, out, nullptr,
                                           nullptr, blk.co, blk.ci, blk.co, W, B, st);
  return launch_conv33<CO_T, M33_RES_ID>(h, "conv2_3x3_res_pool_f32", mid, in, blk.w2, nullptr, blk.b2, out, nullptr,
                                         nullptr, blk.co, blk.co, blk.co, W, B, st);
}

// in (B,ci,23,W) -> mid (B,co,23,W) scratch -> out (B,co,23,W/3)
int launch_block33_f32(aasist_handle* h, const ConvBlock33F32& blk, const float* in, int B, int W, float* mid,
                       float* out, cudaStream_t st) {
  if (W < 3) {
    set_error("encoder block input width %d < 3 (MaxPool2d((1,3)) would be empty)", W);
    return AASIST_E_INVALID;
  }
  if (blk.co <= 32) return launch_block33_t<4>(h, blk, in, B, W, mid, out, st);
  return launch_block33_t<8>(h, blk, in, B, W, mid, out, st);
}

}  // namespace aasist
