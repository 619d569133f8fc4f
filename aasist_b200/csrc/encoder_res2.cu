// Fork-only encoders, fp32 CUDA-core path:
//   * Res2NetBlock + SELayer                      reference models/AASIST.py:506-669
//       x -> [bn1 -> SELU] -> split -> per-split 3x3 convs (split i with i % scale == 0, i > 0 also adds the RAW
//       output of split i-1) -> cat -> bn2 -> SELU -> conv_cat 3x3 -> SE gate (global mean -> FC -> ReLU -> FC ->
//       sigmoid) -> + identity | conv_downsample k(1,3) -> MaxPool2d((1,3))
//     Unlike the (2,3) Residual_block, bn1 + SELU on the input is LIVE here (:611-613).
//   * the fork's 3x3 Residual_block               reference models/AASIST.py:672-725 (AASIST-Robust encoder)
// NCHW fp32 throughout.  The split convs have 1..12 channels and the SE gate needs a global mean before the
// residual add, so these are direct convolutions; conv_cat (the only large contraction, K = 9*Ci) is the
// shared-memory tiled kernel below.  Every reduction has a fixed order: results are run-to-run deterministic.
#include "common.cuh"

namespace aasist {

constexpr int kTW3 = 96;     // output columns per CTA (multiple of 3)
constexpr int kCK3 = 8;      // input channels per shared-memory chunk
constexpr int kInLd3 = kTW3 + 2;

enum { M33_SELU = 0, M33_SUM = 1, M33_RES_ID = 2, M33_RES_DS = 3 };

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
//   M33_SUM    : out (B,Co,23,W)   = conv + bias;  partial[b][co][row][tile] = sum of the tile's valid columns
//                                                                               (conv_cat; feeds the SE mean)
//   M33_RES_ID : out (B,Co,23,W/3) = maxpool3(conv + bias + side)               side = block input (B,Co,23,W)
//   M33_RES_DS : out (B,Co,23,W/3) = maxpool3(conv + bias + conv_downsample(side)),  wside [Cs][3][Cop]
template <int CO_T, int MODE>
__global__ void __launch_bounds__(256)
conv3x3_f32_kernel(const float* __restrict__ in, const float* __restrict__ side,
                   const float* __restrict__ wmain, const float* __restrict__ wside,
                   const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ partial,
                   int Ci, int Cs, int Co, int W) {
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
  if (MODE == M33_RES_DS) {
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
            for (int q = 0; q < CO_T; ++q) acc[j][q] = fmaf(v[j + dw], w[q], acc[j][q]);
        }
      }
    }
  }

  const int wbase = w0 + 3 * tx;
  if (MODE == M33_SELU || MODE == M33_SUM) {
    float* ob = out + (size_t)b * Co * H * W;
#pragma unroll
    for (int q = 0; q < CO_T; ++q) {
      const int co = ty * CO_T + q;              // warp-uniform
      float s = 0.f;
      if (co < Co) {
        const float bq = bias[co];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (wbase + j < W) {
            float v = acc[j][q] + bq;
            if (MODE == M33_SELU) v = selu(v);
            ob[((size_t)co * H + row) * W + wbase + j] = v;
            s += v;
          }
      }
      if (MODE == M33_SUM) {
        // fixed-order butterfly over the 32 column triples of this tile (deterministic)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tx == 0 && co < Co)
          partial[(((size_t)b * Co + co) * H + row) * gridDim.x + blockIdx.x] = s;
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
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float v = acc[j][q] + bq;
        if (MODE == M33_RES_ID) v += sb[((size_t)co * H + row) * W + wbase + j];
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
// ---------------------------------------------------------------------------------------
constexpr int kGW = 128;
constexpr int kGLd = kGW + 2;

__global__ void __launch_bounds__(kGW)
res2_group_conv_kernel(const float* __restrict__ x, float* __restrict__ raw, float* __restrict__ mid,
                       const Res2Group* __restrict__ groups, int n_groups, int level,
                       const float* __restrict__ bn1, const float* __restrict__ gw, const int* __restrict__ gw_off,
                       const float* __restrict__ gb, const float* __restrict__ bn2, int Ci, int W) {
  extern __shared__ float s_t[];                 // [n][3][kGLd]
  constexpr int H = kSpecNodes;
  const int g = blockIdx.z % n_groups, b = blockIdx.z / n_groups;
  const Res2Group G = groups[g];
  if (G.level != level) return;
  const int row = blockIdx.y, w0 = blockIdx.x * kGW;
  const int n = G.n;
  const size_t plane = (size_t)H * W;
  const float* xb = x + (size_t)b * Ci * plane;
  const float* rb = raw + (size_t)b * Ci * plane;
  const int cprev = G.level > 0 ? groups[g - 1].c0 : 0;
  for (int i = threadIdx.x; i < n * 3 * kGLd; i += kGW) {
    const int col = i % kGLd, r = (i / kGLd) % 3, ic = i / (3 * kGLd);
    const int gr = row - 1 + r, gc = w0 - 1 + col;
    float v = 0.f;                               // zero padding applies to the conv INPUT (after bn1/SELU/addend)
    if (gr >= 0 && gr < H && gc >= 0 && gc < W) {
      const size_t off = (size_t)gr * W + gc;
      v = xb[(size_t)(G.c0 + ic) * plane + off];
      if (bn1) v = selu(fmaf(v, bn1[G.c0 + ic], bn1[Ci + G.c0 + ic]));
      if (G.level > 0) v += rb[(size_t)(cprev + ic) * plane + off];
    }
    s_t[i] = v;
  }
  __syncthreads();
  const int col = w0 + threadIdx.x;
  if (col >= W) return;
  const float* wg = gw + gw_off[g];
  float* rawb = raw + (size_t)b * Ci * plane;
  float* midb = mid + (size_t)b * Ci * plane;
  for (int oc = 0; oc < n; ++oc) {
    float acc = 0.f;
    const float* wo = wg + (size_t)oc * n * 9;
    for (int ic = 0; ic < n; ++ic) {
      const float* t = s_t + ic * 3 * kGLd + threadIdx.x;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh)
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) acc = fmaf(t[dh * kGLd + dw], __ldg(wo + ic * 9 + dh * 3 + dw), acc);
    }
    const int c = G.c0 + oc;
    const float r = acc + __ldg(gb + c);
    const size_t o = (size_t)c * plane + (size_t)row * W + col;
    if (G.feeds_next) rawb[o] = r;
    midb[o] = selu(fmaf(r, __ldg(bn2 + c), __ldg(bn2 + Ci + c)));
  }
}

// SELayer gate (models/AASIST.py:518-522): mean over (H,W) from the conv_cat partial sums (fixed order),
// FC (co/16 x co, no bias) -> ReLU -> FC (co x co/16) -> sigmoid.  One CTA per utterance.
__global__ void __launch_bounds__(256)
se_gate_kernel(const float* __restrict__ partial, int n_part, float inv_count, const float* __restrict__ w0,
               const float* __restrict__ w2, int Co, int hidden, float* __restrict__ gate) {
  __shared__ float s_mean[64];
  __shared__ float s_hid[8];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < Co; c += 8) {
    const float* p = partial + ((size_t)b * Co + c) * n_part;
    float s = 0.f;
    for (int i = lane; i < n_part; i += 32) s += p[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_mean[c] = s * inv_count;
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

// out = MaxPool2d((1,3))( y * gate + (conv_downsample(x) | x) )        (models/AASIST.py:658-668)
__global__ void __launch_bounds__(128)
res2_finish_kernel(const float* __restrict__ y, const float* __restrict__ gate, const float* __restrict__ x,
                   const float* __restrict__ wd, const float* __restrict__ bd, float* __restrict__ out,
                   int Ci, int Co, int W) {
  constexpr int H = kSpecNodes;
  const int Wo = W / 3;
  const int p = blockIdx.x * 128 + threadIdx.x;
  if (p >= Wo) return;
  const int row = blockIdx.y;
  const int co = blockIdx.z % Co, b = blockIdx.z / Co;
  const float g = gate[(size_t)b * Co + co];
  const float* yr = y + (((size_t)b * Co + co) * H + row) * W + 3 * p;
  float id[3];
  if (wd) {
    id[0] = id[1] = id[2] = __ldg(bd + co);
    for (int ci = 0; ci < Ci; ++ci) {
      const float* xr = x + (((size_t)b * Ci + ci) * H + row) * W;
      float v[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int c = 3 * p - 1 + q;
        v[q] = (c >= 0 && c < W) ? xr[c] : 0.f;
      }
      const float* wp = wd + ((size_t)co * Ci + ci) * 3;
      const float k0 = __ldg(wp), k1 = __ldg(wp + 1), k2 = __ldg(wp + 2);
#pragma unroll
      for (int j = 0; j < 3; ++j) id[j] = fmaf(k2, v[j + 2], fmaf(k1, v[j + 1], fmaf(k0, v[j], id[j])));
    }
  } else {
    const float* xr = x + (((size_t)b * Ci + co) * H + row) * W + 3 * p;
#pragma unroll
    for (int j = 0; j < 3; ++j) id[j] = xr[j];
  }
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 3; ++j) m = fmaxf(m, fmaf(yr[j], g, id[j]));
  out[(((size_t)b * Co + co) * H + row) * Wo + p] = m;
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
static inline size_t al64(size_t v) { return (v + 63) & ~(size_t)63; }

// scratch of one Res2Net block for B utterances: mid, raw (B,Ci,23,W), y (B,Co,23,W), SE partials, gates
size_t res2_block_scratch_floats(const Res2BlockF32& blk, int B, int W) {
  const size_t plane = (size_t)kSpecNodes * W;
  const size_t ntile = (W + kTW3 - 1) / kTW3;
  return 2 * al64((size_t)B * blk.ci * plane) + al64((size_t)B * blk.co * plane) +
         al64((size_t)B * blk.co * kSpecNodes * ntile) + al64((size_t)B * blk.co);
}

template <int CO_T, int MODE>
static int launch_conv33(aasist_handle* h, const char* name, const float* in, const float* side, const float* wmain,
                         const float* wside, const float* bias, float* out, float* partial, int Ci, int Cs, int Co,
                         int W, int B, cudaStream_t st) {
  dim3 grid((W + kTW3 - 1) / kTW3, kSpecNodes, B);
  {
    LaunchSpan span(h, name, st);
    conv3x3_f32_kernel<CO_T, MODE><<<grid, 256, 0, st>>>(in, side, wmain, wside, bias, out, partial, Ci, Cs, Co, W);
  }
  AASIST_CUDA(cudaGetLastError());
  return 0;
}

int launch_res2_block(aasist_handle* h, const Res2BlockF32& blk, const float* in, int B, int W, float* ws,
                      float* out, cudaStream_t st) {
  if (W < 3) {
    set_error("encoder block input width %d < 3 (MaxPool2d((1,3)) would be empty)", W);
    return AASIST_E_INVALID;
  }
  const size_t plane = (size_t)kSpecNodes * W;
  const int ntile = (W + kTW3 - 1) / kTW3;
  float* mid = ws;
  float* raw = mid + al64((size_t)B * blk.ci * plane);
  float* y = raw + al64((size_t)B * blk.ci * plane);
  float* partial = y + al64((size_t)B * blk.co * plane);
  float* gate = partial + al64((size_t)B * blk.co * kSpecNodes * ntile);
  int nmax = 1;
  for (const Res2Group& g : blk.groups) nmax = std::max(nmax, g.n);
  const size_t smem = sizeof(float) * (size_t)nmax * 3 * kGLd;
  AASIST_CUDA(cudaFuncSetAttribute(res2_group_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int level = 0; level < blk.n_levels; ++level) {
    dim3 grid((W + kGW - 1) / kGW, kSpecNodes, B * blk.n_groups);
    LaunchSpan span(h, "res2_split_convs_f32", st);
    res2_group_conv_kernel<<<grid, kGW, smem, st>>>(in, raw, mid, blk.groups_dev, blk.n_groups, level, blk.bn1,
                                                   blk.gw, blk.gw_off, blk.gb, blk.bn2, blk.ci, W);
  }
  AASIST_CUDA(cudaGetLastError());
  int rc;
  if (blk.co <= 32)
    rc = launch_conv33<4, M33_SUM>(h, "res2_conv_cat_f32", mid, nullptr, blk.wcat, nullptr, blk.bcat, y, partial,
                                   blk.ci, 0, blk.co, W, B, st);
  else
    rc = launch_conv33<8, M33_SUM>(h, "res2_conv_cat_f32", mid, nullptr, blk.wcat, nullptr, blk.bcat, y, partial,
                                   blk.ci, 0, blk.co, W, B, st);
  if (rc) return rc;
  {
    LaunchSpan span(h, "se_gate", st);
    se_gate_kernel<<<B, 256, 0, st>>>(partial, kSpecNodes * ntile, 1.f / (float)plane, blk.se0, blk.se2, blk.co,
                                     blk.se_hidden, gate);
  }
  AASIST_CUDA(cudaGetLastError());
  {
    dim3 grid((W / 3 + 127) / 128, kSpecNodes, B * blk.co);
    LaunchSpan span(h, "res2_gate_res_pool", st);
    res2_finish_kernel<<<grid, 128, 0, st>>>(y, gate, in, blk.downsample ? blk.wd : nullptr, blk.bd, out, blk.ci,
                                            blk.co, W);
  }
  AASIST_CUDA(cudaGetLastError());
  return 0;
}

template <int CO_T>
static int launch_block33_t(aasist_handle* h, const ConvBlock33F32& blk, const float* in, int B, int W, float* mid,
                            float* out, cudaStream_t st) {
  int rc = launch_conv33<CO_T, M33_SELU>(h, "conv1_3x3_f32", in, nullptr, blk.w1, nullptr, blk.b1, mid, nullptr,
                                         blk.ci, 0, blk.co, W, B, st);
  if (rc) return rc;
  if (blk.downsample)
    return launch_conv33<CO_T, M33_RES_DS>(h, "conv2_3x3_res_pool_f32", mid, in, blk.w2, blk.wd, blk.b2, out, nullptr,
                                           blk.co, blk.ci, blk.co, W, B, st);
  return launch_conv33<CO_T, M33_RES_ID>(h, "conv2_3x3_res_pool_f32", mid, in, blk.w2, nullptr, blk.b2, out, nullptr,
                                         blk.co, blk.co, blk.co, W, B, st);
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
