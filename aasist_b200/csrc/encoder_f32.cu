// Residual_block encoder, fp32 CUDA-core path (reference models/RawNetGatSpoofST.py:225-278).
//   conv1 k(2,3) pad(1,1) + bn2 (folded) + SELU            -> mid (B,Co,24,W)
//   conv2 k(2,3) pad(0,1) + identity | conv_downsample k(1,3) pad(0,1) + MaxPool2d((1,3))
//                                                          -> out (B,Co,23,W/3)
// `bn1`+SELU on the block input is dead code in the reference (:260-265) and is not computed.
// Direct convolution, NCHW fp32: CTA = one (utterance, row) x 96 columns x all output channels;
// thread = 3 adjacent columns (one pool window) x CO_T channels, input/weights staged through
// shared memory in chunks of 8 input channels.
#include "common.cuh"

namespace aasist {

constexpr int kTW = 96;    // output columns per CTA (multiple of 3)
constexpr int kCK = 8;     // input channels per shared-memory chunk
constexpr int kInLd = kTW + 2;

enum { MODE_CONV1 = 0, MODE_CONV2_ID = 1, MODE_CONV2_DS = 2 };

template <int CO_T>
__device__ __forceinline__ void load_w(float (&w)[CO_T], const float* p) {
#pragma unroll
  for (int q = 0; q < CO_T; q += 4) {
    float4 t = *reinterpret_cast<const float4*>(p + q);
    w[q] = t.x; w[q + 1] = t.y; w[q + 2] = t.z; w[q + 3] = t.w;
  }
}

// in:   (B, Ci, Hin, W)    main input  (MODE_CONV1: x, Hin=23; else: mid, Hin=24)
// side: (B, Cs, 23, W)     block input x (identity or conv_downsample source); null for CONV1
// wmain [Ci][2][3][Cop], wside [Cs][3][Cop], bias [Cop]; Cop = 8*CO_T >= Co
template <int CO_T, int MODE>
__global__ void __launch_bounds__(256)
conv23_f32_kernel(const float* __restrict__ in, const float* __restrict__ side,
                  const float* __restrict__ wmain, const float* __restrict__ wside,
                  const float* __restrict__ bias, float* __restrict__ out, int Ci, int Cs, int Co,
                  int W) {
  constexpr int Cop = 8 * CO_T;
  __shared__ __align__(16) float s_in[kCK * 2 * kInLd];
  __shared__ __align__(16) float s_w[kCK * 6 * Cop];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int w0 = blockIdx.x * kTW;
  const int row = blockIdx.y;              // output row (CONV1: 0..23, CONV2: 0..22)
  const int b = blockIdx.z;
  const int Hin = (MODE == MODE_CONV1) ? 23 : 24;
  const int row_off = (MODE == MODE_CONV1) ? -1 : 0;   // input row of tap dh=0

  float acc[3][CO_T];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int c = 0; c < CO_T; ++c) acc[j][c] = 0.f;

  const float* inb = in + (size_t)b * Ci * Hin * W;
  for (int c0 = 0; c0 < Ci; c0 += kCK) {
    const int nc = min(kCK, Ci - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < nc * 2 * kInLd; i += 256) {
      int col = i % kInLd, r = (i / kInLd) & 1, c = i / (2 * kInLd);
      int gr = row + row_off + r, gw = w0 - 1 + col;
      float v = 0.f;
      if (gr >= 0 && gr < Hin && gw >= 0 && gw < W) v = inb[((size_t)(c0 + c) * Hin + gr) * W + gw];
      s_in[i] = v;
    }
    for (int i = threadIdx.x; i < nc * 6 * Cop / 4; i += 256)
      reinterpret_cast<float4*>(s_w)[i] =
          reinterpret_cast<const float4*>(wmain + (size_t)c0 * 6 * Cop)[i];
    __syncthreads();
    for (int c = 0; c < nc; ++c) {
#pragma unroll
      for (int dh = 0; dh < 2; ++dh) {
        const float* ip = s_in + (c * 2 + dh) * kInLd + 3 * tx;
        float v[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) v[q] = ip[q];
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          float w[CO_T];
          load_w<CO_T>(w, s_w + ((c * 2 + dh) * 3 + dw) * Cop + ty * CO_T);
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int q = 0; q < CO_T; ++q) acc[j][q] = fmaf(v[j + dw], w[q], acc[j][q]);
        }
      }
    }
  }
  if (MODE == MODE_CONV2_DS) {
    const float* sb = side + (size_t)b * Cs * 23 * W;
    for (int c0 = 0; c0 < Cs; c0 += kCK) {
      const int nc = min(kCK, Cs - c0);
      __syncthreads();
      for (int i = threadIdx.x; i < nc * kInLd; i += 256) {
        int col = i % kInLd, c = i / kInLd;
        int gw = w0 - 1 + col;
        float v = 0.f;
        if (gw >= 0 && gw < W) v = sb[((size_t)(c0 + c) * 23 + row) * W + gw];
        s_in[i] = v;
      }
      for (int i = threadIdx.x; i < nc * 3 * Cop / 4; i += 256)
        reinterpret_cast<float4*>(s_w)[i] =
            reinterpret_cast<const float4*>(wside + (size_t)c0 * 3 * Cop)[i];
      __syncthreads();
      for (int c = 0; c < nc; ++c) {
        const float* ip = s_in + c * kInLd + 3 * tx;
        float v[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) v[q] = ip[q];
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          float w[CO_T];
          load_w<CO_T>(w, s_w + (c * 3 + dw) * Cop + ty * CO_T);
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int q = 0; q < CO_T; ++q) acc[j][q] = fmaf(v[j + dw], w[q], acc[j][q]);
        }
      }
    }
  }

  const int wbase = w0 + 3 * tx;
  if (MODE == MODE_CONV1) {
    float* ob = out + (size_t)b * Co * 24 * W;
#pragma unroll
    for (int q = 0; q < CO_T; ++q) {
      int co = ty * CO_T + q;
      if (co >= Co) continue;
      float bq = bias[co];
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (wbase + j < W) ob[((size_t)co * 24 + row) * W + wbase + j] = selu(acc[j][q] + bq);
    }
  } else {
    const int Wo = W / 3;
    const int po = blockIdx.x * (kTW / 3) + tx;
    if (po >= Wo) return;
    float* ob = out + (size_t)b * Co * 23 * Wo;
    const float* sb = side + (size_t)b * Cs * 23 * W;
#pragma unroll
    for (int q = 0; q < CO_T; ++q) {
      int co = ty * CO_T + q;
      if (co >= Co) continue;
      float bq = bias[co];
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float v = acc[j][q] + bq;
        if (MODE == MODE_CONV2_ID) v += sb[((size_t)co * 23 + row) * W + wbase + j];
        m = fmaxf(m, v);
      }
      ob[((size_t)co * 23 + row) * Wo + po] = m;
    }
  }
}

template <int CO_T>
static int launch_block_t(aasist_handle* h, const ConvBlockF32& blk, const float* in, int B, int W,
                          float* mid, float* out, cudaStream_t st) {
  dim3 g1((W + kTW - 1) / kTW, 24, B);
  {
    LaunchSpan span(h, blk.ci == 1 ? "conv1_f32[1->C]" : "conv1_f32", st);
    conv23_f32_kernel<CO_T, MODE_CONV1><<<g1, 256, 0, st>>>(in, nullptr, blk.w1, nullptr, blk.b1, mid,
                                                           blk.ci, 0, blk.co, W);
  }
  AASIST_CUDA(cudaGetLastError());
  int Wo = W / 3;
  dim3 g2((3 * Wo + kTW - 1) / kTW, 23, B);
  {
    LaunchSpan span(h, "conv2_res_pool_f32", st);
    if (blk.downsample)
      conv23_f32_kernel<CO_T, MODE_CONV2_DS><<<g2, 256, 0, st>>>(mid, in, blk.w2, blk.wd, blk.b2, out,
                                                                blk.co, blk.ci, blk.co, W);
    else
      conv23_f32_kernel<CO_T, MODE_CONV2_ID><<<g2, 256, 0, st>>>(mid, in, blk.w2, nullptr, blk.b2, out,
                                                                blk.co, blk.co, blk.co, W);
  }
  AASIST_CUDA(cudaGetLastError());
  return 0;
}

// in (B,ci,23,W) -> mid (B,co,24,W) scratch -> out (B,co,23,W/3)
int launch_block_f32(aasist_handle* h, const ConvBlockF32& blk, const float* in, int B, int W,
                     float* mid, float* out, cudaStream_t st) {
  if (W < 3) {
    set_error("encoder block input width %d < 3 (MaxPool2d((1,3)) would be empty)", W);
    return AASIST_E_INVALID;
  }
  if (blk.co <= 32) return launch_block_t<4>(h, blk, in, B, W, mid, out, st);
  if (blk.co <= 64) return launch_block_t<8>(h, blk, in, B, W, mid, out, st);
  set_error("encoder blocks with more than 64 output channels are not supported (got %d)", blk.co);
  return AASIST_E_INVALID;
}

}  // namespace aasist
