// Sinc front end, fp32 CUDA-core path.
//   a1  filter bank built on device            (reference models/AASIST.py:460-482)
//   a2  valid cross-correlation, stride 1      (models/AASIST.py:497-503)
//   a3  |.| -> 3x3 max-pool -> first_bn -> SELU (models/AASIST.py:829-831)
// fused into one kernel: the (B,70,L-128) conv output (18 MB/utterance) never exists.
#include "common.cuh"

namespace aasist {

// ---------------------------------------------------------------------------------------
// filter bank: one thread per (filter, tap); same dtype chain as the reference expression
// (fp32 sinc argument, fp64 scale/difference, fp32 Hamming product).
// ---------------------------------------------------------------------------------------
__global__ void sinc_filterbank_kernel(float* __restrict__ bank, int n_filters, int taps,
                                       int sample_rate) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_filters * taps) return;
  int i = idx / taps, j = idx % taps;
  const double PI = 3.141592653589793;
  double fnyq = (double)(sample_rate / 2);
  double mel_max = 2595.0 * log10(1.0 + fnyq / 700.0);           // to_mel(f[-1]); to_mel(0) == 0
  double step = mel_max / (double)n_filters;                      // np.linspace(0, mel_max, F+1)
  double mel_lo = (double)i * step;
  double mel_hi = (i + 1 == n_filters) ? mel_max : (double)(i + 1) * step;
  double fmin = 700.0 * (pow(10.0, mel_lo / 2595.0) - 1.0);       // to_hz
  double fmax = 700.0 * (pow(10.0, mel_hi / 2595.0) - 1.0);
  float n = (float)(j - (taps - 1) / 2);                          // hsupp, float32
  float sr = (float)sample_rate;
  float pif = (float)PI;
  float ideal_parts[2];
  double scale[2] = {2.0 * fmax / (double)sample_rate, 2.0 * fmin / (double)sample_rate};
  float two_f[2] = {(float)(2.0 * fmax), (float)(2.0 * fmin)};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float arg = __fdiv_rn(__fmul_rn(n, two_f[q]), sr);            // 2*f*n/sr in fp32
    float y = __fmul_rn(pif, arg == 0.f ? 1.0e-20f : arg);        // np.sinc: pi*where(x==0,1e-20,x)
    float s = (float)sin((double)y);                              // correctly rounded fp32 sin(y)
    ideal_parts[q] = __fdiv_rn(s, y);
  }
  double ideal = scale[0] * (double)ideal_parts[0] - scale[1] * (double)ideal_parts[1];
  float window = (float)(0.54 - 0.46 * cos(2.0 * PI * (double)j / (double)(taps - 1)));
  bank[idx] = __fmul_rn(window, (float)ideal);
}

__global__ void transpose_bank_kernel(const float* __restrict__ bank, float* __restrict__ bank_t, int n_filters,
                                      int taps) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_filters * taps) return;
  int f = idx / taps, k = idx % taps;
  bank_t[k * n_filters + f] = bank[idx];
}

int build_filterbank(aasist_handle* h) {
  int n = h->cfg.n_filters * h->taps;
  if (!h->bank) AASIST_CUDA(cudaMalloc(&h->bank, sizeof(float) * n));
  sinc_filterbank_kernel<<<(n + 127) / 128, 128>>>(h->bank, h->cfg.n_filters, h->taps,
                                                   h->cfg.sample_rate);
  h->launches++;
  AASIST_CUDA(cudaGetLastError());
  if (h->cfg.kind == AASIST_KIND_ROBUST) {     // the strided front end reads the bank tap-major
    if (!h->bank_t) AASIST_CUDA(cudaMalloc(&h->bank_t, sizeof(float) * n));
    transpose_bank_kernel<<<(n + 127) / 128, 128>>>(h->bank, h->bank_t, h->cfg.n_filters, h->taps);
    h->launches++;
    AASIST_CUDA(cudaGetLastError());
  }
  AASIST_CUDA(cudaDeviceSynchronize());
  return 0;
}

// ---------------------------------------------------------------------------------------
// fused sinc conv + abs + 3x3 max-pool + BN(scalar affine) + SELU
// CTA: 64 pooled time steps x all 23 pooled bands of one utterance.
// thread (fi, tg): bands 3fi..3fi+2, pooled steps tg + 16m (m<4) -> 36 accumulators.
// ---------------------------------------------------------------------------------------
constexpr int kFrontTile = 64;   // pooled outputs per CTA
constexpr int kFrontTG = 16;     // time groups
constexpr int kFrontM = kFrontTile / kFrontTG;

__global__ void __launch_bounds__(kSpecNodes* kFrontTG)
sinc_frontend_f32_kernel(const float* __restrict__ x, const float* __restrict__ bank,
                         float* __restrict__ out, int L, int taps, int Wp, int n_bands,
                         float bn_scale, float bn_shift, int mask_start, int mask_count) {
  extern __shared__ float smem[];
  float* fs = smem;                               // [3*n_bands][taps]
  float* xs = smem + 3 * n_bands * taps;          // [3*kFrontTile + taps - 1 (+pad)]
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kFrontTile;         // first pooled step of this tile
  const int t0 = 3 * p0;                          // first conv output time
  const int nx = 3 * kFrontTile + taps - 1;
  for (int i = threadIdx.x; i < 3 * n_bands * taps; i += blockDim.x) {
    const int f = i / taps;                       // Freq_aug (AASIST.py:486-490): masked filters are all-zero
    fs[i] = (f >= mask_start && f < mask_start + mask_count) ? 0.f : bank[i];
  }
  const float* xb = x + (size_t)b * L;
  for (int i = threadIdx.x; i < nx + 3; i += blockDim.x) {
    int g = t0 + i;
    xs[i] = g < L ? xb[g] : 0.f;
  }
  __syncthreads();
  const int fi = threadIdx.x / kFrontTG, tg = threadIdx.x % kFrontTG;
  if (fi >= n_bands) return;
  float acc[3][kFrontM][3];
#pragma unroll
  for (int f = 0; f < 3; ++f)
#pragma unroll
    for (int m = 0; m < kFrontM; ++m)
#pragma unroll
      for (int d = 0; d < 3; ++d) acc[f][m][d] = 0.f;
  const float* f0 = fs + (3 * fi) * taps;
  float win[kFrontM][3];
#pragma unroll
  for (int m = 0; m < kFrontM; ++m) {
    win[m][0] = xs[3 * (tg + kFrontTG * m) + 0];
    win[m][1] = xs[3 * (tg + kFrontTG * m) + 1];
    win[m][2] = xs[3 * (tg + kFrontTG * m) + 2];
  }
  for (int k = 0; k < taps; ++k) {
    float c0 = f0[k], c1 = f0[taps + k], c2 = f0[2 * taps + k];
#pragma unroll
    for (int m = 0; m < kFrontM; ++m) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        acc[0][m][d] = fmaf(c0, win[m][d], acc[0][m][d]);
        acc[1][m][d] = fmaf(c1, win[m][d], acc[1][m][d]);
        acc[2][m][d] = fmaf(c2, win[m][d], acc[2][m][d]);
      }
      win[m][0] = win[m][1];
      win[m][1] = win[m][2];
      win[m][2] = xs[3 * (tg + kFrontTG * m) + k + 3];
    }
  }
#pragma unroll
  for (int m = 0; m < kFrontM; ++m) {
    int p = p0 + tg + kFrontTG * m;
    if (p >= Wp) continue;
    float v = 0.f;
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
      for (int d = 0; d < 3; ++d) v = fmaxf(v, fabsf(acc[f][m][d]));
    out[((size_t)b * n_bands + fi) * Wp + p] = selu(fmaf(v, bn_scale, bn_shift));
  }
}

int launch_frontend_f32(aasist_handle* h, const float* x, int B, int L, float* out, int mask_start,
                        int mask_count, cudaStream_t st) {
  const int taps = h->taps;
  const int Wp = (L - taps + 1) / 3;
  const int n_bands = h->cfg.n_filters / 3;
  if (Wp < 1) {
    set_error("input length %d too short for a %d-tap filter bank", L, taps);
    return AASIST_E_INVALID;
  }
  size_t smem = sizeof(float) * (3 * n_bands * taps + 3 * kFrontTile + taps + 8);
  if (smem > 227 * 1024) {
    // 3 * 23 * taps fp32 filter rows are resident per CTA: ~840 taps is the most that fits on an SM
    set_error("first_conv=%d: a %d-tap filter bank needs %zu bytes of shared memory per CTA in the fp32 front end "
              "(max 232448)", h->cfg.first_conv, taps, smem);
    return AASIST_E_INVALID;
  }
  // the attribute is per device (a second GPU in the same process needs it too): set it on every launch
  AASIST_CUDA(cudaFuncSetAttribute(sinc_frontend_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  dim3 grid((Wp + kFrontTile - 1) / kFrontTile, B);
  {
    LaunchSpan span(h, "sinc_frontend_f32", st);
    sinc_frontend_f32_kernel<<<grid, n_bands * kFrontTG, smem, st>>>(
        x, h->bank, out, L, taps, Wp, n_bands, h->bn0_scale, h->bn0_shift, mask_start, mask_count);
  }
  AASIST_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
// AASIST-Robust front end (models/AASIST_Robust.py:96-102, 217-221): `first_conv` sinc filters of 1025 taps
// at stride 256, then the same |.| -> 3x3 max-pool -> first_bn -> SELU.  0.07 MMAC per output frame and
// filter: one CTA per (utterance, pooled step) = 3 frames x all filters, signal span (2*256 + taps samples)
// in shared memory, bank read tap-major (coalesced over filters).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sinc_frontend_strided_f32_kernel(const float* __restrict__ x, const float* __restrict__ bank_t,
                                 float* __restrict__ out, int L, int taps, int stride, int Wp, int n_filters,
                                 int n_bands, float bn_scale, float bn_shift, int mask_start, int mask_count) {
  extern __shared__ float smem[];
  float* xs = smem;                                // [2*stride + taps]
  float* ys = smem + 2 * stride + taps;            // [3 frames][3*n_bands] |conv|
  const int b = blockIdx.y, p = blockIdx.x;
  const int t0 = 3 * p * stride;                   // first sample of frame 3p
  const int span = 2 * stride + taps;
  const float* xb = x + (size_t)b * L;
  for (int i = threadIdx.x; i < span; i += blockDim.x) xs[i] = (t0 + i < L) ? xb[t0 + i] : 0.f;
  __syncthreads();
  const int nf = 3 * n_bands;                      // filters that survive the floor-mode pool
  for (int o = threadIdx.x; o < 3 * nf; o += blockDim.x) {
    const int fr = o / nf, f = o % nf;
    const float* xp = xs + fr * stride;
    float acc = 0.f;
    for (int k = 0; k < taps; ++k) acc = fmaf(xp[k], __ldg(bank_t + (size_t)k * n_filters + f), acc);
    if (f >= mask_start && f < mask_start + mask_count) acc = 0.f;
    ys[fr * nf + f] = fabsf(acc);
  }
  __syncthreads();
  for (int band = threadIdx.x; band < n_bands; band += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int fr = 0; fr < 3; ++fr)
#pragma unroll
      for (int d = 0; d < 3; ++d) v = fmaxf(v, ys[fr * nf + 3 * band + d]);
    out[((size_t)b * n_bands + band) * Wp + p] = selu(fmaf(v, bn_scale, bn_shift));
  }
}

int launch_frontend_strided_f32(aasist_handle* h, const float* x, int B, int L, float* out, int mask_start,
                                int mask_count, cudaStream_t st) {
  const int taps = h->taps, stride = h->stride;
  if (L < taps) {
    set_error("input length %d too short for a %d-tap filter bank", L, taps);
    return AASIST_E_INVALID;
  }
  const int frames = (L - taps) / stride + 1;
  const int Wp = frames / 3;
  const int n_bands = h->cfg.n_filters / 3;
  if (Wp < 1) {
    set_error("input length %d gives %d frames: too short for the 3x3 max-pool", L, frames);
    return AASIST_E_INVALID;
  }
  const size_t smem = sizeof(float) * (2 * stride + taps + 9 * n_bands + 8);
  AASIST_CUDA(cudaFuncSetAttribute(sinc_frontend_strided_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  dim3 grid(Wp, B);
  {
    LaunchSpan span(h, "sinc_frontend_strided_f32", st);
    sinc_frontend_strided_f32_kernel<<<grid, 256, smem, st>>>(x, h->bank_t, out, L, taps, stride, Wp, h->cfg.n_filters,
                                                              n_bands, h->bn0_scale, h->bn0_shift, mask_start,
                                                              mask_count);
  }
  AASIST_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace aasist
