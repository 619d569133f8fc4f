// Sinc front end on the tensor cores (reference models/AASIST.py:497-503, 829-831):
// valid 129-tap cross-correlation of the waveform with the 70-filter bank, |.|, 3x3 max-pool
// (band x time), first_bn, SELU -- one kernel, the (B,70,L-128) conv output never exists.
//
// GEMM formulation.  A CTA tile is 128 POOLED time steps ti (384 conv outputs).  The pool phase
// s = t mod 3 is folded into the filter operand instead of the signal operand:
//     D[ti][n] = sum_k' A[ti][k'] * H[n][k'],     A[ti][k'] = x[3*(ti0+ti) + k'],  k' < 131 (pad 144)
//     n = 9*fi + 3*df + s  <->  H[n][k'] = h[3*fi+df][k' - s]   (zero outside 0..128)
// so ONE Hankel-like A tile (rows overlap with stride 3 samples: TMA cannot express it, producer
// warps build it in shared memory from a staged fp16 copy of the waveform segment) is multiplied by
// a resident 208 x 144 filter operand, and the 9 columns of a pooled band fi -- 3 filters x 3
// phases -- are adjacent, so |.|/max-pool is register-local in the epilogue.  N = 23*9 = 207 -> 208.
// Operands are fp16 hi/lo pairs (x and h pre-scaled by 2^10, undone exactly in the epilogue),
// 3 products per K chunk, fp32 accumulation in TMEM (same scheme as encoder_tc.cu).
//
// Both operands use the no-swizzle K-major canonical layout (8-row x 16-byte core matrices):
//   addr(row, kbyte) = chunk*CHUNK + (row/8)*256 + (kbyte/16)*128 + (row%8)*16 + kbyte%16
// i.e. SBO = 256 B (next 8-row group), LBO = 128 B (second 16-byte K half of a UMMA_K=16 chunk).
// The A operand lives in TENSOR MEMORY (tcgen05.mma with A in TMEM, ptx.cuh umma_f16_ts): a producer thread owns
// a tile row = a TMEM lane and writes the 16 (hi,lo) words of a K chunk straight from the staged segment with one
// tcgen05.st -- the tile is never written to or read from shared memory (it was 73 KB of writes + 108 KB of
// operand reads per tile next to the 175 KB of filter-operand reads).  The chunks form a 6-slot ring in the 96
// TMEM columns the two 208-column accumulators leave free; the MMA warp releases a slot with tcgen05.commit.
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "ptx.cuh"
#include "tc.cuh"

namespace aasist {

using namespace ptx;

constexpr int kFtTile = 128;                  // pooled steps per tile (UMMA M)
constexpr int kFtN = 208;                     // 23 bands x 9 (3 filters x 3 phases), padded to /16
constexpr int kFtKC = 9;                      // K chunks of 16: 131 taps+phases -> 144
constexpr int kFtASlots = 6;                  // A-operand ring: K chunks of 16 TMEM columns [hi 8 | lo 8]
constexpr int kFtBChunk = kFtN * 32;          // bytes of one B K-chunk (hi or lo)
constexpr int kFtSeg = 3 * kFtTile + 16 * kFtKC;   // staged samples per tile (527 used)
static_assert(kFtKC % 3 == 0, "K chunks split over three producer groups");
constexpr int kFtProdWarps = 12;              // 384 producer threads: three per A row (K chunks kc % 3); the producers bound this kernel
constexpr int kFtThreads = 64 + 32 * 8 + 32 * kFtProdWarps;
constexpr int kFtBufCols = 256;               // TMEM column stride between the two accumulators
constexpr float kFtScale = 1024.f;            // 2^10 on both operands

struct FrontTcParams {
  const float* x;          // (B, L)
  float* out;              // (B, 23, Wp)
  int collector;           // A-operand collector reuse between the two a_hi products
  int* range_flag;         // host-mapped: set to 1 when a sample leaves the fp16 operand range (|x| * 2^10 > 65504)
  const uint8_t* bimg;     // filter operand image: [hi|lo][kc][208 rows x 32 B], no-swizzle canonical
  int B, L, Wp, n_tiles_per_utt;
  float bn_scale, bn_shift;
  long long* stats;        // optional: MMA-warp wait cycles per CTA [total, afull(producers), tempty(epilogue)]
};

// TMEM column of A-ring slot i: the columns the two accumulators ([0,208) and [256,464)) leave free
__device__ __forceinline__ uint32_t ft_a_col(int slot) {
  return (uint32_t)(slot < 3 ? kFtN + 16 * slot : kFtBufCols + kFtN + 16 * (slot - 3));
}
// no-swizzle K-major descriptor: LBO = 128 B, SBO = 256 B, version 1, layout 0
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) |
         ((uint64_t)1 << 46);
}

__device__ __forceinline__ float ft_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ft_selu(float v) {
  const float e = ft_ex2(v * 1.4426950408889634f);
  const float n = fminf(fmaf(e, kSeluScale * kSeluAlpha, -(kSeluScale * kSeluAlpha)), 0.f);
  return fmaf(fmaxf(v, 0.f), kSeluScale, n);
}

// epilogue of one half of the bands: HALF 0 -> fi 0..11 (columns 0..107), HALF 1 -> fi 12..22
// (columns 108..206).  Loads are issued at 16-aligned column offsets.
template <int HALF>
__device__ __forceinline__ void front_epilogue(uint32_t t_row, const FrontTcParams& p, int b, int ti,
                                               uint64_t* tempty_bar, int lane) {
  constexpr int COL0 = HALF == 0 ? 0 : 96;           // first loaded column (16-aligned)
  constexpr int FI0 = HALF == 0 ? 0 : 12, NFI = HALF == 0 ? 12 : 11;
  // two batches of accumulator columns (4 + 3 chunks of 16) keep the register count low:
  // batch 0 covers bands whose nine columns end before local column 64, batch 1 the rest
  float* o = p.out + ((size_t)b * kSpecNodes) * p.Wp + ti;
  const bool live = ti < p.Wp;
  constexpr int SPLIT = 7;                                       // bands 0..6 of this half live in columns < 63
  {
    uint32_t acc[4][16];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld16_async(t_row + (uint32_t)(COL0 + 16 * c), acc[c]);
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld_wait16(acc[c]);
#pragma unroll
    for (int f = 0; f < NFI; ++f) {
      const int c0 = 9 * (FI0 + f) - COL0;
      if (c0 + 8 < 64) {                                         // compile-time after unrolling
        float m = 0.f;
#pragma unroll
        for (int q = 0; q < 9; ++q) m = fmaxf(m, fabsf(__uint_as_float(acc[(c0 + q) >> 4][(c0 + q) & 15])));
        m *= 1.f / (kFtScale * kFtScale);                        // undo the 2^10 operand scales (exact)
        if (live) o[(size_t)(FI0 + f) * p.Wp] = ft_selu(fmaf(m, p.bn_scale, p.bn_shift));
      }
    }
  }
  {
    uint32_t acc[4][16];                                         // local columns 48 .. 111
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld16_async(t_row + (uint32_t)(COL0 + 48 + 16 * c), acc[c]);
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld_wait16(acc[c]);
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty_bar);
#pragma unroll
    for (int f = 0; f < NFI; ++f) {
      const int c0 = 9 * (FI0 + f) - COL0;
      if (c0 + 8 >= 64) {
        float m = 0.f;
#pragma unroll
        for (int q = 0; q < 9; ++q) m = fmaxf(m, fabsf(__uint_as_float(acc[(c0 + q - 48) >> 4][(c0 + q - 48) & 15])));
        m *= 1.f / (kFtScale * kFtScale);
        if (live) o[(size_t)(FI0 + f) * p.Wp] = ft_selu(fmaf(m, p.bn_scale, p.bn_shift));
      }
    }
  }
  (void)SPLIT;
}

__global__ void __launch_bounds__(kFtThreads, 1)
sinc_frontend_tc_kernel(const FrontTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_b = smem;                                         // [2][9][kFtBChunk]
  // staged waveform segment, double buffered (tile t+1 is staged while tile t is being built): 4 copies each --
  // [0] hi, xh0[i] = hi(x[g0+i]); [1] hi shifted by one sample, xh1[i] = xh0[i+1]; [2], [3] the same for lo
  __half* s_x = reinterpret_cast<__half*>(s_b + 2 * kFtKC * kFtBChunk);
  constexpr int XLEN = kFtSeg + 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_x + 2 * 4 * XLEN);
  uint64_t* afull = bars;            // [kFtASlots] (kFtKC entries reserved)
  uint64_t* aempty = bars + kFtKC;   // [kFtASlots]
  uint64_t* tfull = bars + 2 * kFtKC;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.B * p.n_tiles_per_utt;

  for (int i = threadIdx.x; i < 2 * kFtKC * kFtBChunk / 16; i += kFtThreads)
    reinterpret_cast<uint4*>(s_b)[i] = __ldg(reinterpret_cast<const uint4*>(p.bimg) + i);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kFtASlots; ++i) {
      mbar_init(&afull[i], 4);                     // the four warps (one per TMEM lane quadrant) that own this chunk
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 1) {
    // =============================== MMA issuer ==================================
    const bool leader = elect_one();   // the whole schedule runs in this one lane
    if (leader) {
    const uint32_t b_base = smem_u32(s_b);
    constexpr uint32_t IDESC = umma_idesc_f16(128, kFtN);
    int tcount = 0;
    int g = 0;                           // K chunks consumed so far: chunk g lives in ring slot g % kFtASlots
    long long w_te = 0, w_af = 0;
    const long long t_begin = AASIST_CLOCK();
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      AASIST_TIMED_WAIT(&tempty[buf], ((tcount >> 1) & 1) ^ 1, w_te);
      tc_fence_after_sync();
      const uint32_t d = tmem_base + (uint32_t)(buf * kFtBufCols);
      for (int kc = 0; kc < kFtKC; ++kc, ++g) {
        const int slot = g % kFtASlots;
        AASIST_TIMED_WAIT(&afull[slot], (g / kFtASlots) & 1, w_af);
        tc_fence_after_sync();
        {
          const uint32_t a_hi = tmem_base + ft_a_col(slot), a_lo = a_hi + 8;
          const uint64_t b_hi = umma_desc_noswz(b_base + (uint32_t)(kc * kFtBChunk));
          const uint64_t b_lo = umma_desc_noswz(b_base + (uint32_t)((kFtKC + kc) * kFtBChunk));
          umma_f16_ts(d, a_hi, b_hi, IDESC, kc > 0 ? 1u : 0u);
          umma_f16_ts(d, a_lo, b_hi, IDESC, 1);
          umma_f16_ts(d, a_hi, b_lo, IDESC, 1);
          umma_commit(&aempty[slot]);
        }
      }
      umma_commit(&tfull[buf]);
    }
    if (p.stats) {
      long long* stt = p.stats + (size_t)blockIdx.x * 4;
      stt[0] = AASIST_CLOCK() - t_begin; stt[1] = w_af; stt[2] = w_te;
    }
    }
  } else if (warp >= 2 && warp < 10) {
    // =============================== epilogue ====================================
    const int quad = warp & 3, half = (warp - 2) >> 2;
    int tcount = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
      const int b = t / p.n_tiles_per_utt, tile = t % p.n_tiles_per_utt;
      const int buf = tcount & 1;
      const int ti = tile * kFtTile + quad * 32 + lane;
      mbar_wait(&tfull[buf], (tcount >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * kFtBufCols);
      if (half == 0) front_epilogue<0>(t_row, p, b, ti, &tempty[buf], lane);
      else front_epilogue<1>(t_row, p, b, ti, &tempty[buf], lane);
    }
  } else if (warp >= 10) {
    // ============ A-operand producers: stage the waveform segment, build the Hankel-like tile ============
    const int ptid = threadIdx.x - 320;          // 0..255 (staging of the segment)
    // tile row = TMEM lane: a warp reaches the 32 lanes of quadrant warp % 4; K-chunk parity = warp group
    const int prow = (warp & 3) * 32 + lane, ppart = (warp - 10) >> 2;
    int tcount = 0;
    constexpr int PER = (kFtSeg + 1 + 32 * kFtProdWarps - 1) / (32 * kFtProdWarps);   // samples per thread (5)
    float pre[PER];
    // fetch the waveform segment of tile t into registers (issued one tile ahead: the global-load
    // latency overlaps the construction of the previous tile)
    auto fetch = [&](int t) {
      const int b = t / p.n_tiles_per_utt, tile = t % p.n_tiles_per_utt;
      const size_t g0 = (size_t)3 * tile * kFtTile;
      const float* xb = p.x + (size_t)b * p.L;
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int i = ptid + 32 * kFtProdWarps * q;
        const size_t g = g0 + i;
        pre[q] = (i < kFtSeg + 1 && g < (size_t)p.L) ? __ldg(xb + g) : 0.f;
      }
    };
    // registers -> fp16 pairs in segment buffer `buf`
    auto stage = [&](int buf) {
      __half* xh0 = s_x + buf * 4 * XLEN;
      __half* xh1 = xh0 + XLEN;
      __half* xl0 = xh0 + 2 * XLEN;
      __half* xl1 = xh0 + 3 * XLEN;
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int i = ptid + 32 * kFtProdWarps * q;
        if (i < kFtSeg + 1) {
          const float vc = fminf(fmaxf(pre[q] * kFtScale, -65504.f), 65504.f);
          // un-normalised input (e.g. int16-scale samples): the operand saturates and the logits are wrong without
          // any error -- raise the host-visible flag (a store over PCIe on this rare path only)
          if (fabsf(pre[q]) * kFtScale > 65504.f) *reinterpret_cast<volatile int*>(p.range_flag) = 1;
          const __half h = __float2half_rn(vc);
          const __half l = __float2half_rn(vc - __half2float(h));
          xh0[i] = h;
          xl0[i] = l;
          if (i > 0) { xh1[i - 1] = h; xl1[i - 1] = l; }
        }
      }
    };
    if ((int)blockIdx.x < n_tiles) {
      fetch(blockIdx.x);
      stage(0);
      if ((int)(blockIdx.x + gridDim.x) < n_tiles) fetch(blockIdx.x + gridDim.x);
    }
    asm volatile("bar.sync 2, %0;" ::"n"(32 * kFtProdWarps) : "memory");
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
      // tile t+1 into the other buffer (its last readers, the builders of tile t-1, are behind the barrier that
      // ended the previous iteration), then the samples of tile t+2 into registers
      if (t + (int)gridDim.x < n_tiles) {
        stage((tcount + 1) & 1);
        if (t + 2 * (int)gridDim.x < n_tiles) fetch(t + 2 * gridDim.x);
      }
      const __half* xh0 = s_x + (tcount & 1) * 4 * XLEN;
      const __half* xh1 = xh0 + XLEN;
      const __half* xl0 = xh0 + 2 * XLEN;
      const __half* xl1 = xh0 + 3 * XLEN;
      // row ptid, K halves [8q, 8q+8) = samples 3*ptid + 8q ...; pick the copy that makes the start even
      const int start0 = 3 * prow;
      const bool odd = start0 & 1;
      const uint32_t* srch = reinterpret_cast<const uint32_t*>(odd ? xh1 : xh0) + ((start0 - (odd ? 1 : 0)) >> 1);
      const uint32_t* srcl = reinterpret_cast<const uint32_t*>(odd ? xl1 : xl0) + ((start0 - (odd ? 1 : 0)) >> 1);
      const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      for (int kc = ppart; kc < kFtKC; kc += kFtProdWarps / 4) {
        const int g = tcount * kFtKC + kc, slot = g % kFtASlots;
        mbar_wait(&aempty[slot], ((g / kFtASlots) & 1) ^ 1);
        tc_fence_after_sync();
        uint32_t wv[16];                                 // [hi: K 0..15 | lo: K 0..15] of this row's chunk
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          wv[i] = srch[8 * kc + i];
          wv[8 + i] = srcl[8 * kc + i];
        }
        tmem_st16(t_lane + ft_a_col(slot), wv);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&afull[slot]);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(32 * kFtProdWarps) : "memory");   // tile t built, tile t+1 staged
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// filter operand image from the device-built bank (copied back once at finalize)
int tc_front_finalize(aasist_handle* h, uint8_t** bimg_dev) {
  const int F = h->cfg.n_filters, K = h->taps;
  if (K > 16 * kFtKC - 2 || F / 3 != kSpecNodes) {
    set_error("f16x3 sinc front end supports up to %d taps and 69..71 filters", 16 * kFtKC - 2);
    return AASIST_E_INVALID;
  }
  std::vector<float> bank((size_t)F * K);
  AASIST_CUDA(cudaMemcpy(bank.data(), h->bank, sizeof(float) * bank.size(), cudaMemcpyDeviceToHost));
  std::vector<uint8_t> img((size_t)2 * kFtKC * kFtBChunk, 0);
  for (int n = 0; n < 9 * kSpecNodes; ++n) {
    const int fi = n / 9, df = (n % 9) / 3, s = n % 3;
    const int f = 3 * fi + df;
    for (int k = 0; k < K; ++k) {
      const int kp = k + s;                                   // H[n][k'] = h[f][k' - s]
      const float w = bank[(size_t)f * K + k] * kFtScale;
      const __half hi = __float2half_rn(w);
      const __half lo = __float2half_rn(w - __half2float(hi));
      const int kc = kp / 16, kb = (kp % 16) * 2;
      const size_t off = (size_t)kc * kFtBChunk + (size_t)(n / 8) * 256 + (size_t)(kb / 16) * 128 +
                         (size_t)(n % 8) * 16 + (size_t)(kb % 16);
      memcpy(&img[off], &hi, 2);
      memcpy(&img[(size_t)kFtKC * kFtBChunk + off], &lo, 2);
    }
  }
  if (*bimg_dev) cudaFree(*bimg_dev);
  *bimg_dev = nullptr;
  AASIST_CUDA(cudaMalloc(bimg_dev, img.size()));
  AASIST_CUDA(cudaMemcpy(*bimg_dev, img.data(), img.size(), cudaMemcpyHostToDevice));
  return 0;
}

// Freq_aug (models/AASIST.py:486-490): copy of the filter operand image with the rows of the masked filters zeroed.
// One thread per 16-byte unit; unit u of a K chunk holds row n = (u/16)*8 + u%8 (see tc_front_finalize).
__global__ void mask_front_image_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n_units,
                                        int mask_start, int mask_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_units) return;
  const int u = i % (kFtBChunk / 16);
  const int n = (u / 16) * 8 + (u % 8);
  const int f = 3 * (n / 9) + (n % 9) / 3;
  const bool masked = n < 9 * kSpecNodes && f >= mask_start && f < mask_start + mask_count;
  dst[i] = masked ? make_uint4(0u, 0u, 0u, 0u) : src[i];
}

int tc_front_mask(aasist_handle* h, const uint8_t* bimg, int mask_start, int mask_count, cudaStream_t st,
                  const uint8_t** out) {
  const size_t bytes = (size_t)2 * kFtKC * kFtBChunk;
  if (!h->front_bimg_masked) AASIST_CUDA(cudaMalloc(&h->front_bimg_masked, bytes));
  const int n_units = (int)(bytes / 16);
  {
    LaunchSpan span(h, "freq_mask_filter_image", st);
    mask_front_image_kernel<<<(n_units + 255) / 256, 256, 0, st>>>(reinterpret_cast<const uint4*>(bimg),
                                                                  reinterpret_cast<uint4*>(h->front_bimg_masked),
                                                                  n_units, mask_start, mask_count);
  }
  AASIST_CUDA(cudaGetLastError());
  *out = h->front_bimg_masked;
  return 0;
}

int launch_frontend_tc(aasist_handle* h, const uint8_t* bimg, int sm_count, const float* x, int B, int L,
                       float* out, cudaStream_t st) {
  const int Wp = (L - h->taps + 1) / 3;
  if (Wp < 1) {
    set_error("input length %d too short for a %d-tap filter bank", L, h->taps);
    return AASIST_E_INVALID;
  }
  FrontTcParams p;
  p.x = x; p.out = out; p.bimg = bimg; p.B = B; p.L = L; p.Wp = Wp;
  p.n_tiles_per_utt = (Wp + kFtTile - 1) / kFtTile;
  p.bn_scale = h->bn0_scale; p.bn_shift = h->bn0_shift;
  const size_t smem = 1024 + 2 * kFtKC * kFtBChunk + 2 * 4 * (kFtSeg + 16) * 2 + 256;
  AASIST_CUDA(cudaFuncSetAttribute(sinc_frontend_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(B * p.n_tiles_per_utt, sm_count);
  static int want_stats = -1;
  if (want_stats < 0) { const char* e = getenv("AASIST_TC_STATS"); want_stats = e ? atoi(e) : 0; }
#ifndef AASIST_KERNEL_STATS
  want_stats = 0;   // the instrumentation is compiled in only by tools/variant_build.sh -DAASIST_KERNEL_STATS
#endif
  p.stats = nullptr;
  p.collector = collector_mask() & 1;
  if (!h->range_flag) {
    AASIST_CUDA(cudaHostAlloc(&h->range_flag, sizeof(int), cudaHostAllocMapped));
    *h->range_flag = 0;
  }
  p.range_flag = h->range_flag;
  if (want_stats) {
    AASIST_CUDA(cudaMalloc(&p.stats, sizeof(long long) * 4 * grid));
    AASIST_CUDA(cudaMemset(p.stats, 0, sizeof(long long) * 4 * grid));
  }
  {
    LaunchSpan span(h, "sinc_frontend_tc", st);
    sinc_frontend_tc_kernel<<<grid, kFtThreads, smem, st>>>(p);
  }
  AASIST_CUDA(cudaGetLastError());
  if (want_stats) {   // debugging aid: where the MMA warp waits (cycles per tile, mean over CTAs)
    std::vector<long long> hst((size_t)4 * grid);
    AASIST_CUDA(cudaStreamSynchronize(st));
    AASIST_CUDA(cudaMemcpy(hst.data(), p.stats, sizeof(long long) * hst.size(), cudaMemcpyDeviceToHost));
    double acc[3] = {0, 0, 0};
    for (int c = 0; c < grid; ++c)
      for (int k = 0; k < 3; ++k) acc[k] += (double)hst[(size_t)c * 4 + k] / grid;
    const double tiles = (double)B * p.n_tiles_per_utt / grid;
    fprintf(stderr, "[sinc_frontend_tc stats] per tile cycles: total %.0f | wait afull(producers) %.0f tempty(epilogue) %.0f | "
            "issuing %.0f\n", acc[0] / tiles, acc[1] / tiles, acc[2] / tiles, (acc[0] - acc[1] - acc[2]) / tiles);
    cudaFree(p.stats);
  }
  return 0;
}

}  // namespace aasist
