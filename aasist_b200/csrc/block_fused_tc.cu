// A whole 32 -> 32 channel Residual_block (identity shortcut) in ONE kernel, its intermediate kept on chip
// (reference models/RawNetGatSpoofST.py:258-278 with nb_filts[0] == nb_filts[1]):
//
//   x pairs [B][23][3][J][64] --conv1 k(2,3) pad(1,1) + bn2 + SELU--> v (24 rows)   [TMEM -> smem, never in HBM]
//                             --conv2 k(2,3) pad(0,1) + x + max-pool 3--> pairs [B][23][3][Jn][64]  (or fp32 NCHW)
//
// Unfused, this block moves 93 MB per utterance for 4.1 GFLOP and is HBM-bound (profiles/README.md);
// fused it reads x once from HBM (21 MB; the second tap row and the identity hit L2) and writes 7 MB.
// Structure = block0_tc.cu with conv1 fed by TMA instead of an im2col producer:
//   warp 0 TMA producer (x tiles, 130-row boxes) | warp 1 MMA issuer: conv1 of v row r+1 (merged wider-N
//   groups, accumulators D1[r&1]) one row ahead of conv2 of v row r (two passes, accumulators D2) |
//   warps 2-9 epilogue | warps 10-17 transformers (D1 -> bias, SELU, zero-pad mask, fp16 pairs -> v ring).
// Work item: (utterance, strip of 126 pooled columns).
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "ptx.cuh"
#include "tc.cuh"

namespace aasist {

using namespace ptx;

constexpr int kBfStrip = 126;
constexpr int kBfSlab = 17 * 1024;
constexpr int kBfRows = 130;                // TMA box rows (j0-2 .. j0+127)
constexpr int kBfThreads = 576;
constexpr int kBfWBytes = 6 * 32 * 128;     // one 32->32 weight image (taps stored dw = 2,1,0 per dh)
constexpr int kBfNX = 4;                    // x ring slots (TMA)

struct BlockFusedParams {
  const __half* x;         // block input pairs [B][23][3][J][64] (identity operand)
  __half* out;             // pairs [B][23][3][Jn][64]   (or out_f32)
  float* out_f32;          // last block: (B,Co,23,Wo) fp32 NCHW
  const uint8_t* w1img;    // conv1 image (bn2 folded)
  const uint8_t* w2img;    // conv2 image
  const float* b1;         // [32]
  const float* b2;         // [32]
  int B, W, J, Wo, Jn, n_jt, n_vslots, Co;
  long long* stats;        // optional: MMA-warp wait cycles per CTA [total, d1empty, xfull, vfull, tempty]
};

__device__ __forceinline__ float bf_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float bf_selu(float v) {
  const float e = bf_ex2(v * 1.4426950408889634f);
  const float n = fminf(fmaf(e, kSeluScale * kSeluAlpha, -(kSeluScale * kSeluAlpha)), 0.f);
  return fmaf(fmaxf(v, 0.f), kSeluScale, n);
}
template <bool LOWER_BOUNDED>
__device__ __forceinline__ void bf_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = fminf(a, 65504.f);
  b = fminf(b, 65504.f);
  if (!LOWER_BOUNDED) {
    a = fmaxf(a, -65504.f);
    b = fmaxf(b, -65504.f);
  }
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(kBfThreads, 1)
block_fused_tc_kernel(const __grid_constant__ CUtensorMap tmX, const BlockFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_w1 = smem;
  uint8_t* s_w2 = smem + kBfWBytes;
  uint8_t* s_v = smem + 2 * kBfWBytes;                         // v ring
  uint8_t* s_x = s_v + (size_t)p.n_vslots * kBfSlab;           // x ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_x + kBfNX * kBfSlab);
  uint64_t* vfull = bars;                  // [8]  v tile written (4 transformer warps)
  uint64_t* vempty = bars + 8;             // [8]
  uint64_t* tfull = bars + 16;             // [2]  conv2 accumulators complete
  uint64_t* tempty = bars + 18;            // [2]  (8 epilogue warps)
  uint64_t* xfull = bars + 20;             // [4]  TMA
  uint64_t* xempty = bars + 24;            // [4]
  uint64_t* d1full = bars + 28;            // [2]  conv1 accumulators of a v row complete
  uint64_t* d1empty = bars + 30;           // [2]  drained (3 phase tiles x 4 warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 32);
  float* s_b1 = reinterpret_cast<float*>(bars + 34);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_strips = p.B * p.n_jt;

  for (int i = threadIdx.x; i < kBfWBytes / 16; i += kBfThreads) {
    reinterpret_cast<uint4*>(s_w1)[i] = __ldg(reinterpret_cast<const uint4*>(p.w1img) + i);
    reinterpret_cast<uint4*>(s_w2)[i] = __ldg(reinterpret_cast<const uint4*>(p.w2img) + i);
  }
  for (int i = threadIdx.x; i < p.n_vslots * kBfSlab / 16; i += kBfThreads)
    reinterpret_cast<uint4*>(s_v)[i] = make_uint4(0, 0, 0, 0);   // rows 128..135 are read by discarded rows only
  if (threadIdx.x < 32) s_b1[threadIdx.x] = __ldg(p.b1 + threadIdx.x);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&vfull[i], 4); mbar_init(&vempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    for (int i = 0; i < kBfNX; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&d1full[i], 1); mbar_init(&d1empty[i], 12); }
    fence_barrier_init();
    prefetch_tensormap(&tmX);
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr int D1_COL0 = 192;             // TMEM: [0,192) conv2 accumulators, [192,384) conv1 accumulators

  if (warp == 0) {
    // ======================================= TMA producer ======================================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < n_strips; t += gridDim.x) {
        const int jt = t % p.n_jt, b = t / p.n_jt;
        const int jbox = jt * kBfStrip - 2;                      // x tile row a <-> j = j0 - 2 + a
        for (int r = 0; r < 24; ++r)                             // v row r <- x rows r-1 (dh=0), r (dh=1)
          for (int dh = 0; dh < 2; ++dh) {
            const int xr = r + dh - 1;
            if (xr < 0 || xr > 22) continue;                     // conv1 zero padding: no tile, no MMA
            for (int phi = 0; phi < 3; ++phi) {
              mbar_wait(&xempty[slot], phase ^ 1);
              mbar_arrive_expect_tx(&xfull[slot], kBfRows * 128);
              tma_load_5d(s_x + (size_t)slot * kBfSlab, &tmX, &xfull[slot], 0, jbox, phi, xr, b);
              if (++slot == kBfNX) { slot = 0; phase ^= 1; }
            }
          }
      }
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer =======================================
    const bool leader = elect_one();
    const uint32_t w1_base = smem_u32(s_w1), w2_base = smem_u32(s_w2);
    const uint32_t v_base = smem_u32(s_v), x_base = smem_u32(s_x);
    int xslot = 0, vslot = 0;
    uint32_t xphase = 0, vphase = 0;
    int nrow1 = 0;                         // v rows whose conv1 has been issued (D1 buffer = nrow1 & 1)
    int nstart = 0;                        // conv2 output rows started
    long long w_d1 = 0, w_x = 0, w_v = 0, w_t = 0;
    const long long t_begin = clock64();

    // merged wider-N groups of one input tile (see conv_tc_kernel::issue_group)
    auto mma3 = [&](uint32_t d_tmem, uint32_t a_row, uint32_t w_row, int ntaps, bool fresh) {
      const uint32_t idesc = ntaps == 3 ? umma_idesc_f16(128, 96)
                                        : (ntaps == 2 ? umma_idesc_f16(128, 64) : umma_idesc_f16(128, 32));
      const uint64_t a_hi = umma_desc_sw128(a_row), a_lo = umma_desc_sw128(a_row + 64);
      const uint64_t w_hi = umma_desc_sw128(w_row), w_lo = umma_desc_sw128(w_row + 64);
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        umma_f16(d_tmem, a_hi + 2 * kc, w_hi + 2 * kc, idesc, (kc > 0 || !fresh) ? 1u : 0u);
        umma_f16(d_tmem, a_lo + 2 * kc, w_hi + 2 * kc, idesc, 1);
        umma_f16(d_tmem, a_hi + 2 * kc, w_lo + 2 * kc, idesc, 1);
      }
    };
    auto issue_group = [&](uint32_t a_slot, uint32_t wb, int phi, uint32_t d0, bool fresh) {
      if (phi == 1) {
        mma3(d0, a_slot + 128, wb, 3, fresh);
      } else if (phi == 0) {
        mma3(d0, a_slot + 128, wb + 4096, 2, fresh);
        mma3(d0 + 64, a_slot + 256, wb, 1, fresh);
      } else {
        mma3(d0 + 32, a_slot + 128, wb, 2, fresh);
        mma3(d0, a_slot, wb + 8192, 1, fresh);
      }
    };
    auto conv1_row = [&](int r) {          // all x tiles of v row r into D1[nrow1 & 1]
      const int buf = nrow1 & 1;
      { long long c0 = clock64(); mbar_wait(&d1empty[buf], ((nrow1 >> 1) & 1) ^ 1); w_d1 += clock64() - c0; }
      tc_fence_after_sync();
      const uint32_t d0 = tmem_base + (uint32_t)(D1_COL0 + buf * 96);
      bool fresh = true;
      for (int dh = 0; dh < 2; ++dh) {
        const int xr = r + dh - 1;
        if (xr < 0 || xr > 22) continue;
        for (int phi = 0; phi < 3; ++phi) {
          { long long c0 = clock64(); mbar_wait(&xfull[xslot], xphase); w_x += clock64() - c0; }
          tc_fence_after_sync();
          if (leader) {
            issue_group(x_base + (uint32_t)xslot * kBfSlab, w1_base + (uint32_t)(dh * 3 * 4096), phi, d0,
                        fresh && phi == 0);
            umma_commit(&xempty[xslot]);
          }
          __syncwarp();
          if (++xslot == kBfNX) { xslot = 0; xphase ^= 1; }
        }
        fresh = false;
      }
      if (leader) umma_commit(&d1full[buf]);
      __syncwarp();
      ++nrow1;
    };
    auto vadvance = [&](int& sl, uint32_t& ph) {
      if (++sl == p.n_vslots) { sl = 0; ph ^= 1; }
    };

    for (int t = blockIdx.x; t < n_strips; t += gridDim.x) {
      int buf_open = 0;
      conv1_row(0);
      for (int r = 0; r < 24; ++r) {
        if (r < 23) conv1_row(r + 1);                            // one v row ahead of conv2
        const bool has_o1 = r >= 1, has_o0 = r <= 22;
        int sl[3];
        uint32_t ph[3];
        {
          int s2 = vslot;
          uint32_t p2 = vphase;
          for (int i = 0; i < 3; ++i) { sl[i] = s2; ph[i] = p2; vadvance(s2, p2); }
        }
        for (int phi = 0; phi < 3; ++phi) {                      // pass 1: dh=1 completes output row r-1
          { long long c0 = clock64(); mbar_wait(&vfull[sl[phi]], ph[phi]); w_v += clock64() - c0; }
          tc_fence_after_sync();
          if (has_o1 && leader)
            issue_group(v_base + (uint32_t)sl[phi] * kBfSlab, w2_base + 3 * 4096, phi,
                        tmem_base + (uint32_t)(buf_open * 96), false);
          __syncwarp();
        }
        if (has_o1) {
          if (leader) umma_commit(&tfull[buf_open]);
          __syncwarp();
        }
        if (has_o0) {                                            // pass 2: dh=0 starts output row r
          buf_open = nstart & 1;
          { long long c0 = clock64(); mbar_wait(&tempty[buf_open], ((nstart >> 1) & 1) ^ 1); w_t += clock64() - c0; }
          tc_fence_after_sync();
          ++nstart;
          for (int phi = 0; phi < 3; ++phi) {
            if (leader) {
              issue_group(v_base + (uint32_t)sl[phi] * kBfSlab, w2_base, phi, tmem_base + (uint32_t)(buf_open * 96),
                          phi == 0);
              umma_commit(&vempty[sl[phi]]);
            }
            __syncwarp();
          }
        } else {
          for (int phi = 0; phi < 3; ++phi) {
            if (leader) umma_commit(&vempty[sl[phi]]);
            __syncwarp();
          }
        }
        for (int i = 0; i < 3; ++i) vadvance(vslot, vphase);
      }
    }
    if (p.stats && leader) {
      long long* stt = p.stats + (size_t)blockIdx.x * 8;
      stt[0] = clock64() - t_begin; stt[1] = w_d1; stt[2] = w_x; stt[3] = w_v; stt[4] = w_t;
    }
  } else if (warp >= 2 && warp < 10) {
    // ======================================= epilogue =========================================
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int col0 = half * 16;
    float bias[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bias[i] = __ldg(p.b2 + col0 + i);
    int tcount = 0;
    for (int t = blockIdx.x; t < n_strips; t += gridDim.x) {
      const int jt = t % p.n_jt, b = t / p.n_jt;
      const int j = jt * kBfStrip + m;
      const bool live = m < kBfStrip;
      const bool valid = j < p.Wo;
      for (int h = 0; h < 23; ++h, ++tcount) {
        const int buf = tcount & 1;
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 96 + col0);
        mbar_wait(&tfull[buf], (tcount >> 1) & 1);
        tc_fence_after_sync();
        uint32_t acc[3][16];
#pragma unroll
        for (int s = 0; s < 3; ++s) tmem_ld16_async(t_row + (uint32_t)(s * 32), acc[s]);
#pragma unroll
        for (int s = 0; s < 3; ++s) tmem_ld_wait16(acc[s]);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        if (!live) continue;
        float mx[16];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          // identity operand x[b][h][s][j][col0..col0+16) (L2-resident: conv1 just read this row)
          uint32_t ih[8], il[8];
          if (j < p.J) {
            const __half* xs = p.x + ((((size_t)b * 23 + h) * 3 + s) * p.J + j) * 64 + col0;
            ld_global_nc_256(xs, ih);
            ld_global_nc_256(xs + 32, il);                       // lo half-row: 32 halves = 64 B later
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) ih[k] = il[k] = 0u;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&ih[k]));
            const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&il[k]));
            const float v0 = __uint_as_float(acc[s][2 * k]) + (fa.x + fb.x);
            const float v1 = __uint_as_float(acc[s][2 * k + 1]) + (fa.y + fb.y);
            mx[2 * k] = s == 0 ? v0 : fmaxf(mx[2 * k], v0);
            mx[2 * k + 1] = s == 0 ? v1 : fmaxf(mx[2 * k + 1], v1);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) mx[i] = valid ? mx[i] + bias[i] : 0.f;
        if (p.out_f32) {
          if (valid)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int ch = col0 + i;
              if (ch < p.Co) p.out_f32[(((size_t)b * p.Co + ch) * 23 + h) * p.Wo + j] = mx[i];
            }
        } else if (j / 3 < p.Jn) {
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) bf_split2<false>(mx[2 * i], mx[2 * i + 1], hw[i], lw[i]);
          __half* o = p.out + ((((size_t)b * 23 + h) * 3 + (j % 3)) * p.Jn + j / 3) * 64 + col0;
          st_global_256(o, hw);
          st_global_256(o + 32, lw);
        }
      }
    }
  } else if (warp >= 10) {
    // ============ transformers: D1 (TMEM) -> bias, SELU, zero-pad mask, fp16 pairs -> swizzled v tile ============
    const int quad = warp & 3, grp = (warp - 10) >> 2;           // two groups of four warps take alternate tiles
    const int jj = quad * 32 + lane;
    const uint32_t row_off = (uint32_t)jj * 128;
    const uint32_t sw = (uint32_t)(jj & 7);
    int n = 0, slot = 0, nrow = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < n_strips; t += gridDim.x) {
      const int jt = t % p.n_jt;
      const int j = jt * kBfStrip - 1 + jj;
      for (int r = 0; r < 24; ++r, ++nrow) {
        const int buf = nrow & 1;
        for (int s = 0; s < 3; ++s, ++n) {
          if ((n & 1) == grp) {
            const int pos = 3 * j + s;
            const bool valid = j >= 0 && pos < p.W;              // conv2 zero-pads v itself
            mbar_wait(&d1full[buf], (nrow >> 1) & 1);
            tc_fence_after_sync();
            uint32_t acc[2][16];
            const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(D1_COL0 + buf * 96 + s * 32);
            tmem_ld16_async(ta, acc[0]);
            tmem_ld16_async(ta + 16, acc[1]);
            tmem_ld_wait16(acc[0]);
            tmem_ld_wait16(acc[1]);
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d1empty[buf]);
            uint32_t hw[16], lw[16];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 bb = *reinterpret_cast<const float2*>(s_b1 + c * 16 + 2 * i);
                float x0 = bf_selu(__uint_as_float(acc[c][2 * i]) + bb.x);
                float x1 = bf_selu(__uint_as_float(acc[c][2 * i + 1]) + bb.y);
                if (!valid) { x0 = 0.f; x1 = 0.f; }
                bf_split2<true>(x0, x1, hw[c * 8 + i], lw[c * 8 + i]);
              }
            mbar_wait(&vempty[slot], phase ^ 1);
            uint8_t* row = s_v + (size_t)slot * kBfSlab + row_off;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              *reinterpret_cast<uint4*>(row + (((uint32_t)q ^ sw) << 4)) =
                  make_uint4(hw[4 * q], hw[4 * q + 1], hw[4 * q + 2], hw[4 * q + 3]);
              *reinterpret_cast<uint4*>(row + (((uint32_t)(4 + q) ^ sw) << 4)) =
                  make_uint4(lw[4 * q], lw[4 * q + 1], lw[4 * q + 2], lw[4 * q + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&vfull[slot]);
          }
          if (++slot == p.n_vslots) { slot = 0; phase ^= 1; }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

int launch_block_fused_tc(aasist_handle* h, int sm_count, const char* name, const CUtensorMap& tmX,
                          const uint8_t* w1img, const uint8_t* w2img, const float* b1, const float* b2,
                          const __half* x, int nb, int W, int Co, __half* out, float* out_f32, cudaStream_t st) {
  BlockFusedParams p;
  p.x = x; p.out = out; p.out_f32 = out_f32; p.w1img = w1img; p.w2img = w2img; p.b1 = b1; p.b2 = b2;
  p.B = nb; p.W = W; p.J = (W + 2) / 3; p.Wo = W / 3; p.Jn = (p.Wo + 2) / 3; p.Co = Co;
  p.n_jt = (std::max(p.J, 3 * p.Jn) + kBfStrip - 1) / kBfStrip;
  const int fixed = 1024 + 2 * kBfWBytes + kBfNX * kBfSlab + 512;
  p.n_vslots = std::min(8, (227 * 1024 - fixed) / kBfSlab);
  if (p.n_vslots < 6) {
    set_error("block_fused_tc: shared memory budget allows only %d v-ring slots", p.n_vslots);
    return AASIST_E_INVALID;
  }
  const size_t smem = (size_t)fixed + (size_t)p.n_vslots * kBfSlab;
  AASIST_CUDA(cudaFuncSetAttribute(block_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(nb * p.n_jt, sm_count);
  static int want_stats = -1;
  if (want_stats < 0) { const char* e = getenv("AASIST_BF_STATS"); want_stats = e ? atoi(e) : 0; }
  p.stats = nullptr;
  if (want_stats) {
    AASIST_CUDA(cudaMalloc(&p.stats, sizeof(long long) * 8 * grid));
    AASIST_CUDA(cudaMemset(p.stats, 0, sizeof(long long) * 8 * grid));
  }
  {
    LaunchSpan span(h, name, st);
    block_fused_tc_kernel<<<grid, kBfThreads, smem, st>>>(tmX, p);
  }
  AASIST_CUDA(cudaGetLastError());
  if (want_stats) {   // debugging aid: where the MMA warp waits (cycles per row-tile, mean over CTAs)
    std::vector<long long> hst((size_t)8 * grid);
    AASIST_CUDA(cudaStreamSynchronize(st));
    AASIST_CUDA(cudaMemcpy(hst.data(), p.stats, sizeof(long long) * hst.size(), cudaMemcpyDeviceToHost));
    double acc[5] = {0, 0, 0, 0, 0};
    for (int c = 0; c < grid; ++c)
      for (int k = 0; k < 5; ++k) acc[k] += (double)hst[(size_t)c * 8 + k] / grid;
    const double rows = (double)nb * p.n_jt * 23 / grid;
    fprintf(stderr, "[%s stats] per row-tile cycles: total %.0f | wait d1empty %.0f xfull %.0f vfull %.0f tempty %.0f | "
            "issuing %.0f\n", name, acc[0] / rows, acc[1] / rows, acc[2] / rows, acc[3] / rows, acc[4] / rows,
            (acc[0] - acc[1] - acc[2] - acc[3] - acc[4]) / rows);
    cudaFree(p.stats);
  }
  return 0;
}

}  // namespace aasist
