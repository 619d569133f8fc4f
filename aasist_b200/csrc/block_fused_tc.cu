// A whole 32 -> 32 channel Residual_block (identity shortcut) in ONE kernel, its intermediate kept on chip
// (reference models/RawNetGatSpoofST.py:258-278 with nb_filts[0] == nb_filts[1]):
//
//   x pairs [B][23][3][J][64] --conv1 k(2,3) pad(1,1) + bn2 + SELU--> v (24 rows)   [TMEM -> smem, never in HBM]
//                             --conv2 k(2,3) pad(0,1) + x + max-pool 3--> pairs [B][23][3][Jn][64]  (or fp32 NCHW)
//
// Unfused, this block moves 93 MB per utterance for 4.1 GFLOP and is HBM-bound (profiles/README.md);
// fused it reads x once from HBM (21 MB; the second tap row and the identity hit L2) and writes 7 MB.
// Structure = block0_tc.cu with conv1 fed by TMA instead of an im2col producer:
//   warp 0 TMA producer (x tiles, 130-row boxes, every tile loaded once) | warp 1 MMA issuer: conv1 one v row
//   ahead of conv2 | warps 2-9 epilogue | warps 10-17 transformers (D1 -> bias, SELU, zero-pad mask, fp16
//   pairs -> v ring).  Work item: (utterance, strip of 126 pooled columns).
//
// The MMAs are bound by their ~53-cycle issue/operand-fetch floor, not by math, so every A tile is used by as FEW
// instructions as possible: an input row feeds the output row it completes (tap row dh=1) and the one it
// starts (dh=0), and the two accumulator rows sit side by side in TMEM -- columns [phase s][slot][32 ch] --
// so ONE wider-N MMA (N = 64 / 128 / 192 for 1 / 2 / 3 merged column taps) serves both.  The B rows follow
// the same order; because a row changes role (new -> old) between consecutive steps while it stays in
// its slot, shared memory holds the weight image in both slot orders.  A merged MMA has one accumulate flag
// for all its columns, so accumulators are never overwritten: whoever drains a slot (epilogue / transformer
// warps) stores zeros back with tcgen05.st before releasing it, and every MMA accumulates.
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
constexpr int kBfThreads = 576;              // 18 warps: 96 registers per thread
constexpr int kBfWBytes = 6 * 32 * 128;     // one 32->32 weight image (taps stored dw = 2,1,0 per dh)
constexpr int kBfMaxX = 4;                  // x ring slots (TMA), upper bound

struct BlockFusedParams {
  const __half* x;         // block input pairs [B][23][3][J][64] (identity operand)
  __half* out;             // pairs [B][23][3][Jn][64]   (or out_f32)
  float* out_f32;          // last block: (B,Co,23,Wo) fp32 NCHW
  const uint8_t* w1img;    // conv1 image (bn2 folded)
  const uint8_t* w2img;    // conv2 image
  const float* b1;         // [32]
  const float* b2;         // [32]
  int B, W, J, Wo, Jn, n_jt, n_vslots, n_xslots, Co;
  long long* stats;        // optional: MMA-warp wait cycles per CTA [total, d1empty, xfull, vfull, tempty]
  int products;            // 3 (f16x3) or 2 (f16x2: no a_hi * w_lo weight-correction product)
  int collector;           // A-operand collector reuse between the two a_hi products (tc.cuh collector_mask)
};

__device__ __forceinline__ float bf_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// SELU of v given y = v * log2(e): conv1's weights and bias are pre-scaled by log2(e) on the host and the bias
// is what the conv1 accumulators start from (the transformers store it back when they drain a slot)
__device__ __forceinline__ float bf_selu_scaled(float y) {
  const float e = bf_ex2(y);
  const float n = fminf(fmaf(e, kSeluScale * kSeluAlpha, -(kSeluScale * kSeluAlpha)), 0.f);
  return fmaf(fmaxf(y, 0.f), kSeluScale * 0.6931471805599453f, n);
}

__global__ void __launch_bounds__(kBfThreads, 1)
block_fused_tc_kernel(const __grid_constant__ CUtensorMap tmX, const BlockFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_w1 = smem;                                        // conv1 image, slot order 0 then 1
  uint8_t* s_w2 = smem + 2 * kBfWBytes;                        // conv2 image, slot order 0 then 1
  uint8_t* s_v = smem + 4 * kBfWBytes;                         // v ring
  uint8_t* s_x = s_v + (size_t)p.n_vslots * kBfSlab;           // x ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_x + (size_t)p.n_xslots * kBfSlab);
  uint64_t* vfull = bars;                  // [8]  v tile written (8 transformer warps)
  uint64_t* vempty = bars + 8;             // [8]
  uint64_t* tfull = bars + 16;             // [2]  conv2 accumulators complete
  uint64_t* tempty = bars + 18;            // [2]  (8 epilogue warps)
  uint64_t* xfull = bars + 20;             // [4]  TMA
  uint64_t* xempty = bars + 24;            // [4]
  uint64_t* d1full = bars + 28;            // [2]  conv1 accumulators of a v row complete
  uint64_t* d1empty = bars + 30;           // [2]  drained (8 transformer warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 32);
  float* s_b1 = reinterpret_cast<float*>(bars + 34);
  float* s_b2 = s_b1 + 32;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_strips = p.B * p.n_jt;

  // weight images: global [dh][tap 2,1,0][32 rows] -> shared [conv][order o][tap][slot][32 rows], where
  // slot sigma of order o holds tap row dh = sigma ^ o (4 KB blocks, swizzle pattern preserved)
  for (int i = threadIdx.x; i < 4 * kBfWBytes / 16; i += kBfThreads) {
    const int blk = i >> 8, within = i & 255;
    const int conv = blk / 12, o = (blk / 6) & 1, tap = (blk % 6) >> 1, sigma = blk & 1;
    const uint8_t* src = (conv ? p.w2img : p.w1img) + (size_t)(((sigma ^ o) * 3 + tap) * 4096);
    reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(src) + within);
  }
  for (int i = threadIdx.x; i < p.n_vslots * kBfSlab / 16; i += kBfThreads)
    reinterpret_cast<uint4*>(s_v)[i] = make_uint4(0, 0, 0, 0);   // rows 128..135 are read by discarded rows only
  if (threadIdx.x < 32) {
    s_b1[threadIdx.x] = __ldg(p.b1 + threadIdx.x);
    s_b2[threadIdx.x] = __ldg(p.b2 + threadIdx.x);
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&vfull[i], 8); mbar_init(&vempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    for (int i = 0; i < kBfMaxX; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&d1full[i], 1); mbar_init(&d1empty[i], 8); }
    fence_barrier_init();
    prefetch_tensormap(&tmX);
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  // TMEM: [0,192) conv2 accumulators, [192,384) conv1 accumulators; column = s*64 + slot*32 + channel
  constexpr int D1_COL0 = 192;
  if (warp >= 2) {   // conv2 accumulators start at zero, conv1 accumulators at the bias (and return there after every drain)
    const int part = (warp - 2) >> 2;      // 0,1: D2 halves; 2,3: D1 halves
    const uint32_t tz = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(part * 96);
    if (part < 2) {
#pragma unroll
      for (int c = 0; c < 6; ++c) tmem_st16_zero(tz + (uint32_t)(c * 16));
    } else {
      uint32_t bb[2][16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        bb[0][i] = __float_as_uint(s_b1[i]);
        bb[1][i] = __float_as_uint(s_b1[16 + i]);
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) tmem_st16(tz + (uint32_t)(c * 16), bb[c & 1]);   // 32-column [slot] blocks
    }
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  if (warp == 0) {
    // ======================================= TMA producer ======================================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < n_strips; t += gridDim.x) {
        const int jt = t % p.n_jt, b = t / p.n_jt;
        const int jbox = jt * kBfStrip - 2;                      // x tile row a <-> j = j0 - 2 + a
        for (int xr = 0; xr < 23; ++xr)                          // x row xr: completes v row xr, starts v row xr+1
          for (int phi = 0; phi < 3; ++phi) {
            mbar_wait(&xempty[slot], phase ^ 1);
            mbar_arrive_expect_tx(&xfull[slot], kBfRows * 128);
            tma_load_5d(s_x + (size_t)slot * kBfSlab, &tmX, &xfull[slot], 0, jbox, phi, xr, b);
            if (++slot == p.n_xslots) { slot = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ======================================= MMA issuer =======================================
    // the whole schedule runs in ONE elected lane: tcgen05.mma / tcgen05.commit are single-thread instructions
    // and a single-lane loop pays neither divergence nor __syncwarp per step
    const bool leader = elect_one();
    if (leader) {
    const uint32_t w1_base = smem_u32(s_w1), w2_base = smem_u32(s_w2);
    const uint32_t v_base = smem_u32(s_v), x_base = smem_u32(s_x);
    int xslot = 0, vslot = 0;
    uint32_t xphase = 0, vphase = 0;
    int vbase = 0;                         // global index of this strip's v row 0 (v row n lives in D1 slot n & 1)
    int g2 = 0;                            // conv2 steps issued (step g starts an output row in D2 slot g & 1)
    long long w_d1 = 0, w_x = 0, w_v = 0, w_t = 0;
    const long long t_begin = AASIST_CLOCK();

    // one A tile (128 rows x [hi|lo]) against `ntaps` merged column taps x both slots: N = 64 * ntaps,
    // three fp16 products, everything accumulates (the slots were zeroed when they were drained)
    const bool three = p.products == 3;
    const bool coll1 = (p.collector & 1) != 0, coll2 = (p.collector & 2) != 0;   // conv1 (TMA x tiles) / conv2 (v tiles)
    auto mma3 = [&](uint32_t d_tmem, uint32_t a_row, uint32_t w_row, int ntaps, bool coll) {
      const uint32_t idesc = ntaps == 3 ? umma_idesc_f16(128, 192)
                                        : (ntaps == 2 ? umma_idesc_f16(128, 128) : umma_idesc_f16(128, 64));
      const uint64_t a_hi = umma_desc_sw128(a_row), a_lo = umma_desc_sw128(a_row + 64);
      const uint64_t w_hi = umma_desc_sw128(w_row), w_lo = umma_desc_sw128(w_row + 64);
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        if (three && coll) {     // a_hi * w_hi and a_hi * w_lo back to back: the second takes A from the collector
          umma_f16_keep(d_tmem, a_hi + 2 * kc, w_hi + 2 * kc, idesc, 1);
          umma_f16_reuse(d_tmem, a_hi + 2 * kc, w_lo + 2 * kc, idesc, 1);
          umma_f16(d_tmem, a_lo + 2 * kc, w_hi + 2 * kc, idesc, 1);
        } else {
          umma_f16(d_tmem, a_hi + 2 * kc, w_hi + 2 * kc, idesc, 1);
          umma_f16(d_tmem, a_lo + 2 * kc, w_hi + 2 * kc, idesc, 1);
          if (three) umma_f16(d_tmem, a_hi + 2 * kc, w_lo + 2 * kc, idesc, 1);
        }
      }
    };
    // pool phase s is served by column tap dw with (s + dw - 1) == phi (mod 3), from A rows shifted by
    // floor((s + dw - 1) / 3); taps are stored dw = 2,1,0 (8 KB each: both slots), D columns s*64
    auto issue_group = [&](uint32_t a_slot, uint32_t wb, int phi, uint32_t d0, bool coll) {
      if (phi == 1) {
        mma3(d0, a_slot + 128, wb, 3, coll);
      } else if (phi == 0) {
        mma3(d0, a_slot + 128, wb + 8192, 2, coll);
        mma3(d0 + 128, a_slot + 256, wb, 1, coll);
      } else {
        mma3(d0 + 64, a_slot + 128, wb, 2, coll);
        mma3(d0, a_slot, wb + 16384, 1, coll);
      }
    };
    auto d1_claim = [&](int n) {           // v row n is about to receive its first MMA: its slot must be drained
      AASIST_TIMED_WAIT(&d1empty[n & 1], ((n >> 1) & 1) ^ 1, w_d1);
    };
    auto conv1_step = [&](int h) {         // x row h: dh=1 completes v row h, dh=0 starts v row h+1
      const int n_old = vbase + h, n_new = n_old + 1;
      if (h == 0) d1_claim(n_old);         // v row 0 has no dh=0 contribution (zero padding): it starts here
      d1_claim(n_new);
      tc_fence_after_sync();
      const uint32_t wb = w1_base + (uint32_t)((n_new & 1) * kBfWBytes);   // order o: slot (n_new & 1) gets dh=0
      for (int phi = 0; phi < 3; ++phi) {
        AASIST_TIMED_WAIT(&xfull[xslot], xphase, w_x);
        tc_fence_after_sync();
        {
          issue_group(x_base + (uint32_t)xslot * kBfSlab, wb, phi, tmem_base + (uint32_t)D1_COL0, coll1);
          umma_commit(&xempty[xslot]);
        }
        if (++xslot == p.n_xslots) { xslot = 0; xphase ^= 1; }
      }
      {
        umma_commit(&d1full[n_old & 1]);
        if (h == 22) umma_commit(&d1full[n_new & 1]);                // v row 23 has no dh=1 contribution
      }
    };
    auto conv2_step = [&]() {              // v row r: dh=1 completes output row r-1, dh=0 starts output row r
      const int g = g2++;
      // slot g & 1 was last completed by step g-1 (the epilogue's (g-1)>>1-th drain of that slot); g = 0: none
      AASIST_TIMED_WAIT(&tempty[g & 1], (uint32_t)(((g - 1) >> 1) & 1), w_t);
      tc_fence_after_sync();
      const uint32_t wb = w2_base + (uint32_t)((g & 1) * kBfWBytes);
      for (int phi = 0; phi < 3; ++phi) {
        AASIST_TIMED_WAIT(&vfull[vslot], vphase, w_v);
        tc_fence_after_sync();
        {
          issue_group(v_base + (uint32_t)vslot * kBfSlab, wb, phi, tmem_base, coll2);
          umma_commit(&vempty[vslot]);
        }
        if (++vslot == p.n_vslots) { vslot = 0; vphase ^= 1; }
      }
      // the row in the other slot is complete.  For v row 0 that "row" is a dummy (output row -1: it holds the
      // dh=0 product of the previous strip's v row 23 plus this dh=1 product); the epilogue just clears it.
      umma_commit(&tfull[(g & 1) ^ 1]);
    };

    for (int t = blockIdx.x; t < n_strips; t += gridDim.x) {
      conv1_step(0);
      for (int r = 0; r < 24; ++r) {
        if (r + 1 <= 22) conv1_step(r + 1);                        // one v row ahead of conv2
        conv2_step();
      }
      vbase += 24;
    }
    if (p.stats) {
      long long* stt = p.stats + (size_t)blockIdx.x * 16;
      stt[0] = AASIST_CLOCK() - t_begin; stt[1] = w_d1; stt[2] = w_x; stt[3] = w_v; stt[4] = w_t;
    }
    }
  } else if (warp >= 2 && warp < 10) {
    // ======================================= epilogue =========================================
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int col0 = half * 16;
    int tcount = 0;                        // completed conv2 steps (24 per strip; the first is the dummy row -1)
    for (int t = blockIdx.x; t < n_strips; t += gridDim.x) {
      const int jt = t % p.n_jt, b = t / p.n_jt;
      const int j = jt * kBfStrip + m;
      const bool live = m < kBfStrip;
      const bool valid = j < p.Wo;
      const int jc = min(j, p.J - 1);                              // j >= J is never stored: any address will do
      // The identity operand x[b][h][s][j][col0..col0+16) (hi + lo) of the NEXT row is fetched from L2 while this
      // thread would otherwise idle, summed to fp32 and parked in spare TMEM columns [384,480): it costs no
      // registers across the wait for the accumulators, and its latency is off the drain path.
      const uint32_t t_idn = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(384 + col0);
      for (int h = -1; h < 23; ++h, ++tcount) {
        const int buf = (tcount & 1) ^ 1;
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 32 + col0);
        mbar_wait(&tfull[buf], (tcount >> 1) & 1);
        tc_fence_after_sync();
        uint32_t acc[3][16];
        if (h >= 0) {
#pragma unroll
          for (int s = 0; s < 3; ++s) tmem_ld16_async(t_row + (uint32_t)(s * 64), acc[s]);
#pragma unroll
          for (int s = 0; s < 3; ++s) tmem_ld_wait16(acc[s]);
        }
#pragma unroll
        for (int s = 0; s < 3; ++s) tmem_st16_zero(t_row + (uint32_t)(s * 64));
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        float mx[16];
        if (h >= 0) {
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            uint32_t idn[16];
            tmem_ld16_async(t_idn + (uint32_t)(s * 32), idn);
            tmem_ld_wait16(idn);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float v = __uint_as_float(acc[s][i]) + __uint_as_float(idn[i]);
              mx[i] = s == 0 ? v : fmaxf(mx[i], v);
            }
          }
        }
        if (live && h >= 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx[i] = valid ? mx[i] + s_b2[col0 + i] : 0.f;
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
            for (int i = 0; i < 8; ++i) split2_sat(mx[2 * i], mx[2 * i + 1], hw[i], lw[i]);
            __half* o = p.out + ((((size_t)b * 23 + h) * 3 + (j % 3)) * p.Jn + j / 3) * 64 + col0;
            st_global_256(o, hw);
            st_global_256(o + 32, lw);
          }
        }
        if (h < 22) {                                              // park the next row's identity operand
          const __half* xs = p.x + (((size_t)b * 23 + (h + 1)) * 3 * p.J + jc) * 64 + col0;
          uint32_t ih[3][8], il[3][8];
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            // (no L1 allocation: the L1 data array is the shared-memory array the MMAs fetch operands from)
            ld_global_na_256(xs + (size_t)s * p.J * 64, ih[s]);
            ld_global_na_256(xs + (size_t)s * p.J * 64 + 32, il[s]);   // lo half-row: 32 halves = 64 B later
          }
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            uint32_t sum[16];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&ih[s][k]));
              const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&il[s][k]));
              sum[2 * k] = __float_as_uint(fa.x + fb.x);
              sum[2 * k + 1] = __float_as_uint(fa.y + fb.y);
            }
            tmem_st16(t_idn + (uint32_t)(s * 32), sum);
          }
          tmem_st_wait();
        }
      }
    }
  } else if (warp >= 10) {
    // ============ transformers: D1 (TMEM) -> bias, SELU, zero-pad mask, fp16 pairs -> swizzled v tiles ============
    // warp = (TMEM lane quadrant, 16-channel half); the whole v row (3 phase tiles) is read and its slot
    // cleared and released at once, so conv1 can start the row after next while the math is still running
    const int quad = warp & 3, half = (warp - 10) >> 2;
    const int jj = quad * 32 + lane;
    const int col0 = half * 16;
    const uint32_t row_off = (uint32_t)jj * 128;
    const uint32_t sw = (uint32_t)(jj & 7);
    const uint32_t c_hi = (uint32_t)(2 * half), c_lo = (uint32_t)(4 + 2 * half);   // 16-byte chunks of this half
    int slot = 0, nrow = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < n_strips; t += gridDim.x) {
      const int jt = t % p.n_jt;
      const int j = jt * kBfStrip - 1 + jj;
      // warp-uniform: every row of this warp lies inside [0, W) for all three phases
      const bool valid_all = jt * kBfStrip - 1 + quad * 32 >= 0 && 3 * (jt * kBfStrip - 1 + quad * 32 + 31) + 2 < p.W;
      for (int r = 0; r < 24; ++r, ++nrow) {
        const int buf = nrow & 1;
        const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(D1_COL0 + buf * 32 + col0);
        mbar_wait(&d1full[buf], (nrow >> 1) & 1);
        tc_fence_after_sync();
        uint32_t acc[3][16];
#pragma unroll
        for (int s = 0; s < 3; ++s) tmem_ld16_async(ta + (uint32_t)(s * 64), acc[s]);
#pragma unroll
        for (int s = 0; s < 3; ++s) tmem_ld_wait16(acc[s]);
        {
          uint32_t bb[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 q = *reinterpret_cast<const uint4*>(s_b1 + col0 + 4 * i);
            bb[4 * i] = q.x; bb[4 * i + 1] = q.y; bb[4 * i + 2] = q.z; bb[4 * i + 3] = q.w;
          }
#pragma unroll
          for (int s = 0; s < 3; ++s) tmem_st16(ta + (uint32_t)(s * 64), bb);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d1empty[buf]);
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          uint32_t hw[8], lw[8];
          if (valid_all) {                                         // interior strip: no zero-padding mask needed
#pragma unroll
            for (int i = 0; i < 8; ++i)
              split2_sat(bf_selu_scaled(__uint_as_float(acc[s][2 * i])),
                              bf_selu_scaled(__uint_as_float(acc[s][2 * i + 1])), hw[i], lw[i]);
          } else {
            const bool valid = j >= 0 && 3 * j + s < p.W;          // conv2 zero-pads v itself
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float x0 = bf_selu_scaled(__uint_as_float(acc[s][2 * i]));
              float x1 = bf_selu_scaled(__uint_as_float(acc[s][2 * i + 1]));
              if (!valid) { x0 = 0.f; x1 = 0.f; }
              split2_sat(x0, x1, hw[i], lw[i]);
            }
          }
          mbar_wait(&vempty[slot], phase ^ 1);
          uint8_t* row = s_v + (size_t)slot * kBfSlab + row_off;
          *reinterpret_cast<uint4*>(row + ((c_hi ^ sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(row + (((c_hi + 1) ^ sw) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
          *reinterpret_cast<uint4*>(row + ((c_lo ^ sw) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          *reinterpret_cast<uint4*>(row + (((c_lo + 1) ^ sw) << 4)) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&vfull[slot]);
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
  static int xslots = -1;
  if (xslots < 0) { const char* e = getenv("AASIST_BF_XSLOTS"); xslots = e ? std::min(kBfMaxX, std::max(2, atoi(e))) : 3; }
  p.n_xslots = xslots;
  const int fixed = 1024 + 4 * kBfWBytes + p.n_xslots * kBfSlab + 1024;
  p.n_vslots = std::min(8, (227 * 1024 - fixed) / kBfSlab);
  if (p.n_vslots < 3) {
    set_error("block_fused_tc: shared memory budget allows only %d v-ring slots", p.n_vslots);
    return AASIST_E_INVALID;
  }
  const size_t smem = (size_t)fixed + (size_t)p.n_vslots * kBfSlab;
  AASIST_CUDA(cudaFuncSetAttribute(block_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(nb * p.n_jt, sm_count);
  static int want_stats = -1;
  if (want_stats < 0) { const char* e = getenv("AASIST_BF_STATS"); want_stats = e ? atoi(e) : 0; }
#ifndef AASIST_KERNEL_STATS
  want_stats = 0;   // the instrumentation is compiled in only by tools/variant_build.sh -DAASIST_KERNEL_STATS
#endif
  p.stats = nullptr;
  p.products = h->cfg.precision == AASIST_PREC_F16X2 ? 2 : 3;
  p.collector = ((collector_mask() >> 2) & 1) | (((collector_mask() >> 6) & 1) << 1);
  if (want_stats) {
    AASIST_CUDA(cudaMalloc(&p.stats, sizeof(long long) * 16 * grid));
    AASIST_CUDA(cudaMemset(p.stats, 0, sizeof(long long) * 16 * grid));
  }
  {
    LaunchSpan span(h, name, st);
    block_fused_tc_kernel<<<grid, kBfThreads, smem, st>>>(tmX, p);
  }
  AASIST_CUDA(cudaGetLastError());
  if (want_stats) {   // debugging aid: where the MMA warp waits (cycles per row-tile, mean over CTAs)
    std::vector<long long> hst((size_t)16 * grid);
    AASIST_CUDA(cudaStreamSynchronize(st));
    AASIST_CUDA(cudaMemcpy(hst.data(), p.stats, sizeof(long long) * hst.size(), cudaMemcpyDeviceToHost));
    double acc[16] = {0};
    for (int c = 0; c < grid; ++c)
      for (int k = 0; k < 16; ++k) acc[k] += (double)hst[(size_t)c * 16 + k] / grid;
    const double rows = (double)nb * p.n_jt * 23 / grid;
    fprintf(stderr, "[%s stats] per row-tile cycles: total %.0f | wait d1empty %.0f xfull %.0f vfull %.0f tempty %.0f | "
            "issuing %.0f\n", name, acc[0] / rows, acc[1] / rows, acc[2] / rows, acc[3] / rows, acc[4] / rows,
            (acc[0] - acc[1] - acc[2] - acc[3] - acc[4]) / rows);
    cudaFree(p.stats);
  }
  return 0;
}

}  // namespace aasist
