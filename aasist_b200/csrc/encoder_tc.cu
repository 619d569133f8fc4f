// Residual_block encoder on the 5th-generation tensor cores (reference
// models/RawNetGatSpoofST.py:225-278), fp16 hi/lo split operands, 3 products, fp32 accumulation
// in tensor memory:   a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo   (a = a_hi + a_lo, fp16 each).
//
// Activation layout in HBM ("phase-split pairs"):  act[b][h][phi][j][2*Cp] fp16, where the
// time position is w = 3*j + phi and the innermost vector is [hi(Cp) | lo(Cp)].  Splitting the
// time axis by w mod 3 makes the MaxPool2d((1,3)) lane-local: a CTA tile is 128 pooled
// columns j of one (utterance, row); the three pool phases s are three accumulators in TMEM and
// the epilogue takes max_s in registers.  A conv tap (dh,dw) for phase s reads input phase
// (s+dw-1) mod 3 shifted by floor((s+dw-1)/3) rows -- expressed as a +-128-byte offset of the
// UMMA shared-memory descriptor into ONE TMA-loaded tile of 130 rows, so each input tile is
// loaded once and feeds three accumulators.  Zero padding (conv pad (1,1)/(0,1)) is TMA
// out-of-bounds fill; positions >= W are kept zero in memory by every producer.
//
// Kernel structure (persistent, 1 CTA/SM, 320 threads): warp 0 = TMA producer, warp 1 = TMEM
// allocator + MMA issuer (warp-uniform schedule, one elected lane issues tcgen05.mma/commit),
// warps 2-9 = epilogue (two warps per TMEM lane quadrant; batched tcgen05.ld -> bias / SELU /
// residual / max-pool -> fp16-pair split -> vector stores).  smem ring of input tiles (full/empty
// mbarriers), double-buffered accumulators (tmem_full/tmem_empty mbarriers).
// This file serves the blocks that change the channel count (conv_downsample) and the 64-channel
// blocks; block 0 lives in block0_tc.cu and the 32->32 identity blocks in block_fused_tc.cu.
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "ptx.cuh"
#include "tc.cuh"

namespace aasist {

using namespace ptx;

constexpr int kTileJ = 128;                 // pooled columns per CTA tile (UMMA M)
constexpr int kBoxRows = kTileJ + 2;        // rows j0-1 .. j0+128
constexpr int kSlabBytes = 17 * 1024;       // 130 rows x 128 B, rounded up to the 1024-B swizzle atom
// epilogue warps: every warp owns 16 accumulator columns of one TMEM lane quadrant (COP/16 warps per quadrant)
constexpr int kMaxSlots = 8;
// utterances per encoder pass: as many as the batch has, up to 512 and up to 40 GB of scratch (≈ 59 MB per
// 4 s utterance) -- larger passes mean fewer launches and shorter tails (512 vs 256: +3 % at batch 512);
// AASIST_TC_CHUNK overrides
static int tc_chunk(size_t bytes_per_utt) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AASIST_TC_CHUNK"); v = e ? std::max(1, atoi(e)) : 0; }
  if (v > 0) return v;
  const size_t budget = (size_t)40 << 30;
  return (int)std::max<size_t>(1, std::min<size_t>(512, budget / std::max<size_t>(bytes_per_utt, 1)));
}

enum { TC_CONV1 = 0, TC_CONV2_ID = 1, TC_CONV2_DS = 2 };

struct ConvTcParams {
  int B, H_out, J, W_in, Co, Wo, Jn;
  int n_slots;
  // Work items: B * n_full regular strips (128 columns each), then n_tail packed tiles.  The last, partial strip
  // of a row ("tail": rows - 128*n_full columns) is short in the deeper blocks; when it fits a 64-row segment,
  // `segs` utterances' tails share one 128-row tile, each in a `seg_rows`-row segment (columns 128*n_full - 1 ..,
  // a multiple of 8 rows so the swizzle atoms line up).  A tail wider than one segment is cut into `pieces` equal
  // pieces of `piece_cols` columns, one segment each (block 4 at 4 s: 90 columns = 3 x 30, four segments per tile =
  // 0.75 tiles per utterance instead of one).  n_tail == 0: every strip is regular.
  int n_full, segs, seg_rows, pieces, piece_cols, n_tail, n_strips;
  const float* bias;       // [COP]
  __half* out;             // CONV1: [B][24][3][J][2*COP]; CONV2: [B][23][3][Jn][2*COP]
  float* out_f32;          // last block: (B,Co,23,Wo) fp32 NCHW (then `out` is unused)
  const __half* idn;       // CONV2_ID: block input, [B][23][3][J][2*COP]
  const uint8_t* wimg;     // pre-swizzled weight image (shared-memory layout)
  int wimg_bytes;
  long long* stats;        // optional: MMA-warp wait cycles per CTA [total, full(TMA), tempty(epilogue)]
  int products;            // 3 = a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (f16x3); 2 = without the weight correction (f16x2)
  int collector;           // A-operand collector reuse between the two a_hi products (tc.cuh collector_mask)
};

struct ConvTc {            // one convolution's packed device state
  uint8_t* wimg = nullptr;
  int wimg_bytes = 0;
  float* bias = nullptr;
};
struct BlockTc {
  int ci = 0, co = 0, cpi = 0, cop = 0;
  bool downsample = false;
  ConvTc c1, c2;
  std::vector<float> w1_host, wd_host;   // block 0 only: conv1 (bn2 folded) [6][32], conv_downsample [3][32]
  uint8_t* b0_img = nullptr; // block 0 only: conv2 + conv1/downsample K=16 operand images (block0_tc.cu)
};
struct TcState {
  BlockTc blocks[2][6];
  int sm_count = 148;
  void* encode_fn = nullptr;  // cuTensorMapEncodeTiled
  uint8_t* front_bimg = nullptr;  // sinc filter operand image (frontend_tc.cu)
  // second stream of an encoder pass (tc_encode): the two halves of a pass run on two streams so that the
  // partial last wave of one half's kernel is filled by the other half's kernels
  cudaStream_t st2[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
};

__device__ __forceinline__ float ex2_approx(float x) {   // one MUFU.EX2, flush-to-zero, no range branches
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float selu_scaled(float y) {
  // SELU of v given y = v*log2(e) (conv1 weights and bias are pre-scaled by log2(e) when they are packed):
  // scale*max(v,0) + min(0, scale*alpha*(2^y - 1)); the negative branch goes through MUFU.EX2:
  // |abs error| <= ~2.5e-7, the same order as the fp16-pair representation error
  const float e = ex2_approx(y);
  const float n = fminf(fmaf(e, kSeluScale * kSeluAlpha, -(kSeluScale * kSeluAlpha)), 0.f);
  return fmaf(fmaxf(y, 0.f), kSeluScale * 0.6931471805599453f, n);
}

// (a,b) fp32 -> packed hi half2 and lo half2 (a ~= hi+lo to 2^-22), saturating at the fp16 range inside the
// conversion instruction (ptx.cuh split2_sat)
template <bool LOWER_BOUNDED = false>
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  split2_sat(a, b, hi, lo);
}
// store 16 fp32 channels as hi (32 B) and lo (32 B) fp16 vectors
template <bool LOWER_BOUNDED = false>
__device__ __forceinline__ void store_pair16(__half* hi_dst, __half* lo_dst, const float (&v)[16]) {
  uint32_t hw[8], lw[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_pack2<LOWER_BOUNDED>(v[2 * i], v[2 * i + 1], hw[i], lw[i]);
  st_global_256(hi_dst, hw);
  st_global_256(lo_dst, lw);
}
__device__ __forceinline__ void store_pair32(__half* hi_dst, __half* lo_dst, const float (&v)[32]) {
  float a[16], b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = v[i]; b[i] = v[16 + i]; }
  store_pair16(hi_dst, lo_dst, a);
  store_pair16(hi_dst + 16, lo_dst + 16, b);
}
struct Pair16 { uint4 h[2], l[2]; };   // 16 channels of an activation: hi and lo fp16 vectors
__device__ __forceinline__ Pair16 load_pair16(const __half* hi_src, const __half* lo_src) {
  Pair16 p;
  uint32_t a[8], b[8];
#ifdef TC_IDN_ALLOC
  ld_global_nc_256(hi_src, a);
  ld_global_nc_256(lo_src, b);
#else
  ld_global_na_256(hi_src, a);   // no L1 allocation: the L1 data array is the shared memory the MMAs fetch from
  ld_global_na_256(lo_src, b);
#endif
  p.h[0] = make_uint4(a[0], a[1], a[2], a[3]); p.h[1] = make_uint4(a[4], a[5], a[6], a[7]);
  p.l[0] = make_uint4(b[0], b[1], b[2], b[3]); p.l[1] = make_uint4(b[4], b[5], b[6], b[7]);
  return p;
}
__device__ __forceinline__ Pair16 zero_pair16() {
  Pair16 p;
  p.h[0] = p.h[1] = p.l[0] = p.l[1] = make_uint4(0, 0, 0, 0);
  return p;
}
// v[i] += float(hi[i]) + float(lo[i])
__device__ __forceinline__ void add_pair16(const Pair16& p, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint32_t aw[4] = {p.h[i].x, p.h[i].y, p.h[i].z, p.h[i].w};
    const uint32_t bw[4] = {p.l[i].x, p.l[i].y, p.l[i].z, p.l[i].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[k]));
      const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&bw[k]));
      v[8 * i + 2 * k] += fa.x + fb.x;
      v[8 * i + 2 * k + 1] += fa.y + fb.y;
    }
  }
}

// ------------------------------------------------------------------------------------------
// implicit-GEMM (2,3)/(1,3) convolution on tcgen05, strip-mined
//
// Work item = one (utterance, 128-column strip); the CTA walks the input rows of the strip and
// every input tile (row r, phase phi) is brought into shared memory ONCE: it feeds the output
// row it is the upper tap of (dh=0) and the output row it is the lower tap of (dh=1).
// Two output rows are therefore in flight, one per TMEM accumulator buffer.  With >= 6 ring
// slots the three phase tiles of a row stay resident and the MMAs run in two passes (dh=1 for
// all phases -> completes an output row and hands it to the epilogue; then dh=0 -> starts the
// next one), which gives the epilogue half a row of slack to drain the buffer being recycled.
//
// Encoder block 0 and the 32 -> 32 identity blocks do not use this kernel: their conv1 -> conv2
// intermediate stays on chip (block0_tc.cu, block_fused_tc.cu).
// ------------------------------------------------------------------------------------------
template <int COP>
struct ConvCfg {
  static constexpr int kEpiWarps = COP / 4;                 // 8 (COP = 32) or 16 (COP = 64)
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  template <int CPI, int MODE>
  static constexpr bool tma_out() { return MODE == 0 /*TC_CONV1*/ && CPI == 32; }
  // staging for the TMA-store epilogue: 4 quadrants x 2 buffers x 32 rows x (4*COP) bytes
  template <int CPI, int MODE>
  static constexpr int stage_bytes() { return tma_out<CPI, MODE>() ? 4 * 2 * 32 * 4 * COP : 0; }
};

template <int CPI, int COP, int MODE>
__global__ void __launch_bounds__(ConvCfg<COP>::kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmS,
               const __grid_constant__ CUtensorMap tmAt, const __grid_constant__ CUtensorMap tmSt,
               const __grid_constant__ CUtensorMap tmO, const ConvTcParams p) {
  constexpr int SLABS = CPI / 32;                    // 128-byte-wide K slabs per input tile
  constexpr int SLOT_BYTES = SLABS * kSlabBytes;
  constexpr int KC = CPI / 16;                       // K chunks (UMMA_K = 16) per product
  constexpr int TAP_BYTES = SLABS * COP * 128;       // weight image bytes per tap (main input)
  constexpr int SIDE_TAP_BYTES = COP * 128;          // side input is 32 channels wide (1 slab)
  constexpr int TMEM_COLS = (6 * COP <= 256) ? 256 : 512;
  constexpr uint32_t IDESC = umma_idesc_f16(128, COP);
  constexpr bool HAS_SIDE = MODE == TC_CONV2_DS;
  constexpr int NCH = 1;                             // 16-column chunks per epilogue warp
  constexpr int kEpiWarps = ConvCfg<COP>::kEpiWarps;
  constexpr int R_IN = (MODE == TC_CONV1) ? 23 : 24; // input rows walked per strip
  constexpr int NTHREADS = ConvCfg<COP>::kThreads;
  // CONV1 of a 32-channel input (block 2: the store-heaviest kernel, 14.7 MB per utterance): the epilogue stages
  // a quadrant's 32 rows in shared memory (SWIZZLE_128B sub-tiles, conflict-free 16-byte stores) and ONE thread
  // hands them to the TMA (cp.async.bulk.tensor store) -- full 128-byte lines leave the SM without touching the
  // L1/LSU, which thread-per-row 32-byte st.global kept 87 % busy.  64-channel inputs have no room for the staging.
  constexpr bool TMA_OUT = ConvCfg<COP>::template tma_out<CPI, MODE>();
  constexpr int STAGE_SUB = COP / 32;                // 128-byte-wide sub-tiles of a staged row (hi, lo)
  constexpr int STAGE_BYTES = STAGE_SUB * 32 * 128;  // one quadrant, one buffer

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_w = smem;
  uint8_t* s_ring = smem + p.wimg_bytes;
  uint8_t* s_stage = s_ring + (size_t)p.n_slots * SLOT_BYTES;    // 1024-byte aligned: every term is a multiple of 1 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + ConvCfg<COP>::template stage_bytes<CPI, MODE>());
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxSlots;
  uint64_t* tfull = bars + 2 * kMaxSlots;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_bias = reinterpret_cast<float*>(tempty + 4);          // [COP] (registers are scarce in the epilogue)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_strips = p.n_strips;
  const bool two_pass = p.n_slots >= 6;
  if (threadIdx.x < COP) s_bias[threadIdx.x] = __ldg(p.bias + threadIdx.x);

  // weights: global image -> shared (generic proxy), then make visible to the async proxy
  for (int i = threadIdx.x; i < p.wimg_bytes / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(s_w)[i] = __ldg(reinterpret_cast<const uint4*>(p.wimg) + i);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_slots; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kEpiWarps);   // one arrival per epilogue warp
    }
    fence_barrier_init();
    prefetch_tensormap(&tmA);
    if (HAS_SIDE) prefetch_tensormap(&tmS);
    if (TMA_OUT) prefetch_tensormap(&tmO);
    if (p.n_tail) {
      prefetch_tensormap(&tmAt);
      if (HAS_SIDE) prefetch_tensormap(&tmSt);
    }
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_ptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      const int n_regular = p.B * p.n_full;
      for (int t = blockIdx.x; t < n_strips; t += gridDim.x) {
        const bool tail = t >= n_regular;                          // packed tile of `segs` tail pieces
        const int q0 = tail ? (t - n_regular) * p.segs : 0;        // first (utterance, piece) of a packed tile
        const int b = tail ? 0 : t / p.n_full;
        const int nseg = tail ? min(p.segs, p.B * p.pieces - q0) : 1;   // segments in this tile
        const int j0 = (tail ? p.n_full : t % p.n_full) * kTileJ - 1;
        const uint32_t box_bytes = (uint32_t)(tail ? p.seg_rows : kBoxRows) * 128u;
        const CUtensorMap* mA = tail ? &tmAt : &tmA;
        const CUtensorMap* mS = tail ? &tmSt : &tmS;
        for (int r = 0; r < R_IN; ++r) {
          for (int phi = 0; phi < 3; ++phi) {
            mbar_wait(&empty[slot], phase ^ 1);
            uint8_t* dst = s_ring + (size_t)slot * SLOT_BYTES;
            mbar_arrive_expect_tx(&full[slot], SLABS * nseg * box_bytes);
            for (int g = 0; g < nseg; ++g) {
              const int bg = tail ? (q0 + g) / p.pieces : b;
              const int jg = tail ? j0 + ((q0 + g) % p.pieces) * p.piece_cols : j0;
#pragma unroll
              for (int sl = 0; sl < SLABS; ++sl)
                tma_load_5d(dst + sl * kSlabBytes + (size_t)g * box_bytes, mA, &full[slot], sl * 64, jg, phi, r, bg);
            }
            if (++slot == p.n_slots) { slot = 0; phase ^= 1; }
          }
          if (HAS_SIDE && r < 23) {   // conv_downsample input: block input row h = r (output row r)
            for (int phi = 0; phi < 3; ++phi) {
              mbar_wait(&empty[slot], phase ^ 1);
              mbar_arrive_expect_tx(&full[slot], nseg * box_bytes);
              for (int g = 0; g < nseg; ++g) {
                const int bg = tail ? (q0 + g) / p.pieces : b;
                const int jg = tail ? j0 + ((q0 + g) % p.pieces) * p.piece_cols : j0;
                tma_load_5d(s_ring + (size_t)slot * SLOT_BYTES + (size_t)g * box_bytes, mS, &full[slot], 0, jg, phi, r, bg);
              }
              if (++slot == p.n_slots) { slot = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ==================================
    // the whole warp walks the (warp-uniform) schedule; one elected lane issues tcgen05 ops
    int slot = 0;
    uint32_t phase = 0;
    int nstart = 0;                       // output rows started so far (accumulator buffer = nstart & 1)
    const uint32_t w_base = smem_u32(s_w);
    const uint32_t ring_base = smem_u32(s_ring);
    const bool leader = elect_one();
    long long w_te = 0, w_fu = 0;
    const long long t_begin = AASIST_CLOCK();

    // All MMAs of one input tile (phase phi) for tap row dh into accumulator buffer `buf`.
    // Pool phase s is served by tap dw with (s + dw - 1) == phi (mod 3), from A rows shifted by
    // floor((s+dw-1)/3).  The MMA is operand-fetch bound (4 KB of A per instruction), so phases that
    // read the SAME A rows are merged into one wider-N MMA over contiguous B rows (taps stored 2,1,0):
    //   phi 1: s=0,1,2 <- dw=2,1,0, shift 0             -> one N = 3*COP group
    //   phi 0: s=0,1   <- dw=1,0,   shift 0 (N = 2*COP);  s=2 <- dw=2, shift +1 (N = COP)
    //   phi 2: s=1,2   <- dw=2,1,   shift 0 (N = 2*COP);  s=0 <- dw=0, shift -1 (N = COP)
    auto mma3 = [&](uint32_t d_tmem, uint32_t a_row, uint32_t w_row, int ntaps, bool fresh, bool side) {
      const uint32_t idesc = ntaps == 3 ? umma_idesc_f16(128, 3 * COP)
                                        : (ntaps == 2 ? umma_idesc_f16(128, 2 * COP) : umma_idesc_f16(128, COP));
      const bool one_slab = side || SLABS == 1;
      const uint64_t a_hi = umma_desc_sw128(a_row);
      const uint64_t a_lo = umma_desc_sw128(one_slab ? a_row + 64 : a_row + kSlabBytes);
      const uint64_t w_hi = umma_desc_sw128(w_row);
      const uint64_t w_lo = umma_desc_sw128(one_slab ? w_row + 64 : w_row + 3 * COP * 128);
      const int nkc = one_slab ? 2 : KC;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {               // +32 B per K chunk == +2 in the descriptor
        if (kc < nkc) {
          const uint32_t acc0 = (kc > 0 || !fresh) ? 1u : 0u;
          if (p.products == 3 && p.collector) {   // a_hi * w_hi, a_hi * w_lo back to back: A from the collector
            umma_f16_keep(d_tmem, a_hi + 2 * kc, w_hi + 2 * kc, idesc, acc0);
            umma_f16_reuse(d_tmem, a_hi + 2 * kc, w_lo + 2 * kc, idesc, 1);
            umma_f16(d_tmem, a_lo + 2 * kc, w_hi + 2 * kc, idesc, 1);
          } else {
            umma_f16(d_tmem, a_hi + 2 * kc, w_hi + 2 * kc, idesc, acc0);
            umma_f16(d_tmem, a_lo + 2 * kc, w_hi + 2 * kc, idesc, 1);
            if (p.products == 3) umma_f16(d_tmem, a_hi + 2 * kc, w_lo + 2 * kc, idesc, 1);
          }
        }
      }
    };
    auto issue_group = [&](int slot_i, int dh, int phi, int buf, bool fresh, bool side) {
      const uint32_t a_slot = ring_base + (uint32_t)slot_i * SLOT_BYTES;
      const uint32_t wb = side ? w_base + (uint32_t)(6 * TAP_BYTES) : w_base + (uint32_t)(dh * 3 * TAP_BYTES);
      const uint32_t d0 = tmem_base + (uint32_t)(buf * 3 * COP);
      constexpr uint32_t ROWS = COP * 128;            // bytes of one tap's rows
      if (phi == 1) {
        mma3(d0, a_slot + 128, wb, 3, fresh, side);
      } else if (phi == 0) {
        mma3(d0, a_slot + 128, wb + ROWS, 2, fresh, side);
        mma3(d0 + 2 * COP, a_slot + 256, wb, 1, fresh, side);
      } else {
        mma3(d0 + COP, a_slot + 128, wb, 2, fresh, side);
        mma3(d0, a_slot, wb + 2 * ROWS, 1, fresh, side);
      }
    };
    auto advance = [&](int& sl, uint32_t& ph) {
      if (++sl == p.n_slots) { sl = 0; ph ^= 1; }
    };
    auto begin_row = [&]() -> int {       // claim the next accumulator buffer (waits for its drain)
      const int buf = nstart & 1;
      AASIST_TIMED_WAIT(&tempty[buf], ((nstart >> 1) & 1) ^ 1, w_te);
      tc_fence_after_sync();
      ++nstart;
      return buf;
    };

    for (int t = blockIdx.x; t < n_strips; t += gridDim.x) {
      int buf_open = 0;                   // buffer of the output row that is currently half done
      for (int r = 0; r < R_IN; ++r) {
        // input row r: dh=1 completes output row o1, dh=0 starts output row o0
        const bool has_o1 = (MODE == TC_CONV1) ? true : (r >= 1);
        const bool has_o0 = (MODE == TC_CONV1) ? true : (r <= 22);
        const bool o1_fresh = (MODE == TC_CONV1) && r == 0;          // conv1 row 0 has no dh=0 tap (zero pad)
        const bool o0_closes = (MODE == TC_CONV1) && r == R_IN - 1;  // conv1 row 23 has no dh=1 tap
        int sl[3];
        uint32_t ph[3];
        {
          int s2 = slot;
          uint32_t p2 = phase;
          for (int i = 0; i < 3; ++i) { sl[i] = s2; ph[i] = p2; advance(s2, p2); }
        }
        if (two_pass) {
          if (o1_fresh) buf_open = begin_row();
          for (int phi = 0; phi < 3; ++phi) {
            AASIST_TIMED_WAIT(&full[sl[phi]], ph[phi], w_fu);
            tc_fence_after_sync();
            if (has_o1 && leader) issue_group(sl[phi], 1, phi, buf_open, o1_fresh && phi == 0, false);
            __syncwarp();
          }
          if (has_o1) {
            if (leader) umma_commit(&tfull[buf_open]);               // output row o1 complete
            __syncwarp();
          }
          if (has_o0) {
            buf_open = begin_row();
            for (int phi = 0; phi < 3; ++phi) {
              if (leader) {
                issue_group(sl[phi], 0, phi, buf_open, phi == 0, false);
                umma_commit(&empty[sl[phi]]);
              }
              __syncwarp();
            }
          } else {
            for (int phi = 0; phi < 3; ++phi) {
              if (leader) umma_commit(&empty[sl[phi]]);
              __syncwarp();
            }
          }
        } else {
          int buf_new = -1;
          if (o1_fresh) buf_open = begin_row();
          for (int phi = 0; phi < 3; ++phi) {
            AASIST_TIMED_WAIT(&full[sl[phi]], ph[phi], w_fu);
            tc_fence_after_sync();
            if (has_o1 && leader) issue_group(sl[phi], 1, phi, buf_open, o1_fresh && phi == 0, false);
            __syncwarp();
            if (has_o0) {
              if (phi == 0) buf_new = begin_row();
              if (leader) issue_group(sl[phi], 0, phi, buf_new, phi == 0, false);
            }
            if (leader) umma_commit(&empty[sl[phi]]);
            __syncwarp();
          }
          if (has_o1) {
            if (leader) umma_commit(&tfull[buf_open]);
            __syncwarp();
          }
          if (has_o0) buf_open = buf_new;
        }
        for (int i = 0; i < 3; ++i) advance(slot, phase);
        if (HAS_SIDE && has_o0) {
          for (int phi = 0; phi < 3; ++phi) {
            AASIST_TIMED_WAIT(&full[slot], phase, w_fu);
            tc_fence_after_sync();
            if (leader) {
              issue_group(slot, 0, phi, buf_open, false, true);
              umma_commit(&empty[slot]);
            }
            __syncwarp();
            advance(slot, phase);
          }
        }
        if (has_o0 && o0_closes) {
          if (leader) umma_commit(&tfull[buf_open]);
          __syncwarp();
        }
      }
    }
    if (p.stats && leader) {
      long long* stt = p.stats + (size_t)blockIdx.x * 4;
      stt[0] = AASIST_CLOCK() - t_begin; stt[1] = w_fu; stt[2] = w_te;
    }
  } else if (warp < 2 + kEpiWarps) {
    // =============================== epilogue (warps 2..9) ========================
    // warp -> (TMEM lane quadrant, column half): thread = one pooled column j, COP/2 channels
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;                 // which 16-column slice of the accumulators
    const int r = quad * 32 + lane;
    const int col0 = part * 16;
    const float* bias0 = s_bias + col0;
    int tcount = 0;
    int nstage = 0;                       // TMA-store epilogue: staging steps of this quadrant so far (buffer = nstage & 1)
    for (int t = blockIdx.x; t < n_strips; t += gridDim.x) {
      const bool tail = t >= p.B * p.n_full;
      int jt, b, j;
      if (!tail) {
        jt = t % p.n_full;
        b = t / p.n_full;
        j = jt * kTileJ + r;
      } else {                              // packed tile: accumulator row r = (utterance, column) of segment g
        const int g = r / p.seg_rows, lc = r - g * p.seg_rows;
        const int q = (t - p.B * p.n_full) * p.segs + g;           // (utterance, piece)
        jt = p.n_full;
        b = q / p.pieces;
        j = p.n_full * kTileJ + (q % p.pieces) * p.piece_cols + lc;
        // rows past the piece (the next piece's columns, halo rows) and empty segments: every guard below fails
        if (g >= p.segs || b >= p.B || lc >= p.piece_cols) { b = p.B - 1; j = 1 << 28; }
      }
      for (int h = 0; h < p.H_out; ++h, ++tcount) {
        const int buf = tcount & 1;
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 3 * COP + col0);
        if (MODE == TC_CONV1) {
          mbar_wait(&tfull[buf], (tcount >> 1) & 1);
          tc_fence_after_sync();
          // out[b][h][s][j][:] = selu(acc + b1), zero beyond the valid width
          uint32_t acc[3][NCH][16];
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int c = 0; c < NCH; ++c) tmem_ld16_async(t_row + (uint32_t)(s * COP + c * 16), acc[s][c]);
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int c = 0; c < NCH; ++c) tmem_ld_wait16(acc[s][c]);
          // the accumulators are in registers: release the TMEM buffer before the math / stores
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[buf]);
          // warp-uniform: all three phases of every row of this warp lie inside [0, W_in)
          const bool valid_all = !tail && 3 * (jt * kTileJ + quad * 32 + 31) + 2 < p.W_in;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const bool valid = valid_all || 3 * j + s < p.W_in;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              float v[16];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 bb = *reinterpret_cast<const float4*>(bias0 + c * 16 + 4 * q);
                v[4 * q] = selu_scaled(__uint_as_float(acc[s][c][4 * q]) + bb.x);
                v[4 * q + 1] = selu_scaled(__uint_as_float(acc[s][c][4 * q + 1]) + bb.y);
                v[4 * q + 2] = selu_scaled(__uint_as_float(acc[s][c][4 * q + 2]) + bb.z);
                v[4 * q + 3] = selu_scaled(__uint_as_float(acc[s][c][4 * q + 3]) + bb.w);
              }
              if (!valid) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
              }
              if (TMA_OUT && !tail) {
                // stage this thread's 32 B of hi and 32 B of lo in the quadrant's buffer, then one thread issues
                // the stores of the quadrant's 32 rows (rows >= J are clipped by the tensor map)
                uint32_t hw[8], lw[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) split_pack2<true>(v[2 * i], v[2 * i + 1], hw[i], lw[i]);
                uint8_t* sb = s_stage + (size_t)((quad * 2 + (nstage & 1)) * STAGE_BYTES);
                const uint32_t sw = (uint32_t)(lane & 7);
                // COP = 64: sub-tile 0 = hi (128 B rows), sub-tile 1 = lo;  COP = 32: one sub-tile, row = [hi | lo]
                uint8_t* rh = sb + lane * 128;
                uint8_t* rl = sb + (STAGE_SUB == 2 ? 32 * 128 : 0) + lane * 128;
                const uint32_t ch = (uint32_t)(col0 / 8), cl = (uint32_t)((STAGE_SUB == 2 ? 0 : COP / 8) + col0 / 8);
                *reinterpret_cast<uint4*>(rh + ((ch ^ sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(rh + (((ch + 1) ^ sw) << 4)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                *reinterpret_cast<uint4*>(rl + ((cl ^ sw) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                *reinterpret_cast<uint4*>(rl + (((cl + 1) ^ sw) << 4)) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                fence_proxy_async_smem();
                const bool issuer = part == 0 && lane == 0;
                // the OTHER buffer was handed to the TMA one step ago: its reads must be over before anybody
                // writes it in the next step, i.e. before this barrier releases
                if (issuer) tma_store_wait_read<0>();
                asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * (COP / 16)) : "memory");
                if (issuer) {
#pragma unroll
                  for (int u = 0; u < STAGE_SUB; ++u)
                    tma_store_5d(&tmO, sb + u * 32 * 128, u * 64, jt * kTileJ + quad * 32, s, h, b);
                  tma_store_commit();
                }
                ++nstage;
              } else if (j < p.J) {
                __half* o = p.out + ((((size_t)b * 24 + h) * 3 + s) * p.J + j) * (2 * COP) + col0 + c * 16;
                store_pair16<true>(o, o + COP, v);
              }
            }
          }
        } else {
          mbar_wait(&tfull[buf], (tcount >> 1) & 1);
          tc_fence_after_sync();
          const bool valid = j < p.Wo;
          uint32_t acc[3][NCH][16];
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int c = 0; c < NCH; ++c) tmem_ld16_async(t_row + (uint32_t)(s * COP + c * 16), acc[s][c]);
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int c = 0; c < NCH; ++c) tmem_ld_wait16(acc[s][c]);
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[buf]);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            float m[16];
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(acc[s][c][i]);
              if (MODE == TC_CONV2_ID) {
                // identity operand of this pool phase (L2-resident: conv1 of this block just read it);
                // loaded here rather than prefetched: 16 epilogue warps hide the latency and the
                // kernel stays inside its 96-register budget
                if (j < p.J) {
                  const __half* x = p.idn + ((((size_t)b * 23 + h) * 3 + s) * p.J + j) * (2 * COP) + col0 + c * 16;
                  add_pair16(load_pair16(x, x + COP), v);
                }
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) m[i] = s == 0 ? v[i] : fmaxf(m[i], v[i]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) m[i] = valid ? m[i] + bias0[c * 16 + i] : 0.f;
            if (p.out_f32) {
              if (valid)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int ch = col0 + c * 16 + i;
                  if (ch < p.Co) p.out_f32[(((size_t)b * p.Co + ch) * 23 + h) * p.Wo + j] = m[i];
                }
            } else if (j / 3 < p.Jn) {
              __half* o = p.out + ((((size_t)b * 23 + h) * 3 + (j % 3)) * p.Jn + j / 3) * (2 * COP) + col0 + c * 16;
              store_pair16(o, o + COP, m);
            }
          }
        }
      }
    }
  }
  if (TMA_OUT) tma_store_wait_read<0>();   // (only the issuing threads have groups pending) smem must outlive the reads
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// layout converters (stage entry points / tests only)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float clamp_h(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }
__global__ void pack_nchw_to_pairs_kernel(const float* __restrict__ in, __half* __restrict__ out, int C, int H,
                                          int W, int J, int Cp, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over (b,h,phi,j,c<Cp)
  if (i >= total) return;
  int c = i % Cp;
  size_t t = i / Cp;
  int j = t % J; t /= J;
  int phi = t % 3; t /= 3;
  int h = t % H;
  size_t b = t / H;
  int w = 3 * j + phi;
  float v = (c < C && w < W) ? in[((b * C + c) * H + h) * W + w] : 0.f;
  __half hi, lo;
  split_f16(clamp_h(v), hi, lo);
  __half* o = out + ((((b * H + h) * 3 + phi) * J + j) * (size_t)(2 * Cp));
  o[c] = hi;
  o[Cp + c] = lo;
}
__global__ void unpack_pairs_to_nchw_kernel(const __half* __restrict__ in, float* __restrict__ out, int C, int H,
                                            int W, int J, int Cp, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over (b,c,h,w)
  if (i >= total) return;
  int w = i % W;
  size_t t = i / W;
  int h = t % H; t /= H;
  int c = t % C;
  size_t b = t / C;
  const __half* p = in + ((((b * H + h) * 3 + (w % 3)) * J + w / 3) * (size_t)(2 * Cp));
  out[i] = __half2float(p[c]) + __half2float(p[Cp + c]);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_act_tmap(aasist_handle* h, CUtensorMap* m, const void* base, int Cp, int J, int H, int B,
                         int box_rows = kBoxRows) {
  EncodeTiledFn fn = (EncodeTiledFn)h->tc->encode_fn;
  cuuint64_t dims[5] = {(cuuint64_t)(2 * Cp), (cuuint64_t)J, 3, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[4];
  strides[0] = (cuuint64_t)2 * Cp * 2;
  strides[1] = strides[0] * J;
  strides[2] = strides[1] * 3;
  strides[3] = strides[2] * H;
  cuuint32_t box[5] = {64, (cuuint32_t)box_rows, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (Cp=%d J=%d H=%d B=%d)", (int)r, Cp, J, H, B);
    return AASIST_E_CUDA;
  }
  return 0;
}

// weight image: per tap, SLABS slabs of [cop rows][128 B], rows 128-byte swizzled like a TMA SW128 tile
//   cpi == 32: one slab, row = [w_hi(32) | w_lo(32)];  cpi == 64: slab 0 = w_hi(64), slab 1 = w_lo(64)
// Weight image.  Rows are 128-byte-swizzled like a TMA SW128 tile.  For each tap row dh the three
// column taps are stored in the order dw = 2, 1, 0 with `cop` rows each, so that the B rows of pool
// phases that share an A tile are contiguous and one wider-N tcgen05.mma serves them (see issue_group).
//   cpi == 32: row = [w_hi(32) | w_lo(32)];  block of dh = 3*cop rows
//   cpi == 64: block of dh = [hi: 3*cop rows of w_hi(64)] [lo: 3*cop rows of w_lo(64)]
static void put_tap(std::vector<uint8_t>& img, size_t base, int cpi, int cop, int dh, int dw, int n, int k, float w) {
  __half hi = __float2half_rn(w);
  __half lo = __float2half_rn(w - __half2float(hi));
  auto put = [&](size_t slab_off, int kk, __half v) {
    int byte_in_row = kk * 2;
    int chunk = byte_in_row / 16, within = byte_in_row % 16;
    size_t off = slab_off + (size_t)n * 128 + (size_t)((chunk ^ (n & 7)) * 16 + within);
    memcpy(&img[off], &v, 2);
  };
  const size_t tap = (size_t)(2 - dw) * cop * 128;
  if (cpi == 32) {
    const size_t b = base + (size_t)dh * 3 * cop * 128 + tap;
    put(b, k, hi);
    put(b, 32 + k, lo);
  } else {
    const size_t b = base + (size_t)dh * 6 * cop * 128 + tap;
    put(b, k, hi);
    put(b + (size_t)3 * cop * 128, k, lo);
  }
}

static int upload_bytes(uint8_t** dst, const std::vector<uint8_t>& v) {
  if (*dst) cudaFree(*dst);
  *dst = nullptr;
  AASIST_CUDA(cudaMalloc(dst, std::max<size_t>(v.size(), 16)));
  AASIST_CUDA(cudaMemcpy(*dst, v.data(), v.size(), cudaMemcpyHostToDevice));
  return 0;
}
static int upload_f(float** dst, const std::vector<float>& v) {
  if (*dst) cudaFree(*dst);
  *dst = nullptr;
  AASIST_CUDA(cudaMalloc(dst, sizeof(float) * std::max<size_t>(v.size(), 4)));
  AASIST_CUDA(cudaMemcpy(*dst, v.data(), sizeof(float) * v.size(), cudaMemcpyHostToDevice));
  return 0;
}

static inline int pad_ch(int c) { return c <= 32 ? 32 : 64; }
// blocks whose conv1 -> conv2 intermediate stays on chip in block_fused_tc.cu
static inline bool is_fused_block(const BlockTc& b) { return !b.downsample && b.cpi == 32 && b.cop == 32; }

static int pack_block_tc(aasist_handle* h, const std::string& pfx, int index, BlockTc& blk) {
  const int ci = h->cfg.enc_channels[index][0], co = h->cfg.enc_channels[index][1];
  blk.ci = ci;
  blk.co = co;
  blk.cpi = pad_ch(ci);
  blk.cop = pad_ch(co);
  blk.downsample = ci != co;
  auto P = [&](const std::string& n) -> const std::vector<float>& { return h->params.at(pfx + n); };
  const auto &w1 = P(".conv1.weight"), &b1 = P(".conv1.bias"), &w2 = P(".conv2.weight"), &b2 = P(".conv2.bias");
  const auto &g = P(".bn2.weight"), &be = P(".bn2.bias"), &mu = P(".bn2.running_mean"), &var = P(".bn2.running_var");
  std::vector<double> sc(co), sh(co);
  for (int o = 0; o < co; ++o) {
    sc[o] = (double)g[o] / sqrt((double)var[o] + kBnEps);
    sh[o] = (double)be[o] - (double)mu[o] * sc[o];
  }
  int rc;
  std::vector<float> bias1(blk.cop, 0.f), bias2(blk.cop, 0.f);
  for (int o = 0; o < co; ++o) {
    bias1[o] = (float)((double)b1[o] * sc[o] + sh[o]);
    bias2[o] = b2[o];
  }
  if (index == 0) {
    if (ci != 1 || blk.cop != 32) {
      set_error("f16x3 path: encoder block 0 must be 1 -> <=32 channels");
      return AASIST_E_INVALID;
    }
    std::vector<float> w(6 * 32, 0.f), wd(3 * 32, 0.f);
    for (int o = 0; o < co; ++o)
      for (int t = 0; t < 6; ++t) w[t * 32 + o] = (float)((double)w1[(size_t)o * 6 + t] * sc[o]);
    blk.w1_host = w;
    const auto &wdv = P(".conv_downsample.weight"), &bd = P(".conv_downsample.bias");
    for (int o = 0; o < co; ++o) {
      for (int t = 0; t < 3; ++t) wd[t * 32 + o] = wdv[(size_t)o * 3 + t];
      bias2[o] = (float)((double)b2[o] + (double)bd[o]);
    }
    blk.wd_host = wd;
  } else {
    const int tap_bytes = (blk.cpi / 32) * blk.cop * 128;
    std::vector<uint8_t> img((size_t)6 * tap_bytes, 0);
    // conv1's SELU is evaluated from y = v * log2(e) everywhere on this path (conv_tc_kernel's CONV1 epilogue,
    // the transformers of block_fused_tc.cu): conv1 and its bias are pre-scaled
    const double pre = 1.4426950408889634;
    for (int o = 0; o < co; ++o) {
      bias1[o] = (float)((double)bias1[o] * pre);
      for (int i = 0; i < ci; ++i)
        for (int t = 0; t < 6; ++t)
          put_tap(img, 0, blk.cpi, blk.cop, t / 3, t % 3, o, i,
                  (float)((double)w1[((size_t)o * ci + i) * 6 + t] * sc[o] * pre));
    }
    blk.c1.wimg_bytes = (int)img.size();
    if ((rc = upload_bytes(&blk.c1.wimg, img))) return rc;
  }
  if ((rc = upload_f(&blk.c1.bias, bias1))) return rc;
  {
    // conv2: K = cop (its input is this block's conv1 output)
    const int tap_bytes = (blk.cop / 32) * blk.cop * 128;
    const bool side = blk.downsample && index != 0;
    if (side && blk.cpi != 32) {
      set_error("f16x3 path: conv_downsample from a %d-channel input is not supported", ci);
      return AASIST_E_INVALID;
    }
    std::vector<uint8_t> img((size_t)6 * tap_bytes + (side ? 3 * blk.cop * 128 : 0), 0);
    for (int o = 0; o < co; ++o)
      for (int i = 0; i < co; ++i)
        for (int t = 0; t < 6; ++t)
          put_tap(img, 0, blk.cop, blk.cop, t / 3, t % 3, o, i, w2[((size_t)o * co + i) * 6 + t]);
    if (side) {
      const auto &wdv = P(".conv_downsample.weight"), &bd = P(".conv_downsample.bias");
      for (int o = 0; o < co; ++o) {
        for (int i = 0; i < ci; ++i)
          for (int t = 0; t < 3; ++t)
            put_tap(img, (size_t)6 * tap_bytes, 32, blk.cop, 0, t, o, i, wdv[((size_t)o * ci + i) * 3 + t]);
        bias2[o] = (float)((double)b2[o] + (double)bd[o]);
      }
    }
    blk.c2.wimg_bytes = (int)img.size();
    if ((rc = upload_bytes(&blk.c2.wimg, img))) return rc;
    if (index == 0) {
      std::vector<uint8_t> img0 = img;            // conv2 taps first, then the small K=16 operands
      block0_pack_small(img0, blk.w1_host, blk.wd_host, bias1, co);
      if ((rc = upload_bytes(&blk.b0_img, img0))) return rc;
    }
  }
  if ((rc = upload_f(&blk.c2.bias, bias2))) return rc;
  return 0;
}

int tc_finalize(aasist_handle* h) {
  if (!h->tc) h->tc = new TcState();
  TcState* tc = h->tc;
  cudaDeviceProp prop;
  AASIST_CUDA(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10) {
    set_error("the f16x3 path needs an sm_100a GPU (tcgen05); device is sm_%d%d", prop.major, prop.minor);
    return AASIST_E_CUDA;
  }
  tc->sm_count = prop.multiProcessorCount;
  cudaDriverEntryPointQueryResult q;
  AASIST_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &tc->encode_fn, cudaEnableDefault, &q));
  if (!tc->encode_fn || q != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return AASIST_E_CUDA;
  }
  const char* enc_names[2] = {h->cfg.kind == AASIST_KIND_AASIST ? "encoder" : "encoder_T", "encoder_S"};
  // the Res2Net encoder runs on fp32 kernels: only the sinc front end uses the tensor cores there
  for (int e = 0; e < (h->cfg.encoder == AASIST_ENC_RESIDUAL23 ? h->n_encoders : 0); ++e)
    for (int i = 0; i < 6; ++i) {
      std::string p = std::string(enc_names[e]) + "." + std::to_string(i) + ".0";
      int rc = pack_block_tc(h, p, i, tc->blocks[e][i]);
      if (rc) return rc;
    }
  return tc_front_finalize(h, &tc->front_bimg);
}

void tc_destroy(aasist_handle* h) {
  if (!h->tc) return;
  for (int e = 0; e < 2; ++e)
    for (int i = 0; i < 6; ++i) {
      BlockTc& b = h->tc->blocks[e][i];
      cudaFree(b.c1.wimg); cudaFree(b.c1.bias); cudaFree(b.c2.wimg); cudaFree(b.c2.bias);
      cudaFree(b.b0_img);
    }
  cudaFree(h->tc->front_bimg);
  for (int k = 0; k < 3; ++k)
    if (h->tc->st2[k]) { cudaStreamDestroy(h->tc->st2[k]); cudaEventDestroy(h->tc->ev_join[k]); }
  if (h->tc->ev_fork) cudaEventDestroy(h->tc->ev_fork);
  delete h->tc;
  h->tc = nullptr;
}

struct TcPlan {
  int W[7], J[7];         // W[i], J[i] = ceil(W[i]/3): width / phase length of block i's input
  size_t z, mid, act;     // bytes per utterance
};
static void make_tc_plan(const aasist_handle* h, int L, TcPlan& pl) {
  pl.W[0] = (L - h->taps + 1) / 3;
  for (int i = 0; i < 6; ++i) pl.W[i + 1] = pl.W[i] / 3;
  for (int i = 0; i < 7; ++i) pl.J[i] = (pl.W[i] + 2) / 3;
  pl.z = sizeof(float) * (size_t)kSpecNodes * pl.W[0];
  pl.mid = pl.act = 0;
  for (int i = 0; i < 6; ++i) {
    const int ci = h->cfg.enc_channels[i][0], co = h->cfg.enc_channels[i][1];
    const int cop = pad_ch(co);
    // blocks whose conv1 -> conv2 intermediate stays on chip need no `mid` buffer:
    // block 0 (block0_tc.cu) and 32->32 identity blocks (block_fused_tc.cu)
    const bool fused = i == 0 || (ci == co && pad_ch(ci) == 32 && cop == 32);
    if (!fused) pl.mid = std::max(pl.mid, (size_t)24 * 3 * pl.J[i] * 2 * cop * 2);
    pl.act = std::max(pl.act, (size_t)23 * 3 * pl.J[i + 1] * 2 * cop * 2);
  }
  pl.mid = std::max<size_t>(pl.mid, 256);
}
static inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t tc_workspace_bytes(const aasist_handle* h, int B, int L) {
  TcPlan pl;
  make_tc_plan(h, L, pl);
  size_t nb = std::min(B, tc_chunk(pl.z + pl.mid + 2 * pl.act));
  return al256(pl.z * nb) + al256(pl.mid * nb) + 2 * al256(pl.act * nb) + 1024;
}

static const char* kConvNames[2][6] = {
    {"enc0.conv1", "enc1.conv1_tc", "enc2.conv1_tc", "enc3.conv1_tc", "enc4.conv1_tc", "enc5.conv1_tc"},
    {"enc0.conv2_tc", "enc1.conv2_tc", "enc2.conv2_tc", "enc3.conv2_tc", "enc4.conv2_tc", "enc5.conv2_tc"}};

template <int CPI, int COP, int MODE>
static int launch_conv(aasist_handle* h, const char* name, const CUtensorMap& tmA, const CUtensorMap& tmS,
                       const CUtensorMap& tmAt, const CUtensorMap& tmSt, const CUtensorMap& tmO, ConvTcParams p,
                       cudaStream_t st) {
  constexpr int SLOT = (CPI / 32) * kSlabBytes;
  constexpr int STAGE = ConvCfg<COP>::template stage_bytes<CPI, MODE>();
  const int budget = 227 * 1024 - 1024 /*align*/ - p.wimg_bytes - STAGE - 512 /*barriers, bias*/;
  int n_slots = std::min(kMaxSlots, budget / SLOT);
  if (n_slots < 2) {
    set_error("conv_tc: not enough shared memory for the input ring (weights %d bytes)", p.wimg_bytes);
    return AASIST_E_INVALID;
  }
  {
    static int slots_override = -1;
    if (slots_override < 0) { const char* e = getenv("AASIST_TC_SLOTS"); slots_override = e ? atoi(e) : 0; }
    if (slots_override >= 2 && slots_override <= n_slots) n_slots = slots_override;
  }
  p.n_slots = n_slots;
  size_t smem = 1024 + (size_t)p.wimg_bytes + (size_t)n_slots * SLOT + STAGE + 512;
  auto kern = conv_tc_kernel<CPI, COP, MODE>;
  AASIST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_strips = p.B * p.n_full + p.n_tail;
  p.n_strips = n_strips;
  const int grid = std::min(n_strips, h->tc->sm_count);
  static int want_stats = -1;
  if (want_stats < 0) { const char* e = getenv("AASIST_TC_STATS"); want_stats = e ? atoi(e) : 0; }
#ifndef AASIST_KERNEL_STATS
  want_stats = 0;   // the instrumentation is compiled in only by tools/variant_build.sh -DAASIST_KERNEL_STATS
#endif
  p.stats = nullptr;
  p.products = h->cfg.precision == AASIST_PREC_F16X2 ? 2 : 3;
  p.collector = (collector_mask() >> (CPI == 32 ? 3 : 4)) & 1;
  if (want_stats) {
    AASIST_CUDA(cudaMalloc(&p.stats, sizeof(long long) * 4 * grid));
    AASIST_CUDA(cudaMemset(p.stats, 0, sizeof(long long) * 4 * grid));
  }
  {
    LaunchSpan span(h, name, st);
    kern<<<grid, ConvCfg<COP>::kThreads, smem, st>>>(tmA, tmS, tmAt, tmSt, tmO, p);
  }
  AASIST_CUDA(cudaGetLastError());
  if (want_stats) {   // debugging aid: where the MMA warp waits (cycles per output row-tile, mean over CTAs)
    std::vector<long long> hst((size_t)4 * grid);
    AASIST_CUDA(cudaStreamSynchronize(st));
    AASIST_CUDA(cudaMemcpy(hst.data(), p.stats, sizeof(long long) * hst.size(), cudaMemcpyDeviceToHost));
    double acc[3] = {0, 0, 0};
    for (int c = 0; c < grid; ++c)
      for (int k = 0; k < 3; ++k) acc[k] += (double)hst[(size_t)c * 4 + k] / grid;
    const double rows = (double)n_strips * p.H_out / grid;
    fprintf(stderr, "[%s stats] per row-tile cycles: total %.0f | wait full(TMA) %.0f tempty(epilogue) %.0f | issuing %.0f"
            " | slots %d strips/CTA %.2f\n", name, acc[0] / rows, acc[1] / rows, acc[2] / rows,
            (acc[0] - acc[1] - acc[2]) / rows, p.n_slots, (double)n_strips / grid);
    cudaFree(p.stats);
  }
  return 0;
}

// one residual block on the tensor-core path.
//   in_pairs: block input in pair layout [nb][23][3][J][2*cpi]   (index >= 1)
//   z:        block-0 input (nb,23,W) fp32                       (index == 0)
//   out_pairs [nb][23][3][Jn][2*cop]  or  out_f32 (nb,co,23,Wo)
static int run_block_tc(aasist_handle* h, int enc, int index, const __half* in_pairs, const float* z, int nb,
                        int W, __half* mid, __half* out_pairs, float* out_f32, cudaStream_t st) {
  const BlockTc& blk = h->tc->blocks[enc][index];
  const int J = (W + 2) / 3, Wo = W / 3, Jn = (Wo + 2) / 3;
  if (Wo < 1) {
    set_error("encoder block %d input width %d < 3", index, W);
    return AASIST_E_INVALID;
  }
  int rc;
  // tile plans: columns a row's strips must cover (conv2 also writes the zero rows up to 3*Jn of the next block's
  // layout); regular 128-column strips, and a packed tile for the tails when a tail fits a 64-row segment
  const int rows1 = J, rows2 = std::max(J, std::min(3 * Jn, Wo + 2));
  struct StripPlan { int n_full, segs, seg_rows, pieces, piece_cols, n_tail; };
  auto plan_strips = [nb](int rows) {
    StripPlan sp;
    sp.n_full = rows / kTileJ;
    const int rem = rows - sp.n_full * kTileJ;
    sp.segs = 1; sp.seg_rows = kTileJ; sp.pieces = 1; sp.piece_cols = kTileJ; sp.n_tail = 0;
    if (rem > 0) {
      // cut the tail into 1..4 equal pieces of one segment each; keep the cut that needs the fewest tiles per
      // utterance, if it beats a regular strip (1 tile)
      int best_num = 1, best_den = 1, best_p = 0, best_rows = 0;
      for (int pc = 1; pc <= 4; ++pc) {
        const int cols = (rem + pc - 1) / pc;
        const int seg_rows = ((cols + 2 + 7) / 8) * 8;
        if (seg_rows > 64) continue;
        const int segs = kTileJ / seg_rows;
        if (pc * best_den < best_num * segs) { best_num = pc; best_den = segs; best_p = pc; best_rows = seg_rows; }
      }
      if (best_p) {
        sp.pieces = best_p;
        sp.piece_cols = (rem + best_p - 1) / best_p;
        sp.seg_rows = best_rows;
        sp.segs = kTileJ / best_rows;
        sp.n_tail = (nb * sp.pieces + sp.segs - 1) / sp.segs;
      } else {
        sp.n_full += 1;                                  // the tail is a strip like the others
      }
    }
    return sp;
  };
  const StripPlan sp1 = plan_strips(rows1), sp2 = plan_strips(rows2);
  CUtensorMap tmIn, tmMid, tmInT1, tmMidT, tmInT2, tmMidO;
  memset(&tmMidO, 0, sizeof(tmMidO));
  memset(&tmIn, 0, sizeof(tmIn));
  memset(&tmMid, 0, sizeof(tmMid));
  memset(&tmInT1, 0, sizeof(tmInT1));
  memset(&tmMidT, 0, sizeof(tmMidT));
  memset(&tmInT2, 0, sizeof(tmInT2));
  const bool fused_path = index == 0 || is_fused_block(blk);
  if (index > 0 && (rc = make_act_tmap(h, &tmMid, mid, blk.cop, J, 24, nb))) return rc;
  if (index > 0 && (rc = make_act_tmap(h, &tmIn, in_pairs, blk.cpi, J, 23, nb))) return rc;
  if (!fused_path) {
    if ((rc = make_act_tmap(h, &tmMidO, mid, blk.cop, J, 24, nb, 32))) return rc;   // conv1's TMA stores: 32-row boxes
    if ((rc = make_act_tmap(h, &tmInT1, in_pairs, blk.cpi, J, 23, nb, sp1.seg_rows))) return rc;
    if ((rc = make_act_tmap(h, &tmMidT, mid, blk.cop, J, 24, nb, sp2.seg_rows))) return rc;
    if ((rc = make_act_tmap(h, &tmInT2, in_pairs, blk.cpi, J, 23, nb, sp2.seg_rows))) return rc;
  }
  // ---- block 0 and 32->32 identity blocks: the whole block is one kernel, intermediate kept on chip ----
  if (index == 0)
    return launch_block0_tc(h, h->tc->sm_count, blk.b0_img, blk.c1.bias, blk.c2.bias, z, nb, W, out_pairs, st);
  if (is_fused_block(blk)) {
    static const char* names[6] = {"", "enc1.fused_tc", "enc2.fused_tc", "enc3.fused_tc", "enc4.fused_tc", "enc5.fused_tc"};
    return launch_block_fused_tc(h, h->tc->sm_count, names[index], tmIn, blk.c1.wimg, blk.c2.wimg, blk.c1.bias,
                                 blk.c2.bias, in_pairs, nb, W, blk.co, out_pairs, out_f32, st);
  }
  // ---- conv1 ----
  {
    ConvTcParams p;
    memset(&p, 0, sizeof(p));
    p.B = nb; p.H_out = 24; p.J = J; p.W_in = W; p.Co = blk.co;
    p.n_full = sp1.n_full; p.segs = sp1.segs; p.seg_rows = sp1.seg_rows; p.pieces = sp1.pieces;
    p.piece_cols = sp1.piece_cols; p.n_tail = sp1.n_tail;
    p.bias = blk.c1.bias; p.out = mid; p.wimg = blk.c1.wimg; p.wimg_bytes = blk.c1.wimg_bytes;
    if (blk.cpi == 32 && blk.cop == 32) rc = launch_conv<32, 32, TC_CONV1>(h, kConvNames[0][index], tmIn, tmIn, tmInT1, tmInT1, tmMidO, p, st);   // 32 -> 24 (AASIST-L)
    else if (blk.cpi == 32 && blk.cop == 64) rc = launch_conv<32, 64, TC_CONV1>(h, kConvNames[0][index], tmIn, tmIn, tmInT1, tmInT1, tmMidO, p, st);
    else if (blk.cpi == 64 && blk.cop == 64) rc = launch_conv<64, 64, TC_CONV1>(h, kConvNames[0][index], tmIn, tmIn, tmInT1, tmInT1, tmMidO, p, st);
    else { set_error("f16x3 path: unsupported conv1 shape %d->%d", blk.ci, blk.co); rc = AASIST_E_INVALID; }
    if (rc) return rc;
  }
  AASIST_CUDA(cudaGetLastError());
  // ---- conv2 (+ identity / downsample) + max-pool ----
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.B = nb; p.H_out = 23; p.J = J; p.W_in = W; p.Co = blk.co; p.Wo = Wo; p.Jn = Jn;
  p.n_full = sp2.n_full; p.segs = sp2.segs; p.seg_rows = sp2.seg_rows; p.pieces = sp2.pieces;
  p.piece_cols = sp2.piece_cols; p.n_tail = sp2.n_tail;
  p.bias = blk.c2.bias; p.out = out_pairs; p.out_f32 = out_f32; p.wimg = blk.c2.wimg; p.wimg_bytes = blk.c2.wimg_bytes;
  if (!blk.downsample) {
    p.idn = in_pairs;   // 64 -> 64 identity block (the 32 -> 32 ones took the fused path above)
    rc = launch_conv<64, 64, TC_CONV2_ID>(h, kConvNames[1][index], tmMid, tmMid, tmMidT, tmMidT, tmMidO, p, st);
  } else {
    if (blk.cop == 32) rc = launch_conv<32, 32, TC_CONV2_DS>(h, kConvNames[1][index], tmMid, tmIn, tmMidT, tmInT2, tmMidO, p, st);
    else rc = launch_conv<64, 64, TC_CONV2_DS>(h, kConvNames[1][index], tmMid, tmIn, tmMidT, tmInT2, tmMidO, p, st);
  }
  return rc;
}

int tc_encode(aasist_handle* h, const float* x, int B, int L, float** enc_out, void* ws, int mask_start,
              int mask_count, cudaStream_t st) {
  const uint8_t* bimg = h->tc->front_bimg;
  if (mask_count > 0) {
    int rc = tc_front_mask(h, bimg, mask_start, mask_count, st, &bimg);
    if (rc) return rc;
  }
  TcPlan pl;
  make_tc_plan(h, L, pl);
  const int nbmax = std::min(B, tc_chunk(pl.z + pl.mid + 2 * pl.act));
  char* w = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  float* z = (float*)w;
  w += al256(pl.z * nbmax);
  __half* mid = (__half*)w;
  w += al256(pl.mid * nbmax);
  __half* act[2];
  act[0] = (__half*)w;
  w += al256(pl.act * nbmax);
  act[1] = (__half*)w;
  const int C = h->cfg.enc_channels[5][1];
  const size_t enc_per = (size_t)C * kSpecNodes * pl.W[6];
  static int f32_front = -1;   // AASIST_TC_F32_FRONT=1: timing experiments with the CUDA-core sinc stage
  if (f32_front < 0) { const char* e = getenv("AASIST_TC_F32_FRONT"); f32_front = e ? atoi(e) : 0; }
  // The two halves of a pass run on two streams: every kernel here is persistent with one CTA per SM, so the
  // partial last wave of one half's kernel (blocks 3-5: 86-92 % wave efficiency at 512 x 4 s) is filled by the other
  // half's kernels instead of idling (+1.6 % on the step).  AASIST_TC_STREAMS=1 runs the pass on one stream, 3 / 4
  // split it further.
  static int n_streams = -1;
  if (n_streams < 0) { const char* e = getenv("AASIST_TC_STREAMS"); n_streams = e ? std::min(4, std::max(1, atoi(e))) : 2; }
  // utterances [u0, u0+n) of the pass that starts at b0, in the slice of every scratch buffer that starts at
  // utterance u0 of the pass
  auto run_range = [&](int b0, int u0, int n, cudaStream_t s) -> int {
    float* zz = (float*)((char*)z + pl.z * u0);
    __half* mm = (__half*)((char*)mid + pl.mid * u0);
    __half* aa[2] = {(__half*)((char*)act[0] + pl.act * u0), (__half*)((char*)act[1] + pl.act * u0)};
    const float* xx = x + (size_t)(b0 + u0) * L;
    int rc = f32_front ? launch_frontend_f32(h, xx, n, L, zz, mask_start, mask_count, s)
                       : launch_frontend_tc(h, bimg, h->tc->sm_count, xx, n, L, zz, s);
    if (rc) return rc;
    for (int e = 0; e < h->n_encoders; ++e) {
      const __half* in = nullptr;
      for (int i = 0; i < 6; ++i) {
        __half* out = aa[i & 1];
        float* of32 = i == 5 ? enc_out[e] + (size_t)(b0 + u0) * enc_per : nullptr;
        if ((rc = run_block_tc(h, e, i, in, zz, n, pl.W[i], mm, out, of32, s))) return rc;
        in = out;
      }
    }
    return 0;
  };
  for (int b0 = 0; b0 < B; b0 += nbmax) {
    const int nb = std::min(nbmax, B - b0);
    int rc;
    // per-kernel profiling wants kernels that do not overlap; tiny passes gain nothing
    if (n_streams > 1 && !h->profiling && nb >= 32 * n_streams && (pl.mid % 16 == 0) && (pl.act % 16 == 0)) {   // (TMA bases: 16-byte aligned)
      TcState* tc = h->tc;
      if (!tc->ev_fork) AASIST_CUDA(cudaEventCreateWithFlags(&tc->ev_fork, cudaEventDisableTiming));
      for (int k = 0; k < n_streams - 1; ++k)
        if (!tc->st2[k]) {
          AASIST_CUDA(cudaStreamCreateWithFlags(&tc->st2[k], cudaStreamNonBlocking));
          AASIST_CUDA(cudaEventCreateWithFlags(&tc->ev_join[k], cudaEventDisableTiming));
        }
      AASIST_CUDA(cudaEventRecord(tc->ev_fork, st));
      int u0 = 0;
      for (int k = 0; k < n_streams; ++k) {
        const int n = (nb - u0) / (n_streams - k);
        cudaStream_t s = k == 0 ? st : tc->st2[k - 1];
        if (k > 0) AASIST_CUDA(cudaStreamWaitEvent(s, tc->ev_fork, 0));
        if ((rc = run_range(b0, u0, n, s))) return rc;
        if (k > 0) AASIST_CUDA(cudaEventRecord(tc->ev_join[k - 1], s));
        u0 += n;
      }
      for (int k = 0; k < n_streams - 1; ++k) AASIST_CUDA(cudaStreamWaitEvent(st, tc->ev_join[k], 0));
    } else if ((rc = run_range(b0, 0, nb, st))) {
      return rc;
    }
  }
  return 0;
}

int tc_frontend_to_f32(aasist_handle* h, const float* x, int B, int L, float* out, int mask_start, int mask_count,
                       cudaStream_t st) {
  const uint8_t* bimg = h->tc->front_bimg;
  if (mask_count > 0) {
    int rc = tc_front_mask(h, bimg, mask_start, mask_count, st, &bimg);
    if (rc) return rc;
  }
  return launch_frontend_tc(h, bimg, h->tc->sm_count, x, B, L, out, st);
}

int tc_block_f32io(aasist_handle* h, int enc, int index, const float* in, int B, int W, float* out, void* ws,
                   int64_t ws_bytes, cudaStream_t st) {
  const BlockTc& blk = h->tc->blocks[enc][index];
  const int J = (W + 2) / 3, Wo = W / 3, Jn = (Wo + 2) / 3;
  size_t in_b = al256((size_t)B * 23 * 3 * J * 2 * blk.cpi * 2), mid_b = al256((size_t)B * 24 * 3 * J * 2 * blk.cop * 2),
         out_b = al256((size_t)B * 23 * 3 * Jn * 2 * blk.cop * 2);
  if (!ws || (size_t)ws_bytes < in_b + mid_b + out_b + 1024) {
    set_error("aasist_encoder_block (f16x3): workspace needs %zu bytes", in_b + mid_b + out_b + 1024);
    return AASIST_E_WORKSPACE;
  }
  char* w = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  __half* pin = (__half*)w;
  __half* mid = (__half*)(w + in_b);
  __half* pout = (__half*)(w + in_b + mid_b);
  if (index > 0) {
    size_t total = (size_t)B * 23 * 3 * J * blk.cpi;
    LaunchSpan span(h, "pack_pairs", st);
    pack_nchw_to_pairs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, pin, blk.ci, 23, W, J, blk.cpi, total);
  }
  int rc = run_block_tc(h, enc, index, pin, in, B, W, mid, pout, nullptr, st);
  if (rc) return rc;
  size_t total = (size_t)B * blk.co * 23 * Wo;
  {
    LaunchSpan span(h, "unpack_pairs", st);
    unpack_pairs_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pout, out, blk.co, 23, Wo, Jn, blk.cop, total);
  }
  AASIST_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace aasist
