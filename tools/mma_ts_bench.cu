// Microbenchmark + layout check: tcgen05.mma (kind::f16, M=128, K=16, cta_group::1) with the A operand in TENSOR
// MEMORY ("TS" form) instead of shared memory.  (1) correctness: A written by four warps with tcgen05.st 32x32b
// (lane = row, 32-bit column c = K elements 2c, 2c+1), B K-major SWIZZLE_128B in shared memory, D compared with the
// host; (2) cycles per MMA vs N for A-in-TMEM against A-in-smem.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_ts_bench tools/mma_ts_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include "../aasist_b200/csrc/ptx.cuh"
using namespace aasist::ptx;

__global__ void __launch_bounds__(128, 1) ts_check(const __half* A /*[128][16]*/, const __half* B /*[64][16]*/,
                                                    float* D /*[128][64]*/) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;                                  // 64 rows x 128 B, SW128
  uint64_t* bar = (uint64_t*)(sB + 64 * 128);
  uint32_t* tptr = (uint32_t*)(bar + 1);
  for (int i = threadIdx.x; i < 64 * 128 / 16; i += 128) ((uint4*)sB)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 16; i += 128) {
    const int n = i / 16, k = i % 16;
    const int chunk = (k / 8) ^ (n & 7);
    *reinterpret_cast<__half*>(sB + n * 128 + chunk * 16 + (k % 8) * 2) = B[i];
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<128>(tptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = *tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = threadIdx.x;
  // A operand at columns [64, 72): lane = row m, column c holds (A[m][2c], A[m][2c+1])
  uint32_t w[8];
  for (int c = 0; c < 8; ++c) {
    const __half2 h = __halves2half2(A[m * 16 + 2 * c], A[m * 16 + 2 * c + 1]);
    w[c] = *reinterpret_cast<const uint32_t*>(&h);
  }
  tmem_st8(tm + ((uint32_t)(warp * 32) << 16) + 64u, w);
  tmem_st_wait();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x == 0) {
    umma_f16_ts(tm, tm + 64u, umma_desc_sw128(smem_u32(sB)), umma_idesc_f16(128, 64), 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after_sync();
  float v[32];
  for (int half = 0; half < 2; ++half) {
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 32), v);
    for (int i = 0; i < 32; ++i) D[m * 64 + half * 32 + i] = v[i];
  }
  (void)lane;
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc<128>(tm); }
}

// TS: A at TMEM columns (4 different operands), B smem.  SS: the mma_bench mode-1 pattern.
template <int N, int TS>
__global__ void __launch_bounds__(64, 1) mma_bench(int trips, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 4 * 17408;
  uint64_t* bar = (uint64_t*)(sB + 256 * 128);
  uint32_t* tptr = (uint32_t*)(bar + 1);
  for (int i = threadIdx.x; i < (4 * 17408 + 256 * 128) / 16; i += 64) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(tptr);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = *tptr;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    uint64_t ad[4], bd[2];
    for (int i = 0; i < 4; ++i) ad[i] = umma_desc_sw128(a0 + i * 17408 + 128 * (i % 3));
    bd[0] = umma_desc_sw128(b0); bd[1] = umma_desc_sw128(b0 + 64);
    long long t0 = clock64();
    for (int t = 0; t < trips; ++t) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t d = tm + (uint32_t)((i % 2) * N > 256 - N ? 0 : (i % 2) * N);
        if (TS) {
          const uint32_t at = tm + 384u + (uint32_t)(i * 32);       // [hi kc0 | lo kc0 | hi kc1 | lo kc1] x 8 columns
          umma_f16_ts(d, at, bd[0], idesc, 1);
          umma_f16_ts(d, at + 8, bd[0], idesc, 1);
          umma_f16_ts(d, at, bd[1], idesc, 1);
        } else {
          umma_f16(d, ad[i], bd[0], idesc, 1);
          umma_f16(d, ad[i] + 4, bd[0], idesc, 1);
          umma_f16(d, ad[i], bd[1], idesc, 1);
        }
      }
    }
    long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc<512>(tm); }
}

template <int N, int TS> void run(long long* d, size_t smem) {
  const int trips = 400;
  cudaFuncSetAttribute(mma_bench<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mma_bench<N, TS><<<148, 64, smem>>>(trips, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  double per = (double)h[1] / (trips * 12);
  printf("N=%3d A in %s | issue %6.1f  total %6.1f cyc/mma | %5.0f MAC/clk\n", N, TS ? "TMEM" : "smem",
         (double)h[0] / (trips * 12), per, 128.0 * N * 16 / per);
}

int main() {
  // ---- layout check
  std::vector<__half> A(128 * 16), B(64 * 16);
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 16; ++k) A[m * 16 + k] = __float2half((float)((m * 7 + k * 3) % 11 - 5));
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 16; ++k) B[n * 16 + k] = __float2half((float)((n * 5 + k * 2) % 9 - 4));
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(ts_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  ts_check<<<1, 128, 16384>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ts_check error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> D(128 * 64);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      float r = 0.f;
      for (int k = 0; k < 16; ++k) r += __half2float(A[m * 16 + k]) * __half2float(B[n * 16 + k]);
      if (r != D[m * 64 + n] && bad++ < 5) printf("mismatch m=%d n=%d ref=%g got=%g\n", m, n, r, D[m * 64 + n]);
    }
  printf("TS layout check: %d mismatches of %d\n", bad, 128 * 64);
  // ---- timing
  long long* d; cudaMalloc(&d, 16);
  size_t smem = 1024 + 4 * 17408 + 256 * 128 + 64;
  run<64, 0>(d, smem); run<64, 1>(d, smem);
  run<128, 0>(d, smem); run<128, 1>(d, smem);
  run<192, 0>(d, smem); run<192, 1>(d, smem);
  return bad != 0;
}
