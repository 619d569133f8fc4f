// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, cta_group::1) vs N, with minimal issue
// overhead (descriptors precomputed, 12 MMAs unrolled per loop trip).  One CTA per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include "../aasist_b200/csrc/ptx.cuh"
using namespace aasist::ptx;

// MODE 0: same A tile/rows, same D;  1: A alternates between 4 tiles + row shifts (like the conv);  2: + 3 D targets
template <int N, int MODE>
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
    for (int i = 0; i < 4; ++i) ad[i] = umma_desc_sw128(a0 + (MODE ? i * 17408 + 128 * (i % 3) : 0));
    bd[0] = umma_desc_sw128(b0); bd[1] = umma_desc_sw128(b0 + 64);
    long long t0 = clock64();
    for (int t = 0; t < trips; ++t) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t d = tm + (MODE == 2 ? (i % 3) * N : 0);
        umma_f16(d, ad[i], bd[0], idesc, 1);
        umma_f16(d, ad[i] + 4, bd[0], idesc, 1);
        umma_f16(d, ad[i], bd[1], idesc, 1);
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

template <int N, int MODE> void run(long long* d, size_t smem) {
  const int trips = 400;
  cudaFuncSetAttribute(mma_bench<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mma_bench<N, MODE><<<148, 64, smem>>>(trips, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  double per = (double)h[1] / (trips * 12);
  printf("N=%3d mode=%d | issue %6.1f  total %6.1f cyc/mma | %5.0f MAC/clk | operand bytes/clk %5.1f\n", N, MODE,
         (double)h[0] / (trips * 12), per, 128.0 * N * 16 / per, (4096.0 + N * 32) / per);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  size_t smem = 1024 + 4 * 17408 + 256 * 128 + 64;
  run<32, 0>(d, smem); run<32, 1>(d, smem); run<32, 2>(d, smem);
  run<64, 0>(d, smem); run<64, 1>(d, smem); run<64, 2>(d, smem);
  run<96, 0>(d, smem); run<96, 1>(d, smem); run<96, 2>(d, smem);
  run<128, 1>(d, smem); run<128, 2>(d, smem);
  run<192, 1>(d, smem); run<192, 2>(d, smem);
  run<256, 1>(d, smem);
  return 0;
}
