// Input staging, the step before the hot path: the reference pads / crops every
// evaluation utterance to 64 600 samples on the host by repeat-tiling (data_utils.py:45-52, called
// from Dataset_ASVspoof2019_devNeval.__getitem__ :208).  Here a batch of ragged utterances that is
// already in device memory (one concatenated buffer + offsets) becomes the (B, max_len) model input
// in one launch:  out[b][i] = x_b[i mod len_b]  (for len_b >= max_len this is the crop x_b[:max_len]).
#include "common.cuh"

namespace aasist {

__global__ void __launch_bounds__(256)
pad_tile_kernel(const float* __restrict__ samples, const int64_t* __restrict__ offsets,
                const int32_t* __restrict__ lengths, float* __restrict__ out, int max_len) {
  const int b = blockIdx.y;
  const int len = lengths[b];
  const float* src = samples + offsets[b];
  float* dst = out + (size_t)b * max_len;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max_len; i += gridDim.x * blockDim.x)
    dst[i] = __ldg(src + (i < len ? i : i % len));
}

}  // namespace aasist

using namespace aasist;

#pragma GCC visibility push(default)
extern "C" int aasist_pad_batch(aasist_handle* h, const float* samples_dev, const int64_t* offsets_host,
                                const int32_t* lengths_host, int32_t B, int32_t max_len, float* out_dev,
                                void* stream) {
  if (!h || !samples_dev || !offsets_host || !lengths_host || !out_dev || B < 1 || max_len < 1) {
    set_error("aasist_pad_batch: invalid arguments");
    return AASIST_E_INVALID;
  }
  for (int b = 0; b < B; ++b)
    if (lengths_host[b] < 1) {
      // reference: int(max_len / x_len) raises ZeroDivisionError for an empty utterance
      set_error("aasist_pad_batch: utterance %d is empty", b);
      return AASIST_E_INVALID;
    }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t meta = sizeof(int64_t) * B + sizeof(int32_t) * B;
  if (h->stage_meta_bytes < meta) {
    cudaFree(h->stage_meta);
    h->stage_meta = nullptr;
    h->stage_meta_bytes = 0;
    AASIST_CUDA(cudaMalloc(&h->stage_meta, meta));
    h->stage_meta_bytes = meta;
  }
  int64_t* d_off = (int64_t*)h->stage_meta;
  int32_t* d_len = (int32_t*)((char*)h->stage_meta + sizeof(int64_t) * B);
  AASIST_CUDA(cudaMemcpyAsync(d_off, offsets_host, sizeof(int64_t) * B, cudaMemcpyHostToDevice, st));
  AASIST_CUDA(cudaMemcpyAsync(d_len, lengths_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
  dim3 grid((max_len + 256 * 8 - 1) / (256 * 8), B);
  {
    LaunchSpan span(h, "pad_tile", st);
    pad_tile_kernel<<<grid, 256, 0, st>>>(samples_dev, d_off, d_len, out_dev, max_len);
  }
  AASIST_CUDA(cudaGetLastError());
  return AASIST_OK;
}
#pragma GCC visibility pop
