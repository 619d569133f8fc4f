// Input staging, the step before the hot path: the reference pads / crops every
// evaluation utterance to 64 600 samples on the host by repeat-tiling (data_utils.py:45-52, called
// from Dataset_ASVspoof2019_devNeval.__getitem__ :208).  Here a batch of ragged utterances that is
// already in device memory (one concatenated buffer + offsets) becomes the (B, max_len) model input
// in one launch:  out[b][i] = x_b[i mod len_b]  (for len_b >= max_len this is the crop x_b[:max_len]).
// The same kernel, with a per-utterance start and target length, covers pad_random, dynamic_chunk_size and
// pad_sequence (data_utils.py:55-119): see aasist_stage_batch in include/aasist_b200.h.
#include "common.cuh"

namespace aasist {

// out[b][i] = i < target_b ? x_b[(start_b + i) mod len_b] : 0     (starts / targets may be null: 0 / row_len)
__global__ void __launch_bounds__(256)
stage_rows_kernel(const float* __restrict__ samples, const int64_t* __restrict__ offsets,
                  const int32_t* __restrict__ lengths, const int32_t* __restrict__ starts,
                  const int32_t* __restrict__ targets, float* __restrict__ out, int row_len) {
  const int b = blockIdx.y;
  const int len = lengths[b];
  const int start = starts ? starts[b] : 0;
  const int target = targets ? min(targets[b], row_len) : row_len;
  const float* src = samples + offsets[b];
  float* dst = out + (size_t)b * row_len;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_len; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < target) {
      const long long j = (long long)start + i;
      v = __ldg(src + (j < len ? j : j % len));
    }
    dst[i] = v;
  }
}

}  // namespace aasist

using namespace aasist;

#pragma GCC visibility push(default)
extern "C" {

int aasist_stage_batch(aasist_handle* h, const float* samples_dev, const int64_t* offsets_host,
                       const int32_t* lengths_host, const int32_t* starts_host, const int32_t* targets_host,
                       int32_t B, int32_t row_len, float* out_dev, void* stream) {
  if (!h || !samples_dev || !offsets_host || !lengths_host || !out_dev || B < 1 || row_len < 1) {
    set_error("aasist_stage_batch: invalid arguments");
    return AASIST_E_INVALID;
  }
  for (int b = 0; b < B; ++b) {
    if (lengths_host[b] < 1) {
      // reference: int(max_len / x_len) raises ZeroDivisionError for an empty utterance
      set_error("aasist_stage_batch: utterance %d is empty", b);
      return AASIST_E_INVALID;
    }
    if (starts_host && (starts_host[b] < 0 || starts_host[b] >= lengths_host[b])) {
      set_error("aasist_stage_batch: start %d of utterance %d outside its %d samples", starts_host[b], b,
                lengths_host[b]);
      return AASIST_E_INVALID;
    }
    if (targets_host && targets_host[b] < 0) {
      set_error("aasist_stage_batch: negative target length for utterance %d", b);
      return AASIST_E_INVALID;
    }
  }
  int dev_prev = -1;
  cudaGetDevice(&dev_prev);
  if (h->device >= 0 && dev_prev != h->device) AASIST_CUDA(cudaSetDevice(h->device));
  struct Restore {
    int prev, cur;
    ~Restore() { if (prev >= 0 && prev != cur) cudaSetDevice(prev); }
  } restore{dev_prev, h->device};
  cudaStream_t st = (cudaStream_t)stream;
  const size_t meta = sizeof(int64_t) * B + 3 * sizeof(int32_t) * B;
  if (h->stage_meta_bytes < meta) {
    cudaFree(h->stage_meta);
    h->stage_meta = nullptr;
    h->stage_meta_bytes = 0;
    AASIST_CUDA(cudaMalloc(&h->stage_meta, meta));
    h->stage_meta_bytes = meta;
  }
  int64_t* d_off = (int64_t*)h->stage_meta;
  int32_t* d_len = (int32_t*)((char*)h->stage_meta + sizeof(int64_t) * B);
  int32_t* d_start = d_len + B;
  int32_t* d_target = d_start + B;
  AASIST_CUDA(cudaMemcpyAsync(d_off, offsets_host, sizeof(int64_t) * B, cudaMemcpyHostToDevice, st));
  AASIST_CUDA(cudaMemcpyAsync(d_len, lengths_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
  if (starts_host) AASIST_CUDA(cudaMemcpyAsync(d_start, starts_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
  if (targets_host)
    AASIST_CUDA(cudaMemcpyAsync(d_target, targets_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
  dim3 grid((row_len + 256 * 8 - 1) / (256 * 8), B);
  {
    LaunchSpan span(h, "stage_rows", st);
    stage_rows_kernel<<<grid, 256, 0, st>>>(samples_dev, d_off, d_len, starts_host ? d_start : nullptr,
                                           targets_host ? d_target : nullptr, out_dev, row_len);
  }
  AASIST_CUDA(cudaGetLastError());
  return AASIST_OK;   // (copies from pageable host arrays return once staged: the caller may reuse them)
}

int aasist_pad_batch(aasist_handle* h, const float* samples_dev, const int64_t* offsets_host,
                     const int32_t* lengths_host, int32_t B, int32_t max_len, float* out_dev, void* stream) {
  return aasist_stage_batch(h, samples_dev, offsets_host, lengths_host, nullptr, nullptr, B, max_len, out_dev, stream);
}

int32_t aasist_pad_sequence_length(const int32_t* lengths_host, int32_t B) {
  if (!lengths_host || B < 1) {
    set_error("aasist_pad_sequence_length: invalid arguments");
    return AASIST_E_INVALID;
  }
  int32_t m = 0;
  for (int b = 0; b < B; ++b) m = lengths_host[b] > m ? lengths_host[b] : m;
  return ((m + 3) / 4) * 4;                       // data_utils.py:108-110
}

}  // extern "C"
#pragma GCC visibility pop
