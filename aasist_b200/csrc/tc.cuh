// Tensor-core (tcgen05, fp16 hi/lo split x3) path: entry points used by api.cu.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include "common.cuh"

constexpr int kCollectorDefault = 0x54;   // fused 32->32 block + conv_tc with 64 input channels (measured: profiles/README.md)

namespace aasist {
// A-operand collector reuse (ptx.cuh umma_f16_keep / _reuse) per kernel family; bit 0 sinc front end, 1 block 0
// (conv2 MMAs; bit 5: its conv1 / downsample MMAs), 2 fused 32->32 block conv1 (bit 6: its conv2), 3 conv_tc with
// 32 input channels, 4 conv_tc with 64 input channels.  The default is what
// measured faster on B200 (profiles/README.md); AASIST_COLLECTOR=<mask> overrides it for A/B runs.
inline int collector_mask() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AASIST_COLLECTOR");
    v = e ? atoi(e) : kCollectorDefault;
  }
  return v;
}
int tc_finalize(aasist_handle* h);
void tc_destroy(aasist_handle* h);
size_t tc_workspace_bytes(const aasist_handle* h, int B, int L);
// x (B,L) -> enc_out[e] (B,C,23,NT) fp32 NCHW for every encoder of the model
// mask_count > 0: Freq_aug, filters [mask_start, mask_start+mask_count) are zero for this call
int tc_encode(aasist_handle* h, const float* x, int B, int L, float** enc_out, void* ws, int mask_start,
              int mask_count, cudaStream_t st);
int tc_frontend_to_f32(aasist_handle* h, const float* x, int B, int L, float* out, int mask_start, int mask_count,
                       cudaStream_t st);
int tc_front_mask(aasist_handle* h, const uint8_t* bimg, int mask_start, int mask_count, cudaStream_t st,
                  const uint8_t** out);
int tc_block_f32io(aasist_handle* h, int enc, int index, const float* in, int B, int W, float* out,
                   void* ws, int64_t ws_bytes, cudaStream_t st);
// tensor-core sinc front end (frontend_tc.cu)
int tc_front_finalize(aasist_handle* h, uint8_t** bimg_dev);
int launch_frontend_tc(aasist_handle* h, const uint8_t* bimg, int sm_count, const float* x, int B, int L,
                       float* out, cudaStream_t st);
// encoder block 0 fully on tensor cores (block0_tc.cu)
void block0_pack_small(std::vector<uint8_t>& img, const std::vector<float>& w1, const std::vector<float>& wd,
                       const std::vector<float>& bias1, int co);
int block0_image_bytes();
int block0_w2_bytes();
int launch_block0_tc(aasist_handle* h, int sm_count, const uint8_t* wimg, const float* b1, const float* b2,
                     const float* z, int nb, int W, __half* out, cudaStream_t st);
// 32->32 residual block with identity shortcut, conv1 -> conv2 fused on chip (block_fused_tc.cu)
int launch_block_fused_tc(aasist_handle* h, int sm_count, const char* name, const CUtensorMap& tmX,
                          const uint8_t* w1img, const uint8_t* w2img, const float* b1, const float* b2,
                          const __half* x, int nb, int W, int Co, __half* out, float* out_f32, cudaStream_t st);
}  // namespace aasist
