// Shared definitions for libaasist_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/aasist_b200.h"

namespace aasist {

constexpr float kSeluAlpha = 1.6732632423543772f;
constexpr float kSeluScale = 1.0507009873554805f;
constexpr double kBnEps = 1e-5;
constexpr int kSpecNodes = 23;  // pooled sinc bands == spectral graph nodes (AASIST.py:774)

__device__ __forceinline__ float selu(float v) {
  // torch SELU: scale * (max(0,x) + min(0, alpha*(exp(x)-1)))
  return v > 0.f ? kSeluScale * v : (kSeluScale * kSeluAlpha) * expm1f(v);
}

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define AASIST_CUDA(expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return cuda_fail(_e, #expr);     \
  } while (0)

// ---------------------------------------------------------------------------------------
// packed parameters
// ---------------------------------------------------------------------------------------
struct ConvBlockF32 {  // one Residual_block, fp32 CUDA-core path
  int ci = 0, co = 0;
  bool downsample = false;
  float* w1 = nullptr;  // [ci][2][3][co]  conv1 with bn2 folded
  float* b1 = nullptr;  // [co]
  float* w2 = nullptr;  // [co][2][3][co]
  float* b2 = nullptr;  // [co]  conv2 bias (+ conv_downsample bias)
  float* wd = nullptr;  // [ci][3][co]     conv_downsample (k(1,3)) or null
};

// Res2NetBlock + SELayer (models/AASIST.py:506-669), fp32 CUDA-core path (encoder_res2.cu)
struct Res2Group {
  int c0, n;        // channel range [c0, c0+n) of the block input / of `convs[i]`'s output
  int level;        // 0 = reads only the block input; l > 0 = also adds the raw output of the previous split
  int feeds_next;   // its raw conv output is the next split's addend (keep it in the `raw` scratch)
};
struct Res2Launch {   // one split-conv launch: the splits of one dependency level that share an accumulator width
  int level, nreg, n_max, first, count;     // lvl_groups[first .. first+count)
  int n_tot;                                // channels of all its splits (shared-memory tile rows)
};
struct Res2BlockF32 {
  int index = 0;           // encoder block number (kernel names in the per-launch profile)
  int ci = 0, co = 0;
  bool first = false, downsample = false;
  int n_groups = 0, n_levels = 0;
  std::vector<Res2Group> groups;     // host copy
  std::vector<Res2Launch> launches;  // in dependency order
  int* lvl_groups_dev = nullptr;     // split indices, launch by launch
  Res2Group* groups_dev = nullptr;
  float* bn1 = nullptr;    // [2][ci]  scale, shift of bn1 (LIVE in this block, AASIST.py:611-613); null when first
  float* gw = nullptr;     // split convs, concatenated: group g at gw_off[g]: [n][n][3][3] (out,in,kh,kw)
  float* gb = nullptr;     // [ci] split conv biases, by absolute channel
  int* gw_off = nullptr;   // [n_groups] (device)
  float* bn2 = nullptr;    // [2][ci]  scale, shift of bn2
  float* wcat = nullptr;   // conv_cat [ci][3][3][cop]
  float* bcat = nullptr;   // [cop]
  float* se0 = nullptr;    // se.fc.0.weight (co/16, co)
  float* se2 = nullptr;    // se.fc.2.weight (co, co/16)
  int se_hidden = 0;
  float* wd = nullptr;     // conv_downsample [ci][3][cop] or null
  float* bd = nullptr;     // [cop]
};

// the fork's 3x3 Residual_block (models/AASIST.py:672-725), fp32 CUDA-core path
struct ConvBlock33F32 {
  int ci = 0, co = 0;
  bool downsample = false;
  float* w1 = nullptr;  // [ci][3][3][cop]  conv1 with bn2 folded
  float* b1 = nullptr;
  float* w2 = nullptr;  // [co][3][3][cop]
  float* b2 = nullptr;  // conv2 bias (+ conv_downsample bias)
  float* wd = nullptr;  // [ci][3][cop] or null
};

// SpeakerConditioningModule (models/AASIST.py:325-415) on the gat_dims[1]-wide fused node features
struct SpkParams {
  int emb_dim;            // 0 = module absent
  int use_attention;
  const float *projW, *projB;     // proj: (g1, emb_dim), (g1)
  const float *att0Wt, *att0B;    // attention.0: weight^T [2*g1][g1], (g1)
  const float* att2W; float att2B;  // attention.2: (g1), scalar
  const float *fusWt, *fusB;      // fusion.0: weight^T [2*g1][g1], (g1)
};

struct GatParams {  // GraphAttentionLayer (AASIST.py:17-110), eval BN folded into the projections
  int D, Do;
  const float* attWt;  // [D][Do]  att_proj.weight^T
  const float* attImg; // att_proj.weight as an fp16 hi/lo tensor-core operand image (graph.cu att_image_floats)
  const float* attB;   // [Do]
  const float* attW;   // [Do]     att_weight
  const float* pWt;    // [D][Do]  proj_with_att.weight^T    * bn_scale
  const float* qWt;    // [D][Do]  proj_without_att.weight^T * bn_scale
  const float* bias;   // [Do]     (b_with + b_without - bn_mean) * bn_scale + bn_bias
  float temp;
};

struct HtrgParams {  // HtrgGraphAttentionLayer (AASIST.py:113-282)
  int D, Do;
  const float *t1Wt, *t1B, *t2Wt, *t2B;   // proj_type1/2: [D][D], [D]
  const float *attWt, *attB;              // att_proj
  const float* attImg;                    // att_proj.weight as an fp16 hi/lo tensor-core operand image
  const float *w11, *w22, *w12;           // att_weight11/22/12 [Do]
  const float *attMWt, *attMB, *wM;       // att_projM, att_weightM
  const float *pWt, *qWt, *bias;          // node projection, BN folded
  const float *pMWt, *qMWt, *biasM;       // master projection (no BN): bias = b_withM + b_withoutM
  float temp;
};

struct PoolParams {  // GraphPool (AASIST.py:285-322)
  int D;
  const float* w;  // [D]
  float b;
};

struct GraphArgsAasist {
  // dims
  int B;             // utterances in this launch (a CTA walks b = blockIdx.x, blockIdx.x + gridDim.x, ...)
  int C, NT, g0, g1, nS, nT, nS2, nT2;
  int ld;            // smem row stride (floats), odd
  int nmax;          // max node count of any layer
  const float* e;    // (B,C,23,NT)
  const float* posS; // (23,C)
  const float *master1, *master2;  // (g0)
  GatParams gatS, gatT;
  HtrgParams st11, st12, st21, st22;
  PoolParams poolS, poolT, poolhS1, poolhT1, poolhS2, poolhT2;
  const float* outWt;  // [5*g1][2]
  float outB0, outB1;
  float* last_hidden;  // (B,5*g1)   [robust: (B,2) ensemble logits]
  float* logits;       // (B,2)
  int32_t* topk_idx;   // (B,topk_total) or null
  float* pool_scores;  // (B,score_total) or null
  int topk_total, score_total;
  // speaker conditioning (AASIST.py:895-900): applied to the fused T and S nodes when spk_emb != null
  SpkParams spk;
  const float* spk_emb;    // (B, spk.emb_dim) or null
  // AASIST-Robust variant (AASIST_Robust.py:248-301): one branch (st11 = ST1, st12 = ST2, poolhS1/poolhT1 =
  // pool_hS/pool_hT), readout without the master node, auxiliary head on mean(e), softmax-weighted ensemble
  int robust;
  int tc;                  // attention maps on the tensor cores (precision f16x3 / f16x2); 0 = CUDA-core fp32
  const float* auxWt;      // aux_out_layer.weight^T [C][2]
  float auxB0, auxB1, ens0, ens1;
};

struct GraphArgsRawGat {
  int NT;            // temporal nodes of encoder_S output (29)
  int tc;            // attention maps on the tensor cores (precision f16x3 / f16x2)
  int ld, nmax;
  const float* eT;   // encoder_T output (B,64,23,NT) -> 23 nodes (max over time)
  const float* eS;   // encoder_S output (B,64,23,NT) -> NT nodes (max over freq)
  GatParams gatT, gatS, gatST;
  PoolParams poolT, poolS, poolST;
  int nT, nS, nST;   // pooled node counts (14, 23, 7)
  const float *projTW, *projTB;   // Linear(14,12): [12][14], [12]
  const float *projSW, *projSB;   // Linear(23,12): [12][23], [12]
  const float* projSTW; float projSTB;  // Linear(16,1)
  const float *outW, *outB;       // Linear(7,2): [2][7], [2]
  float* last_hidden;  // (B,7)
  float* logits;
  int32_t* topk_idx;
  float* pool_scores;
  int topk_total, score_total;
};

struct TcState;  // tensor-core path state (encoder_tc.cu)

}  // namespace aasist

struct aasist_handle {
  aasist_config cfg;
  int device = 0;
  int taps = 129;
  bool finalized = false;
  int64_t launches = 0;
  std::vector<std::pair<std::string, int64_t>> expected;     // strict state_dict layout
  std::map<std::string, std::vector<float>> params;          // host copies
  // device-side packed state
  float* bank = nullptr;                 // (n_filters, taps)
  float bn0_scale = 1.f, bn0_shift = 0.f;  // first_bn folded to y = scale*p + shift
  aasist::ConvBlockF32 blocks[2][6];     // [encoder][block]
  aasist::Res2BlockF32 res2[6];          // AASIST_ENC_RES2NET
  aasist::ConvBlock33F32 blocks33[6];    // AASIST_ENC_RESIDUAL33
  float* bank_t = nullptr;               // robust front end: bank transposed to (taps, n_filters)
  int stride = 1;                        // sinc conv stride (robust: 256)
  int n_encoders = 1;
  float* graph_buf = nullptr;            // all packed graph parameters
  std::vector<float> graph_host;         // staging while packing
  aasist::GraphArgsAasist ga;            // pointers filled at finalize (io pointers per call)
  aasist::GraphArgsRawGat gr;
  aasist::TcState* tc = nullptr;
  // per-kernel event timing (aasist_profile_*)
  bool profiling = false;
  struct ProfSpan { const char* name; cudaEvent_t a, b; };
  std::vector<ProfSpan> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  std::map<std::string, std::pair<int64_t, double>> prof_totals;
  // pinned staging for aasist_forward_host
  float* pin_x = nullptr; size_t pin_x_bytes = 0;
  float* pin_out = nullptr; size_t pin_out_bytes = 0;
  void* dev_stage = nullptr; size_t dev_stage_bytes = 0;
  void* stage_meta = nullptr; size_t stage_meta_bytes = 0;   // offsets/lengths of aasist_pad_batch
  cudaStream_t copy_stream = nullptr;     // H2D of chunk c+1 overlaps the forward of chunk c
  cudaEvent_t copy_done[2] = {nullptr, nullptr}, start_ev = nullptr;
  // scratch owned by the handle (aasist_forward_ex with workspace == NULL, forward_host, scoring stream)
  void* own_ws = nullptr; size_t own_ws_bytes = 0;
  // input-range guard of the f16x3 front end: pinned, device-mapped flag (aasist_input_range_exceeded)
  int* range_flag = nullptr;
  // Freq_aug: masked copy of the tensor-core filter operand (frontend_tc.cu)
  uint8_t* front_bimg_masked = nullptr;
  // scoring stream (aasist_score_begin / submit / finish)
  struct ScoreStream {
    bool active = false;
    int64_t capacity = 0, n = 0;
    int max_batch = 0, L = 0, slot = 0;
    cudaStream_t st = nullptr;
    float* pin[2] = {nullptr, nullptr}; size_t pin_bytes = 0;     // pinned staging (pageable sources)
    float* dx[2] = {nullptr, nullptr}; size_t dx_bytes = 0;       // device input double buffer
    float* d_logits = nullptr; float* d_hidden = nullptr; size_t out_cap = 0;
    cudaEvent_t h2d_done[2] = {nullptr, nullptr};   // H2D of the slot finished  (host may refill the pinned slot)
    cudaEvent_t fwd_done[2] = {nullptr, nullptr};   // forward that read the slot finished (device slot reusable)
    bool used[2] = {false, false};
  } score;
};

namespace aasist {
// RAII span around one kernel launch: counts it and, when profiling, brackets it with events.
struct LaunchSpan {
  aasist_handle* h;
  cudaStream_t st;
  int idx = -1;
  LaunchSpan(aasist_handle* h_, const char* name, cudaStream_t st_) : h(h_), st(st_) {
    h->launches++;
    if (!h->profiling) return;
    auto get = [&]() {
      cudaEvent_t e;
      if (!h->prof_pool.empty()) { e = h->prof_pool.back(); h->prof_pool.pop_back(); }
      else cudaEventCreate(&e);
      return e;
    };
    aasist_handle::ProfSpan sp{name, get(), get()};
    cudaEventRecord(sp.a, st);
    h->prof_pending.push_back(sp);
    idx = (int)h->prof_pending.size() - 1;
  }
  ~LaunchSpan() {
    if (idx >= 0) cudaEventRecord(h->prof_pending[idx].b, st);
  }
};
// kernels' host launchers (each returns 0 / AASIST_E_*; counts launches into h->launches)
int build_filterbank(aasist_handle* h);
// mask_count > 0: rows [mask_start, mask_start+mask_count) of the bank are read as zero (Freq_aug)
int launch_frontend_f32(aasist_handle* h, const float* x, int B, int L, float* out, int mask_start, int mask_count,
                        cudaStream_t st);
// strided sinc front end of AASIST-Robust (AASIST_Robust.py:96-102,217-221): out (B, n_filters/3, Wp)
int launch_frontend_strided_f32(aasist_handle* h, const float* x, int B, int L, float* out, int mask_start,
                                int mask_count, cudaStream_t st);
// Res2Net+SE block / 3x3 residual block, fp32 NCHW in -> out (B,co,23,W/3); scratch carved from `ws`
size_t res2_block_scratch_floats(const Res2BlockF32& blk, int B, int W);
int launch_res2_block(aasist_handle* h, const Res2BlockF32& blk, const float* in, int B, int W, float* ws,
                      float* out, cudaStream_t st);
int launch_block33_f32(aasist_handle* h, const ConvBlock33F32& blk, const float* in, int B, int W, float* mid,
                       float* out, cudaStream_t st);
int launch_block_f32(aasist_handle* h, const ConvBlockF32& blk, const float* in, int B, int W,
                     float* mid, float* out, cudaStream_t st);
// att_proj.weight (Do, D) -> fp16 hi/lo operand image for the tensor-core attention maps, as raw bytes in floats
std::vector<float> att_image_floats(const std::vector<float>& w, int D, int Do);
int launch_graph_aasist(aasist_handle* h, const float* e, int B, int NT, const float* spk_emb,
                        float* last_hidden, float* logits, int32_t* topk, float* scores, cudaStream_t st);
int launch_graph_rawgat(aasist_handle* h, const float* eT, const float* eS, int B, int NT,
                        float* last_hidden, float* logits, int32_t* topk, float* scores,
                        cudaStream_t st);
int pooled_count(int n, double ratio, int min_nodes);
}  // namespace aasist
