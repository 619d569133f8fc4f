/*
 * aasist_b200.h -- C ABI of libaasist_b200.so: the B200 (sm_100a) implementation of the
 * AASIST / RawGAT-ST batched utterance-scoring forward pass.
 *
 * This is the drop-in boundary for the reference's model plug-in interface
 * (reference main.py:251-259 `get_model`, README.md:69-77): everything the reference does
 * between `model(batch_x)` (main.py:376) and the returned `(last_hidden, output)` tuple
 * (models/AASIST.py:806-921, models/RawNetGatSpoofST.py:324-356) happens behind these
 * entry points.  Plain C types only: raw pointers, sizes, an opaque handle and a
 * `cudaStream_t` passed as `void*`.  No torch types, no exceptions.  Every function returns
 * 0 on success and a negative `AASIST_E_*` code on failure; `aasist_last_error()` gives the
 * message of the calling thread's most recent failure.  A handle is bound to the CUDA
 * device that was current at `aasist_create` and is not thread-safe.
 *
 * There is NO CPU fallback: every compute entry point fails with AASIST_E_CUDA when no
 * CUDA device is usable.
 */
#ifndef AASIST_B200_H_
#define AASIST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AASIST_B200_ABI_VERSION 2

enum {
  AASIST_OK = 0,
  AASIST_E_INVALID = -1,   /* bad argument / shape (reference: RuntimeError / ValueError)   */
  AASIST_E_STATE = -2,     /* call order violated (e.g. forward before finalize)            */
  AASIST_E_PARAM = -3,     /* unknown / missing / mis-sized parameter (strict state_dict)   */
  AASIST_E_CUDA = -4,      /* CUDA runtime/driver error, or no device                        */
  AASIST_E_WORKSPACE = -5  /* workspace too small                                           */
};

/* model kinds: which reference `Model.forward` is reproduced */
enum {
  AASIST_KIND_AASIST = 0,   /* models/AASIST.py:728-921 with the (2,3) Residual_block encoder
                               (models/RawNetGatSpoofST.py:225-278) the shipped weights fit   */
  AASIST_KIND_RAWGAT_ST = 1,/* models/RawNetGatSpoofST.py:281-356                            */
  AASIST_KIND_ROBUST = 2    /* models/AASIST_Robust.py:90-303 (eval): `first_conv` sinc filters of 1025 taps at
                               stride 256 (:96-102), the fork's 3x3 Residual_block encoder, one heterogeneous
                               branch, auxiliary head on the mean encoder feature, softmax-weighted ensemble */
};

/* encoder block type (AASIST / ROBUST kinds) */
enum {
  AASIST_ENC_RESIDUAL23 = 0, /* (2,3) Residual_block, models/RawNetGatSpoofST.py:225-278: what the shipped
                                checkpoints hold (tensor-core path available)                          */
  AASIST_ENC_RES2NET = 1,    /* Res2NetBlock + SELayer, models/AASIST.py:506-669: what the fork's
                                `Model(d_args)` builds (fp32 CUDA-core kernels)                        */
  AASIST_ENC_RESIDUAL33 = 2  /* the fork's 3x3 Residual_block, models/AASIST.py:672-725 (AASIST-Robust) */
};

/* arithmetic of the sinc / encoder convolutions and of the graph stages' attention projections (everything else
 * in the graph stages is always fp32) */
enum {
  AASIST_PREC_FP32 = 0,   /* fp32 FFMA on CUDA cores                                         */
  AASIST_PREC_F16X3 = 1,  /* tcgen05 tensor cores: fp16 hi/lo operand split, 3 products,
                             fp32 accumulation in TMEM (logits within ~1e-5 of fp32)         */
  AASIST_PREC_F16X2 = 2   /* EXPERIMENT, opt-in only: like F16X3 in the sinc front end and encoder block 0, but
                             encoder blocks 1..5 drop the weight-correction product (a_hi*w_hi + a_lo*w_hi:
                             weights rounded to fp16).  +8 % throughput on AASIST.  Measured against the reference
                             goldens (tests/test_gpu_tc.py, profiles/README.md) it does NOT hold the shipping bar
                             (logits <= 1e-3 and 100 % ordered GraphPool-index match): AASIST reaches 1.0e-3 with
                             index flips on long utterances, AASIST-L flips top-k.  Never a default.          */
};

/* Mirrors the reference's `model_config` dict (config/AASIST.conf:13-21) that
 * `Model.__init__(d_args)` reads (models/AASIST.py:729-804). */
typedef struct aasist_config {
  int32_t kind;              /* AASIST_KIND_*                                                */
  int32_t precision;         /* AASIST_PREC_*                                                */
  int32_t first_conv;        /* d_args["first_conv"] (128 -> 129 taps, AASIST.py:449-450)    */
  int32_t n_filters;         /* d_args["filts"][0]  (70; must give 23 pooled bands)          */
  int32_t enc_channels[6][2];/* (in,out) of the six residual blocks: filts[1],[2],[3],[4]x3  */
  int32_t gat_dims[2];       /* d_args["gat_dims"]            (AASIST kind only)             */
  double pool_ratios[4];     /* d_args["pool_ratios"]         (AASIST kind only)             */
  double temperatures[4];    /* d_args["temperatures"]        (AASIST kind only)             */
  int32_t sample_rate;       /* 16000 (CONV.__init__ default, AASIST.py:430)                 */
  int32_t encoder;           /* AASIST_ENC_* (ROBUST kind: must be RESIDUAL33)               */
  int32_t res2net_width;     /* d_args.get("res2net_width", 14)   (AASIST.py:739)            */
  int32_t res2net_scale;     /* d_args.get("res2net_scale", 8)    (AASIST.py:740)            */
  int32_t spk_emb_dim;       /* d_args["spk_emb_dim"] when d_args["speaker_conditioning"], else 0
                                (SpeakerConditioningModule on gat_dims[1], AASIST.py:743-755) */
  int32_t spk_level;         /* 0 = "frame", 1 = "utterance" (AASIST.py:746)                 */
  int32_t spk_use_attention; /* d_args.get("use_attention", True) (AASIST.py:747)            */
  int32_t reserved[1];
} aasist_config;

/* Per-call options of `Model.forward(x, Freq_aug=..., speaker_embedding=...)` (AASIST.py:806). */
typedef struct aasist_forward_opts {
  const float* speaker_embedding; /* DEVICE (B, spk_emb_dim) fp32, or NULL (reference main.py:375 passes None) */
  int32_t freq_mask_start;        /* Freq_aug (AASIST.py:486-490): rows [start, start+count) of the sinc bank   */
  int32_t freq_mask_count;        /* are zeroed for this call; count 0 = no masking.  The CALLER draws A, A0    */
                                  /* (the reference uses numpy's and Python's global RNGs)                      */
  int32_t reserved[4];
} aasist_forward_opts;

typedef struct aasist_handle aasist_handle;

/* ---- life cycle ------------------------------------------------------------------------ */
int aasist_abi_version(void);
const char* aasist_last_error(void);

/* Replaces `Model(d_args)` (reference main.py:255, models/AASIST.py:729). */
int aasist_create(const aasist_config* cfg, aasist_handle** out);
int aasist_destroy(aasist_handle* h);

/* Replaces `model.load_state_dict(...)` (reference main.py:104-105), one tensor per call.
 * `name` is the reference state_dict key (e.g. "encoder.0.0.conv1.weight"); `data` may be a
 * host or a device pointer to `numel` fp32 values; the handle keeps its own copy.
 * `*.num_batches_tracked` keys are accepted and ignored. */
int aasist_set_param(aasist_handle* h, const char* name, const float* data, int64_t numel);
/* Number of state_dict tensors the configuration expects / their names (for strict loading). */
int aasist_num_params(const aasist_handle* h);
const char* aasist_param_name(const aasist_handle* h, int index, int64_t* numel);
/* Checks that every expected tensor was set, folds the eval-mode batch norms into the
 * adjacent weights, repacks for the kernels and builds the sinc filter bank ON DEVICE
 * (models/AASIST.py:460-482). */
int aasist_finalize(aasist_handle* h);

/* ---- the hot path ---------------------------------------------------------------------- */
/* Bytes of scratch `aasist_forward` needs for a batch of B utterances of L samples. */
int64_t aasist_workspace_bytes(const aasist_handle* h, int32_t B, int32_t L);

/* Width of `last_hidden`: 5*gat_dims[1] (AASIST.py:909-910) or 7 (RawNetGatSpoofST.py:353). */
int aasist_hidden_dim(const aasist_handle* h);
/* Total number of top-k indices written per utterance for length L (all GraphPools, in the
 * order pool_S, pool_T, pool_hS1, pool_hT1, pool_hS2, pool_hT2  /  pool_T, pool_S, pool_ST),
 * and per-pool (n_nodes_in, k) pairs written to `nk` (2*n_pools ints) when not NULL. */
int aasist_topk_layout(const aasist_handle* h, int32_t L, int32_t* n_pools, int32_t* nk);

/* Replaces `Model.forward(x)` in eval mode (models/AASIST.py:806-921; Freq_aug=False,
 * speaker_embedding=None).  All pointers are DEVICE pointers:
 *   x            (B, L) fp32 waveforms, row-major
 *   last_hidden  (B, hidden_dim) fp32 out
 *   logits       (B, 2) fp32 out          -- `output`; the score is logits[:,1] (main.py:377)
 *   topk_idx     (B, topk_total) int32 out, or NULL -- GraphPool indices in descending score
 *                order (AASIST.py:316); exact ties are broken towards the lower node index
 *   pool_scores  (B, sum of n_nodes_in) fp32 out, or NULL -- pre-sigmoid GraphPool weights
 *   workspace    >= aasist_workspace_bytes(h,B,L) bytes, 256-byte aligned
 *   stream       cudaStream_t (NULL = legacy default stream)
 * Asynchronous with respect to the host. */
int aasist_forward(aasist_handle* h, const float* x, int32_t B, int32_t L,
                   float* last_hidden, float* logits, int32_t* topk_idx, float* pool_scores,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* `Model.forward(x, Freq_aug, speaker_embedding)` with its optional arguments; `opts` may be NULL
 * (= aasist_forward).  For AASIST_KIND_ROBUST `last_hidden` receives the (B,2) ensemble logits and
 * `logits` the main head's logits -- the reference returns `(ensemble_logits, logits)`
 * (AASIST_Robust.py:303).  `workspace` may be NULL: the handle then uses (and grows on demand) a
 * scratch buffer of its own, shared with aasist_forward_host and the scoring stream. */
int aasist_forward_ex(aasist_handle* h, const float* x, int32_t B, int32_t L, const aasist_forward_opts* opts,
                      float* last_hidden, float* logits, int32_t* topk_idx, float* pool_scores,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* Same call with HOST buffers: pinned staging + H2D of x, forward, D2H of logits and
 * last_hidden (either may be NULL), stream-synchronised before return.  This is what
 * reference main.py:372-377 does per batch (`.to(device)` ... `.cpu()`). */
int aasist_forward_host(aasist_handle* h, const float* x_host, int32_t B, int32_t L,
                        float* last_hidden_host, float* logits_host, void* stream);

/* ---- input staging (the step before the path) ------------------------------------------------ */
/* Replaces the host-side `pad` (data_utils.py:45-52, used for every evaluation utterance at :208):
 * B ragged utterances, already in device memory as one concatenated fp32 buffer, become the
 * (B, max_len) model input: out[b][i] = x_b[i mod len_b] (repeat-tile; crop when len_b >= max_len).
 * `offsets_host[b]` is the element offset of utterance b in `samples_dev`, `lengths_host[b]` its
 * length (>= 1; an empty utterance is an error like the reference's ZeroDivisionError). */
int aasist_pad_batch(aasist_handle* h, const float* samples_dev, const int64_t* offsets_host,
                     const int32_t* lengths_host, int32_t B, int32_t max_len, float* out_dev, void* stream);

/* Generalisation used by the other staging functions of data_utils.py: row b of the (B, row_len) output is
 *   out[b][i] = i < target_b ? x_b[(start_b + i) mod len_b] : 0
 * `starts_host` NULL = all 0; `targets_host` NULL = row_len for every row.  With the reference's random draws made
 * by the caller this reproduces, bit-exactly,
 *   pad               (data_utils.py:45-52)    start 0, target = row_len = max_len
 *   pad_random        (data_utils.py:55-65)    start = np.random.randint(len - max_len) when len >= max_len
 *   dynamic_chunk_size(data_utils.py:68-97)    target = randint(min, max+1), start = randint(0, len-target+1)
 *   pad_sequence      (data_utils.py:100-119)  start 0, target = min(len, row_len), row_len = the batch maximum
 *                                              rounded up to a multiple of 4 (aasist_pad_sequence_length) */
int aasist_stage_batch(aasist_handle* h, const float* samples_dev, const int64_t* offsets_host,
                       const int32_t* lengths_host, const int32_t* starts_host, const int32_t* targets_host,
                       int32_t B, int32_t row_len, float* out_dev, void* stream);
/* ((max(lengths) + 3) / 4) * 4  (data_utils.py:108-110) */
int32_t aasist_pad_sequence_length(const int32_t* lengths_host, int32_t B);

/* ---- the scoring loop (the caller of the path) --------------------------------------------------- */
/* Replaces the per-batch `batch_x.to(device)` -> `model(batch_x)` -> `batch_out[:,1].cpu()` round trip of
 * produce_evaluation_file (main.py:364-378) with a pipeline: `aasist_score_submit` copies one batch of HOST
 * utterances through one of two pinned staging buffers, issues its H2D on a copy stream and enqueues its forward
 * behind it -- the H2D of batch n+1 runs under the forward of batch n, and the call returns without waiting for
 * either.  Scores (logits, and last_hidden when requested) accumulate in device memory;
 * `aasist_score_finish` waits once and copies them to the host.  One scratch buffer serves every forward.
 *   begin : capacity = total utterances of the session; max_batch = largest B of any submit; L = samples each
 *   submit: x_host (B, L) fp32, pageable or pinned; returns after the staging copy (no device wait unless both
 *           staging buffers are still in flight)
 *   finish: logits_out (n, 2) and last_hidden_out (n, hidden_dim) may each be NULL and may be HOST or DEVICE
 *           pointers (a multi-GPU caller all-gathers the device copy without a host round trip); `logits_dev_out`,
 *           when not NULL, receives the device pointer of the accumulated logits (valid until the next begin /
 *           destroy).  Waits for the stream once.  Returns n, the number of utterances scored. */
int aasist_score_begin(aasist_handle* h, int64_t capacity, int32_t max_batch, int32_t L, void* stream);
int aasist_score_submit(aasist_handle* h, const float* x_host, int32_t B);
int64_t aasist_score_finish(aasist_handle* h, float* logits_out, float* last_hidden_out,
                            const float** logits_dev_out);

/* ---- detection metrics (the step after the path) ---------------------------------------------- */
/* Replaces compute_det_curve / compute_eer (evaluation.py:120-154) and the t-DCF curve of compute_tDCF
 * (evaluation.py:266-282) for scores that are already in device memory (float64, as numpy holds them):
 * stable ascending sort of the concatenation [targets, nontargets], running target count,
 *   frr[i] = cum_i / n_target,  far[i] = (n_nontarget - (i - cum_i)) / n_nontarget,  frr[0] = 0, far[0] = 1,
 *   thr[0] = min score - 0.001, thr[i] = i-th smallest score,
 *   tdcf[i] = (c1 * frr[i] + c2 * far[i]) / min(c1, c2)     (skipped when c1 < 0 and c2 < 0),
 * all in IEEE float64 with the reference's operation order: every value is bit-identical to numpy's.
 * results_host[8] = { EER = (frr[k] + far[k]) / 2 at k = first argmin |frr - far|, thr[k], k,
 *                     min t-DCF, its threshold, its index (first argmin), number of distinct scores,
 *                     flags (bit 0: a score is NaN or infinite) }.
 * The four curve outputs are optional device arrays of n_target + n_nontarget + 1 doubles (NULL to skip).
 * Stateless (no handle); synchronises `stream` before returning.  c1/c2 are the caller's
 * C1 = Ptar*(Cmiss_cm - Cmiss_asv*Pmiss_asv) - Pnon*Cfa_asv*Pfa_asv, C2 = Cfa_cm*Pspoof*(1 - Pmiss_spoof_asv). */
int64_t aasist_det_workspace_bytes(int64_t n_total);
int aasist_det_metrics(const double* target_dev, int64_t n_target, const double* nontarget_dev,
                       int64_t n_nontarget, double c1, double c2, double* results_host,
                       double* frr_dev, double* far_dev, double* thr_dev, double* tdcf_dev,
                       void* workspace_dev, int64_t workspace_bytes, void* stream);

/* ---- per-stage entry points (parity tests; same kernels as aasist_forward) -------------- */
/* Device copy of the sinc filter bank, (n_filters, taps) fp32, into `bank_dev`. */
int aasist_get_filterbank(aasist_handle* h, float* bank_dev, int32_t* n_filters, int32_t* taps);
/* Sinc conv + |.| + 3x3 max-pool + first_bn + SELU (AASIST.py:823-831):
 * x (B,L) -> out (B,23,floor((L-taps+1)/3)) fp32 NCHW (one channel). */
int aasist_frontend(aasist_handle* h, const float* x, int32_t B, int32_t L, float* out,
                    void* workspace, int64_t workspace_bytes, void* stream);
/* Residual block `index` (0..5) of encoder `enc` (0; RawGAT-ST: 0 = encoder_T, 1 = encoder_S)
 * (RawNetGatSpoofST.py:258-278): in (B,Cin,23,W) -> out (B,Cout,23,floor(W/3)), fp32 NCHW. */
int aasist_encoder_block(aasist_handle* h, int32_t enc, int32_t index, const float* in, int32_t B,
                         int32_t W, float* out, void* workspace, int64_t workspace_bytes,
                         void* stream);
/* Everything after the encoder (AASIST.py:841-921 / RawNetGatSpoofST.py:338-356):
 * e (B,C,23,NT) fp32 NCHW (RawGAT-ST: e = encoder_T output, e2 = encoder_S output). */
int aasist_graph(aasist_handle* h, const float* e, const float* e2, int32_t B, int32_t NT,
                 float* last_hidden, float* logits, int32_t* topk_idx, float* pool_scores,
                 void* stream);

/* Input-range guard of the tensor-core path.  The f16x3 front end holds a sample as an fp16 pair of x * 2^10, which
 * saturates for |x| > 63.96: waveforms are expected in [-1, 1] (what soundfile returns), and an un-normalised one
 * (e.g. int16-scale samples) would give wrong logits silently, where the reference's fp32 forward would not.  The
 * kernel raises a host-visible flag when a sample saturates; this returns 1 if that happened in any forward that has
 * COMPLETED since the last reset (call it after synchronising the stream), else 0.  fp32 handles always return 0. */
int aasist_input_range_exceeded(aasist_handle* h, int32_t reset);

/* Number of kernels this library has launched through the handle so far. */
int64_t aasist_launch_count(const aasist_handle* h);

/* Per-kernel device timing: while enabled, every launch made through the handle is bracketed
 * by CUDA events recorded on the launching stream.  `aasist_profile_report` synchronises the
 * device, folds the pending events into per-kernel totals and writes them as JSON
 * (`[{"kernel": "...", "launches": n, "ms": total}, ...]`) into `buf`; `reset` != 0 clears the
 * totals afterwards.  Returns the number of bytes written (excluding the NUL) or < 0. */
int aasist_profile_enable(aasist_handle* h, int32_t enable);
int aasist_profile_report(aasist_handle* h, char* buf, int64_t buf_bytes, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* AASIST_B200_H_ */
