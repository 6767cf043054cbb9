/* embclip_b200 -- C ABI of the B200-native EmbCLIP hot path (libembclip_b200.so).
 *
 * The reference (allenai/embodied-clip) has NO native / FFI boundary: its hot path is Python calling
 * PyTorch modules of two pinned dependencies (openai/CLIP @ 40f5484c, allenai/allenact @ v0.5.0; pins at
 * /root/reference/primitive_probing/environment.yml:22 and readme_files/baselines_robothor_objectnav.md:6).
 * Each entry point below therefore cites the *Python call* it replaces; the ctypes binding a maintainer
 * adds on the reference side is shown in INTEGRATION.md.
 *
 * Conventions: every call returns 0 on success, a negative EMBCLIP_E* code otherwise (no C++ exception
 * crosses the ABI; `embclip_last_error()` gives the message of the calling thread's last failure).  The
 * CALLER owns every tensor and the workspace (device memory, e.g. torch allocations); the library owns only
 * the handle, its layer table and TMA descriptors.  All work is enqueued on the `cudaStream_t` passed as
 * `void* stream`; no call synchronises the device or allocates device memory.  A handle is used from one
 * host thread at a time.
 */
#ifndef EMBCLIP_B200_H_
#define EMBCLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMBCLIP_OK 0
#define EMBCLIP_EINVAL (-1)   /* bad argument / unsupported shape            */
#define EMBCLIP_ECUDA (-2)    /* a CUDA runtime / driver call failed         */
#define EMBCLIP_ESTATE (-3)   /* call order violated (e.g. weights not bound) */
#define EMBCLIP_ENOSPC (-4)   /* workspace too small                         */

#define EMBCLIP_DTYPE_F16 0
#define EMBCLIP_DTYPE_F32 1

const char* embclip_last_error(void);
/* ABI version of this library (bumped on any signature change). */
int embclip_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * CLIP ModifiedResNet encoder (RN50): replaces `clip_model(clip_input)` + `clip_pool(...)` +
 * `clip_avgpool(...)` at primitive_probing/generate_data/thor_image_features.py:109,112,113
 * (reachable_image_features.py:88,91,92) and `ClipResNetEmbedder.forward` of
 * allenact_plugins/clip_plugin/clip_preprocessors.py.
 * ------------------------------------------------------------------------------------------------ */
typedef struct embclip_rn50* embclip_rn50_t;

typedef struct {
  int32_t layers[4];         /* (3,4,6,3) for RN50                         */
  int32_t width;             /* 64                                         */
  int32_t heads;             /* 32 attention-pool heads                    */
  int32_t output_dim;        /* 1024 (RN50x16: 768); 0 = plan without the attention-pool head (trunk / avg-pool only) */
  int32_t input_resolution;  /* 224                                        */
  int32_t arch;              /* 0: CLIP ModifiedResNet (clip/model.py).  1: torchvision ResNet with v1.5 Bottlenecks
                              * (models.resnet50: the reference's ImageNet baseline encoder, thor_image_features.py:46-49) --
                              * conv 7x7/2 + BN + ReLU + MaxPool(3,2,1) stem, stride on the 3x3 conv, strided 1x1 downsample;
                              * width 64, output_dim 0; heads unused */
} embclip_rn50_cfg;

/* One packed parameter of the device weight blob.  Python (embclip_b200/packing.py) folds BatchNorm into
 * the conv weights in fp32, reorders to the layouts named here and writes each tensor at `offset`. */
typedef struct {
  char name[64];     /* e.g. "layer1.0.conv2.w"  */
  int32_t dtype;     /* EMBCLIP_DTYPE_*          */
  int32_t ndim;
  int64_t shape[4];
  uint64_t offset;   /* bytes from blob start, 256-B aligned */
  uint64_t nbytes;
} embclip_param_info;

/* Build the layer table.  Needs no GPU.  width 64 (RN50, RN101) or 96 (RN50x16: the stem's 48 channels are carried as 64,
 * 16 of them zero); the attention-pool head is planned only for (input_resolution / 32)^2 + 1 <= 64 tokens. */
int embclip_rn50_create(const embclip_rn50_cfg* cfg, embclip_rn50_t* out);
int embclip_rn50_destroy(embclip_rn50_t h);
int embclip_rn50_num_params(embclip_rn50_t h);
int embclip_rn50_param_info(embclip_rn50_t h, int index, embclip_param_info* out);
uint64_t embclip_rn50_blob_bytes(embclip_rn50_t h);
/* Point the handle at a caller-owned device blob laid out per embclip_rn50_param_info. */
int embclip_rn50_bind_weights(embclip_rn50_t h, const void* device_blob, uint64_t nbytes);
/* Bytes of caller-provided scratch a forward at this batch size needs. */
uint64_t embclip_rn50_workspace_bytes(embclip_rn50_t h, int batch);
/* frames: fp32 NHWC [batch, R, R, 3], already mean/std normalised (what the AllenAct RGB sensor hands to
 * ClipResNetPreprocessor.process).  Any output pointer may be NULL (that head is skipped):
 *   out_trunk_nchw  fp32 [batch, 32*width, R/32, R/32]   (pool=False / 'clip_conv')
 *   out_avgpool     fp32 [batch, 32*width]               (pool=True  / 'clip_avgpool')
 *   out_attnpool    fp32 [batch, output_dim]             ('clip_attnpool', AttentionPool2d row 0) */
int embclip_rn50_forward(embclip_rn50_t h, const float* frames_nhwc, int batch, float* out_trunk_nchw,
                         float* out_avgpool, float* out_attnpool, void* workspace, uint64_t workspace_bytes,
                         void* stream);
/* Same, from RAW frames: uint8 NHWC [batch, R, R, 3]; (v / 255 - mean[c]) / std[c] -- the normalisation the AllenAct
 * RGB sensor does on the host (ClipResNetPreprocessor.CLIP_RGB_MEANS / CLIP_RGB_STDS) -- is applied inside the stem
 * kernel, so the host-to-device copy is 4x smaller.  mean3 / std3: host arrays. */
int embclip_rn50_forward_u8(embclip_rn50_t h, const uint8_t* frames_nhwc_u8, const float* mean3, const float* std3, int batch,
                            float* out_trunk_nchw, float* out_avgpool, float* out_attnpool, void* workspace,
                            uint64_t workspace_bytes, void* stream);
/* Introspection for layer-by-layer parity tests: intermediate activations live in the workspace. */
typedef struct {
  char name[64];
  int32_t dtype;
  int32_t n, h, w, c;   /* NHWC; n == 0: the tensor is not materialised (its producer feeds fused consumers on chip) */
  uint64_t offset;      /* bytes from workspace start */
} embclip_act_info;
int embclip_rn50_num_acts(embclip_rn50_t h);
int embclip_rn50_act_info(embclip_rn50_t h, int batch, int index, embclip_act_info* out);
/* Per-op device timing of one forward (CUDA events on `stream`; synchronises). op_ms[i] <- milliseconds,
 * names[i*64..] <- op name.  Returns the number of ops, or a negative error. */
int embclip_rn50_profile(embclip_rn50_t h, const float* frames_nhwc, int batch, float* out_trunk_nchw,
                         float* out_avgpool, float* out_attnpool, void* workspace, uint64_t workspace_bytes,
                         void* stream, float* op_ms, char* names, int max_ops);
/* The same for raw uint8 frames (arguments as embclip_rn50_forward_u8). */
int embclip_rn50_profile_u8(embclip_rn50_t h, const uint8_t* frames_nhwc_u8, const float* mean3, const float* std3, int batch,
                            float* out_trunk_nchw, float* out_avgpool, float* out_attnpool, void* workspace,
                            uint64_t workspace_bytes, void* stream, float* op_ms, char* names, int max_ops);
/* Trunk forward whose ONLY result is the actor-critic path's input: fp16 NHWC pixel rows [batch * fres * fres, embed]
 * (the layout of embclip_ac_pack_features, bit-identical to packing the fp32 NCHW trunk output).  The last conv rounds to fp16
 * in its own epilogue: no fp32 trunk tensor, no cast launch.  frames_are_u8 = 0: fp32 NHWC normalised frames (mean3 / std3
 * ignored); 1: raw uint8 NHWC frames, normalised in the stem kernel.  Replaces ClipResNetEmbedder.forward (trunk only)
 * + ResnetTensorNavActorCritic's view of the features for a rollout loop that keeps its storage on the device. */
int embclip_rn50_encode_rows_f16(embclip_rn50_t h, const void* frames_nhwc, int frames_are_u8, const float* mean3, const float* std3,
                                 int batch, void* out_rows_f16, void* workspace, uint64_t workspace_bytes, void* stream);
/* fp16 copy of the trunk output of the LAST forward on this workspace, as NHWC pixel rows [batch * fres * fres, embed] -- the
 * layout the actor-critic path consumes (embclip_ac_pack_features output), bit-identical to packing the fp32 NCHW trunk
 * output.  Lets a rollout loop skip the fp32 NCHW round trip. */
int embclip_rn50_export_rows_f16(embclip_rn50_t h, int batch, const void* workspace, uint64_t workspace_bytes, void* out_rows_f16,
                                 void* stream);
/* How many kernels one forward launches (for bench.py's gpu_launches). */
int embclip_rn50_launches_per_forward(embclip_rn50_t h, int want_trunk, int want_avgpool, int want_attnpool);

/* ------------------------------------------------------------------------------------------------
 * Primitive ops (the kernels the plans above are made of), exposed for unit tests and reuse.
 * All tensors fp16 unless noted, device pointers, row-major.
 * ------------------------------------------------------------------------------------------------ */
/* out[M,N] = act( [A0 | A1][M, K0+K1] . W[N, K0+K1]^T + bias[N] + residual[M,N] ).  A1/bias/residual may
 * be NULL (K1 = 0).  out_f32 != 0 writes fp32.  K0, K1 multiples of 32; N multiple of 32.
 * Replaces nn.Conv2d(k=1) / nn.Linear of clip/model.py and allenact basic_models. */
int embclip_gemm_f16(const void* a0, const void* a1, const void* w, const float* bias, const void* residual,
                     void* out, int M, int N, int K0, int K1, int relu, int out_f32, void* stream);
/* Grouped variant: N-group g = n / grp_n reads A columns [g*grp_a_koff, +K0) and W columns
 * [g*grp_b_koff, +K0), W rows n % grp_b_nmod when grp_b_nmod != 0 (per-head contractions of AttentionPool2d).
 * lda / ldw = row pitches (elements) of A and W. */
int embclip_gemm_grouped_f16(const void* a, int lda, const void* w, int ldw, int w_rows, const float* bias,
                             void* out, int M, int N, int K, int grp_n, int grp_a_koff, int grp_b_koff,
                             int grp_b_nmod, int relu, int out_f32, void* stream);
/* 3x3 / pad 1 / stride 1 conv, NHWC: in [B,H,W,Cin], w [Cout, 9*Cin] (tap-major: kh, kw, cin), out [B,H,W,Cout];
 * pool == 1 fuses the nn.AvgPool2d(2) that follows it in the anti-aliased strided Bottleneck / the stem:
 * out [B,H/2,W/2,Cout] = avgpool2(act(conv)) (H, W even); pool == 2 keeps act(conv)[:, ::2, ::2] instead, i.e. the conv
 * runs with STRIDE 2 (torchvision Bottleneck.conv2).  Replaces nn.Conv2d(k=3, padding=1) + folded
 * BatchNorm + ReLU (+ AvgPool2d) of clip/model.py Bottleneck / ModifiedResNet stem. */
int embclip_conv3x3_f16(const void* in, const void* w, const float* bias, void* out, int B, int H, int W,
                        int Cin, int Cout, int relu, int pool, void* stream);
/* bneck_tail for the LAST block of a stage: the same two fused convs, but instead of x' the launch writes pool(x'), the input of
 * the next stage's downsample branch (clip/model.py Bottleneck.downsample [UPSTREAM]: AvgPool2d(2) before the 1x1 conv;
 * torchvision's stride-2 1x1 conv reads x'[:, ::2, ::2]).  The M rows are pixels of images `width` wide (width even, <= 64;
 * M a multiple of 2 * width); pool_mode 1 = 2x2 average, 2 = top-left pixel of each window; pool_out is [M / 4, 256] fp16.
 * Identity residual and n1 == 128 only. */
int embclip_bneck_tail_pool_f16(const void* y2, const void* w3, const float* b3, const void* residual, void* pool_out, int pool_mode,
                                int width, const void* w1, const float* b1, void* y1, int64_t M, int n1, void* stream);
/* The same fusion with STREAMED weights, for stages whose conv3 / next-conv1 matrices do not fit shared memory (layer 2):
 *   out[M,N3] = relu(y2[M,K3] . w3[N3,K3]^T + b3 + residual[M,N3]);   y1[M,n1] = relu(out . w1[n1,N3]^T + b1)
 * Built for K3 = 128, n1 = 128, N3 a multiple of 64 in [256, 1024]. */
int embclip_bneck_tail_stream_f16(const void* y2, const void* w3, const float* b3, const void* residual, void* out, const void* w1,
                                  const float* b1, void* y1, int64_t M, int K3, int N3, int n1, void* stream);
/* 2x2-window pools on NHWC fp16 [B,H,W,C] -> [B,H/2,W/2,C] (H, W even, C % 8 == 0).  mode 1: nn.AvgPool2d(2) (CLIP's
 * anti-aliasing pool); 2: x[:, ::2, ::2] (what a stride-2 1x1 conv reads); 3: nn.MaxPool2d(3, stride 2, padding 1)
 * (torchvision ResNet stem). */
int embclip_pool2_f16(const void* in, void* out, int B, int H, int W, int C, int mode, void* stream);
/* Bottleneck tail + next head in one pass (clip/model.py Bottleneck.forward [UPSTREAM]: out = relu(bn3(conv3(y2)) + identity),
 * then the next block's relu(bn1(conv1(out)))):
 *   out[M,256] = relu([y2 | x0] . w3^T + b3 (+ residual));   y1[M,n1] = relu(out . w1^T + b1)
 * y2 fp16 [M,64]; exactly one of x0 (fp16 [M,64]: downsample conv K-concatenated, w3 = [256,128]) and residual (fp16 [M,256],
 * w3 = [256,64]) is non-null; w1 fp16 [n1,256], n1 = 64 or 128; biases fp32 (BN folded). */
int embclip_bneck_tail_f16(const void* y2, const void* x0, const void* w3, const float* b3, const void* residual, void* out,
                           const void* w1, const float* b1, void* y1, int64_t M, int n1, void* stream);
/* 2x2 average pool, NHWC fp16 (nn.AvgPool2d(2) of clip/model.py). */
int embclip_avgpool2_f16(const void* in, void* out, int B, int H, int W, int C, void* stream);
/* Stem conv1: fp32 NHWC [B,R,R,3] -> fp16 NHWC [B,R/2,R/2,Cout]; w fp32 [27, Cout] (kh,kw,cin major), bias fp32. */
int embclip_stem_conv1(const float* frames, const float* w, const float* bias, void* out, int B, int R, int Cout,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * ObjectNav actor-critic + PPO update: replaces, on the update hot path,
 *   ResnetTensorNavActorCritic.forward   (allenact projects/objectnav_baselines/models/object_nav_models.py)
 *   RNNStateEncoder.forward              (allenact/embodiedai/models/basic_models.py)
 *   LinearActorHead / LinearCriticHead   (allenact/algorithms/onpolicy_sync/policy.py)
 *   PPO.loss_per_step                    (allenact/algorithms/onpolicy_sync/losses/ppo.py)
 *   RolloutStorage.compute_returns       (allenact/algorithms/onpolicy_sync/storage.py)
 *   OnPolicyTrainer.backprop_step: total_loss.backward(), clip_grad_norm_, Adam.step  (.../engine.py)
 * of allenai/allenact v0.5.0 (pin: /root/reference/readme_files/baselines_robothor_objectnav.md:6; the
 * experiment that instantiates them is named at :51).  All tensors are [steps T, samplers N, ...] row-major.
 * ------------------------------------------------------------------------------------------------ */
typedef struct embclip_ac* embclip_ac_t;

typedef struct {
  int32_t feat_channels;    /* 2048: channels of the CLIP-RN50 trunk feature            */
  int32_t feat_pixels;      /* 49 = 7 x 7                                                */
  int32_t compress_hidden;  /* 128  resnet_compressor.0 out                              */
  int32_t compress_out;     /* 32   resnet_compressor.2 out                              */
  int32_t goal_dims;        /* 32   embed_class width                                    */
  int32_t combine_hidden;   /* 128  target_obs_combiner.0 out                            */
  int32_t combine_out;      /* 32   target_obs_combiner.2 out; GRU input = 32 * 49       */
  int32_t hidden;           /* 512  GRU hidden size                                      */
  int32_t num_actions;      /* 6                                                         */
  int32_t num_goals;        /* 12                                                        */
  int32_t trainable_masked_hidden_state;  /* 0 (the ObjectNav config): an episode starts from h = 0.  1: from the learned
                             * parameter "state_encoder.init_hidden_state" [1,1,hidden] (RNNStateEncoder kwarg of the same name) */
} embclip_ac_cfg;

int embclip_ac_create(const embclip_ac_cfg* cfg, embclip_ac_t* out);   /* needs no GPU */
int embclip_ac_destroy(embclip_ac_t h);
/* Parameters live in ONE flat fp32 buffer (and gradients / Adam moments in buffers of the same layout):
 * param_info(i).name is the upstream state_dict key, .offset/.nbytes its slot (256-B aligned, padding zero). */
int embclip_ac_num_params(embclip_ac_t h);
int embclip_ac_param_info(embclip_ac_t h, int index, embclip_param_info* out);
uint64_t embclip_ac_param_floats(embclip_ac_t h);
uint64_t embclip_ac_workspace_bytes(embclip_ac_t h, int T, int N);
/* Introspection for layer-by-layer parity tests: the intermediates of the last forward / backward on a [T, N]
 * block live in the workspace as row-major 2-D tensors (info.w rows x info.c columns, info.offset bytes in). */
int embclip_ac_num_acts(embclip_ac_t h);
int embclip_ac_act_info(embclip_ac_t h, int T, int N, int index, embclip_act_info* out);
/* Rollout features fp32 [frames, C, H*W] (what ClipResNetPreprocessor wrote into RolloutStorage) -> fp16
 * [frames*H*W, C] rows, the layout every update pass reads.  Once per rollout. */
int embclip_ac_pack_features(embclip_ac_t h, const float* feats_nchw, long long frames, void* feats_f16, void* stream);
/* ActorCriticModel.forward on a [T, N] block: goals int64 [T,N]; masks fp32 [T,N] (0 = episode start);
 * h0 fp32 [N, hidden] -> logits fp32 [T,N,A], values fp32 [T,N], h_last fp32 [N, hidden] (may be NULL).
 * save_for_backward != 0 keeps what embclip_ac_backward needs in the workspace. */
int embclip_ac_forward(embclip_ac_t h, const float* params, const void* feats_f16, const long long* goals,
                       const float* masks, const float* h0, int T, int N, float* logits, float* values, float* h_last,
                       void* workspace, uint64_t workspace_bytes, int save_for_backward, void* stream);
/* One rollout step (OnPolicyRLEngine.act: actor_critic(...) with steps = 1, then distributions.sample() and log_prob):
 * embclip_ac_forward(T = 1, no save) followed by CategoricalDistr sampling by inverse CDF on `uniforms` fp32 [N] in [0,1).
 * -> actions int64 [N], action_log_probs fp32 [N], values fp32 [N], h_out fp32 [N, hidden], logits fp32 [N, A].
 * params_version: 0 = always rebuild the fp16 weight layouts; otherwise the caller promises that equal versions mean equal
 * parameter values, and consecutive calls with the same (version, params, workspace, N) reuse the layouts in the workspace. */
int embclip_ac_act(embclip_ac_t h, const float* params, uint64_t params_version, const void* feats_f16, const long long* goals,
                   const float* masks, const float* h0, int N, const float* uniforms, long long* actions,
                   float* action_log_probs, float* values, float* h_out, float* logits, void* workspace,
                   uint64_t workspace_bytes, void* stream);
/* PPO.loss_per_step on the block embclip_ac_forward just evaluated (reads its hidden states from the workspace):
 * loss_sums[3] <- sums over the block of {action loss, value loss, entropy}; the gradient of
 *   grad_scale * sum(action + value_loss_coef * value - entropy_coef * entropy)
 * w.r.t. logits / values is left in the workspace for embclip_ac_backward(dlogits = NULL). */
int embclip_ac_ppo_loss(embclip_ac_t h, const float* params, int T, int N, const long long* actions,
                        const float* old_action_log_probs, const float* norm_adv, const float* old_values,
                        const float* returns, float clip_param, float value_loss_coef, float entropy_coef,
                        float grad_scale, float* logits, float* values, float* loss_sums, void* workspace,
                        uint64_t workspace_bytes, void* stream);
/* Backward of embclip_ac_forward(save_for_backward = 1): ACCUMULATES into `grads` (flat, caller zeroes it).
 * dlogits [T,N,A] / dvalues [T,N] / dh_last [N,hidden] are the output gradients (autograd path); pass
 * dlogits = dvalues = NULL to use the ones embclip_ac_ppo_loss left in the workspace. */
int embclip_ac_backward(embclip_ac_t h, const float* params, const void* feats_f16, const long long* goals,
                        const float* masks, const float* h0, int T, int N, const float* dlogits, const float* dvalues,
                        const float* dh_last, float* grads, void* workspace, uint64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CLIP transformer towers + cosine-similarity logits (zero-shot ObjectNav path, BASELINE.json config 5):
 * replaces VisionTransformer.forward (ViT-B/32), CLIP.encode_text and CLIP.forward of openai/CLIP clip/model.py
 * (pin: /root/reference/primitive_probing/environment.yml:22; used by the zeroshot-objectnav branch named at
 * /root/reference/readme_files/zeroshot_objectnav.md:5,17,27).  ResidualAttentionBlock = LN -> QKV -> 12-head
 * attention over 50 (image) / 77 (text, causal) tokens -> out-proj -> +res -> LN -> fc -> QuickGELU -> proj -> +res.
 * ------------------------------------------------------------------------------------------------ */
typedef struct embclip_tf* embclip_tf_t;
#define EMBCLIP_TF_VISION 0
#define EMBCLIP_TF_TEXT 1
typedef struct {
  int32_t kind;              /* EMBCLIP_TF_VISION or EMBCLIP_TF_TEXT                           */
  int32_t width;             /* 768 (ViT-B/32) / 512 (text)                                    */
  int32_t layers;            /* 12                                                             */
  int32_t heads;             /* width / 64                                                     */
  int32_t output_dim;        /* 512                                                            */
  int32_t patch_size;        /* vision: 32                                                     */
  int32_t input_resolution;  /* vision: 224                                                    */
  int32_t context_length;    /* text: 77                                                       */
  int32_t vocab_size;        /* text: 49408                                                    */
} embclip_tf_cfg;
int embclip_tf_create(const embclip_tf_cfg* cfg, embclip_tf_t* out);   /* needs no GPU */
int embclip_tf_destroy(embclip_tf_t h);
int embclip_tf_num_params(embclip_tf_t h);
int embclip_tf_param_info(embclip_tf_t h, int index, embclip_param_info* out);
uint64_t embclip_tf_blob_bytes(embclip_tf_t h);
int embclip_tf_bind_weights(embclip_tf_t h, const void* device_blob, uint64_t nbytes);
uint64_t embclip_tf_workspace_bytes(embclip_tf_t h, int batch);
int embclip_tf_launches_per_forward(embclip_tf_t h);
/* CLIP.encode_image for the ViT tower: frames fp32 NHWC [batch, R, R, 3] (mean/std normalised) -> fp32 [batch, output_dim]. */
int embclip_vit_forward(embclip_tf_t h, const float* frames_nhwc, int batch, float* out, void* workspace,
                        uint64_t workspace_bytes, void* stream);
/* CLIP.encode_text: token ids int64 [prompts, context_length] -> fp32 [prompts, output_dim] (row at argmax(ids) = EOT). */
int embclip_text_forward(embclip_tf_t h, const long long* token_ids, int prompts, float* out, void* workspace,
                         uint64_t workspace_bytes, void* stream);
/* CLIP.forward's logits_per_image: exp(logit_scale) * normalize(image_features) @ normalize(text_features)^T -> [batch, prompts]. */
int embclip_clip_logits(const float* image_features, const float* text_features, int batch, int prompts, int embed_dim,
                        float logit_scale, float* logits, void* stream);

/* RolloutStorage.compute_returns(use_gae): rewards [T,N]; values [T+1,N] (row T = next value); masks [T+1,N]
 * -> returns [T,N], advantages [T,N] and (if non-NULL) norm_advantages = (A - mean) / (std + eps). */
int embclip_gae(const float* rewards, const float* values, const float* masks, int T, int N, float gamma, float tau,
                float* returns, float* advantages, float* norm_advantages, float eps, void* stream);
/* out[0] = sum x^2 (the squared global gradient norm of clip_grad_norm_), DETERMINISTIC: the same input gives the same bits on
 * every launch / rank (block partials are combined in index order, not arrival order), so data-parallel replicas that hold the
 * same all-reduced gradient apply the same clip coefficient.  `out` must hold EMBCLIP_SUMSQ_FLOATS floats ([1 ..] is scratch). */
#define EMBCLIP_SUMSQ_FLOATS 1024
int embclip_sumsq_f32(const float* x, long long n, float* out, void* stream);
/* clip_grad_norm_(max_grad_norm) from *grad_sumsq (skipped when max_grad_norm <= 0), then torch.optim.Adam
 * (no weight decay / amsgrad) step number `step` (1-based) on flat buffers; grads are scaled in place. */
int embclip_adam_clip_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                           const float* grad_sumsq, float max_grad_norm, float lr, float beta1, float beta2, float eps,
                           int step, void* stream);

/* Primitives of the plan above, exposed for unit tests. */
/* out[m*ldo_m + n*ldo_n] += alpha * sum_k A[k][m] * B[k][n]; A fp16 [Kdim][lda], B fp16 [Kdim][ldb]; N1 % 32 == 0;
 * alpha = device scalar or NULL.  (weight gradients: both operands row-major activations, contraction over rows) */
int embclip_wgrad_f16(const void* a, int lda, int M1, const void* b, int ldb, int N1, long long Kdim, float* out,
                      long long ldo_m, long long ldo_n, const float* alpha, void* stream);
/* nn.GRU (1 layer) with RNNStateEncoder's episode masking: gi = x W_ih^T + b_ih precomputed [T,N,3H];
 * out [T,N,H]; save_* [T,N,H] (all NULL for inference); scratch32 = 256 B of device scratch
 * (uint32 [0,32): one barrier counter per sampler group, zeroed by the call; [32]: max |dgi| bits written by the backward).
 * h_init fp32 [H] or NULL: the state an episode starts from where masks == 0 (trainable_masked_hidden_state); the backward
 * ADDS its gradient into dh_init [H] (both NULL or both set). */
int embclip_gru_forward(const float* gi, const float* w_hh, const float* b_hh, const float* h0, const float* masks,
                        const float* h_init, int T, int N, int H, float* out, float* save_r, float* save_z, float* save_n,
                        float* save_hn, void* scratch32, void* stream);
/* Launch geometry of the two GRU kernels for (N samplers, hidden H) on the current device: out5 = {CTAs per thread-block cluster
 * (H / 32; 0 = the cooperative grid-barrier kernels run instead), clusters, samplers per cluster, resident clusters (forward),
 * resident clusters (backward)}. */
int embclip_gru_geometry(int N, int H, int* out5);
/* BPTT of the above: dout [T,N,H], dh_last [N,H] or NULL -> dgi, dgh [T,N,3H] fp32, hm_f16 [T,N,H] fp16
 * (masked previous hidden state), dh0 [N,H] or NULL. */
int embclip_gru_backward(const float* w_hh, const float* h0, const float* masks, const float* out, const float* save_r,
                         const float* save_z, const float* save_n, const float* save_hn, const float* dout,
                         const float* dh_last, const float* h_init, int T, int N, int H, float* dgi, float* dgh, void* hm_f16,
                         float* dh0, float* dh_init, void* scratch32, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EMBCLIP_B200_H_ */
