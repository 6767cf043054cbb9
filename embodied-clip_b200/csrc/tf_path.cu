// libembclip_b200.so -- CLIP transformer towers (ViT-B/32 image tower, causal text tower) and the cosine-similarity
// logits of CLIP.forward: the zero-shot ObjectNav path (BASELINE.json config 5).
// Every GEMM is conv_gemm_kernel (tcgen05); the fp32 residual stream is carried by its fp32-residual epilogue.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host.h"
#include "tf_kernels.cuh"

using namespace embclip;

namespace {
struct TfBlock { int ln1w, ln1b, qkvw, qkvb, outw, outb, ln2w, ln2b, fcw, fcb, projw, projb; };
}

struct embclip_tf {
  embclip_tf_cfg cfg;
  std::vector<embclip_param_info> params;
  uint64_t blob_bytes = 0;
  const uint8_t* blob = nullptr;
  std::vector<TfBlock> blocks;
  int p_patch = -1, p_cls = -1, p_pos = -1, p_lnpre_w = -1, p_lnpre_b = -1, p_lnpost_w = -1, p_lnpost_b = -1, p_head = -1, p_tok = -1;
  int L = 0;       // tokens per sequence
};

static int tf_add_param(embclip_tf* m, const std::string& name, int dtype, std::initializer_list<int64_t> shape) {
  embclip_param_info pi;
  memset(&pi, 0, sizeof pi);
  snprintf(pi.name, sizeof pi.name, "%s", name.c_str());
  pi.dtype = dtype;
  pi.ndim = (int)shape.size();
  uint64_t n = 1;
  int i = 0;
  for (int64_t s : shape) { pi.shape[i++] = s; n *= (uint64_t)s; }
  pi.nbytes = n * (dtype == EMBCLIP_DTYPE_F16 ? 2 : 4);
  pi.offset = m->blob_bytes;
  m->blob_bytes += (pi.nbytes + 255) & ~uint64_t(255);
  m->params.push_back(pi);
  return (int)m->params.size() - 1;
}

extern "C" int embclip_tf_create(const embclip_tf_cfg* cfg, embclip_tf_t* out) {
  if (!cfg || !out) return fail(EMBCLIP_EINVAL, "tf_create: null argument");
  const embclip_tf_cfg& c = *cfg;
  if (c.kind != EMBCLIP_TF_VISION && c.kind != EMBCLIP_TF_TEXT) return fail(EMBCLIP_EINVAL, "tf_create: kind must be vision (0) or text (1)");
  if (c.width != 512 && c.width != 768) return fail(EMBCLIP_EINVAL, "tf_create: width %d not built (512 and 768 are)", c.width);
  if (c.heads <= 0 || c.width / c.heads != 64 || c.width % c.heads) return fail(EMBCLIP_EINVAL, "tf_create: head dim must be 64");
  if (c.layers < 1 || c.output_dim <= 0 || c.output_dim % 32) return fail(EMBCLIP_EINVAL, "tf_create: bad layers / output_dim");
  int L;
  if (c.kind == EMBCLIP_TF_VISION) {
    if (c.patch_size <= 0 || c.input_resolution % c.patch_size || (c.patch_size * c.patch_size * 3) % 64)
      return fail(EMBCLIP_EINVAL, "tf_create: bad patch size / resolution");
    const int g = c.input_resolution / c.patch_size;
    L = g * g + 1;
  } else {
    if (c.context_length <= 0 || c.vocab_size <= 0) return fail(EMBCLIP_EINVAL, "tf_create: text tower needs context_length and vocab_size");
    L = c.context_length;
  }
  if (L > kAttnMaxL) return fail(EMBCLIP_EINVAL, "tf_create: %d tokens per sequence exceed the attention kernel's %d", L, kAttnMaxL);
  embclip_tf* m = new embclip_tf();
  m->cfg = c;
  m->L = L;
  const int64_t D = c.width;
  if (c.kind == EMBCLIP_TF_VISION) {
    m->p_patch = tf_add_param(m, "patch.w", EMBCLIP_DTYPE_F16, {D, (int64_t)c.patch_size * c.patch_size * 3});
    m->p_cls = tf_add_param(m, "cls", EMBCLIP_DTYPE_F32, {D});
    m->p_pos = tf_add_param(m, "pos", EMBCLIP_DTYPE_F32, {L, D});
    m->p_lnpre_w = tf_add_param(m, "ln_pre.w", EMBCLIP_DTYPE_F32, {D});
    m->p_lnpre_b = tf_add_param(m, "ln_pre.b", EMBCLIP_DTYPE_F32, {D});
  } else {
    m->p_tok = tf_add_param(m, "tok_emb", EMBCLIP_DTYPE_F32, {c.vocab_size, D});
    m->p_pos = tf_add_param(m, "pos", EMBCLIP_DTYPE_F32, {L, D});
  }
  for (int i = 0; i < c.layers; ++i) {
    const std::string P = "blk" + std::to_string(i);
    TfBlock b;
    b.ln1w = tf_add_param(m, P + ".ln1.w", EMBCLIP_DTYPE_F32, {D});
    b.ln1b = tf_add_param(m, P + ".ln1.b", EMBCLIP_DTYPE_F32, {D});
    b.qkvw = tf_add_param(m, P + ".qkv.w", EMBCLIP_DTYPE_F16, {3 * D, D});
    b.qkvb = tf_add_param(m, P + ".qkv.b", EMBCLIP_DTYPE_F32, {3 * D});
    b.outw = tf_add_param(m, P + ".out.w", EMBCLIP_DTYPE_F16, {D, D});
    b.outb = tf_add_param(m, P + ".out.b", EMBCLIP_DTYPE_F32, {D});
    b.ln2w = tf_add_param(m, P + ".ln2.w", EMBCLIP_DTYPE_F32, {D});
    b.ln2b = tf_add_param(m, P + ".ln2.b", EMBCLIP_DTYPE_F32, {D});
    b.fcw = tf_add_param(m, P + ".fc.w", EMBCLIP_DTYPE_F16, {4 * D, D});
    b.fcb = tf_add_param(m, P + ".fc.b", EMBCLIP_DTYPE_F32, {4 * D});
    b.projw = tf_add_param(m, P + ".proj.w", EMBCLIP_DTYPE_F16, {D, 4 * D});
    b.projb = tf_add_param(m, P + ".proj.b", EMBCLIP_DTYPE_F32, {D});
    m->blocks.push_back(b);
  }
  m->p_lnpost_w = tf_add_param(m, "ln_post.w", EMBCLIP_DTYPE_F32, {D});
  m->p_lnpost_b = tf_add_param(m, "ln_post.b", EMBCLIP_DTYPE_F32, {D});
  m->p_head = tf_add_param(m, "head.w", EMBCLIP_DTYPE_F16, {c.output_dim, D});
  *out = m;
  return 0;
}
extern "C" int embclip_tf_destroy(embclip_tf_t h) { delete h; return 0; }
extern "C" int embclip_tf_num_params(embclip_tf_t h) { return h ? (int)h->params.size() : fail(EMBCLIP_EINVAL, "null handle"); }
extern "C" int embclip_tf_param_info(embclip_tf_t h, int index, embclip_param_info* out) {
  if (!h || !out || index < 0 || index >= (int)h->params.size()) return fail(EMBCLIP_EINVAL, "tf_param_info: bad argument");
  *out = h->params[index];
  return 0;
}
extern "C" uint64_t embclip_tf_blob_bytes(embclip_tf_t h) { return h ? h->blob_bytes : 0; }
extern "C" int embclip_tf_bind_weights(embclip_tf_t h, const void* device_blob, uint64_t nbytes) {
  if (!h || !device_blob) return fail(EMBCLIP_EINVAL, "tf_bind_weights: null argument");
  if (nbytes < h->blob_bytes) return fail(EMBCLIP_EINVAL, "tf_bind_weights: blob too small");
  if (reinterpret_cast<uintptr_t>(device_blob) % 256) return fail(EMBCLIP_EINVAL, "tf_bind_weights: blob must be 256-B aligned");
  h->blob = reinterpret_cast<const uint8_t*>(device_blob);
  return 0;
}

namespace {

struct TfWs {
  float* x;          // residual stream [S*L][D]
  float* patch_out;  // vision: patch embeddings [S*(L-1)][D]
  __half* patches;   // vision: fp16 patch rows [S*(L-1)][ps*ps*3]
  __half* xn;        // LayerNorm output [S*L][D]
  __half* qkv;       // [S*L][3D]
  __half* attn;      // [S*L][D]
  __half* hid;       // [S*L][4D]
  __half* pooled;    // [S][D] ln_post of the pooled token
  long long* eot;    // text: [S] row of the end-of-text token
  uint64_t total;
};
uint64_t tf_carve(uint64_t& off, uint64_t bytes) {
  const uint64_t o = off;
  off += (bytes + 1023) & ~uint64_t(1023);
  return o;
}
void tf_workspace(const embclip_tf* m, int S, uint8_t* base, TfWs* w) {
  const uint64_t D = m->cfg.width, M = (uint64_t)S * m->L;
  uint64_t off = 0;
  w->x = reinterpret_cast<float*>(base + tf_carve(off, M * D * 4));
  if (m->cfg.kind == EMBCLIP_TF_VISION) {
    const uint64_t Mp = (uint64_t)S * (m->L - 1), Kp = (uint64_t)m->cfg.patch_size * m->cfg.patch_size * 3;
    w->patch_out = reinterpret_cast<float*>(base + tf_carve(off, Mp * D * 4));
    w->patches = reinterpret_cast<__half*>(base + tf_carve(off, Mp * Kp * 2));
  } else {
    w->patch_out = nullptr; w->patches = nullptr;
  }
  w->xn = reinterpret_cast<__half*>(base + tf_carve(off, M * D * 2));
  w->qkv = reinterpret_cast<__half*>(base + tf_carve(off, M * 3 * D * 2));
  w->attn = reinterpret_cast<__half*>(base + tf_carve(off, M * D * 2));
  w->hid = reinterpret_cast<__half*>(base + tf_carve(off, M * 4 * D * 2));
  w->pooled = reinterpret_cast<__half*>(base + tf_carve(off, (uint64_t)S * D * 2));
  w->eot = reinterpret_cast<long long*>(base + tf_carve(off, (uint64_t)S * 8));
  w->total = off;
}

template <typename T>
const T* TP(const embclip_tf* m, int id) { return reinterpret_cast<const T*>(m->blob + m->params[id].offset); }

int tf_gemm(const void* a, int K, const void* w, const float* bias, void* out, int rows, int N, int act, int out_f32, const float* res_f32,
            cudaStream_t st) {
  GemmOp g;
  g.a0 = a; g.n = 1; g.h = 1; g.w = rows; g.c0 = K; g.lda0 = K;
  g.wgt = w; g.ldw = K; g.w_rows = N;
  g.bias = bias; g.out = out; g.cout = N; g.relu = act; g.out_f32 = out_f32; g.res_f32 = res_f32;
  return launch_gemm(g, st);
}

int tf_layernorm(const embclip_tf* m, const float* x, int wid, int bid, __half* y, int rows, long long stride, const long long* gather,
                 cudaStream_t st) {
  const int D = m->cfg.width;
  const int blocks = (rows + 7) / 8;
  if (D == 768) layernorm_rows_kernel<6><<<blocks, 256, 0, st>>>(x, TP<float>(m, wid), TP<float>(m, bid), y, rows, D, stride, gather);
  else layernorm_rows_kernel<4><<<blocks, 256, 0, st>>>(x, TP<float>(m, wid), TP<float>(m, bid), y, rows, D, stride, gather);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int tf_attention(embclip_tf* m, const TfWs& w, int S, int causal, cudaStream_t st) {
  const int D = m->cfg.width, L = m->L, rows = S * L;
  static const bool cuda_core = getenv("EMBCLIP_ATTN_CUDA_CORE") != nullptr;      // first version, kept for A/B timing
  if (cuda_core || L > 128) {
    attention_kernel<<<dim3(S, m->cfg.heads), 128, 0, st>>>(w.qkv, w.attn, L, D, causal);
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  { const int rc_ = ensure_smem((const void*)attention_tc_kernel, (size_t)(kAttnTcSmem)); if (rc_) return rc_; }
  CUtensorMap tm;
  int rc;
  if ((rc = make_map_2d(&tm, w.qkv, rows, 3 * D, 3 * D, 64, 128))) return rc;
  AttnTcParams p;
  p.rows = rows; p.L = L; p.D = D; p.nseq = 128 / L; p.causal = causal; p.out = w.attn;
  const int tiles = (S + p.nseq - 1) / p.nseq;
  attention_tc_kernel<<<dim3(tiles, m->cfg.heads), 128, kAttnTcSmem, st>>>(tm, p);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// the 12 (or `layers`) ResidualAttentionBlocks on the fp32 stream w.x, then ln_post / ln_final of the pooled rows and the head GEMM
int tf_blocks_and_head(embclip_tf* m, const TfWs& w, int S, int causal, long long pool_stride, const long long* pool_rows, float* out,
                       cudaStream_t st) {
  const int D = m->cfg.width, M = S * m->L;
  int rc;
  for (const TfBlock& b : m->blocks) {
    if ((rc = tf_layernorm(m, w.x, b.ln1w, b.ln1b, w.xn, M, 1, nullptr, st))) return rc;
    if ((rc = tf_gemm(w.xn, D, TP<__half>(m, b.qkvw), TP<float>(m, b.qkvb), w.qkv, M, 3 * D, 0, 0, nullptr, st))) return rc;
    if ((rc = tf_attention(m, w, S, causal, st))) return rc;
    if ((rc = tf_gemm(w.attn, D, TP<__half>(m, b.outw), TP<float>(m, b.outb), w.x, M, D, 0, 1, w.x, st))) return rc;
    if ((rc = tf_layernorm(m, w.x, b.ln2w, b.ln2b, w.xn, M, 1, nullptr, st))) return rc;
    if ((rc = tf_gemm(w.xn, D, TP<__half>(m, b.fcw), TP<float>(m, b.fcb), w.hid, M, 4 * D, 2, 0, nullptr, st))) return rc;
    if ((rc = tf_gemm(w.hid, 4 * D, TP<__half>(m, b.projw), TP<float>(m, b.projb), w.x, M, D, 0, 1, w.x, st))) return rc;
  }
  if ((rc = tf_layernorm(m, w.x, m->p_lnpost_w, m->p_lnpost_b, w.pooled, S, pool_stride, pool_rows, st))) return rc;
  return tf_gemm(w.pooled, D, TP<__half>(m, m->p_head), nullptr, out, S, m->cfg.output_dim, 0, 1, nullptr, st);
}

int tf_check(embclip_tf* m, int kind, int S, const void* in, const void* out, const void* ws, uint64_t ws_bytes) {
  if (!m || !in || !out || !ws || S <= 0) return fail(EMBCLIP_EINVAL, "tf_forward: null argument or empty batch");
  if (m->cfg.kind != kind) return fail(EMBCLIP_EINVAL, "tf_forward: handle is a %s tower", m->cfg.kind == EMBCLIP_TF_VISION ? "vision" : "text");
  if (!m->blob) return fail(EMBCLIP_ESTATE, "tf_forward: weights not bound (call embclip_tf_bind_weights first)");
  if (reinterpret_cast<uintptr_t>(ws) % 1024) return fail(EMBCLIP_EINVAL, "tf_forward: workspace must be 1024-B aligned");
  if (S > 65535) return fail(EMBCLIP_EINVAL, "tf_forward: at most 65535 sequences per call");
  TfWs w;
  tf_workspace(m, S, nullptr, &w);
  if (ws_bytes < w.total) return fail(EMBCLIP_ENOSPC, "tf_forward: workspace %llu B < required %llu B", (unsigned long long)ws_bytes, (unsigned long long)w.total);
  return 0;
}

}  // namespace

extern "C" uint64_t embclip_tf_workspace_bytes(embclip_tf_t h, int batch) {
  if (!h || batch <= 0) return 0;
  TfWs w;
  tf_workspace(h, batch, nullptr, &w);
  return w.total;
}

extern "C" int embclip_vit_forward(embclip_tf_t h, const float* frames_nhwc, int batch, float* out, void* workspace,
                                   uint64_t workspace_bytes, void* stream) {
  EMBCLIP_TRACE();
  int rc;
  if ((rc = tf_check(h, EMBCLIP_TF_VISION, batch, frames_nhwc, out, workspace, workspace_bytes))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  TfWs w;
  tf_workspace(h, batch, reinterpret_cast<uint8_t*>(workspace), &w);
  const embclip_tf_cfg& c = h->cfg;
  const int D = c.width, L = h->L, Mp = batch * (L - 1), Kp = c.patch_size * c.patch_size * 3;
  long long nb = (long long)Mp;
  if (nb > (long long)num_sms() * 8) nb = (long long)num_sms() * 8;
  vit_patchify_kernel<<<(int)nb, 256, 0, st>>>(frames_nhwc, w.patches, batch, c.input_resolution, c.patch_size);
  CUDA_TRY(cudaGetLastError());
  if ((rc = tf_gemm(w.patches, Kp, TP<__half>(h, h->p_patch), nullptr, w.patch_out, Mp, D, 0, 1, nullptr, st))) return rc;
  const int rows = batch * L;
  if (D == 768)
    vit_embed_ln_pre_kernel<6><<<(rows + 7) / 8, 256, 0, st>>>(w.patch_out, TP<float>(h, h->p_cls), TP<float>(h, h->p_pos), TP<float>(h, h->p_lnpre_w),
                                                               TP<float>(h, h->p_lnpre_b), w.x, rows, L, D);
  else
    vit_embed_ln_pre_kernel<4><<<(rows + 7) / 8, 256, 0, st>>>(w.patch_out, TP<float>(h, h->p_cls), TP<float>(h, h->p_pos), TP<float>(h, h->p_lnpre_w),
                                                               TP<float>(h, h->p_lnpre_b), w.x, rows, L, D);
  CUDA_TRY(cudaGetLastError());
  return tf_blocks_and_head(h, w, batch, 0, L, nullptr, out, st);          // pooled row of image b = its class token, row b*L
}

extern "C" int embclip_text_forward(embclip_tf_t h, const long long* token_ids, int prompts, float* out, void* workspace,
                                    uint64_t workspace_bytes, void* stream) {
  EMBCLIP_TRACE();
  int rc;
  if ((rc = tf_check(h, EMBCLIP_TF_TEXT, prompts, token_ids, out, workspace, workspace_bytes))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  TfWs w;
  tf_workspace(h, prompts, reinterpret_cast<uint8_t*>(workspace), &w);
  const embclip_tf_cfg& c = h->cfg;
  const int D = c.width, L = h->L, rows = prompts * L;
  long long nb = ((long long)rows * (D / 4) + 255) / 256;
  if (nb > (long long)num_sms() * 16) nb = (long long)num_sms() * 16;
  text_embed_kernel<<<(int)nb, 256, 0, st>>>(token_ids, TP<float>(h, h->p_tok), TP<float>(h, h->p_pos), w.x, rows, L, D, c.vocab_size);
  text_eot_rows_kernel<<<(prompts + 127) / 128, 128, 0, st>>>(token_ids, w.eot, prompts, L);
  CUDA_TRY(cudaGetLastError());
  return tf_blocks_and_head(h, w, prompts, 1, 0, w.eot, out, st);
}

extern "C" int embclip_clip_logits(const float* image_features, const float* text_features, int batch, int prompts, int embed_dim,
                                   float logit_scale, float* logits, void* stream) {
  EMBCLIP_TRACE();
  if (!image_features || !text_features || !logits || batch <= 0 || prompts <= 0 || embed_dim <= 0)
    return fail(EMBCLIP_EINVAL, "clip_logits: bad argument");
  const int total = batch * prompts;
  clip_logits_kernel<<<(total + 7) / 8, 256, 0, (cudaStream_t)stream>>>(image_features, text_features, logits, batch, prompts, embed_dim,
                                                                       expf(logit_scale));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int embclip_tf_launches_per_forward(embclip_tf_t h) {
  if (!h) return fail(EMBCLIP_EINVAL, "null handle");
  return (h->cfg.kind == EMBCLIP_TF_VISION ? 3 : 2) + 7 * (int)h->blocks.size() + 2;
}
