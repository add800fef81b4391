// Moshi LM handle: weights (reference state-dict names), streaming state, and the per-frame decode
// step (LMGen._step, lm.py:668-783): token ring -> Temporal transformer -> text sample ->
// Depformer (dep_q dependent sub-steps) -> token ring.
#include "gemm_sk.cuh"
#include "lm_kernels.cuh"

using namespace b200;
using namespace b200::lm;

namespace {

struct TLayer {
  const bf16 *in_w, *out_w, *n1, *n2, *lin_in, *lin_out;
  const float *in_s = nullptr, *out_s = nullptr, *lin_in_s = nullptr, *lin_out_s = nullptr;   // int8 path: weight-row scales
  bf16 *kc = nullptr, *vc = nullptr;
  uint8_t *kc8 = nullptr, *vc8 = nullptr;      // opt-in fp8 ring: e4m3 bytes [B][H][cap][D] ...
  float *ks = nullptr, *vs = nullptr;          // ... and one scale per (session, head, slot)
};
struct DLayer {
  std::vector<const bf16*> in_w, out_w, lin_in, lin_out;   // one per depformer step
  std::vector<const float*> in_s, out_s, lin_in_s, lin_out_s;   // int8 path: weight-row scales
  const bf16 *n1, *n2;
  bf16 *kc = nullptr, *vc = nullptr;
};

}  // namespace

struct b200_lm {
  b200_lm_config cfg;
  TensorStore store;
  Arena weights, state;
  bool finalized = false;
  int device = -1;                     // CUDA device the handle lives on (made current by every entry point)
  int batch = 0;                       // sessions (LMGen batch)
  int MB = 0;                          // rows the model runs on: batch, or 2 * batch under classifier-free guidance (lm.py:646-647)
  cudaStream_t stream = nullptr;       // caller's stream (inputs/outputs are ordered on it)
  cudaStream_t gstream = nullptr;      // private stream: graphs cannot be captured on the legacy default stream
  cudaStream_t body = nullptr;         // stream the step body is currently being enqueued on
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  int Kc = 0, max_delay = 0, CT = 0;
  // sampling (LMGen defaults, lm.py:556-571)
  int use_sampling = 1, top_k = 250, top_k_text = 25;
  float temp = 0.8f, temp_text = 0.7f;
  int pdl = 1;                                 // programmatic dependent launch of the GEMMs (B200_PDL=0 disables)
  int sk_smem = 0;                             // bytes of pipeline stages per GEMM CTA (B200_SK_SMEM_KB; 0 = 200 KB, one CTA per SM)
  float* sk_ws = nullptr;                      // stream-K partial-accumulator slots (L2-resident)
  int* sk_counters = nullptr;                  // per-tile arrival counters (zero between launches)
  // weights
  EmbedTables emb;
  std::vector<TLayer> layers;
  std::vector<DLayer> dlayers;
  const bf16 *out_norm = nullptr, *text_linear = nullptr;
  const float *text_linear_s = nullptr, *dep_in_s = nullptr;
  std::vector<const float*> dep_heads_s;
  int8_t* xq = nullptr; float* xq_scale = nullptr;     // int8 path: the current GEMM's quantised activations [B][max K] and row scales
  bf16* dep_in_all = nullptr;                  // [dep_q * dd][dim]
  std::vector<const bf16*> dep_tables;         // [0] = depformer_text_emb, [k] = depformer_emb[k-1]
  std::vector<const bf16*> dep_heads;          // linears[k]
  int* delays_dev = nullptr;
  // state
  long long *cache = nullptr, *offsets = nullptr, *pos = nullptr;
  uint8_t* exec_mask = nullptr;                // [batch] LMGen's mask
  uint8_t* exec_mask_m = nullptr;              // [MB] the model's mask (= exec_mask, or exec_mask.repeat(2) with CFG, lm.py:658-661)
  long long offset_cpu = 0;
  // classifier-free guidance (lm.py:596-604, 714-732, 820-833) and conditioning by sum (lm.py:398-399, 616-626)
  float cfg_coef = 1.f; int cfg_is_no_text = 0;
  std::vector<long long> cfg_until_host; long long* cfg_until = nullptr;
  bf16 *text_logits_cfg = nullptr, *dep_logits_cfg = nullptr;       // guided logits [batch][card] / [dep_q][batch][card]
  bf16* cond_sum = nullptr; int cond_on = 0;                          // [MB][dim]
  // extra heads of the STT models (lm.py:224-226, 793-807)
  std::vector<const bf16*> extra_w; const bf16** extra_w_dev = nullptr; bf16* extra_out = nullptr;
  // Exp(1) noise drawn inside the step (when the caller passes none): Philox keyed by (seed, step counter)
  unsigned long long noise_seed = 0x9E3779B97F4A7C15ull; unsigned long long* noise_ctr = nullptr; int noise_internal = 0;
  int* err = nullptr;                          // device error flags (lm::ERR_*)
  // activations
  long long *in_codes = nullptr, *input_tokens = nullptr, *text_token = nullptr, *audio_tokens = nullptr, *out_tokens = nullptr;
  float* noise = nullptr;
  bf16 *x = nullptr, *xn = nullptr, *qkv = nullptr, *ao = nullptr, *hbuf = nullptr, *tout = nullptr;
  bf16 *text_logits = nullptr, *din = nullptr, *dx = nullptr, *dxn = nullptr, *dqkv = nullptr, *dao = nullptr,
       *dh = nullptr, *dep_logits = nullptr;
  float* attn_part = nullptr;
  int* attn_counters = nullptr;                // split arrival counters [B*H]
  tc::DepFused* depf = nullptr;                // the depformer of a frame as one persistent kernel (B200_DEP_FUSED=0: off)
  int dep_fused = 1;
  int fuse_norm = 1;                           // B200_FUSE_NORM=0: keep rmsnorm_kernel in front of the GEMV path too (diagnostics)
  int kv_cap = 0;                              // b200_lm_set_kv_capacity: slots per ring (0 = cfg.context, the reference's ring)
  int kv_fp8 = 0;                              // b200_lm_set_kv_dtype / B200_KV_DTYPE: opt-in 8-bit KV ring (B200_KV_FP8_E4M3 or B200_KV_INT8; 0 = bf16)
  float *dep_part0 = nullptr, *dep_part1 = nullptr;
  unsigned* dep_bar = nullptr;
  unsigned long long* dep_trace = nullptr;     // B200_DEP_TRACE=1: barrier timestamps of the fused depformer (diagnostics)
  int nsplit = 1;
  int n_in_static = 0;
  int noise_static = 0;                        // 1: the captured step draws its own noise
  int replace_static = 0;                      // 1: this step's audio tokens are given (depformer_replace_tokens, lm.py:751-755)
  long long* replace_tokens = nullptr;         // [B][dep_q] staging of those tokens
  // graph
  int graph_enabled = 1;
  cudaGraphExec_t graph_exec = nullptr;
  int64_t graph_kernels = 0;                   // kernel nodes in the captured step (for b200_launch_count)
  // host staging
  long long *pin_in = nullptr, *pin_out = nullptr;
  float* pin_noise = nullptr;
  int64_t weight_bytes = 0;
  tc::GemmPlanCache plans;
};

namespace {

// slots per temporal KV ring: the reference's context unless b200_lm_set_kv_capacity shortened it
inline int ring_cap(const b200_lm* h) { return h->kv_cap > 0 && h->kv_cap < h->cfg.context ? h->kv_cap : h->cfg.context; }

int get_bf16(b200_lm* h, const std::string& name, std::vector<int64_t> shape, const bf16** out) {
  const Tensor* t = h->store.find(name);
  if (!t) B200_FAIL(B200_ERR_MISSING, "lm finalize: tensor '%s' was not loaded", name.c_str());
  if (t->dtype != B200_BF16) B200_FAIL(B200_ERR_SHAPE, "lm tensor '%s' must be bfloat16", name.c_str());
  if (t->shape != shape) {
    std::string got;
    for (auto s : t->shape) got += std::to_string(s) + ",";
    B200_FAIL(B200_ERR_SHAPE, "lm tensor '%s' has shape [%s]", name.c_str(), got.c_str());
  }
  *out = static_cast<const bf16*>(t->data);
  h->weight_bytes += t->numel() * 2;
  return B200_OK;
}

bool cfg_on(const b200_lm* h) { return h->cfg_coef != 1.f; }

int noise_per_row(const b200_lm* h) {
  const int kt = h->top_k_text < h->cfg.text_card ? h->top_k_text : h->cfg.text_card;
  const int ka = h->top_k < h->cfg.card ? h->top_k : h->cfg.card;
  return kt + h->cfg.dep_q * ka;
}

// y[M][N] = epi(x[M][K] . w[N][K]^T); `w` is the packed-tile form of gemm_sk.cu.
int linear(b200_lm* h, const bf16* x, long long ldx, const bf16* w, bf16* y, long long ldy, const bf16* res,
           long long ldr, int M, int N, int K, int epi, int gate_rows, const float* w_scales = nullptr,
           const bf16* norm_alpha = nullptr) {
  tc::SkTuning t;
  t.pdl = h->pdl;
  t.smem_budget = h->sk_smem;
  t.norm_alpha = norm_alpha;
  if (h->cfg.quantize) {       // QLinear.forward (utils/quantize.py:22-40): row-wise int8 activations, int8 x int8 -> int32
    if (!w_scales) B200_FAIL(B200_ERR_STATE, "quantised LM: linear without weight scales");
    B200_TRY(tc::sk_quantize_rows(x, ldx, h->xq, h->xq_scale, M, K, h->body, h->pdl));
    t.xq = h->xq; t.sa = h->xq_scale; t.sw = w_scales;
  }
  return tc::sk_linear(h->plans, x, ldx, w, y, ldy, res, ldr, M, N, K, epi, gate_rows, h->sk_ws, h->sk_counters, t, h->body);
}

// y = epi(rmsnorm(x, alpha) . w^T).  One or two sessions: the norm is folded into the GEMV's activation staging (one kernel
// and one dependency less per linear); otherwise rmsnorm_kernel writes xn and the GEMM reads it.
int norm_linear(b200_lm* h, const bf16* x, const bf16* alpha, bf16* xn, const bf16* w, bf16* y, long long ldy, int M, int N, int K,
                int epi, int gate_rows, const float* w_scales) {
  if (!h->cfg.quantize && h->fuse_norm && M <= tc::sk_gemv_max_m())
    return linear(h, x, K, w, y, ldy, nullptr, 0, M, N, K, epi, gate_rows, w_scales, alpha);
  B200_LAUNCH(rmsnorm_kernel, M, 256, 0, h->body, x, alpha, xn, K, 1e-8f);
  return linear(h, xn, K, w, y, ldy, nullptr, 0, M, N, K, epi, gate_rows, w_scales);
}

// A QLinear as model.q8.safetensors stores it (utils/quantize.py:16-21): "<name>" int8 [rows][K] + "<name>_scb" fp32 [rows]
int get_prequantised(b200_lm* h, const std::string& name, int rows, int K, const int8_t** q, const float** scb) {
  const Tensor* t = h->store.find(name);
  const Tensor* s = h->store.find(name + "_scb");
  if (!h->cfg.quantize) B200_FAIL(B200_ERR_INVALID, "lm tensor '%s' is int8 but the model was not built with quantize=True", name.c_str());
  if (!s || s->dtype != B200_F32) B200_FAIL(B200_ERR_MISSING, "lm finalize: '%s_scb' (float32 row scales of the int8 weight) was not loaded; "
                                            "care should be taken not to change its dtype (utils/quantize.py:31-35)", name.c_str());
  if (t->shape != std::vector<int64_t>{rows, K} || s->shape != std::vector<int64_t>{rows})
    B200_FAIL(B200_ERR_SHAPE, "lm tensor '%s' / its _scb: expected [%d, %d] / [%d]", name.c_str(), rows, K, rows);
  *q = static_cast<const int8_t*>(t->data);
  *scb = static_cast<const float*>(s->data);
  return B200_OK;
}

// Linear weight from the store -> packed tiles (gemm_sk.cu), which replace the row-major
// tensor (which is released), so the 15.4 GB checkpoint is resident once.
int get_linear(b200_lm* h, const std::string& name, int N, int K, int epi, int gate_rows, const bf16** out,
               const float** scales_out) {
  const int w_rows = epi == LIN_GATE ? 2 * gate_rows : N;
  *scales_out = nullptr;
  void* packed = nullptr;
  const Tensor* t = h->store.find(name);
  if (t && t->dtype == B200_I8) {      // pre-quantised checkpoint: tile CB as it is, keep SCB
    const int8_t* q = nullptr; const float* scb = nullptr;
    B200_TRY(get_prequantised(h, name, w_rows, K, &q, &scb));
    float* scales = nullptr;
    B200_TRY(h->weights.alloc(&packed, tc::sk_packed_bytes_i8(N, K, epi, gate_rows), false));
    B200_TRY(h->weights.alloc_t(&scales, (size_t)w_rows, false));
    B200_TRY(tc::sk_pack_weights_i8(q, packed, N, K, epi, gate_rows, nullptr));
    B200_CUDA(cudaMemcpy(scales, scb, (size_t)w_rows * 4, cudaMemcpyDeviceToDevice));
    *scales_out = scales;
    h->weight_bytes += (int64_t)w_rows * K + (int64_t)w_rows * 4;
    B200_CUDA(cudaStreamSynchronize(nullptr));
    h->store.release(name);
    h->store.release(name + "_scb");
    *out = static_cast<const bf16*>(packed);
    return B200_OK;
  }
  const bf16* w = nullptr;
  B200_TRY(get_bf16(h, name, {w_rows, K}, &w));
  if (h->cfg.quantize) {        // QLinear.__init__ (utils/quantize.py:16-21): row-wise absmax int8 of weight.to(float16)
    float* scales = nullptr;
    B200_TRY(h->weights.alloc(&packed, tc::sk_packed_bytes_i8(N, K, epi, gate_rows), false));
    B200_TRY(h->weights.alloc_t(&scales, (size_t)w_rows, false));
    B200_TRY(tc::sk_quant_pack_weights(w, packed, scales, N, K, epi, gate_rows, nullptr));
    *scales_out = scales;
    h->weight_bytes -= (int64_t)w_rows * K;        // one byte per weight instead of two (get_bf16 counted two) ...
    h->weight_bytes += (int64_t)w_rows * 4;        // ... plus the row scales
  } else {
    B200_TRY(h->weights.alloc(&packed, tc::sk_packed_bytes(N, K, epi, gate_rows), false));
    B200_TRY(tc::sk_pack_weights(w, packed, N, K, epi, gate_rows, nullptr));
  }
  B200_CUDA(cudaStreamSynchronize(nullptr));
  h->store.release(name);
  *out = static_cast<const bf16*>(packed);
  return B200_OK;
}

// logits [MB][card] of the model -> the [batch][card] rows the sampler reads: the conditioned rows themselves, or (CFG) the
// guided combination written to `guided`
const bf16* guide(b200_lm* h, const bf16* logits, bf16* guided, int card, bool combine) {
  if (!cfg_on(h) || !combine) return logits;
  const long long n = (long long)h->batch * card;
  B200_LAUNCH(cfg_combine_kernel, (unsigned)ceil_div64(n, 256), 256, 0, h->body, logits, guided, h->batch, card, h->cfg_coef);
  return guided;
}

int sample(b200_lm* h, const bf16* logits, int card, const float* noise, long long* out, float temp, int top_k) {
  B200_LAUNCH(sample_kernel, h->batch, SAMPLE_THREADS, 0, h->body, logits, (long long)card, noise,
              (long long)noise_per_row(h), out, card, h->use_sampling, temp, top_k);
  return check_launch("sample");
}

TokenRing ring(b200_lm* h) {
  TokenRing r;
  r.cache = h->cache; r.offsets = h->offsets; r.exec_mask = h->exec_mask; r.delays = h->delays_dev;
  r.Kc = h->Kc; r.CT = h->CT; r.dep_q = h->cfg.dep_q; r.n_q = h->cfg.n_q; r.card = h->cfg.card;
  r.text_card = h->cfg.text_card; r.max_delay = h->max_delay;
  r.cfg = cfg_on(h) ? 1 : 0; r.cfg_is_no_text = h->cfg_is_no_text; r.cfg_masked_until = h->cfg_until; r.err = h->err;
  return r;
}

// The whole frame as a fixed launch sequence over static buffers (captured into one CUDA graph).
int step_body(b200_lm* h) {
  const auto& c = h->cfg;
  const int B = h->batch, MB = h->MB, d = c.dim, H = c.num_heads, D = d / H, F = c.ffn_hidden;
  const int dd = c.depformer_dim, dH = c.depformer_num_heads, dF = c.depformer_ffn_hidden;
  cudaStream_t st = h->body;
  const float nl = -logf(c.max_period) * 2.f / (float)D;

  B200_LAUNCH(lm_prepare_kernel, ceil_div(B * h->Kc, 128), 128, 0, st, ring(h), h->in_codes, h->n_in_static,
              h->input_tokens, B);
  {
    EmbedTables t = h->emb;
    t.condition_sum = h->cond_on ? h->cond_sum : nullptr;
    t.err = h->err;
    B200_LAUNCH(lm_embed_sum_kernel, ceil_div(MB * d / 2, 256), 256, 0, st, t, h->input_tokens, h->x, MB, d);
  }
  if (h->noise_static) {
    const long long n = (long long)B * noise_per_row(h);
    B200_LAUNCH(lm_noise_kernel, (unsigned)ceil_div64(ceil_div64(n, 4), 256), 256, 0, st, h->noise, n, h->noise_seed, h->noise_ctr);
  }
  if (ring_cap(h) < c.context)      // a shortened ring serves a session only until it is full: flag instead of silently forgetting
    B200_LAUNCH(kv_capacity_check_kernel, ceil_div(MB, 128), 128, 0, st, h->pos, h->exec_mask_m, MB, ring_cap(h), h->err);
  for (auto& L : h->layers) {
    B200_TRY(norm_linear(h, h->x, L.n1, h->xn, L.in_w, h->qkv, 3 * d, MB, 3 * d, d, LIN_STORE, 0, L.in_s));
    if (h->kv_fp8) {
      AttnStepQ8 a;
      a.qkv = h->qkv; a.kc = L.kc8; a.vc = L.vc8; a.ks = L.ks; a.vs = L.vs; a.out = h->ao; a.part = h->attn_part;
      a.counters = h->attn_counters; a.pos = h->pos; a.exec_mask = h->exec_mask_m; a.H = H; a.cap = ring_cap(h); a.nsplit = h->nsplit;
      a.neg_log_period_2_over_d = nl;
      dim3 grid(MB * H, h->nsplit);
      if (h->kv_fp8 == B200_KV_INT8) B200_LAUNCH(attn_step_q8_kernel<KV_INT8>, grid, ATT_THREADS, 0, st, a);
      else B200_LAUNCH(attn_step_q8_kernel<KV_E4M3>, grid, ATT_THREADS, 0, st, a);
    } else {   // RoPE + ring append + split-KV attention + split merge in one launch
      AttnStep a;
      a.qkv = h->qkv; a.kc = L.kc; a.vc = L.vc; a.out = h->ao; a.part = h->attn_part; a.counters = h->attn_counters;
      a.pos = h->pos; a.exec_mask = h->exec_mask_m; a.H = H; a.cap = ring_cap(h); a.nsplit = h->nsplit;
      a.neg_log_period_2_over_d = nl;
      dim3 grid(MB * H, h->nsplit);
      if (attn_group_keys() == 2) B200_LAUNCH(attn_step_kernel<2>, grid, ATT_THREADS, 0, st, a);
      else B200_LAUNCH(attn_step_kernel<4>, grid, ATT_THREADS, 0, st, a);
    }
    B200_TRY(linear(h, h->ao, d, L.out_w, h->x, d, h->x, d, MB, d, d, LIN_RESADD, 0, L.out_s));
    B200_TRY(norm_linear(h, h->x, L.n2, h->xn, L.lin_in, h->hbuf, F, MB, F, d, LIN_GATE, F, L.lin_in_s));
    B200_TRY(linear(h, h->hbuf, F, L.lin_out, h->x, d, h->x, d, MB, d, F, LIN_RESADD, 0, L.lin_out_s));
  }
  B200_LAUNCH(rmsnorm_kernel, MB, 256, 0, st, h->x, h->out_norm, h->tout, d, 1e-8f);
  B200_TRY(linear(h, h->tout, d, h->text_linear, h->text_logits, c.text_card, nullptr, 0, MB, c.text_card, d, LIN_STORE, 0, h->text_linear_s));
  // lm.py:728-732: with cfg_is_no_text the text logits are the conditioned rows, otherwise the guided combination
  const bf16* tl = guide(h, h->text_logits, h->text_logits_cfg, c.text_card, !h->cfg_is_no_text);
  B200_TRY(sample(h, tl, c.text_card, h->noise, h->text_token, h->temp_text, h->top_k_text));
  if (!h->extra_w.empty())
    B200_LAUNCH(extra_heads_kernel, ceil_div((int)h->extra_w.size() * MB * 32, 128), 128, 0, st, h->tout, h->extra_w_dev, h->extra_out, MB, d,
                c.extra_heads_dim, (int)h->extra_w.size());

  // Depformer (lm.py:809-850): fresh KV state every frame, all rows advance together.
  const int kt = h->top_k_text < c.text_card ? h->top_k_text : c.text_card;
  const int ka = h->top_k < c.card ? h->top_k : c.card;
  if (h->replace_static || c.dep_q == 0) {
    // depformer_replace_tokens (lm.py:751-755): the caller supplies this frame's audio tokens, the depformer does not run;
    // dep_q == 0 (lm.py:219-222, "No-Depformer --- e.g., an ASR model"): there is none
    if (c.dep_q > 0)
      B200_LAUNCH(replace_audio_kernel, ceil_div(B * c.dep_q, 128), 128, 0, st, h->replace_tokens, h->audio_tokens, B, c.dep_q);
    B200_LAUNCH(advance_pos_kernel, ceil_div(MB, 128), 128, 0, st, h->pos, h->exec_mask_m, MB);
    B200_LAUNCH(lm_finish_kernel, ceil_div(B, 128), 128, 0, st, ring(h), h->text_token, h->audio_tokens, h->out_tokens, B, h->noise_ctr);
    return check_launch("lm step (no depformer)");
  }
  B200_TRY(linear(h, h->tout, d, h->dep_in_all, h->din, (long long)c.dep_q * dd, nullptr, 0, MB, c.dep_q * dd, d, LIN_STORE, 0, h->dep_in_s));
  if (h->depf) {
    B200_TRY(tc::dep_fused_launch(h->depf, st));
  } else {
    for (int k = 0; k < c.dep_q; ++k) {
      const long long* prev = k == 0 ? h->text_token : h->audio_tokens + (long long)(k - 1) * B;
      B200_LAUNCH(dep_input_kernel, ceil_div(MB * dd, 256), 256, 0, st, h->din, (long long)c.dep_q * dd, k * dd,
                  h->dep_tables[k], prev, h->dx, MB, dd, B, k == 0 ? c.text_card : c.card, h->err);
      for (auto& L : h->dlayers) {
        B200_TRY(norm_linear(h, h->dx, L.n1, h->dxn, L.in_w[k], h->dqkv, 3 * dd, MB, 3 * dd, dd, LIN_STORE, 0, L.in_s[k]));
        B200_LAUNCH(dep_attn_step_kernel, ceil_div(MB * dH * 32, 128), 128, 0, st, h->dqkv, L.kc, L.vc, h->dao, MB, dH, c.dep_q, k);
        B200_TRY(linear(h, h->dao, dd, L.out_w[k], h->dx, dd, h->dx, dd, MB, dd, dd, LIN_RESADD, 0, L.out_s[k]));
        B200_TRY(norm_linear(h, h->dx, L.n2, h->dxn, L.lin_in[k], h->dh, dF, MB, dF, dd, LIN_GATE, dF, L.lin_in_s[k]));
        B200_TRY(linear(h, h->dh, dF, L.lin_out[k], h->dx, dd, h->dx, dd, MB, dd, dF, LIN_RESADD, 0, L.lin_out_s[k]));
      }
      bf16* logits = h->dep_logits + (long long)k * MB * c.card;
      B200_TRY(linear(h, h->dx, dd, h->dep_heads[k], logits, c.card, nullptr, 0, MB, c.card, dd, LIN_STORE, 0, h->dep_heads_s[k]));
      const bf16* gl = guide(h, logits, cfg_on(h) ? h->dep_logits_cfg + (long long)k * B * c.card : nullptr, c.card, true);   // lm.py:828-833
      B200_TRY(sample(h, gl, c.card, h->noise + kt + (long long)k * ka, h->audio_tokens + (long long)k * B, h->temp, h->top_k));
    }
  }
  B200_LAUNCH(advance_pos_kernel, ceil_div(MB, 128), 128, 0, st, h->pos, h->exec_mask_m, MB);
  B200_LAUNCH(lm_finish_kernel, ceil_div(B, 128), 128, 0, st, ring(h), h->text_token, h->audio_tokens, h->out_tokens, B, h->noise_ctr);
  return check_launch("lm step");
}

int ensure_streaming(b200_lm* h, const char* what) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "%s: null handle", what);
  if (!h->finalized) B200_FAIL(B200_ERR_STATE, "%s: handle not finalized", what);
  if (h->batch <= 0) B200_FAIL(B200_ERR_STATE, "%s: not streaming (wrap calls in streaming())", what);   // lm.py:673-676
  return B200_OK;
}

void drop_graph(b200_lm* h) {
  if (h->graph_exec) {
    cudaGraphExecDestroy(h->graph_exec);
    h->graph_exec = nullptr;
  }
}

}  // namespace

extern "C" {

int b200_lm_create(const b200_lm_config* cfg, b200_lm** out) {
  if (!cfg || !out) B200_FAIL(B200_ERR_INVALID, "lm_create: null argument");
  if (cfg->n_q < 1 || cfg->n_q > 32 || cfg->dep_q < 0 || cfg->dep_q > cfg->n_q || cfg->dep_q > DEP_MAX_Q)
    B200_FAIL(B200_ERR_INVALID, "lm_create: n_q=%d dep_q=%d unsupported (n_q <= 32, dep_q <= %d)", cfg->n_q, cfg->dep_q, DEP_MAX_Q);
  if (cfg->extra_heads_num_heads < 0 || cfg->extra_heads_num_heads > 8 || (cfg->extra_heads_num_heads > 0 && (cfg->extra_heads_dim < 1 || cfg->extra_heads_dim > 32)))
    B200_FAIL(B200_ERR_INVALID, "lm_create: extra heads %d x %d unsupported (<= 8 heads of <= 32 outputs)", cfg->extra_heads_num_heads, cfg->extra_heads_dim);
  if (cfg->dim % cfg->num_heads || cfg->dim / cfg->num_heads != ATT_D)
    B200_FAIL(B200_ERR_INVALID, "lm_create: temporal head dim must be %d", ATT_D);
  if (cfg->dep_q > 0 && (cfg->depformer_dim % cfg->depformer_num_heads || cfg->depformer_dim / cfg->depformer_num_heads != 64))
    B200_FAIL(B200_ERR_INVALID, "lm_create: depformer head dim must be 64");
  if (cfg->dim % 8 || cfg->ffn_hidden % 8 || (cfg->dep_q > 0 && (cfg->depformer_dim % 8 || cfg->depformer_ffn_hidden % 8)))
    B200_FAIL(B200_ERR_INVALID, "lm_create: feature sizes must be multiples of 8");
  if (cfg->text_card + 1 > 65535 || cfg->card + 1 > 65535)
    B200_FAIL(B200_ERR_INVALID, "lm_create: vocabularies above 65535 unsupported by the sampler");
  b200_lm* h = new b200_lm();
  h->cfg = *cfg;
  cudaGetDevice(&h->device);
  h->Kc = cfg->n_q + 1;
  h->max_delay = 0;
  for (int k = 0; k < h->Kc; ++k) h->max_delay = cfg->delays[k] > h->max_delay ? cfg->delays[k] : h->max_delay;
  h->CT = h->max_delay + 2;     // lm.py:606-611
  if (const char* e = getenv("B200_PDL")) h->pdl = atoi(e) != 0;
  if (const char* e = getenv("B200_SK_SMEM_KB")) h->sk_smem = atoi(e) * 1024;
  if (const char* e = getenv("B200_DEP_FUSED")) h->dep_fused = atoi(e);
  if (const char* e = getenv("B200_FUSE_NORM")) h->fuse_norm = atoi(e) != 0;
  if (const char* e = getenv("B200_KV_DTYPE")) {
    const std::string v = e;
    h->kv_fp8 = (v == "fp8" || v == "fp8_e4m3") ? B200_KV_FP8_E4M3 : v == "int8" ? B200_KV_INT8 : 0;
  }
  *out = h;
  return B200_OK;
}

int b200_lm_load_tensor(b200_lm* h, const char* name, const void* data_dev, int dtype, int ndim, const int64_t* shape) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "lm_load_tensor: null handle");
  if (h->finalized) B200_FAIL(B200_ERR_STATE, "lm_load_tensor: already finalized");
  DeviceGuard g(h->device);
  return h->store.put(name, data_dev, dtype, ndim, shape);
}

int b200_lm_finalize(b200_lm* h) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "lm_finalize: null handle");
  if (h->finalized) return B200_OK;
  const auto& c = h->cfg;
  const int d = c.dim, dd = c.depformer_dim, F = c.ffn_hidden, dF = c.depformer_ffn_hidden;
  h->emb.n_q = c.n_q; h->emb.card = c.card; h->emb.text_card = c.text_card; h->emb.condition_sum = nullptr; h->emb.err = nullptr;
  DeviceGuard guard_dev(h->device);
  for (int k = 0; k < c.n_q; ++k)
    B200_TRY(get_bf16(h, "emb." + std::to_string(k) + ".weight", {c.card + 1, d}, &h->emb.audio[k]));
  B200_TRY(get_bf16(h, "text_emb.weight", {c.text_card + 1, d}, &h->emb.text));
  B200_TRY(get_linear(h, "text_linear.weight", c.text_card, d, LIN_STORE, 0, &h->text_linear, &h->text_linear_s));
  B200_TRY(get_bf16(h, "out_norm.alpha", {1, 1, d}, &h->out_norm));
  h->layers.resize(c.num_layers);
  for (int l = 0; l < c.num_layers; ++l) {
    const std::string p = "transformer.layers." + std::to_string(l);
    TLayer& L = h->layers[l];
    B200_TRY(get_linear(h, p + ".self_attn.in_projs.0.weight", 3 * d, d, LIN_STORE, 0, &L.in_w, &L.in_s));
    B200_TRY(get_linear(h, p + ".self_attn.out_projs.0.weight", d, d, LIN_RESADD, 0, &L.out_w, &L.out_s));
    B200_TRY(get_bf16(h, p + ".norm1.alpha", {1, 1, d}, &L.n1));
    B200_TRY(get_bf16(h, p + ".norm2.alpha", {1, 1, d}, &L.n2));
    B200_TRY(get_linear(h, p + ".gating.linear_in.weight", 2 * F, d, LIN_GATE, F, &L.lin_in, &L.lin_in_s));
    B200_TRY(get_linear(h, p + ".gating.linear_out.weight", d, F, LIN_RESADD, 0, &L.lin_out, &L.lin_out_s));
  }
  for (int i = 0; i < c.extra_heads_num_heads; ++i) {      // lm.py:224-226
    const bf16* w = nullptr;
    B200_TRY(get_bf16(h, "extra_heads." + std::to_string(i) + ".weight", {c.extra_heads_dim, d}, &w));
    h->extra_w.push_back(w);
  }
  if (!h->extra_w.empty()) {
    B200_TRY(h->weights.alloc_t(&h->extra_w_dev, h->extra_w.size(), false));
    B200_CUDA(cudaMemcpy(h->extra_w_dev, h->extra_w.data(), h->extra_w.size() * sizeof(void*), cudaMemcpyHostToDevice));
  }
  // depformer_in.{k} stacked so that all dep_q projections of transformer_out are one GEMM
  const Tensor* din0 = c.dep_q > 0 ? h->store.find("depformer_in.0.weight") : nullptr;
  if (din0 && din0->dtype == B200_I8) {          // pre-quantised: stack CB and SCB
    int8_t* stacked = nullptr; float* scales = nullptr; void* packed = nullptr;
    B200_CUDA(cudaMalloc(&stacked, (size_t)c.dep_q * dd * d));
    struct Guard { void* p; ~Guard() { cudaFree(p); } } guard{stacked};
    B200_TRY(h->weights.alloc_t(&scales, (size_t)c.dep_q * dd, false));
    for (int k = 0; k < c.dep_q; ++k) {
      const std::string name = "depformer_in." + std::to_string(k) + ".weight";
      const int8_t* q = nullptr; const float* scb = nullptr;
      if (!h->store.find(name)) B200_FAIL(B200_ERR_MISSING, "lm finalize: tensor '%s' was not loaded", name.c_str());
      B200_TRY(get_prequantised(h, name, dd, d, &q, &scb));
      B200_CUDA(cudaMemcpy(stacked + (size_t)k * dd * d, q, (size_t)dd * d, cudaMemcpyDeviceToDevice));
      B200_CUDA(cudaMemcpy(scales + (size_t)k * dd, scb, (size_t)dd * 4, cudaMemcpyDeviceToDevice));
      h->store.release(name);
      h->store.release(name + "_scb");
    }
    B200_TRY(h->weights.alloc(&packed, tc::sk_packed_bytes_i8(c.dep_q * dd, d, LIN_STORE, 0), false));
    B200_TRY(tc::sk_pack_weights_i8(stacked, packed, c.dep_q * dd, d, LIN_STORE, 0, nullptr));
    B200_CUDA(cudaStreamSynchronize(nullptr));
    h->dep_in_s = scales;
    h->dep_in_all = static_cast<bf16*>(packed);
    h->weight_bytes += (int64_t)c.dep_q * dd * d + (int64_t)c.dep_q * dd * 4;
  } else if (c.dep_q > 0) {
    bf16* stacked = nullptr;
    B200_CUDA(cudaMalloc(&stacked, (size_t)c.dep_q * dd * d * 2));
    struct Guard { void* p; ~Guard() { cudaFree(p); } } guard{stacked};      // released on every path out of this block
    for (int k = 0; k < c.dep_q; ++k) {
      const bf16* w = nullptr;
      const std::string name = "depformer_in." + std::to_string(k) + ".weight";
      B200_TRY(get_bf16(h, name, {dd, d}, &w));
      B200_CUDA(cudaMemcpy(stacked + (size_t)k * dd * d, w, (size_t)dd * d * 2, cudaMemcpyDeviceToDevice));
      h->store.release(name);
    }
    void* packed = nullptr;
    int rc = B200_OK;
    if (c.quantize) {
      float* scales = nullptr;
      rc = h->weights.alloc(&packed, tc::sk_packed_bytes_i8(c.dep_q * dd, d, LIN_STORE, 0), false);
      if (rc == B200_OK) rc = h->weights.alloc_t(&scales, (size_t)c.dep_q * dd, false);
      if (rc == B200_OK) rc = tc::sk_quant_pack_weights(stacked, packed, scales, c.dep_q * dd, d, LIN_STORE, 0, nullptr);
      h->dep_in_s = scales;
      h->weight_bytes -= (int64_t)c.dep_q * dd * d;
      h->weight_bytes += (int64_t)c.dep_q * dd * 4;
    } else {
      rc = h->weights.alloc(&packed, tc::sk_packed_bytes(c.dep_q * dd, d, LIN_STORE, 0), false);
      if (rc == B200_OK) rc = tc::sk_pack_weights(stacked, packed, c.dep_q * dd, d, LIN_STORE, 0, nullptr);
    }
    B200_TRY(rc);
    B200_CUDA(cudaStreamSynchronize(nullptr));
    h->dep_in_all = static_cast<bf16*>(packed);
  }
  h->dep_tables.resize(c.dep_q);
  if (c.dep_q > 0) B200_TRY(get_bf16(h, "depformer_text_emb.weight", {c.text_card + 1, dd}, &h->dep_tables[0]));
  for (int k = 1; k < c.dep_q; ++k)
    B200_TRY(get_bf16(h, "depformer_emb." + std::to_string(k - 1) + ".weight", {c.card + 1, dd}, &h->dep_tables[k]));
  h->dep_heads.resize(c.dep_q);
  h->dep_heads_s.resize(c.dep_q);
  for (int k = 0; k < c.dep_q; ++k)
    B200_TRY(get_linear(h, "linears." + std::to_string(k) + ".weight", c.card, dd, LIN_STORE, 0, &h->dep_heads[k], &h->dep_heads_s[k]));
  h->dlayers.resize(c.dep_q > 0 ? c.depformer_num_layers : 0);
  for (int l = 0; l < (int)h->dlayers.size(); ++l) {
    const std::string p = "depformer.layers." + std::to_string(l);
    DLayer& L = h->dlayers[l];
    L.in_w.resize(c.dep_q); L.out_w.resize(c.dep_q); L.lin_in.resize(c.dep_q); L.lin_out.resize(c.dep_q);
    L.in_s.resize(c.dep_q); L.out_s.resize(c.dep_q); L.lin_in_s.resize(c.dep_q); L.lin_out_s.resize(c.dep_q);
    for (int k = 0; k < c.dep_q; ++k) {
      const std::string ks = std::to_string(k);
      B200_TRY(get_linear(h, p + ".self_attn.in_projs." + ks + ".weight", 3 * dd, dd, LIN_STORE, 0, &L.in_w[k], &L.in_s[k]));
      B200_TRY(get_linear(h, p + ".self_attn.out_projs." + ks + ".weight", dd, dd, LIN_RESADD, 0, &L.out_w[k], &L.out_s[k]));
      B200_TRY(get_linear(h, p + ".gating." + ks + ".linear_in.weight", 2 * dF, dd, LIN_GATE, dF, &L.lin_in[k], &L.lin_in_s[k]));
      B200_TRY(get_linear(h, p + ".gating." + ks + ".linear_out.weight", dd, dF, LIN_RESADD, 0, &L.lin_out[k], &L.lin_out_s[k]));
    }
    B200_TRY(get_bf16(h, p + ".norm1.alpha", {1, 1, dd}, &L.n1));
    B200_TRY(get_bf16(h, p + ".norm2.alpha", {1, 1, dd}, &L.n2));
  }
  B200_TRY(h->weights.alloc_t(&h->delays_dev, h->Kc, false));
  B200_CUDA(cudaMemcpy(h->delays_dev, c.delays, h->Kc * sizeof(int), cudaMemcpyHostToDevice));
  B200_CUDA(cudaDeviceSynchronize());
  h->finalized = true;
  return B200_OK;
}

int b200_lm_destroy(b200_lm* h) {
  if (!h) return B200_OK;
  DeviceGuard g(h->device);
  b200_lm_streaming_end(h);
  h->store.release_all();
  h->weights.free_all();
  delete h;
  return B200_OK;
}

int b200_lm_set_sampling(b200_lm* h, int use_sampling, float temp, float temp_text, int top_k, int top_k_text) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "lm_set_sampling: null handle");
  if (top_k < 1 || top_k_text < 1 || top_k > SAMPLE_MAX_K || top_k_text > SAMPLE_MAX_K)
    B200_FAIL(B200_ERR_INVALID, "lm_set_sampling: top_k must be in [1, %d]", SAMPLE_MAX_K);
  if (h->batch > 0 && (top_k != h->top_k || top_k_text != h->top_k_text))
    B200_FAIL(B200_ERR_STATE, "lm_set_sampling: top_k cannot change while streaming");
  if (use_sampling == h->use_sampling && temp == h->temp && temp_text == h->temp_text && top_k == h->top_k && top_k_text == h->top_k_text)
    return B200_OK;
  h->use_sampling = use_sampling; h->temp = temp; h->temp_text = temp_text; h->top_k = top_k; h->top_k_text = top_k_text;
  tc::dep_fused_set_sampling(h->depf, use_sampling, temp, top_k);
  drop_graph(h);         // the sampling parameters are kernel arguments of the captured step
  return B200_OK;
}

int b200_lm_noise_per_row(b200_lm* h) { return h ? noise_per_row(h) : 0; }

int b200_lm_set_graph(b200_lm* h, int enable) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "lm_set_graph: null handle");
  h->graph_enabled = enable;
  if (!enable) drop_graph(h);
  return B200_OK;
}

int b200_lm_set_cfg(b200_lm* h, float cfg_coef, int cfg_is_no_text, const int64_t* masked_until_host, int n) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "lm_set_cfg: null handle");
  if (h->batch > 0) B200_FAIL(B200_ERR_STATE, "lm_set_cfg: classifier-free guidance doubles the model's rows; set it before streaming_begin");
  if (masked_until_host && n < 1) B200_FAIL(B200_ERR_INVALID, "lm_set_cfg: empty cfg_is_masked_until");
  h->cfg_coef = cfg_coef;
  h->cfg_is_no_text = cfg_is_no_text ? 1 : 0;
  h->cfg_until_host.clear();
  if (masked_until_host) h->cfg_until_host.assign(masked_until_host, masked_until_host + n);
  return B200_OK;
}

int b200_lm_seed_noise(b200_lm* h, uint64_t seed) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "lm_seed_noise: null handle");
  h->noise_seed = seed;
  DeviceGuard g(h->device);
  if (h->noise_ctr) B200_CUDA(cudaMemsetAsync(h->noise_ctr, 0, 8, h->stream));
  drop_graph(h);         // the seed is a kernel argument of the captured step
  return B200_OK;
}

int b200_lm_set_stream(b200_lm* h, void* stream) {
  B200_TRY(ensure_streaming(h, "lm_set_stream"));
  h->stream = static_cast<cudaStream_t>(stream);
  return B200_OK;
}

int b200_lm_streaming_begin(b200_lm* h, int batch, void* stream) {
  if (!h || !h->finalized) B200_FAIL(B200_ERR_STATE, "lm_streaming_begin: handle not finalized");
  if (h->batch > 0) B200_FAIL(B200_ERR_STATE, "lm_streaming_begin: already streaming");
  if (batch < 1) B200_FAIL(B200_ERR_INVALID, "lm_streaming_begin: batch %d", batch);
  DeviceGuard guard_dev(h->device);
  const auto& c = h->cfg;
  const bool cfg = cfg_on(h);
  const int B = batch, MB = cfg ? 2 * batch : batch, d = c.dim, H = c.num_heads, D = d / H, dd = c.depformer_dim;
  if (MB > 256) B200_FAIL(B200_ERR_INVALID, "lm_streaming_begin: %d model rows (the linears take at most 256)", MB);
  if (cfg && !h->cfg_until_host.empty() && (int)h->cfg_until_host.size() != B)      // lm.py:717 view(-1, 1, 1) broadcast
    B200_FAIL(B200_ERR_SHAPE, "lm_streaming_begin: cfg_is_masked_until has %zu entries for %d sessions", h->cfg_until_host.size(), B);
  h->stream = static_cast<cudaStream_t>(stream);
  Arena& A = h->state;
  struct Cleanup { b200_lm* h; bool armed = true; ~Cleanup() { if (armed) { h->state.free_all(); h->cfg_until = nullptr; h->noise_ctr = nullptr; } } } cleanup{h};
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const int cap = ring_cap(h);
  const size_t kv_bytes = (size_t)c.num_layers * 2 * MB * H * cap * (h->kv_fp8 ? (size_t)D + 4 : (size_t)D * 2);
  if (kv_bytes + (1ull << 30) > free_b)
    B200_FAIL(B200_ERR_INVALID, "lm_streaming_begin: %d sessions need %.1f GB of KV ring, %.1f GB free", B, kv_bytes / 1e9,
              free_b / 1e9);
  B200_TRY(A.alloc_t(&h->exec_mask, B, false));
  B200_CUDA(cudaMemset(h->exec_mask, 1, B));
  if (cfg) {
    B200_TRY(A.alloc_t(&h->exec_mask_m, MB, false));
    B200_CUDA(cudaMemset(h->exec_mask_m, 1, MB));
  } else {
    h->exec_mask_m = h->exec_mask;
  }
  B200_TRY(A.alloc_t(&h->err, 1));
  B200_TRY(A.alloc_t(&h->noise_ctr, 1));
  h->cfg_until = nullptr;
  if (cfg && !h->cfg_until_host.empty()) {
    B200_TRY(A.alloc_t(&h->cfg_until, B, false));
    B200_CUDA(cudaMemcpy(h->cfg_until, h->cfg_until_host.data(), (size_t)B * 8, cudaMemcpyHostToDevice));
  }
  B200_TRY(A.alloc_t(&h->cache, (size_t)B * h->Kc * h->CT, false));
  {
    std::vector<long long> init((size_t)B * h->Kc * h->CT, -2);   // ungenerated_token_id, lm.py:606-611
    B200_CUDA(cudaMemcpy(h->cache, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
  }
  B200_TRY(A.alloc_t(&h->offsets, B));
  B200_TRY(A.alloc_t(&h->pos, MB));
  h->offset_cpu = 0;
  for (auto& L : h->layers) {
    if (h->kv_fp8) {
      B200_TRY(A.alloc_t(&L.kc8, (size_t)MB * H * cap * D));
      B200_TRY(A.alloc_t(&L.vc8, (size_t)MB * H * cap * D));
      B200_TRY(A.alloc_t(&L.ks, (size_t)MB * H * cap));
      B200_TRY(A.alloc_t(&L.vs, (size_t)MB * H * cap));
    } else {
      B200_TRY(A.alloc_t(&L.kc, (size_t)MB * H * cap * D));
      B200_TRY(A.alloc_t(&L.vc, (size_t)MB * H * cap * D));
    }
  }
  for (auto& L : h->dlayers) {
    B200_TRY(A.alloc_t(&L.kc, (size_t)MB * dd * c.dep_q));
    B200_TRY(A.alloc_t(&L.vc, (size_t)MB * dd * c.dep_q));
  }
  const int n_in_max = c.n_q;   // callers may pass more columns than needed (lm.py:688-689)
  const int dq1 = c.dep_q > 0 ? c.dep_q : 1;
  B200_TRY(A.alloc_t(&h->in_codes, (size_t)B * n_in_max));
  B200_TRY(A.alloc_t(&h->input_tokens, (size_t)MB * h->Kc));
  B200_TRY(A.alloc_t(&h->text_token, B));
  B200_TRY(A.alloc_t(&h->audio_tokens, (size_t)B * dq1));
  B200_TRY(A.alloc_t(&h->replace_tokens, (size_t)B * dq1));
  B200_TRY(A.alloc_t(&h->out_tokens, (size_t)B * (c.dep_q + 1)));
  B200_TRY(A.alloc_t(&h->noise, (size_t)B * noise_per_row(h), false));
  {
    std::vector<float> ones((size_t)B * noise_per_row(h), 1.f);
    B200_CUDA(cudaMemcpy(h->noise, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice));
  }
  B200_TRY(A.alloc_t(&h->x, (size_t)MB * d));
  B200_TRY(A.alloc_t(&h->xn, (size_t)MB * d));
  B200_TRY(A.alloc_t(&h->qkv, (size_t)MB * 3 * d));
  B200_TRY(A.alloc_t(&h->ao, (size_t)MB * d));
  B200_TRY(A.alloc_t(&h->hbuf, (size_t)MB * c.ffn_hidden));
  B200_TRY(A.alloc_t(&h->tout, (size_t)MB * d));
  B200_TRY(A.alloc_t(&h->cond_sum, (size_t)MB * d));
  h->cond_on = 0;
  B200_TRY(A.alloc_t(&h->text_logits, (size_t)MB * c.text_card));
  if (cfg) {
    B200_TRY(A.alloc_t(&h->text_logits_cfg, (size_t)B * c.text_card));
    B200_TRY(A.alloc_t(&h->dep_logits_cfg, (size_t)B * dq1 * c.card));
  }
  if (!h->extra_w.empty()) B200_TRY(A.alloc_t(&h->extra_out, h->extra_w.size() * (size_t)MB * c.extra_heads_dim));
  B200_TRY(A.alloc_t(&h->din, (size_t)MB * dq1 * dd));
  B200_TRY(A.alloc_t(&h->dx, (size_t)MB * dd));
  B200_TRY(A.alloc_t(&h->dxn, (size_t)MB * dd));
  B200_TRY(A.alloc_t(&h->dqkv, (size_t)MB * 3 * dd));
  B200_TRY(A.alloc_t(&h->dao, (size_t)MB * dd));
  B200_TRY(A.alloc_t(&h->dh, (size_t)MB * c.depformer_ffn_hidden));
  B200_TRY(A.alloc_t(&h->dep_logits, (size_t)MB * dq1 * c.card));
  B200_TRY(A.alloc(reinterpret_cast<void**>(&h->sk_ws), tc::sk_workspace_bytes(MB), false));
  B200_TRY(A.alloc_t(&h->sk_counters, tc::SK_MAX_TILES));
  const int ns = attn_pick_splits(MB, H, cap);
  h->nsplit = ns;
  B200_TRY(A.alloc_t(&h->attn_part, (size_t)MB * H * ns * (ATT_D + 2)));
  B200_TRY(A.alloc_t(&h->attn_counters, (size_t)MB * H));
  B200_CUDA(cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking));
  B200_CUDA(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
  B200_CUDA(cudaMallocHost(&h->pin_in, (size_t)B * n_in_max * 8));
  B200_CUDA(cudaMallocHost(&h->pin_out, (size_t)B * (c.dep_q + 1) * 8));
  B200_CUDA(cudaMallocHost(&h->pin_noise, (size_t)B * noise_per_row(h) * 4));
  if (c.quantize) {     // activations are re-quantised row-wise before every linear (QLinear.forward, utils/quantize.py:28-36)
    const int kmax = std::max(std::max(d, c.ffn_hidden), std::max(dd, c.depformer_ffn_hidden));
    B200_TRY(A.alloc_t(&h->xq, (size_t)MB * kmax, false));
    B200_TRY(A.alloc_t(&h->xq_scale, (size_t)MB, false));
  }
  // One or two sessions: the depformer's linears take the GEMV path and the PDL-chained launches beat the persistent kernel
  // (B=1: 5.5 vs 6.4 ms per LM step, B=2: 6.2 vs 7.1; profiles/r01_r_*).  B200_DEP_FUSED=2 forces the kernel at any batch.
  // Under classifier-free guidance the two halves of the batch meet before every sampler (lm.py:828-833): launch chain.
  const bool dep_small = B <= tc::sk_gemv_max_m() && h->dep_fused != 2;
  h->depf = nullptr;
  if (c.dep_q > 0 && !cfg && !c.quantize && h->dep_fused && !dep_small && B <= 256 && dd <= 1024 && dd % 64 == 0) {
    tc::DepFusedConfig fc;
    memset(&fc, 0, sizeof(fc));
    fc.B = B; fc.dd = dd; fc.H = c.depformer_num_heads; fc.F = c.depformer_ffn_hidden; fc.card = c.card; fc.text_card = c.text_card;
    fc.dep_q = c.dep_q; fc.L = c.depformer_num_layers; fc.err = h->err;
    std::vector<const void*> in_w, out_w, lin_in, lin_out, heads, tables, n1, n2;
    std::vector<void*> kc, vc;
    for (int k = 0; k < c.dep_q; ++k)
      for (auto& L : h->dlayers) {
        in_w.push_back(L.in_w[k]); out_w.push_back(L.out_w[k]); lin_in.push_back(L.lin_in[k]); lin_out.push_back(L.lin_out[k]);
      }
    for (int k = 0; k < c.dep_q; ++k) { heads.push_back(h->dep_heads[k]); tables.push_back(h->dep_tables[k]); }
    for (auto& L : h->dlayers) { n1.push_back(L.n1); n2.push_back(L.n2); kc.push_back(L.kc); vc.push_back(L.vc); }
    fc.in_w = in_w.data(); fc.out_w = out_w.data(); fc.lin_in = lin_in.data(); fc.lin_out = lin_out.data();
    fc.heads = heads.data(); fc.tables = tables.data(); fc.n1 = n1.data(); fc.n2 = n2.data(); fc.kc = kc.data(); fc.vc = vc.data();
    fc.din = h->din; fc.din_ld = (long long)c.dep_q * dd; fc.text_token = h->text_token;
    fc.x = h->dx; fc.xn = h->dxn; fc.ao = h->dao; fc.hbuf = h->dh;
    const size_t pf = tc::dep_fused_partial_floats(fc);
    B200_TRY(A.alloc_t(&h->dep_part0, pf, false));
    B200_TRY(A.alloc_t(&h->dep_part1, pf, false));
    B200_TRY(A.alloc_t(&h->dep_bar, 1));
    fc.part0 = h->dep_part0; fc.part1 = h->dep_part1; fc.bar = h->dep_bar;
    if (const char* e = getenv("B200_DEP_TRACE")) {
      if (atoi(e)) B200_TRY(A.alloc_t(&h->dep_trace, (size_t)tc::DEP_TRACE_SLOTS));
    }
    fc.trace = h->dep_trace;
    fc.logits = h->dep_logits; fc.audio_tokens = h->audio_tokens;
    fc.noise = h->noise; fc.noise_ld = noise_per_row(h);
    fc.noise_off = h->top_k_text < c.text_card ? h->top_k_text : c.text_card;
    fc.ka = h->top_k < c.card ? h->top_k : c.card;
    fc.use_sampling = h->use_sampling; fc.top_k = h->top_k; fc.temp = h->temp;
    B200_TRY(tc::dep_fused_create(fc, &h->depf));
  }
  // streaming-state snapshot (lm.py:527-542 _LMGenState + the temporal transformer's ring caches, transformer.py:196-288)
  A.mark_state(h->exec_mask, B, "exec_mask", B200_U8, {B});
  if (cfg) A.mark_state(h->exec_mask_m, MB, "model.exec_mask", B200_U8, {MB});
  A.mark_state(h->cache, (size_t)B * h->Kc * h->CT * 8, "cache", B200_I64, {B, h->Kc, h->CT});
  A.mark_state(h->offsets, (size_t)B * 8, "offsets", B200_I64, {B});
  A.mark_state(h->pos, (size_t)MB * 8, "model.offset", B200_I64, {MB});
  A.mark_state(h->noise_ctr, 8, "noise_counter", B200_I64, {1});
  for (size_t l = 0; l < h->layers.size(); ++l) {
    auto& L = h->layers[l];
    const std::string p = "layers." + std::to_string(l);
    if (h->kv_fp8) {
      A.mark_state(L.kc8, (size_t)MB * H * cap * D, p + ".k8", B200_U8, {MB, H, cap, D});
      A.mark_state(L.vc8, (size_t)MB * H * cap * D, p + ".v8", B200_U8, {MB, H, cap, D});
      A.mark_state(L.ks, (size_t)MB * H * cap * 4, p + ".k_scale", B200_F32, {MB, H, cap});
      A.mark_state(L.vs, (size_t)MB * H * cap * 4, p + ".v_scale", B200_F32, {MB, H, cap});
    } else {
      A.mark_state(L.kc, (size_t)MB * H * cap * D * 2, p + ".k", B200_BF16, {MB, H, cap, D});
      A.mark_state(L.vc, (size_t)MB * H * cap * D * 2, p + ".v", B200_BF16, {MB, H, cap, D});
    }
  }
  B200_CUDA(cudaDeviceSynchronize());
  cleanup.armed = false;
  h->batch = B;
  h->MB = MB;
  return B200_OK;
}

/* get_streaming_state / set_streaming_state: device blob = marked buffers + the host step counter (last 256 B) */
int64_t b200_lm_state_bytes(b200_lm* h) { return (h && h->batch > 0) ? (int64_t)h->state.state_bytes() + 256 : 0; }

int b200_lm_get_state(b200_lm* h, void* dst_dev, int64_t capacity) {
  B200_TRY(ensure_streaming(h, "lm_get_state"));
  DeviceGuard g(h->device);
  const size_t n = h->state.state_bytes();
  if (!dst_dev || capacity < (int64_t)n + 256) B200_FAIL(B200_ERR_SHAPE, "lm_get_state: destination too small");
  B200_TRY(h->state.save(dst_dev, h->stream));
  B200_CUDA(cudaMemcpyAsync(static_cast<char*>(dst_dev) + n, &h->offset_cpu, 8, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200_OK;
}

int b200_lm_set_state(b200_lm* h, const void* src_dev, int64_t nbytes) {
  B200_TRY(ensure_streaming(h, "lm_set_state"));
  DeviceGuard g(h->device);
  const size_t n = h->state.state_bytes();
  if (!src_dev || nbytes != (int64_t)n + 256) B200_FAIL(B200_ERR_SHAPE, "lm_set_state: snapshot does not fit this session layout");
  B200_TRY(h->state.load(src_dev, h->stream));
  B200_CUDA(cudaMemcpyAsync(&h->offset_cpu, static_cast<const char*>(src_dev) + n, 8, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200_OK;
}

/* the same state, entry by entry (the Python shim groups them into the reference's per-module State objects) */
int b200_lm_state_count(b200_lm* h) { return (h && h->batch > 0) ? (int)h->state.snap.size() : 0; }
int b200_lm_state_entry(b200_lm* h, int index, const char** name, int* dtype, int* ndim, int64_t* shape8, int64_t* nbytes) {
  B200_TRY(ensure_streaming(h, "lm_state_entry"));
  return state_entry_info(h->state, index, name, dtype, ndim, shape8, nbytes);
}
int b200_lm_state_read(b200_lm* h, const char* name, void* dst_dev, int64_t nbytes) {
  B200_TRY(ensure_streaming(h, "lm_state_read"));
  DeviceGuard g(h->device);
  if (!dst_dev) B200_FAIL(B200_ERR_INVALID, "lm_state_read: null destination");
  return state_entry_copy(h->state, name, dst_dev, nullptr, nbytes, h->stream);
}
int b200_lm_state_write(b200_lm* h, const char* name, const void* src_dev, int64_t nbytes) {
  B200_TRY(ensure_streaming(h, "lm_state_write"));
  DeviceGuard g(h->device);
  if (!src_dev) B200_FAIL(B200_ERR_INVALID, "lm_state_write: null source");
  return state_entry_copy(h->state, name, nullptr, src_dev, nbytes, h->stream);
}
int64_t b200_lm_get_offset_cpu(b200_lm* h) { return h ? h->offset_cpu : 0; }
int b200_lm_set_offset_cpu(b200_lm* h, int64_t v) {
  B200_TRY(ensure_streaming(h, "lm_set_offset_cpu"));
  h->offset_cpu = v;
  return B200_OK;
}

int b200_lm_streaming_end(b200_lm* h) {
  if (!h) return B200_OK;
  DeviceGuard g(h->device);
  if (h->batch > 0) {
    cudaStreamSynchronize(h->stream);
    if (h->gstream) cudaStreamSynchronize(h->gstream);
  }
  drop_graph(h);
  if (h->gstream) cudaStreamDestroy(h->gstream);
  if (h->ev_in) cudaEventDestroy(h->ev_in);
  if (h->ev_out) cudaEventDestroy(h->ev_out);
  h->gstream = nullptr; h->ev_in = h->ev_out = nullptr;
  h->plans.clear();
  tc::dep_fused_destroy(h->depf);
  h->depf = nullptr;
  h->state.free_all();
  h->cfg_until = nullptr; h->noise_ctr = nullptr; h->err = nullptr; h->cond_sum = nullptr; h->cond_on = 0;
  if (h->pin_in) cudaFreeHost(h->pin_in);
  if (h->pin_out) cudaFreeHost(h->pin_out);
  if (h->pin_noise) cudaFreeHost(h->pin_noise);
  h->pin_in = h->pin_out = nullptr;
  h->pin_noise = nullptr;
  h->batch = 0;
  h->MB = 0;
  return B200_OK;
}

int b200_lm_reset(b200_lm* h, const uint8_t* reset_mask_dev) {
  B200_TRY(ensure_streaming(h, "lm_reset"));
  DeviceGuard g(h->device);
  const int B = h->batch;
  // per-row device offsets (lm.py:539, transformer.py:229-234, 331-333) ...
  B200_LAUNCH(lm_reset_kernel, ceil_div(B, 128), 128, 0, h->stream, h->offsets, h->pos, h->exec_mask, h->exec_mask_m,
              reset_mask_dev, B, cfg_on(h) ? 1 : 0);
  // ... and the global host counter (lm.py:540): after any reset, step() reports "not ready" for max_delay calls
  h->offset_cpu = 0;
  return check_launch("lm_reset");
}

int b200_lm_set_exec_mask(b200_lm* h, const uint8_t* exec_mask_dev) {
  B200_TRY(ensure_streaming(h, "lm_set_exec_mask"));
  DeviceGuard g(h->device);
  if (!exec_mask_dev) B200_FAIL(B200_ERR_INVALID, "lm_set_exec_mask: null mask");
  B200_CUDA(cudaMemcpyAsync(h->exec_mask, exec_mask_dev, h->batch, cudaMemcpyDeviceToDevice, h->stream));
  if (cfg_on(h)) {      // exec_mask.repeat(2) (lm.py:658-661)
    B200_CUDA(cudaMemcpyAsync(h->exec_mask_m, exec_mask_dev, h->batch, cudaMemcpyDeviceToDevice, h->stream));
    B200_CUDA(cudaMemcpyAsync(h->exec_mask_m + h->batch, exec_mask_dev, h->batch, cudaMemcpyDeviceToDevice, h->stream));
  }
  return B200_OK;
}

int b200_lm_set_condition_sum(b200_lm* h, const void* sum_bf16_dev, int rows) {
  B200_TRY(ensure_streaming(h, "lm_set_condition_sum"));
  DeviceGuard g(h->device);
  const int on = sum_bf16_dev ? 1 : 0;
  if (on && rows != h->MB)      // lm.py:648-651: "cfg requires 2x more conditions"
    B200_FAIL(B200_ERR_SHAPE, "lm_set_condition_sum: %d rows given, the model runs on %d", rows, h->MB);
  if (on) B200_CUDA(cudaMemcpyAsync(h->cond_sum, sum_bf16_dev, (size_t)h->MB * h->cfg.dim * 2, cudaMemcpyDeviceToDevice, h->stream));
  if (on != h->cond_on) drop_graph(h);
  h->cond_on = on;
  return B200_OK;
}

int b200_lm_set_kv_dtype(b200_lm* h, int kv_dtype) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "lm_set_kv_dtype: null handle");
  if (h->batch > 0) B200_FAIL(B200_ERR_STATE, "lm_set_kv_dtype: the rings are allocated at streaming_begin; set the dtype before it");
  if (kv_dtype != B200_KV_BF16 && kv_dtype != B200_KV_FP8_E4M3 && kv_dtype != B200_KV_INT8)
    B200_FAIL(B200_ERR_INVALID, "lm_set_kv_dtype: unknown dtype %d", kv_dtype);
  h->kv_fp8 = kv_dtype;
  return B200_OK;
}

int b200_lm_set_kv_capacity(b200_lm* h, int slots) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "lm_set_kv_capacity: null handle");
  if (h->batch > 0) B200_FAIL(B200_ERR_STATE, "lm_set_kv_capacity: the rings are allocated at streaming_begin; set the capacity before it");
  if (slots < 0) B200_FAIL(B200_ERR_INVALID, "lm_set_kv_capacity: negative capacity");
  h->kv_cap = slots >= h->cfg.context ? 0 : slots;
  return B200_OK;
}

int b200_lm_assume_fill(b200_lm* h, int fill) {
  B200_TRY(ensure_streaming(h, "lm_assume_fill"));
  DeviceGuard g(h->device);
  if (fill < 0) B200_FAIL(B200_ERR_INVALID, "lm_assume_fill: negative fill");
  std::vector<long long> v(h->MB, fill);
  B200_CUDA(cudaMemcpyAsync(h->offsets, v.data(), (size_t)h->batch * 8, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaMemcpyAsync(h->pos, v.data(), v.size() * 8, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  h->offset_cpu = fill;
  return B200_OK;
}

/* Synchronises the handle's stream and returns (and clears) the device error flags: B200_FLAG_TOKEN_RANGE = a token id outside
 * its embedding table reached the model (the reference would hit a device assert in F.embedding). */
int b200_lm_error_flags(b200_lm* h, int* flags_out) {
  B200_TRY(ensure_streaming(h, "lm_error_flags"));
  DeviceGuard g(h->device);
  int v = 0;
  B200_CUDA(cudaMemcpyAsync(&v, h->err, 4, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  if (v) B200_CUDA(cudaMemsetAsync(h->err, 0, 4, h->stream));
  if (flags_out) *flags_out = v;
  return B200_OK;
}

static int run_step(b200_lm* h, int n_in, int replace, int internal_noise) {
  if (n_in != h->n_in_static || replace != h->replace_static || internal_noise != h->noise_static) {     // the captured graph is specific to all three
    h->n_in_static = n_in;
    h->replace_static = replace;
    h->noise_static = internal_noise;
    drop_graph(h);
  }
  B200_TRY(tc::prepare_plans(h->plans));
  if (!h->graph_enabled) {
    h->body = h->stream;
    return step_body(h);
  }
  // the step runs on the private stream, fenced against the caller's stream on both sides
  B200_CUDA(cudaEventRecord(h->ev_in, h->stream));
  B200_CUDA(cudaStreamWaitEvent(h->gstream, h->ev_in, 0));
  if (!h->graph_exec) {
    cudaGraph_t graph = nullptr;
    const int64_t before = g_launches.load();
    h->body = h->gstream;
    B200_CUDA(cudaStreamBeginCapture(h->gstream, cudaStreamCaptureModeRelaxed));
    const int rc = step_body(h);
    cudaError_t e = cudaStreamEndCapture(h->gstream, &graph);
    h->graph_kernels = g_launches.load() - before;
    g_launches.fetch_sub(h->graph_kernels);      // recorded, not executed
    if (rc != B200_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (e != cudaSuccess) B200_FAIL(B200_ERR_CUDA, "graph capture of the LM step failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) B200_FAIL(B200_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
  }
  B200_CUDA(cudaGraphLaunch(h->graph_exec, h->gstream));
  g_launches.fetch_add(h->graph_kernels);
  B200_CUDA(cudaEventRecord(h->ev_out, h->gstream));
  B200_CUDA(cudaStreamWaitEvent(h->stream, h->ev_out, 0));
  return B200_OK;
}

static bool sampling_on(const b200_lm* h) { return h->use_sampling && (h->temp > 0.f || h->temp_text > 0.f); }

int b200_lm_step(b200_lm* h, const int64_t* in_codes_dev, int n_in, const float* noise_dev, int64_t* out_tokens_dev,
                 int support_out_of_sync, int* ready_host) {
  return b200_lm_step_ex(h, in_codes_dev, n_in, noise_dev, nullptr, out_tokens_dev, support_out_of_sync, ready_host);
}

int b200_lm_step_ex(b200_lm* h, const int64_t* in_codes_dev, int n_in, const float* noise_dev, const int64_t* replace_audio_dev,
                    int64_t* out_tokens_dev, int support_out_of_sync, int* ready_host) {
  B200_TRY(ensure_streaming(h, "lm_step"));
  DeviceGuard g(h->device);
  const auto& c = h->cfg;
  const int B = h->batch, needed = c.n_q - c.dep_q;
  if (!in_codes_dev || !out_tokens_dev) B200_FAIL(B200_ERR_SHAPE, "lm_step: null buffers");
  if (n_in < needed || n_in > c.n_q)   // lm.py:683-686 assertion
    B200_FAIL(B200_ERR_SHAPE, "lm_step: expected at least %d user codebooks, got %d", needed, n_in);
  if (replace_audio_dev && c.dep_q == 0) B200_FAIL(B200_ERR_INVALID, "lm_step: depformer_replace_tokens on a model without depformer");
  B200_CUDA(cudaMemcpyAsync(h->in_codes, in_codes_dev, (size_t)B * n_in * 8, cudaMemcpyDeviceToDevice, h->stream));
  if (noise_dev)
    B200_CUDA(cudaMemcpyAsync(h->noise, noise_dev, (size_t)B * noise_per_row(h) * 4, cudaMemcpyDeviceToDevice, h->stream));
  if (replace_audio_dev)
    B200_CUDA(cudaMemcpyAsync(h->replace_tokens, replace_audio_dev, (size_t)B * c.dep_q * 8, cudaMemcpyDeviceToDevice, h->stream));
  B200_TRY(run_step(h, n_in, replace_audio_dev ? 1 : 0, (!noise_dev && sampling_on(h)) ? 1 : 0));
  B200_CUDA(cudaMemcpyAsync(out_tokens_dev, h->out_tokens, (size_t)B * (c.dep_q + 1) * 8, cudaMemcpyDeviceToDevice,
                            h->stream));
  h->offset_cpu += 1;
  if (ready_host) *ready_host = (support_out_of_sync || h->offset_cpu > h->max_delay) ? 1 : 0;   // lm.py:774-776
  return B200_OK;
}

int b200_lm_step_host(b200_lm* h, const int64_t* in_codes_host, int n_in, const float* noise_host,
                      int64_t* out_tokens_host, int support_out_of_sync, int* ready_host) {
  B200_TRY(ensure_streaming(h, "lm_step_host"));
  DeviceGuard g(h->device);
  const auto& c = h->cfg;
  const int B = h->batch, needed = c.n_q - c.dep_q;
  if (!in_codes_host || !out_tokens_host) B200_FAIL(B200_ERR_SHAPE, "lm_step_host: null buffers");
  if (n_in < needed || n_in > c.n_q) B200_FAIL(B200_ERR_SHAPE, "lm_step_host: expected at least %d user codebooks, got %d", needed, n_in);
  memcpy(h->pin_in, in_codes_host, (size_t)B * n_in * 8);
  B200_CUDA(cudaMemcpyAsync(h->in_codes, h->pin_in, (size_t)B * n_in * 8, cudaMemcpyHostToDevice, h->stream));
  if (noise_host) {
    memcpy(h->pin_noise, noise_host, (size_t)B * noise_per_row(h) * 4);
    B200_CUDA(cudaMemcpyAsync(h->noise, h->pin_noise, (size_t)B * noise_per_row(h) * 4, cudaMemcpyHostToDevice, h->stream));
  }
  B200_TRY(run_step(h, n_in, 0, (!noise_host && sampling_on(h)) ? 1 : 0));
  B200_CUDA(cudaMemcpyAsync(h->pin_out, h->out_tokens, (size_t)B * (c.dep_q + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
  int flags = 0;
  B200_CUDA(cudaMemcpyAsync(&flags, h->err, 4, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  memcpy(out_tokens_host, h->pin_out, (size_t)B * (c.dep_q + 1) * 8);
  h->offset_cpu += 1;
  if (ready_host) *ready_host = (support_out_of_sync || h->offset_cpu > h->max_delay) ? 1 : 0;
  if (flags) {
    cudaMemsetAsync(h->err, 0, 4, h->stream);
    if (flags & lm::ERR_KV_CAPACITY)
      B200_FAIL(B200_ERR_STATE, "lm_step_host: a session stepped past its %d-slot KV ring (b200_lm_set_kv_capacity; flags %d)", ring_cap(h), flags);
    B200_FAIL(B200_ERR_INVALID, "lm_step_host: a token id outside its embedding table reached the model (flags %d)", flags);
  }
  return B200_OK;
}

int b200_lm_read_buffer(b200_lm* h, const char* name, void* dst_dev, int64_t capacity_bytes, int64_t* nbytes) {
  B200_TRY(ensure_streaming(h, "lm_read_buffer"));
  DeviceGuard g(h->device);
  const auto& c = h->cfg;
  const std::string n = name ? name : "";
  const int B = h->batch, MB = h->MB;
  const bool cfg = cfg_on(h);
  const void* src = nullptr;
  int64_t sz = 0;
  // with CFG "text_logits" / "dep_logits" are what the samplers read (the guided rows, lm.py:728-732, 828-833); the model's
  // own 2B rows are "model_text_logits" / "model_dep_logits"
  if (n == "text_logits") { src = (cfg && !h->cfg_is_no_text) ? h->text_logits_cfg : h->text_logits; sz = (int64_t)B * c.text_card * 2; }
  else if (n == "model_text_logits") { src = h->text_logits; sz = (int64_t)MB * c.text_card * 2; }
  else if (n == "transformer_out") { src = h->tout; sz = (int64_t)MB * c.dim * 2; }
  else if (n == "dep_logits") {
    if (cfg) { src = h->dep_logits_cfg; sz = (int64_t)c.dep_q * B * c.card * 2; }
    else { src = h->dep_logits; sz = (int64_t)c.dep_q * B * c.card * 2; }
  }
  else if (n == "model_dep_logits") { src = h->dep_logits; sz = (int64_t)c.dep_q * MB * c.card * 2; }
  else if (n == "input_tokens") { src = h->input_tokens; sz = (int64_t)MB * h->Kc * 8; }
  else if (n == "text_token") { src = h->text_token; sz = (int64_t)B * 8; }
  else if (n == "audio_tokens") { src = h->audio_tokens; sz = (int64_t)c.dep_q * B * 8; }
  else if (n == "dep_trace") { src = h->dep_trace; sz = h->dep_trace ? (int64_t)tc::DEP_TRACE_SLOTS * 8 : 0; }
  else if (n == "extra_heads") { src = h->extra_out; sz = (int64_t)h->extra_w.size() * MB * c.extra_heads_dim * 2; }
  else B200_FAIL(B200_ERR_INVALID, "lm_read_buffer: unknown buffer '%s'", n.c_str());
  if (nbytes) *nbytes = sz;
  if (!dst_dev) return B200_OK;
  if (capacity_bytes < sz) B200_FAIL(B200_ERR_SHAPE, "lm_read_buffer: destination too small");
  if (sz > 0) B200_CUDA(cudaMemcpyAsync(dst_dev, src, (size_t)sz, cudaMemcpyDeviceToDevice, h->stream));
  return B200_OK;
}

int64_t b200_lm_algorithmic_bytes(b200_lm* h, int kv_fill) {
  if (!h || h->batch <= 0) return 0;
  const auto& c = h->cfg;
  const int64_t B = h->MB;
  if (kv_fill > c.context) kv_fill = c.context;
  // weights streamed once per step (embedding tables are gathers, counted per row below)
  int64_t w = 0;
  w += (int64_t)c.num_layers * ((int64_t)3 * c.dim * c.dim + (int64_t)c.dim * c.dim + (int64_t)3 * c.ffn_hidden * c.dim) * 2;
  w += (int64_t)c.text_card * c.dim * 2;
  w += (int64_t)c.dep_q * c.depformer_dim * c.dim * 2;
  w += (int64_t)c.dep_q * c.depformer_num_layers *
       ((int64_t)4 * c.depformer_dim * c.depformer_dim + (int64_t)3 * c.depformer_ffn_hidden * c.depformer_dim) * 2;
  w += (int64_t)c.dep_q * c.card * c.depformer_dim * 2;
  if (c.quantize) w /= 2;        // one byte per weight (the fp32 row scales are < 0.1 % of that)
  // per session: KV ring read (valid slots only) + append, embedding rows, logits written + read by the sampler
  const int64_t kv_row = h->kv_fp8 ? (int64_t)c.dim + 4 * c.num_heads : (int64_t)c.dim * 2;      // bytes of one K (or V) row of a layer
  int64_t per = (int64_t)c.num_layers * 2 * kv_fill * kv_row + (int64_t)c.num_layers * 2 * kv_row;
  per += (int64_t)(c.n_q + 1) * c.dim * 2 + (int64_t)c.dep_q * c.depformer_dim * 2;
  per += (int64_t)2 * (c.text_card + (int64_t)c.dep_q * c.card) * 2;
  return w + per * B;
}

}  // extern "C"
