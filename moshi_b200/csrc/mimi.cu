// Mimi handle: weight ingestion (reference state-dict names), streaming state, encode / decode.
// Reference orchestration: moshi/moshi/models/compression.py:338-433.
//
// Everything between the PCM frame and the latent is TOKEN-MAJOR ([session][step][channel]): every SEANet conv / transposed
// conv and every transformer linear is one launch of mimi_tc_kernel (TMA + tcgen05 kind::tf32 with 3xTF32 split products,
// mimi_tc.cuh); a layer's epilogue writes its consumer's activation (ELU applied, hi / lo pair) straight behind the
// consumer's carried left context, so that `cat(previous, x)` (conv.py:261) is an address, the streaming-state update
// (conv.py:263-267) is a P-row shift inside that buffer, and ProjectedTransformer.conv_layout (transformer.py:972-981)
// costs nothing.  The degenerate convs (Cin = 1, Cout = 1), the learnt 2x down-sampling conv (replicate padding), the
// depth-wise up-sampling, LayerNorm / RoPE / ring attention and the residual VQ are small SIMT kernels.
#include "mimi_kernels.cuh"
#include "mimi_tc.cuh"
#include "mimi_tm_kernels.cuh"

using namespace b200;
using namespace b200::mimi;

namespace {

struct SeaLayer {
  int kind = 0;              // 0 = StreamingConv1d, 1 = StreamingConvTranspose1d
  std::string key;           // state-dict prefix of the nn.Conv1d / nn.ConvTranspose1d
  std::string tap;           // debug name of the raw output (empty: not stored)
  int cin = 0, cout = 0, k = 0, stride = 1, dil = 1;
  bool elu_in = false;
  int res_from = -1;         // index of the layer whose raw output is added in the epilogue (residual block second conv)
  bool simt_first = false, simt_last = false;   // Cin = 1 / Cout = 1: memory-bound SIMT kernels
  // per-frame geometry
  int t_in = 0, t_out = 0;   // steps per frame in / out
  int P = 0;                 // carried input rows: conv (k-1)*dil + 1 - stride; convtr 1 (the previous input step)
  int N = 0;                 // GEMM output features: conv Cout; convtr S * Cout, ordered (phase, channel)
  // weights
  float* w_simt = nullptr;   // first conv [Cout][K]; last conv [K][Cin]
  float* bias = nullptr;
  // streaming buffers
  float *ext_hi = nullptr, *ext_lo = nullptr;   // [B][P + t_in][cin] (SIMT layers: ext_hi is plain fp32, ext_lo null)
  float* y = nullptr;                           // raw output, token-major [B][t_out * (convtr: S)][cout]
  mtc::TcLayer tc;
};

struct TrLayer {
  const float *n1w, *n1b, *n2w, *n2b, *ls1, *ls2;
  mtc::TcLayer in_proj, out_proj, l1, l2;
  float *kc = nullptr, *vc = nullptr;
};

struct Transformer {
  std::vector<TrLayer> layers;
  long long* offset = nullptr;   // [B] tokens seen (== RingKVCache.end_offset == _MHAState.offset)
};

struct Tap { const float* p; int C, T; bool transpose; };

struct DownLayer {             // ConvDownsample1d (resample.py:14-65): k = 2 * stride, no bias, replicate padding
  std::string key;
  int cin = 0, cout = 0, k = 0, stride = 1;
  int t_in = 0, t_out = 0, P = 0;
  float* w = nullptr;          // [Cout][K * Cin]
  float* state = nullptr;      // previous [B][Cin][P]
  uint8_t* first = nullptr;    // replicate flags [B]
};     // transpose: stored [B][T][C], reported [B][C][T]

}  // namespace

struct b200_mimi {
  b200_mimi_config cfg;
  TensorStore store;
  Arena weights, state;
  bool finalized = false;
  int device = -1;
  int batch = 0;
  int num_codebooks = 8;
  cudaStream_t stream = nullptr;
  int frame_size = 0, hop = 0, rs = 0;       // rs = resample stride (2)
  int sms = 148;

  std::vector<SeaLayer> enc, dec;
  DownLayer down;                            // learnt down-sampling conv (replicate padding): first-generation SIMT kernel
  // up-sampling (depth-wise convtr)
  float *up_w = nullptr, *up_partial = nullptr, *up_scratch = nullptr;
  Transformer enc_tr, dec_tr;
  // quantizer
  float *wT[2] = {nullptr, nullptr}, *woT[2] = {nullptr, nullptr};
  float *cb[2] = {nullptr, nullptr}, *cbT[2] = {nullptr, nullptr}, *cnorm[2] = {nullptr, nullptr};
  int stored_levels[2] = {0, 0};
  // per-batch buffers
  float *tok_in_enc = nullptr, *tok_enc = nullptr, *latent = nullptr;     // [B][T][C], [B][T][C], [B][C]
  float *latent_q = nullptr, *tok_in_dec = nullptr, *tok_dec = nullptr;
  float *tr_xn_hi = nullptr, *tr_xn_lo = nullptr, *tr_qkv = nullptr, *tr_ao_hi = nullptr,
        *tr_ao_lo = nullptr, *tr_h_hi = nullptr, *tr_h_lo = nullptr;
  uint8_t* exec_mask = nullptr;
  uint8_t* first_flags = nullptr; int n_first = 0;
  ConvCommit* enc_commits = nullptr; int n_enc_commits = 0, max_enc_rows = 0;
  ConvTrCommit* dec_tr_commits = nullptr; int n_dec_tr_commits = 0; long long max_tr_rows = 0;
  TmCommit *enc_tm = nullptr, *dec_tm = nullptr; int n_enc_tm = 0, n_dec_tm = 0; long long max_enc_tm = 0, max_dec_tm = 0;
  long long* scratch_codes = nullptr;          // [B][K][1] for the host variants
  float* rvq_res[2] = {nullptr, nullptr};      // RVQ workspace: residuals, per-chunk partial argmin
  float* rvq_best[2] = {nullptr, nullptr};
  int* rvq_idx[2] = {nullptr, nullptr};
  int rvq_cap = 0;
  float* splitk_ws = nullptr; size_t splitk_bytes = 0;   // split-K partial sums of the deep / skinny layers
  int* err = nullptr;                                     // device error flags (B200_FLAG_CODE_RANGE)
  // one-frame encode / decode as CUDA graphs over static buffers (frame in enc[0].ext -> enc_codes, dec_codes -> out_frame)
  cudaStream_t body = nullptr;                 // stream the kernels are being enqueued on (caller's, or gstream in capture)
  cudaStream_t gstream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaGraphExec_t enc_graph = nullptr, dec_graph = nullptr;
  int64_t enc_graph_kernels = 0, dec_graph_kernels = 0;
  int dec_graph_ncb = 0;
  int graph_enabled = 1;
  long long *enc_codes = nullptr, *dec_codes = nullptr;   // [B][K], [B][q_n_q]
  float* out_frame = nullptr;                              // [B][frame_size]
  float *pin_pcm = nullptr; long long* pin_codes = nullptr; size_t pin_pcm_n = 0, pin_codes_n = 0;
  float* dev_pcm = nullptr; long long* dev_codes = nullptr; size_t dev_pcm_n = 0, dev_codes_n = 0;
  float* tap_scratch = nullptr; size_t tap_scratch_n = 0;
  std::map<std::string, Tap> taps;
  int64_t weight_bytes = 0;
};

namespace {

// ---------------------------------------------------------------------------------------------
// structure (seanet.py:170-236, 323-388; resample.py; loaders.py:38-88)
// ---------------------------------------------------------------------------------------------
void plan_seanet(b200_mimi* h) {
  const auto& c = h->cfg;
  h->enc.clear();
  h->dec.clear();
  auto conv = [](const std::string& key, int cin, int cout, int k, int stride, int dil, bool elu) {
    SeaLayer l;
    l.kind = 0; l.key = key; l.cin = cin; l.cout = cout; l.k = k; l.stride = stride; l.dil = dil; l.elu_in = elu;
    return l;
  };
  int idx = 0, mult = 1;
  {
    auto l = conv("encoder.model.0.conv.conv", c.channels, mult * c.n_filters, c.kernel_size, 1, 1, false);
    l.tap = "enc.0";
    h->enc.push_back(l);
  }
  idx = 1;
  for (int ri = c.n_ratios - 1; ri >= 0; --ri) {
    const int ratio = c.ratios[ri];
    const int ch = mult * c.n_filters;
    for (int j = 0; j < c.n_residual_layers; ++j) {
      int dil = 1;
      for (int q = 0; q < j; ++q) dil *= c.dilation_base;
      const std::string base = "encoder.model." + std::to_string(idx);
      auto a = conv(base + ".block.1.conv.conv", ch, ch / c.compress, c.residual_kernel_size, 1, dil, true);
      auto b = conv(base + ".block.3.conv.conv", ch / c.compress, ch, 1, 1, 1, true);
      b.res_from = (int)h->enc.size() - 1;      // skip connection: the input of the block = the previous module's raw output
      b.tap = "enc." + std::to_string(idx);
      h->enc.push_back(a);
      h->enc.push_back(b);
      ++idx;
    }
    ++idx;  // ELU
    auto l = conv("encoder.model." + std::to_string(idx) + ".conv.conv", ch, ch * 2, 2 * ratio, ratio, 1, true);
    l.tap = "enc." + std::to_string(idx);
    h->enc.push_back(l);
    ++idx;
    mult *= 2;
  }
  ++idx;  // ELU
  {
    auto l = conv("encoder.model." + std::to_string(idx) + ".conv.conv", mult * c.n_filters, c.dimension,
                  c.last_kernel_size, 1, 1, true);
    l.tap = "enc." + std::to_string(idx);
    h->enc.push_back(l);
  }

  idx = 0;
  mult = 1 << c.n_ratios;
  {
    auto l = conv("decoder.model.0.conv.conv", c.dimension, mult * c.n_filters, c.kernel_size, 1, 1, false);
    l.tap = "dec.0";
    h->dec.push_back(l);
  }
  idx = 1;
  for (int ri = 0; ri < c.n_ratios; ++ri) {
    const int ratio = c.ratios[ri];
    const int ch = mult * c.n_filters;
    ++idx;  // ELU
    SeaLayer t;
    t.kind = 1; t.key = "decoder.model." + std::to_string(idx) + ".convtr.convtr";
    t.cin = ch; t.cout = ch / 2; t.k = 2 * ratio; t.stride = ratio; t.elu_in = true;
    t.tap = "dec." + std::to_string(idx);
    h->dec.push_back(t);
    ++idx;
    for (int j = 0; j < c.n_residual_layers; ++j) {
      int dil = 1;
      for (int q = 0; q < j; ++q) dil *= c.dilation_base;
      const std::string base = "decoder.model." + std::to_string(idx);
      auto a = conv(base + ".block.1.conv.conv", ch / 2, ch / 2 / c.compress, c.residual_kernel_size, 1, dil, true);
      auto b = conv(base + ".block.3.conv.conv", ch / 2 / c.compress, ch / 2, 1, 1, 1, true);
      b.res_from = (int)h->dec.size() - 1;
      b.tap = "dec." + std::to_string(idx);
      h->dec.push_back(a);
      h->dec.push_back(b);
      ++idx;
    }
    mult /= 2;
  }
  ++idx;  // ELU
  {
    auto l = conv("decoder.model." + std::to_string(idx) + ".conv.conv", c.n_filters, c.channels,
                  c.last_kernel_size, 1, 1, true);
    l.tap = "dec." + std::to_string(idx);
    h->dec.push_back(l);
  }
  for (auto* layers : {&h->enc, &h->dec})
    for (auto& l : *layers) {
      l.simt_first = l.kind == 0 && l.cin == 1;
      l.simt_last = l.kind == 0 && l.cout == 1;
      l.N = l.kind == 1 ? l.stride * l.cout : l.cout;
      l.P = l.kind == 1 ? 1 : (l.k - 1) * l.dil + 1 - l.stride;
    }
  h->down = DownLayer();
  h->down.key = "downsample.conv.conv.conv"; h->down.cin = c.dimension; h->down.cout = c.dimension;
  h->down.k = 2 * h->rs; h->down.stride = h->rs;
}

int get_f32(b200_mimi* h, const std::string& name, std::vector<int64_t> shape, const float** out) {
  const Tensor* t = h->store.find(name);
  if (!t) B200_FAIL(B200_ERR_MISSING, "mimi finalize: tensor '%s' was not loaded", name.c_str());
  if (t->dtype != B200_F32) B200_FAIL(B200_ERR_SHAPE, "mimi tensor '%s' must be float32", name.c_str());
  if (!shape.empty() && t->shape != shape) {
    std::string got;
    for (auto s : t->shape) got += std::to_string(s) + ",";
    B200_FAIL(B200_ERR_SHAPE, "mimi tensor '%s' has shape [%s]", name.c_str(), got.c_str());
  }
  *out = static_cast<const float*>(t->data);
  return B200_OK;
}

struct ScratchFree { void* p = nullptr; ~ScratchFree() { if (p) cudaFree(p); } };

// nn.Linear-shaped weight [N][n_taps * Cin] (fp32, device) -> packed hi / lo tiles owned by the handle
int pack_tc(b200_mimi* h, mtc::TcLayer& L, const float* w_ntc, int N, int Cin, int n_taps) {
  if (Cin % mtc::TC_KB || N % 16 || (N > mtc::TC_MAX_NT && N % mtc::TC_MAX_NT))
    B200_FAIL(B200_ERR_INVALID, "mimi: a %d x %d (x %d taps) contraction does not tile on the tensor-core kernel", N, Cin, n_taps);
  L.Cin = Cin; L.N = N; L.n_taps = n_taps;
  L.NT = N >= mtc::TC_MAX_NT ? mtc::TC_MAX_NT : N;
  L.n_tiles_n = (N + L.NT - 1) / L.NT;
  L.num_kb = n_taps * (Cin / mtc::TC_KB);
  const size_t bytes = mtc::tc_packed_bytes(N, Cin, n_taps);
  void* out = nullptr;
  B200_TRY(h->weights.alloc(&out, bytes, false));
  B200_TRY(mtc::tc_pack_weights(w_ntc, out, N, Cin, n_taps, nullptr));
  L.wt = static_cast<uint8_t*>(out);
  h->weight_bytes += (int64_t)bytes;
  return B200_OK;
}

int pack_sea(b200_mimi* h, SeaLayer& l) {
  const float* w = nullptr;
  const long long n = (long long)l.cout * l.cin * l.k;
  if (l.kind == 0) B200_TRY(get_f32(h, l.key + ".weight", {l.cout, l.cin, l.k}, &w));
  else {
    if (l.k != 2 * l.stride) B200_FAIL(B200_ERR_INVALID, "convtr %s: kernel must be 2*stride", l.key.c_str());
    B200_TRY(get_f32(h, l.key + ".weight", {l.cin, l.cout, l.k}, &w));
  }
  if (l.simt_first) {                       // [Cout][1][K] is already [Cout][K]
    if (l.k > 8 || l.stride != 1 || l.cout % 4) B200_FAIL(B200_ERR_INVALID, "%s: unsupported first conv", l.key.c_str());
    B200_TRY(h->weights.alloc_t(&l.w_simt, (size_t)n, false));
    B200_CUDA(cudaMemcpy(l.w_simt, w, (size_t)n * 4, cudaMemcpyDeviceToDevice));
    h->weight_bytes += n * 4;
  } else if (l.simt_last) {                 // [1][Cin][K] -> [K][Cin]
    if (l.stride != 1 || l.cin % 4) B200_FAIL(B200_ERR_INVALID, "%s: unsupported last conv", l.key.c_str());
    B200_TRY(h->weights.alloc_t(&l.w_simt, (size_t)n, false));
    B200_LAUNCH(tm_weight_conv_kernel, (unsigned)ceil_div64(n, 256), 256, 0, 0, w, l.w_simt, 1, l.cin, l.k);
    h->weight_bytes += n * 4;
  } else {
    ScratchFree tmp;
    B200_CUDA(cudaMalloc(&tmp.p, (size_t)n * 4));
    float* wk = static_cast<float*>(tmp.p);
    if (l.kind == 0) B200_LAUNCH(tm_weight_conv_kernel, (unsigned)ceil_div64(n, 256), 256, 0, 0, w, wk, l.cout, l.cin, l.k);
    else B200_LAUNCH(tm_weight_convtr_kernel, (unsigned)ceil_div64(n, 256), 256, 0, 0, w, wk, l.cin, l.cout, l.stride);
    B200_TRY(pack_tc(h, l.tc, wk, l.N, l.cin, l.kind == 0 ? l.k : 2));
    B200_CUDA(cudaDeviceSynchronize());
    l.tc.kind = l.kind; l.tc.dil = l.kind == 0 ? l.dil : 1; l.tc.stride = l.kind == 0 ? l.stride : 1;
    l.tc.bias_mod = l.cout;
  }
  const float* b = nullptr;
  B200_TRY(get_f32(h, l.key + ".bias", {l.cout}, &b));
  B200_TRY(h->weights.alloc_t(&l.bias, l.cout, false));
  B200_CUDA(cudaMemcpy(l.bias, b, l.cout * 4, cudaMemcpyDeviceToDevice));
  h->weight_bytes += l.cout * 4;
  l.tc.bias = l.bias;
  B200_CUDA(cudaDeviceSynchronize());
  h->store.release(l.key + ".weight");
  h->store.release(l.key + ".bias");
  return check_launch("pack_sea");
}

// the down-sampling conv keeps the first-generation kernel and weight layout ([Cout][K*Cin])
int pack_down(b200_mimi* h, DownLayer& l) {
  const float* w = nullptr;
  B200_TRY(get_f32(h, l.key + ".weight", {l.cout, l.cin, l.k}, &w));
  const long long n = (long long)l.cout * l.cin * l.k;
  B200_TRY(h->weights.alloc_t(&l.w, n, false));
  B200_LAUNCH(pack_conv_w_kernel, (unsigned)ceil_div64(n, 256), 256, 0, 0, w, l.w, l.cout, l.cin, l.k);
  h->weight_bytes += n * 4;
  B200_CUDA(cudaDeviceSynchronize());
  h->store.release(l.key + ".weight");
  return check_launch("pack_down");
}

int keep(b200_mimi* h, const std::string& name, std::vector<int64_t> shape, const float** out) {
  const float* src = nullptr;
  B200_TRY(get_f32(h, name, shape, &src));
  long long n = 1;
  for (auto s : shape) n *= s;
  float* dst = nullptr;
  B200_TRY(h->weights.alloc_t(&dst, n, false));
  B200_CUDA(cudaMemcpy(dst, src, n * 4, cudaMemcpyDeviceToDevice));
  h->store.release(name);
  h->weight_bytes += n * 4;
  *out = dst;
  return B200_OK;
}

int keep_linear(b200_mimi* h, const std::string& name, int M, int K, mtc::TcLayer& L) {
  const float* src = nullptr;
  B200_TRY(get_f32(h, name, {M, K}, &src));
  B200_TRY(pack_tc(h, L, src, M, K, 1));       // nn.Linear weight [out][in] is already [N][K]
  L.kind = 2; L.dil = 1; L.stride = 1; L.bias = nullptr; L.bias_mod = M;
  B200_CUDA(cudaDeviceSynchronize());
  h->store.release(name);
  return B200_OK;
}

int pack_transformer(b200_mimi* h, const std::string& prefix, Transformer& tr) {
  const auto& c = h->cfg;
  const int d = c.tr_d_model, ff = c.tr_dim_feedforward;
  tr.layers.resize(c.tr_num_layers);
  for (int li = 0; li < c.tr_num_layers; ++li) {
    const std::string p = prefix + ".transformer.layers." + std::to_string(li);
    TrLayer& L = tr.layers[li];
    B200_TRY(keep_linear(h, p + ".self_attn.in_projs.0.weight", 3 * d, d, L.in_proj));
    B200_TRY(keep_linear(h, p + ".self_attn.out_projs.0.weight", d, d, L.out_proj));
    B200_TRY(keep(h, p + ".norm1.weight", {d}, &L.n1w));
    B200_TRY(keep(h, p + ".norm1.bias", {d}, &L.n1b));
    B200_TRY(keep(h, p + ".norm2.weight", {d}, &L.n2w));
    B200_TRY(keep(h, p + ".norm2.bias", {d}, &L.n2b));
    B200_TRY(keep_linear(h, p + ".linear1.weight", ff, d, L.l1));
    B200_TRY(keep_linear(h, p + ".linear2.weight", d, ff, L.l2));
    B200_TRY(keep(h, p + ".layer_scale_1.scale", {d}, &L.ls1));
    B200_TRY(keep(h, p + ".layer_scale_2.scale", {d}, &L.ls2));
  }
  return B200_OK;
}

int pack_quantizer(b200_mimi* h) {
  const auto& c = h->cfg;
  const char* names[2] = {"rvq_first", "rvq_rest"};
  h->stored_levels[0] = c.q_n_semantic;
  h->stored_levels[1] = c.q_n_q - c.q_n_semantic;
  for (int w = 0; w < 2; ++w) {
    const std::string p = std::string("quantizer.") + names[w];
    const float* src = nullptr;
    B200_TRY(get_f32(h, p + ".input_proj.weight", {c.q_dimension, c.dimension, 1}, &src));
    B200_TRY(h->weights.alloc_t(&h->wT[w], (size_t)c.q_dimension * c.dimension, false));
    B200_LAUNCH(transpose_kernel, (unsigned)ceil_div64((long long)c.q_dimension * c.dimension, 256), 256, 0, 0, src,
                h->wT[w], c.q_dimension, c.dimension);
    B200_TRY(get_f32(h, p + ".output_proj.weight", {c.dimension, c.q_dimension, 1}, &src));
    B200_TRY(h->weights.alloc_t(&h->woT[w], (size_t)c.q_dimension * c.dimension, false));
    B200_LAUNCH(transpose_kernel, (unsigned)ceil_div64((long long)c.q_dimension * c.dimension, 256), 256, 0, 0, src,
                h->woT[w], c.dimension, c.q_dimension);
    const int L = h->stored_levels[w];
    const size_t per = (size_t)c.q_bins * c.q_dimension;
    B200_TRY(h->weights.alloc_t(&h->cb[w], per * L, false));
    B200_TRY(h->weights.alloc_t(&h->cbT[w], per * L, false));
    B200_TRY(h->weights.alloc_t(&h->cnorm[w], (size_t)c.q_bins * L, false));
    for (int l = 0; l < L; ++l) {
      const std::string cbp = p + ".vq.layers." + std::to_string(l) + "._codebook";
      const float *es = nullptr, *us = nullptr;
      B200_TRY(get_f32(h, cbp + ".embedding_sum", {c.q_bins, c.q_dimension}, &es));
      B200_TRY(get_f32(h, cbp + ".cluster_usage", {c.q_bins}, &us));
      B200_LAUNCH(build_codebook_kernel, c.q_bins, 128, 0, 0, es, us, h->cb[w] + per * l, h->cbT[w] + per * l,
                  h->cnorm[w] + (size_t)c.q_bins * l, c.q_bins, c.q_dimension);
    }
    B200_CUDA(cudaDeviceSynchronize());
    for (int l = 0; l < L; ++l) {
      const std::string cbp = p + ".vq.layers." + std::to_string(l) + "._codebook";
      h->store.release(cbp + ".embedding_sum");
      h->store.release(cbp + ".cluster_usage");
      h->store.release(cbp + "._initialized");
    }
    h->store.release(p + ".input_proj.weight");
    h->store.release(p + ".output_proj.weight");
  }
  return check_launch("pack_quantizer");
}

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
int launch_tc(b200_mimi* h, const mtc::TcLayer& L) { return mtc::tc_launch(L, h->body); }

// StreamingTransformer.forward for T tokens per row (transformer.py:894-929, layer :752-802)
int run_transformer(b200_mimi* h, Transformer& tr, const float* x_in, float* x, int T) {
  const auto& c = h->cfg;
  const int B = h->batch, ntok = B * T, d = c.tr_d_model, H = c.tr_num_heads, D = d / H;
  const float nl = -logf(c.tr_max_period) * 2.f / (float)D;
  const size_t attn_smem = (size_t)(T * D + T * c.tr_context) * sizeof(float);
  for (size_t li = 0; li < tr.layers.size(); ++li) {
    TrLayer& L = tr.layers[li];
    const float* cur = li == 0 ? x_in : x;
    B200_LAUNCH(layernorm_split_kernel, ceil_div(ntok * 32, 256), 256, 0, h->body, cur, L.n1w, L.n1b, h->tr_xn_hi, h->tr_xn_lo, ntok, d, 1e-5f);
    B200_TRY(launch_tc(h, L.in_proj));
    B200_LAUNCH((ring_attn_step_kernel<64>), B * H, 128, attn_smem, h->body, h->tr_qkv, L.kc, L.vc, h->tr_ao_hi, h->tr_ao_lo, tr.offset,
                h->exec_mask, T, H, c.tr_context, c.tr_context, nl);
    B200_TRY(launch_tc(h, L.out_proj));       // x = cur + layer_scale_1 * out_proj(attention)
    B200_LAUNCH(layernorm_split_kernel, ceil_div(ntok * 32, 256), 256, 0, h->body, x, L.n2w, L.n2b, h->tr_xn_hi, h->tr_xn_lo, ntok, d, 1e-5f);
    B200_TRY(launch_tc(h, L.l1));             // h = gelu(linear1(xn)) as a hi / lo pair
    B200_TRY(launch_tc(h, L.l2));             // x = x + layer_scale_2 * linear2(h)
  }
  B200_LAUNCH(advance_offsets_kernel, ceil_div(B, 128), 128, 0, h->body, tr.offset, h->exec_mask, B, T);
  return check_launch("mimi transformer");
}

int run_seanet(b200_mimi* h, std::vector<SeaLayer>& layers) {
  const int B = h->batch;
  for (size_t i = 0; i < layers.size(); ++i) {
    SeaLayer& l = layers[i];
    if (l.simt_first) {
      const SeaLayer& nx = layers[i + 1];
      ConvFirst q;
      q.ext = l.ext_hi; q.E = l.P + l.t_in; q.P = l.P; q.w = l.w_simt; q.bias = l.bias;
      q.y = l.y; q.y_sb = (long long)l.t_out * l.cout;
      q.a_hi = nx.ext_hi + (long long)nx.P * nx.cin; q.a_lo = nx.ext_lo + (long long)nx.P * nx.cin;
      q.a_sb = (long long)(nx.P + nx.t_in) * nx.cin; q.a_elu = nx.elu_in;
      q.B = B; q.Cout = l.cout; q.K = l.k; q.T = l.t_out;
      const long long n = (long long)B * l.t_out;
      if (l.cout % 4 || 128 % (l.cout / 4) || l.k > 8) B200_FAIL(B200_ERR_INVALID, "mimi: first conv with %d channels / %d taps", l.cout, l.k);
      const int per_pass = 128 / (l.cout / 4);
      q.tok_per_block = per_pass * 16;           // 16 passes per CTA: the 4 x K weights of a thread are loaded once
      B200_LAUNCH(conv_first_tm_kernel, (unsigned)ceil_div64(n, q.tok_per_block), 128, 0, h->body, q);
    } else if (l.simt_last) {
      ConvLast q;
      q.ext = l.ext_hi; q.e_sb = (long long)(l.P + l.t_in) * l.cin; q.Cin = l.cin; q.w = l.w_simt; q.bias = l.bias;
      q.y = l.y; q.y_sb = l.t_out; q.B = B; q.K = l.k; q.dil = l.dil; q.T = l.t_out;
      const long long n = (long long)B * l.t_out;
      B200_LAUNCH(conv_last_tm_kernel, (unsigned)ceil_div64(n, 128), 128, (size_t)l.k * l.cin * 4, h->body, q);
    } else {
      B200_TRY(launch_tc(h, l.tc));
    }
  }
  return check_launch("mimi seanet");
}

int commit_states(b200_mimi* h, bool encoder) {
  const int B = h->batch;
  if (encoder) {
    if (h->n_enc_commits) {      // the down-sampling conv (first-generation state layout + replicate flags)
      dim3 grid(ceil_div(h->max_enc_rows, 128), h->n_enc_commits);
      B200_LAUNCH(conv_commit_kernel, grid, 128, 0, h->body, h->enc_commits, h->n_enc_commits, h->exec_mask, B);
      B200_LAUNCH(conv_clear_first_kernel, ceil_div(h->n_enc_commits * B, 128), 128, 0, h->body, h->enc_commits,
                  h->n_enc_commits, h->exec_mask, B);
    }
    if (h->n_enc_tm) {
      dim3 grid((unsigned)ceil_div64(h->max_enc_tm, 128), h->n_enc_tm);
      B200_LAUNCH(tm_commit_kernel, grid, 128, 0, h->body, h->enc_tm, h->exec_mask, B);
    }
  } else {
    if (h->n_dec_tm) {
      dim3 grid((unsigned)ceil_div64(h->max_dec_tm, 128), h->n_dec_tm);
      B200_LAUNCH(tm_commit_kernel, grid, 128, 0, h->body, h->dec_tm, h->exec_mask, B);
    }
    dim3 g2((unsigned)ceil_div64(h->max_tr_rows, 256), h->n_dec_tr_commits);      // the depth-wise up-sampling's overlap-add carry
    B200_LAUNCH(convtr_commit_kernel, g2, 256, 0, h->body, h->dec_tr_commits, h->exec_mask, B);
  }
  return check_launch("commit_states");
}

// frame in enc[0].ext -> latent [B][C]   (static buffers only: this is what gets captured into a graph)
int encode_body(b200_mimi* h) {
  const int d = h->cfg.dimension;
  const int T = h->frame_size / h->hop;   // encoder tokens per frame (2)
  B200_TRY(run_seanet(h, h->enc));        // the last conv's raw output is tok_in_enc, token-major [B][T][C]
  B200_TRY(run_transformer(h, h->enc_tr, h->tok_in_enc, h->tok_enc, T));
  {                                       // learnt down-sampling conv reads token-major, writes latent [B][C][1]
    const DownLayer& l = h->down;
    ConvP p;
    p.x = h->tok_enc; p.xb = (long long)T * d; p.xc = 1; p.xt = d; p.Tin = l.t_in;
    p.st = l.state; p.P = l.P; p.first = l.first;
    p.w = l.w; p.bias = nullptr;
    p.y = h->latent; p.yb = d; p.yc = 1; p.yt = 1;
    p.res = nullptr; p.rb = p.rc = p.rt = 0;
    p.B = h->batch; p.Cin = l.cin; p.Cout = l.cout; p.K = l.k; p.stride = l.stride; p.dil = 1; p.Tout = l.t_out;
    p.elu_in = 0;
    p.M = l.cout; p.N = h->batch * l.t_out; p.Kd = l.cin * l.k; p.cin_aligned = (l.cin % BK) == 0;
    const int KS = ceil_div(p.Kd, DS_KP);
    if ((size_t)KS * p.N * p.M * 4 <= h->splitk_bytes) {
      // deep and skinny (2048 reduction steps, N = sessions): the reduction cut over KS CTAs per 64 outputs, slices added in order
      dim3 g1(ceil_div(p.M, DS_CO), KS);
      B200_LAUNCH(conv_splitk_kernel, g1, 256, 0, h->body, p, h->splitk_ws, KS);
      B200_LAUNCH(conv_splitk_reduce_kernel, (unsigned)ceil_div64((long long)p.N * p.M, 256), 256, 0, h->body, p, h->splitk_ws, KS);
    } else {
      dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM));
      B200_LAUNCH((igemm_f32_kernel<ConvP, false>), grid, 256, 0, h->body, p);
    }
  }
  B200_TRY(commit_states(h, true));
  return B200_OK;
}

// the PCM frame lands right behind the first conv's carried samples
int load_frame(b200_mimi* h, const float* pcm, int n_frames, int f) {
  const int fs = h->frame_size;
  const SeaLayer& l = h->enc[0];
  B200_CUDA(cudaMemcpy2DAsync(l.ext_hi + l.P, (size_t)(l.P + fs) * 4, pcm + (size_t)f * fs, (size_t)fs * n_frames * 4,
                              (size_t)fs * 4, h->batch, cudaMemcpyDeviceToDevice, h->stream));
  return B200_OK;
}

// RVQ workspace (residuals + per-chunk partial argmin); sized at streaming_begin for Q = batch and
// grown only by the multi-frame `quantize` entry point.
int ensure_rvq_workspace(b200_mimi* h, int Q) {
  if (Q <= h->rvq_cap) return B200_OK;
  const int Dq = h->cfg.q_dimension, n_chunks = ceil_div(h->cfg.q_bins, RVQ_CHUNK);
  if (h->stream) B200_CUDA(cudaStreamSynchronize(h->stream));
  for (int w = 0; w < 2; ++w) {
    if (h->rvq_res[w]) cudaFree(h->rvq_res[w]);
    if (h->rvq_best[w]) cudaFree(h->rvq_best[w]);
    if (h->rvq_idx[w]) cudaFree(h->rvq_idx[w]);
    B200_CUDA(cudaMalloc(&h->rvq_res[w], (size_t)Q * Dq * 4));
    B200_CUDA(cudaMalloc(&h->rvq_best[w], (size_t)Q * n_chunks * 4));
    B200_CUDA(cudaMalloc(&h->rvq_idx[w], (size_t)Q * n_chunks * 4));
  }
  h->rvq_cap = Q;
  return B200_OK;
}

int quantize_cols(b200_mimi* h, const float* lat, long long lb, long long lc, long long lt, int n_cols,
                  long long* codes, long long cs_b, long long cs_k, long long cs_f) {
  const auto& c = h->cfg;
  const int Q = h->batch * n_cols, Dq = c.q_dimension, bins = c.q_bins;
  const int n_chunks = ceil_div(bins, RVQ_CHUNK);
  B200_TRY(ensure_rvq_workspace(h, Q));
  int levels[2];
  levels[0] = c.q_n_semantic < h->num_codebooks ? c.q_n_semantic : h->num_codebooks;
  levels[1] = h->num_codebooks - levels[0];
  {
    dim3 grid(ceil_div(Q, RVQ_PQ), 2, ceil_div(Dq, 64));
    B200_LAUNCH(rvq_project_kernel, grid, 256, (size_t)(RVQ_PQ * c.dimension + 4 * RVQ_PQ * 64) * 4, h->body, lat, lb, lc, lt, n_cols, Q,
                h->wT[0], h->wT[1], h->rvq_res[0], h->rvq_res[1], c.dimension, Dq);
  }
  const int max_levels = levels[0] > levels[1] ? levels[0] : levels[1];
  const size_t per = (size_t)bins * Dq;
  for (int level = 0; level < max_levels; ++level) {
    RvqLevelArgs a;
    for (int w = 0; w < 2; ++w) {
      a.res[w] = h->rvq_res[w];
      a.cbT[w] = h->cbT[w] + per * level; a.cb[w] = h->cb[w] + per * level; a.cnorm[w] = h->cnorm[w] + (size_t)bins * level;
      a.part_best[w] = h->rvq_best[w]; a.part_idx[w] = h->rvq_idx[w];
      a.active[w] = level < levels[w];
      a.code_index[w] = (w == 0 ? 0 : levels[0]) + level;
    }
    a.codes = codes; a.cs_b = cs_b; a.cs_k = cs_k; a.cs_f = cs_f;
    a.n_query = Q; a.n_frames = n_cols; a.Dq = Dq; a.bins = bins; a.n_chunks = n_chunks;
    dim3 g1(n_chunks, ceil_div(Q, RVQ_QT), 2);
    B200_LAUNCH(rvq_search_kernel, g1, RVQ_CHUNK, (size_t)RVQ_QT * Dq * 4, h->body, a);
    dim3 g2(Q, 2);
    B200_LAUNCH(rvq_pick_kernel, g2, 128, 0, h->body, a);
  }
  return check_launch("rvq_encode");
}

int dequantize_cols(b200_mimi* h, const long long* codes, long long cs_b, long long cs_k, long long cs_f, int n_codebooks,
                    int n_cols, float* out, long long ob, long long oc, long long ot) {
  const auto& c = h->cfg;
  if (n_codebooks < 1 || n_codebooks > c.q_n_q) B200_FAIL(B200_ERR_SHAPE, "decode: %d codebooks", n_codebooks);
  RvqDecArgs a;
  a.codes = codes; a.cs_b = cs_b; a.cs_k = cs_k; a.cs_f = cs_f; a.n_frames = n_cols;
  for (int w = 0; w < 2; ++w) { a.cb[w] = h->cb[w]; a.woT[w] = h->woT[w]; }
  a.levels[0] = c.q_n_semantic < n_codebooks ? c.q_n_semantic : n_codebooks;
  a.levels[1] = n_codebooks - a.levels[0];
  a.level_offset[0] = 0; a.level_offset[1] = a.levels[0];
  a.out = out; a.ob = ob; a.oc = oc; a.ot = ot;
  a.Dq = c.q_dimension; a.Cout = c.dimension; a.bins = c.q_bins; a.err = h->err;
  dim3 grid(h->batch * n_cols, (c.dimension + 63) / 64);
  B200_LAUNCH(rvq_decode_kernel, grid, 256, (2 * c.q_dimension + 256) * sizeof(float), h->body, a);
  return check_launch("rvq_decode");
}

// latent_q [B][C] -> out_frame [B][frame]   (static buffers only)
int decode_latent_body(b200_mimi* h) {
  const int B = h->batch, d = h->cfg.dimension, S = h->rs;
  const int T = S;   // tokens per frame after up-sampling
  {
    const long long total = (long long)B * 2 * S * d;   // (T_in + 1) * S * C with T_in = 1
    B200_LAUNCH(upsample_dw_kernel, (unsigned)ceil_div64(total, 256), 256, 0, h->body, h->latent_q, (long long)d, 1LL, 1LL, 1,
                h->up_w, h->up_partial, h->up_scratch, h->tok_in_dec, B, d, S);
  }
  B200_TRY(run_transformer(h, h->dec_tr, h->tok_in_dec, h->tok_dec, T));   // its last linear also fills dec[0]'s ext
  B200_TRY(run_seanet(h, h->dec));
  B200_TRY(commit_states(h, false));
  return B200_OK;
}

void drop_graphs(b200_mimi* h) {
  if (h->enc_graph) cudaGraphExecDestroy(h->enc_graph);
  if (h->dec_graph) cudaGraphExecDestroy(h->dec_graph);
  h->enc_graph = h->dec_graph = nullptr;
}

// Runs `body` (a fixed launch sequence over static buffers) as one CUDA graph on the private stream, fenced against
// the caller's stream on both sides; the first call captures it.  With graphs disabled the body runs eagerly.
template <class F>
int run_captured(b200_mimi* h, cudaGraphExec_t* exec, int64_t* n_kernels, F body) {
  if (!h->graph_enabled) {
    h->body = h->stream;
    return body();
  }
  B200_CUDA(cudaEventRecord(h->ev_in, h->stream));
  B200_CUDA(cudaStreamWaitEvent(h->gstream, h->ev_in, 0));
  if (!*exec) {
    cudaGraph_t graph = nullptr;
    const int64_t before = g_launches.load();
    h->body = h->gstream;
    B200_CUDA(cudaStreamBeginCapture(h->gstream, cudaStreamCaptureModeRelaxed));
    const int rc = body();
    cudaError_t e = cudaStreamEndCapture(h->gstream, &graph);
    *n_kernels = g_launches.load() - before;
    g_launches.fetch_sub(*n_kernels);      // recorded, not executed
    h->body = h->stream;
    if (rc != B200_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (e != cudaSuccess) B200_FAIL(B200_ERR_CUDA, "graph capture of a Mimi frame failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) B200_FAIL(B200_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
  }
  B200_CUDA(cudaGraphLaunch(*exec, h->gstream));
  g_launches.fetch_add(*n_kernels);
  B200_CUDA(cudaEventRecord(h->ev_out, h->gstream));
  B200_CUDA(cudaStreamWaitEvent(h->stream, h->ev_out, 0));
  return B200_OK;
}

int ensure_streaming(b200_mimi* h, const char* what) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "%s: null handle", what);
  if (!h->finalized) B200_FAIL(B200_ERR_STATE, "%s: handle not finalized", what);
  if (h->batch <= 0) B200_FAIL(B200_ERR_STATE, "%s: not streaming (call streaming_begin first)", what);
  return B200_OK;
}

void register_tap(b200_mimi* h, const std::string& name, const float* p, int C, int T, bool transpose) { h->taps[name] = Tap{p, C, T, transpose}; }

// Tile shape of a layer whose sessions produce T GEMM rows per frame: the largest power of two that divides T (<= 128) steps
// of bb = 128 / tt sessions; strided convolutions also keep the TMA box (tt * stride rows) within 256.
int tile_shape(int T, int stride, int* tt, int* bb) {
  int t = 128;
  while (t > 1 && (T % t || t * stride > 256)) t >>= 1;
  if (T % t) return B200_ERR_SHAPE;
  *tt = t; *bb = 128 / t;
  return B200_OK;
}

// Plans one tensor-core layer for B sessions of T GEMM rows: A maps over `in_hi / in_lo` ([B][rows][Cin]), epilogue pointers set
// by the caller in L.p beforehand (y / res / act); this fills geometry, schedule and tensor maps.
int plan_layer(b200_mimi* h, mtc::TcLayer& L, const float* in_hi, const float* in_lo, int rows_per_session, int row0, int T) {
  const int B = h->batch;
  int tt = 0, bb = 0;
  if (tile_shape(T, L.stride, &tt, &bb) != B200_OK) B200_FAIL(B200_ERR_INVALID, "mimi: %d steps per frame do not tile", T);
  L.p.tt = tt; L.p.bb = bb; L.p.row0 = row0;
  L.p.ws = h->splitk_ws;
  B200_TRY(mtc::tc_make_map(&L.map_hi, in_hi, L.Cin, rows_per_session, (long long)rows_per_session * L.Cin, B, tt, bb, L.stride));
  B200_TRY(mtc::tc_make_map(&L.map_lo, in_lo, L.Cin, rows_per_session, (long long)rows_per_session * L.Cin, B, tt, bb, L.stride));
  return mtc::tc_plan(L, B, T, h->sms, h->splitk_bytes);
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int b200_mimi_create(const b200_mimi_config* cfg, b200_mimi** out) {
  if (!cfg || !out) B200_FAIL(B200_ERR_INVALID, "mimi_create: null argument");
  if (cfg->n_ratios < 1 || cfg->n_ratios > 8) B200_FAIL(B200_ERR_INVALID, "mimi_create: n_ratios out of range");
  if (cfg->channels != 1) B200_FAIL(B200_ERR_INVALID, "mimi_create: only mono audio is on the hot path");
  if (cfg->q_dimension % 8) B200_FAIL(B200_ERR_INVALID, "mimi_create: codebook dimension must be a multiple of 8");
  if (cfg->tr_d_model / cfg->tr_num_heads != 64) B200_FAIL(B200_ERR_INVALID, "mimi_create: head dim must be 64");
  if (cfg->tr_d_model != cfg->dimension) B200_FAIL(B200_ERR_INVALID, "mimi_create: projected transformer unsupported");
  if (cfg->n_filters % 32 || cfg->compress != 2 || cfg->n_filters / cfg->compress % 32)
    B200_FAIL(B200_ERR_INVALID, "mimi_create: SEANet channel counts must be multiples of 32 (tensor-core k-blocks)");
  b200_mimi* h = new b200_mimi();
  h->cfg = *cfg;
  cudaGetDevice(&h->device);
  h->num_codebooks = cfg->num_codebooks;
  h->hop = 1;
  for (int i = 0; i < cfg->n_ratios; ++i) h->hop *= cfg->ratios[i];
  h->frame_size = (int)(cfg->sample_rate / cfg->frame_rate);
  const double enc_rate = (double)cfg->sample_rate / h->hop;
  h->rs = (int)(enc_rate / cfg->frame_rate);
  if (h->rs < 1 || h->frame_size != h->hop * h->rs) {
    delete h;
    B200_FAIL(B200_ERR_INVALID, "mimi_create: frame size %d is not hop*stride", h->frame_size);
  }
  plan_seanet(h);
  *out = h;
  return B200_OK;
}

int b200_mimi_load_tensor(b200_mimi* h, const char* name, const void* data_dev, int dtype, int ndim,
                          const int64_t* shape) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "mimi_load_tensor: null handle");
  if (h->finalized) B200_FAIL(B200_ERR_STATE, "mimi_load_tensor: already finalized");
  DeviceGuard g(h->device);
  return h->store.put(name, data_dev, dtype, ndim, shape);
}

int b200_mimi_finalize(b200_mimi* h) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "mimi_finalize: null handle");
  if (h->finalized) return B200_OK;
  DeviceGuard g(h->device);
  B200_TRY(mtc::tc_init());
  B200_CUDA(cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, h->device));
  for (auto& l : h->enc) B200_TRY(pack_sea(h, l));
  for (auto& l : h->dec) B200_TRY(pack_sea(h, l));
  B200_TRY(pack_down(h, h->down));
  {
    const float* w = nullptr;
    B200_TRY(keep(h, "upsample.convtr.convtr.convtr.weight", {h->cfg.dimension, 1, 2 * h->rs}, &w));
    h->up_w = const_cast<float*>(w);
  }
  B200_TRY(pack_transformer(h, "encoder_transformer", h->enc_tr));
  B200_TRY(pack_transformer(h, "decoder_transformer", h->dec_tr));
  B200_TRY(pack_quantizer(h));
  B200_CUDA(cudaDeviceSynchronize());
  h->store.release_all();   // anything left is not used on the hot path
  h->finalized = true;
  return B200_OK;
}

int b200_mimi_destroy(b200_mimi* h) {
  if (!h) return B200_OK;
  DeviceGuard g(h->device);
  b200_mimi_streaming_end(h);
  h->store.release_all();
  h->weights.free_all();
  if (h->pin_pcm) cudaFreeHost(h->pin_pcm);
  if (h->pin_codes) cudaFreeHost(h->pin_codes);
  if (h->dev_pcm) cudaFree(h->dev_pcm);
  if (h->dev_codes) cudaFree(h->dev_codes);
  if (h->tap_scratch) cudaFree(h->tap_scratch);
  for (int w = 0; w < 2; ++w) {
    if (h->rvq_res[w]) cudaFree(h->rvq_res[w]);
    if (h->rvq_best[w]) cudaFree(h->rvq_best[w]);
    if (h->rvq_idx[w]) cudaFree(h->rvq_idx[w]);
  }
  delete h;
  return B200_OK;
}

int b200_mimi_set_num_codebooks(b200_mimi* h, int n) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "set_num_codebooks: null handle");
  if (n < h->cfg.q_n_semantic || n > h->cfg.q_n_q)   // vq.py:315-317 assertion
    B200_FAIL(B200_ERR_SHAPE, "set_num_codebooks(%d): must be in [%d, %d]", n, h->cfg.q_n_semantic, h->cfg.q_n_q);
  h->num_codebooks = n;
  if (h->enc_graph) { cudaGraphExecDestroy(h->enc_graph); h->enc_graph = nullptr; }
  return B200_OK;
}

int b200_mimi_streaming_begin(b200_mimi* h, int batch, void* stream) {
  if (!h || !h->finalized) B200_FAIL(B200_ERR_STATE, "streaming_begin: handle not finalized");
  if (h->batch > 0) B200_FAIL(B200_ERR_STATE, "streaming_begin: already streaming");   // streaming.py:112
  if (batch < 1) B200_FAIL(B200_ERR_INVALID, "streaming_begin: batch %d", batch);
  DeviceGuard guard_dev(h->device);
  const auto& c = h->cfg;
  const int B = batch, d = c.dimension;
  h->stream = static_cast<cudaStream_t>(stream);
  h->taps.clear();
  Arena& A = h->state;
  struct Cleanup { b200_mimi* h; bool armed = true; ~Cleanup() { if (armed) h->state.free_all(); } } cleanup{h};
  h->batch = B;                              // plan_layer reads it (reset below on failure paths by the caller's streaming_end)
  struct BatchReset { b200_mimi* h; bool armed = true; ~BatchReset() { if (armed) h->batch = 0; } } batch_reset{h};
  B200_TRY(A.alloc_t(&h->exec_mask, B, false));
  B200_CUDA(cudaMemset(h->exec_mask, 1, B));
  B200_TRY(A.alloc_t(&h->err, 1));
  h->splitk_bytes = (size_t)64 << 20;
  B200_TRY(A.alloc(reinterpret_cast<void**>(&h->splitk_ws), h->splitk_bytes, false));

  // flags for replicate-padded convs (only the down-sampling conv in Mimi)
  h->n_first = 1;
  B200_TRY(A.alloc_t(&h->first_flags, (size_t)h->n_first * B, false));
  B200_CUDA(cudaMemset(h->first_flags, 1, (size_t)h->n_first * B));

  const int T = h->rs;   // transformer tokens per frame on both sides
  B200_TRY(A.alloc_t(&h->tok_in_enc, (size_t)B * T * d));   // written by the last encoder conv
  B200_TRY(A.alloc_t(&h->tok_enc, (size_t)B * T * d));
  B200_TRY(A.alloc_t(&h->latent, (size_t)B * d));
  B200_TRY(A.alloc_t(&h->latent_q, (size_t)B * d));
  B200_TRY(A.alloc_t(&h->tok_in_dec, (size_t)B * T * d));
  B200_TRY(A.alloc_t(&h->tok_dec, (size_t)B * T * d));
  B200_TRY(A.alloc_t(&h->out_frame, (size_t)B * h->frame_size));

  std::vector<TmCommit> enc_tm, dec_tm;
  h->max_enc_tm = h->max_dec_tm = 0;
  // buffers of one SEANet: every layer owns its extended input; raw outputs where something reads them
  auto setup = [&](std::vector<SeaLayer>& layers, int t0, bool is_dec, std::vector<TmCommit>& commits, long long& max_rows) -> int {
    int t = t0;
    for (size_t i = 0; i < layers.size(); ++i) {
      SeaLayer& l = layers[i];
      l.t_in = t;
      if (l.kind == 0) {
        if (t % l.stride) B200_FAIL(B200_ERR_INVALID, "%s: %d samples not divisible by stride %d", l.key.c_str(), t, l.stride);
        l.t_out = t / l.stride;
      } else {
        l.t_out = t;                            // GEMM rows per session; each row holds `stride` output steps
      }
      const size_t ext_n = (size_t)B * (l.P + l.t_in) * l.cin;
      B200_TRY(A.alloc_t(&l.ext_hi, ext_n));
      if (!l.simt_first && !l.simt_last) B200_TRY(A.alloc_t(&l.ext_lo, ext_n));
      else l.ext_lo = nullptr;
      if (l.P > 0) {
        TmCommit cm;
        cm.hi = l.ext_hi; cm.lo = l.ext_lo; cm.P = l.P; cm.T = l.t_in; cm.rowlen = l.cin; cm.sb = (long long)(l.P + l.t_in) * l.cin;
        commits.push_back(cm);
        if ((long long)B * l.cin > max_rows) max_rows = (long long)B * l.cin;
      }
      const int steps_out = l.kind == 1 ? l.t_out * l.stride : l.t_out;
      const bool is_last = i + 1 == layers.size();
      l.y = nullptr;
      if (is_last && !is_dec) l.y = h->tok_in_enc;                       // transformer input
      else if (is_last && is_dec) l.y = h->out_frame;
      else if (!l.tap.empty()) B200_TRY(A.alloc_t(&l.y, (size_t)B * steps_out * l.cout));
      if (!l.tap.empty() && l.y) register_tap(h, l.tap, l.y, l.cout, steps_out, !(is_last && !is_dec));
      t = steps_out;
    }
    return B200_OK;
  };
  B200_TRY(setup(h->enc, h->frame_size, false, enc_tm, h->max_enc_tm));
  if (h->enc.back().t_out != T) B200_FAIL(B200_ERR_INVALID, "encoder yields %d tokens/frame, expected %d", h->enc.back().t_out, T);
  B200_TRY(setup(h->dec, T, true, dec_tm, h->max_dec_tm));
  if (h->dec.back().t_out != h->frame_size) B200_FAIL(B200_ERR_INVALID, "decoder yields %d samples/frame", h->dec.back().t_out);
  // epilogues + plans: layer i writes its raw output and the activation its consumer i + 1 reads
  auto wire = [&](std::vector<SeaLayer>& layers) -> int {
    for (size_t i = 0; i < layers.size(); ++i) {
      SeaLayer& l = layers[i];
      if (l.simt_first || l.simt_last) continue;
      mtc::TcParams& p = l.tc.p;
      memset(&p, 0, sizeof(p));
      p.epi = mtc::TC_EPI_CONV;
      p.y = l.y; p.y_sb = (long long)l.t_out * l.N; p.y_row = l.N;      // convtr: a GEMM row is `stride` consecutive output steps
      if (l.res_from >= 0) {
        const SeaLayer& r = layers[l.res_from];
        if (!r.y) B200_FAIL(B200_ERR_INVALID, "bad residual plan");
        p.res = r.y; p.r_sb = (long long)l.t_out * l.N; p.r_row = l.N;
      }
      if (i + 1 < layers.size()) {
        const SeaLayer& nx = layers[i + 1];
        p.act_mode = nx.simt_last ? mtc::TC_ACT_FULL : mtc::TC_ACT_SPLIT;
        p.act_elu = nx.elu_in ? 1 : 0;
        p.a_hi = nx.ext_hi + (long long)nx.P * nx.cin;
        p.a_lo = nx.ext_lo ? nx.ext_lo + (long long)nx.P * nx.cin : nullptr;
        p.a_sb = (long long)(nx.P + nx.t_in) * nx.cin;
        p.a_row = l.N;                         // = stride * nx.cin for a convtr: `stride` consumer rows per GEMM row
      }
      // conv: GEMM row t reads ext rows t * stride + tap * dil (ext row 0 = oldest carried sample)
      // convtr: GEMM row t reads ext rows t + slot, slot 0 = x[t-1] (ext row 0 = the carried previous step)
      B200_TRY(plan_layer(h, l.tc, l.ext_hi, l.ext_lo, l.P + l.t_in, 0, l.t_out));
    }
    return B200_OK;
  };
  B200_TRY(wire(h->enc));
  B200_TRY(wire(h->dec));
  register_tap(h, "enc.tr", h->tok_enc, d, T, false);
  register_tap(h, "enc.latent", h->latent, d, 1, false);
  register_tap(h, "dec.latent", h->latent_q, d, 1, false);
  register_tap(h, "dec.up", h->tok_in_dec, d, T, false);
  register_tap(h, "dec.tr", h->tok_dec, d, T, false);

  // down-sampling conv (replicate pad): reads tok_enc token-major
  std::vector<ConvCommit> enc_c;
  {
    DownLayer& l = h->down;
    l.t_in = T; l.t_out = T / l.stride; l.P = l.k - l.stride;
    B200_TRY(A.alloc_t(&l.state, (size_t)B * l.cin * l.P));
    l.first = h->first_flags;
    ConvCommit cc;
    cc.x = h->tok_enc; cc.xb = (long long)T * d; cc.xc = 1; cc.xt = d; cc.Tin = T; cc.st = l.state; cc.P = l.P;
    cc.Cin = l.cin; cc.elu_in = 0; cc.first = l.first;
    enc_c.push_back(cc);
    h->max_enc_rows = B * l.cin;
  }
  // up-sampling depth-wise convtr
  std::vector<ConvTrCommit> dec_t;
  B200_TRY(A.alloc_t(&h->up_partial, (size_t)B * d * h->rs));
  B200_TRY(A.alloc_t(&h->up_scratch, (size_t)B * d * h->rs));
  {
    ConvTrCommit tc;
    tc.partial = h->up_partial; tc.scratch = h->up_scratch; tc.per_row = d * h->rs;
    dec_t.push_back(tc);
    h->max_tr_rows = (long long)B * tc.per_row;
  }
  h->n_enc_commits = (int)enc_c.size();
  h->n_dec_tr_commits = (int)dec_t.size();
  h->n_enc_tm = (int)enc_tm.size();
  h->n_dec_tm = (int)dec_tm.size();
  B200_TRY(A.alloc_t(&h->enc_commits, enc_c.size(), false));
  B200_TRY(A.alloc_t(&h->dec_tr_commits, dec_t.size(), false));
  B200_TRY(A.alloc_t(&h->enc_tm, enc_tm.size() ? enc_tm.size() : 1, false));
  B200_TRY(A.alloc_t(&h->dec_tm, dec_tm.size() ? dec_tm.size() : 1, false));
  B200_CUDA(cudaMemcpy(h->enc_commits, enc_c.data(), enc_c.size() * sizeof(ConvCommit), cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemcpy(h->dec_tr_commits, dec_t.data(), dec_t.size() * sizeof(ConvTrCommit), cudaMemcpyHostToDevice));
  if (!enc_tm.empty()) B200_CUDA(cudaMemcpy(h->enc_tm, enc_tm.data(), enc_tm.size() * sizeof(TmCommit), cudaMemcpyHostToDevice));
  if (!dec_tm.empty()) B200_CUDA(cudaMemcpy(h->dec_tm, dec_tm.data(), dec_tm.size() * sizeof(TmCommit), cudaMemcpyHostToDevice));

  // transformers
  const int H = c.tr_num_heads, D = d / H, ff = c.tr_dim_feedforward;
  const size_t ntok = (size_t)B * T;
  B200_TRY(A.alloc_t(&h->tr_xn_hi, ntok * d));
  B200_TRY(A.alloc_t(&h->tr_xn_lo, ntok * d));
  B200_TRY(A.alloc_t(&h->tr_qkv, ntok * 3 * d));
  B200_TRY(A.alloc_t(&h->tr_ao_hi, ntok * d));
  B200_TRY(A.alloc_t(&h->tr_ao_lo, ntok * d));
  B200_TRY(A.alloc_t(&h->tr_h_hi, ntok * ff));
  B200_TRY(A.alloc_t(&h->tr_h_lo, ntok * ff));
  for (int which = 0; which < 2; ++which) {
    Transformer* tr = which == 0 ? &h->enc_tr : &h->dec_tr;
    const float* x_in = which == 0 ? h->tok_in_enc : h->tok_in_dec;
    float* x = which == 0 ? h->tok_enc : h->tok_dec;
    B200_TRY(A.alloc_t(&tr->offset, B));
    for (size_t li = 0; li < tr->layers.size(); ++li) {
      TrLayer& L = tr->layers[li];
      B200_TRY(A.alloc_t(&L.kc, (size_t)B * H * c.tr_context * D));
      B200_TRY(A.alloc_t(&L.vc, (size_t)B * H * c.tr_context * D));
      const float* cur = li == 0 ? x_in : x;
      auto lin = [&](mtc::TcLayer& G, const float* in_hi, const float* in_lo, int K, float* y, int N) -> int {
        memset(&G.p, 0, sizeof(G.p));
        G.p.epi = mtc::TC_EPI_CONV;
        G.p.y = y; G.p.y_sb = (long long)T * N; G.p.y_row = N;
        (void)K;
        return B200_OK;
      };
      // in_proj: qkv = W_in xn
      B200_TRY(lin(L.in_proj, h->tr_xn_hi, h->tr_xn_lo, d, h->tr_qkv, 3 * d));
      B200_TRY(plan_layer(h, L.in_proj, h->tr_xn_hi, h->tr_xn_lo, T, 0, T));
      // out_proj: x = cur + layer_scale_1 * (W_out ao)   (transformer.py:769)
      B200_TRY(lin(L.out_proj, h->tr_ao_hi, h->tr_ao_lo, d, x, d));
      L.out_proj.p.epi = mtc::TC_EPI_RES_SCALE; L.out_proj.p.res = cur; L.out_proj.p.r_sb = (long long)T * d; L.out_proj.p.r_row = d;
      L.out_proj.p.scale = L.ls1;
      B200_TRY(plan_layer(h, L.out_proj, h->tr_ao_hi, h->tr_ao_lo, T, 0, T));
      // linear1 + GELU -> h as a hi / lo pair (never stored raw)
      B200_TRY(lin(L.l1, h->tr_xn_hi, h->tr_xn_lo, d, nullptr, ff));
      L.l1.p.epi = mtc::TC_EPI_GELU; L.l1.p.act_mode = mtc::TC_ACT_SPLIT; L.l1.p.act_elu = 0;
      L.l1.p.a_hi = h->tr_h_hi; L.l1.p.a_lo = h->tr_h_lo; L.l1.p.a_sb = (long long)T * ff; L.l1.p.a_row = ff;
      B200_TRY(plan_layer(h, L.l1, h->tr_xn_hi, h->tr_xn_lo, T, 0, T));
      // linear2: x = x + layer_scale_2 * (W_2 h)   (transformer.py:777); the decoder's last layer also feeds dec[0]
      B200_TRY(lin(L.l2, h->tr_h_hi, h->tr_h_lo, ff, x, d));
      L.l2.p.epi = mtc::TC_EPI_RES_SCALE; L.l2.p.res = x; L.l2.p.r_sb = (long long)T * d; L.l2.p.r_row = d; L.l2.p.scale = L.ls2;
      if (which == 1 && li + 1 == tr->layers.size()) {
        const SeaLayer& nx = h->dec[0];
        L.l2.p.act_mode = mtc::TC_ACT_SPLIT; L.l2.p.act_elu = nx.elu_in ? 1 : 0;
        L.l2.p.a_hi = nx.ext_hi + (long long)nx.P * nx.cin; L.l2.p.a_lo = nx.ext_lo + (long long)nx.P * nx.cin;
        L.l2.p.a_sb = (long long)(nx.P + nx.t_in) * nx.cin; L.l2.p.a_row = d;
      }
      B200_TRY(plan_layer(h, L.l2, h->tr_h_hi, h->tr_h_lo, T, 0, T));
    }
  }
  B200_TRY(A.alloc_t(&h->scratch_codes, (size_t)B * c.q_n_q));
  B200_TRY(A.alloc_t(&h->enc_codes, (size_t)B * c.q_n_q));
  B200_TRY(A.alloc_t(&h->dec_codes, (size_t)B * c.q_n_q));
  B200_CUDA(cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking));
  B200_CUDA(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
  h->body = h->stream;
  // what a streaming-state snapshot holds (streaming.py:158-181): masks, carried conv rows (the transposed convs' carry is
  // their previous input step), both transformers' KV rings and offsets
  A.mark_state(h->exec_mask, B, "exec_mask", B200_U8, {B});
  A.mark_state(h->first_flags, (size_t)h->n_first * B, "downsample.first", B200_U8, {B});
  for (int which = 0; which < 2; ++which)
    for (auto& l : which == 0 ? h->enc : h->dec) {
      const std::string nm = (which == 0 ? "encoder." : "decoder.") + l.key.substr(l.key.find("model."));
      const int64_t E = l.P + l.t_in;
      // the whole extended buffer is registered (the frame part is scratch); the shim slices the first P rows out
      A.mark_state(l.ext_hi, (size_t)B * E * l.cin * 4, nm + ".ext_hi", B200_F32, {B, E, l.cin});
      if (l.ext_lo) A.mark_state(l.ext_lo, (size_t)B * E * l.cin * 4, nm + ".ext_lo", B200_F32, {B, E, l.cin});
    }
  A.mark_state(h->down.state, (size_t)B * h->down.cin * h->down.P * 4, "downsample.previous", B200_F32, {B, h->down.cin, h->down.P});
  A.mark_state(h->up_partial, (size_t)B * d * h->rs * 4, "upsample.partial", B200_F32, {B, d, h->rs});
  for (int which = 0; which < 2; ++which) {
    Transformer* tr = which == 0 ? &h->enc_tr : &h->dec_tr;
    const std::string tp = which == 0 ? "encoder_transformer" : "decoder_transformer";
    A.mark_state(tr->offset, (size_t)B * 8, tp + ".offset", B200_I64, {B});
    for (size_t li = 0; li < tr->layers.size(); ++li) {
      auto& L = tr->layers[li];
      const std::string p = tp + ".layers." + std::to_string(li);
      A.mark_state(L.kc, (size_t)B * H * c.tr_context * D * 4, p + ".k", B200_F32, {B, H, c.tr_context, D});
      A.mark_state(L.vc, (size_t)B * H * c.tr_context * D * 4, p + ".v", B200_F32, {B, H, c.tr_context, D});
    }
  }
  B200_TRY(ensure_rvq_workspace(h, B));
  B200_CUDA(cudaDeviceSynchronize());
  cleanup.armed = false;
  batch_reset.armed = false;
  return B200_OK;
}

int b200_mimi_streaming_end(b200_mimi* h) {
  if (!h) return B200_OK;
  DeviceGuard g(h->device);
  if (h->batch > 0) {
    cudaStreamSynchronize(h->stream);
    if (h->gstream) cudaStreamSynchronize(h->gstream);
  }
  drop_graphs(h);
  if (h->gstream) cudaStreamDestroy(h->gstream);
  if (h->ev_in) cudaEventDestroy(h->ev_in);
  if (h->ev_out) cudaEventDestroy(h->ev_out);
  h->gstream = nullptr; h->ev_in = h->ev_out = nullptr;
  h->state.free_all();
  h->err = nullptr;
  h->batch = 0;
  h->taps.clear();
  return B200_OK;
}

int b200_mimi_reset(b200_mimi* h, const uint8_t* reset_mask_dev) {
  B200_TRY(ensure_streaming(h, "mimi_reset"));
  DeviceGuard g(h->device);
  const int B = h->batch;
  auto zero = [&](float* buf, long long per_row) {
    if (!buf || per_row <= 0) return;
    B200_LAUNCH(zero_rows_kernel, (unsigned)ceil_div64((long long)B * per_row, 256), 256, 0, h->stream, buf, per_row,
                reset_mask_dev, B);
  };
  for (auto* layers : {&h->enc, &h->dec})
    for (auto& l : *layers)
      if (l.P > 0)
        B200_LAUNCH(tm_zero_kernel, (unsigned)ceil_div64((long long)B * l.P * l.cin, 256), 256, 0, h->stream, l.ext_hi, l.ext_lo, l.P,
                    l.cin, (long long)(l.P + l.t_in) * l.cin, reset_mask_dev, B);
  zero(h->down.state, (long long)h->down.cin * h->down.P);
  zero(h->up_partial, (long long)h->cfg.dimension * h->rs);
  B200_LAUNCH(reset_flags_kernel, ceil_div(B, 128), 128, 0, h->stream, h->first_flags, h->n_first, h->enc_tr.offset,
              h->dec_tr.offset, h->exec_mask, reset_mask_dev, B);
  return check_launch("mimi_reset");
}

int b200_mimi_set_exec_mask(b200_mimi* h, const uint8_t* exec_mask_dev) {
  B200_TRY(ensure_streaming(h, "mimi_set_exec_mask"));
  DeviceGuard g(h->device);
  if (!exec_mask_dev) B200_FAIL(B200_ERR_INVALID, "mimi_set_exec_mask: null mask");
  B200_CUDA(cudaMemcpyAsync(h->exec_mask, exec_mask_dev, h->batch, cudaMemcpyDeviceToDevice, h->stream));
  return B200_OK;
}

int b200_mimi_encode_to_latent(b200_mimi* h, const float* pcm_dev, int n_frames, float* latent_dev) {
  B200_TRY(ensure_streaming(h, "mimi_encode_to_latent"));
  if (!pcm_dev || !latent_dev || n_frames < 1) B200_FAIL(B200_ERR_SHAPE, "mimi_encode_to_latent: bad arguments");
  const int d = h->cfg.dimension;
  for (int f = 0; f < n_frames; ++f) {
    B200_TRY(load_frame(h, pcm_dev, n_frames, f));
    h->body = h->stream;
    B200_TRY(encode_body(h));
    B200_CUDA(cudaMemcpy2DAsync(latent_dev + f, (size_t)n_frames * 4, h->latent, 4, 4, (size_t)h->batch * d,
                                cudaMemcpyDeviceToDevice, h->stream));
  }
  return B200_OK;
}

int b200_mimi_quantize(b200_mimi* h, const float* latent_dev, int n_frames, int64_t* codes_dev) {
  B200_TRY(ensure_streaming(h, "mimi_quantize"));
  if (!latent_dev || !codes_dev || n_frames < 1) B200_FAIL(B200_ERR_SHAPE, "mimi_quantize: bad arguments");
  const int d = h->cfg.dimension, K = h->num_codebooks;
  h->body = h->stream;
  return quantize_cols(h, latent_dev, (long long)d * n_frames, n_frames, 1, n_frames,
                       reinterpret_cast<long long*>(codes_dev), (long long)K * n_frames, n_frames, 1);
}

int b200_mimi_encode(b200_mimi* h, const float* pcm_dev, int n_frames, int64_t* codes_dev) {
  B200_TRY(ensure_streaming(h, "mimi_encode"));
  if (!pcm_dev || !codes_dev || n_frames < 1) B200_FAIL(B200_ERR_SHAPE, "mimi_encode: bad arguments");
  const int d = h->cfg.dimension, K = h->num_codebooks;
  for (int f = 0; f < n_frames; ++f) {
    B200_TRY(load_frame(h, pcm_dev, n_frames, f));
    B200_TRY(run_captured(h, &h->enc_graph, &h->enc_graph_kernels, [&]() -> int {
      B200_TRY(encode_body(h));
      return quantize_cols(h, h->latent, d, 1, 1, 1, h->enc_codes, K, 1, 1);
    }));
    // enc_codes [B][K] -> codes[b][k][f]
    B200_CUDA(cudaMemcpy2DAsync(reinterpret_cast<long long*>(codes_dev) + f, (size_t)n_frames * 8, h->enc_codes, 8, 8,
                                (size_t)h->batch * K, cudaMemcpyDeviceToDevice, h->stream));
  }
  return B200_OK;
}

int b200_mimi_decode_latent(b200_mimi* h, const int64_t* codes_dev, int n_codebooks, int n_frames, float* latent_dev) {
  B200_TRY(ensure_streaming(h, "mimi_decode_latent"));
  if (!codes_dev || !latent_dev || n_frames < 1) B200_FAIL(B200_ERR_SHAPE, "mimi_decode_latent: bad arguments");
  const int d = h->cfg.dimension;
  h->body = h->stream;
  return dequantize_cols(h, reinterpret_cast<const long long*>(codes_dev), (long long)n_codebooks * n_frames, n_frames, 1,
                         n_codebooks, n_frames, latent_dev, (long long)d * n_frames, n_frames, 1);
}

int b200_mimi_decode(b200_mimi* h, const int64_t* codes_dev, int n_codebooks, int n_frames, float* pcm_dev) {
  B200_TRY(ensure_streaming(h, "mimi_decode"));
  if (!codes_dev || !pcm_dev || n_frames < 1) B200_FAIL(B200_ERR_SHAPE, "mimi_decode: bad arguments");
  if (n_codebooks < 1 || n_codebooks > h->cfg.q_n_q) B200_FAIL(B200_ERR_SHAPE, "decode: %d codebooks", n_codebooks);
  const int d = h->cfg.dimension, fs = h->frame_size;
  if (h->dec_graph && h->dec_graph_ncb != n_codebooks) {
    cudaGraphExecDestroy(h->dec_graph);
    h->dec_graph = nullptr;
  }
  h->dec_graph_ncb = n_codebooks;
  for (int f = 0; f < n_frames; ++f) {
    // codes[b][k][f] -> dec_codes [B][n_codebooks]
    B200_CUDA(cudaMemcpy2DAsync(h->dec_codes, 8, reinterpret_cast<const long long*>(codes_dev) + f, (size_t)n_frames * 8, 8,
                                (size_t)h->batch * n_codebooks, cudaMemcpyDeviceToDevice, h->stream));
    B200_TRY(run_captured(h, &h->dec_graph, &h->dec_graph_kernels, [&]() -> int {
      B200_TRY(dequantize_cols(h, h->dec_codes, n_codebooks, 1, 1, n_codebooks, 1, h->latent_q, d, 1, 1));
      return decode_latent_body(h);
    }));
    B200_CUDA(cudaMemcpy2DAsync(pcm_dev + (size_t)f * fs, (size_t)fs * n_frames * 4, h->out_frame, (size_t)fs * 4,
                                (size_t)fs * 4, h->batch, cudaMemcpyDeviceToDevice, h->stream));
  }
  return B200_OK;
}

/* get_streaming_state / set_streaming_state (streaming.py:158-181) as one opaque device blob */
int64_t b200_mimi_state_bytes(b200_mimi* h) { return (h && h->batch > 0) ? (int64_t)h->state.state_bytes() : 0; }

int b200_mimi_get_state(b200_mimi* h, void* dst_dev, int64_t capacity) {
  B200_TRY(ensure_streaming(h, "mimi_get_state"));
  if (!dst_dev || capacity < (int64_t)h->state.state_bytes()) B200_FAIL(B200_ERR_SHAPE, "mimi_get_state: destination too small");
  return h->state.save(dst_dev, h->stream);
}

int b200_mimi_set_state(b200_mimi* h, const void* src_dev, int64_t nbytes) {
  B200_TRY(ensure_streaming(h, "mimi_set_state"));
  if (!src_dev || nbytes != (int64_t)h->state.state_bytes())
    B200_FAIL(B200_ERR_SHAPE, "mimi_set_state: snapshot of %lld bytes does not fit this session layout (%lld)",
              (long long)nbytes, (long long)h->state.state_bytes());
  return h->state.load(src_dev, h->stream);
}

int b200_mimi_set_graph(b200_mimi* h, int enable) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "mimi_set_graph: null handle");
  h->graph_enabled = enable;
  if (!enable) drop_graphs(h);
  return B200_OK;
}

static int ensure_host_staging(b200_mimi* h, size_t pcm_n, size_t codes_n) {
  if (pcm_n > h->pin_pcm_n) {
    if (h->pin_pcm) cudaFreeHost(h->pin_pcm);
    if (h->dev_pcm) cudaFree(h->dev_pcm);
    B200_CUDA(cudaMallocHost(&h->pin_pcm, pcm_n * 4));
    B200_CUDA(cudaMalloc(&h->dev_pcm, pcm_n * 4));
    h->pin_pcm_n = h->dev_pcm_n = pcm_n;
  }
  if (codes_n > h->pin_codes_n) {
    if (h->pin_codes) cudaFreeHost(h->pin_codes);
    if (h->dev_codes) cudaFree(h->dev_codes);
    B200_CUDA(cudaMallocHost(&h->pin_codes, codes_n * 8));
    B200_CUDA(cudaMalloc(&h->dev_codes, codes_n * 8));
    h->pin_codes_n = h->dev_codes_n = codes_n;
  }
  return B200_OK;
}

int b200_mimi_encode_host(b200_mimi* h, const float* pcm_host, int n_frames, int64_t* codes_host) {
  B200_TRY(ensure_streaming(h, "mimi_encode_host"));
  if (!pcm_host || !codes_host || n_frames < 1) B200_FAIL(B200_ERR_SHAPE, "mimi_encode_host: bad arguments");
  const size_t pn = (size_t)h->batch * h->frame_size * n_frames, cn = (size_t)h->batch * h->num_codebooks * n_frames;
  B200_TRY(ensure_host_staging(h, pn, cn));
  memcpy(h->pin_pcm, pcm_host, pn * 4);
  B200_CUDA(cudaMemcpyAsync(h->dev_pcm, h->pin_pcm, pn * 4, cudaMemcpyHostToDevice, h->stream));
  B200_TRY(b200_mimi_encode(h, h->dev_pcm, n_frames, reinterpret_cast<int64_t*>(h->dev_codes)));
  B200_CUDA(cudaMemcpyAsync(h->pin_codes, h->dev_codes, cn * 8, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  memcpy(codes_host, h->pin_codes, cn * 8);
  return B200_OK;
}

int b200_mimi_decode_host(b200_mimi* h, const int64_t* codes_host, int n_codebooks, int n_frames, float* pcm_host) {
  B200_TRY(ensure_streaming(h, "mimi_decode_host"));
  if (!pcm_host || !codes_host || n_frames < 1) B200_FAIL(B200_ERR_SHAPE, "mimi_decode_host: bad arguments");
  const size_t pn = (size_t)h->batch * h->frame_size * n_frames, cn = (size_t)h->batch * n_codebooks * n_frames;
  B200_TRY(ensure_host_staging(h, pn, cn));
  memcpy(h->pin_codes, codes_host, cn * 8);
  B200_CUDA(cudaMemcpyAsync(h->dev_codes, h->pin_codes, cn * 8, cudaMemcpyHostToDevice, h->stream));
  B200_TRY(b200_mimi_decode(h, reinterpret_cast<const int64_t*>(h->dev_codes), n_codebooks, n_frames, h->dev_pcm));
  B200_CUDA(cudaMemcpyAsync(h->pin_pcm, h->dev_pcm, pn * 4, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  memcpy(pcm_host, h->pin_pcm, pn * 4);
  return B200_OK;
}

/* the same state entry by entry (names in include/moshi_b200.h) */
int b200_mimi_state_count(b200_mimi* h) { return (h && h->batch > 0) ? (int)h->state.snap.size() : 0; }
int b200_mimi_state_entry(b200_mimi* h, int index, const char** name, int* dtype, int* ndim, int64_t* shape8, int64_t* nbytes) {
  B200_TRY(ensure_streaming(h, "mimi_state_entry"));
  return state_entry_info(h->state, index, name, dtype, ndim, shape8, nbytes);
}
int b200_mimi_state_read(b200_mimi* h, const char* name, void* dst_dev, int64_t nbytes) {
  B200_TRY(ensure_streaming(h, "mimi_state_read"));
  DeviceGuard g(h->device);
  if (!dst_dev) B200_FAIL(B200_ERR_INVALID, "mimi_state_read: null destination");
  return state_entry_copy(h->state, name, dst_dev, nullptr, nbytes, h->stream);
}
int b200_mimi_state_write(b200_mimi* h, const char* name, const void* src_dev, int64_t nbytes) {
  B200_TRY(ensure_streaming(h, "mimi_state_write"));
  DeviceGuard g(h->device);
  if (!src_dev) B200_FAIL(B200_ERR_INVALID, "mimi_state_write: null source");
  return state_entry_copy(h->state, name, nullptr, src_dev, nbytes, h->stream);
}

int b200_mimi_error_flags(b200_mimi* h, int* flags_out) {
  B200_TRY(ensure_streaming(h, "mimi_error_flags"));
  DeviceGuard g(h->device);
  int v = 0;
  B200_CUDA(cudaMemcpyAsync(&v, h->err, 4, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  if (v) B200_CUDA(cudaMemsetAsync(h->err, 0, 4, h->stream));
  if (flags_out) *flags_out = v;
  return B200_OK;
}

int b200_mimi_read_buffer(b200_mimi* h, const char* name, float* dst_dev, int64_t capacity, int64_t* numel) {
  B200_TRY(ensure_streaming(h, "mimi_read_buffer"));
  DeviceGuard g(h->device);
  auto it = h->taps.find(name ? name : "");
  if (it == h->taps.end()) B200_FAIL(B200_ERR_INVALID, "mimi_read_buffer: unknown buffer '%s'", name ? name : "(null)");
  const Tap& t = it->second;
  const int64_t n = (int64_t)h->batch * t.C * t.T;
  if (numel) *numel = n;
  if (!dst_dev) return B200_OK;
  if (capacity < n) B200_FAIL(B200_ERR_SHAPE, "mimi_read_buffer: destination too small");
  if (t.transpose && t.T > 1) {        // stored token-major [B][T][C], reported in the reference's [B][C][T]
    B200_LAUNCH(tm_to_bct_kernel, (unsigned)ceil_div64(n, 256), 256, 0, h->stream, t.p, dst_dev, h->batch, t.C, t.T);
    return check_launch("mimi_read_buffer");
  }
  B200_CUDA(cudaMemcpyAsync(dst_dev, t.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, h->stream));
  return B200_OK;
}

int64_t b200_mimi_algorithmic_bytes(b200_mimi* h) {
  if (!h || h->batch <= 0) return 0;
  const auto& c = h->cfg;
  // weights once per step (encode + decode; the tensor-core layers hold a hi and a lo copy: 8 bytes per parameter) + active
  // codebooks (row-major for decode, transposed for encode)
  int64_t bytes = h->weight_bytes;
  bytes += (int64_t)2 * h->num_codebooks * c.q_bins * c.q_dimension * 4 + (int64_t)4 * c.q_dimension * c.dimension * 4;
  // per session: both transformer KV rings read once, carried conv rows read + written (hi and lo), PCM in/out
  int64_t per_row = (int64_t)2 * c.tr_num_layers * 2 * c.tr_context * c.tr_d_model * 4;
  int64_t st = 0;
  for (auto* layers : {&h->enc, &h->dec})
    for (auto& l : *layers) st += (int64_t)l.cin * l.P * (l.ext_lo ? 2 : 1);
  st += (int64_t)h->down.cin * h->down.P + (int64_t)c.dimension * h->rs;
  per_row += 2 * st * 4 + (int64_t)2 * h->frame_size * 4 + h->num_codebooks * 8;
  return bytes + per_row * h->batch;
}

}  // extern "C"
