// Mimi handle: weight ingestion (reference state-dict names), streaming state, encode / decode.
// Reference orchestration: moshi/moshi/models/compression.py:338-433.
#include "mimi_kernels.cuh"
#include "mimi_gemm.cuh"

using namespace b200;
using namespace b200::mimi;

namespace {

struct ConvLayer {
  int kind = 0;              // 0 = StreamingConv1d, 1 = StreamingConvTranspose1d
  std::string key;           // state-dict prefix of the nn.Conv1d / nn.ConvTranspose1d
  int cin = 0, cout = 0, k = 0, stride = 1, dil = 1;
  bool elu_in = false, replicate = false, has_bias = true;
  int res_from = -1;         // buffer index added in the epilogue (residual block second conv)
  int in_buf = -1, out_buf = -1;
  int t_in = 0, t_out = 0;   // samples per frame at this layer
  float* w = nullptr;        // packed
  float* bias = nullptr;
  // streaming state
  int P = 0;                 // conv: carried samples; convtr: K - S
  float* state = nullptr;    // conv: previous [B][Cin][P]; convtr: partial [B][Cout][S]
  float* scratch = nullptr;  // convtr candidate partial
  uint8_t* first = nullptr;  // replicate flag [B]
  std::string tap;           // debug name of the output buffer
  // mimi_gemm.cuh path: k-major weights and the extended input buffer (carried state | frame), see there
  bool fast = false;
  float* wk = nullptr;
  float* ext = nullptr;      // conv: ext [B][Cin][E]; convtr: zero-padded input [B][Cin][E]
  int E = 0, D0 = 0;         // row length; index of the first sample of the current frame
};

struct Buf {
  float* p = nullptr;
  int c = 0, t = 0;
  bool token_major = false;  // [B][T][C] instead of [B][C][T]
  long long sb(int) const { return (long long)c * t; }
  long long sc() const { return token_major ? 1 : t; }
  long long st() const { return token_major ? c : 1; }
};

struct TrLayer {
  const float *in_w, *out_w, *n1w, *n1b, *n2w, *n2b, *l1, *l2, *ls1, *ls2;   // linear weights k-major [K][M]
  float *kc = nullptr, *vc = nullptr;
};

struct Transformer {
  std::vector<TrLayer> layers;
  long long* offset = nullptr;   // [B] tokens seen (== RingKVCache.end_offset == _MHAState.offset)
};

}  // namespace

struct b200_mimi {
  b200_mimi_config cfg;
  TensorStore store;
  Arena weights, state;
  bool finalized = false;
  int batch = 0;
  int num_codebooks = 8;
  cudaStream_t stream = nullptr;
  int frame_size = 0, hop = 0, rs = 0;       // rs = resample stride (2)

  std::vector<ConvLayer> enc, dec;
  ConvLayer down;
  std::vector<Buf> enc_bufs, dec_bufs;
  // up-sampling (depth-wise convtr)
  float *up_w = nullptr, *up_partial = nullptr, *up_scratch = nullptr;
  Transformer enc_tr, dec_tr;
  // quantizer
  float *wT[2] = {nullptr, nullptr}, *woT[2] = {nullptr, nullptr};
  float *cb[2] = {nullptr, nullptr}, *cbT[2] = {nullptr, nullptr}, *cnorm[2] = {nullptr, nullptr};
  int stored_levels[2] = {0, 0};
  // per-batch buffers
  float* in_frame = nullptr;                   // [B][1920]
  float *tok_in_enc = nullptr, *tok_enc = nullptr, *latent = nullptr;     // [B][T][C], [B][T][C], [B][C]
  float *latent_q = nullptr, *tok_in_dec = nullptr, *tok_dec = nullptr;
  float *tr_xn = nullptr, *tr_qkv = nullptr, *tr_q = nullptr, *tr_ao = nullptr, *tr_h = nullptr;
  uint8_t* exec_mask = nullptr;
  uint8_t* first_flags = nullptr; int n_first = 0;
  ConvCommit *enc_commits = nullptr, *dec_commits = nullptr;
  int n_enc_commits = 0, n_dec_commits = 0, max_enc_rows = 0, max_dec_rows = 0;
  ConvTrCommit* dec_tr_commits = nullptr; int n_dec_tr_commits = 0; long long max_tr_rows = 0;
  ExtCommit *enc_ext_commits = nullptr, *dec_ext_commits = nullptr;
  int n_enc_ext = 0, n_dec_ext = 0, max_enc_ext_rows = 0, max_dec_ext_rows = 0;
  long long* scratch_codes = nullptr;          // [B][K][1] for the host variants
  float* rvq_res[2] = {nullptr, nullptr};      // RVQ workspace: residuals, per-chunk partial argmin
  float* rvq_best[2] = {nullptr, nullptr};
  int* rvq_idx[2] = {nullptr, nullptr};
  int rvq_cap = 0;
  float* splitk_ws = nullptr; size_t splitk_bytes = 0;   // split-K partial sums of the deep / skinny GEMMs
  // one-frame encode / decode as CUDA graphs over static buffers (in_frame -> enc_codes, dec_codes -> out_frame)
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
  std::map<std::string, std::pair<const float*, int64_t>> taps;
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
    ConvLayer l;
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
      b.res_from = -2;  // resolved below: input of the block
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
    ConvLayer t;
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
      b.res_from = -2;
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
  h->down = conv("downsample.conv.conv.conv", c.dimension, c.dimension, 2 * h->rs, h->rs, 1, false);
  h->down.replicate = true;
  h->down.has_bias = false;
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

int pack_conv(b200_mimi* h, ConvLayer& l) {
  const float* w = nullptr;
  // ext-buffer path (mimi_gemm.cuh): GEMM-shaped layers, plus the Cout = 1 tail conv which reduces straight from ext
  l.fast = !l.replicate && (l.cin % GK) == 0 &&
           (l.kind == 0 ? (l.cout % 4 == 0 || (l.cout == 1 && l.stride == 1)) : (l.cout * l.stride) % 4 == 0);
  if (getenv("B200_MIMI_LEGACY")) l.fast = false;          // debug switch: first-generation kernels only
  if (l.kind == 0) {
    B200_TRY(get_f32(h, l.key + ".weight", {l.cout, l.cin, l.k}, &w));
    const long long n = (long long)l.cout * l.cin * l.k;
    if (l.fast) {
      B200_TRY(h->weights.alloc_t(&l.wk, n, false));
      B200_LAUNCH(pack_conv_k_kernel, (unsigned)ceil_div64(n, 256), 256, 0, 0, w, l.wk, l.cout, l.cin, l.k);
    } else {
      B200_TRY(h->weights.alloc_t(&l.w, n, false));
      B200_LAUNCH(pack_conv_w_kernel, (unsigned)ceil_div64(n, 256), 256, 0, 0, w, l.w, l.cout, l.cin, l.k);
    }
    h->weight_bytes += n * 4;
  } else {
    if (l.k != 2 * l.stride) B200_FAIL(B200_ERR_INVALID, "convtr %s: kernel must be 2*stride", l.key.c_str());
    B200_TRY(get_f32(h, l.key + ".weight", {l.cin, l.cout, l.k}, &w));
    const long long n = (long long)l.cin * l.cout * l.k;
    if (l.fast) {
      B200_TRY(h->weights.alloc_t(&l.wk, n, false));
      B200_LAUNCH(pack_convtr_k_kernel, (unsigned)ceil_div64(n, 256), 256, 0, 0, w, l.wk, l.cin, l.cout, l.stride);
    } else {
      B200_TRY(h->weights.alloc_t(&l.w, n, false));
      B200_LAUNCH(pack_convtr_w_kernel, (unsigned)ceil_div64(n, 256), 256, 0, 0, w, l.w, l.cin, l.cout, l.stride);
    }
    h->weight_bytes += n * 4;
  }
  if (l.has_bias) {
    const float* b = nullptr;
    B200_TRY(get_f32(h, l.key + ".bias", {l.cout}, &b));
    B200_TRY(h->weights.alloc_t(&l.bias, l.cout, false));
    B200_CUDA(cudaMemcpy(l.bias, b, l.cout * 4, cudaMemcpyDeviceToDevice));
    h->weight_bytes += l.cout * 4;
  }
  B200_CUDA(cudaDeviceSynchronize());
  h->store.release(l.key + ".weight");
  h->store.release(l.key + ".bias");
  return check_launch("pack_conv");
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

// nn.Linear weight [M][K] -> k-major [K][M] (A operand of mimi_gemm_kernel)
int keep_kmajor(b200_mimi* h, const std::string& name, int M, int K, const float** out) {
  const float* src = nullptr;
  B200_TRY(get_f32(h, name, {M, K}, &src));
  float* dst = nullptr;
  B200_TRY(h->weights.alloc_t(&dst, (size_t)M * K, false));
  B200_LAUNCH(transpose_kernel, (unsigned)ceil_div64((long long)M * K, 256), 256, 0, 0, src, dst, M, K);
  B200_CUDA(cudaDeviceSynchronize());
  h->store.release(name);
  h->weight_bytes += (int64_t)M * K * 4;
  *out = dst;
  return check_launch("keep_kmajor");
}

int pack_transformer(b200_mimi* h, const std::string& prefix, Transformer& tr) {
  const auto& c = h->cfg;
  const int d = c.tr_d_model, ff = c.tr_dim_feedforward;
  tr.layers.resize(c.tr_num_layers);
  for (int li = 0; li < c.tr_num_layers; ++li) {
    const std::string p = prefix + ".transformer.layers." + std::to_string(li);
    TrLayer& L = tr.layers[li];
    B200_TRY(keep_kmajor(h, p + ".self_attn.in_projs.0.weight", 3 * d, d, &L.in_w));
    B200_TRY(keep_kmajor(h, p + ".self_attn.out_projs.0.weight", d, d, &L.out_w));
    B200_TRY(keep(h, p + ".norm1.weight", {d}, &L.n1w));
    B200_TRY(keep(h, p + ".norm1.bias", {d}, &L.n1b));
    B200_TRY(keep(h, p + ".norm2.weight", {d}, &L.n2w));
    B200_TRY(keep(h, p + ".norm2.bias", {d}, &L.n2b));
    B200_TRY(keep_kmajor(h, p + ".linear1.weight", ff, d, &L.l1));
    B200_TRY(keep_kmajor(h, p + ".linear2.weight", d, ff, &L.l2));
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
// tile choice: the largest tile that still gives every SM a couple of CTAs; layers with a long reduction and few
// output tiles (deep SEANet convs, transformer linears at small batch) are cut along K as well.
template <int KIND>
int launch_gemm(b200_mimi* h, GemmArgs a) {
  auto ctas = [&](int bm, int bn) { return (long long)ceil_div(a.M, bm) * ceil_div(a.N, bn); };
  const long long want = 2 * 148;
  a.ksplit = 1; a.ws = nullptr;
  if (a.M > 64 && ctas(128, 128) >= want) {
    dim3 grid(ceil_div(a.N, 128), ceil_div(a.M, 128));
    B200_LAUNCH((mimi_gemm_kernel<128, 128, KIND>), grid, 256, 0, h->body, a);
  } else if (ctas(64, 128) >= want) {
    dim3 grid(ceil_div(a.N, 128), ceil_div(a.M, 64));
    B200_LAUNCH((mimi_gemm_kernel<64, 128, KIND>), grid, 256, 0, h->body, a);
  } else {
    const long long n64 = ctas(64, 64);
    const int nk = ceil_div(a.Kd, GK);
    int ks = 1;
    if (n64 < 148 && nk >= 32) {
      ks = (int)(want / n64);
      if (ks > nk / 8) ks = nk / 8;          // at least 8 k-blocks (128 k) per split
      if (ks > 16) ks = 16;
      if (ks < 1) ks = 1;
      if ((size_t)ks * a.M * a.N * 4 > h->splitk_bytes) ks = 1;
    }
    a.ksplit = ks; a.ws = h->splitk_ws;
    dim3 grid(ceil_div(a.N, 64), ceil_div(a.M, 64), ks);
    B200_LAUNCH((mimi_gemm_kernel<64, 64, KIND>), grid, 256, 0, h->body, a);
    if (ks > 1) {
      const long long n = (long long)a.M * a.N;
      B200_LAUNCH((gemm_splitk_reduce_kernel<KIND>), (unsigned)ceil_div64(n, 256), 256, 0, h->body, a);
    }
  }
  return check_launch("mimi_gemm");
}

// One SEANet layer.  (y, strides) = raw output; `next` (may be null) = the layer that consumes it: when that layer
// runs on the mimi_gemm.cuh path its activated input is written here, behind its carried state.
int launch_conv(b200_mimi* h, const ConvLayer& l, const float* x, long long xb, long long xc, long long xt,
                float* y, long long yb, long long yc, long long yt, const float* res, long long rb, long long rc,
                long long rt, const ConvLayer* next = nullptr) {
  const int B = h->batch;
  float* act = nullptr; long long ab = 0, ac = 0; int a_elu = 0;
  if (next && next->fast) {
    act = next->ext + next->D0; ab = (long long)next->cin * next->E; ac = next->E; a_elu = next->elu_in;
  }
  if (l.fast && l.kind == 0 && l.cout == 1) {     // last decoder conv: memory-bound reduction over (ci, kw)
    ConvCout1 q;
    q.ext = l.ext; q.eb = (long long)l.cin * l.E; q.E = l.E; q.off = l.D0 - l.P;
    q.wk = l.wk; q.bias = l.bias; q.y = y; q.yb = yb; q.yt = yt;
    q.B = B; q.Cin = l.cin; q.K = l.k; q.dil = l.dil; q.T = l.t_out;
    const long long n = (long long)B * l.t_out;
    B200_LAUNCH(conv_cout1_kernel, (unsigned)ceil_div64(n, 256), 256, (size_t)l.k * l.cin * 4, h->body, q);
    return check_launch(l.key.c_str());
  }
  if (!l.fast && l.kind == 0 && l.cin == 1 && l.stride == 1 && l.k <= 8 && !l.first && !l.elu_in && !res && xc == 0 && yt == 1) {
    ConvCin1 q;                                   // first encoder conv: one thread per sample makes all channels
    q.x = x; q.xb = xb; q.xt = xt; q.st = l.state; q.P = l.P; q.w = l.w; q.bias = l.bias;
    q.y = y; q.yb = yb; q.yc = yc; q.a = act; q.ab = ab; q.ac = ac; q.a_elu = a_elu;
    q.B = B; q.Cout = l.cout; q.K = l.k; q.T = l.t_out;
    const long long n = (long long)B * l.t_out;
    B200_LAUNCH(conv_cin1_kernel, (unsigned)ceil_div64(n, 256), 256, (size_t)(l.cout * l.k + l.cout) * 4, h->body, q);
    return check_launch(l.key.c_str());
  }
  if (l.fast) {
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.wk = l.wk; a.in = l.ext; a.in_b = (long long)l.cin * l.E; a.in_E = l.E;
    a.Cin = l.cin; a.bias = l.bias;
    a.y = y; a.yb = yb; a.yc = yc; a.yt = yt;
    a.a = act; a.ab = ab; a.ac = ac; a.at = 1; a.a_elu = a_elu;
    if (l.kind == 0) {
      a.in_off = l.D0 - l.P;
      a.M = l.cout; a.N = B * l.t_out; a.Kd = l.cin * l.k; a.stride = l.stride; a.dil = l.dil; a.T = l.t_out;
      a.res = res; a.rb = rb; a.rc = rc; a.rt = rt;
      const bool t4 = l.t_out % 4 == 0;
      a.vec_y = y && t4 && yt == 1 && yb % 4 == 0 && yc % 4 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
      a.vec_a = act && t4;
      a.vec_r = res && t4 && rt == 1 && rb % 4 == 0 && rc % 4 == 0;
      return launch_gemm<G_CONV>(h, a);
    }
    a.in_off = l.D0;
    a.M = l.cout * l.stride; a.N = B * (l.t_in + 1); a.Kd = 2 * l.cin; a.stride = l.stride; a.T = l.t_in;
    a.partial = l.state; a.scratch = l.scratch;
    // consecutive GEMM rows of one channel are consecutive output samples: move them as float4 / float2 when S allows
    const bool al = (reinterpret_cast<uintptr_t>(y) & 15) == 0 && yb % 4 == 0 && yc % 4 == 0;
    a.vec_y = (l.stride % 4 == 0 && al) ? 4 : (l.stride % 2 == 0 && al) ? 2 : 1;
    return launch_gemm<G_CONVTR>(h, a);
  }
  if (l.kind == 0) {
    ConvP p;
    p.x = x; p.xb = xb; p.xc = xc; p.xt = xt; p.Tin = l.t_in;
    p.st = l.state; p.P = l.P; p.first = l.first;
    p.w = l.w; p.bias = l.bias;
    p.y = y; p.yb = yb; p.yc = yc; p.yt = yt;
    p.res = res; p.rb = rb; p.rc = rc; p.rt = rt;
    p.a = act; p.ab = ab; p.ac = ac; p.at = 1; p.a_elu = a_elu;
    p.B = B; p.Cin = l.cin; p.Cout = l.cout; p.K = l.k; p.stride = l.stride; p.dil = l.dil; p.Tout = l.t_out;
    p.elu_in = l.elu_in;
    p.M = l.cout; p.N = B * l.t_out; p.Kd = l.cin * l.k; p.cin_aligned = (l.cin % BK) == 0;
    dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM));
    B200_LAUNCH((igemm_f32_kernel<ConvP, false>), grid, 256, 0, h->body, p);
  } else {
    if (act) B200_FAIL(B200_ERR_INVALID, "first-generation convtr cannot feed a mimi_gemm layer");
    ConvTrP p;
    p.x = x; p.xb = xb; p.xc = xc; p.xt = xt; p.T = l.t_in;
    p.partial = l.state; p.scratch = l.scratch; p.w = l.w; p.bias = l.bias;
    p.y = y; p.yb = yb; p.yc = yc; p.yt = yt;
    p.B = B; p.Cin = l.cin; p.Cout = l.cout; p.S = l.stride; p.elu_in = l.elu_in;
    p.M = l.cout * l.stride; p.N = B * (l.t_in + 1); p.Kd = 2 * l.cin; p.cin_aligned = (l.cin % BK) == 0;
    dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM));
    B200_LAUNCH((igemm_f32_kernel<ConvTrP, false>), grid, 256, 0, h->body, p);
  }
  return check_launch(l.key.c_str());
}

// y[n][m] = epi(sum_k x[n][k] * w[m][k]) on token-major activations; w is k-major [K][M]
int launch_linear(b200_mimi* h, const float* x, int K, const float* w, float* y, int M, int ntok, int epi,
                  const float* res, const float* scale) {
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  a.wk = w; a.in = x; a.in_E = K; a.M = M; a.N = ntok; a.Kd = K;
  a.y = y; a.yb = M; a.res = res; a.scale = scale;
  a.lin_epi = epi == EPI_GELU ? GL_GELU : epi == EPI_RES_SCALE ? GL_RES_SCALE : GL_NONE;
  return launch_gemm<G_LIN>(h, a);
}

// StreamingTransformer.forward for T tokens per row (transformer.py:894-929, layer :752-802)
int run_transformer(b200_mimi* h, Transformer& tr, const float* x_in, float* x, int T) {
  const auto& c = h->cfg;
  const int B = h->batch, ntok = B * T, d = c.tr_d_model, H = c.tr_num_heads, D = d / H, ff = c.tr_dim_feedforward;
  const float nl = -logf(c.tr_max_period) * 2.f / (float)D;
  const size_t attn_smem = (size_t)(T * D + T * c.tr_context) * sizeof(float);
  for (size_t li = 0; li < tr.layers.size(); ++li) {
    TrLayer& L = tr.layers[li];
    const float* cur = li == 0 ? x_in : x;
    B200_LAUNCH(layernorm_kernel, ceil_div(ntok * 32, 256), 256, 0, h->body, cur, L.n1w, L.n1b, h->tr_xn, ntok, d, 1e-5f);
    B200_TRY(launch_linear(h, h->tr_xn, d, L.in_w, h->tr_qkv, 3 * d, ntok, EPI_NONE, nullptr, nullptr));
    {
      const long long total = (long long)B * T * H * (D / 2);
      B200_LAUNCH(rope_append_f32_kernel, (unsigned)ceil_div64(total, 256), 256, 0, h->body, h->tr_qkv, h->tr_q, L.kc,
                  L.vc, tr.offset, h->exec_mask, B, T, H, D, c.tr_context, nl);
    }
    B200_LAUNCH((ring_attn_f32_kernel<64>), B * H, 128, attn_smem, h->body, h->tr_q, L.kc, L.vc, h->tr_ao, tr.offset,
                h->exec_mask, T, H, c.tr_context, c.tr_context);
    B200_TRY(launch_linear(h, h->tr_ao, d, L.out_w, x, d, ntok, EPI_RES_SCALE, cur, L.ls1));
    B200_LAUNCH(layernorm_kernel, ceil_div(ntok * 32, 256), 256, 0, h->body, x, L.n2w, L.n2b, h->tr_xn, ntok, d, 1e-5f);
    B200_TRY(launch_linear(h, h->tr_xn, d, L.l1, h->tr_h, ff, ntok, EPI_GELU, nullptr, nullptr));
    B200_TRY(launch_linear(h, h->tr_h, ff, L.l2, x, d, ntok, EPI_RES_SCALE, x, L.ls2));
  }
  B200_LAUNCH(advance_offsets_kernel, ceil_div(B, 128), 128, 0, h->body, tr.offset, h->exec_mask, B, T);
  return check_launch("mimi transformer");
}

int run_seanet(b200_mimi* h, std::vector<ConvLayer>& layers, std::vector<Buf>& bufs, const float* x0, long long xb,
               long long xc, long long xt, float* last_out, long long lb, long long lc, long long lt) {
  // bufs[i] is the output of layers[i]; layer 0 reads (x0, strides); the last layer writes last_out.
  if (layers[0].fast) {   // its input comes from outside the SEANet (transformer output): copy it behind the carried state
    const ConvLayer& l = layers[0];
    const long long n = (long long)h->batch * l.cin * l.t_in;
    B200_LAUNCH(fill_act_kernel, (unsigned)ceil_div64(n, 256), 256, 0, h->body, x0, xb, xc, xt, l.ext + l.D0,
                (long long)l.cin * l.E, (long long)l.E, 1LL, h->batch, l.cin, l.t_in, (int)l.elu_in);
  }
  for (size_t i = 0; i < layers.size(); ++i) {
    ConvLayer& l = layers[i];
    const float* x; long long sb, sc, st;
    if (i == 0) { x = x0; sb = xb; sc = xc; st = xt; }
    else { const Buf& b = bufs[i - 1]; x = b.p; sb = b.sb(0); sc = b.sc(); st = b.st(); }
    float* y; long long yb, yc, yt;
    if (i + 1 == layers.size() && last_out) { y = last_out; yb = lb; yc = lc; yt = lt; }
    else { const Buf& b = bufs[i]; y = b.p; yb = b.sb(0); yc = b.sc(); yt = b.st(); }
    const float* res = nullptr; long long rb = 0, rc = 0, rt = 0;
    if (l.res_from == -2) {   // residual block: skip connection is the input of the first conv of the block
      if (i < 2) B200_FAIL(B200_ERR_INVALID, "bad residual plan");
      const Buf& b = bufs[i - 2];
      res = b.p; rb = b.sb(0); rc = b.sc(); rt = b.st();
    }
    const ConvLayer* next = i + 1 < layers.size() ? &layers[i + 1] : nullptr;
    B200_TRY(launch_conv(h, l, x, sb, sc, st, y, yb, yc, yt, res, rb, rc, rt, next));
  }
  return B200_OK;
}

int commit_states(b200_mimi* h, bool encoder) {
  const int B = h->batch;
  if (encoder) {
    if (h->n_enc_commits) {
      dim3 grid(ceil_div(h->max_enc_rows, 128), h->n_enc_commits);
      B200_LAUNCH(conv_commit_kernel, grid, 128, 0, h->body, h->enc_commits, h->n_enc_commits, h->exec_mask, B);
      B200_LAUNCH(conv_clear_first_kernel, ceil_div(h->n_enc_commits * B, 128), 128, 0, h->body, h->enc_commits,
                  h->n_enc_commits, h->exec_mask, B);
    }
    if (h->n_enc_ext) {
      dim3 grid(ceil_div(h->max_enc_ext_rows, 128), h->n_enc_ext);
      B200_LAUNCH(ext_commit_kernel, grid, 128, 0, h->body, h->enc_ext_commits, h->exec_mask, B);
    }
  } else {
    if (h->n_dec_commits) {
      dim3 grid(ceil_div(h->max_dec_rows, 128), h->n_dec_commits);
      B200_LAUNCH(conv_commit_kernel, grid, 128, 0, h->body, h->dec_commits, h->n_dec_commits, h->exec_mask, B);
    }
    if (h->n_dec_ext) {
      dim3 grid(ceil_div(h->max_dec_ext_rows, 128), h->n_dec_ext);
      B200_LAUNCH(ext_commit_kernel, grid, 128, 0, h->body, h->dec_ext_commits, h->exec_mask, B);
    }
    dim3 g2((unsigned)ceil_div64(h->max_tr_rows, 256), h->n_dec_tr_commits);
    B200_LAUNCH(convtr_commit_kernel, g2, 256, 0, h->body, h->dec_tr_commits, h->exec_mask, B);
  }
  return check_launch("commit_states");
}

// in_frame [B][frame] -> latent [B][C]   (static buffers only: this is what gets captured into a graph)
int encode_body(b200_mimi* h) {
  const int fs = h->frame_size, d = h->cfg.dimension;
  const int T = fs / h->hop;   // encoder tokens per frame (2)
  // SEANet encoder; the last conv writes token-major [B][T][C] for the transformer
  B200_TRY(run_seanet(h, h->enc, h->enc_bufs, h->in_frame, fs, 0, 1, h->tok_in_enc, (long long)T * d, 1, d));
  B200_TRY(run_transformer(h, h->enc_tr, h->tok_in_enc, h->tok_enc, T));
  // learnt down-sampling conv reads token-major, writes latent [B][C][1]
  B200_TRY(launch_conv(h, h->down, h->tok_enc, (long long)T * d, 1, d, h->latent, d, 1, 1, nullptr, 0, 0, 0));
  B200_TRY(commit_states(h, true));
  return B200_OK;
}

int load_frame(b200_mimi* h, const float* pcm, int n_frames, int f) {
  const int fs = h->frame_size;
  B200_CUDA(cudaMemcpy2DAsync(h->in_frame, (size_t)fs * 4, pcm + (size_t)f * fs, (size_t)fs * n_frames * 4,
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
    dim3 grid(Q, 2);
    B200_LAUNCH(rvq_project_kernel, grid, 256, (size_t)c.dimension * 4, h->body, lat, lb, lc, lt, n_cols, h->wT[0], h->wT[1],
                h->rvq_res[0], h->rvq_res[1], c.dimension, Dq);
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
  a.Dq = c.q_dimension; a.Cout = c.dimension; a.bins = c.q_bins;
  dim3 grid(h->batch * n_cols, (c.dimension + 63) / 64);
  B200_LAUNCH(rvq_decode_kernel, grid, 256, (2 * c.q_dimension + 256) * sizeof(float), h->body, a);
  return check_launch("rvq_decode");
}

// latent_q [B][C] -> out_frame [B][frame]   (static buffers only)
int decode_latent_body(b200_mimi* h) {
  const int B = h->batch, fs = h->frame_size, d = h->cfg.dimension, S = h->rs;
  const int T = S;   // tokens per frame after up-sampling
  {
    const long long total = (long long)B * 2 * S * d;   // (T_in + 1) * S * C with T_in = 1
    B200_LAUNCH(upsample_dw_kernel, (unsigned)ceil_div64(total, 256), 256, 0, h->body, h->latent_q, (long long)d, 1LL, 1LL, 1,
                h->up_w, h->up_partial, h->up_scratch, h->tok_in_dec, B, d, S);
  }
  B200_TRY(run_transformer(h, h->dec_tr, h->tok_in_dec, h->tok_dec, T));
  B200_TRY(run_seanet(h, h->dec, h->dec_bufs, h->tok_dec, (long long)T * d, 1, d, h->out_frame, (long long)fs, 0, 1));
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

void register_tap(b200_mimi* h, const std::string& name, const float* p, int64_t n) { h->taps[name] = {p, n}; }

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
  b200_mimi* h = new b200_mimi();
  h->cfg = *cfg;
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
  return h->store.put(name, data_dev, dtype, ndim, shape);
}

int b200_mimi_finalize(b200_mimi* h) {
  if (!h) B200_FAIL(B200_ERR_INVALID, "mimi_finalize: null handle");
  if (h->finalized) return B200_OK;
  for (auto& l : h->enc) B200_TRY(pack_conv(h, l));
  for (auto& l : h->dec) B200_TRY(pack_conv(h, l));
  B200_TRY(pack_conv(h, h->down));
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
  b200_mimi_streaming_end(h);
  h->store.release_all();
  h->weights.free_all();
  if (h->pin_pcm) cudaFreeHost(h->pin_pcm);
  if (h->pin_codes) cudaFreeHost(h->pin_codes);
  if (h->dev_pcm) cudaFree(h->dev_pcm);
  if (h->dev_codes) cudaFree(h->dev_codes);
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
  const auto& c = h->cfg;
  const int B = batch, d = c.dimension;
  h->stream = static_cast<cudaStream_t>(stream);
  h->taps.clear();
  Arena& A = h->state;
  B200_TRY(A.alloc_t(&h->exec_mask, B, false));
  B200_CUDA(cudaMemset(h->exec_mask, 1, B));
  B200_TRY(A.alloc_t(&h->in_frame, (size_t)B * h->frame_size));

  // flags for replicate-padded convs (only the down-sampling conv in Mimi)
  h->n_first = 1;
  B200_TRY(A.alloc_t(&h->first_flags, (size_t)h->n_first * B, false));
  B200_CUDA(cudaMemset(h->first_flags, 1, (size_t)h->n_first * B));

  std::vector<ConvCommit> enc_c, dec_c;
  std::vector<ExtCommit> enc_e, dec_e;
  std::vector<ConvTrCommit> dec_t;
  h->max_enc_rows = h->max_dec_rows = 0;
  h->max_enc_ext_rows = h->max_dec_ext_rows = 0;
  h->max_tr_rows = 0;

  auto setup = [&](std::vector<ConvLayer>& layers, std::vector<Buf>& bufs, int t0, bool last_token_major,
                   const float* x0, long long x0b, long long x0c, long long x0t, std::vector<ConvCommit>& commits,
                   std::vector<ExtCommit>& ecommits, bool is_dec) -> int {
    bufs.assign(layers.size(), Buf());
    int t = t0;
    for (size_t i = 0; i < layers.size(); ++i) {
      ConvLayer& l = layers[i];
      l.t_in = t;
      if (l.kind == 0) {
        if (t % l.stride) B200_FAIL(B200_ERR_INVALID, "%s: %d samples not divisible by stride %d", l.key.c_str(), t, l.stride);
        l.t_out = t / l.stride;
        l.P = (l.k - 1) * l.dil + 1 - l.stride;
        if (l.fast) {          // carried samples live inside ext, right in front of the 16-byte aligned frame
          l.D0 = (l.P + 3) / 4 * 4;
          l.E = (l.D0 + t + 3) / 4 * 4;
          B200_TRY(A.alloc_t(&l.ext, (size_t)B * l.cin * l.E));
        } else if (l.P > 0) {
          B200_TRY(A.alloc_t(&l.state, (size_t)B * l.cin * l.P));
        }
      } else {
        l.t_out = t * l.stride;
        l.P = l.k - l.stride;
        B200_TRY(A.alloc_t(&l.state, (size_t)B * l.cout * l.P));
        B200_TRY(A.alloc_t(&l.scratch, (size_t)B * l.cout * l.P));
        if (l.fast) {          // zero column on both sides of the frame: taps that fall outside read 0
          l.D0 = 4;
          l.E = (l.D0 + t + 1 + 3) / 4 * 4;
          B200_TRY(A.alloc_t(&l.ext, (size_t)B * l.cin * l.E));
        }
      }
      t = l.t_out;
      Buf& b = bufs[i];
      b.c = l.cout; b.t = l.t_out;
      const bool is_last = i + 1 == layers.size();
      if (!(is_last && is_dec)) {   // the decoder's last conv writes straight into the caller's PCM buffer
        b.token_major = is_last && last_token_major;
        B200_TRY(A.alloc_t(&b.p, (size_t)B * l.cout * l.t_out));
        if (!l.tap.empty()) register_tap(h, l.tap, b.p, (int64_t)B * l.cout * l.t_out);
      }
      // commit descriptor
      const float* x; long long sb, sc, st;
      if (i == 0) { x = x0; sb = x0b; sc = x0c; st = x0t; }
      else { const Buf& pb = bufs[i - 1]; x = pb.p; sb = pb.sb(0); sc = pb.sc(); st = pb.st(); }
      if (l.kind == 0 && l.P > 0 && l.fast) {
        ExtCommit ec;
        ec.ext = l.ext + (l.D0 - l.P); ec.P = l.P; ec.T = l.t_in; ec.E = l.E; ec.Cin = l.cin;
        ecommits.push_back(ec);
        int& mx = is_dec ? h->max_dec_ext_rows : h->max_enc_ext_rows;
        if (B * l.cin > mx) mx = B * l.cin;
      } else if (l.kind == 0 && l.P > 0) {
        ConvCommit cc;
        cc.x = x; cc.xb = sb; cc.xc = sc; cc.xt = st; cc.Tin = l.t_in; cc.st = l.state; cc.P = l.P; cc.Cin = l.cin;
        cc.elu_in = l.elu_in; cc.first = nullptr;
        commits.push_back(cc);
        int rows = B * l.cin;
        int& mx = is_dec ? h->max_dec_rows : h->max_enc_rows;
        if (rows > mx) mx = rows;
      } else if (l.kind == 1) {
        ConvTrCommit tc;
        tc.partial = l.state; tc.scratch = l.scratch; tc.per_row = l.cout * l.P;
        dec_t.push_back(tc);
        if ((long long)B * tc.per_row > h->max_tr_rows) h->max_tr_rows = (long long)B * tc.per_row;
      }
    }
    return B200_OK;
  };

  const int T = h->rs;   // transformer tokens per frame on both sides
  B200_TRY(A.alloc_t(&h->tok_in_enc, (size_t)B * T * d));   // written by the last encoder conv (token-major)
  B200_TRY(A.alloc_t(&h->tok_enc, (size_t)B * T * d));
  B200_TRY(A.alloc_t(&h->latent, (size_t)B * d));
  B200_TRY(A.alloc_t(&h->latent_q, (size_t)B * d));
  B200_TRY(A.alloc_t(&h->tok_in_dec, (size_t)B * T * d));
  B200_TRY(A.alloc_t(&h->tok_dec, (size_t)B * T * d));
  B200_TRY(setup(h->enc, h->enc_bufs, h->frame_size, true, h->in_frame, h->frame_size, 0, 1, enc_c, enc_e, false));
  // the last encoder conv writes tok_in_enc instead of its own buffer
  register_tap(h, h->enc.back().tap, h->tok_in_enc, (int64_t)B * T * d);
  if (h->enc.back().t_out != T) B200_FAIL(B200_ERR_INVALID, "encoder yields %d tokens/frame, expected %d", h->enc.back().t_out, T);
  B200_TRY(setup(h->dec, h->dec_bufs, T, false, h->tok_dec, (long long)T * d, 1, d, dec_c, dec_e, true));
  if (h->dec.back().t_out != h->frame_size) B200_FAIL(B200_ERR_INVALID, "decoder yields %d samples/frame", h->dec.back().t_out);
  register_tap(h, "enc.tr", h->tok_enc, (int64_t)B * T * d);
  register_tap(h, "enc.latent", h->latent, (int64_t)B * d);
  register_tap(h, "dec.latent", h->latent_q, (int64_t)B * d);
  register_tap(h, "dec.up", h->tok_in_dec, (int64_t)B * T * d);
  register_tap(h, "dec.tr", h->tok_dec, (int64_t)B * T * d);

  // down-sampling conv (replicate pad): reads tok_enc token-major
  {
    ConvLayer& l = h->down;
    l.t_in = T; l.t_out = T / l.stride; l.P = l.k - l.stride;
    B200_TRY(A.alloc_t(&l.state, (size_t)B * l.cin * l.P));
    l.first = h->first_flags;
    ConvCommit cc;
    cc.x = h->tok_enc; cc.xb = (long long)T * d; cc.xc = 1; cc.xt = d; cc.Tin = T; cc.st = l.state; cc.P = l.P;
    cc.Cin = l.cin; cc.elu_in = 0; cc.first = l.first;
    enc_c.push_back(cc);
    if (B * l.cin > h->max_enc_rows) h->max_enc_rows = B * l.cin;
  }
  // up-sampling depth-wise convtr
  B200_TRY(A.alloc_t(&h->up_partial, (size_t)B * d * h->rs));
  B200_TRY(A.alloc_t(&h->up_scratch, (size_t)B * d * h->rs));
  {
    ConvTrCommit tc;
    tc.partial = h->up_partial; tc.scratch = h->up_scratch; tc.per_row = d * h->rs;
    dec_t.push_back(tc);
    if ((long long)B * tc.per_row > h->max_tr_rows) h->max_tr_rows = (long long)B * tc.per_row;
  }
  h->n_enc_commits = (int)enc_c.size();
  h->n_dec_commits = (int)dec_c.size();
  h->n_dec_tr_commits = (int)dec_t.size();
  h->n_enc_ext = (int)enc_e.size();
  h->n_dec_ext = (int)dec_e.size();
  B200_TRY(A.alloc_t(&h->enc_commits, enc_c.size() ? enc_c.size() : 1, false));
  B200_TRY(A.alloc_t(&h->dec_commits, dec_c.size() ? dec_c.size() : 1, false));
  B200_TRY(A.alloc_t(&h->dec_tr_commits, dec_t.size(), false));
  B200_TRY(A.alloc_t(&h->enc_ext_commits, enc_e.size() ? enc_e.size() : 1, false));
  B200_TRY(A.alloc_t(&h->dec_ext_commits, dec_e.size() ? dec_e.size() : 1, false));
  if (!enc_c.empty())
    B200_CUDA(cudaMemcpy(h->enc_commits, enc_c.data(), enc_c.size() * sizeof(ConvCommit), cudaMemcpyHostToDevice));
  if (!dec_c.empty())
    B200_CUDA(cudaMemcpy(h->dec_commits, dec_c.data(), dec_c.size() * sizeof(ConvCommit), cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemcpy(h->dec_tr_commits, dec_t.data(), dec_t.size() * sizeof(ConvTrCommit), cudaMemcpyHostToDevice));
  if (!enc_e.empty())
    B200_CUDA(cudaMemcpy(h->enc_ext_commits, enc_e.data(), enc_e.size() * sizeof(ExtCommit), cudaMemcpyHostToDevice));
  if (!dec_e.empty())
    B200_CUDA(cudaMemcpy(h->dec_ext_commits, dec_e.data(), dec_e.size() * sizeof(ExtCommit), cudaMemcpyHostToDevice));

  // transformers
  const int H = c.tr_num_heads, D = d / H, ff = c.tr_dim_feedforward;
  for (Transformer* tr : {&h->enc_tr, &h->dec_tr}) {
    B200_TRY(A.alloc_t(&tr->offset, B));
    for (auto& L : tr->layers) {
      B200_TRY(A.alloc_t(&L.kc, (size_t)B * H * c.tr_context * D));
      B200_TRY(A.alloc_t(&L.vc, (size_t)B * H * c.tr_context * D));
    }
  }
  const size_t ntok = (size_t)B * T;
  B200_TRY(A.alloc_t(&h->tr_xn, ntok * d));
  B200_TRY(A.alloc_t(&h->tr_qkv, ntok * 3 * d));
  B200_TRY(A.alloc_t(&h->tr_q, ntok * d));
  B200_TRY(A.alloc_t(&h->tr_ao, ntok * d));
  B200_TRY(A.alloc_t(&h->tr_h, ntok * ff));
  B200_TRY(A.alloc_t(&h->scratch_codes, (size_t)B * c.q_n_q));
  B200_TRY(A.alloc_t(&h->enc_codes, (size_t)B * c.q_n_q));
  B200_TRY(A.alloc_t(&h->dec_codes, (size_t)B * c.q_n_q));
  B200_TRY(A.alloc_t(&h->out_frame, (size_t)B * h->frame_size));
  B200_CUDA(cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking));
  B200_CUDA(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
  h->body = h->stream;
  // what a streaming-state snapshot holds (streaming.py:158-181): masks, carried conv samples, overlap-add
  // partials, both transformers' KV rings and offsets
  A.mark_state(h->exec_mask, B);
  A.mark_state(h->first_flags, (size_t)h->n_first * B);
  for (auto* layers : {&h->enc, &h->dec})
    for (auto& l : *layers) {
      if (l.ext && l.kind == 0) A.mark_state(l.ext, (size_t)B * l.cin * l.E * 4);
      if (l.state) A.mark_state(l.state, (size_t)B * (l.kind == 0 ? l.cin : l.cout) * l.P * 4);
    }
  A.mark_state(h->down.state, (size_t)B * h->down.cin * h->down.P * 4);
  A.mark_state(h->up_partial, (size_t)B * d * h->rs * 4);
  for (Transformer* tr : {&h->enc_tr, &h->dec_tr}) {
    A.mark_state(tr->offset, (size_t)B * 8);
    for (auto& L : tr->layers) {
      A.mark_state(L.kc, (size_t)B * H * c.tr_context * D * 4);
      A.mark_state(L.vc, (size_t)B * H * c.tr_context * D * 4);
    }
  }
  B200_TRY(ensure_rvq_workspace(h, B));
  h->splitk_bytes = (size_t)16 << 20;
  B200_TRY(A.alloc(reinterpret_cast<void**>(&h->splitk_ws), h->splitk_bytes, false));
  B200_CUDA(cudaDeviceSynchronize());
  h->batch = B;
  return B200_OK;
}

int b200_mimi_streaming_end(b200_mimi* h) {
  if (!h) return B200_OK;
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
  h->batch = 0;
  h->taps.clear();
  return B200_OK;
}

int b200_mimi_reset(b200_mimi* h, const uint8_t* reset_mask_dev) {
  B200_TRY(ensure_streaming(h, "mimi_reset"));
  const int B = h->batch;
  auto zero = [&](float* buf, long long per_row) {
    if (!buf || per_row <= 0) return;
    B200_LAUNCH(zero_rows_kernel, (unsigned)ceil_div64((long long)B * per_row, 256), 256, 0, h->stream, buf, per_row,
                reset_mask_dev, B);
  };
  for (auto* layers : {&h->enc, &h->dec})
    for (auto& l : *layers) {
      if (l.kind == 0 && l.fast) {
        if (l.P > 0)
          B200_LAUNCH(ext_zero_kernel, (unsigned)ceil_div64((long long)B * l.cin * l.P, 256), 256, 0, h->stream,
                      l.ext + (l.D0 - l.P), l.P, l.E, l.cin, reset_mask_dev, B);
      } else {
        zero(l.state, l.kind == 0 ? (long long)l.cin * l.P : (long long)l.cout * l.P);
      }
    }
  zero(h->down.state, (long long)h->down.cin * h->down.P);
  zero(h->up_partial, (long long)h->cfg.dimension * h->rs);
  B200_LAUNCH(reset_flags_kernel, ceil_div(B, 128), 128, 0, h->stream, h->first_flags, h->n_first, h->enc_tr.offset,
              h->dec_tr.offset, h->exec_mask, reset_mask_dev, B);
  return check_launch("mimi_reset");
}

int b200_mimi_set_exec_mask(b200_mimi* h, const uint8_t* exec_mask_dev) {
  B200_TRY(ensure_streaming(h, "mimi_set_exec_mask"));
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

int b200_mimi_read_buffer(b200_mimi* h, const char* name, float* dst_dev, int64_t capacity, int64_t* numel) {
  B200_TRY(ensure_streaming(h, "mimi_read_buffer"));
  auto it = h->taps.find(name ? name : "");
  if (it == h->taps.end()) B200_FAIL(B200_ERR_INVALID, "mimi_read_buffer: unknown buffer '%s'", name ? name : "(null)");
  if (numel) *numel = it->second.second;
  if (!dst_dev) return B200_OK;
  if (capacity < it->second.second) B200_FAIL(B200_ERR_SHAPE, "mimi_read_buffer: destination too small");
  B200_CUDA(cudaMemcpyAsync(dst_dev, it->second.first, (size_t)it->second.second * 4, cudaMemcpyDeviceToDevice, h->stream));
  return B200_OK;
}

int64_t b200_mimi_algorithmic_bytes(b200_mimi* h) {
  if (!h || h->batch <= 0) return 0;
  const auto& c = h->cfg;
  // weights once per step (encode + decode) + active codebooks (row-major for decode, transposed for encode)
  int64_t bytes = h->weight_bytes;
  bytes += (int64_t)2 * h->num_codebooks * c.q_bins * c.q_dimension * 4 + (int64_t)4 * c.q_dimension * c.dimension * 4;
  // per session: both transformer KV rings read once, conv/convtr carried state read + written, PCM in/out
  int64_t per_row = (int64_t)2 * c.tr_num_layers * 2 * c.tr_context * c.tr_d_model * 4;
  int64_t st = 0;
  for (auto* layers : {&h->enc, &h->dec})
    for (auto& l : *layers) st += (l.kind == 0 ? (int64_t)l.cin * l.P : (int64_t)l.cout * l.P);
  st += (int64_t)h->down.cin * h->down.P + (int64_t)c.dimension * h->rs;
  per_row += 2 * st * 4 + (int64_t)2 * h->frame_size * 4 + h->num_codebooks * 8;
  return bytes + per_row * h->batch;
}

}  // extern "C"
