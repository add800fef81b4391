// Kernel-level C entry points used by the parity tests (same kernels the handles launch).
#include "gemm_sk.cuh"
#include "lm_kernels.cuh"
#include "mimi_kernels.cuh"

using namespace b200;

extern "C" {

static int sk_scratch(float** ws, int** counters) {
  static float* g_ws = nullptr;
  static int* g_counters = nullptr;
  if (!g_ws) {
    B200_CUDA(cudaMalloc(&g_ws, tc::sk_workspace_bytes(256)));
    B200_CUDA(cudaMalloc(&g_counters, tc::SK_MAX_TILES * sizeof(int)));
    B200_CUDA(cudaMemset(g_counters, 0, tc::SK_MAX_TILES * sizeof(int)));
  }
  *ws = g_ws; *counters = g_counters;
  return B200_OK;
}

int b200_op_linear_bf16(const void* x_dev, const void* w_dev, void* y_dev, int M, int N, int K, void* stream) {
  if (!x_dev || !w_dev || !y_dev || M < 1 || N < 1 || K < 8 || K % 8) B200_FAIL(B200_ERR_SHAPE, "op_linear_bf16: bad shape");
  if (!tc::sk_supported(M, N, K, 0)) B200_FAIL(B200_ERR_SHAPE, "op_linear_bf16: unsupported shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static tc::GemmPlanCache cache;
  float* ws = nullptr; int* counters = nullptr;
  B200_TRY(sk_scratch(&ws, &counters));
  void* tiles = nullptr;
  B200_CUDA(cudaMalloc(&tiles, tc::sk_packed_bytes(N, K, 0, 0)));
  int rc = tc::sk_pack_weights(static_cast<const __nv_bfloat16*>(w_dev), tiles, N, K, 0, 0, st);
  tc::SkTuning t;
  if (rc == B200_OK)
    rc = tc::sk_linear(cache, static_cast<const __nv_bfloat16*>(x_dev), K, tiles, static_cast<__nv_bfloat16*>(y_dev), N, nullptr, 0,
                       M, N, K, 0, 0, ws, counters, t, st);
  cudaStreamSynchronize(st);
  cudaFree(tiles);
  return rc;
}

int64_t b200_op_packed_bytes(int N, int K, int epi, int gate_rows) { return (int64_t)tc::sk_packed_bytes(N, K, epi, gate_rows); }

int b200_op_pack_tiles(const void* w_dev, void* out_dev, int N, int K, int epi, int gate_rows, void* stream) {
  if (!w_dev || !out_dev) B200_FAIL(B200_ERR_INVALID, "op_pack_tiles: null pointer");
  return tc::sk_pack_weights(static_cast<const __nv_bfloat16*>(w_dev), out_dev, N, K, epi, gate_rows,
                             static_cast<cudaStream_t>(stream));
}

int b200_op_linear_sk(const void* x_dev, const void* w_tiles_dev, void* y_dev, const void* res_dev, int M, int N, int K,
                      int epi, int gate_rows, int grid, int smem_budget, int stream_only, void* stream) {
  if (!x_dev || !w_tiles_dev || !y_dev) B200_FAIL(B200_ERR_INVALID, "op_linear_sk: null pointer");
  if (!tc::sk_supported(M, N, K, epi)) B200_FAIL(B200_ERR_SHAPE, "op_linear_sk: unsupported shape");
  static tc::GemmPlanCache cache;
  float* ws = nullptr; int* counters = nullptr;
  B200_TRY(sk_scratch(&ws, &counters));
  tc::SkTuning t;
  // grid > 0: stream-K with that many CTAs; grid < 0: cluster split-K with |grid| CTAs per tile; 0: the LM's own choice
  t.grid = grid > 0 ? grid : 0; t.cluster = grid < 0 ? -grid : 0; t.smem_budget = smem_budget; t.stream_only = stream_only;
  t.force_split = grid > 0;     // an explicit grid (parity tests) exercises the tile-cutting path at every M
  const int out_cols = epi == 2 ? gate_rows : N;
  return tc::sk_linear(cache, static_cast<const __nv_bfloat16*>(x_dev), K, w_tiles_dev, static_cast<__nv_bfloat16*>(y_dev),
                       out_cols, static_cast<const __nv_bfloat16*>(res_dev), out_cols, M, N, K, epi, gate_rows, ws, counters, t,
                       static_cast<cudaStream_t>(stream));
}

int b200_op_set_gemv_max_rows(int max_rows) { return tc::sk_set_gemv_max_m(max_rows); }

int64_t b200_op_packed_bytes_i8(int N, int K, int epi, int gate_rows) { return (int64_t)tc::sk_packed_bytes_i8(N, K, epi, gate_rows); }

int b200_op_quant_pack_tiles(const void* w_dev, void* tiles_dev, float* scales_dev, int N, int K, int epi, int gate_rows,
                             void* stream) {
  if (!w_dev || !tiles_dev || !scales_dev) B200_FAIL(B200_ERR_INVALID, "op_quant_pack_tiles: null pointer");
  return tc::sk_quant_pack_weights(static_cast<const __nv_bfloat16*>(w_dev), tiles_dev, scales_dev, N, K, epi, gate_rows,
                                   static_cast<cudaStream_t>(stream));
}

int b200_op_quantize_rows(const void* x_dev, void* xq_dev, float* sa_dev, int M, int K, void* stream) {
  if (!x_dev || !xq_dev || !sa_dev) B200_FAIL(B200_ERR_INVALID, "op_quantize_rows: null pointer");
  return tc::sk_quantize_rows(static_cast<const __nv_bfloat16*>(x_dev), K, xq_dev, sa_dev, M, K, static_cast<cudaStream_t>(stream));
}

int b200_op_linear_i8(const void* xq_dev, const float* sa_dev, const void* w_tiles_dev, const float* sw_dev, void* y_dev,
                      const void* res_dev, int M, int N, int K, int epi, int gate_rows, void* stream) {
  if (!xq_dev || !sa_dev || !w_tiles_dev || !sw_dev || !y_dev) B200_FAIL(B200_ERR_INVALID, "op_linear_i8: null pointer");
  if (!tc::sk_supported(M, N, K, epi)) B200_FAIL(B200_ERR_SHAPE, "op_linear_i8: unsupported shape");
  static tc::GemmPlanCache cache;
  float* ws = nullptr; int* counters = nullptr;
  B200_TRY(sk_scratch(&ws, &counters));      // int32 partials of cut tiles share the bf16 path's workspace
  tc::SkTuning t;
  t.xq = xq_dev; t.sa = sa_dev; t.sw = sw_dev;
  const int out_cols = epi == 2 ? gate_rows : N;
  return tc::sk_linear(cache, nullptr, K, w_tiles_dev, static_cast<__nv_bfloat16*>(y_dev), out_cols,
                       static_cast<const __nv_bfloat16*>(res_dev), out_cols, M, N, K, epi, gate_rows, ws, counters, t,
                       static_cast<cudaStream_t>(stream));
}

int b200_op_conv1d(const float* x_dev, const float* w_dev, const float* bias_dev, float* prev_dev,
                   const uint8_t* exec_mask_dev, float* y_dev, int B, int Cin, int Cout, int T, int K, int stride,
                   int dilation, int elu_in, void* stream) {
  using namespace b200::mimi;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int keff = (K - 1) * dilation + 1, P = keff - stride;
  if (T % stride || P < 0) B200_FAIL(B200_ERR_SHAPE, "op_conv1d: T %% stride != 0");
  float* wp = nullptr;
  ConvCommit* desc = nullptr;
  B200_CUDA(cudaMalloc(&wp, (size_t)Cout * Cin * K * 4));
  const long long n = (long long)Cout * Cin * K;
  B200_LAUNCH(pack_conv_w_kernel, (unsigned)ceil_div64(n, 256), 256, 0, st, w_dev, wp, Cout, Cin, K);
  ConvP p;
  p.x = x_dev; p.xb = (long long)Cin * T; p.xc = T; p.xt = 1; p.Tin = T;
  p.st = prev_dev; p.P = P; p.first = nullptr; p.w = wp; p.bias = bias_dev;
  p.y = y_dev; p.yb = (long long)Cout * (T / stride); p.yc = T / stride; p.yt = 1;
  p.res = nullptr; p.rb = p.rc = p.rt = 0;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.K = K; p.stride = stride; p.dil = dilation; p.Tout = T / stride; p.elu_in = elu_in;
  p.M = Cout; p.N = B * (T / stride); p.Kd = Cin * K; p.cin_aligned = (Cin % BK) == 0;
  auto kern = igemm_f32_kernel<ConvP, false>;
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM));
  B200_LAUNCH(kern, grid, 256, 0, st, p);
  if (P > 0) {
    ConvCommit cc;
    cc.x = x_dev; cc.xb = p.xb; cc.xc = p.xc; cc.xt = 1; cc.Tin = T; cc.st = prev_dev; cc.P = P; cc.Cin = Cin;
    cc.elu_in = elu_in; cc.first = nullptr;
    B200_CUDA(cudaMalloc(&desc, sizeof(ConvCommit)));
    B200_CUDA(cudaMemcpyAsync(desc, &cc, sizeof(cc), cudaMemcpyHostToDevice, st));
    dim3 g2(ceil_div(B * Cin, 128), 1);
    B200_LAUNCH(conv_commit_kernel, g2, 128, 0, st, desc, 1, exec_mask_dev, B);
  }
  B200_CUDA(cudaStreamSynchronize(st));
  cudaFree(wp);
  if (desc) cudaFree(desc);
  return check_launch("op_conv1d");
}

int b200_op_convtr1d(const float* x_dev, const float* w_dev, const float* bias_dev, float* partial_dev,
                     const uint8_t* exec_mask_dev, float* y_dev, int B, int Cin, int Cout, int T, int K, int stride,
                     int elu_in, void* stream) {
  using namespace b200::mimi;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (K != 2 * stride) B200_FAIL(B200_ERR_SHAPE, "op_convtr1d: kernel must be 2*stride");
  float *wp = nullptr, *scratch = nullptr;
  ConvTrCommit* desc = nullptr;
  const long long n = (long long)Cin * Cout * K;
  B200_CUDA(cudaMalloc(&wp, n * 4));
  B200_CUDA(cudaMalloc(&scratch, (size_t)B * Cout * stride * 4));
  B200_LAUNCH(pack_convtr_w_kernel, (unsigned)ceil_div64(n, 256), 256, 0, st, w_dev, wp, Cin, Cout, stride);
  ConvTrP p;
  p.x = x_dev; p.xb = (long long)Cin * T; p.xc = T; p.xt = 1; p.T = T;
  p.partial = partial_dev; p.scratch = scratch; p.w = wp; p.bias = bias_dev;
  p.y = y_dev; p.yb = (long long)Cout * T * stride; p.yc = (long long)T * stride; p.yt = 1;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.S = stride; p.elu_in = elu_in;
  p.M = Cout * stride; p.N = B * (T + 1); p.Kd = 2 * Cin; p.cin_aligned = (Cin % BK) == 0;
  auto kern = igemm_f32_kernel<ConvTrP, false>;
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM));
  B200_LAUNCH(kern, grid, 256, 0, st, p);
  ConvTrCommit tc_;
  tc_.partial = partial_dev; tc_.scratch = scratch; tc_.per_row = Cout * stride;
  B200_CUDA(cudaMalloc(&desc, sizeof(ConvTrCommit)));
  B200_CUDA(cudaMemcpyAsync(desc, &tc_, sizeof(tc_), cudaMemcpyHostToDevice, st));
  dim3 g2((unsigned)ceil_div64((long long)B * tc_.per_row, 256), 1);
  B200_LAUNCH(convtr_commit_kernel, g2, 256, 0, st, desc, exec_mask_dev, B);
  B200_CUDA(cudaStreamSynchronize(st));
  cudaFree(wp);
  cudaFree(scratch);
  cudaFree(desc);
  return check_launch("op_convtr1d");
}

int b200_op_attn_step(const void* qkv_dev, void* k_dev, void* v_dev, void* out_dev, const int64_t* pos_dev,
                      const uint8_t* exec_mask_dev, int B, int H, int cap, int nsplit, float max_period, void* stream) {
  using namespace b200::lm;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B < 1 || H < 1 || cap < 1) B200_FAIL(B200_ERR_SHAPE, "op_attn_step: bad shape");
  if (nsplit <= 0) nsplit = attn_pick_splits(B, H, cap);
  static float* part = nullptr;
  static int* counters = nullptr;
  static size_t part_n = 0, cnt_n = 0;
  const size_t need = (size_t)B * H * nsplit * (ATT_D + 2);
  if (need > part_n) {
    if (part) cudaFree(part);
    B200_CUDA(cudaMalloc(&part, need * sizeof(float)));
    part_n = need;
  }
  if ((size_t)B * H > cnt_n) {
    if (counters) cudaFree(counters);
    B200_CUDA(cudaMalloc(&counters, (size_t)B * H * sizeof(int)));
    B200_CUDA(cudaMemset(counters, 0, (size_t)B * H * sizeof(int)));
    cnt_n = (size_t)B * H;
  }
  AttnStep a;
  a.qkv = static_cast<const bf16*>(qkv_dev); a.kc = static_cast<bf16*>(k_dev); a.vc = static_cast<bf16*>(v_dev);
  a.out = static_cast<bf16*>(out_dev); a.part = part; a.counters = counters;
  a.pos = reinterpret_cast<const long long*>(pos_dev); a.exec_mask = exec_mask_dev; a.H = H; a.cap = cap; a.nsplit = nsplit;
  a.neg_log_period_2_over_d = -logf(max_period) * 2.f / (float)ATT_D;
  dim3 grid(B * H, nsplit);
  if (attn_group_keys() == 2) B200_LAUNCH(attn_step_kernel<2>, grid, ATT_THREADS, 0, st, a);
  else B200_LAUNCH(attn_step_kernel<4>, grid, ATT_THREADS, 0, st, a);
  return check_launch("op_attn_step");
}

int b200_op_attn_step_q8(const void* qkv_dev, void* k8_dev, void* v8_dev, float* ks_dev, float* vs_dev, void* out_dev,
                         const int64_t* pos_dev, const uint8_t* exec_mask_dev, int B, int H, int cap, int nsplit, float max_period,
                         int kv_dtype, void* stream) {
  using namespace b200::lm;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B < 1 || H < 1 || cap < 1) B200_FAIL(B200_ERR_SHAPE, "op_attn_step_q8: bad shape");
  if (kv_dtype != B200_KV_FP8_E4M3 && kv_dtype != B200_KV_INT8) B200_FAIL(B200_ERR_INVALID, "op_attn_step_q8: kv_dtype %d", kv_dtype);
  if (nsplit <= 0) nsplit = attn_pick_splits(B, H, cap);
  static float* part = nullptr;
  static int* counters = nullptr;
  static size_t part_n = 0, cnt_n = 0;
  const size_t need = (size_t)B * H * nsplit * (ATT_D + 2);
  if (need > part_n) {
    if (part) cudaFree(part);
    B200_CUDA(cudaMalloc(&part, need * sizeof(float)));
    part_n = need;
  }
  if ((size_t)B * H > cnt_n) {
    if (counters) cudaFree(counters);
    B200_CUDA(cudaMalloc(&counters, (size_t)B * H * sizeof(int)));
    B200_CUDA(cudaMemset(counters, 0, (size_t)B * H * sizeof(int)));
    cnt_n = (size_t)B * H;
  }
  AttnStepQ8 a;
  a.qkv = static_cast<const bf16*>(qkv_dev); a.kc = static_cast<uint8_t*>(k8_dev); a.vc = static_cast<uint8_t*>(v8_dev);
  a.ks = ks_dev; a.vs = vs_dev; a.out = static_cast<bf16*>(out_dev); a.part = part; a.counters = counters;
  a.pos = reinterpret_cast<const long long*>(pos_dev); a.exec_mask = exec_mask_dev; a.H = H; a.cap = cap; a.nsplit = nsplit;
  a.neg_log_period_2_over_d = -logf(max_period) * 2.f / (float)ATT_D;
  dim3 grid(B * H, nsplit);
  if (kv_dtype == B200_KV_INT8) B200_LAUNCH(attn_step_q8_kernel<KV_INT8>, grid, ATT_THREADS, 0, st, a);
  else B200_LAUNCH(attn_step_q8_kernel<KV_E4M3>, grid, ATT_THREADS, 0, st, a);
  return check_launch("op_attn_step_q8");
}

int b200_op_sample(const void* logits_bf16_dev, const float* noise_dev, int64_t* out_dev, int B, int card,
                   int use_sampling, float temp, int top_k, void* stream) {
  using namespace b200::lm;
  if (card + 1 > 65535 || top_k < 1 || top_k > SAMPLE_MAX_K) B200_FAIL(B200_ERR_SHAPE, "op_sample: bad card/top_k");
  const int k = top_k < card ? top_k : card;
  B200_LAUNCH(sample_kernel, B, SAMPLE_THREADS, 0, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(logits_bf16_dev),
              (long long)card, noise_dev, (long long)k, reinterpret_cast<long long*>(out_dev), card, use_sampling, temp, top_k);
  return check_launch("op_sample");
}

}  // extern "C"
