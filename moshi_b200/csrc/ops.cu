// Kernel-level C entry points used by the parity tests (same kernels the handles launch).
#include "gemm_sk.cuh"
#include "lm_kernels.cuh"
#include "mimi_kernels.cuh"
#include "mimi_tc.cuh"

using namespace b200;

extern "C" {

// split-K workspace + arrival counters of the kernel-level entry points (a handle owns its own); separate sets for the bf16 and
// the int8 kernels, like the handles keep them
static int sk_scratch(float** ws, int** counters, int which = 0) {
  static float* g_ws[2] = {nullptr, nullptr};
  static int* g_counters[2] = {nullptr, nullptr};
  if (!g_ws[which]) {
    B200_CUDA(cudaMalloc(&g_ws[which], tc::sk_workspace_bytes(256)));
    B200_CUDA(cudaMalloc(&g_counters[which], tc::SK_MAX_TILES * sizeof(int)));
    B200_CUDA(cudaMemset(g_counters[which], 0, tc::SK_MAX_TILES * sizeof(int)));
  }
  *ws = g_ws[which]; *counters = g_counters[which];
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
  t.ns = smem_budget == -1 ? -1 : 0;   // smem_budget -1: the swap-AB kernels at any M (comparison row of tools/kbench.py)
  if (smem_budget < 0) t.smem_budget = 0;
  const int out_cols = epi == 2 ? gate_rows : N;
  return tc::sk_linear(cache, static_cast<const __nv_bfloat16*>(x_dev), K, w_tiles_dev, static_cast<__nv_bfloat16*>(y_dev),
                       out_cols, static_cast<const __nv_bfloat16*>(res_dev), out_cols, M, N, K, epi, gate_rows, ws, counters, t,
                       static_cast<cudaStream_t>(stream));
}

// the non-swapped kernel (33..256 sessions in the LM) at any M <= 256; cluster = K-splits (0 = the LM's own choice)
int b200_op_linear_ns(const void* x_dev, const void* w_tiles_dev, void* y_dev, const void* res_dev, int M, int N, int K, int epi,
                      int gate_rows, int cluster, void* stream) {
  if (!x_dev || !w_tiles_dev || !y_dev) B200_FAIL(B200_ERR_INVALID, "op_linear_ns: null pointer");
  if (!tc::ns_supported(M, N, K, epi)) B200_FAIL(B200_ERR_SHAPE, "op_linear_ns: unsupported shape");
  static tc::GemmPlanCache cache;
  const int out_cols = epi == 2 ? gate_rows : N;
  // cluster >= 100: single-tile units (N = 128 per instruction) with cluster - 100 K-splits
  const int unit_tiles = cluster >= 100 ? 1 : (cluster > 0 ? 2 : 0);
  if (cluster >= 100) cluster -= 100;
  return tc::ns_linear(cache, static_cast<const __nv_bfloat16*>(x_dev), K, w_tiles_dev, static_cast<__nv_bfloat16*>(y_dev), out_cols,
                       static_cast<const __nv_bfloat16*>(res_dev), out_cols, M, N, K, epi, gate_rows, cluster, 0,
                       static_cast<cudaStream_t>(stream), unit_tiles);
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
  B200_TRY(sk_scratch(&ws, &counters, 1));   // int32 partials of cut tiles
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

}  // extern "C"

// ---- mimi_tc_kernel on the reference's [B, C, T] layout (test scaffolding: the handle keeps everything token-major) --------
namespace {

// ext[b][j][ci] (hi / lo) = cat(previous, act(x))[b][ci][j]  for j < P + T
__global__ void tm_pack_in_kernel(const float* __restrict__ x, const float* __restrict__ prev, float* __restrict__ hi, float* __restrict__ lo,
                                  int B, int Cin, int T, int P, int elu) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long E = P + T;
  if (i >= (long long)B * E * Cin) return;
  const int ci = (int)(i % Cin);
  const long long r = i / Cin;
  const int j = (int)(r % E), b = (int)(r / E);
  float v;
  if (j < P) v = prev ? prev[((long long)b * Cin + ci) * P + j] : 0.f;
  else { v = x[((long long)b * Cin + ci) * T + (j - P)]; v = elu ? elu1(v) : v; }
  float h, l;
  b200::mtc::split_tf32(v, h, l);
  hi[i] = h; lo[i] = l;
}
// y[b][c][t] = ytm[b][t][c]
__global__ void tm_unpack_out_kernel(const float* __restrict__ ytm, float* __restrict__ y, int B, int C, int T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * T) return;
  const int t = (int)(i % T);
  const long long r = i / T;
  const int c = (int)(r % C), b = (int)(r / C);
  y[i] = ytm[((long long)b * T + t) * C + c];
}
// previous[b][ci][j] <- (hi + lo)[b][T + j][ci] for rows with exec_mask (conv.py:263-267)
__global__ void tm_commit_kernel(const float* __restrict__ hi, const float* __restrict__ lo, float* __restrict__ prev,
                                 const uint8_t* __restrict__ exec_mask, int B, int Cin, int T, int P) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * Cin * P) return;
  const int j = (int)(i % P);
  const long long r = i / P;
  const int ci = (int)(r % Cin), b = (int)(r / Cin);
  if (exec_mask && !exec_mask[b]) return;
  const long long o = ((long long)b * (P + T) + T + j) * Cin + ci;
  prev[i] = hi[o] + lo[o];
}
// conv [Cout][Cin][K] -> [Cout][K][Cin];  convtr [Cin][Cout][2S] -> [(r, co)][tap'][ci] with tap' = 0: x[t-1] (kernel taps S + r), 1: x[t]
__global__ void tm_weight_conv_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int K) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * Cin * K) return;
  const int kw = (int)(i % K);
  const long long r = i / K;
  const int ci = (int)(r % Cin), co = (int)(r / Cin);
  out[((long long)co * K + kw) * Cin + ci] = w[i];
}
__global__ void tm_weight_convtr_kernel(const float* __restrict__ w, float* __restrict__ out, int Cin, int Cout, int S) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cin * Cout * 2 * S) return;
  const int k = (int)(i % (2 * S));
  const long long r = i / (2 * S);
  const int co = (int)(r % Cout), ci = (int)(r / Cout);
  const int tap = k / S, ph = k - tap * S;            // tap 0 multiplies x[t], tap 1 multiplies x[t-1]
  const int n = ph * Cout + co, slot = 1 - tap;       // rows of the A operand: slot 0 = x[t-1], slot 1 = x[t]
  out[((long long)n * 2 + slot) * Cin + ci] = w[i];
}

struct DevBufs {
  std::vector<void*> p;
  ~DevBufs() { for (void* q : p) cudaFree(q); }
  template <class T> int get(T** out, size_t n) {
    void* q = nullptr;
    B200_CUDA(cudaMalloc(&q, n ? n : 1));
    p.push_back(q);
    *out = static_cast<T*>(q);
    return B200_OK;
  }
};

int tc_tile_shape(int T, int* tt, int* bb) {
  for (int c = 128; c >= 1; c >>= 1)
    if (T % c == 0) { *tt = c; *bb = 128 / c; return B200_OK; }
  return B200_ERR_SHAPE;
}

}  // namespace

extern "C" {

int b200_op_tc_linear_f32(const float* x_dev, const float* w_dev, float* y_dev, int M, int N, int K, void* stream) {
  using namespace b200::mtc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!x_dev || !w_dev || !y_dev || M < 1 || K % 32 || N % 16) B200_FAIL(B200_ERR_SHAPE, "op_tc_linear_f32: bad shape");
  B200_TRY(tc_init());
  DevBufs d;
  float *hi, *lo, *ws; uint8_t* wt;
  B200_TRY(d.get(&hi, (size_t)M * K * 4));
  B200_TRY(d.get(&lo, (size_t)M * K * 4));
  B200_TRY(d.get(&wt, tc_packed_bytes(N, K, 1)));
  const size_t ws_bytes = (size_t)64 << 20;
  B200_TRY(d.get(&ws, ws_bytes));
  const long long n = (long long)M * K;
  // x is already [M][K] token-major: a flat hi/lo split is tm_pack_in with B = M "sessions" of Cin = K, T = 1, P = 0
  B200_LAUNCH(tm_pack_in_kernel, (unsigned)ceil_div64(n, 256), 256, 0, st, x_dev, (const float*)nullptr, hi, lo, M, K, 1, 0, 0);
  B200_TRY(tc_pack_weights(w_dev, wt, N, K, 1, st));
  TcLayer L;
  L.kind = 2; L.Cin = K; L.N = N; L.n_taps = 1; L.wt = wt;
  L.NT = N >= TC_MAX_NT ? TC_MAX_NT : N; L.n_tiles_n = (N + L.NT - 1) / L.NT; L.num_kb = K / TC_KB;
  memset(&L.p, 0, sizeof(L.p));
  L.p.tt = 128; L.p.bb = 1; L.p.row0 = 0;
  L.p.epi = TC_EPI_CONV; L.p.y = y_dev; L.p.y_sb = 0; L.p.y_row = N; L.p.ws = ws;
  B200_TRY(tc_make_map(&L.map_hi, hi, K, M, (long long)M * K, 1, 128, 1, 1));
  B200_TRY(tc_make_map(&L.map_lo, lo, K, M, (long long)M * K, 1, 128, 1, 1));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  B200_TRY(tc_plan(L, 1, M, sms, ws_bytes));
  B200_TRY(tc_launch(L, st));
  B200_CUDA(cudaStreamSynchronize(st));
  return B200_OK;
}

int b200_op_tc_conv1d(const float* x_dev, const float* w_dev, const float* bias_dev, float* prev_dev, const uint8_t* exec_mask_dev,
                      float* y_dev, int B, int Cin, int Cout, int T, int K, int stride, int dilation, int elu_in, int transposed,
                      void* stream) {
  using namespace b200::mtc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_TRY(tc_init());
  // transposed: ConvTranspose1d with K == 2 * stride as two taps over [x[t-1], x[t]] (P = 1 carried input row, N = S * Cout)
  const int n_taps = transposed ? 2 : K, S = stride;
  const int P = transposed ? 1 : (K - 1) * dilation + 1 - stride;
  const int Tout = transposed ? T : T / stride;               // GEMM rows per session
  const int N = transposed ? S * Cout : Cout;
  if (Cin % 32 || N % 16 || (N > 128 && N % 128) || P < 0 || (!transposed && T % stride) || (transposed && K != 2 * S))
    B200_FAIL(B200_ERR_SHAPE, "op_tc_conv1d: unsupported shape");
  DevBufs d;
  float *hi, *lo, *ws, *wk, *ytm; uint8_t* wt;
  const size_t E = (size_t)P + T;
  B200_TRY(d.get(&hi, (size_t)B * E * Cin * 4));
  B200_TRY(d.get(&lo, (size_t)B * E * Cin * 4));
  B200_TRY(d.get(&wk, (size_t)N * n_taps * Cin * 4));
  B200_TRY(d.get(&wt, tc_packed_bytes(N, Cin, n_taps)));
  B200_TRY(d.get(&ytm, (size_t)B * Tout * N * 4));
  const size_t ws_bytes = (size_t)64 << 20;
  B200_TRY(d.get(&ws, ws_bytes));
  const long long n_in = (long long)B * E * Cin, n_w = (long long)N * n_taps * Cin;
  B200_LAUNCH(tm_pack_in_kernel, (unsigned)ceil_div64(n_in, 256), 256, 0, st, x_dev, (const float*)prev_dev, hi, lo, B, Cin, T, P, elu_in);
  if (transposed) B200_LAUNCH(tm_weight_convtr_kernel, (unsigned)ceil_div64(n_w, 256), 256, 0, st, w_dev, wk, Cin, Cout, S);
  else B200_LAUNCH(tm_weight_conv_kernel, (unsigned)ceil_div64(n_w, 256), 256, 0, st, w_dev, wk, Cout, Cin, K);
  B200_TRY(tc_pack_weights(wk, wt, N, Cin, n_taps, st));
  TcLayer L;
  L.kind = transposed ? 1 : 0; L.Cin = Cin; L.N = N; L.n_taps = n_taps; L.dil = transposed ? 1 : dilation; L.stride = transposed ? 1 : stride;
  L.wt = wt; L.bias = bias_dev; L.bias_mod = Cout;
  L.NT = N >= TC_MAX_NT ? TC_MAX_NT : N; L.n_tiles_n = (N + L.NT - 1) / L.NT; L.num_kb = n_taps * (Cin / TC_KB);
  memset(&L.p, 0, sizeof(L.p));
  if (tc_tile_shape(Tout, &L.p.tt, &L.p.bb) != B200_OK) B200_FAIL(B200_ERR_SHAPE, "op_tc_conv1d: T");
  while (L.p.tt * L.stride > 256) { L.p.tt /= 2; L.p.bb *= 2; }
  L.p.row0 = 0;
  L.p.epi = TC_EPI_CONV; L.p.y = ytm; L.p.y_sb = (long long)Tout * N; L.p.y_row = N; L.p.ws = ws;
  B200_TRY(tc_make_map(&L.map_hi, hi, Cin, (long long)E, (long long)E * Cin, B, L.p.tt, L.p.bb, L.stride));
  B200_TRY(tc_make_map(&L.map_lo, lo, Cin, (long long)E, (long long)E * Cin, B, L.p.tt, L.p.bb, L.stride));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  B200_TRY(tc_plan(L, B, Tout, sms, ws_bytes));
  B200_TRY(tc_launch(L, st));
  // token-major [B][Tout][N] -> the reference layout: conv [B][Cout][Tout]; convtr [B][Cout][Tout * S] (row t holds S steps of Cout)
  const int Tr = transposed ? Tout * S : Tout;
  B200_LAUNCH(tm_unpack_out_kernel, (unsigned)ceil_div64((long long)B * Cout * Tr, 256), 256, 0, st, ytm, y_dev, B, Cout, Tr);
  if (P > 0 && prev_dev)
    B200_LAUNCH(tm_commit_kernel, (unsigned)ceil_div64((long long)B * Cin * P, 256), 256, 0, st, hi, lo, prev_dev, exec_mask_dev, B, Cin, T, P);
  B200_CUDA(cudaStreamSynchronize(st));
  return check_launch("op_tc_conv1d");
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
