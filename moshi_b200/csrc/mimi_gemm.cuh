// fp32 implicit-GEMM kernels of the Mimi codec, second generation.
//
// Mimi at serving batch is FLOP-bound on the fp32 SIMT pipes (0.9 GFLOP per session-frame against
// ~20 KB of per-session HBM traffic), and the encode side must stay in exact fp32 FMA arithmetic
// (the RVQ indices downstream have to match the reference), so the design goal is FMA issue rate:
//
//   * 128x128 / 64x128 / 64x64 CTA tiles, 256 threads, 8x8 / 4x8 / 4x4 register micro-tiles:
//     16-64 FMAs per pair of 16-byte shared-memory reads.
//   * Weights are repacked once at load to k-major [Kd][M], so the A tile is a row of 16-byte
//     cp.async copies (no transposition, no register staging).
//   * The im2col operand is read from a per-layer *extended input* buffer ext[b][ci][state | frame]:
//     the producing layer's epilogue writes ELU(y) (the consumer's activation) straight behind the
//     consumer's carried left context, so the gather in the inner loop is one strided load without
//     ELU, concatenation branch or division (conv.py:261 `cat(previous, x)` becomes an address).
//     ConvTranspose reads a zero-padded copy xpad[b][ci][0 | frame | 0] the same way.
//   * StreamingConv1d state update (conv.py:263-267) is a shift inside ext, done by one commit kernel
//     per direction after the frame (rows with exec_mask only).
#pragma once

#include "common.cuh"

namespace b200 {
namespace mimi {

constexpr int GK = 16;                       // k per shared-memory stage
enum { G_CONV = 0, G_CONVTR = 1, G_LIN = 2 };
enum { GL_NONE = 0, GL_GELU = 1, GL_RES_SCALE = 2 };

struct GemmArgs {
  // operands
  const float* wk;                           // weights, k-major [Kd][M]
  const float* in; long long in_b; int in_E; // conv: ext; convtr: xpad; lin: x [N][in_E]
  int in_off;                                // conv: index of the first carried sample; convtr: index of input step 0
  int M, N, Kd;
  int Cin, stride, dil, T;                   // conv: T = Tout per row; convtr: T = input steps per row, stride = S
  // epilogue
  const float* bias;                         // conv [M]; convtr [Cout]
  float* y; long long yb, yc, yt;            // raw output (null = not stored)
  float* a; long long ab, ac, at; int a_elu; // activated copy for the consumer (null = none)
  const float* res; long long rb, rc, rt;    // conv: residual source; lin: residual [N][M]
  const float* partial; float* scratch;      // convtr overlap-add carry and its candidate (conv.py:349-361)
  const float* scale; int lin_epi;           // lin: layer-scale vector, GL_*
  int vec_y, vec_a, vec_r;                   // conv: 4 consecutive t may be moved as one float4
  int ksplit; float* ws;                     // split-K: blockIdx.z owns k-blocks [z*kchunk, ...); partials [z][M][N] in ws
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int sz = pred ? 16 : 0;              // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Epilogue of one output element (used by the split-K reduction; the main kernel has vectorised forms of the same).
template <int KIND>
__device__ __forceinline__ void gemm_store_one(const GemmArgs& p, int m, int n, float accv) {
  if (KIND == G_LIN) {
    float v = accv;
    if (p.lin_epi == GL_GELU) v = gelu_erf(v);
    else if (p.lin_epi == GL_RES_SCALE) v = p.res[(long long)n * p.yb + m] + p.scale[m] * v;
    p.y[(long long)n * p.yb + m] = v;
  } else if (KIND == G_CONV) {
    const int b = n / p.T, t = n - b * p.T;
    float v = accv + (p.bias ? p.bias[m] : 0.f);
    if (p.res) v += p.res[b * p.rb + m * p.rc + t * p.rt];
    if (p.y) p.y[b * p.yb + m * p.yc + t * p.yt] = v;
    if (p.a) p.a[b * p.ab + m * p.ac + t * p.at] = p.a_elu ? elu1(v) : v;
  } else {
    const int S = p.stride;
    const int b = n / (p.T + 1), t = n - b * (p.T + 1);
    const int co = m / S, r = m - co * S;
    const long long sidx = ((long long)b * (p.M / S) + co) * S + r;
    if (t == p.T) { p.scratch[sidx] = accv; return; }
    float v = accv + (p.bias ? p.bias[co] : 0.f);
    if (t == 0) v += p.partial[sidx];
    const long long to = (long long)t * S + r;
    if (p.y) p.y[b * p.yb + co * p.yc + to * p.yt] = v;
    if (p.a) p.a[b * p.ab + co * p.ac + to * p.at] = p.a_elu ? elu1(v) : v;
  }
}

// sums the split-K partials in split order (deterministic) and applies the epilogue
template <int KIND>
static __global__ void gemm_splitk_reduce_kernel(const GemmArgs p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.M * p.N) return;
  const int n = i % p.N, m = i / p.N;
  float v = 0.f;
  for (int z = 0; z < p.ksplit; ++z) v += p.ws[((long long)z * p.M + m) * p.N + n];
  gemm_store_one<KIND>(p, m, n, v);
}

// ConvTranspose epilogue for V consecutive GEMM rows m = co*S + r .. (one channel, V consecutive output samples)
template <int V>
__device__ __forceinline__ void convtr_store(const GemmArgs& p, int m, int b, int t, const float* accv) {
  if (m >= p.M) return;
  const int S = p.stride;
  const int co = m / S, r = m - co * S;
  const long long sidx = ((long long)b * (p.M / S) + co) * S + r;
  float v[V];
  if (t == p.T) {                                            // tail of y minus bias: the next frame's carry (conv.py:352-356)
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = accv[i];
    if (V == 4) *reinterpret_cast<float4*>(p.scratch + sidx) = make_float4(v[0], v[1], v[2], v[3]);
    else if (V == 2) *reinterpret_cast<float2*>(p.scratch + sidx) = make_float2(v[0], v[1]);
    else p.scratch[sidx] = v[0];
    return;
  }
  const float bias = p.bias ? p.bias[co] : 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = accv[i] + bias;
  if (t == 0) {                                              // y[..., :PT] += partial (conv.py:351)
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] += p.partial[sidx + i];
  }
  const long long to = (long long)t * S + r;
  if (p.y) {
    float* y = p.y + b * p.yb + co * p.yc + to * p.yt;
    if (V == 4 && p.yt == 1) *reinterpret_cast<float4*>(y) = make_float4(v[0], v[1], v[2], v[3]);
    else if (V == 2 && p.yt == 1) *reinterpret_cast<float2*>(y) = make_float2(v[0], v[1]);
    else {
#pragma unroll
      for (int i = 0; i < V; ++i) y[i * p.yt] = v[i];
    }
  }
  if (p.a) {
    if (p.a_elu) {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = elu1(v[i]);
    }
    float* a = p.a + b * p.ab + co * p.ac + to * p.at;
    if (V == 4 && p.at == 1) *reinterpret_cast<float4*>(a) = make_float4(v[0], v[1], v[2], v[3]);
    else if (V == 2 && p.at == 1) *reinterpret_cast<float2*>(a) = make_float2(v[0], v[1]);
    else {
#pragma unroll
      for (int i = 0; i < V; ++i) a[i * p.at] = v[i];
    }
  }
}

template <int BM, int BN, int KIND>
static __global__ void __launch_bounds__(256, (BM * BN >= 64 * 128) ? 2 : 3) mimi_gemm_kernel(const GemmArgs p) {
  constexpr int TM = BM / 16, TN = BN / 16;          // micro-tile, in groups of 4
  constexpr int GM = TM / 4, GN = TN / 4;
  constexpr int LDA = BM + 4, LDB = BN + 4;
  __shared__ __align__(16) float As[2][GK][LDA];
  __shared__ __align__(16) float Bs[2][GK][LDB];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  // ---- B loader set-up -----------------------------------------------------------------------
  constexpr int KG = 256 / BN;                       // k groups (gather form): 2 or 4
  constexpr int BPT = GK / KG;                       // elements per thread per stage: 8 or 4
  const int ln = tid % BN, lk = tid / BN;
  long long col_off = 0;
  bool col_ok = false;
  if (KIND != G_LIN) {
    const int n = n0 + ln;
    col_ok = n < p.N;
    if (col_ok) {
      if (KIND == G_CONV) {
        const int b = n / p.T, t = n - b * p.T;
        col_off = (long long)b * p.in_b + p.in_off + (long long)t * p.stride;
      } else {
        const int b = n / (p.T + 1), t = n - b * (p.T + 1);
        col_off = (long long)b * p.in_b + p.in_off + t;
      }
    }
  }
  const bool aligned = (p.Cin % GK) == 0;
  // lin: thread -> (token = id % BN, k quad = id / BN), BN * 4 quads per stage
  constexpr int LQ = (BN * 4) / 256;                 // quads per thread per stage: 1 or 2

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float rb[KIND == G_LIN ? 4 * LQ : BPT];

  auto load_a = [&](int buf, int k0) {
    constexpr int CH = GK * BM / 4;                  // 16-byte chunks per stage
#pragma unroll
    for (int i = 0; i < CH / 256; ++i) {
      const int id = tid + 256 * i;
      const int kk = id / (BM / 4), m4 = id % (BM / 4);
      const int m = m0 + 4 * m4, k = k0 + kk;
      const bool ok = m < p.M && k < p.Kd;           // M % 4 == 0 (host-checked)
      cp_async16(&As[buf][kk][4 * m4], p.wk + (ok ? (long long)k * p.M + m : 0), ok);
    }
    cp_async_commit();
  };
  auto fetch_b = [&](int k0) {
    if (KIND == G_LIN) {
#pragma unroll
      for (int i = 0; i < LQ; ++i) {
        const int id = tid + 256 * i;
        const int n = n0 + id % BN, kq = id / BN;
        const int k = k0 + 4 * kq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < p.N && k < p.Kd) v = *reinterpret_cast<const float4*>(p.in + (long long)n * p.in_E + k);   // Kd % 4 == 0
        rb[4 * i] = v.x; rb[4 * i + 1] = v.y; rb[4 * i + 2] = v.z; rb[4 * i + 3] = v.w;
      }
    } else {
      int kw0 = 0, ci0 = 0;
      if (aligned) { kw0 = k0 / p.Cin; ci0 = k0 - kw0 * p.Cin; }
#pragma unroll
      for (int i = 0; i < BPT; ++i) {
        const int kk = lk + KG * i, k = k0 + kk;
        float v = 0.f;
        if (col_ok && k < p.Kd) {
          int kw, ci;
          if (aligned) { kw = kw0; ci = ci0 + kk; }
          else { kw = k / p.Cin; ci = k - kw * p.Cin; }
          const long long off = col_off + (long long)ci * p.in_E + (KIND == G_CONV ? kw * p.dil : -kw);
          v = p.in[off];
        }
        rb[i] = v;
      }
    }
  };
  auto stash_b = [&](int buf) {
    if (KIND == G_LIN) {
#pragma unroll
      for (int i = 0; i < LQ; ++i) {
        const int id = tid + 256 * i;
        const int nn = id % BN, kq = id / BN;
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[buf][4 * kq + j][nn] = rb[4 * i + j];
      }
    } else {
#pragma unroll
      for (int i = 0; i < BPT; ++i) Bs[buf][lk + KG * i][ln] = rb[i];
    }
  };

  const int nk_all = (p.Kd + GK - 1) / GK;
  const int per = (nk_all + p.ksplit - 1) / p.ksplit;          // k-blocks per split (ksplit == 1: all of them)
  const int kb0 = blockIdx.z * per;
  const int nk = min(per, nk_all - kb0);                          // may be <= 0 for a trailing split: contributes zeros
  if (nk > 0) {
    load_a(0, kb0 * GK);
    fetch_b(kb0 * GK);
    stash_b(0);
    cp_async_wait_all();
  }
  __syncthreads();
  for (int it = 0; it < nk; ++it) {
    const int cur = it & 1;
    if (it + 1 < nk) {
      load_a(cur ^ 1, (kb0 + it + 1) * GK);
      fetch_b((kb0 + it + 1) * GK);
    }
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < GM; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4 + g * (BM / GM)]);
        a[4 * g] = v.x; a[4 * g + 1] = v.y; a[4 * g + 2] = v.z; a[4 * g + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < GN; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4 + g * (BN / GN)]);
        b[4 * g] = v.x; b[4 * g + 1] = v.y; b[4 * g + 2] = v.z; b[4 * g + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < nk) {
      stash_b(cur ^ 1);
      cp_async_wait_all();
    }
    __syncthreads();
  }

  // ---- epilogue ------------------------------------------------------------------------------
  // thread owns rows m = m0 + ty*4 + gi*(BM/GM) + ii and columns n = n0 + tx*4 + gj*(BN/GN) + jj
  if (p.ksplit > 1) {                                             // partial sums only; gemm_splitk_reduce_kernel finishes
    float* wz = p.ws + (long long)blockIdx.z * p.M * p.N;
#pragma unroll
    for (int gj = 0; gj < GN; ++gj)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int n = n0 + tx * 4 + gj * (BN / GN) + jj;
        if (n >= p.N) continue;
#pragma unroll
        for (int gi = 0; gi < GM; ++gi)
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int m = m0 + ty * 4 + gi * (BM / GM) + ii;
            if (m < p.M) wz[(long long)m * p.N + n] = acc[4 * gi + ii][4 * gj + jj];
          }
      }
    return;
  }
#pragma unroll
  for (int gj = 0; gj < GN; ++gj) {
    const int nb = n0 + tx * 4 + gj * (BN / GN);                 // first of 4 consecutive columns
    if (nb >= p.N) continue;
    if (KIND == G_LIN) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int n = nb + jj;
        if (n >= p.N) break;
#pragma unroll
        for (int gi = 0; gi < GM; ++gi) {
          const int mb = m0 + ty * 4 + gi * (BM / GM);
          if (mb >= p.M) continue;
          float v[4];
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) v[ii] = acc[4 * gi + ii][4 * gj + jj];
          if (p.lin_epi == GL_GELU) {
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) v[ii] = gelu_erf(v[ii]);
          } else if (p.lin_epi == GL_RES_SCALE) {                  // x + layer_scale(update)  (transformer.py:769,777)
            const float4 r = *reinterpret_cast<const float4*>(p.res + (long long)n * p.yb + mb);
            const float4 s = *reinterpret_cast<const float4*>(p.scale + mb);
            v[0] = r.x + s.x * v[0]; v[1] = r.y + s.y * v[1]; v[2] = r.z + s.z * v[2]; v[3] = r.w + s.w * v[3];
          }
          *reinterpret_cast<float4*>(p.y + (long long)n * p.yb + mb) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    } else if (KIND == G_CONV) {
      const int b = nb / p.T, t = nb - b * p.T;                    // vec paths: T % 4 == 0, so the 4 columns share b
#pragma unroll
      for (int gi = 0; gi < GM; ++gi) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int m = m0 + ty * 4 + gi * (BM / GM) + ii;
          if (m >= p.M) continue;
          const float bias = p.bias ? p.bias[m] : 0.f;
          float v[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) v[jj] = acc[4 * gi + ii][4 * gj + jj] + bias;
          const bool full = nb + 3 < p.N;
          if (p.res) {                                             // SEANetResnetBlock: u + v (seanet.py:90-93)
            if (p.vec_r && full) {
              const float4 r = *reinterpret_cast<const float4*>(p.res + b * p.rb + m * p.rc + t);
              v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
            } else {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int n = nb + jj;
                if (n < p.N) { const int b2 = n / p.T, t2 = n - b2 * p.T; v[jj] += p.res[b2 * p.rb + m * p.rc + t2 * p.rt]; }
              }
            }
          }
          if (p.y) {
            if (p.vec_y && full) *reinterpret_cast<float4*>(p.y + b * p.yb + m * p.yc + t) = make_float4(v[0], v[1], v[2], v[3]);
            else {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int n = nb + jj;
                if (n < p.N) { const int b2 = n / p.T, t2 = n - b2 * p.T; p.y[b2 * p.yb + m * p.yc + t2 * p.yt] = v[jj]; }
              }
            }
          }
          if (p.a) {
            if (p.a_elu) {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) v[jj] = elu1(v[jj]);
            }
            if (p.vec_a && full) *reinterpret_cast<float4*>(p.a + b * p.ab + m * p.ac + t) = make_float4(v[0], v[1], v[2], v[3]);
            else {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int n = nb + jj;
                if (n < p.N) { const int b2 = n / p.T, t2 = n - b2 * p.T; p.a[b2 * p.ab + m * p.ac + t2 * p.at] = v[jj]; }
              }
            }
          }
        }
      }
    } else {   // G_CONVTR: m = co*S + r, n = (b, t), t in [0, T]; t == T is the new carry
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int n = nb + jj;
        if (n >= p.N) break;
        const int b = n / (p.T + 1), t = n - b * (p.T + 1);
#pragma unroll
        for (int gi = 0; gi < GM; ++gi) {
          const int mb = m0 + ty * 4 + gi * (BM / GM);
          if (mb >= p.M) continue;
          float v[4];
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) v[ii] = acc[4 * gi + ii][4 * gj + jj];
          // vec_y = how many consecutive m (= consecutive output samples of one channel) go out as one vector: S % vec_y == 0
          if (p.vec_y == 4) convtr_store<4>(p, mb, b, t, v);
          else if (p.vec_y == 2) { convtr_store<2>(p, mb, b, t, v); convtr_store<2>(p, mb + 2, b, t, v + 2); }
          else {
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) convtr_store<1>(p, mb + ii, b, t, v + ii);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The two degenerate SEANet convolutions, which are memory-bound and do not fit a GEMM tile:
//   conv_cin1_kernel  : first encoder conv (Cin = 1, K = 7): one thread per (b, t) produces all Cout channels
//   conv_cout1_kernel : last decoder conv (Cout = 1, K = 3): one thread per (b, t) reduces over (ci, kw) from ext
// ---------------------------------------------------------------------------------------------
struct ConvCin1 {
  const float* x; long long xb, xt;          // input [B][1][T]
  const float* st; int P;                    // carried left context [B][P]
  const float* w;                            // [Cout][K]
  const float* bias;
  float* y; long long yb, yc;                // raw [B][Cout][T]
  float* a; long long ab, ac; int a_elu;     // activated copy for the consumer (may be null)
  int B, Cout, K, T;
};
static __global__ void __launch_bounds__(256) conv_cin1_kernel(const ConvCin1 p) {
  extern __shared__ float sw[];              // [Cout][K] then bias [Cout]
  for (int i = threadIdx.x; i < p.Cout * p.K; i += blockDim.x) sw[i] = p.w[i];
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) sw[p.Cout * p.K + i] = p.bias ? p.bias[i] : 0.f;
  __syncthreads();
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= (long long)p.B * p.T) return;
  const int b = n / p.T, t = n - (long long)b * p.T;
  float xin[8];                              // K <= 8
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int j = t + k;                     // index into cat(previous, x)
    float v = 0.f;
    if (k < p.K) v = j < p.P ? p.st[(long long)b * p.P + j] : p.x[b * p.xb + (long long)(j - p.P) * p.xt];
    xin[k] = v;
  }
  for (int co = 0; co < p.Cout; ++co) {
    float acc = sw[p.Cout * p.K + co];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < p.K) acc = fmaf(sw[co * p.K + k], xin[k], acc);
    p.y[b * p.yb + co * p.yc + t] = acc;
    if (p.a) p.a[b * p.ab + co * p.ac + t] = p.a_elu ? elu1(acc) : acc;
  }
}

struct ConvCout1 {
  const float* ext; long long eb; int E, off; // ext[b*eb + ci*E + off + t + kw*dil]  (activation already applied)
  const float* wk;                            // [K*Cin] (tap-major)
  const float* bias;
  float* y; long long yb, yt;                 // [B][1][T] via strides
  int B, Cin, K, dil, T;
};
static __global__ void __launch_bounds__(256) conv_cout1_kernel(const ConvCout1 p) {
  extern __shared__ float sw[];               // [K*Cin]
  for (int i = threadIdx.x; i < p.K * p.Cin; i += blockDim.x) sw[i] = p.wk[i];
  __syncthreads();
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= (long long)p.B * p.T) return;
  const int b = n / p.T, t = n - (long long)b * p.T;
  const float* e = p.ext + b * p.eb + p.off + t;
  float acc = p.bias ? p.bias[0] : 0.f;
  for (int kw = 0; kw < p.K; ++kw) {
    const float* ek = e + kw * p.dil;
#pragma unroll 8
    for (int ci = 0; ci < p.Cin; ++ci) acc = fmaf(sw[kw * p.Cin + ci], ek[(long long)ci * p.E], acc);
  }
  p.y[b * p.yb + (long long)t * p.yt] = acc;
}

// ext / xpad filler for layers whose producer is not one of the kernels above (PCM frame, transformer output):
// dst[b*db + c*dc + t*dt] = act(src[b*sb + c*sc + t*st])
static __global__ void fill_act_kernel(const float* __restrict__ src, long long sb, long long sc, long long st,
                                float* __restrict__ dst, long long db, long long dc, long long dt, int B, int C, int T,
                                int elu) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * T) return;
  const int t = i % T;
  const long long r = i / T;
  const int c = r % C, b = r / C;
  const float v = src[b * sb + c * sc + t * st];
  dst[b * db + c * dc + t * dt] = elu ? elu1(v) : v;
}

// StreamingConv1d state update inside ext (conv.py:263-267): the P samples that precede the next frame are the
// last P of cat(previous, x) = ext[T .. T+P)  ->  ext[0 .. P), for rows with exec_mask.
struct ExtCommit { float* ext; int P, T, E, Cin; };
static __global__ void ext_commit_kernel(const ExtCommit* descs, const uint8_t* exec_mask, int B) {
  const ExtCommit d = descs[blockIdx.y];
  const int row = blockIdx.x * blockDim.x + threadIdx.x;     // (b, ci)
  if (row >= B * d.Cin) return;
  if (!exec_mask[row / d.Cin]) return;
  float* e = d.ext + (long long)row * d.E;
  for (int j = 0; j < d.P; ++j) e[j] = e[d.T + j];           // ascending: reads stay ahead of writes
}
// reset: zero the carried samples of the rows in mask (conv.py:166-169)
static __global__ void ext_zero_kernel(float* ext, int P, int E, int Cin, const uint8_t* mask, int B) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * Cin * P) return;
  const int j = i % P;
  const long long row = i / P;
  if (mask == nullptr || mask[row / Cin]) ext[row * E + j] = 0.f;
}

// load-time repacking to k-major
static __global__ void pack_conv_k_kernel(const float* w /*[Cout][Cin][K]*/, float* out /*[K*Cin][Cout]*/, int Cout, int Cin, int K) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * Cin * K) return;
  const int kw = i % K; long long r = i / K;
  const int ci = r % Cin; const int co = r / Cin;
  out[((long long)kw * Cin + ci) * Cout + co] = w[i];
}
static __global__ void pack_convtr_k_kernel(const float* w /*[Cin][Cout][2S]*/, float* out /*[2*Cin][Cout*S]*/, int Cin, int Cout, int S) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cin * Cout * 2 * S) return;
  const int k = i % (2 * S); long long r = i / (2 * S);
  const int co = r % Cout; const int ci = r / Cout;
  const int tap = k / S, ph = k % S;
  out[((long long)tap * Cin + ci) * ((long long)Cout * S) + (long long)co * S + ph] = w[i];
}

}  // namespace mimi
}  // namespace b200
