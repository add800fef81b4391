// Token-major helper kernels around mimi_tc_kernel: the two degenerate SEANet convolutions (Cin = 1 at the encoder's
// mouth, Cout = 1 at the decoder's), streaming-state commits / resets inside the extended input buffers, LayerNorm with a
// split (hi / lo) output, load-time weight re-layouts and the layout conversions of the debug taps.
#pragma once

#include "common.cuh"
#include "mimi_tc.cuh"

namespace b200 {
namespace mimi {

// lets a following mimi_tc_kernel (launched with programmatic stream serialization) start its prologue and weight prefetch
// while this kernel runs; that kernel waits (griddepcontrol.wait) before it touches activations
__device__ __forceinline__ void pdl_trigger_next() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// First encoder conv (Cin = 1, K <= 8; seanet.py:170-178): memory-bound, one thread per (b, t) produces all Cout channels.
//   in: ext[b][P + T] fp32 (carried samples | frame);  out: raw y[b][t][co] and the consumer's ELU(y) as a hi / lo pair
// ---------------------------------------------------------------------------------------------
struct ConvFirst {
  const float* ext; int E, P;                // [B][E]
  const float* w; const float* bias;         // [Cout][K], [Cout]
  float* y; long long y_sb;                  // [B][T][Cout]
  float *a_hi, *a_lo; long long a_sb; int a_elu;      // consumer ext (+ its carried rows), row stride Cout
  int B, Cout, K, T;
  int tok_per_block;          // tokens a CTA walks (a multiple of 128 / (Cout / 4))
};
static __global__ void __launch_bounds__(128) conv_first_tm_kernel(const ConvFirst p) {
  pdl_trigger_next();
  // A thread owns 4 consecutive channels of one token: the Cout / 4 threads of a token write its 3 x Cout floats as contiguous
  // 16-byte pieces (a warp stores 512 contiguous bytes per instruction; one thread per token wrote 32 scattered sectors per
  // instruction and ran at a sixth of the write bandwidth).  Weights of the 4 channels live in registers.
  const int tpt = p.Cout >> 2;                                   // threads per token (host: Cout % 4 == 0, 128 % tpt == 0)
  const int q = threadIdx.x % tpt, slot = threadIdx.x / tpt, tok_per_pass = 128 / tpt;
  float w[8][4], bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bias[j] = p.bias ? p.bias[4 * q + j] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k][j] = k < p.K ? p.w[(4 * q + j) * p.K + k] : 0.f;
  }
  const long long total = (long long)p.B * p.T;
  for (long long n = (long long)blockIdx.x * p.tok_per_block + slot; n < min(total, (long long)(blockIdx.x + 1) * p.tok_per_block); n += tok_per_pass) {
    const int b = (int)(n / p.T), t = (int)(n - (long long)b * p.T);
    float xin[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) xin[k] = k < p.K ? p.ext[(long long)b * p.E + t + k] : 0.f;     // cat(previous, x)[t + k]
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = bias[j];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < p.K) acc = fmaf(w[k][j], xin[k], acc);
      v[j] = acc;
    }
    const long long o = (long long)t * p.Cout + 4 * q;
    *reinterpret_cast<float4*>(p.y + (long long)b * p.y_sb + o) = make_float4(v[0], v[1], v[2], v[3]);
    float hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) mtc::split_tf32(p.a_elu ? elu1(v[j]) : v[j], hi[j], lo[j]);
    *reinterpret_cast<float4*>(p.a_hi + (long long)b * p.a_sb + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(p.a_lo + (long long)b * p.a_sb + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Last decoder conv (Cout = 1; seanet.py:380-388): one thread per (b, t) reduces over (tap, ci) from a plain fp32 ext
struct ConvLast {
  const float* ext; long long e_sb; int Cin;  // [B][P + T][Cin], activation already applied
  const float* w;                             // [K][Cin]
  const float* bias;
  float* y; long long y_sb;                   // [B][T]
  int B, K, dil, T;
};
static __global__ void __launch_bounds__(128) conv_last_tm_kernel(const ConvLast p) {
  extern __shared__ float sw[];               // [K * Cin]
  for (int i = threadIdx.x; i < p.K * p.Cin; i += blockDim.x) sw[i] = p.w[i];
  __syncthreads();
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= (long long)p.B * p.T) return;
  const int b = (int)(n / p.T), t = (int)(n - (long long)b * p.T);
  float acc = p.bias ? p.bias[0] : 0.f;
  for (int kw = 0; kw < p.K; ++kw) {
    const float4* e = reinterpret_cast<const float4*>(p.ext + (long long)b * p.e_sb + (long long)(t + kw * p.dil) * p.Cin);
    const float* wk = sw + kw * p.Cin;
    for (int c4 = 0; c4 < p.Cin / 4; ++c4) {
      const float4 v = e[c4];
      acc = fmaf(wk[4 * c4], v.x, acc); acc = fmaf(wk[4 * c4 + 1], v.y, acc);
      acc = fmaf(wk[4 * c4 + 2], v.z, acc); acc = fmaf(wk[4 * c4 + 3], v.w, acc);
    }
  }
  p.y[(long long)b * p.y_sb + t] = acc;
}

// ---------------------------------------------------------------------------------------------
// Streaming state inside the extended input buffers.  StreamingConv1d (conv.py:263-267): previous <- the last P rows of
// cat(previous, x) = rows [T, T + P) -> rows [0, P), for sessions with exec_mask.  StreamingConvTranspose1d's overlap-add carry
// (conv.py:349-361) is the same thing with P = 1 (the last input step).  Buffers are described once per direction.
// ---------------------------------------------------------------------------------------------
struct TmCommit { float* hi; float* lo; int P, T, rowlen; long long sb; };       // rowlen = Cin floats per row
static __global__ void tm_commit_kernel(const TmCommit* descs, const uint8_t* exec_mask, int B) {
  pdl_trigger_next();
  const TmCommit d = descs[blockIdx.y];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (b, ci)
  if (i >= (long long)B * d.rowlen) return;
  const int b = (int)(i / d.rowlen), ci = (int)(i - (long long)b * d.rowlen);
  if (!exec_mask[b]) return;
  float* h = d.hi + (long long)b * d.sb + ci;
  for (int j = 0; j < d.P; ++j) h[(long long)j * d.rowlen] = h[(long long)(d.T + j) * d.rowlen];      // ascending: reads stay ahead of writes
  if (d.lo != nullptr) {
    float* l = d.lo + (long long)b * d.sb + ci;
    for (int j = 0; j < d.P; ++j) l[(long long)j * d.rowlen] = l[(long long)(d.T + j) * d.rowlen];
  }
}
// reset (conv.py:166-169, 281-286): zero the carried rows of the sessions in mask (null = all)
static __global__ void tm_zero_kernel(float* hi, float* lo, int P, int rowlen, long long sb, const uint8_t* mask, int B) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)P * rowlen;
  if (i >= (long long)B * per) return;
  const int b = (int)(i / per);
  if (mask != nullptr && !mask[b]) return;
  const long long o = (long long)b * sb + (i - (long long)b * per);
  hi[o] = 0.f;
  if (lo != nullptr) lo[o] = 0.f;
}

// LayerNorm (eps 1e-5, affine; transformer.py:126) -> the following linear's input as a hi / lo pair
static __global__ void layernorm_split_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                                       float* __restrict__ y_hi, float* __restrict__ y_lo, int n_tok, int C, float eps) {
  pdl_trigger_next();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_tok) return;
  const float* xr = x + (long long)warp * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum(s) / C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { float d = xr[c] - mean; v += d * d; }
  const float rstd = rsqrtf(warp_sum(v) / C + eps);
  for (int c = lane; c < C; c += 32) {
    float hi, lo;
    mtc::split_tf32((xr[c] - mean) * rstd * g[c] + bta[c], hi, lo);
    y_hi[(long long)warp * C + c] = hi;
    y_lo[(long long)warp * C + c] = lo;
  }
}

// One attention step of a bottleneck transformer layer in ONE launch (transformer.py:557-597 for T tokens per frame):
// RoPE(q, k) (rope.py:45-82, interleaved pairs, fp32) -> ring append of the T new keys / values for executing sessions
// (RingKVCache.complete, transformer.py:236-288) -> attention of the T queries over the ring under the causal / context mask ->
// the out_proj's input as a hi / lo pair.  One CTA per (session, head), 128 threads; qkv [n_tok][3C] (rows q | k | v, each (h d)).
template <int D>
static __global__ void __launch_bounds__(128) ring_attn_step_kernel(const float* __restrict__ qkv, float* __restrict__ kc,
                                                             float* __restrict__ vc, float* __restrict__ out_hi,
                                                             float* __restrict__ out_lo, const long long* __restrict__ offset,
                                                             const uint8_t* __restrict__ exec_mask, int T, int H, int cap,
                                                             int context, float neg_log_period_2_over_d) {
  pdl_trigger_next();
  extern __shared__ float sm[];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int C = H * D;
  float* sq = sm;                 // [T][D] rotated queries
  float* sc = sm + T * D;         // [T][cap] scores, then probabilities
  const int tid = threadIdx.x;
  const long long off = offset[b];
  const bool exec = exec_mask[b] != 0;
  float* kb = kc + ((long long)b * H + h) * cap * D;
  float* vb = vc + ((long long)b * H + h) * cap * D;
  for (int i = tid; i < T * (D / 2); i += blockDim.x) {
    const int t = i / (D / 2), pr = i - t * (D / 2);
    const float pos = (float)off + (float)t;                         // rope.py:48
    const float freq = expf((float)pr * neg_log_period_2_over_d);    // rope.py:46
    float sn, cs;
    sincosf(freq * pos, &sn, &cs);
    const float* base = qkv + ((long long)b * T + t) * 3 * C + h * D + 2 * pr;
    const float qr = base[0], qi = base[1], kr = base[C], ki = base[C + 1];
    sq[t * D + 2 * pr] = qr * cs - qi * sn;
    sq[t * D + 2 * pr + 1] = qr * sn + qi * cs;
    if (exec) {                                                      // masked sessions do not write (their offsets do not move)
      const int slot = (int)((off + t) % cap);
      float* kd = kb + (long long)slot * D + 2 * pr;
      float* vd = vb + (long long)slot * D + 2 * pr;
      kd[0] = kr * cs - ki * sn; kd[1] = kr * sn + ki * cs;
      vd[0] = base[2 * C]; vd[1] = base[2 * C + 1];
    }
  }
  __syncthreads();                 // the new keys are read back below by other threads of this CTA (no other CTA touches this ring)
  // end_offset after the append: only advanced for executing sessions (transformer.py:279-284)
  const long long end_after = exec ? off + T : off;
  const long long last = off + T - 1;
  const int end_index = (int)(last % cap);
  const float scale = rsqrtf((float)D);
  for (int s = tid; s < cap; s += blockDim.x) {
    const int delta = s - end_index;
    long long pos = delta <= 0 ? last + delta : last + delta - cap;
    if (s >= end_after) pos = -1;
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
    if (pos >= 0) {
      const float4* kr = reinterpret_cast<const float4*>(kb + (long long)s * D);
#pragma unroll 4
      for (int d4 = 0; d4 < D / 4; ++d4) {
        const float4 kv = kr[d4];
        for (int t = 0; t < T; ++t) {
          const float* qq = sq + t * D + d4 * 4;
          dot[t] += kv.x * qq[0] + kv.y * qq[1] + kv.z * qq[2] + kv.w * qq[3];
        }
      }
    }
    for (int t = 0; t < T; ++t) {
      const long long dq = (off + t) - pos;
      const bool ok = pos >= 0 && dq >= 0 && dq < context;     // transformer.py:576-580
      sc[t * cap + s] = ok ? dot[t] * scale : -INFINITY;
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < T) {                  // softmax per query: warp t handles query t
    float* row = sc + warp * cap;
    float mx = -INFINITY;
    for (int s = lane; s < cap; s += 32) mx = fmaxf(mx, row[s]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < cap; s += 32) {
      const float e = (row[s] == -INFINITY) ? 0.f : expf(row[s] - mx);
      row[s] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    for (int s = lane; s < cap; s += 32) row[s] *= inv;
  }
  __syncthreads();
  for (int i = tid; i < T * D; i += blockDim.x) {
    const int t = i / D, d = i % D;
    const float* row = sc + t * cap;
    float acc = 0.f;
    for (int s = 0; s < cap; ++s) {
      const float pw = row[s];
      if (pw != 0.f) acc = fmaf(pw, vb[(long long)s * D + d], acc);
    }
    float hi, lo;
    mtc::split_tf32(acc, hi, lo);
    const long long o = ((long long)b * T + t) * C + h * D + d;
    out_hi[o] = hi;
    out_lo[o] = lo;
  }
}

// ---------------------------------------------------------------------------------------------
// load-time weight re-layouts to [N][tap][Cin] (the order mimi_tc's k-blocks walk)
//   conv   [Cout][Cin][K]  -> [Cout][K][Cin]
//   convtr [Cin][Cout][2S] -> [(r, co)][slot][Cin], slot 0 = x[t-1] (kernel taps S + r), slot 1 = x[t] (kernel taps r)
// ---------------------------------------------------------------------------------------------
static __global__ void tm_weight_conv_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int K) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * Cin * K) return;
  const int kw = (int)(i % K);
  const long long r = i / K;
  const int ci = (int)(r % Cin), co = (int)(r / Cin);
  out[((long long)co * K + kw) * Cin + ci] = w[i];
}
static __global__ void tm_weight_convtr_kernel(const float* __restrict__ w, float* __restrict__ out, int Cin, int Cout, int S) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cin * Cout * 2 * S) return;
  const int k = (int)(i % (2 * S));
  const long long r = i / (2 * S);
  const int co = (int)(r % Cout), ci = (int)(r / Cout);
  const int tap = k / S, ph = k - tap * S;            // kernel tap 0 multiplies x[t], tap 1 multiplies x[t-1]
  out[((long long)(ph * Cout + co) * 2 + (1 - tap)) * Cin + ci] = w[i];
}

// debug taps / reference layouts: y[b][c][t] = ytm[b][t][c]
static __global__ void tm_to_bct_kernel(const float* __restrict__ ytm, float* __restrict__ y, int B, int C, int T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * T) return;
  const int t = (int)(i % T);
  const long long r = i / T;
  const int c = (int)(r % C), b = (int)(r / C);
  y[i] = ytm[((long long)b * T + t) * C + c];
}

}  // namespace mimi
}  // namespace b200
