// bf16 kernels of the Moshi LM decode step (one 12.5 Hz frame for B concurrent sessions).
//
// Cast points follow the reference exactly: every nn.Linear output is a bf16
// tensor, RMSNorm / RoPE / softmax / SiLU are evaluated in fp32 on bf16 inputs and rounded once,
// residual adds and the 17-way embedding sum are bf16 adds.
#pragma once

#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"

namespace b200 {
namespace lm {

typedef __nv_bfloat16 bf16;

// The GEMM that follows each of these kernels is launched with programmatic stream serialization
// (gemm_sk.cu): triggering at entry lets its CTAs take free SM resources and start streaming weights
// while this kernel runs; the GEMM itself waits (griddepcontrol.wait) before it touches activations.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// ---------------------------------------------------------------------------------------------
// token ring (LMGen._step bookkeeping, lm.py:691-702 and :759-783)
// ---------------------------------------------------------------------------------------------
struct TokenRing {
  long long* cache;        // [B][Kc][CT]
  long long* offsets;      // [B]
  const uint8_t* exec_mask;
  const int* delays;       // [Kc]
  int Kc, CT, dep_q, n_q, card, text_card, max_delay;
  // classifier-free guidance (lm.py:596-604, 714-726): the model runs on 2B rows, rows [B, 2B) are the "null" copies
  int cfg;                          // 0 = off, 1 = on
  int cfg_is_no_text;               // null rows: text stream zeroed
  const long long* cfg_masked_until;   // [B] or null: null rows: every stream zeroed while offset <= delay + masked_until[b]
  int* err;                         // device error flags (ERR_*), may be null
};

// device error flags: raised instead of reading out of bounds; surfaced by b200_lm_error_flags / b200_mimi_error_flags and by
// the host-synchronising entry points
enum { ERR_TOKEN_RANGE = 1, ERR_CODE_RANGE = 2, ERR_KV_CAPACITY = 4 };

// One thread per (b, k): write the user's codes, then build the model input row(s) [MB][Kc].
static __global__ void lm_prepare_kernel(const TokenRing r, const long long* __restrict__ in_codes, int n_in,
                                  long long* __restrict__ input_tokens, int B) {
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * r.Kc) return;
  const int b = i / r.Kc, k = i % r.Kc;
  const long long off = r.offsets[b];
  const bool exec = r.exec_mask[b] != 0;
  long long* row = r.cache + ((long long)b * r.Kc + k) * r.CT;
  if (k > r.dep_q && exec) {                                    // lm.py:691-696
    const int pos = (int)((off + r.delays[k]) % r.CT);
    row[pos] = in_codes[(long long)b * n_in + (k - r.dep_q - 1)];
  }
  const bool is_init = off <= r.delays[k] || !exec;             // lm.py:698-699
  const long long initial = k == 0 ? r.text_card : r.card;      // lm.py:297-311
  const long long tok = is_init ? initial : row[off % r.CT];
  input_tokens[i] = tok;
  if (r.cfg) {                                                  // lm.py:714-726
    long long t2 = tok;
    if (r.cfg_masked_until != nullptr && off <= r.delays[k] + r.cfg_masked_until[b] && !is_init) t2 = -1;
    if (r.cfg_is_no_text && k == 0 && !is_init) t2 = -1;
    input_tokens[(long long)B * r.Kc + i] = t2;
  }
}

// offsets += exec; store sampled tokens; gather the delay-aligned output (lm.py:759-783)
static __global__ void lm_finish_kernel(const TokenRing r, const long long* __restrict__ text_token,
                                 const long long* __restrict__ audio_tokens /*[dep_q][B]*/,
                                 long long* __restrict__ out /*[B][dep_q+1]*/, int B, unsigned long long* noise_ctr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0 && noise_ctr != nullptr) *noise_ctr += 1;        // the next step draws fresh Exp(1) noise (lm_noise_kernel)
  if (b >= B) return;
  const bool exec = r.exec_mask[b] != 0;
  long long off = r.offsets[b];
  if (exec) off += 1;
  r.offsets[b] = off;
  const int pos = (int)(off % r.CT);
  if (exec) {
    r.cache[((long long)b * r.Kc + 0) * r.CT + pos] = text_token[b];
    for (int k = 0; k < r.dep_q; ++k)
      r.cache[((long long)b * r.Kc + 1 + k) * r.CT + pos] = audio_tokens[(long long)k * B + b];
  }
  const bool not_ready = off <= r.max_delay || !exec;
  for (int k = 0; k <= r.dep_q; ++k) {
    long long idx = (off - r.max_delay + r.delays[k]) % r.CT;
    if (idx < 0) idx += r.CT;                                    // python modulo
    const long long v = r.cache[((long long)b * r.Kc + k) * r.CT + idx];
    out[(long long)b * (r.dep_q + 1) + k] = not_ready ? -2 : v;
  }
}

// ---------------------------------------------------------------------------------------------
// embeddings (lm.py:390-397; ScaledEmbedding: token -1 -> zero row, lm_utils.py:103-121)
// ---------------------------------------------------------------------------------------------
struct EmbedTables {
  const bf16* audio[32];   // [card+1][dim]
  const bf16* text;        // [text_card+1][dim]
  int n_q, card, text_card;
  const bf16* condition_sum;   // [rows][dim] or null: fuser.get_sum(condition_tensors), added last (lm.py:398-399)
  int* err;
};
// valid ids: [0, vocab] (vocab = the initial token) and -1 (zero row); anything else (-2 "ungenerated", ids past the table)
// would be a device assert in the reference (F.embedding); here it raises ERR_TOKEN_RANGE and embeds as the zero row
__device__ __forceinline__ bool embed_id_ok(long long id, int vocab, int* err) {
  if (id >= 0 && id <= vocab) return true;
  if (id != -1 && err != nullptr) atomicOr(err, ERR_TOKEN_RANGE);
  return false;
}
static __global__ void lm_embed_sum_kernel(const EmbedTables t, const long long* __restrict__ tokens /*[B][n_q+1]*/,
                                    bf16* __restrict__ x /*[B][dim]*/, int B, int dim) {
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (b, pair)
  const int half = dim / 2;
  if (i >= B * half) return;
  const int b = i / half, c = (i % half) * 2;
  const long long* tok = tokens + (long long)b * (t.n_q + 1);
  float s0 = 0.f, s1 = 0.f;
  for (int k = 0; k < t.n_q; ++k) {
    const long long id = tok[k + 1];
    float e0 = 0.f, e1 = 0.f;
    if (embed_id_ok(id, t.card, t.err)) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(t.audio[k] + id * dim + c);
      e0 = __low2float(v); e1 = __high2float(v);
    }
    if (k == 0) { s0 = e0; s1 = e1; }
    else { s0 = rbf(s0 + e0); s1 = rbf(s1 + e1); }       // bf16 add per table
  }
  {
    const long long id = tok[0];
    float e0 = 0.f, e1 = 0.f;
    if (embed_id_ok(id, t.text_card, t.err)) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(t.text + id * dim + c);
      e0 = __low2float(v); e1 = __high2float(v);
    }
    if (t.n_q == 0) { s0 = e0; s1 = e1; }
    else { s0 = rbf(s0 + e0); s1 = rbf(s1 + e1); }
  }
  if (t.condition_sum != nullptr) {
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(t.condition_sum + (long long)b * dim + c);
    s0 = rbf(s0 + __low2float(v)); s1 = rbf(s1 + __high2float(v));
  }
  *reinterpret_cast<__nv_bfloat162*>(x + (long long)b * dim + c) = __floats2bfloat162_rn(s0, s1);
}

// depformer input: x = depformer_in[k](transformer_out) + emb(prev)   (lm.py:475-486); with CFG the rows [Bt, 2*Bt) reuse the
// tokens of rows [0, Bt) (lm.py:824-826)
static __global__ void dep_input_kernel(const bf16* __restrict__ din /*[B][ld]*/, long long ld, int col0,
                                 const bf16* __restrict__ table /*[V][dd]*/, const long long* __restrict__ prev /*[Bt]*/,
                                 bf16* __restrict__ x /*[B][dd]*/, int B, int dd, int Bt, int vocab, int* err) {
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * dd) return;
  const int b = i / dd, c = i % dd;
  const long long id = prev[b % Bt];
  const float e = embed_id_ok(id, vocab, err) ? bf2f(table[id * dd + c]) : 0.f;
  x[i] = f2bf(bf2f(din[(long long)b * ld + col0 + c]) + e);
}

// classifier-free guidance on logits (lm.py:728-732, 828-833): out = null + (cond - null) * coef, every operation a bf16
// tensor op like the reference's; rows [0, B) = conditioned, [B, 2B) = null
static __global__ void cfg_combine_kernel(const bf16* __restrict__ logits /*[2B][card]*/, bf16* __restrict__ out /*[B][card]*/,
                                   int B, int card, float coef) {
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * card) return;
  const float l = bf2f(logits[i]), n = bf2f(logits[(long long)B * card + i]);
  out[i] = f2bf(n + rbf(rbf(l - n) * coef));
}

// extra heads of the STT models (lm.py:224-226, 803-806): softmax(extra_head(transformer_out)) in the model dtype.
// One warp per (head, row); E <= 32 outputs per head.
static __global__ void extra_heads_kernel(const bf16* __restrict__ tout /*[B][dim]*/, const bf16* const* __restrict__ w /*[n][E][dim]*/,
                                   bf16* __restrict__ out /*[n][B][E]*/, int B, int dim, int E, int n_heads) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_heads * B) return;
  const int hd = warp / B, b = warp - hd * B;
  const bf16* wh = w[hd];
  float mine = -INFINITY;                      // lane e keeps logit e
  for (int e = 0; e < E; ++e) {
    float acc = 0.f;
    for (int k = lane; k < dim; k += 32) acc = fmaf(bf2f(tout[(long long)b * dim + k]), bf2f(wh[(long long)e * dim + k]), acc);
    acc = rbf(warp_sum(acc));                  // nn.Linear output is a bf16 tensor
    if (lane == e) mine = acc;
  }
  const float mx = warp_max(mine);
  const float ex = lane < E ? expf(mine - mx) : 0.f;
  const float sum = warp_sum(ex);
  if (lane < E) out[((long long)hd * B + b) * E + lane] = f2bf(ex / sum);
}

// Exp(1) noise for one step's samplers (sampling.py:44: torch.empty_like(probs).exponential_(1)), drawn inside the step's graph:
// Philox4x32-10 keyed by the session seed, counter = (element / 4, step counter); -log(u), u in (0, 1].
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static __global__ void lm_noise_kernel(float* __restrict__ noise, long long n, unsigned long long seed,
                                const unsigned long long* __restrict__ step_ctr) {
  pdl_trigger();
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one Philox block = 4 values
  if (q * 4 >= n) return;
  const unsigned long long ctr = *step_ctr;
  uint32_t r[4];
  philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long i = q * 4 + j;
    if (i < n) noise[i] = -__logf(((float)(r[j] >> 8) + 1.0f) * (1.0f / 16777216.0f));     // u in (2^-24, 1]
  }
}

// ---------------------------------------------------------------------------------------------
// RMSNorm with fp32 math, eps 1e-8 (transformer.py:45-58): y = bf16(x * (alpha * rsqrt(eps + mean(x^2))))
// Optionally fuses the preceding residual add: x <- bf16(x + upd) first (transformer.py:769,777).
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) rmsnorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ alpha,
                                                      bf16* __restrict__ y, int dim, float eps) {
  pdl_trigger();
  const int row = blockIdx.x;
  const bf16* xr = x + (long long)row * dim;
  float ss = 0.f;
  for (int c = threadIdx.x * 2; c < dim; c += blockDim.x * 2) {
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xr + c));
    ss += v.x * v.x + v.y * v.y;
  }
  __shared__ float red[32];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float r = rsqrtf(eps + red[0] / (float)dim);
  for (int c = threadIdx.x * 2; c < dim; c += blockDim.x * 2) {
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xr + c));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(alpha + c));
    *reinterpret_cast<__nv_bfloat162*>(y + (long long)row * dim + c) =
        __floats2bfloat162_rn(v.x * (a.x * r), v.y * (a.y * r));
  }
}

// epilogues of the LM linears (gemm_sk.cu): plain store, residual add, gated SiLU (gating.py:18-20)
enum { LIN_STORE = 0, LIN_RESADD = 1, LIN_GATE = 2 };

constexpr int ATT_D = 128;
constexpr int ATT_THREADS = 128;

// B200_ATT_U=2|4 (diagnostics): register-buffer depth of attn_step_kernel; default 4
inline int attn_group_keys() {
  static const int u = [] { const char* e = getenv("B200_ATT_U"); return (e && atoi(e) == 2) ? 2 : 4; }();
  return u;
}

// split-KV so that B*H*nsplit CTAs cover the 148 SMs a few times over even at B = 1
inline int attn_pick_splits(int B, int H, int cap) {
  int ns = (148 * 4 + B * H - 1) / (B * H);
  if (ns < 1) ns = 1;
  if (ns > 16) ns = 16;
  const int max_by_len = (cap + 63) / 64;
  if (ns > max_by_len) ns = max_by_len;
  return ns;
}

// ---------------------------------------------------------------------------------------------
// Fused temporal attention step (T = 1, D = 128): RoPE(q, k) + ring append + split-KV attention +
// split combine in ONE launch (transformer.py:557-597).
//   grid = (B*H, nsplit), 128 threads.  qkv [B][3C] bf16 (rows q | k | v, each (h d)).
//   * the CTA whose key range holds the slot of the new key rotates k, writes K/V there, syncs, and then
//     reads it back like any other key; no other CTA touches that slot (ranges are disjoint), so the
//     append needs no separate kernel and the oldest key it replaces is never attended;
//   * q is rotated in registers (rounded to bf16 like the reference's apply_rope output);
//   * with nsplit > 1 the last CTA to arrive for a (b, h) merges the partials in split order.
// ---------------------------------------------------------------------------------------------
struct AttnStep {
  const bf16* qkv; bf16* kc; bf16* vc; bf16* out;
  float* part; int* counters;                  // [B*H][nsplit][ATT_D + 2], [B*H] (zero between launches)
  const long long* pos; const uint8_t* exec_mask;
  int H, cap, nsplit; float neg_log_period_2_over_d;
};

template <int ATT_U>       // keys per half-warp per register buffer
static __global__ void __launch_bounds__(ATT_THREADS) attn_step_kernel(const AttnStep a) {
  pdl_trigger();
  const int bh = blockIdx.x, split = blockIdx.y;
  const int b = bh / a.H, h = bh - b * a.H;
  const int C = a.H * ATT_D;
  const int tid = threadIdx.x, lane = tid & 31, l16 = lane & 15;
  const int hw = tid >> 4;                          // half-warp id 0..7
  const bool exec = a.exec_mask[b] != 0;
  const long long p = a.pos[b];
  long long n_valid = p + (exec ? 1 : 0);
  if (n_valid > a.cap) n_valid = a.cap;
  const int per = (int)((n_valid + a.nsplit - 1) / a.nsplit);
  const int s0 = split * per;
  const int s1 = (int)min((long long)(s0 + per), n_valid);
  const int slot_new = (int)(p % a.cap);
  const bf16* base = a.qkv + (long long)b * 3 * C + h * ATT_D;
  const float fpos = (float)p;

  if (exec && slot_new >= s0 && slot_new < s1) {    // CTA-uniform: this CTA owns the new key's slot
    if (tid < ATT_D / 2) {
      const int pr = tid;
      float kr = bf2f(base[C + 2 * pr]), ki = bf2f(base[C + 2 * pr + 1]);
      const float freq = expf((float)pr * a.neg_log_period_2_over_d);
      float sn, cs;
      sincosf(freq * fpos, &sn, &cs);
      const float c2 = kr * cs - ki * sn, d2 = kr * sn + ki * cs;
      const long long o = ((long long)bh * a.cap + slot_new) * ATT_D + 2 * pr;
      *reinterpret_cast<__nv_bfloat162*>(a.kc + o) = __floats2bfloat162_rn(c2, d2);
      *reinterpret_cast<__nv_bfloat162*>(a.vc + o) = *reinterpret_cast<const __nv_bfloat162*>(base + 2 * C + 2 * pr);
    }
    __syncthreads();
  }

  float qf[8];
  {
    float raw[8];
    unpack8(*reinterpret_cast<const uint4*>(base + l16 * 8), raw);
    const float scale = 0.08838834764831845f;       // 1/sqrt(128)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pr = l16 * 4 + j;
      const float freq = expf((float)pr * a.neg_log_period_2_over_d);
      float sn, cs;
      sincosf(freq * fpos, &sn, &cs);
      const float qr = raw[2 * j], qi = raw[2 * j + 1];
      qf[2 * j] = rbf(qr * cs - qi * sn) * scale;
      qf[2 * j + 1] = rbf(qr * sn + qi * cs) * scale;
    }
  }

  const bf16* kb = a.kc + (long long)bh * a.cap * ATT_D + l16 * 8;
  const bf16* vb = a.vc + (long long)bh * a.cap * ATT_D + l16 * 8;
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;

  // Groups of U keys per half-warp, double-buffered in registers: the loads of group i + 1 are in flight while group i
  // is reduced (same structure as attn_step_q8_kernel below).
  constexpr int U = ATT_U;
  struct Group { uint4 k[U], v[U]; };
  auto load_group = [&](int kb0, Group& gp) {
    const int s = kb0 + hw * U;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ss = s + u < s1 ? s + u : s1 - 1;
      gp.k[u] = *reinterpret_cast<const uint4*>(kb + (long long)ss * ATT_D);
      gp.v[u] = *reinterpret_cast<const uint4*>(vb + (long long)ss * ATT_D);
    }
  };
  auto reduce_group = [&](int kb0, const Group& gp) {
    const int s = kb0 + hw * U;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float kf[8];
      unpack8(gp.k[u], kf);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) d = fmaf(qf[i], kf[i], d);
      d += __shfl_xor_sync(0xffffffffu, d, 8);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      if (s + u < s1) {
        const float mn = fmaxf(m, d);
        const float corr = __expf(m - mn), pw = __expf(d - mn);
        float vf[8];
        unpack8(gp.v[u], vf);
        l = l * corr + pw;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(pw, vf[i], acc[i] * corr);
        m = mn;
      }
    }
  };
  Group ga, gb;
  if (s0 < s1) load_group(s0, ga);
  for (int kb0 = s0; kb0 < s1; kb0 += 16 * U) {     // CTA-uniform trip count and branches (shuffles inside reduce_group)
    const int k1 = kb0 + 8 * U, k2 = kb0 + 16 * U;
    if (k1 < s1) load_group(k1, gb);
    reduce_group(kb0, ga);
    if (k1 < s1) {
      if (k2 < s1) load_group(k2, ga);
      reduce_group(k1, gb);
    }
  }
  // merge the 8 half-warps of the CTA
  __shared__ float sm_m[8], sm_l[8], sm_acc[8][ATT_D];
  __shared__ int s_last;
  if (l16 == 0) { sm_m[hw] = m; sm_l[hw] = l; }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm_acc[hw][l16 * 8 + i] = acc[i];
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < 8; ++w) M = fmaxf(M, sm_m[w]);
  float L = 0.f, A = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const float c = sm_m[w] == -INFINITY ? 0.f : __expf(sm_m[w] - M);
    L += sm_l[w] * c;
    A += sm_acc[w][tid] * c;
  }
  if (a.nsplit == 1) {
    a.out[(long long)bh * ATT_D + tid] = f2bf(L > 0.f ? A / L : 0.f);
    return;
  }
  float* o = a.part + ((long long)bh * a.nsplit + split) * (ATT_D + 2);
  __stcg(o + tid, A);
  if (tid == 0) { __stcg(o + ATT_D, M); __stcg(o + ATT_D + 1, L); }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int old = atomicAdd(a.counters + bh, 1);
    s_last = old == a.nsplit - 1;
    if (s_last) a.counters[bh] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* pp = a.part + (long long)bh * a.nsplit * (ATT_D + 2);
  float MM = -INFINITY;
  for (int s = 0; s < a.nsplit; ++s) MM = fmaxf(MM, __ldcg(pp + s * (ATT_D + 2) + ATT_D));
  float LL = 0.f, AA = 0.f;
  for (int s = 0; s < a.nsplit; ++s) {
    const float ms = __ldcg(pp + s * (ATT_D + 2) + ATT_D);
    const float c = ms == -INFINITY ? 0.f : __expf(ms - MM);
    LL += __ldcg(pp + s * (ATT_D + 2) + ATT_D + 1) * c;
    AA += __ldcg(pp + s * (ATT_D + 2) + tid) * c;
  }
  a.out[(long long)bh * ATT_D + tid] = f2bf(LL > 0.f ? AA / LL : 0.f);
}

// ---------------------------------------------------------------------------------------------
// Opt-in 8-bit KV rings (NOT the reference's numerics): the same fused step as attn_step_kernel
// over rings of one byte per element with one fp32 scale per (session, head, slot).  Halves the ring (0.81 GB instead
// of 1.57 GB per session at context 3000), i.e. doubles the sessions a GPU can hold.  Two encodings of a row x[128]:
//   KV_E4M3: e4m3(x * 448 / absmax), scale = absmax / 448   (keeps relative precision under outlier channels)
//   KV_INT8: round(x * 127 / absmax) + 128, scale = absmax / 127   (4x smaller error on smooth rows, cheaper decode:
//            PRMT places the byte under the fp16 exponent 0x64 (= 1024 + u exactly) and one HADD2 removes 1152)
// Append: the rotated key (rounded to bf16 first, like the bf16 ring) and the value are scaled and rounded to nearest
// even.  Attention: a half-warp owns one key (16 lanes x 8 bytes); keys and values are decoded to packed fp16 (exact),
// the dot product and the per-group weighted value sums run as HFMA2, and everything that accumulates over the ring
// (softmax statistics, output accumulator) stays fp32; the online softmax is updated once per group of keys.
// ---------------------------------------------------------------------------------------------
constexpr int KV_E4M3 = 1, KV_INT8 = 2;       // = B200_KV_FP8_E4M3, B200_KV_INT8

struct AttnStepQ8 {
  const bf16* qkv; uint8_t* kc; uint8_t* vc; float* ks; float* vs; bf16* out;
  float* part; int* counters;
  const long long* pos; const uint8_t* exec_mask;
  int H, cap, nsplit; float neg_log_period_2_over_d;
};

__device__ __forceinline__ __half2 e4m3x2_to_half2(uint32_t two) {
  const __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(two & 0xFFFFu), __NV_E4M3);    // exact
  __half2 h2;
  memcpy(&h2, &hr, sizeof(h2));
  return h2;
}
__device__ __forceinline__ __half2 u8x2_to_half2(uint32_t w, uint32_t selector) {
  const uint32_t bits = __byte_perm(w, 0x64646464u, selector);     // [b, 0x64, b', 0x64] = (1024 + b, 1024 + b')
  __half2 h;
  memcpy(&h, &bits, sizeof(h));
  return __hsub2(h, __floats2half2_rn(1152.f, 1152.f));            // exact: values in [-128, 127]
}
template <int FMT>
__device__ __forceinline__ void unpack8_q8(const uint2& v, __half2* h) {
  if (FMT == KV_INT8) {
    h[0] = u8x2_to_half2(v.x, 0x4140u);
    h[1] = u8x2_to_half2(v.x, 0x4342u);
    h[2] = u8x2_to_half2(v.y, 0x4140u);
    h[3] = u8x2_to_half2(v.y, 0x4342u);
  } else {
    h[0] = e4m3x2_to_half2(v.x);
    h[1] = e4m3x2_to_half2(v.x >> 16);
    h[2] = e4m3x2_to_half2(v.y);
    h[3] = e4m3x2_to_half2(v.y >> 16);
  }
}
// two scaled values -> two ring bytes
template <int FMT>
__device__ __forceinline__ uint16_t quant2_q8(float x0, float x1) {
  if (FMT == KV_INT8) {
    const float r0 = fminf(fmaxf(rintf(x0), -127.f), 127.f), r1 = fminf(fmaxf(rintf(x1), -127.f), 127.f);
    return (uint16_t)((uint32_t)((int)r0 + 128) | ((uint32_t)((int)r1 + 128) << 8));
  }
  return (uint16_t)__nv_cvt_float2_to_fp8x2(make_float2(x0, x1), __NV_SATFINITE, __NV_E4M3);
}

template <int FMT>
static __global__ void __launch_bounds__(ATT_THREADS) attn_step_q8_kernel(const AttnStepQ8 a) {
  pdl_trigger();
  const int bh = blockIdx.x, split = blockIdx.y;
  const int b = bh / a.H, h = bh - b * a.H;
  const int C = a.H * ATT_D;
  const int tid = threadIdx.x, lane = tid & 31, l16 = lane & 15;
  const int hw = tid >> 4;                          // half-warp id 0..7
  const bool exec = a.exec_mask[b] != 0;
  const long long p = a.pos[b];
  long long n_valid = p + (exec ? 1 : 0);
  if (n_valid > a.cap) n_valid = a.cap;
  const int per = (int)((n_valid + a.nsplit - 1) / a.nsplit);
  const int s0 = split * per;
  const int s1 = (int)min((long long)(s0 + per), n_valid);
  const int slot_new = (int)(p % a.cap);
  const bf16* base = a.qkv + (long long)b * 3 * C + h * ATT_D;
  const float fpos = (float)p;

  if (exec && slot_new >= s0 && slot_new < s1) {    // CTA-uniform: this CTA owns the new key's slot
    __shared__ float s_amax[2][2];
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    if (tid < ATT_D / 2) {
      const int pr = tid;
      const float kr = bf2f(base[C + 2 * pr]), ki = bf2f(base[C + 2 * pr + 1]);
      const float freq = expf((float)pr * a.neg_log_period_2_over_d);
      float sn, cs;
      sincosf(freq * fpos, &sn, &cs);
      k0 = rbf(kr * cs - ki * sn);
      k1 = rbf(kr * sn + ki * cs);
      v0 = bf2f(base[2 * C + 2 * pr]);
      v1 = bf2f(base[2 * C + 2 * pr + 1]);
    }
    const float ak = warp_max(fmaxf(fabsf(k0), fabsf(k1))), av = warp_max(fmaxf(fabsf(v0), fabsf(v1)));
    if (tid < ATT_D / 2 && lane == 0) { s_amax[tid >> 5][0] = ak; s_amax[tid >> 5][1] = av; }
    __syncthreads();
    if (tid < ATT_D / 2) {
      const float amk = fmaxf(s_amax[0][0], s_amax[1][0]), amv = fmaxf(s_amax[0][1], s_amax[1][1]);
      constexpr float QMAX = FMT == KV_INT8 ? 127.f : 448.f;
      const float ik = amk > 0.f ? QMAX / amk : 0.f, iv = amv > 0.f ? QMAX / amv : 0.f;
      const long long o = ((long long)bh * a.cap + slot_new) * ATT_D + 2 * tid;
      *reinterpret_cast<uint16_t*>(a.kc + o) = quant2_q8<FMT>(k0 * ik, k1 * ik);
      *reinterpret_cast<uint16_t*>(a.vc + o) = quant2_q8<FMT>(v0 * iv, v1 * iv);
      if (tid == 0) {
        a.ks[(long long)bh * a.cap + slot_new] = amk * (1.f / QMAX);
        a.vs[(long long)bh * a.cap + slot_new] = amv * (1.f / QMAX);
      }
    }
    __syncthreads();
  }

  __half2 qh[4];
  {
    float raw[8];
    unpack8(*reinterpret_cast<const uint4*>(base + l16 * 8), raw);
    const float scale = 0.08838834764831845f;       // 1/sqrt(128)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pr = l16 * 4 + j;
      const float freq = expf((float)pr * a.neg_log_period_2_over_d);
      float sn, cs;
      sincosf(freq * fpos, &sn, &cs);
      const float qr = raw[2 * j], qi = raw[2 * j + 1];
      qh[j] = __floats2half2_rn(rbf(qr * cs - qi * sn) * scale, rbf(qr * sn + qi * cs) * scale);
    }
  }

  const uint8_t* kb = a.kc + (long long)bh * a.cap * ATT_D + l16 * 8;
  const uint8_t* vb = a.vc + (long long)bh * a.cap * ATT_D + l16 * 8;
  const float* ksb = a.ks + (long long)bh * a.cap;
  const float* vsb = a.vs + (long long)bh * a.cap;
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;

  // Groups of U keys per half-warp, double-buffered in registers: the loads of group i + 1 are in flight while group i is
  // reduced, so a warp's time per group is max(memory latency, arithmetic) instead of their sum (the bf16 kernel can
  // afford the sum: it moves twice the bytes per key).
  constexpr int U = 4;
  struct Group { uint2 k[U], v[U]; float ks[U], vs[U]; };
  auto load_group = [&](int kb0, Group& gp) {
    const int s = kb0 + hw * U;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ss = s + u < s1 ? s + u : s1 - 1;
      gp.k[u] = *reinterpret_cast<const uint2*>(kb + (long long)ss * ATT_D);
      gp.v[u] = *reinterpret_cast<const uint2*>(vb + (long long)ss * ATT_D);
      gp.ks[u] = ksb[ss];
      gp.vs[u] = vsb[ss];
    }
  };
  auto reduce_group = [&](int kb0, const Group& gp) {
    const int s = kb0 + hw * U;
    float d[U];
    float mn = m;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      __half2 kh[4];
      unpack8_q8<FMT>(gp.k[u], kh);
      __half2 t2 = __hmul2(qh[0], kh[0]);
      t2 = __hfma2(qh[1], kh[1], t2);
      t2 = __hfma2(qh[2], kh[2], t2);
      t2 = __hfma2(qh[3], kh[3], t2);
      const float2 tf = __half22float2(t2);
      float t = tf.x + tf.y;
      t += __shfl_xor_sync(0xffffffffu, t, 8);
      t += __shfl_xor_sync(0xffffffffu, t, 4);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      d[u] = s + u < s1 ? t * gp.ks[u] : -INFINITY;
      mn = fmaxf(mn, d[u]);
    }
    if (mn > -INFINITY) {                           // uniform over the half-warp; no shuffles inside
      const float corr = __expf(m - mn);            // m = -inf -> 0
      __half2 g[4];                                 // this group's weighted value sum, packed fp16
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = __floats2half2_rn(0.f, 0.f);
      float lsum = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float pw = __expf(d[u] - mn);         // padding keys: exp(-inf) = 0
        lsum += pw;
        const __half2 pv = __float2half2_rn(pw * gp.vs[u]);
        __half2 vh[4];
        unpack8_q8<FMT>(gp.v[u], vh);
#pragma unroll
        for (int i = 0; i < 4; ++i) g[i] = __hfma2(pv, vh[i], g[i]);
      }
      l = fmaf(l, corr, lsum);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 gf = __half22float2(g[i]);
        acc[2 * i] = fmaf(acc[2 * i], corr, gf.x);
        acc[2 * i + 1] = fmaf(acc[2 * i + 1], corr, gf.y);
      }
      m = mn;
    }
  };
  Group ga, gb;
  if (s0 < s1) load_group(s0, ga);
  for (int kb0 = s0; kb0 < s1; kb0 += 16 * U) {     // CTA-uniform trip count and branches (shuffles inside reduce_group)
    const int k1 = kb0 + 8 * U, k2 = kb0 + 16 * U;
    if (k1 < s1) load_group(k1, gb);
    reduce_group(kb0, ga);
    if (k1 < s1) {
      if (k2 < s1) load_group(k2, ga);
      reduce_group(k1, gb);
    }
  }
  // merge the 8 half-warps of the CTA, then the splits (same protocol as attn_step_kernel)
  __shared__ float sm_m[8], sm_l[8], sm_acc[8][ATT_D];
  __shared__ int s_last;
  if (l16 == 0) { sm_m[hw] = m; sm_l[hw] = l; }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm_acc[hw][l16 * 8 + i] = acc[i];
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < 8; ++w) M = fmaxf(M, sm_m[w]);
  float L = 0.f, A = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const float c = sm_m[w] == -INFINITY ? 0.f : __expf(sm_m[w] - M);
    L += sm_l[w] * c;
    A += sm_acc[w][tid] * c;
  }
  if (a.nsplit == 1) {
    a.out[(long long)bh * ATT_D + tid] = f2bf(L > 0.f ? A / L : 0.f);
    return;
  }
  float* o = a.part + ((long long)bh * a.nsplit + split) * (ATT_D + 2);
  __stcg(o + tid, A);
  if (tid == 0) { __stcg(o + ATT_D, M); __stcg(o + ATT_D + 1, L); }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int old = atomicAdd(a.counters + bh, 1);
    s_last = old == a.nsplit - 1;
    if (s_last) a.counters[bh] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* pp = a.part + (long long)bh * a.nsplit * (ATT_D + 2);
  float MM = -INFINITY;
  for (int s = 0; s < a.nsplit; ++s) MM = fmaxf(MM, __ldcg(pp + s * (ATT_D + 2) + ATT_D));
  float LL = 0.f, AA = 0.f;
  for (int s = 0; s < a.nsplit; ++s) {
    const float ms = __ldcg(pp + s * (ATT_D + 2) + ATT_D);
    const float c = ms == -INFINITY ? 0.f : __expf(ms - MM);
    LL += __ldcg(pp + s * (ATT_D + 2) + ATT_D + 1) * c;
    AA += __ldcg(pp + s * (ATT_D + 2) + tid) * c;
  }
  a.out[(long long)bh * ATT_D + tid] = f2bf(LL > 0.f ? AA / LL : 0.f);
}

constexpr int DEP_MAX_Q = 16;              // dep_q of the largest family member (configs/moshi_dev_2b.json)

// Depformer attention step: the new key/value of sub-step `step` is appended to the per-frame cache and the query
// attends over step + 1 keys; no positional embedding (depformer_pos_emb = "none").  One warp per (b, h), D = 64.
static __global__ void dep_attn_step_kernel(const bf16* __restrict__ qkv /*[B][3*H*D]*/, bf16* __restrict__ kc,
                                     bf16* __restrict__ vc, bf16* __restrict__ out, int B, int H, int cap, int step) {
  pdl_trigger();
  constexpr int D = 64;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * H) return;
  const int b = warp / H, h = warp - b * H;
  const int C = H * D;
  const bf16* base = qkv + (long long)b * 3 * C + h * D + 2 * lane;
  const long long row = (long long)warp * cap;
  *reinterpret_cast<__nv_bfloat162*>(kc + (row + step) * D + 2 * lane) = *reinterpret_cast<const __nv_bfloat162*>(base + C);
  *reinterpret_cast<__nv_bfloat162*>(vc + (row + step) * D + 2 * lane) = *reinterpret_cast<const __nv_bfloat162*>(base + 2 * C);
  __syncwarp();
  const float2 q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(base));
  const float scale = 0.125f;                       // 1/sqrt(64)
  float sc[DEP_MAX_Q];
  float mx = -INFINITY;
  const int n_keys = step + 1;
  for (int j = 0; j < n_keys; ++j) {
    const float2 kk = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(kc + (row + j) * D + 2 * lane));
    const float d = warp_sum(q.x * kk.x + q.y * kk.y) * scale;
    sc[j] = d;
    mx = fmaxf(mx, d);
  }
  float sum = 0.f;
  for (int j = 0; j < n_keys; ++j) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
  const float inv = 1.f / sum;
  float a0 = 0.f, a1 = 0.f;
  for (int j = 0; j < n_keys; ++j) {
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vc + (row + j) * D + 2 * lane));
    a0 = fmaf(sc[j] * inv, v.x, a0);
    a1 = fmaf(sc[j] * inv, v.y, a1);
  }
  *reinterpret_cast<__nv_bfloat162*>(out + (long long)warp * D + 2 * lane) = __floats2bfloat162_rn(a0, a1);
}

// _LMGenState.reset / _MHAState.reset / RingKVCache.reset / State.reset for the rows in `mask` (null = all)
// (with CFG the model rows b and B + b are reset together, lm.py:653-656)
static __global__ void lm_reset_kernel(long long* offsets, long long* pos, uint8_t* exec_mask, uint8_t* exec_mask_m, const uint8_t* mask,
                                int B, int cfg) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (mask != nullptr && !mask[b]) return;
  offsets[b] = 0;
  pos[b] = 0;
  exec_mask[b] = 1;
  exec_mask_m[b] = 1;
  if (cfg) { pos[B + b] = 0; exec_mask_m[B + b] = 1; }
}

// depformer_replace_tokens [B][dep_q] -> this frame's audio tokens [dep_q][B] (lm.py:751-755)
static __global__ void replace_audio_kernel(const long long* __restrict__ given, long long* __restrict__ audio_tokens, int B, int dep_q) {
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * dep_q) return;
  const int b = i / dep_q, k = i - b * dep_q;
  audio_tokens[(long long)k * B + b] = given[i];
}

// rings shorter than the model's context (b200_lm_set_kv_capacity): a row about to write position >= cap would overwrite keys the
// reference still attends to -> ERR_KV_CAPACITY (the step still runs, as a ring of `cap` slots)
static __global__ void kv_capacity_check_kernel(const long long* __restrict__ pos, const uint8_t* __restrict__ exec_mask, int B, int cap,
                                                int* err) {
  pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && exec_mask[b] && pos[b] >= cap && err != nullptr) atomicOr(err, ERR_KV_CAPACITY);
}

static __global__ void advance_pos_kernel(long long* pos, const uint8_t* exec_mask, int B) {
  pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && exec_mask[b]) pos[b] += 1;
}

// ---------------------------------------------------------------------------------------------
// sample_token (sampling.py:86-106): softmax(logits/temp) -> top-k -> argmax(p / Exp(1) noise).
// Candidates are ranked by (logit desc, index asc); the noise is indexed by rank like the
// reference indexes it by the position in torch.topk's sorted output.
// ---------------------------------------------------------------------------------------------
constexpr int SAMPLE_THREADS = 256;
constexpr int SAMPLE_MAX_K = 1024;

__device__ __forceinline__ unsigned sample_key(const bf16* logits, int i) {
  const unsigned short raw = __bfloat16_as_ushort(logits[i]);
  const unsigned short mono = (raw & 0x8000u) ? (unsigned short)~raw : (unsigned short)(raw | 0x8000u);
  return ((unsigned)mono << 16) | (unsigned)(0xFFFFu - (unsigned)i);
}

static __device__ float block_reduce(float v, bool is_max, float* red) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = is_max ? -INFINITY : 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = is_max ? fmaxf(t, red[w]) : t + red[w];
  return t;
}

// One row per call; every thread of a SAMPLE_THREADS-wide CTA must call it (block-wide barriers inside).
static __device__ void sample_row(const bf16* __restrict__ logits, const float* __restrict__ noise, long long* __restrict__ out,
                                  int card, int use_sampling, float temp, int top_k) {
  __shared__ float red[SAMPLE_THREADS / 32];
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_remaining, s_count;
  __shared__ unsigned sel[SAMPLE_MAX_K];
  __shared__ float s_best[SAMPLE_THREADS / 32];
  __shared__ int s_brank[SAMPLE_THREADS / 32];
  const int tid = threadIdx.x;

  if (!(use_sampling && temp > 0.f)) {        // greedy: argmax, first maximum (sampling.py:102-103)
    unsigned best = 0;
    for (int i = tid; i < card; i += SAMPLE_THREADS) best = max(best, sample_key(logits, i));
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((tid & 31) == 0) hist[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
      unsigned bb = 0;
      for (int w = 0; w < SAMPLE_THREADS / 32; ++w) bb = max(bb, hist[w]);
      *out = 0xFFFFu - (bb & 0xFFFFu);
    }
    return;
  }
  const int k = min(min(top_k, card), SAMPLE_MAX_K);
  // softmax statistics in fp32 on logits / temp
  float mx = -INFINITY;
  for (int i = tid; i < card; i += SAMPLE_THREADS) mx = fmaxf(mx, bf2f(logits[i]) / temp);
  mx = block_reduce(mx, true, red);
  float sum = 0.f;
  for (int i = tid; i < card; i += SAMPLE_THREADS) sum += expf(bf2f(logits[i]) / temp - mx);
  sum = block_reduce(sum, false, red);

  // radix select of the k-th largest 32-bit key (keys are unique)
  if (tid == 0) { s_prefix = 0; s_remaining = (unsigned)k; }
  __syncthreads();
  for (int pass = 3; pass >= 0; --pass) {
    hist[tid] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned mask = pass == 3 ? 0u : (0xFFFFFFFFu << ((pass + 1) * 8));
    for (int i = tid; i < card; i += SAMPLE_THREADS) {
      const unsigned key = sample_key(logits, i);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> (pass * 8)) & 0xFFu], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      // the bin of the rem-th largest key, scanning the 256 counts from the top: lane l owns bins 8l .. 8l+7, a suffix sum over
      // the lanes finds the one lane whose bins contain it (the serial scan by one thread cost ~4 us per pass)
      const unsigned rem = s_remaining;
      unsigned h8[8], own = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { h8[i] = hist[tid * 8 + i]; own += h8[i]; }
      unsigned incl = own;                         // counts in this lane's bins and every higher lane's
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_down_sync(0xffffffffu, incl, o);
        if (tid + o < 32) incl += v;
      }
      unsigned cum = incl - own;                   // counts in bins above this lane's
      if (cum < rem && rem <= incl) {
        int bin = tid * 8;
#pragma unroll
        for (int i = 7; i >= 0; --i) {
          if (cum + h8[i] >= rem) { bin = tid * 8 + i; break; }
          cum += h8[i];
        }
        s_remaining = rem - cum;
        s_prefix = prefix | ((unsigned)bin << (pass * 8));
      }
    }
    __syncthreads();
  }
  const unsigned thr = s_prefix;               // exactly k keys are >= thr
  if (tid == 0) s_count = 0;
  // pad the sort buffer
  int p2 = 1;
  while (p2 < k) p2 <<= 1;
  for (int i = tid; i < p2; i += SAMPLE_THREADS) sel[i] = 0u;
  __syncthreads();
  for (int i = tid; i < card; i += SAMPLE_THREADS) {
    const unsigned key = sample_key(logits, i);
    if (key >= thr) {
      const unsigned slot = atomicAdd(&s_count, 1u);
      if (slot < (unsigned)p2) sel[slot] = key;
    }
  }
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= p2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < p2 / 2; i += SAMPLE_THREADS) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned a = sel[lo], b = sel[hi];
        if ((a < b) == desc) { sel[lo] = b; sel[hi] = a; }
      }
      __syncthreads();
    }
  }
  // argmax over ranks of p / q (first maximum wins, like torch.argmax)
  float best = -INFINITY;
  int brank = 0x7fffffff;
  for (int j = tid; j < k; j += SAMPLE_THREADS) {
    const int idx = 0xFFFF - (int)(sel[j] & 0xFFFFu);
    const float p = expf(bf2f(logits[idx]) / temp - mx) / sum;
    const float s = p / noise[j];
    if (s > best || (s == best && j < brank)) { best = s; brank = j; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, brank, o);
    if (ob > best || (ob == best && oj < brank)) { best = ob; brank = oj; }
  }
  if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_brank[tid >> 5] = brank; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < SAMPLE_THREADS / 32; ++w)
      if (s_best[w] > best || (s_best[w] == best && s_brank[w] < brank)) { best = s_best[w]; brank = s_brank[w]; }
    if (brank == 0x7fffffff) brank = 0;
    *out = 0xFFFF - (int)(sel[brank] & 0xFFFFu);
  }
}

static __global__ void __launch_bounds__(SAMPLE_THREADS) sample_kernel(const bf16* __restrict__ logits_all, long long ld,
                                                                const float* __restrict__ noise, long long noise_ld,
                                                                long long* __restrict__ out, int card, int use_sampling,
                                                                float temp, int top_k) {
  pdl_trigger();
  const int row = blockIdx.x;
  sample_row(logits_all + (long long)row * ld, noise + (long long)row * noise_ld, out + row, card, use_sampling, temp, top_k);
}

}  // namespace lm
}  // namespace b200
