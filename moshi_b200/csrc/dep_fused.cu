// The Depth transformer of one frame as ONE persistent kernel (LMGen.depformer_step, lm.py:809-850, with
// LMModel.forward_depformer lm.py:450-493 and the weights-per-step layers of transformer.py:291-318 inside).
//
// The depformer is 8 dependent sub-steps x 6 layers of tiny GEMMs (1.23 GB of weights per frame, M = sessions):
// as separate launches that is ~430 kernels whose launch / set-up / drain latency is 10-20x their HBM time.
// Here one CTA per SM stays resident for the whole frame and the ~50 phases of a sub-step are separated by
// grid barriers (all CTAs are co-resident: cooperative launch, 1 CTA/SM by shared-memory size, grid <= #SMs):
//
//   per sub-step k:   [row phase]   x = depformer_in[k](transformer_out) + emb_k(prev token); xn = rmsnorm(x)
//     per layer l:    [GEMM]        in_proj partials            (tcgen05, split-K units over all CTAs)
//                     [warp phase]  q,k,v = bf16(sum partials); KV append; attention over k+1 keys -> ao
//                     [GEMM]        out_proj partials
//                     [row phase]   x = bf16(x + bf16(sum)); xn = rmsnorm(x)
//                     [GEMM]        gated-MLP input partials (gate and value accumulators)
//                     [elem phase]  h = bf16(bf16(silu(g)) * u)
//                     [GEMM]        linear_out partials
//                     [row phase]   x = bf16(x + bf16(sum)); xn = rmsnorm(x) for the next layer
//                     ...
//                     [GEMM]        logits partials (linears[k])
//                     [row phase]   logits = bf16(sum) ; token = sample(logits) ; next sub-step's input row
//
// GEMM phases use the same machinery as gemm_sk.cu (pre-tiled SWIZZLE_128B weights streamed by cp.async.bulk, the
// activation box by 2-D TMA, single-thread tcgen05.mma into TMEM, tcgen05.ld epilogue); a unit = (128-row weight
// tile, k-split) and every unit writes its fp32 partial [M x 128] to an L2-resident workspace; the consumer phase
// sums the splits in split order (deterministic, independent of the batch) and applies the reference's cast points.
#include "gemm_sk.cuh"
#include "tc_prims.cuh"
#include "lm_kernels.cuh"

namespace b200 {
namespace tc {

namespace {

using lm::bf16;

constexpr int BLOCK_ROWS = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int THREADS = 256;                 // warp 0 producer, 1 MMA, 2-5 epilogue, 6-7 extra hands in the SIMT phases
constexpr int TILE_BYTES = BLOCK_ROWS * BLOCK_K * 2;
constexpr int MAX_STAGES = 8;
constexpr int DD = 64;                       // depformer head dim
static_assert(THREADS == lm::SAMPLE_THREADS, "sample_row needs a SAMPLE_THREADS-wide CTA");

struct Gemm {                                // one GEMM phase
  const uint8_t* wt;                         // packed tiles [n_tile][kb][A][16 KB]
  int n_tiles, num_kb, kbps, S, A, N;        // S splits of kbps k-blocks; N = output columns per accumulator
};

using namespace tcp;
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

}  // namespace

struct DepLayerW { const uint8_t *in_w, *out_w, *lin_in, *lin_out; };
struct DepParams {
  int B, Mpad, dd, H, F, card, text_card, dep_q, L, n_tables;
  int* err;                                                  // device error flags (lm::ERR_*)
  int stages; uint32_t stage_bytes, tmem_cols, acc_cols; int n_acc;
  Gemm g_in, g_out, g_lin_in, g_lin_out, g_head;              // shapes (wt filled per (k, l) from the tables below)
  const DepLayerW* w;                                        // [dep_q][L] packed weight pointers (device)
  const uint8_t* const* heads;                               // [dep_q]
  const bf16* const* tables;                                 // [dep_q] embedding tables ([0] = text)
  const bf16* const* n1; const bf16* const* n2;              // [L] RMSNorm alphas
  const bf16* din; long long din_ld;                         // depformer_in_all output [B][dep_q*dd]
  const long long* text_token;                               // [B]
  bf16 *x, *xn, *ao, *hbuf;                                  // [B][dd] x3, [B][F]
  bf16* const* kc; bf16* const* vc;                          // [L] per-frame KV [B][H][dep_q][64]
  float *part0, *part1;                                      // split partials [S][B][Nmax]
  bf16* logits;                                              // [dep_q][B][card]
  long long* audio_tokens;                                   // [dep_q][B]
  const float* noise; long long noise_ld; int noise_off, ka; // Exp(1) noise, row stride, offset of sub-step 0, per-step width
  int use_sampling, top_k; float temp;
  unsigned* bar;                                             // grid barrier counter (zero at launch)
  unsigned long long* trace;                                 // optional [DEP_TRACE_SLOTS]: %globaltimer of CTA 0 after every grid barrier
};

namespace {

struct Pipe { int s; uint32_t ph; int acc; uint32_t acc_bits; int pre; };   // per-role pipeline cursor, carried across phases
// (pre: producer only — stages of the upcoming GEMM phase whose weight copy is already in flight)

// Grid barrier: one release-reduction per CTA and acquire polling by one thread (all CTAs are co-resident).
// The CTA's own writes are ordered before the reduction by __syncthreads (cumulativity of the gpu-scope release).
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& epoch, unsigned long long* trace = nullptr) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    const unsigned target = epoch * gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned spins = 0;
    uint64_t t0 = 0;
    while (ld_acquire(bar) < target) {               // all CTAs are co-resident (cooperative launch): bounded by wall clock only
      if ((++spins & 4095u) == 0u) {
        const uint64_t now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > WAIT_LIMIT_NS) __trap();
      }
    }
    if (trace != nullptr && blockIdx.x == 0 && epoch < (unsigned)DEP_TRACE_SLOTS) trace[epoch] = global_timer_ns();
  }
  __syncthreads();
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) t += red[w];
  return t;
}

// rows handled by this CTA: m = blockIdx.x, blockIdx.x + gridDim.x, ...   (same mapping in every row phase, so a row's
// sampled token is consumed by the CTA that produced it without a barrier)
//   x_new = has_sum ? bf16(x + bf16(sum_s part[s][m][:])) : x (already written);  xn = rmsnorm(x_new, alpha)
template <int NV, class P>                                    // dd <= NV * THREADS
__device__ void row_phase(const P& p, int dd, int S, const float* part, int N, bool has_sum, const bf16* alpha, float* red) {
  for (int m = blockIdx.x; m < p.B; m += gridDim.x) {
    float vals[NV];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = threadIdx.x + i * THREADS;
      float v = 0.f;
      if (j < dd) {
        v = bf2f(p.x[(long long)m * dd + j]);
        if (has_sum) {
          float a = 0.f;
          for (int s = 0; s < S; ++s) a += __ldcg(part + ((long long)s * p.B + m) * N + j);
          v = rbf(v + rbf(a));                                // x_orig + update, both bf16 (transformer.py:769,777)
          p.x[(long long)m * dd + j] = f2bf(v);
        }
        ss += v * v;
      }
      vals[i] = v;
    }
    if (alpha != nullptr) {
      const float tot = block_sum(ss, red);
      const float r = rsqrtf(1e-8f + tot / (float)dd);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int j = threadIdx.x + i * THREADS;
        if (j < dd) p.xn[(long long)m * dd + j] = f2bf(vals[i] * (bf2f(alpha[j]) * r));
      }
    }
  }
}

// x = depformer_in[k](transformer_out) + emb_k(prev)   (lm.py:475-486; token -1 -> zero row, lm_utils.py:103-121)
__device__ void input_rows(const DepParams& p, int k) {
  const long long* prev = k == 0 ? p.text_token : p.audio_tokens + (long long)(k - 1) * p.B;
  const bf16* table = p.tables[k];
  for (int m = blockIdx.x; m < p.B; m += gridDim.x) {
    const long long id = prev[m];
    const bool ok = lm::embed_id_ok(id, k == 0 ? p.text_card : p.card, p.err);
    for (int j = threadIdx.x; j < p.dd; j += THREADS) {
      const float e = ok ? bf2f(table[id * p.dd + j]) : 0.f;
      p.x[(long long)m * p.dd + j] = f2bf(bf2f(p.din[(long long)m * p.din_ld + (long long)k * p.dd + j]) + e);
    }
  }
  __syncthreads();                                            // row_phase re-reads x written by other threads of this CTA
}

// q,k,v = bf16(sum of in_proj partials); append k,v at slot `step`; attention over step+1 keys (no RoPE in the depformer)
__device__ void attn_phase(const DepParams& p, int S, int layer, int step) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.dd, N = 3 * C;
  bf16* kc = p.kc[layer];
  bf16* vc = p.vc[layer];
  for (int it = blockIdx.x * (THREADS / 32) + warp; it < p.B * p.H; it += gridDim.x * (THREADS / 32)) {
    const int b = it / p.H, h = it - b * p.H;
    float q[2] = {0.f, 0.f}, kk[2] = {0.f, 0.f}, vv[2] = {0.f, 0.f};
    for (int s = 0; s < S; ++s) {
      const float* row = p.part0 + ((long long)s * p.B + b) * N + h * DD + 2 * lane;
      const float2 a = __ldcg(reinterpret_cast<const float2*>(row));
      const float2 c = __ldcg(reinterpret_cast<const float2*>(row + C));
      const float2 d = __ldcg(reinterpret_cast<const float2*>(row + 2 * C));
      q[0] += a.x; q[1] += a.y; kk[0] += c.x; kk[1] += c.y; vv[0] += d.x; vv[1] += d.y;
    }
    const long long rowo = ((long long)b * p.H + h) * p.dep_q;
    *reinterpret_cast<__nv_bfloat162*>(kc + (rowo + step) * DD + 2 * lane) = __floats2bfloat162_rn(kk[0], kk[1]);
    *reinterpret_cast<__nv_bfloat162*>(vc + (rowo + step) * DD + 2 * lane) = __floats2bfloat162_rn(vv[0], vv[1]);
    __syncwarp();
    const float qx = rbf(q[0]), qy = rbf(q[1]);
    float sc[lm::DEP_MAX_Q];
    float mx = -INFINITY;
    for (int j = 0; j <= step; ++j) {
      const float2 kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(kc + (rowo + j) * DD + 2 * lane));
      const float d = warp_sum(qx * kv.x + qy * kv.y) * 0.125f;
      sc[j] = d;
      mx = fmaxf(mx, d);
    }
    float sum = 0.f;
    for (int j = 0; j <= step; ++j) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
    const float inv = 1.f / sum;
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j <= step; ++j) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vc + (rowo + j) * DD + 2 * lane));
      a0 = fmaf(sc[j] * inv, v.x, a0);
      a1 = fmaf(sc[j] * inv, v.y, a1);
    }
    *reinterpret_cast<__nv_bfloat162*>(p.ao + (long long)b * C + h * DD + 2 * lane) = __floats2bfloat162_rn(a0, a1);
  }
}

// h = bf16(bf16(silu(bf16 gate)) * bf16 value)   (gating.py:18-20)
template <class P>
__device__ void gate_phase(const P& p, int S) {
  const long long total = (long long)p.B * p.F;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * THREADS) {
    float g = 0.f, u = 0.f;
    for (int s = 0; s < S; ++s) {
      g += __ldcg(p.part0 + (long long)s * total + i);
      u += __ldcg(p.part1 + (long long)s * total + i);
    }
    g = rbf(g); u = rbf(u);
    p.hbuf[i] = f2bf(rbf(g / (1.f + expf(-g))) * u);
  }
}

// logits = bf16(sum partials) -> sample -> audio_tokens[k]
__device__ void sample_phase(const DepParams& p, int S, int k) {
  for (int m = blockIdx.x; m < p.B; m += gridDim.x) {
    bf16* lg = p.logits + ((long long)k * p.B + m) * p.card;
    for (int j = threadIdx.x; j < p.card; j += THREADS) {
      float a = 0.f;
      for (int s = 0; s < S; ++s) a += __ldcg(p.part0 + ((long long)s * p.B + m) * p.card + j);
      lg[j] = f2bf(a);
    }
    __syncthreads();
    lm::sample_row(lg, p.noise + (long long)m * p.noise_ld + p.noise_off + (long long)k * p.ka, p.audio_tokens + (long long)k * p.B + m,
                   p.card, p.use_sampling, p.temp, p.top_k);
    __syncthreads();
  }
}

// One GEMM phase: units u = blockIdx.x, += gridDim.x; unit = (tile, split); partial [M x 128] -> part[split][m][tile*128 + row]
template <class P>
__device__ void gemm_phase(const P& p, const Gemm& g, const Gemm* next, const CUtensorMap* xmap, uint32_t base, uint32_t full0,
                           uint32_t empty0, uint32_t tfull0, uint32_t tempty0, uint32_t tmem_base, Pipe& pipe) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = g.n_tiles * g.S;
  const uint32_t a_bytes = (uint32_t)g.A * TILE_BYTES;
  const uint32_t x_bytes = (uint32_t)p.Mpad * BLOCK_K * 2;
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");        // activations were written with generic stores by other CTAs
      int item = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int tile = u / g.S, sp = u - tile * g.S;
        const int kb0 = sp * g.kbps, kb1 = min(g.num_kb, kb0 + g.kbps);
        const uint8_t* src = g.wt + ((size_t)tile * g.num_kb + kb0) * a_bytes;
        for (int kb = kb0; kb < kb1; ++kb, src += a_bytes, ++item) {
          const uint32_t sa = base + (uint32_t)pipe.s * p.stage_bytes;
          if (item >= pipe.pre) {                              // (else: requested at the end of the previous GEMM phase)
            mbar_wait(empty0 + 8 * pipe.s, pipe.ph ^ 1u);
            mbar_expect_tx(full0 + 8 * pipe.s, a_bytes + x_bytes);
            bulk_load(sa, src, a_bytes, full0 + 8 * pipe.s);
          }
          tma_load_2d(sa + 2 * TILE_BYTES, xmap, full0 + 8 * pipe.s, kb * BLOCK_K, 0);
          if (++pipe.s == p.stages) { pipe.s = 0; pipe.ph ^= 1u; }
        }
      }
      // Weights never depend on the phases in between: request the next GEMM phase's first tiles now, so its HBM
      // latency (and, at large batch, most of its stream) overlaps the SIMT phase and the two grid barriers ahead.
      pipe.pre = 0;
      if (next != nullptr) {
        const uint32_t nb = (uint32_t)next->A * TILE_BYTES;
        int s2 = pipe.s; uint32_t ph2 = pipe.ph;
        const int nunits = next->n_tiles * next->S;
        for (int u = blockIdx.x; u < nunits && pipe.pre < p.stages; u += gridDim.x) {
          const int tile = u / next->S, sp = u - tile * next->S;
          const int kb0 = sp * next->kbps, kb1 = min(next->num_kb, kb0 + next->kbps);
          const uint8_t* src = next->wt + ((size_t)tile * next->num_kb + kb0) * nb;
          for (int kb = kb0; kb < kb1 && pipe.pre < p.stages; ++kb, src += nb, ++pipe.pre) {
            mbar_wait(empty0 + 8 * s2, ph2 ^ 1u);
            mbar_expect_tx(full0 + 8 * s2, nb + x_bytes);
            bulk_load(base + (uint32_t)s2 * p.stage_bytes, src, nb, full0 + 8 * s2);
            if (++s2 == p.stages) { s2 = 0; ph2 ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_ROWS, p.Mpad);
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int tile = u / g.S, sp = u - tile * g.S;
        const int kb0 = sp * g.kbps, kb1 = min(g.num_kb, kb0 + g.kbps);
        mbar_wait(tempty0 + 8 * pipe.acc, ((pipe.acc_bits >> pipe.acc) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(pipe.acc * p.acc_cols);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full0 + 8 * pipe.s, pipe.ph);
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)pipe.s * p.stage_bytes;
          const uint32_t sb = sa + 2 * TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t db = make_desc(sb + k * UMMA_K * 2);
            const uint32_t accum = (kb == kb0 && k == 0) ? 0u : 1u;
            umma_bf16(d0, make_desc(sa + k * UMMA_K * 2), db, idesc, accum);
            if (g.A == 2) umma_bf16(d0 + (uint32_t)p.Mpad, make_desc(sa + TILE_BYTES + k * UMMA_K * 2), db, idesc, accum);
          }
          umma_commit(empty0 + 8 * pipe.s);
          if (++pipe.s == p.stages) { pipe.s = 0; pipe.ph ^= 1u; }
        }
        umma_commit(tfull0 + 8 * pipe.acc);
        pipe.acc_bits ^= 1u << pipe.acc;
        if (p.n_acc == 2) pipe.acc ^= 1;
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int tile = u / g.S, sp = u - tile * g.S;
      const int n = tile * BLOCK_ROWS + row;
      mbar_wait(tfull0 + 8 * pipe.acc, (pipe.acc_bits >> pipe.acc) & 1u);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pipe.acc * p.acc_cols);
      float* o0 = p.part0 + (long long)sp * p.B * g.N + n;
      float* o1 = p.part1 + (long long)sp * p.B * g.N + n;
      for (int c0 = 0; c0 < p.Mpad; c0 += 16) {
        if (c0 >= p.B) break;
        uint32_t r0[16], r1[16];
        tmem_ld16(lane_addr + (uint32_t)c0, r0);
        if (g.A == 2) tmem_ld16(lane_addr + (uint32_t)(p.Mpad + c0), r1);
        tmem_ld_wait();
        if (n < g.N) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = c0 + j;
            if (m < p.B) {
              __stcg(o0 + (long long)m * g.N, __uint_as_float(r0[j]));
              if (g.A == 2) __stcg(o1 + (long long)m * g.N, __uint_as_float(r1[j]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * pipe.acc);
      pipe.acc_bits ^= 1u << pipe.acc;
      if (p.n_acc == 2) pipe.acc ^= 1;
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
dep_fused_kernel(const __grid_constant__ CUtensorMap map_xn, const __grid_constant__ CUtensorMap map_ao,
                 const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_x, const DepParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float red[THREADS / 32];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + (uint32_t)p.stages * p.stage_bytes;
  const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES, tfull0 = bars + 16 * MAX_STAGES, tempty0 = tfull0 + 16;
  const uint32_t tptr = tempty0 + 16;
  uint32_t* tptr_generic = reinterpret_cast<uint32_t*>(smem_raw + (tptr - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xn) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ao) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_h) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tptr, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tptr_generic;

  Pipe pipe{0, 0u, 0, 0u, 0};
  unsigned epoch = 0;
  if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.trace[0] = global_timer_ns();
  // GEMM phases in execution order: per sub-step k, per layer l: in_proj, out_proj, linear_in, linear_out; then the head
  auto gemm_at = [&](int k, int l, int which) -> Gemm {
    Gemm g;
    if (which == 4) { g = p.g_head; g.wt = p.heads[k]; return g; }
    const DepLayerW w = p.w[k * p.L + l];
    if (which == 0) { g = p.g_in; g.wt = w.in_w; }
    else if (which == 1) { g = p.g_out; g.wt = w.out_w; }
    else if (which == 2) { g = p.g_lin_in; g.wt = w.lin_in; }
    else { g = p.g_lin_out; g.wt = w.lin_out; }
    return g;
  };
  Gemm cur = gemm_at(0, 0, 0);
  for (int k = 0; k < p.dep_q; ++k) {
    input_rows(p, k);
    row_phase<4>(p, p.dd, 0, nullptr, 0, false, p.n1[0], red);
    grid_sync(p.bar, epoch, p.trace);
    for (int l = 0; l < p.L; ++l) {
      Gemm nxt = gemm_at(k, l, 1);
      gemm_phase(p, cur, &nxt, &map_xn, base, full0, empty0, tfull0, tempty0, tmem_base, pipe);
      grid_sync(p.bar, epoch, p.trace);
      attn_phase(p, cur.S, l, k);
      grid_sync(p.bar, epoch, p.trace);
      cur = nxt; nxt = gemm_at(k, l, 2);
      gemm_phase(p, cur, &nxt, &map_ao, base, full0, empty0, tfull0, tempty0, tmem_base, pipe);
      grid_sync(p.bar, epoch, p.trace);
      row_phase<4>(p, p.dd, cur.S, p.part0, cur.N, true, p.n2[l], red);
      grid_sync(p.bar, epoch, p.trace);
      cur = nxt; nxt = gemm_at(k, l, 3);
      gemm_phase(p, cur, &nxt, &map_xn, base, full0, empty0, tfull0, tempty0, tmem_base, pipe);
      grid_sync(p.bar, epoch, p.trace);
      gate_phase(p, cur.S);
      grid_sync(p.bar, epoch, p.trace);
      cur = nxt; nxt = l + 1 < p.L ? gemm_at(k, l + 1, 0) : gemm_at(k, 0, 4);
      gemm_phase(p, cur, &nxt, &map_h, base, full0, empty0, tfull0, tempty0, tmem_base, pipe);
      grid_sync(p.bar, epoch, p.trace);
      // depformer_norms is Identity (lm.py:197-198): after the last layer the head reads x itself
      row_phase<4>(p, p.dd, cur.S, p.part0, cur.N, true, l + 1 < p.L ? p.n1[l + 1] : nullptr, red);
      grid_sync(p.bar, epoch, p.trace);
      cur = nxt;
    }
    const bool last = k + 1 == p.dep_q;
    Gemm nxt = last ? cur : gemm_at(k + 1, 0, 0);
    gemm_phase(p, cur, last ? nullptr : &nxt, &map_x, base, full0, empty0, tfull0, tempty0, tmem_base, pipe);
    grid_sync(p.bar, epoch, p.trace);
    sample_phase(p, cur.S, k);       // the next sub-step's input rows are the same rows of the same CTA: no barrier needed
    cur = nxt;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(EncodeTiledFn enc, CUtensorMap* m, const void* ptr, int rows, int cols, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) B200_FAIL(B200_ERR_CUDA, "depformer: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return B200_OK;
}

Gemm plan(int N, int K, int A, int M, int grid) {
  Gemm g;
  g.wt = nullptr; g.A = A; g.N = N;
  g.n_tiles = (N + BLOCK_ROWS - 1) / BLOCK_ROWS;
  g.num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  // splits: enough units for ~2 per CTA, but the fp32 partials (written + re-read through L2) stay under ~6 MB (12 MB gated)
  long long want = (2LL * grid + g.n_tiles - 1) / g.n_tiles;
  const long long per_split = (long long)M * N * 4 * A;
  long long cap = ((long long)(A == 2 ? 12 : 6) << 20) / (per_split > 0 ? per_split : 1);
  if (cap < 1) cap = 1;
  long long S = want < cap ? want : cap;
  if (S > g.num_kb) S = g.num_kb;
  if (S < 1) S = 1;
  g.kbps = (int)((g.num_kb + S - 1) / S);
  g.S = (g.num_kb + g.kbps - 1) / g.kbps;
  return g;
}

}  // namespace

// ---- host side -------------------------------------------------------------------------------------------------------
struct DepFused {
  DepParams p;
  CUtensorMap map_xn, map_ao, map_h, map_x;
  int grid = 0; size_t smem = 0;
  void* dev_tables = nullptr;     // one allocation holding the pointer tables
};

size_t dep_fused_partial_floats(const DepFusedConfig& c) {
  const int grid = sk_num_sms();
  size_t mx = 0;
  const Gemm gs[5] = {plan(3 * c.dd, c.dd, 1, c.B, grid), plan(c.dd, c.dd, 1, c.B, grid), plan(c.F, c.dd, 2, c.B, grid),
                      plan(c.dd, c.F, 1, c.B, grid), plan(c.card, c.dd, 1, c.B, grid)};
  for (const Gemm& g : gs) {
    const size_t n = (size_t)g.S * c.B * g.N;
    if (n > mx) mx = n;
  }
  return mx;
}

int dep_fused_create(const DepFusedConfig& c, DepFused** out) {
  if (c.dd % 64 || c.dd > 4 * THREADS || c.dd / c.H != DD || c.dep_q > lm::DEP_MAX_Q || c.dep_q < 1 || c.F % 8 || c.B < 1 || c.B > 256)
    B200_FAIL(B200_ERR_INVALID, "fused depformer: unsupported configuration");
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  B200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) B200_FAIL(B200_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  DepFused* d = new DepFused();
  DepParams& p = d->p;
  memset(&p, 0, sizeof(p));
  d->grid = sk_num_sms();
  p.B = c.B; p.Mpad = ((c.B + 15) / 16) * 16; p.dd = c.dd; p.H = c.H; p.F = c.F; p.card = c.card; p.text_card = c.text_card; p.err = c.err; p.dep_q = c.dep_q; p.L = c.L;
  p.stage_bytes = (uint32_t)(2 * TILE_BYTES + p.Mpad * BLOCK_K * 2);
  int stages = (200 * 1024) / (int)p.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) { delete d; B200_FAIL(B200_ERR_INVALID, "fused depformer: batch too large for the stage ring"); }
  p.stages = stages;
  p.acc_cols = 2 * p.Mpad;
  p.n_acc = 2 * p.acc_cols <= 512 ? 2 : 1;
  uint32_t cols = (uint32_t)(p.n_acc * p.acc_cols), pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  p.tmem_cols = pow2;
  p.g_in = plan(3 * c.dd, c.dd, 1, c.B, d->grid);
  p.g_out = plan(c.dd, c.dd, 1, c.B, d->grid);
  p.g_lin_in = plan(c.F, c.dd, 2, c.B, d->grid);
  p.g_lin_out = plan(c.dd, c.F, 1, c.B, d->grid);
  p.g_head = plan(c.card, c.dd, 1, c.B, d->grid);
  // pointer tables -> one device allocation
  const size_t n_w = (size_t)c.dep_q * c.L;
  const size_t bytes = n_w * sizeof(DepLayerW) + (size_t)c.dep_q * 8 * 2 + (size_t)c.L * 8 * 4;
  std::vector<uint8_t> host(bytes);
  uint8_t* hp = host.data();
  size_t off = 0;
  auto put = [&](const void* src, size_t n) { memcpy(hp + off, src, n); const size_t o = off; off += n; return o; };
  std::vector<DepLayerW> w(n_w);
  for (size_t i = 0; i < n_w; ++i) w[i] = DepLayerW{(const uint8_t*)c.in_w[i], (const uint8_t*)c.out_w[i], (const uint8_t*)c.lin_in[i], (const uint8_t*)c.lin_out[i]};
  const size_t o_w = put(w.data(), n_w * sizeof(DepLayerW));
  const size_t o_heads = put(c.heads, (size_t)c.dep_q * 8);
  const size_t o_tables = put(c.tables, (size_t)c.dep_q * 8);
  const size_t o_n1 = put(c.n1, (size_t)c.L * 8);
  const size_t o_n2 = put(c.n2, (size_t)c.L * 8);
  const size_t o_kc = put(c.kc, (size_t)c.L * 8);
  const size_t o_vc = put(c.vc, (size_t)c.L * 8);
  if (cudaMalloc(&d->dev_tables, bytes) != cudaSuccess) { delete d; B200_FAIL(B200_ERR_CUDA, "fused depformer: cudaMalloc failed"); }
  B200_CUDA(cudaMemcpy(d->dev_tables, hp, bytes, cudaMemcpyHostToDevice));
  uint8_t* dv = static_cast<uint8_t*>(d->dev_tables);
  p.w = reinterpret_cast<const DepLayerW*>(dv + o_w);
  p.heads = reinterpret_cast<const uint8_t* const*>(dv + o_heads);
  p.tables = reinterpret_cast<const bf16* const*>(dv + o_tables);
  p.n1 = reinterpret_cast<const bf16* const*>(dv + o_n1);
  p.n2 = reinterpret_cast<const bf16* const*>(dv + o_n2);
  p.kc = reinterpret_cast<bf16* const*>(dv + o_kc);
  p.vc = reinterpret_cast<bf16* const*>(dv + o_vc);
  p.din = static_cast<const bf16*>(c.din); p.din_ld = c.din_ld;
  p.text_token = c.text_token;
  p.x = static_cast<bf16*>(c.x); p.xn = static_cast<bf16*>(c.xn); p.ao = static_cast<bf16*>(c.ao); p.hbuf = static_cast<bf16*>(c.hbuf);
  p.part0 = c.part0; p.part1 = c.part1;
  p.logits = static_cast<bf16*>(c.logits); p.audio_tokens = c.audio_tokens;
  p.noise = c.noise; p.noise_ld = c.noise_ld; p.noise_off = c.noise_off; p.ka = c.ka;
  p.use_sampling = c.use_sampling; p.top_k = c.top_k; p.temp = c.temp;
  p.bar = c.bar; p.trace = c.trace;
  B200_TRY(make_map(enc, &d->map_xn, c.xn, c.B, c.dd, p.Mpad));
  B200_TRY(make_map(enc, &d->map_ao, c.ao, c.B, c.dd, p.Mpad));
  B200_TRY(make_map(enc, &d->map_h, c.hbuf, c.B, c.F, p.Mpad));
  B200_TRY(make_map(enc, &d->map_x, c.x, c.B, c.dd, p.Mpad));
  d->smem = (size_t)stages * p.stage_bytes + 1024 + 16 * MAX_STAGES + 64;
  B200_CUDA(cudaFuncSetAttribute(dep_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  int per_sm = 0, coop = 0, dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dep_fused_kernel, THREADS, d->smem));
  if (!coop || per_sm < 1) {
    dep_fused_destroy(d);
    B200_FAIL(B200_ERR_CUDA, "fused depformer: the device cannot co-schedule %d CTAs of %zu B shared memory", d->grid, d->smem);
  }
  *out = d;
  return B200_OK;
}

void dep_fused_set_sampling(DepFused* d, int use_sampling, float temp, int top_k) {
  if (!d) return;
  d->p.use_sampling = use_sampling; d->p.temp = temp; d->p.top_k = top_k;
}

void dep_fused_destroy(DepFused* d) {
  if (!d) return;
  if (d->dev_tables) cudaFree(d->dev_tables);
  delete d;
}

// Cooperative launch: the driver guarantees that all CTAs of the grid are co-resident (or fails the launch), so the grid
// barrier cannot deadlock against another handle's persistent kernel, an MPS neighbour or a second replica on the same GPU.
int dep_fused_launch(DepFused* d, cudaStream_t stream) {
  B200_CUDA(cudaMemsetAsync(d->p.bar, 0, sizeof(unsigned), stream));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)d->grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = d->smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, dep_fused_kernel, d->map_xn, d->map_ao, d->map_h, d->map_x, d->p);
  if (le != cudaSuccess) B200_FAIL(B200_ERR_CUDA, "fused depformer: cooperative launch failed: %s", cudaGetErrorString(le));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch("dep_fused");
}

}  // namespace tc
}  // namespace b200
