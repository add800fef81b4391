// The Depth transformer of one frame as ONE persistent kernel, second generation (LMGen.depformer_step, lm.py:809-850, with
// LMModel.forward_depformer lm.py:450-493 and the weights-per-step layers of transformer.py:291-318 inside).
//
// dep_fused.cu (first generation) separates every GEMM from its consumer by TWO grid barriers and an L2 round trip of fp32
// split-K partials (8 barriers per layer, ~9 us per GEMM at 104 sessions, profiles/r02_dep_trace_*).  Here the split-K
// reduction never leaves the chip and the element-wise / row-wise work rides on the GEMMs:
//
//   * grid = NC clusters x 4 CTAs (cooperative + cluster launch, one CTA per SM).  A cluster owns a 128-row weight tile (tile
//     pair for the gated MLP); rank r streams k-blocks [r*kbps, (r+1)*kbps) into its own TMEM accumulator.
//   * the four partial accumulators are reduce-scattered over distributed shared memory: rank q receives session columns
//     [q*Wc, (q+1)*Wc) of every peer (st.shared::cluster), sums them in rank order (deterministic) and runs the EPILOGUE for
//     its column block with the reference's cast points:  bf16 store (in_proj -> qkv, head -> logits) | residual add
//     x = bf16(x + bf16(acc)) + per-tile sum of squares of the new x (out_proj, linear_out) | h = bf16(bf16(silu(g)) * u).
//   * RMSNorm is folded into the STAGING of the consuming GEMM: a CTA builds its own UMMA B operand in shared memory from x,
//     alpha and r[m] = rsqrt(eps + sum_tiles ssq[tile][m] / dd) (SWIZZLE_128B layout written by hand), no xn buffer, no norm
//     phase.  GEMMs that read ao / h / x unnormalised take their activation boxes by 2-D TMA as before.
//   * per layer: in_proj | attention | out_proj | linear_in | linear_out = 5 grid barriers (8 before), 32 per sub-step.
//   * weights are immutable: one producer thread per CTA walks the CTA's frame-long list of 16 KB weight tiles and keeps a
//     shared-memory ring full across phase boundaries (blocking only for the unit being computed), so HBM latency is hidden
//     behind the barriers and the other clusters' GEMMs; consecutive GEMMs are mapped to disjoint clusters where they fit.
//
// Sessions: 1 <= B <= 128 (receive buffers for B > 128 do not fit beside the rings: dep_fused.cu serves 129..256).
#include "gemm_sk.cuh"
#include "tc_prims.cuh"
#include "lm_kernels.cuh"

namespace b200 {
namespace tc {

namespace {

using lm::bf16;
using namespace tcp;

constexpr int BLOCK_ROWS = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int THREADS = 256;                 // warp 0 weight producer, 1 MMA, 2-5 epilogue, 6 activation TMA; 2-7 staging
constexpr int TILE_BYTES = BLOCK_ROWS * BLOCK_K * 2;
constexpr int CS = 4;                        // CTAs per cluster = K-splits of every GEMM
constexpr int MAX_WST = 12, MAX_XST = 8;
constexpr int DD = 64;                       // depformer head dim
constexpr int MAX_MPAD = 128;
static_assert(THREADS == lm::SAMPLE_THREADS, "sample_row needs a SAMPLE_THREADS-wide CTA");

struct Gm {                                  // one GEMM of the schedule
  const uint8_t* wt;                         // packed tiles [n_tile][kb][A][16 KB]
  int n_tiles, num_kb, kbps, A, N, off;      // kbps k-blocks per rank; N = output rows; off = first cluster of tile 0
};

struct DcLayerW { const uint8_t *in_w, *out_w, *lin_in, *lin_out; };

struct DcParams {
  int B, Mpad, Wc, ldr, dd, H, F, card, text_card, dep_q, L, NC, n_ssq;
  int* err;
  int wst, xst; uint32_t xkb, xregion, tmem_cols;
  Gm g_in, g_out, g_lin_in, g_lin_out, g_head;
  const DcLayerW* w;                                         // [dep_q][L]
  const uint8_t* const* heads;                               // [dep_q]
  const bf16* const* tables;                                 // [dep_q] embedding tables ([0] = text)
  const bf16* const* n1; const bf16* const* n2;              // [L] RMSNorm alphas
  const bf16* din; long long din_ld;                         // depformer_in_all output [B][dep_q*dd]
  const long long* text_token;                               // [B]
  bf16 *x, *qkv, *ao, *hbuf;                                 // [B][dd], [B][3*dd], [B][dd], [B][F]
  bf16* const* kc; bf16* const* vc;                          // [L] per-frame KV [B][H][dep_q][64]
  float* ssq;                                                // [n_ssq][B] sum of squares of x per 128-column tile
  bf16* logits;                                              // [dep_q][B][card]
  long long* audio_tokens;                                   // [dep_q][B]
  const float* noise; long long noise_ld; int noise_off, ka;
  int use_sampling, top_k; float temp;
  unsigned* bar;
  unsigned long long* trace;
};

enum { EPI_QKV = 0, EPI_RESADD = 1, EPI_GATE = 2, EPI_LOGITS = 3 };

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0u;
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ float ldcg_bf(const bf16* p) {
  return bf2f(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}

__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& epoch, unsigned long long* trace) {
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

// ---- the frame-long list of weight tiles of one CTA --------------------------------------------------------------------
__device__ __forceinline__ Gm gemm_at(const DcParams& p, int k, int l, int which) {
  Gm g;
  if (which == 4) { g = p.g_head; g.wt = p.heads[k]; return g; }
  const DcLayerW w = p.w[k * p.L + l];
  if (which == 0) { g = p.g_in; g.wt = w.in_w; }
  else if (which == 1) { g = p.g_out; g.wt = w.out_w; }
  else if (which == 2) { g = p.g_lin_in; g.wt = w.lin_in; }
  else { g = p.g_lin_out; g.wt = w.lin_out; }
  return g;
}
__device__ __forceinline__ int first_tile(const Gm& g, int cl, int NC) {
  int t = cl - g.off % NC;
  return t < 0 ? t + NC : t;
}
__device__ __forceinline__ int gemm_index(int L, int k, int l, int which) { return k * (4 * L + 1) + (which == 4 ? 4 * L : 4 * l + which); }

struct WCursor { int k, l, which, tile, kb, kb1, a, valid; Gm g; };

// moves the cursor to the next (GEMM, tile) of cluster `cl` in which rank `rank` has k-blocks; tile < 0 = "first of this GEMM"
__device__ void cursor_seek(const DcParams& p, WCursor& c, int cl, int rank) {
  while (c.k < p.dep_q) {
    c.g = gemm_at(p, c.k, c.l, c.which);
    if (c.tile < 0) c.tile = first_tile(c.g, cl, p.NC);
    if (c.tile < c.g.n_tiles) {
      const int kb0 = rank * c.g.kbps;
      c.kb1 = min(c.g.num_kb, kb0 + c.g.kbps);
      if (kb0 < c.kb1) { c.kb = kb0; c.a = 0; c.valid = 1; return; }
    }
    c.tile = -1;
    if (c.which == 4) { c.which = 0; c.l = 0; c.k += 1; }
    else if (c.which == 3) { if (c.l + 1 < p.L) { c.l += 1; c.which = 0; } else c.which = 4; }
    else c.which += 1;
  }
  c.valid = 0;
}

struct Ctx {
  uint32_t wbase, xbase, full_w0, empty_w0, full_x0, empty_x0, tfull, xraw, tmem_base;
  uint8_t* xgen;                      // generic pointer to the X region (= DSMEM receive buffer after the MMAs)
  float* rs; float* red;              // [MAX_MPAD] r[m] of the folded RMSNorm; [4][32] cross-warp sums
  int cl, rank;
  int ws; uint32_t wph;               // MMA thread: weight ring cursor
  int xs; uint32_t xph;               // activation ring cursor (every thread keeps it)
  uint32_t tph;                       // parity of the accumulator-ready barrier
  uint32_t rph;                       // parity of the raw-x barrier (GEMMs with a folded RMSNorm)
  int p_ws; uint32_t p_wph;           // producer thread: weight ring cursor
  WCursor cur;                        // producer thread: next weight tile to request
};

// Producer (one thread): requests weight tiles in list order.  Tiles of the unit being computed are waited for (their
// ring slots free up as this CTA's own MMAs retire); tiles of later units are requested only while a slot is free.
__device__ void produce(const DcParams& p, Ctx& c, int cur_gemm, int cur_tile) {
  while (c.cur.valid) {
    const bool mine = gemm_index(p.L, c.cur.k, c.cur.l, c.cur.which) == cur_gemm && c.cur.tile == cur_tile;
    const uint32_t eb = c.empty_w0 + 8 * c.p_ws;
    if (mine) mbar_wait(eb, c.p_wph ^ 1u);
    else if (!mbar_test(eb, c.p_wph ^ 1u)) break;
    const uint32_t fb = c.full_w0 + 8 * c.p_ws;
    const uint8_t* src = c.cur.g.wt + (((size_t)c.cur.tile * c.cur.g.num_kb + c.cur.kb) * c.cur.g.A + c.cur.a) * TILE_BYTES;
    mbar_expect_tx(fb, TILE_BYTES);
    bulk_load(c.wbase + (uint32_t)c.p_ws * TILE_BYTES, src, TILE_BYTES, fb);
    if (++c.p_ws == p.wst) { c.p_ws = 0; c.p_wph ^= 1u; }
    if (++c.cur.a == c.cur.g.A) {
      c.cur.a = 0;
      if (++c.cur.kb == c.cur.kb1) { c.cur.tile += p.NC; cursor_seek(p, c.cur, c.cl, c.rank); }
    }
  }
}

// ---- one GEMM phase ------------------------------------------------------------------------------------------------------
// xmap != nullptr: the activations are read as they are (2-D TMA boxes);  xmap == nullptr: the input is rmsnorm(x, alpha),
// built in shared memory by warps 2-7 from x and the per-tile sums of squares of the producing GEMM.
template <int EPI>
__device__ void gemm_phase(const DcParams& p, Ctx& c, const Gm& g, int gidx, const CUtensorMap* xmap, const CUtensorMap* rawmap,
                           const bf16* alpha, bf16* y, long long ldy) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb0 = c.rank * g.kbps, kb1 = min(g.num_kb, kb0 + g.kbps);
  const int n_kb = kb1 > kb0 ? kb1 - kb0 : 0;
  int tile = first_tile(g, c.cl, p.NC);
  if (tile >= g.n_tiles) {                        // no tile for this cluster: keep the weight ring full and go to the barrier
    if (threadIdx.x == 0) produce(p, c, -1, -1);
    __syncwarp();
    return;
  }
  bool rs_ready = false;
  for (; tile < g.n_tiles; tile += p.NC) {
    if (warp == 0) {
      if (lane == 0) produce(p, c, gidx, tile);
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0 && n_kb > 0) {
        const uint32_t idesc = make_idesc(BLOCK_ROWS, p.Mpad);
        int xs = c.xs; uint32_t xph = c.xph;
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(c.full_x0 + 8 * xs, xph);
          const uint32_t sb = c.xbase + (uint32_t)xs * p.xkb;
          for (int a = 0; a < g.A; ++a) {
            mbar_wait(c.full_w0 + 8 * c.ws, c.wph);
            tc_fence_after();
            const uint32_t sa = c.wbase + (uint32_t)c.ws * TILE_BYTES;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16(c.tmem_base + (uint32_t)(a * p.Mpad), make_desc(sa + k * UMMA_K * 2), make_desc(sb + k * UMMA_K * 2), idesc,
                        (kb == kb0 && k == 0) ? 0u : 1u);
            umma_commit(c.empty_w0 + 8 * c.ws);
            if (++c.ws == p.wst) { c.ws = 0; c.wph ^= 1u; }
          }
          umma_commit(c.empty_x0 + 8 * xs);
          if (++xs == p.xst) { xs = 0; xph ^= 1u; }
        }
      }
      if (lane == 0) umma_commit(c.tfull);          // (a rank without k-blocks: nothing outstanding, arrives at once)
      __syncwarp();
    } else {
      if (xmap != nullptr) {
        if (warp == 6 && lane == 0) {
          fence_async_proxy();                      // the activations were written with generic stores by other CTAs
          int xs = c.xs; uint32_t xph = c.xph;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(c.empty_x0 + 8 * xs, xph ^ 1u);
            mbar_expect_tx(c.full_x0 + 8 * xs, p.xkb);
            tma_load_2d(c.xbase + (uint32_t)xs * p.xkb, xmap, c.full_x0 + 8 * xs, kb * BLOCK_K, 0);
            if (++xs == p.xst) { xs = 0; xph ^= 1u; }
          }
        }
        __syncwarp();
      } else {
        // xn = bf16(x * (alpha * r)), r = rsqrt(eps + mean x^2)  (transformer.py:45-58), built in place as the UMMA B operand:
        // warp 6 brings the raw x boxes of this rank's k-range in by TMA (rows >= B and columns >= dd arrive as zeros), then
        // warps 2-7 scale every 16-byte chunk where it lies: the chunk at position pp of row m holds columns
        // (kb*64 + ((pp ^ (m & 7)) << 3)) .. +7 (SWIZZLE_128B)
        const int t2 = threadIdx.x - 64;
        if (warp == 6 && lane == 0 && n_kb > 0) {
          fence_async_proxy();                      // x was written with generic stores by other CTAs
          mbar_expect_tx(c.xraw, (uint32_t)n_kb * p.xkb);
          for (int i = 0; i < n_kb; ++i) {
            int st = c.xs + i;
            if (st >= p.xst) st -= p.xst;
            tma_load_2d(c.xbase + (uint32_t)st * p.xkb, rawmap, c.xraw, (kb0 + i) * BLOCK_K, 0);
          }
        }
        __syncwarp();
        if (!rs_ready) {
          if (t2 < p.Mpad) {
            float tot = 0.f;
            if (t2 < p.B) {
              for (int i0 = 0; i0 < p.n_ssq; i0 += 8) {      // the loads of a group are independent: one L2 round trip, not eight
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = i0 + i < p.n_ssq ? __ldcg(p.ssq + (long long)(i0 + i) * p.B + t2) : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) tot += v[i];
              }
            }
            c.rs[t2] = t2 < p.B ? rsqrtf(1e-8f + tot / (float)p.dd) : 0.f;
          }
          named_bar(2, 192);
          rs_ready = true;
        }
        if (n_kb > 0) {
          mbar_wait(c.xraw, c.rph);
          const int per_row = n_kb * 8, total = p.B * per_row;
          for (int ci = t2; ci < total; ci += 192) {
            const int m = ci / per_row, j = ci - m * per_row;
            const int i = j >> 3, pp = j & 7;
            const int k = (kb0 + i) * BLOCK_K + ((pp ^ (m & 7)) << 3);
            if (k < p.dd) {
              int st = c.xs + i;
              if (st >= p.xst) st -= p.xst;
              uint4* cell = reinterpret_cast<uint4*>(c.xgen + (size_t)st * p.xkb + (size_t)m * 128 + (pp << 4));
              const uint4 xv = *cell;
              const uint4 av = *reinterpret_cast<const uint4*>(alpha + k);
              const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xv);
              const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&av);
              const float r = c.rs[m];
              __nv_bfloat162 o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 xf = __bfloat1622float2(xh[e]), af = __bfloat1622float2(ah[e]);
                o[e] = __floats2bfloat162_rn(xf.x * (af.x * r), xf.y * (af.y * r));
              }
              *cell = *reinterpret_cast<const uint4*>(o);
            }
          }
          fence_async_proxy();                      // generic-proxy writes -> visible to the tensor core's async-proxy reads
          named_bar(2, 192);
          if (t2 == 0) {
            for (int i = 0; i < n_kb; ++i) {
              int st = c.xs + i;
              if (st >= p.xst) st -= p.xst;
              mbar_arrive(c.full_x0 + 8 * st);
            }
          }
        }
      }
    }
    if (xmap == nullptr && n_kb > 0) c.rph ^= 1u;
    // every thread: advance the activation ring cursor past this unit
    {
      int adv = c.xs + n_kb;
      while (adv >= p.xst) { adv -= p.xst; c.xph ^= 1u; }
      c.xs = adv;
    }

    // ---- reduce-scatter of the four partial accumulators over DSMEM, epilogue on the owned session columns ----
    // (Wc is a multiple of 8: every TMEM access is a batch of 8-column loads behind ONE wait; every global load a thread
    // needs is issued before the cluster barriers it can overlap with)
    const bool epi_warp = warp >= 2 && warp < 6;
    const int q = warp & 3;                          // TMEM lane quadrant of an epilogue warp
    const int row = q * 32 + lane;
    const uint32_t lane_addr = c.tmem_base + ((uint32_t)(q * 32) << 16);
    const int n = tile * BLOCK_ROWS + row;
    const bool n_ok = n < g.N;
    uint32_t resp[16];                               // x_orig of the owned columns, two bf16 per register
    if (epi_warp) {
      if (EPI == EPI_RESADD) {                       // in flight across the barriers below
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          unsigned lo = 0u, hi = 0u;
          const int m = c.rank * p.Wc + j;
          if (j < p.Wc && n_ok) {
            if (m < p.B) lo = __ldcg(reinterpret_cast<const unsigned short*>(p.x + (long long)m * p.dd + n));
            if (m + 1 < p.B) hi = __ldcg(reinterpret_cast<const unsigned short*>(p.x + (long long)(m + 1) * p.dd + n));
          }
          resp[j >> 1] = lo | (hi << 16);
        }
      }
      mbar_wait(c.tfull, c.tph);
      tc_fence_after();
    }
    c.tph ^= 1u;
    cluster_sync_all();                              // every rank's MMAs have retired: the X regions are free to receive
    if (epi_warp) {
      for (int pr = 0; pr < CS; ++pr) {
        if (pr == c.rank) continue;
        const int slot = c.rank < pr ? c.rank : c.rank - 1;
        for (int a = 0; a < g.A; ++a) {
          const uint32_t dst = map_to_rank(c.xbase + (uint32_t)(((slot * 2 + a) * BLOCK_ROWS + row) * p.ldr) * 4u, (uint32_t)pr);
          uint32_t r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
          if (n_kb > 0) {
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 8)
              if (c0 < p.Wc) tmem_ld8(lane_addr + (uint32_t)(a * p.Mpad + pr * p.Wc + c0), r + c0);
            tmem_ld_wait();
          }
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 4)
            if (c0 < p.Wc) st_cluster_v4(dst + (uint32_t)c0 * 4u, r[c0], r[c0 + 1], r[c0 + 2], r[c0 + 3]);
        }
      }
    }
    cluster_sync_all();                              // every partial has landed in its owner's receive buffer
    if (epi_warp) {
      const float* recv = reinterpret_cast<const float*>(c.xgen);
#pragma unroll
      for (int hb = 0; hb < 32; hb += 16) {          // the owned column block in halves of 16 (register pressure)
        if (hb < p.Wc) {
          float acc[2][16];
#pragma unroll
          for (int a = 0; a < 2; ++a) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[a][j] = 0.f;
            if (a < g.A) {
              uint32_t own[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) own[j] = 0u;
              if (n_kb > 0) {
#pragma unroll
                for (int c0 = 0; c0 < 16; c0 += 8)
                  if (hb + c0 < p.Wc) tmem_ld8(lane_addr + (uint32_t)(a * p.Mpad + c.rank * p.Wc + hb + c0), own + c0);
                tmem_ld_wait();
              }
              for (int r = 0; r < CS; ++r) {          // rank order: the sum does not depend on which rank does it
                if (r == c.rank) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) acc[a][j] += __uint_as_float(own[j]);
                } else {
                  const int slot = r < c.rank ? r : r - 1;
                  const float* src = recv + (size_t)((slot * 2 + a) * BLOCK_ROWS + row) * p.ldr + hb;
#pragma unroll
                  for (int c0 = 0; c0 < 16; c0 += 4) {
                    if (hb + c0 < p.Wc) {
                      const float4 v = *reinterpret_cast<const float4*>(src + c0);
                      acc[a][c0] += v.x; acc[a][c0 + 1] += v.y; acc[a][c0 + 2] += v.z; acc[a][c0 + 3] += v.w;
                    }
                  }
                }
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float sqv = 0.f;
            const int m = c.rank * p.Wc + hb + j;
            if (hb + j < p.Wc && m < p.B && n_ok) {
              if (EPI == EPI_QKV || EPI == EPI_LOGITS) {
                y[(long long)m * ldy + n] = f2bf(acc[0][j]);
              } else if (EPI == EPI_RESADD) {
                const unsigned short rb16 = (unsigned short)((resp[(hb + j) >> 1] >> (((hb + j) & 1) * 16)) & 0xFFFFu);
                const float v = rbf(bf2f(__ushort_as_bfloat16(rb16)) + rbf(acc[0][j]));   // x_orig + update, both bf16 (transformer.py:769,777)
                p.x[(long long)m * p.dd + n] = f2bf(v);
                sqv = v * v;
              } else {
                const float gt = rbf(acc[0][j]), u = rbf(acc[1][j]);
                y[(long long)m * ldy + n] = f2bf(rbf(gt / (1.f + expf(-gt))) * u);         // gating.py:18-20
              }
            }
            if (EPI == EPI_RESADD && hb + j < p.Wc) {
              // sum of squares of the new x over this tile's 128 columns, per owned session (rows = threads): warp, then quadrant order
              const float sv = warp_sum(sqv);
              if (lane == 0) c.red[q * 32 + hb + j] = sv;
            }
          }
        }
      }
      if (EPI == EPI_RESADD) {
        named_bar(3, 128);
        const int t = threadIdx.x - 64;
        if (t < p.Wc) {
          const int m = c.rank * p.Wc + t;
          if (m < p.B) __stcg(p.ssq + (long long)tile * p.B + m, c.red[t] + c.red[32 + t] + c.red[64 + t] + c.red[96 + t]);
        }
      }
      tc_fence_before();
    }
    __syncthreads();                                 // the receive buffer is the next unit's activation region
  }
}

// attention of one sub-step: one warp per (b, h); k, v appended to the per-frame cache (no positional embedding)
__device__ void attn_phase(const DcParams& p, int layer, int step) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.dd;
  bf16* kc = p.kc[layer];
  bf16* vc = p.vc[layer];
  for (int it = blockIdx.x * (THREADS / 32) + warp; it < p.B * p.H; it += gridDim.x * (THREADS / 32)) {
    const int b = it / p.H, h = it - b * p.H;
    const unsigned* base = reinterpret_cast<const unsigned*>(p.qkv + (long long)b * 3 * C + h * DD + 2 * lane);
    const unsigned qr = __ldcg(base), kr = __ldcg(base + C / 2), vr = __ldcg(base + C);
    const long long rowo = ((long long)b * p.H + h) * p.dep_q;
    *reinterpret_cast<unsigned*>(kc + (rowo + step) * DD + 2 * lane) = kr;
    *reinterpret_cast<unsigned*>(vc + (rowo + step) * DD + 2 * lane) = vr;
    const float2 qf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&qr));
    // all keys and values of the frame so far in flight at once (<= 16 independent loads instead of 2*(step+1) dependent ones)
    unsigned kraw[lm::DEP_MAX_Q], vraw[lm::DEP_MAX_Q];
#pragma unroll
    for (int j = 0; j < lm::DEP_MAX_Q; ++j) {
      kraw[j] = j < step ? *reinterpret_cast<const unsigned*>(kc + (rowo + j) * DD + 2 * lane) : (j == step ? kr : 0u);
      vraw[j] = j < step ? *reinterpret_cast<const unsigned*>(vc + (rowo + j) * DD + 2 * lane) : (j == step ? vr : 0u);
    }
    float sc[lm::DEP_MAX_Q];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < lm::DEP_MAX_Q; ++j) {
      if (j <= step) {
        const float2 kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&kraw[j]));
        const float d = warp_sum(qf.x * kv.x + qf.y * kv.y) * 0.125f;
        sc[j] = d;
        mx = fmaxf(mx, d);
      } else sc[j] = -INFINITY;
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < lm::DEP_MAX_Q; ++j) {
      if (j <= step) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
    }
    const float inv = 1.f / sum;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int j = 0; j < lm::DEP_MAX_Q; ++j) {
      if (j <= step) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vraw[j]));
        a0 = fmaf(sc[j] * inv, v.x, a0);
        a1 = fmaf(sc[j] * inv, v.y, a1);
      }
    }
    *reinterpret_cast<__nv_bfloat162*>(p.ao + (long long)b * C + h * DD + 2 * lane) = __floats2bfloat162_rn(a0, a1);
  }
}

// rows m = blockIdx.x, += gridDim.x:  token of sub-step k-1 from its logits, then the input row of sub-step k:
// x = depformer_in[k](transformer_out) + emb_k(prev)   (lm.py:475-486; token -1 -> zero row) and its sum of squares
__device__ void sample_input_phase(const DcParams& p, int k, float* red) {
  for (int m = blockIdx.x; m < p.B; m += gridDim.x) {
    if (k > 0) {
      lm::sample_row(p.logits + ((long long)(k - 1) * p.B + m) * p.card, p.noise + (long long)m * p.noise_ld + p.noise_off + (long long)(k - 1) * p.ka,
                     p.audio_tokens + (long long)(k - 1) * p.B + m, p.card, p.use_sampling, p.temp, p.top_k);
      __syncthreads();
    }
    if (k < p.dep_q) {
      const long long id = k == 0 ? p.text_token[m] : p.audio_tokens[(long long)(k - 1) * p.B + m];
      const bool ok = lm::embed_id_ok(id, k == 0 ? p.text_card : p.card, p.err);
      const bf16* table = p.tables[k];
      float ss = 0.f;
      for (int j = threadIdx.x; j < p.dd; j += THREADS) {
        const float e = ok ? bf2f(table[id * p.dd + j]) : 0.f;
        const float v = rbf(bf2f(p.din[(long long)m * p.din_ld + (long long)k * p.dd + j]) + e);
        p.x[(long long)m * p.dd + j] = f2bf(v);
        ss += v * v;
      }
      const float tot = block_sum(ss, red);
      if (threadIdx.x < p.n_ssq) __stcg(p.ssq + (long long)threadIdx.x * p.B + m, threadIdx.x == 0 ? tot : 0.f);
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
dep_cluster_kernel(const __grid_constant__ CUtensorMap map_ao, const __grid_constant__ CUtensorMap map_h,
                   const __grid_constant__ CUtensorMap map_x, const DcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_rs[MAX_MPAD];
  __shared__ float s_red[128];
  __shared__ float s_bsum[THREADS / 32];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  Ctx c;
  c.wbase = base;
  c.xbase = base + (uint32_t)p.wst * TILE_BYTES;
  const uint32_t bars = c.xbase + p.xregion;
  c.full_w0 = bars; c.empty_w0 = bars + 8 * MAX_WST;
  c.full_x0 = bars + 16 * MAX_WST; c.empty_x0 = c.full_x0 + 8 * MAX_XST;
  c.tfull = c.empty_x0 + 8 * MAX_XST;
  c.xraw = c.tfull + 8;
  const uint32_t tptr = c.xraw + 8;
  uint32_t* tptr_generic = reinterpret_cast<uint32_t*>(smem_raw + (tptr - raw));
  c.xgen = smem_raw + (c.xbase - raw);
  c.rs = s_rs; c.red = s_red;
  c.rank = (int)cluster_ctarank();
  c.cl = blockIdx.x / CS;
  c.ws = 0; c.wph = 0u; c.xs = 0; c.xph = 0u; c.tph = 0u; c.rph = 0u; c.p_ws = 0; c.p_wph = 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ao) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_h) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < MAX_WST; ++s) { mbar_init(c.full_w0 + 8 * s, 1); mbar_init(c.empty_w0 + 8 * s, 1); }
    for (int s = 0; s < MAX_XST; ++s) { mbar_init(c.full_x0 + 8 * s, 1); mbar_init(c.empty_x0 + 8 * s, 1); }
    mbar_init(c.tfull, 1);
    mbar_init(c.xraw, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tptr, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = *tptr_generic;
  c.cur.k = 0; c.cur.l = 0; c.cur.which = 0; c.cur.tile = -1; c.cur.valid = 0;
  if (threadIdx.x == 0) {
    cursor_seek(p, c.cur, c.cl, c.rank);
    produce(p, c, -1, -1);                           // the first weight tiles are on their way before the first input row is built
  }
  __syncwarp();
  cluster_sync_all();                                // no DSMEM traffic before every CTA of the cluster has started

  unsigned epoch = 0;
  if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.trace[0] = global_timer_ns();
  for (int k = 0; k < p.dep_q; ++k) {
    sample_input_phase(p, k, s_bsum);
    grid_sync(p.bar, epoch, p.trace);
    for (int l = 0; l < p.L; ++l) {
      const int gi = gemm_index(p.L, k, l, 0);
      Gm g = gemm_at(p, k, l, 0);
      gemm_phase<EPI_QKV>(p, c, g, gi, nullptr, &map_x, p.n1[l], p.qkv, 3LL * p.dd);
      grid_sync(p.bar, epoch, p.trace);
      if (threadIdx.x == 0) produce(p, c, -1, -1);
      __syncwarp();
      attn_phase(p, l, k);
      grid_sync(p.bar, epoch, p.trace);
      g = gemm_at(p, k, l, 1);
      gemm_phase<EPI_RESADD>(p, c, g, gi + 1, &map_ao, nullptr, nullptr, nullptr, 0);
      grid_sync(p.bar, epoch, p.trace);
      g = gemm_at(p, k, l, 2);
      gemm_phase<EPI_GATE>(p, c, g, gi + 2, nullptr, &map_x, p.n2[l], p.hbuf, (long long)p.F);
      grid_sync(p.bar, epoch, p.trace);
      g = gemm_at(p, k, l, 3);
      gemm_phase<EPI_RESADD>(p, c, g, gi + 3, &map_h, nullptr, nullptr, nullptr, 0);
      grid_sync(p.bar, epoch, p.trace);
    }
    // depformer_norms is Identity (lm.py:197-198): the head reads x itself
    const Gm gh = gemm_at(p, k, 0, 4);
    gemm_phase<EPI_LOGITS>(p, c, gh, gemm_index(p.L, k, 0, 4), &map_x, nullptr, nullptr, p.logits + (long long)k * p.B * p.card, (long long)p.card);
    grid_sync(p.bar, epoch, p.trace);
    if (threadIdx.x == 0) produce(p, c, -1, -1);
    __syncwarp();
  }
  sample_input_phase(p, p.dep_q, s_bsum);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(c.tmem_base, p.tmem_cols);
  cluster_sync_all();                                // no CTA of the cluster exits while a peer could still address its shared memory
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

Gm plan(int N, int K, int A, int off) {
  Gm g;
  g.wt = nullptr; g.A = A; g.N = N; g.off = off;
  g.n_tiles = (N + BLOCK_ROWS - 1) / BLOCK_ROWS;
  g.num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  g.kbps = (g.num_kb + CS - 1) / CS;
  return g;
}

}  // namespace

struct DepCluster {
  DcParams p;
  CUtensorMap map_ao, map_h, map_x;
  int grid = 0; size_t smem = 0; int coop = 1;
  void* dev_tables = nullptr;
};

// 0 = this configuration runs on the cluster kernel; otherwise why not (the caller falls back to dep_fused / the launch chain)
static const char* dep_cluster_unsupported(const DepFusedConfig& c) {
  if (c.B < 1 || c.B > MAX_MPAD) return "1..128 sessions";
  if (c.dd % 64 || c.dd > 4096 || c.dd / c.H != DD) return "depformer dim (multiple of 64, head dim 64)";
  if (c.dep_q < 1 || c.dep_q > lm::DEP_MAX_Q || c.L < 1) return "dep_q / layers";
  if (c.F % 8 || c.dd % 8) return "k-extents must be multiples of 8";
  if ((c.dd + 127) / 128 > THREADS) return "too many column tiles";
  return nullptr;
}

int dep_cluster_create(const DepFusedConfig& c, DepCluster** out) {
  *out = nullptr;
  if (const char* why = dep_cluster_unsupported(c)) B200_FAIL(B200_ERR_INVALID, "cluster depformer: unsupported configuration (%s)", why);
  if (!c.qkv || !c.ssq) B200_FAIL(B200_ERR_INVALID, "cluster depformer: qkv / ssq buffers missing");
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  B200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) B200_FAIL(B200_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  DepCluster* d = new DepCluster();
  DcParams& p = d->p;
  memset(&p, 0, sizeof(p));
  p.B = c.B; p.Mpad = ((c.B + 31) / 32) * 32; p.Wc = p.Mpad / CS;          // Wc: a multiple of 8
  p.ldr = p.Wc + 4;
  if (((p.ldr / 4) & 1) == 0) p.ldr += 4;              // row stride = odd multiple of 16 bytes: conflict-free float4 rows
  p.dd = c.dd; p.H = c.H; p.F = c.F; p.card = c.card; p.text_card = c.text_card; p.err = c.err; p.dep_q = c.dep_q; p.L = c.L;
  p.n_ssq = (c.dd + BLOCK_ROWS - 1) / BLOCK_ROWS;
  p.xkb = (uint32_t)p.Mpad * 128u;
  const int kb_norm = ((c.dd + BLOCK_K - 1) / BLOCK_K + CS - 1) / CS;          // k-blocks per rank of the GEMMs with a folded RMSNorm
  const size_t recv_bytes = (size_t)(CS - 1) * 2 * BLOCK_ROWS * p.ldr * 4;
  int xst = kb_norm > 4 ? kb_norm : 4;
  if (xst > MAX_XST) { delete d; B200_FAIL(B200_ERR_INVALID, "cluster depformer: depformer dim too large for the activation ring"); }
  size_t xregion = (size_t)xst * p.xkb;
  if (xregion < recv_bytes) xregion = recv_bytes;
  xregion = (xregion + 1023) & ~(size_t)1023;
  xst = (int)(xregion / p.xkb);
  if (xst > MAX_XST) xst = MAX_XST;
  const size_t budget = 216 * 1024;
  int wst = xregion < budget ? (int)((budget - xregion) / TILE_BYTES) : 0;
  if (wst > MAX_WST) wst = MAX_WST;
  if (wst < 4) { delete d; B200_FAIL(B200_ERR_INVALID, "cluster depformer: no room for the weight ring"); }
  p.wst = wst; p.xst = xst; p.xregion = (uint32_t)xregion;
  uint32_t pow2 = 32;
  while (pow2 < (uint32_t)(2 * p.Mpad)) pow2 <<= 1;
  p.tmem_cols = pow2;
  d->smem = (size_t)wst * TILE_BYTES + xregion + 1024 + 16 * MAX_WST + 16 * MAX_XST + 96;
  B200_CUDA(cudaFuncSetAttribute(dep_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem));
  // how many clusters of 4 are co-resident (one CTA per SM by shared-memory size)
  if (const char* e = getenv("B200_DEP_COOP")) d->coop = atoi(e);
  int dev = 0, coop = 0, sms = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)(sms / CS * CS)); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = d->smem;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int max_clusters = 0;
  const cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, dep_cluster_kernel, &cfg);
  if (oe != cudaSuccess || !coop || max_clusters < 8) {
    cudaGetLastError();
    const size_t smem = d->smem;
    delete d;
    B200_FAIL(B200_ERR_CUDA, "cluster depformer: the device cannot co-schedule clusters of %d CTAs with %zu B shared memory (%d)", CS, smem, max_clusters);
  }
  int NC = max_clusters;
  if (NC > sms / CS) NC = sms / CS;
  if (const char* e = getenv("B200_DEP_CLUSTERS")) { const int v = atoi(e); if (v >= 1 && v < NC) NC = v; }
  p.NC = NC; d->grid = NC * CS;
  // consecutive GEMMs on disjoint clusters where they fit: in_proj / linear_in from cluster 0, out_proj / linear_out / head behind in_proj's tiles
  p.g_in = plan(3 * c.dd, c.dd, 1, 0);
  const int off2 = p.g_in.n_tiles % NC;
  p.g_out = plan(c.dd, c.dd, 1, off2);
  p.g_lin_in = plan(c.F, c.dd, 2, 0);
  p.g_lin_out = plan(c.dd, c.F, 1, off2);
  p.g_head = plan(c.card, c.dd, 1, off2);
  const size_t n_w = (size_t)c.dep_q * c.L;
  const size_t bytes = n_w * sizeof(DcLayerW) + (size_t)c.dep_q * 8 * 2 + (size_t)c.L * 8 * 4;
  std::vector<uint8_t> host(bytes);
  uint8_t* hp = host.data();
  size_t off = 0;
  auto put = [&](const void* src, size_t n) { memcpy(hp + off, src, n); const size_t o = off; off += n; return o; };
  std::vector<DcLayerW> w(n_w);
  for (size_t i = 0; i < n_w; ++i) w[i] = DcLayerW{(const uint8_t*)c.in_w[i], (const uint8_t*)c.out_w[i], (const uint8_t*)c.lin_in[i], (const uint8_t*)c.lin_out[i]};
  const size_t o_w = put(w.data(), n_w * sizeof(DcLayerW));
  const size_t o_heads = put(c.heads, (size_t)c.dep_q * 8);
  const size_t o_tables = put(c.tables, (size_t)c.dep_q * 8);
  const size_t o_n1 = put(c.n1, (size_t)c.L * 8);
  const size_t o_n2 = put(c.n2, (size_t)c.L * 8);
  const size_t o_kc = put(c.kc, (size_t)c.L * 8);
  const size_t o_vc = put(c.vc, (size_t)c.L * 8);
  if (cudaMalloc(&d->dev_tables, bytes) != cudaSuccess) { delete d; B200_FAIL(B200_ERR_CUDA, "cluster depformer: cudaMalloc failed"); }
  B200_CUDA(cudaMemcpy(d->dev_tables, hp, bytes, cudaMemcpyHostToDevice));
  uint8_t* dv = static_cast<uint8_t*>(d->dev_tables);
  p.w = reinterpret_cast<const DcLayerW*>(dv + o_w);
  p.heads = reinterpret_cast<const uint8_t* const*>(dv + o_heads);
  p.tables = reinterpret_cast<const bf16* const*>(dv + o_tables);
  p.n1 = reinterpret_cast<const bf16* const*>(dv + o_n1);
  p.n2 = reinterpret_cast<const bf16* const*>(dv + o_n2);
  p.kc = reinterpret_cast<bf16* const*>(dv + o_kc);
  p.vc = reinterpret_cast<bf16* const*>(dv + o_vc);
  p.din = static_cast<const bf16*>(c.din); p.din_ld = c.din_ld;
  p.text_token = c.text_token;
  p.x = static_cast<bf16*>(c.x); p.qkv = static_cast<bf16*>(c.qkv); p.ao = static_cast<bf16*>(c.ao); p.hbuf = static_cast<bf16*>(c.hbuf);
  p.ssq = c.ssq;
  p.logits = static_cast<bf16*>(c.logits); p.audio_tokens = c.audio_tokens;
  p.noise = c.noise; p.noise_ld = c.noise_ld; p.noise_off = c.noise_off; p.ka = c.ka;
  p.use_sampling = c.use_sampling; p.top_k = c.top_k; p.temp = c.temp;
  p.bar = c.bar; p.trace = c.trace;
  B200_TRY(make_map(enc, &d->map_ao, c.ao, c.B, c.dd, p.Mpad));
  B200_TRY(make_map(enc, &d->map_h, c.hbuf, c.B, c.F, p.Mpad));
  B200_TRY(make_map(enc, &d->map_x, c.x, c.B, c.dd, p.Mpad));
  *out = d;
  return B200_OK;
}

bool dep_cluster_supported(const DepFusedConfig& c) { return dep_cluster_unsupported(c) == nullptr; }

void dep_cluster_set_sampling(DepCluster* d, int use_sampling, float temp, int top_k) {
  if (!d) return;
  d->p.use_sampling = use_sampling; d->p.temp = temp; d->p.top_k = top_k;
}

void dep_cluster_destroy(DepCluster* d) {
  if (!d) return;
  if (d->dev_tables) cudaFree(d->dev_tables);
  delete d;
}

// Cooperative cluster launch: the driver guarantees that every cluster of the grid is co-resident (or fails the launch), so
// the grid barrier cannot deadlock against another handle's persistent kernel or a neighbour on the same GPU.
int dep_cluster_launch(DepCluster* d, cudaStream_t stream) {
  B200_CUDA(cudaMemsetAsync(d->p.bar, 0, sizeof(unsigned), stream));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)d->grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = d->smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = d->coop ? 2 : 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, dep_cluster_kernel, d->map_ao, d->map_h, d->map_x, d->p);
  if (le != cudaSuccess) B200_FAIL(B200_ERR_CUDA, "cluster depformer: launch failed: %s", cudaGetErrorString(le));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch("dep_cluster");
}

int dep_cluster_info(const DepCluster* d, int* clusters, int* w_stages, int* x_stages) {
  if (!d) return B200_ERR_INVALID;
  if (clusters) *clusters = d->p.NC;
  if (w_stages) *w_stages = d->p.wst;
  if (x_stages) *x_stages = d->p.xst;
  return B200_OK;
}

}  // namespace tc
}  // namespace b200
