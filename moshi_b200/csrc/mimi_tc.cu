// mimi_tc_kernel: persistent TMA + tcgen05 kind::tf32 implicit GEMM with 3xTF32 split products (see mimi_tc.cuh).
#include <cstdlib>
#include <cstring>

#include "mimi_tc.cuh"
#include "tc_prims.cuh"

namespace b200 {
namespace mtc {

namespace {

using namespace tcp;

constexpr uint32_t A_TILE_BYTES = TC_ROWS * TC_KB * 4;      // 16 KB

struct Tile { int m, n, split, b0, t0, kb0, kb1; };

__device__ __forceinline__ Tile tile_at(const TcParams& p, int id) {
  Tile t;
  t.split = id % p.ksplit;
  const int mn = id / p.ksplit;
  t.n = mn % p.n_tiles_n;
  t.m = mn / p.n_tiles_n;
  const int tpg = (p.T + p.tt - 1) / p.tt;                  // tiles along the steps of one session group
  const int group = t.m / tpg, j = t.m - group * tpg;
  t.b0 = group * p.bb;
  t.t0 = j * p.tt;
  t.kb0 = t.split * p.kb_per_split;
  t.kb1 = min(p.num_kb, t.kb0 + p.kb_per_split);
  return t;
}

// Epilogue of V consecutive output features n0.. of GEMM row (b, t): shared by the kernel and the split-K reduction.
template <int V>
__device__ __forceinline__ void store_row_chunk(const TcParams& p, int b, int t, int n0, float* v) {
  if (p.epi == TC_EPI_GELU) {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = gelu_erf(v[i]);
  } else if (p.epi == TC_EPI_RES_SCALE) {                       // x + layer_scale * update (transformer.py:769,777)
    const float* r = p.res + (long long)b * p.r_sb + (long long)t * p.r_row + n0;
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      const float4 rv = *reinterpret_cast<const float4*>(r + i);
      const float4 sv = *reinterpret_cast<const float4*>(p.scale + n0 + i);
      v[i] = rv.x + sv.x * v[i]; v[i + 1] = rv.y + sv.y * v[i + 1]; v[i + 2] = rv.z + sv.z * v[i + 2]; v[i + 3] = rv.w + sv.w * v[i + 3];
    }
  } else {
    if (p.bias != nullptr) {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] += p.bias[(n0 + i) % p.bias_mod];
    }
    if (p.res != nullptr) {                                     // SEANetResnetBlock: u + v (seanet.py:90-93)
      const float* r = p.res + (long long)b * p.r_sb + (long long)t * p.r_row + n0;
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        const float4 rv = *reinterpret_cast<const float4*>(r + i);
        v[i] += rv.x; v[i + 1] += rv.y; v[i + 2] += rv.z; v[i + 3] += rv.w;
      }
    }
  }
  if (p.y != nullptr) {
    float* y = p.y + (long long)b * p.y_sb + (long long)t * p.y_row + n0;
#pragma unroll
    for (int i = 0; i < V; i += 4) *reinterpret_cast<float4*>(y + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  if (p.act_mode != TC_ACT_NONE) {
    if (p.act_elu) {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = elu1(v[i]);
    }
    const long long o = (long long)b * p.a_sb + (long long)t * p.a_row + n0;
    if (p.act_mode == TC_ACT_FULL) {
#pragma unroll
      for (int i = 0; i < V; i += 4) *reinterpret_cast<float4*>(p.a_hi + o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
      float lo[V];
#pragma unroll
      for (int i = 0; i < V; ++i) split_tf32(v[i], v[i], lo[i]);
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        *reinterpret_cast<float4*>(p.a_hi + o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        *reinterpret_cast<float4*>(p.a_lo + o + i) = make_float4(lo[i], lo[i + 1], lo[i + 2], lo[i + 3]);
      }
    }
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
mimi_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t bars = base + (uint32_t)p.stages * p.stage_bytes;
  const uint32_t full0 = bars, empty0 = bars + 8 * TC_MAX_STAGES, tfull0 = bars + 16 * TC_MAX_STAGES, tempty0 = tfull0 + 16;
  const uint32_t tptr = tempty0 + 16;
  uint32_t* tptr_generic = reinterpret_cast<uint32_t*>(smem_raw + (tptr - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.m_tiles * p.n_tiles_n * p.ksplit;
  const uint32_t w_bytes = (uint32_t)p.NT * 128u;               // one weight tile (hi or lo)
  const int cb_per_tap = p.Cin / TC_KB;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4);           // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tptr, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tptr_generic;
  pdl_trigger();                                 // the next kernel of the frame may begin its own prologue and weight prefetch

  if (warp == 0) {
    if (lane == 0) {
      // ===== producer =====
      // Programmatic dependent launch: the weights do not depend on the preceding kernel, so the first `stages` weight pairs of
      // this CTA's first tile are requested before the dependency wait (their HBM latency overlaps the predecessor's tail and
      // this kernel's own launch); the activation boxes, and everything the epilogue touches, come after it.
      int pre = 0;
      if ((int)blockIdx.x < n_tiles) {
        const Tile t0 = tile_at(p, blockIdx.x);
        const uint8_t* wsrc = p.wt + ((size_t)t0.n * p.num_kb + t0.kb0) * (size_t)(2 * w_bytes);
        for (int kb = t0.kb0; kb < t0.kb1 && pre < p.stages; ++kb, wsrc += 2 * w_bytes, ++pre) {
          mbar_expect_tx(full0 + 8 * pre, 2 * A_TILE_BYTES + 2 * w_bytes);
          bulk_load(base + (uint32_t)pre * p.stage_bytes + 2 * A_TILE_BYTES, wsrc, 2 * w_bytes, full0 + 8 * pre);
        }
      }
      pdl_wait();
      int s = 0; uint32_t ph = 0;
      int item = 0;
      for (int id = blockIdx.x; id < n_tiles; id += gridDim.x) {
        const Tile tl = tile_at(p, id);
        const uint8_t* wsrc = p.wt + ((size_t)tl.n * p.num_kb + tl.kb0) * (size_t)(2 * w_bytes);
        for (int kb = tl.kb0; kb < tl.kb1; ++kb, wsrc += 2 * w_bytes, ++item) {
          const int tap = kb / cb_per_tap, cb = kb - tap * cb_per_tap;
          const uint32_t sa = base + (uint32_t)s * p.stage_bytes;
          if (item >= pre) {
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            mbar_expect_tx(full0 + 8 * s, 2 * A_TILE_BYTES + 2 * w_bytes);
            bulk_load(sa + 2 * A_TILE_BYTES, wsrc, 2 * w_bytes, full0 + 8 * s);
          }
          const int row = p.row0 + tl.t0 * p.stride + tap * p.dil;
          tma_load_3d(sa, &map_hi, full0 + 8 * s, cb * TC_KB, row, tl.b0);
          tma_load_3d(sa + A_TILE_BYTES, &map_lo, full0 + 8 * s, cb * TC_KB, row, tl.b0);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: per k-block 4 k-steps x (hi*hi + hi*lo + lo*hi) =====
      const uint32_t idesc = make_idesc_tf32(TC_ROWS, p.NT);
      int s = 0; uint32_t ph = 0;
      int acc = 0; uint32_t acc_bits = 0u;
      for (int id = blockIdx.x; id < n_tiles; id += gridDim.x) {
        const Tile tl = tile_at(p, id);
        mbar_wait(tempty0 + 8 * acc, ((acc_bits >> acc) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * p.NT);
        for (int kb = tl.kb0; kb < tl.kb1; ++kb) {
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t a_hi = base + (uint32_t)s * p.stage_bytes, a_lo = a_hi + A_TILE_BYTES;
          const uint32_t w_hi = a_lo + A_TILE_BYTES, w_lo = w_hi + w_bytes;
#pragma unroll
          for (int k = 0; k < TC_KB / 8; ++k) {
            const uint32_t off = (uint32_t)k * 32u;          // 8 tf32 = 32 bytes of K per MMA
            const uint32_t first = (kb == tl.kb0 && k == 0) ? 0u : 1u;
            umma_tf32(d0, make_desc(a_lo + off), make_desc(w_hi + off), idesc, first);      // small terms first
            umma_tf32(d0, make_desc(a_hi + off), make_desc(w_lo + off), idesc, 1u);
            umma_tf32(d0, make_desc(a_hi + off), make_desc(w_hi + off), idesc, 1u);
          }
          umma_commit(empty0 + 8 * s);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        umma_commit(tfull0 + 8 * acc);
        acc_bits ^= 1u << acc;
        acc ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: TMEM lane = GEMM row = one (session, step) =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int bi = row / p.tt, ti = row - bi * p.tt;
    int acc = 0; uint32_t acc_bits = 0u;
    pdl_wait();                                  // outputs / residuals / the split-K workspace may still be in use by the predecessor
    for (int id = blockIdx.x; id < n_tiles; id += gridDim.x) {
      const Tile tl = tile_at(p, id);
      const int b = tl.b0 + bi, t = tl.t0 + ti;
      const bool valid = b < p.n_sessions && t < p.T;
      mbar_wait(tfull0 + 8 * acc, (acc_bits >> acc) & 1u);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.NT);
      const int n_base = tl.n * p.NT;
      for (int c0 = 0; c0 < p.NT; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(lane_addr + (uint32_t)c0, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        if (p.ksplit > 1) {
          float* w = p.ws + ((size_t)tl.split * p.m_tiles * TC_ROWS + (size_t)tl.m * TC_ROWS + row) * (size_t)p.N + n_base + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) __stcg(reinterpret_cast<float4*>(w + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        } else if (valid) {
          store_row_chunk<16>(p, b, t, n_base + c0, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      acc_bits ^= 1u << acc;
      acc ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// split-K: sum the partials in split order (deterministic) and run the epilogue; one thread per (row, 4 features)
__global__ void __launch_bounds__(256) mimi_tc_reduce_kernel(const TcParams p) {
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n4 = p.N / 4;
  const long long rows = (long long)p.m_tiles * TC_ROWS;
  if (i >= rows * n4) return;
  const long long grow = i / n4;
  const int n0 = (int)(i - grow * n4) * 4;
  const int m = (int)(grow / TC_ROWS), row = (int)(grow - (long long)m * TC_ROWS);
  const int tpg = (p.T + p.tt - 1) / p.tt;
  const int group = m / tpg, j = m - group * tpg;
  const int bi = row / p.tt, ti = row - bi * p.tt;
  const int b = group * p.bb + bi, t = j * p.tt + ti;
  if (b >= p.n_sessions || t >= p.T) return;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < p.ksplit; ++s) {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(p.ws + ((size_t)s * rows + grow) * (size_t)p.N + n0));
    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
  }
  store_row_chunk<4>(p, b, t, n0, v);
}

// load-time: w fp32 [N][n_taps * Cin] -> [n_tile][kb][hi | lo][NT rows][32] in the SWIZZLE_128B layout
// (16-byte chunk c of row r at r * 128 + ((c ^ (r & 7)) << 4)); rows beyond N are zero
__global__ void tc_pack_kernel(const float* __restrict__ w, float4* __restrict__ out, int N, int K, int NT, int n_tiles_n, int num_kb) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk (4 floats) of a hi tile
  const long long per_tile = (long long)NT * 8;
  const long long total = (long long)n_tiles_n * num_kb * per_tile;
  if (idx >= total) return;
  const int chunk = (int)(idx % per_tile);
  long long t = idx / per_tile;
  const int kb = (int)(t % num_kb);
  const int nt = (int)(t / num_kb);
  const int r = chunk >> 3, cpos = chunk & 7;
  const int csrc = cpos ^ (r & 7);
  const int n = nt * NT + r, k = kb * TC_KB + csrc * 4;
  float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
  if (n < N) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (k + i < K) split_tf32(w[(long long)n * K + k + i], hi[i], lo[i]);
  }
  const long long o = (((long long)nt * num_kb + kb) * 2) * per_tile + chunk;
  out[o] = make_float4(hi[0], hi[1], hi[2], hi[3]);
  out[o + per_tile] = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_attr = false;

}  // namespace

int tc_init() {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    B200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) B200_FAIL(B200_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  if (!g_attr) {
    B200_CUDA(cudaFuncSetAttribute(mimi_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    g_attr = true;
  }
  return B200_OK;
}

static int pick_nt(int N) { return N >= TC_MAX_NT ? TC_MAX_NT : N; }

size_t tc_packed_bytes(int N, int Cin, int n_taps) {
  const int NT = pick_nt(N), n_tiles = (N + NT - 1) / NT, num_kb = n_taps * (Cin / TC_KB);
  return (size_t)n_tiles * num_kb * 2 * NT * 128;
}

int tc_pack_weights(const float* w_dev, void* out_dev, int N, int Cin, int n_taps, cudaStream_t st) {
  if (Cin % TC_KB || N % 16 || (N > TC_MAX_NT && N % TC_MAX_NT))
    B200_FAIL(B200_ERR_SHAPE, "tc_pack_weights: Cin %d must be a multiple of 32, N %d a multiple of 16 (of 128 above 128)", Cin, N);
  const int NT = pick_nt(N), n_tiles = (N + NT - 1) / NT, num_kb = n_taps * (Cin / TC_KB);
  const long long total = (long long)n_tiles * num_kb * NT * 8;
  tc_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w_dev, static_cast<float4*>(out_dev), N, n_taps * Cin, NT, n_tiles, num_kb);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch("tc_pack");
}

int tc_make_map(CUtensorMap* m, const float* base, int Cin, long long rows_per_session, long long sb_elems, int n_sessions, int tt,
                int bb, int stride) {
  B200_TRY(tc_init());
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)rows_per_session, (cuuint64_t)n_sessions};
  cuuint64_t strides[2] = {(cuuint64_t)Cin * 4, (cuuint64_t)sb_elems * 4};
  cuuint32_t box[3] = {(cuuint32_t)TC_KB, (cuuint32_t)(tt * stride), (cuuint32_t)bb};
  cuuint32_t estr[3] = {1, (cuuint32_t)stride, 1};
  if (box[1] > 256 || box[2] > 256) B200_FAIL(B200_ERR_SHAPE, "tc_make_map: box %u x %u exceeds the TMA limit of 256", box[1], box[2]);
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    B200_FAIL(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): Cin %d rows %lld sessions %d box %u/%u/%u stride %d", (int)r, Cin,
              rows_per_session, n_sessions, box[0], box[1], box[2], stride);
  return B200_OK;
}

// Fills the launch plan of a layer for `n_sessions` sessions of T output steps each (linear: n_sessions = 1, T = tokens).
// The caller has set the static fields, the tensor maps and the epilogue pointers of L.p.
int tc_plan(TcLayer& L, int n_sessions, int T, int sms, size_t ws_bytes) {
  TcParams& p = L.p;
  p.n_sessions = n_sessions; p.T = T;
  if (p.tt * p.bb != TC_ROWS) B200_FAIL(B200_ERR_INVALID, "tc_plan: tile %d x %d is not 128 rows", p.tt, p.bb);
  const int tpg = (T + p.tt - 1) / p.tt, groups = (n_sessions + p.bb - 1) / p.bb;
  p.m_tiles = tpg * groups;
  p.n_taps = L.n_taps; p.dil = L.dil; p.stride = L.stride; p.Cin = L.Cin;
  p.wt = L.wt; p.NT = L.NT; p.n_tiles_n = L.n_tiles_n; p.num_kb = L.num_kb; p.N = L.N;
  p.bias = L.bias; p.bias_mod = L.bias_mod > 0 ? L.bias_mod : L.N;
  // deep-and-skinny layers: cut K so that the chip has a few hundred tiles, at least 4 k-blocks each
  int ks = 1;
  const long long mn = (long long)p.m_tiles * p.n_tiles_n;
  if (mn < sms && p.num_kb >= 8) {
    ks = (int)((2LL * sms + mn - 1) / mn);
    if (ks > p.num_kb / 4) ks = p.num_kb / 4;
    while (ks > 1 && (size_t)ks * p.m_tiles * TC_ROWS * (size_t)p.N * 4 > ws_bytes) --ks;
    if (ks < 1) ks = 1;
  }
  p.kb_per_split = (p.num_kb + ks - 1) / ks;
  p.ksplit = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.stage_bytes = 2 * A_TILE_BYTES + 2 * (uint32_t)p.NT * 128u;
  int stages = (200 * 1024) / (int)p.stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  p.stages = stages;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * p.NT)) cols <<= 1;
  p.tmem_cols = cols;
  const long long tiles = mn * p.ksplit;
  L.grid = (int)(tiles < sms ? tiles : sms);
  L.smem = (size_t)stages * p.stage_bytes + 1024 + 16 * TC_MAX_STAGES + 64;
  return B200_OK;
}

int tc_launch(const TcLayer& L, cudaStream_t st) {
  // programmatic dependent launch (B200_MIMI_PDL=0: off): the kernel may start while its predecessor drains; it requests its
  // first weight tiles, then waits (griddepcontrol.wait) before it reads an activation or writes anything
  static const int pdl = [] { const char* e = getenv("B200_MIMI_PDL"); return e ? atoi(e) : 1; }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)L.grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = L.smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, mimi_tc_kernel, L.map_hi, L.map_lo, L.p);
  if (le != cudaSuccess) B200_FAIL(B200_ERR_CUDA, "mimi_tc launch failed: %s", cudaGetErrorString(le));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (L.p.ksplit > 1) {
    const long long n = (long long)L.p.m_tiles * TC_ROWS * (L.p.N / 4);
    mimi_tc_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L.p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  return check_launch("mimi_tc");
}

}  // namespace mtc
}  // namespace b200
