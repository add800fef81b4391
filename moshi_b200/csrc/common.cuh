// Shared helpers for the moshi_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/moshi_b200.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define B200_CUDA(expr)                                                                      \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      b200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

#define B200_TRY(expr)             \
  do {                             \
    int _rc = (expr);              \
    if (_rc != B200_OK) return _rc; \
  } while (0)

#define B200_FAIL(code, ...)      \
  do {                            \
    b200::set_error(__VA_ARGS__); \
    return code;                  \
  } while (0)

// Every kernel launch goes through this so that `gpu_launches` is a count, not a guess.
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...)                    \
  do {                                                                         \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                \
    b200::g_launches.fetch_add(1, std::memory_order_relaxed);                  \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// named tensor store (the state-dict side of the ABI)
// ---------------------------------------------------------------------------------------------
struct Tensor {
  void* data = nullptr;   // device, owned by the store
  int dtype = B200_F32;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

inline size_t dtype_size(int dtype) {
  switch (dtype) {
    case B200_F32: return 4;
    case B200_BF16: return 2;
    case B200_F16: return 2;
    case B200_I64: return 8;
    case B200_U8: return 1;
    case B200_I8: return 1;
  }
  return 0;
}

struct TensorStore {
  std::map<std::string, Tensor> items;
  int put(const char* name, const void* dev, int dtype, int ndim, const int64_t* shape);
  const Tensor* find(const std::string& name) const;
  void release(const std::string& name);
  void release_all();
};

// Device allocation bookkeeping for a handle: everything freed in one place.
struct StateEntry {
  std::string name;        // handle-local name; the Python shims map these onto the reference's per-module State fields
  void* ptr; size_t bytes; int dtype; std::vector<int64_t> shape;
};
struct Arena {
  std::vector<void*> ptrs;
  // buffers that make up the streaming state (get/set_streaming_state, streaming.py:158-181), in registration order
  std::vector<StateEntry> snap;
  void mark_state(void* p, size_t bytes, const std::string& name = "", int dtype = B200_U8, std::vector<int64_t> shape = {}) {
    if (!p || !bytes) return;
    if (shape.empty()) shape = {(int64_t)bytes};
    snap.push_back({name, p, bytes, dtype, shape});
  }
  const StateEntry* find_state(const std::string& name) const {
    for (auto& e : snap) if (e.name == name) return &e;
    return nullptr;
  }
  size_t state_bytes() const;                 // each segment padded to 256 B
  int save(void* dst, cudaStream_t st) const;
  int load(const void* src, cudaStream_t st) const;
  int alloc(void** out, size_t bytes, bool zero = true);
  template <typename T>
  int alloc_t(T** out, size_t count, bool zero = true) {
    return alloc(reinterpret_cast<void**>(out), count * sizeof(T), zero);
  }
  void free_all();
};

// Makes the handle's device current for the duration of an entry point (the reference's modules carry their device; a caller
// whose current device is another GPU must still be able to drive this handle).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// shared body of b200_{lm,mimi}_state_{count,entry,read,write}
int state_entry_info(const Arena& a, int index, const char** name, int* dtype, int* ndim, int64_t* shape8, int64_t* nbytes);
int state_entry_copy(const Arena& a, const char* name, void* dst_dev, const void* src_dev, int64_t nbytes, cudaStream_t st);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }
// round-trip through bf16: the reference materialises most intermediates as bf16 tensors
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

}  // namespace b200
