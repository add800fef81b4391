#include "common.cuh"

namespace b200 {

static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int TensorStore::put(const char* name, const void* dev, int dtype, int ndim, const int64_t* shape) {
  if (!name || !dev || ndim < 0 || ndim > 8) B200_FAIL(B200_ERR_INVALID, "load_tensor: bad arguments");
  size_t es = dtype_size(dtype);
  if (es == 0) B200_FAIL(B200_ERR_INVALID, "load_tensor(%s): unknown dtype %d", name, dtype);
  Tensor t;
  t.dtype = dtype;
  t.shape.assign(shape, shape + ndim);
  size_t bytes = (size_t)t.numel() * es;
  B200_CUDA(cudaMalloc(&t.data, bytes ? bytes : 1));
  B200_CUDA(cudaMemcpy(t.data, dev, bytes, cudaMemcpyDeviceToDevice));
  auto it = items.find(name);
  if (it != items.end()) {
    cudaFree(it->second.data);
    items.erase(it);
  }
  items[name] = t;
  return B200_OK;
}

const Tensor* TensorStore::find(const std::string& name) const {
  auto it = items.find(name);
  return it == items.end() ? nullptr : &it->second;
}

void TensorStore::release(const std::string& name) {
  auto it = items.find(name);
  if (it != items.end()) {
    cudaFree(it->second.data);
    items.erase(it);
  }
}

void TensorStore::release_all() {
  for (auto& kv : items) cudaFree(kv.second.data);
  items.clear();
}

int Arena::alloc(void** out, size_t bytes, bool zero) {
  *out = nullptr;
  B200_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  ptrs.push_back(*out);
  if (zero) B200_CUDA(cudaMemset(*out, 0, bytes));
  return B200_OK;
}

void Arena::free_all() {
  for (void* p : ptrs) cudaFree(p);
  ptrs.clear();
  snap.clear();
}

static size_t pad256(size_t n) { return (n + 255) / 256 * 256; }

size_t Arena::state_bytes() const {
  size_t n = 0;
  for (auto& s : snap) n += pad256(s.bytes);
  return n;
}

int Arena::save(void* dst, cudaStream_t st) const {
  size_t off = 0;
  for (auto& s : snap) {
    B200_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + off, s.ptr, s.bytes, cudaMemcpyDeviceToDevice, st));
    off += pad256(s.bytes);
  }
  return B200_OK;
}

int Arena::load(const void* src, cudaStream_t st) const {
  size_t off = 0;
  for (auto& s : snap) {
    B200_CUDA(cudaMemcpyAsync(s.ptr, static_cast<const char*>(src) + off, s.bytes, cudaMemcpyDeviceToDevice, st));
    off += pad256(s.bytes);
  }
  return B200_OK;
}

int state_entry_info(const Arena& a, int index, const char** name, int* dtype, int* ndim, int64_t* shape8, int64_t* nbytes) {
  if (index < 0 || index >= (int)a.snap.size()) B200_FAIL(B200_ERR_INVALID, "state_entry: index %d out of range", index);
  const StateEntry& e = a.snap[index];
  if (name) *name = e.name.c_str();
  if (dtype) *dtype = e.dtype;
  if (ndim) *ndim = (int)e.shape.size();
  if (shape8) for (size_t i = 0; i < e.shape.size() && i < 8; ++i) shape8[i] = e.shape[i];
  if (nbytes) *nbytes = (int64_t)e.bytes;
  return B200_OK;
}

int state_entry_copy(const Arena& a, const char* name, void* dst_dev, const void* src_dev, int64_t nbytes, cudaStream_t st) {
  const StateEntry* e = a.find_state(name ? name : "");
  if (!e) B200_FAIL(B200_ERR_INVALID, "state entry '%s' does not exist", name ? name : "(null)");
  if (nbytes != (int64_t)e->bytes) B200_FAIL(B200_ERR_SHAPE, "state entry '%s' holds %zu bytes, caller passed %lld", e->name.c_str(), e->bytes, (long long)nbytes);
  if (dst_dev) B200_CUDA(cudaMemcpyAsync(dst_dev, e->ptr, e->bytes, cudaMemcpyDeviceToDevice, st));
  else if (src_dev) B200_CUDA(cudaMemcpyAsync(e->ptr, src_dev, e->bytes, cudaMemcpyDeviceToDevice, st));
  else B200_FAIL(B200_ERR_INVALID, "state entry copy: null buffer");
  return B200_OK;
}

}  // namespace b200

extern "C" {
const char* b200_last_error(void) { return b200::g_err; }
int b200_abi_version(void) { return B200_ABI_VERSION; }
int64_t b200_launch_count(void) { return b200::g_launches.load(); }
}
