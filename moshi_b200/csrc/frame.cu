// One dialogue frame for every session slot of a GPU, host buffers in and out, one host synchronisation.
//
// This is the per-frame body of the reference's callers, moved below the C ABI so that the masks, the
// Mimi -> LM -> Mimi hand-offs and the "is this row's output ready" decision stay on the device:
//   * moshi/moshi/server.py:120-147 (recv_loop: chunk -> mimi.encode -> lm_gen.step -> mimi.decode -> .cpu()),
//     where `main_pcm.cpu()` and `tokens[0, 0, 0].item()` are two blocking D2H reads per frame (:82-86);
//   * rust/moshi-server/batched_asr.py:138-215 (ASRService.step: per-slot update flags -> reset / exec masks ->
//     encode -> step -> host copies), the batched form the Rust server binds through py_basr_module.rs.
// Everything is enqueued on the stream both handles stream on; the only wait is the one before the host reads.
#include <algorithm>
#include <string>

#include "common.cuh"

namespace b200 {
namespace {

// Slot updates -> reset / exec masks (batched_asr.py:146-177 builds them on the host and copies three tensors).
__global__ void frame_masks_kernel(const int32_t* __restrict__ updates, uint8_t* __restrict__ active, uint8_t* __restrict__ reset,
                                   uint8_t* __restrict__ exec, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int u = updates[b];
  uint8_t a = active[b];
  uint8_t r = 0;
  if (u == B200_SLOT_IDLE) a = 0;
  else if (u == B200_SLOT_ACTIVE) a = 1;
  else if (u == B200_SLOT_RESET) { a = 1; r = 1; }
  // u > 0: an end-of-stream marker (batched_asr.py:161-170) leaves the slot's activity as it is
  active[b] = a;
  reset[b] = r;
  exec[b] = a;
}

// LMGen.step output -> the decoder's codes and its exec mask.  A row whose output is not ready yet (-2, lm.py:779-782:
// fewer than max_delay + 1 steps since its reset, or not executing) must not advance the decoder's streaming state:
// the reference server simply does not call mimi.decode until step() stops returning None (server.py:139-142).
__global__ void frame_decode_inputs_kernel(const int64_t* __restrict__ tokens, const uint8_t* __restrict__ exec, int64_t* __restrict__ codes,
                                           uint8_t* __restrict__ dec_exec, uint8_t* __restrict__ ready, int B, int dep_q) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t* t = tokens + (size_t)b * (dep_q + 1);
  bool ok = exec[b] != 0;
  for (int k = 0; k <= dep_q; ++k) ok = ok && t[k] >= 0;
  for (int k = 0; k < dep_q; ++k) codes[(size_t)b * dep_q + k] = ok ? t[1 + k] : 0;
  dec_exec[b] = ok ? 1 : 0;
  ready[b] = ok ? 1 : 0;
}

// PCM of rows without a decoded frame is silence, not whatever the masked decoder left in the buffer.
__global__ void frame_silence_kernel(float* __restrict__ pcm, const uint8_t* __restrict__ ready, int frame) {
  if (ready[blockIdx.x]) return;
  for (int i = threadIdx.x; i < frame; i += blockDim.x) pcm[(size_t)blockIdx.x * frame + i] = 0.f;
}

}  // namespace
}  // namespace b200

using namespace b200;

struct b200_frame {
  b200_mimi* mimi = nullptr;
  b200_lm* lm = nullptr;
  int B = 0, dep_q = 0, n_in = 0, frame = 0, noise_row = 0;
  cudaStream_t stream = nullptr;
  // device
  float* pcm_in = nullptr; float* pcm_out = nullptr; float* noise = nullptr;
  int64_t* codes_in = nullptr; int64_t* tokens = nullptr; int64_t* codes_out = nullptr;
  int32_t* updates = nullptr;
  uint8_t* active = nullptr; uint8_t* reset = nullptr; uint8_t* exec = nullptr; uint8_t* dec_exec = nullptr; uint8_t* ready = nullptr;
  // pinned host staging
  float* h_pcm_in = nullptr; float* h_pcm_out = nullptr; float* h_noise = nullptr;
  int64_t* h_tokens = nullptr; int32_t* h_updates = nullptr; uint8_t* h_ready = nullptr;
  cudaEvent_t done = nullptr;
  std::vector<uint8_t> h_active;     // host mirror of `active` (only to skip a frame in which no slot executes)
};

extern "C" {

int b200_frame_create(b200_mimi* mimi, b200_lm* lm, int batch, int n_codebooks, int dep_q, int frame_size, void* stream,
                      b200_frame** out) {
  if (!mimi || !lm || !out || batch <= 0 || n_codebooks <= 0 || dep_q <= 0 || frame_size <= 0)
    B200_FAIL(B200_ERR_INVALID, "frame_create: bad arguments");
  if (n_codebooks != dep_q)
    B200_FAIL(B200_ERR_SHAPE, "frame_create: the decoder takes the dep_q generated codebooks (%d), the codec has %d", dep_q, n_codebooks);
  auto* f = new b200_frame();
  f->mimi = mimi; f->lm = lm; f->B = batch; f->dep_q = dep_q; f->n_in = n_codebooks; f->frame = frame_size;
  f->stream = static_cast<cudaStream_t>(stream);
  f->noise_row = b200_lm_noise_per_row(lm);
  f->h_active.assign((size_t)batch, 1);
  const size_t B = (size_t)batch;
  B200_CUDA(cudaMalloc(&f->pcm_in, B * frame_size * 4));
  B200_CUDA(cudaMalloc(&f->pcm_out, B * frame_size * 4));
  B200_CUDA(cudaMalloc(&f->noise, B * (size_t)std::max(f->noise_row, 1) * 4));
  B200_CUDA(cudaMalloc(&f->codes_in, B * n_codebooks * 8));
  B200_CUDA(cudaMalloc(&f->tokens, B * (dep_q + 1) * 8));
  B200_CUDA(cudaMalloc(&f->codes_out, B * dep_q * 8));
  B200_CUDA(cudaMalloc(&f->updates, B * 4));
  B200_CUDA(cudaMalloc(&f->active, B));
  B200_CUDA(cudaMalloc(&f->reset, B));
  B200_CUDA(cudaMalloc(&f->exec, B));
  B200_CUDA(cudaMalloc(&f->dec_exec, B));
  B200_CUDA(cudaMalloc(&f->ready, B));
  B200_CUDA(cudaMemsetAsync(f->active, 1, B, f->stream));          // streaming() starts with every row executing (streaming.py:31-33)
  B200_CUDA(cudaMemsetAsync(f->exec, 1, B, f->stream));
  B200_CUDA(cudaMemsetAsync(f->reset, 0, B, f->stream));
  B200_CUDA(cudaMallocHost(&f->h_pcm_in, B * frame_size * 4));
  B200_CUDA(cudaMallocHost(&f->h_pcm_out, B * frame_size * 4));
  B200_CUDA(cudaMallocHost(&f->h_noise, B * (size_t)std::max(f->noise_row, 1) * 4));
  B200_CUDA(cudaMallocHost(&f->h_tokens, B * (dep_q + 1) * 8));
  B200_CUDA(cudaMallocHost(&f->h_updates, B * 4));
  B200_CUDA(cudaMallocHost(&f->h_ready, B));
  B200_CUDA(cudaEventCreateWithFlags(&f->done, cudaEventDisableTiming));
  *out = f;
  return B200_OK;
}

int b200_frame_destroy(b200_frame* f) {
  if (!f) return B200_OK;
  cudaFree(f->pcm_in); cudaFree(f->pcm_out); cudaFree(f->noise); cudaFree(f->codes_in); cudaFree(f->tokens); cudaFree(f->codes_out);
  cudaFree(f->updates); cudaFree(f->active); cudaFree(f->reset); cudaFree(f->exec); cudaFree(f->dec_exec); cudaFree(f->ready);
  cudaFreeHost(f->h_pcm_in); cudaFreeHost(f->h_pcm_out); cudaFreeHost(f->h_noise); cudaFreeHost(f->h_tokens);
  cudaFreeHost(f->h_updates); cudaFreeHost(f->h_ready);
  if (f->done) cudaEventDestroy(f->done);
  delete f;
  return B200_OK;
}

int b200_frame_read_buffer(b200_frame* f, const char* name, void* dst_dev, int64_t capacity_bytes, int64_t* nbytes) {
  if (!f || !name) B200_FAIL(B200_ERR_INVALID, "frame_read_buffer: null argument");
  const std::string n = name;
  const size_t B = (size_t)f->B;
  const void* src = nullptr; int64_t sz = 0;
  if (n == "codes_in") { src = f->codes_in; sz = (int64_t)(B * f->n_in * 8); }
  else if (n == "codes_out") { src = f->codes_out; sz = (int64_t)(B * f->dep_q * 8); }
  else if (n == "tokens") { src = f->tokens; sz = (int64_t)(B * (f->dep_q + 1) * 8); }
  else if (n == "exec") { src = f->exec; sz = (int64_t)B; }
  else if (n == "decoder_exec") { src = f->dec_exec; sz = (int64_t)B; }
  else B200_FAIL(B200_ERR_INVALID, "frame_read_buffer: unknown buffer '%s'", name);
  if (nbytes) *nbytes = sz;
  if (!dst_dev) return B200_OK;
  if (capacity_bytes < sz) B200_FAIL(B200_ERR_SHAPE, "frame_read_buffer: destination too small");
  B200_CUDA(cudaMemcpyAsync(dst_dev, src, (size_t)sz, cudaMemcpyDeviceToDevice, f->stream));
  return B200_OK;
}

int b200_frame_step(b200_frame* f, const float* pcm_in_host, const int32_t* updates_host, const float* noise_host,
                    const float* noise_dev, float* pcm_out_host, int64_t* tokens_out_host, uint8_t* ready_out_host) {
  if (!f || !pcm_in_host || !pcm_out_host || !tokens_out_host) B200_FAIL(B200_ERR_INVALID, "frame_step: null buffers");
  const size_t B = (size_t)f->B;
  cudaStream_t st = f->stream;
  memcpy(f->h_pcm_in, pcm_in_host, B * f->frame * 4);
  B200_CUDA(cudaMemcpyAsync(f->pcm_in, f->h_pcm_in, B * f->frame * 4, cudaMemcpyHostToDevice, st));
  if (noise_host && f->noise_row > 0) {
    memcpy(f->h_noise, noise_host, B * f->noise_row * 4);
    B200_CUDA(cudaMemcpyAsync(f->noise, f->h_noise, B * f->noise_row * 4, cudaMemcpyHostToDevice, st));
  }
  if (updates_host) {
    bool any_reset = false, all_idle = true;
    for (size_t b = 0; b < B; ++b) {
      const int32_t u = updates_host[b];
      if (u < B200_SLOT_RESET) B200_FAIL(B200_ERR_INVALID, "frame_step: unknown slot update %d for slot %zu", (int)u, b);
      any_reset |= u == B200_SLOT_RESET;
      if (u == B200_SLOT_IDLE) f->h_active[b] = 0;
      else if (u < 0) f->h_active[b] = 1;
      all_idle &= f->h_active[b] == 0;
      f->h_updates[b] = u;
    }
    B200_CUDA(cudaMemcpyAsync(f->updates, f->h_updates, B * 4, cudaMemcpyHostToDevice, st));
    B200_LAUNCH(frame_masks_kernel, (unsigned)((B + 127) / 128), 128, 0, st, f->updates, f->active, f->reset, f->exec, (int)B);
    B200_TRY(check_launch("frame_masks"));
    if (any_reset) {                  // batched_asr.py:179-181
      B200_TRY(b200_lm_reset(f->lm, f->reset));
      B200_TRY(b200_mimi_reset(f->mimi, f->reset));
    }
    if (all_idle) {                   // skip_exec (batched_asr.py:183-184): nothing runs, nothing is ready
      B200_CUDA(cudaStreamSynchronize(st));
      if (ready_out_host) memset(ready_out_host, 0, B);
      memset(pcm_out_host, 0, B * f->frame * 4);
      for (size_t i = 0; i < B * (f->dep_q + 1); ++i) tokens_out_host[i] = -2;      // ungenerated_token_id (lm.py:779-782)
      return B200_OK;
    }
  }
  B200_TRY(b200_lm_set_exec_mask(f->lm, f->exec));
  B200_TRY(b200_mimi_set_exec_mask(f->mimi, f->exec));
  B200_TRY(b200_mimi_encode(f->mimi, f->pcm_in, 1, f->codes_in));
  int ready_all = 0;
  B200_TRY(b200_lm_step(f->lm, f->codes_in, f->n_in, noise_host ? f->noise : noise_dev, f->tokens, /*support_out_of_sync=*/1, &ready_all));
  B200_LAUNCH(frame_decode_inputs_kernel, (unsigned)((B + 127) / 128), 128, 0, st, f->tokens, f->exec, f->codes_out, f->dec_exec, f->ready,
              (int)B, f->dep_q);
  B200_TRY(check_launch("frame_decode_inputs"));
  B200_TRY(b200_mimi_set_exec_mask(f->mimi, f->dec_exec));
  B200_TRY(b200_mimi_decode(f->mimi, f->codes_out, f->dep_q, 1, f->pcm_out));
  B200_LAUNCH(frame_silence_kernel, (unsigned)B, 256, 0, st, f->pcm_out, f->ready, f->frame);
  B200_TRY(check_launch("frame_silence"));
  B200_CUDA(cudaMemcpyAsync(f->h_tokens, f->tokens, B * (f->dep_q + 1) * 8, cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaMemcpyAsync(f->h_pcm_out, f->pcm_out, B * f->frame * 4, cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaMemcpyAsync(f->h_ready, f->ready, B, cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaEventRecord(f->done, st));
  B200_CUDA(cudaEventSynchronize(f->done));           // the frame's only host wait
  memcpy(tokens_out_host, f->h_tokens, B * (f->dep_q + 1) * 8);
  memcpy(pcm_out_host, f->h_pcm_out, B * f->frame * 4);
  if (ready_out_host) memcpy(ready_out_host, f->h_ready, B);
  return B200_OK;
}

}  // extern "C"
