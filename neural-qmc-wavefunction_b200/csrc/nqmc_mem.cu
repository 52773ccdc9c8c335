// nqmc_mem.cu -- device memory, pinned host memory and streams behind the C ABI, so a host process can drive the
// walker loop without PyTorch / XLA owning the buffers (BASELINE north_star: "no PyTorch").  These are thin,
// validated wrappers over the CUDA runtime; every hot-path entry point still takes caller-owned raw pointers, so a
// caller that already has device memory (XLA buffers through the FFI shim, torch tensors in the tests) never needs
// them.
#include "nqmc_internal.cuh"

static thread_local std::string g_mem_err;

#define MEM_CUDA(call)                                                       \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess) {                                                 \
      g_mem_err = std::string(#call) + ": " + cudaGetErrorString(_e);        \
      return NQMC_ERR_CUDA;                                                  \
    }                                                                        \
  } while (0)

extern "C" {

const char* nqmc_mem_last_error(void) { return g_mem_err.c_str(); }

int nqmc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int nqmc_dev_alloc(int device, size_t bytes, void** out) {
  if (!out) return NQMC_ERR_ARG;
  *out = nullptr;
  MEM_CUDA(cudaSetDevice(device));
  MEM_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  return NQMC_OK;
}

int nqmc_dev_free(int device, void* p) {
  if (!p) return NQMC_OK;
  MEM_CUDA(cudaSetDevice(device));
  MEM_CUDA(cudaFree(p));
  return NQMC_OK;
}

int nqmc_host_alloc(size_t bytes, void** out) {
  if (!out) return NQMC_ERR_ARG;
  *out = nullptr;
  MEM_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return NQMC_OK;
}

int nqmc_host_free(void* p) {
  if (!p) return NQMC_OK;
  MEM_CUDA(cudaFreeHost(p));
  return NQMC_OK;
}

int nqmc_dev_memcpy(int device, void* dst, const void* src, size_t bytes, int kind, nqmc_stream st) {
  if ((!dst || !src) && bytes) return NQMC_ERR_ARG;
  cudaMemcpyKind k;
  switch (kind) {
    case NQMC_COPY_H2D: k = cudaMemcpyHostToDevice; break;
    case NQMC_COPY_D2H: k = cudaMemcpyDeviceToHost; break;
    case NQMC_COPY_D2D: k = cudaMemcpyDeviceToDevice; break;
    default: g_mem_err = "unknown copy kind"; return NQMC_ERR_ARG;
  }
  MEM_CUDA(cudaSetDevice(device));
  if (bytes) MEM_CUDA(cudaMemcpyAsync(dst, src, bytes, k, reinterpret_cast<cudaStream_t>(st)));
  return NQMC_OK;
}

int nqmc_dev_copy_rows(int device, void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                       size_t rows, nqmc_stream st) {
  if (!dst || !src) return NQMC_ERR_ARG;
  MEM_CUDA(cudaSetDevice(device));
  MEM_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows, cudaMemcpyDeviceToDevice,
                             reinterpret_cast<cudaStream_t>(st)));
  return NQMC_OK;
}

int nqmc_dev_memset(int device, void* dst, int value, size_t bytes, nqmc_stream st) {
  if (!dst && bytes) return NQMC_ERR_ARG;
  MEM_CUDA(cudaSetDevice(device));
  if (bytes) MEM_CUDA(cudaMemsetAsync(dst, value, bytes, reinterpret_cast<cudaStream_t>(st)));
  return NQMC_OK;
}

int nqmc_stream_create(int device, nqmc_stream* out) {
  if (!out) return NQMC_ERR_ARG;
  MEM_CUDA(cudaSetDevice(device));
  cudaStream_t s;
  MEM_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *out = reinterpret_cast<nqmc_stream>(s);
  return NQMC_OK;
}

int nqmc_stream_wait(int device, nqmc_stream waiter, nqmc_stream signal) {
  MEM_CUDA(cudaSetDevice(device));
  cudaEvent_t ev;
  MEM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  MEM_CUDA(cudaEventRecord(ev, reinterpret_cast<cudaStream_t>(signal)));
  MEM_CUDA(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(waiter), ev, 0));
  MEM_CUDA(cudaEventDestroy(ev));            // released once the wait has been satisfied
  return NQMC_OK;
}

int nqmc_stream_destroy(int device, nqmc_stream st) {
  if (!st) return NQMC_OK;
  MEM_CUDA(cudaSetDevice(device));
  MEM_CUDA(cudaStreamDestroy(reinterpret_cast<cudaStream_t>(st)));
  return NQMC_OK;
}

int nqmc_stream_sync(int device, nqmc_stream st) {
  MEM_CUDA(cudaSetDevice(device));
  MEM_CUDA(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(st)));
  return NQMC_OK;
}

}  // extern "C"
