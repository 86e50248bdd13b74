// TEST INFRASTRUCTURE ONLY -- the sliver of the CUDA RUNTIME API the engine's host code (diffsheg_b200/csrc/engine.cu) uses,
// modelled on the host so that the whole engine -- handle, packed-weight resolution, workspace carving, the per-step launch
// sequence of Runner::denoise and every kernel it launches -- runs on the thread-level emulator of emu_cuda.h.
//   "device memory" = 1024-byte aligned host allocations; streams are ignored (every launch runs to completion before it returns,
//   which is a legal execution of a single stream); CUDA graphs report "not supported" so the engine keeps its eager path;
//   a failed emulated launch (deadlock, out-of-bounds shared-memory access, mismatched barrier ...) becomes the sticky error that
//   cudaGetLastError() returns, with the emulator's message behind emu_rt::last_launch_error().
// Nothing here is linked into libdiffsheg_b200.so.
#pragma once
#include <tuple>
#include <utility>

#include "emu_cuda.h"

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaStreamNonBlocking = 1, cudaStreamCaptureModeThreadLocal = 1, cudaErrorNotSupported = 801 };
typedef void* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 148; };

namespace emu_rt {
inline cudaError_t& sticky() { static cudaError_t e = cudaSuccess; return e; }
inline std::string& last_launch_error() { static std::string s; return s; }
inline int& num_sms() { static int n = 8; return n; }   // a small "device": persistent kernels walk several tiles per CTA

// kernel<<<grid, block, smem, stream>>>(args...) on the emulator; a kernel with STATIC shared memory passes its size as `smem`
template <typename... KA, typename... A>
inline void launch(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t, A... a) {
  std::tuple<typename std::decay<KA>::type...> args(static_cast<typename std::decay<KA>::type>(a)...);
  std::string err;
  if (grid.z != 1 || block.y != 1 || block.z != 1) { sticky() = cudaErrorInvalidValue; last_launch_error() = "emulated launches are 2-D grids of 1-D blocks"; return; }
  const bool ok = emu::run_grid(grid.x, block.x, 1, smem, [&] { std::apply(kern, args); }, &err, grid.y);
  if (!ok) { sticky() = cudaErrorLaunchFailure; last_launch_error() = err; }
}
}  // namespace emu_rt

inline const char* cudaGetErrorString(cudaError_t e) {
  if (e == cudaSuccess) return "no error";
  if (e == cudaErrorLaunchFailure) return emu_rt::last_launch_error().c_str();
  return e == cudaErrorInvalidValue ? "invalid argument" : "emulated runtime error";
}
inline cudaError_t cudaGetLastError() { const cudaError_t e = emu_rt::sticky(); emu_rt::sticky() = cudaSuccess; return e; }

template <class T> inline cudaError_t cudaMalloc(T** p, size_t bytes) {
  void* q = nullptr;
  if (posix_memalign(&q, 1024, (bytes + 1023) / 1024 * 1024 + 1024) != 0) return cudaErrorInvalidValue;
  memset(q, 0xCD, bytes);   // poison like the emulated shared memory
  *p = static_cast<T*>(q);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset2DAsync(void* p, size_t pitch, int v, size_t width, size_t height, cudaStream_t) {
  for (size_t r = 0; r < height; ++r) memset(static_cast<char*>(p) + r * pitch, v, width);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; ++r) memmove(static_cast<char*>(d) + r * dpitch, static_cast<const char*>(s) + r * spitch, width);
  return cudaSuccess;
}
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); p->multiProcessorCount = emu_rt::num_sms(); return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = emu_rt::num_sms(); return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaErrorNotSupported; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* x, cudaGraph_t, unsigned long long) { *x = nullptr; return cudaErrorNotSupported; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
