// TEST INFRASTRUCTURE ONLY -- the sliver of the CUDA RUNTIME API the engine's host code (diffsheg_b200/csrc/engine.cu) uses,
// modelled on the host so that the whole engine -- handle, packed-weight resolution, workspace carving, the per-step launch
// sequence of Runner::denoise and every kernel it launches -- runs on the thread-level emulator of emu_cuda.h.
//   "device memory" = 1024-byte aligned host allocations; streams are ignored (every launch runs to completion before it returns,
//   which is a legal execution of a single stream); stream capture records launches / async copies / memsets with their by-value
//   arguments and cudaGraphLaunch replays them (emu_cuda.h: Capture), so the engine's graph-replay path runs too;
//   a failed emulated launch (deadlock, out-of-bounds shared-memory access, mismatched barrier ...) becomes the sticky error that
//   cudaGetLastError() returns, with the emulator's message behind emu_rt::last_launch_error().
// Nothing here is linked into libdiffsheg_b200.so.
#pragma once
#include <memory>
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
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type = cudaMemoryTypeDevice; int device = 0; };

namespace emu_rt {
inline cudaError_t& sticky() { static cudaError_t e = cudaSuccess; return e; }
inline std::string& last_launch_error() { static std::string s; return s; }
inline long long& graph_launches() { static long long n = 0; return n; }   // cudaGraphLaunch calls (replays of a captured denoiser call)
inline int& num_sms() { static int n = 8; return n; }   // a small "device": persistent kernels walk several tiles per CTA

// kernel<<<grid, block, smem, stream>>>(args...) on the emulator; a kernel with STATIC shared memory passes its size as `smem`
template <typename... KA, typename... A>
inline void launch(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t, A... a) {
  typedef std::tuple<typename std::decay<KA>::type...> Args;
  const std::shared_ptr<Args> args = std::make_shared<Args>(static_cast<typename std::decay<KA>::type>(a)...);   // owned by the (possibly recorded) launch
  std::string err;
  if (grid.z != 1 || block.y != 1 || block.z != 1) { sticky() = cudaErrorInvalidValue; last_launch_error() = "emulated launches are 2-D grids of 1-D blocks"; return; }
  const bool ok = emu::run_grid(grid.x, block.x, 1, smem, [kern, args] { std::apply(kern, *args); }, &err, grid.y);
  if (!ok) { sticky() = cudaErrorLaunchFailure; last_launch_error() = err; }
}

// a stream-ordered host-side operation (copy / memset): recorded while a capture is active, done at once otherwise
template <class F> inline cudaError_t stream_op(F f) {
  if (emu::Capture* cap = emu::active_capture()) cap->ops.push_back([f](std::string*) { f(); return true; });
  else f();
  return cudaSuccess;
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
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { return emu_rt::stream_op([=] { memset(p, v, n); }); }
inline cudaError_t cudaMemset2DAsync(void* p, size_t pitch, int v, size_t width, size_t height, cudaStream_t) {
  return emu_rt::stream_op([=] { for (size_t r = 0; r < height; ++r) memset(static_cast<char*>(p) + r * pitch, v, width); });
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { return emu_rt::stream_op([=] { memmove(d, s, n); }); }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
  return emu_rt::stream_op([=] { for (size_t r = 0; r < height; ++r) memmove(static_cast<char*>(d) + r * dpitch, static_cast<const char*>(s) + r * spitch, width); });
}
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* at, const void*) { *at = cudaPointerAttributes(); return cudaSuccess; }   // every pointer is "device 0" memory
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
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { static char token; *s = &token; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
// capture: the engine launches on nothing but its capture stream between Begin and End (single host thread), so "a capture is active"
// is all the state there is
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) {
  if (emu::active_capture()) return cudaErrorInvalidValue;
  emu::active_capture() = new emu::Capture;
  return cudaSuccess;
}
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = emu::active_capture(); emu::active_capture() = nullptr; return *g ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* x, cudaGraph_t g, unsigned long long) { *x = new emu::Capture(*static_cast<emu::Capture*>(g)); return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t x, cudaStream_t) {
  ++emu_rt::graph_launches();
  for (auto& op : static_cast<emu::Capture*>(x)->ops) {
    std::string err;
    if (!op(&err)) { emu_rt::sticky() = cudaErrorLaunchFailure; emu_rt::last_launch_error() = err; return cudaErrorLaunchFailure; }
  }
  return cudaSuccess;
}
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete static_cast<emu::Capture*>(g); return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t x) { delete static_cast<emu::Capture*>(x); return cudaSuccess; }
