// TEST INFRASTRUCTURE ONLY -- host emulation of the small slice of CUDA the SIMT kernels use.
//
// Compiling a kernel source with g++ -DDSHEG_EMU includes this header instead of <cuda_runtime.h> / <cuda_bf16.h>.
// Every CUDA thread of a thread-block cluster becomes a fiber (ucontext) inside ONE OS thread; fibers run until they
// reach a synchronising operation (warp collective, named barrier, __syncthreads, cluster barrier), where they
// rendezvous.  Scheduling is deterministic round-robin, a pass without progress is reported as a deadlock.
// What this checks: indexing, swizzles, mma / ldmatrix fragment maps, barrier protocols, DSMEM addressing, arithmetic
// (fp32 with bf16 roundings where the kernel rounds).  What it cannot check: memory-model / async-proxy ordering, bank
// conflicts, performance.  Nothing here is linked into libdiffsheg_b200.so.
#pragma once
#if !defined(__x86_64__) || defined(EMU_USE_UCONTEXT)
#include <ucontext.h>
#define EMU_UCONTEXT 1
#else
#define EMU_UCONTEXT 0
// Fiber switch without the two rt_sigprocmask system calls of glibc's swapcontext (a third of the emulator's run time): push the
// callee-saved registers, swap stack pointers, pop, return.  The fibers never touch the signal mask or the FP control words.
extern "C" void emu_ctx_switch(void** save_sp, void* load_sp);
asm(".text\n"
    ".hidden emu_ctx_switch\n"
    ".globl emu_ctx_switch\n"
    ".type emu_ctx_switch,@function\n"
    "emu_ctx_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size emu_ctx_switch, .-emu_ctx_switch\n");
#endif

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __cluster_dims__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

// ---- vector types ---------------------------------------------------------------------------------------------------
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
struct uint3 { uint32_t x, y, z; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return {x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }

// ---- bf16 (round-to-nearest-even; x is the low half of a bf16x2 word like on the device) ----------------------------------
struct __nv_bfloat16 { uint16_t bits; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  __nv_bfloat16 h;
  if ((u & 0x7fffffffu) > 0x7f800000u) { h.bits = 0x7fff; return h; }   // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  h.bits = (uint16_t)(u >> 16);
  return h;
}
static inline float __bfloat162float(__nv_bfloat16 h) {
  const uint32_t u = (uint32_t)h.bits << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline __nv_bfloat162 __floats2bfloat162_rn(float lo, float hi) { return {__float2bfloat16_rn(lo), __float2bfloat16_rn(hi)}; }
static inline __nv_bfloat16 emu_hmax(__nv_bfloat16 a, __nv_bfloat16 b) {
  const float fa = __bfloat162float(a), fb = __bfloat162float(b);
  if (fa != fa) return b;
  if (fb != fb) return a;
  return fa > fb ? a : b;
}
static inline __nv_bfloat162 __hmax2(__nv_bfloat162 a, __nv_bfloat162 b) { return {emu_hmax(a.x, b.x), emu_hmax(a.y, b.y)}; }

// ---- math ------------------------------------------------------------------------------------------------------------------
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#define __expf(x) expf(x)   // glibc declares a __expf of its own
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
template <class T> static inline T __ldg(const T* p) { return *p; }

// ---- the sliver of the runtime API the host-side launch code mentions ------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorLaunchFailure = 719 };

// ---- fibers, CTAs, clusters -----------------------------------------------------------------------------------------------
namespace emu {

struct Rendezvous {
  int expected = 0, arrived = 0;
  uint64_t gen = 0;
};
struct Warp {
  Rendezvous rv;
  alignas(16) uint8_t slot[2][32][64];   // double-buffered per-lane exchange area of the warp collectives
};
struct Cluster;
struct Cta {
  std::vector<uint8_t> smem_store;
  uint8_t* smem = nullptr;   // 1024-aligned
  size_t smem_bytes = 0;
  int nthreads = 0;
  uint3 bid{0, 0, 0};
  uint32_t rank = 0;
  Cluster* cluster = nullptr;
  Rendezvous named[16];
  std::vector<Warp> warps;
  std::shared_ptr<void> ext;   // per-CTA state of optional models (emu_tc_prims.h: mbarriers, TMEM); fresh for every CTA
};
struct Cluster {
  std::vector<Cta> ctas;
  Rendezvous rv;          // barrier.cluster (arrive + wait as one rendezvous, or split: see cluster_arrive / cluster_wait)
  uint64_t arrive_gen_seen = 0;
};
constexpr size_t kStackBytes = 192 * 1024;
struct Thread {
#if EMU_UCONTEXT
  ucontext_t ctx;
#else
  void* sp = nullptr;
#endif
  std::unique_ptr<uint8_t[]> stack;   // kStackBytes, uninitialised
  Cta* cta = nullptr;
  uint3 tid{0, 0, 0};
  int lane = 0, warp = 0;
  int parity = 0;          // exchange-slot parity of the next warp collective
  uint64_t cluster_wait_gen = 0;
  bool done = false;
  const char* waiting_on = "";
};
struct Launch {
  uint3 grid{1, 1, 1}, block{1, 1, 1};
  uint32_t cluster_size = 1;
};

struct Runtime {
#if EMU_UCONTEXT
  ucontext_t sched;
#else
  void* sched_sp = nullptr;
#endif
  Thread* cur = nullptr;
  Launch launch;
  uint64_t progress = 0;
  uint64_t grids_run = 0;   // number of run_grid calls = kernel launches the emulator has executed
  std::string error;
  std::function<void()> body;
};
inline Runtime& rt() { static Runtime r; return r; }
inline Thread& self() { return *rt().cur; }
#if EMU_UCONTEXT
inline void yield() { Thread* t = rt().cur; swapcontext(&t->ctx, &rt().sched); }
#else
inline void yield() { Thread* t = rt().cur; emu_ctx_switch(&t->sp, rt().sched_sp); }
#endif

inline void rendezvous(Rendezvous& r, int expected, const char* what) {
  if (r.arrived == 0) r.expected = expected;
  else if (r.expected != expected) { rt().error = std::string("mismatched participant count at ") + what; }
  const uint64_t g = r.gen;
  if (++r.arrived == r.expected) {
    r.arrived = 0;
    ++r.gen;
    ++rt().progress;
  } else {
    self().waiting_on = what;
    while (r.gen == g) yield();
    self().waiting_on = "";
  }
}

// exchange `bytes` (<= 64) per lane among the 32 lanes of the calling warp; returns the warp's slot array of this collective
inline uint8_t (*warp_exchange(const void* mine, int bytes, const char* what))[64] {
  Thread& t = self();
  Warp& w = t.cta->warps[t.warp];
  const int p = t.parity;
  t.parity ^= 1;
  memcpy(w.slot[p][t.lane], mine, bytes);
  rendezvous(w.rv, 32, what);
  return w.slot[p];
}

// Stream capture (emu_runtime.h: cudaStreamBeginCapture ... cudaGraphLaunch): while a capture is active, run_grid RECORDS the launch --
// grid shape and the body with its by-value kernel arguments, exactly what a CUDA graph kernel node bakes in -- instead of running it;
// replaying the recorded list re-executes it.  A host value that changes from step to step and was passed by value to a captured launch
// is therefore stale on replay, as on the device.
struct Capture { std::vector<std::function<bool(std::string*)>> ops; };
inline Capture*& active_capture() { static Capture* c = nullptr; return c; }

inline void trampoline() {
  rt().body();
  self().done = true;
  ++rt().progress;
  yield();
}

// Run `body` (a call of the kernel with its arguments) for every thread of a grid; clusters execute one after another.
// A 2-D grid (grid_y > 1; clusters of one CTA only) is walked x-fastest: blockIdx = {i % grid_x, i / grid_x, 0}.
inline bool run_grid(uint32_t grid_x, uint32_t block_x, uint32_t cluster_size, size_t smem_bytes, std::function<void()> body,
                     std::string* err, uint32_t grid_y = 1) {
  if (Capture* cap = active_capture()) {
    cap->ops.push_back([=](std::string* e) { return run_grid(grid_x, block_x, cluster_size, smem_bytes, body, e, grid_y); });
    return true;
  }
  Runtime& R = rt();
  R.body = body;
  R.error.clear();
  ++R.grids_run;
  R.launch.grid = {grid_x, grid_y, 1};
  R.launch.block = {block_x, 1, 1};
  R.launch.cluster_size = cluster_size;
  if (grid_x % cluster_size || block_x % 32) { *err = "grid / block not a multiple of the cluster size / warp size"; return false; }
  if (grid_y > 1 && cluster_size != 1) { *err = "2-D grids are emulated for clusters of one CTA only"; return false; }
  const uint32_t grid_total = grid_x * grid_y;
  const size_t nthr = (size_t)cluster_size * block_x;
  // fiber objects and their 192 KB stacks are pooled across launches (a 768-thread kernel would otherwise allocate and zero 144 MB
  // per launch); every field a launch reads is re-initialised below, the stacks need no clearing
  static std::vector<std::unique_ptr<Thread>> pool;
  while (pool.size() < nthr) { pool.emplace_back(new Thread); pool.back()->stack.reset(new uint8_t[kStackBytes]); }
  std::vector<std::unique_ptr<Thread>>& threads = pool;
  for (uint32_t c0 = 0; c0 < grid_total; c0 += cluster_size) {
    Cluster cl;
    cl.ctas.resize(cluster_size);
    for (uint32_t r = 0; r < cluster_size; ++r) {
      Cta& c = cl.ctas[r];
      c.smem_store.assign(smem_bytes + 2048, 0xCD);   // poison: reads of never-written smem show up as garbage
      c.smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(c.smem_store.data()) + 1023) & ~(uintptr_t)1023);
      c.smem_bytes = smem_bytes;
      c.nthreads = (int)block_x;
      c.bid = {(c0 + r) % grid_x, (c0 + r) / grid_x, 0};
      c.rank = r;
      c.cluster = &cl;
      c.warps.resize(block_x / 32);
      for (uint32_t i = 0; i < block_x; ++i) {
        Thread& t = *threads[(size_t)r * block_x + i];
        t.cta = &c;
        t.tid = {i, 0, 0};
        t.lane = (int)(i & 31);
        t.warp = (int)(i >> 5);
        t.parity = 0;
        t.cluster_wait_gen = 0;
        t.done = false;
        t.waiting_on = "";
#if EMU_UCONTEXT
        getcontext(&t.ctx);
        t.ctx.uc_stack.ss_sp = t.stack.get();
        t.ctx.uc_stack.ss_size = kStackBytes;
        t.ctx.uc_link = &R.sched;
        makecontext(&t.ctx, (void (*)())trampoline, 0);
#else
        {   // initial frame: six zeroed callee-saved registers, the entry point as the return address, then a null return slot at an
            // address = 8 mod 16 (what a function sees right after being called); trampoline() never returns
          uintptr_t top = (reinterpret_cast<uintptr_t>(t.stack.get()) + kStackBytes) & ~(uintptr_t)15;
          void** frame = reinterpret_cast<void**>(top - 64);
          for (int q = 0; q < 8; ++q) frame[q] = nullptr;
          frame[6] = reinterpret_cast<void*>(&trampoline);
          t.sp = frame;
        }
#endif
      }
    }
    // Scheduling order of a pass.  Any order is a legal execution (threads only rendezvous at synchronising operations), so code
    // that is correct must give the same results under all of them; EMU_SCHED=reverse runs the threads last-to-first, EMU_SCHED=shuffle
    // draws a new pseudo-random order for every pass and lets random warps sit passes out (deterministic LCG) -- the emulator's stand-in for compute-sanitizer racecheck:
    // a missing barrier between a producer and a consumer shows up as a wrong result or a NaN under at least one of the orders.
    const char* sched_env = getenv("EMU_SCHED");
    const int sched_mode = !sched_env ? 0 : (!strcmp(sched_env, "reverse") ? 1 : (!strcmp(sched_env, "shuffle") ? 2 : 0));
    std::vector<uint32_t> order(nthr);
    for (uint32_t i = 0; i < nthr; ++i) order[i] = sched_mode == 1 ? (uint32_t)(nthr - 1 - i) : i;
    uint64_t lcg = 0x9E3779B97F4A7C15ull;
    if (const char* seed_env = getenv("EMU_SCHED_SEED")) lcg ^= strtoull(seed_env, nullptr, 10) * 0xD1342543DE82EF95ull;   // other interleavings
    size_t remaining = nthr;
    while (remaining) {
      const uint64_t before = R.progress;
      remaining = 0;
      if (sched_mode == 2)
        for (size_t i = nthr - 1; i > 0; --i) {
          lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
          std::swap(order[i], order[(size_t)((lcg >> 33) % (i + 1))]);
        }
      // shuffle mode also lets every WARP sit out a pass with probability 1/2 (a fresh draw per pass), so a warp can fall arbitrarily far
      // behind its neighbours: stores before a collective of one warp vs. loads after a collective of another are then really unordered
      bool skipped = false;
      std::vector<uint8_t> sit_out;
      if (sched_mode == 2) {
        sit_out.resize((nthr + 31) / 32);
        for (auto& b : sit_out) { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; b = (uint8_t)((lcg >> 40) & 1); }
      }
      for (uint32_t oi = 0; oi < nthr; ++oi) {
        auto& tp = threads[order[oi]];
        if (tp->done) continue;
        if (sched_mode == 2 && sit_out[order[oi] / 32]) { skipped = true; ++remaining; continue; }
        R.cur = tp.get();
#if EMU_UCONTEXT
        swapcontext(&R.sched, &tp->ctx);
#else
        emu_ctx_switch(&R.sched_sp, tp->sp);
#endif
        if (!tp->done) ++remaining;
      }
      if (skipped && R.progress == before) continue;   // nobody who ran made progress, but some warps sat out: not a deadlock
      if (!R.error.empty()) { *err = R.error; return false; }
      if (remaining && R.progress == before) {
        std::string msg = "deadlock: ";
        int shown = 0;
        for (size_t ti = 0; ti < nthr; ++ti) {
          auto& tp = threads[ti];
          if (!tp->done && shown++ < 6)
            msg += "[cta " + std::to_string(tp->cta->bid.x) + " thread " + std::to_string(tp->tid.x) + " waits on " + tp->waiting_on + "] ";
        }
        *err = msg;
        return false;
      }
    }
  }
  return true;
}

}  // namespace emu

#define threadIdx (emu::self().tid)
#define blockIdx (emu::self().cta->bid)
#define blockDim (emu::rt().launch.block)
#define gridDim (emu::rt().launch.grid)

static inline void __syncthreads() { emu::rendezvous(emu::self().cta->named[0], emu::self().cta->nthreads, "__syncthreads"); }
static inline void __syncwarp(uint32_t = 0xffffffffu) {
  emu::Thread& t = emu::self();
  emu::rendezvous(t.cta->warps[t.warp].rv, 32, "__syncwarp");
}
template <class T> static inline T __shfl_xor_sync(uint32_t, T v, int lane_mask) {
  static_assert(sizeof(T) <= 8, "shuffle of a 32/64-bit value");
  const int lane = emu::self().lane;
  uint8_t(*slots)[64] = emu::warp_exchange(&v, sizeof(T), "__shfl_xor_sync");
  T r;
  memcpy(&r, slots[(lane ^ lane_mask) & 31], sizeof(T));
  return r;
}
template <class T> static inline T __shfl_sync(uint32_t, T v, int src_lane) {
  uint8_t(*slots)[64] = emu::warp_exchange(&v, sizeof(T), "__shfl_sync");
  T r;
  memcpy(&r, slots[src_lane & 31], sizeof(T));
  return r;
}
static inline int __all_sync(uint32_t, int pred) {
  uint8_t(*slots)[64] = emu::warp_exchange(&pred, sizeof(int), "__all_sync");
  int all = 1;
  for (int l = 0; l < 32; ++l) { int v; memcpy(&v, slots[l], sizeof(int)); all &= (v != 0); }
  return all;
}
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { const uint32_t old = *p; *p = old + v; ++emu::rt().progress; return old; }   // fibers never preempt
static inline void __threadfence_block() {}
static inline size_t __cvta_generic_to_shared(const void* p) {
  emu::Cta* c = emu::self().cta;
  const uint8_t* b = static_cast<const uint8_t*>(p);
  if (b < c->smem || b >= c->smem + c->smem_bytes) { emu::rt().error = "__cvta_generic_to_shared: pointer outside this CTA's shared memory"; return 0; }
  return (size_t)(b - c->smem);
}
