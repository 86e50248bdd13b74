// TEST INFRASTRUCTURE ONLY -- host bodies of diffsheg_b200/csrc/simt_prims.cuh (same names, same signatures), modelling
// the PTX ISA's documented semantics:
//   ldmatrix.m8n8.x4[.trans].b16 : lane i supplies the address of row i%8 of matrix i/8; lane (g = lane/4, q = lane%4)
//                                  receives word q of row g of each matrix (.trans: elements [2q][g], [2q+1][g])
//   mma.m16n8k16.row.col bf16    : A regs {(g, 2q..), (g+8, 2q..), (g, 2q+8..), (g+8, 2q+8..)}, B regs {(k=2q.., n=g), (k=2q+8.., n=g)},
//                                  C/D {(g, 2q), (g, 2q+1), (g+8, 2q), (g+8, 2q+1)}
// Shared-memory "addresses" are byte offsets into the CTA's emulated window; cluster addresses carry the target rank.
#pragma once
#include <unordered_map>
#include <vector>

#include "emu_cuda.h"

#define DSHEG_DYN_SMEM(name, align) uint8_t* name = emu::self().cta->smem

namespace dsheg {
namespace prims {

// dynamic operation counts (per LANE; the emulator is one OS thread, so plain counters): what a kernel variant executes per sample
// on the pipes that bound the attention kernels -- MUFU (ex2 / tanh / rcp), tensor (mma), shared-memory matrix loads, cp.async
struct OpCounters { unsigned long long cp_async16, ldsm, mma, ex2, tanh, rcp, packed_fp32; };
inline OpCounters& op_counters() { static OpCounters c{}; return c; }

inline uint8_t* smem_ptr(uint32_t addr, size_t bytes, const char* what) {
  emu::Cta* c = emu::self().cta;
  if ((size_t)addr + bytes > c->smem_bytes) {
    emu::rt().error = std::string(what) + ": shared-memory access out of bounds (offset " + std::to_string(addr) + ")";
    return c->smem;   // keep going on valid memory; the launch reports the error
  }
  return c->smem + addr;
}
inline uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// cp.async is ASYNCHRONOUS: a copy is only guaranteed to have landed when the issuing thread's cp.async.wait_group / wait_all
// returns.  The model performs it at that LATEST legal moment (per-thread commit groups, oldest first), so a tile that is read before
// the wait -- or by another thread before a barrier that follows the owner's wait -- shows its stale contents and fails parity.
struct PendingCopy { uint32_t dst; const void* src; };
struct AsyncCopies { std::vector<PendingCopy> open; std::vector<std::vector<PendingCopy>> groups; };
inline std::unordered_map<emu::Thread*, AsyncCopies>& async_copies() {
  static std::unordered_map<emu::Thread*, AsyncCopies> m;
  return m;
}
inline void cp_async16(uint32_t dst, const void* src) {
  ++op_counters().cp_async16;
  if (dst & 15u) emu::rt().error = "cp.async 16: misaligned shared destination";
  if (reinterpret_cast<uintptr_t>(src) & 15u) emu::rt().error = "cp.async 16: misaligned global source";
  smem_ptr(dst, 16, "cp.async");   // bounds check at issue
  async_copies()[&emu::self()].open.push_back(PendingCopy{dst, src});
}
inline void cp_async_land(const std::vector<PendingCopy>& g) {
  for (const PendingCopy& c : g) memcpy(smem_ptr(c.dst, 16, "cp.async"), c.src, 16);
}
inline void cp_async_commit() {
  AsyncCopies& a = async_copies()[&emu::self()];
  a.groups.push_back(std::move(a.open));
  a.open.clear();
}
template <int N> inline void cp_async_wait_group() {   // at most N of the most recent groups may still be in flight
  auto it = async_copies().find(&emu::self());
  if (it == async_copies().end()) return;
  AsyncCopies& a = it->second;
  while ((int)a.groups.size() > N) { cp_async_land(a.groups.front()); a.groups.erase(a.groups.begin()); }
}
inline void cp_async_wait_all() {
  auto it = async_copies().find(&emu::self());
  if (it == async_copies().end()) return;
  for (auto& g : it->second.groups) cp_async_land(g);
  cp_async_land(it->second.open);
  async_copies().erase(it);
}

inline void ldsm_common(uint32_t addr, bool trans, uint32_t (&r)[4]) {
  ++op_counters().ldsm;
  if (addr & 15u) emu::rt().error = "ldmatrix: row address not 16-byte aligned";
  const int lane = emu::self().lane, g = lane >> 2, q = lane & 3;
  uint8_t(*slots)[64] = emu::warp_exchange(&addr, 4, "ldmatrix");
  for (int j = 0; j < 4; ++j) {
    if (!trans) {
      uint32_t a;
      memcpy(&a, slots[8 * j + g], 4);
      memcpy(&r[j], smem_ptr(a + 4 * q, 4, "ldmatrix"), 4);
    } else {
      uint32_t a0, a1;
      memcpy(&a0, slots[8 * j + 2 * q], 4);
      memcpy(&a1, slots[8 * j + 2 * q + 1], 4);
      uint16_t lo, hi;
      memcpy(&lo, smem_ptr(a0 + 2 * g, 2, "ldmatrix.trans"), 2);
      memcpy(&hi, smem_ptr(a1 + 2 * g, 2, "ldmatrix.trans"), 2);
      r[j] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
  }
}
inline void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  uint32_t r[4];
  ldsm_common(addr, false, r);
  r0 = r[0]; r1 = r[1]; r2 = r[2]; r3 = r[3];
}
inline void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  uint32_t r[4];
  ldsm_common(addr, true, r);
  r0 = r[0]; r1 = r[1]; r2 = r[2]; r3 = r[3];
}

inline float bf16_half(uint32_t w, int hi) {
  const uint32_t u = hi ? (w & 0xffff0000u) : (w << 16);
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  ++op_counters().mma;
  const int lane = emu::self().lane, g = lane >> 2, q = lane & 3;
  uint32_t mine[6] = {a[0], a[1], a[2], a[3], b0, b1};
  uint8_t(*slots)[64] = emu::warp_exchange(mine, sizeof(mine), "mma.sync");
  auto A = [&](int row, int k) {   // row 0..15, k 0..15
    uint32_t regs[6];
    memcpy(regs, slots[(row & 7) * 4 + ((k & 7) >> 1)], sizeof(regs));
    return bf16_half(regs[(row >> 3) + 2 * (k >> 3)], k & 1);
  };
  auto B = [&](int k, int n) {     // k 0..15, n 0..7
    uint32_t regs[6];
    memcpy(regs, slots[n * 4 + ((k & 7) >> 1)], sizeof(regs));
    return bf16_half(regs[4 + (k >> 3)], k & 1);
  };
  for (int i = 0; i < 4; ++i) {
    const int row = g + 8 * (i >> 1), col = 2 * q + (i & 1);
    float s = c[i];
    for (int k = 0; k < 16; ++k) s = fmaf(A(row, k), B(k, col), s);
    c[i] = s;
  }
}
template <int NTHREADS> inline void named_bar_sync(int id) {
  if (id < 1 || id > 15) { emu::rt().error = "bar.sync: barrier id out of range"; return; }
  emu::rendezvous(emu::self().cta->named[id], NTHREADS, "bar.sync (named)");
}
inline float ex2f(float x) { ++op_counters().ex2; return exp2f(x); }
inline void prefetch_l1(const void*) {}
inline float rcp_approx(float x) { ++op_counters().rcp; return 1.0f / x; }
inline float tanh_approx(float x) { ++op_counters().tanh; return tanhf(x); }
inline float2 ffma2(float2 a, float2 b, float2 c) { ++op_counters().packed_fp32; return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline float2 fadd2(float2 a, float2 b) { ++op_counters().packed_fp32; return make_float2(a.x + b.x, a.y + b.y); }
inline float2 fmul2(float2 a, float2 b) { ++op_counters().packed_fp32; return make_float2(a.x * b.x, a.y * b.y); }

inline uint32_t cluster_rank() { return emu::self().cta->rank; }
inline void cluster_arrive() {
  emu::Thread& t = emu::self();
  emu::Cluster& cl = *t.cta->cluster;
  const int expected = (int)cl.ctas.size() * t.cta->nthreads;
  t.cluster_wait_gen = cl.rv.gen;
  if (++cl.rv.arrived == expected) { cl.rv.arrived = 0; ++cl.rv.gen; ++emu::rt().progress; }
}
inline void cluster_arrive_relaxed() { cluster_arrive(); }
inline void cluster_wait() {
  emu::Thread& t = emu::self();
  emu::Cluster& cl = *t.cta->cluster;
  t.waiting_on = "barrier.cluster.wait";
  while (cl.rv.gen == t.cluster_wait_gen) emu::yield();
  t.waiting_on = "";
}
inline void cluster_barrier() { cluster_arrive(); cluster_wait(); }
inline void st_peer_f32x2(uint32_t local_addr, uint32_t peer, float a, float b) {
  emu::Cluster& cl = *emu::self().cta->cluster;
  if (peer >= cl.ctas.size()) { emu::rt().error = "mapa: peer rank outside the cluster"; return; }
  emu::Cta& c = cl.ctas[peer];
  if ((size_t)local_addr + 8 > c.smem_bytes || (local_addr & 7u)) { emu::rt().error = "st.shared::cluster.v2.f32: bad address"; return; }
  memcpy(c.smem + local_addr, &a, 4);
  memcpy(c.smem + local_addr + 4, &b, 4);
}

}  // namespace prims
}  // namespace dsheg
