// TEST INFRASTRUCTURE ONLY -- host models of diffsheg_b200/csrc/tc_prims.cuh (same names and signatures) for the thread-level
// emulator (emu_cuda.h).  What is modelled, from the PTX ISA / CUDA documentation the kernel was written against:
//   mbarrier      : {expected arrivals, pending arrivals, pending tx bytes, phase}; a phase completes when pending == 0 and
//                   tx == 0; try_wait.parity(P) succeeds once the phase of parity P has completed (phase bit != P)
//   TMA 2-D tiles : box [rows x cols] of a row-major bf16 tensor, out-of-bounds reads are zero and writes are clipped;
//                   SWIZZLE_128B / 64B = XOR of smem address bits [4,7) with bits [7,10) (64B: [4,6) with [7,9));
//                   a load credits box-bytes to its mbarrier (complete_tx); the cta_group::2 form credits the LEADER's
//   tcgen05.mma   : kind::f16, D[M x N] (+)= A[M x 16] . B[N x 16]^T, fp32 accumulate in TMEM; operands from smem through
//                   K-major SWIZZLE_128B descriptors (start address >> 4, SBO = 1024 B per 8 rows, the same address-bit
//                   XOR); cta_group::2: M = 256 -- rows 0..127 from the leader's smem / TMEM, 128..255 from the peer's;
//                   each CTA supplies HALF of the N rows of B.  Executed synchronously at issue, so tcgen05.commit
//                   arrives immediately (the real ordering guarantees are a superset).
//   kind::tf32    : the same with fp32 operands read as TF32 (19 bits), K = 8 per instruction; TFLOAT32 tensor maps round on load
//   tcgen05.ld    : 32x32b.x32 -- thread `lane` of the warp reads 32 consecutive columns of TMEM lane (base + lane)
//   TMEM          : 128 lanes x 512 fp32 columns per CTA; alloc returns base 0
// Not modelled: async-proxy / generic-proxy ordering (fence.proxy.async), TMEM allocation contention, timing.
#pragma once
#include <unordered_map>

#include "emu_cuda.h"

struct CUtensorMap {   // emulated tensor map: row-major bf16 [rows, cols], leading dimension in bytes
  const void* base = nullptr;
  uint64_t cols = 0, rows = 0, ld_bytes = 0;
  uint32_t box_cols = 0, box_rows = 0;
  int swizzle_bytes = 128;
  uint64_t n2 = 1, ld2_bytes = 0;   // 3-D maps (column, frame, sample): number of samples and their stride; `rows` = frames per sample
  int elem_bytes = 2;               // 2: bf16; 4: fp32 loaded as TF32 (CU_TENSOR_MAP_DATA_TYPE_TFLOAT32: gemm_tf32.cuh)
};

#define DSHEG_TC_DYN_SMEM(name) uint8_t* name = emu::self().cta->smem

namespace emu {
struct MBar { int expected = 0, pending = 0; long long tx = 0; uint32_t phase = 0; bool init = false; };
struct TcState {
  std::unordered_map<uint32_t, MBar> bars;   // keyed by smem offset
  std::vector<float> tmem;                    // [128][512]
};
inline TcState& tc_state(const Cta* c) {
  Cta* cc = const_cast<Cta*>(c);
  if (!cc->ext) {
    auto s = std::make_shared<TcState>();
    s->tmem.assign(128 * 512, 0.0f);
    cc->ext = s;
  }
  return *static_cast<TcState*>(cc->ext.get());
}
// resolve a shared::cta offset or a mapa-encoded shared::cluster address to (CTA, offset)
inline Cta* resolve(uint32_t addr, uint32_t* off) {
  Cta* me = self().cta;
  if (addr & 0x80000000u) {
    const uint32_t rank = (addr >> 24) & 0x7Fu;
    *off = addr & 0x00FFFFFFu;
    if (rank >= me->cluster->ctas.size()) { rt().error = "cluster address: rank outside the cluster"; return me; }
    return &me->cluster->ctas[rank];
  }
  *off = addr;
  return me;
}
inline MBar& bar_at(uint32_t addr, const char* what) {
  uint32_t off;
  Cta* c = resolve(addr, &off);
  if (off & 7u) rt().error = std::string(what) + ": mbarrier address not 8-byte aligned";
  MBar& b = tc_state(c).bars[off];
  if (!b.init && std::string(what) != "mbarrier.init") rt().error = std::string(what) + ": mbarrier used before mbarrier.init";
  return b;
}
inline void bar_check_complete(MBar& b) {
  if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.expected; }
  ++rt().progress;
}
inline void bar_arrive(uint32_t addr, const char* what) {
  MBar& b = bar_at(addr, what);
  if (b.pending <= 0) { rt().error = std::string(what) + ": more arrivals than the mbarrier expects"; return; }
  --b.pending;
  bar_check_complete(b);
}
inline void bar_complete_tx(uint32_t addr, long long bytes) {
  MBar& b = bar_at(addr, "complete_tx");
  b.tx -= bytes;
  bar_check_complete(b);
}
// Adversarial timing: EMU_DELAY_TMEM_LD / EMU_DELAY_TMA / EMU_DELAY_MMA = number of scheduler passes the calling thread
// sits out before the operation.  The round-robin schedule makes every role equally "fast"; slowing one role down exposes
// protocol bugs that need a particular interleaving (a TMEM stage or smem slot handed back before its last read, ...).
inline void delay(const char* env_name) {
  const char* e = getenv(env_name);
  const int n = e ? atoi(e) : 0;
  for (int i = 0; i < n; ++i) { ++rt().progress; yield(); }
}
inline uint32_t swizzle_addr(uint32_t a, int swizzle_bytes) {
  if (swizzle_bytes == 128) return a ^ (((a >> 7) & 7u) << 4);
  if (swizzle_bytes == 64) return a ^ (((a >> 7) & 3u) << 4);
  return a;
}
}  // namespace emu

namespace dsheg {
namespace tc {

inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
inline void mbar_init(uint32_t bar, uint32_t count) {
  emu::MBar& b = emu::bar_at(bar, "mbarrier.init");
  b.expected = b.pending = (int)count; b.tx = 0; b.phase = 0; b.init = true;
}
inline void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  emu::bar_at(bar, "mbarrier.arrive.expect_tx").tx += bytes;
  emu::bar_arrive(bar, "mbarrier.arrive.expect_tx");
}
inline void mbar_arrive(uint32_t bar) { emu::bar_arrive(bar, "mbarrier.arrive"); }
inline uint64_t globaltimer_ns() { return 0; }   // the watchdog never fires: deadlocks are reported by the scheduler
inline uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  emu::MBar& b = emu::bar_at(bar, "mbarrier.try_wait");
  if (((b.phase ^ parity) & 1u) != 0) return 1;
  emu::self().waiting_on = "mbarrier.try_wait";
  emu::yield();
  emu::self().waiting_on = "";
  return 0;
}
inline uint32_t mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t) { return mbar_try_wait(bar, parity); }

// ---- TMA ------------------------------------------------------------------------------------------------------------------
inline void tma_copy(const CUtensorMap* m, emu::Cta* cta, uint32_t smem_off, int c0, int c1, bool load) {
  if (smem_off % (m->swizzle_bytes == 128 ? 1024 : (m->swizzle_bytes == 64 ? 512 : 16)))
    emu::rt().error = "TMA: shared-memory box not aligned to its swizzle atom";
  const uint32_t eb = (uint32_t)m->elem_bytes, row_bytes = m->box_cols * eb;
  if ((size_t)smem_off + (size_t)row_bytes * m->box_rows > cta->smem_bytes) { emu::rt().error = "TMA: box outside shared memory"; return; }
  for (uint32_t r = 0; r < m->box_rows; ++r) {
    for (uint32_t c = 0; c < m->box_cols; ++c) {
      const long long gr = (long long)c1 + r, gc = (long long)c0 + c;
      const bool inb = gr >= 0 && gc >= 0 && (uint64_t)gr < m->rows && (uint64_t)gc < m->cols;
      uint8_t* sp = cta->smem + emu::swizzle_addr(smem_off + r * row_bytes + c * eb, m->swizzle_bytes);
      uint8_t* gp = const_cast<uint8_t*>(static_cast<const uint8_t*>(m->base)) + (size_t)gr * m->ld_bytes + (size_t)gc * eb;
      if (load) {
        if (!inb) { memset(sp, 0, eb); continue; }
        if (eb == 4) {   // TFLOAT32 load: the 13 low mantissa bits are rounded away (nearest even); Inf / NaN pass through
          uint32_t u;
          memcpy(&u, gp, 4);
          if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x0FFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
          memcpy(sp, &u, 4);
        } else {
          memcpy(sp, gp, eb);
        }
      } else if (inb) {
        memcpy(gp, sp, eb);
      }
    }
  }
}
inline long long tma_box_bytes(const CUtensorMap* m) { return (long long)m->box_cols * m->elem_bytes * m->box_rows; }
inline void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  emu::delay("EMU_DELAY_TMA");
  tma_copy(map, emu::self().cta, dst, c0, c1, true);
  emu::bar_complete_tx(bar, tma_box_bytes(map));
}
inline void tma_prefetch_l2_2d(const CUtensorMap*, int, int) {}
// 3-D (column, frame, sample): the box covers one sample; frames beyond the sample's `rows` (and samples beyond n2) arrive as zeros
inline void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  emu::delay("EMU_DELAY_TMA");
  CUtensorMap m2 = *map;
  if (c2 < 0 || (uint64_t)c2 >= map->n2) m2.rows = 0;   // everything out of bounds
  else m2.base = static_cast<const uint8_t*>(map->base) + (size_t)c2 * map->ld2_bytes;
  tma_copy(&m2, emu::self().cta, dst, c0, c1, true);
  emu::bar_complete_tx(bar, tma_box_bytes(map));
}
inline void tma_load_3d_hint(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, uint64_t) { tma_load_3d(map, bar, dst, c0, c1, c2); }
inline void tma_prefetch_l2_3d(const CUtensorMap*, int, int, int) {}
constexpr uint64_t L2_EVICT_NORMAL = 0, L2_EVICT_FIRST = 1, L2_EVICT_LAST = 2;   // cache hints are not modelled
inline void tma_load_2d_pair(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst, int c0, int c1, uint64_t = 0) {
  emu::delay("EMU_DELAY_TMA");
  tma_copy(map, emu::self().cta, dst, c0, c1, true);
  emu::bar_complete_tx(leader_bar, tma_box_bytes(map));
}
// TMA stores are ASYNCHRONOUS: the engine may read the shared-memory box at any time until the issuing thread's
// cp.async.bulk.wait_group(.read) returns.  The model reads it at the LATEST legal moment -- at that wait -- so a kernel that reuses
// a box without waiting stores the overwritten data and fails its parity check; a store that is never waited for before the CTA
// exits is reported (launch_variant checks tma_stores_drained after the grid).
struct PendingStore { CUtensorMap map; emu::Cta* cta; uint32_t src; int c0, c1; };
inline std::unordered_map<emu::Thread*, std::vector<PendingStore>>& pending_stores() {
  static std::unordered_map<emu::Thread*, std::vector<PendingStore>> p;
  return p;
}
inline void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  pending_stores()[&emu::self()].push_back(PendingStore{*map, emu::self().cta, src, c0, c1});
}
inline void bulk_commit() {}
inline void tma_flush_stores() {
  auto it = pending_stores().find(&emu::self());
  if (it == pending_stores().end()) return;
  for (const PendingStore& s : it->second) tma_copy(&s.map, s.cta, s.src, s.c0, s.c1, false);
  pending_stores().erase(it);
}
inline void bulk_wait_read0() { tma_flush_stores(); }
inline void bulk_wait0() { tma_flush_stores(); }
inline bool tma_stores_drained() {
  const bool ok = pending_stores().empty();
  pending_stores().clear();
  return ok;
}
inline void fence_async_smem() {}

// ---- tcgen05 ---------------------------------------------------------------------------------------------------------------
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void tc_commit(uint32_t bar) { emu::bar_arrive(bar, "tcgen05.commit"); }
inline void tc_commit_pair(uint32_t bar) {   // multicast mask 0b11: the same barrier offset in both CTAs of the pair
  uint32_t off;
  emu::resolve(bar, &off);
  for (uint32_t r = 0; r < 2; ++r) emu::bar_arrive(0x80000000u | (r << 24) | off, "tcgen05.commit (pair)");
}
inline float emu_bf16_at(const emu::Cta* c, uint32_t addr) {
  if ((size_t)addr + 2 > c->smem_bytes) { emu::rt().error = "tcgen05.mma: operand read outside shared memory"; return 0.f; }
  uint16_t h;
  memcpy(&h, c->smem + addr, 2);
  const uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline float emu_f32_at(const emu::Cta* c, uint32_t addr) {
  if ((size_t)addr + 4 > c->smem_bytes) { emu::rt().error = "tcgen05.mma: operand read outside shared memory"; return 0.f; }
  uint32_t u;
  memcpy(&u, c->smem + addr, 4);
  u &= 0xFFFFE000u;   // kind::tf32 reads 19 bits (the TMA load already rounded them)
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// element (row, k) of a K-major SWIZZLE_128B operand described by `desc`; eb = 2: bf16, k < 16; eb = 4: tf32, k < 8 (one UMMA_K slice = 32 bytes)
inline float emu_operand(const emu::Cta* c, uint64_t desc, int row, int k, int eb = 2) {
  const uint32_t start = (uint32_t)(desc & 0x3FFFu) << 4;
  const uint32_t sbo = (uint32_t)((desc >> 32) & 0x3FFFu) << 4;
  const uint32_t layout = (uint32_t)(desc >> 61) & 7u;
  if (layout != 2 || sbo != 1024) emu::rt().error = "tcgen05.mma: the emulator models K-major SWIZZLE_128B descriptors with SBO = 1024 only";
  const uint32_t lin = start + (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * 128u + (uint32_t)(k * eb);
  return eb == 4 ? emu_f32_at(c, emu::swizzle_addr(lin, 128)) : emu_bf16_at(c, emu::swizzle_addr(lin, 128));
}
// kind: 0 = kind::f16 with bf16 operands (K = 16 per instruction), 1 = kind::tf32 (K = 8)
inline void emu_mma(int cg, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate, int kind = 0) {
  const int N = (int)((idesc >> 17) & 0x3Fu) << 3, M = (int)((idesc >> 24) & 0x1Fu) << 4;
  const int K = kind == 1 ? 8 : 16, eb = kind == 1 ? 4 : 2;
  const uint32_t fmt_a = (idesc >> 7) & 7u, fmt_b = (idesc >> 10) & 7u;
  if (kind == 1 && (fmt_a != 2 || fmt_b != 2)) emu::rt().error = "tcgen05.mma.kind::tf32: the instruction descriptor does not say TF32 operands";
  emu::delay("EMU_DELAY_MMA");
  emu::Cta* me = emu::self().cta;
  if (cg == 2 && me->rank != 0) { emu::rt().error = "tcgen05.mma.cta_group::2 issued by the non-leader CTA"; return; }
  if (M != 128 * cg || N % 16 || N < 16 || N > 256 || (tmem_d >> 16) != 0 || (tmem_d & 0xFFFFu) + (uint32_t)N > 512u) {
    emu::rt().error = "tcgen05.mma: unsupported instruction descriptor / TMEM address in the emulator";
    return;
  }
  const int col0 = (int)(tmem_d & 0xFFFFu);
  std::vector<float> a((size_t)M * K), b((size_t)N * K);
  for (int m = 0; m < M; ++m) {
    const emu::Cta* c = cg == 2 ? &me->cluster->ctas[m / 128] : me;
    for (int k = 0; k < K; ++k) a[(size_t)m * K + k] = emu_operand(c, desc_a, m % 128, k, eb);
  }
  const int n_per_cta = N / cg;
  for (int n = 0; n < N; ++n) {
    const emu::Cta* c = cg == 2 ? &me->cluster->ctas[n / n_per_cta] : me;
    for (int k = 0; k < K; ++k) b[(size_t)n * K + k] = emu_operand(c, desc_b, n % n_per_cta, k, eb);
  }
  for (int m = 0; m < M; ++m) {
    emu::Cta* c = cg == 2 ? &me->cluster->ctas[m / 128] : me;
    float* d = emu::tc_state(c).tmem.data() + (size_t)(m % 128) * 512 + col0;
    const float* am = &a[(size_t)m * K];
    for (int n = 0; n < N; ++n) {
      const float* bn = &b[(size_t)n * K];
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += am[k] * bn[k];
      d[n] = accumulate ? d[n] + s : s;
    }
  }
  ++emu::rt().progress;
}
inline void tc_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { emu_mma(1, tmem_d, da, db, idesc, acc); }
inline void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { emu_mma(2, tmem_d, da, db, idesc, acc); }
inline void tc_mma_tf32_emu(int cg, uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { emu_mma(cg, tmem_d, da, db, idesc, acc, 1); }
inline void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  emu::delay("EMU_DELAY_TMEM_LD");
  emu::Thread& t = emu::self();
  const uint32_t lane_base = taddr >> 16, col = taddr & 0xFFFFu;
  if (lane_base != (uint32_t)(t.warp & 3) * 32u) emu::rt().error = "tcgen05.ld: a warp may only touch the TMEM lane quadrant (warp id % 4)";
  if (lane_base + 32 > 128 || col + 32 > 512) { emu::rt().error = "tcgen05.ld: TMEM address out of range"; return; }
  const float* src = emu::tc_state(t.cta).tmem.data() + (size_t)(lane_base + t.lane) * 512 + col;
  memcpy(r, src, 32 * 4);
  __syncwarp();   // .sync.aligned: the warp executes it together
}
inline void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  emu::Thread& t = emu::self();
  const uint32_t lane_base = taddr >> 16, col = taddr & 0xFFFFu;
  if (lane_base != (uint32_t)(t.warp & 3) * 32u) emu::rt().error = "tcgen05.st: a warp may only touch the TMEM lane quadrant (warp id % 4)";
  if (lane_base + 32 > 128 || col + 32 > 512) { emu::rt().error = "tcgen05.st: TMEM address out of range"; return; }
  memcpy(emu::tc_state(t.cta).tmem.data() + (size_t)(lane_base + t.lane) * 512 + col, r, 32 * 4);
  __syncwarp();   // .sync.aligned
}
template <int N> inline float* emu_tmem_row(uint32_t taddr, const char* what) {
  emu::Thread& t = emu::self();
  const uint32_t lane_base = taddr >> 16, col = taddr & 0xFFFFu;
  if (lane_base != (uint32_t)(t.warp & 3) * 32u) emu::rt().error = std::string(what) + ": a warp may only touch the TMEM lane quadrant (warp id % 4)";
  if (lane_base + 32 > 128 || col + N > 512) { emu::rt().error = std::string(what) + ": TMEM address out of range"; return nullptr; }
  return emu::tc_state(t.cta).tmem.data() + (size_t)(lane_base + t.lane) * 512 + col;
}
inline void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  if (float* p = emu_tmem_row<16>(taddr, "tcgen05.st")) memcpy(p, r, 16 * 4);
  __syncwarp();
}
inline void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  emu::delay("EMU_DELAY_TMEM_LD");
  if (const float* p = emu_tmem_row<16>(taddr, "tcgen05.ld")) memcpy(r, p, 16 * 4);
  __syncwarp();
}
inline void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  if (float* p = emu_tmem_row<8>(taddr, "tcgen05.st")) memcpy(p, r, 8 * 4);
  __syncwarp();
}
inline void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  emu::delay("EMU_DELAY_TMEM_LD");
  if (const float* p = emu_tmem_row<8>(taddr, "tcgen05.ld")) memcpy(r, p, 8 * 4);
  __syncwarp();
}
// split load: the destination registers hold poison until the thread's tcgen05.wait::ld (a use before the wait fails parity)
struct PendingTmemLd { uint32_t* dst; const float* src; int n; };
inline std::unordered_map<emu::Thread*, std::vector<PendingTmemLd>>& pending_tmem_lds() {
  static std::unordered_map<emu::Thread*, std::vector<PendingTmemLd>> p;
  return p;
}
inline void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  emu::delay("EMU_DELAY_TMEM_LD");
  for (int i = 0; i < 16; ++i) r[i] = 0x7FC0DEADu;   // NaN poison
  if (const float* p = emu_tmem_row<16>(taddr, "tcgen05.ld")) pending_tmem_lds()[&emu::self()].push_back(PendingTmemLd{r, p, 16});
  __syncwarp();
}
inline void tmem_wait_ld() {
  auto it = pending_tmem_lds().find(&emu::self());
  if (it != pending_tmem_lds().end()) {
    for (const PendingTmemLd& l : it->second) memcpy(l.dst, l.src, (size_t)l.n * 4);
    pending_tmem_lds().erase(it);
  }
  __syncwarp();
}
template <int CG, int COLS> inline void tmem_alloc(uint32_t slot) {
  static_assert(COLS == 32 || COLS == 64 || COLS == 128 || COLS == 256 || COLS == 512, "TMEM allocations are powers of two >= 32 columns");
  emu::Cta* c = emu::self().cta;
  if (emu::self().lane == 0) { const uint32_t base = 0; memcpy(c->smem + slot, &base, 4); }
  __syncwarp();
}
template <int CG, int COLS> inline void tmem_dealloc(uint32_t) { __syncwarp(); }

// ---- cluster ----------------------------------------------------------------------------------------------------------------
inline uint32_t cluster_ctarank() { return emu::self().cta->rank; }
inline void cluster_sync_all() {
  emu::Thread& t = emu::self();
  emu::Cluster& cl = *t.cta->cluster;
  emu::rendezvous(cl.rv, (int)cl.ctas.size() * t.cta->nthreads, "barrier.cluster");
}
inline uint32_t mapa_rank(uint32_t saddr, uint32_t rank) { return 0x80000000u | (rank << 24) | (saddr & 0x00FFFFFFu); }
inline void mbar_arrive_cluster(uint32_t cluster_addr) { emu::bar_arrive(cluster_addr, "mbarrier.arrive (cluster)"); }
template <int NTHREADS> inline void epi_bar_sync() { emu::rendezvous(emu::self().cta->named[1], NTHREADS, "bar.sync 1 (epilogue)"); }
inline void fence_mbarrier_init() {}
inline void prefetch_tensormap(const CUtensorMap*) {}
inline void trap() {
  emu::rt().error = "__trap() executed";
  emu::self().done = true;
  ++emu::rt().progress;
  emu::yield();
}
inline float ex2_fast(float x) { return exp2f(x); }
inline float tanh_fast(float x) { return tanhf(x); }
inline float2 ffma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline float2 fmul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 fadd2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }

}  // namespace tc
}  // namespace dsheg
