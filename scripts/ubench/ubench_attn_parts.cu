// Micro-benchmarks that de-risk the attention redesign (attn_rows): (1) legacy mma.sync m16n8k16 bf16 issue rate per SM
// sub-partition at 1 / 2 / 4 / 6 warps per scheduler, (2) TMEM used as a per-thread parking lot by ordinary (mma.sync) warps:
// tcgen05.st / tcgen05.ld 32x32b from 16 warps of one CTA, round-trip checked, cycles per park / unpark.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_variants/ubench_attn_parts scripts/ubench/ubench_attn_parts.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void hmma_rate(float* out, long long* cyc, int iters) {
  float acc[8][4];
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  uint32_t a[4] = {0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u};
  uint32_t b0 = 0x3F803F80u + threadIdx.x, b1 = 0x3F803F80u;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) mma_bf16(acc[i], a, b0, b1);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 17 warps like the planned kernel: warps 0..15 park two 32-column sets each, warp 16 allocates / frees 256 columns
__global__ void __launch_bounds__(544, 1) tmem_park(int* bad, long long* cyc, int rounds) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  int nbad = 0;
  long long tst = 0, tld = 0;
  if (warp < 16) {
    const uint32_t taddr = base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 64);
    for (int r = 0; r < rounds; ++r) {
      uint32_t a[32], b[32], c[32];
      for (int i = 0; i < 32; ++i) { a[i] = (blockIdx.x << 24) ^ (warp << 16) ^ (lane << 8) ^ i ^ (r << 20); b[i] = ~a[i]; }
      long long t0 = clock64();
      tmem_st32(taddr, a);
      tmem_st32(taddr + 32, b);
      tmem_wait_st();
      long long t1 = clock64();
      tmem_ld32(taddr + 32, c);
      for (int i = 0; i < 32; ++i) nbad += (c[i] != b[i]);
      tmem_ld32(taddr, c);
      long long t2 = clock64();
      for (int i = 0; i < 32; ++i) nbad += (c[i] != a[i]);
      tst += t1 - t0; tld += t2 - t1;
    }
    if (nbad) atomicAdd(bad, nbad);
    if (lane == 0) { cyc[(blockIdx.x * 16 + warp) * 2] = tst / rounds; cyc[(blockIdx.x * 16 + warp) * 2 + 1] = tld / rounds; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base) : "memory");
}

int main() {
  float* out; long long* cyc; int* bad;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 64 * 8); cudaMalloc(&bad, 4);
  cudaMemset(bad, 0, 4);
  const int iters = 2000;
  for (int warps : {4, 8, 16, 24, 32}) {
    hmma_rate<<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double per_smsp = (double)iters * 8 * warps / 4;
    printf("hmma m16n8k16 bf16: %2d warps/CTA: %.0f cycles for %d HMMA per warp -> %.2f cycles per HMMA per SMSP, %.0f FLOP/clk/SM\n",
           warps, c, iters * 8, c / per_smsp, 4096.0 * per_smsp * 4 / c);
  }
  tmem_park<<<148, 544>>>(bad, cyc, 50);
  cudaError_t e = cudaDeviceSynchronize();
  int hb = -1; long long h[148 * 32];
  cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double st = 0, ld = 0; for (int i = 0; i < 148 * 16; ++i) { st += h[2 * i]; ld += h[2 * i + 1]; }
  printf("tmem park: %s, mismatches = %d, 2 x st.x32 + wait = %.0f cycles, 2 x (ld.x32 + wait) = %.0f cycles (16 warps concurrently)\n",
         cudaGetErrorString(e), hb, st / (148 * 16), ld / (148 * 16));
  return (e != cudaSuccess) || hb != 0;
}
