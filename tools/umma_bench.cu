// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, fp16 -> fp32) as a function of N and of the smem operand
// layout (128B/64B/32B swizzle vs the no-swizzle "core matrix" layout used by conv_halo.cu).  Build & run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench tools/umma_bench.cu && ./umma_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

struct Cfg { int n, layout, a_lbo, a_sbo, b_lbo, b_sbo, iters, a_stride, d_rot, m; };

__global__ void __launch_bounds__(128) bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && wid < c.d_rot) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(c.n >> 3) << 17) | ((uint32_t)(c.m >> 4) << 24);
    const uint64_t da0 = mk_desc(base, c.a_lbo, c.a_sbo, c.layout);
    const uint64_t db = mk_desc(base + 96 * 1024, c.b_lbo, c.b_sbo, c.layout);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      const uint32_t t1m = tmem + (uint32_t)(wid * c.n);
      for (int i = 0; i < c.iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint64_t da = da0 + (uint64_t)(u * 2);
          asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;"
                       ::"r"(t1m), "l"(da), "l"(db), "r"(idesc) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[wid])) : "memory");
      mbar_wait(smem_u32(&bar[wid]), phase);
      phase ^= 1;
      const long long t1 = clock64();
      if (wid == 0) out[rep] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2048;
  struct L { const char* name; int layout, a_lbo, a_sbo, b_sbo_mode; } layouts[] = {
      {"SW128 (K-major, 128B rows)", 2, 16, 1024, 0},
      {"NOSWZ halo 3x3 (LBO=2880,SBO=160)", 0, 2880, 160, 1},
  };
  printf("%-40s %6s %10s %10s\n", "layout", "N", "cyc/MMA", "floor N/2");
  for (auto& l : layouts) {
    for (int n : {16, 32, 64, 128, 256}) {
      Cfg c{};
      c.n = n; c.layout = l.layout; c.a_lbo = l.a_lbo; c.a_sbo = l.a_sbo; c.iters = iters;
      if (l.b_sbo_mode == 0) { c.b_lbo = 16; c.b_sbo = l.a_sbo; } else { c.b_lbo = n * 16; c.b_sbo = 128; }
      c.a_stride = 2;  // +32 bytes between consecutive MMAs (like k-steps / taps)
      for (int m : {128}) for (int rot : {1, 2, 4}) {
        if (512 / n < rot) continue;
        c.d_rot = rot; c.m = m;
        bench<<<1, 128, 170 * 1024>>>(c, d);
        long long h[3];
        cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        printf("%-36s M=%3d N=%3d issuers=%d %8.1f cyc per MMA-per-issuer (aggregate %.1f)\n", l.name, m, n, rot, (double)h[2] / iters, (double)h[2] / iters / rot);
      }
    }
  }
  return 0;
}
