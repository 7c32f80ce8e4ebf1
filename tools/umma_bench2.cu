// Micro-benchmark 2: the exact operand access pattern of conv_halo.cu's ResConv (c=64, MT=2): 72 MMAs per step with
// distinct A (shifted halo views) and B (per-tap weight tiles) descriptors.  Prints cycles per MMA.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
template <int KSTEPS, int NE, int MT, int MODE, int E = 0, int KK = 0>
__device__ __forceinline__ void issue(uint32_t acc, uint64_t a_base, uint64_t b_base, uint32_t idesc) {
  if constexpr (E < 9) {
    constexpr int pitch = 10, plane16 = (16 * MT + 2) * pitch;
    // MODE 0: real pattern; 1: same B for every MMA; 2: same A for every MMA; 3: both fixed
    constexpr int a_off = (MODE == 2 || MODE == 3) ? 0 : 2 * KK * plane16 + (E / 3) * pitch + (E % 3);
    constexpr int b_off = (MODE == 1 || MODE == 3) ? 0 : E * (2 * KSTEPS) * NE + 2 * KK * NE;
    asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(acc), "l"(a_base + (uint64_t)a_off),
                 "l"(b_base + (uint64_t)b_off), "r"(idesc) : "memory");
    if constexpr (KK + 1 < KSTEPS) issue<KSTEPS, NE, MT, MODE, E, KK + 1>(acc, a_base, b_base, idesc);
    else issue<KSTEPS, NE, MT, MODE, E + 1, 0>(acc, a_base, b_base, idesc);
  }
}
template <int KSTEPS, int NE, int MODE>
__global__ void __launch_bounds__(128) bench(long long* out, int steps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    constexpr int MT = 2;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NE >> 3) << 17) | (8u << 24);
    const uint32_t w_smem = base, a_smem = base + 76 * 1024;
    const uint64_t a0 = mk_desc(a_smem, (16 * MT + 2) * 10 * 16, 160), b0 = mk_desc(w_smem, NE * 16, 128);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int s = 0; s < steps; ++s)
        for (int m = 0; m < MT; ++m)
          issue<KSTEPS, NE, MT, MODE>(tmem + ((s & 1) * MT + m) * NE, a0 + (uint64_t)(m * 160) + (uint64_t)((s % 2) * 2720), b0, idesc);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      out[rep] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int KSTEPS, int NE, int MODE>
void run(const char* name, long long* d) {
  cudaFuncSetAttribute(bench<KSTEPS, NE, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int steps = 64;
  bench<KSTEPS, NE, MODE><<<1, 128, 205 * 1024>>>(d, steps);
  long long h[3];
  cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  printf("%-44s KSTEPS=%d NE=%3d : %7.1f cyc/MMA\n", name, KSTEPS, NE, (double)h[2] / (steps * 2 * 9 * KSTEPS));
}
int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<4, 64, 0>("real pattern (distinct A views, distinct B)", d);
  run<4, 64, 1>("same B tile every MMA", d);
  run<4, 64, 2>("same A view every MMA", d);
  run<4, 64, 3>("same A and B", d);
  run<2, 32, 0>("real pattern c=32", d);
  run<2, 32, 3>("same A and B c=32", d);
  run<1, 16, 0>("real pattern c=16", d);
  run<1, 16, 3>("same A and B c=16", d);
  return 0;
}
