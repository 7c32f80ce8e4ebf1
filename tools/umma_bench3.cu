// Micro-benchmark 4 (round 2): cost of an M = 128, N = 64, K = 16 tcgen05.mma whose A operand is the row-stacked view of
// conv_head.cu -- 8-row groups (core matrices) SBO bytes apart, the two K chunks LBO bytes apart -- for several SBO / LBO.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/umma_bench3.cu -o tools/umma_bench3
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
template <int NE>
__global__ void __launch_bounds__(128) mma_k(long long* out, int steps, uint32_t a_lbo, uint32_t a_sbo, uint32_t pitch16,
                                             uint32_t b_lbo, int commit_every) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)) : "memory");
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
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NE >> 3) << 17) | (8u << 24);
    const uint32_t w_smem = base, a_smem = base + 32 * 1024;
    const uint64_t a0 = mk_desc(a_smem, a_lbo, a_sbo), b0 = mk_desc(w_smem, b_lbo, 128);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int s = 0; s < steps; ++s) {
#pragma unroll
        for (int xi = 0; xi < 3; ++xi)
#pragma unroll
          for (int ji = 0; ji < 6; ++ji)
            asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem + (uint32_t)((s & 3) * 64)),
                         "l"(a0 + (uint64_t)((s & 3) * 8 + ji * pitch16 + xi)), "l"(b0 + (uint64_t)(xi * 288 + (5 - ji) * 16)),
                         "r"(idesc) : "memory");
        if (commit_every && (s + 1) % commit_every == 0)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
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
template <int NE>
void run(long long* d, const char* what, uint32_t a_lbo, uint32_t a_sbo, uint32_t pitch16, uint32_t b_lbo, int commit_every = 0) {
  cudaFuncSetAttribute(mma_k<NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int steps = 128;
  mma_k<NE><<<1, 128, 205 * 1024>>>(d, steps, a_lbo, a_sbo, pitch16, b_lbo, commit_every);
  long long h[3];
  cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  printf("N=%3d %-44s A LBO %6u SBO %5u  B LBO %5u : %6.1f cyc/MMA\n", NE, what, a_lbo, a_sbo, b_lbo, (double)h[2] / (steps * 18));
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
// the same loop executed by a whole converged warp, the MMA issued by the elected lane: the descriptors stay warp-uniform
template <int NE>
__global__ void __launch_bounds__(128) mma_k_elect(long long* out, int steps, uint32_t a_lbo, uint32_t a_sbo, uint32_t pitch16,
                                                   uint32_t b_lbo) {
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
  if (threadIdx.x < 32) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NE >> 3) << 17) | (8u << 24);
    const uint32_t w_smem = base, a_smem = base + 32 * 1024;
    const uint64_t a0 = mk_desc(a_smem, a_lbo, a_sbo), b0 = mk_desc(w_smem, b_lbo, 128);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int s = 0; s < steps; ++s) {
#pragma unroll
        for (int xi = 0; xi < 3; ++xi)
#pragma unroll
          for (int ji = 0; ji < 6; ++ji)
            if (elect_one())
              asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem + (uint32_t)((s & 3) * 64)),
                           "l"(a0 + (uint64_t)((s & 3) * 8 + ji * pitch16 + xi)), "l"(b0 + (uint64_t)(xi * 288 + (5 - ji) * 16)),
                           "r"(idesc) : "memory");
      }
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      if (threadIdx.x == 0) out[rep] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int NE>
void run_elect(long long* d, const char* what, uint32_t a_lbo, uint32_t a_sbo, uint32_t pitch16, uint32_t b_lbo) {
  cudaFuncSetAttribute(mma_k_elect<NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int steps = 128;
  mma_k_elect<NE><<<1, 128, 205 * 1024>>>(d, steps, a_lbo, a_sbo, pitch16, b_lbo);
  long long h[3];
  cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  printf("N=%3d %-44s A LBO %6u SBO %5u  B LBO %5u : %6.1f cyc/MMA\n", NE, what, a_lbo, a_sbo, b_lbo, (double)h[2] / (steps * 18));
}
int main() {
  long long* d;
  cudaMalloc(&d, 256);
  run<64>(d, "conv_head.cu (R = 4, pitch 36)", 40320, 2304, 36, 2304);
  run<64>(d, "R = 4, pitch 35", 39200, 2240, 35, 2304);
  run<64>(d, "R = 4, pitch 37", 41440, 2368, 37, 2304);
  run<64>(d, "R = 4, pitch 36, plane + 64 B", 40384, 2304, 36, 2304);
  run<64>(d, "R = 4, pitch 36, B LBO 2048", 40320, 2304, 36, 2048);
  run<64>(d, "halo view (pitch 10)", 40320, 160, 10, 2304);
  run<64>(d, "linear strip (SBO 128)", 40320, 128, 32, 2304);
  run<16>(d, "N = 16, R = 4 view", 40320, 2304, 36, 2304);
  run<16>(d, "N = 16, halo view", 40320, 160, 10, 2304);
  run<32>(d, "N = 32, R = 2 view (pitch 36)", 40320, 1152, 36, 2304);
  run<32>(d, "N = 32, halo view", 40320, 160, 10, 2304);
  run<128>(d, "N = 128, R = 8 view (pitch 36)", 40320, 4608, 36, 2304);
  run<64>(d, "conv_head.cu view, commit every 18 MMAs", 40320, 2304, 36, 2304, 1);
  run<64>(d, "conv_head.cu view, commit every 36 MMAs", 40320, 2304, 36, 2304, 2);
  run<64>(d, "conv_head.cu view, commit every 72 MMAs", 40320, 2304, 36, 2304, 4);
  run_elect<16>(d, "whole warp + elect.sync issuer, halo view", 40320, 160, 10, 2304);
  run_elect<32>(d, "whole warp + elect.sync issuer, halo view", 40320, 160, 10, 2304);
  run_elect<64>(d, "whole warp + elect.sync issuer, halo view", 40320, 160, 10, 2304);
  run_elect<64>(d, "whole warp + elect.sync issuer, R = 4 view", 40320, 2304, 36, 2304);
  run_elect<96>(d, "whole warp + elect.sync issuer, halo view", 40320, 160, 10, 2304);
  run_elect<128>(d, "whole warp + elect.sync issuer, halo view", 40320, 160, 10, 2304);
  return 0;
}
