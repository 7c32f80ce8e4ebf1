// Micro-benchmark 3 (round 2): (a) tcgen05.ld throughput per SM with 4 / 8 / 12 / 16 reader warps (32x32b .x8/.x16/.x32),
// (b) tcgen05.mma M=128 K=16 cost for the "linear strip" A view (SBO = 128 B: 128 consecutive 16-byte pixels) at
// N = 16..192, i.e. the kx-stacked weight tile N = 3 * Cout of DESIGN.md section 8.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/tmem_bench.cu -o tools/tmem_bench
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

template <int X>
__device__ __forceinline__ uint32_t tld(uint32_t addr) {
  uint32_t acc = 0;
  if constexpr (X == 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i];
  } else if constexpr (X == 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
  } else {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i];
  }
  return acc;
}

// two loads in flight before one wait (what an unrolled epilogue can do)
__device__ __forceinline__ uint32_t tld16x2(uint32_t addr) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(addr + 16));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc ^= r[i];
  return acc;
}

template <int X, int MODE>
__global__ void __launch_bounds__(512) tmem_read(long long* out, uint32_t* sink, int iters) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)((i * 32 + (warp >> 2) * 64) & 255);
    if constexpr (MODE == 1) acc ^= tld16x2(row + col);
    else acc ^= tld<X>(row + col);
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
  sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int X, int MODE>
void run_read(int warps, long long* d, uint32_t* sink) {
  const int iters = 2000;
  tmem_read<X, MODE><<<1, warps * 32>>>(d, sink, iters);
  long long h[16];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  long long mx = 0;
  for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
  const int cols = MODE == 1 ? 32 : X;
  const double bytes = (double)warps * iters * 32 * cols * 4;
  printf("tcgen05.ld 32x32b.x%-2d%s %2d warps: %6.1f cyc per load per warp, %7.1f B/clk/SM\n", MODE == 1 ? 16 : X,
         MODE == 1 ? " (2 in flight)" : "              ", warps, (double)mx / iters, bytes / (double)mx);
}

template <int NE>
__global__ void __launch_bounds__(128) mma_strip(long long* out, int steps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
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
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NE >> 3) << 17) | (8u << 24);
    // A: planes of [20 rows][32 px][8 ch] = 10240 B each (LBO), 128 consecutive pixels per M-tile (SBO = 128 B)
    const uint32_t w_smem = base, a_smem = base + 64 * 1024;
    const uint64_t a0 = mk_desc(a_smem, 10240, 128), b0 = mk_desc(w_smem, NE * 16, 128);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int s = 0; s < steps; ++s) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem + (uint32_t)((s & 1) * 256)),
                         "l"(a0 + (uint64_t)(((s & 3) * 128 + ky * 32) + kk * 1280)), "l"(b0 + (uint64_t)((ky * 2 + kk) * 2 * NE)),
                         "r"(idesc) : "memory");
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
void run_mma(long long* d) {
  cudaFuncSetAttribute(mma_strip<NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  const int steps = 256;
  mma_strip<NE><<<1, 128, 165 * 1024>>>(d, steps);
  long long h[3];
  cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  printf("tcgen05.mma M=128 K=16 linear-strip A, N=%3d : %6.1f cyc/MMA (%.0f %% of the tensor peak)\n", NE,
         (double)h[2] / (steps * 6), 100.0 * (NE / 2.0) / ((double)h[2] / (steps * 6)));
}

int main() {
  long long* d;
  uint32_t* sink;
  cudaMalloc(&d, 256);
  cudaMalloc(&sink, 4096);
  for (int w : {1, 4, 8, 12, 16}) run_read<8, 0>(w, d, sink);
  for (int w : {1, 4, 8, 12, 16}) run_read<16, 0>(w, d, sink);
  for (int w : {1, 4, 8, 16}) run_read<32, 0>(w, d, sink);
  for (int w : {4, 8, 16}) run_read<16, 1>(w, d, sink);
  run_mma<16>(d); run_mma<32>(d); run_mma<48>(d); run_mma<64>(d); run_mma<96>(d); run_mma<128>(d); run_mma<192>(d);
  run_mma<256>(d);
  return 0;
}
