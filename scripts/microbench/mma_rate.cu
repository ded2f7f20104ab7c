// Microbenchmark: issue rate of tcgen05.mma kind::f16 (bf16, M128 x N x K16, both operands in shared
// memory, SWIZZLE_NONE K-major) as a function of N.  Answers: is a small-N MMA (N = 48 / 96) paced
// by the tensor pipe (N/2 cycles) or by the shared-memory read of the 128-row A operand?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) mma_kernel(int N, int n_mma, int n_acc, int a_tiles, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* data = smem + 1024;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(data)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(N);
    const uint32_t a0 = smem_u32(data), b0 = a0 + 128 * 1024;
    long long t0 = clock64();
    // 8 precomputed descriptor pairs, straight-line issue (the loop itself must not be the limit)
    uint64_t ad[8], bd[8];
    uint32_t dd[8];
    for (int i = 0; i < 8; ++i) {
      ad[i] = make_desc(a0 + (uint32_t)(i % a_tiles) * 4096, 2048, 128);
      bd[i] = make_desc(b0 + (uint32_t)(i % 4) * 32, (uint32_t)N * 16, 128);
      dd[i] = tmem + (uint32_t)((i % n_acc) * N);
    }
    for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dd[u]),
            "l"(ad[u]), "l"(bd[u]), "r"(idesc)
            : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
  }
}

int main() {
  long long* d_cyc;
  cudaMalloc(&d_cyc, 148 * 8);
  const int smem = 1024 + 160 * 1024;
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_mma = 4096;
  const int Ns[] = {16, 32, 48, 64, 96, 128, 144, 192, 256};
  printf("tcgen05.mma kind::f16 M128 x N x K16, SS, no swizzle; %d MMAs per CTA; ideal = N/2 cycles\n", n_mma);
  for (int grid : {1, 148}) {
    for (int N : Ns) {
      for (int n_acc : {1, 2}) {
        if (n_acc * N > 512) continue;
        for (int a_tiles : {1, 8}) {
          mma_kernel<<<grid, 128, smem>>>(N, n_mma, n_acc, a_tiles, d_cyc);
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          mma_kernel<<<grid, 128, smem>>>(N, n_mma, n_acc, a_tiles, d_cyc);
          cudaEventRecord(e1);
          cudaError_t err = cudaDeviceSynchronize();
          float ms = 0;
          cudaEventElapsedTime(&ms, e0, e1);
          long long h[148];
          cudaMemcpy(h, d_cyc, grid * 8, cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          const double cyc = (double)mx / n_mma;
          const double tflops = 2.0 * 128 * N * 16 * n_mma * grid / (ms * 1e-3) / 1e12;
          printf("grid %3d  N %3d  acc %d  a_tiles %d : %7.2f cyc/MMA (ideal %5.1f, %5.1f%%)  kernel %.3f ms  %.0f TFLOP/s  %s\n",
                 grid, N, n_acc, a_tiles, cyc, N / 2.0, 100.0 * (N / 2.0) / cyc, ms, tflops, cudaGetErrorString(err));
        }
      }
    }
  }
  return 0;
}
