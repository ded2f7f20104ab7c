// Microbenchmark: how fast does TMA fill shared memory for different box shapes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_box tma_box.cu && ./tma_box
// One elected thread per CTA issues `depth` boxes in flight into a smem ring; 148 persistent CTAs.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ int g_arrive_mode = 0;  // 0: arrive.expect_tx (release), 1: arrive.expect_tx.relaxed.cta, 2: copy first, then arrive.expect_tx
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_relaxed(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ int g_wait_mode = 0;  // 0: try_wait (may suspend), 1: test_wait spin
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  if (g_wait_mode == 0) {
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
  } else {
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
  }
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// Each "tile" = nbox boxes of box_bytes each landing in one ring slot; tiles walk dim `walk` of the map.
__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap map, int n_tiles, int nbox, int box_bytes,
                                                    int depth, int step1, int step2, int n1, int n2, int n3, long long* cycles, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint8_t* ring = smem + 1024;
  const int box_pitch = (box_bytes + 127) / 128 * 128;
  const int slot_bytes = (nbox * box_pitch + 1023) / 1024 * 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    int issued = 0, waited = 0;
    uint32_t ph = 0;
    int my = 0;
    for (int tile = blockIdx.x; tile < n_tiles * reps; tile += gridDim.x) ++my;
    int tile = blockIdx.x;
    while (waited < my) {
      while (issued < my && issued - waited < depth) {
        const int s = issued % depth;
        mbar_expect_tx(&bar[s], (uint32_t)(nbox * box_bytes));
        // tile -> coordinates: c1 = (tile % n1) * step1 ; c3 = (tile / n1) % n3 ; c4 = tile / (n1*n3)
        const int tl = tile % n_tiles;
        const int a = tl % n1, q = tl / n1;
        for (int b = 0; b < nbox; ++b)
          tma_load_5d(ring + (size_t)s * slot_bytes + (size_t)b * box_pitch, &map, &bar[s], 0, a * step1, b * step2, q % n3, q / n3);
        tile += gridDim.x;
        ++issued;
      }
      const int s = waited % depth;
      mbar_wait(&bar[s], ph);
      ++waited;
      if (waited % depth == 0) ph ^= 1;
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

// 1-D bulk copies: each tile = nchunk contiguous chunks of chunk_bytes (plain cp.async.bulk, no tensor map)
__global__ void __launch_bounds__(128, 1) bulk_kernel(const uint8_t* src, size_t span_bytes, int n_tiles, int nchunk, int chunk_bytes,
                                                     size_t chunk_pitch, int depth, long long* cycles, int reps, int issuers) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint8_t* ring = smem + 1024;
  const int slot_bytes = (nchunk * chunk_bytes + 1023) / 1024 * 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    int issued = 0, waited = 0;
    uint32_t ph = 0;
    int my = 0;
    for (int tile = blockIdx.x; tile < n_tiles * reps; tile += gridDim.x) ++my;
    int tile = blockIdx.x;
    while (waited < my) {
      while (issued < my && issued - waited < depth) {
        const int s = issued % depth;
        mbar_expect_tx(&bar[s], (uint32_t)(nchunk * chunk_bytes));
        const size_t base = ((size_t)(tile % n_tiles) * 2080) % (span_bytes - (size_t)nchunk * chunk_pitch);
        for (int b = 0; b < nchunk; ++b)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(ring + (size_t)s * slot_bytes + (size_t)b * chunk_bytes)),
                       "l"(src + (base / 16 * 16) + (size_t)b * chunk_pitch), "r"(chunk_bytes), "r"(smem_u32(&bar[s]))
                       : "memory");
        tile += gridDim.x;
        ++issued;
      }
      const int s = waited % depth;
      mbar_wait(&bar[s], ph);
      ++waited;
      if (waited % depth == 0) ph ^= 1;
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

// reference: plain 16-byte LDG streaming over the same (L2-resident) buffer
__global__ void __launch_bounds__(1024, 1) ldg_kernel(const uint4* src, size_t n16, int reps, uint32_t* sink) {
  uint32_t acc = 0;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
      uint4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i));
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678u) *sink = acc;
}
// several independent issuing threads (one per warp), each with its own ring of `depth` 1-D bulk copies
__global__ void __launch_bounds__(256, 1) bulk_multi_kernel(const uint8_t* src, size_t span_bytes, int chunk_bytes, int depth,
                                                           int per_issuer, int issuers, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint8_t* ring = smem + 1024;
  const int w = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth * issuers; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && w < issuers) {
    long long t0 = clock64();
    int issued = 0, waited = 0;
    uint32_t ph = 0;
    uint64_t* mybar = bar + w * depth;
    uint8_t* myring = ring + (size_t)w * depth * chunk_bytes;
    size_t off = ((size_t)(blockIdx.x * issuers + w) * 104729u * 16) % (span_bytes - chunk_bytes);
    while (waited < per_issuer) {
      while (issued < per_issuer && issued - waited < depth) {
        const int s = issued % depth;
        mbar_expect_tx(&mybar[s], (uint32_t)chunk_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(myring + (size_t)s * chunk_bytes)),
                     "l"(src + off / 16 * 16), "r"(chunk_bytes), "r"(smem_u32(&mybar[s]))
                     : "memory");
        off = (off + chunk_bytes * 7) % (span_bytes - chunk_bytes);
        ++issued;
      }
      mbar_wait(&mybar[waited % depth], ph);
      ++waited;
      if (waited % depth == 0) ph ^= 1;
    }
    if (w == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}

// (a) pure issue + transfer: `n` bulk copies of chunk_bytes into a small smem ring, ONE barrier phase per `group` copies
__global__ void __launch_bounds__(128, 1) bulk_group_kernel(const uint8_t* src, size_t span_bytes, int chunk_bytes, int group, int n_groups,
                                                           int depth, long long* cycles, int amode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint8_t* ring = smem + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    int issued = 0, waited = 0;
    uint32_t ph = 0;
    uint32_t off = (uint32_t)blockIdx.x * 65536u;
    const uint32_t mask = 8u * 1024 * 1024 - 1;  // stay inside the first 8 MB
    while (waited < n_groups) {
      while (issued < n_groups && issued - waited < depth) {
        const int s = issued % depth;
        if (amode == 0) mbar_expect_tx(&bar[s], (uint32_t)(group * chunk_bytes));
        else if (amode == 1) mbar_expect_tx_relaxed(&bar[s], (uint32_t)(group * chunk_bytes));
        for (int g = 0; g < group; ++g) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(ring + ((size_t)s * group + g) * chunk_bytes)),
                       "l"(src + (off & mask)), "r"(chunk_bytes), "r"(smem_u32(&bar[s]))
                       : "memory");
          off += (uint32_t)chunk_bytes * 3;
        }
        if (amode == 2) mbar_expect_tx(&bar[s], (uint32_t)(group * chunk_bytes));
        ++issued;
      }
      mbar_wait(&bar[waited % depth], ph);
      ++waited;
      if (waited % depth == 0) ph ^= 1;
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  // usage: tma_box [T rows (default 256)] [B (default 4)] [reps (default 1)]: small T*B keeps the tensor L2 resident
  const long long Targ = argc > 1 ? atoll(argv[1]) : 256, Barg = argc > 2 ? atoll(argv[2]) : 4;
  const int reps = argc > 3 ? atoi(argv[3]) : 1;
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  // activation tensor: B=4, T=256, F=3072, C=48 bf16 = 302 MB
  const long long Bn = Barg, T = Targ, F = 3072, C = 48;
  printf("tensor %.1f MB, %d passes\n", (double)Bn * T * F * C * 2 / 1e6, reps);
  const size_t bytes = (size_t)Bn * T * F * C * 2;
  void* d;
  cudaMalloc(&d, bytes);
  cudaMemset(d, 1, bytes);
  long long* d_cyc;
  cudaMalloc(&d_cyc, 148 * 8);
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);

  struct Case { const char* name; cuuint64_t dims[5]; cuuint64_t strides[4]; cuuint32_t box[5]; int nbox, step1, step2, n1, n3; };
  // every case moves ~the same bytes: per tile 6 channel groups x ~130 positions x 16 B
  Case cases[] = {
      // (a) today: channels-last [B][T][F][C]; box = 8 ch x 130 pos (16-B rows, 96-B pitch); 6 boxes (one per channel group)
      {"NHWC  box{8,130} pitch 96B, 6 boxes", {(cuuint64_t)8, (cuuint64_t)F, (cuuint64_t)(C / 8), (cuuint64_t)T, (cuuint64_t)Bn},
       {(cuuint64_t)C * 2, 16, (cuuint64_t)F * C * 2, (cuuint64_t)T * F * C * 2}, {8, 130, 1, 1, 1}, 6, 128, 1, (int)(F / 128), (int)T},
      // (a2) same memory, one 3-D box {8,130,6}
      {"NHWC  box{8,130,6} pitch 96B, 1 box", {(cuuint64_t)8, (cuuint64_t)F, (cuuint64_t)(C / 8), (cuuint64_t)T, (cuuint64_t)Bn},
       {(cuuint64_t)C * 2, 16, (cuuint64_t)F * C * 2, (cuuint64_t)T * F * C * 2}, {8, 130, 6, 1, 1}, 1, 128, 0, (int)(F / 128), (int)T},
      // (b) planar [B][T][C/8][F][8]; box = 8 ch x 130 pos contiguous (16-B rows, 16-B pitch), 6 boxes
      {"CG8   box{8,130} pitch 16B, 6 boxes", {(cuuint64_t)8, (cuuint64_t)F, (cuuint64_t)(C / 8), (cuuint64_t)T, (cuuint64_t)Bn},
       {16, (cuuint64_t)F * 16, (cuuint64_t)F * C * 2, (cuuint64_t)T * F * C * 2}, {8, 130, 1, 1, 1}, 6, 128, 1, (int)(F / 128), (int)T},
      // (b2) planar, one box {8,130,6}
      {"CG8   box{8,130,6} pitch 16B, 1 box", {(cuuint64_t)8, (cuuint64_t)F, (cuuint64_t)(C / 8), (cuuint64_t)T, (cuuint64_t)Bn},
       {16, (cuuint64_t)F * 16, (cuuint64_t)F * C * 2, (cuuint64_t)T * F * C * 2}, {8, 130, 6, 1, 1}, 1, 128, 0, (int)(F / 128), (int)T},
      // (c) planar viewed with 128-B inner rows: dims {64, F/8, C/8, T, B}; box {64,18,6}
      {"CG8   box{64,18,6} 128-B rows, 1 box", {(cuuint64_t)64, (cuuint64_t)(F / 8), (cuuint64_t)(C / 8), (cuuint64_t)T, (cuuint64_t)Bn},
       {128, (cuuint64_t)F * 16, (cuuint64_t)F * C * 2, (cuuint64_t)T * F * C * 2}, {64, 18, 6, 1, 1}, 1, 16, 0, (int)(F / 128), (int)T},
      // (d) planar, 64-B inner rows: dims {32, F/4, ...}; box {32,34,6}
      {"CG8   box{32,34,6} 64-B rows, 1 box", {(cuuint64_t)32, (cuuint64_t)(F / 4), (cuuint64_t)(C / 8), (cuuint64_t)T, (cuuint64_t)Bn},
       {64, (cuuint64_t)F * 16, (cuuint64_t)F * C * 2, (cuuint64_t)T * F * C * 2}, {32, 34, 6, 1, 1}, 1, 32, 0, (int)(F / 128), (int)T},
      // (e) planar, 256-element inner (512 B): dims {256, F/32, ...}; box {256,5,6}  (160 positions)
      {"CG8   box{256,5,6} 512-B rows, 1 box", {(cuuint64_t)256, (cuuint64_t)(F / 32), (cuuint64_t)(C / 8), (cuuint64_t)T, (cuuint64_t)Bn},
       {512, (cuuint64_t)F * 16, (cuuint64_t)F * C * 2, (cuuint64_t)T * F * C * 2}, {256, 5, 6, 1, 1}, 1, 4, 0, (int)(F / 128), (int)T},
  };
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  {
    uint32_t* sink;
    cudaMalloc(&sink, 4);
    for (int grid : {148, 296}) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      ldg_kernel<<<grid, 1024>>>((const uint4*)d, bytes / 16, reps, sink);
      cudaEventRecord(e0);
      ldg_kernel<<<grid, 1024>>>((const uint4*)d, bytes / 16, reps, sink);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("LDG.128 streaming, grid %d x 1024 thr: %8.3f ms  %7.1f GB/s\n", grid, ms, (double)bytes * reps / ms / 1e6);
    }
    cudaFuncSetAttribute(bulk_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(bulk_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int amode : {0, 1, 2})
    for (int chunk : {8192})
      for (int group : {1, 4})
        for (int depth : {1, 3, 6}) {
          const int n_groups = 400;
          const int smem = 1024 + depth * group * chunk;
          if (smem > 227 * 1024) continue;
          bulk_group_kernel<<<148, 128, smem>>>((const uint8_t*)d, bytes, chunk, group, n_groups, depth, d_cyc, amode);
          cudaError_t err = cudaDeviceSynchronize();
          long long cyc[148];
          cudaMemcpy(cyc, d_cyc, 148 * 8, cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
          printf("bulk group (arrive mode %d): %5d B x %d per phase, depth %d: %7.1f cycles/phase  %6.1f B/clk/SM (%s)\n", amode, chunk, group, depth,
                 (double)mx / n_groups, (double)chunk * group * n_groups / (double)mx, cudaGetErrorString(err));
        }
    for (int mode : {0}) {
    cudaMemcpyToSymbol(g_wait_mode, &mode, sizeof(int));
    printf("wait mode %d (0 try_wait, 1 test_wait)\n", mode);
    for (int issuers : {1, 4})
      for (int chunk : {4096, 16384})
      for (int depth : {1, 2, 3, 6, 12}) {
        const int per = 400;
        const int smem = 1024 + issuers * depth * chunk;
        if (smem > 227 * 1024) continue;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        bulk_multi_kernel<<<148, 256, smem>>>((const uint8_t*)d, bytes, chunk, depth, per, issuers, d_cyc);
        cudaEventRecord(e0);
        bulk_multi_kernel<<<148, 256, smem>>>((const uint8_t*)d, bytes, chunk, depth, per, issuers, d_cyc);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        long long cyc[148];
        cudaMemcpy(cyc, d_cyc, 148 * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
        const double total = 148.0 * issuers * per * chunk;
        printf("bulk 1-D %5d B x depth %2d, %d issuing warps: %8.3f ms  %7.1f GB/s  %6.1f B/clk/SM (%s)\n", chunk, depth, issuers, ms,
               total / ms / 1e6, total / 148.0 / (double)mx, cudaGetErrorString(err));
      }
    }
    { int mode = 0; cudaMemcpyToSymbol(g_wait_mode, &mode, sizeof(int)); }
  }
  struct BCase { const char* name; int nchunk, chunk_bytes; size_t pitch; };
  BCase bcases[] = {{"bulk 1-D 6 x 2080 B (plane pitch)", 6, 2080, (size_t)F * 16}, {"bulk 1-D 1 x 12480 B", 1, 12480, 0},
                    {"bulk 1-D 3 x 4160 B", 3, 4160, (size_t)F * 16}, {"bulk 1-D 24 x 512 B", 24, 512, (size_t)F * 16}};
  for (const BCase& bc : bcases) {
    const int n_tiles = 24 * (int)T * (int)Bn;
    for (int grid : {37, 74, 148}) {
      for (int depth : {4, 8}) {
        const int slot = (bc.nchunk * bc.chunk_bytes + 1023) / 1024 * 1024;
        const int smem = 1024 + depth * slot;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        bulk_kernel<<<grid, 128, smem>>>((const uint8_t*)d, bytes, n_tiles, bc.nchunk, bc.chunk_bytes, bc.pitch, depth, d_cyc, reps, 1);
        cudaEventRecord(e0);
        bulk_kernel<<<grid, 128, smem>>>((const uint8_t*)d, bytes, n_tiles, bc.nchunk, bc.chunk_bytes, bc.pitch, depth, d_cyc, reps, 1);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        long long cyc[148];
        cudaMemcpy(cyc, d_cyc, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
        const double total = (double)n_tiles * reps * bc.nchunk * bc.chunk_bytes;
        printf("%-40s grid %3d depth %d: %8.3f ms  %7.1f GB/s  %6.1f B/clk/SM  (%s)\n", bc.name, grid, depth, ms, total / ms / 1e6,
               total / grid / (double)mx, cudaGetErrorString(err));
      }
    }
  }
  {
    // GEMM-style operand tiles: row-major matrix [rows][4096] bf16, box {64 cols (128 B), 128 rows} = 16 KB
    const long long cols = 4096, rows = (long long)(bytes / (cols * 2));
    for (int sw = 0; sw < 2; ++sw) {
      CUtensorMap map;
      const cuuint64_t dims[5] = {(cuuint64_t)cols, (cuuint64_t)rows, 1, 1, 1};
      const cuuint64_t strides[4] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * rows * 2, (cuuint64_t)cols * rows * 2, (cuuint64_t)cols * rows * 2};
      const cuuint32_t box[5] = {64, 128, 1, 1, 1};
      const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("gemm-like encode failed %d\n", (int)r); continue; }
      const int n1 = (int)(rows / 128), n_tiles = n1 * 8;  // walk 128-row blocks (c1) for 8 column blocks... (c0 fixed at 0)
      for (int grid : {37, 148})
        for (int depth : {4, 8}) {
          const int box_bytes = 64 * 128 * 2;
          const int smem = 1024 + depth * 16384;
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          tma_kernel<<<grid, 128, smem>>>(map, n1, 1, box_bytes, depth, 128, 0, n1, 1, 1, d_cyc, reps * 8);
          cudaEventRecord(e0);
          tma_kernel<<<grid, 128, smem>>>(map, n1, 1, box_bytes, depth, 128, 0, n1, 1, 1, d_cyc, reps * 8);
          cudaEventRecord(e1);
          cudaError_t err = cudaDeviceSynchronize();
          float ms = 0;
          cudaEventElapsedTime(&ms, e0, e1);
          long long cyc[148];
          cudaMemcpy(cyc, d_cyc, grid * 8, cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < grid; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
          const double total = (double)n1 * reps * 8 * box_bytes;
          printf("GEMM-like box{64,128} %-18s grid %3d depth %d: %8.3f ms  %7.1f GB/s  %6.1f B/clk/SM  (%s)\n",
                 sw ? "SWIZZLE_128B" : "no swizzle", grid, depth, ms, total / ms / 1e6, total / grid / (double)mx, cudaGetErrorString(err));
          (void)n_tiles;
        }
    }
  }
  for (const Case& cs : cases) {
    CUtensorMap map;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, cs.dims, cs.strides, cs.box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-40s encode failed %d\n", cs.name, (int)r); continue; }
    int box_bytes = 2;
    for (int i = 0; i < 5; ++i) box_bytes *= cs.box[i];
    const int n_tiles = cs.n1 * cs.n3 * (int)Bn;
    for (int grid : {37, 148})
    for (int depth : {4, 8}) {
      const int slot = (cs.nbox * ((box_bytes + 127) / 128 * 128) + 1023) / 1024 * 1024;
      const int smem = 1024 + depth * slot;
      if (smem > 227 * 1024) continue;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      tma_kernel<<<grid, 128, smem>>>(map, n_tiles, cs.nbox, box_bytes, depth, cs.step1, cs.step2, cs.n1, 1, cs.n3, d_cyc, reps);
      cudaEventRecord(e0);
      tma_kernel<<<grid, 128, smem>>>(map, n_tiles, cs.nbox, box_bytes, depth, cs.step1, cs.step2, cs.n1, 1, cs.n3, d_cyc, reps);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      long long cyc[148];
      cudaMemcpy(cyc, d_cyc, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
      const double total = (double)n_tiles * reps * cs.nbox * box_bytes;
      printf("%-40s grid %3d depth %d: %8.3f ms  %7.1f GB/s  %6.1f B/clk/SM  (%s)\n", cs.name, grid, depth, ms, total / ms / 1e6,
             total / grid / (double)mx, cudaGetErrorString(err));
    }
  }
  return 0;
}
