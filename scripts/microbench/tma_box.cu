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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
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
                                                    int depth, int step1, int step2, int n1, int n2, int n3, long long* cycles) {
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
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++my;
    int tile = blockIdx.x;
    while (waited < my) {
      while (issued < my && issued - waited < depth) {
        const int s = issued % depth;
        mbar_expect_tx(&bar[s], (uint32_t)(nbox * box_bytes));
        // tile -> coordinates: c1 = (tile % n1) * step1 ; c3 = (tile / n1) % n3 ; c4 = tile / (n1*n3)
        const int a = tile % n1, q = tile / n1;
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  // activation tensor: B=4, T=256, F=3072, C=48 bf16 = 302 MB
  const long long Bn = 4, T = 256, F = 3072, C = 48;
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
  for (const Case& cs : cases) {
    CUtensorMap map;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, cs.dims, cs.strides, cs.box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-40s encode failed %d\n", cs.name, (int)r); continue; }
    int box_bytes = 2;
    for (int i = 0; i < 5; ++i) box_bytes *= cs.box[i];
    const int n_tiles = cs.n1 * cs.n3 * (int)Bn;
    for (int depth : {2, 4, 8}) {
      const int slot = (cs.nbox * ((box_bytes + 127) / 128 * 128) + 1023) / 1024 * 1024;
      const int smem = 1024 + depth * slot;
      if (smem > 227 * 1024) continue;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      tma_kernel<<<148, 128, smem>>>(map, n_tiles, cs.nbox, box_bytes, depth, cs.step1, cs.step2, cs.n1, 1, cs.n3, d_cyc);
      cudaEventRecord(e0);
      tma_kernel<<<148, 128, smem>>>(map, n_tiles, cs.nbox, box_bytes, depth, cs.step1, cs.step2, cs.n1, 1, cs.n3, d_cyc);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      long long cyc[148];
      cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
      const double total = (double)n_tiles * cs.nbox * box_bytes;
      printf("%-40s depth %d: %8.3f ms  %7.1f GB/s  %6.1f B/clk/SM  (%s)\n", cs.name, depth, ms, total / ms / 1e6,
             total / 148.0 / (double)mx, cudaGetErrorString(err));
    }
  }
  return 0;
}
