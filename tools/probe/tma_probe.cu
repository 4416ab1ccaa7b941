// Stand-alone probe of the TMA tile load used by k_fast_tiles: tma_probe <box_w> <box_h> <pitch> <x0> <y0>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int bw, int bh, int x0, int y0, uint8_t* out) {
  extern __shared__ __align__(128) uint8_t s[];
  uint64_t* barp = reinterpret_cast<uint64_t*>(s + ((bw * bh + 127) & ~127));
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(barp), dst = (uint32_t)__cvta_generic_to_shared(s);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bw * bh) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(x0), "r"(y0), "r"(bar) : "memory");
  }
  __syncthreads();
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar) : "memory");
  } while (!done);
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv) {
  const int bw = atoi(argv[1]), bh = atoi(argv[2]), pitch = atoi(argv[3]), x0 = atoi(argv[4]), y0 = atoi(argv[5]);
  const int rows = 1000;
  std::vector<uint8_t> img((size_t)pitch * rows);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 24);
  uint8_t *d, *o;
  cudaMalloc(&d, img.size()); cudaMalloc(&o, bw * bh);
  cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto enc = (PFN_cuTensorMapEncodeTiled)p;
  CUtensorMap m;
  cuuint64_t gd[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}; cuuint64_t gs[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %dx%d pitch %d at (%d,%d): encode=%d ", bw, bh, pitch, x0, y0, (int)r);
  const int smem = ((bw * bh + 127) & ~127) + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<1, 128, smem>>>(m, bw, bh, x0, y0, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run=%s ", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<uint8_t> out(bw * bh);
    cudaMemcpy(out.data(), o, out.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
      const int gx = x0 + x, gy = y0 + y;
      const uint8_t ref = (gx < pitch && gy < rows && gx >= 0 && gy >= 0) ? img[(size_t)gy * pitch + gx] : 0;
      bad += out[y * bw + x] != ref;
    }
    printf("mismatches=%d", bad);
  }
  printf("\n");
  return 0;
}
