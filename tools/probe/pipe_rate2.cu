// Microbenchmark (round 2): issue rate of the instructions the FAST kernel is made of, as warp instructions per clock per SM
// (8 independent chains per thread, 1024 threads per SM), alone and mixed, to see which share a pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rate2 pipe_rate2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm volatile("lop3.b32 %0, %1, %2, %3, 0xb2;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t vabs4(uint32_t a, uint32_t b) {
  uint32_t d;
  asm volatile("vabsdiff4.u32.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0));
  return d;
}
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t iadd(uint32_t a, uint32_t b) {
  uint32_t d;
  asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t shf(uint32_t a, uint32_t b) {
  uint32_t d;
  asm volatile("shf.r.wrap.b32 %0, %1, %2, 8;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }

template <int OP>
__global__ void k(unsigned* out, unsigned a, unsigned b, int iters) {
  unsigned r[8];
  __shared__ unsigned sm[1024];
  sm[threadIdx.x] = threadIdx.x;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) r[i] = lop3(r[i], a, b);
      if (OP == 1) r[i] = vabs4(r[i], a);
      if (OP == 2) r[i] = imad(r[i], a, b);
      if (OP == 3) r[i] = iadd(r[i], a);
      if (OP == 4) r[i] = shf(r[i], a);
      if (OP == 5) r[i] = vmax3(r[i], a, b);
      if (OP == 6) r[i] = (i & 1) ? lop3(r[i], a, b) : imad(r[i], a, b);          // ALU + FMA pipes
      if (OP == 7) r[i] = (i & 1) ? lop3(r[i], a, b) : vabs4(r[i], a);            // same pipe?
      if (OP == 8) r[i] = (i & 1) ? imad(r[i], a, b) : vabs4(r[i], a);
      if (OP == 9) r[i] = (i & 3) == 3 ? sm[(r[i] + i) & 1023] : lop3(r[i], a, b);  // 1 LDS per 3 LOP3
      if (OP == 10) r[i] = (i & 1) ? lop3(r[i], a, b) : __popc(r[i]) + a;
      if (OP == 11) r[i] = (i & 1) ? lop3(r[i], a, b) : vmax3(r[i], a, b);
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, unsigned* d) {
  int sms = 148, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<sms * 4, 256>>>(d, 0x00010001u, 3u, 16);
  cudaEventRecord(e0);
  k<OP><<<sms * 4, 256>>>(d, 0x00010001u, 3u, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double winst = (double)sms * 4 * 8 * iters * 8;   // warps * chains * iters
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-22s %.3f ms  %.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, winst / (ms * 1e-3) / (clk * 1e3) / sms, clk / 1000);
}
int main() {
  unsigned* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  run<0>("lop3", d); run<1>("vabsdiff4", d); run<2>("imad", d); run<3>("iadd", d); run<4>("shf", d); run<5>("vimnmx3.u16x2", d);
  run<6>("lop3+imad", d); run<7>("lop3+vabsdiff4", d); run<8>("imad+vabsdiff4", d); run<9>("3 lop3 + 1 lds", d);
  run<10>("lop3+popc(+iadd)", d); run<11>("lop3+vimnmx3", d);
  return 0;
}
