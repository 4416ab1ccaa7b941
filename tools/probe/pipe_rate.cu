// Microbenchmark: issue rate of IDP.2A (dp2a), IDP.4A (dp4a), IMAD, LOP3, VIMNMX3 and PRMT on sm_100a, as warp instructions per
// clock per SM (8 independent chains per thread, 1024 threads per SM). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(unsigned* out, unsigned a, unsigned b, int iters) {
  unsigned r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) r[i] = __dp2a_lo(a, r[i], r[i]);
      if (OP == 1) r[i] = __dp4a(a, r[i], r[i]);
      if (OP == 2) r[i] = r[i] * a + b;
      if (OP == 3) r[i] = (r[i] & a) ^ b;
      if (OP == 4) r[i] = max(max(r[i], a), b) ^ 1u;
      if (OP == 5) r[i] = __byte_perm(r[i], a, b + i);
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
  printf("%-8s %.3f ms  %.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, winst / (ms * 1e-3) / (clk * 1e3) / sms, clk / 1000);
}
int main() {
  unsigned* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  run<0>("dp2a", d); run<1>("dp4a", d); run<2>("imad", d); run<3>("lop3", d); run<4>("vimnmx", d); run<5>("prmt", d);
  return 0;
}
