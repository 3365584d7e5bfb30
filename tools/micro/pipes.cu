// Micro-benchmark: per-SM throughput of the integer instructions the F1 kernel leans on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the B200.
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256
template <int OP>
__global__ void k(int* out, int a0, int b0, int iters) {
  int a = a0 + threadIdx.x, b = b0, c = threadIdx.x * 3, d = threadIdx.x ^ 5;
  int e = a + 1, f = b + 2, g = c + 3, h = d + 4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < REP; ++r) {
      if (OP == 0) { a = a * 19595 + b; e = e * 38469 + f; c = c * 7471 + d; g = g * 13 + h; }            // IMAD
      if (OP == 1) { a = __mulhi(a, 0x6a0a0000) + b; e = __mulhi(e, 0x32ec0000) + f; c = __mulhi(c, 0x5a820000) + d; g = __mulhi(g, 0x12340000) + h; }  // IMAD.HI
      if (OP == 2) { a = __dp4a(a, 0x01020304, b); e = __dp4a(e, 0x05060708, f); c = __dp4a(c, 0x090a0b0c, d); g = __dp4a(g, 0x0d0e0f01, h); }  // IDP4A
      if (OP == 3) { a = __byte_perm(a, b, 0x5410) + 1; e = __byte_perm(e, f, 0x7531) + 1; c = __byte_perm(c, d, 0x6420) + 1; g = __byte_perm(g, h, 0x3210) + 1; }  // PRMT + IADD
      if (OP == 4) { a = (a >> 3) ^ b; e = (e >> 5) ^ f; c = (c >> 7) ^ d; g = (g >> 9) ^ h; }       // SHF + LOP3
      if (OP == 5) { a = (a * 19595 + b) >> 16; e = (e * 38469 + f) >> 16; c = (c * 7471 + d) >> 16; g = (g * 13 + h) >> 16; }  // IMAD + SHF mix
      if (OP == 6) { a = abs(a) + b; e = abs(e) + f; c = abs(c) + d; g = abs(g) + h; }                 // IABS + IADD
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + e + c + g + b + f + d + h;
}

template <int OP>
void run(const char* name, int* out, int ops_per_rep) {
  const int iters = 200, blocks = 148 * 4, threads = 256;
  k<OP><<<blocks, threads>>>(out, 1, 2, 2);
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0); cudaEventCreate(&t1);
  cudaEventRecord(t0);
  k<OP><<<blocks, threads>>>(out, 1, 2, iters);
  cudaEventRecord(t1);
  cudaEventSynchronize(t1);
  float ms; cudaEventElapsedTime(&ms, t0, t1);
  double warp_inst = (double)blocks * (threads / 32) * iters * REP * 4 * ops_per_rep;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-16s %.3f ms  %.2f warp-inst/clk/SM (at %d MHz nominal)\n", name, ms, warp_inst / (ms * 1e-3) / (clk * 1e3) / 148, clk / 1000);
}

int main() {
  int* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
  run<0>("IMAD", out, 1);
  run<1>("IMAD.HI+add", out, 1);
  run<2>("IDP4A", out, 1);
  run<3>("PRMT+IADD", out, 2);
  run<4>("SHF+LOP3", out, 2);
  run<5>("IMAD+SHF", out, 2);
  run<6>("IABS+IADD", out, 2);
  return 0;
}
