// FP64 tensor-core throughput on this GPU: mma.sync m8n8k4 (DMMA.8x8x4) and m16n8k16, N independent accumulators per warp, no memory
// traffic.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k884(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k16816(double* out, int iters, double a0, double b0) {
  double c[NACC][4];
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = a0 + threadIdx.x + i;
  for (int i = 0; i < 4; i++) b[i] = b0 - threadIdx.x - i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7, %8, %9, %10, %11}, {%12, %13, %14, %15}, {%0, %1, %2, %3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void kdfma(double* out, int iters, double a0, double b0) {
  double c[NACC];
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  const int iters = 20000;
  for (int warps : {4, 8, 16}) {
    float ms = timeit([&] { k884<18><<<148, warps * 32>>>(out, iters, 1.0, 2.0); });
    printf("DMMA.884   %2d warps/SM, 18 acc: %.1f TFLOP/s\n", warps, 148.0 * warps * iters * 18 * 512.0 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k16816<9><<<148, warps * 32>>>(out, iters, 1.0, 2.0); });
    printf("DMMA.16816 %2d warps/SM,  9 acc: %.1f TFLOP/s\n", warps, 148.0 * warps * iters * 9 * 4096.0 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { kdfma<16><<<148, warps * 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DFMA       %2d warps/SM, 16 acc: %.1f TFLOP/s\n", warps, 148.0 * warps * 32 * iters * 16 * 2.0 / (ms * 1e-3) / 1e12);
  }
  return 0;
}
