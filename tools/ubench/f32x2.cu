// Micro-benchmark: issue throughput of packed FP32 (FMUL2 / FFMA2) against scalar FMUL / FADD on sm_100a, and a check that
// mul.rn.f32x2 followed by fma.rn.f32x2(p, ONE, q) with a run-time ONE is NOT contracted (bit-identical to scalar mul + add).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 up(u64 r) { float2 c; asm("mov.b64 {%0,%1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(r)); return c; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 c; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 d) { u64 c; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(c) : "l"(a), "l"(b), "l"(d)); return c; }

constexpr int ILP = 8, ITERS = 4096;
__global__ void k_scalar(float* out, float m, float a) {
  float x[2 * ILP];
  for (int i = 0; i < 2 * ILP; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < 2 * ILP; i++) x[i] = __fadd_rn(__fmul_rn(x[i], m), a);
  float s = 0; for (int i = 0; i < 2 * ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float m, float a, float one) {
  u64 x[ILP]; const u64 M = pk(m, m), A = pk(a, a), ONE = pk(one, one);
  for (int i = 0; i < ILP; i++) x[i] = pk(threadIdx.x + 2 * i, threadIdx.x + 2 * i + 1);
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma2(mul2(x[i], M), ONE, A);
  float s = 0; for (int i = 0; i < ILP; i++) { float2 v = up(x[i]); s += v.x; s += v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: packed FP32 plus an independent integer (ALU pipe) stream, to see whether freed issue slots are usable
__global__ void k_packed_mix(float* out, float m, float a, float one, int seed) {
  u64 x[ILP]; const u64 M = pk(m, m), A = pk(a, a), ONE = pk(one, one);
  int z[ILP];
  for (int i = 0; i < ILP; i++) { x[i] = pk(threadIdx.x + 2 * i, threadIdx.x + 2 * i + 1); z[i] = seed + i; }
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = fma2(mul2(x[i], M), ONE, A); z[i] = (z[i] ^ (z[i] << 1)) + it; z[i] = (z[i] ^ (z[i] >> 3)) + seed; }
  float s = 0; for (int i = 0; i < ILP; i++) { float2 v = up(x[i]); s += v.x; s += v.y; s += z[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_scalar_mix(float* out, float m, float a, int seed) {
  float x[2 * ILP]; int z[ILP];
  for (int i = 0; i < 2 * ILP; i++) x[i] = threadIdx.x + i;
  for (int i = 0; i < ILP; i++) z[i] = seed + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 2 * ILP; i++) x[i] = __fadd_rn(__fmul_rn(x[i], m), a);
#pragma unroll
    for (int i = 0; i < ILP; i++) { z[i] = (z[i] ^ (z[i] << 1)) + it; z[i] = (z[i] ^ (z[i] >> 3)) + seed; }
  }
  float s = 0; for (int i = 0; i < 2 * ILP; i++) s += x[i];
  for (int i = 0; i < ILP; i++) s += z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_check(const float* a, const float* b, const float* c, float* o_s, float* o_p, float one, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * i + 1 < n) {
    o_s[2 * i] = __fadd_rn(__fmul_rn(a[2 * i], b[2 * i]), c[2 * i]);
    o_s[2 * i + 1] = __fadd_rn(__fmul_rn(a[2 * i + 1], b[2 * i + 1]), c[2 * i + 1]);
    float2 r = up(fma2(mul2(pk(a[2 * i], a[2 * i + 1]), pk(b[2 * i], b[2 * i + 1])), pk(one, one), pk(c[2 * i], c[2 * i + 1])));
    o_p[2 * i] = r.x; o_p[2 * i + 1] = r.y;
  }
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int i = 0; i < 5; i++) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
  const int blocks = 148 * 8, threads = 256;
  float* out; cudaMalloc(&out, blocks * threads * 4);
  const double flop_pairs = (double)blocks * threads * ITERS * 2 * ILP;   // (mul, add) pairs
  float t1 = timeit([&] { k_scalar<<<blocks, threads>>>(out, 0.999f, 0.5f); });
  float t2 = timeit([&] { k_packed<<<blocks, threads>>>(out, 0.999f, 0.5f, 1.0f); });
  float t3 = timeit([&] { k_scalar_mix<<<blocks, threads>>>(out, 0.999f, 0.5f, 3); });
  float t4 = timeit([&] { k_packed_mix<<<blocks, threads>>>(out, 0.999f, 0.5f, 1.0f, 3); });
  printf("scalar FMUL+FADD   %.3f ms  %.1f G(mul,add)/s\n", t1, flop_pairs / t1 * 1e-6);
  printf("packed FMUL2+FFMA2 %.3f ms  %.1f G(mul,add)/s\n", t2, flop_pairs / t2 * 1e-6);
  printf("scalar + int mix   %.3f ms\npacked + int mix   %.3f ms\n", t3, t4);
  // contraction check on awkward values
  const int n = 1 << 20; float *a, *b, *c, *os, *op;
  cudaMallocManaged(&a, n * 4); cudaMallocManaged(&b, n * 4); cudaMallocManaged(&c, n * 4); cudaMallocManaged(&os, n * 4); cudaMallocManaged(&op, n * 4);
  srand(1);
  for (int i = 0; i < n; i++) {
    a[i] = (float)rand() / RAND_MAX * 3.f - 1.5f; b[i] = (float)rand() / RAND_MAX * 3.f - 1.5f; c[i] = -a[i] * b[i] * (1.f + ((i & 7) - 3) * 1e-7f);
    if (i % 1000 == 0) { a[i] = 1e-20f; b[i] = 3e-20f; c[i] = 1e-41f; }   // subnormal product
  }
  k_check<<<n / 2 / 256, 256>>>(a, b, c, os, op, 1.0f, n); cudaDeviceSynchronize();
  int diff = 0; for (int i = 0; i < n; i++) diff += (os[i] != op[i]) || (*(unsigned*)&os[i] != *(unsigned*)&op[i]);
  printf("packed vs scalar bit differences: %d of %d  (%s)\n", diff, n, diff ? "CONTRACTED / MISMATCH" : "bit-identical");
  return 0;
}
