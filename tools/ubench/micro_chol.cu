// The serial core of the direct solver's panel factorisation, timed alone on one SM: the 16 x 16 micro block (factor + inverse, one warp)
// (one warp factors, a second one inverts), and the whole 96 x 96 diagonal block (factor_core) with its phases.  Checks L L^T = A and W L = I on the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I lv_slam_b200/csrc -I include -o tools/ubench/micro_chol tools/ubench/micro_chol.cu
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../../lv_slam_b200/csrc/pgo_chol.cu"
namespace lvs {
int fail(int status, const char* fmt, ...) { va_list a; va_start(a, fmt); vfprintf(stderr, fmt, a); va_end(a); fputc('\n', stderr); return status; }
int cuda_fail(cudaError_t e, const char* what, const char*, int) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return -1; }
}
using namespace lvs;

__global__ void k_micro(const double* A0, double* Lout, double* Wout, long long* cyc, int reps, int* flag) {
  __shared__ __align__(16) double S[16 * 20], W[16 * 20], il[16], P[16 * 20];
  __shared__ unsigned long long bars[16];
  if (threadIdx.x < 16) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars + threadIdx.x)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (int t = threadIdx.x; t < 256; t += 64) P[(t / 16) * 20 + t % 16] = A0[t];
  __syncthreads();
  long long tot = 0;
  for (int r = 0; r < reps; r++) {
    for (int t = threadIdx.x; t < 320; t += 64) S[t] = P[t];
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x < 32) micro_factor(S, 20, il, flag, bars); else micro_invert(S, 20, il, W, 20, bars, r & 1);
    __syncthreads();
    tot += clock64() - t0;
  }
  if (threadIdx.x == 0) *cyc = tot / reps;
  for (int t = threadIdx.x; t < 256; t += 64) { Lout[t] = (t % 16 <= t / 16) ? S[(t / 16) * 20 + t % 16] : 0.0; Wout[t] = W[(t / 16) * 20 + t % 16]; }
}

__global__ void __launch_bounds__(256, 1) k_core(const double* A0, double* Lout, long long* tm, int reps, int* flag) {
  using Cfg = FrontCfg<true>;
  extern __shared__ __align__(16) double sm[];
  double *s_D = sm + Cfg::oD, *s_W = sm + Cfg::oW, *s_il = sm + Cfg::oIl;
  __shared__ long long s_tm[4];
  __shared__ unsigned long long bars[16];
  unsigned uses = 0;
  if (threadIdx.x < 16) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars + threadIdx.x)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (threadIdx.x < 4) s_tm[threadIdx.x] = 0;
  long long tot = 0;
  for (int r = 0; r < reps; r++) {
    for (int t = threadIdx.x; t < 96 * 96; t += 256) { const int i = t / 96, j = t % 96; s_D[i * Cfg::LDD + j] = (j <= i) ? A0[i * 96 + j] : 0.0; }
    __syncthreads();
    const long long t0 = clock64();
    factor_core<true>(96, s_D, s_W, s_il, flag, bars, uses, s_tm);
    tot += clock64() - t0;
  }
  if (threadIdx.x == 0) { for (int k = 0; k < 4; k++) tm[k] = s_tm[k] / reps; tm[4] = tot / reps; }
  for (int t = threadIdx.x; t < 96 * 96; t += 256) { const int i = t / 96, j = t % 96; Lout[t] = (j <= i) ? s_D[i * Cfg::LDD + j] : 0.0; }
}

static double check_llt(const std::vector<double>& A, const std::vector<double>& L, int n) {
  double worst = 0, scale = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double s = 0;
      for (int k = 0; k <= j; k++) s += L[i * n + k] * L[j * n + k];
      worst = std::fmax(worst, std::fabs(s - A[i * n + j])); scale = std::fmax(scale, std::fabs(A[i * n + j]));
    }
  return worst / scale;
}

int main() {
  const int n = 96;
  std::vector<double> A(n * n);
  unsigned long long x = 88172645463325252ull;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (double)(x >> 11) / 9007199254740992.0 - 0.5; };
  std::vector<double> B(n * n);
  for (auto& v : B) v = rnd();
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) { double s = (i == j) ? 4.0 : 0.0; for (int k = 0; k < n; k++) s += B[i * n + k] * B[j * n + k]; A[i * n + j] = s; }
  std::vector<double> A16(256);
  for (int i = 0; i < 16; i++) for (int j = 0; j < 16; j++) A16[i * 16 + j] = A[i * n + j];
  double *dA, *dA16, *dL, *dW; long long* dT; int* dF;
  cudaMalloc(&dA, n * n * 8); cudaMalloc(&dA16, 256 * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&dW, 256 * 8); cudaMalloc(&dT, 64); cudaMalloc(&dF, 4);
  cudaMemset(dF, 0, 4);
  cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dA16, A16.data(), 256 * 8, cudaMemcpyHostToDevice);
  {
    k_micro<<<1, 64>>>(dA16, dL, dW, dT, 200, dF);
    std::vector<double> L(256), W(256); long long cyc;
    cudaMemcpy(L.data(), dL, 256 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(W.data(), dW, 256 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dT, 8, cudaMemcpyDeviceToHost);
    double werr = 0;
    for (int i = 0; i < 16; i++) for (int j = 0; j < 16; j++) { double s = 0; for (int k = 0; k < 16; k++) s += W[i * 16 + k] * L[k * 16 + j]; werr = std::fmax(werr, std::fabs(s - (i == j))); }
    printf("micro block (factor warp + inverse warp): %lld cycles per 16 x 16 block, |L L^T - A| / |A| = %.2e, |W L - I| = %.2e, %s\n", cyc, check_llt(A16, L, 16), werr, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFuncSetAttribute(k_core, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FrontCfg<true>::bytes);
  {
    k_core<<<1, 256, FrontCfg<true>::bytes>>>(dA, dL, dT, 50, dF);
    std::vector<double> L(n * n); long long tm[5];
    cudaMemcpy(L.data(), dL, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(tm, dT, 40, cudaMemcpyDeviceToHost);
    printf("factor_core<96>: %lld cycles (micro blocks %lld, barrier wait %lld, row solve %lld, update %lld), |L L^T - A| / |A| = %.2e, %s\n", tm[4], tm[0], tm[1], tm[2], tm[3],
           check_llt(A, L, n), cudaGetErrorString(cudaGetLastError()));
  }
  int f; cudaMemcpy(&f, dF, 4, cudaMemcpyDeviceToHost);
  printf("fail flag %d\n", f);
  return 0;
}
