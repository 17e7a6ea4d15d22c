/* Which arithmetic does the host's expf perform?  Sweeps EVERY finite float in [-104, 0] and [0, 88] and compares the glibc
 * e_expf.c restatement used by oracle/ndt_oracle.cpp (expf_restated) and lv_slam_b200/csrc/lvs_math.cuh (glibc_expf), in the
 * 16 possible placements of fused multiply-adds, with the libm that is installed.
 *   gcc -O2 -fopenmp -ffp-contract=off -o /tmp/expf_sweep tools/expf_sweep.c -lm && /tmp/expf_sweep      (about a minute on 8 cores)
 * Result in this image (glibc 2.39, x86-64 with FMA; 1 120 927 745 + 1 118 830 593 inputs): variants 8-15 (r = fma(InvLn2N, x, -kd))
 * match bit for bit everywhere, variants 0-7 differ in exactly one input per sign.  The polynomial's own contractions never
 * change a result, so the restatement uses fma throughout (variant 15). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#define N 32
static uint64_t T[N];
static double asd(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static uint64_t asu(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static uint32_t asu32(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static const double SHIFT = 0x1.8p+52, INVLN2N = 0x1.71547652b82fep+0 * N;
static const double C0 = 0x1.c6af84b912394p-5 / N / N / N, C1 = 0x1.ebfce50fac4f3p-3 / N / N, C2 = 0x1.62e42ff0c52d6p-1 / N;
static inline float restated(float x, int v) {
  double xd = x, z = INVLN2N * xd, kd = z + SHIFT;
  uint64_t ki = asu(kd);
  kd -= SHIFT;
  double r = (v & 8) ? fma(INVLN2N, xd, -kd) : z - kd;
  double s = asd(T[ki % N] + (ki << (52 - 5)));
  double zz = (v & 1) ? fma(C0, r, C1) : C0 * r + C1, r2 = r * r, y = (v & 2) ? fma(C2, r, 1.0) : C2 * r + 1;
  y = (v & 4) ? fma(zz, r2, y) : zz * r2 + y;
  return (float)(y * s);
}
int main(void) {
  for (int i = 0; i < N; i++) T[i] = asu(exp2((double)i / N)) - ((uint64_t)i << (52 - 5));
  long bad[16] = {0}, badp[16] = {0}, n = 0, np = 0;
  const uint32_t lo = asu32(-0.0f), hi = asu32(-104.0f);
#pragma omp parallel for reduction(+ : bad[:16]) reduction(+ : n)
  for (uint32_t u = lo; u <= hi; u++) {
    volatile float x = asf(u);
    const float r = expf(x);
    n++;
    for (int v = 0; v < 16; v++) if (asu32(restated(x, v)) != asu32(r)) bad[v]++;
  }
#pragma omp parallel for reduction(+ : badp[:16]) reduction(+ : np)
  for (uint32_t u = 0; u <= asu32(88.0f); u++) {
    volatile float x = asf(u);
    const float r = expf(x);
    np++;
    for (int v = 0; v < 16; v++) if (asu32(restated(x, v)) != asu32(r)) badp[v]++;
  }
  printf("inputs: %ld negative, %ld positive\n", n, np);
  for (int v = 0; v < 16; v++) printf("variant %2d: mismatches %ld (x <= 0), %ld (x >= 0)\n", v, bad[v], badp[v]);
  return 0;
}
