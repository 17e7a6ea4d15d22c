/* Synthetic Velodyne-like scan generator (bench + test input).
 *
 * Not part of the reference and not part of the oracle: it only manufactures inputs, identically for the
 * CUDA path, the oracle and the CPU baseline.  A procedural street canyon (infinite along +x) is ray-cast
 * analytically from a sensor pose:
 *   ground plane z = -1.73 m (sensor height of the KITTI rig the reference targets),
 *   one building box per 20 m segment per side (footprint 6-18 m x 8-20 m, height 3-15 m, setback 6-20 m),
 *   one pole / tree trunk (r = 0.15-0.4 m, h = 3-8 m) per 8 m per side, one car-sized box per 25 m per side,
 *   one kerb-side clutter box (0.6-2.5 m) per 10 m per side.
 * Beams: n_beams elevations linearly from +2.0 deg to -24.9 deg (the reference's own HDL-64E constants,
 * include/ndt_pca/voxel_grid_covariance_pca.h:97,113), n_az azimuth steps per revolution, azimuth-major
 * order.  Returns with range in (0.5, 100) m are kept (launch/dlo_kitti.launch:31-32); Gaussian range
 * noise sigma is applied along the ray.  The RNG is counter based (splitmix64 of seed/frame/ray) so the
 * output does not depend on the thread count.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
static inline double u01(uint64_t h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline double hash_u(uint64_t seed, int64_t a, int64_t b, int64_t c) {
  uint64_t h = splitmix64(seed ^ splitmix64((uint64_t)a * 0x100000001B3ull + 17));
  h = splitmix64(h ^ splitmix64((uint64_t)b + 0x51ED27ull));
  h = splitmix64(h ^ (uint64_t)c);
  return u01(h);
}

typedef struct { double lo[3], hi[3]; } Box;
typedef struct { double cx, cy, r, z0, z1; } Cyl;

#define GROUND_Z (-1.73)
#define MAX_RANGE 100.0
#define MIN_RANGE 0.5

static int build_world(uint64_t seed, double sx, Box* boxes, int* nb, Cyl* cyls, int* nc) {
  int b = 0, c = 0;
  const double reach = MAX_RANGE + 45.0;
  for (int side = -1; side <= 1; side += 2) {
    int s0 = (int)floor((sx - reach) / 20.0), s1 = (int)floor((sx + reach) / 20.0);
    for (int s = s0; s <= s1; s++) {
      double w = 6.0 + 12.0 * hash_u(seed, s, side, 1);
      double d = 8.0 + 12.0 * hash_u(seed, s, side, 2);
      double h = 3.0 + 12.0 * hash_u(seed, s, side, 3);
      double sb = 6.0 + 14.0 * hash_u(seed, s, side, 4);
      double cx = 20.0 * s + 10.0 + 2.0 * (hash_u(seed, s, side, 5) - 0.5);
      Box* B = &boxes[b++];
      B->lo[0] = cx - 0.5 * w; B->hi[0] = cx + 0.5 * w;
      if (side > 0) { B->lo[1] = sb; B->hi[1] = sb + d; } else { B->lo[1] = -sb - d; B->hi[1] = -sb; }
      B->lo[2] = GROUND_Z; B->hi[2] = GROUND_Z + h;
    }
    s0 = (int)floor((sx - reach) / 25.0); s1 = (int)floor((sx + reach) / 25.0);
    for (int s = s0; s <= s1; s++) {
      double cx = 25.0 * s + 25.0 * hash_u(seed, s, side, 11);
      double cy = side * (2.5 + 1.0 * hash_u(seed, s, side, 12));
      Box* B = &boxes[b++];
      B->lo[0] = cx - 2.25; B->hi[0] = cx + 2.25;
      B->lo[1] = cy - 0.9; B->hi[1] = cy + 0.9;
      B->lo[2] = GROUND_Z; B->hi[2] = GROUND_Z + 1.5;
    }
    s0 = (int)floor((sx - reach) / 10.0); s1 = (int)floor((sx + reach) / 10.0);
    for (int s = s0; s <= s1; s++) {
      double e = 0.6 + 1.9 * hash_u(seed, s, side, 31);
      double cx = 10.0 * s + 10.0 * hash_u(seed, s, side, 32);
      double cy = side * (4.2 + 1.6 * hash_u(seed, s, side, 33));
      Box* B = &boxes[b++];
      B->lo[0] = cx - 0.5 * e; B->hi[0] = cx + 0.5 * e;
      B->lo[1] = cy - 0.5 * e; B->hi[1] = cy + 0.5 * e;
      B->lo[2] = GROUND_Z; B->hi[2] = GROUND_Z + 0.4 + 1.6 * hash_u(seed, s, side, 34);
    }
    s0 = (int)floor((sx - reach) / 8.0); s1 = (int)floor((sx + reach) / 8.0);
    for (int s = s0; s <= s1; s++) {
      Cyl* C = &cyls[c++];
      C->cx = 8.0 * s + 8.0 * hash_u(seed, s, side, 21);
      C->cy = side * (4.5 + 1.5 * hash_u(seed, s, side, 22));
      C->r = 0.15 + 0.25 * hash_u(seed, s, side, 23); C->z0 = GROUND_Z; C->z1 = GROUND_Z + 3.0 + 5.0 * hash_u(seed, s, side, 24);
    }
  }
  *nb = b; *nc = c;
  return 0;
}

static inline double ray_box(const double o[3], const double d[3], const Box* B) {
  double t0 = 0.0, t1 = 1e30;
  for (int a = 0; a < 3; a++) {
    if (fabs(d[a]) < 1e-12) { if (o[a] < B->lo[a] || o[a] > B->hi[a]) return 1e30; continue; }
    double inv = 1.0 / d[a];
    double ta = (B->lo[a] - o[a]) * inv, tb = (B->hi[a] - o[a]) * inv;
    if (ta > tb) { double t = ta; ta = tb; tb = t; }
    if (ta > t0) t0 = ta;
    if (tb < t1) t1 = tb;
    if (t0 > t1) return 1e30;
  }
  return t0 > 0.0 ? t0 : 1e30;
}

static inline double ray_cyl(const double o[3], const double d[3], const Cyl* C) {
  double ox = o[0] - C->cx, oy = o[1] - C->cy;
  double a = d[0] * d[0] + d[1] * d[1];
  if (a < 1e-14) return 1e30;
  double b = ox * d[0] + oy * d[1], c = ox * ox + oy * oy - C->r * C->r;
  double disc = b * b - a * c;
  if (disc < 0) return 1e30;
  double t = (-b - sqrt(disc)) / a;
  if (t <= 0) return 1e30;
  double z = o[2] + t * d[2];
  return (z >= C->z0 && z <= C->z1) ? t : 1e30;
}

/* pose6 = x y z yaw pitch roll (world <- sensor, R = Rz(yaw) Ry(pitch) Rx(roll)).  out_xyz holds up to
 * n_beams*n_az points (sensor frame, packed xyz float).  Returns the number of points written. */
int synth_scan(uint64_t seed, int frame, const double* pose6, int n_beams, int n_az, double noise_sigma, float* out_xyz, int nthreads) {
  Box boxes[160]; Cyl cyls[96]; int nb = 0, nc = 0;
  build_world(seed, pose6[0], boxes, &nb, cyls, &nc);
  double cy = cos(pose6[3]), sy = sin(pose6[3]), cp = cos(pose6[4]), sp = sin(pose6[4]), cr = cos(pose6[5]), sr = sin(pose6[5]);
  double R[3][3] = {{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr},
                    {sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr},
                    {-sp, cp * sr, cp * cr}};
  const long total = (long)n_beams * n_az;
  float* tmp = (float*)malloc(sizeof(float) * 3 * (size_t)total);
  unsigned char* ok = (unsigned char*)malloc((size_t)total);
  const double el_top = 2.0 * M_PI / 180.0, el_bot = -24.9 * M_PI / 180.0;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (long k = 0; k < total; k++) {
    int j = (int)(k / n_beams), i = (int)(k % n_beams);
    double el = n_beams > 1 ? el_top + (el_bot - el_top) * (double)i / (double)(n_beams - 1) : el_top;
    double az = 2.0 * M_PI * (double)j / (double)n_az;
    double ds[3] = {cos(el) * cos(az), cos(el) * sin(az), sin(el)};
    double d[3], o[3] = {pose6[0], pose6[1], pose6[2]};
    for (int a = 0; a < 3; a++) d[a] = R[a][0] * ds[0] + R[a][1] * ds[1] + R[a][2] * ds[2];
    double t = 1e30;
    if (d[2] < -1e-9) t = (GROUND_Z - o[2]) / d[2];
    for (int b = 0; b < nb; b++) { double tb = ray_box(o, d, &boxes[b]); if (tb < t) t = tb; }
    for (int c = 0; c < nc; c++) { double tc = ray_cyl(o, d, &cyls[c]); if (tc < t) t = tc; }
    ok[k] = 0;
    if (t < 1e29) {
      uint64_t h1 = splitmix64(seed ^ splitmix64(((uint64_t)(uint32_t)frame << 32) ^ (uint64_t)k));
      uint64_t h2 = splitmix64(h1 ^ 0xD1B54A32D192ED03ull);
      double g = sqrt(-2.0 * log(u01(h1))) * cos(2.0 * M_PI * u01(h2));
      double r = t + noise_sigma * g;
      if (r > MIN_RANGE && r < MAX_RANGE) {
        tmp[3 * k] = (float)(r * ds[0]); tmp[3 * k + 1] = (float)(r * ds[1]); tmp[3 * k + 2] = (float)(r * ds[2]);
        ok[k] = 1;
      }
    }
  }
  long n = 0;
  for (long k = 0; k < total; k++)
    if (ok[k]) { out_xyz[3 * n] = tmp[3 * k]; out_xyz[3 * n + 1] = tmp[3 * k + 1]; out_xyz[3 * n + 2] = tmp[3 * k + 2]; n++; }
  free(tmp); free(ok);
  return (int)n;
}

/* Stream trajectory (configs 2 and 5): 1.2 m per frame along +x, a +-2 m lateral weave and +-15 deg yaw
 * oscillation with a 200-frame period, +-0.5 deg pitch ripple. */
void synth_traj_pose(int frame, double* pose6) {
  double ph = 2.0 * M_PI * (double)frame / 200.0;
  pose6[0] = 1.2 * (double)frame;
  pose6[1] = 2.0 * sin(ph);
  pose6[2] = 0.0;
  pose6[3] = (15.0 * M_PI / 180.0) * sin(ph);
  pose6[4] = (0.5 * M_PI / 180.0) * sin(2.0 * M_PI * (double)frame / 37.0);
  pose6[5] = 0.0;
}
