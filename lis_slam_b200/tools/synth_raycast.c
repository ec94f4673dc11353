/* Synthetic-workload generator (NOT part of the registration path, NOT an oracle): a C restatement of
 * lis_slam_b200/synth.py Scene.raycast so that the 600-frame HDL-64 / 1000-frame VLP-16 streams of
 * BASELINE.json configs[1] / configs[4] can be generated in seconds instead of minutes.  Same geometry, same
 * primitive order, same strict "nearer hit wins" rule; double precision like the numpy code.
 * Built by __graft_entry__.build() into lis_slam_b200/libsynth.so and loaded by synth.py (scan_fast). */
#include <math.h>
#include <stdint.h>

static inline double nmin(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }   /* np.minimum: NaN propagates */
static inline double nmax(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }

void synth_raycast(const double* origin, const double* dirs, int64_t n,
                   const double* boxes /* nb x 6 */, const uint16_t* box_labels, int32_t nb,
                   const double* poles /* np x 2 */, int32_t np_, double pole_r, double pole_h,
                   double ground_z, double extent, double* out_range, uint16_t* out_label) {
  const double reach = 80.0;
  const double ox0 = origin[0], oy0 = origin[1], oz0 = origin[2];
  /* boxes / poles farther than the sensor reach are culled once */
  int keep_box[4096]; int nkb = 0;
  for (int b = 0; b < nb && nkb < 4096; b++) {
    const double* B = boxes + 6 * b;
    double c[3];
    for (int k = 0; k < 3; k++) { double v = origin[k]; if (v < B[k]) v = B[k]; if (v > B[3 + k]) v = B[3 + k]; c[k] = v - origin[k]; }
    if (sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) > reach) continue;
    keep_box[nkb++] = b;
  }
  int keep_pole[8192]; int nkp = 0;
  for (int p = 0; p < np_ && nkp < 8192; p++) {
    const double ox = ox0 - poles[2 * p], oy = oy0 - poles[2 * p + 1];
    const double dist = hypot(ox, oy);
    if (dist > reach || dist <= pole_r) continue;
    keep_pole[nkp++] = p;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    const double dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
    double best = INFINITY; uint16_t lab = 0;
    {
      const double t = (ground_z - oz0) / dz;
      const double hx = ox0 + t * dx, hy = oy0 + t * dy;
      if (t > 0.5 && fabs(hx) < extent + 5 && fabs(hy) < extent + 5) { best = t; lab = 9; }
    }
    const double inv[3] = {1.0 / dx, 1.0 / dy, 1.0 / dz};
    for (int kb = 0; kb < nkb; kb++) {
      const double* B = boxes + 6 * keep_box[kb];
      double tmin = -INFINITY, tmax = INFINITY;
      for (int k = 0; k < 3; k++) {
        const double t1 = (B[k] - origin[k]) * inv[k], t2 = (B[3 + k] - origin[k]) * inv[k];
        tmin = nmax(tmin, nmin(t1, t2)); tmax = nmin(tmax, nmax(t1, t2));
      }
      if (tmax >= tmin && tmin > 0.5 && tmin < best) { best = tmin; lab = box_labels[keep_box[kb]]; }
    }
    const double a = dx * dx + dy * dy;
    for (int kp = 0; kp < nkp; kp++) {
      const int p = keep_pole[kp];
      const double ox = ox0 - poles[2 * p], oy = oy0 - poles[2 * p + 1];
      const double bq = ox * dx + oy * dy;
      const double cq = ox * ox + oy * oy - pole_r * pole_r;
      const double disc = bq * bq - a * cq;
      if (!(disc > 0)) continue;
      const double t = (-bq - sqrt(disc)) / a;
      const double z = oz0 + t * dz;
      if (t > 0.5 && t < best && z >= ground_z && z <= ground_z + pole_h) { best = t; lab = 18; }
    }
    out_range[i] = best; out_label[i] = lab;
  }
}
