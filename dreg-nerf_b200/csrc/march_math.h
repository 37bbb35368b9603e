/* Per-ray arithmetic of the surface-field marcher (sample_grid.py:245-318 + nerfacc 0.3.5's ray_marching kernel,
 * conerf/utils/nerfacc_utils.py:84-222), ONE definition compiled into both
 *   - the CUDA kernel   dreg-nerf_b200/csrc/ngp.cu      (surface_mask_kernel), and
 *   - the C oracle      oracle/extract_c.c              (march), the checker of the mask parity tests,
 * so that sample positions, the empty-space skip rule, the transmittance recurrence and the termination tests
 * are bit-identical by construction (single precision, fused multiply-add where written, IEEE division).
 * What is NOT shared is the density network behind sigma: the kernel evaluates its first layer on the tensor
 * cores (3xTF32, ~2^-21), the oracle sums sequentially in fp32 - the only source of differing mask decisions,
 * confined to |alpha T - cut_off| of a few 1e-6.
 *
 * Plain C99 + optional CUDA qualifiers. */
#ifndef DRB_MARCH_MATH_H_
#define DRB_MARCH_MATH_H_

#include <math.h>

#ifdef __CUDACC__
#define DRB_MM __host__ __device__ __forceinline__
#else
#define DRB_MM static inline
#endif

#define DRB_EARLY_STOP_EPS 1e-4f      /* nerfacc_utils.py:209 */

/* a / b correctly rounded.  Device: y = RN(1 / b) is precomputed and q = a y, r = fma(-b, q, a), q + r y is the
 * IEEE quotient (Markstein; exact unless b's significand is all ones, which the host refuses) - same bits as the
 * host's division, without the slow path. */
DRB_MM float drb_div(float a, float b, float b_rcp) {
#ifdef __CUDA_ARCH__
  const float q = a * b_rcp;
  const float r = fmaf(-b, q, a);
  return fmaf(r, b_rcp, q);
#else
  (void)b_rcp;
  return a / b;
#endif
}

/* Ray from camera o to point p clipped to the scene box: unit direction, its reciprocal, t_max = |p - o| and the
 * first interval [t0, t1] with its midpoint tm.  Returns 0 when nothing is to be marched. */
DRB_MM int drb_ray_begin(const float o[3], const float p[3], const float scene_min[3], const float scene_max[3],
                         float step, float dir[3], float inv[3], float* len, float* t0, float* t1, float* tm) {
  for (int d = 0; d < 3; ++d) dir[d] = p[d] - o[d];
  /* explicit fused multiply-adds: the same bits whether or not the compiler contracts a * b + c */
  const float l = sqrtf(fmaf(dir[2], dir[2], fmaf(dir[1], dir[1], dir[0] * dir[0])));
  if (!(l > 0.f)) return 0;
  for (int d = 0; d < 3; ++d) { dir[d] = dir[d] / l; inv[d] = 1.f / dir[d]; }
  float tn = -1e30f, tf = 1e30f;
  for (int d = 0; d < 3; ++d) {
    float ta = (scene_min[d] - o[d]) * inv[d], tb = (scene_max[d] - o[d]) * inv[d];
    if (ta > tb) { const float t = ta; ta = tb; tb = t; }
    tn = fmaxf(tn, ta); tf = fminf(tf, tb);
  }
  if (tn > tf) return 0;
  *len = l;
  *t0 = fmaxf(tn, 0.f); *t1 = *t0 + step; *tm = 0.5f * (*t0 + *t1);
  return 1;
}

/* interval after a sample that did not end the ray */
DRB_MM void drb_chain_next(float step, float* t0, float* t1, float* tm) {
  *t0 = *t1; *t1 = *t0 + step; *tm = 0.5f * (*t0 + *t1);
}

/* sample position x = o + tm dir (one fused multiply-add per axis) */
DRB_MM void drb_sample_pos(float tm, const float dir[3], const float o[3], float x[3]) {
  for (int d = 0; d < 3; ++d) x[d] = fmaf(tm, dir[d], o[d]);
}

/* unit coordinates u = (x - roi_min) / extent and the voxel that contains them; returns 0 outside the ROI */
DRB_MM int drb_voxel_of(const float x[3], const float roi_min[3], const float roi_ext[3], const float roi_rcp[3],
                        int res, float u[3], int idx[3]) {
  int in = 1;
  for (int d = 0; d < 3; ++d) {
    u[d] = drb_div(x[d] - roi_min[d], roi_ext[d], roi_rcp[d]);
    in = in && (u[d] >= 0.f) && (u[d] < 1.f);
  }
  if (!in) return 0;
  for (int d = 0; d < 3; ++d) {
    const int i = (int)(u[d] * (float)res);
    idx[d] = i < 0 ? 0 : (i > res - 1 ? res - 1 : i);
  }
  return 1;
}

/* Empty voxel at tm: advance to the first sample position past the voxel's exit face (nerfacc's
 * distance_to_next_voxel + the while loop of its marching kernel). */
DRB_MM void drb_skip_empty(const float u[3], const float dir[3], const float inv[3], int res, float res_rcp,
                           const float roi_ext[3], float step, float* t0, float* t1, float* tm) {
  float dist = 1e30f;
  for (int d = 0; d < 3; ++d) {
    const float ur = u[d] * (float)res;
    const float sgn = dir[d] > 0.f ? 1.f : (dir[d] < 0.f ? -1.f : 0.f);
    const float td = drb_div((floorf(ur + 0.5f + 0.5f * sgn) - ur) * inv[d], (float)res, res_rcp);
    dist = fminf(dist, td * roi_ext[d]);
  }
  const float tt = *tm + fmaxf(dist, 0.f);
  do { *tm += step; } while (*tm < tt);
  *t0 = *tm - 0.5f * step; *t1 = *tm + 0.5f * step;
}

DRB_MM float drb_alpha(float sigma, float dt) { return 1.f - expf(-sigma * dt); }

/* One sample of a ray in marching order.  Returns 0: go on to the next interval, 1: the surface-field value
 * reached cut_off (the point is seen), 2: transmittance below early_stop_eps (the sample is dropped, the ray ends). */
DRB_MM int drb_accumulate(float alpha, float cut_off, float* T, float* best) {
  if (*T < DRB_EARLY_STOP_EPS) return 2;
  *best = fmaxf(*best, alpha * *T);
  if (*best >= cut_off) return 1;
  *T *= 1.f - alpha;
  return 0;
}

#endif /* DRB_MARCH_MATH_H_ */
