#include "../../oracle/extract_c.c"
/* per (point, camera): best, samples marched with the exact early-out (T < cut_off), occupied samples count total along ray */
void exp_stats(const float* table, const float* w1, const float* w2, const float* aabb, const uint8_t* occ,
               int res, const float* roi, const float* scene, const float* pts, int n, const float* cams,
               int ncams, float step, float cut_off, float* best_out, int* samples_out, int* occlen_out, float* firstocc_out) {
  field_t f;
  f.table = table; f.w1 = w1; f.w2 = w2;
  memcpy(f.aabb, aabb, sizeof(f.aabb));
  level_table(&f);
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < n; ++i) {
    for (int c = 0; c < ncams; ++c) {
      const float* o = cams + 3 * c; const float* p = pts + 3 * (size_t)i;
      float dir[3], inv[3], len, t0, t1, tm;
      const float roi_ext[3] = {roi[3] - roi[0], roi[4] - roi[1], roi[5] - roi[2]};
      const float roi_rcp[3] = {0.f, 0.f, 0.f};
      float best = 0.f; int ns = 0, nocc = 0; float first = -1.f;
      if (drb_ray_begin(o, p, scene, scene + 3, step, dir, inv, &len, &t0, &t1, &tm)) {
        float T = 1.f; int done = 0;
        while (tm < len) {
          float x[3], u[3]; int idx[3];
          drb_sample_pos(tm, dir, o, x);
          if (drb_voxel_of(x, roi, roi_ext, roi_rcp, res, u, idx) && occ[((size_t)idx[0] * res + idx[1]) * res + idx[2]] != 0) {
            ++nocc; if (first < 0) first = tm;
            if (!done) {
              const float sigma = density_at(&f, x);
              const float alpha = drb_alpha(sigma, t1 - t0);
              ++ns;
              if (drb_accumulate(alpha, cut_off, &T, &best)) done = 1;
              else if (T < cut_off) done = 1;
            }
            drb_chain_next(step, &t0, &t1, &tm);
          } else {
            drb_skip_empty(u, dir, inv, res, 0.f, roi_ext, step, &t0, &t1, &tm);
          }
        }
      }
      best_out[(size_t)i * ncams + c] = best; samples_out[(size_t)i * ncams + c] = ns; occlen_out[(size_t)i * ncams + c] = nocc;
      firstocc_out[(size_t)i * ncams + c] = len - first;
    }
  }
}
