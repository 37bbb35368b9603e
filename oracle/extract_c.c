/* CPU restatement in C of the extract half's arithmetic (TEST INFRASTRUCTURE, PARITY UNPINNED):
 *   A1  NGPradianceField.query_density      conerf/radiance_fields/ngp.py:148-176  (tiny-cuda-nn hash grid +
 *       MLP, restated from the published algorithm exactly as oracle/ngp.py does)
 *   A5  the surface-field mask             conerf/register/sample_grid.py:245-318, the marching of
 *       conerf/utils/nerfacc_utils.py:84-222 (nerfacc 0.3.5 ray_marching: fixed step, occupancy-grid
 *       skipping) and conerf/loss/confidence_loss.py:93-157
 * Single precision throughout, one operation per line of the scalar Python oracle (oracle/extract.py,
 * oracle/ngp.py); the sample position uses one fused multiply-add per axis as nerfacc's compiled kernel
 * does.  OpenMP over points.  Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline legs may
 * load this library; the product never does.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC -> oracle/_c/liboracle_extract.so)
 */
#include <math.h>
#include <stdint.h>

#include "march_math.h"   /* dreg-nerf_b200/csrc: the per-ray arithmetic shared with the CUDA kernel */
#include <stdlib.h>
#include <string.h>

#define N_LEVELS 16
#define T_SIZE (1u << 19)

typedef struct {
  const float* table; /* [entries][2] */
  const float* w1;    /* [64][32] */
  const float* w2;    /* [16][64] */
  float aabb[6];
  float scale[N_LEVELS];
  uint32_t res[N_LEVELS], size[N_LEVELS], offset[N_LEVELS];
} field_t;

/* level table of oracle/ngp.py level_table(): scale evaluated in double, rounded once to fp32 */
static void level_table(field_t* f) {
  const double b = 1.4472692012786865;
  uint32_t off = 0;
  for (int l = 0; l < N_LEVELS; ++l) {
    const float scale = (float)(exp2((double)l * log2(b)) * 16.0 - 1.0);
    const uint32_t res = (uint32_t)ceilf(scale) + 1u;
    uint64_t n = (uint64_t)res * res * res;
    n = (n + 7) / 8 * 8;
    if (n > T_SIZE) n = T_SIZE;
    f->scale[l] = scale; f->res[l] = res; f->size[l] = (uint32_t)n; f->offset[l] = off;
    off += (uint32_t)n;
  }
}

static uint32_t grid_index(const uint32_t g[3], uint32_t res, uint32_t size) {
  uint32_t stride = 1, index = 0;
  for (int d = 0; d < 3; ++d) {
    if (stride <= size) { index += g[d] * stride; stride *= res; }
  }
  if (size < stride) index = (g[0] * 1u) ^ (g[1] * 2654435761u) ^ (g[2] * 805459861u);
  return index % size;
}

/* ngp.py:148-176: density = exp(out[0] - 1) * all(0 < xn < 1) */
static float density_at(const field_t* f, const float x[3]) {
  float xn[3];
  int inside = 1;
  for (int d = 0; d < 3; ++d) {
    xn[d] = (x[d] - f->aabb[d]) / (f->aabb[3 + d] - f->aabb[d]);
    inside = inside && (xn[d] > 0.f) && (xn[d] < 1.f);
  }
  if (!inside) return 0.f;
  float feat[2 * N_LEVELS];
  for (int l = 0; l < N_LEVELS; ++l) {
    float frac[3];
    uint32_t g0[3];
    for (int d = 0; d < 3; ++d) {
      const float pos = fmaf(xn[d], f->scale[l], 0.5f);
      const float fl = floorf(pos);
      frac[d] = pos - fl;
      g0[d] = (uint32_t)(int)fl;
    }
    float a0 = 0.f, a1 = 0.f;
    for (int c = 0; c < 8; ++c) {
      uint32_t g[3];
      float w = 1.f;
      for (int d = 0; d < 3; ++d) {
        if (c & (1 << d)) { g[d] = g0[d] + 1u; w *= frac[d]; }
        else { g[d] = g0[d]; w *= 1.f - frac[d]; }
      }
      const float* e = f->table + 2 * (size_t)(f->offset[l] + grid_index(g, f->res[l], f->size[l]));
      a0 = fmaf(w, e[0], a0);
      a1 = fmaf(w, e[1], a1);
    }
    feat[2 * l] = a0;
    feat[2 * l + 1] = a1;
  }
  float out0 = 0.f;
  for (int j = 0; j < 64; ++j) {
    float h = 0.f;
    for (int i = 0; i < 32; ++i) h = fmaf(f->w1[j * 32 + i], feat[i], h);
    h = h > 0.f ? h : 0.f;
    out0 = fmaf(f->w2[j], h, out0);
  }
  return expf(out0 - 1.f);
}

void orc_density(const float* table, const float* w1, const float* w2, const float* aabb, const float* x, int n,
                 float* out) {
  field_t f;
  f.table = table; f.w1 = w1; f.w2 = w2;
  memcpy(f.aabb, aabb, sizeof(f.aabb));
  level_table(&f);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) out[i] = density_at(&f, x + 3 * (size_t)i);
}

/* One ray camera -> point: returns the largest alpha * T reached (early exit at cut_off), counts samples.
 * Every arithmetic step is a function of dreg-nerf_b200/csrc/march_math.h, the header the CUDA kernel compiles
 * too: positions, skips and the transmittance recurrence are bit-identical by construction. */
static float march(const field_t* f, const uint8_t* occ, int res, const float roi[6], const float scene[6],
                   const float o[3], const float p[3], float step, float cut_off, long long* samples) {
  float dir[3], inv[3], len, t0, t1, tm;
  const float roi_ext[3] = {roi[3] - roi[0], roi[4] - roi[1], roi[5] - roi[2]};
  const float roi_rcp[3] = {0.f, 0.f, 0.f};     /* unused on the host: drb_div divides */
  if (!drb_ray_begin(o, p, scene, scene + 3, step, dir, inv, &len, &t0, &t1, &tm)) return 0.f;
  float T = 1.f, best = 0.f;
  while (tm < len) {
    float x[3], u[3];
    int idx[3];
    drb_sample_pos(tm, dir, o, x);
    if (drb_voxel_of(x, roi, roi_ext, roi_rcp, res, u, idx) && occ[((size_t)idx[0] * res + idx[1]) * res + idx[2]] != 0) {
      const float sigma = density_at(f, x);
      const float alpha = drb_alpha(sigma, t1 - t0);
      ++*samples;
      if (drb_accumulate(alpha, cut_off, &T, &best)) break;
      drb_chain_next(step, &t0, &t1, &tm);
    } else {
      drb_skip_empty(u, dir, inv, res, 0.f, roi_ext, step, &t0, &t1, &tm);
    }
  }
  return best;
}

/* surface[p] = any camera sees max_s(alpha_s * T_s) >= cut_off on its ray to point p.
 * active: optional per-point flags (rays only for flagged points).  all_rays != 0: march every (camera, point)
 * ray as the reference does (timing baseline); otherwise stop at the first camera that sees the point.
 * best_out (optional): largest surface-field value any marched ray of the point reached. */
void orc_surface_mask(const float* table, const float* w1, const float* w2, const float* aabb, const uint8_t* occ,
                      int res, const float* roi, const float* scene, const float* pts, int n, const float* cams,
                      int ncams, float step, float cut_off, const uint8_t* active, int all_rays, uint8_t* surface,
                      float* best_out, long long* n_samples) {
  field_t f;
  f.table = table; f.w1 = w1; f.w2 = w2;
  memcpy(f.aabb, aabb, sizeof(f.aabb));
  level_table(&f);
  long long total = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
  for (int i = 0; i < n; ++i) {
    surface[i] = 0;
    float bmax = 0.f;
    if (!active || active[i]) {
      for (int c = 0; c < ncams; ++c) {
        if (surface[i] && !all_rays) break;
        long long s = 0;
        const float b = march(&f, occ, res, roi, scene, cams + 3 * (size_t)c, pts + 3 * (size_t)i, step, cut_off, &s);
        total += s;
        bmax = fmaxf(bmax, b);
        if (b >= cut_off) surface[i] = 1;
      }
    }
    if (best_out) best_out[i] = bmax;
  }
  if (n_samples) *n_samples = total;
}
