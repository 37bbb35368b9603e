// R8 + R9: weighted Procrustes / Kabsch over the stacked correspondences
// (conerf/register/nerf_regtr.py:208-230, conerf/register/se3.py:89-140).
// One block per decoder layer: warp-shuffle reductions of the weighted centroids and the 3x3
// covariance in double, then a one-sided Jacobi SVD of the 3x3 matrix on one thread (replaces
// the cuSOLVER batched SVD + host-synchronising weight assert of the reference).
#include "common.cuh"
#include "svd3.cuh"

namespace drb {

struct ProcArgs {
  const float *a1, *b1, *w1, *a2, *b2, *w2;
  long long a1_ls, b1_ls, w1_ls, a2_ls, b2_ls, w2_ls;
  int n1, n2, ld;
  float* out;
};

__device__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) t += sh[i];
  return t;
}

__global__ void __launch_bounds__(256) procrustes_kernel(ProcArgs p) {
  __shared__ double sh[8];
  const int layer = blockIdx.x;
  const float* a1 = p.a1 + layer * p.a1_ls; const float* b1 = p.b1 + layer * p.b1_ls;
  const float* w1 = p.w1 + layer * p.w1_ls;
  const float* a2 = p.a2 + layer * p.a2_ls; const float* b2 = p.b2 + layer * p.b2_ls;
  const float* w2 = p.w2 + layer * p.w2_ls;
  const int n = p.n1 + p.n2;
  auto fetch = [&](int i, float& w, float a[3], float b[3]) {
    const float *pa, *pb;
    if (i < p.n1) { pa = a1 + (long long)i * p.ld; pb = b1 + (long long)i * p.ld; w = w1[i]; }
    else { const int j = i - p.n1; pa = a2 + (long long)j * p.ld; pb = b2 + (long long)j * p.ld; w = w2[j]; }
    a[0] = pa[0]; a[1] = pa[1]; a[2] = pa[2];
    b[0] = pb[0]; b[1] = pb[1]; b[2] = pb[2];
  };
  double sw = 0, sa[3] = {0, 0, 0}, sb[3] = {0, 0, 0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float w, a[3], b[3];
    fetch(i, w, a, b);
    sw += w;
    for (int d = 0; d < 3; ++d) { sa[d] += (double)w * a[d]; sb[d] += (double)w * b[d]; }
  }
  sw = block_sum(sw, sh);
  for (int d = 0; d < 3; ++d) { sa[d] = block_sum(sa[d], sh); sb[d] = block_sum(sb[d], sh); }
  const double denom = sw > 1e-6 ? sw : 1e-6;   // clamp_min(sum w, eps), se3.py:113-114
  double ca[3], cb[3];
  for (int d = 0; d < 3; ++d) { ca[d] = sa[d] / denom; cb[d] = sb[d] / denom; }
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float w, a[3], b[3];
    fetch(i, w, a, b);
    const double wn = (double)w / denom;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) cov[r * 3 + c] += ((double)a[r] - ca[r]) * (((double)b[c] - cb[c]) * wn);
  }
  __shared__ double scov[9];
  for (int i = 0; i < 9; ++i) {
    const double tot = block_sum(cov[i], sh);
    if (threadIdx.x == 0) scov[i] = tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // the reduced covariance goes through shared memory and svd3 stays out of line: the fully
    // inlined register-only variant was observed to miscompile (wrong rotation angle) with nvcc 12.9
    for (int i = 0; i < 9; ++i) cov[i] = scov[i];
    double U[9], S[3], V[9];
    svd3(cov, U, S, V);
    double R[9];
    auto build = [&](double sign) {
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          R[r * 3 + c] = V[r * 3 + 0] * U[c * 3 + 0] + V[r * 3 + 1] * U[c * 3 + 1] +
                         sign * V[r * 3 + 2] * U[c * 3 + 2];
    };
    build(1.0);
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) +
                       R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (!(det > 0)) build(-1.0);
    float* o = p.out + layer * 12;
    for (int r = 0; r < 3; ++r) {
      const double t = -(R[r * 3 + 0] * ca[0] + R[r * 3 + 1] * ca[1] + R[r * 3 + 2] * ca[2]) + cb[r];
      o[r * 4 + 0] = (float)R[r * 3 + 0];
      o[r * 4 + 1] = (float)R[r * 3 + 1];
      o[r * 4 + 2] = (float)R[r * 3 + 2];
      o[r * 4 + 3] = (float)t;
    }
  }
}

extern "C" int drb_procrustes(const float* a1, long long a1_ls, const float* b1, long long b1_ls,
                              const float* w1, long long w1_ls, int n1, const float* a2,
                              long long a2_ls, const float* b2, long long b2_ls, const float* w2,
                              long long w2_ls, int n2, int ld_pts, int layers, float* out,
                              cudaStream_t stream) {
  DRB_REQUIRE(out && layers > 0 && ld_pts >= 3, "drb_procrustes: bad arguments");
  DRB_REQUIRE(n1 == 0 || (a1 && b1 && w1), "drb_procrustes: segment 1 pointers");
  DRB_REQUIRE(n2 == 0 || (a2 && b2 && w2), "drb_procrustes: segment 2 pointers");
  ProcArgs p;
  p.a1 = a1; p.b1 = b1; p.w1 = w1; p.a2 = a2; p.b2 = b2; p.w2 = w2;
  p.a1_ls = a1_ls; p.b1_ls = b1_ls; p.w1_ls = w1_ls; p.a2_ls = a2_ls; p.b2_ls = b2_ls; p.w2_ls = w2_ls;
  p.n1 = n1; p.n2 = n2; p.ld = ld_pts; p.out = out;
  procrustes_kernel<<<layers, 256, 0, stream>>>(p);
  DRB_LAUNCH_OK();
  return 0;
}

}  // namespace drb
