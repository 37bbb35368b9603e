// R8 + R9: weighted Procrustes / Kabsch over the stacked correspondences
// (conerf/register/nerf_regtr.py:208-230, conerf/register/se3.py:89-140).
// One block per decoder layer: warp-shuffle reductions of the weighted centroids and the 3x3
// covariance in double, then a one-sided Jacobi SVD of the 3x3 matrix on one thread (replaces
// the cuSOLVER batched SVD + host-synchronising weight assert of the reference).
#include "common.cuh"

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

// A (3x3, row major) = U diag(S) V^T, S descending.
__host__ __device__ __noinline__ void svd3(const double A_in[9], double U[9], double S[3], double V[9]) {
  double A[9];
  for (int i = 0; i < 9; ++i) { A[i] = A_in[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int r = 0; r < 3; ++r) {
          alpha += A[r * 3 + p] * A[r * 3 + p];
          beta += A[r * 3 + q] * A[r * 3 + q];
          gamma += A[r * 3 + p] * A[r * 3 + q];
        }
        if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 1e-17 * sqrt(alpha * beta)) continue;
        off += fabs(gamma);
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int r = 0; r < 3; ++r) {
          const double ap = A[r * 3 + p], aq = A[r * 3 + q];
          A[r * 3 + p] = c * ap - s * aq;
          A[r * 3 + q] = s * ap + c * aq;
          const double vp = V[r * 3 + p], vq = V[r * 3 + q];
          V[r * 3 + p] = c * vp - s * vq;
          V[r * 3 + q] = s * vp + c * vq;
        }
      }
    if (off == 0.0) break;
  }
  double sig[3];
  for (int j = 0; j < 3; ++j)
    sig[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
  int order[3] = {0, 1, 2};
  for (int i = 0; i < 2; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (sig[order[j]] > sig[order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
  double Vs[9];
  for (int j = 0; j < 3; ++j) {
    const int src = order[j];
    S[j] = sig[src];
    for (int r = 0; r < 3; ++r) {
      Vs[r * 3 + j] = V[r * 3 + src];
      U[r * 3 + j] = sig[src] > 0 ? A[r * 3 + src] / sig[src] : 0.0;
    }
  }
  for (int i = 0; i < 9; ++i) V[i] = Vs[i];
  // complete null directions so that U and V stay orthonormal when the covariance is rank deficient
  const double tiny = 1e-12 * (S[0] > 0 ? S[0] : 1.0);
  auto cross_col = [](double* M, int a, int b, int c) {
    M[0 * 3 + c] = M[1 * 3 + a] * M[2 * 3 + b] - M[2 * 3 + a] * M[1 * 3 + b];
    M[1 * 3 + c] = M[2 * 3 + a] * M[0 * 3 + b] - M[0 * 3 + a] * M[2 * 3 + b];
    M[2 * 3 + c] = M[0 * 3 + a] * M[1 * 3 + b] - M[1 * 3 + a] * M[0 * 3 + b];
  };
  if (S[1] <= tiny) {
    // rank <= 1: pick any unit vector orthogonal to U[:,0]
    if (S[0] <= 0) { U[0] = 1; U[3] = 0; U[6] = 0; }
    int ax = 0;
    if (fabs(U[3]) < fabs(U[ax * 3])) ax = 1;
    if (fabs(U[6]) < fabs(U[ax * 3])) ax = 2;
    double e[3] = {0, 0, 0};
    e[ax] = 1.0;
    const double dot = U[0] * e[0] + U[3] * e[1] + U[6] * e[2];
    double y[3] = {e[0] - dot * U[0], e[1] - dot * U[3], e[2] - dot * U[6]};
    const double ny = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
    U[1] = y[0] / ny; U[4] = y[1] / ny; U[7] = y[2] / ny;
  }
  if (S[2] <= tiny) cross_col(U, 0, 1, 2);
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
