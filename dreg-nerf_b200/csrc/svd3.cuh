// One-sided Jacobi SVD of a 3x3 matrix in double (shared by the Procrustes forward and backward kernels).
#pragma once
#include <math.h>

namespace drb {

// A (3x3, row major) = U diag(S) V^T, S descending.
static __host__ __device__ __noinline__ void svd3(const double A_in[9], double U[9], double S[3], double V[9]) {
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

}  // namespace drb
