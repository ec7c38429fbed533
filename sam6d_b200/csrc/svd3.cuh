// svd3.cuh -- rotation of the weighted Procrustes problem from the 3x3 cross-covariance H.
//
// Reference: PEM/utils/model_utils.py:352-358
//     U, _, V = svd(H);  R = V diag(1, 1, sign(det(V U^T))) U^T
// With V = [v1 v2 v3], U = [u1 u2 u3] the determinant correction makes
//     R = v1 u1^T + v2 u2^T + (v1 x v2)(u1 x u2)^T
// independent of the handedness the SVD routine happened to return, so only the two leading
// singular pairs are needed.  They come from a cyclic Jacobi eigen-decomposition of H^T H in
// double precision (fixed sweep count: no data-dependent control flow across a warp).
#pragma once
#include "common.cuh"

__device__ __forceinline__ void jacobi_rot(double S[3][3], double V[3][3], int p, int q) {
  double apq = S[p][q];
  if (fabs(apq) < 1e-300) return;
  double app = S[p][p], aqq = S[q][q];
  double tau = (aqq - app) / (2.0 * apq);
  double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
  double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // S <- S J
    double skp = S[k][p], skq = S[k][q];
    S[k][p] = c * skp - s * skq;
    S[k][q] = s * skp + c * skq;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // S <- J^T S
    double spk = S[p][k], sqk = S[q][k];
    S[p][k] = c * spk - s * sqk;
    S[q][k] = s * spk + c * sqk;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double vkp = V[k][p], vkq = V[k][q];
    V[k][p] = c * vkp - s * vkq;
    V[k][q] = s * vkp + c * vkq;
  }
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double norm3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// any unit vector orthogonal to unit vector a
__device__ __forceinline__ void any_orth(const double a[3], double o[3]) {
  double ax = fabs(a[0]), ay = fabs(a[1]), az = fabs(a[2]);
  double e[3] = {0, 0, 0};
  if (ax <= ay && ax <= az) e[0] = 1; else if (ay <= az) e[1] = 1; else e[2] = 1;
  cross3(a, e, o);
  double n = norm3(o);
  o[0] /= n; o[1] /= n; o[2] /= n;
}

// Rank-1 cross-covariance (a hypothesis whose three sampled correspondences repeat a point: its centred points are
// collinear).  Only the leading singular pair (u1 on the source side, v1 = direction of H^T u1 on the reference side) is
// defined by the data; the reference's rotation about that axis is LAPACK / cuSOLVER rounding noise.  The single
// documented deviation from model_utils.py:352-358: complete with the LEAST rotation that takes u1 to v1,
//     R = c I + [w]x + w w^T / (1 + c),   w = u1 x v1,  c = u1 . v1
// (invariant under the joint sign flip of the pair; for c -> -1 the half turn about a fixed axis orthogonal to u1).
// oracle/pem_oracle.py: coarse_Rt(completion="deterministic") restates exactly this rule.
__device__ inline void rank1_rotation(const double u1[3], const double v1[3], double R[3][3]) {
  const double c = u1[0] * v1[0] + u1[1] * v1[1] + u1[2] * v1[2];
  if (1.0 + c < 1e-9) {
    double a[3];
    any_orth(u1, a);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) R[i][j] = 2.0 * a[i] * a[j] - ((i == j) ? 1.0 : 0.0);
    return;
  }
  double w[3];
  cross3(u1, v1, w);
  const double k = 1.0 / (1.0 + c);
  R[0][0] = c + w[0] * w[0] * k;     R[0][1] = -w[2] + w[0] * w[1] * k; R[0][2] = w[1] + w[0] * w[2] * k;
  R[1][0] = w[2] + w[1] * w[0] * k;  R[1][1] = c + w[1] * w[1] * k;     R[1][2] = -w[0] + w[1] * w[2] * k;
  R[2][0] = -w[1] + w[2] * w[0] * k; R[2][1] = w[0] + w[2] * w[1] * k;  R[2][2] = c + w[2] * w[2] * k;
}

// H (row-major 3x3, double) -> R (row-major 3x3) as the reference's V diag(1,1,d) U^T.
// rank1: the caller knows (from repeated sample indices) that H has a single singular value above rounding noise.
__device__ inline void procrustes_rotation(const double H[3][3], double R[3][3], bool rank1 = false) {
  double S[3][3], V[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      S[i][j] = H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j];  // H^T H
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 8; ++sweep) {
    jacobi_rot(S, V, 0, 1);
    jacobi_rot(S, V, 0, 2);
    jacobi_rot(S, V, 1, 2);
  }
  // order the eigenvalues: i0 >= i1 >= i2
  int i0 = 0, i1 = 1, i2 = 2;
  double l0 = S[0][0], l1 = S[1][1], l2 = S[2][2];
  if (l1 > l0) { double t = l0; l0 = l1; l1 = t; int ti = i0; i0 = i1; i1 = ti; }
  if (l2 > l0) { double t = l0; l0 = l2; l2 = t; int ti = i0; i0 = i2; i2 = ti; }
  if (l2 > l1) { double t = l1; l1 = l2; l2 = t; int ti = i1; i1 = i2; i2 = ti; }
  double v1[3] = {V[0][i0], V[1][i0], V[2][i0]};
  double v2[3] = {V[0][i1], V[1][i1], V[2][i1]};
  double u1[3], u2[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    u1[r] = H[r][0] * v1[0] + H[r][1] * v1[1] + H[r][2] * v1[2];
    u2[r] = H[r][0] * v2[0] + H[r][1] * v2[1] + H[r][2] * v2[2];
  }
  double n1 = norm3(u1);
  if (!(n1 > 0.0)) {  // H == 0: the reference's SVD returns U = V = I, so R = I
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) R[i][j] = (i == j) ? 1.0 : 0.0;
    return;
  }
  u1[0] /= n1; u1[1] /= n1; u1[2] /= n1;
  if (rank1) {  // R maps the source direction u1 onto the reference direction v1
    rank1_rotation(u1, v1, R);
    return;
  }
  // Gram-Schmidt u2 against u1 (exactly orthogonal in exact arithmetic)
  double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
  u2[0] -= d12 * u1[0]; u2[1] -= d12 * u1[1]; u2[2] -= d12 * u1[2];
  double n2 = norm3(u2);
  if (n2 > 1e-12 * n1) {
    u2[0] /= n2; u2[1] /= n2; u2[2] /= n2;
  } else {  // rank-1 H: the completion is arbitrary in any SVD; pick a deterministic one
    any_orth(u1, u2);
  }
  double v3[3], u3[3];
  cross3(v1, v2, v3);
  cross3(u1, u2, u3);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[i][j] = v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j];
}
