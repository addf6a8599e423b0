// Device-side per-edge / per-vertex math (FP64, registers only) for the supported g2o types.
// Semantics follow the reference (paths relative to the reference root); the BAL Jacobian is hand-derived
// (the reference obtains it by forward-mode autodiff, examples/bal/bal_example.cpp:254-281).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/g2ocu.h"

namespace g2ocu {

#define G2D __device__ __forceinline__

// ---------- tiny 3x3 helpers, matrices column-major m[r + 3c] ----------
G2D void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) C[r + 3 * c] = A[r] * B[3 * c] + A[r + 3] * B[1 + 3 * c] + A[r + 6] * B[2 + 3 * c];
}
G2D void mat3_mulT_left(const double* A, const double* B, double* C) {   // C = A^T B
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) C[r + 3 * c] = A[3 * r] * B[3 * c] + A[1 + 3 * r] * B[1 + 3 * c] + A[2 + 3 * r] * B[2 + 3 * c];
}
G2D void mat3_vec(const double* A, const double* x, double* y) {
#pragma unroll
  for (int r = 0; r < 3; ++r) y[r] = A[r] * x[0] + A[r + 3] * x[1] + A[r + 6] * x[2];
}
G2D void mat3T_vec(const double* A, const double* x, double* y) {
#pragma unroll
  for (int r = 0; r < 3; ++r) y[r] = A[3 * r] * x[0] + A[1 + 3 * r] * x[1] + A[2 + 3 * r] * x[2];
}
G2D void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
// skew(v), se3_ops.hpp:27-40
G2D void skew3(const double* v, double* m) {
  m[0] = 0; m[4] = 0; m[8] = 0;
  m[0 + 3 * 1] = -v[2]; m[0 + 3 * 2] = v[1]; m[1 + 3 * 2] = -v[0];
  m[1 + 3 * 0] = v[2]; m[2 + 3 * 0] = -v[1]; m[2 + 3 * 1] = v[0];
}

// ---------- quaternion (x,y,z,w), Eigen semantics ----------
G2D void quat_to_R(const double* q, double* R) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[3] = txy - twz;       R[6] = txz + twy;
  R[1] = txy + twz;       R[4] = 1 - (txx + tzz); R[7] = tyz - twx;
  R[2] = txz - twy;       R[5] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
G2D void R_to_quat(const double* R, double* q) {   // Eigen Quaternion(Matrix3)
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (R[5] - R[7]) * t; q[1] = (R[6] - R[2]) * t; q[2] = (R[1] - R[3]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i + 3 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i + 3 * i] - R[j + 3 * j] - R[k + 3 * k] + 1.0);
    double c[3];
    c[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (R[k + 3 * j] - R[j + 3 * k]) * t;
    c[j] = (R[j + 3 * i] + R[i + 3 * j]) * t;
    c[k] = (R[k + 3 * i] + R[i + 3 * k]) * t;
    q[0] = c[0]; q[1] = c[1]; q[2] = c[2];
  }
}
G2D void quat_mul(const double* a, const double* b, double* o) {
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
G2D void quat_rotate(const double* q, const double* v, double* o) {   // Eigen _transformVector
  double uv[3], c2[3];
  cross3(q, v, uv);
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  cross3(q, uv, c2);
  o[0] = v[0] + q[3] * uv[0] + c2[0]; o[1] = v[1] + q[3] * uv[1] + c2[1]; o[2] = v[2] + q[3] * uv[2] + c2[2];
}
G2D void quat_normalize_pos(double* q) {   // SE3Quat::normalizeRotation, se3quat.h:270-275
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// ---------- SE3Quat stored as [t(3), q(4)] (SE3Quat::toVector layout) ----------
G2D void se3q_mul(const double* a, const double* b, double* o) {   // se3quat.h:99-105
  double rt[3]; quat_rotate(a + 3, b, rt);
  o[0] = a[0] + rt[0]; o[1] = a[1] + rt[1]; o[2] = a[2] + rt[2];
  quat_mul(a + 3, b + 3, o + 3);
  quat_normalize_pos(o + 3);
}
G2D void se3q_inv(const double* a, double* o) {                    // se3quat.h:118-123
  o[3] = -a[3]; o[4] = -a[4]; o[5] = -a[5]; o[6] = a[6];
  const double nt[3] = {-a[0], -a[1], -a[2]};
  quat_rotate(o + 3, nt, o);
}
G2D void se3q_log(const double* a, double* res) {                  // se3quat.h:173-209
  double R[9]; quat_to_R(a + 3, R);
  const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
  const double dR[3] = {R[5] - R[7], R[6] - R[2], R[1] - R[3]};
  double omega[3], Om[9], Om2[9], Vinv[9];
  double coef;
  if (d > 0.99999) {
    omega[0] = 0.5 * dR[0]; omega[1] = 0.5 * dR[1]; omega[2] = 0.5 * dR[2];
    coef = 1. / 12.;
  } else {
    const double theta = acos(d);
    const double s = theta / (2 * sqrt(1 - d * d));
    omega[0] = s * dR[0]; omega[1] = s * dR[1]; omega[2] = s * dR[2];
    coef = (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
  }
  skew3(omega, Om); mat3_mul(Om, Om, Om2);
#pragma unroll
  for (int i = 0; i < 9; ++i) Vinv[i] = -0.5 * Om[i] + coef * Om2[i];
  Vinv[0] += 1; Vinv[4] += 1; Vinv[8] += 1;
  mat3_vec(Vinv, a, res + 3);
  res[0] = omega[0]; res[1] = omega[1]; res[2] = omega[2];
}
G2D void se3q_exp(const double* u, double* o) {                    // se3quat.h:218-257
  const double theta = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  double Om[9], Om2[9], R[9], V[9];
  skew3(u, Om); mat3_mul(Om, Om, Om2);
  double a1, a2, b1, b2;
  if (theta < 0.00001) { a1 = 1; a2 = 0.5; b1 = 0.5; b2 = 1. / 6.; }
  else { a1 = sin(theta) / theta; a2 = (1 - cos(theta)) / (theta * theta); b1 = a2; b2 = (theta - sin(theta)) / (theta * theta * theta); }
#pragma unroll
  for (int i = 0; i < 9; ++i) { R[i] = a1 * Om[i] + a2 * Om2[i]; V[i] = b1 * Om[i] + b2 * Om2[i]; }
  R[0] += 1; R[4] += 1; R[8] += 1; V[0] += 1; V[4] += 1; V[8] += 1;
  mat3_vec(V, u + 3, o);
  R_to_quat(R, o + 3);
  quat_normalize_pos(o + 3);
}
G2D void se3q_adj(const double* a, double sgn, double* J /*6x6 col-major*/) {   // se3quat.h:259-268
  double R[9], T[9], TR[9];
  quat_to_R(a + 3, R); skew3(a, T); mat3_mul(T, R, TR);
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      J[r + 6 * c] = sgn * R[r + 3 * c]; J[(r + 3) + 6 * (c + 3)] = sgn * R[r + 3 * c];
      J[(r + 3) + 6 * c] = sgn * TR[r + 3 * c]; J[r + 6 * (c + 3)] = 0;
    }
}

// ---------- Isometry3 stored as [R col-major(9), t(3)] ----------
G2D void iso_mul(const double* A, const double* B, double* O) {
  mat3_mul(A, B, O);
  double rt[3]; mat3_vec(A, B + 9, rt);
  O[9] = rt[0] + A[9]; O[10] = rt[1] + A[10]; O[11] = rt[2] + A[11];
}
G2D void iso_inv(const double* A, double* O) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) O[r + 3 * c] = A[c + 3 * r];
  double rt[3]; mat3_vec(O, A + 9, rt);
  O[9] = -rt[0]; O[10] = -rt[1]; O[11] = -rt[2];
}
// isometry3d_mappings.cpp:78-100 toVectorMQT
G2D void iso_to_mqt(const double* T, double* v) {
  double q[4]; R_to_quat(T, q);
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; }
  v[0] = T[9]; v[1] = T[10]; v[2] = T[11]; v[3] = q[0]; v[4] = q[1]; v[5] = q[2];
}
// isometry3d_mappings.cpp:85-92,110-115 fromVectorMQT
G2D void iso_from_mqt(const double* v, double* T) {
  double w = 1 - (v[3] * v[3] + v[4] * v[4] + v[5] * v[5]);
  if (w < 0) { T[0] = 1; T[1] = 0; T[2] = 0; T[3] = 0; T[4] = 1; T[5] = 0; T[6] = 0; T[7] = 0; T[8] = 1; }
  else { const double q[4] = {v[3], v[4], v[5], sqrt(w)}; quat_to_R(q, T); }
  T[9] = v[0]; T[10] = v[1]; T[11] = v[2];
}
// d q_xyz / d vec(R) (3x9, vec column-major); dquat2mat.cpp:35-85.  Written from the closed form
// q_c = N_c / S4, S4 = 2 sqrt(1 + s0 r00 + s1 r11 + s2 r22), dominant component = S4 / 4.
G2D void dq_dR_dev(const double* R, double* D /*3x9 col-major: D[row + 3*col]*/) {
  const double r00 = R[0], r11 = R[4], r22 = R[8];
  int which; double s0, s1, s2;
  if (r00 + r11 + r22 > 0) { which = 3; s0 = 1; s1 = 1; s2 = 1; }
  else if ((r00 > r11) & (r00 > r22)) { which = 0; s0 = 1; s1 = -1; s2 = -1; }
  else if (r11 > r22) { which = 1; s0 = -1; s1 = 1; s2 = -1; }
  else { which = 2; s0 = -1; s1 = -1; s2 = 1; }
  const double S4 = 2 * sqrt(1.0 + s0 * r00 + s1 * r11 + s2 * r22);
  const double iS = 1.0 / S4, g = 2 * iS * iS * iS;   // -N * g * s_i is the diagonal derivative
  const double sg[3] = {s0, s1, s2};
#pragma unroll
  for (int i = 0; i < 27; ++i) D[i] = 0;
  double qw;
  if (which == 3) {
    qw = 0.25 * S4;
    const double N0 = R[5] - R[7], N1 = R[6] - R[2], N2 = R[1] - R[3];
    D[0 + 3 * 5] = iS; D[0 + 3 * 7] = -iS;   // r21, r12
    D[1 + 3 * 6] = iS; D[1 + 3 * 2] = -iS;   // r02, r20
    D[2 + 3 * 1] = iS; D[2 + 3 * 3] = -iS;   // r10, r01
#pragma unroll
    for (int d = 0; d < 3; ++d) { D[0 + 3 * (4 * d)] -= N0 * g * sg[d]; D[1 + 3 * (4 * d)] -= N1 * g * sg[d]; D[2 + 3 * (4 * d)] -= N2 * g * sg[d]; }
  } else {
    const int i = which, j = (i + 1) % 3, k = (j + 1) % 3;
    qw = (R[k + 3 * j] - R[j + 3 * k]) * iS;
    const double Nj = R[j + 3 * i] + R[i + 3 * j], Nk = R[k + 3 * i] + R[i + 3 * k];
    D[j + 3 * (j + 3 * i)] += iS; D[j + 3 * (i + 3 * j)] += iS;
    D[k + 3 * (k + 3 * i)] += iS; D[k + 3 * (i + 3 * k)] += iS;
    for (int d = 0; d < 3; ++d) {
      D[i + 3 * (4 * d)] += 0.5 * iS * sg[d];
      D[j + 3 * (4 * d)] -= Nj * g * sg[d];
      D[k + 3 * (4 * d)] -= Nk * g * sg[d];
    }
  }
  if (qw <= 0) {
#pragma unroll
    for (int i = 0; i < 27; ++i) D[i] = -D[i];
  }
}

// stuff/misc.h:114-127
G2D double normalize_theta(double theta) {
  const double pi = 3.14159265358979323846;
  if (theta >= -pi && theta < pi) return theta;
  const double multiplier = floor(theta / (2 * pi));
  theta = theta - multiplier * 2 * pi;
  if (theta >= pi) theta -= 2 * pi;
  if (theta < -pi) theta += 2 * pi;
  return theta;
}

// ---------- robust kernels, core/robust_kernel_impl.cpp:65-170 (rho[0], rho[1] only: rho[2] is unused, base_edge.h:117-123) ----------
G2D void robustify_dev(int kind, double delta, double e2, double& rho0, double& rho1) {
  switch (kind) {
    case G2OCU_KERNEL_HUBER: {
      const double dsqr = delta * delta;
      if (e2 <= dsqr) { rho0 = e2; rho1 = 1.; } else { const double sq = sqrt(e2); rho0 = 2 * sq * delta - dsqr; rho1 = delta / sq; }
      break; }
    case G2OCU_KERNEL_PSEUDO_HUBER: { const double dsqr = delta * delta, aux2 = sqrt(e2 / dsqr + 1.0); rho0 = 2 * dsqr * (aux2 - 1); rho1 = 1. / aux2; break; }
    case G2OCU_KERNEL_CAUCHY: { const double dsqr = delta * delta, aux = e2 / dsqr + 1.0; rho0 = dsqr * log(aux); rho1 = 1. / aux; break; }
    case G2OCU_KERNEL_GEMAN_MCCLURE: { const double aux = delta / (delta + e2); rho0 = e2 * aux; rho1 = aux * aux; break; }
    case G2OCU_KERNEL_WELSCH: { const double dsqr = delta * delta, aux2 = exp(-e2 / dsqr); rho0 = dsqr * (1. - aux2); rho1 = aux2; break; }
    case G2OCU_KERNEL_FAIR: { const double sq = sqrt(e2), aux = sq / delta; rho0 = 2. * delta * delta * (aux - log(1. + aux)); rho1 = 1. / (1. + aux); break; }
    case G2OCU_KERNEL_TUKEY: {
      const double e = sqrt(e2), d2 = delta * delta;
      if (e <= delta) { const double a = 1. - e2 / d2; rho0 = d2 * (1. - a * a * a) / 3.; rho1 = a * a; } else { rho0 = d2 / 3.; rho1 = 0; }
      break; }
    case G2OCU_KERNEL_SATURATED: { const double dsqr = delta * delta; if (e2 <= dsqr) { rho0 = e2; rho1 = 1.; } else { rho0 = dsqr; rho1 = 0.; } break; }
    case G2OCU_KERNEL_DCS: { double sc = (2.0 * delta) / (delta + e2); if (sc >= 1.0) sc = 1.0; rho0 = sc * e2 * sc; rho1 = sc * sc; break; }
    default: rho0 = e2; rho1 = 1.;
  }
}

// ---------- vertex ⊞ ----------
template <int VT> struct VertexT;
template <> struct VertexT<G2OCU_VERTEX_SE2> { static constexpr int S = 3, D = 3;
  G2D static void oplus(double* x, const double* u, int*) { x[0] += u[0]; x[1] += u[1]; x[2] = normalize_theta(x[2] + u[2]); } };   // vertex_se2.h:51-58
template <> struct VertexT<G2OCU_VERTEX_POINT_XY> { static constexpr int S = 2, D = 2;
  G2D static void oplus(double* x, const double* u, int*) { x[0] += u[0]; x[1] += u[1]; } };
template <> struct VertexT<G2OCU_VERTEX_SE3> { static constexpr int S = 12, D = 6;
  G2D static void oplus(double* x, const double* u, int* counter) {                    // vertex_se3.h:105-114
    double inc[12], o[12]; iso_from_mqt(u, inc); iso_mul(x, inc, o);
    if (++(*counter) > 1000) {                                                         // isometry3d_mappings.h:81-86
      *counter = 0;
      double E[9], RE[9]; mat3_mulT_left(o, o, E); E[0] -= 1; E[4] -= 1; E[8] -= 1; mat3_mul(o, E, RE);
#pragma unroll
      for (int i = 0; i < 9; ++i) o[i] -= 0.5 * RE[i];
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = o[i];
  } };
template <> struct VertexT<G2OCU_VERTEX_SE3_EXPMAP> { static constexpr int S = 7, D = 6;
  G2D static void oplus(double* x, const double* u, int*) {                            // types_six_dof_expmap.h:98-101
    double ex[7], o[7]; se3q_exp(u, ex); se3q_mul(ex, x, o);
#pragma unroll
    for (int i = 0; i < 7; ++i) x[i] = o[i];
  } };
template <> struct VertexT<G2OCU_VERTEX_POINT_XYZ> { static constexpr int S = 3, D = 3;
  G2D static void oplus(double* x, const double* u, int*) { x[0] += u[0]; x[1] += u[1]; x[2] += u[2]; } };
template <> struct VertexT<G2OCU_VERTEX_POINT_BAL> { static constexpr int S = 3, D = 3;
  G2D static void oplus(double* x, const double* u, int*) { x[0] += u[0]; x[1] += u[1]; x[2] += u[2]; } };
template <> struct VertexT<G2OCU_VERTEX_CAM_BAL> { static constexpr int S = 9, D = 9;
  G2D static void oplus(double* x, const double* u, int*) {
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] += u[i];
  } };

// ---------- edges: error and (error + Jacobians) ----------
// Jacobians column-major E x D:  J[r + E*c].   WANT_J = false skips the Jacobian work.
template <int ET> struct EdgeT;

template <> struct EdgeT<G2OCU_EDGE_SE2> {   // slam2d/edge_se2.h:45-52, edge_se2.cpp:77-103
  static constexpr int E = 3, D0 = 3, D1 = 3, S0 = 3, S1 = 3, M = 3, NP = 0, VT0 = G2OCU_VERTEX_SE2, VT1 = G2OCU_VERTEX_SE2;
  template <bool WANT_J> G2D static void eval(const double* a, const double* b, const double* z, const double*, double* e, double* J0, double* J1) {
    // Zinv = z.inverse(); delta = Zinv * (a.inverse() * b)
    const double ith = normalize_theta(-a[2]); double si, ci; sincos(ith, &si, &ci);
    const double aix = ci * (-a[0]) - si * (-a[1]), aiy = si * (-a[0]) + ci * (-a[1]);
    const double abx = aix + ci * b[0] - si * b[1], aby = aiy + si * b[0] + ci * b[1], abth = normalize_theta(ith + b[2]);
    const double zth = normalize_theta(-z[2]); double sz, cz; sincos(zth, &sz, &cz);
    const double zix = cz * (-z[0]) - sz * (-z[1]), ziy = sz * (-z[0]) + cz * (-z[1]);
    e[0] = zix + cz * abx - sz * aby; e[1] = ziy + sz * abx + cz * aby; e[2] = normalize_theta(zth + abth);
    if (WANT_J) {
      double s, c; sincos(a[2], &s, &c);
      const double dtx = b[0] - a[0], dty = b[1] - a[1];
      const double A[9] = {-c, s, 0, -s, -c, 0, -s * dtx + c * dty, -c * dtx - s * dty, -1};
      const double B[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
      const double Z[9] = {cz, sz, 0, -sz, cz, 0, 0, 0, 1};
      mat3_mul(Z, A, J0); mat3_mul(Z, B, J1);
    }
  }
};

template <> struct EdgeT<G2OCU_EDGE_SE2_POINT_XY> {   // slam2d/edge_se2_pointxy.h:45-50, .cpp:68-93
  static constexpr int E = 2, D0 = 3, D1 = 2, S0 = 3, S1 = 2, M = 2, NP = 0, VT0 = G2OCU_VERTEX_SE2, VT1 = G2OCU_VERTEX_POINT_XY;
  template <bool WANT_J> G2D static void eval(const double* a, const double* l, const double* z, const double*, double* e, double* J0, double* J1) {
    const double ith = normalize_theta(-a[2]); double si, ci; sincos(ith, &si, &ci);
    const double aix = ci * (-a[0]) - si * (-a[1]), aiy = si * (-a[0]) + ci * (-a[1]);
    e[0] = (aix + ci * l[0] - si * l[1]) - z[0]; e[1] = (aiy + si * l[0] + ci * l[1]) - z[1];
    if (WANT_J) {
      double s, c; sincos(a[2], &s, &c);
      J0[0] = -c; J0[2] = -s; J0[4] = c * l[1] - c * a[1] - s * l[0] + s * a[0];
      J0[1] = s;  J0[3] = -c; J0[5] = -s * l[1] + s * a[1] - c * l[0] + c * a[0];
      J1[0] = c; J1[2] = s; J1[1] = -s; J1[3] = c;
    }
  }
};

template <> struct EdgeT<G2OCU_EDGE_SE3> {   // slam3d/edge_se3.cpp:77-105, isometry3d_gradients.h:192-255
  static constexpr int E = 6, D0 = 6, D1 = 6, S0 = 12, S1 = 12, M = 12, NP = 0, VT0 = G2OCU_VERTEX_SE3, VT1 = G2OCU_VERTEX_SE3;
  template <bool WANT_J> G2D static void eval(const double* Xi, const double* Xj, const double* Z, const double*, double* e, double* J0, double* J1) {
    double A[12], Xii[12], B[12], ZiXi[12], Ee[12];
    iso_inv(Z, A); iso_inv(Xi, Xii);
    iso_mul(A, Xii, ZiXi); iso_mul(ZiXi, Xj, Ee);          // error path: (Z^-1 Xi^-1) Xj
    iso_to_mqt(Ee, e);
    if (WANT_J) {
      iso_mul(Xii, Xj, B);                                  // Jacobian path: A * (Xi^-1 Xj)
      double Eg[12]; iso_mul(A, B, Eg);
      double D[27]; dq_dR_dev(Eg, D);
#pragma unroll
      for (int i = 0; i < 36; ++i) { J0[i] = 0; J1[i] = 0; }
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) { J0[r + 6 * c] = -A[r + 3 * c]; J1[r + 6 * c] = Eg[r + 3 * c]; }
      {  // dte/dqi = Ra * skewT(tb)   (skewT with the factor 2, isometry3d_gradients.h:49-54)
        const double x = 2 * B[9], y = 2 * B[10], zz = 2 * B[11];
        const double S[9] = {0, zz, -y, -zz, 0, x, y, -x, 0};   // column-major of rows (0,-z,y),(z,0,-x),(-y,x,0)
        double RS[9]; mat3_mul(A, S, RS);
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int r = 0; r < 3; ++r) J0[r + 6 * (c + 3)] = RS[r + 3 * c];
      }
      // dre/dq = dq_dR * [vec(L Sx); vec(L Sy); vec(L Sz)]
      auto dre = [&](const double* Lm, const double* Sx, const double* Sy, const double* Sz, double* J) {
        double Mx[9], My[9], Mz[9];
        mat3_mul(Lm, Sx, Mx); mat3_mul(Lm, Sy, My); mat3_mul(Lm, Sz, Mz);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          double o0 = 0, o1 = 0, o2 = 0;
#pragma unroll
          for (int k = 0; k < 9; ++k) { const double d = D[r + 3 * k]; o0 += d * Mx[k]; o1 += d * My[k]; o2 += d * Mz[k]; }
          J[(3 + r) + 6 * 3] = o0; J[(3 + r) + 6 * 4] = o1; J[(3 + r) + 6 * 5] = o2;
        }
      };
      {  // skewT(Sx,Sy,Sz,Rb), rows listed then stored column-major
        const double r11 = 2 * B[0], r12 = 2 * B[3], r13 = 2 * B[6], r21 = 2 * B[1], r22 = 2 * B[4], r23 = 2 * B[7], r31 = 2 * B[2], r32 = 2 * B[5], r33 = 2 * B[8];
        const double Sx[9] = {0, r31, -r21, 0, r32, -r22, 0, r33, -r23};
        const double Sy[9] = {-r31, 0, r11, -r32, 0, r12, -r33, 0, r13};
        const double Sz[9] = {r21, -r11, 0, r22, -r12, 0, r23, -r13, 0};
        dre(A, Sx, Sy, Sz, J0);
      }
      {  // skew(Sx,Sy,Sz,I)
        const double Sx[9] = {0, 0, 0, 0, 0, 2, 0, -2, 0};
        const double Sy[9] = {0, 0, -2, 0, 0, 0, 2, 0, 0};
        const double Sz[9] = {0, 2, 0, -2, 0, 0, 0, 0, 0};
        dre(Eg, Sx, Sy, Sz, J1);
      }
    }
  }
};

template <> struct EdgeT<G2OCU_EDGE_SE3_EXPMAP> {   // sba/types_six_dof_expmap.h:117-124, .cpp:278-293
  static constexpr int E = 6, D0 = 6, D1 = 6, S0 = 7, S1 = 7, M = 7, NP = 0, VT0 = G2OCU_VERTEX_SE3_EXPMAP, VT1 = G2OCU_VERTEX_SE3_EXPMAP;
  template <bool WANT_J> G2D static void eval(const double* T0, const double* T1, const double* C, const double*, double* e, double* J0, double* J1) {
    double T1i[7], t[7], err[7];
    se3q_inv(T1, T1i); se3q_mul(T1i, C, t); se3q_mul(t, T0, err);
    se3q_log(err, e);
    if (WANT_J) {
      double Ci[7], T0i[7], u[7];
      se3q_adj(t, 1.0, J0);                       // adj(Tj^-1 * Tij)
      se3q_inv(C, Ci); se3q_inv(T0, T0i); se3q_mul(T0i, Ci, u);
      se3q_adj(u, -1.0, J1);                      // -adj(Ti^-1 * Tij^-1)
    }
  }
};

template <int ET> struct ProjectBase {   // sba/types_six_dof_expmap.cpp:295-331 (f,cx,cy) and :395-455 (fx,fy,cx,cy); v0 = point, v1 = pose
  static constexpr int E = 2, D0 = 3, D1 = 6, S0 = 3, S1 = 7, M = 2, NP = (ET == G2OCU_EDGE_PROJECT_XYZ2UV ? 3 : 4), VT0 = G2OCU_VERTEX_POINT_XYZ, VT1 = G2OCU_VERTEX_SE3_EXPMAP;
  template <bool WANT_J> G2D static void eval(const double* X, const double* T, const double* z, const double* prm, double* e, double* J0, double* J1) {
    const double fx = prm[0], fy = (ET == G2OCU_EDGE_PROJECT_XYZ2UV) ? prm[0] : prm[1];
    const double cx = (ET == G2OCU_EDGE_PROJECT_XYZ2UV) ? prm[1] : prm[2], cy = (ET == G2OCU_EDGE_PROJECT_XYZ2UV) ? prm[2] : prm[3];
    double Pm[3]; quat_rotate(T + 3, X, Pm); Pm[0] += T[0]; Pm[1] += T[1]; Pm[2] += T[2];
    const double x = Pm[0], y = Pm[1], zz = Pm[2];
    e[0] = z[0] - (x / zz * fx + cx); e[1] = z[1] - (y / zz * fy + cy);
    if (WANT_J) {
      const double z_2 = zz * zz;
      double R[9]; quat_to_R(T + 3, R);
      const double t02 = -x / zz * fx, t12 = -y / zz * fy, miz = -1. / zz;
#pragma unroll
      for (int c = 0; c < 3; ++c) { J0[0 + 2 * c] = miz * (fx * R[0 + 3 * c] + t02 * R[2 + 3 * c]); J0[1 + 2 * c] = miz * (fy * R[1 + 3 * c] + t12 * R[2 + 3 * c]); }
      J1[0 + 2 * 0] = x * y / z_2 * fx;      J1[0 + 2 * 1] = -(1 + (x * x / z_2)) * fx; J1[0 + 2 * 2] = y / zz * fx;  J1[0 + 2 * 3] = -1. / zz * fx; J1[0 + 2 * 4] = 0;             J1[0 + 2 * 5] = x / z_2 * fx;
      J1[1 + 2 * 0] = (1 + y * y / z_2) * fy; J1[1 + 2 * 1] = -x * y / z_2 * fy;         J1[1 + 2 * 2] = -x / zz * fy; J1[1 + 2 * 3] = 0;             J1[1 + 2 * 4] = -1. / zz * fy; J1[1 + 2 * 5] = y / z_2 * fy;
    }
  }
};
template <> struct EdgeT<G2OCU_EDGE_PROJECT_XYZ2UV> : ProjectBase<G2OCU_EDGE_PROJECT_XYZ2UV> {};
template <> struct EdgeT<G2OCU_EDGE_SE3_PROJECT_XYZ> : ProjectBase<G2OCU_EDGE_SE3_PROJECT_XYZ> {};

template <> struct EdgeT<G2OCU_EDGE_BAL> {   // examples/bal/bal_example.cpp:192-281; v0 = camera, v1 = point
  static constexpr int E = 2, D0 = 9, D1 = 3, S0 = 9, S1 = 3, M = 2, NP = 0, VT0 = G2OCU_VERTEX_CAM_BAL, VT1 = G2OCU_VERTEX_POINT_BAL;
  template <bool WANT_J> G2D static void eval(const double* cam, const double* X, const double* z, const double*, double* e, double* J0, double* J1) {
    double p[3], R[9], dpw[9];   // R = dp/dX, dpw = dp/d(omega), both column-major
    const double th = sqrt(cam[0] * cam[0] + cam[1] * cam[1] + cam[2] * cam[2]);
    if (th > 0) {
      const double v[3] = {cam[0] / th, cam[1] / th, cam[2] / th};
      double s, c; sincos(th, &s, &c);
      double vxp[3]; cross3(v, X, vxp);
      const double vdp = v[0] * X[0] + v[1] * X[1] + v[2] * X[2], omc = 1.0 - c;
#pragma unroll
      for (int i = 0; i < 3; ++i) p[i] = X[i] * c + vxp[i] * s + v[i] * vdp * omc;
      if (WANT_J) {
        // dp/dX = c I + s [v]x + (1-c) v v^T
        double K[9]; skew3(v, K);
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
          for (int r = 0; r < 3; ++r) R[r + 3 * cc] = s * K[r + 3 * cc] + omc * v[r] * v[cc] + (r == cc ? c : 0.0);
        // dp/dw_j with d(theta)/dw_j = v_j and dv/dw_j = (e_j - v v_j) / theta
        const double ith = 1.0 / th;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double dv[3] = {-v[0] * v[j] * ith, -v[1] * v[j] * ith, -v[2] * v[j] * ith};
          dv[j] += ith;
          double dvxp[3]; cross3(dv, X, dvxp);
          const double dvdp = dv[0] * X[0] + dv[1] * X[1] + dv[2] * X[2];
#pragma unroll
          for (int i = 0; i < 3; ++i)
            dpw[i + 3 * j] = -s * v[j] * X[i] + dvxp[i] * s + vxp[i] * c * v[j] + dv[i] * vdp * omc + v[i] * dvdp * omc + v[i] * vdp * s * v[j];
        }
      }
    } else {   // first-order rotation, bal_example.cpp:218-224
      double aux[3]; cross3(cam, X, aux);
      p[0] = X[0] + aux[0]; p[1] = X[1] + aux[1]; p[2] = X[2] + aux[2];
      if (WANT_J) {
        skew3(cam, R); R[0] += 1; R[4] += 1; R[8] += 1;
        double Xs[9]; skew3(X, Xs);
#pragma unroll
        for (int i = 0; i < 9; ++i) dpw[i] = -Xs[i];
      }
    }
    p[0] += cam[3]; p[1] += cam[4]; p[2] += cam[5];
    const double ipz = 1.0 / p[2];
    const double u0 = -p[0] * ipz, u1 = -p[1] * ipz;
    const double r2 = u0 * u0 + u1 * u1;
    const double f = cam[6], k1 = cam[7], k2 = cam[8];
    const double rp = 1.0 + k1 * r2 + k2 * r2 * r2;
    e[0] = f * rp * u0 - z[0]; e[1] = f * rp * u1 - z[1];
    if (WANT_J) {
      const double g = 2 * k1 + 4 * k2 * r2;
      // de/du (2x2)
      const double a00 = f * (rp + u0 * g * u0), a01 = f * (u0 * g * u1), a10 = a01, a11 = f * (rp + u1 * g * u1);
      // du/dp (2x3): [-1/pz 0 p0/pz^2; 0 -1/pz p1/pz^2]
      const double d02 = p[0] * ipz * ipz, d12 = p[1] * ipz * ipz;
      double dep[6];   // de/dp 2x3 col-major
      dep[0] = -a00 * ipz; dep[1] = -a10 * ipz;
      dep[2] = -a01 * ipz; dep[3] = -a11 * ipz;
      dep[4] = a00 * d02 + a01 * d12; dep[5] = a10 * d02 + a11 * d12;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        J0[0 + 2 * c] = dep[0] * dpw[0 + 3 * c] + dep[2] * dpw[1 + 3 * c] + dep[4] * dpw[2 + 3 * c];
        J0[1 + 2 * c] = dep[1] * dpw[0 + 3 * c] + dep[3] * dpw[1 + 3 * c] + dep[5] * dpw[2 + 3 * c];
        J0[0 + 2 * (3 + c)] = dep[2 * c]; J0[1 + 2 * (3 + c)] = dep[2 * c + 1];
        J1[0 + 2 * c] = dep[0] * R[0 + 3 * c] + dep[2] * R[1 + 3 * c] + dep[4] * R[2 + 3 * c];
        J1[1 + 2 * c] = dep[1] * R[0 + 3 * c] + dep[3] * R[1 + 3 * c] + dep[5] * R[2 + 3 * c];
      }
      J0[0 + 2 * 6] = rp * u0;          J0[1 + 2 * 6] = rp * u1;
      J0[0 + 2 * 7] = f * r2 * u0;      J0[1 + 2 * 7] = f * r2 * u1;
      J0[0 + 2 * 8] = f * r2 * r2 * u0; J0[1 + 2 * 8] = f * r2 * r2 * u1;
    }
  }
};

}  // namespace g2ocu
