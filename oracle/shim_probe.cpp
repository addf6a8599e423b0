// TEST INFRASTRUCTURE.  The compiled reference under oracle/_ref is built against oracle/eigen_shim, a stand-in for the absent Eigen3 that
// restates the few Eigen ALGORITHMS the hot path calls (fixed and dynamic inverse(), llt().solve(), determinant(), Quaternion <-> rotation
// matrix, quaternion product and vector rotation, AngleAxis, Rotation2D, Isometry3 product / inverse).  The oracle restates the same
// algorithms, so "oracle == compiled reference" cannot see a mistake the two restatements share.  This file exports those stand-in routines
// one by one; tests/test_reference_leaves.py compares each with an INDEPENDENT implementation (LAPACK through numpy, scipy's Rotation) on
// random inputs.  It needs no reference tree (the stand-in has no dependency) and is built next to liboracle.so.
#include <Eigen/Core>
#include <Eigen/Geometry>

#include <cstring>

namespace {
template <int N> void inverseFixed(const double* A, double* out) {
  Eigen::Matrix<double, N, N> M; std::memcpy(M.data(), A, sizeof(double) * N * N);
  const Eigen::Matrix<double, N, N> X = M.inverse(); std::memcpy(out, X.data(), sizeof(double) * N * N);
}
template <int N> double detFixed(const double* A) { Eigen::Matrix<double, N, N> M; std::memcpy(M.data(), A, sizeof(double) * N * N); return M.determinant(); }
template <int N> int lltSolveFixed(const double* A, const double* b, double* x) {
  Eigen::Matrix<double, N, N> M; std::memcpy(M.data(), A, sizeof(double) * N * N);
  Eigen::Matrix<double, N, 1> rhs; std::memcpy(rhs.data(), b, sizeof(double) * N);
  auto f = M.llt();
  const Eigen::Matrix<double, N, 1> sol = f.solve(rhs); std::memcpy(x, sol.data(), sizeof(double) * N);
  return f.info() == Eigen::Success ? 1 : 0;
}
Eigen::Isometry3d isoFrom(const double* v12) {      // column-major R (9), then t (3)
  Eigen::Isometry3d T; Eigen::Matrix3d R; std::memcpy(R.data(), v12, sizeof(double) * 9);
  T = R; T.translation() = Eigen::Vector3d(v12[9], v12[10], v12[11]);
  return T;
}
void isoTo(const Eigen::Isometry3d& T, double* v12) {
  const Eigen::Matrix3d R = T.rotation(); std::memcpy(v12, R.data(), sizeof(double) * 9);
  const Eigen::Vector3d t = T.translation(); for (int i = 0; i < 3; ++i) v12[9 + i] = t[i];
}
}  // namespace

extern "C" {
// matrices column-major
int shim_inverse(int n, const double* A, double* out) {
  switch (n) {
    case 2: inverseFixed<2>(A, out); return 1;
    case 3: inverseFixed<3>(A, out); return 1;
    case 6: inverseFixed<6>(A, out); return 1;
    case 7: inverseFixed<7>(A, out); return 1;
    case 9: inverseFixed<9>(A, out); return 1;
    default: return 0;
  }
}
int shim_inverse_dynamic(int n, const double* A, double* out) {
  Eigen::MatrixXd M(n, n); std::memcpy(M.data(), A, sizeof(double) * (size_t)n * n);
  const Eigen::MatrixXd X = M.inverse(); std::memcpy(out, X.data(), sizeof(double) * (size_t)n * n);
  return 1;
}
double shim_determinant(int n, const double* A) {
  switch (n) { case 2: return detFixed<2>(A); case 3: return detFixed<3>(A); case 6: return detFixed<6>(A); default: return 0.0; }
}
int shim_llt_solve(int n, const double* A, const double* b, double* x) {
  switch (n) { case 2: return lltSolveFixed<2>(A, b, x); case 3: return lltSolveFixed<3>(A, b, x); case 6: return lltSolveFixed<6>(A, b, x); default: return -1; }
}
// quaternions as (x, y, z, w)
void shim_quat_from_R(const double* R9, double* q4) {
  Eigen::Matrix3d R; std::memcpy(R.data(), R9, sizeof(double) * 9);
  const Eigen::Quaterniond q(R); q4[0] = q.x(); q4[1] = q.y(); q4[2] = q.z(); q4[3] = q.w();
}
void shim_R_from_quat(const double* q4, double* R9) {
  const Eigen::Quaterniond q(q4[3], q4[0], q4[1], q4[2]); const Eigen::Matrix3d R = q.toRotationMatrix(); std::memcpy(R9, R.data(), sizeof(double) * 9);
}
void shim_quat_mul(const double* a4, const double* b4, double* out4) {
  const Eigen::Quaterniond a(a4[3], a4[0], a4[1], a4[2]), b(b4[3], b4[0], b4[1], b4[2]); const Eigen::Quaterniond c = a * b;
  out4[0] = c.x(); out4[1] = c.y(); out4[2] = c.z(); out4[3] = c.w();
}
void shim_quat_rotate(const double* q4, const double* v3, double* out3) {
  const Eigen::Quaterniond q(q4[3], q4[0], q4[1], q4[2]); const Eigen::Vector3d v(v3[0], v3[1], v3[2]); const Eigen::Vector3d r = q * v;
  for (int i = 0; i < 3; ++i) out3[i] = r[i];
}
void shim_angle_axis_R(double angle, const double* axis3, double* R9) {
  const Eigen::AngleAxisd aa(angle, Eigen::Vector3d(axis3[0], axis3[1], axis3[2])); const Eigen::Matrix3d R = aa.toRotationMatrix(); std::memcpy(R9, R.data(), sizeof(double) * 9);
}
void shim_rotation2d(double angle, const double* v2, double* out2, double* backAngle) {
  const Eigen::Rotation2D<double> r(angle); const Eigen::Vector2d v = r * Eigen::Vector2d(v2[0], v2[1]); out2[0] = v[0]; out2[1] = v[1];
  Eigen::Rotation2D<double> b(0); b.fromRotationMatrix(r.toRotationMatrix()); *backAngle = b.angle();
}
// out = A^-1 * B for two isometries given as (R column-major, t)
void shim_iso_inverse_times(const double* A12, const double* B12, double* out12) { isoTo(isoFrom(A12).inverse() * isoFrom(B12), out12); }
void shim_iso_apply(const double* A12, const double* v3, double* out3) {
  const Eigen::Vector3d r = isoFrom(A12) * Eigen::Vector3d(v3[0], v3[1], v3[2]); for (int i = 0; i < 3; ++i) out3[i] = r[i];
}
}  // extern "C"
