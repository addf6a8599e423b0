// TEST INFRASTRUCTURE — CPU oracle for the g2o LM/BlockSolver hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/.  The product (g2o_b200/) never links or calls it.
//
// orc_math.hpp: Eigen-free restatement of the small fixed-size algebra the reference
// gets from Eigen3 (header-only dependency, not vendored, version unpinned — see
// cmake_modules/FindEigen3.cmake:18-30).  Every routine names the Eigen call it stands for
// and the reference call site that needs it.
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>

namespace orc {

// ---- column-major dense helpers on raw pointers (Eigen::ColMajor is the reference default) ----
// C(r x c) = A(r x k) * B(k x c)
inline void mm(const double* A, const double* B, double* C, int r, int k, int c) {
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) {
      double s = 0;
      for (int t = 0; t < k; ++t) s += A[i + t * r] * B[t + j * k];
      C[i + j * r] = s;
    }
}
// C(r x c) += A^T(r x k) * B(k x c) where A is stored k x r
inline void mtm_add(const double* A, const double* B, double* C, int r, int k, int c) {
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) {
      double s = 0;
      for (int t = 0; t < k; ++t) s += A[t + i * k] * B[t + j * k];
      C[i + j * r] += s;
    }
}
// y(r) += A(r x c) x(c)
inline void mv_add(const double* A, const double* x, double* y, int r, int c) {
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) y[i] += A[i + j * r] * x[j];
}
// y(c) += A^T x, A is r x c
inline void mtv_add(const double* A, const double* x, double* y, int r, int c) {
  for (int j = 0; j < c; ++j) {
    double s = 0;
    for (int i = 0; i < r; ++i) s += A[i + j * r] * x[i];
    y[j] += s;
  }
}

// ---- 3-vector / 3x3 (col-major) ----
struct V3 { double v[3]; double& operator[](int i){return v[i];} double operator[](int i) const {return v[i];} };
inline V3 cross(const V3& a, const V3& b) { return {{a[1]*b[2]-a[2]*b[1], a[2]*b[0]-a[0]*b[2], a[0]*b[1]-a[1]*b[0]}}; }
inline double dot(const V3& a, const V3& b) { return a[0]*b[0]+a[1]*b[1]+a[2]*b[2]; }
inline V3 operator+(const V3& a, const V3& b) { return {{a[0]+b[0],a[1]+b[1],a[2]+b[2]}}; }
inline V3 operator-(const V3& a, const V3& b) { return {{a[0]-b[0],a[1]-b[1],a[2]-b[2]}}; }
inline V3 operator*(double s, const V3& a) { return {{s*a[0],s*a[1],s*a[2]}}; }

struct M3 {  // column-major like Eigen::Matrix3d
  double m[9];
  double& operator()(int r, int c) { return m[r + 3*c]; }
  double operator()(int r, int c) const { return m[r + 3*c]; }
  static M3 identity() { M3 I; std::memset(I.m, 0, sizeof(I.m)); I(0,0)=I(1,1)=I(2,2)=1; return I; }
  static M3 zero() { M3 Z; std::memset(Z.m, 0, sizeof(Z.m)); return Z; }
};
inline M3 operator*(const M3& A, const M3& B) { M3 C; mm(A.m, B.m, C.m, 3, 3, 3); return C; }
inline V3 operator*(const M3& A, const V3& x) { V3 y{{0,0,0}}; mv_add(A.m, x.v, y.v, 3, 3); return y; }
inline M3 operator+(const M3& A, const M3& B) { M3 C; for (int i=0;i<9;++i) C.m[i]=A.m[i]+B.m[i]; return C; }
inline M3 operator-(const M3& A, const M3& B) { M3 C; for (int i=0;i<9;++i) C.m[i]=A.m[i]-B.m[i]; return C; }
inline M3 operator*(double s, const M3& A) { M3 C; for (int i=0;i<9;++i) C.m[i]=s*A.m[i]; return C; }
inline M3 transpose(const M3& A) { M3 T; for (int r=0;r<3;++r) for (int c=0;c<3;++c) T(r,c)=A(c,r); return T; }

// g2o/types/slam3d/se3_ops.hpp:27-40  skew(v)
inline M3 skew(const V3& v) {
  M3 m = M3::zero();
  m(0,1) = -v[2]; m(0,2) = v[1]; m(1,2) = -v[0];
  m(1,0) = v[2];  m(2,0) = -v[1]; m(2,1) = v[0];
  return m;
}
// se3_ops.hpp:42-49  deltaR(R)
inline V3 deltaR(const M3& R) { return {{R(2,1)-R(1,2), R(0,2)-R(2,0), R(1,0)-R(0,1)}}; }

// ---- Quaternion with Eigen::Quaterniond semantics (coeff order x,y,z,w) ----
struct Quat {
  double x, y, z, w;
  static Quat identity() { return {0,0,0,1}; }
  double squaredNorm() const { return x*x+y*y+z*z+w*w; }
  double norm() const { return std::sqrt(squaredNorm()); }
  void normalize() { double n = norm(); x/=n; y/=n; z/=n; w/=n; }   // Eigen: coeffs() /= norm()
  Quat conjugate() const { return {-x,-y,-z,w}; }
};
// Eigen quaternion product (QuaternionBase::operator*)
inline Quat operator*(const Quat& a, const Quat& b) {
  return { a.w*b.x + a.x*b.w + a.y*b.z - a.z*b.y,
           a.w*b.y + a.y*b.w + a.z*b.x - a.x*b.z,
           a.w*b.z + a.z*b.w + a.x*b.y - a.y*b.x,
           a.w*b.w - a.x*b.x - a.y*b.y - a.z*b.z };
}
// Eigen QuaternionBase::_transformVector: uv = 2 (q.vec x v); v + w uv + q.vec x uv
inline V3 rotate(const Quat& q, const V3& v) {
  V3 qv{{q.x,q.y,q.z}};
  V3 uv = cross(qv, v);
  uv = uv + uv;
  return v + q.w * uv + cross(qv, uv);
}
// Eigen QuaternionBase::toRotationMatrix
inline M3 toRotationMatrix(const Quat& q) {
  const double tx=2*q.x, ty=2*q.y, tz=2*q.z;
  const double twx=tx*q.w, twy=ty*q.w, twz=tz*q.w;
  const double txx=tx*q.x, txy=ty*q.x, txz=tz*q.x;
  const double tyy=ty*q.y, tyz=tz*q.y, tzz=tz*q.z;
  M3 R;
  R(0,0)=1-(tyy+tzz); R(0,1)=txy-twz;     R(0,2)=txz+twy;
  R(1,0)=txy+twz;     R(1,1)=1-(txx+tzz); R(1,2)=tyz-twx;
  R(2,0)=txz-twy;     R(2,1)=tyz+twx;     R(2,2)=1-(txx+tyy);
  return R;
}
// Eigen Quaternion(Matrix3) — the branch logic the reference's own unit test restates at
// unit_test/slam3d/jacobians_slam3d.cpp:141-187 (there with the sign folded in).
inline Quat fromRotationMatrix(const M3& R) {
  Quat q;
  double t = R(0,0)+R(1,1)+R(2,2);
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5*t;
    t = 0.5/t;
    q.x = (R(2,1)-R(1,2))*t; q.y = (R(0,2)-R(2,0))*t; q.z = (R(1,0)-R(0,1))*t;
  } else {
    int i = 0;
    if (R(1,1) > R(0,0)) i = 1;
    if (R(2,2) > R(i,i)) i = 2;
    int j = (i+1)%3, k = (j+1)%3;
    t = std::sqrt(R(i,i)-R(j,j)-R(k,k)+1.0);
    double c[3];
    c[i] = 0.5*t;
    t = 0.5/t;
    q.w = (R(k,j)-R(j,k))*t;
    c[j] = (R(j,i)+R(i,j))*t;
    c[k] = (R(k,i)+R(i,k))*t;
    q.x=c[0]; q.y=c[1]; q.z=c[2];
  }
  return q;
}

// ---- inverses: Eigen Matrix::inverse() as used at block_solver.hpp:350 and linear_solver_pcg.hpp:94 ----
// 2x2 / 3x3: closed-form cofactor (Eigen compute_inverse<2>,<3>);  larger: PartialPivLU.
inline void inverse2(const double* A, double* X) {
  double det = A[0]*A[3]-A[2]*A[1];
  double id = 1.0/det;
  X[0]= A[3]*id; X[1]=-A[1]*id; X[2]=-A[2]*id; X[3]= A[0]*id;
}
inline void inverse3(const double* A, double* X) {
  auto a=[&](int r,int c){return A[r+3*c];};
  double c00 = a(1,1)*a(2,2)-a(1,2)*a(2,1);
  double c10 = a(1,2)*a(2,0)-a(1,0)*a(2,2);   // cofactor (1,0)
  double c20 = a(1,0)*a(2,1)-a(1,1)*a(2,0);
  double det = c00*a(0,0)+c10*a(0,1)+c20*a(0,2);
  double id = 1.0/det;
  X[0+3*0]=c00*id; X[0+3*1]=(a(0,2)*a(2,1)-a(0,1)*a(2,2))*id; X[0+3*2]=(a(0,1)*a(1,2)-a(0,2)*a(1,1))*id;
  X[1+3*0]=c10*id; X[1+3*1]=(a(0,0)*a(2,2)-a(0,2)*a(2,0))*id; X[1+3*2]=(a(0,2)*a(1,0)-a(0,0)*a(1,2))*id;
  X[2+3*0]=c20*id; X[2+3*1]=(a(0,1)*a(2,0)-a(0,0)*a(2,1))*id; X[2+3*2]=(a(0,0)*a(1,1)-a(0,1)*a(1,0))*id;
}
// general n x n inverse through LU with partial pivoting (Eigen PartialPivLU::inverse)
inline bool inverseLU(const double* A, double* X, int n) {
  double lu[16*16]; int piv[16];
  if (n > 16) return false;
  std::memcpy(lu, A, sizeof(double)*n*n);
  for (int i=0;i<n;++i) piv[i]=i;
  for (int k=0;k<n;++k) {
    int p=k; double best=std::fabs(lu[k+k*n]);
    for (int i=k+1;i<n;++i) { double v=std::fabs(lu[i+k*n]); if (v>best){best=v;p=i;} }
    if (best==0) return false;
    if (p!=k) { for (int j=0;j<n;++j) std::swap(lu[k+j*n], lu[p+j*n]); std::swap(piv[k],piv[p]); }
    for (int i=k+1;i<n;++i) {
      lu[i+k*n] /= lu[k+k*n];
      double l = lu[i+k*n];
      for (int j=k+1;j<n;++j) lu[i+j*n] -= l*lu[k+j*n];
    }
  }
  for (int c=0;c<n;++c) {
    double y[16];
    for (int i=0;i<n;++i) y[i] = (piv[i]==c) ? 1.0 : 0.0;
    for (int i=0;i<n;++i) for (int j=0;j<i;++j) y[i] -= lu[i+j*n]*y[j];
    for (int i=n-1;i>=0;--i) { for (int j=i+1;j<n;++j) y[i] -= lu[i+j*n]*y[j]; y[i] /= lu[i+i*n]; }
    for (int i=0;i<n;++i) X[i+c*n]=y[i];
  }
  return true;
}
inline bool inverseN(const double* A, double* X, int n) {
  if (n==1) { X[0]=1.0/A[0]; return true; }
  if (n==2) { inverse2(A,X); return true; }
  if (n==3) { inverse3(A,X); return true; }
  return inverseLU(A,X,n);
}

// g2o/stuff/misc.h:114-127
inline double normalize_theta(double theta) {
  const double pi = 3.14159265358979323846;
  if (theta >= -pi && theta < pi) return theta;
  double multiplier = std::floor(theta / (2*pi));
  theta = theta - multiplier*2*pi;
  if (theta >= pi) theta -= 2*pi;
  if (theta < -pi) theta += 2*pi;
  return theta;
}

// ---- forward-mode dual numbers: stands for EXTERNAL/ceres/jet.h (used by bal_example.cpp:254-281) ----
template <int N>
struct Jet {
  double a; double v[N];
  Jet() : a(0) { for (int i=0;i<N;++i) v[i]=0; }
  explicit Jet(double s) : a(s) { for (int i=0;i<N;++i) v[i]=0; }
  Jet(double s, int k) : a(s) { for (int i=0;i<N;++i) v[i]=0; v[k]=1; }
};
template<int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g){ Jet<N> h; h.a=f.a+g.a; for(int i=0;i<N;++i)h.v[i]=f.v[i]+g.v[i]; return h; }
template<int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g){ Jet<N> h; h.a=f.a-g.a; for(int i=0;i<N;++i)h.v[i]=f.v[i]-g.v[i]; return h; }
template<int N> inline Jet<N> operator-(const Jet<N>& f){ Jet<N> h; h.a=-f.a; for(int i=0;i<N;++i)h.v[i]=-f.v[i]; return h; }
template<int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g){ Jet<N> h; h.a=f.a*g.a; for(int i=0;i<N;++i)h.v[i]=f.a*g.v[i]+f.v[i]*g.a; return h; }
template<int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g){
  Jet<N> h; const double gi=1.0/g.a; const double fg=f.a*gi; h.a=fg; for(int i=0;i<N;++i)h.v[i]=(f.v[i]-fg*g.v[i])*gi; return h; }
template<int N> inline Jet<N> sqrt(const Jet<N>& f){ Jet<N> h; h.a=std::sqrt(f.a); const double t=1.0/(2.0*h.a); for(int i=0;i<N;++i)h.v[i]=f.v[i]*t; return h; }
template<int N> inline Jet<N> cos(const Jet<N>& f){ Jet<N> h; h.a=std::cos(f.a); const double s=-std::sin(f.a); for(int i=0;i<N;++i)h.v[i]=s*f.v[i]; return h; }
template<int N> inline Jet<N> sin(const Jet<N>& f){ Jet<N> h; h.a=std::sin(f.a); const double c=std::cos(f.a); for(int i=0;i<N;++i)h.v[i]=c*f.v[i]; return h; }
template<int N> inline bool operator>(const Jet<N>& f, const Jet<N>& g){ return f.a>g.a; }
inline double sqrt(double x) { return std::sqrt(x); }
inline double cos(double x) { return std::cos(x); }
inline double sin(double x) { return std::sin(x); }

}  // namespace orc
