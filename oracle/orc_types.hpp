// TEST INFRASTRUCTURE — CPU oracle (see orc_math.hpp header note).
//
// orc_types.hpp: restatement of the reference's vertex ⊞ operators, edge error functions,
// edge Jacobians and robust kernels for the edge types on the north-star path.
// Each function cites the reference lines it follows (paths relative to /root/reference).
#pragma once
#include "orc_math.hpp"
#include <limits>

namespace orc {

// Numeric values are shared with include/g2ocu.h by convention (documented there too).
enum VertexType { VT_SE2 = 1, VT_POINT_XY = 2, VT_SE3 = 3, VT_SE3_EXPMAP = 4, VT_POINT_XYZ = 5, VT_CAM_BAL = 6, VT_POINT_BAL = 7 };
enum EdgeType { ET_SE2 = 1, ET_SE2_POINT_XY = 2, ET_SE3 = 3, ET_SE3_EXPMAP = 4, ET_PROJECT_XYZ2UV = 5, ET_SE3_PROJECT_XYZ = 6, ET_BAL = 7 };
enum KernelType { RK_NONE = 0, RK_HUBER = 1, RK_PSEUDO_HUBER = 2, RK_CAUCHY = 3, RK_GEMAN_MCCLURE = 4, RK_WELSCH = 5, RK_FAIR = 6, RK_TUKEY = 7, RK_SATURATED = 8, RK_DCS = 9 };

inline int vertexEstimateDim(int t) { static const int d[] = {0, 3, 2, 12, 7, 3, 9, 3}; return (t >= 1 && t <= 7) ? d[t] : -1; }
inline int vertexDim(int t)         { static const int d[] = {0, 3, 2, 6, 6, 3, 9, 3}; return (t >= 1 && t <= 7) ? d[t] : -1; }
inline int edgeDim(int t)           { static const int d[] = {0, 3, 2, 6, 6, 2, 2, 2}; return (t >= 1 && t <= 7) ? d[t] : -1; }
inline int edgeMeasDim(int t)       { static const int d[] = {0, 3, 2, 12, 7, 2, 2, 2}; return (t >= 1 && t <= 7) ? d[t] : -1; }
inline int edgeParamDim(int t)      { static const int d[] = {0, 0, 0, 0, 0, 3, 4, 0}; return (t >= 1 && t <= 7) ? d[t] : -1; }
inline int edgeVertexType(int t, int k) {
  static const int v[8][2] = {{0,0},{VT_SE2,VT_SE2},{VT_SE2,VT_POINT_XY},{VT_SE3,VT_SE3},{VT_SE3_EXPMAP,VT_SE3_EXPMAP},
                              {VT_POINT_XYZ,VT_SE3_EXPMAP},{VT_POINT_XYZ,VT_SE3_EXPMAP},{VT_CAM_BAL,VT_POINT_BAL}};
  return (t >= 1 && t <= 7) ? v[t][k] : -1;
}

// ---------------- robust kernels: g2o/core/robust_kernel_impl.cpp:65-170 ----------------
inline void robustify(int kind, double delta, double e2, double rho[3]) {
  switch (kind) {
    case RK_HUBER: {                                   // :65-78
      double dsqr = delta * delta;
      if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.; rho[2] = 0.; }
      else { double sqrte = std::sqrt(e2); rho[0] = 2*sqrte*delta - dsqr; rho[1] = delta / sqrte; rho[2] = -0.5 * rho[1] / e2; }
      break; }
    case RK_PSEUDO_HUBER: {                            // :80-89
      double dsqr = delta*delta, dsqrReci = 1./dsqr, aux1 = dsqrReci*e2 + 1.0, aux2 = std::sqrt(aux1);
      rho[0] = 2*dsqr*(aux2-1); rho[1] = 1./aux2; rho[2] = -0.5*dsqrReci*rho[1]/aux1; break; }
    case RK_CAUCHY: {                                  // :91-99
      double dsqr = delta*delta, dsqrReci = 1./dsqr, aux = dsqrReci*e2 + 1.0;
      rho[0] = dsqr*std::log(aux); rho[1] = 1./aux; rho[2] = -dsqrReci*std::pow(rho[1], 2); break; }
    case RK_GEMAN_MCCLURE: {                           // :101-107
      const double aux = delta/(delta+e2);
      rho[0] = e2*aux; rho[1] = aux*aux; rho[2] = -2.*rho[1]*aux; break; }
    case RK_WELSCH: {                                  // :109-117
      const double dsqr = delta*delta, aux = e2/dsqr, aux2 = std::exp(-aux);
      rho[0] = dsqr*(1.-aux2); rho[1] = aux2; rho[2] = -aux2/dsqr; break; }
    case RK_FAIR: {                                    // :119-126
      const double sqrte = std::sqrt(e2), aux = sqrte/delta;
      rho[0] = 2.*delta*delta*(aux-std::log(1.+aux)); rho[1] = 1./(1.+aux); rho[2] = -0.5/(sqrte*(1.+aux)); break; }
    case RK_TUKEY: {                                   // :128-143
      const double e = std::sqrt(e2), delta2 = delta*delta;
      if (e <= delta) { const double aux = e2/delta2; rho[0] = delta2*(1.-std::pow((1.-aux),3))/3.; rho[1] = std::pow((1.-aux),2); rho[2] = -2.*(1.-aux)/delta2; }
      else { rho[0] = delta2/3.; rho[1] = 0; rho[2] = 0; }
      break; }
    case RK_SATURATED: {                               // :145-158
      double dsqr = delta*delta;
      if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.; rho[2] = 0.; } else { rho[0] = dsqr; rho[1] = 0.; rho[2] = 0.; }
      break; }
    case RK_DCS: {                                     // :160-171
      double scale = (2.0*delta)/(delta+e2); if (scale >= 1.0) scale = 1.0;
      rho[0] = scale*e2*scale; rho[1] = scale*scale; rho[2] = 0; break; }
    default: rho[0] = e2; rho[1] = 1.; rho[2] = 0.;
  }
}

// ---------------- SE3Quat: g2o/types/slam3d/se3quat.h ----------------
struct SE3Quat {
  Quat r; V3 t;
  SE3Quat() : r(Quat::identity()), t{{0,0,0}} {}
  SE3Quat(const Quat& q, const V3& t_) : r(q), t(t_) { normalizeRotation(); }        // :56-58
  void normalizeRotation() { if (r.w < 0) { r.x=-r.x; r.y=-r.y; r.z=-r.z; r.w=-r.w; } r.normalize(); }   // :270-275
  SE3Quat operator*(const SE3Quat& o) const {                                        // :99-105
    SE3Quat res(*this);
    res.t = res.t + rotate(r, o.t);
    res.r = res.r * o.r;
    res.normalizeRotation();
    return res;
  }
  SE3Quat inverse() const {                                                           // :118-123
    SE3Quat ret; ret.r = r.conjugate(); ret.t = rotate(ret.r, -1.0 * t); return ret;
  }
  V3 map(const V3& xyz) const { return rotate(r, xyz) + t; }                           // :211-214
  void log(double res[6]) const {                                                     // :173-209
    M3 R = toRotationMatrix(r);
    double d = 0.5*(R(0,0)+R(1,1)+R(2,2)-1);
    V3 omega; V3 dR = deltaR(R); M3 V_inv;
    if (d > 0.99999) {
      omega = 0.5*dR;
      M3 Omega = skew(omega);
      V_inv = M3::identity() - 0.5*Omega + (1./12.)*(Omega*Omega);
    } else {
      double theta = std::acos(d);
      omega = (theta/(2*std::sqrt(1-d*d)))*dR;
      M3 Omega = skew(omega);
      V_inv = M3::identity() - 0.5*Omega + ((1-theta/(2*std::tan(theta/2)))/(theta*theta))*(Omega*Omega);
    }
    V3 upsilon = V_inv * t;
    for (int i=0;i<3;++i) { res[i]=omega[i]; res[i+3]=upsilon[i]; }
  }
  static SE3Quat exp(const double update[6]) {                                        // :218-257
    V3 omega{{update[0],update[1],update[2]}}, upsilon{{update[3],update[4],update[5]}};
    double theta = std::sqrt(dot(omega, omega));
    M3 Omega = skew(omega), Omega2 = Omega*Omega, R, V;
    if (theta < 0.00001) {
      R = M3::identity() + Omega + 0.5*Omega2;
      V = M3::identity() + 0.5*Omega + (1./6.)*Omega2;
    } else {
      R = M3::identity() + (std::sin(theta)/theta)*Omega + ((1-std::cos(theta))/(theta*theta))*Omega2;
      V = M3::identity() + ((1-std::cos(theta))/(theta*theta))*Omega + ((theta-std::sin(theta))/(std::pow(theta,3)))*Omega2;
    }
    return SE3Quat(fromRotationMatrix(R), V*upsilon);
  }
  void adj(double res[36]) const {                                                    // :259-268 (6x6 col-major)
    M3 R = toRotationMatrix(r); M3 tR = skew(t)*R;
    std::memset(res, 0, sizeof(double)*36);
    for (int i=0;i<3;++i) for (int j=0;j<3;++j) { res[i+6*j]=R(i,j); res[(i+3)+6*(j+3)]=R(i,j); res[(i+3)+6*j]=tR(i,j); }
  }
  // exchange layout = SE3Quat::toVector (:132-142): tx ty tz qx qy qz qw
  void toVector(double v[7]) const { v[0]=t[0];v[1]=t[1];v[2]=t[2];v[3]=r.x;v[4]=r.y;v[5]=r.z;v[6]=r.w; }
  static SE3Quat fromVectorRaw(const double v[7]) { SE3Quat s; s.t={{v[0],v[1],v[2]}}; s.r={v[3],v[4],v[5],v[6]}; return s; }   // :144-147 (no normalisation)
};

// ---------------- Isometry3 as (R col-major, t): the VertexSE3 / EdgeSE3 estimate type ----------------
struct Iso3 {
  M3 R; V3 t;
  Iso3 operator*(const Iso3& o) const { Iso3 r; r.R = R*o.R; r.t = R*o.t + t; return r; }   // Eigen Isometry product
  Iso3 inverse() const { Iso3 r; r.R = transpose(R); r.t = -1.0*(r.R*t); return r; }          // Eigen Transform::inverse(Isometry)
  void store(double* d) const { std::memcpy(d, R.m, 72); d[9]=t[0]; d[10]=t[1]; d[11]=t[2]; }
  static Iso3 load(const double* d) { Iso3 r; std::memcpy(r.R.m, d, 72); r.t={{d[9],d[10],d[11]}}; return r; }
};
// isometry3d_mappings.cpp:40-46 normalize(q); :78-83 toCompactQuaternion; :93-98 toVectorMQT
inline void toVectorMQT(const Iso3& T, double v[6]) {
  Quat q = fromRotationMatrix(T.R);
  q.normalize();
  if (q.w < 0) { q.x=-q.x; q.y=-q.y; q.z=-q.z; q.w=-q.w; }
  v[0]=T.t[0]; v[1]=T.t[1]; v[2]=T.t[2]; v[3]=q.x; v[4]=q.y; v[5]=q.z;
}
// isometry3d_mappings.cpp:85-92 fromCompactQuaternion; :110-115 fromVectorMQT
inline Iso3 fromVectorMQT(const double v[6]) {
  Iso3 T;
  double w = 1 - (v[3]*v[3]+v[4]*v[4]+v[5]*v[5]);
  if (w < 0) T.R = M3::identity();
  else { w = std::sqrt(w); T.R = toRotationMatrix(Quat{v[3],v[4],v[5],w}); }
  T.t = {{v[0],v[1],v[2]}};
  return T;
}

// dquat2mat.cpp:35-85 + dquat2mat_maxima_generated.cpp:27-190 — d(q_xyz)/d(vec R), R vectorised column-major.
// Restated from the defining formulas rather than the generated tables: in every branch the quaternion
// components are  q_c = N_c / S4  with S4 = 2 sqrt(1 + s0 r00 + s1 r11 + s2 r22)  (= 4 * dominant component)
// and N_c linear in R; the dominant component itself is S4 / 4.
inline void compute_dq_dR(double dq_dR[27] /*3x9 col-major*/, const M3& R) {
  const double r00=R(0,0), r11=R(1,1), r22=R(2,2);
  const double tr = r00+r11+r22;
  int which; double sg[3];
  if (tr > 0) { which = 3; sg[0]=sg[1]=sg[2]=1; }
  else if ((r00 > r11) & (r00 > r22)) { which = 0; sg[0]=1; sg[1]=-1; sg[2]=-1; }
  else if (r11 > r22) { which = 1; sg[0]=-1; sg[1]=1; sg[2]=-1; }
  else { which = 2; sg[0]=-1; sg[1]=-1; sg[2]=1; }
  const double S4 = 2*std::sqrt(1.0 + sg[0]*r00 + sg[1]*r11 + sg[2]*r22);
  // numerators: N[c] = sum coef * R(entry); for the dominant slot N is unused.
  // dS4/dr_ii = sg[i] * 2 / S4
  double qw;
  std::memset(dq_dR, 0, sizeof(double)*27);
  auto D = [&](int row, int r, int c) -> double& { return dq_dR[row + 3*(r + 3*c)]; };
  struct Lin { int r1,c1; double s1; int r2,c2; double s2; };
  auto applyLin = [&](int row, const Lin& L) {
    const double N = L.s1*R(L.r1,L.c1) + L.s2*R(L.r2,L.c2);
    D(row, L.r1, L.c1) += L.s1 / S4;
    D(row, L.r2, L.c2) += L.s2 / S4;
    for (int i=0;i<3;++i) D(row, i, i) += -N/(S4*S4) * (sg[i]*2/S4);
  };
  if (which == 3) {
    qw = 0.25*S4;
    applyLin(0, {2,1, 1, 1,2,-1});
    applyLin(1, {0,2, 1, 2,0,-1});
    applyLin(2, {1,0, 1, 0,1,-1});
  } else {
    const int i = which, j = (i+1)%3, k = (j+1)%3;
    qw = (R(k,j)-R(j,k))/S4;
    for (int d=0; d<3; ++d) D(i, d, d) += 0.25*(sg[d]*2/S4);
    applyLin(j, {j,i, 1, i,j, 1});
    applyLin(k, {k,i, 1, i,k, 1});
  }
  if (qw <= 0) for (int n=0;n<27;++n) dq_dR[n] = -dq_dR[n];
}

// ---------------- SE2: g2o/types/slam2d/se2.h:39-120 ----------------
struct SE2 {
  double x, y, th;
  SE2 operator*(const SE2& o) const {                                                  // :62-75
    SE2 r(*this); const double c=std::cos(th), s=std::sin(th);
    r.x += c*o.x - s*o.y; r.y += s*o.x + c*o.y;
    r.th = normalize_theta(th + o.th); return r;
  }
  SE2 inverse() const {                                                                // :83-93
    SE2 r; r.th = normalize_theta(-th);
    const double c=std::cos(r.th), s=std::sin(r.th);
    const double tx=-x, ty=-y;
    r.x = c*tx - s*ty; r.y = s*tx + c*ty; return r;
  }
};

// ======================= vertex ⊞ =======================
// returns nothing; `counter` is VertexSE3::_numOplusCalls (vertex_se3.h:105-114)
inline void oplus(int vtype, double* est, const double* u, int* counter) {
  switch (vtype) {
    case VT_SE2: {                                       // vertex_se2.h:51-58
      est[0] += u[0]; est[1] += u[1]; est[2] = normalize_theta(est[2] + u[2]); break; }
    case VT_POINT_XY: est[0]+=u[0]; est[1]+=u[1]; break;  // vertex_point_xy.h:77-81
    case VT_SE3: {                                       // vertex_se3.h:105-114
      Iso3 X = Iso3::load(est);
      X = X * fromVectorMQT(u);
      if (++(*counter) > 1000) {                         // orthogonalizeAfter = 1000 (vertex_se3.h:54)
        *counter = 0;
        M3 E = transpose(X.R)*X.R;                        // isometry3d_mappings.h:81-86
        E(0,0)-=1; E(1,1)-=1; E(2,2)-=1;
        X.R = X.R - 0.5*(X.R*E);
      }
      X.store(est); break; }
    case VT_SE3_EXPMAP: {                                // types_six_dof_expmap.h:98-101
      SE3Quat T = SE3Quat::fromVectorRaw(est);
      T = SE3Quat::exp(u) * T;
      T.toVector(est); break; }
    case VT_POINT_XYZ: case VT_POINT_BAL: est[0]+=u[0]; est[1]+=u[1]; est[2]+=u[2]; break;   // types_sba.h:137-154; bal_example.cpp:127-131
    case VT_CAM_BAL: for (int i=0;i<9;++i) est[i]+=u[i]; break;                              // bal_example.cpp:90-94
  }
}

// ======================= edge error functions =======================
// BAL projection functor, bal_example.cpp:192-244, templated so that Jet<12> reproduces the autodiff Jacobian
template <typename T>
inline void balError(const T* camera, const T* point, const double* meas, T* error) {
  T p[3];
  T theta = sqrt(camera[0]*camera[0] + camera[1]*camera[1] + camera[2]*camera[2]);
  if (theta > T(0.0)) {
    T v[3] = { camera[0]/theta, camera[1]/theta, camera[2]/theta };
    T cth = cos(theta), sth = sin(theta);
    T vXp[3] = { v[1]*point[2]-v[2]*point[1], v[2]*point[0]-v[0]*point[2], v[0]*point[1]-v[1]*point[0] };
    T vDotp = v[0]*point[0]+v[1]*point[1]+v[2]*point[2];
    T oneMinusCth = T(1.0) - cth;
    for (int i=0;i<3;++i) p[i] = point[i]*cth + vXp[i]*sth + v[i]*vDotp*oneMinusCth;
  } else {
    T aux[3] = { camera[1]*point[2]-camera[2]*point[1], camera[2]*point[0]-camera[0]*point[2], camera[0]*point[1]-camera[1]*point[0] };
    for (int i=0;i<3;++i) p[i] = point[i] + aux[i];
  }
  p[0] = p[0] + camera[3]; p[1] = p[1] + camera[4]; p[2] = p[2] + camera[5];
  T pp[2] = { -p[0]/p[2], -p[1]/p[2] };
  T radiusSqr = pp[0]*pp[0] + pp[1]*pp[1];
  T f = camera[6], k1 = camera[7], k2 = camera[8];
  T r_p = T(1.0) + k1*radiusSqr + k2*radiusSqr*radiusSqr;
  error[0] = f*r_p*pp[0] - T(meas[0]);
  error[1] = f*r_p*pp[1] - T(meas[1]);
}

// e: output (edgeDim); x0,x1: vertex estimates in exchange/internal layout; z: measurement; prm: parameters
inline void computeError(int etype, const double* x0, const double* x1, const double* z, const double* prm, double* e) {
  switch (etype) {
    case ET_SE2: {                                       // edge_se2.h:45-52 (uses _inverseMeasurement = m.inverse(), :57-60)
      SE2 a{x0[0],x0[1],x0[2]}, b{x1[0],x1[1],x1[2]}, m{z[0],z[1],z[2]};
      SE2 d = m.inverse() * (a.inverse()*b);
      e[0]=d.x; e[1]=d.y; e[2]=d.th; break; }
    case ET_SE2_POINT_XY: {                              // edge_se2_pointxy.h:45-50
      SE2 a{x0[0],x0[1],x0[2]}; SE2 ai = a.inverse();
      const double c=std::cos(ai.th), s=std::sin(ai.th);
      e[0] = (ai.x + c*x1[0] - s*x1[1]) - z[0];
      e[1] = (ai.y + s*x1[0] + c*x1[1]) - z[1]; break; }
    case ET_SE3: {                                       // edge_se3.cpp:77-82
      Iso3 Xi = Iso3::load(x0), Xj = Iso3::load(x1), Z = Iso3::load(z);
      Iso3 delta = (Z.inverse() * Xi.inverse()) * Xj;
      toVectorMQT(delta, e); break; }
    case ET_SE3_EXPMAP: {                                // types_six_dof_expmap.h:117-124
      SE3Quat v1 = SE3Quat::fromVectorRaw(x0), v2 = SE3Quat::fromVectorRaw(x1), C = SE3Quat::fromVectorRaw(z);
      SE3Quat err = v2.inverse()*C*v1;
      err.log(e); break; }
    case ET_PROJECT_XYZ2UV: {                            // types_six_dof_expmap.h:140-147, .cpp:74-80; v0 = point, v1 = pose
      SE3Quat T = SE3Quat::fromVectorRaw(x1); V3 P = T.map(V3{{x0[0],x0[1],x0[2]}});
      e[0] = z[0] - (P[0]/P[2]*prm[0] + prm[1]);
      e[1] = z[1] - (P[1]/P[2]*prm[0] + prm[2]); break; }
    case ET_SE3_PROJECT_XYZ: {                           // types_six_dof_expmap.h:211-216, .cpp:449-455
      SE3Quat T = SE3Quat::fromVectorRaw(x1); V3 P = T.map(V3{{x0[0],x0[1],x0[2]}});
      e[0] = z[0] - (P[0]/P[2]*prm[0] + prm[2]);
      e[1] = z[1] - (P[1]/P[2]*prm[1] + prm[3]); break; }
    case ET_BAL: balError<double>(x0, x1, z, e); break;   // bal_example.cpp:246-252
  }
}

// J0 (E x D0), J1 (E x D1), column-major — the analytic linearizeOplus of each type
inline void linearizeOplus(int etype, const double* x0, const double* x1, const double* z, const double* prm, double* J0, double* J1) {
  switch (etype) {
    case ET_SE2: {                                       // edge_se2.cpp:77-103
      const double thetai = x0[2], dtx = x1[0]-x0[0], dty = x1[1]-x0[1];
      const double si=std::sin(thetai), ci=std::cos(thetai);
      const double A[9] = { -ci, si, 0,  -si, -ci, 0,  -si*dtx+ci*dty, -ci*dtx-si*dty, -1 };   // col-major
      const double B[9] = { ci, -si, 0,  si, ci, 0,  0, 0, 1 };
      SE2 m{z[0],z[1],z[2]}; SE2 mi = m.inverse();
      const double c=std::cos(mi.th), s=std::sin(mi.th);
      const double Zm[9] = { c, s, 0,  -s, c, 0,  0, 0, 1 };
      mm(Zm, A, J0, 3, 3, 3); mm(Zm, B, J1, 3, 3, 3); break; }
    case ET_SE2_POINT_XY: {                              // edge_se2_pointxy.cpp:68-93
      const double x1_=x0[0], y1_=x0[1], th1=x0[2], x2=x1[0], y2=x1[1];
      const double aux_1=std::cos(th1), aux_2=-aux_1, aux_3=std::sin(th1);
      J0[0+2*0]=aux_2; J0[0+2*1]=-aux_3; J0[0+2*2]=aux_1*y2-aux_1*y1_-aux_3*x2+aux_3*x1_;
      J0[1+2*0]=aux_3; J0[1+2*1]=aux_2;  J0[1+2*2]=-aux_3*y2+aux_3*y1_-aux_1*x2+aux_1*x1_;
      J1[0+2*0]=aux_1; J1[0+2*1]=aux_3; J1[1+2*0]=-aux_3; J1[1+2*1]=aux_1; break; }
    case ET_SE3: {                                       // edge_se3.cpp:92-105 -> isometry3d_gradients.h:192-255
      Iso3 Xi = Iso3::load(x0), Xj = Iso3::load(x1), Z = Iso3::load(z);
      const Iso3 A = Z.inverse(), B = Xi.inverse()*Xj, E = A*B;
      const M3& Re = E.R; const M3& Ra = A.R; const M3& Rb = B.R; const V3& tb = B.t;
      double dq_dR[27]; compute_dq_dR(dq_dR, Re);
      std::memset(J0, 0, sizeof(double)*36); std::memset(J1, 0, sizeof(double)*36);
      auto setBlock = [](double* J, int r0, int c0, const M3& M, double sgn) { for (int r=0;r<3;++r) for (int c=0;c<3;++c) J[(r0+r)+6*(c0+c)] = sgn*M(r,c); };
      setBlock(J0, 0, 0, Ra, -1.0);                      // dte/dti
      setBlock(J1, 0, 0, Re, 1.0);                       // dte/dtj
      { M3 S; const double x=2*tb[0], y=2*tb[1], zz=2*tb[2];   // skewT(S,tb) :49-54 (row-major comma initialiser)
        S(0,0)=0; S(0,1)=-zz; S(0,2)=y; S(1,0)=zz; S(1,1)=0; S(1,2)=-x; S(2,0)=-y; S(2,1)=x; S(2,2)=0;
        setBlock(J0, 0, 3, Ra*S, 1.0); }
      auto rowsToM3 = [](const double r[9]) { M3 S; for (int i=0;i<3;++i) for (int j=0;j<3;++j) S(i,j)=r[3*i+j]; return S; };
      auto dre = [&](const M3& L, const M3& Sx, const M3& Sy, const M3& Sz, double* J) {
        double buf[27]; M3 Mx=L*Sx, My=L*Sy, Mz=L*Sz;
        std::memcpy(buf, Mx.m, 72); std::memcpy(buf+9, My.m, 72); std::memcpy(buf+18, Mz.m, 72);   // 9x3 col-major
        double out[9]; mm(dq_dR, buf, out, 3, 9, 3);
        for (int r=0;r<3;++r) for (int c=0;c<3;++c) J[(3+r)+6*(3+c)] = out[r+3*c];
      };
      { const double r11=2*Rb(0,0), r12=2*Rb(0,1), r13=2*Rb(0,2), r21=2*Rb(1,0), r22=2*Rb(1,1), r23=2*Rb(1,2), r31=2*Rb(2,0), r32=2*Rb(2,1), r33=2*Rb(2,2);
        const double sx[9]={0,0,0, r31,r32,r33, -r21,-r22,-r23};        // skewT(Sx,Sy,Sz,R) :73-85
        const double sy[9]={-r31,-r32,-r33, 0,0,0, r11,r12,r13};
        const double sz[9]={r21,r22,r23, -r11,-r12,-r13, 0,0,0};
        dre(Ra, rowsToM3(sx), rowsToM3(sy), rowsToM3(sz), J0); }
      { const double sx[9]={0,0,0, 0,0,-2, 0,2,0};                       // skew(Sx,Sy,Sz,I) :57-70
        const double sy[9]={0,0,2, 0,0,0, -2,0,0};
        const double sz[9]={0,-2,0, 2,0,0, 0,0,0};
        dre(Re, rowsToM3(sx), rowsToM3(sy), rowsToM3(sz), J1); }
      break; }
    case ET_SE3_EXPMAP: {                                // types_six_dof_expmap.cpp:278-293
      SE3Quat Ti = SE3Quat::fromVectorRaw(x0), Tj = SE3Quat::fromVectorRaw(x1), Tij = SE3Quat::fromVectorRaw(z);
      SE3Quat invTij = Tij.inverse();
      SE3Quat invTj_Tij = Tj.inverse()*Tij;
      SE3Quat infTi_invTij = Ti.inverse()*invTij;
      invTj_Tij.adj(J0);
      infTi_invTij.adj(J1); for (int i=0;i<36;++i) J1[i] = -J1[i]; break; }
    case ET_PROJECT_XYZ2UV: case ET_SE3_PROJECT_XYZ: {   // types_six_dof_expmap.cpp:295-331 / :395-430
      const double fx = prm[0], fy = (etype==ET_PROJECT_XYZ2UV) ? prm[0] : prm[1];
      SE3Quat T = SE3Quat::fromVectorRaw(x1); V3 P = T.map(V3{{x0[0],x0[1],x0[2]}});
      const double x=P[0], y=P[1], zz=P[2], z_2=zz*zz;
      double tmp[6] = { fx, 0,  0, fy,  -x/zz*fx, -y/zz*fy };   // 2x3 col-major
      M3 R = toRotationMatrix(T.r);
      double tR[6]; mm(tmp, R.m, tR, 2, 3, 3);
      for (int i=0;i<6;++i) J0[i] = -1./zz * tR[i];
      J1[0+2*0]= x*y/z_2*fx;        J1[0+2*1]=-(1+(x*x/z_2))*fx; J1[0+2*2]= y/zz*fx; J1[0+2*3]=-1./zz*fx; J1[0+2*4]=0;          J1[0+2*5]=x/z_2*fx;
      J1[1+2*0]=(1+y*y/z_2)*fy;     J1[1+2*1]=-x*y/z_2*fy;       J1[1+2*2]=-x/zz*fy; J1[1+2*3]=0;         J1[1+2*4]=-1./zz*fy;  J1[1+2*5]=y/z_2*fy;
      break; }
    case ET_BAL: {                                       // bal_example.cpp:254-281 (ceres AutoDiff over 9+3 parameters)
      typedef Jet<12> J;
      J cam[9], pt[3], err[2];
      for (int i=0;i<9;++i) cam[i] = J(x0[i], i);
      for (int i=0;i<3;++i) pt[i] = J(x1[i], 9+i);
      balError<J>(cam, pt, z, err);
      for (int r=0;r<2;++r) { for (int c=0;c<9;++c) J0[r+2*c] = err[r].v[c]; for (int c=0;c<3;++c) J1[r+2*c] = err[r].v[9+c]; }
      break; }
  }
}

// numeric Jacobian of the reference's fallback: base_binary_edge.hpp:199-266 (delta 1e-9, central differences)
inline void linearizeOplusNumeric(int etype, const double* x0, const double* x1, const double* z, const double* prm, double* J0, double* J1) {
  const double delta = 1e-9, scalar = 1.0/(2*delta);
  const int E = edgeDim(etype);
  for (int side=0; side<2; ++side) {
    const int vt = edgeVertexType(etype, side), D = vertexDim(vt), S = vertexEstimateDim(vt);
    double* Jout = side==0 ? J0 : J1;
    for (int d=0; d<D; ++d) {
      double add[9] = {0}, e1[6], e2[6], buf[12]; int counter = 0;
      const double* xs = side==0 ? x0 : x1;
      std::memcpy(buf, xs, sizeof(double)*S); add[d] = delta; oplus(vt, buf, add, &counter);
      computeError(etype, side==0?buf:x0, side==0?x1:buf, z, prm, e1);
      std::memcpy(buf, xs, sizeof(double)*S); add[d] = -delta; oplus(vt, buf, add, &counter);
      computeError(etype, side==0?buf:x0, side==0?x1:buf, z, prm, e2);
      for (int r=0;r<E;++r) Jout[r+E*d] = scalar*(e1[r]-e2[r]);
    }
  }
}

}  // namespace orc
