#include "g2o_io.hpp"

#include <cmath>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <vector>

namespace g2o {
namespace {

// unit quaternion (x, y, z, w) -> rotation matrix, column-major (Eigen::Quaternion::toRotationMatrix)
void quatToR(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[3] = txy - twz;       R[6] = txz + twy;
  R[1] = txy + twz;       R[4] = 1 - (txx + tzz); R[7] = tyz - twx;
  R[2] = txz - twy;       R[5] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
// rotation matrix (column-major) -> unit quaternion (x, y, z, w), Eigen's branch on the trace / largest diagonal entry
void rToQuat(const double R[9], double q[4]) {
  auto m = [&](int r, int c) { return R[r + 3 * c]; };
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t; q[1] = (m(0, 2) - m(2, 0)) * t; q[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0; if (m(1, 1) > m(0, 0)) i = 1; if (m(2, 2) > m(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t; q[j] = (m(j, i) + m(i, j)) * t; q[k] = (m(k, i) + m(i, k)) * t;
  }
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int a = 0; a < 4; ++a) q[a] /= n;
}
bool readUpperInformation(std::istream& is, OptimizableGraph::Edge* e) {   // rows i, columns j >= i; mirrored (edge_se3.cpp:53-59)
  const int d = e->dimension(); double* I = e->informationData();
  for (int i = 0; i < d; ++i)
    for (int j = i; j < d; ++j) { double v; if (!(is >> v)) return false; I[i + (size_t)d * j] = v; I[j + (size_t)d * i] = v; }
  return true;
}
void writeUpperInformation(std::ostream& os, OptimizableGraph::Edge* e) {
  const int d = e->dimension(); const double* I = e->informationData();
  for (int i = 0; i < d; ++i) for (int j = i; j < d; ++j) os << " " << I[i + (size_t)d * j];
}
// SE3Quat (q, t) helpers for the EXPMAP file convention: files store camera-to-world, the vertex holds world-to-camera
// (VertexSE3Expmap::read / write, types_six_dof_expmap.cpp:92-112)
void invertQT(const double t[3], const double q[4], double ti[3], double qi[4]) {
  qi[0] = -q[0]; qi[1] = -q[1]; qi[2] = -q[2]; qi[3] = q[3];
  double R[9]; quatToR(qi, R);
  for (int r = 0; r < 3; ++r) ti[r] = -(R[r] * t[0] + R[r + 3] * t[1] + R[r + 6] * t[2]);
}

}  // namespace

bool loadG2o(std::istream& is, SparseOptimizer& opt, LoadReport* rep) {
  LoadReport local; LoadReport& R = rep ? *rep : local;
  std::map<int, std::vector<double>> cameraParams;   // PARAMS_CAMERAPARAMETERS id -> f, cx, cy (baseline ignored by EdgeProjectXYZ2UV)
  std::string line;
  while (std::getline(is, line)) {
    std::istringstream ls(line);
    std::string tag;
    if (!(ls >> tag) || tag[0] == '#') continue;
    if (tag == "FIX") { int id; while (ls >> id) { auto* v = opt.vertex(id); if (v) { v->setFixed(true); ++R.fixed; } else std::cerr << "Warning: Unable to fix vertex with id " << id << ". Not found in the graph." << std::endl; } continue; }
    if (tag == "PARAMS_CAMERAPARAMETERS") { int id; double f, cx, cy, b; if (ls >> id >> f >> cx >> cy >> b) cameraParams[id] = {f, cx, cy}; continue; }
    OptimizableGraph::Vertex* v = nullptr;
    if (tag == "VERTEX_SE2") { int id; double x, y, th; if (!(ls >> id >> x >> y >> th)) return false; auto* p = new VertexSE2; p->setId(id); p->setEstimate(x, y, th); v = p; }
    else if (tag == "VERTEX_XY") { int id; double x, y; if (!(ls >> id >> x >> y)) return false; auto* p = new VertexPointXY; p->setId(id); p->setEstimate(x, y); v = p; }
    else if (tag == "VERTEX_SE3:QUAT") { int id; double e[7]; if (!(ls >> id)) return false; for (double& a : e) if (!(ls >> a)) return false;
      double Rm[9]; quatToR(e + 3, Rm); auto* p = new VertexSE3; p->setId(id); p->setEstimate(Rm, e); v = p; }
    else if (tag == "VERTEX_SE3:EXPMAP") { int id; double e[7]; if (!(ls >> id)) return false; for (double& a : e) if (!(ls >> a)) return false;
      double n = std::sqrt(e[3] * e[3] + e[4] * e[4] + e[5] * e[5] + e[6] * e[6]); for (int a = 3; a < 7; ++a) e[a] /= n;
      double ti[3], qi[4]; invertQT(e, e + 3, ti, qi); if (qi[3] < 0) for (double& a : qi) a = -a;
      auto* p = new VertexSE3Expmap; p->setId(id); p->setEstimate(ti, qi); v = p; }
    else if (tag == "VERTEX_XYZ") { int id; double x, y, z; if (!(ls >> id >> x >> y >> z)) return false; auto* p = new VertexSBAPointXYZ; p->setId(id); p->setEstimate(x, y, z); v = p; }
    if (v) { if (!opt.addVertex(v)) { std::cerr << "loadG2o: duplicate vertex id " << v->id() << std::endl; delete v; return false; } ++R.vertices; continue; }

    OptimizableGraph::Edge* e = nullptr; int i0 = -1, i1 = -1;
    if (tag == "EDGE_SE2") { double x, y, th; if (!(ls >> i0 >> i1 >> x >> y >> th)) return false; auto* p = new EdgeSE2; p->setMeasurement(x, y, th); e = p; if (!readUpperInformation(ls, e)) { delete e; return false; } }
    else if (tag == "EDGE_SE2_XY") { double x, y; if (!(ls >> i0 >> i1 >> x >> y)) return false; auto* p = new EdgeSE2PointXY; p->setMeasurement(x, y); e = p; if (!readUpperInformation(ls, e)) { delete e; return false; } }
    else if (tag == "EDGE_SE3:QUAT") { double m[7]; if (!(ls >> i0 >> i1)) return false; for (double& a : m) if (!(ls >> a)) return false;
      const double n = std::sqrt(m[3] * m[3] + m[4] * m[4] + m[5] * m[5] + m[6] * m[6]); for (int a = 3; a < 7; ++a) m[a] /= n;   // edge_se3.cpp:44-48
      double Rm[9]; quatToR(m + 3, Rm); auto* p = new EdgeSE3; p->setMeasurement(Rm, m); e = p; if (!readUpperInformation(ls, e)) { delete e; return false; } }
    else if (tag == "EDGE_PROJECT_XYZ2UV:EXPMAP") { int pid; double u, w; if (!(ls >> i0 >> i1 >> pid >> u >> w)) return false;
      auto it = cameraParams.find(pid); if (it == cameraParams.end()) { std::cerr << "loadG2o: EDGE_PROJECT_XYZ2UV:EXPMAP refers to unknown PARAMS_CAMERAPARAMETERS " << pid << std::endl; return false; }
      auto* p = new EdgeProjectXYZ2UV; p->setMeasurement(u, w); p->setCameraParameters(it->second[0], it->second[1], it->second[2]); e = p; if (!readUpperInformation(ls, e)) { delete e; return false; } }
    if (e) {
      auto* a = opt.vertex(i0); auto* b = opt.vertex(i1);
      e->setVertex(0, a); e->setVertex(1, b);
      if (!a || !b || !opt.addEdge(e)) { std::cerr << "loadG2o: edge " << tag << " " << i0 << " " << i1 << " refers to a missing or wrongly typed vertex" << std::endl; delete e; return false; }
      ++R.edges; continue;
    }
    if (R.firstUnknownTag.empty()) { R.firstUnknownTag = tag; std::cerr << "loadG2o: unknown type: " << tag << " (skipped, as OptimizableGraph::load does)" << std::endl; }
    ++R.skippedLines;
  }
  return true;
}

bool saveG2o(std::ostream& os, const SparseOptimizer& opt) {
  os << std::setprecision(17);
  std::map<std::vector<double>, int> paramIds;
  for (auto* e : opt.edgeList())
    if (e->typeCode() == G2OCU_EDGE_PROJECT_XYZ2UV) {
      std::vector<double> prm(e->parameterData(), e->parameterData() + 3);
      if (!paramIds.count(prm)) { const int id = (int)paramIds.size(); paramIds[prm] = id; os << "PARAMS_CAMERAPARAMETERS " << id << " " << prm[0] << " " << prm[1] << " " << prm[2] << " 0\n"; }
    }
  for (auto* v : opt.vertexList()) {
    const auto& x = v->estimateVector();
    switch (v->typeCode()) {
      case G2OCU_VERTEX_SE2: os << "VERTEX_SE2 " << v->id() << " " << x[0] << " " << x[1] << " " << x[2] << "\n"; break;
      case G2OCU_VERTEX_POINT_XY: os << "VERTEX_XY " << v->id() << " " << x[0] << " " << x[1] << "\n"; break;
      case G2OCU_VERTEX_SE3: { double q[4]; rToQuat(x.data(), q); os << "VERTEX_SE3:QUAT " << v->id() << " " << x[9] << " " << x[10] << " " << x[11] << " " << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << "\n"; break; }
      case G2OCU_VERTEX_SE3_EXPMAP: { double ti[3], qi[4]; invertQT(x.data(), x.data() + 3, ti, qi); os << "VERTEX_SE3:EXPMAP " << v->id() << " " << ti[0] << " " << ti[1] << " " << ti[2] << " " << qi[0] << " " << qi[1] << " " << qi[2] << " " << qi[3] << "\n"; break; }
      case G2OCU_VERTEX_POINT_XYZ: os << "VERTEX_XYZ " << v->id() << " " << x[0] << " " << x[1] << " " << x[2] << "\n"; break;
      default: std::cerr << "saveG2o: vertex type " << v->typeCode() << " has no .g2o tag" << std::endl; return false;
    }
    if (v->fixed()) os << "FIX " << v->id() << "\n";
  }
  for (auto* e : opt.edgeList()) {
    const double* m = e->measurementData();
    const int a = e->vertex(0)->id(), b = e->vertex(1)->id();
    switch (e->typeCode()) {
      case G2OCU_EDGE_SE2: os << "EDGE_SE2 " << a << " " << b << " " << m[0] << " " << m[1] << " " << m[2]; break;
      case G2OCU_EDGE_SE2_POINT_XY: os << "EDGE_SE2_XY " << a << " " << b << " " << m[0] << " " << m[1]; break;
      case G2OCU_EDGE_SE3: { double q[4]; rToQuat(m, q); os << "EDGE_SE3:QUAT " << a << " " << b << " " << m[9] << " " << m[10] << " " << m[11] << " " << q[0] << " " << q[1] << " " << q[2] << " " << q[3]; break; }
      case G2OCU_EDGE_PROJECT_XYZ2UV: { std::vector<double> prm(e->parameterData(), e->parameterData() + 3); os << "EDGE_PROJECT_XYZ2UV:EXPMAP " << a << " " << b << " " << paramIds[prm] << " " << m[0] << " " << m[1]; break; }
      default: std::cerr << "saveG2o: edge type " << e->typeCode() << " has no .g2o tag here" << std::endl; return false;
    }
    writeUpperInformation(os, e);
    os << "\n";
  }
  return os.good();
}

// examples/bal/bal_example.cpp:336-414: header "numCameras numPoints numObservations", observations "cam point u v", then 9 values per
// camera (angle-axis, t, f, k1, k2) and 3 per point; cameras get ids 0.., points follow and are marginalized; information = identity
bool loadBal(std::istream& is, SparseOptimizer& opt, LoadReport* rep) {
  LoadReport local; LoadReport& R = rep ? *rep : local;
  long nc = 0, np = 0, no = 0;
  if (!(is >> nc >> np >> no) || nc <= 0 || np <= 0 || no <= 0) return false;
  std::vector<long> oc(no), op(no); std::vector<double> ou(no), ov(no);
  for (long k = 0; k < no; ++k) if (!(is >> oc[k] >> op[k] >> ou[k] >> ov[k])) return false;
  for (long c = 0; c < nc; ++c) { double p[9]; for (double& a : p) if (!(is >> a)) return false; auto* v = new VertexCameraBAL; v->setId((int)c); v->setEstimate(p); if (!opt.addVertex(v)) return false; ++R.vertices; }
  for (long q = 0; q < np; ++q) { double x, y, z; if (!(is >> x >> y >> z)) return false; auto* v = new VertexPointBAL; v->setId((int)(nc + q)); v->setEstimate(x, y, z); v->setMarginalized(true); if (!opt.addVertex(v)) return false; ++R.vertices; }
  for (long k = 0; k < no; ++k) {
    if (oc[k] < 0 || oc[k] >= nc || op[k] < 0 || op[k] >= np) return false;
    auto* e = new EdgeObservationBAL; e->setVertex(0, opt.vertex((int)oc[k])); e->setVertex(1, opt.vertex((int)(nc + op[k]))); e->setMeasurement(ou[k], ov[k]);
    if (!opt.addEdge(e)) { delete e; return false; }
    ++R.edges;
  }
  return true;
}

}  // namespace g2o
