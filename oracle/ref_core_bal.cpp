// TEST INFRASTRUCTURE.  The BAL types of the reference (VertexCameraBAL, VertexPointBAL, EdgeObservationBAL with its ceres-autodiff Jacobian)
// live in g2o/examples/bal/bal_example.cpp next to that program's main(); the file is compiled here as it lies, its main() renamed by the
// preprocessor and never called, and three factory functions hand the types to oracle/ref_core.cpp.
#define main bal_example_main_unused
#include "g2o/examples/bal/bal_example.cpp"
#undef main

extern "C" {
g2o::OptimizableGraph::Vertex* refbal_new_camera(const double* est9) {
  VertexCameraBAL* v = new VertexCameraBAL;
  Eigen::VectorXd e(9); for (int i = 0; i < 9; ++i) e[i] = est9[i];
  v->setEstimate(e);
  return v;
}
g2o::OptimizableGraph::Vertex* refbal_new_point(const double* est3) {
  VertexPointBAL* v = new VertexPointBAL;
  v->setEstimate(Eigen::Vector3d(est3[0], est3[1], est3[2]));
  return v;
}
g2o::OptimizableGraph::Edge* refbal_new_edge(const double* z2, const double* info4) {
  EdgeObservationBAL* e = new EdgeObservationBAL;
  e->setMeasurement(Eigen::Vector2d(z2[0], z2[1]));
  Eigen::Matrix2d I; for (int c = 0; c < 2; ++c) for (int r = 0; r < 2; ++r) I(r, c) = info4[r + 2 * c];
  e->setInformation(I);
  return e;
}
void refbal_camera_estimate(const g2o::OptimizableGraph::Vertex* v, double* out9) { const Eigen::VectorXd& e = static_cast<const VertexCameraBAL*>(v)->estimate(); for (int i = 0; i < 9; ++i) out9[i] = e[i]; }
void refbal_point_estimate(const g2o::OptimizableGraph::Vertex* v, double* out3) { const Eigen::Vector3d& e = static_cast<const VertexPointBAL*>(v)->estimate(); for (int i = 0; i < 3; ++i) out3[i] = e[i]; }
}
