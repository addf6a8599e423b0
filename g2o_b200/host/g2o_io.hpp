// Text I/O of the host mirror: the .g2o graph format for the vertex / edge classes the backend supports and the BAL problem format.
//   OptimizableGraph::load / save        g2o/core/optimizable_graph.cpp:397-640 (tags, FIX, '#' comments, unknown tags skipped with a warning)
//   per-type read / write                types/slam2d/{vertex_se2,vertex_point_xy,edge_se2,edge_se2_pointxy}.cpp, types/slam3d/{vertex_se3,edge_se3}.cpp
//                                        (t + unit quaternion qx qy qz qw, information as upper-triangular rows), types/sba/types_six_dof_expmap.cpp
//   BAL files                            examples/bal/bal_example.cpp:336-414
#pragma once
#include <iosfwd>
#include <string>

#include "g2o_mirror.hpp"

namespace g2o {

struct LoadReport { size_t vertices = 0, edges = 0, fixed = 0, skippedLines = 0; std::string firstUnknownTag; };

bool loadG2o(std::istream& is, SparseOptimizer& optimizer, LoadReport* report = nullptr);
bool saveG2o(std::ostream& os, const SparseOptimizer& optimizer);
bool loadBal(std::istream& is, SparseOptimizer& optimizer, LoadReport* report = nullptr);   // cameras 0..Nc-1, points Nc.., points marginalized

}  // namespace g2o
