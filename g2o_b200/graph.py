"""Flat (structure-of-arrays) description of a g2o graph, as the C-ABI takes it.

The layout mirrors what a g2o-side adapter reads off ``SparseOptimizer``:
vertices in insertion order (``OptimizableGraph::addVertex``), edges in ``internalId`` order
(``optimizable_graph.cpp:267-292``), per-vertex ``fixed``/``marginalized`` flags
(``optimizable_graph.h:103-346``), per-edge measurement / information / robust kernel /
parameters (``optimizable_graph.h:348-480``).  Numeric codes are the ones in ``include/g2ocu.h``.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

# vertex types (g2ocu.h: G2OCU_VERTEX_*)
VERTEX_SE2 = 1            # g2o/types/slam2d/vertex_se2.h            estimate (x, y, theta)
VERTEX_POINT_XY = 2       # g2o/types/slam2d/vertex_point_xy.h       (x, y)
VERTEX_SE3 = 3            # g2o/types/slam3d/vertex_se3.h            Isometry3: R col-major (9) + t (3)
VERTEX_SE3_EXPMAP = 4     # g2o/types/sba/types_six_dof_expmap.h:84  SE3Quat::toVector: t (3) + q xyzw (4)
VERTEX_POINT_XYZ = 5      # g2o/types/sba/types_sba.h:137            (x, y, z)
VERTEX_CAM_BAL = 6        # g2o/examples/bal/bal_example.cpp:65      (rx ry rz tx ty tz f k1 k2)
VERTEX_POINT_BAL = 7      # g2o/examples/bal/bal_example.cpp:102     (x, y, z)

# edge types (g2ocu.h: G2OCU_EDGE_*); vertex order inside the edge is the reference's
EDGE_SE2 = 1              # (VertexSE2, VertexSE2)               meas (x, y, theta)
EDGE_SE2_POINT_XY = 2     # (VertexSE2, VertexPointXY)           meas (x, y)
EDGE_SE3 = 3              # (VertexSE3, VertexSE3)               meas Isometry3 (12)
EDGE_SE3_EXPMAP = 4       # (VertexSE3Expmap, VertexSE3Expmap)   meas SE3Quat (7)
EDGE_PROJECT_XYZ2UV = 5   # (VertexSBAPointXYZ, VertexSE3Expmap) meas (u, v); param (f, cx, cy)
EDGE_SE3_PROJECT_XYZ = 6  # (VertexSBAPointXYZ, VertexSE3Expmap) meas (u, v); param (fx, fy, cx, cy)
EDGE_BAL = 7              # (VertexCameraBAL, VertexPointBAL)    meas (u, v)

# robust kernels (g2o/core/robust_kernel_impl.cpp)
KERNEL_NONE, KERNEL_HUBER, KERNEL_PSEUDO_HUBER, KERNEL_CAUCHY, KERNEL_GEMAN_MCCLURE, \
    KERNEL_WELSCH, KERNEL_FAIR, KERNEL_TUKEY, KERNEL_SATURATED, KERNEL_DCS = range(10)
KERNEL_BY_NAME = {"": 0, "Huber": 1, "PseudoHuber": 2, "Cauchy": 3, "GemanMcClure": 4, "Welsch": 5,
                  "Fair": 6, "Tukey": 7, "Saturated": 8, "DCS": 9}

VERTEX_ESTIMATE_DIM = np.array([0, 3, 2, 12, 7, 3, 9, 3], dtype=np.int64)
VERTEX_DIM = np.array([0, 3, 2, 6, 6, 3, 9, 3], dtype=np.int64)
EDGE_DIM = np.array([0, 3, 2, 6, 6, 2, 2, 2], dtype=np.int64)
EDGE_MEAS_DIM = np.array([0, 3, 2, 12, 7, 2, 2, 2], dtype=np.int64)
EDGE_PARAM_DIM = np.array([0, 0, 0, 0, 0, 3, 4, 0], dtype=np.int64)
EDGE_VERTEX_TYPES = {
    EDGE_SE2: (VERTEX_SE2, VERTEX_SE2), EDGE_SE2_POINT_XY: (VERTEX_SE2, VERTEX_POINT_XY),
    EDGE_SE3: (VERTEX_SE3, VERTEX_SE3), EDGE_SE3_EXPMAP: (VERTEX_SE3_EXPMAP, VERTEX_SE3_EXPMAP),
    EDGE_PROJECT_XYZ2UV: (VERTEX_POINT_XYZ, VERTEX_SE3_EXPMAP), EDGE_SE3_PROJECT_XYZ: (VERTEX_POINT_XYZ, VERTEX_SE3_EXPMAP),
    EDGE_BAL: (VERTEX_CAM_BAL, VERTEX_POINT_BAL),
}


class CGraph(ctypes.Structure):
    """``g2ocu_graph`` (include/g2ocu.h).  The CPU oracle takes the same layout."""
    _fields_ = [
        ("n_vertices", ctypes.c_int32),
        ("v_id", ctypes.POINTER(ctypes.c_int32)),
        ("v_type", ctypes.POINTER(ctypes.c_int32)),
        ("v_fixed", ctypes.POINTER(ctypes.c_uint8)),
        ("v_marginalized", ctypes.POINTER(ctypes.c_uint8)),
        ("v_estimate", ctypes.POINTER(ctypes.c_double)),
        ("n_edges", ctypes.c_int32),
        ("e_type", ctypes.POINTER(ctypes.c_int32)),
        ("e_v0", ctypes.POINTER(ctypes.c_int32)),
        ("e_v1", ctypes.POINTER(ctypes.c_int32)),
        ("e_level", ctypes.POINTER(ctypes.c_int32)),
        ("e_measurement", ctypes.POINTER(ctypes.c_double)),
        ("e_information", ctypes.POINTER(ctypes.c_double)),
        ("e_kernel", ctypes.POINTER(ctypes.c_int32)),
        ("e_kernel_delta", ctypes.POINTER(ctypes.c_double)),
        ("e_param", ctypes.POINTER(ctypes.c_double)),
    ]


def _ptr(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


@dataclass
class Graph:
    """Host-side graph.  All arrays are contiguous numpy arrays; packed arrays (``v_estimate``,
    ``e_measurement``, ``e_information``, ``e_param``) are concatenated in vertex / edge order with the
    per-type strides above (information is E x E column-major)."""
    v_id: np.ndarray
    v_type: np.ndarray
    v_fixed: np.ndarray
    v_marginalized: np.ndarray
    v_estimate: np.ndarray
    e_type: np.ndarray
    e_v0: np.ndarray
    e_v1: np.ndarray
    e_measurement: np.ndarray
    e_information: np.ndarray
    e_level: np.ndarray | None = None
    e_kernel: np.ndarray | None = None
    e_kernel_delta: np.ndarray | None = None
    e_param: np.ndarray | None = None
    name: str = ""
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        c = np.ascontiguousarray
        self.v_id = c(self.v_id, dtype=np.int32)
        self.v_type = c(self.v_type, dtype=np.int32)
        self.v_fixed = c(self.v_fixed, dtype=np.uint8)
        self.v_marginalized = c(self.v_marginalized, dtype=np.uint8)
        self.v_estimate = c(self.v_estimate, dtype=np.float64)
        self.e_type = c(self.e_type, dtype=np.int32)
        self.e_v0 = c(self.e_v0, dtype=np.int32)
        self.e_v1 = c(self.e_v1, dtype=np.int32)
        self.e_measurement = c(self.e_measurement, dtype=np.float64)
        self.e_information = c(self.e_information, dtype=np.float64)
        ne = self.e_type.shape[0]
        self.e_level = c(np.zeros(ne) if self.e_level is None else self.e_level, dtype=np.int32)
        self.e_kernel = c(np.zeros(ne) if self.e_kernel is None else self.e_kernel, dtype=np.int32)
        self.e_kernel_delta = c(np.ones(ne) if self.e_kernel_delta is None else self.e_kernel_delta, dtype=np.float64)
        self.e_param = c(np.zeros(0) if self.e_param is None else self.e_param, dtype=np.float64)
        self.validate()

    # ---- sizes -------------------------------------------------------------------------------
    @property
    def n_vertices(self) -> int:
        return int(self.v_id.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.e_type.shape[0])

    def estimate_offsets(self) -> np.ndarray:
        off = np.zeros(self.n_vertices + 1, dtype=np.int64)
        np.cumsum(VERTEX_ESTIMATE_DIM[self.v_type], out=off[1:])
        return off

    def validate(self) -> None:
        nv, ne = self.n_vertices, self.n_edges
        for a in (self.v_type, self.v_fixed, self.v_marginalized):
            if a.shape != (nv,):
                raise ValueError("vertex arrays must all have n_vertices entries")
        for a in (self.e_v0, self.e_v1, self.e_level, self.e_kernel, self.e_kernel_delta):
            if a.shape != (ne,):
                raise ValueError("edge arrays must all have n_edges entries")
        if nv and (self.v_type.min() < 1 or self.v_type.max() > 7):
            raise ValueError("unsupported vertex type code")
        if ne and (self.e_type.min() < 1 or self.e_type.max() > 7):
            raise ValueError("unsupported edge type code")
        if self.v_estimate.shape[0] != int(VERTEX_ESTIMATE_DIM[self.v_type].sum()):
            raise ValueError("v_estimate length does not match the vertex types")
        if self.e_measurement.shape[0] != int(EDGE_MEAS_DIM[self.e_type].sum()):
            raise ValueError("e_measurement length does not match the edge types")
        if self.e_information.shape[0] != int((EDGE_DIM[self.e_type] ** 2).sum()):
            raise ValueError("e_information length does not match the edge types")
        if self.e_param.shape[0] != int(EDGE_PARAM_DIM[self.e_type].sum()):
            raise ValueError("e_param length does not match the edge types")
        if ne and (min(self.e_v0.min(), self.e_v1.min()) < 0 or max(self.e_v0.max(), self.e_v1.max()) >= nv):
            raise ValueError("edge vertex index out of range")

    def as_c(self) -> CGraph:
        """ctypes view; keeps ``self`` alive only as long as the caller does."""
        g = CGraph()
        g.n_vertices = self.n_vertices
        g.v_id = _ptr(self.v_id, ctypes.c_int32)
        g.v_type = _ptr(self.v_type, ctypes.c_int32)
        g.v_fixed = _ptr(self.v_fixed, ctypes.c_uint8)
        g.v_marginalized = _ptr(self.v_marginalized, ctypes.c_uint8)
        g.v_estimate = _ptr(self.v_estimate, ctypes.c_double)
        g.n_edges = self.n_edges
        g.e_type = _ptr(self.e_type, ctypes.c_int32)
        g.e_v0 = _ptr(self.e_v0, ctypes.c_int32)
        g.e_v1 = _ptr(self.e_v1, ctypes.c_int32)
        g.e_level = _ptr(self.e_level, ctypes.c_int32)
        g.e_measurement = _ptr(self.e_measurement, ctypes.c_double)
        g.e_information = _ptr(self.e_information, ctypes.c_double)
        g.e_kernel = _ptr(self.e_kernel, ctypes.c_int32)
        g.e_kernel_delta = _ptr(self.e_kernel_delta, ctypes.c_double)
        g.e_param = _ptr(self.e_param, ctypes.c_double)
        return g

    def copy(self) -> "Graph":
        return Graph(**{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()
                        if k not in ("name", "meta")}, name=self.name, meta=dict(self.meta))

    def set_robust_kernel(self, kind: int | str, delta: float = 1.0) -> None:
        """What ``g2o -robustKernel <name> -robustKernelWidth <delta>`` does (apps/g2o_cli/g2o.cpp:333-357)."""
        k = KERNEL_BY_NAME[kind] if isinstance(kind, str) else int(kind)
        self.e_kernel[:] = k
        self.e_kernel_delta[:] = delta
