"""g2o_b200 — B200-native solver backend for g2o's LM / BlockSolver hot path."""
from . import graph, workloads  # noqa: F401

__all__ = ["graph", "workloads"]
