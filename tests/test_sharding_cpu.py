"""Host-side logic of the landmark-sharded (multi-GPU) path, on the CPU: the partition of edges and landmarks over ranks,
and the allreduce hook over ``gloo`` with world_size 2."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from g2o_b200 import _lib
from g2o_b200 import workloads as W
from g2o_b200.binding import CudaSolver, G2oCudaError


def _host(graph, name, rank, world):
    s = CudaSolver(graph, name)
    if world > 1:
        s.set_shard(rank, world, lambda *a: 0)
    s.initialize_optimization()
    try:
        s.build_structure()
    except G2oCudaError as e:
        assert e.code == _lib.E_CUDA
    return s


def test_partition_covers_every_edge_once():
    for g, name in [(W.bal_synthetic(n_cameras=40, n_points=3000, n_obs=14000, seed=3, k_max=30, min_window=4), "lm_fix9_3_cuda"),
                    (W.slam2d(n_poses=600, n_landmarks=150, world_size=30.0), "lm_fix3_2_cuda"),
                    (W.sphere(nodes_per_level=12, laps=6), "lm_var_cuda")]:
        ref = _host(g, name, 0, 1)
        n_active = len(ref.get_i32("active_edges"))
        dims = ref.get_i32("dims")
        for world in (2, 3, 8):
            seen = np.zeros(n_active, dtype=np.int64)
            ranges = []
            for rank in range(world):
                s = _host(g, name, rank, world)
                seen[s.get_i32("shard_edge_positions")] += 1
                ranges.append(tuple(s.get_i32("shard_landmark_range")))
                if dims[1]:   # slab PCG: equal, contiguous block ranges of the reduced system, covering it exactly once
                    lo, hi = s.get_i32("slab_block_range")
                    nnz = int(ref.get_i32("hschur_colptr")[-1]); c = -(-nnz // world)
                    assert (lo, hi) == (min(nnz, rank * c), min(nnz, (rank + 1) * c))
                # the global structure is identical on every rank
                for arr in ("hessian_index", "hpp_colptr", "hpp_rowidx"):
                    assert np.array_equal(s.get_i32(arr), ref.get_i32(arr))
                if dims[1]:
                    for arr in ("hschur_colptr", "hschur_rowidx", "hpl_colptr"):
                        assert np.array_equal(s.get_i32(arr), ref.get_i32(arr))
            assert np.all(seen == 1)
            if dims[1]:
                assert ranges[0][0] == 0 and ranges[-1][1] == dims[1]
                assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
                # balanced by the cost model (0.40 ns per observation + 0.085 ns per Schur block product), not by landmark count
                cp = ref.get_i32("hpl_colptr").astype(np.int64)
                k = np.diff(cp)
                cost = np.concatenate([[0], np.cumsum(400 * k + 85 * (k * (k + 1) // 2))])
                loads = [cost[b] - cost[a] for a, b in ranges]
                assert max(loads) - min(loads) <= max(2 * int((400 * k + 85 * (k * (k + 1) // 2)).max()), 0.2 * cost[-1] / world)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from g2o_b200.dist import make_allreduce
    fn = make_allreduce(backend_is_cuda=False)
    a = np.arange(5, dtype=np.float64) + 10 * rank
    assert fn(a.ctypes.data, a.size, 0, 0) == 0
    b = np.array([rank, -rank, 3.5], dtype=np.float64)
    assert fn(b.ctypes.data, b.size, 1, 0) == 0
    c = np.arange(3 * world, dtype=np.float64) * (rank + 1)          # in-place reduce-scatter: rank r keeps the sum of range r
    assert fn(c.ctypes.data, 3, 2, 0) == 0
    expect = np.arange(3 * world, dtype=np.float64) * sum(range(1, world + 1))
    assert np.array_equal(c[3 * rank:3 * rank + 3], expect[3 * rank:3 * rank + 3])
    q.put((rank, a.tolist(), b.tolist()))
    dist.destroy_process_group()


def test_allreduce_hook_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, a, b in out:
        assert a == [10.0, 12.0, 14.0, 16.0, 18.0]
        assert b == [1.0, 0.0, 3.5]
