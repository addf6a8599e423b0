"""torchrun worker: landmark-sharded LM over all ranks vs. the single-GPU run of the same problem (tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from g2o_b200 import workloads as W  # noqa: E402
from g2o_b200.binding import CudaSolver  # noqa: E402
from g2o_b200.dist import install_nccl, install_p2p, install_torch_allreduce  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [("bal", W.bal_synthetic(n_cameras=60, n_points=6000, n_obs=30000, seed=5, k_max=40, min_window=4), "lm_fix9_3_cuda"),
             ("slam2d", W.slam2d(n_poses=800, n_landmarks=200, world_size=30.0), "lm_fix3_2_cuda"),
             ("sphere", W.sphere(nodes_per_level=16, laps=8), "lm_var_cuda")]
    for ci, (name, g, solver) in enumerate(cases):
        s = CudaSolver(g, solver, device=local)
        # both transports of the collectives: the library's own NCCL communicator and the host callback over torch.distributed
        if ci % 2 == 0:
            install_nccl(s, rank, world)
        else:
            install_torch_allreduce(s, rank, world)
        s.initialize_optimization()
        if name == "bal":                      # slab PCG with the peer-memory exchange of q (CUDA IPC over NVLink)
            s.init(); s.build_structure()
            install_p2p(s, rank, world)
        n, st = s.optimize(6)
        est = s.get_estimates()
        if rank == 0:
            r = CudaSolver(g, solver, device=local)
            r.initialize_optimization()
            nr, str_ = r.optimize(6)
            estr = r.get_estimates()
            assert n == nr, (name, n, nr)
            for a, b in zip(st, str_):
                assert abs(a["chi2"] - b["chi2"]) <= 1e-6 * abs(b["chi2"]), (name, a["chi2"], b["chi2"])
                assert a["levenberg_iterations"] == b["levenberg_iterations"]
                assert abs(a["lambda"] - b["lambda"]) <= 1e-6 * abs(b["lambda"])
            err = np.max(np.abs(est - estr) / (1 + np.abs(estr)))
            assert err < 1e-5, (name, err)   # both sides stop PCG at a relative residual of 1e-6
            print(f"{name}: sharded x{world} matches single GPU, chi2 {st[-1]['chi2']:.6f}, max rel est diff {err:.2e}", flush=True)
        # the shard-restricted transfers: the ranks' read-backs, each into a buffer of NaNs, tile the complete vector (poses on every rank, a
        # landmark on exactly one); uploading them again changes nothing
        own = np.full_like(est, np.nan); s.get_estimates_owned(own)
        have = ~np.isnan(own)
        assert np.array_equal(own[have], est[have]), name
        cover = torch.from_numpy(have.astype(np.int32)).cuda(); dist.all_reduce(cover)
        marg = np.repeat(np.asarray(g.v_marginalized, dtype=bool), np.diff(g.estimate_offsets()))
        restricted = torch.tensor([int((~have).any())], device="cuda"); dist.all_reduce(restricted, op=dist.ReduceOp.MIN)
        if name == "bal":
            assert int(restricted) == 1, "the BAL case must take the shard-restricted path"
        if int(restricted):                    # (graphs with inactive vertices fall back to the complete transfer)
            assert bool((cover.cpu().numpy()[marg] == 1).all()) and bool((cover.cpu().numpy()[~marg] == world).all()), name
        s.set_estimates_owned(np.where(have, own, 0.0)); s.compute_active_errors()
        chi_after = s.active_robust_chi2()
        assert abs(chi_after - st[-1]["chi2"]) <= 1e-9 * abs(st[-1]["chi2"]), (name, chi_after, st[-1]["chi2"])
        # all ranks hold the same complete estimate vector
        t = torch.from_numpy(est.copy()).cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), name
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK", flush=True)


if __name__ == "__main__":
    main()
