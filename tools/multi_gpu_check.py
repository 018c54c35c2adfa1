"""Real multi-GPU check of the sharded entry points (run under torchrun on 2+ GPUs):
epoch-sharded synthesis with an NCCL gather and row-block-sharded covariance propagation with a
broadcast of Sigma, both compared with the same call done by one GPU alone.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import grates_b200 as gb
from grates_b200 import distributed as gd
from oracle import sh_oracle as orc


def main():
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N, E, d = 96, 48, 0.5
    grid = gb.GeographicGrid(d, d)
    anm = np.stack([orc.synthetic_coefficients(N, e) for e in range(E)])
    gd.synthesis_sharded(anm, grid, "ewh", gather=True)                       # warm-up (plans, NCCL)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    full, _ = gd.synthesis_sharded(anm, grid, "ewh", gather=True)
    torch.cuda.synchronize(); t_syn = time.perf_counter() - t0
    alone = gb.to_grid_batch(torch.as_tensor(anm).cuda(), grid, "ewh")
    err_syn = float((full - alone).abs().max() / alone.abs().max())
    sigma = orc.synthetic_covariance(N) if rank == 0 else None
    gd.covariance_propagation_sharded(sigma, grid, 0, N, "ewh", src=0)          # warm-up
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    std = gd.covariance_propagation_sharded(sigma, grid, 0, N, "ewh", src=0)
    torch.cuda.synchronize(); t_cov = time.perf_counter() - t0
    out = None
    if rank == 0:
        ref = gb.get_plan(grid, N, "ewh").covariance_propagation(torch.as_tensor(sigma).cuda(), 0).reshape(-1)
        out = {"world_size": world, "synthesis_sharded_gather_s": t_syn, "synthesis_vs_single_gpu": err_syn,
               "covprop_sharded_s_incl_708MB_broadcast": t_cov,
               "covprop_vs_single_gpu": float((std - ref).abs().max() / ref.abs().max())}
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
