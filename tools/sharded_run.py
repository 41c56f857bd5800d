#!/usr/bin/env python
"""One large pair split over the ranks of a torchrun job (BASELINE.json configs[4]: N = 50000,
5 % inliers): `python -m torch.distributed.run --nproc-per-node G tools/sharded_run.py [--n N]`.

Each rank owns one GPU and calls Registrar.register_sharded (three library phases, two NCCL
exchanges over NVLink).  Rank 0 also runs the unsharded path on its GPU and checks that every
rank's result equals it bit for bit, then prints one JSON line with the timings (max over ranks).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sac_cot_b200 import synth  # noqa: E402
from sac_cot_b200.api import Registrar  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=50000)
    ap.add_argument("--ratio", type=float, default=0.05)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    p = synth.make_pair(args.points, args.ratio, 7000)
    reg = Registrar(device=local)
    times = []
    for _ in range(args.reps):
        dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        R, t, inl = reg.register_sharded(p.src, p.dst)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times.append(float(dt.item()))
    # unsharded reference on every rank's own GPU (same library, one device)
    t0 = time.perf_counter()
    R1, t1, i1 = reg.register(p.src, p.dst)
    t_single = time.perf_counter() - t0
    t0 = time.perf_counter()
    R1, t1, i1 = reg.register(p.src, p.dst)
    t_single = min(t_single, time.perf_counter() - t0)
    same = bool((R == R1).all() and (t == t1).all() and inl == i1)
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ang, dtr = synth.pose_error(R, t, p.R_gt, p.t_gt)
    if rank == 0:
        print(json.dumps({
            "workload": f"single pair N={args.points}, {args.ratio:.0%} inliers, sharded over {world} GPUs",
            "world": world, "sharded_s": times, "sharded_best_s": min(times), "unsharded_1gpu_s": t_single,
            "all_ranks_bit_identical_to_unsharded": bool(flag.item()), "inliers": int(inl),
            "rot_err_deg": float(np.degrees(ang)), "trans_err": dtr,
        }))
    reg.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
