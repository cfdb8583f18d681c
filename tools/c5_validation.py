#!/usr/bin/env python
"""Validation of the sizes the CPU reference cannot reach or confirm (BASELINE.json configs[4]).

Run with one process per GPU (>= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/c5_validation.py [--n 32768]

 1. sharded vs single GPU at N = 32768 (the largest size one B200 holds comfortably; the reference's `int`
    indices end there): one implicit time step to tol 1e-6 from the reference initial conditions on P ranks and on
    rank 0 alone -- cycle counts, residual histories and the field digest must be IDENTICAL (same arithmetic,
    different partitioning);
 2. why the cycle count grows with N: the same step on one GPU with the reference's coarse velocity towers
    (multigrid.cpp:148-160, never halving n: SURVEY.md 8 P1) and with true injection (options.correct_towers).
    The convergence factor per cycle is printed for both.  (Measured: the two tower variants converge at the same
    rate; the factor grows with N because dt = dx/10 at fixed nu makes the diffusion number 4 r |nu| = |nu| / (5 dx)
    grow with N -- 0.33 / 0.66 / 1.31 at N = 16384 / 32768 / 65536 -- and the reference's injection restriction
    loses efficiency as the operator turns from identity-dominated to Laplacian-dominated.)
Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hpcclassmultigridproject_b200 as mg  # noqa: E402
from hpcclassmultigridproject_b200 import digest  # noqa: E402


def describe(s, n, rank, world):
    info = s.timestep(1)[0]
    lo, hi, u = s.get_u_host_rows()
    d = digest.slab_digests(u, n, lo, hi, base_row=lo)
    parts = [d]
    if world > 1:
        parts = [None] * world if rank == 0 else None
        dist.gather_object(d, parts, dst=0)
    if rank != 0:
        return None
    allp = {}
    for q in parts:
        allp.update(q)
    hist = info.history()
    return {"cycles": info.cycles, "hist": hist, "factors": [hist[k + 1] / hist[k] for k in range(info.cycles)],
            "u_sha256": digest.combine(allp)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32768)
    ap.add_argument("--tol", type=float, default=1e-6)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n; dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    box = [mg.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    with mg.Solver(n, nu, dt, dx, args.tol, device=local, rank=rank, nranks=world, unique_id=box[0]) as s:
        s.set_fields_reference_ic(1.0)
        sharded = describe(s, n, rank, world)
    out = {"n": n, "tol": args.tol, "gpus": world, "sharded": sharded}
    if rank == 0:
        for key, ct in (("single_gpu", 0), ("single_gpu_corrected_towers", 1)):
            with mg.Solver(n, nu, dt, dx, args.tol, device=local, correct_towers=ct) as s:
                s.set_fields_reference_ic(1.0)
                out[key] = describe(s, n, 0, 1)
        a, b = out["sharded"], out["single_gpu"]
        # the field must be bit-identical; the norms are sums in a different (rank-ordered) order: last digits only
        out["hist_max_rel_diff"] = max(abs(x - y) / y for x, y in zip(a["hist"], b["hist"]))
        out["sharded_equals_single_gpu"] = bool(a["cycles"] == b["cycles"] and a["u_sha256"] == b["u_sha256"] and out["hist_max_rel_diff"] <= 1e-12)
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not out["sharded_equals_single_gpu"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
