"""Worker of tests/test_gpu_sharded.py: one process per GPU (torchrun), sharded solve of a small and a
mid-size problem, checked on rank 0 against the single-GPU result of the same build and the oracle."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hpcclassmultigridproject_b200 as mg  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {}
    for n, steps, shard_min, vscale in ((1024, 2, 32, 2.0), (4096, 1, 128, 1.0)):
        dx = 1.0 / n; dt = dx / 10; nu = -4e-4; tol = 1e-10
        box = [mg.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        with mg.Solver(n, nu, dt, dx, tol, arith=mg.ARITH_EXACT, device=local, rank=rank, nranks=world, unique_id=box[0],
                       shard_min_rows=shard_min) as s:
            s.set_fields_reference_ic(vscale)
            infos = s.timestep(steps)
            mine = s.get_u_host()
            w = s.slab(0)
            nshard = sum(1 for l in range(s.maxlvl) if s.slab(l)["sharded"])
        # stitch the owned rows on rank 0
        rows = torch.from_numpy(np.nan_to_num(mine, nan=0.0)).cuda()
        dist.reduce(rows, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            got = rows.cpu().numpy()
            with mg.Solver(n, nu, dt, dx, tol, arith=mg.ARITH_EXACT, device=local) as ref:
                ref.set_fields_reference_ic(vscale)
                rinfos = ref.timestep(steps)
                want = ref.get_u_host()
            out[str(n)] = dict(bitwise=bool(np.array_equal(got, want)), rel=float(np.linalg.norm(got - want) / np.linalg.norm(want)),
                               cycles=[i.cycles for i in infos], ref_cycles=[i.cycles for i in rinfos],
                               hist_rel=float(max(abs(a - b) / b for i, j in zip(infos, rinfos) for a, b in zip(i.history(), j.history()))),
                               sharded_levels=nshard, own=[w["own_lo"], w["own_hi"]])
    if rank == 0:
        print("SHARDED_RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
