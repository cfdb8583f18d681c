#!/usr/bin/env python3
"""Scaling harnesses of the reference, on libmgb200.

  nsweep   grid-size sweep of mg_timer.cu:210-268: for N = 32, 64, ... run the full `timestepper`
           (reference initial conditions, T = 100 dt, V-cycle, tol 1e-6) on device arrays plus the
           device->host copy of uT, timed with CUDA events; prints the reference's
           "Time elapsed for grid size N: X ms" lines and writes `cudatime.txt`
           ("N<TAB>seconds", the table speedupplot.py:38-50 reads).
  strong   GPU-count sweep in the spirit of multigrid_strongsc.cpp:246-262 (which sweeps OpenMP
           threads): one fixed problem on 1, 2, 4, ... GPUs, one process per GPU through torchrun;
           writes `strong_scale.txt` ("count<TAB>seconds", the table strongsc_plot.py:51-60 reads).

The product path only: every number comes from the C ABI on a B200; nothing here touches oracle/.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def nsweep(args):
    import torch
    import hpcclassmultigridproject_b200 as mg

    torch.cuda.set_device(args.device)
    arith = mg.ARITH_EXACT if args.arith == "exact" else mg.ARITH_FAST
    rows = []
    n = args.nmin
    while n <= args.nmax:
        dx = 1.0 / n; dt = dx / 10
        maxlvl = mg.maxlvl_for(n)
        u0, v1, v2, uT = (torch.empty(n + 1, n + 1, dtype=torch.float64, device="cuda") for _ in range(4))
        mg.ops.initial_conditions(u0, v1, v2, n, 1.0)                     # multigrid.cpp:206-233 on the device
        host = torch.empty(n + 1, n + 1, dtype=torch.float64).pin_memory()
        best = None
        for rep in range(1 + args.reps):                                   # first repetition builds the handle
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            info = mg.timestepper_device(uT, u0, v1, v2, args.nu, maxlvl, n, dt, args.steps * dt, dx, args.tol, args.shape,
                                         arith=arith, device=args.device)
            host.copy_(uT, non_blocking=True)                              # mg_timer.cu:258
            ev1.record()
            ev1.synchronize()
            ms = ev0.elapsed_time(ev1)
            if rep >= 1:
                best = ms if best is None else min(best, ms)
        print(f"Time elapsed for grid size {n}: {best:g} ms   (last step: {info.cycles} cycle(s))", flush=True)
        rows.append((n, best * 1e-3))
        mg.release_cached()
        del u0, v1, v2, uT, host
        torch.cuda.empty_cache()
        n *= 2
    with open(args.out, "w") as f:
        for n, s in rows:
            f.write(f"{n}\t{s:f}\n")
    print(f"wrote {args.out}")


def strong(args):
    rows = []
    g = 1
    while g <= args.max_gpus:
        cmd = [sys.executable]
        if g > 1:
            cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={g}", "--master-addr", "127.0.0.1",
                    "--master-port", str(29600 + g)]
        cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(g), "--steps", str(args.steps), "--warmup", "3", "--size", str(args.n),
                "--no-cpu", "--no-e2e"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            print(f"{g} GPU(s): failed\n{r.stderr[-2000:]}", file=sys.stderr)
            break
        d = json.loads(line[-1])
        sec = d["ms_per_step"] * 1e-3                                      # one implicit time step
        print(f"{g} GPU(s): {d['ms_per_step']:.3f} ms per time step, {d['value']:.3f} ms per V-cycle", flush=True)
        rows.append((g, sec))
        g *= 2
    with open(args.out, "w") as f:
        for g, s in rows:
            f.write(f"{g}\t{s:f}\n")
    print(f"wrote {args.out}")


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("nsweep")
    a.add_argument("--nmin", type=int, default=32)
    a.add_argument("--nmax", type=int, default=4096)                       # mg_timer.cu:212
    a.add_argument("--steps", type=int, default=100)                       # T = 100 dt, mg_timer.cu:245
    a.add_argument("--nu", type=float, default=-4e-4)
    a.add_argument("--tol", type=float, default=1e-6)
    a.add_argument("--shape", type=int, default=1)
    a.add_argument("--arith", choices=["fast", "exact"], default="fast")
    a.add_argument("--reps", type=int, default=2)
    a.add_argument("--device", type=int, default=0)
    a.add_argument("--out", default="cudatime.txt")
    a.set_defaults(fn=nsweep)
    b = sub.add_parser("strong")
    b.add_argument("--n", type=int, default=16384)
    b.add_argument("--max-gpus", type=int, default=8)
    b.add_argument("--steps", type=int, default=5)
    b.add_argument("--out", default="strong_scale.txt")
    b.set_defaults(fn=strong)
    args = ap.parse_args()
    args.fn(args)


if __name__ == "__main__":
    main()
