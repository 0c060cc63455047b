#!/usr/bin/env python
"""Sharded Jacobian / Hessian with the result completed on every rank: peer-memory exchange fused into the
sweep kernels (fdb_exchange_*) against an NCCL all-gather after the sweeps.  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/exchange_timing.py

Prints one JSON line per workload from rank 0 (ms = max over ranks of the CUDA-event time of sweep + exchange).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "forwarddiff.jl_b200"))
import numpy as np
import torch
import torch.distributed as dist

import fdb200
from fdb200 import functions as F


def timed(step, reps, dev):
    for _ in range(3):
        step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        e0.record(); step(); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.median(ms))], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    ctx = fdb200.Context(local)
    ctx.set_async(True)
    rng = np.random.default_rng(2)
    reps = int(os.environ.get("REPS", "10"))
    out = []

    # C3: jacobian of tanh.(A*x .+ b), n = m = 8192, Chunk{12}: J is 512 MB, each rank produces 1/world of its columns
    n = m = int(os.environ.get("C3_N", "8192"))
    A = rng.uniform(-1, 1, (m, n)) / np.sqrt(n); b = rng.uniform(-1, 1, m); x = rng.uniform(-1, 1, n)
    f = F.tanh_affine.with_params(A=A, b=b)
    row = {"workload": f"C3 tanh.(A*x .+ b) Jacobian {m}x{n} Chunk{{12}}", "world": world, "result_MB": 8 * m * n / 1e6}
    J = {}
    for mode in ("nccl", "p2p"):
        sj = fdb200.dist.ShardedJacobian(ctx, f, n, m, fdb200.JacobianConfig(f, x, fdb200.Chunk(12)), exchange=mode)
        sj.x_t.copy_(torch.from_numpy(x))

        def step():
            sj.sweep(); sj.gather()
        row[mode + "_ms"] = timed(step, reps, dev)
        row[mode + "_sweep_only_ms"] = timed(sj.sweep, reps, dev)
        J[mode] = sj.gather()[: m * n].clone()
        sj.close(); del sj
    row["identical"] = bool(torch.equal(J["nccl"], J["p2p"]))
    del J
    out.append(row)

    # C5: hessian of Rosenbrock, n = 2048, Chunk{8}: H is 33.5 MB
    nh = 2048
    xh = np.random.default_rng(5).random(nh)
    row = {"workload": f"C5 Rosenbrock Hessian n={nh} Chunk{{8}}", "world": world, "result_MB": 8 * nh * nh / 1e6}
    H = {}
    for mode in ("nccl", "p2p"):
        sh = fdb200.dist.ShardedHessian(ctx, F.rosenbrock, nh, fdb200.HessianConfig(F.rosenbrock, xh, fdb200.Chunk(8)), exchange=mode)
        sh.x_t.copy_(torch.from_numpy(xh))

        def step():
            sh.sweep(); sh.gather(); sh.value_and_gradient()
        row[mode + "_ms"] = timed(step, reps, dev)
        H[mode] = sh.gather()[: nh * nh].clone()
        sh.close(); del sh
    row["identical"] = bool(torch.equal(H["nccl"], H["p2p"]))
    out.append(row)

    # C4 (reduced so that the gathered matrix stays small): Brusselator Jacobian n = 16384, Chunk{8}: J is 2.1 GB
    M = int(os.environ.get("C4_M", "8192")); nb = 2 * M
    xb = np.concatenate([1 + np.sin(2 * np.pi * np.arange(1, M + 1) / (M + 1)), 3 * np.ones(M)])
    fb = F.brusselator.with_params(alpha=(M + 1) ** 2 / 50.0)
    row = {"workload": f"C4-like Brusselator Jacobian n={nb} Chunk{{8}}", "world": world, "result_MB": 8 * nb * nb / 1e6}
    for mode in ("nccl", "p2p"):
        sb = fdb200.dist.ShardedJacobian(ctx, fb, nb, nb, fdb200.JacobianConfig(fb, xb, fdb200.Chunk(8), y=np.zeros(nb)), exchange=mode)
        sb.x_t.copy_(torch.from_numpy(xb))

        def step():
            sb.sweep(); sb.gather()
        row[mode + "_ms"] = timed(step, reps, dev)
        sb.close(); del sb
    out.append(row)

    if rank == 0:
        for r in out:
            print(json.dumps(r), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
