#!/usr/bin/env python
"""Multi-GPU parity check: sharded gradient / Jacobian / Hessian (one rank per GPU, NCCL) against the same
result computed by a single rank.  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py
(also works as a plain `python tools/check_sharded.py` with one GPU: world size 1).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "forwarddiff.jl_b200"))
import numpy as np
import torch
import torch.distributed as dist

import fdb200
from fdb200 import functions as F


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    ctx = fdb200.Context(local)
    modes = ["nccl", "p2p"]
    if world > 1:
        sup = [None] * world
        dist.all_gather_object(sup, fdb200.dist.multicast_supported(ctx))
        if all(sup):
            modes.append("multicast")          # NVSwitch multicast stores (csrc/fdb_multicast.cu)
    for exchange in modes:
        check(ctx, rank, world, exchange)
    if world > 1:
        dist.barrier()
    print(f"rank {rank}/{world}: sharded gradient / Jacobian / Hessian match the single-rank results ({', '.join(modes)} exchange)", flush=True)
    if world > 1:
        dist.destroy_process_group()


def check(ctx, rank, world, exchange):
    """exchange="nccl": all-gather after the sweeps; "p2p": the sweep kernels store into every peer's arena; "multicast": they
    store once through the NVSwitch multicast mapping of all arenas."""
    rng = np.random.default_rng(3)

    # gradient: Ackley n = 100 003 (remainder chunk), Chunk{12}
    n, N = 100_003, 12
    x = rng.uniform(0.5, 1.5, n)
    sg = fdb200.dist.ShardedGradient(ctx, F.ackley, n, fdb200.GradientConfig(F.ackley, x, fdb200.Chunk(N)), exchange=exchange)
    assert sg.exchange == exchange
    g = sg(torch.from_numpy(x).pin_memory()).cpu().numpy()
    ref = fdb200.GradientResult(x)
    fdb200.gradient_b(ref, F.ackley, x, fdb200.GradientConfig(F.ackley, x, fdb200.Chunk(N)), ctx=ctx)
    assert np.array_equal(g, ref.derivs[0]), "sharded gradient differs from the single-rank result"
    assert sg.v_t.item() == ref.value

    # Jacobian: Brusselator n = 4096, Chunk{8}; tanh.(A*x .+ b) 256 x 160, Chunk{12} (DMMA path)
    M = 2048; nb = 2 * M
    xb = np.concatenate([1 + np.sin(2 * np.pi * np.arange(1, M + 1) / (M + 1)), 3 * np.ones(M)])
    fb = F.brusselator.with_params(alpha=(M + 1) ** 2 / 50.0)
    sj = fdb200.dist.ShardedJacobian(ctx, fb, nb, nb, fdb200.JacobianConfig(fb, xb, fdb200.Chunk(8), y=np.zeros(nb)), exchange=exchange)
    sj.x_t.copy_(torch.from_numpy(xb)); sj.sweep()
    Jfull = sj.full_matrix(); yv = sj.values().cpu().numpy()
    yref = np.zeros(nb)
    Jref = fdb200.jacobian(fb, xb, fdb200.JacobianConfig(fb, xb, fdb200.Chunk(8), y=yref), ctx=ctx, y=yref)
    assert np.array_equal(Jfull, Jref) and np.array_equal(yv, yref), "sharded Brusselator Jacobian differs"
    blk = sj.local_block()
    assert np.array_equal(blk, Jref[:, sj.plan["elem_lo"]:sj.plan["elem_hi"]])

    m, k = 256, 160
    A = rng.uniform(-1, 1, (m, k)) / np.sqrt(k); b = rng.uniform(-1, 1, m); xt = rng.uniform(-1, 1, k)
    ft = F.tanh_affine.with_params(A=A, b=b)
    st = fdb200.dist.ShardedJacobian(ctx, ft, k, m, fdb200.JacobianConfig(ft, xt, fdb200.Chunk(12)), exchange=exchange)
    st.x_t.copy_(torch.from_numpy(xt)); st.sweep()
    Jt = st.full_matrix()
    Jtref = fdb200.jacobian(ft, xt, fdb200.JacobianConfig(ft, xt, fdb200.Chunk(12)), ctx=ctx, m=m)
    assert np.array_equal(Jt, Jtref), "sharded tanh_affine Jacobian differs"

    # Hessian: Rosenbrock n = 515, Chunk{8}
    nh = 515
    xh = rng.uniform(0.5, 1.5, nh)
    sh = fdb200.dist.ShardedHessian(ctx, F.rosenbrock, nh, fdb200.HessianConfig(F.rosenbrock, xh, fdb200.Chunk(8)), exchange=exchange)
    sh.x_t.copy_(torch.from_numpy(xh)); sh.sweep()
    H = sh.full_matrix(); v, gh = sh.value_and_gradient()
    href = fdb200.HessianResult(xh)
    fdb200.hessian_b(href, F.rosenbrock, xh, fdb200.HessianConfig(F.rosenbrock, xh, fdb200.Chunk(8)), ctx=ctx)
    assert np.array_equal(H, href.derivs[1]) and np.array_equal(gh.cpu().numpy(), href.derivs[0]) and v.item() == href.value

    # a second sweep reuses the arenas (pre-sweep barrier protects the previous result) and must reproduce the first
    sg.x_t.copy_(torch.from_numpy(x)); sg.sweep(); sg.gather()
    assert np.array_equal(sg.g_t[:n].cpu().numpy(), ref.derivs[0])
    if exchange in ("p2p", "multicast"):
        # with an exchange enabled a result outside the arena is refused, not silently left un-mirrored
        sg.xch.enable()
        try:
            stray = ctx.zeros(n)
            rc = ctx.lib.fdb_gradient_chunked(ctx.h, sg.fid, sg.x_a.h, N, 0, 1, stray.h, None)
            assert rc != 0 and b"exchange arena" in ctx.lib.fdb_last_error(ctx.h)
        finally:
            sg.xch.disable()
        assert sg.xch.timeouts() == 0 and sj.xch.timeouts() == 0
    if exchange in ("p2p", "multicast"):
        # all-gather of data no driver mirrored: every rank fills its slice of a fresh arena, pushes it, barrier
        per = 1000
        xa = (fdb200.dist.PeerExchange if exchange == "p2p" else fdb200.dist.MulticastExchange)(ctx, per * world)
        mine = xa.torch_view(rank * per, per)
        mine.copy_(torch.arange(per, dtype=torch.float64, device=mine.device) + 1000.0 * rank)
        xa.push(rank * per, per); xa.barrier()
        allv = xa.torch_view(0, per * world).cpu().numpy()
        assert np.array_equal(allv, np.concatenate([np.arange(per) + 1000.0 * r for r in range(world)]))
        xa.close()
    for s in (sg, sj, st, sh):
        s.close()


if __name__ == "__main__":
    main()
