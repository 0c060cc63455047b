#!/usr/bin/env python
"""All ranks push into all peers at once: outgoing GB/s per rank for the store patterns of fdb_exchange_measure_store_bw.
Run under torchrun (one rank per GPU)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "forwarddiff.jl_b200"))
import torch
import torch.distributed as dist

import fdb200

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = fdb200.Context(local)
n = int(os.environ.get("PROBE_MB", "256")) * 1024 * 1024 // 8
xch = fdb200.dist.PeerExchange(ctx, n)
res = {}
for variant, name in ((0, "st8_per_peer"), (1, "st16_per_peer"), (2, "peer_per_block")):
    for rotate in (0, 1):
        g = C.c_double()
        ctx.check(ctx.lib.fdb_exchange_measure_store_bw(xch.h, n, variant, rotate, C.byref(g)))
        t = torch.tensor([g.value], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        res[f"{name}{'_rotated' if rotate else ''}"] = round(t.item(), 1)
if rank == 0:
    print(json.dumps({"world": world, "MB_per_peer": n * 8 / 1e6, "egress_GBps_min_over_ranks": res}), flush=True)
xch.close()
dist.destroy_process_group()
