"""Multi-GPU: one process per GPU, chunk sweeps sharded by contiguous chunk ranges.

The reference runs its ceil(n/N) chunk sweeps sequentially on one thread
(src/gradient.jl:141-147); they are independent (chunk c only produces gradient
entries / Jacobian columns c*N+1 .. min((c+1)*N, n)), so rank r of W owns chunks
[r*ceil(C/W), min((r+1)*ceil(C/W), C)) and therefore a contiguous slice of the
result.  There is no collective inside the sweep; the only exchange is one
all-gather of the (padded, equal-sized) slices, done with torch.distributed
(NCCL over NVLink on GPUs, gloo in the CPU tests).  The value f(x) is taken from
the rank that owns the last chunk, as the reference does (src/gradient.jl:155).

torch is only the plumbing here (device memory for the gathered buffer, the
process group); all compute is the C-ABI library.
"""
import numpy as np

from . import _lib
from .api import GradientConfig, HessianConfig, JacobianConfig, chunk_plan, checktag, _catalogue_id


def slice_layout(xlen, N, world):
    """Elements per rank (padded to whole chunks) and total padded length of the gathered buffer."""
    p = chunk_plan(xlen, N, 0, world)
    per_chunks = -(-p["nchunks"] // world)
    per = per_chunks * N
    return dict(nchunks=p["nchunks"], per_chunks=per_chunks, per=per, padded=per * world)


def owner_of_last_chunk(xlen, N, world):
    lay = slice_layout(xlen, N, world)
    return (lay["nchunks"] - 1) // lay["per_chunks"]


def allgather_slices(local_padded, world, group=None):
    """local_padded: this rank's 1-D torch tensor of `per * width` values -> gathered tensor of world * that."""
    import torch
    import torch.distributed as dist
    out = torch.empty(local_padded.numel() * world, dtype=local_padded.dtype, device=local_padded.device)
    dist.all_gather_into_tensor(out, local_padded.contiguous(), group=group)
    return out


def _rank_world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class ShardedGradient:
    """Reusable state for ForwardDiff.gradient!(out, f, x, cfg) sharded over the process group:
    device buffers are allocated once (like a GradientConfig's work buffers)."""

    def __init__(self, ctx, f, xlen, cfg=None, group=None):
        import torch
        self.torch = torch
        self.ctx, self.f, self.n, self.group = ctx, f, int(xlen), group
        self.cfg = cfg if cfg is not None else GradientConfig(f, np.empty(self.n))
        self.N = self.cfg.N
        self.rank, self.world = _rank_world(group)
        self.fid = _catalogue_id(f, _lib.FN)
        if self.fid is None:
            raise _lib.UnsupportedOp("sharded gradient needs a catalogue target")
        self.plan = chunk_plan(self.n, self.N, self.rank, self.world)
        self.lay = slice_layout(self.n, self.N, self.world)
        dev = torch.device("cuda", ctx.device)
        self.x_t = torch.empty(self.n, dtype=torch.float64, device=dev)
        # gathered buffer: rank r's slice sits at [r*per, (r+1)*per) == its absolute gradient positions
        self.g_t = torch.zeros(max(self.lay["padded"], self.n), dtype=torch.float64, device=dev)
        self.v_t = torch.zeros(1, dtype=torch.float64, device=dev)
        self.x_a = ctx.wrap(self.x_t.data_ptr(), self.n, keepalive=self.x_t)
        self.g_a = ctx.wrap(self.g_t.data_ptr(), self.n, ld=self.g_t.numel(), keepalive=self.g_t)
        self.v_a = ctx.wrap(self.v_t.data_ptr(), 1, keepalive=self.v_t)
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)     # kernels and NCCL on one stream

    def sweep(self):
        """This rank's chunk sweeps (device-resident inputs); no host sync."""
        c = self.ctx
        c.check(c.lib.fdb_gradient_chunked(c.h, self.fid, self.x_a.h, self.N, self.plan["chunk_lo"], self.plan["chunk_hi"],
                                           self.g_a.h, self.v_a.h))

    def gather(self):
        """All-gather the gradient slices (in place in g_t) and broadcast the value from the owner of the last chunk."""
        import torch.distributed as dist
        if self.world > 1:
            per = self.lay["per"]
            mine = self.g_t[self.rank * per:(self.rank + 1) * per].clone()
            dist.all_gather_into_tensor(self.g_t[:per * self.world], mine, group=self.group)
            dist.broadcast(self.v_t, src=owner_of_last_chunk(self.n, self.N, self.world), group=self.group)

    def __call__(self, x_host_pinned, out_host_pinned=None):
        """End to end: H2D of x, sweeps, gather, D2H of the full gradient (rank-local copy)."""
        self.x_t.copy_(x_host_pinned, non_blocking=True)
        self.sweep()
        self.gather()
        if out_host_pinned is not None:
            out_host_pinned.copy_(self.g_t[:self.n], non_blocking=True)
            self.torch.cuda.current_stream().synchronize()
            return out_host_pinned
        return self.g_t[:self.n]
