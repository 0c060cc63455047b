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


class _ShardedColumns:
    """Shared plumbing of the sharded Jacobian / Hessian: rank r owns the contiguous column block
    [elem_lo, elem_hi) of the column-major m x n result (its chunks' columns, src/jacobian.jl:109-118).
    Only that block is allocated per rank (`local`, m x per); the kernels address columns absolutely, so
    the wrapped base pointer is shifted by -elem_lo columns.  `gather()` all-gathers the padded blocks
    into the full matrix on every rank (optional: C4's 34 GB Jacobian is normally left sharded)."""

    def _setup(self, ctx, n, m, N, group):
        import torch
        self.torch = torch
        self.ctx, self.n, self.m, self.N, self.group = ctx, int(n), int(m), int(N), group
        self.rank, self.world = _rank_world(group)
        self.plan = chunk_plan(self.n, self.N, self.rank, self.world)
        self.lay = slice_layout(self.n, self.N, self.world)
        self.dev = torch.device("cuda", ctx.device)
        per = self.lay["per"]
        self.x_t = torch.empty(self.n, dtype=torch.float64, device=self.dev)
        self.local = torch.zeros(self.m * per, dtype=torch.float64, device=self.dev)
        self.full = None
        self.x_a = ctx.wrap(self.x_t.data_ptr(), self.n, keepalive=self.x_t)
        shifted = self.local.data_ptr() - 8 * self.m * self.plan["elem_lo"]
        self.J_a = ctx.wrap(shifted, self.m * self.n, keepalive=self.local)      # virtual m x n view, only own columns backed
        ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)

    def local_block(self):
        """This rank's column block as an (m, ncols) column-major numpy array."""
        ncols = self.plan["elem_hi"] - self.plan["elem_lo"]
        return self.local[: self.m * ncols].cpu().numpy().reshape((self.m, ncols), order="F")

    def gather(self):
        import torch.distributed as dist
        per = self.lay["per"]
        if self.full is None:
            self.full = self.torch.empty(self.m * per * self.world, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_gather_into_tensor(self.full, self.local, group=self.group)
        else:
            self.full.copy_(self.local)
        return self.full[: self.m * self.n]

    def full_matrix(self):
        return self.gather().cpu().numpy().reshape((self.m, self.n), order="F")


class ShardedJacobian(_ShardedColumns):
    """ForwardDiff.jacobian!(J, f, x, cfg) / jacobian!(J, f!, y, x, cfg) with chunk sweeps sharded over the ranks."""

    def __init__(self, ctx, f, xlen, m, cfg=None, group=None):
        self.f = f
        self.cfg = cfg if cfg is not None else JacobianConfig(f, np.empty(int(xlen)))
        self.fid = _catalogue_id(f, _lib.JFN)
        if self.fid is None:
            raise _lib.UnsupportedOp("sharded Jacobian needs a catalogue target")
        self._setup(ctx, xlen, m, self.cfg.N, group)
        torch = self.torch
        p = getattr(f, "params", {}) or {}
        self.alpha = float(p.get("alpha", 0.0))
        self.A_t = self.b_t = self.A_a = self.b_a = None
        if p.get("A") is not None:
            self.A_t = torch.from_numpy(np.asfortranarray(p["A"]).ravel(order="F").copy()).to(self.dev)
            self.b_t = torch.from_numpy(np.ascontiguousarray(p["b"], dtype=np.float64)).to(self.dev)
            self.A_a = ctx.wrap(self.A_t.data_ptr(), self.A_t.numel(), keepalive=self.A_t)
            self.b_a = ctx.wrap(self.b_t.data_ptr(), self.b_t.numel(), keepalive=self.b_t)
        self.y_t = torch.zeros(self.m, dtype=torch.float64, device=self.dev)
        self.y_a = ctx.wrap(self.y_t.data_ptr(), self.m, keepalive=self.y_t)

    def sweep(self):
        c = self.ctx
        c.check(c.lib.fdb_jacobian_chunked(c.h, self.fid, self.x_a.h, self.m, self.N, self.A_a.h if self.A_a else None, self.m,
                                           self.b_a.h if self.b_a else None, self.alpha, self.plan["chunk_lo"], self.plan["chunk_hi"],
                                           self.J_a.h, self.m, self.y_a.h))

    def values(self):
        """y = value.(ydual): taken from the rank owning the last chunk (src/jacobian.jl:229)."""
        import torch.distributed as dist
        if self.world > 1:
            dist.broadcast(self.y_t, src=owner_of_last_chunk(self.n, self.N, self.world), group=self.group)
        return self.y_t


class ShardedHessian(_ShardedColumns):
    """ForwardDiff.hessian!(result, f, x, cfg): outer chunks (column blocks of H) sharded over the ranks;
    every rank runs all inner chunks for its columns (src/hessian.jl:14-19)."""

    def __init__(self, ctx, f, xlen, cfg=None, group=None):
        self.f = f
        self.cfg = cfg if cfg is not None else HessianConfig(f, np.empty(int(xlen)))
        self.fid = _catalogue_id(f, _lib.FN)
        if self.fid is None:
            raise _lib.UnsupportedOp("sharded Hessian needs a catalogue target")
        self._setup(ctx, xlen, xlen, self.cfg.N, group)
        torch = self.torch
        self.g_t = torch.zeros(self.n, dtype=torch.float64, device=self.dev)
        self.v_t = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.g_a = ctx.wrap(self.g_t.data_ptr(), self.n, keepalive=self.g_t)
        self.v_a = ctx.wrap(self.v_t.data_ptr(), 1, keepalive=self.v_t)

    def sweep(self):
        c = self.ctx
        c.check(c.lib.fdb_hessian_chunked(c.h, self.fid, self.x_a.h, self.N, self.plan["chunk_lo"], self.plan["chunk_hi"],
                                          self.J_a.h, self.n, self.g_a.h, self.v_a.h))

    def value_and_gradient(self):
        """DiffResult value and gradient fall out of the last outer chunk's inner sweeps (src/hessian.jl:50-55)."""
        import torch.distributed as dist
        if self.world > 1:
            src = owner_of_last_chunk(self.n, self.N, self.world)
            dist.broadcast(self.g_t, src=src, group=self.group)
            dist.broadcast(self.v_t, src=src, group=self.group)
        return self.v_t, self.g_t


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
