"""Multi-GPU: one process per GPU, chunk sweeps sharded by contiguous chunk ranges.

The reference runs its ceil(n/N) chunk sweeps sequentially on one thread
(src/gradient.jl:141-147); they are independent (chunk c only produces gradient
entries / Jacobian columns c*N+1 .. min((c+1)*N, n)), so rank r of W owns chunks
[r*ceil(C/W), min((r+1)*ceil(C/W), C)) and therefore a contiguous slice of the
result.  There is no collective inside the sweep.  The exchange that makes the result
complete on every rank comes in two forms:

  exchange="p2p"  (default on GPUs that can map each other's memory): the results
      live in a symmetric arena (fdb_exchange_*), the ranks swap its CUDA IPC
      handle once, and the sweep kernels store every extracted entry into all
      peers' arenas over NVLink as they produce it -- the extraction IS the
      gather; what remains is a device-side flag barrier.
  exchange="nccl": one all-gather of the (padded, equal-sized) slices with
      torch.distributed after the sweeps (also what the gloo CPU tests exercise).

The value f(x) is taken from the rank that owns the last chunk, as the reference
does (src/gradient.jl:155).

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


class _CudaView:
    """__cuda_array_interface__ shim: lets torch address arena memory owned by the library (for D2H copies)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr="<f8", data=(int(ptr), False), version=3, strides=None)


class PeerExchange:
    """A symmetric result arena mapped by every peer (fdb_exchange_*).  The 64-byte IPC handles travel through
    the process group once at construction; afterwards no host-side collective is involved."""

    def __init__(self, ctx, n_doubles, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.ctx, self.group = ctx, group
        self.rank, self.world = _rank_world(group)
        self.n = int(n_doubles)
        h = C.c_void_p()
        ctx.check(ctx.lib.fdb_exchange_create(ctx.h, self.n, self.rank, self.world, C.byref(h)))
        self.h = h
        try:
            buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
            ctx.check(ctx.lib.fdb_exchange_handle(h, buf))
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(buf.raw), group=group)
                ok = ctx.lib.fdb_exchange_connect(h, b"".join(handles))
                # every rank must agree that the mapping worked before anyone stores into a peer
                flags = [None] * self.world
                dist.all_gather_object(flags, int(ok), group=group)
                if any(flags):
                    ctx.check(ok)
                    raise _lib.CudaError(f"peer arena mapping failed on rank(s) {[r for r, f in enumerate(flags) if f]}")
        except Exception:
            ctx.lib.fdb_exchange_destroy(h)
            self.h = None
            raise
        v = self.view(0, 0)
        self.base_ptr = v.ptr
        self._views = []

    def view(self, offset, n, ld=None):
        import ctypes as C
        from .array import B200Array
        ld = int(n if ld is None else ld)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.fdb_exchange_view(self.h, int(offset), int(n), ld, C.byref(h)))
        ptr = C.c_void_p()
        self.ctx.check(self.ctx.lib.fdb_array_info(h, None, None, None, None, C.byref(ptr)))
        a = B200Array.__new__(B200Array)
        a._init(self.ctx, h, int(n), 0, 0, ld, ptr.value or 0, keepalive=self)
        return a

    def torch_view(self, offset, n):
        import torch
        return torch.as_tensor(_CudaView(self.base_ptr + 8 * int(offset), n), device=torch.device("cuda", self.ctx.device))

    def enable(self):
        self.ctx.check(self.ctx.lib.fdb_exchange_enable(self.ctx.h, self.h))

    def disable(self):
        self.ctx.check(self.ctx.lib.fdb_exchange_enable(self.ctx.h, None))

    def push(self, offset, count):
        """Copy [offset, offset+count) of the local arena into every peer's arena (all-gather / broadcast building block)."""
        self.ctx.check(self.ctx.lib.fdb_exchange_push(self.h, int(offset), int(count)))

    def barrier(self):
        """Device-side barrier on the context's stream (no host synchronisation)."""
        self.ctx.check(self.ctx.lib.fdb_exchange_barrier(self.h))

    def timeouts(self):
        import ctypes as C
        c = C.c_int64()
        self.ctx.check(self.ctx.lib.fdb_exchange_timeouts(self.h, C.byref(c)))
        return c.value

    def close(self):
        """Collective: every rank must have stopped storing into the peers before an arena is unmapped."""
        if self.h is None:
            return
        import torch.distributed as dist
        self.ctx.sync()
        if self.world > 1:
            dist.barrier(group=self.group)
        self.ctx.lib.fdb_exchange_destroy(self.h)
        self.h = None


# "auto" picks the multicast exchange from this many ranks up (None: never); set from the measurements in DESIGN.md section 6
MULTICAST_AUTO_MIN_RANKS = None


class MulticastExchange(PeerExchange):
    """The same arena bound into an NVSwitch multicast object (fdb_exchange_create_multicast, csrc/fdb_multicast.cu): the
    drivers store each result entry ONCE through the multicast mapping and the switch replicates it into every rank's arena
    (egress per rank = its own slice instead of (W-1) x it).  The {pid, fd} handles travel through the process group; every
    rank must agree on each step before the next one (a device that was not added, or an arena that is not bound yet, would
    make the first multicast store fault)."""

    def __init__(self, ctx, n_doubles, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.ctx, self.group = ctx, group
        self.rank, self.world = _rank_world(group)
        self.n = int(n_doubles)
        self.h = None
        lib = ctx.lib

        def agree(rc, what):
            """all ranks learn whether the step worked everywhere; the local error (if any) is raised, else a summary"""
            flags = [None] * self.world
            dist.all_gather_object(flags, int(rc), group=group)
            if any(flags):
                if self.h is not None:
                    lib.fdb_exchange_destroy(self.h)
                    self.h = None
                ctx.check(rc)
                raise _lib.UnsupportedOp(f"multicast exchange: {what} failed on rank(s) {[r for r, f in enumerate(flags) if f]}")

        h = C.c_void_p()
        rc = lib.fdb_exchange_create_multicast(ctx.h, self.n, self.rank, self.world, C.byref(h))
        if rc == 0:
            self.h = h
        agree(rc, "arena allocation")
        buf = C.create_string_buffer(_lib.SHARE_HANDLE_BYTES)
        rc = lib.fdb_exchange_share_handle(self.h, buf)
        mcb = C.create_string_buffer(_lib.SHARE_HANDLE_BYTES)
        if rc == 0 and self.rank == 0:
            rc = lib.fdb_exchange_multicast_create(self.h, mcb)
        agree(rc, "handle export")
        handles = [None] * self.world
        dist.all_gather_object(handles, (bytes(buf.raw), bytes(mcb.raw)), group=group)
        rc = lib.fdb_exchange_connect_multicast(self.h, b"".join(a for a, _ in handles), handles[0][1])
        agree(rc, "peer mapping / cuMulticastAddDevice")          # doubles as the host barrier before binding
        rc = lib.fdb_exchange_multicast_bind(self.h)
        agree(rc, "cuMulticastBindMem")                            # ... and as the barrier before the first multicast store
        v = self.view(0, 0)
        self.base_ptr = v.ptr
        self._views = []

    def set_multicast(self, on):
        """on=False: keep the mapping but store once per peer again (comparison runs)."""
        self.ctx.check(self.ctx.lib.fdb_exchange_set_multicast(self.h, 1 if on else 0))


def multicast_supported(ctx):
    import ctypes as C
    v = C.c_int(0)
    ctx.check(ctx.lib.fdb_multicast_supported(ctx.h, C.byref(v)))
    return bool(v.value)


def _make_exchange(ctx, n_doubles, group, mode):
    """mode: "p2p" (per-peer stores; raise if the peers cannot be mapped), "multicast" (NVSwitch multicast stores; raise if
    unsupported), "nccl" (no arena), "auto" (multicast from four ranks up where the box supports it, else p2p, else nccl)."""
    if mode == "nccl":
        return None
    _, world = _rank_world(group)
    if mode == "multicast" or (mode == "auto" and MULTICAST_AUTO_MIN_RANKS is not None and world >= MULTICAST_AUTO_MIN_RANKS):
        import torch.distributed as dist
        # all ranks must take the same branch: decide on the AND of the per-rank capability
        sup = [None] * world
        if world > 1:
            dist.all_gather_object(sup, bool(multicast_supported(ctx)), group=group)
        else:
            sup = [multicast_supported(ctx)]
        if all(sup):
            try:
                return MulticastExchange(ctx, n_doubles, group)
            except _lib.FdbError as e:
                if mode == "multicast":
                    raise
                import sys
                print(f"fdb200.dist: multicast exchange unavailable ({e}); using per-peer stores", file=sys.stderr)
        elif mode == "multicast":
            raise _lib.UnsupportedOp("multicast exchange: not supported on every rank's device")
    try:
        return PeerExchange(ctx, n_doubles, group)
    except _lib.FdbError as e:
        if mode == "p2p":
            raise
        import sys
        print(f"fdb200.dist: peer-memory exchange unavailable ({e}); using the NCCL all-gather", file=sys.stderr)
        return None


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

    def _setup(self, ctx, n, m, N, group, exchange="auto", extra=0):
        """`extra` doubles are reserved behind the matrix in the arena for the driver's vector outputs (y, gradient, value)."""
        import torch
        self.torch = torch
        self.ctx, self.n, self.m, self.N, self.group = ctx, int(n), int(m), int(N), group
        self.rank, self.world = _rank_world(group)
        self.plan = chunk_plan(self.n, self.N, self.rank, self.world)
        self.lay = slice_layout(self.n, self.N, self.world)
        self.dev = torch.device("cuda", ctx.device)
        per = self.lay["per"]
        self.x_t = torch.empty(self.n, dtype=torch.float64, device=self.dev)
        self.x_a = ctx.wrap(self.x_t.data_ptr(), self.n, keepalive=self.x_t)
        # kernels, torch copies and NCCL on ONE stream.  torch's default stream has handle 0 = the legacy default stream;
        # Context.set_stream passes that as cudaStreamLegacy (NULL would mean the context's own non-blocking stream,
        # which does not order against torch's copies at all)
        ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        self.full = None
        self._mat = self.m * per * self.world                     # padded full matrix (whole chunks per rank)
        self.xch = _make_exchange(ctx, self._mat + extra, group, exchange) if self.world > 1 or exchange == "p2p" else None
        self.exchange = "nccl" if self.xch is None else ("multicast" if isinstance(self.xch, MulticastExchange) else "p2p")
        if self.xch is not None:
            # every rank holds the full matrix in its arena; the kernels fill the peers' copies of this rank's columns
            self.full = self.xch.torch_view(0, self._mat)
            self.local = self.full[self.m * self.plan["elem_lo"]: self.m * self.plan["elem_lo"] + self.m * per]
            self.J_a = self.xch.view(0, self.m * self.n)
        else:
            # only this rank's column block is backed; the kernels address columns absolutely, so shift the base
            self.local = torch.zeros(self.m * per, dtype=torch.float64, device=self.dev)
            shifted = self.local.data_ptr() - 8 * self.m * self.plan["elem_lo"]
            self.J_a = ctx.wrap(shifted, self.m * self.n, keepalive=self.local)

    def _vector(self, offset, n):
        """A vector output: arena view (p2p) or a torch tensor (nccl).  Returns (torch tensor, device array)."""
        if self.xch is not None:
            return self.xch.torch_view(self._mat + offset, n), self.xch.view(self._mat + offset, n)
        t = self.torch.zeros(n, dtype=self.torch.float64, device=self.dev)
        return t, self.ctx.wrap(t.data_ptr(), n, keepalive=t)

    def _sweep_begin(self):
        if self.xch is not None:
            self.xch.barrier()          # all peers have consumed the previous result before it is overwritten
            self.xch.enable()

    def _sweep_end(self):
        if self.xch is not None:
            self.xch.disable()

    def local_block(self):
        """This rank's column block as an (m, ncols) column-major numpy array."""
        ncols = self.plan["elem_hi"] - self.plan["elem_lo"]
        return self.local[: self.m * ncols].cpu().numpy().reshape((self.m, ncols), order="F")

    def gather(self):
        """The full matrix on every rank: a device-side barrier (p2p: the sweeps already stored into every peer)
        or one all-gather of the padded column blocks (nccl)."""
        import torch.distributed as dist
        if self.xch is not None:
            self.xch.barrier()                    # asynchronous; a barrier that times out poisons the context: the next
            return self.full[: self.m * self.n]   # ctx.sync() / download / full_matrix() raises instead of returning stale data
        per = self.lay["per"]
        if self.full is None:
            self.full = self.torch.empty(self.m * per * self.world, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_gather_into_tensor(self.full, self.local, group=self.group)
        else:
            self.full.copy_(self.local)
        return self.full[: self.m * self.n]

    def full_matrix(self):
        host = self.gather().cpu().numpy().reshape((self.m, self.n), order="F")
        self.ctx.sync()                           # raises if an exchange barrier gave up on a peer (the matrix would be stale)
        return host

    def close(self):
        if self.xch is not None:
            self.xch.close()
            self.xch = None


class ShardedJacobian(_ShardedColumns):
    """ForwardDiff.jacobian!(J, f, x, cfg) / jacobian!(J, f!, y, x, cfg) with chunk sweeps sharded over the ranks."""

    def __init__(self, ctx, f, xlen, m, cfg=None, group=None, exchange="auto"):
        self.f = f
        self.cfg = cfg if cfg is not None else JacobianConfig(f, np.empty(int(xlen)))
        self.fid = _catalogue_id(f, _lib.JFN)
        if self.fid is None:
            raise _lib.UnsupportedOp("sharded Jacobian needs a catalogue target")
        self._setup(ctx, xlen, m, self.cfg.N, group, exchange, extra=-(-int(m) // 32) * 32)
        torch = self.torch
        p = getattr(f, "params", {}) or {}
        self.alpha = float(p.get("alpha", 0.0))
        self.A_t = self.b_t = self.A_a = self.b_a = None
        if p.get("A") is not None:
            self.A_t = torch.from_numpy(np.asfortranarray(p["A"]).ravel(order="F").copy()).to(self.dev)
            self.b_t = torch.from_numpy(np.ascontiguousarray(p["b"], dtype=np.float64)).to(self.dev)
            self.A_a = ctx.wrap(self.A_t.data_ptr(), self.A_t.numel(), keepalive=self.A_t)
            self.b_a = ctx.wrap(self.b_t.data_ptr(), self.b_t.numel(), keepalive=self.b_t)
        self.y_t, self.y_a = self._vector(0, self.m)

    def sweep(self):
        c = self.ctx
        self._sweep_begin()
        try:
            c.check(c.lib.fdb_jacobian_chunked(c.h, self.fid, self.x_a.h, self.m, self.N, self.A_a.h if self.A_a else None, self.m,
                                               self.b_a.h if self.b_a else None, self.alpha, self.plan["chunk_lo"], self.plan["chunk_hi"],
                                               self.J_a.h, self.m, self.y_a.h))
        finally:
            self._sweep_end()

    def values(self):
        """y = value.(ydual): taken from the rank owning the last chunk (src/jacobian.jl:229); with the peer exchange
        that rank's kernel has already stored it everywhere (valid after gather())."""
        import torch.distributed as dist
        if self.world > 1 and self.xch is None:
            dist.broadcast(self.y_t, src=owner_of_last_chunk(self.n, self.N, self.world), group=self.group)
        return self.y_t


class ShardedHessian(_ShardedColumns):
    """ForwardDiff.hessian!(result, f, x, cfg): outer chunks (column blocks of H) sharded over the ranks;
    every rank runs all inner chunks for its columns (src/hessian.jl:14-19)."""

    def __init__(self, ctx, f, xlen, cfg=None, group=None, exchange="auto", peer_stores="auto"):
        """peer_stores: "mirror" = the sweep kernel stores every entry into all peers (one launch), "push" = local stores, then one
        streaming copy of the column block to the peers; "auto" = mirror for two ranks, push beyond (measured mirror / push at C5:
        0.110 / 0.140 ms at 2 ranks, 0.149 / 0.124 at 4, 0.173 / 0.112 at 8 -- see DESIGN.md section 6)."""
        self.f = f
        self.peer_stores = peer_stores
        self.cfg = cfg if cfg is not None else HessianConfig(f, np.empty(int(xlen)))
        self.fid = _catalogue_id(f, _lib.FN)
        if self.fid is None:
            raise _lib.UnsupportedOp("sharded Hessian needs a catalogue target")
        npad = -(-int(xlen) // 32) * 32
        self._npad = npad
        self._setup(ctx, xlen, xlen, self.cfg.N, group, exchange, extra=npad + 32)
        self.g_t, self.g_a = self._vector(0, self.n)
        self.v_t, self.v_a = self._vector(npad, 1)

    def sweep(self):
        """This rank's outer chunks.  With the peer exchange the column block is written locally and then PUSHED to the peers in
        one streaming kernel (fdb_exchange_push): the Hessian kernel's own stores are N(N+1) scattered 8-byte entries per block,
        which cross NVLink at a fraction of the link rate (measured at 8 ranks: 0.169 ms with the stores mirrored from the sweep
        kernel, the 29 MB push runs at the link rate); same bits either way."""
        c = self.ctx
        mode = self.peer_stores if self.peer_stores != "auto" else ("mirror" if self.world <= 2 else "push")
        if mode == "mirror":
            self._sweep_begin()
            try:
                c.check(c.lib.fdb_hessian_chunked(c.h, self.fid, self.x_a.h, self.N, self.plan["chunk_lo"], self.plan["chunk_hi"],
                                                  self.J_a.h, self.n, self.g_a.h, self.v_a.h))
            finally:
                self._sweep_end()
            return
        if self.xch is not None:
            self.xch.barrier()          # all peers have consumed the previous result before it is overwritten
        c.check(c.lib.fdb_hessian_chunked(c.h, self.fid, self.x_a.h, self.N, self.plan["chunk_lo"], self.plan["chunk_hi"],
                                          self.J_a.h, self.n, self.g_a.h, self.v_a.h))
        if self.xch is not None and self.world > 1:
            lo, hi = self.plan["elem_lo"], self.plan["elem_hi"]
            if hi > lo:
                self.xch.push(self.m * lo, self.m * (hi - lo))
            if self.rank == owner_of_last_chunk(self.n, self.N, self.world):       # value and gradient come from the last outer chunk
                self.xch.push(self._mat, self._npad + 1)

    def value_and_gradient(self):
        """DiffResult value and gradient fall out of the last outer chunk's inner sweeps (src/hessian.jl:50-55)."""
        import torch.distributed as dist
        if self.world > 1 and self.xch is None:
            src = owner_of_last_chunk(self.n, self.N, self.world)
            dist.broadcast(self.g_t, src=src, group=self.group)
            dist.broadcast(self.v_t, src=src, group=self.group)
        return self.v_t, self.g_t


class ShardedGradient:
    """Reusable state for ForwardDiff.gradient!(out, f, x, cfg) sharded over the process group:
    device buffers are allocated once (like a GradientConfig's work buffers)."""

    def __init__(self, ctx, f, xlen, cfg=None, group=None, exchange="auto"):
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
        self.x_a = ctx.wrap(self.x_t.data_ptr(), self.n, keepalive=self.x_t)
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)     # kernels, copies and NCCL on one stream (0 = legacy default stream)
        glen = -(-max(self.lay["padded"], self.n) // 32) * 32
        slot = glen + 32                                               # one result: gradient | value
        self.xch = _make_exchange(ctx, 2 * slot, group, exchange) if self.world > 1 or exchange == "p2p" else None
        self.exchange = "nccl" if self.xch is None else ("multicast" if isinstance(self.xch, MulticastExchange) else "p2p")
        if self.xch is not None:
            # Gradient and value live in the arena every peer maps: the sweep kernel stores each entry everywhere.
            # Two result slots used alternately: a peer overwrites slot s again only two sweeps later, after the
            # barrier that follows the sweep in between -- which this rank reaches only after it has consumed slot s.
            # So ONE barrier per sweep (after it) is enough; no barrier before the sweep.
            self._slots = [(self.xch.torch_view(b * slot, glen), self.xch.torch_view(b * slot + glen, 1),
                            self.xch.view(b * slot, self.n, ld=glen), self.xch.view(b * slot + glen, 1, ld=32)) for b in (0, 1)]
            self._cur = 1
            self.g_t, self.v_t, self.g_a, self.v_a = self._slots[self._cur]
        else:
            # gathered buffer: rank r's slice sits at [r*per, (r+1)*per) == its absolute gradient positions
            self.g_t = torch.zeros(glen, dtype=torch.float64, device=dev)
            self.v_t = torch.zeros(1, dtype=torch.float64, device=dev)
            self.g_a = ctx.wrap(self.g_t.data_ptr(), self.n, ld=self.g_t.numel(), keepalive=self.g_t)
            self.v_a = ctx.wrap(self.v_t.data_ptr(), 1, keepalive=self.v_t)

    def sweep(self):
        """This rank's chunk sweeps (device-resident inputs); no host sync.  With the peer exchange the result of
        this sweep is in `g_t` / `v_t` (which alternate between the two arena slots) after `gather()`."""
        c = self.ctx
        if self.xch is not None:
            self._cur ^= 1
            self.g_t, self.v_t, self.g_a, self.v_a = self._slots[self._cur]
            self.xch.enable()
        try:
            c.check(c.lib.fdb_gradient_chunked(c.h, self.fid, self.x_a.h, self.N, self.plan["chunk_lo"], self.plan["chunk_hi"],
                                               self.g_a.h, self.v_a.h))
        finally:
            if self.xch is not None:
                self.xch.disable()

    def gather(self):
        """Complete gradient and value on every rank: a device-side barrier (p2p; the sweep kernel already stored
        into every peer), or an all-gather of the slices plus a broadcast of the value from the owner of the last
        chunk (nccl)."""
        import torch.distributed as dist
        if self.xch is not None:
            self.xch.barrier()
            return
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
            self.ctx.sync()                       # raises if an exchange barrier gave up on a peer (the result would be stale)
            return out_host_pinned
        return self.g_t[:self.n]

    def close(self):
        if self.xch is not None:
            self.xch.close()
            self.xch = None
