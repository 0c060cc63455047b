"""One host process driving several GPUs (fdb_multi_* of include/fdb200.h).

This is what a single Julia session calling `ForwardDiff.gradient!(out, f, x, cfg)`
(src/gradient.jl:35-44), `jacobian!` (src/jacobian.jl:57-87) or `hessian!`
(src/hessian.jl:31-37) binds to: the group owns one context per device, device i of n sweeps
the contiguous chunk range `fdb_plan_chunks(xlen, N, i, n)`, and

  * `gradient` / `jacobian` / `hessian` take host arrays and return the complete host result
    (each device copies its own slice; no inter-GPU traffic);
  * `ResidentArena` keeps inputs and results on the devices: the sweep kernels store every extracted
    entry into every peer's arena (linked in-process with cudaDeviceEnablePeerAccess, no CUDA IPC), so the
    complete result is resident on EVERY device on return.

The one-process-per-GPU form (torch.distributed, CUDA IPC) is fdb200.dist.
"""
import ctypes as C

import numpy as np

from . import _lib
from .api import ChunkSizeError, DimensionMismatch, _catalogue_id, chunk_plan
from .array import Context

_dp = C.POINTER(C.c_double)
_ERR = {1: _lib.FdbError, 2: DimensionMismatch, 3: _lib.CudaError, 4: MemoryError, 5: _lib.UnsupportedOp, 6: ChunkSizeError}


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class DeviceGroup:
    """fdb_multi_init: one context per device, all owned by this process.  `devices=None`: all visible devices."""

    def __init__(self, devices=None):
        self.lib = _lib.load()
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.fdb_multi_init(0, None, C.byref(h))
        else:
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.lib.fdb_multi_init(len(devices), devs, C.byref(h))
        if rc != 0:
            msg = self.lib.fdb_multi_last_error(None)
            raise _lib.CudaError((msg.decode() if msg else "fdb_multi_init failed") + f" (status {rc})")
        self.h = h
        self.size = int(self.lib.fdb_multi_size(h))
        # per-device contexts as borrowed Context objects (the group owns and shuts them down)
        self.contexts = []
        for i in range(self.size):
            c = Context.__new__(Context)
            c.lib, c.h, c._capturing, c._graph_handles = self.lib, C.c_void_p(self.lib.fdb_multi_ctx(h, i)), False, []
            c.device = int(devices[i]) if devices is not None else i
            c.close = lambda: None
            self.contexts.append(c)

    def check(self, rc):
        if rc != 0:
            msg = self.lib.fdb_multi_last_error(self.h)
            raise _ERR.get(rc, _lib.FdbError)(msg.decode() if msg else f"fdb200 error {rc}")

    def close(self):
        if getattr(self, "h", None):
            for c in self.contexts:
                c.h = None
            self.lib.fdb_multi_shutdown(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self.check(self.lib.fdb_multi_sync(self.h))

    def plan(self, xlen, N):
        """chunk range and result slice of every device (fdb_plan_chunks)"""
        return [chunk_plan(xlen, N, i, self.size) for i in range(self.size)]

    # ---- host arrays in, complete host result out ------------------------------------------------------
    def gradient(self, f, x, N, out=None):
        """gradient!(out, f, x, GradientConfig(f, x, Chunk{N}())) -> (out, f(x))"""
        fid = self._fid(f, _lib.FN)
        x = _f64(x).ravel()
        out = np.zeros(x.size) if out is None else out
        if out.size != x.size:
            raise DimensionMismatch(f"gradient result has {out.size} entries, x has {x.size}")
        val = C.c_double(0.0)
        self.check(self.lib.fdb_multi_gradient_host(self.h, fid, x.ctypes.data_as(_dp), x.size, int(N), out.ctypes.data_as(_dp), C.byref(val)))
        return out, val.value

    def jacobian(self, f, x, N, m=None, out=None, reuse_params=False):
        """jacobian!(out, f, x, cfg) / jacobian!(out, f!, y, x, cfg) for a catalogue target -> (J, y); J column-major m x n"""
        fid = self._fid(f, _lib.JFN)
        x = _f64(x).ravel()
        p = getattr(f, "params", {}) or {}
        A, b, alpha = p.get("A"), p.get("b"), float(p.get("alpha", 0.0))
        if m is None:
            m = np.shape(A)[0] if A is not None else f.output_length(x.size)
        Af = None if A is None else np.asfortranarray(A, dtype=np.float64)
        bf = None if b is None else _f64(b)
        out = np.zeros((m, x.size), order="F") if out is None else out
        if out.shape != (m, x.size) or not out.flags.f_contiguous:
            raise DimensionMismatch("jacobian!: result must be a column-major m x n float64 matrix")
        y = np.zeros(m)
        self.check(self.lib.fdb_multi_jacobian_host(self.h, fid, x.ctypes.data_as(_dp), x.size, m, int(N),
                                                    None if Af is None else Af.ctypes.data_as(_dp), m,
                                                    None if bf is None else bf.ctypes.data_as(_dp), alpha, int(reuse_params),
                                                    out.ctypes.data_as(_dp), m, y.ctypes.data_as(_dp)))
        return out, y

    def hessian(self, f, x, N, out=None):
        """hessian!(result, f, x, cfg) -> (H, gradient, value): outer chunks (column blocks) sharded over the devices"""
        fid = self._fid(f, _lib.FN)
        x = _f64(x).ravel()
        n = x.size
        out = np.zeros((n, n), order="F") if out is None else out
        if out.shape != (n, n) or not out.flags.f_contiguous:
            raise DimensionMismatch("hessian!: result must be a column-major n x n float64 matrix")
        g = np.zeros(n); val = C.c_double(0.0)
        self.check(self.lib.fdb_multi_hessian_host(self.h, fid, x.ctypes.data_as(_dp), n, int(N), out.ctypes.data_as(_dp), n,
                                                   g.ctypes.data_as(_dp), C.byref(val)))
        return out, g, val.value

    @staticmethod
    def _fid(f, table):
        fid = _catalogue_id(f, table)
        if fid is None:
            raise _lib.UnsupportedOp("the multi-device drivers take catalogue targets")
        return fid

    def arena(self, n_doubles):
        return ResidentArena(self, n_doubles)


class ResidentArena:
    """fdb_multi_exchange_create: one symmetric arena per device, linked in-process.  Offsets are in doubles and the
    same on every device.  Results of the `*_resident` drivers are complete on every device on return."""

    def __init__(self, group, n_doubles):
        self.g, self.n = group, int(n_doubles)
        self.xs = (C.c_void_p * group.size)()
        group.check(group.lib.fdb_multi_exchange_create(group.h, self.n, self.xs))

    def close(self):
        if self.xs is not None and self.g.h:
            self.g.lib.fdb_multi_exchange_destroy(self.g.h, self.xs)
        self.xs = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, offset, host):
        """replicate a host vector at `offset` of every device's arena"""
        host = _f64(host).ravel()
        self.g.check(self.g.lib.fdb_multi_upload(self.g.h, self.xs, int(offset), host.ctypes.data_as(_dp), host.size))

    def download(self, device_index, offset, count):
        out = np.empty(int(count))
        self.g.check(self.g.lib.fdb_multi_download(self.g.h, self.xs, int(device_index), int(offset), out.ctypes.data_as(_dp), out.size))
        return out

    def barrier(self):
        self.g.check(self.g.lib.fdb_multi_barrier(self.g.h, self.xs))

    def set_timeout(self, seconds):
        for x in self.xs:
            self.g.check(self.g.lib.fdb_exchange_set_timeout(x, float(seconds)))

    def gradient(self, f, x_off, xlen, N, grad_off, value_off=-1):
        self.g.check(self.g.lib.fdb_multi_gradient_resident(self.g.h, self.xs, DeviceGroup._fid(f, _lib.FN), int(x_off), int(xlen), int(N),
                                                            int(grad_off), int(value_off)))

    def jacobian(self, f, x_off, xlen, m, N, J_off, y_off=-1, A=None, b=None, lda=None):
        """A, b: one device array per device (allocated on group.contexts[i]) or None"""
        alpha = float((getattr(f, "params", {}) or {}).get("alpha", 0.0))
        Ah = (C.c_void_p * self.g.size)(*[a.h for a in A]) if A is not None else None
        bh = (C.c_void_p * self.g.size)(*[a.h for a in b]) if b is not None else None
        self.g.check(self.g.lib.fdb_multi_jacobian_resident(self.g.h, self.xs, DeviceGroup._fid(f, _lib.JFN), int(x_off), int(xlen), int(m), int(N),
                                                            Ah, int(m if lda is None else lda), bh, alpha, int(J_off), int(m), int(y_off)))

    def hessian(self, f, x_off, xlen, N, H_off, grad_off=-1, value_off=-1):
        self.g.check(self.g.lib.fdb_multi_hessian_resident(self.g.h, self.xs, DeviceGroup._fid(f, _lib.FN), int(x_off), int(xlen), int(N),
                                                           int(H_off), int(xlen), int(grad_off), int(value_off)))
