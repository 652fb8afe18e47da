"""Host-side mirror of the 1D hot path: batched planar-EDL problems on one GPU.

One :class:`Solver1D` = one mesh + ``batch`` independent problems (sweep points) living on
one device.  The methods map one-to-one on the C-ABI (include/gmpnp.h) and on the reference
call sites they replace:

* ``assemble``  -> FFC kernels + SystemAssembler for the forms 1D/MPNP_CO2ER_EDL.py:381-595
* ``newton``    -> ``solve(F + J_OH*v_OH*ds + J_H*v_H*ds == 0, u, bcs, ...)`` 1D:737-742
* ``march``     -> the pseudo-time loop 1D:633-796 (incl. the H_OHP ladder 1D:766-793)
* ``steady``    -> the new steady solve with voltage continuation (BASELINE.json north_star)
* ``field``     -> ``project(-grad(u_np), W)`` 1D:802-803

PyTorch is used for device buffers and streams only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import NewtonOpts, check, ptr
from . import params as _params

NC = 7


def bulk_state(batch: int, n: int, device) -> torch.Tensor:
    """u = (1,...,1,0) everywhere: the reference's initial u_n (1D:322-326)."""
    u = torch.ones(batch, n, NC, dtype=torch.float64, device=device)
    u[:, :, NC - 1] = 0.0
    return u


def pack_1d(p) -> np.ndarray:
    """ProblemParams -> packed record, including the H_OHP controller slots (1D:766-793)."""
    P = p.pack()
    e = p.extras
    P[51] = e.get("J_OH_prefactor", 0.0) * e.get("current_OHP_ss", 0.0)
    P[52] = e.get("J_H_prefactor", 0.0) * e.get("current_OHP_ss", 0.0)
    H = e.get("H_OHP", None)
    P[53] = -1.0 if H is None else float(H)
    P[54] = e.get("current_H_frac", 0.0)
    return P


class Solver1D:
    def __init__(self, x: np.ndarray, batch: int, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.GmpnpError("CUDA device required: the GMPNP hot path has no CPU fallback")
        self.device = torch.device("cuda", int(device))
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
        self.x = x
        self.n = int(x.shape[0])
        self.batch = int(batch)
        self._h = C.c_void_p()
        check(self.lib.gmpnp_create_1d(C.byref(self._h), self.device.index, x.ctypes.data_as(C.POINTER(C.c_double)),
                                       self.n, 6, self.batch), self._h)
        self.packed = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.gmpnp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def set_params(self, plist):
        """``plist``: list of ProblemParams (len batch) or a packed [batch, NPAR] array."""
        if isinstance(plist, np.ndarray):
            P = np.ascontiguousarray(plist, dtype=np.float64)
        else:
            P = np.stack([pack_1d(p) for p in plist])
        assert P.shape == (self.batch, _params.NPAR), P.shape
        self.packed = P
        check(self.lib.gmpnp_set_params(self._h, P.ctypes.data_as(C.POINTER(C.c_double)), self.batch), self._h)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, t, shape):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous(), "need contiguous cuda float64"
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))

    def assemble(self, u, un, want_F=True, want_J=True):
        self._chk(u, (self.batch, self.n, NC))
        self._chk(un, (self.batch, self.n, NC))
        F = torch.empty(self.batch, self.n, NC, dtype=torch.float64, device=self.device) if want_F else None
        J = torch.empty(self.batch, self.n, 3, NC, NC, dtype=torch.float64, device=self.device) if want_J else None
        check(self.lib.gmpnp_assemble_1d(self._h, ptr(u), ptr(un), ptr(F), ptr(J), self._stream()), self._h)
        return F, J

    def newton(self, u, un, opts: NewtonOpts | None = None):
        """In-place Newton solve of every problem; returns dict(iters, r0, r, status) tensors."""
        opts = opts or NewtonOpts.reference_1d()
        self._chk(u, (self.batch, self.n, NC))
        self._chk(un, (self.batch, self.n, NC))
        it = torch.zeros(self.batch, dtype=torch.int32, device=self.device)
        st = torch.zeros(self.batch, dtype=torch.int32, device=self.device)
        r0 = torch.zeros(self.batch, dtype=torch.float64, device=self.device)
        r = torch.zeros(self.batch, dtype=torch.float64, device=self.device)
        check(self.lib.gmpnp_newton_1d(self._h, ptr(u), ptr(un), C.byref(opts), ptr(it), ptr(r0), ptr(r), ptr(st),
                                       self._stream()), self._h)
        return dict(iters=it, r0=r0, r=r, status=st)

    def march(self, u, un, n_steps: int, opts: NewtonOpts | None = None, history=False):
        opts = opts or NewtonOpts.reference_1d()
        self._chk(u, (self.batch, self.n, NC))
        self._chk(un, (self.batch, self.n, NC))
        hist = (torch.zeros(self.batch, n_steps, self.n, NC, dtype=torch.float64, device=self.device)
                if history else None)
        it = torch.zeros(self.batch, n_steps, dtype=torch.int32, device=self.device)
        st = torch.zeros(self.batch, dtype=torch.int32, device=self.device)
        hf = torch.zeros(self.batch, dtype=torch.float64, device=self.device)
        check(self.lib.gmpnp_march_1d(self._h, ptr(u), ptr(un), int(n_steps), C.byref(opts), ptr(hist), ptr(it),
                                      ptr(hf), ptr(st), self._stream()), self._h)
        return dict(iters=it, status=st, hfrac=hf, history=hist)

    def steady(self, u, Vpath, opts: NewtonOpts | None = None):
        """Steady equations with voltage continuation.  ``Vpath``: [batch, n_V] (tensor/array)."""
        opts = opts or NewtonOpts.steady()
        self._chk(u, (self.batch, self.n, NC))
        if not torch.is_tensor(Vpath):
            Vpath = torch.as_tensor(np.asarray(Vpath, dtype=np.float64))
        Vpath = Vpath.to(self.device, torch.float64).contiguous()
        assert Vpath.shape[0] == self.batch and Vpath.dim() == 2
        nV = int(Vpath.shape[1])
        it = torch.zeros(self.batch, nV, dtype=torch.int32, device=self.device)
        st = torch.zeros(self.batch, dtype=torch.int32, device=self.device)
        sg = torch.zeros(self.batch, dtype=torch.int32, device=self.device)
        dx = torch.zeros(self.batch, dtype=torch.float64, device=self.device)
        check(self.lib.gmpnp_steady_continuation_1d(self._h, ptr(u), ptr(Vpath), nV, C.byref(opts), ptr(it), ptr(sg),
                                                    ptr(st), ptr(dx), self._stream()), self._h)
        return dict(iters=it, status=st, stages=sg, dx=dx)

    def field(self, u):
        self._chk(u, (self.batch, self.n, NC))
        f = torch.empty(self.batch, self.n, dtype=torch.float64, device=self.device)
        check(self.lib.gmpnp_field_1d(self._h, ptr(u), ptr(f), self._stream()), self._h)
        return f

    def field_ohp(self, u):
        """The projected field at the OHP (node 0 of :meth:`field`) only: [batch]."""
        self._chk(u, (self.batch, self.n, NC))
        f = torch.empty(self.batch, dtype=torch.float64, device=self.device)
        check(self.lib.gmpnp_field_ohp_1d(self._h, ptr(u), ptr(f), self._stream()), self._h)
        return f

    def launch_count(self) -> int:
        return int(self.lib.gmpnp_launch_count(self._h))
