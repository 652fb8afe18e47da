"""ctypes binding of libgmpnp.so (the C-ABI declared in include/gmpnp.h).

There is no CPU fallback: if the CUDA library is missing the import of any solver
raises, loudly.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GMPNP_LIB", os.path.join(_HERE, "libgmpnp.so"))   # override: experiments only

NPAR = 64


class NewtonOpts(C.Structure):
    """Mirror of ``gmpnp_newton_opts`` (dolfin NewtonSolver parameters, 1D:357-364, 3D:789-798)."""
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("relax", C.c_double), ("xtol", C.c_double),
                ("maxit", C.c_int), ("criterion", C.c_int), ("pivot", C.c_int), ("lin_maxit", C.c_int),
                ("lin_restart", C.c_int), ("lin_rtol", C.c_double), ("xtol_path", C.c_double),
                ("jac_rule", C.c_int), ("partitions", C.c_int), ("xtol_floor", C.c_double)]

    @classmethod
    def reference_1d(cls):
        return cls(1e-4, 1e-4, 1.0, 1e-12, 50, 0, 1, 0, 0, 0.0, 0.0, 0, 2, 0.0)

    @classmethod
    def reference_3d(cls):
        return cls(1e-4, 1e-4, 0.9, 1e-12, 50, 0, 1, 2000, 100, 1e-10, 0.0, 0, 2, 0.0)

    @classmethod
    def sweep_3d(cls):
        """reference_3d with the linear-solver settings used for batched sweeps: GMRES(40) to 1e-8.  Measured on a
        batch of 64 config-3 problems: same Newton counts, solutions within 3e-11 of GMRES(100)/1e-10, 2.3x faster
        (the CGS2 orthogonalisation against a long basis dominates the iteration's HBM traffic)."""
        o = cls.reference_3d()
        o.lin_restart, o.lin_rtol = 40, 1e-8
        return o

    @classmethod
    def sweep_3d_inexact(cls, eta: float = 1e-4):
        """sweep_3d with a constant forcing term: GMRES stops at ||J dx - F|| <= eta ||F||.  The damped iteration of
        the reference (relaxation 0.9, 3D:796) contracts the error by 0.1 per iteration whatever the accuracy of the
        linear solve beyond ~1e-2, so the Newton counts do not change while GMRES needs about a third of the
        iterations (profiles/r01_inexact_newton_cpu_study.md; GPU-measured distance to the 1e-8 iterate:
        tests/test_gpu_3d.py::test_inexact_linear_solves_same_march, bench.py -> pore3d.inexact).  A throughput
        setting for sweeps, not the parity path.  Measured on a B200, 128 config-3 problems (tools/steady_time.py), as
        steady solves/s and worst per-field relative L2 distance to the GMRES(40)/1e-10 iterate:
        lin_rtol 1e-8: 96, 6.6e-11;  eta 1e-6: 146, 2.8e-9;  eta 1e-5: 198, 7.5e-8;  eta 1e-4: 269, 1.6e-6 (4.1e-8 in
        the max norm over all fields).  These are distances between STEADY states (fixed points, which do not depend on
        the linear-solve accuracy); the transient states of the time-accurate march are more sensitive: with eta 1e-6
        they differ from the oracle's six-step march by 1.6e-6 (test_forcing_term_1e_6_distance_to_the_oracle_march),
        because the reference's residual criterion (1e-4) stops Newton before a loose linear solve is corrected."""
        o = cls.sweep_3d()
        o.lin_rtol = eta
        return o

    @classmethod
    def steady(cls, xtol=1e-12, maxit=50, relax=1.0, xtol_path=0.0, jac_rule=0, xtol_floor=0.0):
        return cls(1e-4, 1e-4, relax, xtol, maxit, 1, 1, 2000, 100, 1e-12, xtol_path, jac_rule, 2, xtol_floor)


STATUS_NAMES = {0: "converged", 1: "maxit", 2: "not_finite", 3: "linear_failed", 4: "stagnated"}

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_pd, _pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
_po = C.POINTER(NewtonOpts)

# name -> (restype, argtypes); every symbol include/gmpnp.h declares
SIGNATURES = {
    "gmpnp_strerror": (C.c_char_p, [_i]),
    "gmpnp_last_cuda_error": (C.c_char_p, [_vp]),
    "gmpnp_version": (_i, []),
    "gmpnp_fp64_peak": (_i, [_i, _pd]),
    "gmpnp_destroy": (None, [_vp]),
    "gmpnp_launch_count": (C.c_longlong, [_vp]),
    "gmpnp_create_1d": (_i, [C.POINTER(_vp), _i, _pd, _i, _i, _i]),
    "gmpnp_set_params": (_i, [_vp, _pd, _i]),
    "gmpnp_assemble_1d": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "gmpnp_newton_1d": (_i, [_vp, _vp, _vp, _po, _vp, _vp, _vp, _vp, _vp]),
    "gmpnp_march_1d": (_i, [_vp, _vp, _vp, _i, _po, _vp, _vp, _vp, _vp, _vp]),
    "gmpnp_steady_continuation_1d": (_i, [_vp, _vp, _vp, _i, _po, _vp, _vp, _vp, _vp, _vp]),
    "gmpnp_field_1d": (_i, [_vp, _vp, _vp, _vp]),
    "gmpnp_field_ohp_1d": (_i, [_vp, _vp, _vp, _vp]),
    "gmpnp_create_3d": (_i, [C.POINTER(_vp), _i, _pd, _i, _pi, _i, _pi, _i, _i, _i]),
    "gmpnp_set_dirichlet_3d": (_i, [_vp, _pd, _i]),
    "gmpnp_pattern_3d": (_i, [_vp, _pi, _pi, _pi]),
    "gmpnp_assemble_3d": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "gmpnp_spmv_3d": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "gmpnp_newton_3d": (_i, [_vp, _vp, _vp, _po, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gmpnp_median_3d": (_i, [_vp, _vp, _i, _vp, _vp]),
    "gmpnp_set_march_data_3d": (_i, [_vp, C.POINTER(C.c_byte), _pd, _pd, _i]),
    "gmpnp_march_3d": (_i, [_vp, _vp, _vp, _i, _po, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gmpnp_steady_3d": (_i, [_vp, _vp, _vp, _po, _d, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _pi, _vp]),
    "gmpnp_spmv_rows_3d": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "gmpnp_vec_multi_dot": (_i, [_vp, _vp, C.c_longlong, _i, _vp, C.c_longlong, _vp, _vp]),
    "gmpnp_vec_lincomb": (_i, [_vp, _vp, C.c_longlong, _i, _vp, _d, _vp, _vp, C.c_longlong, _vp]),
    "gmpnp_bjacobi_setup_3d": (_i, [_vp, _vp, _vp]),
    "gmpnp_bjacobi_apply_3d": (_i, [_vp, _vp, _vp, _i, _vp]),
    "gmpnp_set_facet_terms_3d": (_i, [_vp, _pd, _pi, _pd, _i, _pd, _pd, _i]),
    "gmpnp_grad_project_3d": (_i, [_vp, _vp, _vp, _i, _vp]),
    "gmpnp_set_aggregates_3d": (_i, [_vp, _pi]),
    "gmpnp_coarse_accumulate_3d": (_i, [_vp, _vp, _i, _vp, _vp]),
    "gmpnp_coarse_invert_3d": (_i, [_vp, _vp, _vp]),
    "gmpnp_coarse_restrict_3d": (_i, [_vp, _vp, _i, _vp, _vp]),
    "gmpnp_coarse_prolong_3d": (_i, [_vp, _vp, _vp, _i, _vp]),
}

_lib = None


class GmpnpError(RuntimeError):
    pass


def load():
    """Load libgmpnp.so and declare the signatures.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GmpnpError(
            f"{LIB_PATH} not found: the CUDA library is not built (run __graft_entry__.build()). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, handle=None):
    if rc != 0:
        lib = load()
        msg = lib.gmpnp_strerror(rc).decode()
        if handle is not None and rc == -2:
            msg += ": " + lib.gmpnp_last_cuda_error(handle).decode()
        raise GmpnpError(f"libgmpnp error {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
