"""The C-ABI library loads and exports every symbol include/gmpnp.h declares (no compute)."""
import ctypes
import os
import re

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "gmpnp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gmpnp_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from gmpnp_b200 import _lib
    names = header_functions()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"libgmpnp.so does not export {n}"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.gmpnp_version() >= 100
    assert lib.gmpnp_strerror(0) == b"ok"
    assert b"argument" in lib.gmpnp_strerror(-1)


def test_struct_layout_matches_header():
    from gmpnp_b200 import _lib, params
    o = _lib.NewtonOpts.reference_3d()
    assert ctypes.sizeof(o) == 4 * 8 + 5 * 4 + 4 + 2 * 8 + 2 * 4 + 8   # 4 doubles, 5 ints (+pad), 2 doubles, 2 ints, 1 double
    assert _lib.NewtonOpts.xtol_floor.offset == 80 and o.xtol_floor == 0.0
    assert (o.rtol, o.atol, o.relax, o.maxit) == (1e-4, 1e-4, 0.9, 50)
    src = open(os.path.join(ROOT, "include", "gmpnp.h")).read()
    defs = dict(re.findall(r"#define\s+GMPNP_P_(\w+)\s+(\d+)", src))
    assert int(defs["NS"]) == params.P_NS and int(defs["Z"]) == params.P_Z and int(defs["NU"]) == params.P_NU
    assert int(defs["ZC0"]) == params.P_ZC0 and int(defs["S"]) == params.P_S and int(defs["KAPPA"]) == params.P_KAPPA
    assert int(defs["V"]) == params.P_V and int(defs["JFLUX"]) == params.P_JFLUX and int(defs["Q"]) == params.P_Q
    assert int(re.search(r"#define\s+GMPNP_NPAR\s+(\d+)", src).group(1)) == params.NPAR


def test_api_argument_errors_without_gpu(lib):
    # argument validation happens before any CUDA call
    h = ctypes.c_void_p()
    x = (ctypes.c_double * 3)(0.0, 0.5, 0.25)          # not sorted
    assert lib.gmpnp_create_1d(ctypes.byref(h), 0, x, 3, 6, 1) == -1
    assert lib.gmpnp_create_1d(ctypes.byref(h), 0, x, 1, 6, 1) == -1
    assert lib.gmpnp_set_params(None, None, 1) == -1
