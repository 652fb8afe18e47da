"""Quadrature rules FFC 2019.1 selects for the GMPNP forms.

The reference never names a rule: FFC estimates the polynomial degree of the summed
``dx`` integrand (residual: 3, Jacobian: 4 -- division adds numerator and denominator
degrees) and asks FIAT ``create_quadrature(cell, degree, "default")``:

* interval: Gauss-Jacobi(=Gauss-Legendre) with m = (degree+2)//2 points -> 2 (F), 3 (J);
* tetrahedron: FIAT ``quadrature_schemes._tetrahedron_scheme``: degree 3 -> the 5-point
  Zienkiewicz-Taylor rule (one negative weight), degree 4 -> the 14-point Keast rule.
  [upstream: FIAT is not vendored in the reference; tables restated from the published
  rules (Zienkiewicz & Taylor; Keast 1986, rule "KEAST5").]

Rules are returned as barycentric coordinates ``lam[q, nv]`` and weights that sum to 1
(multiply by the cell volume).
"""
import numpy as np


def interval_rule(npts: int):
    xg, wg = np.polynomial.legendre.leggauss(npts)
    xi = 0.5 * (xg + 1.0)
    lam = np.stack([1.0 - xi, xi], axis=1)
    return lam, 0.5 * wg


def _bary3(x):
    x = np.asarray(x, dtype=np.float64)
    return np.concatenate([1.0 - x.sum(axis=1, keepdims=True), x], axis=1)


def tet_rule_degree3():
    x = [[0.25, 0.25, 0.25],
         [0.5, 1.0 / 6.0, 1.0 / 6.0],
         [1.0 / 6.0, 0.5, 1.0 / 6.0],
         [1.0 / 6.0, 1.0 / 6.0, 0.5],
         [1.0 / 6.0, 1.0 / 6.0, 1.0 / 6.0]]
    w = np.array([-0.8, 0.45, 0.45, 0.45, 0.45])
    return _bary3(x), w


def tet_rule_degree4():
    a1, b1 = 0.6984197043243866, 0.1005267652252045
    a2, b2 = 0.0568813795204234, 0.3143728734931922
    x = [[0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0],
         [0.5, 0.0, 0.0], [0.0, 0.5, 0.0], [0.0, 0.0, 0.5],
         [a1, b1, b1], [b1, b1, b1], [b1, b1, a1], [b1, a1, b1],
         [a2, b2, b2], [b2, b2, b2], [b2, b2, a2], [b2, a2, b2]]
    w = np.array([0.0190476190476190] * 6 + [0.0885898247429807] * 4 + [0.1328387466855907] * 4)
    return _bary3(x), w


def rules_for_dim(dim: int):
    """(residual rule, Jacobian rule) for P1 GMPNP forms on an interval / tetrahedron."""
    if dim == 1:
        return interval_rule(2), interval_rule(3)
    if dim == 3:
        return tet_rule_degree3(), tet_rule_degree4()
    raise ValueError(dim)
