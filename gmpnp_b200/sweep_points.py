"""Sweep-point enumeration of BASELINE config 2, static sharding and continuation paths (NumPy only: the CPU arm
of bench.py imports this module in its worker processes, which must not pay for ``import torch``)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CONFIG2_CATIONS = ("K", "Cs")
CONFIG2_CONCS = (0.1, 0.5, 1.0)
CONFIG2_LN = (1e-6, 5e-6, 10e-6, 50e-6, 200e-6)
CONFIG2_NV = 256
CONFIG2_VMAX = -12.5


@dataclass
class SweepPoint:
    cation: str
    conc: float
    L_n: float
    V: float
    index: int = 0


def config2_points(n_voltages: int = CONFIG2_NV, meshes=CONFIG2_LN, concs=CONFIG2_CONCS,
                   cations=CONFIG2_CATIONS, vmax: float = CONFIG2_VMAX):
    """{K, Cs} x n_voltages (V_k = vmax (k+1)/n) x {0.1, 0.5, 1.0} M x 5 meshes (SURVEY 8d cfg 2)."""
    pts = []
    for L_n in meshes:
        for k in range(n_voltages):
            V = vmax * (k + 1) / n_voltages
            for conc in concs:
                for cat in cations:
                    pts.append(SweepPoint(cat, conc, L_n, V, len(pts)))
    return pts


def shard(points, rank: int, world: int):
    """Static shard by sweep point: every rank gets the same mix of meshes and voltages."""
    return points[rank::world]


def voltage_paths(Vs: np.ndarray, dv_max: float) -> np.ndarray:
    """Ragged continuation paths, NaN-terminated: point b walks 0 -> V_b in ceil(|V_b|/dv_max) equal steps."""
    Vs = np.asarray(Vs, dtype=np.float64)
    nst = np.maximum(1, np.ceil(np.abs(Vs) / dv_max - 1e-12).astype(int))
    nV = int(nst.max())
    path = np.full((len(Vs), nV), np.nan)
    for b, (V, n) in enumerate(zip(Vs, nst)):
        path[b, :n] = V * np.arange(1, n + 1) / n
    return path
