"""Bulk electrolyte composition before / after CO2 saturation -- the pre-processing step of the reference
(``utilities/bulk_soln.py``), as a function and a CLI instead of an edit-and-run script.

The reference integrates the homogeneous carbonate kinetics with ``scipy.integrate.odeint`` (BS:21-31, 56-64, 121-131,
189-193) from the salt's nominal composition and writes ``bulk_soln_<conc><electrolyte>.yaml`` with a ``pre_CO2`` and a
``post_CO2`` block -- the files every solver script of the repository reads.  This module follows it step by step
(same ODE right-hand sides, same output times, the same integrator with its default tolerances) so that the
checked-in YAML files ARE its golden vectors (``tests/test_bulk_stern.py``):

* ``pre_tmax``: the checked-in script integrates the CO2-free stage for 10 s (BS:117), but the checked-in YAML files
  were produced with 1000 s (reproduced here to 2e-6 at 0.1 / 0.5 / 1.0 M; with 10 s the hydroxide is off by 20x).
  Default 1.0e+3 = what the files hold; ``pre_tmax=1.0e+1`` = the script as checked in;
* the Sechenov correction of the post-CO2 stage uses the PRE-CO2 ion concentrations (BS:57, 183-187: "has not been solved
  self consistently") while the reported ``C0_CO2`` is the ion-free Henry value (BS:206) -- kept;
* cations: the script hard-codes K (``h_ion_K``); other monovalent cations take their own Sechenov constant
  (``h_ion_Li`` / ``h_ion_Na``), or K's when the parameter file has none (Cs: SURVEY finding 5), and the block lists
  ``C0_<cation>`` for every cation the solvers can ask for, as the checked-in files do.

Host-side by design: four unknowns, run once per electrolyte -- nothing to accelerate.
"""
from __future__ import annotations

import argparse
import json
import math
import os

import numpy as np

CATIONS = ("K", "Li", "Cs", "Na")


def kinetics(y, t, ka1, ka2, kb1, kb2):
    """BS:21-31."""
    C_HCO3, C_OH, C_CO32, C_CO2 = y
    return [kb1 * C_CO2 * C_OH - kb2 * C_HCO3 - ka1 * C_HCO3 * C_OH + ka2 * C_CO32,
            ka2 * C_CO32 - ka1 * C_HCO3 * C_OH + kb2 * C_HCO3 - kb1 * C_CO2 * C_OH,
            ka1 * C_HCO3 * C_OH - ka2 * C_CO32,
            kb2 * C_HCO3 - kb1 * C_CO2 * C_OH]


def CO2_conc(temp, fugacity_CO2, ions, sechenov_const):
    """Sechenov-corrected CO2 solubility in mol/m3 (BS:33-54); ``ions``: name -> mol/m3."""
    h_CO2 = sechenov_const["h_CO2_0"] + sechenov_const["h_CO2_T"] * (temp - 298.15)
    lnK_H_CO2 = 93.4517 * (100 / temp) - 60.2409 + 23.3585 * math.log(temp / 100)
    sechenov = 0.0
    for ion in ions.keys():
        sechenov += (sechenov_const["h_ion_" + ion] + h_CO2) * (ions[ion] / 1000)
    return fugacity_CO2 * math.exp(lnK_H_CO2) * 1000 * 10 ** (-sechenov)


def initial_composition(conc, electrolyte):
    """BS:79-109 (mol/m3): cation, HCO3, OH, CO32, CO2, Cl."""
    c = conc * 1000
    table = {"KHCO3": (c, c, 1.0e-7 * 1000, 0.0, 0.0, 0.0),
             "KOH": (c, 0.0, c, 0.0, 0.0, 0.0),
             "K2CO3": (2 * c, 0.0, 1.0e-7 * 1000, c, 0.0, 0.0),
             "KCl": (c, 0.0, 1.0e-7 * 1000, 0.0, 0.0, c)}
    if electrolyte not in table:
        raise ValueError("Electrolyte type not yet supported. Sorry!")            # BS:108-109
    return table[electrolyte]


def bulk_solution(conc=0.1, electrolyte="KHCO3", cation="K", T=298.15, f_CO2=1, params_file="parameters",
                  utilities_dir=None, pre_tmax=1.0e+3, dt=1.0e-2):
    """Returns the dictionary the reference dumps to YAML: ``{'bulk_conc_pre_CO2': {...}, 'bulk_conc_post_CO2': {...}}``."""
    from scipy.integrate import odeint
    from . import params as _params
    data = _params._load_inputs(params_file, utilities_dir)
    rc = data["rate_constants"]
    ka1, ka2, kb1, kb2 = (float(rc[k]) for k in ("ka1", "ka2", "kb1", "kb2"))
    sc = {k: float(v) for k, v in data["sechonov_const"].items()}
    if "h_ion_" + cation not in sc:
        sc["h_ion_" + cation] = sc["h_ion_K"]
    C_cat, C_HCO3, C_OH, C_CO32, C_CO2, C_Cl = initial_composition(conc, electrolyte)

    def ions(hco3, oh, co32):
        return {cation: C_cat, "HCO3": hco3, "OH": oh, "CO32": co32, "Cl": C_Cl}

    def block(sol_last, co2, extra):
        pH = -math.log10(1.0e-14 / (sol_last[1] / 1000))                          # BS:126, 195
        concs = {"C0_H": (10 ** (-pH)) * 1000, "C0_OH": float(sol_last[1]), "C0_CO2": float(co2),
                 "C0_HCO3": float(sol_last[0]), "C0_CO32": float(sol_last[2]), "C0_Cl": C_Cl}
        for c in CATIONS:
            concs["C0_" + c] = C_cat
        return dict(conc_electrolyte=conc, electrolyte=electrolyte, final_pH=pH, concentrations=concs, **extra)

    t = np.linspace(0, pre_tmax, int(pre_tmax / dt))                              # BS:116-119
    sol = odeint(kinetics, [C_HCO3, C_OH, C_CO32, C_CO2], t, args=(ka1, ka2, kb1, kb2))
    pre = sol[-1]
    C_CO2_sechenov = CO2_conc(T, f_CO2, ions(pre[0], pre[1], pre[2]), sc)         # BS:133
    saturated = pre[3] > C_CO2_sechenov                                           # BS:149-150, 177-180
    out = {}
    if saturated:
        out["bulk_conc_pre_CO2"] = ("Concentrations before adding CO2 will be same as on adding CO2 since solution "
                                    "is already saturated")
        y0 = [C_HCO3, C_OH, C_CO32]
    else:
        out["bulk_conc_pre_CO2"] = block(pre, pre[3], {})
        y0 = [pre[0], pre[1], pre[2]]
    C0_CO2 = C_CO2_sechenov                                                       # BS:57: constant during stage 2

    def kinetics_const_CO2(y, t_):                                                # BS:56-64
        h, o, c = y
        return [kb1 * C0_CO2 * o - kb2 * h - ka1 * h * o + ka2 * c,
                ka2 * c - ka1 * h * o + kb2 * h - kb1 * C0_CO2 * o,
                ka1 * h * o - ka2 * c]

    tmax = 1.0e+3 if conc <= 1 else (1.0e+4 if conc <= 5 else 5.0e+4)             # BS:182-187
    t = np.linspace(0, tmax, int(tmax / dt))
    sol = odeint(kinetics_const_CO2, y0, t)
    out["bulk_conc_post_CO2"] = block(sol[-1], CO2_conc(T, f_CO2, {}, sc), {"CO2_pressure": f_CO2})   # BS:206
    return out


def write_yaml(data, path):
    import yaml
    with open(path, "w") as f:
        yaml.dump({"bulk_conc_pre_CO2": data["bulk_conc_pre_CO2"]}, f)            # two dumps, BS:171, 211
        yaml.dump({"bulk_conc_post_CO2": data["bulk_conc_post_CO2"]}, f)


def main(argv=None):
    p = argparse.ArgumentParser(description="bulk electrolyte composition (utilities/bulk_soln.py)")
    p.add_argument("--conc", default=0.1, type=float, help="electrolyte concentration in M")
    p.add_argument("--electrolyte", default="KHCO3", type=str, help="KHCO3 / KOH / K2CO3 / KCl")
    p.add_argument("--cation", default="K", type=str)
    p.add_argument("--T", default=298.15, type=float)
    p.add_argument("--f_CO2", default=1, type=float)
    p.add_argument("--pre_tmax", default=1.0e+3, type=float)
    p.add_argument("--params_file", default="parameters", type=str)
    p.add_argument("--utilities_dir", default=None)
    p.add_argument("--out_dir", default=None)
    a = p.parse_args(argv)
    data = bulk_solution(a.conc, a.electrolyte, a.cation, a.T, a.f_CO2, a.params_file, a.utilities_dir, a.pre_tmax)
    if a.out_dir:
        os.makedirs(a.out_dir, exist_ok=True)
        write_yaml(data, os.path.join(a.out_dir, "bulk_soln_" + str(a.conc) + a.electrolyte + ".yaml"))   # BS:147
    print(json.dumps(data["bulk_conc_post_CO2"]))


if __name__ == "__main__":
    main()
