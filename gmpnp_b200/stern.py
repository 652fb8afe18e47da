"""Drop-in for ``1D/Stern_CO2ER.py``: potential and field across the ion-free Stern layer between the outer Helmholtz
plane (OHP) and the electrode, from the OHP potential, OHP field and OHP permittivity that the 1D GMPNP solve
delivers (``edl1d`` metadata: ``field_OHP`` in V/nm, ``eps_rel_OHP``).

Two models (ST:88-165): ``BDM`` -- Poisson with a permittivity that varies linearly across the layer, integrated by the
reference with ``odeint`` on 40 points over L_stern = 0.4 nm -- and ``Stern_linear`` (constant field).  The BDM
equation ``E' = -E a / (a x + b)`` has the closed form ``E = E_0 b / (a x + b)``,
``phi = phi_0 + E_0 (b / a) ln((a x + b) / b)``, which is what is evaluated here (the unit test integrates the
reference's right-hand side with the same ``odeint`` call and agrees to 1e-9).

Reference behaviour kept by default (``as_executed=True``), both flagged in DESIGN.md:

* ``odeint(BDM, y0, x, args=(eps_rel_OHP, eps_rel_surface, L_stern))`` passes the two permittivities in the opposite
  order of ``def BDM(Y, x, eps_rel_surface, eps_rel_OHP, L_stern_scaled)`` (ST:91, 109), so the profile runs from
  6 at the OHP to eps_OHP at the electrode and the surface field comes out as ``E_OHP * 6 / eps_OHP`` instead of
  ``E_OHP * eps_OHP / 6``;
* the field is carried in V/nm while x runs in metres (ST:101-107), so the potential drop across the layer is
  scaled by 1e-9 and the "voltage at the electrode" equals the OHP voltage to nine digits.

``as_executed=False`` uses the argument order of the signature and integrates the potential over nanometres.
The five (field_OHP, eps_rel_OHP) pairs of ST:66-68 -- the only result values the reference holds -- are the defaults
of the CLI loop, exactly like the reference's ``main`` (ST:167-172), which ignores its own ``--field_OHP`` flags.
"""
from __future__ import annotations

import argparse
import math
import os
from datetime import datetime

import numpy as np

L_STERN = 4.0e-10                         # ST:57
EPS_REL_SURFACE = 6.0                     # ST:80
OHP_DICT = {-2.5: {"E": -0.08032108300135771, "eps": 74.56149297894756},           # ST:66-68
            -5.0: {"E": -0.2524415478848975, "eps": 57.64572780716129},
            -7.5: {"E": -0.4612956299192668, "eps": 50.16243860179017},
            -10.0: {"E": -0.6149631587776277, "eps": 49.311548142969336},
            -12.5: {"E": -0.7310301485096051, "eps": 49.2556833480052}}


def thermal_voltage(params_file="parameters", utilities_dir=None):
    from . import params as _params
    nc = _params._load_inputs(params_file, utilities_dir)["nat_const"]
    return (float(nc["k_B"]) * float(nc["T"])) / float(nc["e_0"])                  # ST:60


def stern_bdm(voltage_OHP, field_OHP, eps_rel_OHP, as_executed=True, L_stern=L_STERN, dx=1.0e-11):
    """Returns dict(x [m], sol [n, 2] = (potential, -field) as odeint returns them, x_scaled [nm], y1_scaled [V],
    y2_scaled [V/nm], voltage_electrode, field_surf)."""
    xmax = -L_stern                                                                # going backwards in length
    x = np.linspace(0, xmax, abs(int(xmax / dx)))                                  # ST:101-103
    if as_executed:
        a, b = EPS_REL_SURFACE - eps_rel_OHP, EPS_REL_SURFACE * L_stern            # the swapped arguments, ST:91, 109
    else:
        a, b = eps_rel_OHP - EPS_REL_SURFACE, eps_rel_OHP * L_stern
    y2_0 = -field_OHP                                                              # ST:105
    y2 = y2_0 * b / (a * x + b)
    integral = y2_0 * (b / a) * np.log((a * x + b) / b) if a != 0.0 else y2_0 * x
    if not as_executed:
        integral = integral * 1.0e+9                                               # field in V/nm, x in m
    y1 = voltage_OHP + integral
    sol = np.stack([y1, y2], axis=1)
    y1_scaled, y2_scaled = sol[:, 0], sol[:, 1] * -1                               # ST:112-113
    return dict(x=x, sol=sol, x_scaled=x * 1.0e+9, y1_scaled=y1_scaled, y2_scaled=y2_scaled,
                voltage_electrode=float(y1_scaled[-1]), field_surf=float(y2_scaled[-1]))


def stern_linear(voltage_OHP, field_OHP, L_stern=L_STERN):
    """ST:141-165."""
    y1_surf = voltage_OHP - (-field_OHP * (L_stern * 1.0e+9))
    dx = 1.0e-2
    xmax = -L_stern * 1.0e+9
    x = np.linspace(0, xmax, abs(int(xmax / dx)))
    y1_x = -field_OHP * x + voltage_OHP
    return dict(x=x, y1_x=y1_x, voltage_electrode=float(y1_surf), field_surf=float(field_OHP))


def write_metadata(f, model, voltage_OHP, field_OHP, L_stern, field_surf, eps_rel_OHP, voltage_electrode):
    """ST:31-43, line for line (including its units)."""
    f.write("model=" + model + "\n")
    f.write("voltage_OHP=" + str(voltage_OHP) + "V\n")
    f.write("field_OHP=" + str(field_OHP) + "V/nm\n")
    f.write(f"Relative permittivity at the OHP is {eps_rel_OHP} \n")
    f.write(f"voltage at the electrode is {voltage_electrode} \n")
    f.write(f"Electric field at the surface is {field_surf} m\n")
    f.write(f"Stern length is {L_stern} m\n")


def Stern(voltage_scaled_OHP=-1.0, field_OHP=-0.5, eps_rel_OHP=80.0, model="BDM", *, as_executed=True, out_dir=None,
          stamp=None, params_file="parameters", utilities_dir=None, write=True):
    """One Stern-layer evaluation (ST:70-165).  Returns the result dict (+ ``output_dir`` when written)."""
    voltage_OHP = voltage_scaled_OHP * thermal_voltage(params_file, utilities_dir)  # ST:79
    if model == "BDM":
        res = stern_bdm(voltage_OHP, field_OHP, eps_rel_OHP, as_executed)
    elif model == "Stern_linear":
        res = stern_linear(voltage_OHP, field_OHP)
    else:
        raise ValueError(model)
    res.update(model=model, voltage_OHP=voltage_OHP, field_OHP=field_OHP, eps_rel_OHP=eps_rel_OHP)
    if write:
        stamp = stamp or datetime.now().strftime("%y-%m-%d-%H-%M-%S")
        newpath = os.path.join(out_dir or os.path.join(os.getcwd(), "out"), stamp + "_experiment",
                               "voltage_scaled_OHP" + str(voltage_scaled_OHP))     # ST:74-77
        os.makedirs(newpath, exist_ok=True)
        v = str(voltage_scaled_OHP)
        if model == "BDM":
            np.savez(os.path.join(newpath, "stern_unscaled_BDM" + v + ".npz"), res["sol"])                   # ST:118
            np.savez(os.path.join(newpath, "stern_scaled_BDM" + v + ".npz"), res["x_scaled"], res["y1_scaled"],
                     res["y2_scaled"])                                                                        # ST:119
        else:
            np.savez(os.path.join(newpath, "stern_scaled_linear" + v + ".npz"), res["x"], res["y1_x"])       # ST:158
        with open(os.path.join(newpath, "metadata.txt"), "w") as f:
            write_metadata(f, model, voltage_OHP, field_OHP, L_STERN, res["field_surf"], eps_rel_OHP,
                           res["voltage_electrode"])
        res["output_dir"] = newpath
    return res


def from_edl_metadata(meta: dict, model="BDM", **kw):
    """Feed the Stern layer from a 1D GMPNP run (``edl1d.solve_EDL`` metadata): the self-made version of the table
    the reference pasted in by hand (ST:64-68)."""
    return Stern(voltage_scaled_OHP=float(meta["voltage_multiplier"]), field_OHP=float(meta["field_OHP"]),
                 eps_rel_OHP=float(meta["eps_rel_OHP"]), model=model, **kw)


def main(argv=None):
    p = argparse.ArgumentParser(description="experiment parameters")               # ST:175-193
    p.add_argument("--voltage_scaled_OHP", default=-2.5, type=float, help="float val")
    p.add_argument("--model", default="BDM", type=str, help="str, BDM/Stern_linear")
    p.add_argument("--field_OHP", default=-0.5, type=float, help="float val, -0.5")
    p.add_argument("--eps_rel_OHP", default=80.0, type=float, help="float, 80.0")
    p.add_argument("--single", action="store_true",
                   help="evaluate the flags' point instead of the reference's built-in table (ST:167-172 always loops "
                        "over its table)")
    p.add_argument("--edl_metadata", default=None, help="metadata.json of a gmpnp_b200.edl1d run to take the OHP values from")
    p.add_argument("--intended", action="store_true", help="signature argument order and consistent units")
    p.add_argument("--out_dir", default=None)
    a = p.parse_args(argv)
    stamp = datetime.now().strftime("%y-%m-%d-%H-%M-%S")
    kw = dict(model=a.model, as_executed=not a.intended, out_dir=a.out_dir, stamp=stamp)
    if a.edl_metadata:
        import json
        with open(a.edl_metadata) as f:
            results = [from_edl_metadata(json.load(f), **kw)]
    elif a.single:
        results = [Stern(a.voltage_scaled_OHP, a.field_OHP, a.eps_rel_OHP, **kw)]
    else:
        results = [Stern(v, d["E"], d["eps"], **kw) for v, d in OHP_DICT.items()]
    for r in results:
        print(r["voltage_OHP"], r["voltage_electrode"], r["field_surf"])


if __name__ == "__main__":
    main()
