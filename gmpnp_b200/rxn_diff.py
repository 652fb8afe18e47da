"""Drop-in for ``1D/rxn_diff_planar.py`` (the reaction-diffusion comparison model of the reference: five carbonate
species, no potential, no steric term) on the SAME CUDA kernels as the GMPNP path ("a strict subset
of the hot-path kernels: nu = 0, z = 0, no phi").

Mapping onto the 7-component 1D kernel: charges and steric volumes are zero, so the potential decouples (it stays at
its Dirichlet values 0) and the sixth species is an inert passenger that starts and stays at its bulk value; the rows
of both are exactly zero in every residual, so Newton counts and iterates are those of the 5-species system.
Reference details kept: ``time_constant = L_n^2 / D_CO32`` and the time term ``(u - u_n)/del_t`` (RD1:152, 299-313),
fluxes ``J_OH v_OH ds + J_CO2 v_CO2 ds`` (RD1:260-261, 314), Dirichlet bulk values at x = 1 (RD1:255), Newton
``maximum_iterations 100, rtol = atol = 1e-6`` (RD1:331-339), ``time_step = 2e-2 s`` for 10 s (RD1:200-201).
FFC integrates this Jacobian with the residual's own 2-point rule (no rational term: degree 3), i.e. ``jac_rule = 1``.

    python -m gmpnp_b200.rxn_diff --L_n 50e-6 --n_steps 20
"""
from __future__ import annotations

import argparse
import json
import math
import os
from datetime import datetime

import numpy as np


def params_rxn_diff(concentration_KHCO3=0.1, H2_FE=0.2, L_n=50.0e-6, current_OHP_ss=10.0, cation="K",
                    params_file="parameters", utilities_dir=None, time_step=2.0e-2):
    from . import params as _params
    p = _params.params_1d(concentration_elec=concentration_KHCO3, model="PNP", voltage_multiplier=0.0, H2_FE=H2_FE,
                          current_OHP_ss=current_OHP_ss, L_n=L_n, cation=cation, params_file=params_file,
                          utilities_dir=utilities_dir)
    time_constant = L_n ** 2 / p.D[3]                      # RD1:152 (smallest diffusion coefficient: CO32)
    dt = time_step / time_constant                         # RD1:204
    jflux = np.zeros(6)
    jflux[1] = p.extras["J_OH_prefactor"] * current_OHP_ss * (-1.0)                  # RD1:261
    jflux[4] = p.extras["J_CO2_prefactor"] * current_OHP_ss * 0.5 * (1 - H2_FE)      # RD1:260
    return p.with_(z=np.zeros(6), nu=np.zeros(6), kappa=1.0 / dt, V=0.0, jflux=jflux, time_constant=time_constant,
                   dt_scaled=dt)


def solve_rxn_diff(concentration_KHCO3=0.1, H2_FE=0.2, L_n=50.0e-6, mesh_structure="variable", current_OHP_ss=10.0,
                   cation="K", params_file="parameters", *, utilities_dir=None, out_dir=None, n_steps=None, device=0,
                   write=True):
    import torch
    from . import meshio, params as _params, solver1d
    from ._lib import NewtonOpts
    stamp = datetime.now().strftime("%y-%m-%d-%H-%M-%S")
    total_sim_time, time_step = 10, 2.0e-2                 # RD1:200-201
    prm = params_rxn_diff(concentration_KHCO3, H2_FE, L_n, current_OHP_ss, cation, params_file, utilities_dir, time_step)
    if mesh_structure != "variable":
        raise NotImplementedError("the uniform 1D mesh file is absent from the reference (SURVEY App. E)")
    mesh = meshio.load_mesh(_params.mesh_name_1d(L_n), utilities_dir)
    x = mesh.x[:, 0]
    T = total_sim_time / prm.time_constant
    num_steps = int(T / prm.dt_scaled) if n_steps is None else int(n_steps)
    s = solver1d.Solver1D(x, batch=1, device=device)
    s.set_params([prm])
    dev = s.device
    u = torch.zeros(1, s.n, 7, dtype=torch.float64, device=dev)           # u = Function(V), RD1:231
    u[:, :, 5] = 1.0                                                      # the passenger species sits at its bulk value
    un = solver1d.bulk_state(1, s.n, dev)                                 # u_n = project(u_0), RD1:233-235
    o = NewtonOpts.reference_1d()
    o.rtol, o.atol, o.maxit, o.jac_rule = 1.0e-6, 1.0e-6, 100, 1          # RD1:331-339
    out = s.march(u, un, num_steps, o, history=True)
    if int(out["status"][0]) != 0:
        raise RuntimeError("Newton solver did not converge")              # what dolfin raises
    hist = out["history"][0].cpu().numpy()                                # [steps, n, 7]
    names = ["H", "OH", "HCO3", "CO32", "CO2"]
    arrays = {nm: np.vstack((np.ones(s.n), hist[:, :, i])) for i, nm in enumerate(names)}     # RD1:316-364
    tau_array = np.linspace(0, T, num_steps)                                                  # RD1:371
    c0 = dict(zip(prm.species, prm.c0))
    D = dict(zip(prm.species, prm.D))
    scaled = {}
    for nm in names:
        scaled["t_" + nm] = (tau_array * L_n ** 2) / D[nm]                                    # RD1:50-64
        scaled["c_" + nm] = arrays[nm] * c0[nm]
    scaled["c_cat"] = scaled["c_HCO3"] + 2 * scaled["c_CO32"] + scaled["c_OH"] - scaled["c_H"]   # RD1:421
    pH_OHP = -math.log10(scaled["c_H"][-1][0] / 1000)
    CO2_surf = scaled["c_CO2"][-1][0]
    bulk_pH = prm.extras["bulk_pH"]
    meta = {"concentration_KHCO3": concentration_KHCO3, "L_n": L_n, "bulk_pH": bulk_pH,
            "time_constant": prm.time_constant, "total_sim_time": total_sim_time, "time_step": time_step,
            "mesh_structure": "variable_" + str(int(L_n * 1.0e+6)) + "um", "H2_FE": H2_FE, "CO_FE": 1 - H2_FE,
            "current_OHP_ss": current_OHP_ss, "pH_OHP": pH_OHP,
            "pH_overpotential": -0.059 * (bulk_pH - pH_OHP) * 1.0e+3,
            "CO2_overpotential": (0.059 / 2) * math.log10(c0["CO2"] / CO2_surf) * 1.0e+3,
            "CO2_OHP_frac": CO2_surf / c0["CO2"], "newton_iterations": out["iters"][0].tolist()}
    if write:
        identifier = "H2_FE_" + str(H2_FE) + "_current_" + str(current_OHP_ss) + "_L_n_" + str(L_n) + "_cation_" + cation
        newpath = os.path.join(out_dir or os.path.join(os.getcwd(), "out"), stamp + "_experiment", identifier)
        os.makedirs(newpath, exist_ok=True)
        np.savez(os.path.join(newpath, "arrays_unscaled.npz"), coor_array=mesh.x, tau_array=tau_array, **arrays)
        np.savez(os.path.join(newpath, "arrays_scaled.npz"), x=mesh.x * L_n, **scaled)
        with open(os.path.join(newpath, "metadata.json"), "w") as f:
            f.write(json.dumps(meta, indent=0))
        meta["output_dir"] = newpath
    s.close()
    return meta


def main(argv=None):
    p = argparse.ArgumentParser(description="experiment parameters")          # RD1:496-548
    p.add_argument("--concentration_KHCO3", default=0.1, type=float, help="float val, 0.1 M")
    p.add_argument("--mesh_structure", default="variable", type=str)
    p.add_argument("--H2_FE", default=0.2, type=float)
    p.add_argument("--L_n", default=50.0e-6, type=float)
    p.add_argument("--current_OHP_ss", default=10.0, type=float)
    p.add_argument("--params_file", default="parameters", type=str)
    p.add_argument("--utilities_dir", default=None)
    p.add_argument("--out_dir", default=None)
    p.add_argument("--n_steps", default=None, type=int)
    a = p.parse_args(argv)
    meta = solve_rxn_diff(**vars(a))
    print(json.dumps({k: meta[k] for k in ("pH_OHP", "CO2_OHP_frac", "output_dir")}))


if __name__ == "__main__":
    main()
