"""Drop-in for ``3D/rxn_diff_CO2ER_pore.py`` (the reaction-diffusion comparison model of the reference in the pore:
seven species H, OH, HCO3, CO32, CO2, CO, H2; no potential, no cation, no steric term; wall fluxes and pore-exit Robin
terms LIVE here, RD3:513-548) on the SAME CUDA kernels as the GMPNP pore path.

Mapping onto the 9-component kernels: charges and steric volumes are zero and the wall voltage is 0, so the potential
stays at its Dirichlet value 0 and the cation is an inert passenger that starts and stays at its bulk value (its
wall flux and exit coefficient are zero); the rows of both are exactly zero in every residual and decouple in the
Jacobian, so Newton counts and iterates are those of the 7-species system.  All dx integrands are polynomials of
degree <= 3, for which both quadrature rules of the GMPNP kernels (degree 3 / degree 4) are exact -- the rule pair
does not matter here.  Reference details kept: same dimensionless groups as the MPNP script (RD3:115-324 equals
3D:115-324 without the electrostatics), gases pinned at the pore entry (RD3:408-412), ``relaxation_parameter 0.9``
(RD3:563-571), Sechenov update from the medians with the ELECTRONEUTRAL cation estimate (RD3:575-601), gradients of
the final state (RD3:634-653), output keys of RD3:659-676 / 742-766 / 770-790.

    python -m gmpnp_b200.rxn_diff3d --L 50e-9 --R 5e-9 --n_steps 5
"""
from __future__ import annotations

import argparse
import json
import os
from datetime import datetime

import numpy as np

NAMES = ["H", "OH", "HCO3", "CO32", "CO2", "CO", "H2"]


def params_rxn_diff_3d(concentration_elec=1.0, H2_FE=0.05, current_rough=3000.0, L=100.0e-9, cation="K", R=5.0e-9,
                       press_gas=1.0, pore_geom_multiplier=1.0, porosity_eff=0.5, tortuosity_eff=1.5,
                       constrictivity_eff=0.9, params_file="parameters_pore", y_CO2=0.95,
                       electrolyte_flow_geom_multiplier=1.0, roughness_factor=150.0, utilities_dir=None):
    """RD3:115-324 computes the groups of 3D:115-324; the electrostatics is switched off on top of them."""
    from . import params as _params
    p = _params.params_3d(concentration_elec=concentration_elec, voltage_multiplier=0.0, H2_FE=H2_FE,
                          current_rough=current_rough, L=L, cation=cation, R=R, press_gas=press_gas,
                          pore_geom_multiplier=pore_geom_multiplier, porosity_eff=porosity_eff,
                          tortuosity_eff=tortuosity_eff, constrictivity_eff=constrictivity_eff,
                          params_file=params_file, y_CO2=y_CO2,
                          electrolyte_flow_geom_multiplier=electrolyte_flow_geom_multiplier,
                          roughness_factor=roughness_factor, utilities_dir=utilities_dir)
    ex = dict(p.extras)
    kx = np.array(ex["k_exit"], dtype=np.float64)
    kx[7] = 0.0                                            # no cation equation in the reference model
    ex["k_exit"] = kx
    ex["current_planar"] = current_rough / roughness_factor
    return p.with_(z=np.zeros(8), nu=np.zeros(8), V=0.0, extras=ex)


def solveEDL(concentration_elec=1.0, H2_FE=0.05, current_rough=3000.0, L=100.0e-9, cation="K", R=5.0e-9,
             press_gas=1.0, pore_geom_multiplier=1.0, porosity_eff=0.5, tortuosity_eff=1.5, constrictivity_eff=0.9,
             params_file="parameters_pore", y_CO2=0.95, electrolyte_flow_geom_multiplier=1.0, roughness_factor=150.0,
             *, utilities_dir=None, out_dir=None, n_steps=None, mesh_file=None, device=0, write=True, pvd=True):
    """Signature of RD3:96-111 plus the keyword-only additions of ``gmpnp_b200.pore3d.solveEDL``."""
    from . import meshio, params as _params, solver3d
    from .pore3d import scale_conc_time

    stamp = datetime.now().strftime("%y-%m-%d-%H-%M-%S")
    prm = params_rxn_diff_3d(concentration_elec, H2_FE, current_rough, L, cation, R, press_gas, pore_geom_multiplier,
                             porosity_eff, tortuosity_eff, constrictivity_eff, params_file, y_CO2,
                             electrolyte_flow_geom_multiplier, roughness_factor, utilities_dir)
    mesh = meshio.load_mesh(mesh_file or _params.mesh_name_3d(L, R), utilities_dir)       # RD3:296-299
    time_step, total_sim_time = 1.0e-3, 1.0                                               # RD3:329-330
    tot_num_steps = int(total_sim_time / time_step) if n_steps is None else int(n_steps)
    T = total_sim_time / prm.time_constant

    pp = solver3d.PoreProblem(mesh, L, R, [prm], device=device, intended_bcs=True, rxn_diff=True)
    out = pp.march(tot_num_steps, history=True)
    end_time = datetime.now().strftime("%y-%m-%d-%H-%M-%S")
    hist = out["history"][:, 0]                                   # [steps + 1, nvert, 9]
    arrays = {n: hist[:, :, i] for i, n in enumerate(NAMES)}
    G = pp.solver.grad_project(out["u"])[0].cpu().numpy()        # [nvert, 9, 3]; project(grad(u_n_i), W), RD3:634-653
    grads = {n + "_grad": np.ascontiguousarray(G[:, i, :].T).ravel() for i, n in enumerate(NAMES)}
    tau_array = np.linspace(0, T, tot_num_steps)                  # RD3:656
    bulk_conc = dict(zip(prm.species, prm.c0))
    diff_eff = dict(zip(prm.species, prm.D))
    scaled = {}
    for n in NAMES:
        c, t, gs = scale_conc_time(species=n, C=arrays[n], grad_c=grads[n + "_grad"], bulk_conc=bulk_conc,
                                   tau=tau_array, diff_coeff_eff=diff_eff, L=L)
        scaled["c_" + n], scaled["t_" + n], scaled[n + "_grad"] = c, t, gs
    scaled["c_cat"] = scaled["c_HCO3"] + 2 * scaled["c_CO32"] + scaled["c_OH"] - scaled["c_H"]       # RD3:740
    eq = prm.extras["eq_conc"]
    metadata = {
        "concentration_elec": concentration_elec, "cation": cation, "H2_FE": H2_FE, "L": L, "R": R,
        "time_step": time_step, "total_sim_time": total_sim_time, "porosity": porosity_eff,
        "tortuosity": tortuosity_eff, "constrictivity": constrictivity_eff, "y_CO2": y_CO2, "press_gas": press_gas,
        "pore_geom_multiplier": pore_geom_multiplier,
        "electrolyte_flow_geom_multiplier": electrolyte_flow_geom_multiplier, "end_time": end_time,
        "eq_conc_CO": float(eq[1]), "eq_conc_H2": float(eq[2]), "current_planar": prm.extras["current_planar"],
        "CO2_min": float(arrays["CO2"][-1].min()),                                                   # RD3:770-790
        "newton_iterations": out["iters"][:, 0].tolist(), "gmres_iterations": out["lin_iters"][:, 0].tolist(),
        "CO2_entry_scaled": out["co2_entry"][:, 0].tolist(),
        # the two components the kernels carry along and the reference model does not have
        "passenger_drift": float(max(np.abs(hist[1:, :, 7] - 1.0).max(), np.abs(hist[1:, :, 8]).max()))}
    if write:
        identifier = "L_" + str(int(L * 1e+9)) + "_R_" + str(int(R * 1e+9)) + "_P_g_" + str(press_gas) + \
            "_D_eff_" + str(pore_geom_multiplier) + "_Re_" + str(electrolyte_flow_geom_multiplier) + \
            "_rough_" + str(roughness_factor)                                                        # RD3:354-359
        newpath = os.path.join(out_dir or os.path.join(os.getcwd(), "out"), stamp + "_experiment", identifier)
        os.makedirs(newpath, exist_ok=True)
        np.savez(os.path.join(newpath, "arrays_unscaled.npz"), coor=mesh.x, tau=tau_array, **arrays, **grads)
        np.savez(os.path.join(newpath, "arrays_scaled.npz"), coor_scaled=mesh.x * L, **scaled)
        with open(os.path.join(newpath, "metadata.json"), "w") as f:
            f.write(json.dumps(metadata, indent=0))
        if pvd:                                                    # File(newpath + '/solution_X.pvd') << _u_X, RD3:619-632
            from . import vtkio
            for n in ("CO", "H2", "CO2", "OH", "H", "HCO3", "CO32"):
                vtkio.write_pvd(os.path.join(newpath, "solution_" + n + ".pvd"), mesh.x, mesh.cells, arrays[n][-1], name=n)
        metadata["output_dir"] = newpath
    metadata["final_state"] = hist[-1][:, :7]
    pp.solver.close()
    return metadata


def build_parser():
    """The reference's argparse (RD3:793-924), flag for flag, plus the additions."""
    p = argparse.ArgumentParser(description="experiment parameters")
    p.add_argument("--concentration_elec", default=1.0, type=float, help="float val, 1.0 M")
    p.add_argument("--H2_FE", default=0.05, type=float)
    p.add_argument("--current_rough", default=3000.0, type=float)
    p.add_argument("--L", default=100.0e-9, type=float)
    p.add_argument("--R", default=5.0e-9, type=float)
    p.add_argument("--cation", default="K", type=str)
    p.add_argument("--porosity_eff", default=0.5, type=float)
    p.add_argument("--tortuosity_eff", default=1.5, type=float)
    p.add_argument("--constrictivity_eff", default=0.9, type=float)
    p.add_argument("--press_gas", default=1.0, type=float)
    p.add_argument("--pore_geom_multiplier", default=1.0, type=float)
    p.add_argument("--electrolyte_flow_geom_multiplier", default=1.0, type=float)
    p.add_argument("--params_file", default="parameters_pore", type=str)
    p.add_argument("--y_CO2", default=0.95, type=float)
    p.add_argument("--roughness_factor", default=150.0, type=float)
    p.add_argument("--utilities_dir", default=None)
    p.add_argument("--out_dir", default=None)
    p.add_argument("--n_steps", default=None, type=int)
    p.add_argument("--mesh_file", default=None)
    p.add_argument("--device", default=0, type=int)
    p.add_argument("--no_pvd", dest="pvd", action="store_false", help="do not write the solution_*.pvd files")
    return p


def main(argv=None):
    a = build_parser().parse_args(argv)
    meta = solveEDL(**vars(a))
    print(json.dumps({k: meta[k] for k in ("newton_iterations", "gmres_iterations", "CO2_min", "output_dir")}))


if __name__ == "__main__":
    main()
