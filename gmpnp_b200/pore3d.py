"""Drop-in for ``3D/MPNP_CO2ER_pore.py``: same function signature, CLI flags and defaults, same YAML / dolfin-XML
inputs, same output arrays -- with the FEniCS hot path (``solve(F == 0, u, bcs, {newton, mumps, relaxation 0.9})``
inside the pseudo-time loop, 3D:782-858) replaced by the CUDA library.

    python -m gmpnp_b200.pore3d --L 50e-9 --R 5e-9 --n_steps 10

Behaviour kept on purpose (SURVEY findings 3, 4, 6): the residual has no boundary integrals (the reference's
``+ J_... * ds(k)`` lines are dead code), the wall marker uses the absolute tolerance on r^2, the mesh file name is
built with int() truncation.  Additions: ``--utilities_dir`` / ``--out_dir`` instead of hard-coded paths,
``--n_steps`` (the reference always runs 1000 steps, of which all but the first few are no-ops once the residual is
below its 1e-4 tolerance), ``--mesh_file`` override.  The P1 gradient projections (``field_values``, ``*_grad``,
3D:884-909) are computed on the device (``gmpnp_grad_project_3d``) and written with the reference's keys and
component-major layout; the final fields also go to ``solution_{CO,K,H2,CO2,OH,H,HCO3,CO32,p}.pvd`` (3D:863-880,
``gmpnp_b200/vtkio.py``; ``--no_pvd`` skips them).
"""
from __future__ import annotations

import argparse
import json
import os
from datetime import datetime

import numpy as np


def scale_conc_time(species="H", C=None, grad_c=None, bulk_conc=None, tau=None, diff_coeff_eff=None, L=0.0):
    """Dimensionless -> SI (3D:56-67)."""
    c = C * bulk_conc[species]
    t = tau * (L ** 2) / diff_coeff_eff[species]
    grad_c_scaled = grad_c * bulk_conc[species] / L
    return c, t, grad_c_scaled


def solveEDL(concentration_elec=1.0, voltage_multiplier=-1.0, H2_FE=0.05, current_rough=3000.0, L=100.0e-9,
             cation="K", R=5.0e-9, press_gas=1.0, pore_geom_multiplier=1.0, porosity_eff=0.5, tortuosity_eff=1.5,
             constrictivity_eff=0.9, params_file="parameters_pore", y_CO2=0.95, electrolyte_flow_geom_multiplier=1.0,
             roughness_factor=150.0, *, utilities_dir=None, out_dir=None, n_steps=None, mesh_file=None, device=0,
             write=True, intended_bcs=False, pvd=True):
    from . import meshio, params as _params, solver3d

    stamp = datetime.now().strftime("%y-%m-%d-%H-%M-%S")
    prm = _params.params_3d(concentration_elec=concentration_elec, voltage_multiplier=voltage_multiplier,
                            H2_FE=H2_FE, current_rough=current_rough, L=L, cation=cation, R=R, press_gas=press_gas,
                            pore_geom_multiplier=pore_geom_multiplier, porosity_eff=porosity_eff,
                            tortuosity_eff=tortuosity_eff, constrictivity_eff=constrictivity_eff,
                            params_file=params_file, y_CO2=y_CO2,
                            electrolyte_flow_geom_multiplier=electrolyte_flow_geom_multiplier,
                            roughness_factor=roughness_factor, utilities_dir=utilities_dir)
    mesh = meshio.load_mesh(mesh_file or _params.mesh_name_3d(L, R), utilities_dir)       # 3D:329-332
    time_step, total_sim_time = 1.0e-3, 1.0                                               # 3D:358-359
    tot_num_steps = int(total_sim_time / time_step) if n_steps is None else int(n_steps)
    T = total_sim_time / prm.time_constant

    pp = solver3d.PoreProblem(mesh, L, R, [prm], device=device, intended_bcs=intended_bcs)
    out = pp.march(tot_num_steps, history=True)
    end_time = datetime.now().strftime("%y-%m-%d-%H-%M-%S")
    hist = out["history"][:, 0]                                   # [steps+1, nvert, 9]
    names = ["H", "OH", "HCO3", "CO32", "CO2", "CO", "H2", "cat", "p"]
    arrays = {n: hist[:, :, i] for i, n in enumerate(names)}
    # project(grad(u_i), W).compute_vertex_values(): flat, component-major [all x | all y | all z] (3D:884-909)
    G = pp.solver.grad_project(out["u"])[0].cpu().numpy()        # [nvert, 9, 3]
    grads = {n + "_grad": np.ascontiguousarray(G[:, i, :].T).ravel() for i, n in enumerate(names[:8])}
    field_values = -np.ascontiguousarray(G[:, 8, :].T).ravel()   # project(-grad(u_np), W), 3D:884-885
    tau_array = np.linspace(0, T, tot_num_steps)                  # 3D:912
    species = prm.species
    bulk_conc = dict(zip(species, prm.c0))
    diff_eff = dict(zip(species, prm.D))
    scaled = {}
    for n, sp in zip(names[:8], species):
        c, t, gs = scale_conc_time(species=sp, C=arrays[n], grad_c=grads[n + "_grad"], bulk_conc=bulk_conc,
                                   tau=tau_array, diff_coeff_eff=diff_eff, L=L)
        scaled["c_" + n], scaled["t_" + n], scaled[n + "_grad"] = c, t, gs
    psi = arrays["p"] * prm.thermal_voltage
    scaled["field_values"] = field_values * prm.thermal_voltage / L                                  # 3D:1016
    metadata = {
        "concentration_elec": concentration_elec, "cation": cation, "voltage_multiplier": voltage_multiplier,
        "H2_FE": H2_FE, "current_rough": current_rough, "L": L, "R": R, "press_gas": press_gas,
        "pore_geom_multiplier": pore_geom_multiplier, "porosity_eff": porosity_eff, "tortuosity_eff": tortuosity_eff,
        "constrictivity_eff": constrictivity_eff, "y_CO2": y_CO2,
        "electrolyte_flow_geom_multiplier": electrolyte_flow_geom_multiplier, "roughness_factor": roughness_factor,
        "time_constant": prm.time_constant, "time_step": time_step, "total_sim_time": total_sim_time,
        "num_vertices": mesh.num_vertices, "end_time": end_time,
        "newton_iterations": out["iters"][:, 0].tolist(), "gmres_iterations": out["lin_iters"][:, 0].tolist(),
        "CO2_entry_scaled": out["co2_entry"][:, 0].tolist(), "dirichlet_info": pp.info}
    if write:
        identifier = "v_" + str(voltage_multiplier) + "_L_" + str(int(L * 1e+9)) + "_R_" + str(int(R * 1e+9)) + \
            "_P_g_" + str(press_gas) + "_D_eff_" + str(pore_geom_multiplier) + "_Re_" + \
            str(electrolyte_flow_geom_multiplier) + "_rough_" + str(roughness_factor)               # 3D:389-395
        newpath = os.path.join(out_dir or os.path.join(os.getcwd(), "out"), stamp + "_experiment", identifier)
        os.makedirs(newpath, exist_ok=True)
        np.savez(os.path.join(newpath, "arrays_unscaled.npz"), coor=mesh.x, tau=tau_array, field_values=field_values,
                 **arrays, **grads)                                                                    # 3D:916-937
        np.savez(os.path.join(newpath, "arrays_scaled.npz"), x=mesh.x * L, psi=psi, **scaled)          # 3D:1026-1056
        with open(os.path.join(newpath, "metadata.json"), "w") as f:
            f.write(json.dumps(metadata, indent=0))
        if pvd:                                                    # File(newpath + '/solution_X.pvd') << _u_X, 3D:863-880
            from . import vtkio
            for fname, key in (("CO", "CO"), ("K", "cat"), ("H2", "H2"), ("CO2", "CO2"), ("OH", "OH"), ("H", "H"),
                               ("HCO3", "HCO3"), ("CO32", "CO32"), ("p", "p")):
                vtkio.write_pvd(os.path.join(newpath, "solution_" + fname + ".pvd"), mesh.x, mesh.cells,
                                arrays[key][-1], name=key)
        metadata["output_dir"] = newpath
    pp.solver.close()
    return metadata


def build_parser():
    """The reference's argparse (3D:1088-1233), flag for flag, plus the additions."""
    p = argparse.ArgumentParser(description="experiment parameters")
    p.add_argument("--concentration_elec", default=1.0, type=float, help="float val, 1.0 M")
    p.add_argument("--voltage_multiplier", default=-1.0, type=float, help="float val, -1.0")
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
    p.add_argument("--intended_bcs", action="store_true",
                   help="add the wall-flux / pore-exit Robin boundary integrals that the reference writes but Python "
                        "discards (3D:474-499, 560-750); default: as executed")
    return p


def main(argv=None):
    a = build_parser().parse_args(argv)
    kw = vars(a)
    meta = solveEDL(**kw)
    print(json.dumps({k: meta[k] for k in ("newton_iterations", "gmres_iterations", "output_dir")}))


if __name__ == "__main__":
    main()
