"""Drop-in for ``1D/MPNP_CO2ER_EDL.py``: same function signature, same CLI flags and defaults, same
YAML / dolfin-XML inputs, same output files and keys -- with the FEniCS hot path (``solve(F == 0, u,
bcs)`` inside the pseudo-time loop, 1D:633-796) replaced by the CUDA library.

    python -m gmpnp_b200.edl1d --voltage_multiplier=-10.0 --cation=Cs

Differences from the reference, all behind new keyword-only arguments / flags:
``--utilities_dir`` / ``--out_dir`` replace the hard-coded absolute paths (1D:85, 293); ``--dry_run``
is parsed as a real boolean (the reference's ``type=bool`` turns any string into True, 1D:1094-1101);
the non-dry staged run defines ``time_step`` / ``total_sim_time`` for the metadata (the reference
crashes with NameError at 1D:971-972); ``--mode steady`` selects the new steady solve with voltage
continuation (BASELINE.json north_star) instead of the march.

Staging of the non-dry run (1D:271-290, 641-648): the reference means to switch from dt = 1e-5 s to 1e-3 s at
t = 0.1 s, but it only rebinds the Python name ``del_t`` to a second ``Constant``; the form ``F`` was built with the
first ``Constant`` object (1D:458 ff.) and is never rebuilt, so all 10 000 + 10 000 steps are integrated with 1e-5 s
(physical end time 0.2 s) while the stored time axis pretends 10.1 s.  ``staging="as_executed"`` (default) reproduces
that; ``--staging intended`` uses the two step sizes.  The time axis written to the files is the reference's in both.
"""
from __future__ import annotations

import argparse
import json
import math
import os
from datetime import datetime

import numpy as np


def scale(species="H", tau=None, C=None, initial_conc=None, diff_coeff=None, L_n=0.0, L_debye=0.0):
    """Dimensionless -> SI (1D:51-63)."""
    t = (tau * L_debye * L_n) / diff_coeff[species]
    c = C * initial_conc[species]
    return t, c


def solve_EDL(concentration_elec=0.1, model="MPNP", voltage_multiplier=-1.0, H2_FE=0.2,
              mesh_structure="variable", current_OHP_ss=10.0, L_n=50.0e-6, stabilization="N", H_OHP=None,
              cation="K", params_file="parameters", dry_run=True, *, utilities_dir=None, out_dir=None,
              mode="march", device=0, n_steps=None, write=True, staging="as_executed"):
    import torch
    from . import meshio, params as _params, solver1d
    from ._lib import NewtonOpts

    stamp = datetime.now().strftime("%y-%m-%d-%H-%M-%S")
    cat_str = cation
    if stabilization == "Y":
        if model == "PNP":
            raise NotImplementedError("SUPG stabilisation (1D:597-621, 650-722) is outside the GMPNP hot path")
        print("Warning:stabilization not implemented for MPNP!")        # 1D:724-727: solves the plain form

    mesh_name = _params.mesh_name_1d(L_n, mesh_structure)
    mesh = meshio.load_mesh(mesh_name, utilities_dir)
    x = mesh.x[:, 0]
    num_vertices = mesh.num_vertices
    mesh_number = int(mesh_name.rsplit("_", 1)[1])
    mesh_structure_out = mesh_structure + ("_" + str(int(L_n * 1.0e+6)) + "um" if mesh_structure == "variable" else "")

    # time staging (1D:256-290)
    if dry_run:
        stages = [(1.0e-5, 1.0e-3)]
    else:
        stages = [(1.0e-5, 0.1), (1.0e-3, 10.1)]
    prm = _params.params_1d(concentration_elec=concentration_elec, model=model,
                            voltage_multiplier=voltage_multiplier, H2_FE=H2_FE, current_OHP_ss=current_OHP_ss,
                            L_n=L_n, H_OHP=H_OHP, cation=cation, params_file=params_file,
                            utilities_dir=utilities_dir, time_step=stages[0][0])
    prm.extras["H_OHP"] = H_OHP
    time_constant = prm.time_constant

    solver = solver1d.Solver1D(x, batch=1, device=device)
    dev = solver.device
    hist_rows = [np.tile(np.array([1.0] * 6 + [0.0]), (num_vertices, 1))]      # row 0 = initial state (1D:623-629)
    newton_its = []
    current_H_frac = prm.extras["current_H_frac"]

    if mode == "march":
        u = torch.zeros(1, num_vertices, 7, dtype=torch.float64, device=dev)     # u = Function(V) (1D:320)
        un = solver1d.bulk_state(1, num_vertices, dev)                           # u_n (1D:322-326)
        t_prev = 0.0
        tau_parts = []
        for si, (time_step, t_end) in enumerate(stages):
            ns = int((t_end - t_prev) / time_step) if si else int(t_end / time_step)
            if n_steps is not None:
                ns = min(ns, n_steps - sum(len(p) for p in tau_parts)) if tau_parts else min(ns, n_steps)
            if ns <= 0:
                break
            # the step size that is IN THE FORM: the first stage's for every stage as executed (module docstring)
            dt_form = time_step if staging == "intended" else stages[0][0]
            p_stage = _params.params_1d(concentration_elec=concentration_elec, model=model,
                                        voltage_multiplier=voltage_multiplier, H2_FE=H2_FE,
                                        current_OHP_ss=current_OHP_ss, L_n=L_n, H_OHP=H_OHP, cation=cation,
                                        params_file=params_file, utilities_dir=utilities_dir, time_step=dt_form,
                                        current_H_frac=current_H_frac)
            p_stage.extras["H_OHP"] = H_OHP
            solver.set_params([p_stage])
            o_ref = NewtonOpts.reference_1d()
            o_ref.partitions = 0          # one problem: the partitioned elimination (8 sweeps) halves the latency
            out = solver.march(u, un, ns, o_ref, history=True)
            status = int(out["status"][0])
            if status != 0:
                raise RuntimeError("Newton solver did not converge (status %d)" % status)   # dolfin raises too
            hist_rows += list(out["history"][0].cpu().numpy())
            newton_its += out["iters"][0].cpu().tolist()
            current_H_frac = float(out["hfrac"][0])
            T0 = t_prev / time_constant
            T1 = t_end / time_constant
            # tau arrays replicate np.linspace(0, T, steps) / the staged concatenation (1D:807-815)
            tau_parts.append(np.linspace(T0 + (time_step / time_constant if si else 0.0), T1, ns))
            t_prev = t_end
        tau_array = np.concatenate(tau_parts)
        time_step_meta, total_sim_time_meta = stages[-1][0] if not dry_run else stages[0][0], stages[-1][1]
        if dry_run:
            time_step_meta, total_sim_time_meta = stages[0]
    elif mode == "steady":
        solver.set_params([prm])
        u = solver1d.bulk_state(1, num_vertices, dev)
        nst = max(1, int(math.ceil(abs(voltage_multiplier) / 0.5 - 1e-12)))
        Vpath = (voltage_multiplier * np.arange(1, nst + 1) / nst)[None, :]
        o_st = NewtonOpts.steady(xtol=1e-12, xtol_path=1e-3, jac_rule=1)
        o_st.partitions = 0
        out = solver.steady(u, Vpath, o_st)
        if int(out["status"][0]) != 0:
            raise RuntimeError("steady Newton did not converge (status %d)" % int(out["status"][0]))
        hist_rows.append(u[0].cpu().numpy())
        newton_its = out["iters"][0].cpu().tolist()
        tau_array = np.array([0.0])
        time_step_meta, total_sim_time_meta = float("inf"), float("inf")
    else:
        raise ValueError(mode)

    end_time = datetime.now().strftime("%y-%m-%d-%H-%M-%S")
    hist = np.array(hist_rows)                                     # [steps+1, nvert, 7]
    H, OH, HCO3, CO32, CO2, cat, p = (hist[:, :, i] for i in range(7))
    coor_array = mesh.x.copy()

    # electric field of the last potential profile: project(-grad(u_np), W) (1D:802-805)
    field_values = solver.field(u)[0].cpu().numpy()
    thermal_voltage = prm.thermal_voltage
    field_values_rescaled = field_values * thermal_voltage / L_n
    field_OHP = field_values_rescaled[0] * 1.0e-9                  # V/nm

    species = prm.species
    initial_conc = dict(zip(species, prm.c0))
    diff_coeff = dict(zip(species, prm.D))
    L_debye = prm.extras["L_debye"]
    sc = {}
    for name, arr in zip(species, (H, OH, HCO3, CO32, CO2, cat)):
        sc[name] = scale(species=name, tau=tau_array, C=arr, initial_conc=initial_conc, diff_coeff=diff_coeff,
                         L_n=L_n, L_debye=L_debye)
    (t_H, c_H), (t_OH, c_OH), (t_HCO3, c_HCO3) = sc["H"], sc["OH"], sc["HCO3"]
    (t_CO32, c_CO32), (t_CO2, c_CO2), (t_cat, c_cat) = sc["CO32"], sc["CO2"], sc[cat_str]
    coor_scaled = coor_array * L_n
    psi = p * thermal_voltage
    pH_OHP = -math.log10(c_H[-1][0] / 1000)
    eps_rel = prm.eps_w
    n_w_cat, n_w_H = prm.n_water_cat, prm.n_water_H
    eps_rel_conc_ss = eps_rel * ((55 - (n_w_cat * c_cat + n_w_H * c_H) * 1.0e-3) / 55) \
        + 6 * (((n_w_cat * c_cat + n_w_H * c_H) * 1.0e-3) / 55)                                  # 1D:895-898
    eps_rel_OHP = eps_rel_conc_ss[-1][0]
    charge_density = c_cat[-1] - c_HCO3[-1] - 2 * c_CO32[-1] - c_OH[-1] + c_H[-1]                # 1D:903-904
    potential_OHP = psi[-1][0]
    CO2_OHP_frac = c_CO2[-1][0] / initial_conc["CO2"]
    bulk_pH = prm.extras["bulk_pH"]
    pH_overpotential = -0.059 * (bulk_pH - pH_OHP) * 1.0e+3
    CO2_overpotential = (0.059 / 2) * math.log10(1 / CO2_OHP_frac) * 1.0e+3
    current_H = current_H_frac * current_OHP_ss

    metadata_dict = {
        "concentration_elec": concentration_elec, "cation": cation, "model": model,
        "stabilization": stabilization, "voltage_multiplier": voltage_multiplier, "H2_FE": H2_FE,
        "L_n_EDL": L_n, "time_constant": time_constant, "time_step": time_step_meta,
        "total_sim_time": total_sim_time_meta, "mesh_number": mesh_number, "mesh_structure": mesh_structure_out,
        "eps_rel_OHP": float(eps_rel_OHP), "field_OHP": float(field_OHP), "current_OHP_ss": current_OHP_ss,
        "current_H": current_H, "H_OHP_vs_bulk": H_OHP, "potential_OHP": float(potential_OHP),
        "pH_OHP": float(pH_OHP), "CO2_OHP_frac": float(CO2_OHP_frac), "pH_overpotential": float(pH_overpotential),
        "CO2_overpotential": float(CO2_overpotential), "end_time": end_time,
        # additions (not in the reference's file): solver bookkeeping
        "mode": mode, "newton_iterations": [int(k) for k in newton_its]}

    if write:
        identifier = "voltage_" + str(voltage_multiplier) + "_H2_FE_" + str(H2_FE) + "_current_" + \
            str(current_OHP_ss) + "_H_OHP_" + str(H_OHP) + "_cation_" + cat_str                     # 1D:211-213
        basepath = os.path.join(out_dir or os.path.join(os.getcwd(), "out"), model)
        newpath = os.path.join(basepath, stamp + "_experiment", identifier)
        os.makedirs(newpath, exist_ok=True)
        np.savez(os.path.join(newpath, "arrays_unscaled.npz"), H=H, OH=OH, HCO3=HCO3, CO32=CO32, CO2=CO2, cat=cat,
                 p=p, coor=coor_array, tau=tau_array, field_values=field_values)                     # 1D:821-832
        np.savez(os.path.join(newpath, "arrays_scaled.npz"), x=coor_scaled, psi=psi, t_H=t_H, c_H=c_H, t_OH=t_OH,
                 c_OH=c_OH, t_HCO3=t_HCO3, c_HCO3=c_HCO3, t_CO32=t_CO32, c_CO32=c_CO32, t_CO2=t_CO2, c_CO2=c_CO2,
                 t_cat=t_cat, c_cat=c_cat, eps_rel=eps_rel_conc_ss, field_values=field_values_rescaled,
                 charge_density=charge_density)                                                      # 1D:906-924
        with open(os.path.join(newpath, "metadata.json"), "w") as f:
            f.write(json.dumps(metadata_dict, indent=0))
        metadata_dict["output_dir"] = newpath
    solver.close()
    return metadata_dict


def _bool(s):
    if isinstance(s, bool):
        return s
    if s.lower() in ("1", "true", "t", "yes", "y"):
        return True
    if s.lower() in ("0", "false", "f", "no", "n"):
        return False
    raise argparse.ArgumentTypeError("boolean expected")


def _opt_float(s):
    return None if s in (None, "None", "none", "") else float(s)


def build_parser():
    """The reference's argparse (1D:993-1101), flag for flag, plus the path/mode additions."""
    parser = argparse.ArgumentParser(description="experiment parameters")
    parser.add_argument("--concentration_elec", metavar="electrolyte_concentration", required=False,
                        help="float val, 0.1 M", default=0.1, type=float)
    parser.add_argument("--model", metavar="model_type", required=False, help="str, PNP/MPNP", default="MPNP", type=str)
    parser.add_argument("--voltage_multiplier", metavar="thermal_voltage_multiplier", required=False,
                        help="float val, -1.0", default=-1.0, type=float)
    parser.add_argument("--mesh_structure", metavar="bias in mesh structure", required=False,
                        help="str, uniform/variable", default="variable", type=str)
    parser.add_argument("--H2_FE", metavar="faradaic efficiency for hydrogen in fraction", required=False,
                        help="float val, 0.2", default=0.2, type=float)
    parser.add_argument("--current_OHP_ss", metavar="steady state current in A/m2", required=False,
                        help="float val, 10.0", default=10.0, type=float)
    parser.add_argument("--L_n", metavar="system size", required=False, help="float val, 50.0e-6", default=50.0e-6,
                        type=float)
    parser.add_argument("--stabilization", metavar="SUPG", required=False, help="str, Y/N", default="N", type=str)
    parser.add_argument("--H_OHP", metavar="build up of protons at the OHP relative to the bulk", required=False,
                        help="float val, None/1.1/2.0", default=None, type=_opt_float)
    parser.add_argument("--cation", metavar="monovalent cation in solution", required=False, help="str, K/Cs/Li",
                        default="K", type=str)
    parser.add_argument("--params_file", metavar="yaml file with parameter values", required=False,
                        help="str, parameters", default="parameters", type=str)
    parser.add_argument("--dry_run", metavar="run 100 time steps as test", required=False, help="boolean value",
                        default=True, type=_bool)
    # additions
    parser.add_argument("--utilities_dir", default=None, help="folder with the reference's utilities/ files")
    parser.add_argument("--out_dir", default=None, help="output base folder (default ./out)")
    parser.add_argument("--mode", default="march", choices=["march", "steady"])
    parser.add_argument("--staging", default="as_executed", choices=["as_executed", "intended"],
                        help="non-dry run: keep dt = 1e-5 s in the form for all 20000 steps as the reference does "
                             "(its del_t rebinding never reaches the form), or switch to 1e-3 s at t = 0.1 s")
    parser.add_argument("--device", default=0, type=int)
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    meta = solve_EDL(concentration_elec=args.concentration_elec, model=args.model,
                     voltage_multiplier=args.voltage_multiplier, mesh_structure=args.mesh_structure,
                     H2_FE=args.H2_FE, current_OHP_ss=args.current_OHP_ss, L_n=args.L_n,
                     stabilization=args.stabilization, H_OHP=args.H_OHP, cation=args.cation,
                     params_file=args.params_file, dry_run=args.dry_run, utilities_dir=args.utilities_dir,
                     out_dir=args.out_dir, mode=args.mode, device=args.device, staging=args.staging)
    print(json.dumps({k: meta[k] for k in ("field_OHP", "eps_rel_OHP", "pH_OHP", "potential_OHP", "output_dir")}))


if __name__ == "__main__":
    main()
