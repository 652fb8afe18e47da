"""The drop-in scripts keep the reference's CLI surface (flags, defaults) -- no GPU needed for this."""
import inspect

from gmpnp_b200 import edl1d, pore3d, rxn_diff3d


def test_1d_cli_flags_and_defaults_match_reference():
    # 1D/MPNP_CO2ER_EDL.py:993-1101
    a = edl1d.build_parser().parse_args([])
    assert (a.concentration_elec, a.model, a.voltage_multiplier, a.mesh_structure) == (0.1, "MPNP", -1.0, "variable")
    assert (a.H2_FE, a.current_OHP_ss, a.L_n, a.stabilization, a.H_OHP) == (0.2, 10.0, 50.0e-6, "N", None)
    assert (a.cation, a.params_file, a.dry_run) == ("K", "parameters", True)
    assert a.staging == "as_executed" and inspect.signature(edl1d.solve_EDL).parameters["staging"].default == "as_executed"
    b = edl1d.build_parser().parse_args(["--voltage_multiplier=-10.0", "--cation=Cs", "--dry_run", "False", "--H_OHP", "1.1"])
    assert b.voltage_multiplier == -10.0 and b.cation == "Cs" and b.dry_run is False and b.H_OHP == 1.1
    sig = inspect.signature(edl1d.solve_EDL)
    ref = ["concentration_elec", "model", "voltage_multiplier", "H2_FE", "mesh_structure", "current_OHP_ss", "L_n",
           "stabilization", "H_OHP", "cation", "params_file", "dry_run"]            # 1D:66-79
    assert list(sig.parameters)[:len(ref)] == ref
    assert sig.parameters["dry_run"].default is True and sig.parameters["L_n"].default == 50.0e-6


def test_3d_cli_flags_and_defaults_match_reference():
    # 3D/MPNP_CO2ER_pore.py:1088-1233
    a = pore3d.build_parser().parse_args([])
    assert (a.concentration_elec, a.voltage_multiplier, a.H2_FE, a.current_rough) == (1.0, -1.0, 0.05, 3000.0)
    assert (a.L, a.R, a.cation, a.porosity_eff, a.tortuosity_eff, a.constrictivity_eff) == (100.0e-9, 5.0e-9, "K", 0.5, 1.5, 0.9)
    assert (a.press_gas, a.pore_geom_multiplier, a.electrolyte_flow_geom_multiplier) == (1.0, 1.0, 1.0)
    assert (a.params_file, a.y_CO2, a.roughness_factor) == ("parameters_pore", 0.95, 150.0)
    sig = inspect.signature(pore3d.solveEDL)
    ref = ["concentration_elec", "voltage_multiplier", "H2_FE", "current_rough", "L", "cation", "R", "press_gas",
           "pore_geom_multiplier", "porosity_eff", "tortuosity_eff", "constrictivity_eff", "params_file", "y_CO2",
           "electrolyte_flow_geom_multiplier", "roughness_factor"]                     # 3D:96-113
    assert list(sig.parameters)[:len(ref)] == ref


def test_3d_rxn_diff_cli_flags_and_defaults_match_reference():
    # 3D/rxn_diff_CO2ER_pore.py:793-942
    a = rxn_diff3d.build_parser().parse_args([])
    assert (a.concentration_elec, a.H2_FE, a.current_rough, a.L, a.R, a.cation) == (1.0, 0.05, 3000.0, 100e-9, 5e-9, "K")
    assert (a.porosity_eff, a.tortuosity_eff, a.constrictivity_eff, a.press_gas) == (0.5, 1.5, 0.9, 1.0)
    assert (a.pore_geom_multiplier, a.electrolyte_flow_geom_multiplier, a.params_file) == (1.0, 1.0, "parameters_pore")
    assert (a.y_CO2, a.roughness_factor) == (0.95, 150.0)
    assert not hasattr(a, "voltage_multiplier")                       # the model has no potential
    sig = inspect.signature(rxn_diff3d.solveEDL)
    ref = ["concentration_elec", "H2_FE", "current_rough", "L", "cation", "R", "press_gas", "pore_geom_multiplier",
           "porosity_eff", "tortuosity_eff", "constrictivity_eff", "params_file", "y_CO2",
           "electrolyte_flow_geom_multiplier", "roughness_factor"]                     # RD3:96-111
    assert list(sig.parameters)[:len(ref)] == ref


def test_3d_rxn_diff_parameters_switch_the_electrostatics_off():
    import numpy as np
    from gmpnp_b200 import params
    p = rxn_diff3d.params_rxn_diff_3d(L=50e-9, R=5e-9)
    q = params.params_3d(L=50e-9, R=5e-9)
    assert not p.z.any() and not p.nu.any() and p.V == 0.0 and p.extras["k_exit"][7] == 0.0
    assert np.array_equal(p.extras["scale_R"], q.extras["scale_R"]) and p.kappa == q.kappa      # RD3:266-277 = 3D:270-277
    assert np.array_equal(p.extras["J_wall"], q.extras["J_wall"])                               # RD3:422-431 = 3D:474-483
    assert np.array_equal(p.extras["k_exit"][:7], q.extras["k_exit"][:7])
    assert q.extras["k_exit"][7] != 0.0                                # the MPNP parameters are left untouched
