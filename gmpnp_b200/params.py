"""Parameter layer: YAML scalars -> dimensionless groups of the GMPNP system.

Host-side, once per sweep point.  Follows the reference's arithmetic literally
(same operation order, so the known-answer constants of SURVEY App. D match to
the last bit):

* 1D planar EDL:  1D/MPNP_CO2ER_EDL.py:81-213 (inputs, L_debye, L_D, Vt,
  time_constant, scale_R, q, scale_vol, flux prefactors), :256-268 (time step),
  :368-375 (OHP fluxes).
* 3D pore:        3D/MPNP_CO2ER_pore.py:115-324 (inputs, D_eff, Henry gas
  concentrations, scale_R, q, scale_vol, J prefactors, Re/Sc/Sh/k_elec),
  :70-93 (Sechenov CO2 solubility), :358-365 (time step), :469-499 (fluxes).

The result is a :class:`ProblemParams`; ``pack()`` lays it out as the flat
``double[GMPNP_NPAR]`` record that ``include/gmpnp.h`` documents and the CUDA
kernels read.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field

import numpy as np

_PKG_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

MAXS = 8           # max number of species slots in the packed record
NPAR = 64          # doubles per packed problem record (see include/gmpnp.h)

# offsets inside the packed record (keep in sync with include/gmpnp.h)
P_NS, P_Z, P_NU, P_ZC0, P_S = 0, 1, 9, 17, 25
P_KW, P_KA, P_KB, P_KA2, P_KB2, P_KW1 = 30, 31, 32, 33, 34, 35
P_EPSW, P_EPSH, P_EPSC, P_KAPPA, P_V = 36, 37, 38, 39, 40
P_JFLUX, P_ICAT, P_Q = 41, 49, 50

SPECIES_1D = ["H", "OH", "HCO3", "CO32", "CO2"]          # + cation  (1D:117)
SPECIES_3D = ["H", "OH", "HCO3", "CO32", "CO2", "CO", "H2"]  # + cation  (3D:138)


def _load_inputs(name: str, utilities_dir: str | None):
    """Read ``<name>.yaml`` from the reference's utilities folder if given, else the
    packaged JSON copy (tools/make_data.py)."""
    if utilities_dir:
        p = os.path.join(utilities_dir, name + ".yaml")
        if os.path.exists(p):
            import yaml
            with open(p) as f:
                return yaml.safe_load(f)
    with open(os.path.join(_PKG_DATA, "reference_inputs.json")) as f:
        allin = json.load(f)
    if name not in allin:
        raise FileNotFoundError(f"input '{name}.yaml' not found (utilities_dir={utilities_dir})")
    return allin[name]


def _bulk_conc(block: dict, sp: str, cation: str, allow_cation_fallback: bool) -> float:
    """``C0_<sp>`` from a bulk_soln block.  Documented extension (SURVEY finding 5):
    bulk_soln_0.5/1.0 only list C0_K; for another monovalent cation fall back to
    C0_K, which is what bulk_soln_0.1KHCO3.yaml itself does for every cation."""
    conc = block["concentrations"]
    key = "C0_" + sp
    if key in conc:
        return float(conc[key])
    if allow_cation_fallback and sp == cation and "C0_K" in conc:
        return float(conc["C0_K"])
    raise KeyError(key)


@dataclass
class ProblemParams:
    """Dimensionless description of one sweep point (1D or 3D)."""
    dim: int
    species: list           # names, cation last
    z: np.ndarray           # charges
    c0: np.ndarray          # bulk concentrations [mol/m3]
    D: np.ndarray           # (effective) diffusion coefficients used for scaling
    nu: np.ndarray          # scale_vol = a^3 c0 N_A   (0 for model 'PNP')
    s: np.ndarray           # scale_R of H, OH, HCO3, CO32, CO2
    rate: dict              # kw1 kw2 ka1 ka2 kb1 kb2
    q: float
    eps_w: float
    n_water_H: float
    n_water_cat: float
    kappa: float            # coefficient of (u-u_n) v dx; 0 = steady equations
    V: float                # scaled potential at the OHP (1D) / pore wall (3D)
    jflux: np.ndarray       # 1D: constants added to the node-0 rows (J_i of `J_i v ds`)
    length: float           # L_n (1D) or L (3D)
    thermal_voltage: float
    time_constant: float
    dt_scaled: float
    extras: dict = field(default_factory=dict)

    @property
    def ns(self) -> int:
        return len(self.species)

    @property
    def ncomp(self) -> int:
        return self.ns + 1

    def with_(self, **kw) -> "ProblemParams":
        import copy
        p = copy.copy(self)
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def pack(self) -> np.ndarray:
        """Flat double[NPAR] record for the C-ABI (include/gmpnp.h)."""
        ns = self.ns
        r = self.rate
        c0 = self.c0
        P = np.zeros(NPAR, dtype=np.float64)
        P[P_NS] = ns
        P[P_Z:P_Z + ns] = self.z
        P[P_NU:P_NU + ns] = self.nu
        P[P_ZC0:P_ZC0 + ns] = self.z * c0
        P[P_S:P_S + 5] = self.s
        iH, iOH, iHCO3, iCO32, iCO2 = 0, 1, 2, 3, 4
        P[P_KW] = r["kw2"] * c0[iH] * c0[iOH]
        P[P_KA] = r["ka1"] * c0[iOH] * c0[iHCO3]
        P[P_KB] = r["kb1"] * c0[iCO2] * c0[iOH]
        P[P_KA2] = r["ka2"] * c0[iCO32]
        P[P_KB2] = r["kb2"] * c0[iHCO3]
        P[P_KW1] = r["kw1"]
        P[P_EPSW] = self.eps_w
        P[P_EPSH] = self.n_water_H * c0[iH] * 1.0e-3
        P[P_EPSC] = self.n_water_cat * c0[ns - 1] * 1.0e-3
        P[P_KAPPA] = self.kappa
        P[P_V] = self.V
        P[P_JFLUX:P_JFLUX + ns] = self.jflux
        P[P_ICAT] = ns - 1
        P[P_Q] = self.q
        return P


def params_1d(concentration_elec=0.1, model="MPNP", voltage_multiplier=-1.0, H2_FE=0.2,
              current_OHP_ss=10.0, L_n=50.0e-6, H_OHP=None, cation="K",
              params_file="parameters", utilities_dir=None, time_step=1.0e-5,
              current_H_frac=None) -> ProblemParams:
    """Dimensionless groups of the 1D planar problem (1D/MPNP_CO2ER_EDL.py:81-213, 368-375)."""
    data = _load_inputs(params_file, utilities_dir)
    rate = {k: float(v) for k, v in data["rate_constants"].items()}
    cat = cation
    # hydration numbers are hard-coded in the 1D script (1D:106-115)
    n_w = {"K": 4, "Li": 5, "Cs": 3, "Na": 5}.get(cat, 0.0)
    species = SPECIES_1D + [cat]
    D = np.array([float(data["diff_coef"]["D_" + i]) for i in species])
    a = np.array([float(data["solv_size"]["a_" + i]) for i in species])
    nc = data["nat_const"]
    farad, temp, k_B, e_0 = float(nc["F"]), float(nc["T"]), float(nc["k_B"]), float(nc["e_0"])
    eps_0, eps_rel, R, N_A = float(nc["eps_0"]), float(nc["eps_rel"]), float(nc["R"]), float(nc["N_A"])

    bulk = _load_inputs("bulk_soln_" + str(concentration_elec) + "KHCO3", utilities_dir)
    post = bulk["bulk_conc_post_CO2"]              # 1D reads the post-CO2 block (1D:151,161)
    bulk_pH = float(post["final_pH"])
    c0 = np.array([_bulk_conc(post, i, cat, True) for i in species])
    z = np.array([1, -1, -1, -2, 0, 1], dtype=np.float64)   # 1D:158

    if current_H_frac is None:
        current_H_frac = 0.0 if H_OHP is None else 0.001    # 1D:167-170

    L_debye = math.sqrt((eps_0 * eps_rel * k_B * temp) /
                        (2 * e_0 ** 2 * concentration_elec * 1.0e+3 * N_A))   # 1D:173-176
    L_D = L_debye / L_n
    thermal_voltage = (k_B * temp) / e_0
    time_constant = L_debye * L_n / D[3]                    # D_CO32, 1D:183
    scale_R = (L_n ** 2) / (D * c0)                         # 1D:190
    q = (farad ** 2 * L_n ** 2) / (eps_0 * R * temp)        # 1D:193
    scale_vol = a ** 3 * c0 * N_A                           # 1D:200
    J_H_pref = L_n / (D[0] * c0[0] * farad)
    J_OH_pref = L_n / (D[1] * c0[1] * farad)
    J_CO2_pref = L_n / (D[4] * c0[4] * farad)
    CO_FE = 1 - H2_FE
    J_CO2 = J_CO2_pref * current_OHP_ss * 0.5 * (CO_FE)                       # 1D:371
    J_OH = J_OH_pref * current_OHP_ss * (1 - current_H_frac) * (-1.0)         # 1D:372-374
    J_H = J_H_pref * current_OHP_ss * current_H_frac                          # 1D:375
    jflux = np.zeros(6)
    jflux[0], jflux[1], jflux[4] = J_H, J_OH, J_CO2

    dt = time_step / time_constant                          # 1D:264
    kappa = 1.0 / (dt * L_D)                                # 1D:458  (u-u_n)/(del_t*L_D)
    nu = scale_vol if model == "MPNP" else np.zeros(6)      # PNP = steric term removed (1D:429-455)
    return ProblemParams(
        dim=1, species=species, z=z, c0=c0, D=D, nu=np.asarray(nu, dtype=np.float64),
        s=scale_R[:5].copy(), rate=rate, q=q, eps_w=eps_rel, n_water_H=10.0, n_water_cat=float(n_w),
        kappa=kappa, V=float(voltage_multiplier), jflux=jflux, length=L_n,
        thermal_voltage=thermal_voltage, time_constant=time_constant, dt_scaled=dt,
        extras=dict(L_debye=L_debye, L_D=L_D, bulk_pH=bulk_pH, scale_R=scale_R, scale_vol=scale_vol,
                    J_H_prefactor=J_H_pref, J_OH_prefactor=J_OH_pref, J_CO2_prefactor=J_CO2_pref,
                    current_H_frac=current_H_frac, current_OHP_ss=current_OHP_ss, farad=farad,
                    model=model, cation=cat, concentration_elec=concentration_elec, H2_FE=H2_FE))


def CO2_conc(temp, fugacity_CO2, conc_ions, h_sechenov):
    """Sechenov-corrected CO2 solubility [mol/m3] (3D/MPNP_CO2ER_pore.py:70-93)."""
    lnK_H_CO2 = 93.4517 * (100 / temp) - 60.2409 + 23.3585 * math.log(temp / 100)
    h_CO2 = h_sechenov["CO2_0"] + h_sechenov["CO2_T"] * (temp - 298.15)
    sechenov = 0.0
    for ion in conc_ions.keys():
        add = (h_sechenov[ion] + h_CO2) * (conc_ions[ion] / 1000)
        sechenov += add
    K_H_CO2 = math.exp(lnK_H_CO2)
    return fugacity_CO2 * K_H_CO2 * 1000 * 10 ** (-sechenov)


def params_3d(concentration_elec=1.0, voltage_multiplier=-1.0, H2_FE=0.05, current_rough=3000.0,
              L=100.0e-9, cation="K", R=5.0e-9, press_gas=1.0, pore_geom_multiplier=1.0,
              porosity_eff=0.5, tortuosity_eff=1.5, constrictivity_eff=0.9,
              params_file="parameters_pore", y_CO2=0.95, electrolyte_flow_geom_multiplier=1.0,
              roughness_factor=150.0, utilities_dir=None, time_step=1.0e-3) -> ProblemParams:
    """Dimensionless groups of the 3D pore problem (3D/MPNP_CO2ER_pore.py:115-324, 469-499)."""
    data = _load_inputs(params_file, utilities_dir)
    rate = {k: float(v) for k, v in data["rate_constants"].items()}
    cat = cation
    species = SPECIES_3D + [cat]
    D = np.array([float(data["diff_coef"]["D_" + i]) for i in species])
    D_eff = (D * porosity_eff * constrictivity_eff * pore_geom_multiplier) / tortuosity_eff ** 2  # 3D:156-158
    n_w_H = float(data["Hydration_number"]["w_H"])
    n_w_cat = float(data["Hydration_number"]["w_" + cat])
    a = np.array([float(data["solv_size"]["a_" + i]) for i in species])
    nc = data["nat_const"]
    farad, k_B, e_0 = float(nc["F"]), float(nc["k_B"]), float(nc["e_0"])
    eps_0, eps_rel, R_gas, N_A = float(nc["eps_0"]), float(nc["eps_rel"]), float(nc["R"]), float(nc["N_A"])
    H_CO2, H_CO, H_H2 = (float(data["Henrys_const"][k]) for k in ("H_CO2", "H_CO", "H_H2"))
    sp = data["sys_params"]
    temp, density_e, viscosity_e = float(sp["T"]), float(sp["density_e"]), float(sp["viscosity_e"])
    L_electrode, vel_e = float(sp["L_electrode"]), float(sp["vel_e"])
    A_cross_e, L_cross_e = float(sp["A_cross_e"]), float(sp["L_cross_e"])
    sc = data["sechonov_const"]
    if "h_ion_" + cat not in sc:
        raise KeyError("h_ion_" + cat)         # same failure as 3D:210 (SURVEY finding 5)
    h_sechenov = {"OH": float(sc["h_ion_OH"]), "HCO3": float(sc["h_ion_HCO3"]),
                  "CO32": float(sc["h_ion_CO32"]), cat: float(sc["h_ion_" + cat]),
                  "CO2_0": float(sc["h_CO2_0"]), "CO2_T": float(sc["h_CO2_T"])}

    y_CO = 0.9 * (1 - y_CO2)
    y_H2 = 1 - y_CO2 - y_CO
    fugacity_CO2 = y_CO2 * press_gas

    bulk = _load_inputs("bulk_soln_" + str(concentration_elec) + "KHCO3", utilities_dir)
    pre = bulk["bulk_conc_pre_CO2"]                # 3D reads the pre-CO2 block (3D:238)
    c0 = np.array([_bulk_conc(pre, i, cat, True) for i in species])
    z = np.array([1, -1, -1, -2, 0, 0, 0, 1], dtype=np.float64)   # 3D:233-234

    eq_conc_CO2 = H_CO2 * press_gas * y_CO2 * density_e    # 3D:253-255
    eq_conc_CO = H_CO * press_gas * y_CO * density_e
    eq_conc_H2 = H_H2 * press_gas * y_H2 * density_e
    c0[5] = 0.01 * eq_conc_CO                              # 3D:258-259
    c0[6] = 0.01 * eq_conc_H2
    eq_scaled = np.array([eq_conc_CO2 / c0[4], eq_conc_CO / c0[5], eq_conc_H2 / c0[6]])

    thermal_voltage = (k_B * temp) / e_0
    time_constant = L ** 2 / D_eff[3]                      # 3D:270
    scale_R = (L ** 2) / (D_eff * c0)                      # 3D:277
    q = (farad ** 2 * L ** 2) / (eps_0 * R_gas * temp)     # 3D:280
    scale_vol = a ** 3 * c0 * N_A                          # 3D:287
    J_prefactor = L / (D_eff * c0)                         # 3D:295

    Re = (density_e * (vel_e / A_cross_e) * L_electrode * electrolyte_flow_geom_multiplier) / viscosity_e
    Sc = viscosity_e / (density_e * D)
    Sh = 1.017 * ((L_electrode * 2 / L_cross_e) * Re * Sc) ** (1.0 / 3)
    k_elec = (D / L_electrode) * Sh

    CO_FE = 1 - H2_FE
    current_planar = current_rough / roughness_factor
    # wall fluxes and pore-exit Robin coefficients (3D:474-499).  As executed by the
    # reference these never enter the residual (SURVEY finding 3); kept for --intended_bcs.
    J_wall = np.zeros(8)
    J_wall[4] = (J_prefactor[4] / farad) * current_planar * 0.5 * (CO_FE)
    J_wall[5] = (J_prefactor[5] / farad) * current_planar * 0.5 * (CO_FE) * (-1.0)
    J_wall[6] = (J_prefactor[6] / farad) * current_planar * 0.5 * (H2_FE) * (-1.0)
    J_wall[1] = (J_prefactor[1] / farad) * current_planar * (-1.0)
    k_exit = J_prefactor * k_elec * c0

    dt = time_step / time_constant                         # 3D:362
    kappa = 1.0 / dt                                       # 3D:534  (u-u_n)/del_t
    return ProblemParams(
        dim=3, species=species, z=z, c0=c0, D=D_eff, nu=scale_vol.copy(), s=scale_R[:5].copy(),
        rate=rate, q=q, eps_w=eps_rel, n_water_H=n_w_H, n_water_cat=n_w_cat, kappa=kappa,
        V=float(voltage_multiplier), jflux=np.zeros(8), length=L, thermal_voltage=thermal_voltage,
        time_constant=time_constant, dt_scaled=dt,
        extras=dict(R=R, aspect_pore=R / L, eq_conc=(eq_conc_CO2, eq_conc_CO, eq_conc_H2),
                    eq_scaled=eq_scaled, temp=temp, fugacity_CO2=fugacity_CO2, h_sechenov=h_sechenov,
                    Re=Re, Sc=Sc, Sh=Sh, k_elec=k_elec, J_wall=J_wall, k_exit=k_exit,
                    scale_R=scale_R, scale_vol=scale_vol, J_prefactor=J_prefactor, D_free=D,
                    cation=cat, concentration_elec=concentration_elec, farad=farad))


def sechenov_co2_scaled(p: ProblemParams, med_OH, med_HCO3, med_CO32, med_cat) -> float:
    """Scaled CO2 entry value from median scaled ion concentrations
    (3D/MPNP_CO2ER_pore.py:817-835)."""
    cat = p.species[-1]
    conc_ions = {"OH": med_OH * p.c0[1], "HCO3": med_HCO3 * p.c0[2],
                 "CO32": med_CO32 * p.c0[3], cat: med_cat * p.c0[-1]}
    eq = CO2_conc(p.extras["temp"], p.extras["fugacity_CO2"], conc_ions, p.extras["h_sechenov"])
    return eq / p.c0[4]


def mesh_name_1d(L_n: float, mesh_structure: str = "variable") -> str:
    """Mesh file stem chosen by 1D/MPNP_CO2ER_EDL.py:216-234 (+ the 200 um extension,
    SURVEY App. H)."""
    L_sys = int(L_n * 1.0e+6)
    if mesh_structure == "variable":
        number = {1: 1090, 5: 1490, 10: 1990, 50: 5990, 200: 4998}.get(L_sys)
        if number is None:
            raise ValueError(f"no variable mesh for L_n={L_n}")
        return f"1D_variable_{L_sys}um_mesh_{number}"
    if mesh_structure == "uniform":
        return "1D_uniform_mesh_1000"
    raise ValueError(mesh_structure)


def mesh_name_3d(L: float, R: float) -> str:
    """3D/MPNP_CO2ER_pore.py:329-332 (int() truncation replicated, SURVEY finding 6)."""
    return "L_" + str(int(L * 1e+9)) + "_R_" + str(int(R * 1e+9))
