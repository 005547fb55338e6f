"""Pins of the vlasov-1d2v oracle (oracle/vlasov1d2v.py) on the reference's own identities, restated at reduced size:

* tests/test_vlasov1d2v/test_1d_limit.py:21-46 -- for a v_perp-separable initial condition the 2V marginal F(t, x, v_par)
  reproduces the vlasov-1d solve (distribution to 1e-9 of its maximum, field to 1e-8), dougherty and dougherty_nodrag,
  sixth-order, driven, self-consistent beta on;
* tests/test_vlasov1d2v/test_1d_limit.py:49-85 -- the cumulative diagnostics telescope to F(t) - F(0), and the collision
  accumulator carries no density.
"""

import numpy as np
import pytest

from oracle import vlasov1d as O
from oracle import vlasov1d2v as O2


def base_config(nx=16, nv=64, nvperp=8, tmax=10.0, dt=0.25, nu=1.0e-3, fp_type="dougherty", a0=1.0e-2, time="sixth"):
    """Small pawl-style deck (the layout of the reference's tests/test_vlasov1d2v/utils.py:8-101)."""
    env = {"baseline": 1.0, "bump_or_trough": "bump", "center": 0.0, "rise": 25.0, "slope": 0.0, "bump_height": 0.0,
           "width": 1.0e5}
    return {
        "units": {"normalizing_temperature": "2000eV", "normalizing_density": "1.5e21/cc"},
        "density": {"quasineutrality": True,
                    "species-background": {"noise_seed": 420, "noise_type": "gaussian", "noise_val": 0.0, "v0": 0.0,
                                           "T0": 1.0, "m": 2.0, "basis": "uniform", "baseline": 1.0,
                                           "bump_or_trough": "bump", "center": 0.0, "rise": 25.0, "bump_height": 0.0,
                                           "width": 1.0e5}},
        "grid": {"dt": dt, "nv": nv, "nx": nx, "tmin": 0.0, "tmax": tmax, "vmax": 6.4, "xmax": 20.94, "xmin": 0.0,
                 "nvperp": nvperp, "vperp_max": 6.4},
        "save": {},
        "solver": "vlasov-1d2v",
        "drivers": {"ex": {"0": {"params": {"a0": a0, "k0": 0.3, "w0": 1.1598, "dw0": 0.0},
                                 "envelope": {"time": {"center": 4.0, "rise": 0.5, "width": 6.0},
                                              "space": {"center": 0.0, "rise": 10.0, "width": 4.0e6}}}}, "ey": {}},
        "diagnostics": {"diag-vlasov-cumulative": True, "diag-fp-cumulative": True, "diag-vlasov-dfdt": False,
                        "diag-fp-dfdt": False},
        "terms": {"field": "poisson", "edfdv": "exponential", "time": time,
                  "fokker_planck": {"is_on": nu > 0.0, "type": fp_type,
                                    "self_consistent_beta": {"enabled": True, "max_steps": 3},
                                    "time": dict(env, baseline=nu), "space": dict(env, baseline=1.0)},
                  "krook": {"is_on": False, "time": dict(env), "space": dict(env)}},
    }


@pytest.mark.parametrize("fp_type", ["dougherty", "dougherty_nodrag"])
def test_oracle_marginal_matches_the_1d_oracle(fp_type):
    deck = base_config(fp_type=fp_type)
    cfg2 = O2.build_cfg(deck)
    cfg1 = O.build_cfg(deck)
    w = cfg2["grid"]["species_grids"]["electron"]["wperp"]
    # the initial marginal is the 1-D initial condition (helpers.py:38-43)
    y2, y1 = O2.init_state(cfg2), O.init_state(cfg1)
    assert np.max(np.abs(O2.marginal(y2["electron"], w) - y1["electron"])) <= 1e-14 * np.max(y1["electron"])
    vf2, vf1 = O2.VlasovMaxwell2V(cfg2), O.VlasovMaxwell(cfg1)
    dt = cfg1["grid"]["dt"]
    emax = 0.0
    for n in range(30):
        y2, y1 = vf2(n * dt, y2, None), vf1(n * dt, y1, None)
        emax = max(emax, np.max(np.abs(y1["e"])))
        assert np.max(np.abs(y2["e"] - y1["e"])) <= 1e-8 * max(emax, 1e-12)
    assert emax > 1e-4, "driver did not couple"
    F = O2.marginal(y2["electron"], w)
    assert np.max(np.abs(F - y1["electron"])) / np.max(np.abs(y1["electron"])) < 1e-9


def test_oracle_cumulative_diags_telescope():
    cfg = O2.build_cfg(base_config(nx=8, nv=64, nvperp=4, nu=1e-2, time="leapfrog"))
    g = cfg["grid"]["species_grids"]["electron"]
    y0 = O2.init_state(cfg)
    y, _ = O2.run(cfg, 20)
    F0, F1 = O2.marginal(y0["electron"], g["wperp"]), O2.marginal(y["electron"], g["wperp"])
    acc = y["diag-vlasov-cumulative"] + y["diag-fp-cumulative"]
    scale = np.max(np.abs(F1 - F0))
    assert scale > 0
    assert np.max(np.abs(acc - (F1 - F0))) / scale < 1e-10
    assert np.max(np.abs(np.sum(y["diag-fp-cumulative"], axis=-1) * g["dv"])) < 1e-10
