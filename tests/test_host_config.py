"""Host logic without a GPU: the deck completion of the product (adept_b200.config.build_cfg, mirroring
adept/_vlasov1d/modules.py:190-317 and helpers.py:37-161) against the oracle's and against the reference's golden
arrays, and the save-time axes."""

from copy import deepcopy
from pathlib import Path

import numpy as np
import pytest
import yaml

from adept_b200.config import build_cfg
from oracle import vlasov1d as O

GOLD = Path(__file__).parent / "golden"
DECKS = ["epw", "resonance", "fokker_planck_conservation", "multispecies_ion_acoustic"]


def load(name):
    with open(GOLD / f"{name}.yaml") as fh:
        return yaml.safe_load(fh)


@pytest.mark.parametrize("name", DECKS)
def test_host_cfg_equals_oracle_cfg(name):
    deck = load(name)
    cfg, grid = build_cfg(deepcopy(deck))
    ref = O.build_cfg(deepcopy(deck))
    g, r = cfg["grid"], ref["grid"]
    for key in ("nx", "nt", "max_steps"):
        assert int(g[key]) == int(r[key]), key
    for key in ("dt", "dx", "tmax", "beta"):
        assert float(g[key]) == float(r[key]), key
    for key in ("x", "x_a", "kx", "kxr", "one_over_kx", "one_over_kxr", "t", "ion_charge", "n_prof_total"):
        np.testing.assert_array_equal(np.asarray(g[key]), np.asarray(r[key]), err_msg=key)
    assert list(g["species_grids"]) == list(r["species_grids"])
    for s in g["species_grids"]:
        for key in ("v", "kv", "kvr", "one_over_kv", "one_over_kvr"):
            np.testing.assert_array_equal(np.asarray(g["species_grids"][s][key]), np.asarray(r["species_grids"][s][key]))
        assert float(g["species_grids"][s]["dv"]) == float(r["species_grids"][s]["dv"])
        assert g["species_params"][s] == r["species_params"][s]
        np.testing.assert_array_equal(g["species_distributions"][s][1], r["species_distributions"][s][1])


@pytest.mark.parametrize("name", ["resonance", "fokker_planck_conservation", "multispecies_ion_acoustic"])
def test_host_cfg_matches_reference_golden_arrays(name):
    """The reference's regression fixtures (tests/test_vlasov1d/test_config_regression/*_array_config.yml, committed as
    tests/golden/*.npz) pin grids and the full initial distribution to 14 significant figures."""
    gold = np.load(GOLD / f"{name}.npz")
    cfg, _ = build_cfg(load(name))
    g = cfg["grid"]
    checked = 0
    for key in gold.files:
        parts = key.split(".")
        if parts[0] == "grid" and parts[1] in ("x", "x_a", "kx", "kxr", "one_over_kx", "one_over_kxr", "t", "ion_charge",
                                               "n_prof_total"):
            np.testing.assert_allclose(np.asarray(g[parts[1]]), gold[key], rtol=1e-13, atol=1e-300, err_msg=key)
            checked += 1
        elif parts[0] == "species_grids" and parts[2] in ("v", "kv", "kvr", "one_over_kv", "one_over_kvr"):
            np.testing.assert_allclose(np.asarray(g["species_grids"][parts[1]][parts[2]]), gold[key], rtol=1e-13,
                                       atol=1e-300, err_msg=key)
            checked += 1
        elif parts[0] == "species_distributions" and parts[2] == "f0":
            np.testing.assert_allclose(g["species_distributions"][parts[1]][1], gold[key], rtol=1e-13, atol=1e-300)
            checked += 1
    assert checked >= 12, gold.files


def test_save_axis_matches_oracle():
    from adept_b200.module import save_axis

    cfg, grid = build_cfg(load("epw"))
    ref = O.build_cfg(load("epw"))
    for tcfg in ({"nt": 11}, {"nt": 7, "tmin": 1.0, "tmax": 5.0}):
        np.testing.assert_array_equal(save_axis(tcfg, grid), O.save_axis(tcfg, ref["grid"]))


def test_stochastic_driver_matches_oracle_realisation():
    """simulation.py:95-148: same numpy Generator stream, same OU series; the host hands every mode to the device as
    an amplitude/phase pair (A sin(kx - phase)), which must reproduce ar cos(kx) - ai sin(kx)."""
    from adept_b200 import pushers

    class G:
        xmin, xmax, tmin, tmax = 0.0, 20.94, 0.0, 50.0

    sc = {"modes": [1, 2, 5], "amplitude": 1e-3, "tau": 3.0, "seed": 7}
    o = O.StochasticDriver(sc, G.xmin, G.xmax, G.tmin, G.tmax)
    h = pushers.StochasticDriver(sc, G)
    np.testing.assert_array_equal(o.amp_real, h.amp_real)
    np.testing.assert_array_equal(o.t_grid, h.t_grid)
    x = np.linspace(0.1, 20.8, 64)
    for t in (0.0, 0.37, 12.3, 49.9):
        tot = np.zeros_like(x)
        for d in h.modes():
            tot += d.envelope(x, t) * (d.w0 + d.dw0) * d.a0 * np.sin(d.k0 * x - d.phase(t))
        ref = o(t, x)
        assert np.linalg.norm(tot - ref) <= 1e-14 * np.linalg.norm(ref)
    # stationary RMS of the OU process is `amplitude` per mode (|a_m|^2 averages to amplitude^2)
    long = O.StochasticDriver({"modes": [1], "amplitude": 0.5, "tau": 1.0, "seed": 1}, 0.0, 1.0, 0.0, 4000.0)
    rms = np.sqrt(np.mean(long.amp_real**2 + long.amp_imag**2))
    assert abs(rms / 0.5 - 1.0) < 0.05
    with pytest.raises(ValueError):
        pushers.StochasticDriver({"modes": [0], "amplitude": 1.0, "tau": 1.0}, G)


def test_nonuniform_density_quasineutral_init():
    """tests/test_vlasov1d/test_quasineutral_init.py: with a tanh density profile the static ion background cancels the
    electron charge of the initial state, so the charge density and the Poisson field vanish (atol 1e-6) - for the
    oracle's initialiser and for the host's (adept_b200.config), whose state feeds the kernels."""
    with open(GOLD / "resonance.yaml") as fh:
        deck = yaml.safe_load(fh)
    deck["density"]["species-background"].update(basis="tanh", baseline=1.0, bump_height=0.1, bump_or_trough="bump",
                                                 width=10.0, center=10.0, rise=2.0)
    deck["drivers"]["ex"] = {}
    for build in (lambda d: O.build_cfg(d), lambda d: build_cfg(d)[0]):
        cfg = build(deepcopy(deck))
        g = cfg["grid"]
        ion = np.asarray(g["ion_charge"])
        assert float(np.max(ion) - np.min(ion)) > 0.01  # the background is non-uniform
        f_dict = {name: np.asarray(d[1]) for name, d in g["species_distributions"].items()}
        rho = O.charge_density(f_dict, g["species_grids"], g["species_params"], ion)
        np.testing.assert_allclose(rho, 0.0, atol=1e-6)
        np.testing.assert_allclose(O.poisson(rho, np.asarray(g["one_over_kx"])), 0.0, atol=1e-6)


# ---- tests/test_vlasov1d/test_asymmetric_velocity_grid.py, for the oracle's and the host's config builders ------------
def _assert_grid(sg, vmin, vmax, nv):
    dv = (vmax - vmin) / nv
    v = np.asarray(sg["v"])
    assert sg["nv"] == nv and len(v) == nv
    assert np.isclose(sg["vmin"], vmin) and np.isclose(sg["vmax"], vmax) and np.isclose(sg["dv"], dv)
    assert np.isclose(v[0], vmin + dv / 2.0) and np.isclose(v[-1], vmax - dv / 2.0)
    assert np.allclose(np.diff(v), dv)


BUILDERS = {"oracle": lambda d: O.build_cfg(d), "host": lambda d: build_cfg(d)[0]}


@pytest.mark.parametrize("which", list(BUILDERS))
def test_asymmetric_velocity_grids(which):
    """:45-106: grid-level vmin/vmax, the symmetric default, a per-species override, and the normalisation of f."""
    build = BUILDERS[which]
    with open(GOLD / "resonance.yaml") as fh:
        res = yaml.safe_load(fh)
    with open(GOLD / "multispecies_ion_acoustic.yaml") as fh:
        multi = yaml.safe_load(fh)
    d = deepcopy(res)
    d["grid"].update(vmin=-4.0, vmax=8.0)
    g = build(d)["grid"]
    _assert_grid(g["species_grids"]["electron"], -4.0, 8.0, d["grid"]["nv"])
    np.testing.assert_allclose(np.asarray(g["v"]), np.asarray(g["species_grids"]["electron"]["v"]))
    d = deepcopy(res)
    d["grid"].pop("vmin", None)
    _assert_grid(build(d)["grid"]["species_grids"]["electron"], -d["grid"]["vmax"], d["grid"]["vmax"], d["grid"]["nv"])
    d = deepcopy(multi)
    el = next(s for s in d["terms"]["species"] if s["name"] == "electron")
    ion = next(s for s in d["terms"]["species"] if s["name"] == "ion")
    el.update(vmin=-3.0, vmax=9.0)
    sgs = build(d)["grid"]["species_grids"]
    _assert_grid(sgs["electron"], -3.0, 9.0, el["nv"])
    _assert_grid(sgs["ion"], -ion["vmax"], ion["vmax"], ion["nv"])
    d = deepcopy(res)
    d["grid"].update(vmin=-5.0, vmax=10.0)
    g = build(d)["grid"]
    f = np.asarray(g["species_distributions"]["electron"][1])
    np.testing.assert_allclose(f.sum(axis=1) * g["species_grids"]["electron"]["dv"], 1.0, rtol=1e-3)


def test_collisions_conserve_density_on_asymmetric_grid_oracle():
    """:109-137: Fokker-Planck + Krook of resonance.yaml on vmin = -5, vmax = 8 conserve density to 1e-6."""
    with open(GOLD / "resonance.yaml") as fh:
        d = yaml.safe_load(fh)
    d["grid"].update(vmin=-5.0, vmax=8.0)
    cfg = O.build_cfg(d)
    g = cfg["grid"]
    dv = g["species_grids"]["electron"]["dv"]
    f0 = np.asarray(g["species_distributions"]["electron"][1])
    nu = np.ones(f0.shape[0])
    f1 = O.Collisions(cfg)(nu, nu, f0, g["dt"])
    assert np.all(np.isfinite(f1))
    np.testing.assert_allclose(f1.sum(axis=1) * dv, f0.sum(axis=1) * dv, rtol=1e-6)


@pytest.mark.parametrize("which", list(BUILDERS))
def test_multispecies_and_single_species_state_structures(which):
    """tests/test_vlasov1d/test_multispecies_init.py:11-104: species grids / params / shapes of the two-species deck and
    the backward-compatible single-species structure (grid-level v, ion background = total density profile)."""
    build = BUILDERS[which]
    with open(GOLD / "multispecies_ion_acoustic.yaml") as fh:
        cfg = build(yaml.safe_load(fh))
    g = cfg["grid"]
    assert set(g["species_grids"]) == {"electron", "ion"} and set(g["species_params"]) == {"electron", "ion"}
    assert np.asarray(g["species_distributions"]["electron"][1]).shape == (32, 512)
    assert np.asarray(g["species_distributions"]["ion"][1]).shape == (32, 256)
    eg, ig = g["species_grids"]["electron"], g["species_grids"]["ion"]
    assert eg["nv"] == 512 and eg["vmax"] == 6.4 and len(eg["v"]) == 512
    assert ig["nv"] == 256 and ig["vmax"] == 0.005 and len(ig["v"]) == 256
    ep, ip = g["species_params"]["electron"], g["species_params"]["ion"]
    assert ep["charge"] == -1.0 and ep["mass"] == 1.0 and ep["charge_to_mass"] == -1.0
    assert ip["charge"] == 10.0 and ip["mass"] == 18360.0 and np.isclose(ip["charge_to_mass"], 10.0 / 18360.0)
    assert np.allclose(g["ion_charge"], 0.0)
    with open(GOLD / "resonance.yaml") as fh:
        cfg = build(yaml.safe_load(fh))
    g = cfg["grid"]
    assert np.asarray(g["species_distributions"]["electron"][1]).shape == (g["nx"], g["nv"])
    assert len(g["v"]) == g["nv"]
    assert np.allclose(g["ion_charge"], g["n_prof_total"])
    if which == "oracle":  # the state dict of modules.py:279-317
        y = O.init_state(cfg)
        assert {"electron", "e", "de", "a", "da", "prev_a"} <= set(y)


def test_two_component_species_like_twostream():
    """configs/vlasov-1d/twostream.yaml: one electron species built from two drifting density components
    (helpers.py:136-153 sums them); host and oracle initialisers agree, the two beams are there, density is 1."""
    with open(GOLD / "epw.yaml") as fh:
        deck = yaml.safe_load(fh)
    comp = {"noise_seed": 420, "noise_type": "gaussian", "noise_val": 0.0, "T0": 0.2, "m": 2.0, "basis": "sine",
            "baseline": 0.5, "wavenumber": 0.3}
    deck["density"] = {"quasineutrality": True, "species-electron1": dict(comp, v0=-1.5, amplitude=1.0e-4),
                       "species-electron2": dict(comp, v0=1.5, amplitude=-1.0e-4)}
    deck["grid"].update(nx=64, nv=512)
    host_cfg, _ = build_cfg(deepcopy(deck))
    ora_cfg = O.build_cfg(deepcopy(deck))
    fh_, fo = (np.asarray(c["grid"]["species_distributions"]["electron"][1]) for c in (host_cfg, ora_cfg))
    np.testing.assert_allclose(fh_, fo, rtol=1e-14, atol=1e-300)
    g = host_cfg["grid"]
    v, dv = np.asarray(g["species_grids"]["electron"]["v"]), g["species_grids"]["electron"]["dv"]
    np.testing.assert_allclose(fh_.sum(axis=1) * dv, 1.0, rtol=1e-3)
    row = fh_[0]
    assert abs(v[np.argmax(np.where(v < 0, row, 0))] + 1.5) < 2 * dv and abs(v[np.argmax(np.where(v > 0, row, 0))] - 1.5) < 2 * dv
    assert row[np.argmin(np.abs(v))] < 1e-2 * row.max()  # the beams are separated: exp(-1.5^2 / 0.4) = 3.6e-3
    np.testing.assert_allclose(g["ion_charge"], g["n_prof_total"])
