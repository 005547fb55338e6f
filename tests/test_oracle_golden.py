"""Pin the oracle's grids and initial state against the reference's golden vectors.

Fixtures: tests/golden/*.npz, copied verbatim by tests/golden/make_golden.py from
/root/reference/tests/test_vlasov1d/test_config_regression/*_array_config.yml (14 s.f.).
"""

from pathlib import Path

import numpy as np
import pytest
import yaml

from oracle import vlasov1d as O

GOLD = Path(__file__).parent / "golden"
NAMES = ["resonance", "fokker_planck_conservation", "multispecies_ion_acoustic"]
# golden values are rounded to 14 significant figures (test_config_regression.py:30-43)
RTOL = 6e-14


@pytest.mark.parametrize("name", NAMES)
def test_grid_and_f0_match_reference_golden(name):
    gold = np.load(GOLD / f"{name}.npz")
    with open(GOLD / f"{name}.yaml") as fh:
        cfg = O.build_cfg(yaml.safe_load(fh))
    g = cfg["grid"]
    for k in ["x", "x_a", "t", "kx", "kxr", "one_over_kx", "one_over_kxr", "ion_charge", "n_prof_total"]:
        np.testing.assert_allclose(g[k], gold[f"grid.{k}"], rtol=RTOL, atol=1e-300, err_msg=k)
    for k in ["beta", "dt", "dx", "tmax", "tmin", "xmax", "xmin"]:
        np.testing.assert_allclose(g[k], gold[f"grid.{k}"], rtol=1e-8 if k == "beta" else RTOL, err_msg=k)
    for k in ["nt", "nx", "max_steps"]:
        assert int(g[k]) == int(gold[f"grid.{k}"]), k
    for sp, sg in g["species_grids"].items():
        for k in ["v", "kv", "kvr", "one_over_kv", "one_over_kvr", "dv", "vmax", "vmin"]:
            np.testing.assert_allclose(sg[k], gold[f"species_grids.{sp}.{k}"], rtol=RTOL, atol=1e-300, err_msg=k)
        assert sg["nv"] == int(gold[f"species_grids.{sp}.nv"])
        for k in ["T0", "charge", "charge_to_mass", "mass"]:
            np.testing.assert_allclose(g["species_params"][sp][k], gold[f"species_params.{sp}.{k}"], rtol=RTOL)
        n_prof, f0, vax = g["species_distributions"][sp]
        np.testing.assert_allclose(n_prof, gold[f"species_distributions.{sp}.n_prof"], rtol=RTOL)
        np.testing.assert_allclose(vax, gold[f"species_distributions.{sp}.v"], rtol=RTOL)
        # f0 spans 1e-9 .. 0.4; 14 s.f. rounding is relative
        np.testing.assert_allclose(f0, gold[f"species_distributions.{sp}.f0"], rtol=RTOL, atol=1e-300)


def test_state_layout_matches_reference():
    """modules.py:291-316: state keys and shapes."""
    with open(GOLD / "epw.yaml") as fh:
        cfg = O.build_cfg(yaml.safe_load(fh))
    y = O.init_state(cfg)
    nx, nv = cfg["grid"]["nx"], cfg["grid"]["nv"]
    assert y["electron"].shape == (nx, nv)
    assert y["e"].shape == y["de"].shape == (nx,)
    assert y["a"].shape == y["da"].shape == y["prev_a"].shape == (nx + 2,)
    assert y["diag-vlasov-dfdt"].shape == y["diag-fp-dfdt"].shape == (nx, nv)
