"""Extract the reference's golden grid / initial-state vectors into compact fixtures.

Run in the build container (needs /root/reference, which does NOT exist on the GPU box):

    python tests/golden/make_golden.py

Source: /root/reference/tests/test_vlasov1d/test_config_regression/*_array_config.yml
(pytest-regressions dumps of ``array_config.pkl`` written by the reference's
``ergoExo._setup_``, rounded to 14 significant figures) together with the input decks
/root/reference/tests/test_vlasov1d/configs/*.yaml.  Nothing is computed here: values are
copied verbatim into ``tests/golden/<name>.npz`` and the deck into ``<name>.yaml``.
"""

import shutil
from pathlib import Path

import numpy as np
import yaml

REF = Path("/root/reference/tests/test_vlasov1d")
OUT = Path(__file__).parent
NAMES = ["resonance", "fokker_planck_conservation", "multispecies_ion_acoustic"]


def main():
    for name in NAMES:
        with open(REF / "test_config_regression" / f"{name}_array_config.yml") as fh:
            d = yaml.safe_load(fh)
        g = d["grid"]
        arrs = {}
        for k in ["x", "x_a", "t", "kx", "kxr", "one_over_kx", "one_over_kxr", "ion_charge", "n_prof_total"]:
            arrs[f"grid.{k}"] = np.asarray(g[k], dtype=np.float64)
        for k in ["beta", "dt", "dx", "tmax", "tmin", "xmax", "xmin"]:
            arrs[f"grid.{k}"] = np.float64(g[k])
        for k in ["nt", "nx", "max_steps"]:
            arrs[f"grid.{k}"] = np.int64(g[k])
        for sp, sg in g["species_grids"].items():
            for k in ["v", "kv", "kvr", "one_over_kv", "one_over_kvr"]:
                arrs[f"species_grids.{sp}.{k}"] = np.asarray(sg[k], dtype=np.float64)
            for k in ["dv", "vmax", "vmin"]:
                arrs[f"species_grids.{sp}.{k}"] = np.float64(sg[k])
            arrs[f"species_grids.{sp}.nv"] = np.int64(sg["nv"])
        for sp, p in g["species_params"].items():
            for k in ["T0", "charge", "charge_to_mass", "mass"]:
                arrs[f"species_params.{sp}.{k}"] = np.float64(p[k])
        for sp, (n_prof, f0, vax) in g["species_distributions"].items():
            arrs[f"species_distributions.{sp}.n_prof"] = np.asarray(n_prof, dtype=np.float64)
            arrs[f"species_distributions.{sp}.f0"] = np.asarray(f0, dtype=np.float64)
            arrs[f"species_distributions.{sp}.v"] = np.asarray(vax, dtype=np.float64)
        np.savez_compressed(OUT / f"{name}.npz", **arrs)
        shutil.copyfile(REF / "configs" / f"{name}.yaml", OUT / f"{name}.yaml")
        print(name, {k: np.shape(v) for k, v in arrs.items() if np.ndim(v) > 0})


if __name__ == "__main__":
    main()
