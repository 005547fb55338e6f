"""GPU parity of the full vlasov-1d step (VlasovMaxwell) against the numpy oracle, through the drop-in classes.

Decks: BASELINE.json configs[0] (configs/vlasov-1d/epw.yaml: sixth + cubic-spline + Dougherty + dfdt diags, 32x256),
configs[1] (EPW + collisions, 64x512, leapfrog + exponential), and the reference's test decks in tests/golden/.
Bars: per-step relative L2 <= 1e-12 on every state entry; field-energy history to 1e-9 (relative to its peak).
"""

from copy import deepcopy
from pathlib import Path

import numpy as np
import pytest
import torch
import yaml

from oracle import vlasov1d as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def rel_l2(a, b):
    a, b = np.asarray(a), np.asarray(b)
    nb = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / nb) if nb > 0 else float(np.linalg.norm(a.ravel()))


def load(name):
    with open(GOLD / f"{name}.yaml") as fh:
        return yaml.safe_load(fh)


def c2_deck():
    """configs[1]: EPW with Lenard-Bernstein/Dougherty collisions, nx=64, nv=512 (from the stock epw.yaml)."""
    d = load("epw")
    d["grid"].update(nx=64, nv=512)
    d["terms"].update(time="leapfrog", edfdv="exponential")
    d["terms"]["fokker_planck"]["time"]["baseline"] = 1.0e-3
    d["diagnostics"] = {"diag-vlasov-dfdt": False, "diag-fp-dfdt": False}
    return d


def variants():
    out = {"C1-epw": load("epw"), "C2-epw-fp": c2_deck()}
    d = c2_deck()
    d["terms"]["fokker_planck"]["type"] = "chang_cooper"
    d["terms"]["krook"]["is_on"] = True
    d["terms"]["krook"]["time"]["baseline"] = 1e-3
    out["C2-cc-lb-krook"] = d
    d = c2_deck()
    d["terms"].update(field="ampere")
    d["grid"]["dt"] = 0.025
    out["C2-ampere"] = d
    d = c2_deck()  # Hamiltonian Ampere solver (leapfrog only, single species)
    d["terms"].update(field="hampere")
    out["C2-hampere"] = d
    d = c2_deck()  # Hou-Li spectral filter after the collisions
    d["terms"]["hou_li_filter"] = {"is_on": True, "alpha": 36.0, "order": 8}
    out["C2-houli"] = d
    out["resonance"] = load("resonance")
    out["fp-conservation"] = load("fokker_planck_conservation")
    d = load("multispecies_ion_acoustic")
    out["multispecies"] = d
    d = c2_deck()  # driven transverse wave: wave solver + ponderomotive force on
    d["drivers"]["ey"] = {"0": deepcopy(d["drivers"]["ex"]["0"])}
    d["drivers"]["ey"]["0"]["params"].update(a0=1.0e-2, k0=1.0, w0=2.0)
    d["drivers"]["ey"]["0"]["envelope"]["time"].update(center=1.0, width=2.0, rise=0.2)
    out["C2-ey-wave"] = d
    # large single grids: TMA x-advection with fused charge density, ONE fused field launch (driver + ponderomotive
    # force + density + Poisson), fused v-push + collisions
    d = c2_deck()
    d["grid"].update(nx=1024, nv=1024)
    out["L-1024x1024-fp"] = d
    d = c2_deck()  # collisions strong enough that the reduced tridiagonal system needs every PCR step
    d["grid"].update(nx=1024, nv=512)
    d["terms"]["fokker_planck"]["time"]["baseline"] = 0.5
    out["L-1024x512-strong-fp"] = d
    d = c2_deck()  # configs/vlasov-1d/wavepacket.yaml in miniature: nv = 3 x 2^k, cubic-spline v-push, FP + Krook
    d["grid"].update(nx=256, nv=384)
    d["terms"].update(edfdv="cubic-spline")
    d["terms"]["krook"]["is_on"] = True
    d["terms"]["krook"]["time"]["baseline"] = 1e-3
    out["wavepacket-like-256x384"] = d
    d = c2_deck()  # configs/vlasov-1d/srs-debug-small.yaml's grid: nx = 1028 = 4 x 257 (chirp-z transforms in x)
    d["grid"].update(nx=1028, nv=64)
    out["srs-debug-small-like-1028x64"] = d
    d = c2_deck()  # non-power-of-two in both directions, sixth-order integrator
    d["grid"].update(nx=96, nv=384)
    d["terms"].update(time="sixth")
    out["any-length-96x384-sixth"] = d
    d = c2_deck()  # Chang-Cooper weighting in the fused v-push + collision kernel
    d["grid"].update(nx=1024, nv=1024)
    d["terms"]["fokker_planck"]["type"] = "chang_cooper_dougherty"
    out["L-1024x1024-cc"] = d
    d = c2_deck()  # sixth-order integrator: the fused field launch serves substeps 2..6 (density from the x-pushes)
    d["grid"].update(nx=2048, nv=512)
    d["terms"].update(time="sixth")
    out["L-2048x512-sixth"] = d
    d = load("multispecies_ion_acoustic")  # two species with different nv feed one fused field solve
    d["grid"].update(nx=1024)
    out["L-multispecies-1024"] = d
    d = c2_deck()  # no Ex driver at all on a large grid: the fused field launch must still leave dex = 0
    d["grid"].update(nx=1024, nv=512)
    d["drivers"]["ex"] = {}
    d["density"]["species-background"].update(basis="sine", baseline=1.0, amplitude=1.0e-2, wavenumber=0.3)
    out["L-1024x512-no-driver"] = d
    for nx in (64, 1024):  # configs/vlasov-1d/iaw-turbulence.yaml in miniature: kinetic ions, Boltzmann electrons,
        d = c2_deck()      # cubic-spline v-push, sixth-order integrator, stochastic box-scale forcing, no collisions
        d["grid"].update(nx=nx, nv=256, xmax=45.8, dt=0.25)
        d["density"] = {"quasineutrality": True,
                        "species-ion-background": {"noise_seed": 416, "noise_type": "gaussian", "noise_val": 0.0,
                                                   "v0": 0.0, "T0": 1.0, "m": 2.0, "basis": "sine", "baseline": 1.0,
                                                   "amplitude": 1.0e-3, "wavenumber": 2 * np.pi / 45.8}}
        d["drivers"] = {"ex": {}, "ey": {}, "ex_stochastic": {"modes": [1], "amplitude": 2.0e-3, "tau": 45.8, "seed": 42}}
        d["terms"].update(field="poisson-boltzmann", boltzmann_electrons={"Te": 0.05}, edfdv="cubic-spline", time="sixth",
                          species=[{"name": "ion", "charge": 1.0, "mass": 1.0, "vmax": 6.4, "nv": 256,
                                    "density_components": ["species-ion-background"]}])
        d["terms"]["fokker_planck"]["is_on"] = False
        out[f"iaw-like-{nx}x256"] = d
    d = deepcopy(out["iaw-like-64x256"])  # configs/vlasov-1d/iaw-turbulence-big-bench.yaml's x-grid: nx = 17280 = 128 x 135
    d["grid"].update(nx=17280, nv=64, xmax=966.0)
    d["density"]["species-ion-background"]["wavenumber"] = 2 * np.pi / 966.0
    d["drivers"]["ex_stochastic"]["tau"] = 966.0
    d["terms"]["species"][0]["nv"] = 64
    out["iaw-big-bench-like-17280x64"] = d
    d = deepcopy(d)  # the same grid with the spectral v-push and leapfrog (scratch = f_tmp allocated for the x-push)
    d["terms"].update(edfdv="exponential", time="leapfrog")
    out["iaw-big-leapfrog-exp-17280x64"] = d
    d = c2_deck()  # self-consistent beta (Newton on the discrete temperature) in a driven step
    d["terms"]["fokker_planck"].update(type="chang_cooper_dougherty",
                                       self_consistent_beta={"enabled": True, "max_steps": 3})
    d["terms"]["fokker_planck"]["time"]["baseline"] = 0.1
    out["C2-cc-sc-beta"] = d
    d = c2_deck()  # Ornstein-Uhlenbeck forcing of three box modes on top of the deterministic driver
    d["drivers"]["ex_stochastic"] = {"modes": [1, 2, 5], "amplitude": 2.0e-3, "tau": 1.5, "seed": 7}
    out["C2-ex-stochastic"] = d
    d = c2_deck()
    d["terms"].update(time="sixth")
    d["drivers"]["ex_stochastic"] = {"modes": [3], "amplitude": 1.0e-3, "tau": 0.4, "dt_update": 0.05}
    out["C2-ex-stochastic-sixth"] = d
    d = c2_deck()  # the same on a large grid: the deck leaves the fused v-row kernel for the general collision kernel
    d["grid"].update(nx=1024, nv=1024)
    d["terms"]["fokker_planck"]["self_consistent_beta"] = {"enabled": True, "max_steps": 2}
    out["L-1024x1024-sc-beta"] = d
    return out


VARIANTS = variants()


@pytest.mark.parametrize("name", list(VARIANTS))
def test_step_by_step_parity(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200.module import Vlasov1D

    deck = VARIANTS[name]
    sim = Vlasov1D(deck)
    cfg = O.build_cfg(deck)
    vf = O.VlasovMaxwell(cfg)
    y = O.init_state(cfg)
    dt = cfg["grid"]["dt"]
    assert abs(dt - sim.grid.dt) == 0.0
    nsteps = 12
    # start the comparison where the driver is on so that every term is exercised
    t0 = 30.0 if (cfg["drivers"].get("ex") and not cfg["drivers"].get("ey")) else 0.0
    sim.t, sim.step_index = t0, int(round(t0 / dt))
    g_ = cfg["grid"]
    # E is the k-space integral of rho = sum_s q_s n_s + n_ion, a difference of O(1) moments: two correct
    # implementations differ by a few eps * |q n| / k_min in E whatever its size, so E gets an absolute floor.
    rho_scale = sum(abs(g_["species_params"][s]["charge"]) * np.max(np.abs(d[0]))
                    for s, d in g_["species_distributions"].items())
    k1 = 2 * np.pi / (g_["xmax"] - g_["xmin"])
    e_floor = 50 * np.finfo(float).eps * rho_scale * max(1.0, 1.0 / k1)
    worst, worst_abs = {}, {}
    for n in range(nsteps):
        t = t0 + n * dt
        # per-step parity: both sides start each step from the oracle's state
        sim.state = {k: torch.as_tensor(v, device="cuda").contiguous() for k, v in y.items()}
        y_gpu = sim.vector_field(t, sim.state, None)
        y = vf(t, y, None)
        assert set(y_gpu) == set(y)
        for k in y:
            g = y_gpu[k].cpu().numpy()
            assert g.shape == y[k].shape, k
            if k.startswith("diag-"):
                # (f1 - f0)/dt amplifies rounding of f by 1/dt: compare against the scale of f/dt
                err = np.linalg.norm(g - y[k]) / (np.linalg.norm(y["electron"]) / dt)
            elif k == "e":
                err = max(0.0, np.max(np.abs(g - y[k])) - e_floor) / max(np.max(np.abs(y[k])), 1e-300)
            else:
                err = rel_l2(g, y[k])
            worst[k] = max(worst.get(k, 0.0), err)
    for k, err in worst.items():
        assert err <= 1e-12, (name, k, err)
    if name == "C2-ey-wave":
        assert np.max(np.abs(y["a"])) > 0 and np.max(np.abs(y["e"])) > 0  # the wave + ponderomotive path really ran


@pytest.mark.parametrize("name", ["C2-epw-fp", "C1-epw"])
def test_free_running_energy_history(name):
    """Run both sides freely (no re-synchronisation) and compare the field-energy history (diagnostic parity)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200.module import Vlasov1D, default_scalars

    deck = deepcopy(VARIANTS[name])
    deck["grid"]["tmax"] = 60.0
    sim = Vlasov1D(deck)
    cfg = O.build_cfg(deck)
    nsteps = 500
    ts = cfg["grid"]["dt"] * np.arange(1, nsteps + 1)
    _, ref = O.run(cfg, nsteps=nsteps, save={"s": (ts, O.default_scalars)})
    _, got = sim.run(nsteps=nsteps, save={"s": (ts, default_scalars)})
    for key in ["mean_e2", "mean_field_energy", "mean_kinetic_energy", "mean_total_energy", "mean_n_electron",
                "mean_P_electron", "mean_f2_electron", "mean_-flogf_electron"]:
        r = np.array([s[key] for s in ref["s"]])
        g = np.array([float(s[key]) for s in got["s"]])
        assert len(r) == len(g) == nsteps
        np.testing.assert_allclose(g, r, rtol=1e-9, atol=1e-9 * np.max(np.abs(r)), err_msg=key)
    f_ref = O.run(cfg, nsteps=0)[0]  # shapes only
    assert sim.state["electron"].shape == f_ref["electron"].shape


def test_c3_full_size_free_running_parity():
    """BASELINE.json configs[2] at its FULL size (4096 x 4096, the bench workload, two launches per step: x-advection
    with the field solve in its tail, fused v-push + collisions): three free-running steps from t = 30 (driver on)
    against the oracle; every state entry to 1e-12, the field energy to 1e-9."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sys

    sys.path.insert(0, str(GOLD.parents[1]))
    from bench import c3_deck

    from adept_b200.module import Vlasov1D

    deck = c3_deck(4096, 4096)
    sim = Vlasov1D(deepcopy(deck))
    cfg = O.build_cfg(deepcopy(deck))
    vf = O.VlasovMaxwell(cfg)
    y = O.init_state(cfg)
    dt = cfg["grid"]["dt"]
    t0 = 30.0
    sim.t, sim.step_index = t0, int(round(t0 / dt))
    for n in range(3):
        y = vf(t0 + n * dt, y, None)
        sim.step()
    for k in ("electron", "de", "a", "prev_a"):
        assert rel_l2(sim.state[k].cpu().numpy(), y[k]) <= 1e-12, k
    e_gpu = sim.state["e"].cpu().numpy()
    assert np.max(np.abs(e_gpu - y["e"])) <= 1e-12 * max(np.max(np.abs(y["e"])), 1e-3)
    assert abs(np.mean(e_gpu**2) - np.mean(y["e"] ** 2)) <= 1e-9 * np.mean(y["e"] ** 2)
    from adept_b200 import ops

    assert ops.LAUNCHES > 0


def test_iaw_dispersion_boltzmann_on_gpu():
    """The reference's tests/test_vlasov1d/test_boltzmann_electrons.py:109-135 on the GPU path: kinetic ions with
    Boltzmann electrons, 8000 sixth-order steps; the ion-acoustic frequency matches the dispersion relation to 5 %, and
    the density history matches the oracle's first 200 steps to 1e-9."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from test_oracle_operators import boltzmann_iaw_deck, iaw_expected_omega, measure_frequency

    from adept_b200.module import Vlasov1D

    deck = boltzmann_iaw_deck()
    sim = Vlasov1D(deepcopy(deck))
    dv = float(sim.cfg["grid"]["species_grids"]["ion"]["dv"])
    dt = sim.grid.dt
    hist = [sim.state["ion"].sum(dim=1) * dv]
    for n in range(8000):
        sim.step()
        if (n + 1) % 10 == 0:
            hist.append(sim.state["ion"].sum(dim=1) * dv)
    n_hist = torch.stack(hist).cpu().numpy()
    t_hist = np.arange(len(hist)) * 10 * dt
    want = iaw_expected_omega(deck)
    got = measure_frequency(n_hist, t_hist, want)
    np.testing.assert_allclose(got, want, rtol=0.05)
    # and against the oracle over the first 200 steps
    cfg = O.build_cfg(deepcopy(deck))
    vf = O.VlasovMaxwell(cfg)
    y = O.init_state(cfg)
    for n in range(200):
        y = vf(n * dt, y, None)
        if (n + 1) % 10 == 0:
            ref = np.sum(y["ion"], axis=1) * dv
            assert np.max(np.abs(n_hist[(n + 1) // 10] - ref)) <= 1e-9 * np.max(np.abs(ref - 1.0)) + 1e-13


def test_ex_driver_quiver_on_gpu():
    """tests/test_vlasov1d/test_ex_driver_quiver.py:146-203 on the GPU path: the quiver velocity amplitude of electrons
    in the driver field is a0 to 5 %."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from test_oracle_operators import ex_quiver_deck, quiver_amplitude

    from adept_b200.module import Vlasov1D

    deck, a0, k0, w0 = ex_quiver_deck()
    sim = Vlasov1D(deepcopy(deck))
    sg = sim.cfg["grid"]["species_grids"]["electron"]
    v, dv, dt = torch.as_tensor(np.asarray(sg["v"]), device="cuda"), float(sg["dv"]), sim.grid.dt
    mv = lambda f: (f * v[None, :]).sum(dim=1) / f.sum(dim=1)  # noqa: E731  (dv cancels)
    hist, ts = [mv(sim.state["electron"])], [0.0]
    for n in range(2000):
        sim.step()
        if (n + 1) % 4 == 0:
            hist.append(mv(sim.state["electron"]))
            ts.append((n + 1) * dt)
    amp = quiver_amplitude(torch.stack(hist).cpu().numpy(), np.array(ts), w0)
    assert abs(amp - a0) / a0 < 0.05


def test_ion_acoustic_dispersion_multispecies_on_gpu():
    """tests/test_vlasov1d/test_ion_acoustic_wave.py (test_ion_acoustic_dispersion): kinetic electrons AND ions
    (m_i = 18360), 10001 sixth-order steps of multispecies_ion_acoustic.yaml at nx = 64; the frequency of the electron
    density's k = 1 mode matches omega = k cs / sqrt(1 + k^2 lambda_D^2) to 15 %."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200.module import Vlasov1D

    deck = load("multispecies_ion_acoustic")
    deck["grid"]["nx"] = 64
    deck["terms"]["time"] = "sixth"
    k = deck["density"]["species-electron-background"]["wavenumber"]
    ion = deck["terms"]["species"][1]
    T_e = deck["density"]["species-electron-background"]["T0"]
    want = np.sqrt(k**2 * (ion["charge"] * T_e / ion["mass"]) / (1 + k**2))
    sim = Vlasov1D(deepcopy(deck))
    dv = float(sim.cfg["grid"]["species_grids"]["electron"]["dv"])
    dt = sim.grid.dt
    nsteps = sim.grid.nt
    hist, ts = [sim.state["electron"].sum(dim=1) * dv], [0.0]
    for n in range(nsteps):
        sim.step()
        if (n + 1) % 10 == 0:  # save.fields.t: 1001 points over [0, 5000]
            hist.append(sim.state["electron"].sum(dim=1) * dv)
            ts.append((n + 1) * dt)
    dens, t = torch.stack(hist).cpu().numpy(), np.array(ts)
    assert np.all(np.isfinite(dens)) and np.std(sim.state["e"].cpu().numpy()) > 0
    nk1 = 2.0 / dens.shape[1] * np.fft.fft(dens, axis=1)[:, 1]
    late = nk1[len(t) // 2:]
    omega = 2 * np.pi * np.fft.fftfreq(len(late), t[1] - t[0])
    spec = np.abs(np.fft.fft(late))
    search = (omega > 0) & (omega < 5 * want) & (omega > want / 5)
    got = omega[search][np.argmax(spec[search])]
    np.testing.assert_allclose(got, want, rtol=0.15)


@pytest.mark.parametrize("operator_type", ["chang_cooper_dougherty", "chang_cooper", "Dougherty", "Lenard_Bernstein",
                                           "dougherty_nodrag"])
def test_fokker_planck_conservation_run_on_gpu(operator_type):
    """tests/test_vlasov1d/test_fokker_planck_conservation.py on the GPU path: the whole deck, every operator type,
    density to 1e-10 and energy to 1e-6 at every grid point and step."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from test_oracle_operators import check_fp_conservation

    from adept_b200.module import Vlasov1D

    deck = load("fokker_planck_conservation")
    deck["terms"]["fokker_planck"]["type"] = operator_type
    sim = Vlasov1D(deepcopy(deck))
    hist = [sim.state["electron"].cpu().numpy()]
    for _ in range(sim.grid.nt - 1):
        sim.step()
        hist.append(sim.state["electron"].cpu().numpy())
    check_fp_conservation(hist, sim.cfg)


@pytest.mark.parametrize("name", ["C1-epw", "C2-epw-fp", "C2-cc-lb-krook", "L-1024x1024-fp", "L-2048x512-sixth",
                                  "L-multispecies-1024", "C2-ex-stochastic-sixth"])
def test_cuda_graph_replay_equals_eager_steps(name):
    """adept_b200_step_f64 captured in a CUDA graph (cooperative field launches included) with the time factors read
    from the device-resident time row: 100 replayed steps give bit-identical state to 100 host-issued steps, and the
    replayed run agrees with the oracle after 6 steps to 1e-12."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200.module import Vlasov1D

    deck = VARIANTS[name]
    dt = O.build_cfg(deepcopy(deck))["grid"]["dt"]
    t0 = 30.0
    i0 = int(round(t0 / dt))
    eager, graph = Vlasov1D(deepcopy(deck)), Vlasov1D(deepcopy(deck))
    for sim in (eager, graph):
        sim.t, sim.step_index = i0 * dt, i0
    gs = graph.graph_stepper(200)  # takes one eager step itself
    eager.step()
    cfg = O.build_cfg(deepcopy(deck))
    vf = O.VlasovMaxwell(cfg)
    y = O.init_state(cfg)
    for n in range(7):
        y = vf((i0 + n) * dt, y, None)
    gs.run(6)
    for _ in range(6):
        eager.step()
    for k in y:
        if k in ("e",) or k.startswith("diag-"):
            continue
        assert rel_l2(graph.state[k].cpu().numpy(), y[k]) <= 1e-12, (name, k)
    gs.run(94)
    for _ in range(94):
        eager.step()
    assert graph.step_index == eager.step_index and graph.t == eager.t
    for k, v in eager.state.items():
        assert torch.equal(graph.state[k], v), (name, k)
