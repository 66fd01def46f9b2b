"""End-to-end FCIQMC runs of the CUDA engine through the C ABI (driver.FciMC = the outer loop of FciMCPar,
src/FciMCPar.F90:394-854): projected energy, shift and trial-wavefunction energy must agree with exact
diagonalisation within blocking-analysis error bars (src/ErrorAnalysis.F90) -- the north_star's energy criterion."""
import numpy as np
import pytest

import helpers
from neci_stable_b200 import capi, host, driver

pytestmark = pytest.mark.gpu


def _engine(system, **kw):
    hii = driver.diag_energy(system, system.ref_orbs)
    kw.setdefault("max_walkers", 400000); kw.setdefault("max_spawned", 400000); kw.setdefault("seed", 3)
    gpu = capi.Engine(host.make_params(system, hii, **kw))
    system.apply(gpu)
    return gpu, hii


def _exact(gpu, system):
    dets = helpers.all_dets(system)
    return dets, np.linalg.eigh(helpers.hamiltonian_matrix(gpu, system, dets))


def _estimates(run, skip=100):
    hist = [h for h in run.history if h["varying"]][skip:]
    e, err = driver.ratio_estimate([h["enum_cyc"] for h in hist], [h["hf_cyc"] for h in hist])
    sm, serr = driver.blocking([h["shift"] for h in hist])
    return hist, e, err, sm, serr


@pytest.mark.parametrize("kind", ["hub_k_2x2", "hub_rs_2x2", "pchb_6e6o"])
def test_projected_energy_and_shift_match_exact_diagonalisation(kind):
    if kind == "hub_k_2x2":
        s, tau = host.hubbard_k_system(2, 2, nel=4, U=1.0), 0.01          # reference suite: -7.29750728 +- 2.8e-4
    elif kind == "hub_rs_2x2":
        s, tau = host.hubbard_rs_system(2, 2, U=4.0), 0.01
    else:
        s, tau = host.random_fcidump_system(6, 6, sparse=0.9, sparse_t=0.9, seed=3), 0.002
    gpu, hii = _engine(s, initiator=False)
    _, (w, _) = _exact(gpu, s)
    e0 = w[0]
    run = driver.FciMC(s, gpu, hii, tau=tau, init_walkers=3000, steps_sft=10, sft_damp=0.1)
    run.seed_reference(10)
    run.run(8000)
    hist, e, err, sm, serr = _estimates(run)
    assert len(hist) > 200
    assert abs(e + hii - e0) < max(5 * err, 2e-3), (e + hii, e0, err)
    assert abs(sm + hii - e0) < max(5 * serr, 2e-2), (sm + hii, e0, serr)
    if kind == "hub_k_2x2":
        assert abs(e0 - (-7.29750728)) < 2e-3          # the reference's published benchmark energy for this system


def test_semi_stochastic_real_coefficient_run_with_trial_energy():
    """Real coefficients, semi-stochastic core space (death after the walker loop, determ_projection_no_death) and the
    trial-wavefunction estimator together: E_trial = E_T + numerator / denominator against exact diagonalisation."""
    s = host.random_fcidump_system(6, 6, sparse=0.9, sparse_t=0.9, seed=3)
    gpu, hii = _engine(s, initiator=False, all_real_coeff=True, semi_stochastic=True)
    dets, (w, v) = _exact(gpu, s)
    e0 = w[0]
    psi0 = v[:, 0]
    order = np.argsort(-np.abs(psi0))
    ref = [int(x) for x in s.ref_orbs]
    core = [dets[i] for i in order[:40]]
    if ref not in core:
        core = [ref] + core[:-1]
    core, sizes, displs, per_rank, H = helpers.build_core_space(gpu, s, core, hii)
    flags = (1 << capi.FLAG_DETERMINISTIC) | (1 << capi.FLAG_INITIATOR)
    recs = np.array([host.record(s, d, 10.0 if d == ref else 0.0, flags) for d in core])
    gpu.upload_walkers(recs)
    c = per_rank[0]
    gpu.set_core_space(c["row_ptr"], c["col"], c["val"], sizes, displs, c["iluts"])
    trial = [dets[i] for i in order[:12]]
    ti, ta, ci, ca, e_t = helpers.build_trial_space(gpu, s, dets, trial)
    gpu.set_trial_space(ti, ta, ci, ca)
    run = driver.FciMC(s, gpu, hii, tau=0.002, init_walkers=3000, steps_sft=10, sft_damp=0.1)
    run.tot_parts = 10.0; run.old_av_walkers = 10.0
    run.run(8000)
    hist, e, err, sm, serr = _estimates(run)
    assert abs(e + hii - e0) < max(5 * err, 2e-3), (e + hii, e0, err)
    et, terr = driver.ratio_estimate([h["trial_num"] for h in hist], [h["trial_den"] for h in hist])
    assert abs(e_t + et - e0) < max(5 * terr, 2e-3), (e_t + et, e0, terr)
    # the larger trial space gives the better estimator
    assert terr <= err * 1.5 + 1e-6


def test_hphf_run_matches_exact_diagonalisation():
    s = host.random_fcidump_system(6, 6, sparse=0.9, sparse_t=0.9, seed=3)
    gpu_det, _ = _engine(s, initiator=False)
    _, (w, _) = _exact(gpu_det, s)
    gpu, hii = _engine(s, initiator=False, hphf=True)
    run = driver.FciMC(s, gpu, hii, tau=0.002, init_walkers=3000, steps_sft=10, sft_damp=0.1)
    run.seed_reference(10)
    run.run(8000)
    hist, e, err, sm, serr = _estimates(run)
    assert abs(e + hii - w[0]) < max(5 * err, 2e-3), (e + hii, w[0], err)
    assert abs(sm + hii - w[0]) < max(5 * serr, 2e-2), (sm + hii, w[0], serr)


def test_reference_regression_case_hehe_ss_doubles():
    """The reference's own regression case test_suite/neci/parallel/HeHe_SS_Doubles run on the CUDA engine with the
    case's own parameters (HPHF, real coefficients with spawn cutoff 0.01, tau 0.001, semi-stochastic doubles core,
    1000 walkers, shift updated every iteration with damping 0.1, initial shift 1.0): the determinant energies the
    reference printed are reproduced to all digits and the projected energy agrees with the reference CPU run's
    -5.76223713 +- 4.0e-5 within the combined error bars -- the north_star's acceptance criterion, against the
    reference's own output."""
    import json, os
    g = json.load(open(os.path.join(helpers.GOLDEN, "hehe_ss_doubles.json")))
    s = host.fcidump_system(g["norb"], g["nelec"], g["h1"], g["eri"], ecore=g["ecore"], ms2=g["ms2"],
                            orbsym=g["orbsym"], eps=g["eps"])
    inp = g["input"]
    gpu, hii = _engine(s, initiator=False, hphf=True, all_real_coeff=True, real_spawn_cutoff=inp["realspawncutoff"],
                       semi_stochastic=True, seed=7, max_walkers=20000, max_spawned=40000)
    il = lambda d: s.ilut(d).reshape(1, -1)
    assert abs(gpu.probe_helement(il(g["reference_det"]), il(g["reference_det"]))[0] - g["reference_energy"]) < 5e-12
    assert abs(gpu.probe_helement(il(g["highest_det"]), il(g["highest_det"]))[0] - g["highest_det_energy"]) < 5e-13
    run = helpers.run_with_doubles_core(gpu, s, hii, tau=inp["tau"], target=inp["totalwalkers"], n_iter=14000,
                                        steps_sft=inp["stepsshift"], sft_damp=inp["shiftdamp"],
                                        start=float(inp["startsinglepart"]), diag_sft=inp["diagshift"])
    hist = [h for h in run.history if h["varying"]][2000:]
    assert len(hist) > 4000
    e, err = driver.ratio_estimate([h["enum_cyc"] for h in hist], [h["hf_cyc"] for h in hist])
    tol = 5 * np.hypot(err, g["total_projected_energy_error"])
    assert abs(e + hii - g["total_projected_energy"]) < max(tol, 3e-4), (e + hii, g["total_projected_energy"], err)
    sm, serr = driver.blocking([h["shift"] for h in hist])
    assert abs(sm + hii - g["total_projected_energy"]) < max(5 * serr, 2e-2), (sm + hii, serr)


@pytest.mark.parametrize("particle_selection", ["FULL-FULL", "UNIF-UNIF"])
def test_reference_regression_case_ne_pchb(particle_selection):
    """The reference's PCHB regression case test_suite/neci/parallel/Ne_FciMCPar_pchb (Ne, 8 active electrons in 22
    orbitals after `freeze 2 0`, i-FCIQMC with integer walkers, addtoinitiator 3, 20000 walkers, shift damping 0.03
    every 25 iterations) on the CUDA engine: reference-determinant energy to the printed digits, and the initiator
    projected energy against the reference CPU run's -128.70959926 +- 6.8e-4.  The reference run picks electron pairs
    with FULL-FULL weighting (its input's `PCHB` block): that is the first case; UNIF-UNIF, a different unbiased
    generator under the same estimator, is the second."""
    import os
    z = np.load(os.path.join(helpers.GOLDEN, "ne_pchb.npz"))
    s = host.fcidump_system(int(z["norb"]), int(z["nelec"]), z["h1"], z["eri"], ecore=float(z["ecore"]), ms2=0,
                            orbsym=[int(x) for x in z["orbsym"]], eps=z["eps"], p_singles=0.2,
                            particle_selection=particle_selection)
    gpu, hii = _engine(s, initiator=True, initiator_walk_no=float(z["input_addtoinitiator"]), seed=8,
                       max_walkers=400000, max_spawned=400000)
    assert [int(x) for x in s.ref_orbs] == [int(x) for x in z["reference_det"]]
    il = s.ilut(s.ref_orbs).reshape(1, -1)
    assert abs(gpu.probe_helement(il, il)[0] - float(z["reference_energy"])) < 5e-10
    run = driver.FciMC(s, gpu, hii, tau=0.01, init_walkers=int(z["input_totalwalkers"]),
                       steps_sft=int(z["input_stepsshift"]), sft_damp=float(z["input_shiftdamp"]), diag_sft=0.5,
                       jump_shift=True)                 # the case's `jump-shift`
    run.seed_reference(100)
    run.run(16000)
    hist = [h for h in run.history if h["varying"]][80:]
    assert len(hist) > 200
    e, err = driver.ratio_estimate([h["enum_cyc"] for h in hist], [h["hf_cyc"] for h in hist])
    ref_e, ref_err = float(z["total_projected_energy"]), float(z["total_projected_energy_error"])
    assert abs(e + hii - ref_e) < max(5 * np.hypot(err, ref_err), 3e-3), (e + hii, ref_e, err)


def test_device_diagonal_elements_match_reference_outputs():
    """The CUDA sltcnd_0 against `Reference Energy set to` of five of the reference's regression runs
    (tests/golden/reference_energies.json: C2, H4, Cr2 24e/30o, H2O, Ne)."""
    import json, os
    for c in json.load(open(os.path.join(helpers.GOLDEN, "reference_energies.json"))):
        n = c["n_spat"]
        s = host.fcidump_system(n, 2 * n, c["h1"], c["eri"], ecore=c["ecore"], ms2=0, ref_spatial=list(range(1, n + 1)))
        gpu, hii = _engine(s, max_walkers=1000, max_spawned=1000)
        il = s.ilut(s.ref_orbs).reshape(1, -1)
        e = gpu.probe_helement(il, il)[0]
        assert abs(e - c["reference_energy"]) < 5e-10 * max(1.0, abs(e) / 100), (c["case"], e, c["reference_energy"])
        gpu.close()


def test_reference_regression_case_determ_doubles():
    """The reference's determ_doubles run (semi-stochastic doubles-core in the determinant basis, set up with the host
    library) on the CUDA engine: projected correlation energy within the combined blocking errors of the reference
    run's -0.065081043 +/- 8.8e-6 and of the exact -0.0650928511 its fci-core twin printed."""
    import test_core_space_cpu as TC
    e, err, hist, ref, g = TC.determ_doubles_run(lambda s, params: capi.Engine(params))
    tol = 4.0 * np.hypot(err, ref["projected_correlation_energy_error"])
    assert err < 5e-5
    assert abs(e - ref["projected_correlation_energy"]) < tol, (e, err)
    assert abs(e - g["fci_core"]["correlation_energy"]) < tol + 2.5e-5
