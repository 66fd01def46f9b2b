"""The oracle against numbers the REFERENCE ITSELF printed: the checked-in regression case
test_suite/neci/parallel/HeHe_SS_Doubles (10 orbitals, 4 electrons, HPHF, semi-stochastic doubles core, real
coefficients).  tests/golden/hehe_ss_doubles.json holds that case's FCIDUMP integrals and the values from its
benchmark output (made by tests/golden/make_hehe_fixture.py in the build container)."""
import json
import os

import numpy as np

import helpers
from neci_stable_b200 import capi, host, driver


def load_case():
    g = json.load(open(os.path.join(helpers.GOLDEN, "hehe_ss_doubles.json")))
    s = host.fcidump_system(g["norb"], g["nelec"], g["h1"], g["eri"], ecore=g["ecore"], ms2=g["ms2"],
                            orbsym=g["orbsym"], eps=g["eps"])
    return g, s


def test_determinant_energies_match_the_reference_output():
    """`Reference Energy set to` and `Highest energy determinant` of the reference's run: FCIDUMP -> UMAT packing
    (UMatInd), TMAT and sltcnd_0 + ECore reproduce them to all printed digits."""
    g, s = load_case()
    assert [int(x) for x in s.ref_orbs] == g["reference_det"]          # the reference chose the same determinant
    assert abs(driver.diag_energy(s, s.ref_orbs) - g["reference_energy"]) < 5e-12
    assert abs(driver.diag_energy(s, g["highest_det"]) - g["highest_det_energy"]) < 5e-13


def test_exact_ground_state_matches_the_reference_fciqmc_energy():
    """Exact diagonalisation of the oracle's Hamiltonian (determinants and HPHF functions) against the final
    projected energy of the reference's FCIQMC run, -5.76223713 +- 4.0e-5."""
    g, s = load_case()
    hii = driver.diag_energy(s, s.ref_orbs)
    o, _ = helpers.make_pair(s, hii, max_walkers=10000, max_spawned=10000)
    dets = helpers.all_dets(s)
    assert len(dets) == 2025
    e0 = np.linalg.eigvalsh(helpers.hamiltonian_matrix(o, s, dets))[0]
    assert abs(e0 - g["total_projected_energy"]) < 5 * g["total_projected_energy_error"], (e0, g["total_projected_energy"])
    # the same state in the HPHF basis
    oh, _ = helpers.make_pair(s, hii, max_walkers=10000, max_spawned=10000, hphf=True)
    A, B = 0xAAAAAAAAAAAAAAAA, 0x5555555555555555
    words = [int(np.uint64(s.ilut(d)[0])) for d in dets]
    reps = [w for w in words if w >= (((w & A) >> 1) | ((w & B) << 1))]
    il = np.array(reps, dtype=np.uint64).view(np.int64).reshape(-1, 1)
    m = len(reps)
    I = np.repeat(np.arange(m), m); J = np.tile(np.arange(m), m)
    eh = np.linalg.eigvalsh(oh.probe_helement(il[I], il[J]).reshape(m, m))[0]
    assert abs(eh - e0) < 1e-10


def test_oracle_run_in_the_reference_configuration_reproduces_its_energy():
    """The reference's own run parameters (neci.inp of the case): HPHF, real coefficients with spawn cutoff 0.01,
    tau 0.001, semi-stochastic with the doubles core space, ~1000 walkers, shift damping 0.1 every iteration.
    Projected energy within the combined error bars of the two runs."""
    g, s = load_case()
    hii = driver.diag_energy(s, s.ref_orbs)
    o, _ = helpers.make_pair(s, hii, max_walkers=20000, max_spawned=40000, hphf=True, all_real_coeff=True,
                             real_spawn_cutoff=g["input"]["realspawncutoff"], semi_stochastic=True, initiator=False, seed=7)
    run = helpers.run_with_doubles_core(o, s, hii, tau=g["input"]["tau"], target=g["input"]["totalwalkers"], n_iter=14000,
                                        steps_sft=g["input"]["stepsshift"], sft_damp=g["input"]["shiftdamp"],
                                        start=float(g["input"]["startsinglepart"]), diag_sft=g["input"]["diagshift"])
    hist = [h for h in run.history if h["varying"]][2000:]
    assert len(hist) > 4000
    e, err = driver.ratio_estimate([h["enum_cyc"] for h in hist], [h["hf_cyc"] for h in hist])
    tol = 5 * np.hypot(err, g["total_projected_energy_error"])
    assert abs(e + hii - g["total_projected_energy"]) < max(tol, 3e-4), (e + hii, g["total_projected_energy"], err)


def load_ne_pchb():
    z = np.load(os.path.join(helpers.GOLDEN, "ne_pchb.npz"))
    s = host.fcidump_system(int(z["norb"]), int(z["nelec"]), z["h1"], z["eri"], ecore=float(z["ecore"]), ms2=0,
                            orbsym=[int(x) for x in z["orbsym"]], eps=z["eps"])
    return z, s


def test_ne_pchb_reference_energy_with_frozen_core():
    """Second regression case of the reference, Ne_FciMCPar_pchb (23 orbitals, `freeze 2 0`, the PCHB generator): the
    frozen-core folding of tests/golden/make_ne_pchb_fixture.py + the oracle's sltcnd_0 reproduce the reference's
    `Reference Energy set to: -128.4963497303` and its choice of reference determinant."""
    z, s = load_ne_pchb()
    assert [int(x) for x in s.ref_orbs] == [int(x) for x in z["reference_det"]]
    assert abs(driver.diag_energy(s, s.ref_orbs) - float(z["reference_energy"])) < 5e-10


def test_ne_pchb_generator_is_unbiased_from_the_reference_determinant():
    """PCHB doubles + uniform singles with the case's eight (spin, irrep) classes: sum 1/pgen over the draws that land
    on a determinant estimates 1 for every connected determinant (the reference's generator criterion)."""
    z, s = load_ne_pchb()
    hii = driver.diag_energy(s, s.ref_orbs)
    o, _ = helpers.make_pair(s, hii, max_walkers=10000, max_spawned=10000)
    il = s.ilut(s.ref_orbs).reshape(1, -1)
    n_draw = 400000
    out = o.probe_gen_excit(np.repeat(il, n_draw, axis=0), np.arange(n_draw, dtype=np.int32), 1)
    ok = out["pgen"] > 0
    tgt = out["ilut_j"][ok, 0]
    inv = 1.0 / out["pgen"][ok]
    keys, idx = np.unique(tgt, return_inverse=True)
    acc = np.bincount(idx, weights=inv) / n_draw
    cnt = np.bincount(idx)
    hel = np.abs(out["hel"][ok])
    nonzero = np.bincount(idx, weights=(hel > 1e-12)) > 0
    big = cnt > 400
    assert big.sum() > 100
    assert np.all(np.abs(acc[big] - 1.0) < 4.5 / np.sqrt(cnt[big]) + 0.01)
    # 24 symmetry-allowed singles + 960 doubles from the reference (the reference's own count, benchmark line 230)
    ic = out["ic"][ok]
    n_singles = len(np.unique(tgt[(ic == 1) & (hel > 1e-12)]))
    n_doubles = len(np.unique(tgt[(ic == 2) & (hel > 1e-12)]))
    assert n_singles <= 24 and n_doubles <= 960 and n_doubles > 800


def test_reference_energies_of_five_regression_cases():
    """`Reference Energy set to` as printed by the reference's own runs of C2_FCIMCPar_CAS (freeze 4), H4, the
    Cr2 case (freeze 24: the 24-electron / 30-orbital system of BASELINE configs[4]), H2O and Ne (freeze 2):
    integrals over the occupied orbitals from tests/golden/reference_energies.json (frozen core folded by the
    generating script), energy from the oracle's UMAT packing + sltcnd_0."""
    cases = json.load(open(os.path.join(helpers.GOLDEN, "reference_energies.json")))
    assert len(cases) == 5
    for c in cases:
        n = c["n_spat"]
        assert c["nalpha"] == c["nbeta"] == n and c["det"] == list(range(1, 2 * n + 1))
        s = host.fcidump_system(n, 2 * n, c["h1"], c["eri"], ecore=c["ecore"], ms2=0, ref_spatial=list(range(1, n + 1)))
        e = driver.diag_energy(s, s.ref_orbs)                      # host-side Python (what seeds Hii)
        tol = 5e-10 * max(1.0, abs(e) / 100)
        assert abs(e - c["reference_energy"]) < tol, (c["case"], e, c["reference_energy"])
        o, _ = helpers.make_pair(s, e, max_walkers=100, max_spawned=100)
        il = s.ilut(s.ref_orbs).reshape(1, -1)
        assert abs(o.probe_helement(il, il)[0] - c["reference_energy"]) < tol        # the oracle's sltcnd_0 + ECore
        o.close()
