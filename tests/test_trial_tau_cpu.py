"""Trial-wavefunction estimator (SumEContrib, src/fcimc_helper.F90:586-648; hash_search_trial,
src/searching.F90:182-223; AddNewHashDet, src/load_balancer.fpp:586-611) and the tau-search hooks
(log_spawn_magnitude, src/tau/tau_search_conventional.F90:138-260) on the oracle."""
import numpy as np

import helpers
from neci_stable_b200 import capi, host, driver
from neci_stable_b200.capi import ST


def _exact_ground_state(o, s):
    dets = helpers.all_dets(s)
    H = helpers.hamiltonian_matrix(o, s, dets)
    w, v = np.linalg.eigh(H)
    psi = v[:, 0] * np.sign(v[np.argmax(np.abs(v[:, 0])), 0])
    return dets, H, w[0], psi


def test_trial_estimator_is_exact_on_the_exact_wavefunction():
    """With walkers = the exact ground state, E = E_T + numerator / denominator = <psi|H|psiT> / <psi|psiT> = E0
    for ANY trial space, because the connected-space amplitudes are (H psiT)_i."""
    s = host.hubbard_k_system(2, 2, nel=4, U=2.0)
    hii = driver.diag_energy(s, s.ref_orbs)
    o, params = helpers.make_pair(s, hii, max_walkers=20000, max_spawned=20000, all_real_coeff=True, initiator=False, seed=3)
    dets, H, e0, psi = _exact_ground_state(o, s)
    order = np.argsort(-np.abs(psi))
    trial = [dets[i] for i in order[:5]]
    ti, ta, ci, ca, e_t = helpers.build_trial_space(o, s, dets, trial)
    recs = np.array([host.record(s, d, 1000.0 * a) for d, a in zip(dets, psi) if abs(a) > 1e-14])
    o.upload_walkers(recs)
    o.set_trial_space(ti, ta, ci, ca)
    st = o.iterate(1e-4, 0.0, 1)
    e = e_t + st[ST["TRIAL_NUMERATOR"]] / st[ST["TRIAL_DENOM"]]
    assert abs(e - e0) < 1e-10, (e, e0)
    # flags: trial and connected are exclusive, and every trial determinant present is flagged
    d, _, _ = o.download_walkers()
    f = d[:, s.nw + 1]
    assert not np.any(((f >> capi.FLAG_TRIAL) & 1) & ((f >> capi.FLAG_CONNECTED) & 1))
    assert int(((f >> capi.FLAG_TRIAL) & 1).sum()) == len(trial)


def test_new_determinants_get_trial_flags_and_amplitudes():
    """Start from the reference only: every determinant inserted later must be looked up in the two tables."""
    s = host.hubbard_k_system(2, 2, nel=4, U=2.0)
    hii = driver.diag_energy(s, s.ref_orbs)
    o, params = helpers.make_pair(s, hii, max_walkers=20000, max_spawned=20000, initiator=False, seed=5)
    dets, H, e0, psi = _exact_ground_state(o, s)
    order = np.argsort(-np.abs(psi))
    trial = [dets[i] for i in order[:4]]
    ti, ta, ci, ca, e_t = helpers.build_trial_space(o, s, dets, trial)
    o.upload_walkers(host.record(s, s.ref_orbs, 200.0).reshape(1, -1))
    o.set_trial_space(ti, ta, ci, ca)
    for it in range(1, 60):
        st = o.iterate(0.02, 0.0, it)
    d, _, _ = o.download_walkers()
    c = helpers.canon(d, nw=s.nw)
    tset = {tuple(int(x) for x in r) for r in ti}; cset = {tuple(int(x) for x in r) for r in ci}
    assert c[0].shape[0] > 10
    for orb, fl in zip(c[0], c[2]):
        key = tuple(int(x) for x in orb)
        assert bool((fl >> capi.FLAG_TRIAL) & 1) == (key in tset)
        assert bool((fl >> capi.FLAG_CONNECTED) & 1) == (key in cset and key not in tset)
    assert st[ST["TRIAL_DENOM"]] != 0.0 and st[ST["TRIAL_NUMERATOR"]] != 0.0


def test_tau_search_gamma_bounds_every_spawn():
    """gamma_class = max |H_ij| / (pgen / p_class), so the largest spawn of the run is tau * max(gamma_class / p_class)
    (the tau search picks tau = p_class / gamma_class to cap it at one walker); the per-class counts add up to the
    valid excitations."""
    s = host.random_fcidump_system(6, 6, sparse=0.9, sparse_t=0.9, seed=3)
    hii = driver.diag_energy(s, s.ref_orbs)
    o, params = helpers.make_pair(s, hii, max_walkers=50000, max_spawned=50000, tau_search=True, seed=9)
    o.upload_walkers(host.record(s, s.ref_orbs, 500.0, 1 << capi.FLAG_INITIATOR).reshape(1, -1))
    g = np.zeros(4); cnt = np.zeros(4); valid = 0; max_spawn = 0.0
    tau = 0.001
    for it in range(1, 30):
        st = o.iterate(tau, 0.0, it)
        g = np.maximum(g, st[ST["TAU_GAMMA_SING"]:ST["TAU_GAMMA_SING"] + 4])
        cnt += st[ST["TAU_CNT_SING"]:ST["TAU_CNT_SING"] + 4]
        valid += st[ST["NVALIDEXCITS"]]; max_spawn = max(max_spawn, st[ST["MAX_CYC_SPAWN"]])
    assert g[1] == 0 and cnt[1] == 0                 # PCHB uses the parallel / opposite split
    assert g[0] > 0 and g[2] > 0 and g[3] > 0
    assert cnt[2] + cnt[3] + cnt[0] <= valid and cnt[2] + cnt[3] > 0.5 * valid
    t = s.tables["pchb"]
    bound = tau * max(g[0] / t["p_singles"], g[2] / (t["p_doubles"] * t["p_parallel"]), g[3] / (t["p_doubles"] * (1 - t["p_parallel"])))
    assert abs(max_spawn - bound) <= 1e-12 * bound, (max_spawn, bound)


def test_tau_search_loop_caps_the_spawns():
    """Host tau search (driver.TauSearch = update_tau) driven by the engine's gamma statistics: starting from a time
    step ten times too large it settles on tau = p_class / gamma_class, after which no spawn exceeds MaxWalkerBloom,
    and the class biases move to the gamma ratios (fed back through set_excit_probs)."""
    s = host.random_fcidump_system(6, 6, sparse=0.9, sparse_t=0.9, seed=3)
    hii = driver.diag_energy(s, s.ref_orbs)
    o, params = helpers.make_pair(s, hii, max_walkers=200000, max_spawned=200000, tau_search=True, seed=9, initiator=False)
    o.upload_walkers(host.record(s, s.ref_orbs, 2000.0, 1 << capi.FLAG_INITIATOR).reshape(1, -1))
    t = s.tables["pchb"]
    ts = driver.TauSearch(0.02, t["p_singles"], t["p_doubles"], t["p_parallel"], consider_par_bias=True)
    tau, sft, it = ts.tau, 0.0, 0
    probs0 = (ts.p_singles, ts.p_parallel)
    late_max = 0.0
    for cyc in range(60):
        for _ in range(10):
            it += 1
            st = o.iterate(tau, sft, it)
            ts.log(st)
            if cyc >= 40:
                late_max = max(late_max, st[ST["MAX_CYC_SPAWN"]])
        sft = -2.0 if st[ST["TOTPARTS"]] > 20000 else 0.0          # crude population control
        tau, ps, pd, pp = ts.update()
        o.set_excit_probs(ps, pd, pp)
    assert ts.enough[0] and ts.enough[2] and ts.enough[3]
    assert tau < 0.02 / 3                                             # the search had to cut the time step
    assert late_max <= 1.0 + 1e-9, late_max                           # MaxWalkerBloom = 1
    g_sing, _, g_par, g_opp = ts.gamma
    assert abs(ts.p_parallel - g_par / (g_par + g_opp)) < 1e-12
    assert (ts.p_singles, ts.p_parallel) != probs0
    assert abs(ts.p_singles + ts.p_doubles - 1.0) < 1e-15
    assert ts.max_death_cpt > 0.0 and tau <= 1.0 / ts.max_death_cpt             # death cap of update_tau


def test_tau_assignment_condition_follows_the_reference():
    """update_tau's final condition as Fortran parses it (src/tau/tau_search_conventional.F90:445-450): tau may be RAISED
    when enough singles have been seen (enough_sing alone), on lattice models (tHub) at every update, on the k-space
    lattice with enough doubles; otherwise only lowered.  The enough_* switches of several ranks are OR-ed."""
    def ts(**kw):
        t = driver.TauSearch(1e-4, 0.1, 0.9, 0.5, consider_par_bias=False, **kw)
        t.gamma = np.array([0.05, 0.0, 0.0, 0.0])        # singles only: tau_new = bloom * pSingles / gamma_sing = 2.0 -> max_tau
        return t
    t = ts(); t.cnt = np.array([10.0, 0, 0, 0])
    assert t.update()[0] == 1e-4                                         # too few singles: tau is not raised
    t = ts(); t.cnt = np.array([60.0, 0, 0, 0])
    assert abs(t.update()[0] - 0.99999) < 1e-12                          # enough_sing alone raises it
    t = ts(t_hub=True); t.cnt = np.array([1.0, 0, 0, 0])
    assert abs(t.update()[0] - 0.99999) < 1e-12                          # tHub: always assigned
    t = ts(); t.cnt = np.array([10.0, 0, 0, 0])
    t.reduce_or = lambda e: np.logical_or(e, np.array([True, False, False, False]))   # another rank has seen enough singles
    assert abs(t.update()[0] - 0.99999) < 1e-12
    t = ts(); t.cnt = np.array([10.0, 0, 0, 0])
    t.reduce_max = lambda v: np.maximum(v, np.array([0.4, 0, 0, 0, 5.0e4]))   # another rank's larger gamma and death component
    t.tau = 1.0
    assert abs(t.update()[0] - 0.99999 / 5.0e4) < 1e-15                  # lowered to 1 / max_death_cpt
