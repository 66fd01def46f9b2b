"""determ_projection (src/semi_stoch_procs.F90:105-241) on the oracle: with the whole Hilbert space as the core
space the iteration is the exact power method psi <- psi - tau (H - S) psi, so the projected energy must converge to the
lowest eigenvalue, and one step must equal the dense mat-vec."""
import numpy as np

import helpers
from neci_stable_b200 import capi, host, driver
from neci_stable_b200.capi import ST


def _setup(system, nranks=1, **kw):
    hii = driver.diag_energy(system, system.ref_orbs)
    orcs = []
    for r in range(nranks):
        o, params = helpers.make_pair(system, hii, max_walkers=20000, max_spawned=20000, nranks=nranks, rank=r,
                                      semi_stochastic=True, seed=3, **kw)
        orcs.append(o)
    return hii, orcs


def test_full_space_core_is_exact_power_iteration():
    s = host.hubbard_k_system(2, 2, nel=4, U=2.0)
    hii, (o,) = _setup(s)
    dets = helpers.all_dets(s)
    dets, sizes, displs, per_rank, H = helpers.build_core_space(o, s, dets, hii)
    n = len(dets)
    rng = np.random.default_rng(0)
    amp = rng.normal(size=n)
    iref = dets.index(sorted(int(x) for x in s.ref_orbs))
    amp[iref] = 5.0
    flags = (1 << capi.FLAG_DETERMINISTIC) | (1 << capi.FLAG_INITIATOR)
    recs = np.array([host.record(s, d, float(a), flags) for d, a in zip(dets, amp)])
    o.upload_walkers(recs)
    c = per_rank[0]
    o.set_core_space(c["row_ptr"], c["col"], c["val"], sizes, displs, c["iluts"])
    tau, S = 0.02, 0.0
    v = amp.copy()
    for it in range(1, 4):
        st = o.iterate(tau, S, it)
        v = v - tau * (H @ v - S * v)                      # H already has Hii subtracted on the diagonal
        d, _, _ = o.download_walkers()
        assert np.allclose(host.signs_of(d, s.nw), v, rtol=1e-13, atol=1e-13)
        assert st[ST["NSPAWNED_SENT"]] == 0                # core -> core spawns are cancelled
    e0 = np.linalg.eigvalsh(H)[0]
    for it in range(4, 2500):                              # shift = E0: the ground-state component is stationary
        st = o.iterate(tau, e0, it)
    proje = st[ST["ENUMCYC"]] / st[ST["HFCYC"]]
    assert abs(proje - e0) < 1e-8, (proje, e0)


def test_core_space_split_over_ranks_matches_single_rank():
    s = host.hubbard_k_system(2, 2, nel=4, U=2.0)
    results = {}
    for nr in (1, 3):
        hii, orcs = _setup(s, nr)
        dets = helpers.all_dets(s)
        dets, sizes, displs, per_rank, H = helpers.build_core_space(orcs[0], s, dets, hii, nranks=nr)
        amp = {tuple(d): float(i % 7) - 3.0 for i, d in enumerate(helpers.all_dets(s))}
        flags = (1 << capi.FLAG_DETERMINISTIC) | (1 << capi.FLAG_INITIATOR)
        for r in range(nr):
            mine = dets[displs[r]:displs[r] + sizes[r]]
            recs = np.array([host.record(s, d, amp[tuple(d)], flags) for d in mine]).reshape(-1, s.W)
            orcs[r].upload_walkers(recs)
            c = per_rank[r]
            orcs[r].set_core_space(c["row_ptr"], c["col"], c["val"], sizes, displs, c["iluts"])
        for it in range(1, 6):
            helpers.world_iterate(orcs, 0.01, 0.1, it, nthreads=1)
        parts = [o.download_walkers() for o in orcs]
        d = np.concatenate([p[0] for p in parts])
        results[nr] = helpers.canon(d, nw=s.nw)
    assert np.array_equal(results[1][0], results[3][0])
    assert np.allclose(results[1][1], results[3][1], rtol=1e-13, atol=1e-13)


def test_no_death_projection_plus_death_equals_fused_projection():
    """tDeathBeforeComms = .false. (the reference's default with real coefficients): determ_projection_no_death
    (src/semi_stoch_procs.F90:285-374) leaves the diagonal and the shift to the death step, which then also acts on
    the core determinants (perform_death_all_walkers, src/fcimc_helper.F90:2253-2277).  With the whole space as
    core space both orders are the same exact power iteration."""
    s = host.hubbard_k_system(2, 2, nel=4, U=2.0)
    out = {}
    for dbc in (True, False):
        hii, (o,) = _setup(s, all_real_coeff=True, death_before_comms=dbc, initiator=False)
        dets = helpers.all_dets(s)
        dets, sizes, displs, per_rank, H = helpers.build_core_space(o, s, dets, hii)
        rng = np.random.default_rng(1)
        amp = rng.normal(size=len(dets)) + 3.0
        flags = (1 << capi.FLAG_DETERMINISTIC) | (1 << capi.FLAG_INITIATOR)
        o.upload_walkers(np.array([host.record(s, d, float(a), flags) for d, a in zip(dets, amp)]))
        c = per_rank[0]
        o.set_core_space(c["row_ptr"], c["col"], c["val"], sizes, displs, c["iluts"])
        v = amp.copy()
        for it in range(1, 6):
            o.iterate(0.02, 0.3, it)
            v = v - 0.02 * (H @ v - 0.3 * v)
        d, _, _ = o.download_walkers()
        out[dbc] = host.signs_of(d, s.nw).copy()
        assert np.allclose(out[dbc], v, rtol=1e-12, atol=1e-12)
    assert np.allclose(out[True], out[False], rtol=1e-12, atol=1e-12)
