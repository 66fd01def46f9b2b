"""HPHF functions (src/HPHFIntegrals.fpp, src/HPHFRandExcit.F90) on the oracle, pinned by an independent
construction: the Hamiltonian in the basis of spin-adapted pairs Phi_I = (|I> + (-1)^open |I_flipped>) / sqrt 2 built
from the determinant Hamiltonian (itself pinned in test_oracle_golden.py) must equal the oracle's HPHF elements, and
its spectrum must be the even-S part of the determinant spectrum."""
import numpy as np

import helpers
from neci_stable_b200 import capi, host, driver
from neci_stable_b200.capi import ST

BETA, ALPHA = 0x5555555555555555, 0xAAAAAAAAAAAAAAAA


def _flip(w):
    return ((w & ALPHA) >> 1) | ((w & BETA) << 1)


def _open_orbs(w):
    return bin(~((w & ALPHA) >> 1) & (w & BETA) & ((1 << 64) - 1)).count("1")


def _setup(hphf, n_spat=5, nel=4, seed=3, **kw):
    s = host.random_fcidump_system(n_spat, nel, sparse=0.9, sparse_t=0.9, seed=seed)
    hii = driver.diag_energy(s, s.ref_orbs)
    o, params = helpers.make_pair(s, hii, max_walkers=60000, max_spawned=60000, hphf=hphf, **kw)
    return s, o, hii


def _hphf_basis(s):
    """(unique representatives, index of every determinant's function, sign of the determinant inside it)"""
    dets = helpers.all_dets(s)
    words = [int(np.uint64(s.ilut(d)[0])) for d in dets]
    index = {w: i for i, w in enumerate(words)}
    reps = []
    for w in words:
        f = _flip(w)
        if f == w or w > f:                 # closed shell, or the member IsAllowedHPHF accepts (the larger one; 1 word, no sign bit)
            reps.append(w)
    return dets, words, index, reps


def test_hphf_matrix_elements_equal_the_projected_determinant_hamiltonian():
    s, o_det, _ = _setup(False)
    _, o_h, _ = _setup(True)
    dets, words, index, reps = _hphf_basis(s)
    H = helpers.hamiltonian_matrix(o_det, s, dets)
    n, m = len(words), len(reps)
    T = np.zeros((n, m))
    for k, w in enumerate(reps):
        f = _flip(w)
        if f == w:
            T[index[w], k] = 1.0
        else:
            sgn = 1.0 if _open_orbs(w) % 2 == 0 else -1.0
            T[index[w], k] = 1 / np.sqrt(2); T[index[f], k] = sgn / np.sqrt(2)
    Hp = T.T @ H @ T
    il = np.array(reps, dtype=np.uint64).view(np.int64).reshape(-1, 1)
    I = np.repeat(np.arange(m), m); J = np.tile(np.arange(m), m)
    Ho = o_h.probe_helement(il[I], il[J]).reshape(m, m)
    assert np.allclose(Ho, Hp, rtol=1e-12, atol=1e-12)
    # the HPHF spectrum is a subset of the determinant spectrum, and holds its (singlet) ground state
    ev_full = np.linalg.eigvalsh(H); ev_h = np.linalg.eigvalsh(Hp)
    assert abs(ev_h[0] - ev_full[0]) < 1e-10
    for e in ev_h:
        assert np.min(np.abs(ev_full - e)) < 1e-9


def test_gen_hphf_excit_is_complete_and_unbiased():
    """The reference's generator criterion (src/unit_test_helpers.F90:438-638) for the HPHF generator: every
    connected HPHF function is reached, sum 1/pgen / n_draws -> 1 per function, and the returned element is the
    HPHF matrix element."""
    s, o, _ = _setup(True, n_spat=5, nel=4, seed=3)
    dets, words, index, reps = _hphf_basis(s)
    il = np.array(reps, dtype=np.uint64).view(np.int64).reshape(-1, 1)
    m = len(reps)
    for src in (0, m // 2, m - 1):
        n_draw = 600000
        out = o.probe_gen_excit(np.repeat(il[src:src + 1], n_draw, axis=0), np.arange(n_draw, dtype=np.int32), 1)
        valid = out["pgen"] > 0
        tgt = out["ilut_j"][valid, 0].view(np.uint64)
        assert set(int(t) for t in tgt) <= set(reps)                   # only allowed representatives come out
        assert int(np.uint64(il[src, 0])) not in set(int(t) for t in tgt)
        hrow = o.probe_helement(np.repeat(il[src:src + 1], m, axis=0), il)
        connected = {reps[k] for k in range(m) if k != src and abs(hrow[k]) > 1e-12}
        acc, cnt = {}, {}
        for t, p, h in zip(tgt, out["pgen"][valid], out["hel"][valid]):
            t = int(t); acc[t] = acc.get(t, 0.0) + 1.0 / p; cnt[t] = cnt.get(t, 0) + 1
            assert abs(h - hrow[reps.index(t)]) < 1e-12
        assert connected <= set(acc)
        for t in connected:                                            # unbiased within 4.5 standard errors of the hit count
            assert abs(acc[t] / n_draw - 1.0) < 4.5 / np.sqrt(cnt[t]) + 0.01, (t, acc[t] / n_draw, cnt[t])


def test_hphf_fciqmc_converges_to_the_ground_state():
    """FCIQMC on HPHF functions: projected energy and shift within blocking error bars of exact diagonalisation
    (the north_star criterion), and only allowed representatives are ever occupied."""
    s, o, hii = _setup(True, n_spat=5, nel=4, seed=3, initiator=False)
    o_det = _setup(False)[1]
    e0 = np.linalg.eigvalsh(helpers.hamiltonian_matrix(o_det, s, helpers.all_dets(s)))[0]
    run = driver.FciMC(s, o, hii, tau=0.002, init_walkers=3000, steps_sft=10, sft_damp=0.1)
    run.seed_reference(10)
    run.run(8000)
    hist = [h for h in run.history if h["varying"]][100:]
    num = np.array([h["enum_cyc"] for h in hist]); den = np.array([h["hf_cyc"] for h in hist])
    e, err = driver.ratio_estimate(num, den)
    assert abs(e + hii - e0) < max(5 * err, 2e-3), (e + hii, e0, err)
    sm, serr = driver.blocking([h["shift"] for h in hist])
    assert abs(sm + hii - e0) < max(5 * serr, 2e-2), (sm + hii, e0, serr)
    d, _, _ = o.download_walkers()
    w = d[np.abs(host.signs_of(d, s.nw)) > 0][:, 0].view(np.uint64)
    assert len(w) > 10
    for x in w:
        x = int(x); f = _flip(x)
        assert f == x or x > f


def test_hphf_core_hamiltonian_matches_the_reference_hehe_ss_doubles_run():
    """Numbers the reference printed for its HPHF regression run HeHe_SS_Doubles (semi-stochastic doubles-core over
    43 HPHF functions, started from the core ground state at 10 walkers): the lowest eigenvalue of the core
    Hamiltonian built from the oracle's hphf_diag_helement / hphf_off_diag_helement is the printed
    `Deterministic subspace correlation energy` -0.0646316671, and the eigenvector's weights on the reference and on
    the doubles are the NoatHF / NoatDoubs of the first line of the iteration table (6.249123, 3.731892)."""
    import test_core_space_cpu as TC
    g, s = TC._hehe_system()
    hii = driver.diag_energy(s, s.ref_orbs)
    o, _ = helpers.make_pair(s, hii, max_walkers=1000, max_spawned=1000, hphf=True)
    sd = host.sing_doub_space(s, orbsym=g["orbsym"])
    A, B = 0xAAAAAAAAAAAAAAAA, 0x5555555555555555
    flip = lambda w: ((w & A) >> 1) | ((w & B) << 1)
    reps = np.array([[r[0]] for r in sd if int(np.uint64(r[0])) >= flip(int(np.uint64(r[0])))], dtype=np.int64)
    n = reps.shape[0]
    assert n == g["doubles_core_size_hphf"] == 43
    I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
    H = o.probe_helement(reps[I], reps[J]).reshape(n, n)
    assert np.allclose(H, H.T, atol=1e-12)
    w, v = np.linalg.eigh(H - hii * np.eye(n))
    ref = g["hphf_run"]
    assert abs(w[0] - ref["core_correlation_energy"]) < 6e-11            # -0.0646316671, printed with 10 decimals
    refw = int(np.uint64(s.ilut(s.ref_orbs)[0]))
    words = [int(np.uint64(r[0])) for r in reps]
    iref = words.index(refw)
    psi = v[:, 0] * np.sign(v[iref, 0]) * ref["start_walkers"] / np.abs(v[:, 0]).sum()       # startsinglepart 10
    level = np.array([min(bin(refw & ~x).count("1"), bin(refw & ~flip(x)).count("1")) for x in words])
    assert abs(psi[iref] - ref["step1_no_at_hf"]) < 6e-7 and abs(np.abs(psi[level == 2]).sum() - ref["step1_no_at_doubs"]) < 6e-7
    # one deterministic step at the run's tau = 0.001 and diagshift 1.0 gives the second line's NoatHF 6.255776
    assert abs(psi[iref] * (1.0 - ref["tau"] * (w[0] - ref["diagshift"])) - ref["step2_no_at_hf"]) < 2e-6


def test_mp1_core_and_trial_spaces_match_the_reference_ne_run():
    """The reference's Ne_SS_Trial_Pops run (22 orbitals and 8 electrons after `freeze 2 0`, HPHF, `mp1-core 50`,
    `mp1-trial 200`; tests/golden/ne_ss_trial.json) against the oracle's HPHF matrix elements:
      * the 985 symmetry-allowed singles and doubles of the reference reduce to 529 HPHF functions;
      * ranked by the first-order amplitude |<D_0|H|D_j> / (F_00 - F_jj)| (return_mp1_amp_and_mp2_energy,
        src/semi_stoch_procs.F90:2143-2218, orbital energies from the FCIDUMP), the first 50 span a space whose lowest
        eigenvalue is the printed `Deterministic subspace correlation energy` -0.1088456879;
      * the 200-function trial space: 188 functions lie above the cut-off and 14 share the cut-off amplitude to all
        digits, so the reference's choice among them depends on its enumeration order; one of the 91 ways of taking 12
        of the 14 gives the printed `Energy eigenvalue(s) of the trial space` -128.68457855852768 to 1e-12;
      * the space connected to that trial space has the printed 59726 members."""
    import itertools
    import json
    import os
    import test_reference_fcidump_cpu as TR
    z, s = TR.load_ne_pchb()
    g = json.load(open(os.path.join(helpers.GOLDEN, "ne_ss_trial.json")))
    hii = driver.diag_energy(s, s.ref_orbs)
    assert abs(hii - g["reference_energy"]) < 6e-11
    o, _ = helpers.make_pair(s, hii, max_walkers=1000, max_spawned=1000, hphf=True)
    sd = host.sing_doub_space(s, orbsym=[int(x) for x in z["orbsym"]])
    assert sd.shape[0] == 985
    A, B = 0xAAAAAAAAAAAAAAAA, 0x5555555555555555
    flip = lambda w: ((w & A) >> 1) | ((w & B) << 1)
    reps = np.array([[r[0]] for r in sd if int(np.uint64(r[0])) >= flip(int(np.uint64(r[0])))], dtype=np.int64)
    assert reps.shape[0] == 529
    eps = np.asarray(z["eps"], dtype=float)
    h0 = lambda w: sum(eps[b >> 1] for b in range(64) if (int(np.uint64(w)) >> b) & 1)
    ref = s.ilut(s.ref_orbs)
    hel = o.probe_helement(np.repeat(ref.reshape(1, 1), reps.shape[0], 0), reps)
    den = np.array([h0(ref[0]) - h0(r[0]) for r in reps])
    a = np.abs(hel / np.where(den == 0, 1.0, den))
    a[int(np.nonzero(reps[:, 0] == ref[0])[0][0])] = np.inf           # the reference heads the list
    order = np.argsort(-a, kind="stable")

    def lowest(idx):
        il = reps[list(idx)]
        n = il.shape[0]
        I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
        return np.linalg.eigvalsh(o.probe_helement(il[I], il[J]).reshape(n, n))[0]
    nc = g["core_size"]
    assert a[order[nc - 1]] - a[order[nc]] > 1e-5                       # no tie at the core cut-off
    assert abs(lowest(order[:nc]) - hii - g["core_correlation_energy"]) < 6e-11
    nt = g["trial_size"]
    cut = a[order[nt - 1]]
    above = [i for i in order if a[i] > cut + 1e-8]
    tied = [i for i in order if abs(a[i] - cut) <= 1e-8]
    assert (len(above), len(tied)) == (188, 14)
    choice = min(itertools.combinations(tied, nt - len(above)), key=lambda c: abs(lowest(above + list(c)) - g["trial_energy"]))
    assert abs(lowest(above + list(choice)) - g["trial_energy"]) < 1e-12
    # `Total size of connected space` 59726 (generate_connected_space_normal, src/enumerate_excitations.F90:158-277, then
    # remove_repeated_states): every symmetry-allowed single and double excitation of each trial representative, mapped
    # to its allowed HPHF representative, duplicates removed -- the trial functions themselves are still in it
    trial = reps[above + list(choice)]
    con = [trial[:, 0].view(np.uint64)]
    for r in trial:
        x = host.sing_doub_space(s, ref_ilut=r, orbsym=[int(v) for v in z["orbsym"]])[1:, 0].view(np.uint64)
        f = ((x & np.uint64(A)) >> np.uint64(1)) | ((x & np.uint64(B)) << np.uint64(1))
        con.append(np.where(x.view(np.int64) >= f.view(np.int64), x, f))
    assert np.unique(np.concatenate(con)).shape[0] == g["connected_size"] == 59726
    # the product function for the same set-up: trial vector, connected space and its vector in one call
    ti, ta, ci, ca, e_t = host.trial_space(s, trial, orbsym=[int(v) for v in z["orbsym"]], hphf=True)
    assert abs(e_t - g["trial_energy"]) < 1e-12
    assert not host.rows_in(ci, ti).any() and np.array_equal(host.hphf_representative(ci), ci)
    n_zero = g["connected_size"] - trial.shape[0] - ci.shape[0]       # connected functions whose vector entry vanishes
    assert 0 <= n_zero < 0.2 * g["connected_size"]
    k = np.random.default_rng(1).integers(0, ci.shape[0], 40)
    I = np.repeat(np.arange(40), ti.shape[0]); J = np.tile(np.arange(ti.shape[0]), 40)
    want = o.probe_helement(ci[k][I], ti[J]).reshape(40, ti.shape[0]) @ ta
    assert np.allclose(ca[k], want, rtol=1e-11, atol=1e-13)
    # the host library's HPHF elements rank the functions the same way
    hel_h = host.get_helement(s, np.repeat(ref.reshape(1, 1), reps.shape[0], 0), reps, hphf=True)
    assert np.allclose(hel_h, hel, rtol=1e-12, atol=1e-14)


def test_host_hphf_elements_and_core_hamiltonian():
    """The host library's HPHF elements (calc_determ_hamil_sparse_hphf for the stand-alone host) equal the oracle's for
    random pairs of representatives (one- and two-word determinants), and its sparse core Hamiltonian over the 43
    HPHF functions of the reference's HeHe_SS_Doubles core space has the printed lowest eigenvalue -0.0646316671."""
    import test_core_space_cpu as TC
    A, B = 0xAAAAAAAAAAAAAAAA, 0x5555555555555555
    for s in (host.random_fcidump_system(6, 6, sparse=0.9, sparse_t=0.9, seed=3),
              host.random_fcidump_system(33, 8, sparse=0.7, sparse_t=0.7, seed=9)):
        hii = driver.diag_energy(s, s.ref_orbs)
        o, _ = helpers.make_pair(s, hii, max_walkers=1000, max_spawned=1000, hphf=True)
        rng = np.random.default_rng(3)
        sd = host.sing_doub_space(s)
        pick = sd[rng.choice(sd.shape[0], min(120, sd.shape[0]), replace=False)]
        reps = {}
        for row in pick:
            w = row.view(np.uint64)
            f = ((w & np.uint64(A)) >> np.uint64(1)) | ((w & np.uint64(B)) << np.uint64(1))
            r = row if tuple(int(x) for x in row) >= tuple(int(x) for x in f.view(np.int64)) else f.view(np.int64)
            reps[tuple(int(x) for x in r)] = r.copy()
        il = np.array(list(reps.values()), dtype=np.int64).reshape(-1, s.nw)
        n = il.shape[0]
        I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
        hh = host.get_helement(s, il[I], il[J], hphf=True)
        ho = o.probe_helement(il[I], il[J])
        assert np.count_nonzero(ho) > 2 * n
        assert np.allclose(hh, ho, rtol=1e-12, atol=1e-13)
        il2, sizes, displs = host.layout_core_space(il, np.zeros(n, dtype=np.int32), 1)
        c = host.core_hamiltonian(s, il2, hii, hphf=True)
        H = np.zeros((n, n))
        for i in range(n):
            sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
            assert c["col"][sl][-1] == i
            H[i, c["col"][sl]] = c["val"][sl]
        I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
        want = o.probe_helement(il2[I], il2[J]).reshape(n, n) - hii * np.eye(n)
        assert np.allclose(H, want, rtol=1e-12, atol=1e-13)
    g, s = TC._hehe_system()
    hii = driver.diag_energy(s, s.ref_orbs)
    sd = host.sing_doub_space(s, orbsym=g["orbsym"])
    flip = lambda w: ((w & A) >> 1) | ((w & B) << 1)
    reps = np.array([[r[0]] for r in sd if int(np.uint64(r[0])) >= flip(int(np.uint64(r[0])))], dtype=np.int64)
    il, sizes, displs = host.layout_core_space(reps, np.zeros(reps.shape[0], dtype=np.int32), 1)
    c = host.core_hamiltonian(s, il, hii, hphf=True)
    n = il.shape[0]
    H = np.zeros((n, n))
    for i in range(n):
        sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
        H[i, c["col"][sl]] = c["val"][sl]
    assert n == 43 and abs(np.linalg.eigvalsh(H)[0] - g["hphf_run"]["core_correlation_energy"]) < 6e-11
