"""Host-side semi-stochastic set-up (neci_stable_b200/csrc/host/core_space.cpp) against the oracle:
get_helement (src/Determinants.F90:508-554), generate_sing_doub_determinants (src/semi_stoch_gen.F90:537-604),
DetermineDetNode (src/load_balance_calcnodes.F90:25-117) and the sparse core Hamiltonian
(src/sparse_arrays.F90:426-572, src/fast_determ_hamil.F90:1421-1507).  All CPU."""
import numpy as np
import pytest

import helpers
from neci_stable_b200 import capi, host, driver
from neci_stable_b200.capi import ST


def _systems():
    return {
        "1word": host.random_fcidump_system(6, 6, sparse=0.7, sparse_t=0.7, seed=3),
        "odd": host.random_fcidump_system(7, 5, sparse=1.0, sparse_t=1.0, seed=5, ms2=1),
        "2words": host.random_fcidump_system(33, 8, sparse=0.7, sparse_t=0.7, seed=9),
    }


def _iluts(system, dets):
    return np.array([system.ilut(d) for d in dets], dtype=np.int64).reshape(len(dets), system.nw)


@pytest.mark.parametrize("name", ["1word", "odd", "2words"])
def test_host_get_helement_equals_oracle(name):
    s = _systems()[name]
    hii = driver.diag_energy(s, s.ref_orbs)
    o, _ = helpers.make_pair(s, hii, max_walkers=1000, max_spawned=1000)
    rng = np.random.default_rng(1)
    # the reference, a sample of the sector and everything within two excitations of the reference: all of
    # sltcnd_0/1/2 and the "more than a double" zero are hit many times
    sd = host.sing_doub_space(s)
    pick = rng.choice(sd.shape[0], min(60, sd.shape[0]), replace=False)
    il = np.concatenate([sd[pick], _iluts(s, helpers.random_dets(s, 40, rng))])
    n = il.shape[0]
    I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
    h_host = host.get_helement(s, il[I], il[J])
    h_orc = o.probe_helement(il[I], il[J])
    assert np.count_nonzero(h_orc) > n                                   # not a trivial comparison
    assert np.allclose(h_host, h_orc, rtol=1e-12, atol=1e-13)
    assert np.array_equal(h_host == 0.0, h_orc == 0.0)
    H = h_host.reshape(n, n)
    assert np.allclose(H, H.T, rtol=0, atol=1e-13)
    assert abs(H[0, 0] - host.get_helement(s, il[:1], il[:1])[0]) == 0.0


@pytest.mark.parametrize("name", ["1word", "odd", "2words"])
def test_sing_doub_space_is_exactly_the_sector_within_two_excitations(name):
    s = _systems()[name]
    sd = host.sing_doub_space(s)
    ref = s.ilut(s.ref_orbs)
    assert np.array_equal(sd[0], ref)
    keys = {tuple(r) for r in sd.tolist()}
    assert len(keys) == sd.shape[0]                                      # no determinant twice
    if name != "2words":
        want = set()
        for d in helpers.all_dets(s):
            il = s.ilut(d)
            if sum(bin(int(x)).count("1") for x in (il ^ ref).view(np.uint64)) <= 4:
                want.add(tuple(il.tolist()))
        assert keys == want
    else:
        na, va = s.nocc_alpha, s.nbasis // 2 - s.nocc_alpha
        nb, vb = s.nocc_beta, s.nbasis // 2 - s.nocc_beta
        pairs = lambda n: n * (n - 1) // 2
        assert sd.shape[0] == 1 + na * va + nb * vb + pairs(na) * pairs(va) + pairs(nb) * pairs(vb) + na * va * nb * vb
        x = (sd ^ ref).view(np.uint64)
        assert max(sum(bin(int(w)).count("1") for w in row) for row in x) == 4
    # only_keep_conn drops exactly the determinants with no matrix element to the reference (:585-594)
    conn = host.sing_doub_space(s, only_keep_conn=True)
    h = host.get_helement(s, np.repeat(ref[None, :], sd.shape[0], 0), sd)
    keep = np.abs(h) >= 1e-12
    keep[0] = True
    assert np.array_equal(conn, sd[keep])


@pytest.mark.parametrize("name,nranks,bpr", [("1word", 1, 1), ("1word", 4, 1), ("2words", 3, 100), ("odd", 8, 100)])
def test_host_det_node_equals_oracle(name, nranks, bpr):
    s = _systems()[name]
    hii = driver.diag_energy(s, s.ref_orbs)
    o, params = helpers.make_pair(s, hii, max_walkers=1000, max_spawned=1000, nranks=nranks, rank=0, blocks_per_rank=bpr)
    il = _iluts(s, helpers.random_dets(s, 500, np.random.default_rng(2)))
    b_o, n_o = o.probe_det_node(il)
    b_h, n_h = host.det_node(params, il, s.nw)
    assert np.array_equal(b_o, b_h) and np.array_equal(n_o, n_h)
    assert len(set(n_h.tolist())) == nranks


@pytest.mark.parametrize("name,nranks", [("1word", 1), ("1word", 3), ("2words", 2)])
def test_core_hamiltonian_rows(name, nranks):
    """Same non-zero pattern and values as the dense matrix of the oracle's elements; diagonal (H_ii - Hii) last in
    its row; off-diagonal columns ascending; rows split over ranks exactly as sizes/displs say."""
    s = _systems()[name]
    hii = driver.diag_energy(s, s.ref_orbs)
    o, params = helpers.make_pair(s, hii, max_walkers=1000, max_spawned=1000, nranks=nranks, rank=0)
    rng = np.random.default_rng(4)
    sd = host.sing_doub_space(s)
    core = sd[np.sort(rng.choice(sd.shape[0], min(300, sd.shape[0]), replace=False))]
    _, nodes = host.det_node(params, core, s.nw)
    il, sizes, displs = host.layout_core_space(core, nodes, nranks)
    # rank-major, sorted by signed words inside a rank (src/semi_stoch_gen.F90:227)
    _, nodes_sorted = host.det_node(params, il, s.nw)
    assert np.all(np.diff(nodes_sorted) >= 0)
    for r in range(nranks):
        seg = il[displs[r]:displs[r] + sizes[r]]
        assert all(tuple(seg[k]) < tuple(seg[k + 1]) for k in range(seg.shape[0] - 1))
    n = il.shape[0]
    I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
    H = o.probe_helement(il[I], il[J]).reshape(n, n)
    H[np.arange(n), np.arange(n)] -= hii
    total = 0
    for r in range(nranks):
        c = host.core_hamiltonian(s, il, hii, displ=int(displs[r]), n_local=int(sizes[r]), threads=3)
        assert c["row_ptr"][0] == 0 and c["row_ptr"][-1] == c["col"].shape[0] == c["val"].shape[0]
        for k in range(int(sizes[r])):
            i = int(displs[r]) + k
            cols = c["col"][c["row_ptr"][k]:c["row_ptr"][k + 1]]
            vals = c["val"][c["row_ptr"][k]:c["row_ptr"][k + 1]]
            assert cols[-1] == i and np.all(np.diff(cols[:-1]) > 0) and i not in cols[:-1]
            row = np.zeros(n); row[cols] = vals
            want = H[i].copy()
            assert np.array_equal(np.nonzero(row)[0], np.nonzero(want)[0]) or want[i] == 0.0
            assert np.allclose(row, want, rtol=1e-12, atol=1e-13)
        total += int(sizes[r])
    assert total == n
    # thread count must not change the result
    a = host.core_hamiltonian(s, il, hii, threads=1)
    b = host.core_hamiltonian(s, il, hii, threads=8)
    assert all(np.array_equal(a[k], b[k]) for k in ("row_ptr", "col", "val"))


def test_doubles_core_run_with_host_built_hamiltonian_matches_helper_built_one():
    """The engine-facing product of this module, end to end on the oracle: a semi-stochastic run whose core
    Hamiltonian comes from host.core_hamiltonian gives the same walker list as one fed by the test helper's dense
    construction (tests/helpers.py:build_core_space), to the 1e-12 of a different summation order."""
    s = host.random_fcidump_system(6, 4, sparse=1.0, sparse_t=1.0, seed=3)
    hii = driver.diag_energy(s, s.ref_orbs)
    sd = host.sing_doub_space(s)
    dets = [[b + 1 for b in range(s.nbasis) if (int(row[0]) >> b) & 1] for row in sd]
    flags = (1 << capi.FLAG_DETERMINISTIC) | (1 << capi.FLAG_INITIATOR)
    lists = []
    for which in ("helper", "host"):
        o, params = helpers.make_pair(s, hii, max_walkers=20000, max_spawned=20000, semi_stochastic=True, seed=5)
        if which == "helper":
            d2, sizes, displs, per_rank, _ = helpers.build_core_space(o, s, dets, hii)
            c = per_rank[0]; il = c["iluts"]
        else:
            il, sizes, displs = host.layout_core_space(sd, np.zeros(sd.shape[0], dtype=np.int64), 1)
            c = host.core_hamiltonian(s, il, hii)
        recs = np.array([host.record(s, [b + 1 for b in range(s.nbasis) if (int(row[0]) >> b) & 1],
                                     10.0 if np.array_equal(row, s.ilut(s.ref_orbs)) else 0.0, flags) for row in il])
        o.upload_walkers(recs)
        o.set_core_space(c["row_ptr"], c["col"], c["val"], sizes, displs, il)
        for it in range(1, 40):
            st = o.iterate(0.002, 0.0, it)
        lists.append(helpers.canon(*o.download_walkers(), nw=s.nw))
        assert st[ST["NORM_SEMISTOCH_SQ"]] > 0
    a, b = lists
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.allclose(a[1], b[1], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("name", ["1word", "odd"])
def test_trial_space_setup_equals_helper_construction(name):
    """host.trial_space (init_trial_wf) against the test helper's dense construction on the oracle's elements; the
    connected space of the helper is restricted to the list it is given, so it gets the whole sector."""
    s = _systems()[name]
    hii = driver.diag_energy(s, s.ref_orbs)
    o, _ = helpers.make_pair(s, hii, max_walkers=1000, max_spawned=1000)
    dets = helpers.all_dets(s)
    sd = host.sing_doub_space(s)
    trial_il = sd[:12]
    trial = [[b + 1 for b in range(s.nbasis) if (int(row[0]) >> b) & 1] for row in trial_il]
    ti, ta, ci, ca, e_t = helpers.build_trial_space(o, s, dets, trial)
    hi, ha, hci, hca, he = host.trial_space(s, trial_il)
    assert np.array_equal(ti, hi) and abs(e_t - he) < 1e-10
    sgn = np.sign(np.dot(ta, ha))
    assert np.allclose(ta, sgn * ha, atol=1e-9)
    a = {tuple(r): v for r, v in zip(ci.tolist(), ca)}
    b = {tuple(r): v for r, v in zip(hci.tolist(), sgn * hca)}
    # determinants whose connected amplitude is zero only to rounding may be in one list and not the other
    for k in set(a) | set(b):
        assert abs(a.get(k, 0.0) - b.get(k, 0.0)) < 1e-9, k
    assert len(set(a) & set(b)) > 50
    # ham_apply itself: thread count does not matter
    x = host.ham_apply(s, hci, hi, ha, threads=1)
    y = host.ham_apply(s, hci, hi, ha, threads=5)
    assert np.array_equal(x, y) and np.array_equal(x, hca)


def test_bench_semistoch_setup_runs_on_the_oracle():
    """bench.py's semi-stochastic workload set-up (two-word determinants, leading singles+doubles as core space,
    trial space of its largest members) through the oracle: the deterministic projection with the host-built rows
    equals the dense mat-vec over the core space, and the trial estimator sums equal sum_i con_i v_i and
    sum_i psiT_i v_i evaluated directly."""
    import bench
    s = host.random_fcidump_system(40, 20, sparse=0.9, sparse_t=0.9, seed=6)
    hii = driver.diag_energy(s, s.ref_orbs)
    o, params = helpers.make_pair(s, hii, max_walkers=50000, max_spawned=50000, semi_stochastic=True, all_real_coeff=True,
                                  death_before_comms=True)
    space = bench.semistoch_space(s, hii, params, 1, 1500, 4)
    il = space["iluts"]
    assert il.shape == (1500, 2) and np.array_equal(space["trial"][0], s.ilut(s.ref_orbs))
    rec = bench.semistoch_records(s, space, 0, l1_total=5000.0)
    v = rec[:, s.nw].view(np.float64).copy()
    assert abs(np.abs(v).sum() - 5000.0) < 1e-9
    o.upload_walkers(rec)
    nnz, _ = bench.semistoch_apply(o, s, hii, space, 0)
    n = il.shape[0]
    I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
    H = host.get_helement(s, il[I], il[J]).reshape(n, n)
    assert nnz == np.count_nonzero(H)
    tau, S = 1e-5, 0.3
    st = o.iterate(tau, S, 1)
    # trial estimator on the signs the walker loop saw: every trial determinant is a core determinant here
    ti, ta, ci, ca, e_t = host.trial_space(s, space["trial"])
    idx = {tuple(r): k for k, r in enumerate(il.tolist())}
    denom = sum(a * v[idx[tuple(r)]] for r, a in zip(ti.tolist(), ta))
    numer = sum(a * v[idx[tuple(r)]] for r, a in zip(ci.tolist(), ca) if tuple(r) in idx)
    # (the trial determinants' own share is E_T * denom, added by the caller: fcimc_helper.F90:600-613)
    assert np.isclose(st[ST["TRIAL_DENOM"]], denom, rtol=1e-11)
    assert np.isclose(st[ST["TRIAL_NUMERATOR"]], numer, rtol=1e-10)
    # one step of the projection: v <- v - tau (H - Hii - S) v on the core space; stochastic spawns onto core
    # determinants are cancelled, those leaving the core space do not touch it
    d, _, _ = o.download_walkers()
    got = {tuple(r[:2]): x for r, x in zip(d.tolist(), host.signs_of(d, s.nw))}
    Hs = H - hii * np.eye(n)
    want = v - tau * (Hs @ v - S * v)
    have = np.array([got[tuple(r)] for r in il.tolist()])
    assert np.allclose(have, want, rtol=1e-11, atol=1e-11)


def test_bench_semistoch_setup_split_over_ranks_matches_single_rank():
    """The same workload set-up hashed over three ranks with DetermineDetNode (rank-major core order, each rank
    building only its rows): after several iterations the union of the ranks' lists equals the single-rank list
    (the iteration is a function of the walker set)."""
    import bench
    s = host.random_fcidump_system(9, 6, sparse=1.0, sparse_t=1.0, seed=4)
    hii = driver.diag_energy(s, s.ref_orbs)
    results = {}
    for nr in (1, 3):
        orcs = []
        for r in range(nr):
            o, params = helpers.make_pair(s, hii, max_walkers=40000, max_spawned=40000, nranks=nr, rank=r,
                                          semi_stochastic=True, all_real_coeff=True, seed=3)
            orcs.append(o)
        space = bench.semistoch_space(s, hii, params, nr, 400, 0)
        assert int(space["sizes"].sum()) == 400 and (nr == 1 or np.all(space["sizes"] > 0))
        for r in range(nr):
            orcs[r].upload_walkers(bench.semistoch_records(s, space, r, l1_total=3000.0))
            bench.semistoch_apply(orcs[r], s, hii, space, r)
        for it in range(1, 8):
            st = helpers.world_iterate(orcs, 2e-4, 0.1, it, nthreads=1)
        assert st[:, ST["NSPAWNED_SENT"]].sum() > 0
        d = np.concatenate([o.download_walkers()[0] for o in orcs])
        results[nr] = helpers.canon(d, nw=s.nw)
    assert np.array_equal(results[1][0], results[3][0]) and np.array_equal(results[1][2], results[3][2])
    assert np.allclose(results[1][1], results[3][1], rtol=1e-12, atol=1e-12)


def test_doubles_core_size_matches_the_reference_runs():
    """`Total size of deterministic space` printed by two of the reference's regression runs on the same FCIDUMP
    (10 orbitals, 4 electrons, D2h labels): 69 determinants for `semi-stochastic doubles-core` in the determinant basis
    (test_suite/neci/determ_and_trial_spaces/determ_doubles) and 43 HPHF functions for the HPHF run
    (test_suite/neci/parallel/HeHe_SS_Doubles) -- generate_sing_doub_determinants with the point-group symmetry of
    GenExcitations3, then one representative per spin-flipped pair."""
    import json
    import os
    from neci_stable_b200 import fcidump
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hehe_ss_doubles.json")))
    d = fcidump.FciDump(norb=g["norb"], nelec=g["nelec"], ms2=g["ms2"], orbsym=g["orbsym"], ecore=g["ecore"], eps=g["eps"],
                        h1=[tuple(x) for x in g["h1"]], eri=[tuple(x) for x in g["eri"]])
    s = d.system()
    assert [int(x) for x in s.ref_orbs] == g["reference_det"]
    sd = host.sing_doub_space(s, orbsym=g["orbsym"])
    assert sd.shape[0] == g["doubles_core_size_determinants"] == 69
    A, B = 0xAAAAAAAAAAAAAAAA, 0x5555555555555555
    reps = [w for w in (int(np.uint64(r[0])) for r in sd) if w >= (((w & A) >> 1) | ((w & B) << 1))]
    assert len(reps) == g["doubles_core_size_hphf"] == 43
    # symmetry-forbidden excitations carry no matrix element to the reference, and every determinant the symmetry
    # keeps is within the unrestricted space
    full = host.sing_doub_space(s)
    keep = host.rows_in(full, sd)
    assert keep.sum() == 69
    h0 = host.get_helement(s, np.repeat(full[:1], full.shape[0], 0), full)
    assert np.all(np.abs(h0[~keep]) < 1e-10)


def _hehe_system():
    import json
    import os
    from neci_stable_b200 import fcidump
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hehe_ss_doubles.json")))
    d = fcidump.FciDump(norb=g["norb"], nelec=g["nelec"], ms2=g["ms2"], orbsym=g["orbsym"], ecore=g["ecore"], eps=g["eps"],
                        h1=[tuple(x) for x in g["h1"]], eri=[tuple(x) for x in g["eri"]])
    return g, d.system()


def test_core_hamiltonian_and_first_iteration_match_the_reference_run():
    """Numbers the reference printed for its `determ_doubles` regression run (semi-stochastic doubles-core on the HeHe
    FCIDUMP, started from the core ground state scaled to 10000 walkers):
      * 60 doubles and 8 singles found from the reference determinant;
      * `Deterministic subspace correlation energy` -0.0646316671 = lowest eigenvalue of the core Hamiltonian
        (host library: symmetry-adapted doubles-core space + sparse core Hamiltonian) to all printed digits;
      * first line of the iteration table: NoatHF 5827.059, NoatDoubs 4147.905 (the weights of that eigenvector at
        10000 walkers) and Proj.E -0.6463167E-01, reproduced by the eigenvector and by the oracle's SumEContrib in
        an iteration started from it."""
    g, s = _hehe_system()
    ref = g["determ_doubles"]
    hii = driver.diag_energy(s, s.ref_orbs)
    sd = host.sing_doub_space(s, orbsym=g["orbsym"])
    refw = s.ilut(s.ref_orbs)
    level = np.array([bin(int(np.uint64(w[0] ^ refw[0]))).count("1") // 2 for w in sd])
    assert (level == 2).sum() == ref["n_doubles_from_reference"] and (level == 1).sum() == ref["n_singles_from_reference"]
    il, sizes, displs = host.layout_core_space(sd, np.zeros(sd.shape[0], dtype=np.int32), 1)
    c = host.core_hamiltonian(s, il, hii)
    n = il.shape[0]
    H = np.zeros((n, n))
    for i in range(n):
        sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
        H[i, c["col"][sl]] = c["val"][sl]
    assert np.allclose(H, H.T, atol=1e-13)
    w, v = np.linalg.eigh(H)
    assert abs(w[0] - ref["core_correlation_energy"]) < 6e-11                      # printed with 10 decimals
    psi = v[:, 0]
    iref = int(np.nonzero((il == refw).all(axis=1))[0][0])
    psi = psi * np.sign(psi[iref]) * ref["start_walkers"] / np.abs(psi).sum()      # scaled to `startsinglepart` walkers
    lvl = np.array([bin(int(np.uint64(x[0] ^ refw[0]))).count("1") // 2 for x in il])
    assert abs(psi[iref] - ref["step1_no_at_hf"]) < 6e-4                           # printed with 7 significant digits
    assert abs(np.abs(psi[lvl == 2]).sum() - ref["step1_no_at_doubs"]) < 6e-4
    # the same through the oracle: one iteration from that state
    o, _ = helpers.make_pair(s, hii, max_walkers=20000, max_spawned=20000, semi_stochastic=True, all_real_coeff=True,
                             real_spawn_cutoff=0.01, initiator_walk_no=2.0, seed=7)
    flags = (1 << capi.FLAG_DETERMINISTIC) | (1 << capi.FLAG_INITIATOR)
    recs = np.zeros((n, s.nw + 2), dtype=np.int64)
    recs[:, :s.nw] = il
    recs[:, s.nw] = psi.view(np.int64)
    recs[:, s.nw + 1] = flags
    o.upload_walkers(recs)
    o.set_core_space(c["row_ptr"], c["col"], c["val"], sizes, displs, il)
    st = o.iterate(ref["tau"], 0.0, 1)
    assert abs(st[ST["HFCYC"]] - ref["step1_no_at_hf"]) < 6e-4
    assert abs(st[ST["NOATDOUBS"]] - ref["step1_no_at_doubs"]) < 6e-4
    assert abs(st[ST["ENUMCYC"]] / st[ST["HFCYC"]] - ref["step1_proj_e"]) < 6e-9  # -0.6463167E-01
    # Second and third line of the table.  Every determinant connected to the reference is a core determinant, so
    # the reference amplitude evolves deterministically; the doubles do too during the first iteration (nothing
    # outside the core space exists yet).  The shift printed in line k is the one computed at the END of iteration k,
    # so iterations 1 and 2 both run at S = 0 (real coefficients: determ_projection_no_death + death of every
    # determinant, FciMCPar.F90:1778-1782, 1842-1846).
    st2 = o.iterate(ref["tau"], 0.0, 2)
    assert abs(st2[ST["HFCYC"]] - ref["step2_no_at_hf"]) < 6e-4                    # 5830.825
    assert abs(st2[ST["NOATDOUBS"]] - ref["step2_no_at_doubs"]) < 6e-4            # 4150.586
    st3 = o.iterate(ref["tau"], ref["step2_shift"], 3)
    assert abs(st3[ST["HFCYC"]] - ref["step3_no_at_hf"]) < 6e-4                    # 5834.594
    # and the closed form behind them: one step multiplies the core ground state by 1 - tau (E_core - S)
    assert abs(ref["step1_no_at_hf"] * (1.0 - ref["tau"] * w[0]) - ref["step2_no_at_hf"]) < 1e-3


def test_fci_core_space_gives_the_reference_fci_correlation_energy():
    """`semi-stochastic fci-core` of the reference on the HeHe FCIDUMP (HeHe_determ): 309 determinants = the Ms = 0
    sector restricted to the irrep of the reference determinant, and `Deterministic subspace correlation energy`
    -0.0650928511 = the exact ground-state energy of that block minus the reference energy.  Reproduced from the
    host library's sparse Hamiltonian over the symmetry-filtered sector: every matrix element of the system takes
    part, to the 10 printed decimals."""
    g, s = _hehe_system()
    ref = g["fci_core"]
    hii = driver.diag_energy(s, s.ref_orbs)
    irr = lambda orbs: int(np.bitwise_xor.reduce([g["orbsym"][(o + 1) // 2 - 1] - 1 for o in orbs]))
    target = irr([int(x) for x in s.ref_orbs])
    dets = [d for d in helpers.all_dets(s) if irr(d) == target]
    assert len(dets) == ref["size"] == 309
    il = np.array([s.ilut(d) for d in dets], dtype=np.int64).reshape(len(dets), s.nw)
    il, sizes, displs = host.layout_core_space(il, np.zeros(len(dets), dtype=np.int32), 1)
    c = host.core_hamiltonian(s, il, hii)
    n = il.shape[0]
    H = np.zeros((n, n))
    for i in range(n):
        sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
        H[i, c["col"][sl]] = c["val"][sl]
    w = np.linalg.eigvalsh(H)
    assert abs(w[0] - ref["correlation_energy"]) < 6e-11
    # the oracle's own elements (the checker of the CUDA path) give the same number
    o, _ = helpers.make_pair(s, hii, max_walkers=1000, max_spawned=1000)
    I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
    Ho = o.probe_helement(il[I], il[J]).reshape(n, n) - hii * np.eye(n)
    assert abs(np.linalg.eigvalsh(Ho)[0] - ref["correlation_energy"]) < 6e-11
    assert np.allclose(Ho, H, rtol=1e-12, atol=1e-13)
    # consistent with the projected energy the HPHF run of the same system converged to (statistical)
    assert abs(hii + w[0] - g["total_projected_energy"]) < 5 * g["total_projected_energy_error"] + 1e-4


def test_read_core_space_gives_the_reference_core_energy():
    """`semi-stochastic read-core` of the reference on the HeHe FCIDUMP (determ_read): the 100 determinants of its
    checked-in CORESPACE file -- up to quadruple excitations of the reference, so the "more than a double excitation
    apart" branch of the builder is exercised -- and the printed lowest eigenvalue -0.0650248789."""
    g, s = _hehe_system()
    ref = g["read_core"]
    hii = driver.diag_energy(s, s.ref_orbs)
    il = np.array(ref["iluts"], dtype=np.int64).reshape(-1, 1)
    assert il.shape[0] == ref["size"] == 100 and len(set(ref["iluts"])) == 100
    assert all(bin(x).count("1") == s.nel for x in ref["iluts"])
    refw = int(s.ilut(s.ref_orbs)[0])
    levels = {bin(x & ~refw).count("1") for x in ref["iluts"]}
    assert max(levels) > 2
    il, sizes, displs = host.layout_core_space(il, np.zeros(100, dtype=np.int32), 1)
    c = host.core_hamiltonian(s, il, hii)
    H = np.zeros((100, 100))
    for i in range(100):
        sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
        H[i, c["col"][sl]] = c["val"][sl]
    assert abs(np.linalg.eigvalsh(H)[0] - ref["correlation_energy"]) < 6e-11


@pytest.mark.parametrize("which", ["trial_doubles", "trial_read"])
def test_trial_space_setup_matches_the_reference_runs(which):
    """init_trial_wf of the reference on the HeHe FCIDUMP: `doubles-trial` (69 determinants) and `read-trial` (the 100
    determinants of the TRIALSPACE file).  host.trial_space gives the printed `Energy eigenvalue(s) of the trial
    space` (-5.7617147894252971 / -5.7621080012308328) to 1e-12; the trial determinants together with their
    symmetry-allowed singles and doubles are the printed 309 members of the connected space (here the whole symmetry
    sector); and the connected-space vector is H psi_T restricted to it."""
    g, s = _hehe_system()
    ref = g["trial_runs"][which]
    if which == "trial_doubles":
        trial = host.sing_doub_space(s, orbsym=g["orbsym"])
    else:
        trial = np.array(g["read_core"]["iluts"], dtype=np.int64).reshape(-1, 1)
    assert trial.shape[0] == ref["trial_size"]
    ti, ta, ci, ca, e_t = host.trial_space(s, trial, orbsym=g["orbsym"])
    assert abs(e_t - ref["trial_energy"]) < 1e-12
    union = np.unique(np.concatenate([trial] + [host.sing_doub_space(s, ref_ilut=r, orbsym=g["orbsym"]) for r in trial]), axis=0)
    assert union.shape[0] == ref["connected_size"] == 309
    # con_space_vecs = sum_j H_ij psiT_j for every connected determinant outside the trial space
    assert not host.rows_in(ci, ti).any() and ci.shape[0] <= 309 - trial.shape[0]
    I = np.repeat(np.arange(ci.shape[0]), ti.shape[0]); J = np.tile(np.arange(ti.shape[0]), ci.shape[0])
    Hct = host.get_helement(s, ci[I], ti[J]).reshape(ci.shape[0], ti.shape[0])
    assert np.allclose(Hct @ ta, ca, rtol=1e-12, atol=1e-14)
    # <psiT|H|psiT> = E_T and (H psiT)_i = E_T psiT_i inside the trial space
    I = np.repeat(np.arange(ti.shape[0]), ti.shape[0]); J = np.tile(np.arange(ti.shape[0]), ti.shape[0])
    Htt = host.get_helement(s, ti[I], ti[J]).reshape(ti.shape[0], ti.shape[0])
    assert np.allclose(Htt @ ta, e_t * ta, atol=1e-11)


def test_cas_core_and_cas_trial_match_the_reference_runs():
    """`cas-core 2 6` and `cas-trial 2 6` of the reference on the HeHe FCIDUMP: 8 determinants, printed core
    correlation energy -0.0095421747, printed trial energy -5.7066252970297464 (to 1e-12) and 187 members of the
    connected space."""
    g, s = _hehe_system()
    ref, tref = g["cas_core"], g["trial_runs"]["trial_cas"]
    hii = driver.diag_energy(s, s.ref_orbs)
    cas = host.cas_space(s, g["eps"], ref["cas"][0], ref["cas"][1], orbsym=g["orbsym"])
    assert cas.shape[0] == ref["size"] == tref["trial_size"] == 8
    assert host.rows_in(s.ilut(s.ref_orbs).reshape(1, -1), cas)[0]
    il, sizes, displs = host.layout_core_space(cas, np.zeros(8, dtype=np.int32), 1)
    c = host.core_hamiltonian(s, il, hii)
    H = np.zeros((8, 8))
    for i in range(8):
        sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
        H[i, c["col"][sl]] = c["val"][sl]
    assert abs(np.linalg.eigvalsh(H)[0] - ref["correlation_energy"]) < 6e-11
    ti, ta, ci, ca, e_t = host.trial_space(s, cas, orbsym=g["orbsym"])
    assert abs(e_t - tref["trial_energy"]) < 1e-12
    union = np.unique(np.concatenate([cas] + [host.sing_doub_space(s, ref_ilut=r, orbsym=g["orbsym"]) for r in cas]), axis=0)
    assert union.shape[0] == tref["connected_size"] == 187


@pytest.mark.parametrize("which", ["determ_opt_num", "determ_opt_amp", "trial_opt_num", "trial_opt_amp"])
def test_optimised_spaces_match_the_reference_runs(which):
    """`optimised-core` / `optimised-trial` of the reference on the HeHe FCIDUMP (connected space -> ground state ->
    keep by number 3/6/60 and 4/20/80, or by amplitude 0.02/0.002 and 0.001): sizes 60 / 44 / 80 / 53 and the printed
    core correlation energies (-0.0647917270, -0.0646041982) and trial energies (-5.7620482067956624,
    -5.7617005825505965, to 1e-12) from host.optimised_space + core_hamiltonian / trial_space."""
    g, s = _hehe_system()
    hii = driver.diag_energy(s, s.ref_orbs)
    ref = g["optimised_core"][which] if which.startswith("determ") else g["trial_runs"][which]
    kind, cuts = ref["cutoff"][0], ref["cutoff"][1:]
    space = host.optimised_space(s, cutoff_num=[int(x) for x in cuts] if kind == "num" else None,
                                 cutoff_amp=cuts if kind == "amp" else None, orbsym=g["orbsym"])
    if which.startswith("determ"):
        assert space.shape[0] == ref["size"]
        il, sizes, displs = host.layout_core_space(space, np.zeros(space.shape[0], dtype=np.int32), 1)
        c = host.core_hamiltonian(s, il, hii)
        n = il.shape[0]
        H = np.zeros((n, n))
        for i in range(n):
            sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
            H[i, c["col"][sl]] = c["val"][sl]
        assert abs(np.linalg.eigvalsh(H)[0] - ref["correlation_energy"]) < 6e-11
    else:
        assert space.shape[0] == ref["trial_size"]
        e_t = host.trial_space(s, space, orbsym=g["orbsym"])[4]
        assert abs(e_t - ref["trial_energy"]) < 1e-12


def test_ras_core_matches_the_reference_run():
    """`ras-core 2 0 8 1 3` of the reference on the HeHe FCIDUMP: 197 determinants and the printed core correlation
    energy -0.0646451087 from host.ras_space + core_hamiltonian."""
    g, s = _hehe_system()
    ref = g["ras_core"]
    hii = driver.diag_energy(s, s.ref_orbs)
    space = host.ras_space(s, g["eps"], *ref["ras"], orbsym=g["orbsym"])
    assert space.shape[0] == ref["size"] == 197
    il, sizes, displs = host.layout_core_space(space, np.zeros(space.shape[0], dtype=np.int32), 1)
    c = host.core_hamiltonian(s, il, hii)
    n = il.shape[0]
    H = np.zeros((n, n))
    for i in range(n):
        sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
        H[i, c["col"][sl]] = c["val"][sl]
    assert abs(np.linalg.eigvalsh(H)[0] - ref["correlation_energy"]) < 6e-11


def test_determ_doubles_run_on_the_oracle_agrees_with_the_reference_run():
    """The reference's determ_doubles run as a whole (semi-stochastic doubles-core, real coefficients, cutoff 0.01,
    initiator threshold 2, tau 0.01, shift damping 0.5 every iteration, 10000 walkers, 400 iterations from the core
    ground state), set up with the host library only and run on the oracle: the projected correlation energy agrees
    with the reference's -0.065081043 +/- 8.8e-6 within the combined blocking errors, and with the exact
    -0.0650928511 of the fci-core run within the initiator bias the reference run itself shows."""
    e, err, hist, ref, g = determ_doubles_run(lambda s, params: helpers.Oracle(params))
    tol = 4.0 * np.hypot(err, ref["projected_correlation_energy_error"])
    assert err < 5e-5
    assert abs(e - ref["projected_correlation_energy"]) < tol, (e, err, ref["projected_correlation_energy"])
    assert abs(e - g["fci_core"]["correlation_energy"]) < tol + 2.5e-5
    assert 0.8 * ref["total_walkers"] < hist[-1]["tot_parts"] < 1.3 * ref["total_walkers"]


def determ_doubles_run(make_engine):
    """The determ_doubles case on an engine made by make_engine(system, params) (oracle here, CUDA engine in
    tests/test_gpu_energies.py): returns (projected correlation energy, blocking error, history, golden entry, golden)."""
    g, s = _hehe_system()
    ref = g["determ_doubles"]
    hii = driver.diag_energy(s, s.ref_orbs)
    sd = host.sing_doub_space(s, orbsym=g["orbsym"])
    il, sizes, displs = host.layout_core_space(sd, np.zeros(sd.shape[0], dtype=np.int32), 1)
    c = host.core_hamiltonian(s, il, hii)
    n = il.shape[0]
    H = np.zeros((n, n))
    for i in range(n):
        sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
        H[i, c["col"][sl]] = c["val"][sl]
    w, v = np.linalg.eigh(H)
    iref = int(np.nonzero((il == s.ilut(s.ref_orbs)).all(axis=1))[0][0])
    psi = v[:, 0] * np.sign(v[iref, 0]) * ref["start_walkers"] / np.abs(v[:, 0]).sum()
    params = host.make_params(s, hii, max_walkers=200000, max_spawned=200000, semi_stochastic=True, all_real_coeff=True,
                              real_spawn_cutoff=ref["real_spawn_cutoff"], initiator_walk_no=ref["add_to_initiator"], seed=7)
    o = make_engine(s, params)
    s.apply(o)
    recs = np.zeros((n, s.nw + 2), dtype=np.int64)
    recs[:, :s.nw] = il
    recs[:, s.nw] = psi.view(np.int64)
    recs[:, s.nw + 1] = (1 << capi.FLAG_DETERMINISTIC) | (1 << capi.FLAG_INITIATOR)
    o.upload_walkers(recs)
    o.set_core_space(c["row_ptr"], c["col"], c["val"], sizes, displs, il)
    run = driver.FciMC(s, o, hii, tau=ref["tau"], init_walkers=ref["total_walkers"], steps_sft=ref["steps_shift"],
                       sft_damp=ref["shift_damp"], diag_sft=0.0)
    run.tot_parts = ref["start_walkers"]; run.old_av_walkers = ref["start_walkers"]
    hist = run.run(4 * ref["nmcyc"])
    rows = [h for h in hist if h["varying"]][100:]
    e, err = driver.ratio_estimate([h["enum_cyc"] for h in rows], [h["hf_cyc"] for h in rows])
    o.close()
    return e, err, hist, ref, g


def test_pops_core_from_a_running_list_and_switch_to_semi_stochastic():
    """`pops-core`: after a stochastic warm-up on the oracle the most populated determinants become the core space
    (host.most_populated_space on the downloaded list), their rows are built by the host library and the run continues
    semi-stochastically -- the dynamic core-space flow of the reference (semistoch-shift-iter) with this repo's host."""
    s = host.random_fcidump_system(6, 6, sparse=0.9, sparse_t=0.9, seed=3)
    hii = driver.diag_energy(s, s.ref_orbs)
    o, _ = helpers.make_pair(s, hii, max_walkers=50000, max_spawned=50000, semi_stochastic=True, all_real_coeff=True, seed=5)
    o.upload_walkers(host.record(s, s.ref_orbs, 2000.0, 1 << capi.FLAG_INITIATOR).reshape(1, -1))
    for it in range(1, 60):
        o.iterate(0.002, 0.0, it)
    d, gd, go = o.download_walkers()
    d = d[np.abs(host.signs_of(d, s.nw)) > 0]
    core, amps = host.most_populated_space(d, 40, nw=s.nw)
    assert core.shape[0] == 40 and np.all(np.diff(np.abs(amps)) <= 0)
    assert np.abs(amps[-1]) >= np.sort(np.abs(host.signs_of(d, s.nw)))[-40]
    il, sizes, displs = host.layout_core_space(core, np.zeros(40, dtype=np.int32), 1)
    c = host.core_hamiltonian(s, il, hii)
    # flag the chosen determinants in the list and hand the core space over
    keys = {tuple(r) for r in il.tolist()}
    for k in range(d.shape[0]):
        if tuple(d[k, :s.nw].tolist()) in keys:
            d[k, s.nw + 1] |= (1 << capi.FLAG_DETERMINISTIC) | (1 << capi.FLAG_INITIATOR)
    o.upload_walkers(d)
    o.set_core_space(c["row_ptr"], c["col"], c["val"], sizes, displs, il)
    tot = np.abs(host.signs_of(d, s.nw)).sum()
    for it in range(60, 80):
        st = o.iterate(0.002, 0.0, it)
    assert st[ST["NORM_SEMISTOCH_SQ"]] > 0.5 * st[ST["NORM_PSI_SQ"]]       # the core space carries most of the norm
    assert 0.5 * tot < st[ST["TOTPARTS"]] < 3.0 * tot


@pytest.mark.parametrize("kind", ["hub_k", "hub_rs"])
def test_host_lattice_elements_and_core_hamiltonian(kind):
    """The host library's elements and sparse Hamiltonian for the lattice models (k-space: get_umat_kspace rule inside
    the Slater-Condon routines; real space: on-site <ii|ii> = U beside the hopping matrix) equal the oracle's
    get_helement_lattice for random determinant pairs and over a random determinant list."""
    s = host.hubbard_k_system(4, 4, U=4.0) if kind == "hub_k" else host.hubbard_rs_system(4, 4, U=4.0)
    hii = driver.diag_energy(s, s.ref_orbs)
    o, _ = helpers.make_pair(s, hii, max_walkers=1000, max_spawned=1000)
    rng = np.random.default_rng(6)
    dets = helpers.random_dets(s, 150, rng)
    ref = [int(x) for x in s.ref_orbs]
    # neighbours of the reference so that singles / doubles occur among the pairs
    near = []
    for _ in range(100):
        d = list(ref)
        for _k in range(int(rng.integers(1, 3))):
            i = int(rng.integers(0, len(d)))
            cand = [x for x in range(1, s.nbasis + 1) if x not in d and (x & 1) == (d[i] & 1)]
            d[i] = int(rng.choice(cand))
        near.append(sorted(d))
    il = np.unique(np.array([s.ilut(d) for d in dets + near + [ref]], dtype=np.int64).reshape(-1, s.nw), axis=0)
    n = il.shape[0]
    I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
    hh = host.get_helement(s, il[I], il[J])
    ho = o.probe_helement(il[I], il[J])
    assert np.count_nonzero(ho) > n
    assert np.allclose(hh, ho, rtol=1e-12, atol=1e-13)
    il2, sizes, displs = host.layout_core_space(il, np.zeros(n, dtype=np.int32), 1)
    c = host.core_hamiltonian(s, il2, hii)
    H = np.zeros((n, n))
    for i in range(n):
        sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
        H[i, c["col"][sl]] = c["val"][sl]
    want = o.probe_helement(il2[I], il2[J]).reshape(n, n) - hii * np.eye(n)
    assert np.allclose(H, want, rtol=1e-12, atol=1e-13)
