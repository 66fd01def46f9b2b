"""Pins the CPU oracle (oracle/, test infrastructure) before anything is compared with it:

  * known-answer vectors transcribed from the reference's own unit tests
    (tests/golden/reference_known_answers.json, file:line cited there),
  * an independent second-quantised construction of H (tests/bruteforce.py),
  * the reference's statistical generator tests (alias-table L1 test, PCHB sum 1/pgen test),
  * exact diagonalisation against energies the reference's regression suite publishes.

CPU only."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import bruteforce
import helpers
from neci_stable_b200 import capi, host, driver

with open(os.path.join(helpers.GOLDEN, "reference_known_answers.json")) as f:
    GOLD = json.load(f)


# ---- hand-built lattices in the reference's own orbital order ---------------------------------
def chain_rs_system(length, periodic, bhub, uhub, nel):
    """lattice('chain', L, ...) + init_tmat: tmat(i,j) = bhub between neighbouring sites, same spin."""
    nb = 2 * length
    neigh = np.zeros((nb, 4), dtype=np.int32)
    tmat = np.zeros((nb, nb))
    for s in range(length):
        cand = []
        for d in (-1, 1):
            x = s + d
            if periodic:
                x %= length
            elif x < 0 or x >= length:
                continue
            if x != s and x not in cand:
                cand.append(x)
        for spin in (0, 1):
            o = 2 * (s + 1) - spin
            for q, s2 in enumerate(sorted(cand)):
                o2 = 2 * (s2 + 1) - spin
                neigh[o - 1, q] = o2
                tmat[o - 1, o2 - 1] = bhub
    return host.System(kind=capi.SYS_HUBBARD_RS, nel=nel, nbasis=nb, nocc_alpha=nel // 2, nocc_beta=nel - nel // 2,
                       t_exch=0, t_no_brillouin=1,
                       tables=dict(max_neigh=4, neighbours=neigh.ravel(), tmat=tmat.T.ravel().copy(), uhub=float(uhub)),
                       ref_orbs=np.arange(1, nel + 1, dtype=np.int32))


def square_rs_system(lx, ly, bhub, uhub, nel, t_scale=1.0):
    s = host.hubbard_rs_system(lx, ly, nel=nel, U=uhub, t=-bhub * t_scale)
    return s


def k_chain_system(kvals, bhub, u_over_n, nel, length):
    """k-space chain in the reference's orbital order: spatial orbital i has momentum kvals[i-1]."""
    nk = len(kvals)
    idx = {k % length: i for i, k in enumerate(kvals)}
    ksum = np.zeros((nk, nk), dtype=np.int32); kdiff = np.zeros((nk, nk), dtype=np.int32)
    for a, ka in enumerate(kvals):
        for b, kb in enumerate(kvals):
            ksum[a, b] = idx[(ka + kb) % length]
            kdiff[a, b] = idx[(ka - kb) % length]
    eps = np.array([2.0 * bhub * np.cos(2 * np.pi * k / length) for k in kvals])
    return host.System(kind=capi.SYS_HUBBARD_K, nel=nel, nbasis=2 * nk, nocc_alpha=nel // 2, nocc_beta=nel - nel // 2,
                       t_exch=1, t_no_brillouin=0,
                       tables=dict(n_k=nk, ksum=ksum.ravel(), kdiff=kdiff.ravel(), eps_k=eps, u_over_n=float(u_over_n)),
                       ref_orbs=np.arange(1, nel + 1, dtype=np.int32))


def oracle_for(system, **kw):
    hii = 0.0
    kw.setdefault("max_walkers", 1000)
    kw.setdefault("max_spawned", 1000)
    o, params = helpers.make_pair(system, hii, **kw)
    return o


def dbl(lib, name, *args):
    out = C.c_double(0.0)
    f = getattr(lib, name); f.restype = C.c_int
    assert f(*args, C.byref(out)) == 0
    return out.value


# ---- Philox -------------------------------------------------------------------------------------
def test_philox_known_answers():
    lib = helpers.oracle_lib()
    for c in GOLD["philox4x32_10"]["cases"]:
        ctr = np.array([int(x, 16) for x in c["ctr"]], dtype=np.uint32)
        key = np.array([int(x, 16) for x in c["key"]], dtype=np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        lib.orc_probe_philox(ctr.ctypes.data_as(C.c_void_p), key.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        assert ["%08x" % v for v in out] == c["out"]


def test_stream_is_uniform_and_in_range():
    lib = helpers.oracle_lib()
    il = np.array([0x5555], dtype=np.int64)
    out = np.zeros(20000)
    lib.orc_probe_stream(C.c_uint64(7), C.c_int64(3), il.ctypes.data_as(C.c_void_p), C.c_int32(1), C.c_int32(0),
                         C.c_int32(1), C.c_int32(out.size), out.ctypes.data_as(C.c_void_p))
    assert out.min() >= 0.0 and out.max() < 1.0
    assert abs(out.mean() - 0.5) < 0.01 and abs(out.var() - 1 / 12) < 0.005


# ---- make_double / parity conventions --------------------------------------------------------------
def test_make_double_reference_cases():
    lib = helpers.oracle_lib()
    for c in GOLD["make_double"]["cases"]:
        nI = np.array(c["nI"], dtype=np.int32)
        nJ = np.zeros_like(nI); ex = np.zeros(4, dtype=np.int32); par = C.c_int32(0)
        lib.orc_probe_make_double(C.c_int32(nI.size), nI.ctypes.data_as(C.c_void_p), C.c_int32(c["elecs"][0]),
                                  C.c_int32(c["elecs"][1]), C.c_int32(c["tgt"][0]), C.c_int32(c["tgt"][1]),
                                  nJ.ctypes.data_as(C.c_void_p), ex.ctypes.data_as(C.c_void_p), C.byref(par))
        assert list(nJ) == c["nJ"], c
        assert list(ex) == c["ex"], c
        assert bool(par.value) == c["tpar"], c


# ---- real-space Hubbard known answers -----------------------------------------------------------------
def _lattice(spec, uhub, nel):
    if spec["type"] == "chain":
        return chain_rs_system(spec["length"], spec["periodic"], spec["bhub"], uhub, nel)
    return square_rs_system(spec["lx"], spec["ly"], spec["bhub"], uhub, nel)


def test_rs_hubbard_offdiag_reference_cases():
    g = GOLD["rs_hubbard_offdiag"]
    o = oracle_for(_lattice(g["lattice"], 0.0, 2))
    for c in g["cases"]:
        v = dbl(o.lib, "orc_probe_offdiag_rs", o.h, C.c_int32(c["ex"][0]), C.c_int32(c["ex"][1]), C.c_int32(int(c["tpar"])))
        assert v == c["hel"], c


def test_rs_hubbard_get_helement_reference_cases():
    for blk in GOLD["rs_hubbard_get_helement"]["blocks"]:
        by_nel = {}
        for c in blk["cases"]:
            by_nel.setdefault(len(c["nI"]), []).append(c)
        for nel, cases in by_nel.items():
            s = _lattice(blk["lattice"], blk["uhub"], nel)
            assert s.nbasis == blk["nbasis"]
            o = oracle_for(s)
            I = np.array([s.ilut(c["nI"]) for c in cases]); J = np.array([s.ilut(c["nJ"]) for c in cases])
            h = o.probe_helement(I, J)
            assert list(h) == [c["hel"] for c in cases], (blk["lattice"], cases, h)


# ---- k-space Hubbard known answers ----------------------------------------------------------------------
def test_k_hubbard_reference_cases():
    g = GOLD["k_hubbard"]
    lat = g["lattice"]
    for c in g["diag"]:
        s = k_chain_system(lat["k_of_spatial_orbital"], lat["bhub"], c["u_over_n"], len(c["nI"]), lat["length"])
        o = oracle_for(s)
        il = s.ilut(c["nI"])
        v = dbl(o.lib, "orc_probe_diag", o.h, il.ctypes.data_as(C.c_void_p))
        assert abs(v - c["hel"]) < 1e-14, (c, v)
    s = k_chain_system(lat["k_of_spatial_orbital"], lat["bhub"], g["uhub"] / g["omega"], 2, lat["length"])
    o = oracle_for(s)
    for c in g["offdiag"]:
        ex = np.array(c["ex"], dtype=np.int32)
        if ex[2] == 0:
            ex[2:] = [2, 4]
        v = dbl(o.lib, "orc_probe_offdiag_k", o.h, ex.ctypes.data_as(C.c_void_p), C.c_int32(int(c["tpar"])))
        assert v == c["hel"], (c, v)


# ---- integral indexing ---------------------------------------------------------------------------------------
def test_umat_ind_eightfold_symmetry_and_packing():
    """UMatInd (src/UMatCache.F90:257-296): <ij|kl> = <kj|il> = <il|kj> = <ji|lk> ..., dense packing 1..N."""
    lib = helpers.oracle_lib()
    lib.orc_probe_umat_ind.restype = C.c_int64
    f = lambda i, j, k, l: lib.orc_probe_umat_ind(C.c_int32(i), C.c_int32(j), C.c_int32(k), C.c_int32(l))
    n = 5
    seen = set()
    for i in range(1, n + 1):
        for j in range(1, n + 1):
            for k in range(1, n + 1):
                for l in range(1, n + 1):
                    a = f(i, j, k, l)
                    assert a == f(k, j, i, l) == f(i, l, k, j) == f(k, l, i, j) == f(j, i, l, k) == f(l, k, j, i)
                    seen.add(a)
    npair = n * (n + 1) // 2
    assert seen == set(range(1, npair * (npair + 1) // 2 + 1))
    # fixed values: <11|11> = 1, <12|12> -> pairs (1,1),(2,2) = tri 1 and 3 -> 3*2/2+1 = 4
    assert f(1, 1, 1, 1) == 1 and f(1, 2, 1, 2) == 4 and f(1, 2, 2, 1) == 3


# ---- Slater-Condon rules vs. independent second quantisation -----------------------------------------------------
def _h_oracle(o, s, dets):
    return helpers.hamiltonian_matrix(o, s, dets)


def test_sltcnd_against_second_quantisation():
    s = host.random_fcidump_system(4, 4, sparse=0.9, sparse_t=0.9, seed=5)
    o = oracle_for(s)
    dets = helpers.all_dets(s)
    H = _h_oracle(o, s, dets)
    t = s.tables
    nb = s.nbasis

    def tri(a, b):
        return a * (a - 1) // 2 + b if a > b else b * (b - 1) // 2 + a

    h1 = lambda p, q: t["tmat"][(p - 1) + nb * (q - 1)]
    eri = lambda i, j, k, l: t["umat"][tri(tri(i, j), tri(k, l)) - 1]     # chemist (ij|kl) = <ik|jl>
    Hb = bruteforce.hamiltonian([tuple(d) for d in dets], nb, h1, eri)
    assert np.allclose(H, H.T, atol=1e-13)
    assert np.allclose(H, Hb, rtol=1e-12, atol=1e-12)
    assert np.count_nonzero(np.abs(Hb) > 1e-9) > len(dets) * 5


def test_sltcnd_two_word_determinants():
    """nbasis > 64 (nIfD = 1): same rules across the word boundary, checked on the subspace reachable from the
    reference via one oracle excitation each (second quantisation on a determinant subset)."""
    s = host.random_fcidump_system(33, 4, sparse=0.8, sparse_t=0.8, seed=4)
    o = oracle_for(s)
    rng = np.random.default_rng(0)
    dets = [list(s.ref_orbs)]
    # singles and doubles that straddle bit 64
    for _ in range(40):
        a = sorted(rng.choice(33, 2, replace=False) + 1); b = sorted(rng.choice(33, 2, replace=False) + 1)
        dets.append(sorted([2 * int(i) for i in a] + [2 * int(i) - 1 for i in b]))
    dets = [list(x) for x in sorted(set(tuple(d) for d in dets))]
    H = _h_oracle(o, s, dets)
    t = s.tables; nb = s.nbasis

    def tri(a, b):
        return a * (a - 1) // 2 + b if a > b else b * (b - 1) // 2 + a

    h1 = lambda p, q: t["tmat"][(p - 1) + nb * (q - 1)]
    eri = lambda i, j, k, l: t["umat"][tri(tri(i, j), tri(k, l)) - 1]
    Hb = bruteforce.hamiltonian([tuple(d) for d in dets], nb, h1, eri)
    assert np.allclose(H, Hb, rtol=1e-12, atol=1e-12)


def test_hubbard_hamiltonians_against_second_quantisation():
    # real space 2x3, 4 electrons
    s = host.hubbard_rs_system(3, 2, nel=4, U=4.0)
    o = oracle_for(s)
    dets = helpers.all_dets(s)
    t = s.tables; nb = s.nbasis
    h1 = lambda p, q: t["tmat"][(p - 1) + nb * (q - 1)]
    eri = lambda i, j, k, l: t["uhub"] if (i == j == k == l) else 0.0
    H = _h_oracle(o, s, dets)
    Hb = bruteforce.hamiltonian([tuple(d) for d in dets], nb, h1, eri)
    assert np.allclose(H, Hb, atol=1e-13)
    # k space 2x2 and a 4-site chain, 4 electrons
    for s in (host.hubbard_k_system(2, 2, nel=4, U=2.0), host.hubbard_k_system(4, 1, nel=4, U=3.0)):
        o = oracle_for(s)
        dets = helpers.all_dets(s)
        t = s.tables; nb = s.nbasis; nk = t["n_k"]
        ks = t["ksum"].reshape(nk, nk)
        h1 = lambda p, q: t["eps_k"][(p - 1) // 2] if p == q else 0.0
        # (ij|kl) = <ik|jl> = U/N [k_i + k_k == k_j + k_l]
        eri = lambda i, j, k, l: t["u_over_n"] if ks[i - 1, k - 1] == ks[j - 1, l - 1] else 0.0
        H = _h_oracle(o, s, dets)
        Hb = bruteforce.hamiltonian([tuple(d) for d in dets], nb, h1, eri)
        assert np.allclose(H, Hb, atol=1e-13)


# ---- exact diagonalisation vs. energies published in the reference's regression suite ---------------------------------
def test_exact_diagonalisation_matches_reference_energies():
    g = GOLD["end_to_end_energies"]
    # k-space 2x2, U = 1, 4 electrons (test_suite/neci/parallel/Hubbard_2x2): -7.29750728 +- 2.8e-4
    k = g["hubbard_2x2_kspace"]
    s = host.hubbard_k_system(k["lx"], k["ly"], nel=k["nel"], U=k["U"], t=k["t"])
    o = oracle_for(s)
    H = _h_oracle(o, s, helpers.all_dets(s))
    e0 = np.linalg.eigvalsh(H)[0]
    assert abs(e0 - k["reference_fciqmc"]) < 4 * k["err"], e0
    # real-space periodic 2x2 of the old lattice code (bonds doubled), U = 16, 4 electrons: shift estimate -2.63628 +- 2.6e-3
    r = g["hubbard_2x2_realspace"]
    s = host.hubbard_rs_system(r["lx"], r["ly"], nel=r["nel"], U=r["U"], t=r["t_effective"])
    o = oracle_for(s)
    H = _h_oracle(o, s, helpers.all_dets(s))
    e0 = np.linalg.eigvalsh(H)[0]
    assert abs(e0 - r["reference_fciqmc"]) < 4 * r["err"], e0


# ---- alias tables: the reference's L1 test (unit_tests/sampler/test_aliasTables.F90:45-110) ------------------------------
def test_alias_table_l1_distance():
    lib = helpers.oracle_lib()
    rng = np.random.default_rng(11)
    n = 10                                              # huge_number-free version of the reference's setup
    w = rng.random(n)
    w[3] = 0.0                                          # a zero-weight entry must never be drawn
    w /= w.sum()
    pchb = host.lib()
    probs = np.zeros(n); bias = np.zeros(n); alias = np.zeros(n, dtype=np.int32)
    pchb.neci_host_alias_build(C.c_int32(n), w.ctypes.data_as(C.c_void_p), probs.ctypes.data_as(C.c_void_p),
                               bias.ctypes.data_as(C.c_void_p), alias.ctypes.data_as(C.c_void_p))
    assert np.allclose(probs, w, rtol=1e-14)
    prev = None
    for ndraw in (1000, 100000, 2000000):
        hist = np.zeros(n, dtype=np.int64)
        lib.orc_probe_alias_hist(bias.ctypes.data_as(C.c_void_p), alias.ctypes.data_as(C.c_void_p), C.c_int32(n),
                                 C.c_uint64(5), C.c_int64(ndraw), hist.ctypes.data_as(C.c_void_p))
        l1 = np.abs(hist / ndraw - w).sum()
        assert hist[3] == 0
        if prev is not None:
            assert l1 < prev
        prev = l1
    assert prev < 1e-2                                  # the reference's threshold after its 8e6 draws


# ---- PCHB: sum(1/pgen) test (unit_tests/excitgen/pchb_excitgen_test_helper.F90:40-118) -------------------------------------
def test_pchb_generator_sum_inverse_pgen():
    """det_I = [1,2,3,7,8,10], 10 spatial orbitals, random FCIDUMP (sparse 0.7): for every connected determinant
    with a non-zero matrix element, sum(1/pgen)/n_iter within [0.85, 1.15]; get_pgen == returned pgen."""
    s = host.random_fcidump_system(10, 6, sparse=0.7, sparse_t=0.7, seed=25, p_singles=0.3)
    o = oracle_for(s)
    det = [1, 2, 3, 7, 8, 10]
    il = s.ilut(det).reshape(1, -1)
    n_iter = 1_500_000
    res = o.probe_gen_excit(np.repeat(il, n_iter, axis=0), np.arange(n_iter, dtype=np.int32), 1)
    valid = res["ilut_j"][:, 0] != 0
    assert 0.2 < valid.mean() < 1.0
    keys, inv, cnt = np.unique(res["ilut_j"][valid, 0], return_inverse=True, return_counts=True)
    contrib = np.bincount(inv, weights=1.0 / res["pgen"][valid]) / n_iter
    # matrix elements to all generated determinants
    h = o.probe_helement(np.repeat(il, keys.size, axis=0), keys.reshape(-1, 1))
    nz = np.abs(h) > 1e-10
    assert nz.sum() > 50
    # the reference draws 5e7 excitations; with 1.5e6 the [0.85, 1.15] window applies to determinants hit
    # often enough (relative std 1/sqrt(hits) <= 4 %), all others must be within 5 standard deviations
    often = nz & (cnt >= 600)
    assert often.sum() > 100
    assert np.all(np.abs(contrib[often] - 1.0) < 0.15), (contrib[often].min(), contrib[often].max())
    z = (contrib[nz] - 1.0) * np.sqrt(cnt[nz])
    assert np.abs(z).max() < 5.0 and 0.7 < z.std() < 1.3, (np.abs(z).max(), z.std())
    # every connected determinant with non-zero element must be generated
    all_conn = []
    occ = set(det)
    for d in helpers.all_dets(s):
        ex = len(occ - set(d))
        if ex in (1, 2):
            all_conn.append(d)
    ilc = np.array([s.ilut(d) for d in all_conn]).reshape(-1, 1)
    hc = o.probe_helement(np.repeat(il, len(all_conn), axis=0), ilc)
    must = set(int(x) for x in ilc[np.abs(hc) > 1e-10, 0])
    assert must.issubset(set(int(x) for x in keys))
    # pgen recomputation for doubles
    dbl_mask = valid & (res["ic"] == 2)
    pg = o.probe_pchb_pgen(res["ex"][dbl_mask][:20000])
    assert np.allclose(pg, res["pgen"][dbl_mask][:20000], rtol=1e-12)


@pytest.mark.parametrize("selection", ["FULL-FULL", "UNIF-FULL"])
def test_pchb_full_full_particle_selection_sum_inverse_pgen(selection):
    """The same acceptance test with PCHB_ParticleSelection FULL-FULL (PC_FullyWeightedParticles_t,
    src/gasci_pchb_doubles_select_particles.fpp:330-438), the selection the reference's own PCHB regression input uses,
    and UNIF-FULL (PC_WeightedParticles_t, :440-506: first particle uniform, second weighted); the reference runs the
    same harness on the same determinant for these particle selections in unit_tests/gasci/gasci_pchb_test_helper.F90:62-125:
    sum(1/pgen)/n_iter within [0.85, 1.15] for every connected determinant with a non-zero element, completeness, and
    get_pgen (which depends on the determinant here) == returned pgen.  Also the tables themselves: p_first and every
    row of p_second are normalised, p(I | I) = 0, the pair weights are symmetric."""
    s = host.random_fcidump_system(10, 6, sparse=0.7, sparse_t=0.7, seed=25, p_singles=0.3, particle_selection=selection)
    t = s.tables["pchb"]
    nb = s.nbasis
    p2 = t["p_second"].reshape(nb, nb)
    assert abs(t["p_first"].sum() - 1.0) < 1e-12 and np.all(t["p_first"] >= 0)
    assert np.allclose(p2.sum(axis=1)[t["p_first"] > 0], 1.0, atol=1e-12) and np.all(np.diag(p2) == 0.0)
    wij = p2 * t["p_first"][:, None]                    # p_first[I] p(J | I) is proportional to IJ_weights(J, I): symmetric
    assert np.allclose(wij, wij.T, rtol=1e-12, atol=1e-15)
    o = oracle_for(s)
    det = [1, 2, 3, 7, 8, 10]
    il = s.ilut(det).reshape(1, -1)
    n_iter = 1_500_000
    res = o.probe_gen_excit(np.repeat(il, n_iter, axis=0), np.arange(n_iter, dtype=np.int32), 1)
    valid = res["ilut_j"][:, 0] != 0
    assert 0.2 < valid.mean() < 1.0
    keys, inv, cnt = np.unique(res["ilut_j"][valid, 0], return_inverse=True, return_counts=True)
    contrib = np.bincount(inv, weights=1.0 / res["pgen"][valid]) / n_iter
    h = o.probe_helement(np.repeat(il, keys.size, axis=0), keys.reshape(-1, 1))
    nz = np.abs(h) > 1e-10
    often = nz & (cnt >= 600)
    assert often.sum() > 100
    assert np.all(np.abs(contrib[often] - 1.0) < 0.15), (contrib[often].min(), contrib[often].max())
    z = (contrib[nz] - 1.0) * np.sqrt(cnt[nz])
    assert np.abs(z).max() < 5.0 and 0.7 < z.std() < 1.3, (np.abs(z).max(), z.std())
    # completeness: every connected determinant with a non-zero element whose probability makes it due (30 expected
    # hits) was generated -- a double with |H_ij| = 2e-4 has p = 1e-7 under weights proportional to |H_ij|
    occ = set(det)
    all_conn = [d for d in helpers.all_dets(s) if len(occ - set(d)) in (1, 2)]
    ilc = np.array([s.ilut(d) for d in all_conn]).reshape(-1, 1)
    hc = o.probe_helement(np.repeat(il, len(all_conn), axis=0), ilc)
    got = set(int(x) for x in keys)
    n_due = 0
    for d, k, hd in zip(all_conn, ilc[:, 0], hc):
        if abs(hd) <= 1e-10:
            continue
        src, tgt = sorted(occ - set(d)), sorted(set(d) - occ)
        if len(src) == 2:
            pg = o.probe_pchb_pgen_det(il, np.array([src + tgt], dtype=np.int32))[0]
            if pg * n_iter < 30.0:
                continue
        n_due += 1
        assert int(k) in got, (src, tgt, hd)
    assert n_due > 150
    dbl_mask = valid & (res["ic"] == 2)
    m = 20000
    pg = o.probe_pchb_pgen_det(np.repeat(il, m, axis=0), res["ex"][dbl_mask][:m])
    assert np.allclose(pg, res["pgen"][dbl_mask][:m], rtol=1e-12)
    # the weighting does what it is for: the spread of |H_ij| / pgen over the doubles is narrower than with UNIF-UNIF
    su = host.random_fcidump_system(10, 6, sparse=0.7, sparse_t=0.7, seed=25, p_singles=0.3)
    ou = oracle_for(su)
    ru = ou.probe_gen_excit(np.repeat(il, 200000, axis=0), np.arange(200000, dtype=np.int32), 1)
    def spread(r):
        d = (r["ilut_j"][:, 0] != 0) & (r["ic"] == 2)
        x = np.abs(r["hel"][d]) / r["pgen"][d]
        return x.std() / x.mean()
    assert spread({k: v[:200000] for k, v in res.items()}) < spread(ru)


def test_hubbard_generators_sum_inverse_pgen():
    """The reference's stochastic generator tests for the lattice models
    (test_real_space_hubbard.F90:1599, test_k_space_hubbard.F90:3804) with the same harness criterion."""
    for s in (host.hubbard_rs_system(3, 2, nel=4, U=4.0), host.hubbard_k_system(3, 2, nel=4, U=4.0)):
        o = oracle_for(s)
        rng = np.random.default_rng(1)
        for det in helpers.random_dets(s, 3, rng):
            il = s.ilut(det).reshape(1, -1)
            n_iter = 300_000
            res = o.probe_gen_excit(np.repeat(il, n_iter, axis=0), np.arange(n_iter, dtype=np.int32), 2)
            valid = res["ilut_j"][:, 0] != 0
            if not valid.any():
                continue
            keys, inv = np.unique(res["ilut_j"][valid, 0], return_inverse=True)
            contrib = np.bincount(inv, weights=1.0 / res["pgen"][valid]) / n_iter
            h = o.probe_helement(np.repeat(il, keys.size, axis=0), keys.reshape(-1, 1))
            assert np.all(np.abs(h) > 1e-12)
            assert np.all(np.abs(contrib - 1.0) < 0.05), (s.kind, contrib.min(), contrib.max())
            # completeness
            occ = set(det)
            conn = [d for d in helpers.all_dets(s) if len(occ - set(d)) in (1, 2)]
            ilc = np.array([s.ilut(d) for d in conn]).reshape(-1, 1)
            hc = o.probe_helement(np.repeat(il, len(conn), axis=0), ilc)
            must = set(int(x) for x in ilc[np.abs(hc) > 1e-12, 0])
            assert must == set(int(x) for x in keys)


# ---- FCIQMC on the oracle against exact diagonalisation (small lattice) ------------------------------------------------------
def test_oracle_fciqmc_energy_matches_exact_diagonalisation():
    s = host.hubbard_k_system(2, 2, nel=4, U=1.0)
    dets = helpers.all_dets(s)
    hii = driver.diag_energy(s, s.ref_orbs)
    o, params = helpers.make_pair(s, hii, max_walkers=20000, max_spawned=20000, seed=3, initiator=False)
    e0 = np.linalg.eigvalsh(_h_oracle(o, s, dets))[0]
    run = driver.FciMC(s, o, hii, tau=0.01, init_walkers=3000, steps_sft=10, sft_damp=0.1)
    run.seed_reference(10)
    run.run(6000)
    hist = [h for h in run.history if h["varying"]][100:]
    num = np.array([h["enum_cyc"] for h in hist]); den = np.array([h["hf_cyc"] for h in hist])
    e, err = driver.ratio_estimate(num, den)
    assert abs(e + hii - e0) < max(5 * err, 2e-3), (e + hii, e0, err)
    sm, serr = driver.blocking([h["shift"] for h in hist])
    assert abs(sm + hii - e0) < max(5 * serr, 1e-2), (sm + hii, e0, serr)


def test_k_hubbard_generator_known_answers():
    """gen_excit_k_space_hub on the reference's 4-site chain from nI = [1,2,3,4]: exactly the six determinants of
    test_k_space_hubbard.F90:2781-2852 are reachable, with the probabilities (p_elec = 1/4, halved where two
    orbital pairs serve one electron pair) and parities asserted there; the pair [3,4] cannot reach [5,6]
    (calc_pgen_k_space_hubbard_test :2755-2779)."""
    g = GOLD["k_hubbard"]; lat = g["lattice"]; k = g["gen_excit"]
    s = k_chain_system(lat["k_of_spatial_orbital"], lat["bhub"], g["uhub"] / g["omega"], 4, lat["length"])
    o = oracle_for(s)
    il = s.ilut(k["nI"]).reshape(1, -1)
    n = 4000
    out = o.probe_gen_excit(np.repeat(il, n, axis=0), np.arange(n, dtype=np.int32), 3)
    seen = {}
    for j in range(n):
        if out["pgen"][j] <= 0:
            continue
        key = int(np.uint64(out["ilut_j"][j, 0]))
        seen.setdefault(key, set()).add((float(out["pgen"][j]), int(out["parity"][j]), tuple(int(x) for x in out["ex"][j])))
    want = {int(np.uint64(s.ilut(c["nJ"])[0])): c for c in k["reachable"]}
    assert set(seen) == set(want)
    for key, vals in seen.items():
        for pg, par, ex in vals:
            assert pg == want[key]["pgen"] and bool(par) == want[key]["tpar"]
    src_34_to_56 = int(np.uint64(s.ilut([1, 2, 5, 6])[0]))
    for pg, par, ex in seen[int(np.uint64(s.ilut([3, 4, 5, 6])[0]))]:
        assert ex[:2] == (1, 2)                       # [3,4,5,6] is reached by exciting the pair (1,2) only
    assert src_34_to_56 not in seen                   # pgen([3,4] -> [5,6]) = 0


def test_rs_hubbard_generator_known_answers():
    """gen_excit_rs_hubbard on the reference's periodic 4-site chain: calc_pgen_rs_hubbard_test
    (test_real_space_hubbard.F90:1808-1860) pins pgen(1->3) = pgen(2->8) = 1/4 from nI = [1,2]; an occupied neighbour
    adds nothing to the cumulative list (create_cum_list_rs_hubbard_test :1673-1700), so from nI = [1,3] the beta
    electron on site 1 reaches only site 4, with pgen 1/2."""
    g = GOLD["rs_hubbard_gen_excit"]
    s = _lattice(g["lattice"], 0.0, 2)
    o = oracle_for(s)
    n = 2000
    il = s.ilut(g["nI"]).reshape(1, -1)
    out = o.probe_gen_excit(np.repeat(il, n, axis=0), np.arange(n, dtype=np.int32), 5)
    want = {int(np.uint64(s.ilut(c["nJ"])[0])): c["pgen"] for c in g["reachable"]}
    got = {}
    for j in range(n):
        assert out["pgen"][j] > 0 and out["ic"][j] == 1
        got.setdefault(int(np.uint64(out["ilut_j"][j, 0])), set()).add(float(out["pgen"][j]))
    assert set(got) == set(want)
    for k, v in got.items():
        assert v == {want[k]}
    b = g["blocked"]
    il = s.ilut(b["nI"]).reshape(1, -1)
    out = o.probe_gen_excit(np.repeat(il, n, axis=0), np.arange(n, dtype=np.int32), 5)
    from_1 = out["ex"][:, 0] == 1
    assert from_1.sum() > 100
    assert set(int(x) for x in out["ex"][from_1, 2]) == set(b["reachable_from_orb_1"])
    assert set(float(x) for x in out["pgen"][from_1]) == {b["pgen"]}


def test_det_node_matches_the_reference_processor_of_its_own_runs():
    """DetermineDetNode / get_det_block (src/load_balance_calcnodes.F90:25-117) + the initial LoadBalanceMapping
    (src/load_balancer.fpp:96-99) against `Reference processor is: N` as printed by the reference's regression runs
    (4 MPI ranks; 4 blocks, or 400 with load-balance-blocks).  RandomOrbIndex of every case was rebuilt with the
    reference's own dSFMT (tests/golden/make_det_node_fixture.py)."""
    cases = json.load(open(os.path.join(helpers.GOLDEN, "reference_det_nodes.json")))
    assert len(cases) >= 40
    assert any(c["balance_blocks"] > c["nprocs"] for c in cases)
    for c in cases:
        nb, nel = c["nbasis"], len(c["det"])
        s = host.System(kind=capi.SYS_HUBBARD_RS, nel=nel, nbasis=nb, nocc_alpha=nel // 2, nocc_beta=nel - nel // 2,
                        t_exch=0, t_no_brillouin=1,
                        tables=dict(max_neigh=1, neighbours=np.zeros(nb, dtype=np.int32), tmat=np.zeros(nb * nb), uhub=0.0),
                        ref_orbs=np.array(c["det"], dtype=np.int32))
        o, params = helpers.make_pair(s, 0.0, max_walkers=100, max_spawned=100 * c["nprocs"], nranks=c["nprocs"], rank=0,
                                      blocks_per_rank=c["balance_blocks"] // c["nprocs"],
                                      random_orb_index=c["random_orb_index"])
        blk, node = o.probe_det_node(s.ilut(c["det"]).reshape(1, -1))
        assert int(blk[0]) == c["block"] and int(node[0]) == c["reference_processor"], (c["case"], blk, node)
        o.close()


@pytest.mark.parametrize("which,ref_orbs", [("total_momentum_6", list(range(1, 13))),
                                            ("total_momentum_0", list(range(1, 11)) + [11, 14])])
def test_k_hubbard_doubles_core_matches_the_reference_runs(which, ref_orbs):
    """Two reference runs on the 12-site k-space Hubbard chain (U = 1, half filling) with `doubles-core`: the
    deterministic space has 205 determinants (reference + every spin- and momentum-conserving double excitation,
    same-spin ones included) and the lowest eigenvalue of H over it, relative to the reference energy, is the printed
    `Deterministic subspace correlation energy` (-0.2580331648 with both open-shell electrons in one eps = 0 orbital,
    -0.2274848837 with one in each); the reference energy -11.9282032303 as printed.  Oracle's
    get_diag_helement_k_sp_hub / get_offdiag_helement_k_sp_hub on a lattice the unit tests do not reach."""
    import itertools
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hubbard_doubles_core.json")))[which]
    assert g["cell"] == [12, 1, 1] and g["b"] == -1.0
    s = host.hubbard_k_system(g["cell"][0], g["cell"][1], nel=g["electrons"], U=g["u"])
    t = s.tables
    nk = t["n_k"]
    ksum = np.array(t["ksum"]).reshape(nk, nk)
    kidx = lambda o: (o + 1) // 2 - 1
    hii = driver.diag_energy(s, ref_orbs)
    assert abs(hii - g["reference_energy"]) < 6e-11
    vir = [o for o in range(1, s.nbasis + 1) if o not in ref_orbs]
    dets = [ref_orbs]
    for i, j in itertools.combinations(ref_orbs, 2):
        for a, b in itertools.combinations(vir, 2):
            if (i & 1) + (j & 1) == (a & 1) + (b & 1) and ksum[kidx(i), kidx(j)] == ksum[kidx(a), kidx(b)]:
                dets.append(sorted([o for o in ref_orbs if o not in (i, j)] + [a, b]))
    n = len(dets)
    assert n == g["core_size"] == 205
    # the host library's enumerate_sing_doub_kpnt gives the same set
    sd = host.sing_doub_space(s, ref_ilut=s.ilut(ref_orbs))
    assert sd.shape[0] == n and {tuple(r) for r in sd.tolist()} == {tuple(s.ilut(d).tolist()) for d in dets}
    o, _ = helpers.make_pair(s, hii, max_walkers=1000, max_spawned=1000)
    il = np.array([s.ilut(d) for d in dets], dtype=np.int64).reshape(n, s.nw)
    I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
    H = o.probe_helement(il[I], il[J]).reshape(n, n) - hii * np.eye(n)
    assert np.allclose(H, H.T, atol=1e-13)
    assert abs(np.linalg.eigvalsh(H)[0] - g["core_correlation_energy"]) < 6e-11
    # the host library's sparse rows for the same space (calc_determ_hamil_sparse of the stand-alone host)
    c = host.core_hamiltonian(s, il, hii)
    Hh = np.zeros((n, n))
    for i in range(n):
        sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
        Hh[i, c["col"][sl]] = c["val"][sl]
    assert np.allclose(Hh, H, rtol=1e-12, atol=1e-13)


def test_particle_selection_tables_equal_the_brute_force_pair_weights():
    """p_first / p_second of the weighted PCHB particle selections against an independent evaluation: the weight of a
    spin-orbital pair (I, J) is the sum of |<IJ||AB>| over all hole pairs (what the reference accumulates sampler by
    sampler, src/gasci_pchb_doubles_spatorb_fastweighted.fpp:374-420), p(J | I) its column normalised, p(I) the
    normalised column sums (init_PC_WeightedParticles_t, src/gasci_pchb_doubles_select_particles.fpp:268-328)."""
    n_spat = 5
    s = host.random_fcidump_system(n_spat, 4, sparse=0.8, sparse_t=0.8, seed=12, particle_selection="FULL-FULL")
    umat = s.tables["umat"]
    nb = 2 * n_spat

    def tri(a, b):
        return a * (a - 1) // 2 + b if a > b else b * (b - 1) // 2 + a

    def um(i, j, k, l):                                   # <ij|kl> over spatial orbitals, UMatInd
        return umat[tri(tri(i, k), tri(j, l)) - 1]

    def sp(o):
        return (o + 1) // 2

    def beta(o):
        return o % 2 == 1

    W = np.zeros((nb + 1, nb + 1))
    for I in range(1, nb + 1):
        for J in range(I + 1, nb + 1):
            tot = 0.0
            for A in range(1, nb + 1):
                for B in range(A + 1, nb + 1):
                    if A in (I, J) or B in (I, J):
                        continue
                    h = 0.0                                # sltcnd_2: <IJ|AB> - <IJ|BA> with spin deltas
                    if beta(I) == beta(A) and beta(J) == beta(B):
                        h += um(sp(I), sp(J), sp(A), sp(B))
                    if beta(I) == beta(B) and beta(J) == beta(A):
                        h -= um(sp(I), sp(J), sp(B), sp(A))
                    tot += abs(h)
            W[I, J] = W[J, I] = tot
    W = W[1:, 1:]
    col = W.sum(axis=0)
    t = s.tables["pchb"]
    assert np.allclose(t["p_first"], col / col.sum(), rtol=1e-12, atol=1e-15)
    p2 = t["p_second"].reshape(nb, nb)
    for I in range(nb):
        assert np.allclose(p2[I], W[:, I] / col[I], rtol=1e-12, atol=1e-15)
