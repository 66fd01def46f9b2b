"""FCIDUMP reader / writer / frozen-core folding (neci_stable_b200/fcidump.py; src/read_fci.F90, src/readint.F90,
IntFreeze in src/Integrals_neci.F90) on CPU: a file written from the integrals of the reference's HeHe_SS_Doubles
case reads back to the same system and to the `Reference Energy` the reference printed for it; freezing a core
orbital leaves every matrix element between determinants with that orbital doubly occupied unchanged."""
import itertools
import json
import os

import numpy as np

from neci_stable_b200 import driver, fcidump, host

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_write_read_round_trip_and_reference_energy(tmp_path):
    g = json.load(open(os.path.join(GOLD, "hehe_ss_doubles.json")))
    d = fcidump.FciDump(norb=g["norb"], nelec=g["nelec"], ms2=g["ms2"], orbsym=g["orbsym"], ecore=g["ecore"], eps=g["eps"],
                        h1=[tuple(x) for x in g["h1"]], eri=[tuple(x) for x in g["eri"]])
    path = str(tmp_path / "FCIDUMP")
    fcidump.write_fcidump(path, d)
    r = fcidump.read_fcidump(path)
    assert (r.norb, r.nelec, r.ms2, r.orbsym) == (d.norb, d.nelec, d.ms2, d.orbsym)
    assert r.ecore == d.ecore and r.eps == d.eps
    assert sorted(r.h1) == sorted((int(i), int(j), float(v)) for i, j, v in d.h1)
    assert sorted(r.eri) == sorted((int(i), int(j), int(k), int(l), float(v)) for i, j, k, l, v in d.eri)
    s = r.system()
    for name in ("umat", "tmat"):
        assert np.array_equal(s.tables[name], d.system().tables[name])
    e_ref = driver.diag_energy(s, g["reference_det"])
    assert abs(e_ref - g["reference_energy"]) < 5e-12          # the reference prints 12 decimals
    # the same through the host library's get_helement
    il = s.ilut(g["reference_det"])
    assert abs(host.get_helement(s, il, il)[0] - g["reference_energy"]) < 5e-12


def test_reader_accepts_fortran_exponents_and_slash_terminator(tmp_path):
    p = tmp_path / "FCIDUMP"
    p.write_text(" &FCI NORB=  2,NELEC= 2,MS2= 0,\n  ORBSYM=1,1,\n  ISYM=1\n /\n"
                 "  0.5D+00   1   1   1   1\n  0.25d0   2   2   1   1\n -1.0E+00   1   1   0   0\n"
                 " -0.5   2   2   0   0\n  0.1   2   1   0   0\n  2.0   0   0   0   0\n")
    d = fcidump.read_fcidump(str(p))
    assert (d.norb, d.nelec, d.ms2, d.orbsym, d.ecore) == (2, 2, 0, [1, 1], 2.0)
    assert d.eri == [(1, 1, 1, 1, 0.5), (2, 2, 1, 1, 0.25)] and d.h1 == [(1, 1, -1.0), (2, 2, -0.5), (2, 1, 0.1)]
    assert d.eps is None


def _random_dump(norb, nelec, rng):
    d = fcidump.FciDump(norb=norb, nelec=nelec, ms2=0, orbsym=[1] * norb, ecore=0.37)
    for i in range(1, norb + 1):
        for j in range(1, i + 1):
            d.h1.append((i, j, float(rng.normal()) + (3.0 * i if i == j else 0.0)))
    pairs = [(i, j) for i in range(1, norb + 1) for j in range(1, i + 1)]
    for a, (i, j) in enumerate(pairs):
        for (k, l) in pairs[:a + 1]:
            d.eri.append((i, j, k, l, float(rng.normal()) * 0.3))
    return d


def test_frozen_core_leaves_the_hamiltonian_of_the_valence_space_unchanged():
    rng = np.random.default_rng(2)
    full = _random_dump(6, 6, rng)
    core = [2, 5]                                          # not the first orbitals: exercises the renumbering
    froz = fcidump.freeze_core(full, core)
    assert (froz.norb, froz.nelec) == (4, 2)
    sf = full.system(ref_spatial=[1, 2, 3])
    sz = froz.system(ref_spatial=[1])
    keep = [p for p in range(1, 7) if p not in core]
    dets_z, dets_f = [], []
    for a in itertools.combinations(range(1, 5), 1):
        for b in itertools.combinations(range(1, 5), 1):
            dz = sorted([2 * i for i in a] + [2 * i - 1 for i in b])
            df = sorted([2 * keep[i - 1] for i in a] + [2 * keep[i - 1] - 1 for i in b] +
                        [2 * c for c in core] + [2 * c - 1 for c in core])
            dets_z.append(sz.ilut(dz)); dets_f.append(sf.ilut(df))
    n = len(dets_z)
    iz, if_ = np.array(dets_z), np.array(dets_f)
    I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
    hz = host.get_helement(sz, iz[I], iz[J])
    hf = host.get_helement(sf, if_[I], if_[J])
    assert np.count_nonzero(hz) > n
    # the frozen electrons sit below / between the valence orbitals: moving a valence electron past an even number
    # of them never changes a parity
    assert np.allclose(hz, hf, rtol=1e-12, atol=1e-12)


def test_frozen_core_open_shell_case_matches_the_reference_run():
    """The reference's BeH_open_shell_explicit regression run (19 orbitals, 5 electrons, Ms = -1/2, C2v labels,
    `freeze 2 0`, `semi-stochastic doubles-core`): after fcidump.freeze_core the reference determinant (1, 2, 3) has
    the printed `Reference Energy` -15.1493554282, the symmetry-adapted singles and doubles are the printed 234
    determinants and the lowest eigenvalue of the host library's sparse Hamiltonian over them is the printed
    `Deterministic subspace correlation energy` -0.0383128036 -- frozen-core folding, open-shell Slater-Condon rules
    and the doubles-core generator on a system with unequal alpha and beta occupations."""
    import helpers
    z = np.load(os.path.join(GOLD, "beh_open_shell.npz"))
    d = fcidump.FciDump(norb=int(z["norb"]), nelec=int(z["nelec"]), ms2=int(z["input_spin_restrict"]),
                        orbsym=[int(x) for x in z["orbsym"]], ecore=float(z["ecore"]), eps=[float(x) for x in z["eps"]],
                        h1=[(int(a), int(b), float(v)) for a, b, v in z["h1"]],
                        eri=[(int(a), int(b), int(c), int(e), float(v)) for a, b, c, e, v in z["eri"]])
    assert list(z["input_freeze"]) == [2, 0]
    f = fcidump.freeze_core(d, [int(np.argmin(d.eps)) + 1])          # the two lowest spin orbitals
    s = f.system()
    assert (f.norb, f.nelec, s.nocc_alpha, s.nocc_beta) == (18, 3, 1, 2)
    assert [int(x) for x in s.ref_orbs] == [int(x) for x in z["reference_det"]]
    hii = driver.diag_energy(s, s.ref_orbs)
    assert abs(hii - float(z["reference_energy"])) < 6e-11
    sd = host.sing_doub_space(s, orbsym=f.orbsym)
    assert sd.shape[0] == int(z["core_size"]) == 234
    il, sizes, displs = host.layout_core_space(sd, np.zeros(sd.shape[0], dtype=np.int32), 1)
    c = host.core_hamiltonian(s, il, hii)
    n = il.shape[0]
    H = np.zeros((n, n))
    for i in range(n):
        sl = slice(c["row_ptr"][i], c["row_ptr"][i + 1])
        H[i, c["col"][sl]] = c["val"][sl]
    assert abs(np.linalg.eigvalsh(H)[0] - float(z["core_correlation_energy"])) < 6e-11
    # and from the oracle's elements
    o, _ = helpers.make_pair(s, hii, max_walkers=1000, max_spawned=1000)
    I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
    Ho = o.probe_helement(il[I], il[J]).reshape(n, n) - hii * np.eye(n)
    assert np.allclose(Ho, H, rtol=1e-12, atol=1e-13)
