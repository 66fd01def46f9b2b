"""The oracle's annihilation (hash merge in C++, oracle/orc_engine.cpp) against a second restatement of the same
reference routines in their own sort-and-merge form (tests/literal_annihilation.py, written from
src/Annihilation.F90:249-634, 965-1479 and src/load_balancer.fpp:514-805): fixed spawned lists with duplicates,
opposite signs, exact cancellations, initiator and non-initiator parents, occupied and unoccupied targets, integer
and real amplitudes (including amplitudes below OccupiedThresh, which exercise stochRoundSpawn and the pruning pass
with the same counter-based random numbers).  Integer lists must agree exactly, real ones to 1e-12."""
import ctypes as C

import numpy as np
import pytest

import helpers
import literal_annihilation as lit
from neci_stable_b200 import capi, host, driver
from neci_stable_b200.capi import ST

SEED = 11


def _draw(system, it, purpose):
    lib = helpers.oracle_lib()

    def f(det):
        il = np.array(det, dtype=np.uint64).view(np.int64)
        out = np.zeros(1)
        lib.orc_probe_stream(C.c_uint64(SEED), C.c_int64(it), il.ctypes.data_as(C.c_void_p), C.c_int32(system.nw),
                             C.c_int32(0), C.c_int32(purpose), C.c_int32(1), out.ctypes.data_as(C.c_void_p))
        return float(out[0])
    return f


def make_case(system, rng, real, n_dets=1500, n_spawn=6000):
    dets = helpers.random_dets(system, n_dets, rng)
    nd = len(dets)
    main = dets[:(2 * nd) // 3]

    def amp(lo_ok):
        if not real:
            return float(rng.integers(1, 6) * rng.choice([-1, 1]))
        x = float(rng.choice([0.3, 0.95, 1.0, 1.7, 2.5]) * rng.random() * 2) if lo_ok else float(1.0 + 3 * rng.random())
        return x * float(rng.choice([-1, 1]))
    recs = np.array([host.record(system, d, amp(False), (1 << capi.FLAG_INITIATOR) if rng.random() < 0.3 else 0)
                     for d in main])
    sp = []
    for _ in range(n_spawn):
        d = dets[int(rng.integers(0, nd))]
        s = float(rng.choice([-2, -1, 1, 1, 2])) if not real else amp(True)
        sp.append(host.record(system, d, s, (1 << capi.FLAG_INITIATOR) if rng.random() < 0.5 else 0))
    # exact cancellations: with a main-list determinant, and inside the spawned list (a block summing to zero)
    sp.append(host.record(system, main[0], -host.signs_of(recs[:1], system.nw)[0], 0))
    sp.append(host.record(system, dets[-1], 1.5, 1 << capi.FLAG_INITIATOR))
    sp.append(host.record(system, dets[-1], -1.5, 0))
    # a single-entry block without amplitude
    sp.append(host.record(system, dets[-2], 0.0, 1 << capi.FLAG_INITIATOR))
    sp = [r for r in sp if not (tuple(r[:system.nw]) == tuple(host.record(system, dets[-2], 0.0)[:system.nw]) and
                                host.signs_of(np.array([r]), system.nw)[0] != 0.0)]
    return recs, np.array(sp)


def literal_run(system, recs, sp, it, initiator, thresh=1.0):
    nw = system.nw
    key = lambda r: tuple(int(x) for x in np.asarray(r[:nw]).view(np.uint64))
    main = {key(r): [float(host.signs_of(np.array([r]), nw)[0]), int(r[nw + 1])] for r in recs}
    spawned = [(key(r), float(host.signs_of(np.array([r]), nw)[0]), int(r[nw + 1])) for r in sp]
    comp, ann1 = lit.compress_spawned_list(spawned, nw, t_trunc_initiator=initiator)
    st = lit.annihilate_spawned_parts(main, comp, _draw(system, it, 3), t_trunc_initiator=initiator, occupied_thresh=thresh)
    st2 = lit.calc_hash_table_stats(main, _draw(system, it, 4), occupied_thresh=thresh)
    return main, dict(merged=len(comp), annihilated=ann1 + st["Annihilated"], aborted=st["NoAborted"],
                      removed=st["NoRemoved"] + st2["NoRemoved"], born=st["NoBorn"] + st2["NoBorn"],
                      inserted=st["inserted"], totparts=st2["TotParts"], norm=st2["norm_psi_squared"],
                      highest=st2["iHighestPop"])


def compare(system, engine_stats, engine_list, main, want, exact):
    eq = (lambda a, b: a == b) if exact else (lambda a, b: np.isclose(a, b, rtol=1e-12, atol=1e-12))
    so = engine_stats
    assert so[ST["NSPAWNED_MERGED"]] == want["merged"]
    assert so[ST["NINSERTED"]] == want["inserted"]
    for name, k in (("ANNIHILATED", "annihilated"), ("NOABORTED", "aborted"), ("NOREMOVED", "removed"),
                    ("NOBORN", "born"), ("TOTPARTS", "totparts"), ("NORM_PSI_SQ", "norm")):
        assert eq(so[ST[name]], want[k]), (name, so[ST[name]], want[k])
    assert so[ST["HIGHEST_POP"]] == want["highest"]
    c = helpers.canon(*engine_list, nw=system.nw)
    got = {tuple(int(x) for x in row.view(np.uint64)): (float(s), int(f)) for row, s, f in zip(c[0], c[1], c[2])}
    assert set(got) == set(main)
    for k, (s, f) in got.items():
        assert eq(s, main[k][0]), (k, s, main[k][0])
        assert (f >> capi.FLAG_INITIATOR) & 1 == (main[k][1] >> capi.FLAG_INITIATOR) & 1, (k, f, main[k][1])


@pytest.mark.parametrize("name,real,initiator", [("1w", False, True), ("1w", False, False), ("1w", True, True),
                                                 ("2w", False, True), ("2w", True, True), ("2w", True, False)])
def test_oracle_annihilation_equals_literal_sort_and_merge(name, real, initiator):
    system = host.random_fcidump_system(6, 6, sparse=0.9, sparse_t=0.9, seed=3) if name == "1w" else \
        host.hubbard_k_system(6, 6, U=4.0)
    hii = driver.diag_energy(system, system.ref_orbs)
    o, _ = helpers.make_pair(system, hii, max_walkers=50000, max_spawned=50000, seed=SEED, initiator=initiator,
                             all_real_coeff=real)
    rng = np.random.default_rng(17 + 2 * real + initiator)
    # few enough spawns per determinant that single, non-initiator spawns (aborted) occur beside merged blocks
    recs, sp = make_case(system, rng, real, n_dets=400 if name == "1w" else 1500, n_spawn=500 if name == "1w" else 3000)
    o.upload_walkers(recs)
    so = o.annihilate(sp, 9)
    main, want = literal_run(system, recs, sp, 9, initiator)
    assert want["merged"] < sp.shape[0] and want["inserted"] > 0 and want["annihilated"] > 0
    if initiator:
        assert want["aborted"] > 0
    if real:
        assert want["removed"] > 0 and want["born"] > 0          # stochRoundSpawn / pruning took both branches
    compare(system, so, o.download_walkers(), main, want, exact=not real)
