"""Test helpers: bind the CPU oracle (oracle/_build/liboracle.so) behind the same
Python wrapper the product uses for the CUDA engine, canonicalise walker lists."""
import ctypes as C
import itertools
import os

import numpy as np

from neci_stable_b200 import capi, host, driver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.environ.get("ORACLE_LIB_OVERRIDE") or os.path.join(ROOT, "oracle", "_build", "liboracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


class Oracle(capi.Engine):
    """The oracle exports orc_* with the shapes of neci_gpu_* (+ phase-level entry points)."""

    def __init__(self, params):
        super().__init__(params, lib_path=ORACLE_LIB, prefix="orc_")

    def spawn_phase(self, tau, sft, it):
        self._check(self._fn("spawn_phase")(self.h, C.c_double(tau), C.c_double(sft), C.c_int64(it)), "spawn_phase")

    def spawned(self, dest):
        f = getattr(self.lib, "orc_spawned_count"); f.restype = C.c_int64
        n = f(self.h, C.c_int32(dest))
        out = np.zeros((max(n, 1), self.W), dtype=np.int64)
        self._fn("spawned_get")(self.h, C.c_int32(dest), out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out[:n]

    def annihilate_phase(self, spawned, it):
        sp = np.ascontiguousarray(spawned, dtype=np.int64).reshape(-1, self.W)
        st = np.zeros(capi.ST_COUNT)
        self._check(self._fn("annihilate_phase")(self.h, sp.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(sp.shape[0]),
                                                  C.c_int64(it), st.ctypes.data_as(C.POINTER(C.c_double))), "annihilate_phase")
        return st

    def partial_vec(self):
        f = getattr(self.lib, "orc_core_local"); f.restype = C.c_int64
        n = f(self.h)
        out = np.zeros(max(n, 1))
        self._fn("partial_vec_get")(self.h, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out[:n]

    def determ_projection(self, full, tau, sft):
        full = np.ascontiguousarray(full, dtype=np.float64)
        self._check(self._fn("determ_projection")(self.h, full.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(tau),
                                                   C.c_double(sft)), "determ_projection")

    def probe_walker_hash(self, iluts, table_len):
        il = np.ascontiguousarray(iluts, dtype=np.int64).reshape(-1, self.nw)
        out = np.zeros(il.shape[0], dtype=np.int32)
        self._fn("probe_walker_hash")(self.h, C.c_int64(il.shape[0]), il.ctypes.data_as(C.POINTER(C.c_int64)),
                                      C.c_int32(table_len), out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def probe_pchb_pgen(self, ex):
        ex = np.ascontiguousarray(ex, dtype=np.int32).reshape(-1, 4)
        out = np.zeros(ex.shape[0])
        self._fn("probe_pchb_pgen")(self.h, C.c_int64(ex.shape[0]), ex.ctypes.data_as(C.POINTER(C.c_int32)),
                                    out.ctypes.data_as(C.POINTER(C.c_double)))
        return out


def _probe_pchb_pgen_det(self, iluts, ex):
    """get_pgen of a PCHB double with a particle selector that depends on the determinant (FULL-FULL)."""
    ex = np.ascontiguousarray(ex, dtype=np.int32).reshape(-1, 4)
    il = np.ascontiguousarray(iluts, dtype=np.int64).reshape(ex.shape[0], -1)
    out = np.zeros(ex.shape[0])
    self._fn("probe_pchb_pgen_det")(self.h, C.c_int64(ex.shape[0]), il.ctypes.data_as(C.POINTER(C.c_int64)),
                                    ex.ctypes.data_as(C.POINTER(C.c_int32)), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


Oracle.probe_pchb_pgen_det = _probe_pchb_pgen_det


def oracle_lib():
    return C.CDLL(ORACLE_LIB)


def world_iterate(oracles, tau, sft, it, nthreads=1):
    """orc_world_iterate over several oracle ranks (threads play the MPI ranks)."""
    lib = oracles[0].lib
    n = len(oracles)
    arr = (C.c_void_p * n)(*[o.h for o in oracles])
    st = np.zeros((n, capi.ST_COUNT))
    rc = lib.orc_world_iterate(arr, C.c_int32(n), C.c_double(tau), C.c_double(sft), C.c_int64(it),
                               st.ctypes.data_as(C.POINTER(C.c_double)), C.c_int32(nthreads))
    assert rc == 0
    return st


def world_rebalance(oracles, new_mapping):
    lib = oracles[0].lib
    n = len(oracles)
    arr = (C.c_void_p * n)(*[o.h for o in oracles])
    m = np.ascontiguousarray(new_mapping, dtype=np.int32)
    rc = lib.orc_world_rebalance(arr, C.c_int32(n), m.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    for o in oracles:
        o.params["load_balance_mapping"] = m.copy()


def make_pair(system, hii, cls_gpu=True, **kw):
    """(oracle, params) for a system; the caller creates the CUDA engine from the same params."""
    params = host.make_params(system, hii, **kw)
    o = Oracle(params)
    system.apply(o)
    return o, params


def canon(dets, gd=None, go=None, nw=1):
    """Order-normalise a walker list: drop empty slots, sort by occupation words.
    Returns (orbital words, signs, flags without the `removed` bit[, gd, go])."""
    dets = np.asarray(dets).reshape(-1, nw + 2)
    sg = np.ascontiguousarray(dets[:, nw]).view(np.float64)
    keep = np.abs(sg) >= 1e-12
    keep |= ((dets[:, nw + 1] >> capi.FLAG_DETERMINISTIC) & 1).astype(bool)
    d = dets[keep]
    order = np.lexsort(tuple(d[:, w].view(np.uint64) for w in reversed(range(nw))))
    d = d[order]
    out = [d[:, :nw].copy(), np.ascontiguousarray(d[:, nw]).view(np.float64).copy(), d[:, nw + 1] & ~1]
    if gd is not None:
        out += [np.asarray(gd)[keep][order], np.asarray(go)[keep][order]]
    return out


def all_dets(system, sector=True):
    """Enumerate all determinants with (nalpha, nbeta) as sorted orbital lists."""
    ns = system.nbasis // 2
    out = []
    for a in itertools.combinations(range(1, ns + 1), system.nocc_alpha):
        for b in itertools.combinations(range(1, ns + 1), system.nocc_beta):
            out.append(sorted([2 * i for i in a] + [2 * i - 1 for i in b]))
    return out


def random_dets(system, n, rng):
    """n distinct random determinants (all of them, shuffled, if the sector holds fewer than 2n)."""
    import math
    ns = system.nbasis // 2
    total = math.comb(ns, system.nocc_alpha) * math.comb(ns, system.nocc_beta)
    if total <= 2 * n:
        dets = all_dets(system)
        order = rng.permutation(len(dets))[:n]
        return [dets[i] for i in sorted(order)]
    out = set()
    while len(out) < n:
        a = rng.choice(ns, system.nocc_alpha, replace=False) + 1
        b = rng.choice(ns, system.nocc_beta, replace=False) + 1
        out.add(tuple(sorted([2 * int(i) for i in a] + [2 * int(i) - 1 for i in b])))
    return [list(t) for t in sorted(out)]


def hamiltonian_matrix(engine, system, dets):
    """Dense H over a determinant list through probe_helement."""
    il = np.array([system.ilut(d) for d in dets], dtype=np.int64).reshape(len(dets), system.nw)
    n = len(dets)
    I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
    h = engine.probe_helement(il[I], il[J])
    return h.reshape(n, n)


def build_core_space(oracle, system, core_dets, hii, nranks=1):
    """Host-side mirror of what init_semi_stochastic hands over (src/semi_stoch_gen.F90:103-350,
    src/fast_determ_hamil.F90:1421-1507): the core determinants grouped by owner rank (core-space order = rank-major),
    and per rank the CSR rows of H over the whole core space with Hii subtracted on the diagonal.
    Returns (ordered_dets, sizes, displs, [dict(row_ptr, col, val, iluts) per rank])."""
    il = np.array([system.ilut(d) for d in core_dets], dtype=np.int64).reshape(len(core_dets), system.nw)
    _, node = oracle.probe_det_node(il)
    if nranks == 1:
        node = np.zeros(len(core_dets), dtype=np.int32)
    order = np.argsort(node, kind="stable")
    dets = [core_dets[i] for i in order]
    il = il[order]; node = node[order]
    sizes = np.bincount(node, minlength=nranks).astype(np.int32)
    displs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    n = len(dets)
    I = np.repeat(np.arange(n), n); J = np.tile(np.arange(n), n)
    H = oracle.probe_helement(il[I], il[J]).reshape(n, n)
    H[np.arange(n), np.arange(n)] -= hii
    per_rank = []
    for r in range(nranks):
        rows = range(displs[r], displs[r] + sizes[r])
        row_ptr = [0]; col = []; val = []
        for i in rows:
            nz = np.nonzero(np.abs(H[i]) > 0)[0]
            nz = nz[nz != i]
            # off-diagonal elements first, the diagonal appended last (fast_determ_hamil.F90:1496-1507)
            col += list(nz) + [i]; val += list(H[i, nz]) + [H[i, i]]
            row_ptr.append(len(col))
        per_rank.append(dict(row_ptr=np.array(row_ptr, dtype=np.int64), col=np.array(col, dtype=np.int32),
                             val=np.array(val, dtype=np.float64), iluts=il.copy()))   # the whole core space, replicated
    return dets, sizes, displs, per_rank, H


class DistOracle:
    """One oracle rank inside a torch.distributed job (gloo on CPU): the spawn exchange (SendProcNewParts,
    src/Annihilation.F90:150-247) and the core-vector gather (src/semi_stoch_procs.F90:127) go through the process
    group.  Same iterate() interface as capi.Engine, so driver.FciMC drives it unchanged."""

    def __init__(self, oracle, dist):
        self.o, self.dist = oracle, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def upload_walkers(self, *a, **k):
        return self.o.upload_walkers(*a, **k)

    def download_walkers(self, *a, **k):
        return self.o.download_walkers(*a, **k)

    def iterate(self, tau, sft, it):
        o = self.o
        o.spawn_phase(tau, sft, it)
        if o.params["t_semi_stochastic"]:
            parts = [None] * self.world
            self.dist.all_gather_object(parts, o.partial_vec())
            o.determ_projection(np.concatenate(parts), tau, sft)
        out = [o.spawned(r) for r in range(self.world)]
        everything = [None] * self.world
        self.dist.all_gather_object(everything, out)
        recv = np.concatenate([everything[s][self.rank] for s in range(self.world)])   # ordered by source rank
        return o.annihilate_phase(recv, it)


def build_trial_space(engine, system, dets, trial_dets):
    """Host-side mirror of init_trial_wf (src/trial_wf_gen.F90): diagonalise H in the trial space, take the lowest
    eigenvector as the trial vector, and form the connected space = determinants of `dets` outside the trial space
    with con_space_vecs_i = sum_j H_ij psiT_j != 0.  Returns (trial_iluts, trial_amps, con_iluts, con_amps,
    trial_energy) with H the full Hamiltonian (ECore included)."""
    tset = {tuple(d) for d in trial_dets}
    rest = [d for d in dets if tuple(d) not in tset]
    Ht = hamiltonian_matrix(engine, system, trial_dets)
    w, v = np.linalg.eigh(Ht)
    psi = v[:, 0]
    il_t = np.array([system.ilut(d) for d in trial_dets], dtype=np.int64).reshape(len(trial_dets), system.nw)
    il_r = np.array([system.ilut(d) for d in rest], dtype=np.int64).reshape(len(rest), system.nw)
    nr, nt = len(rest), len(trial_dets)
    I = np.repeat(np.arange(nr), nt); J = np.tile(np.arange(nt), nr)
    Hct = engine.probe_helement(il_r[I], il_t[J]).reshape(nr, nt)
    con = Hct @ psi
    keep = np.abs(con) > 0
    return il_t, psi.copy(), il_r[keep], con[keep], float(w[0])


def hphf_allowed(system, dets):
    """The determinants IsAllowedHPHF accepts (closed shell, or the larger of a spin-flipped pair; one word)."""
    A, B = 0xAAAAAAAAAAAAAAAA, 0x5555555555555555
    out = []
    for d in dets:
        w = int(np.uint64(system.ilut(d)[0]))
        if w >= (((w & A) >> 1) | ((w & B) << 1)):
            out.append(d)
    return out


def run_with_doubles_core(engine, system, hii, tau, target, n_iter, steps_sft=1, sft_damp=0.1, start=10.0, diag_sft=0.0):
    """`semi-stochastic doubles-core` + HPHF as the reference's HeHe_SS_Doubles case sets it up: the core space is the
    reference plus every HPHF function connected to it (singles and doubles), H over it in the HPHF basis; the
    run starts from `start` walkers on the reference (startsinglepart) and the shift varies once `target` is reached."""
    ref = [int(x) for x in system.ref_orbs]
    dets = hphf_allowed(system, all_dets(system))
    il = np.array([system.ilut(d) for d in dets], dtype=np.int64).reshape(len(dets), system.nw)
    iref = dets.index(ref)
    core = [ref] + [d for k, d in enumerate(dets) if k != iref and _hphf_level(system, ref, d) <= 2]
    core, sizes, displs, per_rank, H = build_core_space(engine, system, core, hii)
    flags = (1 << capi.FLAG_DETERMINISTIC) | (1 << capi.FLAG_INITIATOR)
    recs = np.array([host.record(system, d, start if d == ref else 0.0, flags) for d in core])
    engine.upload_walkers(recs)
    c = per_rank[0]
    engine.set_core_space(c["row_ptr"], c["col"], c["val"], sizes, displs, c["iluts"])
    run = driver.FciMC(system, engine, hii, tau=tau, init_walkers=target, steps_sft=steps_sft, sft_damp=sft_damp,
                       diag_sft=diag_sft)
    run.tot_parts = start; run.old_av_walkers = start
    run.run(n_iter)
    return run


def _hphf_level(system, ref, d):
    """Excitation level between HPHF functions: the smaller of the levels to the determinant and to its spin flip."""
    A, B = 0xAAAAAAAAAAAAAAAA, 0x5555555555555555
    r = int(np.uint64(system.ilut(ref)[0])); w = int(np.uint64(system.ilut(d)[0]))
    f = ((w & A) >> 1) | ((w & B) << 1)
    return min(bin(r & ~w).count("1"), bin(r & ~f).count("1"))
