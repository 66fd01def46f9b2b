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
