"""Independent second-quantised construction of Hamiltonian matrix elements
(Jordan-Wigner ordering = ascending spin-orbital index), used to pin the
oracle's Slater-Condon restatement, parities and integral indexing without
reading any of its code.

    H = sum_pq h_pq a+_p a_q + 1/2 sum_pqrs (pq|rs) a+_p a+_r a_s a_q  (+ E_core)

with spin orbitals p,q,r,s (1-based, odd = beta, even = alpha) and chemist
integrals over the spatial parts, spin conserved in (pq) and (rs).
"""
import numpy as np


def _apply(op_list, det):
    """Apply a string of (creation?, orbital) operators, rightmost first, to a determinant (frozenset
    as sorted tuple).  Returns (sign, det) or (0, None)."""
    occ = list(det)
    sign = 1
    for create, p in reversed(op_list):
        if create:
            if p in occ:
                return 0, None
            k = sum(1 for o in occ if o < p)
            sign *= (-1) ** k
            occ.insert(k, p)
        else:
            if p not in occ:
                return 0, None
            k = occ.index(p)
            sign *= (-1) ** k
            occ.pop(k)
    return sign, tuple(occ)


def hamiltonian(dets, nbasis, h1, eri, ecore=0.0):
    """dets: list of sorted orbital tuples; h1(p,q) spin-orbital one-body; eri(i,j,k,l) chemist (ij|kl) over
    spatial orbitals (1-based)."""
    index = {tuple(d): i for i, d in enumerate(dets)}
    n = len(dets)
    H = np.zeros((n, n))
    sp = lambda o: (o + 1) // 2
    same = lambda a, b: (a - b) % 2 == 0
    orbs = range(1, nbasis + 1)
    for j, dj in enumerate(dets):
        dj = tuple(dj)
        H[j, j] += ecore
        for p in orbs:
            for q in orbs:
                v = h1(p, q)
                if v != 0.0:
                    s, d = _apply([(True, p), (False, q)], dj)
                    if s and d in index:
                        H[index[d], j] += s * v
        for q in dj:
            for s_ in dj:
                if s_ == q:
                    continue
                for p in orbs:
                    if not same(p, q):
                        continue
                    for r in orbs:
                        if not same(r, s_):
                            continue
                        v = eri(sp(p), sp(q), sp(r), sp(s_))
                        if v == 0.0:
                            continue
                        sg, d = _apply([(True, p), (True, r), (False, s_), (False, q)], dj)
                        if sg and d in index:
                            H[index[d], j] += 0.5 * sg * v
    return H
