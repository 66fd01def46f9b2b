"""FCIDUMP files: the integral format the reference's host reads (src/read_fci.F90, src/readint.F90) and that
bench.py's synthetic systems mimic.  Reader, writer and the frozen-core folding of `freeze n 0`
(src/Integrals_neci.F90, IntFreeze), so that the stand-alone driver can run on the reference's own input files.
In a deployment behind NECI these stay in the Fortran host; only the resulting UMAT / TMAT2D cross the C ABI.

Record layout after the &FCI ... &END namelist, one line per value: `value i j k l` with 1-based spatial orbitals:
  i j k l > 0   (ij|kl) in chemist order       i j > 0, k = l = 0   h_ij
  i > 0, rest 0 orbital energy eps_i            all zero             core energy
"""
import re
from dataclasses import dataclass, field

import numpy as np


@dataclass
class FciDump:
    norb: int
    nelec: int
    ms2: int = 0
    orbsym: list = None
    ecore: float = 0.0
    eps: list = None
    h1: list = field(default_factory=list)       # [(i, j, value)]
    eri: list = field(default_factory=list)      # [(i, j, k, l, value)]

    def system(self, **kw):
        """The engine-facing System (UMAT, TMAT2D, PCHB tables, reference determinant)."""
        from . import host
        return host.fcidump_system(self.norb, self.nelec, self.h1, self.eri, ecore=self.ecore, ms2=self.ms2,
                                   orbsym=self.orbsym, eps=self.eps, **kw)


def read_fcidump(path):
    txt = open(path).read()
    m = re.search(r"&END|^\s*/\s*$", txt, flags=re.M)
    if not m:
        raise ValueError("%s: no end of the &FCI namelist" % path)
    head, body = txt[:m.start()], txt[m.end():]

    def num(name, default=None):
        g = re.search(name + r"\s*=\s*(-?\d+)", head, flags=re.I)
        if g is None:
            if default is None:
                raise ValueError("%s: %s missing from the header" % (path, name))
            return default
        return int(g.group(1))
    norb, nelec, ms2 = num("NORB"), num("NELEC"), num("MS2", 0)
    g = re.search(r"ORBSYM\s*=\s*([\d,\s]+)", head, flags=re.I)
    orbsym = [int(x) for x in g.group(1).replace("\n", " ").split(",") if x.strip()][:norb] if g else [1] * norb
    d = FciDump(norb=norb, nelec=nelec, ms2=ms2, orbsym=orbsym)
    eps = {}
    for ln in body.splitlines():
        t = ln.replace("D", "E").replace("d", "e").split()
        if len(t) != 5:
            continue
        v = float(t[0]); i, j, k, l = (int(x) for x in t[1:])
        if i == 0:
            d.ecore = v
        elif j == 0:
            eps[i] = v
        elif k == 0:
            d.h1.append((i, j, v))
        else:
            d.eri.append((i, j, k, l, v))
    if eps:
        d.eps = [eps.get(i, 0.0) for i in range(1, norb + 1)]
    return d


def write_fcidump(path, d):
    with open(path, "w") as f:
        f.write(" &FCI NORB=%d,NELEC=%d,MS2=%d,\n  ORBSYM=%s,\n  ISYM=1,\n &END\n" %
                (d.norb, d.nelec, d.ms2, ",".join(str(int(x)) for x in (d.orbsym or [1] * d.norb))))
        for i, j, k, l, v in d.eri:
            f.write("%28.20E%4d%4d%4d%4d\n" % (v, i, j, k, l))
        for i, j, v in d.h1:
            f.write("%28.20E%4d%4d%4d%4d\n" % (v, i, j, 0, 0))
        for i, e in enumerate(d.eps or [], 1):
            f.write("%28.20E%4d%4d%4d%4d\n" % (e, i, 0, 0, 0))
        f.write("%28.20E%4d%4d%4d%4d\n" % (d.ecore, 0, 0, 0, 0))


def freeze_core(d, core):
    """`freeze`: the spatial orbitals in `core` (1-based) stay doubly occupied.  Their energy goes into ECore,
    their mean field into the one-body integrals: E' = E + sum_c 2 h_cc + sum_cc' [2 (cc|c'c') - (cc'|c'c)],
    h'_pq = h_pq + sum_c [2 (pq|cc) - (pc|cq)]; the remaining orbitals are renumbered in order."""
    core = sorted(int(c) for c in core)
    n = d.norb
    h = np.zeros((n + 1, n + 1))
    for i, j, v in d.h1:
        h[i, j] = h[j, i] = v
    g = {}
    for i, j, k, l, v in d.eri:
        for q in ((i, j, k, l), (j, i, k, l), (i, j, l, k), (j, i, l, k), (k, l, i, j), (l, k, i, j), (k, l, j, i), (l, k, j, i)):
            g[q] = v
    eri = lambda a, b, c, e: g.get((a, b, c, e), 0.0)
    keep = [p for p in range(1, n + 1) if p not in core]
    ecore = d.ecore
    for c in core:
        ecore += 2.0 * h[c, c]
        for c2 in core:
            ecore += 2.0 * eri(c, c, c2, c2) - eri(c, c2, c2, c)
    out = FciDump(norb=len(keep), nelec=d.nelec - 2 * len(core), ms2=d.ms2,
                  orbsym=[d.orbsym[p - 1] for p in keep] if d.orbsym else None, ecore=ecore,
                  eps=[d.eps[p - 1] for p in keep] if d.eps else None)
    for x, p in enumerate(keep, 1):
        for y, q in enumerate(keep[:x], 1):
            v = h[p, q] + sum(2.0 * eri(p, q, c, c) - eri(p, c, c, q) for c in core)
            if v != 0.0:
                out.h1.append((x, y, v))
    m = len(keep)
    seen = set()
    for x in range(1, m + 1):
        for y in range(1, x + 1):
            for z in range(1, x + 1):
                for w in range(1, z + 1):
                    key = tuple(sorted([tuple(sorted((x, y), reverse=True)), tuple(sorted((z, w), reverse=True))], reverse=True))
                    if key in seen:
                        continue
                    seen.add(key)
                    v = eri(keep[x - 1], keep[y - 1], keep[z - 1], keep[w - 1])
                    if v != 0.0:
                        out.eri.append((x, y, z, w, v))
    return out
