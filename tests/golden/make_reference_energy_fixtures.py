"""Generates tests/golden/reference_energies.json: for several checked-in regression cases of the reference
(test_suite/neci/parallel/*) the integrals over the orbitals its reference determinant occupies -- frozen core folded
into ECore and the one-body integrals as the case's `freeze` line asks -- together with the `Reference Energy set to`
value the reference's own run printed.  Known answers for FCIDUMP -> UMAT -> sltcnd_0, closed and open shell.
Run in the build container (/root/reference present); the tests read only the JSON."""
import glob
import json
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
CASES = ["C2_FCIMCPar_CAS", "N_FCIMCPar", "H4_FCIMCPar_InitHF", "Cr2_FCIMCPar_LinAlgo", "H2O_FCIMCPar_InitPOPS",
         "Ne_FCIMCPar"]


def parse_fcidump(path):
    txt = open(path).read()
    m = re.search(r"&END|/\s*\n", txt)
    head, body = txt[:m.start()], txt[m.end():]
    norb = int(re.search(r"NORB\s*=\s*(\d+)", head).group(1))
    h = np.zeros((norb + 1, norb + 1)); eps = np.zeros(norb + 1); ecore = 0.0; g = {}
    for ln in body.strip().splitlines():
        t = ln.split()
        if len(t) != 5:
            continue
        v = float(t[0]); i, j, k, l = (int(x) for x in t[1:])
        if i == 0: ecore = v
        elif j == 0: eps[i] = v
        elif k == 0: h[i, j] = h[j, i] = v
        else:
            for key in ((i, j, k, l), (j, i, k, l), (i, j, l, k), (j, i, l, k), (k, l, i, j), (l, k, i, j), (k, l, j, i), (l, k, j, i)):
                g[key] = v
    return norb, h, eps, ecore, g


def main():
    out = []
    for case in CASES:
        d = os.path.join(REF, "test_suite", "neci", "parallel", case)
        norb, h, eps, ecore, g = parse_fcidump(os.path.join(d, "FCIDUMP"))
        inp = glob.glob(os.path.join(d, "*.inp"))
        nfrz = 0
        if inp:
            m = re.search(r"^\s*freeze\s+(\d+)\s+(\d+)", open(inp[0]).read(), re.M | re.I)
            nfrz = int(m.group(1)) // 2 if m else 0
        bench = open(glob.glob(os.path.join(d, "benchmark*"))[0]).read()
        det = [int(x) for x in re.search(r"Generated reference determinants:\s*\n\(\s*([\d,\s]+)\)", bench).group(1).replace(",", " ").split()]
        eref = float(re.search(r"Reference Energy set to:\s+(-?[\d.]+)", bench).group(1))
        eri = lambda a, b, c, e: g.get((a, b, c, e), 0.0)
        have_eps = np.any(eps[1:] != 0.0)
        order = np.argsort(eps[1:] if have_eps else np.diag(h)[1:], kind="stable") + 1
        core = sorted(int(x) for x in order[:nfrz])
        keep = [p for p in range(1, norb + 1) if p not in core]
        e2 = ecore + sum(2.0 * h[c, c] for c in core) + sum(2.0 * eri(c, c, c2, c2) - eri(c, c2, c2, c) for c in core for c2 in core)
        occ_sp = sorted({(o + 1) // 2 for o in det})                       # kept numbering
        orig = [keep[x - 1] for x in occ_sp]
        hh = lambda p, q: h[p, q] + sum(2.0 * eri(p, q, c, c) - eri(p, c, c, q) for c in core)
        n = len(occ_sp)
        h1 = [(x + 1, y + 1, hh(orig[x], orig[y])) for x in range(n) for y in range(x + 1)]
        er = [(x + 1, y + 1, z + 1, w + 1, eri(orig[x], orig[y], orig[z], orig[w]))
              for x in range(n) for y in range(n) for z in range(n) for w in range(n)
              if eri(orig[x], orig[y], orig[z], orig[w]) != 0.0 and y <= x and w <= z and (z, w) <= (x, y)]
        red = sorted(2 * (occ_sp.index((o + 1) // 2) + 1) - (o % 2) for o in det)   # reduced spin-orbital numbering
        # independent evaluation (numpy) of the determinant energy, to catch a wrong numbering at generation time
        e = e2
        for o in red:
            e += hh(orig[(o + 1) // 2 - 1], orig[(o + 1) // 2 - 1])
        for a in range(len(red)):
            for b in range(a):
                p, q = orig[(red[a] + 1) // 2 - 1], orig[(red[b] + 1) // 2 - 1]
                e += eri(p, p, q, q) - (eri(p, q, q, p) if (red[a] - red[b]) % 2 == 0 else 0.0)
        status = "ok" if abs(e - eref) < 5e-9 else "MISMATCH"
        print("%-24s norb %2d frozen %2d det %s  E %.10f  reference %.10f  %s" % (case, norb, nfrz, det, e, eref, status))
        if status == "ok":
            out.append(dict(case=case, n_spat=n, det=red, nalpha=sum(1 for o in red if o % 2 == 0), nbeta=sum(1 for o in red if o % 2),
                            ecore=e2, h1=h1, eri=er, reference_energy=eref))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_energies.json")
    json.dump(out, open(dst, "w"))
    print("wrote", dst, os.path.getsize(dst), "bytes,", len(out), "cases")


if __name__ == "__main__":
    main()
